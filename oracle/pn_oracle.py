"""ctypes binding of oracle/libpn_oracle.so, the CPU restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpn_oracle.so")
NM = 20


class Params(C.Structure):
    _fields_ = [("box", C.c_double), ("rs", C.c_double), ("cutoff", C.c_double), ("soft", C.c_double),
                ("theta", C.c_double), ("mass", C.c_double), ("maxleaf", C.c_int), ("periodic", C.c_int),
                ("longshort", C.c_int), ("pad", C.c_int)]


class Let(C.Structure):
    _fields_ = [("nnode", C.c_int), ("nbody", C.c_int), ("npart", C.POINTER(C.c_int)), ("son", C.POINTER(C.c_int)),
                ("width", C.POINTER(C.c_double)), ("center", C.POINTER(C.c_double)), ("M", C.POINTER(C.c_double)),
                ("body", C.POINTER(C.c_double)), ("origin", C.POINTER(C.c_int))]


def make_params(box, nside, npart_total, mass, maxleaf=8, theta=0.4, split=-1.0, soft=-1.0, periodic=1, longshort=1):
    """Derived force parameters exactly as src/initial.c:316-345."""
    rs = 1.25 * (box / float(nside))
    eps = 0.03 * box / (float(npart_total) ** 0.3333333)
    if split > 0.0:
        rs = split
    if soft >= 0.0:
        eps = soft
    return Params(box, rs, 4.5 * rs, eps, theta, mass, maxleaf, periodic, longshort, 0)


def build():
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
        L.pno_tree_build.restype = C.c_void_p
        L.pno_tree_build.argtypes = [dp, lp, C.c_int, C.c_int, C.c_int, dp, dp]
        L.pno_treeB_build.restype = C.c_void_p
        L.pno_treeB_build.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, dp, dp, ip]
        L.pno_tree_free.argtypes = [C.c_void_p]
        L.pno_tree_sizes.argtypes = [C.c_void_p, ip, ip, ip, ip, ip]
        L.pno_tree_get_leaves.argtypes = [C.c_void_p, ip, ip, dp, dp, dp, dp]
        L.pno_tree_get_nodes.argtypes = [C.c_void_p, ip, ip, dp, dp, dp, dp, dp]
        L.pno_tree_upward.argtypes = [C.c_void_p, dp, C.c_double]
        L.pno_tree_downward.argtypes = [C.c_void_p, dp, dp]
        pip = C.POINTER(ip)
        L.pno_walk_local.argtypes = [C.c_void_p, C.POINTER(Params), pip, pip, lp, pip, pip, lp]
        L.pno_free.argtypes = [C.c_void_p]
        L.pno_eval_p2p.argtypes = [C.c_void_p, dp, C.POINTER(Params), ip, ip, C.c_long, dp]
        L.pno_eval_m2l.argtypes = [C.c_void_p, C.POINTER(Params), ip, ip, C.c_long]
        L.pno_let_pack.restype = C.POINTER(Let)
        L.pno_let_pack.argtypes = [C.c_void_p, dp, C.POINTER(Params), dp, dp, dp]
        L.pno_let_free.argtypes = [C.POINTER(Let)]
        L.pno_walk_remote.argtypes = [C.c_void_p, C.POINTER(Let), C.POINTER(Params), pip, pip, lp, pip, pip, lp]
        L.pno_eval_p2p_remote.argtypes = [C.c_void_p, dp, C.POINTER(Let), C.POINTER(Params), ip, ip, C.c_long, dp]
        L.pno_eval_m2l_remote.argtypes = [C.c_void_p, C.POINTER(Let), C.POINTER(Params), ip, ip, C.c_long]
        L.pno_domain_boxes.argtypes = [C.c_int, C.c_double, dp, dp, ip, dp]
        L.pno_domain_of.restype = C.c_int
        L.pno_domain_of.argtypes = [dp, C.c_int, dp]
        L.pno_force.argtypes = [dp, C.c_int, C.c_int, C.POINTER(Params), dp, dp]
        L.pno_p2m.argtypes = [dp, C.c_int, C.c_int, dp, C.c_double, dp]
        L.pno_m2m.argtypes = [C.c_double] * 3 + [dp, dp]
        L.pno_m2l.argtypes = [C.c_double] * 3 + [dp, dp, C.c_double, C.c_int]
        L.pno_l2l.argtypes = [C.c_double] * 3 + [dp, dp]
        L.pno_l2p.argtypes = [dp, C.c_int, C.c_int, dp, dp, dp]
        L.pno_acceptance.restype = C.c_int
        L.pno_acceptance.argtypes = [dp, dp, dp, C.c_double, C.c_double, C.c_int]
        L.pno_p2p_pair.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.POINTER(Params), dp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _take_list(L, ps, pt, n):
    n = n.value
    s = np.ctypeslib.as_array(ps, shape=(n,)).copy() if n else np.zeros(0, np.int32)
    t = np.ctypeslib.as_array(pt, shape=(n,)).copy() if n else np.zeros(0, np.int32)
    L.pno_free(ps)
    L.pno_free(pt)
    return s.astype(np.int32), t.astype(np.int32)


class Tree:
    """Local k-d tree of one domain, built exactly like src/fmm.c:30-264."""

    def __init__(self, pos, maxleaf, bl, br, direct0=0):
        L = lib()
        self.pos = np.array(pos, dtype=np.float64, order="C", copy=True)   # permuted in place by the build
        n = self.pos.shape[0]
        self.ids = np.arange(n, dtype=np.int64)
        bl = np.asarray(bl, np.float64)
        br = np.asarray(br, np.float64)
        self.h = L.pno_tree_build(_dp(self.pos), self.ids.ctypes.data_as(C.POINTER(C.c_long)), n, maxleaf, direct0,
                                  _dp(bl), _dp(br))
        v = [C.c_int() for _ in range(5)]
        L.pno_tree_sizes(self.h, *[C.byref(x) for x in v])
        self.n, self.first_leaf, self.last_leaf, self.first_node, self.last_node = [x.value for x in v]
        self.nleaf = self.last_leaf - self.first_leaf
        self.nnode = self.last_node - self.first_node + 1

    def __del__(self):
        if getattr(self, "h", None):
            lib().pno_tree_free(self.h)
            self.h = None

    def leaves(self):
        nl = self.nleaf
        d = {"npart": np.zeros(nl, np.int32), "ipart": np.zeros(nl, np.int32), "center": np.zeros((nl, 3)),
             "width": np.zeros((nl, 3)), "M": np.zeros((nl, NM)), "L": np.zeros((nl, NM))}
        lib().pno_tree_get_leaves(self.h, _ip(d["npart"]), _ip(d["ipart"]), _dp(d["center"]), _dp(d["width"]),
                                  _dp(d["M"]), _dp(d["L"]))
        return d

    def nodes(self):
        nn = self.nnode
        d = {"npart": np.zeros(nn, np.int32), "son": np.zeros((nn, 2), np.int32), "split": np.zeros(nn),
             "center": np.zeros((nn, 3)), "width": np.zeros((nn, 3)), "M": np.zeros((nn, NM)), "L": np.zeros((nn, NM))}
        lib().pno_tree_get_nodes(self.h, _ip(d["npart"]), _ip(d["son"]), _dp(d["split"]), _dp(d["center"]),
                                 _dp(d["width"]), _dp(d["M"]), _dp(d["L"]))
        return d

    def upward(self, mass):
        lib().pno_tree_upward(self.h, _dp(self.pos), mass)

    def downward(self, acc):
        lib().pno_tree_downward(self.h, _dp(self.pos), _dp(acc))

    def walk_local(self, prm):
        L = lib()
        ps, pt, ms, mt = [C.POINTER(C.c_int)() for _ in range(4)]
        n1, n2 = C.c_long(), C.c_long()
        L.pno_walk_local(self.h, C.byref(prm), C.byref(ps), C.byref(pt), C.byref(n1), C.byref(ms), C.byref(mt), C.byref(n2))
        return _take_list(L, ps, pt, n1) + _take_list(L, ms, mt, n2)

    def eval_p2p(self, prm, s, t, acc):
        s = np.ascontiguousarray(s, np.int32)
        t = np.ascontiguousarray(t, np.int32)
        lib().pno_eval_p2p(self.h, _dp(self.pos), C.byref(prm), _ip(s), _ip(t), len(s), _dp(acc))

    def eval_m2l(self, prm, s, t):
        s = np.ascontiguousarray(s, np.int32)
        t = np.ascontiguousarray(t, np.int32)
        lib().pno_eval_m2l(self.h, C.byref(prm), _ip(s), _ip(t), len(s))

    def let_pack(self, prm, tcenter, twidth, displace):
        tc, tw, ds = [np.ascontiguousarray(x, np.float64) for x in (tcenter, twidth, displace)]
        return LetTree(lib().pno_let_pack(self.h, _dp(self.pos), C.byref(prm), _dp(tc), _dp(tw), _dp(ds)))

    def walk_remote(self, let, prm):
        L = lib()
        ps, pt, ms, mt = [C.POINTER(C.c_int)() for _ in range(4)]
        n1, n2 = C.c_long(), C.c_long()
        L.pno_walk_remote(self.h, let.p, C.byref(prm), C.byref(ps), C.byref(pt), C.byref(n1), C.byref(ms), C.byref(mt), C.byref(n2))
        return _take_list(L, ps, pt, n1) + _take_list(L, ms, mt, n2)

    def eval_p2p_remote(self, let, prm, s, t, acc):
        s = np.ascontiguousarray(s, np.int32)
        t = np.ascontiguousarray(t, np.int32)
        lib().pno_eval_p2p_remote(self.h, _dp(self.pos), let.p, C.byref(prm), _ip(s), _ip(t), len(s), _dp(acc))

    def eval_m2l_remote(self, let, prm, s, t):
        s = np.ascontiguousarray(s, np.int32)
        t = np.ascontiguousarray(t, np.int32)
        lib().pno_eval_m2l_remote(self.h, let.p, C.byref(prm), _ip(s), _ip(t), len(s))


class TreeB(Tree):
    """Mode B tree: CPU restatement of the device builder (pn2_tree.cu), in the reference id space."""

    def __init__(self, pos, maxleaf, bl, br, direct0=0):
        L = lib()
        pin = np.ascontiguousarray(pos, np.float64)
        n = pin.shape[0]
        self.pos = np.zeros((n, 3))
        order = np.zeros(n, np.int32)
        bl = np.asarray(bl, np.float64)
        br = np.asarray(br, np.float64)
        self.h = L.pno_treeB_build(_dp(pin), n, maxleaf, direct0, _dp(bl), _dp(br), _dp(self.pos), _ip(order))
        self.ids = order.astype(np.int64)
        v = [C.c_int() for _ in range(5)]
        L.pno_tree_sizes(self.h, *[C.byref(x) for x in v])
        self.n, self.first_leaf, self.last_leaf, self.first_node, self.last_node = [x.value for x in v]
        self.nleaf = self.last_leaf - self.first_leaf
        self.nnode = self.last_node - self.first_node + 1


class LetTree:
    """Pruned, flattened, displaced tree as packed by src/remotes.c:60-169."""

    def __init__(self, p):
        self.p = p
        c = p.contents
        self.nnode, self.nbody = c.nnode, c.nbody

    def arrays(self):
        c = self.p.contents
        nn, nb = c.nnode, c.nbody
        f = np.ctypeslib.as_array
        return {"npart": f(c.npart, (nn,)).copy(), "son": f(c.son, (2 * nn,)).reshape(nn, 2).copy(),
                "width": f(c.width, (3 * nn,)).reshape(nn, 3).copy(), "center": f(c.center, (3 * nn,)).reshape(nn, 3).copy(),
                "M": f(c.M, (NM * nn,)).reshape(nn, NM).copy(), "origin": f(c.origin, (nn,)).copy(),
                "body": f(c.body, (3 * nb,)).reshape(nb, 3).copy() if nb else np.zeros((0, 3))}

    def __del__(self):
        if getattr(self, "p", None):
            lib().pno_let_free(self.p)
            self.p = None


def domain_boxes(nranks, box):
    c = np.zeros((nranks, 3))
    w = np.zeros((nranks, 3))
    d = np.zeros(nranks, np.int32)
    s = np.zeros(2 * nranks - 1)
    lib().pno_domain_boxes(nranks, box, _dp(c), _dp(w), _ip(d), _dp(s))
    return c, w, d, s


def force(pos, prm, nranks=1):
    """One whole short-range force evaluation; returns (acc in input order, counters dict)."""
    pos = np.ascontiguousarray(pos, np.float64)
    n = pos.shape[0]
    acc = np.zeros((n, 3))
    cnt = np.zeros(8)
    lib().pno_force(_dp(pos), n, nranks, C.byref(prm), _dp(acc), _dp(cnt))
    keys = ["p2p_pairs", "m2l_pairs", "int_local", "int_remote", "m2l_calls", "leaves", "nodes", "p2p_pairs_remote"]
    return acc, {k: int(v) for k, v in zip(keys, cnt)}
