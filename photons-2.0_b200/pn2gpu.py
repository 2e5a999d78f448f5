"""ctypes binding of libpn2gpu.so (include/pn2gpu.h) and a thin host-side mirror of the reference's
short-range call sequence (src/photoNs.c:97-116: fmm_prepare -> fmm_task -> fmm_ext).

This is product code: it never imports anything from oracle/.  If the CUDA library is missing or no
sm_100 device is present every entry point raises Pn2Error -- there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PN2GPU_LIB") or os.path.join(HERE, "libpn2gpu.so")      # PN2GPU_LIB: a kernel-parameter variant of the library (tools/build_variants.sh)
NM = 20
FP64, FP32, FP64_LIBM = 0, 1, 2       # include/pn2gpu.h: PN2_FP64 (table-driven, no libm), PN2_FP32, PN2_FP64_LIBM (checker)

# struct layouts of the reference (inc/typesdef.h:25-57, inc/photoNs.h:177-189); sizes 96/376/392/224/32
BODY = np.dtype([("pos", "f8", 3), ("acc", "f8", 3), ("vel", "f8", 3), ("acc_pm", "f8", 3)])
PACK = np.dtype([("npart", "i4"), ("ipart", "i4"), ("width", "f8", 3), ("center", "f8", 3),
                 ("M", "f8", NM), ("L", "f8", NM)])
NODE = np.dtype([("updated", "i4"), ("npart", "i4"), ("son", "i4", 2), ("split", "f8"), ("width", "f8", 3),
                 ("center", "f8", 3), ("M", "f8", NM), ("L", "f8", NM)])
RNODE = np.dtype([("npart", "i4"), ("son", "i4", 2), ("pad", "i4"), ("width", "f8", 3), ("center", "f8", 3),
                  ("M", "f8", NM)])
RBODY = np.dtype([("pos", "f8", 3), ("replenish", "f8")])
assert (BODY.itemsize, PACK.itemsize, NODE.itemsize, RNODE.itemsize, RBODY.itemsize) == (96, 376, 392, 224, 32)

EXPORTS = ["pn2_create", "pn2_destroy", "pn2_set_params", "pn2_sync", "pn2_last_error", "pn2_device_info",
           "pn2_set_particles", "pn2_set_tree", "pn2_set_remote", "pn2_p2m_m2m", "pn2_p2p_batch", "pn2_m2l_batch",
           "pn2_p2p_ext_batch", "pn2_m2l_ext_batch", "pn2_l2l_l2p", "pn2_get_acc", "pn2_zero_acc",
           "pn2_get_multipoles", "pn2_get_locals", "pn2_get_counters", "pn2_force_step", "pn2_force_step_device",
           "pn2_set_comm", "pn2_get_step_info", "pn2_get_order", "pn2_get_cells", "pn2_get_lists", "pn2_fma_peak",
           "pn2_get_timings", "pn2_launch_count", "pn2_timer_start", "pn2_timer_stop", "pn2_comm_unique_id",
           "pn2_comm_init_rank", "pn2_step_begin", "pn2_exchange_local", "pn2_step_finish", "pn2_domain_owner_device",
           "pn2_migrate_begin", "pn2_migrate_exchange_nccl", "pn2_migrate_exchange_local", "pn2_migrate_result",
           "pn2_migrate_device", "pn2_migrate_fetch", "pn2_kick_device", "pn2_drift_device", "pn2_force_step_records",
           "pn2_pm_force_device", "pn2_pm_force_records", "pn2_pm_begin", "pn2_pm_reduce_nccl", "pn2_pm_reduce_local",
           "pn2_pm_finish", "pn2_pm_get_mesh", "pn2_pm_get_timings", "pn2_snapshot_to_body_device", "pn2_body_to_snapshot_device", "pn2_get_timings_ex"]


class Pn2Error(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("box", C.c_double), ("rs", C.c_double), ("cutoff", C.c_double), ("soft", C.c_double),
                ("theta", C.c_double), ("mass", C.c_double), ("maxleaf", C.c_int32), ("periodic", C.c_int32),
                ("longshort", C.c_int32), ("precision", C.c_int32)]


class Domain(C.Structure):
    _fields_ = [("lo", C.c_double * 3), ("hi", C.c_double * 3), ("direct0", C.c_int32), ("pad_", C.c_int32)]


class StepInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("nleaf", C.c_int32), ("nnode", C.c_int32), ("nlevel", C.c_int32),
                ("n_p2p_pairs", C.c_int64), ("n_m2l_pairs", C.c_int64), ("n_interactions", C.c_int64),
                ("n_let_nodes", C.c_int64), ("n_let_bodies", C.c_int64), ("n_walk_visits", C.c_int64),
                ("frontier_bytes", C.c_int64)]


def make_params(box, nside, npart_total, mass, maxleaf=8, theta=0.4, split=-1.0, soft=-1.0, periodic=1, longshort=1,
                precision=FP32):
    """Derived force parameters exactly as the reference computes them (src/initial.c:316-345)."""
    rs = 1.25 * (box / float(nside))
    eps = 0.03 * box / (float(npart_total) ** 0.3333333)
    if split > 0.0:
        rs = split
    if soft >= 0.0:
        eps = soft
    return Params(box, rs, 4.5 * rs, eps, theta, mass, maxleaf, periodic, longshort, precision)


def build_library(verbose=False):
    """Compile libpn2gpu.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-s", "-C", os.path.join(HERE, "csrc"), "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise Pn2Error("building libpn2gpu.so failed:\n" + (r.stdout or "") + (r.stderr or ""))


_lib = None


def lib():
    """Load libpn2gpu.so; raises if it is not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pn2Error(f"{LIB_PATH} is missing: run __graft_entry__.build() (make -C photons-2.0_b200/csrc)")
    if "PN2_NCCL_LIB" not in os.environ:
        # one libnccl per process: prefer the copy bundled with torch (see csrc/pn2_let.cu)
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            if spec and spec.submodule_search_locations:
                cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["PN2_NCCL_LIB"] = cand
        except Exception:
            pass
    L = C.CDLL(LIB_PATH)
    vp, dp, ip, lp = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
    L.pn2_last_error.restype = C.c_char_p
    L.pn2_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params)]
    L.pn2_destroy.argtypes = [vp]
    L.pn2_set_params.argtypes = [vp, C.POINTER(Params)]
    L.pn2_sync.argtypes = [vp]
    L.pn2_device_info.argtypes = [C.c_int, ip, ip, ip, C.POINTER(C.c_size_t)]
    L.pn2_set_particles.argtypes = [vp, vp, C.c_size_t, C.c_int]
    L.pn2_set_tree.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    L.pn2_set_remote.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    for f in ("pn2_p2m_m2m", "pn2_l2l_l2p", "pn2_zero_acc"):
        getattr(L, f).argtypes = [vp]
    for f in ("pn2_p2p_batch", "pn2_m2l_batch", "pn2_p2p_ext_batch", "pn2_m2l_ext_batch"):
        getattr(L, f).argtypes = [vp, vp, vp, C.c_long]
    L.pn2_get_acc.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int]
    L.pn2_get_multipoles.argtypes = [vp, vp, vp]
    L.pn2_get_locals.argtypes = [vp, vp, vp]
    L.pn2_get_counters.argtypes = [vp, dp]
    L.pn2_force_step.argtypes = [vp, vp, C.c_size_t, C.c_int, C.POINTER(Domain), vp, C.c_size_t]
    L.pn2_force_step_device.argtypes = [vp, vp, C.c_int, C.POINTER(Domain), vp]
    L.pn2_set_comm.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Domain), vp]
    L.pn2_get_step_info.argtypes = [vp, C.POINTER(StepInfo)]
    L.pn2_get_order.argtypes = [vp, vp, C.c_int]
    L.pn2_get_cells.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pn2_get_lists.argtypes = [vp, C.c_int, lp, lp, vp, vp, vp]
    L.pn2_fma_peak.argtypes = [vp, C.c_int, dp, dp]
    L.pn2_get_timings.argtypes = [vp, dp]
    L.pn2_get_timings_ex.argtypes = [vp, dp, C.c_int]
    L.pn2_comm_unique_id.argtypes = [vp]
    L.pn2_comm_init_rank.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Domain), vp]
    L.pn2_step_begin.argtypes = [vp, vp, C.c_int, C.POINTER(Domain)]
    L.pn2_exchange_local.argtypes = [C.POINTER(vp), C.c_int]
    L.pn2_step_finish.argtypes = [vp, vp]
    L.pn2_domain_owner_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.pn2_migrate_begin.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp]
    L.pn2_migrate_exchange_nccl.argtypes = [vp]
    L.pn2_migrate_exchange_local.argtypes = [C.POINTER(vp), C.c_int]
    L.pn2_migrate_result.argtypes = [vp, C.POINTER(vp), ip, vp]
    L.pn2_migrate_device.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.POINTER(vp), ip]
    L.pn2_migrate_fetch.argtypes = [vp, vp]
    L.pn2_kick_device.argtypes = [vp, vp, C.c_int, C.c_double, C.c_int]
    L.pn2_drift_device.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double]
    L.pn2_force_step_records.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(Domain)]
    L.pn2_pm_force_device.argtypes = [vp, vp, C.c_int, C.c_int, vp]
    L.pn2_pm_force_records.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pn2_pm_begin.argtypes = [vp, vp, C.c_int, C.c_int]
    L.pn2_pm_reduce_nccl.argtypes = [vp]
    L.pn2_pm_reduce_local.argtypes = [C.POINTER(vp), C.c_int]
    L.pn2_pm_finish.argtypes = [vp, vp]
    L.pn2_pm_get_mesh.argtypes = [vp, vp]
    L.pn2_pm_get_timings.argtypes = [vp, dp]
    L.pn2_snapshot_to_body_device.argtypes = [vp, vp, vp, C.c_int, C.c_double, vp]
    L.pn2_body_to_snapshot_device.argtypes = [vp, vp, C.c_int, C.c_double, vp, vp]
    L.pn2_timer_start.argtypes = [vp, C.c_int]
    L.pn2_timer_stop.argtypes = [vp, C.c_int, dp]
    L.pn2_launch_count.argtypes = [vp]
    L.pn2_launch_count.restype = C.c_long
    _lib = L
    return L


def _ck(rc):
    if rc != 0:
        raise Pn2Error(f"pn2 error {rc}: {lib().pn2_last_error().decode()}")


def _ptr(a):
    return None if a is None else a.ctypes.data


def make_domain(lo, hi, direct0=0):
    d = Domain()
    for k in range(3):
        d.lo[k] = float(lo[k])
        d.hi[k] = float(hi[k])
    d.direct0 = int(direct0)
    return d


class Context:
    """One device context (one rank).  Method names follow the C-ABI; docstrings cite the reference."""

    def __init__(self, params, device=0):
        self.h = C.c_void_p()
        self.params = params
        self.nranks = 1
        self._device = device
        _ck(lib().pn2_create(C.byref(self.h), device, C.byref(params)))

    def close(self):
        if self.h:
            lib().pn2_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Mode A -------------------------------------------------------------------------------
    def set_particles(self, pos, stride=None):
        """pos: (n,3) float64 C-contiguous, or a BODY structured array (stride 96)."""
        if pos.dtype == BODY:
            self._keep = pos
            _ck(lib().pn2_set_particles(self.h, pos.ctypes.data, 96, len(pos)))
        else:
            pos = np.ascontiguousarray(pos, np.float64)
            self._keep = pos
            _ck(lib().pn2_set_particles(self.h, pos.ctypes.data, 24, pos.shape[0]))
        self.n = len(pos)

    def set_tree(self, leaf, first_leaf, btree, first_node):
        """leaf: PACK array (leaf ids first_leaf..), btree: NODE array (node ids first_node..)."""
        assert leaf.dtype == PACK and btree.dtype == NODE
        leaf = np.ascontiguousarray(leaf)
        btree = np.ascontiguousarray(btree)
        _ck(lib().pn2_set_tree(self.h, _ptr(leaf), first_leaf, first_leaf + len(leaf), _ptr(btree), first_node,
                               first_node + len(btree) - 1))
        self.nleaf, self.nnode = len(leaf), len(btree)

    def set_remote(self, rtree, rbody):
        assert rtree.dtype == RNODE and rbody.dtype == RBODY
        rtree = np.ascontiguousarray(rtree)
        rbody = np.ascontiguousarray(rbody)
        _ck(lib().pn2_set_remote(self.h, _ptr(rtree), len(rtree), _ptr(rbody), len(rbody)))

    def p2m_m2m(self):
        """fmm_prepare's P2M loop + walk_m2m (src/fmm.c:741-744)."""
        _ck(lib().pn2_p2m_m2m(self.h))

    def _batch(self, fn, s, t):
        s = np.ascontiguousarray(s, np.int32)
        t = np.ascontiguousarray(t, np.int32)
        assert s.shape == t.shape
        _ck(fn(self.h, _ptr(s), _ptr(t), len(s)))

    def p2p_batch(self, task_s, task_t):
        """task_compute_p2p (src/fmm.c:796-872)."""
        self._batch(lib().pn2_p2p_batch, task_s, task_t)

    def m2l_batch(self, task_s, task_t):
        """task_compute_m2l (src/fmm.c:875-907)."""
        self._batch(lib().pn2_m2l_batch, task_s, task_t)

    def p2p_ext_batch(self, task_s, task_t):
        """task_compute_p2p_ext (src/remotes.c:583-596)."""
        self._batch(lib().pn2_p2p_ext_batch, task_s, task_t)

    def m2l_ext_batch(self, task_s, task_t):
        """task_compute_m2l_ext (src/remotes.c:598-628)."""
        self._batch(lib().pn2_m2l_ext_batch, task_s, task_t)

    def l2l_l2p(self):
        """walk_l2l + L2P loop (src/fmm.c:1054-1057)."""
        _ck(lib().pn2_l2l_l2p(self.h))

    def zero_acc(self):
        _ck(lib().pn2_zero_acc(self.h))

    def get_acc(self):
        acc = np.zeros((self.n, 3))
        _ck(lib().pn2_get_acc(self.h, acc.ctypes.data, 24, self.n, 0))
        return acc

    def get_multipoles(self):
        leaf = np.zeros(self.nleaf, PACK)
        btree = np.zeros(self.nnode, NODE)
        _ck(lib().pn2_get_multipoles(self.h, _ptr(leaf), _ptr(btree)))
        return leaf["M"].copy(), btree["M"].copy()

    def get_locals(self):
        leaf = np.zeros(self.nleaf, PACK)
        btree = np.zeros(self.nnode, NODE)
        _ck(lib().pn2_get_locals(self.h, _ptr(leaf), _ptr(btree)))
        return leaf["L"].copy(), btree["L"].copy()

    def counters(self):
        c = np.zeros(8)
        _ck(lib().pn2_get_counters(self.h, c.ctypes.data_as(C.POINTER(C.c_double))))
        return c

    def sync(self):
        _ck(lib().pn2_sync(self.h))

    # ---- Mode B -------------------------------------------------------------------------------
    def set_comm(self, rank, nranks, domains, nccl_comm):
        arr = (Domain * nranks)(*domains)
        _ck(lib().pn2_set_comm(self.h, rank, nranks, arr, nccl_comm))
        self.nranks = nranks

    def set_comm_torch(self, rank, nranks, domains):
        """Create this context's own NCCL communicator; the 128-byte unique id travels over torch.distributed
        (plumbing), the LET exchange itself is ncclSend/ncclRecv inside libpn2gpu.so."""
        import torch
        import torch.distributed as dist
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            _ck(lib().pn2_comm_unique_id(buf))
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        idb = (C.c_ubyte * 128).from_buffer_copy(raw)
        arr = (Domain * nranks)(*domains)
        _ck(lib().pn2_comm_init_rank(self.h, rank, nranks, arr, idb))
        self.nranks = nranks

    def step_begin(self, d_pos_ptr, n, domain):
        _ck(lib().pn2_step_begin(self.h, d_pos_ptr, n, C.byref(domain)))
        self.n = n

    def step_finish(self, d_acc_ptr):
        _ck(lib().pn2_step_finish(self.h, d_acc_ptr))

    def force_step(self, pos, domain=None):
        """One whole short-range force evaluation (src/photoNs.c:97-116 without PM); pos (n,3) float64 host."""
        pos = np.ascontiguousarray(pos, np.float64)
        n = pos.shape[0]
        if domain is None:
            b = self.params.box
            domain = make_domain([0, 0, 0], [b, b, b], 0)
        acc = np.zeros((n, 3))
        _ck(lib().pn2_force_step(self.h, pos.ctypes.data, 24, n, C.byref(domain), acc.ctypes.data, 24))
        self.n = n
        return acc

    def force_step_device(self, d_pos_ptr, n, d_acc_ptr, domain=None):
        """Same with device pointers (ints); asynchronous on the context stream."""
        if domain is None:
            b = self.params.box
            domain = make_domain([0, 0, 0], [b, b, b], 0)
        _ck(lib().pn2_force_step_device(self.h, d_pos_ptr, n, C.byref(domain), d_acc_ptr))
        self.n = n

    # ---- domain decomposition on the device (src/domains.c:268-375) ----
    def domain_owner_device(self, d_rec_ptr, rec_doubles, n, splits, nranks, d_owner_ptr):
        """Owner rank of n device records (first three doubles = position) under the domain-tree splits."""
        sp = np.ascontiguousarray(splits, np.float64)
        _ck(lib().pn2_domain_owner_device(self.h, d_rec_ptr, rec_doubles, n, sp.ctypes.data, nranks, d_owner_ptr))

    def migrate_begin(self, d_rec_ptr, rec_doubles, n, splits):
        """Classify and order the records by destination rank; returns sendcount[nranks]."""
        sp = np.ascontiguousarray(splits, np.float64)
        sc = np.zeros(max(1, self.nranks), np.int32)
        _ck(lib().pn2_migrate_begin(self.h, d_rec_ptr, rec_doubles, n, sp.ctypes.data, sc.ctypes.data))
        return sc

    def migrate_exchange_nccl(self):
        _ck(lib().pn2_migrate_exchange_nccl(self.h))

    def migrate_result(self):
        """(device pointer, n, recvcount[nranks]) of the records this rank owns after the exchange."""
        ptr, n = C.c_void_p(), C.c_int()
        rc = np.zeros(max(1, self.nranks), np.int32)
        _ck(lib().pn2_migrate_result(self.h, C.byref(ptr), C.byref(n), rc.ctypes.data))
        return ptr.value or 0, n.value, rc

    def migrate_fetch(self, rec_doubles):
        """The received records as a host array (n, rec_doubles)."""
        _, n, _ = self.migrate_result()
        out = np.zeros((n, rec_doubles))
        _ck(lib().pn2_migrate_fetch(self.h, out.ctypes.data))
        return out

    def migrate_device(self, d_rec_ptr, rec_doubles, n, splits):
        """prepare_deliver_realloc_body (src/domains.c:298-375) over NCCL: returns (device pointer, n_new)."""
        self.migrate_begin(d_rec_ptr, rec_doubles, n, splits)
        self.migrate_exchange_nccl()
        ptr, n_new, _ = self.migrate_result()
        return ptr, n_new

    # ---- KDK integrator on device Body records (src/photoNs.c:150-196, 254-268) ----
    def kick_device(self, d_body_ptr, n, dkh, pm_first):
        _ck(lib().pn2_kick_device(self.h, d_body_ptr, n, dkh, 1 if pm_first else 0))

    def drift_device(self, d_body_ptr, n, dd, box):
        _ck(lib().pn2_drift_device(self.h, d_body_ptr, n, dd, box))

    def force_step_records(self, d_rec_ptr, rec_doubles, n, domain=None, acc_offset=3):
        """Mode B force step on device records; accelerations go to doubles acc_offset.. of every record (Body.acc)."""
        if domain is None:
            b = self.params.box
            domain = make_domain([0, 0, 0], [b, b, b], 0)
        _ck(lib().pn2_force_step_records(self.h, d_rec_ptr, rec_doubles, acc_offset, n, C.byref(domain)))
        self.n = n

    # ---- PM long-range force (src/partmesh.c:18-796, src/conv.f90:128-247) ----
    def pm_force_device(self, d_pos_ptr, n, nside, d_acc_pm_ptr):
        """partmesh_thread on the device: packed device positions -> acc_pm (caller order); NCCL all-reduce of the mesh when nranks > 1."""
        _ck(lib().pn2_pm_force_device(self.h, d_pos_ptr, n, nside, d_acc_pm_ptr))

    def pm_force_records(self, d_rec_ptr, rec_doubles, n, nside, acc_pm_offset=9):
        _ck(lib().pn2_pm_force_records(self.h, d_rec_ptr, rec_doubles, acc_pm_offset, n, nside))

    def pm_force(self, pos, nside):
        """Host convenience: pos (n,3) float64 -> acc_pm (n,3)."""
        import torch
        t = torch.from_numpy(np.ascontiguousarray(pos, np.float64)).cuda(self.device_index())
        a = torch.zeros_like(t)
        self.pm_force_device(t.data_ptr(), t.shape[0], nside, a.data_ptr())
        self.sync()
        return a.cpu().numpy()

    def device_index(self):
        return getattr(self, "_device", 0)

    def pm_begin(self, d_pos_ptr, n, nside):
        _ck(lib().pn2_pm_begin(self.h, d_pos_ptr, n, nside))

    def pm_finish(self, d_acc_pm_ptr):
        _ck(lib().pn2_pm_finish(self.h, d_acc_pm_ptr))

    def pm_mesh(self, nside):
        m = np.zeros((nside, nside, nside))
        _ck(lib().pn2_pm_get_mesh(self.h, m.ctypes.data))
        return m

    def pm_timings(self):
        t = np.zeros(4)
        _ck(lib().pn2_pm_get_timings(self.h, t.ctypes.data_as(C.POINTER(C.c_double))))
        return dict(zip(["deposit", "reduce", "fft_green", "gather"], t))

    def step_info(self):
        s = StepInfo()
        _ck(lib().pn2_get_step_info(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in StepInfo._fields_}

    def get_order(self):
        o = np.zeros(self.n, np.int32)
        _ck(lib().pn2_get_order(self.h, o.ctypes.data, self.n))
        return o

    def get_cells(self, with_ml=True):
        si = self.step_info()
        nc = si["nleaf"] + si["nnode"]
        geom = np.zeros((nc, 6))
        son = np.zeros((nc, 2), np.int32)
        rng = np.zeros((nc, 2), np.int32)
        M = np.zeros((nc, NM)) if with_ml else None
        L = np.zeros((nc, NM)) if with_ml else None
        _ck(lib().pn2_get_cells(self.h, _ptr(geom), _ptr(son), _ptr(rng), _ptr(M), _ptr(L)))
        return {"nleaf": si["nleaf"], "nnode": si["nnode"], "geom": geom, "son": son, "range": rng, "M": M, "L": L}

    def get_lists(self, kind):
        nseg, nsrc = C.c_long(), C.c_long()
        _ck(lib().pn2_get_lists(self.h, kind, C.byref(nseg), C.byref(nsrc), None, None, None))
        sink = np.zeros(nseg.value, np.int32)
        off = np.zeros(nseg.value + 1, np.int64)
        src = np.zeros(nsrc.value, np.uint32)
        _ck(lib().pn2_get_lists(self.h, kind, C.byref(nseg), C.byref(nsrc), _ptr(sink), _ptr(off), _ptr(src)))
        return sink, off, src

    def timings(self):
        t = np.zeros(9)
        _ck(lib().pn2_get_timings_ex(self.h, t.ctypes.data_as(C.POINTER(C.c_double)), 9))
        return dict(zip(["tree", "upward", "walk_p2p", "m2l", "downward", "let", "total", "frontier", "m2l_kernel"], t))

    def fma_peak(self, fp64=False):
        ops, ms = C.c_double(), C.c_double()
        _ck(lib().pn2_fma_peak(self.h, int(fp64), C.byref(ops), C.byref(ms)))
        return ops.value

    def timer_start(self, slot=0):
        _ck(lib().pn2_timer_start(self.h, slot))

    def timer_stop(self, slot=0):
        ms = C.c_double()
        _ck(lib().pn2_timer_stop(self.h, slot, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().pn2_launch_count(self.h))


def short_range_force_mode_a(ctx, part_pos, leaf, first_leaf, btree, first_node, p2p, m2l, remotes=()):
    """The reference's call sequence with the device in place of the worker thread (Mode A):
    fmm_prepare (P2M, M2M) -> fmm_task (P2P batches, M2L batches) -> fmm_ext (per received LET:
    remote P2P / M2L batches) -> walk_l2l + L2P.  p2p / m2l = (task_s, task_t); remotes = iterable of
    (rtree, rbody, (p2p_s, p2p_t), (m2l_s, m2l_t)).  Returns accelerations in tree order."""
    ctx.set_particles(part_pos)
    ctx.set_tree(leaf, first_leaf, btree, first_node)
    ctx.p2m_m2m()
    ctx.p2p_batch(*p2p)
    ctx.m2l_batch(*m2l)
    for rtree, rbody, rp2p, rm2l in remotes:
        ctx.set_remote(rtree, rbody)
        ctx.p2p_ext_batch(*rp2p)
        ctx.m2l_ext_batch(*rm2l)
    ctx.l2l_l2p()
    return ctx.get_acc()


def migrate_exchange_local(ctxs):
    """The all-to-all-v of the records between contexts of this process (after migrate_begin on every rank)."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _ck(lib().pn2_migrate_exchange_local(arr, len(ctxs)))


def pm_reduce_local(ctxs):
    """Sum of the density meshes of the contexts of this process (after pm_begin on every rank)."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _ck(lib().pn2_pm_reduce_local(arr, len(ctxs)))


def exchange_local(ctxs):
    """LET exchange between the contexts of one process (ctxs[r] = rank r)."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _ck(lib().pn2_exchange_local(arr, len(ctxs)))


def force_step_local_ranks(ctxs, pos_by_rank, domains):
    """One multi-rank short-range force evaluation with every rank a context of THIS process (device-to-device LET
    exchange instead of NCCL).  pos_by_rank: list of (n_r, 3) float64 host arrays.  Returns the list of accelerations."""
    import torch
    nr = len(ctxs)
    dpos, dacc = [], []
    for r in range(nr):
        ctxs[r].set_comm(r, nr, domains, None)
        t = torch.from_numpy(np.ascontiguousarray(pos_by_rank[r], np.float64)).cuda()
        dpos.append(t)
        dacc.append(torch.zeros_like(t))
    torch.cuda.synchronize()
    for r in range(nr):
        ctxs[r].step_begin(dpos[r].data_ptr(), dpos[r].shape[0], domains[r])
    exchange_local(ctxs)
    for r in range(nr):
        ctxs[r].step_finish(dacc[r].data_ptr())
        ctxs[r].sync()
    return [a.cpu().numpy() for a in dacc]
