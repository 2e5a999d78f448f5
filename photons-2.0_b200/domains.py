"""The reference's domain decomposition rule (src/domains.c, src/initial.c:199-223, src/toptree.c:150-181) as
host-side geometry: a heap-numbered binary domain tree over the ranks, split direction cycling x->y->z from the
root, initial splits proportional to the number of domains on each side (domain_volume_part,
src/domains.c:399-428).  Product code (no oracle imports); checked against the oracle in tests/."""
import numpy as np

import pn2gpu


def mostleft(P):
    """src/initial.c:199-223"""
    m = 1
    while m < 2 * P - 1:
        m *= 2
    m = m // 2 - 1
    return 0 if P == 1 else m


def domain_node_of_rank(rank, P):
    d = rank + mostleft(P)
    if d > 2 * P - 2:
        d -= P
    return d


def _count(node, P):
    return 1 if node >= P - 1 else _count(2 * node + 1, P) + _count(2 * node + 2, P)


def domain_tree(P, box):
    """Returns (splits[2P-1], lo[2P-1,3], hi[2P-1,3], direct[2P-1]) of the heap-numbered domain tree."""
    n = 2 * P - 1
    splits = np.zeros(n)
    lo = np.zeros((n, 3))
    hi = np.zeros((n, 3))
    direct = np.zeros(n, np.int32)

    def rec(node, dim, bl, br):
        lo[node] = bl
        hi[node] = br
        direct[node] = dim
        if node >= P - 1:
            return
        tl, tr = float(_count(2 * node + 1, P)), float(_count(2 * node + 2, P))
        frac = bl[dim] + (br[dim] - bl[dim]) * tl / (tl + tr)
        splits[node] = frac
        b2 = br.copy()
        b2[dim] = frac
        rec(2 * node + 1, (dim + 1) % 3, bl.copy(), b2)
        b1 = bl.copy()
        b1[dim] = frac
        rec(2 * node + 2, (dim + 1) % 3, b1, br.copy())

    rec(0, 0, np.zeros(3), np.full(3, float(box)))
    return splits, lo, hi, direct


def fraction(size):
    """(nleft, nright) domains below a domain-tree node that covers `size` ranks (src/domains.c:42-84)."""
    if size == 1:
        return 1, 0
    if size == 2:
        return 1, 1
    if size == 3:
        return 2, 1
    left, right = 1, 2
    while size - left >= right - size:
        left *= 2
        right *= 2
    left >>= 1
    right = size - left
    if left < right:
        left = right
        right = size - left
    return left, right


def boxes_from_splits(P, box, splits):
    """(lo[2P-1,3], hi[2P-1,3], direct[2P-1]) of the heap-numbered domain tree for given splits
    (center_toptree, src/toptree.c:150-181: the box of a node is its parent's box cut at the parent's split)."""
    n = 2 * P - 1
    lo = np.zeros((n, 3))
    hi = np.zeros((n, 3))
    direct = np.zeros(n, np.int32)

    def rec(node, dim, bl, br):
        lo[node] = bl
        hi[node] = br
        direct[node] = dim
        if node >= P - 1:
            return
        b2 = br.copy()
        b2[dim] = splits[node]
        rec(2 * node + 1, (dim + 1) % 3, bl.copy(), b2)
        b1 = bl.copy()
        b1[dim] = splits[node]
        rec(2 * node + 2, (dim + 1) % 3, b1, br.copy())

    rec(0, 0, np.zeros(3), np.full(3, float(box)))
    return lo, hi, direct


class DomainTree:
    """The reference's adaptive domain decomposition as host state: the splits of domain_initialize()
    (src/domains.c:432-470) and their per-step adjustment by the load every rank reports
    (measure_domain_runtime + determine_split_domtree, src/domains.c:21-38, 86-160; the load is
    DTIME_FRACTION = idxP2P + idxM2L of the rank, normalised by the mean: src/fmm.c:1069, src/photoNs.c:277-283).
    `splits` is what pn2_migrate_begin / pn2_domain_owner_device take."""

    def __init__(self, P, box):
        self.P, self.box = int(P), float(box)
        self.splits = domain_tree(P, box)[0]

    def update(self, load):
        """load[r] = DTIME_FRACTION of rank r.  Returns the new splits."""
        P = self.P
        load = np.asarray(load, np.float64)
        assert load.shape == (P,)
        n = 2 * P - 1
        t_node, t_left, t_right = np.zeros(n), np.zeros(n), np.zeros(n)
        for r in range(P):
            t_node[domain_node_of_rank(r, P)] = load[r]

        def fill(node):                                   # fill_time_domtree, src/domains.c:5-19
            if node >= P - 1:
                return t_node[node]
            t_left[node] = fill(2 * node + 1)
            t_right[node] = fill(2 * node + 2)
            t_node[node] = t_left[node] + t_right[node]
            return t_node[node]

        fill(0)
        relax = 0.3
        splits = self.splits

        def rec(D, nproc, node, bl, br):                  # determine_split_node, src/domains.c:86-144
            if node >= P - 1:
                return
            nleft, nright = fraction(nproc)
            t1 = t_left[node] / nleft
            t2 = t_right[node] / nright
            w0l = splits[node] - bl[D]
            w0r = br[D] - splits[node]
            shift = 0.5 * relax * (t2 - t1) / (t1 * nleft / w0l + t2 * nright / w0r)     # :119 (the branches above it are dead)
            new_split = splits[node] + shift
            b_l, b_r = list(bl), list(br)
            b_r[D] = splits[node]
            rec((D + 1) % 3, nleft, 2 * node + 1, b_l, b_r)
            b_l[D] = splits[node]
            b_r[D] = br[D]
            rec((D + 1) % 3, nright, 2 * node + 2, b_l, b_r)
            splits[node] = new_split

        rec(0, P, 0, [0.0, 0.0, 0.0], [self.box] * 3)
        return splits

    def boxes(self):
        """pn2gpu.Domain of every rank for the current splits."""
        return _domains_from(self.P, *boxes_from_splits(self.P, self.box, self.splits))

    def owner(self, pos):
        return owner_of(pos, self.P, self.splits)


def _domains_from(P, lo, hi, direct):
    out = []
    for r in range(P):
        dn = domain_node_of_rank(r, P)
        c = 0.5 * (hi[dn] + lo[dn])
        w = hi[dn] - lo[dn]
        out.append(pn2gpu.make_domain(c - 0.5 * w, c + 0.5 * w, int(direct[dn])))
    return out


def domain_boxes(P, box):
    """pn2gpu.Domain of every rank (box corners as centre -/+ width/2, like the reference's toptree: src/fmm.c:194-197)."""
    splits, lo, hi, direct = domain_tree(P, box)
    out = []
    for r in range(P):
        dn = domain_node_of_rank(r, P)
        c = 0.5 * (hi[dn] + lo[dn])
        w = hi[dn] - lo[dn]
        out.append(pn2gpu.make_domain(c - 0.5 * w, c + 0.5 * w, int(direct[dn])))
    return out


def domain_of(pos, P, box):
    """Owner rank of every position under the initial splits."""
    return owner_of(pos, P, domain_tree(P, box)[0])


def owner_of(pos, P, splits):
    """Owner rank of every position: pos > split goes right (src/domains.c:163-296).  pos: torch tensor or numpy (N,3)."""
    is_torch = not isinstance(pos, np.ndarray)
    n = pos.shape[0]
    if is_torch:
        import torch
        node = torch.zeros(n, dtype=torch.int64, device=pos.device)
        spl = torch.tensor(splits, dtype=pos.dtype, device=pos.device)
    else:
        node = np.zeros(n, np.int64)
        spl = splits
    dim = 0
    # walk down the heap until every particle sits on a domain (node >= P - 1); depth <= ceil(log2 P) + 1
    for _ in range(max(1, P).bit_length() + 1):
        if is_torch:
            internal = node < P - 1
            s = spl[torch.clamp(node, max=2 * P - 2)]
            nxt = 2 * node + 1 + (pos[:, dim] > s).to(torch.int64)
            node = torch.where(internal, nxt, node)
        else:
            internal = node < P - 1
            s = spl[np.minimum(node, 2 * P - 2)]
            nxt = 2 * node + 1 + (pos[:, dim] > s).astype(np.int64)
            node = np.where(internal, nxt, node)
        dim = (dim + 1) % 3
    ml = mostleft(P)
    return (node - ml + P) % P
