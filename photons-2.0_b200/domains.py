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
    """Owner rank of every position: pos > split goes right (src/domains.c:163-296).  pos: torch tensor or numpy (N,3)."""
    splits, _, _, _ = domain_tree(P, box)
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
