// pn2_operators.cu -- P2M / M2M / M2L / L2L / L2P kernels (FP64) over the cell arrays.
//
// These are <1 % of a force step; they are HBM-bound streaming kernels over
// 160-byte multipole / local-expansion records, one thread per cell, level-synchronous where the
// reference recurses (walk_m2m: src/operator.c:165-194, walk_l2l: src/operator.c:498-528).
#include "pn2_operators.cuh"

using namespace pn2op;

// ---- P2M: one thread per leaf (src/fmm.c:741-742 -> src/operator.c:13-93) ----
__global__ void p2m_kernel(int nleaf, const LeafDesc *__restrict__ desc, const double *__restrict__ pos, double mass,
                           double *__restrict__ M) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nleaf) return;
    LeafDesc d = desc[k];
    double m[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) m[i] = 0.0;
    for (int p = d.first; p < d.first + d.npart; p++)
        p2m_add(pos[3 * p] - d.c[0], pos[3 * p + 1] - d.c[1], pos[3 * p + 2] - d.c[2], mass, m);
#pragma unroll
    for (int i = 0; i < NM; i++) M[(size_t)k * NM + i] = m[i];
}

// ---- M2M: one thread per node of one depth; children are finished (deeper levels ran first) ----
__global__ void m2m_level_kernel(int cnt, const int *__restrict__ nodes, const int *__restrict__ son,
                                 const double *__restrict__ geom, double *__restrict__ M) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    int c = nodes[k];
    double m[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) m[i] = 0.0;
    double cx = geom[6 * (size_t)c], cy = geom[6 * (size_t)c + 1], cz = geom[6 * (size_t)c + 2];
    for (int s = 0; s < 2; s++) {
        int ch = son[2 * (size_t)c + s];
        if (ch < 0) continue;
        double cm[NM];
#pragma unroll
        for (int i = 0; i < NM; i++) cm[i] = M[(size_t)ch * NM + i];
        m2m_add(cx - geom[6 * (size_t)ch], cy - geom[6 * (size_t)ch + 1], cz - geom[6 * (size_t)ch + 2], cm, m);
    }
#pragma unroll
    for (int i = 0; i < NM; i++) M[(size_t)c * NM + i] = m[i];
}

// ---- L2L: one thread per node of one depth, pushes its L into both children ----
__global__ void l2l_level_kernel(int cnt, const int *__restrict__ nodes, const int *__restrict__ son,
                                 const double *__restrict__ geom, double *__restrict__ L) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    int c = nodes[k];
    double l[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) l[i] = L[(size_t)c * NM + i];
    double cx = geom[6 * (size_t)c], cy = geom[6 * (size_t)c + 1], cz = geom[6 * (size_t)c + 2];
    for (int s = 0; s < 2; s++) {
        int ch = son[2 * (size_t)c + s];
        if (ch < 0) return;                         // src/operator.c:524-525 returns at the first missing son
        double cl[NM];
#pragma unroll
        for (int i = 0; i < NM; i++) cl[i] = L[(size_t)ch * NM + i];
        l2l_add(geom[6 * (size_t)ch] - cx, geom[6 * (size_t)ch + 1] - cy, geom[6 * (size_t)ch + 2] - cz, l, cl);
#pragma unroll
        for (int i = 0; i < NM; i++) L[(size_t)ch * NM + i] = cl[i];
    }
}

// ---- L2P: one thread per leaf (src/fmm.c:1056-1057 -> src/operator.c:197-251) ----
__global__ void l2p_kernel(int nleaf, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                           const double *__restrict__ L, double *__restrict__ acc) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nleaf) return;
    LeafDesc d = desc[k];
    double l[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) l[i] = L[(size_t)k * NM + i];
    for (int p = d.first; p < d.first + d.npart; p++) {
        double a[3];
        l2p_eval(pos[3 * p] - d.c[0], pos[3 * p + 1] - d.c[1], pos[3 * p + 2] - d.c[2], l, a);
        acc[3 * p] += a[0]; acc[3 * p + 1] += a[1]; acc[3 * p + 2] += a[2];
    }
}

// ---- M2L: one thread per sink segment, sources in list order (deterministic, no atomics) ----
// task_compute_m2l (src/fmm.c:875-907) / task_compute_m2l_ext (src/remotes.c:598-628)
__global__ void m2l_csr_kernel(long nseg, const int *__restrict__ seg_sink, const long *__restrict__ seg_off,
                               const unsigned *__restrict__ src, const double *__restrict__ sink_geom,
                               const double *__restrict__ src_geom, const double *__restrict__ src_M,
                               double *__restrict__ L, P2PConst pc) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    int t = seg_sink[k];
    double cx = sink_geom[6 * (size_t)t], cy = sink_geom[6 * (size_t)t + 1], cz = sink_geom[6 * (size_t)t + 2];
    double l[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) l[i] = 0.0;
    for (long q = seg_off[k]; q < seg_off[k + 1]; q++) {
        unsigned e = src[q];
        unsigned s = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
        double m[NM];
#pragma unroll
        for (int i = 0; i < NM; i++) m[i] = src_M[(size_t)s * NM + i];
        double sx = src_geom[6 * (size_t)s] + pc.shift[img][0];
        double sy = src_geom[6 * (size_t)s + 1] + pc.shift[img][1];
        double sz = src_geom[6 * (size_t)s + 2] + pc.shift[img][2];
        m2l_add(cx - sx, cy - sy, cz - sz, m, l, pc.rs, pc.longshort);
    }
#pragma unroll
    for (int i = 0; i < NM; i++) L[(size_t)t * NM + i] += l[i];
}

static inline int nblk(long n, int b) { return (int)((n + b - 1) / b); }

int pn2_launch_p2m(pn2_ctx *h) {
    if (h->nleaf == 0) return PN2_OK;
    p2m_kernel<<<nblk(h->nleaf, 128), 128, 0, h->stream>>>(h->nleaf, h->desc.p, h->pos.p, h->prm.mass, h->M.p);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_m2m(pn2_ctx *h) {
    for (int lev = h->nlevel - 1; lev >= 0; lev--) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        m2m_level_kernel<<<nblk(cnt, 128), 128, 0, h->stream>>>(cnt, h->level_nodes.p + h->level_off[lev], h->son.p,
                                                                h->geom.p, h->M.p);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_l2l_l2p(pn2_ctx *h) {
    for (int lev = 0; lev < h->nlevel; lev++) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        l2l_level_kernel<<<nblk(cnt, 128), 128, 0, h->stream>>>(cnt, h->level_nodes.p + h->level_off[lev], h->son.p,
                                                                h->geom.p, h->L.p);
        h->launches++;
    }
    if (h->nleaf > 0) {
        l2p_kernel<<<nblk(h->nleaf, 128), 128, 0, h->stream>>>(h->nleaf, h->desc.p, h->pos.p, h->L.p, h->acc.p);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_m2l(pn2_ctx *h, const CsrList &list, const double *src_geom, const double *src_M) {
    if (list.nseg == 0) return PN2_OK;
    m2l_csr_kernel<<<nblk(list.nseg, 64), 64, 0, h->stream>>>(list.nseg, list.seg_sink, list.seg_off, list.src,
                                                              h->geom.p, src_geom, src_M, h->L.p, h->pc);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}
