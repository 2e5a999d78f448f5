// pn2_operators.cu -- P2M / M2M / M2L / L2L / L2P kernels (FP64) over the cell arrays.
//
// These are a few % of a force step; they are HBM-bound streaming kernels over
// 160-byte multipole / local-expansion records, one thread per cell (records moved by the warp, see
// warp_load_records), level-synchronous where the reference recurses (walk_m2m: src/operator.c:165-194, walk_l2l: src/operator.c:498-528).
#include "pn2_operators.cuh"

using namespace pn2op;

// The 160-byte M / L records are read and written by the WARP, not by the thread: lane l wants record idx_l; the
// warp moves the 32 records as 320 chunks of 16 bytes (lane-consecutive chunks are address-consecutive inside a
// record, and consecutive cells are consecutive records), through a padded shared-memory tile from which every lane
// picks its own 20 values.  A thread-per-record access pattern touches 32 different sectors per load instruction.
#define OP_WARPS 4
#define REC_STRIDE (NM + 1)               // doubles per record in shared memory (+1: bank spread)
__device__ __forceinline__ void warp_load_records(const double *__restrict__ base, int idx, double (&v)[NM], double *tile, int lane) {
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NM / 2; t++) {
        const int c = t * 32 + lane, r = c / (NM / 2), k = c % (NM / 2);
        const int ri = __shfl_sync(0xffffffffu, idx, r);
        if (ri >= 0) {
            const double2 d = *reinterpret_cast<const double2 *>(base + (size_t)ri * NM + 2 * k);
            tile[r * REC_STRIDE + 2 * k] = d.x;
            tile[r * REC_STRIDE + 2 * k + 1] = d.y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NM; i++) v[i] = idx >= 0 ? tile[lane * REC_STRIDE + i] : 0.0;
}
__device__ __forceinline__ void warp_store_records(double *__restrict__ base, int idx, const double (&v)[NM], double *tile, int lane) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NM; i++) tile[lane * REC_STRIDE + i] = v[i];
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NM / 2; t++) {
        const int c = t * 32 + lane, r = c / (NM / 2), k = c % (NM / 2);
        const int ri = __shfl_sync(0xffffffffu, idx, r);
        if (ri >= 0)
            *reinterpret_cast<double2 *>(base + (size_t)ri * NM + 2 * k) = make_double2(tile[r * REC_STRIDE + 2 * k], tile[r * REC_STRIDE + 2 * k + 1]);
    }
}

// ---- P2M: one thread per leaf (src/fmm.c:741-742 -> src/operator.c:13-93) ----
__global__ void __launch_bounds__(OP_WARPS * 32) p2m_kernel(int nleaf, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                                                            double mass, double *__restrict__ M) {
    __shared__ double s_tile[OP_WARPS][32 * REC_STRIDE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = k < nleaf;
    double m[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) m[i] = 0.0;
    if (on) {
        const LeafDesc d = desc[k];
        for (int p = d.first; p < d.first + d.npart; p++)
            p2m_add(pos[3 * (size_t)p] - d.c[0], pos[3 * (size_t)p + 1] - d.c[1], pos[3 * (size_t)p + 2] - d.c[2], mass, m);
    }
    warp_store_records(M, on ? k : -1, m, s_tile[wib], lane);
}

// ---- M2M: one thread per node of one depth; children are finished (deeper levels ran first) ----
__global__ void __launch_bounds__(OP_WARPS * 32) m2m_level_kernel(int cnt, const int *__restrict__ nodes, const int *__restrict__ son,
                                                                  const double *__restrict__ geom, double *__restrict__ M) {
    __shared__ double s_tile[OP_WARPS][32 * REC_STRIDE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = k < cnt;
    const int c = on ? nodes[k] : -1;
    double m[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) m[i] = 0.0;
    double cx = 0, cy = 0, cz = 0;
    int2 ch = make_int2(-1, -1);
    if (on) {
        cx = geom[6 * (size_t)c]; cy = geom[6 * (size_t)c + 1]; cz = geom[6 * (size_t)c + 2];
        ch = *reinterpret_cast<const int2 *>(son + 2 * (size_t)c);
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int cs = s == 0 ? ch.x : ch.y;
        double cm[NM];
        warp_load_records(M, cs, cm, s_tile[wib], lane);
        if (cs >= 0) m2m_add(cx - geom[6 * (size_t)cs], cy - geom[6 * (size_t)cs + 1], cz - geom[6 * (size_t)cs + 2], cm, m);
    }
    warp_store_records(M, c, m, s_tile[wib], lane);
}

// ---- L2L: one thread per node of one depth, pushes its L into both children ----
// has_l (Mode B; may be NULL = every cell carries an expansion): 1 for the cells whose L has been written this step
// (M2L sinks and everything below them).  A node without an expansion is skipped -- no loads, no stores --, a child
// receives its first contribution as a plain store (the L array is never cleared), so the downward pass costs what the
// M2L list makes it cost: next to nothing at NSIDE = particle side, where the walk finds a few hundred M2L pairs.
__global__ void __launch_bounds__(OP_WARPS * 32) l2l_level_kernel(int cnt, const int *__restrict__ nodes, const int *__restrict__ son,
                                                                  const double *__restrict__ geom, double *__restrict__ L,
                                                                  unsigned char *__restrict__ has_l) {
    __shared__ double s_tile[OP_WARPS][32 * REC_STRIDE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = k < cnt;
    int c = on ? nodes[k] : -1;
    if (c >= 0 && has_l && !has_l[c]) c = -1;
    if (!__any_sync(0xffffffffu, c >= 0)) return;
    double l[NM];
    warp_load_records(L, c, l, s_tile[wib], lane);
    double cx = 0, cy = 0, cz = 0;
    int2 ch = make_int2(-1, -1);
    if (c >= 0) {
        cx = geom[6 * (size_t)c]; cy = geom[6 * (size_t)c + 1]; cz = geom[6 * (size_t)c + 2];
        ch = *reinterpret_cast<const int2 *>(son + 2 * (size_t)c);
        if (ch.x < 0) ch.y = -1;                    // src/operator.c:524-525 returns at the first missing son
    }
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int cs = s == 0 ? ch.x : ch.y;
        const bool had = cs >= 0 && (!has_l || has_l[cs]);
        double cl[NM];
        warp_load_records(L, had ? cs : -1, cl, s_tile[wib], lane);          // no expansion yet: starts from zero
        if (cs >= 0) l2l_add(geom[6 * (size_t)cs] - cx, geom[6 * (size_t)cs + 1] - cy, geom[6 * (size_t)cs + 2] - cz, l, cl);
        warp_store_records(L, cs, cl, s_tile[wib], lane);
        if (cs >= 0 && has_l) has_l[cs] = 1;
    }
}

// ---- L2P: one thread per leaf (src/fmm.c:1056-1057 -> src/operator.c:197-251) ----
__global__ void __launch_bounds__(OP_WARPS * 32) l2p_kernel(int nleaf, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                                                            const double *__restrict__ L, double *__restrict__ acc,
                                                            const unsigned char *__restrict__ has_l) {
    __shared__ double s_tile[OP_WARPS][32 * REC_STRIDE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = k < nleaf && (!has_l || has_l[k]);                   // a leaf without an expansion gets exactly 0 from L2P
    if (!__any_sync(0xffffffffu, on)) return;
    double l[NM];
    warp_load_records(L, on ? k : -1, l, s_tile[wib], lane);
    if (!on) return;
    const LeafDesc d = desc[k];
    for (int p = d.first; p < d.first + d.npart; p++) {
        double a[3];
        l2p_eval(pos[3 * (size_t)p] - d.c[0], pos[3 * (size_t)p + 1] - d.c[1], pos[3 * (size_t)p + 2] - d.c[2], l, a);
        acc[3 * (size_t)p] += a[0]; acc[3 * (size_t)p + 1] += a[1]; acc[3 * (size_t)p + 2] += a[2];
    }
}

// ---- M2L: LPS lanes per sink segment (a whole warp, or 8 lanes for short lists), sources in parallel over the lanes ----
// task_compute_m2l (src/fmm.c:875-907) / task_compute_m2l_ext (src/remotes.c:598-628).  Every lane evaluates one
// (sink, source) pair per round into its own 20 accumulators; the 160-byte M records of the 32 sources of a round are
// moved by the warp through shared memory (warp_load_records: coalesced 16-byte chunks); the lanes of a sink are
// summed with xor shuffles in a fixed order, so the result is deterministic (no atomics), and written as one
// coalesced 160-byte record.  (The first version, one thread per sink with a serial source loop, read the M records
// with 32 scattered sectors per load instruction and left a whole warp waiting for its longest list.)
template <int LPS>
__global__ void __launch_bounds__(OP_WARPS * 32) m2l_warp_kernel(long nseg, const int *__restrict__ seg_sink, const long *__restrict__ seg_off,
                                                                 const unsigned *__restrict__ src, const double *__restrict__ sink_geom,
                                                                 const double *__restrict__ src_geom, const double *__restrict__ src_M,
                                                                 double *__restrict__ L, P2PConst pc, unsigned char *__restrict__ has_l) {
    constexpr int SPW = 32 / LPS;                          // sinks per warp
    __shared__ double s_tile[OP_WARPS][32 * REC_STRIDE];
    __shared__ double s_tabE[PN2_M2LTAB_DEG + 1][PN2_M2LTAB_KPAD], s_tabX[PN2_M2LTAB_DEG + 1][PN2_M2LTAB_KPAD];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = lane / LPS, gl = lane % LPS;
    if (pc.longshort) {                                    // the tables of erfc(u) and exp(-u^2)/sqrt(pi), once per CTA (the grid strides)
        for (int i = threadIdx.x; i < (PN2_M2LTAB_DEG + 1) * PN2_M2LTAB_KPAD; i += blockDim.x) {
            (&s_tabE[0][0])[i] = (&PN2_M2LTAB_E[0][0])[i];
            (&s_tabX[0][0])[i] = (&PN2_M2LTAB_X[0][0])[i];
        }
        __syncthreads();
    }
    const pn2op::M2LTab tab = {s_tabE, s_tabX};
    const long ngroup = (nseg + SPW - 1) / SPW;            // one warp per group of SPW sinks
    for (long grp = (long)blockIdx.x * OP_WARPS + wib; grp < ngroup; grp += (long)gridDim.x * OP_WARPS) {
    const long k = grp * SPW + g;
    const bool on = k < nseg;
    int t = 0;
    long o0 = 0, o1 = 0;
    double cx = 0, cy = 0, cz = 0;
    bool had = true;                                       // has_l: the first contribution of the step is stored, not added
    if (on) {
        t = seg_sink[k]; o0 = seg_off[k]; o1 = seg_off[k + 1];
        cx = sink_geom[6 * (size_t)t]; cy = sink_geom[6 * (size_t)t + 1]; cz = sink_geom[6 * (size_t)t + 2];
        if (has_l) had = has_l[t] != 0;
    }
    __syncwarp();                                          // every lane of the sink has read its flag before one of them sets it
    double l[NM];
#pragma unroll
    for (int i = 0; i < NM; i++) l[i] = 0.0;
    long len = o1 - o0;
#pragma unroll
    for (int m = LPS; m < 32; m <<= 1) { const long other = __shfl_xor_sync(0xffffffffu, len, m); len = other > len ? other : len; }   // warp-uniform round count
    for (long r = 0; r < len; r += LPS) {
        const long q = o0 + r + gl;
        const bool valid = on && q < o1;
        unsigned e = 0;
        if (valid) e = src[q];
        const unsigned sc = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
        double m[NM];
        warp_load_records(src_M, valid ? (int)sc : -1, m, s_tile[wib], lane);
        if (valid) {
            const double sx = src_geom[6 * (size_t)sc] + pc.shift[img][0];
            const double sy = src_geom[6 * (size_t)sc + 1] + pc.shift[img][1];
            const double sz = src_geom[6 * (size_t)sc + 2] + pc.shift[img][2];
            m2l_add(cx - sx, cy - sy, cz - sz, m, l, pc.rs, pc.longshort, &tab, pc.inv2rs);
        }
    }
    double mine = 0.0;                                     // lane gl < 20 of the group ends up with component gl
#pragma unroll
    for (int i = 0; i < NM; i++) {
        double v = l[i];
#pragma unroll
        for (int m = 1; m < LPS; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        if (LPS >= NM) { if (gl == i) mine = v; }
        else if (gl == i % LPS) { if (i < LPS) mine = v; else if (on) L[(size_t)t * NM + i] = had ? L[(size_t)t * NM + i] + v : v; }
    }
    if (on && gl < (LPS >= NM ? NM : LPS)) L[(size_t)t * NM + gl] = had ? L[(size_t)t * NM + gl] + mine : mine;
    if (on && gl == 0 && has_l) has_l[t] = 1;
    __syncwarp();
    }   // sink groups of this warp
}

static inline int nblk(long n, int b) { return (int)((n + b - 1) / b); }

int pn2_launch_p2m(pn2_ctx *h) {
    if (h->nleaf == 0) return PN2_OK;
    p2m_kernel<<<nblk(h->nleaf, OP_WARPS * 32), OP_WARPS * 32, 0, h->stream>>>(h->nleaf, h->desc.p, h->pos.p, h->prm.mass, h->M.p);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_m2m(pn2_ctx *h) {
    for (int lev = h->nlevel - 1; lev >= 0; lev--) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        m2m_level_kernel<<<nblk(cnt, OP_WARPS * 32), OP_WARPS * 32, 0, h->stream>>>(cnt, h->level_nodes.p + h->level_off[lev], h->son.p,
                                                                h->geom.p, h->M.p);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_l2l_l2p(pn2_ctx *h) {
    unsigned char *flags = h->use_lflags ? h->has_l.p : nullptr;
    for (int lev = 0; lev < h->nlevel; lev++) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        l2l_level_kernel<<<nblk(cnt, OP_WARPS * 32), OP_WARPS * 32, 0, h->stream>>>(cnt, h->level_nodes.p + h->level_off[lev], h->son.p,
                                                                h->geom.p, h->L.p, flags);
        h->launches++;
    }
    if (h->nleaf > 0) {
        l2p_kernel<<<nblk(h->nleaf, OP_WARPS * 32), OP_WARPS * 32, 0, h->stream>>>(h->nleaf, h->desc.p, h->pos.p, h->L.p, h->acc.p, flags);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_m2l(pn2_ctx *h, const CsrList &list, const double *src_geom, const double *src_M) {
    if (list.nseg == 0) return PN2_OK;
    unsigned char *flags = h->use_lflags ? h->has_l.p : nullptr;
    const long npair = list.npair > 0 ? list.npair : 32 * list.nseg;
    // the grid strides over the sink groups (the tables of the split functions are loaded once per CTA): at most 16 waves of CTAs
    const int gmax = (h->sm_count > 0 ? h->sm_count : 148) * 4 * 16;
    // lanes per sink by the mean list length: a warp per sink for long lists (clustered sets, NSIDE << particle side), 8 or 2
    // lanes for shorter ones, so that the per-sink work (20 accumulators cleared, reduced over the sink's lanes and stored) is
    // shared by 4 / 16 sinks per warp and fewer lanes idle behind the end of a short list (PN2_M2L_LPS = 2, 4, 8, 32 forces a width)
    // (measured at 256^3 / NSIDE 128, 6 <= mean length < 24: 0.97 ms with 2 lanes, 0.98 with 4, 1.08 with 8)
    int lps = npair >= 24 * list.nseg ? 32 : (npair >= 16 * list.nseg ? 8 : 2);
    if (const char *e = getenv("PN2_M2L_LPS")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8 || v == 32) lps = v; }
#define PN2_M2L_LAUNCH(W) do { const int grid = nblk(list.nseg, OP_WARPS * (32 / W)); \
        m2l_warp_kernel<W><<<grid < gmax ? grid : gmax, OP_WARPS * 32, 0, h->stream>>>(list.nseg, list.seg_sink, list.seg_off, list.src, \
                                                                                     h->geom.p, src_geom, src_M, h->L.p, h->pc, flags); } while (0)
    if (lps == 32) PN2_M2L_LAUNCH(32);
    else if (lps == 8) PN2_M2L_LAUNCH(8);
    else if (lps == 4) PN2_M2L_LAUNCH(4);
    else PN2_M2L_LAUNCH(2);
#undef PN2_M2L_LAUNCH
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}
