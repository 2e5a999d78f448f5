// pn2_walk.cu -- Mode B: interaction-list construction fused with the P2P evaluation.
//
// One warp owns one sink leaf l.  It replays the reference's dual-tree walk (walk_task_p2p /
// walk_task_m2l, src/fmm.c:406-712; remote version walk_task_*_ext, src/remotes.c:213-552) RESTRICTED to
// sink-side cells on the path root -> l: whenever the reference would open the sink side, only the child
// that contains l is followed.  Every decision is taken on the same (im, jm) pair with the same
// acceptance() arithmetic (src/fmm.c:267-326, FP64, same expression order, no FMA contraction: this file
// is compiled with -fmad=false), so the set of source leaves found for l is exactly the set of
// (source, l) pairs the reference's walk emits.  An M2L pair (jm -> im) is met by every leaf below im;
// it is emitted once, by the leaf that is the left-most descendant of im.
//
// Periodic images (src/fmm.c:1028-1045) and, on one rank, the self-exchange of 26 displaced pruned
// trees are walked in place: image k is the local tree displaced by shift[k]; the sender-side pruning
// of prepare_sendtree2 (src/remotes.c:97-158) is evaluated on the fly for the node being visited.
//
// 32 (im, jm) pairs are popped per iteration from a per-warp stack in shared memory (one per lane); the
// source leaves found go to a small queue which the same warp drains through the staged P2P pipeline of
// pn2_p2p.cuh.  Nothing is written to HBM except accelerations (and the rare M2L pairs).
#include "pn2_p2p.cuh"

#define WALK_WARPS 4
#define STACK_CAP 768
#define STACK_SOFT 640
#define SRCQ_CAP 64
#define MAX_DEPTH 64

struct WalkArgs {
    int nleaf, ncell, root;
    const double *geom;
    const int *son;
    const LeafDesc *desc;
    const int *parent;
    const int *depth;
    const float4 *rel;
    const double *pos;
    double *acc;
    double cutoff, theta;
    int longshort, nimg, maxleaf;
    double tc[3], tw[3];               // this rank's domain box (pruning target for image trees)
    unsigned *m2l_t, *m2l_s;
    unsigned long long m2l_cap;
    unsigned long long *counters;      // [0] interactions, [1] m2l pairs, [2] p2p leaf pairs, [3] error flags
    long *lst_off;                     // dump mode
    unsigned *lst_src;
    int pass, emit_m2l;
};

// acceptance(), src/fmm.c:267-326: 0 open, 1 accept, -1 drop
__device__ __forceinline__ int accept_dev(const double *wi, const double *wj, double dx, double dy, double dz,
                                          double cutoff, double theta, int longshort) {
    double w0 = (wi[0] + wj[0]) * 0.5, w1 = (wi[1] + wj[1]) * 0.5, w2 = (wi[2] + wj[2]) * 0.5;
    double dd2 = dx * dx + dy * dy + dz * dz;
    double g0 = fabs(dx) - w0, g1 = fabs(dy) - w1, g2 = fabs(dz) - w2;
    if (g0 <= 0.0) g0 = 0.0;
    if (g1 <= 0.0) g1 = 0.0;
    if (g2 <= 0.0) g2 = 0.0;
    if (g0 + g1 + g2 < 0.0001) return 0;
    double dm2 = g0 * g0 + g1 * g1 + g2 * g2;
    if (longshort) {
        double c2 = cutoff * cutoff;
        if (dm2 >= c2) return -1;
        if (dd2 > 1.0 * c2) return 0;
    }
    double wmax = w0;
    if (w1 > wmax) wmax = w1;
    if (w2 > wmax) wmax = w2;
    wmax *= 2;
    return (wmax * wmax < theta * theta * dd2) ? 1 : 0;
}

// prepare_sendtree2's pruning test for a node displaced by sh (src/remotes.c:97-158): 1 = terminal
__device__ __forceinline__ int pruned_dev(const double *c, const double *w, const double *sh, const double *tc,
                                          const double *tw, double cutoff, double theta, int longshort) {
    double dr = 0.0;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double g = tc[d] - c[d] - sh[d];
        if (g < 0.0) g = -g;
        g -= (tw[d] + w[d]) * 0.5;
        if (g > 0.0) dr += g * g;
    }
    dr = sqrt(dr);
    double wmax = w[0];
    if (wmax < w[1]) wmax = w[1];
    if (wmax < w[2]) wmax = w[2];
    if (longshort && dr >= cutoff) return 1;
    if (wmax < 0.95 * theta * dr) return 1;
    return 0;
}

template <int SW, int MODE>     // MODE 0: FP32 P2P, 1: FP64 P2P, 2: dump lists (no arithmetic)
__global__ void __launch_bounds__(WALK_WARPS * 32)
walk_fused_kernel(WalkArgs a, P2PConst pc) {
    using ST = P2PStageF32<SW>;
    constexpr int NSL = 32 / SW;
    __shared__ uint2 s_stack[WALK_WARPS][STACK_CAP];
    __shared__ unsigned s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ int s_anc[WALK_WARPS][MAX_DEPTH];
    __shared__ float4 s_stage[WALK_WARPS][2][ST::STAGE_F4];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int leaf = blockIdx.x * WALK_WARPS + wib;
    if (leaf >= a.nleaf) return;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint2 *stack = s_stack[wib];
    unsigned *srcq = s_srcq[wib];
    int *anc = s_anc[wib];

    // ---- ancestor chain and the depth from which this leaf is the designated M2L emitter ----
    const int ldepth = a.depth[leaf];
    int lm_depth = 0;
    if (lane == 0) {
        int c = leaf, d = ldepth, lm = ldepth;
        bool left = true;
        while (c >= 0) {
            anc[d] = c;
            int p = a.parent[c];
            if (p >= 0 && left) {
                if (a.son[2 * (size_t)p] == c) lm = d - 1; else left = false;
            }
            c = p; d--;
        }
        lm_depth = lm;
    }
    lm_depth = __shfl_sync(0xffffffffu, lm_depth, 0);
    __syncwarp();

    // ---- sink particles ----
    const int q = lane / SW, j = lane % SW;
    const LeafDesc sd = a.desc[leaf];
    float xi = 0.f, yi = 0.f, zi = 0.f, ax = 0.f, ay = 0.f, az = 0.f;
    double xd = 0, yd = 0, zd = 0, axd = 0, ayd = 0, azd = 0;
    if (MODE == 0 && j < sd.npart) { float4 p = a.rel[sd.first + j]; xi = p.x; yi = p.y; zi = p.z; }
    if (MODE == 1 && j < sd.npart) { const double *p = a.pos + 3 * (size_t)(sd.first + j); xd = p[0]; yd = p[1]; zd = p[2]; }
    float qc[PN2_RDEG + 1];
#pragma unroll
    for (int k = 0; k <= PN2_RDEG; k++) qc[k] = pc.q[k];
    const float inv_eps = pc.inv_eps;
    unsigned long long nint = 0;
    long npairs = 0;
    long dump_pos = (MODE == 2 && a.pass == 1) ? a.lst_off[leaf] : 0;

    // ---- drain n entries of the source queue starting at qhead ----
    int qhead = 0, qtail = 0, buf = 0;
    auto load_stage = [&](int pos_, int limit) -> float4 {
        float4 p = make_float4(PN2_PAD_COORD, PN2_PAD_COORD, PN2_PAD_COORD, 0.f);
        int idx = pos_ + q;
        if (idx < limit) {
            unsigned e = srcq[idx & (SRCQ_CAP - 1)];
            unsigned cell = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
            LeafDesc d = a.desc[cell];
            if (j < d.npart) {
                float4 r = a.rel[d.first + j];
                float Dx = (float)(((d.c[0] + pc.shift[img][0]) - sd.c[0]) * pc.inv2rs);
                float Dy = (float)(((d.c[1] + pc.shift[img][1]) - sd.c[1]) * pc.inv2rs);
                float Dz = (float)(((d.c[2] + pc.shift[img][2]) - sd.c[2]) * pc.inv2rs);
                p = make_float4(r.x + Dx, r.y + Dy, r.z + Dz, 1.f);
            }
            if (j == 0) nint += (unsigned long long)(d.npart - ((e == (unsigned)leaf) ? 1 : 0));
        }
        return p;
    };
    auto drain = [&](int limit) {          // consumes [qhead, limit)
        if (MODE == 0) {
            float4 pnext = load_stage(qhead, limit);
            for (int base = qhead; base < limit; base += NSL) {
                s_stage[wib][buf][q * ST::ROW + j] = pnext;
                __syncwarp();
                if (base + NSL < limit) pnext = load_stage(base + NSL, limit);
                const float4 *row = &s_stage[wib][buf][q * ST::ROW];
                if (pc.longshort) {
#pragma unroll
                    for (int k = 0; k < SW; k++) p2p_interact_f32<true>(row[k], xi, yi, zi, ax, ay, az, qc, inv_eps);
                } else {
#pragma unroll
                    for (int k = 0; k < SW; k++) p2p_interact_f32<false>(row[k], xi, yi, zi, ax, ay, az, qc, inv_eps);
                }
                buf ^= 1;
            }
        } else if (MODE == 1) {
            for (int idx = qhead + q; idx < limit; idx += NSL) {
                unsigned e = srcq[idx & (SRCQ_CAP - 1)];
                unsigned cell = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
                LeafDesc d = a.desc[cell];
                const double sx = pc.shift[img][0], sy = pc.shift[img][1], sz = pc.shift[img][2];
                for (int k = 0; k < d.npart; k++) {
                    const double *p = a.pos + 3 * (size_t)(d.first + k);
                    p2p_interact_f64(p[0] + sx, p[1] + sy, p[2] + sz, pc.mass, xd, yd, zd, axd, ayd, azd, pc.soft,
                                     pc.inv2rs, pc.longshort);
                }
                if (j == 0) nint += (unsigned long long)(d.npart - ((e == (unsigned)leaf) ? 1 : 0));
            }
        } else {
            if (a.pass == 1)
                for (int idx = qhead + lane; idx < limit; idx += 32) a.lst_src[dump_pos + (idx - qhead)] = srcq[idx & (SRCQ_CAP - 1)];
            dump_pos += limit - qhead;
        }
        npairs += limit - qhead;
        qhead = limit;
        __syncwarp();
    };

    // ---- the walk ----
    int ssize = 0;
    if (lane < a.nimg) stack[lane] = make_uint2((unsigned)a.root, (unsigned)a.root | ((unsigned)lane << PN2_IMG_SHIFT));
    ssize = a.nimg;
    __syncwarp();
    int err = 0;

    while (ssize > 0) {
        int k = ssize < 32 ? ssize : 32;
        if (ssize > STACK_SOFT) k = 1;                       // depth-first when the stack is nearly full
        const int sbase = ssize - k;
        int npush = 0, emit_p = 0, emit_m = 0;
        unsigned p_im0 = 0, p_jm0 = 0, p_im1 = 0, p_jm1 = 0, m_src = 0, m_snk = 0, p_src = 0;
        if (lane < k) {
            uint2 ent = stack[sbase + lane];
            const int im = (int)ent.x;
            const unsigned jme = ent.y;
            const int jm = (int)(jme & PN2_CELL_MASK);
            const unsigned img = jme >> PN2_IMG_SHIFT, imgbits = jme & ~PN2_CELL_MASK;
            const bool li = im < a.nleaf, lj = jm < a.nleaf;
            if (img == 0 && im == jm) {
                if (li) { emit_p = 1; p_src = jme; }
                else {
                    int t = anc[a.depth[im] + 1];
                    npush = 2;
                    p_im0 = p_im1 = (unsigned)t;
                    p_jm0 = (unsigned)a.son[2 * (size_t)jm]; p_jm1 = (unsigned)a.son[2 * (size_t)jm + 1];
                }
            } else if (li && lj) {
                emit_p = 1; p_src = jme;
            } else {
                const double *gi = a.geom + 6 * (size_t)im, *gj = a.geom + 6 * (size_t)jm;
                double ci[3] = {gi[0], gi[1], gi[2]}, wi[3] = {gi[3], gi[4], gi[5]};
                double cj[3] = {gj[0], gj[1], gj[2]}, wj[3] = {gj[3], gj[4], gj[5]};
                int pruned = 0;
                if (img != 0) {
                    if (!lj) pruned = pruned_dev(cj, wj, pc.shift[img], a.tc, a.tw, a.cutoff, a.theta, a.longshort);
                    cj[0] += pc.shift[img][0]; cj[1] += pc.shift[img][1]; cj[2] += pc.shift[img][2];   // src/remotes.c:73-75
                }
                int f = accept_dev(wi, wj, ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2], a.cutoff, a.theta, a.longshort);
                if (f == -1) {
                } else if (f == 1 || (img != 0 && li && pruned)) {
                    // M2L jm -> im (a pruned remote node met by a local leaf is forced: src/remotes.c:442)
                    if (a.depth[im] >= lm_depth) { emit_m = 1; m_src = jme; m_snk = (unsigned)im; }
                } else {
                    bool open_i;
                    if (li) open_i = false;
                    else if (lj) open_i = true;
                    else open_i = (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) || (img != 0 && pruned);
                    if (open_i) {
                        npush = 1;
                        p_im0 = (unsigned)anc[a.depth[im] + 1]; p_jm0 = jme;
                    } else {
                        npush = 2;
                        p_im0 = p_im1 = (unsigned)im;
                        p_jm0 = (unsigned)a.son[2 * (size_t)jm] | imgbits; p_jm1 = (unsigned)a.son[2 * (size_t)jm + 1] | imgbits;
                    }
                }
            }
        }
        __syncwarp();
        // pushes: second sons first so that the first son is popped first is not required (set semantics)
        const unsigned m1 = __ballot_sync(0xffffffffu, npush >= 1), m2 = __ballot_sync(0xffffffffu, npush == 2);
        int pos0 = sbase + __popc(m1 & lt_mask) + __popc(m2 & lt_mask);
        const int newsize = sbase + __popc(m1) + __popc(m2);
        if (newsize > STACK_CAP) { err = 1; break; }
        if (npush >= 1) stack[pos0] = make_uint2(p_im0, p_jm0);
        if (npush == 2) stack[pos0 + 1] = make_uint2(p_im1, p_jm1);
        ssize = newsize;
        // P2P sources -> queue
        const unsigned mp = __ballot_sync(0xffffffffu, emit_p);
        if (emit_p) srcq[(qtail + __popc(mp & lt_mask)) & (SRCQ_CAP - 1)] = p_src;
        qtail += __popc(mp);
        // M2L pairs -> global list
        const unsigned mm = __ballot_sync(0xffffffffu, emit_m);
        if (mm && a.emit_m2l) {
            unsigned long long basei = 0;
            if (lane == 0) basei = atomicAdd(&a.counters[1], (unsigned long long)__popc(mm));
            basei = __shfl_sync(0xffffffffu, basei, 0);
            if (emit_m) {
                unsigned long long at = basei + __popc(mm & lt_mask);
                if (at < a.m2l_cap) { a.m2l_t[at] = m_snk; a.m2l_s[at] = m_src; }
            }
        }
        __syncwarp();
        if (qtail - qhead >= 32) drain(qhead + ((qtail - qhead) / NSL) * NSL);
    }
    if (qtail > qhead) drain(qtail);
    if (err) { if (lane == 0) atomicOr(&a.counters[3], 1ULL); return; }

    // ---- results ----
    if (MODE == 0) {
#pragma unroll
        for (int m = SW; m < 32; m <<= 1) {
            ax += __shfl_xor_sync(0xffffffffu, ax, m);
            ay += __shfl_xor_sync(0xffffffffu, ay, m);
            az += __shfl_xor_sync(0xffffffffu, az, m);
        }
        if (q == 0 && j < sd.npart) {
            const double sc = pc.mass * pc.inv2rs * pc.inv2rs;
            double *o = a.acc + 3 * (size_t)(sd.first + j);
            o[0] += (double)ax * sc; o[1] += (double)ay * sc; o[2] += (double)az * sc;
        }
    } else if (MODE == 1) {
#pragma unroll
        for (int m = SW; m < 32; m <<= 1) {
            axd += __shfl_xor_sync(0xffffffffu, axd, m);
            ayd += __shfl_xor_sync(0xffffffffu, ayd, m);
            azd += __shfl_xor_sync(0xffffffffu, azd, m);
        }
        if (q == 0 && j < sd.npart) {
            double *o = a.acc + 3 * (size_t)(sd.first + j);
            o[0] += axd; o[1] += ayd; o[2] += azd;
        }
    } else {
        if (a.pass == 0 && lane == 0) a.lst_off[leaf] = npairs;
    }
    if (MODE != 2) {
#pragma unroll
        for (int m = SW; m < 32; m <<= 1) nint += __shfl_xor_sync(0xffffffffu, nint, m);
        if (lane == 0) {
            atomicAdd(&a.counters[0], nint * (unsigned long long)sd.npart);
            atomicAdd(&a.counters[2], (unsigned long long)npairs);
        }
    }
}

template <int SW>
static void launch_mode(pn2_ctx *h, const WalkArgs &a, int mode) {
    unsigned grid = (unsigned)((a.nleaf + WALK_WARPS - 1) / WALK_WARPS);
    if (mode == 0) walk_fused_kernel<SW, 0><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else if (mode == 1) walk_fused_kernel<SW, 1><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else walk_fused_kernel<SW, 2><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    h->launches++;
}

// dump = 0: the product step (P2P evaluated, M2L pairs appended to h->m2l_pairs);
// dump = 1 / 2: list dump passes (count / fill) for pn2_get_lists
int pn2_walk_fused(pn2_ctx *h, int dump) {
    if (h->nleaf == 0) return PN2_OK;
    if (h->nlevel + 1 >= MAX_DEPTH) { pn2_set_error("pn2: tree depth %d exceeds %d", h->nlevel, MAX_DEPTH - 2); return PN2_ERR_ARG; }
    WalkArgs a;
    memset(&a, 0, sizeof a);
    a.nleaf = h->nleaf; a.ncell = h->ncell; a.root = h->nleaf;
    a.geom = h->geom.p; a.son = h->son.p; a.desc = h->desc.p; a.parent = h->parent.p; a.depth = h->depth.p;
    a.rel = h->rel.p; a.pos = h->pos.p; a.acc = h->acc.p;
    a.cutoff = h->prm.cutoff; a.theta = h->prm.theta; a.longshort = h->prm.longshort; a.maxleaf = h->prm.maxleaf;
    a.nimg = h->prm.periodic ? 27 : 1;
    for (int d = 0; d < 3; d++) {
        // the pruning box of prepare_sendtree2 is the target's box as centre / width (src/remotes.c:97-110)
        a.tc[d] = 0.5 * (h->dom.hi[d] + h->dom.lo[d]);
        a.tw[d] = h->dom.hi[d] - h->dom.lo[d];
    }
    a.m2l_t = h->m2l_pairs.p; a.m2l_s = h->m2l_pairs.p + h->m2l_cap; a.m2l_cap = h->m2l_cap;
    a.counters = h->counters.p;
    a.lst_off = h->lst_off.p; a.lst_src = h->lst_src.p;
    a.pass = dump == 2 ? 1 : 0;
    a.emit_m2l = dump == 0;
    int mode = dump ? 2 : (h->prm.precision == PN2_FP64 ? 1 : 0);
    int ml = h->prm.maxleaf;
    if (ml <= 8) launch_mode<8>(h, a, mode);
    else if (ml <= 16) launch_mode<16>(h, a, mode);
    else launch_mode<32>(h, a, mode);
    KERNEL_CHECK();
    return PN2_OK;
}
