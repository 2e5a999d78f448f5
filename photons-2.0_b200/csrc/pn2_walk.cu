// pn2_walk.cu -- Mode B: interaction-list construction on the device, fused with the P2P evaluation.
//
// The reference builds its lists with a recursive dual-tree walk (walk_task_p2p / walk_task_m2l,
// src/fmm.c:406-712; against received trees walk_task_*_ext, src/remotes.c:213-552).  Every call
// walk(im, jm) decides from the PAIR alone (acceptance(), src/fmm.c:267-326) whether to emit, drop,
// open the sink side im or open the source side jm.  The same traversal is organised here BY SINK CELL:
//
//   F(im) = the set of sources jm for which the reference calls walk(im, jm)   ("frontier" of im)
//   F(root) = {root} and, with periodic images, the 26 displaced copies of the root (src/fmm.c:1028-1045).
//   Processing F(im) with the reference's rules gives  M2L pairs (jm -> im), dropped pairs, sources that
//   re-enter F(im) (source side opened) and the list O(im) of sources handed to BOTH sons (sink side
//   opened): F(son) = O(im).
//
// Pass 1 (frontier_node_kernel, one launch per tree level, one warp per sink node) walks the levels top
// down and stores O(im) in a bump-allocated span list.  Pass 2 (walk_fused_kernel, one warp per sink
// leaf) streams F(leaf) = O(parent), resolves it to source leaves and feeds them straight into the staged
// P2P pipeline of pn2_p2p.cuh -- the P2P lists never exist in HBM.  Every pair (im, jm) is visited exactly
// once, as in the reference; decisions use the reference's FP64 expressions in the same order (this file
// is compiled with -fmad=false), so the lists are the reference's lists, bit for bit.
//
// Periodic images / self-exchange: image k is the local tree displaced by shift[k]; the sender-side
// pruning of prepare_sendtree2 (src/remotes.c:97-158) is evaluated on the fly for the visited node.
#include "pn2_p2p.cuh"

#ifndef WALK_WARPS
#define WALK_WARPS 4
#endif
#define STACK_CAP 512
#define SRCQ_CAP 64
#define OBUF_CAP 256              // per-warp staging of O(im) before it is flushed to a span
#ifndef LEAF_MIN_BLOCKS
#define LEAF_MIN_BLOCKS 5          // register cap of the leaf kernel (102 regs, 20 warps/SM): best of 4..8 measured at 256^3
#endif

struct WalkArgs {
    int nleaf, ncell, root;            // LOCAL leaves / cells; root = local root cell
    int rleaf0, rnode0;                // received LET cells: leaves [rleaf0, rnode0), nodes [rnode0, ...)
    unsigned root_head;                // span holding F(root): (root | image) of the local tree and of every peer's tree
    const double *geom;
    const int *son;
    const LeafDesc *desc;
    const int *parent;
    const float *tiles;                // FP32 mode: leaf tiles (see walk_fused_kernel), pad_tile = the all-padding tile
    const double *tiles64;             // FP64 mode: leaf tiles of {x, y, z, w} doubles (walk_fused_f64_kernel)
    const double *gtab;                // FP64 mode: piecewise-polynomial table of g(u) (pn2_gtab.h)
    int pad_tile;
    const double *pos;
    double *acc;
    double cutoff, theta;
    int longshort;
    double tc[3], tw[3];               // this rank's domain box (pruning target for image trees)
    // span lists: 16-byte units; span = {count, next, 0, 0} + entries
    unsigned *spans;
    unsigned long long span_cap16;     // capacity in 16-byte units
    unsigned long long *span_top16;    // bump pointer
    unsigned *o_head;                  // [ncell] first span of O(cell), 0 = empty
    unsigned *m2l_t, *m2l_s;
    unsigned long long m2l_cap;
    unsigned long long *counters;      // [0] interactions, [1] m2l pairs, [2] p2p leaf pairs, [3] error flags, [4] visits
    long *lst_off;                     // dump mode
    unsigned *lst_src;
    int pass, emit_m2l;
    const int *work;                   // node kernel: cells of this level
    int nwork;
    // pass over the received trees: only the part of the sink tree that such a source can reach is visited.  A node that
    // hands a non-empty O(im) to its sons appends them to the next level's active list / the active leaf list; a level's
    // (or the leaf kernel's) launch covers the worst case and CTAs beyond the device-side count exit at once.
    const unsigned *work_count;        // != NULL: work[] holds *work_count entries (device counter)
    int *next_work;                    // node kernel: active list of the next level (or NULL)
    unsigned *next_count;
    int *leaf_work;                    // node kernel: active sink leaves (or NULL)
    unsigned *leaf_count;
};

// acceptance(), src/fmm.c:267-326: 0 open, 1 accept, -1 drop.  Same expressions in the same order (this file is
// compiled with -fmad=false); the reference's early returns are written as selects applied in reverse priority,
// which gives the same result without divergent branches.
__device__ __forceinline__ int accept_dev(const double *wi, const double *wj, double dx, double dy, double dz,
                                          double cutoff, double theta, int longshort) {
    const double w0 = (wi[0] + wj[0]) * 0.5, w1 = (wi[1] + wj[1]) * 0.5, w2 = (wi[2] + wj[2]) * 0.5;
    const double dd2 = dx * dx + dy * dy + dz * dz;
    double g0 = fabs(dx) - w0, g1 = fabs(dy) - w1, g2 = fabs(dz) - w2;
    g0 = g0 <= 0.0 ? 0.0 : g0;
    g1 = g1 <= 0.0 ? 0.0 : g1;
    g2 = g2 <= 0.0 ? 0.0 : g2;
    const double dm2 = g0 * g0 + g1 * g1 + g2 * g2;
    double wmax = w0;
    wmax = w1 > wmax ? w1 : wmax;
    wmax = w2 > wmax ? w2 : wmax;
    wmax *= 2;
    int f = (wmax * wmax < theta * theta * dd2) ? 1 : 0;
    if (longshort) {
        const double c2 = cutoff * cutoff;
        f = (dd2 > 1.0 * c2) ? 0 : f;
        f = (dm2 >= c2) ? -1 : f;
    }
    f = (g0 + g1 + g2 < 0.0001) ? 0 : f;
    return f;
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

__device__ __forceinline__ void load_geom(const double *geom, int cell, double c[3], double w[3]) {
    const double2 *g = reinterpret_cast<const double2 *>(geom + 6 * (size_t)cell);     // 48-byte records, 16-byte aligned
    double2 a = g[0], b = g[1], d = g[2];
    c[0] = a.x; c[1] = a.y; c[2] = b.x; w[0] = b.y; w[1] = d.x; w[2] = d.y;
}

// streams the entries of a span list, 32 at a time
struct SpanReader {
    const unsigned *spans;
    unsigned cur;       // current span (16-byte units), 0 = exhausted
    unsigned cnt, pos;  // entries in the current span / consumed
    __device__ void init(const unsigned *s, unsigned head) {
        spans = s; cur = head; cnt = 0; pos = 0;
        if (cur) cnt = spans[4 * (size_t)cur];
    }
    __device__ bool more() const { return cur != 0; }
    // warp-uniform: returns the number of entries fetched (<= 32); lane i < n gets its entry in e
    __device__ int fetch(int lane, unsigned &e) {
        while (cur && pos >= cnt) {              // next span
            cur = spans[4 * (size_t)cur + 1];
            pos = 0;
            cnt = cur ? spans[4 * (size_t)cur] : 0;
        }
        if (!cur) return 0;
        int n = (int)(cnt - pos);
        if (n > 32) n = 32;
        if (lane < n) e = spans[4 * (size_t)cur + 4 + pos + lane];
        pos += n;
        return n;
    }
};

// M2L pairs -> global list (warp-aggregated append)
__device__ __forceinline__ void emit_m2l_pairs(const WalkArgs &a, int lane, unsigned lt_mask, int emit, unsigned snk, unsigned src) {
    const unsigned mm = __ballot_sync(0xffffffffu, emit);
    if (mm && a.emit_m2l) {
        unsigned long long basei = 0;
        if (lane == 0) basei = atomicAdd(&a.counters[1], (unsigned long long)__popc(mm));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (emit) {
            unsigned long long at = basei + __popc(mm & lt_mask);
            if (at < a.m2l_cap) { a.m2l_t[at] = snk; a.m2l_s[at] = src; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 1: one warp per sink NODE of one level: F(im) -> M2L pairs + O(im)
// ------------------------------------------------------------------------------------------------
#ifndef NODE_MIN_BLOCKS
#define NODE_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(WALK_WARPS * 32, NODE_MIN_BLOCKS) frontier_node_kernel(WalkArgs a, P2PConst pc) {
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ unsigned s_obuf[WALK_WARPS][OBUF_CAP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int wk = blockIdx.x * WALK_WARPS + wib;
    if (wk >= a.nwork || (a.work_count && (unsigned)wk >= *a.work_count)) return;
    const int im = a.work[wk];
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned *stack = s_stack[wib], *obuf = s_obuf[wib];

    double ci[3], wi[3];
    load_geom(a.geom, im, ci, wi);
    const double swi = wi[0] + wi[1] + wi[2];

    SpanReader rd;
    rd.init(a.spans, im == a.root ? a.root_head : a.o_head[a.parent[im]]);
    int ssize = 0;
    unsigned pre_e = 0;                      // the next chunk of F(im), requested one step ahead
    int pre_n = rd.fetch(lane, pre_e);

    int osize = 0;
    unsigned first_span = 0, prev_span = 0;
    unsigned long long visits = 0;
    int err = 0;
    auto flush = [&]() {                     // obuf[0..osize) -> a new span
        if (osize == 0) return;
        unsigned units = 1 + (unsigned)((osize + 3) / 4);
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(a.span_top16, (unsigned long long)units);
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at + units > a.span_cap16) { err = 2; osize = 0; return; }       // span buffer full: the host grows it and redoes the pass
        unsigned *sp = a.spans + 4 * (size_t)at;
        if (lane == 0) {
            sp[0] = (unsigned)osize; sp[1] = 0; sp[2] = 0; sp[3] = 0;
            if (prev_span) a.spans[4 * (size_t)prev_span + 1] = (unsigned)at;
        }
        for (int k = lane; k < osize; k += 32) sp[4 + k] = obuf[k];
        if (!first_span) first_span = (unsigned)at;
        prev_span = (unsigned)at;
        osize = 0;
        __syncwarp();
    };

    while (true) {
        // refill from the parent's list when the stack runs low
        while (ssize < 32 && pre_n > 0) {
            if (lane < pre_n) stack[ssize + lane] = pre_e;
            ssize += pre_n;
            pre_n = rd.fetch(lane, pre_e);
            __syncwarp();
        }
        if (ssize == 0) break;
        int k = ssize < 32 ? ssize : 32;
        if (ssize > STACK_CAP - 64) k = 1;
        const int sbase = ssize - k;
        int npush = 0, emit_o = 0, emit_m = 0;
        unsigned p0 = 0, p1 = 0, o0 = 0, o1 = 0, msrc = 0;
        int no = 0;
        if (lane < k) {
            const unsigned jme = stack[sbase + lane];
            const int jm = (int)(jme & PN2_CELL_MASK);
            const unsigned img = jme >> PN2_IMG_SHIFT, imgbits = jme & ~PN2_CELL_MASK;
            const bool lj = jm < a.nleaf || (jm >= a.rleaf0 && jm < a.rnode0);
            const bool remote = img != 0 || jm >= a.rleaf0;          // walk_task_*_ext rules (src/remotes.c)
            // geometry and sons are requested together, before either is used (one round trip per step)
            double cj[3], wj[3];
            load_geom(a.geom, jm, cj, wj);
            const int2 sons = lj ? make_int2(-1, -1) : *reinterpret_cast<const int2 *>(a.son + 2 * (size_t)jm);
            if (img == 0 && jm == im) {
                // walk(im, im): all four son combinations (src/fmm.c:429-436)
                no = 2; o0 = (unsigned)sons.x; o1 = (unsigned)sons.y;
            } else {
                int pruned = 0;
                if (remote) {
                    // a packed node whose sons were not sent (-1) is terminal whatever this side computes
                    if (!lj) pruned = pruned_dev(cj, wj, pc.shift[img], a.tc, a.tw, a.cutoff, a.theta, a.longshort) | (sons.x < 0) | (sons.y < 0);
                    cj[0] += pc.shift[img][0]; cj[1] += pc.shift[img][1]; cj[2] += pc.shift[img][2];   // src/remotes.c:73-75
                }
                int f = accept_dev(wi, wj, ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2], a.cutoff, a.theta, a.longshort);
                if (f == 1) { emit_m = 1; msrc = jme; }
                else if (f == 0) {
                    // node x leaf: open the node; node x node: the one with the larger width sum, ties -> source
                    // (src/fmm.c:518-527); a pruned remote node cannot be opened (src/remotes.c:351-357)
                    bool open_i = lj || (swi > wj[0] + wj[1] + wj[2]) || pruned;
                    if (open_i) { no = 1; o0 = jme; }
                    else {
                        npush = 2;
                        p0 = (unsigned)sons.x | imgbits; p1 = (unsigned)sons.y | imgbits;
                    }
                }
            }
            emit_o = no;
        }
        visits += k;
        __syncwarp();
        const unsigned m2 = __ballot_sync(0xffffffffu, npush == 2);
        int pos0 = sbase + 2 * __popc(m2 & lt_mask);
        const int newsize = sbase + 2 * __popc(m2);
        if (newsize > STACK_CAP) { err = 1; break; }
        if (npush == 2) { stack[pos0] = p0; stack[pos0 + 1] = p1; }
        ssize = newsize;
        // O(im)
        const unsigned mo1 = __ballot_sync(0xffffffffu, emit_o >= 1), mo2 = __ballot_sync(0xffffffffu, emit_o == 2);
        const int nout = __popc(mo1) + __popc(mo2);
        if (osize + nout > OBUF_CAP) flush();
        if (emit_o >= 1) {
            int at = osize + __popc(mo1 & lt_mask) + __popc(mo2 & lt_mask);
            obuf[at] = o0;
            if (emit_o == 2) obuf[at + 1] = o1;
        }
        osize += nout;
        emit_m2l_pairs(a, lane, lt_mask, emit_m, (unsigned)im, msrc);
        __syncwarp();
        if (err) break;
    }
    flush();
    if (lane == 0) {
        a.o_head[im] = first_span;
        if (err) atomicOr(&a.counters[3], (unsigned long long)err);
        atomicAdd(&a.counters[4], visits);
        if (a.next_work && first_span) {                   // the sons inherit a non-empty frontier: they are active
            const int2 ch = *reinterpret_cast<const int2 *>(a.son + 2 * (size_t)im);
            const int sv[2] = {ch.x, ch.y};
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (sv[q] < 0) continue;
                if (sv[q] < a.nleaf) a.leaf_work[atomicAdd(a.leaf_count, 1u)] = sv[q];
                else a.next_work[atomicAdd(a.next_count, 1u)] = sv[q];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Pass 2: F(leaf) = O(parent) -> source leaves -> P2P (and leaf-level M2L pairs)
// ------------------------------------------------------------------------------------------------
// LeafWalk resolves the frontier of ONE sink leaf, 32 stack entries per step, and appends the source leaves it
// finds to a per-warp queue of 16-byte entries:
//     FP32 mode (MODE 0): {tile, dx, dy, dz}  tile = leaf tile index, d = source leaf centre - sink leaf centre
//                                             (+ image shift) in units of lambda = 2 rs sqrt(ln 2)
//     FP64 libm / dump modes (1, 2): {first, npart, cell | image << 27, 0}
//     FP64 tile mode (MODE 3): two int4 per entry: {tile, 0, dx (double)}, {dy, dz (doubles)}, d in units of 2 rs
template <int MODE, int QCAP>      // QCAP = queue capacity: a power of two > (entries a consumer leaves queued, < 32) + 32
struct LeafWalk {
    SpanReader rd;
    unsigned *stack;
    int4 *queue;
    const double *sink_g;        // shared: centre, width of the sink leaf
    LeafDesc sd;
    int leaf, ssize, qtail, err;
    unsigned nsrc, visits, npairs;
    unsigned pre_e;              // the next chunk of F(leaf), requested one step ahead (its latency overlaps the step)
    int pre_n;

    __device__ __forceinline__ void begin(const WalkArgs &a, int leaf_, unsigned *stack_, int4 *queue_, double *sink_smem, int lane) {
        leaf = leaf_; stack = stack_; queue = queue_; sink_g = sink_smem;
        rd.init(a.spans, a.o_head[a.parent[leaf]]);
        pre_e = 0;
        pre_n = rd.fetch(lane, pre_e);
        sd = a.desc[leaf];
        if (lane < 6) sink_smem[lane] = a.geom[6 * (size_t)leaf + lane];
        ssize = 0; qtail = 0; err = 0; nsrc = 0; visits = 0; npairs = 0;
        __syncwarp();
    }
    // one step = up to 32 entries of the stack; returns false when F(leaf) is exhausted (or the stack overflowed: err)
    __device__ __forceinline__ bool step(const WalkArgs &a, const P2PConst &pc, int lane) {
        const unsigned lt_mask = (1u << lane) - 1u;
        while (ssize < 32 && pre_n > 0) {
            if (lane < pre_n) stack[ssize + lane] = pre_e;
            ssize += pre_n;
            pre_n = rd.fetch(lane, pre_e);
            __syncwarp();
        }
        if (ssize == 0) return false;
        int k = ssize < 32 ? ssize : 32;
        if (k > STACK_CAP - ssize) k = STACK_CAP - ssize > 0 ? STACK_CAP - ssize : 1;     // every entry can grow the stack by one
        const int sbase = ssize - k;
        // ---- phase 1: every global load of the step is requested before any of them is used (one round trip):
        //      a leaf's descriptor {centre, first, npart} or a node's geometry {centre, width} + sons
        unsigned jme = 0;
        bool lj = false;
        double2 r0 = make_double2(0.0, 0.0), r1 = r0, r2 = r0;
        int2 sons = make_int2(0, 0);
        if (lane < k) {
            jme = stack[sbase + lane];
            const int jm = (int)(jme & PN2_CELL_MASK);
            lj = jm < a.nleaf || (jm >= a.rleaf0 && jm < a.rnode0);
            const double2 *rec = lj ? reinterpret_cast<const double2 *>(a.desc + jm) : reinterpret_cast<const double2 *>(a.geom + 6 * (size_t)jm);
            r0 = rec[0]; r1 = rec[1];
            if (!lj) { r2 = rec[2]; sons = *reinterpret_cast<const int2 *>(a.son + 2 * (size_t)jm); }
        }
        visits += k;
        __syncwarp();                      // the popped entries are in registers: the stack may be overwritten from sbase
        // ---- phase 2: decide, then compact pushes / queue entries / M2L pairs with ballots ----
        int npush = 0, emit_p = 0, emit_m = 0;
        unsigned p0 = 0, p1 = 0;
        int4 ent = make_int4(0, 0, 0, 0), ent2 = ent;
        if (lane < k) {
            const int jm = (int)(jme & PN2_CELL_MASK);
            const unsigned img = jme >> PN2_IMG_SHIFT, imgbits = jme & ~PN2_CELL_MASK;
            if (lj) {
                // leaf x leaf: always a P2P pair (src/fmm.c:438-451, src/remotes.c:228-240)
                emit_p = 1;
                const int dfirst = __double2loint(r1.y), dnpart = __double2hiint(r1.y);
                if (MODE == 0) {
                    ent.x = jm < a.nleaf ? jm : jm - a.rleaf0 + a.nleaf;
                    ent.y = __float_as_int((float)(((r0.x + pc.shift[img][0]) - sd.c[0]) * pc.inv_len));
                    ent.z = __float_as_int((float)(((r0.y + pc.shift[img][1]) - sd.c[1]) * pc.inv_len));
                    ent.w = __float_as_int((float)(((r1.x + pc.shift[img][2]) - sd.c[2]) * pc.inv_len));
                } else if (MODE == 3) {
                    const double ox = ((r0.x + pc.shift[img][0]) - sd.c[0]) * pc.inv2rs;
                    const double oy = ((r0.y + pc.shift[img][1]) - sd.c[1]) * pc.inv2rs;
                    const double oz = ((r1.x + pc.shift[img][2]) - sd.c[2]) * pc.inv2rs;
                    ent = make_int4(jm < a.nleaf ? jm : jm - a.rleaf0 + a.nleaf, 0, __double2loint(ox), __double2hiint(ox));
                    ent2 = make_int4(__double2loint(oy), __double2hiint(oy), __double2loint(oz), __double2hiint(oz));
                } else {
                    ent = make_int4(dfirst, dnpart, (int)jme, 0);
                }
                nsrc += (unsigned)(dnpart - ((jme == (unsigned)leaf) ? 1 : 0));
            } else {
                double cj[3] = {r0.x, r0.y, r1.x}, wj[3] = {r1.y, r2.x, r2.y};
                int pruned = 0;
                if (img != 0 || jm >= a.rleaf0) {
                    pruned = pruned_dev(cj, wj, pc.shift[img], a.tc, a.tw, a.cutoff, a.theta, a.longshort) | (sons.x < 0) | (sons.y < 0);
                    cj[0] += pc.shift[img][0]; cj[1] += pc.shift[img][1]; cj[2] += pc.shift[img][2];
                }
                const int f = accept_dev(sink_g + 3, wj, sink_g[0] - cj[0], sink_g[1] - cj[1], sink_g[2] - cj[2], a.cutoff, a.theta,
                                         a.longshort);
                if (f == 1 || (f == 0 && pruned)) emit_m = 1;        // forced M2L on a pruned node: src/remotes.c:442
                else if (f == 0) {
                    npush = 2;
                    p0 = (unsigned)sons.x | imgbits; p1 = (unsigned)sons.y | imgbits;
                }
            }
        }
        const unsigned m2 = __ballot_sync(0xffffffffu, npush == 2);
        const int pos0 = sbase + 2 * __popc(m2 & lt_mask);
        const int top = sbase + 2 * __popc(m2);
        if (top > STACK_CAP) { err = 1; return false; }
        if (npush == 2) { stack[pos0] = p0; stack[pos0 + 1] = p1; }
        const unsigned mp = __ballot_sync(0xffffffffu, emit_p);
        if (MODE == 3) {
            if (emit_p) {
                const int at = 2 * ((qtail + __popc(mp & lt_mask)) & (QCAP - 1));
                queue[at] = ent; queue[at + 1] = ent2;
            }
        } else if (emit_p) queue[(qtail + __popc(mp & lt_mask)) & (QCAP - 1)] = ent;
        qtail += __popc(mp);
        npairs += __popc(mp);
        emit_m2l_pairs(a, lane, lt_mask, emit_m, (unsigned)leaf, jme);
        ssize = top;
        __syncwarp();
        return true;
    }
};

// ---- FP64 parity mode and list dump: one warp per sink leaf, sources read through L1/L2 ----
template <int SW, int MODE>     // MODE 1: FP64 P2P, 2: dump lists (no arithmetic)
__global__ void __launch_bounds__(WALK_WARPS * 32) walk_leaf_kernel(WalkArgs a, P2PConst pc) {
    constexpr int NSL = 32 / SW;
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ int4 s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ double s_sink[WALK_WARPS][6];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int leaf = blockIdx.x * WALK_WARPS + wib;
    if (leaf >= a.nleaf) return;
    if (a.work_count) {
        if ((unsigned)leaf >= *a.work_count) return;
        leaf = a.work[leaf];
    }
    const int q = lane / SW, j = lane % SW;
    LeafWalk<MODE, SRCQ_CAP> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    const LeafDesc sd = w.sd;
    double xd = 0, yd = 0, zd = 0, axd = 0, ayd = 0, azd = 0;
    if (MODE == 1 && j < sd.npart) { const double *p = a.pos + 3 * (size_t)(sd.first + j); xd = p[0]; yd = p[1]; zd = p[2]; }
    long dump_pos = (MODE == 2 && a.pass == 1) ? a.lst_off[leaf] : 0;
    int qhead = 0;
    auto drain = [&](int limit) {          // consumes [qhead, limit)
        if (MODE == 1) {
            for (int idx = qhead + q; idx < limit; idx += NSL) {
                const int4 e = w.queue[idx & (SRCQ_CAP - 1)];
                const unsigned img = (unsigned)e.z >> PN2_IMG_SHIFT;
                const double sx = pc.shift[img][0], sy = pc.shift[img][1], sz = pc.shift[img][2];
                for (int k = 0; k < e.y; k++) {
                    const double *p = a.pos + 3 * (size_t)(e.x + k);
                    p2p_interact_f64(p[0] + sx, p[1] + sy, p[2] + sz, pc.mass, xd, yd, zd, axd, ayd, azd, pc.soft,
                                     pc.inv2rs, pc.longshort);
                }
            }
        } else {
            if (a.pass == 1)
                for (int idx = qhead + lane; idx < limit; idx += 32) a.lst_src[dump_pos + (idx - qhead)] = (unsigned)w.queue[idx & (SRCQ_CAP - 1)].z;
            dump_pos += limit - qhead;
        }
        qhead = limit;
        __syncwarp();
    };
    while (w.step(a, pc, lane))
        if (w.qtail - qhead >= 32) drain(qhead + ((w.qtail - qhead) / NSL) * NSL);
    if (w.qtail > qhead) drain(w.qtail);
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], 1ULL); return; }
    if (MODE == 1) {
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
        unsigned nsrc = w.nsrc;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) nsrc += __shfl_xor_sync(0xffffffffu, nsrc, m);
        if (lane == 0) {
            atomicAdd(&a.counters[0], (unsigned long long)nsrc * (unsigned long long)sd.npart);
            atomicAdd(&a.counters[2], (unsigned long long)w.npairs);
            atomicAdd(&a.counters[4], (unsigned long long)w.visits);
        }
    } else if (a.pass == 0 && lane == 0) {
        a.lst_off[leaf] = (long)w.npairs;
    }
}

// ------------------------------------------------------------------------------------------------
// FP32 product kernel: one warp per sink leaf, list walk fused with the P2P evaluation
// ------------------------------------------------------------------------------------------------
// Sources are read as LEAF TILES: every leaf owns one tile of SW slots in HBM, already in the packed-pair layout of
// the P2P loop (pn2_p2p.cuh: SW/2 pairs of {x0 x1 y0 y1 | z0 z1 w0 w1}, leaf-centre-relative, units of lambda = 2 rs sqrt(ln 2), unused
// slots = far-away zero-weight padding), so staging a source leaf is a plain 16 * SW byte copy: the warp issues
// cp.async (LDGSTS.128: no registers, no arithmetic) for a whole BATCH of queued leaves at once and goes on walking
// while the copies land; the per-leaf centre offset is applied on the SINK side (3 FADD per lane and stage).
// (A warp-specialised persistent variant -- walker warps feeding P2P warps through an mbarrier ring, setmaxnreg --
// was built and measured slower, 61.5 vs 48.1 ms at 256^3: one latency-bound walker cannot feed one P2P warp, see
// DESIGN.md 4.3.)
#ifndef FUSED_NST
#define FUSED_NST 8
#endif
template <int SW>
struct FusedLayout {
    static constexpr int NSL = 32 / SW;
    static constexpr int NST = FUSED_NST;                 // stages per batch
    static constexpr int BATCH = NST * NSL;               // source leaves per batch (32 / 16 / 8)
    static constexpr int TB = 16 * SW;                    // tile bytes
    static constexpr int ROWB = TB + 16;                  // row stride: the 16 spare bytes hold the leaf's {tile, dx, dy, dz}
                                                          // and de-conflict the NSL broadcast rows
    static constexpr int STAGE_BYTES = BATCH * ROWB;
};

__device__ __forceinline__ void cp_async16(unsigned dst_shared, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// (A "ragged" variant for sparsely filled leaves -- batch ordered by occupied slot pairs, a stage running only the pairs
// its fullest source leaf has -- was built and measured slower: the per-stage trip count breaks the straight-line 8-stage
// block the scheduler interleaves; Poisson 256^3 83.8 vs 76.6 ms, profiles/r02n_sweep_ragged.log.  Removed.)
template <int SW, bool LS>      // LS: with the long/short split factor g(r / 2rs)
__global__ void __launch_bounds__(WALK_WARPS * 32, LEAF_MIN_BLOCKS) walk_fused_kernel(WalkArgs a, P2PConst pc) {
    using FL = FusedLayout<SW>;
    constexpr int NSL = FL::NSL, NST = FL::NST, BATCH = FL::BATCH, TB = FL::TB, ROWB = FL::ROWB;
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ int4 s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ __align__(16) unsigned char s_stage[WALK_WARPS][FL::STAGE_BYTES];
    __shared__ double s_sink[WALK_WARPS][6];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int leaf = blockIdx.x * WALK_WARPS + wib;
    if (leaf >= a.nleaf) return;
    if (a.work_count) {                                    // pass over the received trees: active leaves only
        if ((unsigned)leaf >= *a.work_count) return;
        leaf = a.work[leaf];
    }
    const int q = lane / SW, j = lane % SW;
    LeafWalk<0, SRCQ_CAP> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    // sink: slot j of the leaf's own tile (padding slots compute, but are never written)
    float xi, yi, zi;
    {
        const float *t = a.tiles + (size_t)leaf * (4 * SW) + (j >> 1) * 8 + (j & 1);
        xi = t[0]; yi = t[2]; zi = t[4];
    }
    P2PSinkPk sk;
    sk.nx = sk.ny = sk.nz = sk.ax = sk.ay = sk.az = pk2(0.f, 0.f);
    const float inv_eps = pc.inv_eps;

    int qhead = 0, inflight = 0;
    unsigned char *stage = s_stage[wib];
    const unsigned stage_dst = (unsigned)__cvta_generic_to_shared(stage) + q * ROWB + j * 16;   // this lane's 16-byte chunk of row q
    const char *tile_src = reinterpret_cast<const char *>(a.tiles) + j * 16;
    auto issue_batch = [&](int cnt) {      // cnt <= BATCH queue entries from qhead, a multiple of NSL
#pragma unroll
        for (int s = 0; s < NST; s++) {
            if (s * NSL < cnt) {
                const int4 e = w.queue[(qhead + s * NSL + q) & (SRCQ_CAP - 1)];
                cp_async16(stage_dst + s * NSL * ROWB, tile_src + (size_t)e.x * TB);
                if (j == 0) *reinterpret_cast<int4 *>(stage + (s * NSL + q) * ROWB + TB) = e;
            }
        }
        cp_async_commit();
        inflight = cnt;
        qhead += cnt;
    };
    auto compute_stage = [&](int s) {
        const float *row = reinterpret_cast<const float *>(stage + (s * NSL + q) * ROWB);
        const float4 o = *reinterpret_cast<const float4 *>(row + TB / 4);
        const float nx = o.y - xi, ny = o.z - yi, nz = o.w - zi;          // x_j + (centre offset - x_i)
        sk.nx = pk2(nx, nx); sk.ny = pk2(ny, ny); sk.nz = pk2(nz, nz);
        pk_row<SW, LS>(row, 0, sk, inv_eps);
    };
    auto compute_batch = [&]() {
        cp_async_wait_all();
        __syncwarp();
        if (inflight == BATCH) {           // the common case as one straight-line block: stages overlap in the schedule
#pragma unroll
            for (int s = 0; s < NST; s++) compute_stage(s);
        } else {
#pragma unroll 1
            for (int s = 0; s * NSL < inflight; s++) compute_stage(s);
        }
        inflight = 0;
        __syncwarp();                      // the stage is free for the next batch
    };
    // walk until a batch of source leaves is queued, evaluate the batch whose tiles were requested one round
    // earlier, request the tiles of the new batch, walk on
    bool walking = true;
    while (true) {
        while (walking && w.qtail - qhead < BATCH) walking = w.step(a, pc, lane);
        if (inflight) compute_batch();
        const int avail = w.qtail - qhead;
        if (avail == 0 || w.err) break;                              // walking implies avail >= BATCH
        int cnt = BATCH;
        if (avail < BATCH) {                                           // the last batch: pad its last stage with the padding tile
            cnt = ((avail + NSL - 1) / NSL) * NSL;
            if (lane < cnt - avail) w.queue[(w.qtail + lane) & (SRCQ_CAP - 1)] = make_int4(a.pad_tile, 0, 0, 0);
            w.qtail = qhead + cnt;
            __syncwarp();
        }
        issue_batch(cnt);
    }
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], 1ULL); return; }

    float ax, ay, az, hi;
    unpk2(sk.ax, ax, hi); ax += hi;
    unpk2(sk.ay, ay, hi); ay += hi;
    unpk2(sk.az, az, hi); az += hi;
#pragma unroll
    for (int m = SW; m < 32; m <<= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, m);
        ay += __shfl_xor_sync(0xffffffffu, ay, m);
        az += __shfl_xor_sync(0xffffffffu, az, m);
    }
    const LeafDesc sd = w.sd;
    if (q == 0 && j < sd.npart) {
        const double sc = pc.mass * pc.inv_len * pc.inv_len;
        double *o = a.acc + 3 * (size_t)(sd.first + j);
        o[0] += (double)ax * sc; o[1] += (double)ay * sc; o[2] += (double)az * sc;
    }
    unsigned nsrc = w.nsrc;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) nsrc += __shfl_xor_sync(0xffffffffu, nsrc, m);
    if (lane == 0) {
        atomicAdd(&a.counters[0], (unsigned long long)nsrc * (unsigned long long)sd.npart);
        atomicAdd(&a.counters[2], (unsigned long long)w.npairs);
        atomicAdd(&a.counters[4], (unsigned long long)w.visits);
    }
}

// ------------------------------------------------------------------------------------------------
// FP64 product kernel: the same fused walk + P2P with double tiles and no libm
// ------------------------------------------------------------------------------------------------
// Tiles: SW slots of {x, y, z, w} doubles (32 bytes), leaf-centre-relative in units of 2 rs (so u = r), w = 1, padding
// slots far away with w = 0.  Arithmetic per interaction (reference: src/fmm.c:834-852, sqrt + division + erfc + exp
// from libm): 3 DADD + 3 DFMA (r^2) + MUFU.RSQ64H and one Newton step (4 ops, relative error < 1e-12) + 1 DSETP
// (softening) + 2 DMUL (1/r^3) + DMUL (u) + 2 ops and F2I / I2F (interval of the g(u) table) + 8 DFMA (degree-8
// piece of g, |error| < 7e-11: pn2_gtab.h, tools/fit_g64.py) + DMUL + 3 DFMA = 28 FP64-pipe operations.
#include "pn2_gtab.h"
#ifndef F64_MIN_BLOCKS
#define F64_MIN_BLOCKS 4
#endif
#define F64_NST 4
#ifndef F64_MAGIC
#define F64_MAGIC 1
#endif
template <int SW>
struct Fused64Layout {
    static constexpr int NSL = 32 / SW;
    static constexpr int NST = F64_NST;
    static constexpr int BATCH = NST * NSL;               // source leaves per batch (16 / 8 / 4)
    static constexpr int TB = 32 * SW;                    // tile bytes
    static constexpr int ROWB = TB + 32;                  // + {dx, dy, dz, -} of the leaf
    static constexpr int STAGE_BYTES = BATCH * ROWB;
};
__device__ __forceinline__ double pn2_rsqrt64(double x) {   // MUFU.RSQ64H: ~20-bit seed
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
struct P2PSink64 {
    double nx, ny, nz;        // centre offset - x_i
    double ax[2], ay[2], az[2];
};
template <bool LS>
__device__ __forceinline__ void p2p_interact_tab64(const double *slot, P2PSink64 &sk, int par, double eps2, double inv_eps,
                                                   const double (*gtab)[PN2_GTAB_K]) {
    const double2 xy = *reinterpret_cast<const double2 *>(slot);
    double zz, ww = 1.0;
    if (LS) zz = slot[2];
    else { const double2 zw = *reinterpret_cast<const double2 *>(slot + 2); zz = zw.x; ww = zw.y; }
    const double dx = xy.x + sk.nx, dy = xy.y + sk.ny, dz = zz + sk.nz;
    double r2 = fma(dx, dx, 1e-200);
    r2 = fma(dy, dy, r2);
    r2 = fma(dz, dz, r2);
    const double y0 = pn2_rsqrt64(r2);
    const double hh = r2 * y0;
    const double e = fma(-hh, y0, 1.0);
    const double rinv = fma(0.5 * y0, e, y0);                 // one Newton step
    const double ir = r2 < eps2 ? inv_eps : rinv;             // src/fmm.c:839-843
    double s = ir * ir * ir;
    if (LS) {
        const double u = r2 * rinv;
        const double t = fma(u, PN2_GTAB_INVH, -0.5);
#if F64_MAGIC
        // interval of u = round-to-nearest of u / h - 1/2, by the 2^52 + 2^51 trick: two DADD instead of an F2I and an I2F
        // on the quarter-rate conversion pipe (shorter dependent chain in front of the table loads)
        const double tm = t + 6755399441055744.0;
        const int k = __double2loint(tm);
        const double d = t - (tm - 6755399441055744.0);
#else
        const int k = __double2int_rn(t);                     // interval of u (round-to-nearest of u / h - 1/2)
        const double d = t - (double)k;
#endif
        const int kc = k < PN2_GTAB_K - 1 ? k : PN2_GTAB_K - 1;   // u >= 6: the zero entry
        double g = gtab[PN2_GTAB_DEG][kc];
#pragma unroll
        for (int j = PN2_GTAB_DEG - 1; j >= 0; j--) g = fma(g, d, gtab[j][kc]);
        s *= g;
    } else {
        s *= ww;
    }
    sk.ax[par] = fma(dx, s, sk.ax[par]);
    sk.ay[par] = fma(dy, s, sk.ay[par]);
    sk.az[par] = fma(dz, s, sk.az[par]);
}

template <int SW, bool LS>
__global__ void __launch_bounds__(WALK_WARPS * 32, F64_MIN_BLOCKS) walk_fused_f64_kernel(WalkArgs a, P2PConst pc) {
    using FL = Fused64Layout<SW>;
    constexpr int NSL = FL::NSL, NST = FL::NST, BATCH = FL::BATCH, TB = FL::TB, ROWB = FL::ROWB;
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ int4 s_srcq[WALK_WARPS][2 * SRCQ_CAP];
    __shared__ __align__(16) unsigned char s_stage[WALK_WARPS][FL::STAGE_BYTES];
    __shared__ double s_sink[WALK_WARPS][6];
    __shared__ __align__(128) double s_gtab[PN2_GTAB_DEG + 1][PN2_GTAB_K];
    for (int i = threadIdx.x; i < (PN2_GTAB_DEG + 1) * PN2_GTAB_K; i += blockDim.x) (&s_gtab[0][0])[i] = a.gtab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int leaf = blockIdx.x * WALK_WARPS + wib;
    if (leaf >= a.nleaf) return;
    if (a.work_count) {
        if ((unsigned)leaf >= *a.work_count) return;
        leaf = a.work[leaf];
    }
    const int q = lane / SW, j = lane % SW;
    LeafWalk<3, SRCQ_CAP> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    double xi, yi, zi;
    {
        const double *t = a.tiles64 + ((size_t)leaf * SW + j) * 4;
        xi = t[0]; yi = t[1]; zi = t[2];
    }
    P2PSink64 sk;
    sk.nx = sk.ny = sk.nz = 0.0;
    sk.ax[0] = sk.ax[1] = sk.ay[0] = sk.ay[1] = sk.az[0] = sk.az[1] = 0.0;
    const double eps = pc.soft * pc.inv2rs;
    const double eps2 = eps * eps, inv_eps = eps > 0.0 ? 1.0 / eps : 1e100;

    int qhead = 0, inflight = 0;
    unsigned char *stage = s_stage[wib];
    const unsigned stage_dst = (unsigned)__cvta_generic_to_shared(stage) + q * ROWB + j * 16;   // first of this lane's two chunks
    const char *tile_src = reinterpret_cast<const char *>(a.tiles64) + j * 16;
    auto issue_batch = [&](int cnt) {      // cnt <= BATCH queue entries from qhead, a multiple of NSL
#pragma unroll
        for (int s = 0; s < NST; s++) {
            if (s * NSL < cnt) {
                const int qi = 2 * ((qhead + s * NSL + q) & (SRCQ_CAP - 1));
                const int4 e0 = w.queue[qi];
                const char *src = tile_src + (size_t)e0.x * TB;
                cp_async16(stage_dst + s * NSL * ROWB, src);
                cp_async16(stage_dst + s * NSL * ROWB + SW * 16, src + SW * 16);
                if (j == 0) {
                    int4 *hd = reinterpret_cast<int4 *>(stage + (s * NSL + q) * ROWB + TB);
                    hd[0] = e0; hd[1] = w.queue[qi + 1];
                }
            }
        }
        cp_async_commit();
        inflight = cnt;
        qhead += cnt;
    };
    auto compute_stage = [&](int s) {
        const double *row = reinterpret_cast<const double *>(stage + (s * NSL + q) * ROWB);
        const double *hd = row + TB / 8;
        sk.nx = hd[1] - xi; sk.ny = hd[2] - yi; sk.nz = hd[3] - zi;
#pragma unroll
        for (int k = 0; k < SW; k++) p2p_interact_tab64<LS>(row + 4 * k, sk, k & 1, eps2, inv_eps, s_gtab);
    };
    auto compute_batch = [&]() {
        cp_async_wait_all();
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s * NSL < inflight; s++) compute_stage(s);
        inflight = 0;
        __syncwarp();
    };
    bool walking = true;
    while (true) {
        while (walking && w.qtail - qhead < BATCH) walking = w.step(a, pc, lane);
        if (inflight) compute_batch();
        const int avail = w.qtail - qhead;
        if (avail == 0 || w.err) break;
        int cnt = BATCH;
        if (avail < BATCH) {
            cnt = ((avail + NSL - 1) / NSL) * NSL;
            if (lane < cnt - avail) {
                const int at = 2 * ((w.qtail + lane) & (SRCQ_CAP - 1));
                w.queue[at] = make_int4(a.pad_tile, 0, 0, 0); w.queue[at + 1] = make_int4(0, 0, 0, 0);
            }
            w.qtail = qhead + cnt;
            __syncwarp();
        }
        issue_batch(cnt);
    }
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], 1ULL); return; }

    double ax = sk.ax[0] + sk.ax[1], ay = sk.ay[0] + sk.ay[1], az = sk.az[0] + sk.az[1];
#pragma unroll
    for (int m = SW; m < 32; m <<= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, m);
        ay += __shfl_xor_sync(0xffffffffu, ay, m);
        az += __shfl_xor_sync(0xffffffffu, az, m);
    }
    const LeafDesc sd = w.sd;
    if (q == 0 && j < sd.npart) {
        const double sc = pc.mass * pc.inv2rs * pc.inv2rs;
        double *o = a.acc + 3 * (size_t)(sd.first + j);
        o[0] += ax * sc; o[1] += ay * sc; o[2] += az * sc;
    }
    unsigned nsrc = w.nsrc;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) nsrc += __shfl_xor_sync(0xffffffffu, nsrc, m);
    if (lane == 0) {
        atomicAdd(&a.counters[0], (unsigned long long)nsrc * (unsigned long long)sd.npart);
        atomicAdd(&a.counters[2], (unsigned long long)w.npairs);
        atomicAdd(&a.counters[4], (unsigned long long)w.visits);
    }
}

// FP64 tiles: slot j of tile t <- (pos - leaf centre) / (2 rs) of particle j, w = 1; padding far away with w = 0
template <int SW>
__global__ void tile64_kernel(int nt, int nleaf, int rleaf0, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                              double inv2rs, double *__restrict__ tiles) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = (int)(t / SW), j = (int)(t % SW);
    if (tile > nt) return;
    double4 p = make_double4(1e4, 1e4, 1e4, 0.0);
    if (tile < nt) {
        const LeafDesc d = desc[tile < nleaf ? tile : tile - nleaf + rleaf0];
        if (j < d.npart) {
            const double *x = pos + 3 * (size_t)(d.first + j);
            p = make_double4((x[0] - d.c[0]) * inv2rs, (x[1] - d.c[1]) * inv2rs, (x[2] - d.c[2]) * inv2rs, 1.0);
        }
    }
    double2 *o = reinterpret_cast<double2 *>(tiles + ((size_t)tile * SW + j) * 4);
    o[0] = make_double2(p.x, p.y); o[1] = make_double2(p.z, p.w);
}

// Leaf tiles (FP32 mode): slot j of tile t <- particle j of the leaf, packed-pair layout; tile nt = all padding.
// Tiles 0..nleaf-1 are the local leaves, nleaf.. the received LET leaves (cells rleaf0..).
// With the long/short split the tile layout carries no weight (pn2_p2p.cuh): a padding slot is harmless because
// 2^(-r'^2) flushes to 0 at its distance, which holds while every real particle stays within PN2_PAD_SAFE lambda of
// its leaf centre per dimension (|sink - pad| >= sqrt(3) (24 - 3 B - 2.71) >= 11.3 lambda for B <= 4.9; leaf pairs
// are listed only when their boxes are closer than the cut-off, 2.71 lambda).  Wider leaves (NSIDE far finer than the
// particle spacing) are reported instead of being evaluated wrongly: counters[3] |= 4.
#define PN2_PAD_SAFE 4.9f
template <int SW>
__global__ void tile_kernel(int nt, int nleaf, int rleaf0, int t0, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                            const float4 *__restrict__ rel, double inv_len, float *__restrict__ tiles, int longshort,
                            unsigned long long *__restrict__ counters) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = t0 + (int)(t / SW), j = (int)(t % SW);
    if (tile > nt) return;
    float4 p = make_float4(PN2_PAD_COORD, PN2_PAD_COORD, PN2_PAD_COORD, 0.f);
    if (tile < nt) {
        const LeafDesc d = desc[tile < nleaf ? tile : tile - nleaf + rleaf0];
        if (j < d.npart) {
            if (tile < nleaf) {
                // a local leaf: leaf-centre-relative coordinates in units of lambda, straight from the FP64 positions
                const double *x = pos + 3 * (size_t)(d.first + j);
                p = make_float4((float)((x[0] - d.c[0]) * inv_len), (float)((x[1] - d.c[1]) * inv_len), (float)((x[2] - d.c[2]) * inv_len), 1.f);
            } else p = rel[d.first + j];                     // a received leaf: its sender shipped these very values
            if (longshort && fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)) > PN2_PAD_SAFE) atomicOr(&counters[3], 4ULL);
        }
    }
    float *o = tiles + (size_t)tile * (4 * SW) + (j >> 1) * 8 + (j & 1);
    o[0] = p.x; o[2] = p.y; o[4] = p.z; o[6] = p.w;
}

template <int SW>
static void launch_mode(pn2_ctx *h, const WalkArgs &a, int mode) {
    const unsigned grid = (unsigned)((a.nleaf + WALK_WARPS - 1) / WALK_WARPS);
    if (mode == 0 && h->prm.longshort) walk_fused_kernel<SW, true><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else if (mode == 0) walk_fused_kernel<SW, false><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else if (mode == 3 && h->prm.longshort) walk_fused_f64_kernel<SW, true><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else if (mode == 3) walk_fused_f64_kernel<SW, false><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else if (mode == 1) walk_leaf_kernel<SW, 1><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    else walk_leaf_kernel<SW, 2><<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    h->launches++;
}

static void fill_args(pn2_ctx *h, WalkArgs &a) {
    memset(&a, 0, sizeof a);
    a.nleaf = h->nleaf; a.ncell = h->ncell; a.root = h->nleaf;
    a.rleaf0 = h->ncell; a.rnode0 = h->ncell + h->nrl; a.root_head = h->root_head;
    a.geom = h->geom.p; a.son = h->son.p; a.desc = h->desc.p; a.parent = h->parent.p;
    a.pos = h->pos.p; a.acc = h->acc.p;
    a.cutoff = h->prm.cutoff; a.theta = h->prm.theta; a.longshort = h->prm.longshort;
    for (int d = 0; d < 3; d++) {
        // the pruning box of prepare_sendtree2 is the target's box as centre / width (src/remotes.c:97-110)
        a.tc[d] = 0.5 * (h->dom.hi[d] + h->dom.lo[d]);
        a.tw[d] = h->dom.hi[d] - h->dom.lo[d];
    }
    a.spans = h->spans.p; a.span_cap16 = h->span_cap16; a.span_top16 = h->counters.p + 6; a.o_head = h->o_head.p;
    a.m2l_t = h->m2l_pairs.p; a.m2l_s = h->m2l_pairs.p + h->m2l_cap; a.m2l_cap = h->m2l_cap;
    a.counters = h->counters.p;
    a.lst_off = h->lst_off.p; a.lst_src = h->lst_src.p;
}

// Pass 1 over all levels.  counters[6] is the span bump pointer (unit 0 is reserved: 0 = "no span").
// h->walk_active: the pass over the received trees -- levels and leaves are visited through active lists built on the way.
int pn2_walk_frontiers(pn2_ctx *h) {
    if (h->nnode == 0) return PN2_OK;
    WalkArgs a;
    fill_args(h, a);
    a.emit_m2l = 1;
    if (h->walk_active) {
        // act_nodes: one region per level (the level's nodes at most); act_leaf: nleaf; act_count: [nlevel + 1] counters, the last for the leaves
        PN2_TRY(h->act_nodes.ensure((size_t)h->nnode + 1)); PN2_TRY(h->act_leaf.ensure((size_t)h->nleaf + 1));
        PN2_TRY(h->act_count.ensure((size_t)h->nlevel + 2));
        CUDA_TRY(cudaMemsetAsync(h->act_count.p, 0, ((size_t)h->nlevel + 2) * sizeof(unsigned), h->stream));
        const int root = h->nleaf;
        const unsigned one = 1u;
        CUDA_TRY(cudaMemcpyAsync(h->act_nodes.p, &root, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->act_count.p, &one, sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
    }
    for (int lev = 0; lev < h->nlevel; lev++) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        if (h->walk_active) {
            a.work = h->act_nodes.p + h->level_off[lev];
            a.work_count = h->act_count.p + lev;
            a.next_work = lev + 1 < h->nlevel ? h->act_nodes.p + h->level_off[lev + 1] : h->act_nodes.p;     // the last level has no node sons
            a.next_count = h->act_count.p + lev + 1;
            a.leaf_work = h->act_leaf.p;
            a.leaf_count = h->act_count.p + h->nlevel + 1;
        } else {
            a.work = h->level_nodes.p + h->level_off[lev];
        }
        a.nwork = cnt;
        frontier_node_kernel<<<(cnt + WALK_WARPS - 1) / WALK_WARPS, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

// dump = 0: the product step (P2P evaluated, leaf-level M2L pairs appended);
// dump = 1 / 2: list dump passes (count / fill) for pn2_get_lists
int pn2_walk_fused(pn2_ctx *h, int dump) {
    if (h->nleaf == 0) return PN2_OK;
    WalkArgs a;
    fill_args(h, a);
    a.pass = dump == 2 ? 1 : 0;
    a.emit_m2l = dump == 0;
    if (h->walk_active && dump == 0) { a.work = h->act_leaf.p; a.work_count = h->act_count.p + h->nlevel + 1; }
    // PN2_FP64: table-driven FP64 kernel (3); PN2_FP64_LIBM: the reference's expression with libm (1); PN2_FP32: 0
    int mode = dump ? 2 : (h->prm.precision == PN2_FP64 ? 3 : (h->prm.precision == PN2_FP64_LIBM ? 1 : 0));
    int ml = h->prm.maxleaf;
    if (mode == 3) {
        const int sw = ml <= 8 ? 8 : (ml <= 16 ? 16 : 32);
        const int nt = h->nleaf + h->nrl;
        PN2_TRY(h->tiles64.ensure(((size_t)nt + 1) * 4 * sw));
        if (!h->gtab.p) {
            PN2_TRY(h->gtab.ensure((PN2_GTAB_DEG + 1) * PN2_GTAB_K));
            CUDA_TRY(cudaMemcpyAsync(h->gtab.p, PN2_GTAB, sizeof PN2_GTAB, cudaMemcpyHostToDevice, h->stream));
        }
        const long nthr = ((long)nt + 1) * sw;
        const unsigned g = (unsigned)((nthr + 255) / 256);
        if (sw == 8) tile64_kernel<8><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, h->desc.p, h->pos.p, h->pc.inv2rs, h->tiles64.p);
        else if (sw == 16) tile64_kernel<16><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, h->desc.p, h->pos.p, h->pc.inv2rs, h->tiles64.p);
        else tile64_kernel<32><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, h->desc.p, h->pos.p, h->pc.inv2rs, h->tiles64.p);
        h->launches++;
        a.tiles64 = h->tiles64.p; a.gtab = h->gtab.p; a.pad_tile = nt;
    }
    if (mode == 0) {
        // leaf tiles of the local and the received leaves (+ one padding tile).  The local tiles are built once per step
        // (the pass over the local roots); the pass over the received trees appends its tiles and a new padding tile.
        const int sw = ml <= 8 ? 8 : (ml <= 16 ? 16 : 32);
        const int nt = h->nleaf + h->nrl;
        const size_t need = ((size_t)nt + 1) * 4 * sw;
        const bool grows = need > h->tiles.cap;                         // a new buffer: every tile again
        PN2_TRY(h->tiles.ensure(need));
        const int t0 = (!grows && h->tiles_built_for == h->step_serial && h->nrl > 0) ? h->nleaf : 0;
        const long nthr = ((long)nt + 1 - t0) * sw;
        const unsigned g = (unsigned)((nthr + 255) / 256);
        if (sw == 8) tile_kernel<8><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, t0, h->desc.p, h->pos.p, h->rel.p, h->pc.inv_len, h->tiles.p, h->prm.longshort, h->counters.p);
        else if (sw == 16) tile_kernel<16><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, t0, h->desc.p, h->pos.p, h->rel.p, h->pc.inv_len, h->tiles.p, h->prm.longshort, h->counters.p);
        else tile_kernel<32><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, t0, h->desc.p, h->pos.p, h->rel.p, h->pc.inv_len, h->tiles.p, h->prm.longshort, h->counters.p);
        h->launches++;
        h->tiles_built_for = h->step_serial;
        a.tiles = h->tiles.p; a.pad_tile = nt;
    }
    if (ml <= 8) launch_mode<8>(h, a, mode);
    else if (ml <= 16) launch_mode<16>(h, a, mode);
    else launch_mode<32>(h, a, mode);
    KERNEL_CHECK();
    return PN2_OK;
}
