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
#define SRCQ_CAP 256               // per-warp queue of source leaves (4-byte entries), a power of two >= 160 (see LeafWalk)
#define PN2_WALK_WATCHDOG (1u << 20)   // steps of one sink leaf before the walk is declared stuck (counters[3] |= 8)
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
    int item_lo, item_hi;              // leaf kernels: the sink leaves item_lo .. item_hi - 1 (0 .. nleaf - 1 unless set)
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
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned *stack = s_stack[wib], *obuf = s_obuf[wib];
    // a pass over the received trees (work_count != NULL) is launched with a bounded grid that strides over the active
    // list; otherwise the grid covers the level and the loop runs once
    int nwork = a.nwork;
    if (a.work_count && (int)*a.work_count < nwork) nwork = (int)*a.work_count;
    for (int wk = blockIdx.x * WALK_WARPS + wib; wk < nwork; wk += gridDim.x * WALK_WARPS) {
    const int im = a.work[wk];

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
    unsigned nstep = 0;
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
        if (++nstep > PN2_WALK_WATCHDOG) { err = 8; break; }           // a walk that does not end is reported, not waited for
        int k = ssize < 32 ? ssize : 32;
        if (ssize > STACK_CAP - 64) k = 1;
        const int sbase = ssize - k;
        // one straight code path for the whole warp: lanes beyond k repeat entry 0 (same addresses) and are masked out
        const bool act = lane < k;
        const unsigned jme = stack[sbase + (act ? lane : 0)];
        const int jm = (int)(jme & PN2_CELL_MASK);
        const unsigned img = jme >> PN2_IMG_SHIFT, imgbits = jme & ~PN2_CELL_MASK;
        const bool lj = jm < a.nleaf || (jm >= a.rleaf0 && jm < a.rnode0);
        // geometry and sons are requested together, before either is used (one round trip per step)
        double cj[3], wj[3];
        load_geom(a.geom, jm, cj, wj);
        const int2 sons = lj ? make_int2(-1, -1) : *reinterpret_cast<const int2 *>(a.son + 2 * (size_t)jm);
        int pruned = 0;
        if (img != 0 || jm >= a.rleaf0) {                         // walk_task_*_ext rules (src/remotes.c)
            // a packed node whose sons were not sent (-1) is terminal whatever this side computes
            if (!lj) pruned = pruned_dev(cj, wj, pc.shift[img], a.tc, a.tw, a.cutoff, a.theta, a.longshort) | (sons.x < 0) | (sons.y < 0);
            cj[0] += pc.shift[img][0]; cj[1] += pc.shift[img][1]; cj[2] += pc.shift[img][2];   // src/remotes.c:73-75
        }
        const int f = accept_dev(wi, wj, ci[0] - cj[0], ci[1] - cj[1], ci[2] - cj[2], a.cutoff, a.theta, a.longshort);
        // walk(im, im): all four son combinations (src/fmm.c:429-436): both sons inherit both sons
        const bool self = act && img == 0 && jm == im;
        const bool other = act && !self;
        // node x leaf: open the node; node x node: the one with the larger width sum, ties -> source
        // (src/fmm.c:518-527); a pruned remote node cannot be opened (src/remotes.c:351-357)
        const bool open_i = lj || (swi > wj[0] + wj[1] + wj[2]) || pruned;
        const int emit_m = other && f == 1;
        const int npush = (other && f == 0 && !open_i) ? 2 : 0;
        const int emit_o = self ? 2 : ((other && f == 0 && open_i) ? 1 : 0);
        const unsigned o0 = self ? (unsigned)sons.x : jme, o1 = (unsigned)sons.y;
        const unsigned p0 = (unsigned)sons.x | imgbits, p1 = (unsigned)sons.y | imgbits;
        const unsigned msrc = jme;
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
    __syncwarp();
    }   // work items of this warp
}

// ------------------------------------------------------------------------------------------------
// Pass 2: F(leaf) = O(parent) -> source leaves -> P2P (and leaf-level M2L pairs)
// ------------------------------------------------------------------------------------------------
// LeafWalk resolves the frontier of ONE sink leaf.  Source LEAVES are never tested (leaf x leaf is always a P2P pair,
// src/fmm.c:438-451, src/remotes.c:228-240): wherever one turns up -- in the parent's list O(parent) or as the son of an
// opened source node -- it goes straight to the per-warp QUEUE of 4-byte entries (cell | image << 27), untouched.  Only
// source NODES sit on the stack; a step tests up to 32 of them against the sink leaf (all lanes on the same code path),
// pushes the node sons of the opened ones and queues their leaf sons.  What a consumer needs to know about a queued
// leaf (centre, particle count) it reads from the leaf's 32-byte descriptor when it stages the leaf's particles.
__device__ __forceinline__ bool is_leaf_cell(const WalkArgs &a, int jm) { return jm < a.nleaf || (jm >= a.rleaf0 && jm < a.rnode0); }

// TILE: the queue holds TILE indices (tile | image << 27) instead of cell ids: tile = cell for a local leaf, cell - nnode
// for a received one (the tile arrays hold the local leaves, then the received leaves) -- translated here, 32 leaves per
// instruction, so that the staging loop of the fused kernels has nothing to compute.
template <int QCAP, bool TILE>   // QCAP = queue capacity, a power of two: an in-flight batch (<= 32, re-read when it lands) + < batch
                                 // <= 32 entries left queued + <= 32 from the refill + <= 64 from a node step
struct LeafWalk {
    SpanReader rd;
    unsigned *stack;
    unsigned *queue;
    const double *sink_g;        // shared: centre, width of the sink leaf
    LeafDesc sd;
    int leaf, ssize, qtail, err;
    unsigned visits, npairs, nstep;
    unsigned pre_e;              // the next chunk of F(leaf), requested one step ahead (its latency overlaps the step)
    int pre_n;

    __device__ __forceinline__ void begin(const WalkArgs &a, int leaf_, unsigned *stack_, unsigned *queue_, double *sink_smem, int lane) {
        leaf = leaf_; stack = stack_; queue = queue_; sink_g = sink_smem;
        rd.init(a.spans, a.o_head[a.parent[leaf]]);
        pre_e = 0;
        pre_n = rd.fetch(lane, pre_e);
        sd = a.desc[leaf];
        if (lane < 6) sink_smem[lane] = a.geom[6 * (size_t)leaf + lane];
        ssize = 0; qtail = 0; err = 0; visits = 0; npairs = 0; nstep = 0;
        __syncwarp();
    }
    // queue entry of a source leaf
    __device__ __forceinline__ unsigned qentry(const WalkArgs &a, unsigned e) const {
        if (!TILE) return e;
        return (int)(e & PN2_CELL_MASK) >= a.rleaf0 ? e - (unsigned)(a.rleaf0 - a.nleaf) : e;
    }
    // one step; qhead = the consumer's read position, batch = the number of queued leaves it waits for.
    // Returns false when F(leaf) is exhausted (or the stack overflowed: err).
    __device__ __forceinline__ bool step(const WalkArgs &a, const P2PConst &pc, int lane, int qhead, int batch) {
        const unsigned lt_mask = (1u << lane) - 1u;
        if (++nstep > PN2_WALK_WATCHDOG) { err = 8; return false; }      // a walk that does not end is reported, not waited for
        // ---- refill from the parent's list: leaves -> queue, nodes -> stack ----
        while (ssize < 32 && pre_n > 0 && qtail - qhead < batch) {
            const bool valid = lane < pre_n;
            const unsigned e = pre_e;
            const bool lj = valid && is_leaf_cell(a, (int)(e & PN2_CELL_MASK));
            const bool nj = valid && !lj;
            const unsigned ml = __ballot_sync(0xffffffffu, lj), mn = __ballot_sync(0xffffffffu, nj);
            if (lj) queue[(qtail + __popc(ml & lt_mask)) & (QCAP - 1)] = qentry(a, e);
            if (nj) stack[ssize + __popc(mn & lt_mask)] = e;
            qtail += __popc(ml); npairs += __popc(ml); visits += __popc(ml);
            ssize += __popc(mn);
            pre_n = rd.fetch(lane, pre_e);
            __syncwarp();
        }
        if (qtail - qhead >= batch) return true;
        if (ssize == 0) return false;              // the loop above ended with pre_n == 0: nothing left
        int k = ssize < 32 ? ssize : 32;
        if (k > STACK_CAP - ssize) k = STACK_CAP - ssize > 0 ? STACK_CAP - ssize : 1;     // every entry can grow the stack by one
        const int sbase = ssize - k;
        // ---- every global load of the step is requested before any of them is used (one round trip): geometry
        //      {centre, width} + sons of the node.  Lanes beyond k repeat entry 0 (same addresses) and are masked out below,
        //      so the whole step is one straight code path.
        const bool act = lane < k;
        const unsigned jme = stack[sbase + (act ? lane : 0)];
        const int jm = (int)(jme & PN2_CELL_MASK);
        const double2 *rec = reinterpret_cast<const double2 *>(a.geom + 6 * (size_t)jm);
        const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2];
        const int2 sons = *reinterpret_cast<const int2 *>(a.son + 2 * (size_t)jm);
        visits += k;
        __syncwarp();                      // the popped entries are in registers: the stack may be overwritten from sbase
        const unsigned img = jme >> PN2_IMG_SHIFT, imgbits = jme & ~PN2_CELL_MASK;
        double cj[3] = {r0.x, r0.y, r1.x}, wj[3] = {r1.y, r2.x, r2.y};
        int pruned = 0;
        if (img != 0 || jm >= a.rleaf0) {
            pruned = pruned_dev(cj, wj, pc.shift[img], a.tc, a.tw, a.cutoff, a.theta, a.longshort) | (sons.x < 0) | (sons.y < 0);
            cj[0] += pc.shift[img][0]; cj[1] += pc.shift[img][1]; cj[2] += pc.shift[img][2];
        }
        const int f = accept_dev(sink_g + 3, wj, sink_g[0] - cj[0], sink_g[1] - cj[1], sink_g[2] - cj[2], a.cutoff, a.theta, a.longshort);
        const int emit_m = act && (f == 1 || (f == 0 && pruned));        // forced M2L on a pruned node: src/remotes.c:442
        const bool open = act && f == 0 && !pruned;
        const bool l0 = open && is_leaf_cell(a, sons.x), l1 = open && is_leaf_cell(a, sons.y);
        const bool n0 = open && !l0, n1 = open && !l1;
        const unsigned bn0 = __ballot_sync(0xffffffffu, n0), bn1 = __ballot_sync(0xffffffffu, n1);
        const unsigned bl0 = __ballot_sync(0xffffffffu, l0), bl1 = __ballot_sync(0xffffffffu, l1);
        const int top = sbase + __popc(bn0) + __popc(bn1);
        if (top > STACK_CAP) { err = 1; return false; }
        const int spos = sbase + __popc(bn0 & lt_mask) + __popc(bn1 & lt_mask);
        if (n0) stack[spos] = (unsigned)sons.x | imgbits;
        if (n1) stack[spos + (n0 ? 1 : 0)] = (unsigned)sons.y | imgbits;
        const int qpos = qtail + __popc(bl0 & lt_mask) + __popc(bl1 & lt_mask);
        if (l0) queue[qpos & (QCAP - 1)] = qentry(a, (unsigned)sons.x | imgbits);
        if (l1) queue[(qpos + (l0 ? 1 : 0)) & (QCAP - 1)] = qentry(a, (unsigned)sons.y | imgbits);
        const int nl = __popc(bl0) + __popc(bl1);
        qtail += nl; npairs += nl; visits += nl;
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
    __shared__ unsigned s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ double s_sink[WALK_WARPS][6];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int leaf = blockIdx.x * WALK_WARPS + wib;
    if (leaf >= a.nleaf) return;
    if (a.work_count) {
        if ((unsigned)leaf >= *a.work_count) return;
        leaf = a.work[leaf];
    }
    const int q = lane / SW, j = lane % SW;
    LeafWalk<SRCQ_CAP, false> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    const LeafDesc sd = w.sd;
    double xd = 0, yd = 0, zd = 0, axd = 0, ayd = 0, azd = 0;
    if (MODE == 1 && j < sd.npart) { const double *p = a.pos + 3 * (size_t)(sd.first + j); xd = p[0]; yd = p[1]; zd = p[2]; }
    long dump_pos = (MODE == 2 && a.pass == 1) ? a.lst_off[leaf] : 0;
    int qhead = 0;
    unsigned nsrc = 0;
    auto drain = [&](int limit) {          // consumes [qhead, limit)
        if (MODE == 1) {
            for (int idx = qhead + q; idx < limit; idx += NSL) {
                const unsigned e = w.queue[idx & (SRCQ_CAP - 1)];
                const LeafDesc d = a.desc[e & PN2_CELL_MASK];
                const unsigned img = e >> PN2_IMG_SHIFT;
                const double sx = pc.shift[img][0], sy = pc.shift[img][1], sz = pc.shift[img][2];
                if (j == 0) nsrc += (unsigned)(d.npart - ((e == (unsigned)leaf) ? 1 : 0));
                for (int k = 0; k < d.npart; k++) {
                    const double *p = a.pos + 3 * (size_t)(d.first + k);
                    p2p_interact_f64(p[0] + sx, p[1] + sy, p[2] + sz, pc.mass, xd, yd, zd, axd, ayd, azd, pc.soft,
                                     pc.inv2rs, pc.longshort);
                }
            }
        } else {
            if (a.pass == 1)
                for (int idx = qhead + lane; idx < limit; idx += 32) a.lst_src[dump_pos + (idx - qhead)] = w.queue[idx & (SRCQ_CAP - 1)];
            dump_pos += limit - qhead;
        }
        qhead = limit;
        __syncwarp();
    };
    while (w.step(a, pc, lane, qhead, 32))
        if (w.qtail - qhead >= 32) drain(qhead + ((w.qtail - qhead) / NSL) * NSL);
    if (w.qtail > qhead) drain(w.qtail);
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], (unsigned long long)w.err); return; }
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
// Tile layout, FP32: SW/2 pairs of 8 floats {x0 x1 y0 y1 | z0 z1 w0 w1}.  With the long/short split (LS) the weights are
// not used (pn2_p2p.cuh), and the 8 spare bytes of pairs 0..3 carry the leaf's HEADER: its box centre (3 doubles) and its
// particle count -- everything a sink needs to know about a source leaf arrives with the one copy of its tile.  The
// Newtonian build keeps the weights and appends the 32-byte header to the tile.
template <int SW, bool LS>
struct FusedLayout {
    static constexpr int NSL = 32 / SW;
    static constexpr int NST = FUSED_NST;                 // stages per batch
    static constexpr int BATCH = NST * NSL;               // source leaves per batch (32 / 16 / 8)
    static constexpr int TB = 16 * SW;                    // particle bytes of a tile
    static constexpr int TBX = LS ? TB : TB + 32;         // tile stride in HBM and bytes staged per leaf
    static constexpr int OFF = TBX;                       // in the staged row: {-, dx, dy, dz}, the leaf's centre offset from the sink leaf
    static constexpr int ROWB = TBX + 16;                 // row stride (144 B with LS: the NSL broadcast rows fall into distinct banks)
    static constexpr int STAGE_BYTES = BATCH * ROWB;
    __host__ __device__ static constexpr int hdr(int i) { return LS ? 32 * i + 24 : TB + 8 * i; }   // byte offset of header word i
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
    using FL = FusedLayout<SW, LS>;
    constexpr int NSL = FL::NSL, NST = FL::NST, BATCH = FL::BATCH, TB = FL::TB, TBX = FL::TBX, ROWB = FL::ROWB;
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ unsigned s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ __align__(16) unsigned char s_stage[WALK_WARPS][FL::STAGE_BYTES];
    __shared__ double s_sink[WALK_WARPS][6];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int q = lane / SW, j = lane % SW;
    // the pass over the received trees visits the active leaves only (work_count != NULL), with a bounded grid that
    // strides over their list; otherwise one warp per leaf and the loop runs once
    int item0 = a.item_lo, nitem = a.item_hi;
    if (a.work_count && (int)*a.work_count < nitem) nitem = (int)*a.work_count;
    for (int item = item0 + blockIdx.x * WALK_WARPS + wib; item < nitem; item += gridDim.x * WALK_WARPS) {
    const int leaf = a.work_count ? a.work[item] : item;
    LeafWalk<SRCQ_CAP, true> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    const LeafDesc sd = w.sd;
    // sink: slot j of the leaf's own tile (padding slots compute, but are never written)
    float xi, yi, zi;
    {
        const float *t = a.tiles + (size_t)leaf * (TBX / 4) + (j >> 1) * 8 + (j & 1);
        xi = t[0]; yi = t[2]; zi = t[4];
    }
    P2PSinkPk sk;
    sk.nx = sk.ny = sk.nz = sk.ax = sk.ay = sk.az = pk2(0.f, 0.f);
    const float inv_eps = pc.inv_eps;

    int qhead = 0, inflight = 0;
    unsigned nsrc = 0;
    unsigned char *stage = s_stage[wib];
    const unsigned stage_dst = (unsigned)__cvta_generic_to_shared(stage) + q * ROWB + j * 16;   // this lane's 16-byte chunk of row q
    const char *tile_src = reinterpret_cast<const char *>(a.tiles) + j * 16;
    // the tile copies of cnt <= BATCH queue entries from qhead (a multiple of NSL): one LDS, one address, one LDGSTS per stage
    auto issue_stage = [&](int s, unsigned te) {
        const char *src = tile_src + (size_t)(te & PN2_CELL_MASK) * TBX;
        cp_async16(stage_dst + s * NSL * ROWB, src);
        if (!LS && j < 2) cp_async16(stage_dst + s * NSL * ROWB + TB, src + TB);
    };
    auto issue_batch = [&](int cnt) {
        if (cnt == BATCH) {                // the queue entries first (the copies are ordered against shared-memory reads)
            unsigned te[NST];
#pragma unroll
            for (int s = 0; s < NST; s++) te[s] = w.queue[(qhead + s * NSL + q) & (SRCQ_CAP - 1)];
#pragma unroll
            for (int s = 0; s < NST; s++) issue_stage(s, te[s]);
        } else {
#pragma unroll 1
            for (int s = 0; s * NSL < cnt; s++) issue_stage(s, w.queue[(qhead + s * NSL + q) & (SRCQ_CAP - 1)]);
        }
        cp_async_commit();
        inflight = cnt;
        qhead += cnt;
    };
    auto compute_stage = [&](int s) {
        const float *row = reinterpret_cast<const float *>(stage + (s * NSL + q) * ROWB);
        const float4 o = *reinterpret_cast<const float4 *>(row + FL::OFF / 4);
        const float nx = o.y - xi, ny = o.z - yi, nz = o.w - zi;          // x_j + (centre offset - x_i)
        sk.nx = pk2(nx, nx); sk.ny = pk2(ny, ny); sk.nz = pk2(nz, nz);
        pk_row<SW, LS>(row, 0, sk, inv_eps);
    };
    auto compute_batch = [&]() {
        cp_async_wait_all();
        __syncwarp();
        // lane r resolves row r from the header that came with the tile: centre offset of the source leaf from the sink
        // leaf (FP64, rounded once; units of lambda) and the leaf's particle count
        if (lane < inflight) {
            const unsigned te = w.queue[(qhead - inflight + lane) & (SRCQ_CAP - 1)];
            unsigned char *row = stage + lane * ROWB;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (te != (unsigned)a.pad_tile) {
                const unsigned img = te >> PN2_IMG_SHIFT;
                const double cx = *reinterpret_cast<const double *>(row + FL::hdr(0));
                const double cy = *reinterpret_cast<const double *>(row + FL::hdr(1));
                const double cz = *reinterpret_cast<const double *>(row + FL::hdr(2));
                o.y = (float)(((cx + pc.shift[img][0]) - sd.c[0]) * pc.inv_len);
                o.z = (float)(((cy + pc.shift[img][1]) - sd.c[1]) * pc.inv_len);
                o.w = (float)(((cz + pc.shift[img][2]) - sd.c[2]) * pc.inv_len);
                nsrc += (unsigned)(*reinterpret_cast<const int *>(row + FL::hdr(3)) - ((te == (unsigned)leaf) ? 1 : 0));
            }
            *reinterpret_cast<float4 *>(row + FL::OFF) = o;
        }
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
        while (walking && w.qtail - qhead < BATCH) walking = w.step(a, pc, lane, qhead, BATCH);
        if (inflight) compute_batch();
        const int avail = w.qtail - qhead;
        if (avail == 0 || w.err) break;                              // walking implies avail >= BATCH
        int cnt = BATCH;
        if (avail < BATCH) {                                           // the last batch: pad its last stage with the padding tile
            cnt = ((avail + NSL - 1) / NSL) * NSL;
            if (lane < cnt - avail) w.queue[(w.qtail + lane) & (SRCQ_CAP - 1)] = (unsigned)a.pad_tile;
            w.qtail = qhead + cnt;
            __syncwarp();
        }
        issue_batch(cnt);
    }
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], (unsigned long long)w.err); continue; }

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
    if (q == 0 && j < sd.npart) {
        const double sc = pc.mass * pc.inv_len * pc.inv_len;
        double *o = a.acc + 3 * (size_t)(sd.first + j);
        o[0] += (double)ax * sc; o[1] += (double)ay * sc; o[2] += (double)az * sc;
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) nsrc += __shfl_xor_sync(0xffffffffu, nsrc, m);
    if (lane == 0) {
        atomicAdd(&a.counters[0], (unsigned long long)nsrc * (unsigned long long)sd.npart);
        atomicAdd(&a.counters[2], (unsigned long long)w.npairs);
        atomicAdd(&a.counters[4], (unsigned long long)w.visits);
    }
    __syncwarp();
    }   // leaves of this warp
}

// ------------------------------------------------------------------------------------------------
// FP64 product kernel: the same fused walk + P2P with double tiles and no libm
// ------------------------------------------------------------------------------------------------
// Tiles: SW slots of {x, y, z, w} doubles (32 bytes), leaf-centre-relative in units of 2 rs (so u = r), w = 1, padding
// slots far away with w = 0.  Arithmetic per interaction (reference: src/fmm.c:834-852, sqrt + division + erfc + exp
// from libm): 3 DADD + 3 DFMA (r^2) + MUFU.RSQ64H and one Newton step (4 ops, relative error < 1e-12) + 1 DSETP
// (softening) + 2 DMUL (1/r^3) + DMUL (u) + 4 ops for the interval of the g(u) table (2^52 + 2^51 rounding: no F2I / I2F)
// + 6 DFMA (degree-6 piece of g on 31 intervals, |error| < 1.9e-10: pn2_gtab.h, tools/fit_g64.py) + DMUL + 3 DFMA
// = 28 FP64-pipe operations.
#include "pn2_gtab.h"
#ifndef F64_MIN_BLOCKS
#define F64_MIN_BLOCKS 4
#endif
#define F64_NST 4
#ifndef F64_MAGIC
#define F64_MAGIC 1
#endif
// FP64 tiles: SW slots of {x, y, z, w} doubles.  With the long/short split the weight is not read, and the w of slots
// 0..3 carries the leaf's header {centre x, y, z, particle count}; the Newtonian build appends the 32-byte header.
template <int SW, bool LS>
struct Fused64Layout {
    static constexpr int NSL = 32 / SW;
    static constexpr int NST = F64_NST;
    static constexpr int BATCH = NST * NSL;               // source leaves per batch (16 / 8 / 4)
    static constexpr int TB = 32 * SW;                    // particle bytes of a tile
    static constexpr int TBX = LS ? TB : TB + 32;         // tile stride in HBM and bytes staged per leaf
    static constexpr int OFF = TBX;                       // in the staged row: {-, dx, dy, dz} (doubles): centre offset from the sink leaf
    static constexpr int ROWB = LS ? TBX + 32 : TBX + 64; // row stride (288 / 544 / 1056 B with LS: the NSL broadcast rows fall into distinct banks)
    static constexpr int STAGE_BYTES = BATCH * ROWB;
    __host__ __device__ static constexpr int hdr(int i) { return LS ? 32 * i + 24 : TB + 8 * i; }
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
    using FL = Fused64Layout<SW, LS>;
    constexpr int NSL = FL::NSL, NST = FL::NST, BATCH = FL::BATCH, TB = FL::TB, TBX = FL::TBX, ROWB = FL::ROWB;
    __shared__ unsigned s_stack[WALK_WARPS][STACK_CAP];
    __shared__ unsigned s_srcq[WALK_WARPS][SRCQ_CAP];
    __shared__ __align__(16) unsigned char s_stage[WALK_WARPS][FL::STAGE_BYTES];
    __shared__ double s_sink[WALK_WARPS][6];
    __shared__ __align__(128) double s_gtab[PN2_GTAB_DEG + 1][PN2_GTAB_K];
    for (int i = threadIdx.x; i < (PN2_GTAB_DEG + 1) * PN2_GTAB_K; i += blockDim.x) (&s_gtab[0][0])[i] = a.gtab[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int q = lane / SW, j = lane % SW;
    int item0 = a.item_lo, nitem = a.item_hi;
    if (a.work_count && (int)*a.work_count < nitem) nitem = (int)*a.work_count;
    for (int item = item0 + blockIdx.x * WALK_WARPS + wib; item < nitem; item += gridDim.x * WALK_WARPS) {
    const int leaf = a.work_count ? a.work[item] : item;
    LeafWalk<SRCQ_CAP, true> w;
    w.begin(a, leaf, s_stack[wib], s_srcq[wib], s_sink[wib], lane);
    const LeafDesc sd = w.sd;
    double xi, yi, zi;
    {
        const double *t = reinterpret_cast<const double *>(reinterpret_cast<const char *>(a.tiles64) + (size_t)leaf * TBX) + j * 4;
        xi = t[0]; yi = t[1]; zi = t[2];
    }
    P2PSink64 sk;
    sk.nx = sk.ny = sk.nz = 0.0;
    sk.ax[0] = sk.ax[1] = sk.ay[0] = sk.ay[1] = sk.az[0] = sk.az[1] = 0.0;
    const double eps = pc.soft * pc.inv2rs;
    const double eps2 = eps * eps, inv_eps = eps > 0.0 ? 1.0 / eps : 1e100;

    int qhead = 0, inflight = 0;
    unsigned nsrc = 0;
    unsigned char *stage = s_stage[wib];
    const unsigned stage_dst = (unsigned)__cvta_generic_to_shared(stage) + q * ROWB + j * 16;   // first of this lane's two chunks
    const char *tile_src = reinterpret_cast<const char *>(a.tiles64) + j * 16;
    auto issue_batch = [&](int cnt) {      // cnt <= BATCH queue entries from qhead, a multiple of NSL
#pragma unroll
        for (int s = 0; s < NST; s++) {
            if (s * NSL < cnt) {
                const unsigned te = w.queue[(qhead + s * NSL + q) & (SRCQ_CAP - 1)];
                const char *src = tile_src + (size_t)(te & PN2_CELL_MASK) * TBX;
                cp_async16(stage_dst + s * NSL * ROWB, src);
                cp_async16(stage_dst + s * NSL * ROWB + SW * 16, src + SW * 16);
                if (!LS && j < 2) cp_async16(stage_dst + s * NSL * ROWB + TB, src + TB);
            }
        }
        cp_async_commit();
        inflight = cnt;
        qhead += cnt;
    };
    auto compute_stage = [&](int s) {
        const double *row = reinterpret_cast<const double *>(stage + (s * NSL + q) * ROWB);
        const double *hd = row + FL::OFF / 8;
        sk.nx = hd[1] - xi; sk.ny = hd[2] - yi; sk.nz = hd[3] - zi;
#pragma unroll
        for (int k = 0; k < SW; k++) p2p_interact_tab64<LS>(row + 4 * k, sk, k & 1, eps2, inv_eps, s_gtab);
    };
    auto compute_batch = [&]() {
        cp_async_wait_all();
        __syncwarp();
        if (lane < inflight) {             // lane r resolves row r from the header that came with the tile
            const unsigned te = w.queue[(qhead - inflight + lane) & (SRCQ_CAP - 1)];
            unsigned char *row = stage + lane * ROWB;
            double ox = 0.0, oy = 0.0, oz = 0.0;
            if (te != (unsigned)a.pad_tile) {
                const unsigned img = te >> PN2_IMG_SHIFT;
                ox = ((*reinterpret_cast<const double *>(row + FL::hdr(0)) + pc.shift[img][0]) - sd.c[0]) * pc.inv2rs;
                oy = ((*reinterpret_cast<const double *>(row + FL::hdr(1)) + pc.shift[img][1]) - sd.c[1]) * pc.inv2rs;
                oz = ((*reinterpret_cast<const double *>(row + FL::hdr(2)) + pc.shift[img][2]) - sd.c[2]) * pc.inv2rs;
                nsrc += (unsigned)(*reinterpret_cast<const int *>(row + FL::hdr(3)) - ((te == (unsigned)leaf) ? 1 : 0));
            }
            double *hd = reinterpret_cast<double *>(row + FL::OFF);
            hd[1] = ox; hd[2] = oy; hd[3] = oz;
        }
        __syncwarp();
#pragma unroll 1
        for (int s = 0; s * NSL < inflight; s++) compute_stage(s);
        inflight = 0;
        __syncwarp();
    };
    bool walking = true;
    while (true) {
        while (walking && w.qtail - qhead < BATCH) walking = w.step(a, pc, lane, qhead, BATCH);
        if (inflight) compute_batch();
        const int avail = w.qtail - qhead;
        if (avail == 0 || w.err) break;
        int cnt = BATCH;
        if (avail < BATCH) {
            cnt = ((avail + NSL - 1) / NSL) * NSL;
            if (lane < cnt - avail) w.queue[(w.qtail + lane) & (SRCQ_CAP - 1)] = (unsigned)a.pad_tile;
            w.qtail = qhead + cnt;
            __syncwarp();
        }
        issue_batch(cnt);
    }
    if (w.err) { if (lane == 0) atomicOr(&a.counters[3], (unsigned long long)w.err); continue; }

    double ax = sk.ax[0] + sk.ax[1], ay = sk.ay[0] + sk.ay[1], az = sk.az[0] + sk.az[1];
#pragma unroll
    for (int m = SW; m < 32; m <<= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, m);
        ay += __shfl_xor_sync(0xffffffffu, ay, m);
        az += __shfl_xor_sync(0xffffffffu, az, m);
    }
    if (q == 0 && j < sd.npart) {
        const double sc = pc.mass * pc.inv2rs * pc.inv2rs;
        double *o = a.acc + 3 * (size_t)(sd.first + j);
        o[0] += ax * sc; o[1] += ay * sc; o[2] += az * sc;
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) nsrc += __shfl_xor_sync(0xffffffffu, nsrc, m);
    if (lane == 0) {
        atomicAdd(&a.counters[0], (unsigned long long)nsrc * (unsigned long long)sd.npart);
        atomicAdd(&a.counters[2], (unsigned long long)w.npairs);
        atomicAdd(&a.counters[4], (unsigned long long)w.visits);
    }
    __syncwarp();
    }   // leaves of this warp
}

// FP64 tiles: slot j of tile t <- (pos - leaf centre) / (2 rs) of particle j, w = 1; padding far away with w = 0
template <int SW, bool LS>
__global__ void tile64_kernel(int nt, int nleaf, int rleaf0, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                              double inv2rs, double *__restrict__ tiles) {
    using FL = Fused64Layout<SW, LS>;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = (int)(t / SW), j = (int)(t % SW);
    if (tile > nt) return;
    double4 p = make_double4(1e4, 1e4, 1e4, 0.0);
    double hdr = 0.0;                                       // header word j (j < 4): centre x, y, z, particle count
    if (tile < nt) {
        const LeafDesc d = desc[tile < nleaf ? tile : tile - nleaf + rleaf0];
        if (j < d.npart) {
            const double *x = pos + 3 * (size_t)(d.first + j);
            p = make_double4((x[0] - d.c[0]) * inv2rs, (x[1] - d.c[1]) * inv2rs, (x[2] - d.c[2]) * inv2rs, 1.0);
        }
        if (j < 3) hdr = d.c[j];
        else if (j == 3) hdr = __hiloint2double(0, d.npart);
    }
    char *base = reinterpret_cast<char *>(tiles) + (size_t)tile * FL::TBX;
    double *o = reinterpret_cast<double *>(base) + 4 * j;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
    if (!LS) o[3] = p.w;
    if (j < 4) *reinterpret_cast<double *>(base + FL::hdr(j)) = hdr;
}

// Leaf tiles (FP32 mode): slot j of tile t <- particle j of the leaf, packed-pair layout; tile nt = all padding.
// Tiles 0..nleaf-1 are the local leaves, nleaf.. the received LET leaves (cells rleaf0..).
// With the long/short split the tile layout carries no weight (pn2_p2p.cuh): a padding slot is harmless because
// 2^(-r'^2) flushes to 0 at its distance, which holds while every real particle stays within PN2_PAD_SAFE lambda of
// its leaf centre per dimension (|sink - pad| >= sqrt(3) (24 - 3 B - 2.71) >= 11.3 lambda for B <= 4.9; leaf pairs
// are listed only when their boxes are closer than the cut-off, 2.71 lambda).  Wider leaves (NSIDE far finer than the
// particle spacing) are reported instead of being evaluated wrongly: counters[3] |= 4.
#define PN2_PAD_SAFE 4.9f
template <int SW, bool LS>
__global__ void tile_kernel(int nt, int nleaf, int rleaf0, int t0, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                            const float4 *__restrict__ rel, double inv_len, float *__restrict__ tiles,
                            unsigned long long *__restrict__ counters) {
    using FL = FusedLayout<SW, LS>;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int tile = t0 + (int)(t / SW), j = (int)(t % SW);
    if (tile > nt) return;
    float4 p = make_float4(PN2_PAD_COORD, PN2_PAD_COORD, PN2_PAD_COORD, 0.f);
    double hdr = 0.0;                                       // header word j (j < 4): centre x, y, z, particle count; zeros for the padding tile
    if (tile < nt) {
        const LeafDesc d = desc[tile < nleaf ? tile : tile - nleaf + rleaf0];
        if (j < d.npart) {
            if (tile < nleaf) {
                // a local leaf: leaf-centre-relative coordinates in units of lambda, straight from the FP64 positions
                const double *x = pos + 3 * (size_t)(d.first + j);
                p = make_float4((float)((x[0] - d.c[0]) * inv_len), (float)((x[1] - d.c[1]) * inv_len), (float)((x[2] - d.c[2]) * inv_len), 1.f);
            } else p = rel[d.first + j];                     // a received leaf: its sender shipped these very values
            if (LS && fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)) > PN2_PAD_SAFE) atomicOr(&counters[3], 4ULL);
        }
        if (j < 3) hdr = d.c[j];
        else if (j == 3) hdr = __hiloint2double(0, d.npart);
    }
    char *base = reinterpret_cast<char *>(tiles) + (size_t)tile * FL::TBX;
    float *o = reinterpret_cast<float *>(base) + (j >> 1) * 8 + (j & 1);
    o[0] = p.x; o[2] = p.y; o[4] = p.z;
    if (!LS) o[6] = p.w;
    if (j < 4) *reinterpret_cast<double *>(base + FL::hdr(j)) = hdr;
}

template <int SW>
static void launch_mode(pn2_ctx *h, const WalkArgs &a, int mode, cudaStream_t st, unsigned grid) {
    if (mode == 0 && h->prm.longshort) walk_fused_kernel<SW, true><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    else if (mode == 0) walk_fused_kernel<SW, false><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    else if (mode == 3 && h->prm.longshort) walk_fused_f64_kernel<SW, true><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    else if (mode == 3) walk_fused_f64_kernel<SW, false><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    else if (mode == 1) walk_leaf_kernel<SW, 1><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    else walk_leaf_kernel<SW, 2><<<grid, WALK_WARPS * 32, 0, st>>>(a, h->pc);
    h->launches++;
}
// the leaf kernel of `mode` over a's item range on stream st; grid = 0: one warp per item of [item_lo, item_hi)
static void launch_leaves(pn2_ctx *h, const WalkArgs &a, int mode, cudaStream_t st, unsigned grid) {
    if (grid == 0) {
        const int items = a.item_hi - a.item_lo;
        if (items <= 0) return;
        grid = (unsigned)((items + WALK_WARPS - 1) / WALK_WARPS);
        const unsigned gcap = (unsigned)(h->sm_count > 0 ? h->sm_count : 148) * LEAF_MIN_BLOCKS * 4;
        if (a.work_count && (mode == 0 || mode == 3) && grid > gcap) grid = gcap;      // active leaves only: a bounded grid strides over their list
    }
    const int ml = h->prm.maxleaf;
    if (ml <= 8) launch_mode<8>(h, a, mode, st, grid);
    else if (ml <= 16) launch_mode<16>(h, a, mode, st, grid);
    else launch_mode<32>(h, a, mode, st, grid);
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
    a.item_lo = 0; a.item_hi = h->nleaf;
}

// one frontier launch over `cnt` work items (a.work / a.nwork set by the caller); gcap > 0 bounds the grid (the kernel strides)
static void launch_frontier(pn2_ctx *h, WalkArgs &a, int cnt, unsigned gcap) {
    a.nwork = cnt;
    unsigned grid = (unsigned)((cnt + WALK_WARPS - 1) / WALK_WARPS);
    if (gcap > 0 && grid > gcap) grid = gcap;
    frontier_node_kernel<<<grid, WALK_WARPS * 32, 0, h->stream>>>(a, h->pc);
    h->launches++;
}

// Pass 1 over the levels lev0 .. lev1 - 1.  counters[6] is the span bump pointer (unit 0 is reserved: 0 = "no span").
// h->walk_active: the pass over the received trees -- levels and leaves are visited through active lists built on the way.
static int walk_frontier_levels(pn2_ctx *h, WalkArgs &a, int lev0, int lev1) {
    for (int lev = lev0; lev < lev1; lev++) {
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
        // the active list of a pass over the received trees is short (cells near the domain surface): a few waves of CTAs
        // stride over it instead of one CTA per four cells of the level, most of which would exit at once
        launch_frontier(h, a, cnt, h->walk_active ? (unsigned)(h->sm_count > 0 ? h->sm_count : 148) * NODE_MIN_BLOCKS * 2 : 0u);
    }
    return PN2_OK;
}

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
    PN2_TRY(walk_frontier_levels(h, a, 0, h->nlevel));
    KERNEL_CHECK();
    return PN2_OK;
}

// leaf tiles of the leaf kernel of `mode` (FP32: 0, FP64 tables: 3), built on the context's stream; sets a.tiles / tiles64 / pad_tile
static int prepare_tiles(pn2_ctx *h, WalkArgs &a, int mode) {
    const int ml = h->prm.maxleaf;
    if (mode == 3) {
        const int sw = ml <= 8 ? 8 : (ml <= 16 ? 16 : 32);
        const int nt = h->nleaf + h->nrl;
        const bool ls = h->prm.longshort != 0;
        PN2_TRY(h->tiles64.ensure(((size_t)nt + 1) * (4 * sw + (ls ? 0 : 4))));
        if (!h->gtab.p) {
            PN2_TRY(h->gtab.ensure((PN2_GTAB_DEG + 1) * PN2_GTAB_K));
            CUDA_TRY(cudaMemcpyAsync(h->gtab.p, PN2_GTAB, sizeof PN2_GTAB, cudaMemcpyHostToDevice, h->stream));
        }
        const long nthr = ((long)nt + 1) * sw;
        const unsigned g = (unsigned)((nthr + 255) / 256);
#define PN2_TILE64_LAUNCH(SWV, LSV) tile64_kernel<SWV, LSV><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, h->desc.p, h->pos.p, h->pc.inv2rs, h->tiles64.p)
        if (sw == 8) { if (ls) PN2_TILE64_LAUNCH(8, true); else PN2_TILE64_LAUNCH(8, false); }
        else if (sw == 16) { if (ls) PN2_TILE64_LAUNCH(16, true); else PN2_TILE64_LAUNCH(16, false); }
        else { if (ls) PN2_TILE64_LAUNCH(32, true); else PN2_TILE64_LAUNCH(32, false); }
#undef PN2_TILE64_LAUNCH
        h->launches++;
        a.tiles64 = h->tiles64.p; a.gtab = h->gtab.p; a.pad_tile = nt;
    }
    if (mode == 0) {
        // leaf tiles of the local and the received leaves (+ one padding tile).  The local tiles are built once per step
        // (the pass over the local roots); the pass over the received trees appends its tiles and a new padding tile.
        const int sw = ml <= 8 ? 8 : (ml <= 16 ? 16 : 32);
        const int nt = h->nleaf + h->nrl;
        const bool ls = h->prm.longshort != 0;
        const size_t need = ((size_t)nt + 1) * (4 * sw + (ls ? 0 : 8));  // floats: 16 SW bytes per tile (+ the 32-byte header without the split)
        const bool grows = need > h->tiles.cap;                         // a new buffer: every tile again
        PN2_TRY(h->tiles.ensure(need));
        const int t0 = (!grows && h->tiles_built_for == h->step_serial && h->nrl > 0) ? h->nleaf : 0;
        const long nthr = ((long)nt + 1 - t0) * sw;
        const unsigned g = (unsigned)((nthr + 255) / 256);
#define PN2_TILE_LAUNCH(SWV, LSV) tile_kernel<SWV, LSV><<<g, 256, 0, h->stream>>>(nt, h->nleaf, h->ncell, t0, h->desc.p, h->pos.p, h->rel.p, h->pc.inv_len, h->tiles.p, h->counters.p)
        if (sw == 8) { if (ls) PN2_TILE_LAUNCH(8, true); else PN2_TILE_LAUNCH(8, false); }
        else if (sw == 16) { if (ls) PN2_TILE_LAUNCH(16, true); else PN2_TILE_LAUNCH(16, false); }
        else { if (ls) PN2_TILE_LAUNCH(32, true); else PN2_TILE_LAUNCH(32, false); }
#undef PN2_TILE_LAUNCH
        h->launches++;
        h->tiles_built_for = h->step_serial;
        a.tiles = h->tiles.p; a.pad_tile = nt;
    }
    return PN2_OK;
}

// PN2_FP64: table-driven FP64 kernel (3); PN2_FP64_LIBM: the reference's expression with libm (1); PN2_FP32: 0; list dump: 2
static int leaf_mode(const pn2_ctx *h, int dump) {
    return dump ? 2 : (h->prm.precision == PN2_FP64 ? 3 : (h->prm.precision == PN2_FP64_LIBM ? 1 : 0));
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
    const int mode = leaf_mode(h, dump);
    PN2_TRY(prepare_tiles(h, a, mode));
    launch_leaves(h, a, mode, h->stream, 0);
    KERNEL_CHECK();
    return PN2_OK;
}
