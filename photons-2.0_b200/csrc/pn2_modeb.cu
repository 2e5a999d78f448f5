// pn2_modeb.cu -- Mode B orchestration: one whole short-range force evaluation on the device
// (fmm_construct + fmm_prepare + fmm_task + fmm_ext of src/photoNs.c:97-116, PM excluded):
//   tree build (pn2_tree.cu) -> P2M + M2M -> fused walk + P2P incl. periodic images (pn2_walk.cu)
//   -> M2L over the pairs the walk found -> L2L + L2P -> accelerations back in caller order.
#include <cub/cub.cuh>
#include "pn2_common.cuh"
#define PN2_NEV 16

void pn2_modeb_release(pn2_ctx *h) {
    h->order.release(); h->order_alt.release(); h->parent.release(); h->depth.release(); h->b_pay.release(); h->b_pay2.release(); h->b_qc.release(); h->b_qc2.release(); h->n_sum.release(); h->b_idx2.release();
    h->b_seg.release(); h->b_seg2.release(); h->b_q.release(); h->b_key2.release(); h->b_f.release(); h->b_flag.release();
    h->n_start.release(); h->n_count.release(); h->n_son.release(); h->n_depth.release(); h->l_start.release();
    h->l_count.release(); h->n_box.release(); h->n_split.release(); h->l_box.release(); h->b_cnt.release();
    h->b_scal.release(); h->b_lv.release(); h->act_nodes.release(); h->act_leaf.release(); h->act_count.release(); h->stage_in.release(); h->stage_out.release(); h->m2l_pairs.release(); h->spans.release(); h->o_head.release(); h->lst_off.release(); h->lst_src.release(); h->lst_sink.release();
    for (int i = 0; i < PN2_NEV; i++) if (h->ev[i]) { cudaEventDestroy(h->ev[i]); h->ev[i] = nullptr; }
    pn2_let_release(h);
    pn2_migrate_release(h);
    pn2_pm_release(h);
    pn2_comm_release(h);
}

__global__ void scatter_acc_kernel(int n, const double *__restrict__ acc, const int *__restrict__ order,
                                   double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t o = (size_t)order[i];
    out[3 * o] = acc[3 * (size_t)i];
    out[3 * o + 1] = acc[3 * (size_t)i + 1];
    out[3 * o + 2] = acc[3 * (size_t)i + 2];
}
// inspection only: the cells the step never gave an expansion read as zero
__global__ void clear_unset_l_kernel(long total, const unsigned char *__restrict__ has_l, double *__restrict__ L) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total && !has_l[t / NM]) L[t] = 0.0;
}
__global__ void iota_kernel(int n, int *o) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = i;
}

static int ensure_events(pn2_ctx *h) {
    for (int i = 0; i < PN2_NEV; i++)
        if (!h->ev[i]) CUDA_TRY(cudaEventCreate(&h->ev[i]));
    return PN2_OK;
}

// list-builder pools: M2L pair buffer, per-cell list heads, span buffer (96 B per particle to start with; grown on demand)
static int ensure_walk_buffers(pn2_ctx *h) {
    if (h->m2l_cap == 0) {
        size_t cap = 4u << 20;
        PN2_TRY(h->m2l_pairs.ensure(2 * cap));
        h->m2l_cap = cap;
    }
    PN2_TRY(h->o_head.ensure((size_t)h->ncell + (size_t)h->nrl + (size_t)h->nrn + 1));
    if (h->span_cap16 < 1024 + 6ULL * (unsigned long long)h->n) {
        h->span_cap16 = 1024 + 6ULL * (unsigned long long)h->n;
        h->spans.release();
        PN2_TRY(h->spans.ensure(4 * (size_t)h->span_cap16));
    }
    return PN2_OK;
}

// One pass of the list builder + fused P2P over the sink tree for one set of source roots, enqueued on the context's
// stream (no host synchronisation): which = 0: the local tree and its periodic images; which = 1: the received trees
// (and their images).  The walk of a (sink, source root) pair is independent of every other root, so the two passes
// together visit exactly the pairs of one pass over all roots; accelerations, M2L pairs and counters accumulate.
// Pass 0 needs nothing from the peers: it runs while the LET blocks travel (pn2_let.cu).
static int walk_pass_enqueue(pn2_ctx *h, int which, int ev_frontier, int ev_fused) {
    cudaStream_t st = h->stream;
    h->walk_active = which == 1;
    PN2_TRY(pn2_walk_set_roots(h, which));
    if (which == 0) CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), st));
    h->top0_host = 1 + (unsigned long long)h->root_units;               // unit 0 reserved (0 = empty list), then F(root)
    CUDA_TRY(cudaMemcpyAsync(h->counters.p + 6, &h->top0_host, sizeof h->top0_host, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(h->o_head.p, 0, ((size_t)h->ncell + 1) * sizeof(unsigned), st));
    if (h->root_count > 0) PN2_TRY(pn2_walk_frontiers(h));
    CUDA_TRY(cudaEventRecord(h->ev[ev_frontier], st));
    if (h->root_count > 0) PN2_TRY(pn2_walk_fused(h, 0));
    CUDA_TRY(cudaEventRecord(h->ev[ev_fused], st));
    return PN2_OK;
}

// Phase 1 of a step: tree, upward pass, the walk over the local tree and its images (enqueued, not awaited) and
// (multi-rank) the LET packs for every peer on the LET stream
extern "C" int pn2_step_begin(pn2_ctx *h, const double *d_pos, int n, const pn2_domain *dom) {
    if (!h || n < 0 || !dom || (n > 0 && !d_pos)) { pn2_set_error("pn2_step_begin: bad argument"); return PN2_ERR_ARG; }
    for (int d = 0; d < 3; d++)
        if (!(dom->hi[d] > dom->lo[d])) { pn2_set_error("pn2_step_begin: empty domain box"); return PN2_ERR_ARG; }
    if (dom->direct0 < 0 || dom->direct0 > 2) { pn2_set_error("pn2_step_begin: direct0 must be 0..2"); return PN2_ERR_ARG; }
    if (h->nranks > 1) {
        // the LET receiver re-derives the sender's prune decisions from ITS box (pruned_dev, pn2_walk.cu): the box of this
        // step must be the one the peers were given for this rank (pn2_set_comm / pn2_comm_init_rank)
        if ((int)h->all_dom.size() != h->nranks) { pn2_set_error("pn2_step_begin: no domain table (pn2_set_comm)"); return PN2_ERR_STATE; }
        const pn2_domain &me = h->all_dom[h->rank];
        bool same = me.direct0 == dom->direct0;
        for (int d = 0; d < 3; d++) same = same && me.lo[d] == dom->lo[d] && me.hi[d] == dom->hi[d];
        if (!same) { pn2_set_error("pn2_step_begin: dom differs from all_domains[rank] of pn2_set_comm (stale domain table?)"); return PN2_ERR_ARG; }
    }
    CUDA_TRY(cudaSetDevice(h->device));
    PN2_TRY(ensure_events(h));
    cudaStream_t st = h->stream;
    h->have_step = false;
    h->step_open = false;
    h->let_unpacked = false;
    h->step_serial++;
    memset(&h->info, 0, sizeof h->info);
    CUDA_TRY(cudaEventRecord(h->ev[0], st));
    PN2_TRY(pn2_tree_build_device(h, d_pos, n, dom));
    CUDA_TRY(cudaEventRecord(h->ev[1], st));
    h->info.n = n; h->info.nleaf = h->nleaf; h->info.nnode = h->nnode; h->info.nlevel = h->nlevel;
    if (n > 0) {
        PN2_TRY(pn2_launch_p2m(h));
        PN2_TRY(pn2_launch_m2m(h));
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], st));
    if (h->nranks > 1) PN2_TRY(pn2_let_tree_ready(h));                  // what the LET stream waits for: NOT the walk enqueued next
    if (n > 0) {
        PN2_TRY(ensure_walk_buffers(h));
        PN2_TRY(walk_pass_enqueue(h, 0, 10, 11));                       // local tree + images: overlaps the LET exchange
    }
    if (h->nranks > 1) PN2_TRY(pn2_let_pack_all(h));                    // on the LET stream, after ev[2]
    h->step_open = true;
    return PN2_OK;
}

// Phase 2 (after the LET exchange): the walk over the received trees, M2L, downward pass, accelerations in caller order
extern "C" int pn2_step_finish(pn2_ctx *h, double *d_acc) {
    if (!h || !h->step_open || (h->n > 0 && !d_acc)) { pn2_set_error("pn2_step_finish: no open step / bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int n = h->n;
    h->step_open = false;
    if (n == 0) {
        if (h->nranks > 1) PN2_TRY(pn2_let_unpack(h));
        for (int i = 3; i < PN2_NEV; i++) CUDA_TRY(cudaEventRecord(h->ev[i], st));
        h->have_step = true;
        return PN2_OK;
    }
    unsigned long long cnt[8], span_used = 0;
    for (int attempt = 0;; attempt++) {
        if (attempt > 0) {                                             // a pool was too small: both passes again
            CUDA_TRY(cudaMemsetAsync(h->acc.p, 0, 3 * (size_t)n * sizeof(double), st));
            PN2_TRY(walk_pass_enqueue(h, 0, 10, 11));
        }
        CUDA_TRY(cudaMemcpyAsync(cnt, h->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        span_used = cnt[6];
        if (h->nranks > 1) {
            if (!h->let_unpacked) { PN2_TRY(pn2_let_unpack(h)); h->let_unpacked = true; }     // waits for the exchange (LET stream)
            CUDA_TRY(cudaEventRecord(h->ev[8], st));
            PN2_TRY(ensure_walk_buffers(h));                           // o_head covers the received cells too
            PN2_TRY(walk_pass_enqueue(h, 1, 12, 3));
            CUDA_TRY(cudaMemcpyAsync(cnt, h->counters.p, sizeof cnt, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (cnt[6] > span_used) span_used = cnt[6];
        } else {
            CUDA_TRY(cudaEventRecord(h->ev[8], st)); CUDA_TRY(cudaEventRecord(h->ev[12], st)); CUDA_TRY(cudaEventRecord(h->ev[3], st));
        }
        if (cnt[3] & 1) { pn2_set_error("pn2: walk stack overflow (pathological particle distribution)"); return PN2_ERR_NOMEM; }
        if (cnt[3] & 8) { pn2_set_error("pn2: a leaf's list walk did not terminate (watchdog)"); return PN2_ERR_CUDA; }
        if (cnt[3] & 4) {
            pn2_set_error("pn2: a leaf is wider than 9.8 lambda = 13.6 rs: the FP32 tile layout of the long/short split cannot hold it; use PN2_FP64");
            return PN2_ERR_ARG;
        }
        bool redo = false;
        if (cnt[3] & 2) {                                              // span buffer too small
            // a pass that ran out of room truncates the deeper levels' lists, so the counted need can be an underestimate:
            // grow at least geometrically so that the retries converge
            unsigned long long want = span_used + span_used / 4 + 1024;
            if (want < 2 * h->span_cap16) want = 2 * h->span_cap16;
            if (want >= (1ULL << 32)) want = (1ULL << 32) - 1;
            if (want <= h->span_cap16) { pn2_set_error("pn2: frontier lists exceed 64 GB"); return PN2_ERR_NOMEM; }
            h->span_cap16 = want;
            h->spans.release();
            PN2_TRY(h->spans.ensure(4 * (size_t)h->span_cap16));
            redo = true;
        }
        if (cnt[1] > h->m2l_cap) {                                     // M2L pair buffer too small
            size_t cap = (size_t)cnt[1] + (size_t)cnt[1] / 4 + 1024;
            if (redo && cap < 2 * h->m2l_cap) cap = 2 * h->m2l_cap;   // counted on truncated lists: may still be short
            h->m2l_pairs.release();
            PN2_TRY(h->m2l_pairs.ensure(2 * cap));
            h->m2l_cap = cap;
            redo = true;
        }
        if (!redo) break;
        if (attempt >= 8) { pn2_set_error("pn2: interaction lists do not fit"); return PN2_ERR_NOMEM; }
    }
    h->walk_active = false;
    h->span_used16 = span_used;
    h->walk_visits = cnt[4];
    h->info.n_walk_visits = (int64_t)cnt[4]; h->info.frontier_bytes = (int64_t)(16 * span_used);
    h->info.n_interactions = (int64_t)cnt[0]; h->info.n_m2l_pairs = (int64_t)cnt[1]; h->info.n_p2p_pairs = (int64_t)cnt[2];
    // M2L
    if (cnt[1] > 0) {
        CsrList list;
        PN2_TRY(pn2_csr_from_device_pairs(h, (int *)h->m2l_pairs.p, h->m2l_pairs.p + h->m2l_cap, (long)cnt[1], &list));
        CUDA_TRY(cudaEventRecord(h->ev[9], st));                       // the M2L kernel alone (after the list sort)
        PN2_TRY(pn2_launch_m2l(h, list, h->geom.p, h->M.p));
    } else CUDA_TRY(cudaEventRecord(h->ev[9], st));
    CUDA_TRY(cudaEventRecord(h->ev[4], st));
    PN2_TRY(pn2_launch_l2l_l2p(h));
    scatter_acc_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, h->acc.p, h->order.p, d_acc);
    h->launches++;
    CUDA_TRY(cudaEventRecord(h->ev[5], st));
    KERNEL_CHECK();
    h->have_step = true;
    return PN2_OK;
}

extern "C" int pn2_force_step_device(pn2_ctx *h, const double *d_pos, int n, const pn2_domain *dom, double *d_acc) {
    if (h && n > 0 && !d_acc) { pn2_set_error("pn2_force_step_device: bad argument"); return PN2_ERR_ARG; }
    PN2_TRY(pn2_step_begin(h, d_pos, n, dom));
    if (h->nranks > 1) PN2_TRY(pn2_let_exchange_nccl(h));
    return pn2_step_finish(h, d_acc);
}

extern "C" int pn2_force_step(pn2_ctx *h, const double *pos, size_t pos_stride, int n, const pn2_domain *dom, double *acc,
                              size_t acc_stride) {
    if (!h || n < 0 || (n > 0 && (!pos || !acc)) || pos_stride < 24 || acc_stride < 24 || pos_stride % 8 || acc_stride % 8) {
        pn2_set_error("pn2_force_step: bad argument");
        return PN2_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    DBuf<double> &in = h->stage_in, &out = h->stage_out;      // owned by the context: right device, freed by pn2_destroy
    PN2_TRY(in.ensure(3 * (size_t)n + 3)); PN2_TRY(out.ensure(3 * (size_t)n + 3));
    if (n > 0) {
        if (pos_stride == 24) CUDA_TRY(cudaMemcpyAsync(in.p, pos, 24 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        else CUDA_TRY(cudaMemcpy2DAsync(in.p, 24, pos, pos_stride, 24, n, cudaMemcpyHostToDevice, h->stream));
    }
    PN2_TRY(pn2_force_step_device(h, in.p, n, dom, out.p));
    if (n > 0) {
        if (acc_stride == 24) CUDA_TRY(cudaMemcpyAsync(acc, out.p, 24 * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        else CUDA_TRY(cudaMemcpy2DAsync(acc, acc_stride, out.p, 24, 24, n, cudaMemcpyDeviceToHost, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PN2_OK;
}

extern "C" int pn2_set_comm(pn2_ctx *h, int rank, int nranks, const pn2_domain *all, void *comm) {
    if (!h || nranks < 1 || rank < 0 || rank >= nranks || !all) { pn2_set_error("pn2_set_comm: bad argument"); return PN2_ERR_ARG; }
    h->rank = rank; h->nranks = nranks;
    if (comm) { h->nccl = comm; h->own_comm = false; }      // NULL keeps the communicator already attached
    h->all_dom.assign(all, all + nranks);
    return PN2_OK;
}

static int need_step(pn2_ctx *h, const char *who) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    if (!h->have_step) { pn2_set_error("%s: no completed pn2_force_step on this context", who); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    return PN2_OK;
}

extern "C" int pn2_get_step_info(pn2_ctx *h, pn2_step_info *info) {
    PN2_TRY(need_step(h, "pn2_get_step_info"));
    if (!info) { pn2_set_error("pn2_get_step_info: null"); return PN2_ERR_ARG; }
    *info = h->info;
    return PN2_OK;
}

extern "C" int pn2_get_order(pn2_ctx *h, int *order, int n) {
    PN2_TRY(need_step(h, "pn2_get_order"));
    if (n != h->n || (!order && n > 0)) { pn2_set_error("pn2_get_order: n mismatch"); return PN2_ERR_ARG; }
    if (n > 0) CUDA_TRY(cudaMemcpyAsync(order, h->order.p, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PN2_OK;
}

extern "C" int pn2_get_cells(pn2_ctx *h, double *geom, int *son, int *range, double *M, double *L) {
    PN2_TRY(need_step(h, "pn2_get_cells"));
    size_t nc = (size_t)h->ncell;
    cudaStream_t st = h->stream;
    if (nc == 0) return PN2_OK;
    if (geom) CUDA_TRY(cudaMemcpyAsync(geom, h->geom.p, 6 * nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (son) CUDA_TRY(cudaMemcpyAsync(son, h->son.p, 2 * nc * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (M) CUDA_TRY(cudaMemcpyAsync(M, h->M.p, NM * nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (L) {
        if (h->use_lflags) {
            clear_unset_l_kernel<<<(unsigned)((NM * nc + 255) / 256), 256, 0, st>>>((long)(NM * nc), h->has_l.p, h->L.p);
            CUDA_TRY(cudaMemsetAsync(h->has_l.p, 1, nc, st));            // every cell holds a valid expansion now
        }
        CUDA_TRY(cudaMemcpyAsync(L, h->L.p, NM * nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (range) {
        std::vector<LeafDesc> d(nc);
        CUDA_TRY(cudaMemcpyAsync(d.data(), h->desc.p, nc * sizeof(LeafDesc), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (size_t c = 0; c < nc; c++) { range[2 * c] = d[c].first; range[2 * c + 1] = d[c].npart; }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return PN2_OK;
}

extern "C" int pn2_get_lists(pn2_ctx *h, int kind, long *nseg, long *nsrc, int *seg_sink, long *seg_off, int *src) {
    PN2_TRY(need_step(h, "pn2_get_lists"));
    if (!nseg || !nsrc || (kind != 0 && kind != 1)) { pn2_set_error("pn2_get_lists: bad argument"); return PN2_ERR_ARG; }
    cudaStream_t st = h->stream;
    if (kind == 1) {
        long n = (long)h->info.n_m2l_pairs;
        CsrList list;
        list.nseg = 0;
        if (n > 0) {
            // sort a copy: the pair buffer keeps its emission order
            PN2_TRY(h->id_.ensure(n + 2)); PN2_TRY(h->ua.ensure(n));
            CUDA_TRY(cudaMemcpyAsync(h->id_.p, h->m2l_pairs.p, n * sizeof(int), cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(h->ua.p, h->m2l_pairs.p + h->m2l_cap, n * sizeof(unsigned), cudaMemcpyDeviceToDevice, st));
            PN2_TRY(pn2_csr_from_device_pairs(h, h->id_.p, h->ua.p, n, &list));
        }
        *nseg = list.nseg; *nsrc = n;
        if (seg_sink && seg_off && src && list.nseg > 0) {
            CUDA_TRY(cudaMemcpyAsync(seg_sink, list.seg_sink, list.nseg * sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(seg_off, list.seg_off, (list.nseg + 1) * sizeof(long), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(src, list.src, n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        } else if (seg_off) seg_off[0] = 0;
        CUDA_TRY(cudaStreamSynchronize(st));
        return PN2_OK;
    }
    // P2P lists are never materialised by the product step: replay the walk in dump mode (count, scan, fill)
    if (h->nranks > 1) { pn2_set_error("pn2_get_lists: the P2P list dump replays one walk pass; single-rank contexts only"); return PN2_ERR_STATE; }
    int nl = h->nleaf;
    long total = (long)h->info.n_p2p_pairs;
    *nseg = nl; *nsrc = total;
    if (!seg_sink || !seg_off || !src) return PN2_OK;
    PN2_TRY(h->lst_off.ensure(nl + 2)); PN2_TRY(h->lst_src.ensure(total + 1));
    CUDA_TRY(cudaMemsetAsync(h->lst_off.p, 0, (nl + 2) * sizeof(long), st));
    PN2_TRY(pn2_walk_fused(h, 1));
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, h->lst_off.p, h->lst_off.p, nl + 1, st);
    PN2_TRY(h->tmp.ensure(tb + 16));
    cub::DeviceScan::ExclusiveSum(h->tmp.p, tb, h->lst_off.p, h->lst_off.p, nl + 1, st);
    h->launches++;
    PN2_TRY(pn2_walk_fused(h, 2));
    CUDA_TRY(cudaMemcpyAsync(seg_off, h->lst_off.p, (nl + 1) * sizeof(long), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(src, h->lst_src.p, total * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < nl; i++) seg_sink[i] = i;
    if (seg_off[nl] != total) { pn2_set_error("pn2_get_lists: replay found %ld pairs, the step %ld", seg_off[nl], total); return PN2_ERR_STATE; }
    return PN2_OK;
}

extern "C" int pn2_get_timings(pn2_ctx *h, double ms[8]) {
    PN2_TRY(need_step(h, "pn2_get_timings"));
    if (!ms) { pn2_set_error("pn2_get_timings: null"); return PN2_ERR_ARG; }
    for (int i = 0; i < 8; i++) ms[i] = 0.0;
    if (h->n == 0) return PN2_OK;
    CUDA_TRY(cudaEventSynchronize(h->ev[5]));
    float t = 0, t2 = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[1])); ms[0] = t;     // tree
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[1], h->ev[2])); ms[1] = t;     // upward
    // the walk runs in two passes: local tree + images (ev 2 -> 10 -> 11), then, once the LET blocks have arrived and are
    // unpacked (ev 11 -> 8: what of the exchange is NOT hidden behind the first pass), the received trees (ev 8 -> 12 -> 3)
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[11], h->ev[8])); ms[5] = t;    // LET: exposed wait + unpack
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[2], h->ev[10])); CUDA_TRY(cudaEventElapsedTime(&t2, h->ev[8], h->ev[12]));
    ms[7] = t + t2;                                                        // frontier passes (lists by sink node)
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[10], h->ev[11])); CUDA_TRY(cudaEventElapsedTime(&t2, h->ev[12], h->ev[3]));
    ms[2] = t + t2;                                                        // fused leaf walk + P2P
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[3], h->ev[4])); ms[3] = t;     // M2L
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[4], h->ev[5])); ms[4] = t;     // downward
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[5]));
    ms[6] = t;
    return PN2_OK;
}

extern "C" int pn2_get_timings_ex(pn2_ctx *h, double *ms, int cap) {
    if (!ms || cap < 8) { pn2_set_error("pn2_get_timings_ex: need room for at least 8 values"); return PN2_ERR_ARG; }
    PN2_TRY(pn2_get_timings(h, ms));
    for (int i = 8; i < cap; i++) ms[i] = 0.0;
    if (h->n == 0 || cap < 9) return PN2_OK;
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[9], h->ev[4])); ms[8] = t;     // M2L kernel without the list sort
    return PN2_OK;
}
