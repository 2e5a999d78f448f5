// pn2_api.cu -- C-ABI of libpn2gpu.so, Mode A (host tree + host lists, device operators).
// Entry points mirror the reference's batch interface (include/pn2gpu.h cites file:line for each).
#include <cub/cub.cuh>
#include <stdarg.h>
#include "pn2_common.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void pn2_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *pn2_last_error(void) { return g_err; }

void pn2_init_consts(pn2_ctx *h) {
    P2PConst &c = h->pc;
    memset(&c, 0, sizeof c);
    const pn2_params &p = h->prm;
    c.rs = p.rs; c.soft = p.soft; c.mass = p.mass;
    c.inv2rs = 1.0 / (2.0 * p.rs);
    c.inv_len = 1.0 / (2.0 * p.rs * 0.83255461115769775635);
    c.longshort = p.longshort;
    // image displacements in the order of src/fmm.c:1028-1037 (mi, mj, mk in -1..1, (0,0,0) skipped)
    int k = 1;
    for (int a = -1; a <= 1; a++)
        for (int b = -1; b <= 1; b++)
            for (int d = -1; d <= 1; d++) {
                if (a == 0 && b == 0 && d == 0) continue;
                c.shift[k][0] = a * p.box; c.shift[k][1] = b * p.box; c.shift[k][2] = d * p.box;
                k++;
            }
    double ie = (p.soft > 0.0) ? 1.0 / (c.inv_len * p.soft) : 1e12;
    if (ie > 1e12) ie = 1e12;
    c.inv_eps = (float)ie;
}

static int check_params(const pn2_params *p) {
    if (!p) { pn2_set_error("pn2: params is NULL"); return PN2_ERR_ARG; }
    if (p->maxleaf < 1 || p->maxleaf > 32) { pn2_set_error("pn2: maxleaf %d outside 1..32", p->maxleaf); return PN2_ERR_ARG; }
    if (!(p->rs > 0.0) || !(p->box > 0.0)) { pn2_set_error("pn2: rs and box must be positive"); return PN2_ERR_ARG; }
    if (p->precision != PN2_FP64 && p->precision != PN2_FP32 && p->precision != PN2_FP64_LIBM) { pn2_set_error("pn2: unknown precision %d", p->precision); return PN2_ERR_ARG; }
    return PN2_OK;
}

extern "C" int pn2_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *mem_bytes) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        pn2_set_error("pn2: no CUDA device (there is no CPU fallback)");
        return PN2_ERR_NODEVICE;
    }
    cudaDeviceProp pr;
    CUDA_TRY(cudaGetDeviceProperties(&pr, device));
    if (sm_count) *sm_count = pr.multiProcessorCount;
    if (cc_major) *cc_major = pr.major;
    if (cc_minor) *cc_minor = pr.minor;
    if (mem_bytes) *mem_bytes = pr.totalGlobalMem;
    return PN2_OK;
}

extern "C" int pn2_create(pn2_ctx **out, int device, const pn2_params *prm) {
    if (!out) { pn2_set_error("pn2_create: out is NULL"); return PN2_ERR_ARG; }
    *out = nullptr;
    PN2_TRY(check_params(prm));
    int sm = 0, maj = 0, min = 0;
    PN2_TRY(pn2_device_info(device, &sm, &maj, &min, nullptr));
    if (maj != 10) {
        pn2_set_error("pn2_create: device %d is sm_%d%d; this library is built for sm_100a only", device, maj, min);
        return PN2_ERR_NODEVICE;
    }
    CUDA_TRY(cudaSetDevice(device));
    pn2_ctx *h = new pn2_ctx();
    h->device = device;
    h->sm_count = sm;
    h->prm = *prm;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    pn2_init_consts(h);
    if (const char *e = getenv("PN2_TREE_TOP_TARGET")) { long v = atol(e); if (v >= 1) h->tree_top_target = v; }
    PN2_TRY(h->counters.ensure(8));
    CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    *out = h;
    return PN2_OK;
}

extern "C" int pn2_destroy(pn2_ctx *h) {
    if (!h) return PN2_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->pos.release(); h->acc.release(); h->rel.release(); h->tiles.release(); h->tiles64.release(); h->gtab.release(); h->rec_pos.release(); h->rec_acc.release(); h->geom.release(); h->son.release(); h->desc.release();
    h->M.release(); h->L.release(); h->has_l.release(); h->level_nodes.release(); h->r_desc.release(); h->r_geom.release();
    h->r_M.release(); h->r_pos.release(); h->r_rel.release(); h->ia.release(); h->ib.release(); h->ic.release();
    h->id_.release(); h->la.release(); h->ua.release(); h->ub.release(); h->tmp.release(); h->counters.release();
    pn2_modeb_release(h);
    cudaStreamDestroy(h->stream);
    delete h;
    return PN2_OK;
}

extern "C" int pn2_set_params(pn2_ctx *h, const pn2_params *prm) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    PN2_TRY(check_params(prm));
    h->prm = *prm;
    pn2_init_consts(h);
    h->have_tree = false;       // FP32 relative coordinates depend on rs
    return PN2_OK;
}

extern "C" int pn2_sync(pn2_ctx *h) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PN2_OK;
}

extern "C" long pn2_launch_count(pn2_ctx *h) { return h ? h->launches : 0; }

extern "C" int pn2_timer_start(pn2_ctx *h, int slot) {
    if (!h || slot < 0 || slot > 3) { pn2_set_error("pn2_timer_start: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    for (int k = 0; k < 2; k++)
        if (!h->tev[slot][k]) CUDA_TRY(cudaEventCreate(&h->tev[slot][k]));
    CUDA_TRY(cudaEventRecord(h->tev[slot][0], h->stream));
    return PN2_OK;
}
extern "C" int pn2_timer_stop(pn2_ctx *h, int slot, double *ms) {
    if (!h || slot < 0 || slot > 3 || !ms || !h->tev[slot][0]) { pn2_set_error("pn2_timer_stop: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventRecord(h->tev[slot][1], h->stream));
    CUDA_TRY(cudaEventSynchronize(h->tev[slot][1]));
    float t = 0;
    CUDA_TRY(cudaEventElapsedTime(&t, h->tev[slot][0], h->tev[slot][1]));
    *ms = t;
    return PN2_OK;
}

// ------------------------------------------------------------------------------------------------
// Mode A uploads
// ------------------------------------------------------------------------------------------------
extern "C" int pn2_set_particles(pn2_ctx *h, const double *pos, size_t stride_bytes, int n) {
    if (!h || (!pos && n > 0) || n < 0 || stride_bytes < 24 || stride_bytes % 8) {
        pn2_set_error("pn2_set_particles: bad argument");
        return PN2_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    h->n = n;
    PN2_TRY(h->pos.ensure(3 * (size_t)n + 3));
    PN2_TRY(h->acc.ensure(3 * (size_t)n + 3));
    PN2_TRY(h->rel.ensure((size_t)n + 1));
    if (n > 0) {
        if (stride_bytes == 24) CUDA_TRY(cudaMemcpyAsync(h->pos.p, pos, 24 * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        else CUDA_TRY(cudaMemcpy2DAsync(h->pos.p, 24, pos, stride_bytes, 24, n, cudaMemcpyHostToDevice, h->stream));
    }
    CUDA_TRY(cudaMemsetAsync(h->acc.p, 0, 3 * (size_t)n * sizeof(double), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->have_particles = true;
    h->have_tree = false;
    return PN2_OK;
}

extern "C" int pn2_zero_acc(pn2_ctx *h) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->n > 0) CUDA_TRY(cudaMemsetAsync(h->acc.p, 0, 3 * (size_t)h->n * sizeof(double), h->stream));
    CUDA_TRY(cudaMemsetAsync(h->counters.p, 0, 8 * sizeof(unsigned long long), h->stream));
    return PN2_OK;
}

// Convert the reference's AoS Pack/Node arrays to the cell arrays, and group the nodes by depth
// (children have larger ids than their parent: build_kdtree numbers nodes in pre-order, src/fmm.c:101-118).
extern "C" int pn2_set_tree(pn2_ctx *h, const pn2_pack *leaf, int first_leaf, int last_leaf, const pn2_node *btree,
                            int first_node, int last_node) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    if (!h->have_particles) { pn2_set_error("pn2_set_tree: call pn2_set_particles first"); return PN2_ERR_STATE; }
    int nleaf = last_leaf - first_leaf, nnode = last_node - first_node + 1;
    if (h->n == 0) { nleaf = 0; nnode = 0; }
    if (nleaf < 0 || nnode < 0 || (nleaf > 0 && !leaf) || (nnode > 0 && !btree)) {
        pn2_set_error("pn2_set_tree: bad argument");
        return PN2_ERR_ARG;
    }
    if ((size_t)nleaf + nnode >= (1u << PN2_IMG_SHIFT)) { pn2_set_error("pn2_set_tree: more than 2^27 cells"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    h->nleaf = nleaf; h->nnode = nnode; h->ncell = nleaf + nnode;
    h->first_leaf = first_leaf; h->last_leaf = last_leaf; h->first_node = first_node; h->last_node = last_node;
    size_t nc = (size_t)h->ncell;
    std::vector<double> geom(6 * nc), M0;
    std::vector<int> son(2 * nc, -1);
    std::vector<LeafDesc> desc(nc);
    auto cell_of = [&](int id) -> int {
        if (id >= first_leaf && id < last_leaf) return id - first_leaf;
        if (id >= first_node && id <= last_node) return nleaf + (id - first_node);
        return -1;
    };
    for (int k = 0; k < nleaf; k++) {
        const pn2_pack &p = leaf[k];
        if (p.npart < 0 || p.npart > h->prm.maxleaf || p.ipart < 0 || p.ipart + p.npart > h->n) {
            pn2_set_error("pn2_set_tree: leaf %d has npart %d ipart %d (maxleaf %d, n %d)", k, p.npart, p.ipart, h->prm.maxleaf, h->n);
            return PN2_ERR_ARG;
        }
        for (int d = 0; d < 3; d++) { geom[6 * (size_t)k + d] = p.center[d]; geom[6 * (size_t)k + 3 + d] = p.width[d]; desc[k].c[d] = p.center[d]; }
        desc[k].first = p.ipart; desc[k].npart = p.npart;
    }
    std::vector<int> depth(nnode, 0);
    int maxd = 0;
    for (int k = 0; k < nnode; k++) {
        const pn2_node &nd = btree[k];
        size_t c = (size_t)nleaf + k;
        for (int d = 0; d < 3; d++) { geom[6 * c + d] = nd.center[d]; geom[6 * c + 3 + d] = nd.width[d]; desc[c].c[d] = nd.center[d]; }
        desc[c].first = 0; desc[c].npart = nd.npart;
        for (int s = 0; s < 2; s++) {
            int ch = cell_of(nd.son[s]);
            son[2 * c + s] = ch;
            if (ch >= nleaf) {
                if (ch - nleaf <= k) { pn2_set_error("pn2_set_tree: node %d has a son with a smaller id", k); return PN2_ERR_ARG; }
                depth[ch - nleaf] = depth[k] + 1;
                if (depth[k] + 1 > maxd) maxd = depth[k] + 1;
            }
        }
    }
    h->nlevel = nnode > 0 ? maxd + 1 : 0;
    h->level_off.assign(h->nlevel + 1, 0);
    for (int k = 0; k < nnode; k++) h->level_off[depth[k] + 1]++;
    for (int l = 0; l < h->nlevel; l++) h->level_off[l + 1] += h->level_off[l];
    std::vector<int> lv(nnode), fill(h->level_off.begin(), h->level_off.end());
    for (int k = 0; k < nnode; k++) lv[fill[depth[k]]++] = nleaf + k;

    PN2_TRY(h->geom.ensure(6 * nc + 6)); PN2_TRY(h->son.ensure(2 * nc + 2)); PN2_TRY(h->desc.ensure(nc + 1));
    PN2_TRY(h->M.ensure(NM * nc + NM)); PN2_TRY(h->L.ensure(NM * nc + NM)); PN2_TRY(h->level_nodes.ensure(nnode + 1));
    if (nc > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->geom.p, geom.data(), 6 * nc * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->son.p, son.data(), 2 * nc * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->desc.p, desc.data(), nc * sizeof(LeafDesc), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemsetAsync(h->M.p, 0, NM * nc * sizeof(double), h->stream));
        CUDA_TRY(cudaMemsetAsync(h->L.p, 0, NM * nc * sizeof(double), h->stream));
    }
    if (nnode > 0)
        CUDA_TRY(cudaMemcpyAsync(h->level_nodes.p, lv.data(), nnode * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    PN2_TRY(pn2_launch_relpos(h, h->pos.p, h->desc.p, nleaf, h->rel.p, h->n));
    CUDA_TRY(cudaStreamSynchronize(h->stream));     // host vectors go out of scope
    h->have_tree = true;
    h->have_remote = false;
    h->use_lflags = false;          // Mode A: L is cleared here and every cell carries one
    return PN2_OK;
}

extern "C" int pn2_set_remote(pn2_ctx *h, const pn2_remote_node *rt, int nnode, const pn2_remote_body *rb, int nbody) {
    if (!h || nnode < 0 || nbody < 0 || (nnode > 0 && !rt) || (nbody > 0 && !rb)) {
        pn2_set_error("pn2_set_remote: bad argument");
        return PN2_ERR_ARG;
    }
    if (!h->have_tree) { pn2_set_error("pn2_set_remote: call pn2_set_tree first"); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    h->r_nnode = nnode; h->r_nbody = nbody;
    std::vector<LeafDesc> desc(nnode);
    std::vector<double> geom(6 * (size_t)nnode), M(NM * (size_t)nnode), pos(3 * (size_t)nbody);
    int nrleaf = 0;
    for (int k = 0; k < nnode; k++) {
        for (int d = 0; d < 3; d++) { geom[6 * (size_t)k + d] = rt[k].center[d]; geom[6 * (size_t)k + 3 + d] = rt[k].width[d]; desc[k].c[d] = rt[k].center[d]; }
        memcpy(&M[NM * (size_t)k], rt[k].M, NM * sizeof(double));
        // a remote leaf is npart <= MAXLEAF (src/remotes.c:228); its bodies are son[0] .. son[0]+npart (:44-46)
        bool isleaf = rt[k].npart <= h->prm.maxleaf;
        desc[k].first = isleaf ? rt[k].son[0] : 0;
        desc[k].npart = isleaf ? rt[k].npart : 0;
        if (isleaf) {
            nrleaf++;
            if (rt[k].npart < 0 || rt[k].son[0] < 0 || rt[k].son[0] + rt[k].npart > nbody) {
                pn2_set_error("pn2_set_remote: remote leaf %d body range [%d,+%d) outside 0..%d", k, rt[k].son[0], rt[k].npart, nbody);
                return PN2_ERR_ARG;
            }
        }
    }
    for (int k = 0; k < nbody; k++) { pos[3 * (size_t)k] = rb[k].pos[0]; pos[3 * (size_t)k + 1] = rb[k].pos[1]; pos[3 * (size_t)k + 2] = rb[k].pos[2]; }
    PN2_TRY(h->r_desc.ensure(nnode + 1)); PN2_TRY(h->r_geom.ensure(6 * (size_t)nnode + 6)); PN2_TRY(h->r_M.ensure(NM * (size_t)nnode + NM));
    PN2_TRY(h->r_pos.ensure(3 * (size_t)nbody + 3)); PN2_TRY(h->r_rel.ensure(nbody + 1));
    if (nnode > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->r_desc.p, desc.data(), nnode * sizeof(LeafDesc), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->r_geom.p, geom.data(), 6 * (size_t)nnode * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->r_M.p, M.data(), NM * (size_t)nnode * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (nbody > 0) {
        CUDA_TRY(cudaMemcpyAsync(h->r_pos.p, pos.data(), 3 * (size_t)nbody * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        // every flattened node is a candidate "leaf" slot of the relpos kernel; non-leaves have npart 0
        PN2_TRY(pn2_launch_relpos(h, h->r_pos.p, h->r_desc.p, nnode, h->r_rel.p, nbody));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->have_remote = true;
    return PN2_OK;
}

// ------------------------------------------------------------------------------------------------
// (source id, sink id) pair batch -> CSR by sink on the device: radix sort by sink cell, run-length
// encode, exclusive scan.  The reference hands the worker thread the two int arrays task_s / task_t
// (src/fmm.c:796-812).
// ------------------------------------------------------------------------------------------------
__global__ void map_ids_kernel(long n, const int *__restrict__ s, const int *__restrict__ t, int s_leaf0, int s_nleaf,
                               int s_node0, int s_nnode, int s_nleaf_cells, int t_leaf0, int t_nleaf, int t_node0,
                               int t_nnode, unsigned *__restrict__ so, int *__restrict__ to, int *__restrict__ bad) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = s[i], b = t[i];
    int sc = -1, tc = -1;
    if (a >= s_leaf0 && a < s_leaf0 + s_nleaf) sc = a - s_leaf0;
    else if (a >= s_node0 && a < s_node0 + s_nnode) sc = s_nleaf_cells + (a - s_node0);
    if (b >= t_leaf0 && b < t_leaf0 + t_nleaf) tc = b - t_leaf0;
    else if (b >= t_node0 && b < t_node0 + t_nnode) tc = t_nleaf + (b - t_node0);
    if (sc < 0 || tc < 0) { atomicAdd(bad, 1); sc = 0; tc = 0; }
    so[i] = (unsigned)sc;
    to[i] = tc;
}
// o[0..n) = (long)in[0..n), o[n] = 0 (the slot the exclusive scan turns into the total)
__global__ void int_to_long_kernel(int n, const int *__restrict__ in, long *__restrict__ o) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = in[i];
    else if (i == n) o[i] = 0;
}

int pn2_build_csr(pn2_ctx *h, const int *h_s, const int *h_t, long n, int remote_src, int sinks_may_be_nodes, CsrList *out) {
    out->nseg = 0;
    if (n == 0) return PN2_OK;
    if (n < 0 || !h_s || !h_t) { pn2_set_error("pn2: bad batch"); return PN2_ERR_ARG; }
    if (n >= (1L << 31)) { pn2_set_error("pn2: batch of %ld pairs; split it (the reference uses 16384)", n); return PN2_ERR_ARG; }
    PN2_TRY(h->ia.ensure(n)); PN2_TRY(h->ib.ensure(n + 1)); PN2_TRY(h->id_.ensure(n + 2)); PN2_TRY(h->ua.ensure(n));
    cudaStream_t st = h->stream;
    CUDA_TRY(cudaMemcpyAsync(h->ia.p, h_s, n * sizeof(int), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(h->ib.p, h_t, n * sizeof(int), cudaMemcpyHostToDevice, st));
    int *bad = h->id_.p + n + 1;
    CUDA_TRY(cudaMemsetAsync(bad, 0, sizeof(int), st));
    unsigned grid = (unsigned)((n + 255) / 256);
    int *tcell = h->id_.p;     // unsorted sink cells
    if (remote_src)
        map_ids_kernel<<<grid, 256, 0, st>>>(n, h->ia.p, h->ib.p, 0, h->r_nnode, 0, 0, 0, h->first_leaf, h->nleaf, h->first_node,
                                             sinks_may_be_nodes ? h->nnode : 0, h->ua.p, tcell, bad);
    else
        map_ids_kernel<<<grid, 256, 0, st>>>(n, h->ia.p, h->ib.p, h->first_leaf, h->nleaf, h->first_node,
                                             sinks_may_be_nodes ? h->nnode : 0, h->nleaf, h->first_leaf, h->nleaf,
                                             h->first_node, sinks_may_be_nodes ? h->nnode : 0, h->ua.p, tcell, bad);
    h->launches++;
    KERNEL_CHECK();
    CUDA_TRY(cudaStreamSynchronize(st));
    int hb = 0;
    CUDA_TRY(cudaMemcpy(&hb, bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (hb != 0) { pn2_set_error("pn2: %d ids of the batch are outside the tree", hb); return PN2_ERR_ARG; }
    return pn2_csr_from_device_pairs(h, tcell, h->ua.p, n, out);
}

// device pairs (sink cell, source entry) -> CSR by sink.  Stable radix sort by sink cell (sources keep
// their list order within a sink), run-length encode, exclusive scan.  tcell must not alias ia/ib/ic/la/ub.
// out->seg_sink = h->ic, out->seg_off = h->la, out->src = h->ub (valid until the next call).
int pn2_csr_from_device_pairs(pn2_ctx *h, int *tcell, unsigned *scell, long n, CsrList *out) {
    out->nseg = 0;
    if (n == 0) return PN2_OK;
    if (n < 0 || n >= (1L << 31) - 1) { pn2_set_error("pn2: %ld list pairs in one CSR build (limit 2^31 - 2)", n); return PN2_ERR_ARG; }
    cudaStream_t st = h->stream;
    PN2_TRY(h->ia.ensure(n)); PN2_TRY(h->ib.ensure(n + 1)); PN2_TRY(h->ic.ensure(n + 1));
    PN2_TRY(h->ub.ensure(n)); PN2_TRY(h->la.ensure(n + 2)); PN2_TRY(h->b_scal.ensure(16));
    size_t tb = 0, tb2 = 0, tb3 = 0;
    int bits = 1;
    while ((1L << bits) < (long)h->ncell + 1 && bits < 31) bits++;
    int *tsorted = h->ia.p;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, tcell, tsorted, scell, h->ub.p, (int)n, 0, bits, st);
    int *runs = h->ib.p;
    int *nruns = h->b_scal.p + 8;
    cub::DeviceRunLengthEncode::Encode(nullptr, tb2, tsorted, h->ic.p, runs, nruns, (int)n, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, (long *)nullptr, (long *)nullptr, (int)n + 1, st);
    size_t need = tb > tb2 ? tb : tb2;
    if (tb3 > need) need = tb3;
    PN2_TRY(h->tmp.ensure(need + 16));
    cub::DeviceRadixSort::SortPairs(h->tmp.p, tb, tcell, tsorted, scell, h->ub.p, (int)n, 0, bits, st);
    cub::DeviceRunLengthEncode::Encode(h->tmp.p, tb2, tsorted, h->ic.p, runs, nruns, (int)n, st);
    h->launches += 4;
    int nseg = 0;
    CUDA_TRY(cudaMemcpyAsync(&nseg, nruns, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int_to_long_kernel<<<(unsigned)((nseg + 256) / 256), 256, 0, st>>>(nseg, runs, h->la.p);
    cub::DeviceScan::ExclusiveSum(h->tmp.p, tb3, h->la.p, h->la.p, nseg + 1, st);
    h->launches += 2;
    KERNEL_CHECK();
    out->nseg = nseg;
    out->npair = n;
    out->seg_sink = h->ic.p;
    out->seg_off = h->la.p;
    out->src = h->ub.p;
    return PN2_OK;
}

static int need_tree(pn2_ctx *h, const char *who) {
    if (!h) { pn2_set_error("pn2: null context"); return PN2_ERR_ARG; }
    if (!h->have_tree) { pn2_set_error("%s: call pn2_set_particles and pn2_set_tree first", who); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    return PN2_OK;
}

extern "C" int pn2_p2m_m2m(pn2_ctx *h) {
    PN2_TRY(need_tree(h, "pn2_p2m_m2m"));
    PN2_TRY(pn2_launch_p2m(h));
    return pn2_launch_m2m(h);
}

extern "C" int pn2_l2l_l2p(pn2_ctx *h) {
    PN2_TRY(need_tree(h, "pn2_l2l_l2p"));
    return pn2_launch_l2l_l2p(h);
}

extern "C" int pn2_p2p_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n) {
    PN2_TRY(need_tree(h, "pn2_p2p_batch"));
    CsrList list;
    PN2_TRY(pn2_build_csr(h, task_s, task_t, n, 0, 0, &list));
    SourceSet src{h->desc.p, h->rel.p, h->pos.p};
    return pn2_launch_p2p(h, list, src, true);
}

extern "C" int pn2_m2l_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n) {
    PN2_TRY(need_tree(h, "pn2_m2l_batch"));
    CsrList list;
    PN2_TRY(pn2_build_csr(h, task_s, task_t, n, 0, 1, &list));
    return pn2_launch_m2l(h, list, h->geom.p, h->M.p);
}

extern "C" int pn2_p2p_ext_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n) {
    PN2_TRY(need_tree(h, "pn2_p2p_ext_batch"));
    if (!h->have_remote) { pn2_set_error("pn2_p2p_ext_batch: call pn2_set_remote first"); return PN2_ERR_STATE; }
    CsrList list;
    PN2_TRY(pn2_build_csr(h, task_s, task_t, n, 1, 0, &list));
    SourceSet src{h->r_desc.p, h->r_rel.p, h->r_pos.p};
    return pn2_launch_p2p(h, list, src, false);
}

extern "C" int pn2_m2l_ext_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n) {
    PN2_TRY(need_tree(h, "pn2_m2l_ext_batch"));
    if (!h->have_remote) { pn2_set_error("pn2_m2l_ext_batch: call pn2_set_remote first"); return PN2_ERR_STATE; }
    CsrList list;
    PN2_TRY(pn2_build_csr(h, task_s, task_t, n, 1, 1, &list));
    return pn2_launch_m2l(h, list, h->r_geom.p, h->r_M.p);
}

// ------------------------------------------------------------------------------------------------
// results
// ------------------------------------------------------------------------------------------------
extern "C" int pn2_get_acc(pn2_ctx *h, double *acc, size_t stride_bytes, int n, int accumulate) {
    if (!h || (!acc && n > 0) || n != h->n || stride_bytes < 24 || stride_bytes % 8) {
        pn2_set_error("pn2_get_acc: bad argument (n %d, context has %d)", n, h ? h->n : -1);
        return PN2_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    if (n == 0) return pn2_sync(h);
    if (!accumulate) {
        CUDA_TRY(cudaMemcpy2DAsync(acc, stride_bytes, h->acc.p, 24, 24, n, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        return PN2_OK;
    }
    std::vector<double> tmp(3 * (size_t)n);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->acc.p, 3 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n; i++) {
        double *a = (double *)((char *)acc + (size_t)i * stride_bytes);
        a[0] += tmp[3 * (size_t)i]; a[1] += tmp[3 * (size_t)i + 1]; a[2] += tmp[3 * (size_t)i + 2];
    }
    return PN2_OK;
}

static int get_ml(pn2_ctx *h, pn2_pack *leaf, pn2_node *btree, bool wantM) {
    PN2_TRY(need_tree(h, "pn2_get_multipoles/locals"));
    size_t nc = (size_t)h->ncell;
    std::vector<double> buf(NM * nc);
    if (nc == 0) return PN2_OK;
    CUDA_TRY(cudaMemcpyAsync(buf.data(), wantM ? h->M.p : h->L.p, NM * nc * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (leaf)
        for (int k = 0; k < h->nleaf; k++) memcpy(wantM ? leaf[k].M : leaf[k].L, &buf[NM * (size_t)k], NM * sizeof(double));
    if (btree)
        for (int k = 0; k < h->nnode; k++)
            memcpy(wantM ? btree[k].M : btree[k].L, &buf[NM * ((size_t)h->nleaf + k)], NM * sizeof(double));
    return PN2_OK;
}
extern "C" int pn2_get_multipoles(pn2_ctx *h, pn2_pack *leaf, pn2_node *btree) { return get_ml(h, leaf, btree, true); }
extern "C" int pn2_get_locals(pn2_ctx *h, pn2_pack *leaf, pn2_node *btree) { return get_ml(h, leaf, btree, false); }

extern "C" int pn2_get_counters(pn2_ctx *h, double counters[8]) {
    if (!h || !counters) { pn2_set_error("pn2_get_counters: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long c[8];
    CUDA_TRY(cudaMemcpyAsync(c, h->counters.p, sizeof c, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 8; i++) counters[i] = (double)c[i];
    return PN2_OK;
}

// ------------------------------------------------------------------------------------------------
// FMA-pipe issue-rate microbenchmark: 8 independent FFMA (DFMA) chains per thread, register operands
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T *out, int iters, T a, T b) {
    T x0 = (T)threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
            x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
        }
    }
    T s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == (T)123456789) out[0] = s;
}

extern "C" int pn2_fma_peak(pn2_ctx *h, int fp64, double *ops_per_s, double *ms_out) {
    if (!h || !ops_per_s) { pn2_set_error("pn2_fma_peak: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    PN2_TRY(h->tmp.ensure(64));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    int blocks = h->sm_count * 8, iters = fp64 ? 2000 : 8000;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0, h->stream));
        if (fp64) fma_peak_kernel<double><<<blocks, 256, 0, h->stream>>>((double *)h->tmp.p, iters, 0.999999, 1e-7);
        else fma_peak_kernel<float><<<blocks, 256, 0, h->stream>>>((float *)h->tmp.p, iters, 0.999999f, 1e-7f);
        CUDA_TRY(cudaEventRecord(e1, h->stream));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
        h->launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    double ops = (double)blocks * 256.0 * (double)iters * 16.0 * 8.0;
    *ops_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return PN2_OK;
}
