// pn2_let.cu -- Mode B multi-rank: locally-essential-tree (LET) prune, pack, exchange, unpack.
//
// Replaces fmm_remote / prepare_sendtree2 and the (P-1) + 26 P serialised pack -> 4 blocking MPI exchanges
// -> walk rounds of the reference (src/remotes.c:60-169, 684-751; src/fmm.c:1015-1053).
//
// The reference prunes the sender's tree against the target's domain box once per (peer, displacement)
// and ships 27 P - 1 separately flattened trees.  Here a rank packs ONE subtree per peer: the union over
// the 27 displacements of the cells prepare_sendtree2 would send (a cell is sent for displacement s iff no
// ancestor is terminal for s: dr >= cutoff or width < 0.95 theta dr, src/remotes.c:145-158).  The receiver
// re-evaluates the same terminal test on the fly (pruned_dev in pn2_walk.cu, same arithmetic, same
// inputs), so the lists are the ones the reference builds from its 27 copies; what crosses NVLink is each
// needed cell once (224 bytes, the size of the reference's RemoteNode) and each ghost particle once
// (16 bytes in FP32 mode, 24 in FP64 mode; the reference's RemoteBody is 32).
//
// Transport: grouped ncclSend / ncclRecv over the communicator given to pn2_set_comm (NVLink / NVSwitch),
// or device-to-device copies between contexts of one process (pn2_exchange_local; also what the
// single-GPU tests use to drive two ranks).
#include <cub/cub.cuh>
#include <dlfcn.h>
#include "pn2_nccl.cuh"

Pn2NcclApi g_nccl;
bool nccl_load() {
    if (g_nccl.ok) return true;
    void *hd = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    const char *env = getenv("PN2_NCCL_LIB");
    if (!hd && env && *env) hd = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!hd) hd = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!hd) { pn2_set_error("pn2: cannot load libnccl.so.2: %s", dlerror()); return false; }
#define BIND(field, sym) *(void **)(&g_nccl.field) = dlsym(hd, sym); if (!g_nccl.field) { pn2_set_error("pn2: libnccl lacks %s", sym); return false; }
    BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank") BIND(CommDestroy, "ncclCommDestroy")
    BIND(Send, "ncclSend") BIND(Recv, "ncclRecv") BIND(GroupStart, "ncclGroupStart") BIND(GroupEnd, "ncclGroupEnd")
    BIND(GetErrorString, "ncclGetErrorString") BIND(AllReduce, "ncclAllReduce")
#undef BIND
    g_nccl.ok = true;
    return true;
}

struct __align__(16) PackCell {
    double geom[6];
    int a, b;              // leaf: first ghost particle (within the peer block), unused; node: packed sons
    int npart, pad;
    double M[NM];
};
static_assert(sizeof(PackCell) == 224, "PackCell is 224 bytes");

struct LetState {
    int npeer = 0;
    std::vector<int> peers;                 // peer ranks, ascending, self excluded
    DBuf<unsigned> reach;                   // [npeer][ncell] displacement bit-mask of "is sent"
    DBuf<int> lidx, nidx, pidx;             // [npeer][nleaf + 1], [npeer][nnode + 1], [npeer][nleaf + 1] exclusive scans
    DBuf<double> tbox;                      // [npeer][6] target centre / width
    DBuf<PackCell> send_leaf, send_node, recv_leaf, recv_node;
    DBuf<unsigned char> send_part, recv_part;
    DBuf<int> cnt_dev;                      // [2][npeer][4]
    std::vector<long> s_nl, s_nn, s_np, r_nl, r_nn, r_np;   // per peer counts
    int psize = 16;
    // pack and exchange run on their own stream so that the walk over the local tree (pn2_modeb.cu, pass 0) hides them
    cudaStream_t st = nullptr;
    cudaEvent_t ev_tree = nullptr, ev_done = nullptr;       // main stream: tree + multipoles ready; LET stream: blocks received
    bool done_recorded = false;
    int rank_c = -1, nranks_c = -1;
};

// same arithmetic as pruned_dev (pn2_walk.cu) / prepare_sendtree2 (src/remotes.c:97-158); this file is
// compiled with -fmad=false as well
__device__ __forceinline__ int let_terminal(const double *c, const double *w, const double *sh, const double *tc,
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

// one thread per (node of this level, peer): hand the surviving displacement mask to both sons
__global__ void reach_level_kernel(int cnt, const int *__restrict__ nodes, int npeer, int ncell, const double *__restrict__ geom,
                                   const int *__restrict__ son, const double *__restrict__ tbox, unsigned *__restrict__ reach,
                                   P2PConst pc, double cutoff, double theta, int longshort) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cnt * npeer) return;
    int k = t % cnt, p = t / cnt;
    int c = nodes[k];
    unsigned m = reach[(size_t)p * ncell + c];
    if (!m) return;
    double cc[3] = {geom[6 * (size_t)c], geom[6 * (size_t)c + 1], geom[6 * (size_t)c + 2]};
    double ww[3] = {geom[6 * (size_t)c + 3], geom[6 * (size_t)c + 4], geom[6 * (size_t)c + 5]};
    const double *tc = tbox + 6 * p, *tw = tc + 3;
    unsigned keep = 0;
    for (int s = 0; s < 27; s++)
        if ((m >> s) & 1u)
            if (!let_terminal(cc, ww, pc.shift[s], tc, tw, cutoff, theta, longshort)) keep |= 1u << s;
    reach[(size_t)p * ncell + son[2 * (size_t)c]] = keep;
    reach[(size_t)p * ncell + son[2 * (size_t)c + 1]] = keep;
}

__global__ void let_flags_kernel(int npeer, int nleaf, int nnode, int ncell, const unsigned *__restrict__ reach,
                                 const LeafDesc *__restrict__ desc, int *__restrict__ lidx, int *__restrict__ nidx,
                                 int *__restrict__ pidx) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long per = (long)ncell + 2;
    if (t >= per * npeer) return;
    int p = (int)(t / per);
    long c = t % per;
    if (c < nleaf) {
        bool need = reach[(size_t)p * ncell + c] != 0;
        lidx[(size_t)p * (nleaf + 1) + c] = need;
        pidx[(size_t)p * (nleaf + 1) + c] = need ? desc[c].npart : 0;
    } else if (c < ncell) {
        nidx[(size_t)p * (nnode + 1) + (c - nleaf)] = reach[(size_t)p * ncell + c] != 0;
    } else if (c == ncell) {
        lidx[(size_t)p * (nleaf + 1) + nleaf] = 0; pidx[(size_t)p * (nleaf + 1) + nleaf] = 0;
    } else {
        nidx[(size_t)p * (nnode + 1) + nnode] = 0;
    }
}

__global__ void let_pack_cells_kernel(int p, int nleaf, int nnode, int ncell, const unsigned *__restrict__ reach,
                                      const int *__restrict__ lidx, const int *__restrict__ nidx, const int *__restrict__ pidx,
                                      const double *__restrict__ geom, const int *__restrict__ son, const LeafDesc *__restrict__ desc,
                                      const double *__restrict__ M, PackCell *__restrict__ out_leaf, PackCell *__restrict__ out_node) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const unsigned *rp = reach + (size_t)p * ncell;
    if (!rp[c]) return;
    const int *li = lidx + (size_t)p * (nleaf + 1), *ni = nidx + (size_t)p * (nnode + 1), *pi = pidx + (size_t)p * (nleaf + 1);
    PackCell pcell;
#pragma unroll
    for (int d = 0; d < 6; d++) pcell.geom[d] = geom[6 * (size_t)c + d];
#pragma unroll
    for (int d = 0; d < NM; d++) pcell.M[d] = M[(size_t)c * NM + d];
    pcell.npart = desc[c].npart; pcell.pad = 0;
    if (c < nleaf) {
        pcell.a = pi[c]; pcell.b = 0;
        out_leaf[li[c]] = pcell;
    } else {
        int ref[2];
        for (int s = 0; s < 2; s++) {
            int ch = son[2 * (size_t)c + s];
            if (ch < 0 || !rp[ch]) ref[s] = -1;
            else if (ch < nleaf) ref[s] = -(li[ch] + 2);
            else ref[s] = ni[ch - nleaf];
        }
        pcell.a = ref[0]; pcell.b = ref[1];
        out_node[ni[c - nleaf]] = pcell;
    }
}

__global__ void let_pack_parts_kernel(int p, int nleaf, int ncell, int maxleaf, const unsigned *__restrict__ reach,
                                      const int *__restrict__ pidx, const LeafDesc *__restrict__ desc, double inv_len,
                                      const double *__restrict__ pos, int fp64, unsigned char *__restrict__ out) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int k = (int)(t / maxleaf), j = (int)(t % maxleaf);
    if (k >= nleaf) return;
    if (!reach[(size_t)p * ncell + k]) return;
    LeafDesc d = desc[k];
    if (j >= d.npart) return;
    size_t o = (size_t)pidx[(size_t)p * (nleaf + 1) + k] + j;
    if (fp64) {
        double *q = reinterpret_cast<double *>(out) + 3 * o;
        const double *s = pos + 3 * (size_t)(d.first + j);
        q[0] = s[0]; q[1] = s[1]; q[2] = s[2];
    } else {
        // FP32 mode ships what the receiver's tiles hold: leaf-centre-relative coordinates in units of lambda (pn2_walk.cu, tile_kernel)
        const double *s = pos + 3 * (size_t)(d.first + j);
        reinterpret_cast<float4 *>(out)[o] = make_float4((float)((s[0] - d.c[0]) * inv_len), (float)((s[1] - d.c[1]) * inv_len),
                                                         (float)((s[2] - d.c[2]) * inv_len), 1.f);
    }
}

// received block of one peer -> the unified cell / particle arrays
__global__ void let_unpack_kernel(int nl, int nn, const PackCell *__restrict__ in_leaf, const PackCell *__restrict__ in_node,
                                  int leaf_id0, int node_id0, int part0, double *__restrict__ geom, int *__restrict__ son,
                                  LeafDesc *__restrict__ desc, double *__restrict__ M) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nl + nn) return;
    const bool isleaf = t < nl;
    const PackCell pcell = isleaf ? in_leaf[t] : in_node[t - nl];
    const int id = isleaf ? leaf_id0 + t : node_id0 + (t - nl);
    LeafDesc d;
#pragma unroll
    for (int q = 0; q < 6; q++) geom[6 * (size_t)id + q] = pcell.geom[q];
#pragma unroll
    for (int q = 0; q < NM; q++) M[(size_t)id * NM + q] = pcell.M[q];
    d.c[0] = pcell.geom[0]; d.c[1] = pcell.geom[1]; d.c[2] = pcell.geom[2];
    d.npart = pcell.npart;
    if (isleaf) {
        d.first = part0 + pcell.a;
        son[2 * (size_t)id] = -1; son[2 * (size_t)id + 1] = -1;
    } else {
        d.first = 0;
        int r[2] = {pcell.a, pcell.b};
        for (int s = 0; s < 2; s++) son[2 * (size_t)id + s] = r[s] == -1 ? -1 : (r[s] >= 0 ? node_id0 + r[s] : leaf_id0 + (-(r[s] + 2)));
    }
    desc[id] = d;
}

void pn2_let_release(pn2_ctx *h) {
    if (!h->let) return;
    LetState *L = h->let;
    L->reach.release(); L->lidx.release(); L->nidx.release(); L->pidx.release(); L->tbox.release();
    L->send_leaf.release(); L->send_node.release(); L->recv_leaf.release(); L->recv_node.release();
    L->send_part.release(); L->recv_part.release(); L->cnt_dev.release();
    if (L->st) cudaStreamDestroy(L->st);
    if (L->ev_tree) cudaEventDestroy(L->ev_tree);
    if (L->ev_done) cudaEventDestroy(L->ev_done);
    delete L;
    h->let = nullptr;
}

static LetState *let_state(pn2_ctx *h) {
    if (!h->let) h->let = new LetState();
    LetState *L = h->let;
    if (!L->st) {
        // highest priority: its few CTAs (pack kernels, NCCL send / recv) are dispatched ahead of the walk pass that
        // saturates the GPU on the main stream -- without it they would wait for that grid to be fully dispatched
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&L->st, cudaStreamNonBlocking, hi);
        cudaEventCreateWithFlags(&L->ev_tree, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&L->ev_done, cudaEventDisableTiming);
    }
    if (L->rank_c != h->rank || L->nranks_c != h->nranks) {
        L->rank_c = h->rank; L->nranks_c = h->nranks;
        L->peers.clear();
        for (int r = 0; r < h->nranks; r++) if (r != h->rank) L->peers.push_back(r);
        L->npeer = (int)L->peers.size();
    }
    return L;
}

// called on the main stream right after the upward pass: the point the LET stream has to wait for
int pn2_let_tree_ready(pn2_ctx *h) {
    LetState *L = let_state(h);
    CUDA_TRY(cudaEventRecord(L->ev_tree, h->stream));
    return PN2_OK;
}

// Sender side: mark, scan, pack for every peer.  Needs the tree and the multipoles (after P2M / M2M).
int pn2_let_pack_all(pn2_ctx *h) {
    LetState *L = let_state(h);
    cudaStream_t st = L->st;
    L->done_recorded = false;
    CUDA_TRY(cudaStreamWaitEvent(st, L->ev_tree, 0));                  // tree + multipoles (pn2_let_tree_ready), not the walk enqueued behind them
    const int np = L->npeer, nleaf = h->nleaf, nnode = h->nnode, ncell = h->ncell;
    L->psize = h->prm.precision != PN2_FP32 ? 24 : 16;
    L->s_nl.assign(np, 0); L->s_nn.assign(np, 0); L->s_np.assign(np, 0);
    if (np == 0) return PN2_OK;
    std::vector<double> tb(6 * (size_t)np);
    for (int p = 0; p < np; p++) {
        const pn2_domain &d = h->all_dom[L->peers[p]];
        for (int k = 0; k < 3; k++) { tb[6 * p + k] = 0.5 * (d.hi[k] + d.lo[k]); tb[6 * p + 3 + k] = d.hi[k] - d.lo[k]; }
    }
    PN2_TRY(L->tbox.ensure(6 * (size_t)np));
    CUDA_TRY(cudaMemcpyAsync(L->tbox.p, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    if (ncell == 0) { CUDA_TRY(cudaStreamSynchronize(st)); return PN2_OK; }
    PN2_TRY(L->reach.ensure((size_t)np * ncell));
    PN2_TRY(L->lidx.ensure((size_t)np * (nleaf + 1))); PN2_TRY(L->pidx.ensure((size_t)np * (nleaf + 1)));
    PN2_TRY(L->nidx.ensure((size_t)np * (nnode + 1)));
    CUDA_TRY(cudaMemsetAsync(L->reach.p, 0, (size_t)np * ncell * sizeof(unsigned), st));
    const unsigned full = h->prm.periodic ? 0x7ffffffu : 1u;
    std::vector<unsigned> fullv(np, full);
    CUDA_TRY(cudaMemcpy2DAsync(L->reach.p + h->nleaf, (size_t)ncell * sizeof(unsigned), fullv.data(), sizeof(unsigned),
                               sizeof(unsigned), np, cudaMemcpyHostToDevice, st));     // reach[p][root] = all displacements
    for (int lev = 0; lev < h->nlevel; lev++) {
        int cnt = h->level_off[lev + 1] - h->level_off[lev];
        if (cnt == 0) continue;
        long nt = (long)cnt * np;
        reach_level_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(cnt, h->level_nodes.p + h->level_off[lev], np, ncell, h->geom.p,
                                                                          h->son.p, L->tbox.p, L->reach.p, h->pc, h->prm.cutoff,
                                                                          h->prm.theta, h->prm.longshort);
        h->launches++;
    }
    long nt = ((long)ncell + 2) * np;
    let_flags_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(np, nleaf, nnode, ncell, L->reach.p, h->desc.p, L->lidx.p, L->nidx.p, L->pidx.p);
    h->launches++;
    size_t tb1 = 0, tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb1, L->lidx.p, L->lidx.p, nleaf + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, L->nidx.p, L->nidx.p, nnode + 1, st);
    PN2_TRY(h->tmp.ensure((tb1 > tb2 ? tb1 : tb2) + 16));
    for (int p = 0; p < np; p++) {
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tb1, L->lidx.p + (size_t)p * (nleaf + 1), L->lidx.p + (size_t)p * (nleaf + 1), nleaf + 1, st);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tb1, L->pidx.p + (size_t)p * (nleaf + 1), L->pidx.p + (size_t)p * (nleaf + 1), nleaf + 1, st);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tb2, L->nidx.p + (size_t)p * (nnode + 1), L->nidx.p + (size_t)p * (nnode + 1), nnode + 1, st);
        h->launches += 3;
    }
    std::vector<int> tot(3 * (size_t)np);
    for (int p = 0; p < np; p++) {
        CUDA_TRY(cudaMemcpyAsync(&tot[3 * p], L->lidx.p + (size_t)p * (nleaf + 1) + nleaf, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&tot[3 * p + 1], L->nidx.p + (size_t)p * (nnode + 1) + nnode, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(&tot[3 * p + 2], L->pidx.p + (size_t)p * (nleaf + 1) + nleaf, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    long sl = 0, sn = 0, sp = 0;
    for (int p = 0; p < np; p++) { L->s_nl[p] = tot[3 * p]; L->s_nn[p] = tot[3 * p + 1]; L->s_np[p] = tot[3 * p + 2]; sl += L->s_nl[p]; sn += L->s_nn[p]; sp += L->s_np[p]; }
    PN2_TRY(L->send_leaf.ensure(sl + 1)); PN2_TRY(L->send_node.ensure(sn + 1)); PN2_TRY(L->send_part.ensure((size_t)(sp + 1) * L->psize));
    long ol = 0, on = 0, op = 0;
    for (int p = 0; p < np; p++) {
        let_pack_cells_kernel<<<(ncell + 127) / 128, 128, 0, st>>>(p, nleaf, nnode, ncell, L->reach.p, L->lidx.p, L->nidx.p, L->pidx.p, h->geom.p,
                                                                   h->son.p, h->desc.p, h->M.p, L->send_leaf.p + ol, L->send_node.p + on);
        long ntp = (long)nleaf * h->prm.maxleaf;
        let_pack_parts_kernel<<<(unsigned)((ntp + 255) / 256), 256, 0, st>>>(p, nleaf, ncell, h->prm.maxleaf, L->reach.p, L->pidx.p, h->desc.p,
                                                                            h->pc.inv_len, h->pos.p, h->prm.precision != PN2_FP32,
                                                                            L->send_part.p + (size_t)op * L->psize);
        h->launches += 2;
        ol += L->s_nl[p]; on += L->s_nn[p]; op += L->s_np[p];
    }
    KERNEL_CHECK();
    return PN2_OK;
}

static int ensure_recv(LetState *L) {
    long rl = 0, rn = 0, rp = 0;
    for (int p = 0; p < L->npeer; p++) { rl += L->r_nl[p]; rn += L->r_nn[p]; rp += L->r_np[p]; }
    PN2_TRY(L->recv_leaf.ensure(rl + 1)); PN2_TRY(L->recv_node.ensure(rn + 1)); PN2_TRY(L->recv_part.ensure((size_t)(rp + 1) * L->psize));
    return PN2_OK;
}

// counts, then payload, with grouped ncclSend / ncclRecv (an all-to-all-v over NVSwitch)
int pn2_let_exchange_nccl(pn2_ctx *h) {
    LetState *L = let_state(h);
    const int np = L->npeer;
    L->r_nl.assign(np, 0); L->r_nn.assign(np, 0); L->r_np.assign(np, 0);
    if (np == 0) { CUDA_TRY(cudaEventRecord(L->ev_done, L->st)); L->done_recorded = true; return PN2_OK; }
    if (!h->nccl) { pn2_set_error("pn2: no NCCL communicator (pn2_set_comm / pn2_comm_init_rank)"); return PN2_ERR_STATE; }
    if (!nccl_load()) return PN2_ERR_NCCL;
    ncclComm_t comm = (ncclComm_t)h->nccl;
    cudaStream_t st = L->st;
    PN2_TRY(L->cnt_dev.ensure(8 * (size_t)np));
    std::vector<int> sc(4 * (size_t)np), rc(4 * (size_t)np);
    for (int p = 0; p < np; p++) { sc[4 * p] = (int)L->s_nl[p]; sc[4 * p + 1] = (int)L->s_nn[p]; sc[4 * p + 2] = (int)L->s_np[p]; sc[4 * p + 3] = L->psize; }
    CUDA_TRY(cudaMemcpyAsync(L->cnt_dev.p, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(ncclGroupStart());
    for (int p = 0; p < np; p++) {
        NCCL_TRY(ncclSend(L->cnt_dev.p + 4 * p, 4, ncclInt, L->peers[p], comm, st));
        NCCL_TRY(ncclRecv(L->cnt_dev.p + 4 * np + 4 * p, 4, ncclInt, L->peers[p], comm, st));
    }
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaMemcpyAsync(rc.data(), L->cnt_dev.p + 4 * np, rc.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int p = 0; p < np; p++) {
        L->r_nl[p] = rc[4 * p]; L->r_nn[p] = rc[4 * p + 1]; L->r_np[p] = rc[4 * p + 2];
        if (rc[4 * p + 3] != L->psize) { pn2_set_error("pn2: peer %d runs another precision mode", L->peers[p]); return PN2_ERR_ARG; }
    }
    PN2_TRY(ensure_recv(L));
    long sl = 0, sn = 0, sp = 0, rl = 0, rn = 0, rp = 0;
    NCCL_TRY(ncclGroupStart());
    for (int p = 0; p < np; p++) {
        int peer = L->peers[p];
        if (L->s_nl[p]) NCCL_TRY(ncclSend(L->send_leaf.p + sl, (size_t)L->s_nl[p] * sizeof(PackCell), ncclChar, peer, comm, st));
        if (L->s_nn[p]) NCCL_TRY(ncclSend(L->send_node.p + sn, (size_t)L->s_nn[p] * sizeof(PackCell), ncclChar, peer, comm, st));
        if (L->s_np[p]) NCCL_TRY(ncclSend(L->send_part.p + (size_t)sp * L->psize, (size_t)L->s_np[p] * L->psize, ncclChar, peer, comm, st));
        if (L->r_nl[p]) NCCL_TRY(ncclRecv(L->recv_leaf.p + rl, (size_t)L->r_nl[p] * sizeof(PackCell), ncclChar, peer, comm, st));
        if (L->r_nn[p]) NCCL_TRY(ncclRecv(L->recv_node.p + rn, (size_t)L->r_nn[p] * sizeof(PackCell), ncclChar, peer, comm, st));
        if (L->r_np[p]) NCCL_TRY(ncclRecv(L->recv_part.p + (size_t)rp * L->psize, (size_t)L->r_np[p] * L->psize, ncclChar, peer, comm, st));
        sl += L->s_nl[p]; sn += L->s_nn[p]; sp += L->s_np[p]; rl += L->r_nl[p]; rn += L->r_nn[p]; rp += L->r_np[p];
    }
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaEventRecord(L->ev_done, st));
    L->done_recorded = true;
    return PN2_OK;
}

// all ranks live in this process: device-to-device copies instead of NCCL
extern "C" int pn2_exchange_local(pn2_ctx **hs, int nranks) {
    if (!hs || nranks < 1) { pn2_set_error("pn2_exchange_local: bad argument"); return PN2_ERR_ARG; }
    for (int r = 0; r < nranks; r++) {
        if (!hs[r] || hs[r]->nranks != nranks || hs[r]->rank != r) { pn2_set_error("pn2_exchange_local: context %d is not rank %d of %d", r, r, nranks); return PN2_ERR_ARG; }
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaStreamSynchronize(let_state(hs[r])->st));          // the packs (the walk's pass 0 may still run on the main streams)
    }
    for (int r = 0; r < nranks; r++) {
        pn2_ctx *h = hs[r];
        LetState *L = let_state(h);
        const int np = L->npeer;
        L->r_nl.assign(np, 0); L->r_nn.assign(np, 0); L->r_np.assign(np, 0);
        for (int p = 0; p < np; p++) {
            LetState *S = let_state(hs[L->peers[p]]);
            int me = -1;
            for (int k = 0; k < S->npeer; k++) if (S->peers[k] == r) me = k;
            L->r_nl[p] = S->s_nl[me]; L->r_nn[p] = S->s_nn[me]; L->r_np[p] = S->s_np[me];
            if (S->psize != L->psize) { pn2_set_error("pn2_exchange_local: mixed precision modes"); return PN2_ERR_ARG; }
        }
        CUDA_TRY(cudaSetDevice(h->device));
        PN2_TRY(ensure_recv(L));
        long rl = 0, rn = 0, rp = 0;
        for (int p = 0; p < np; p++) {
            LetState *S = let_state(hs[L->peers[p]]);
            long sl = 0, sn = 0, sp = 0;
            for (int k = 0; k < S->npeer && S->peers[k] != r; k++) { sl += S->s_nl[k]; sn += S->s_nn[k]; sp += S->s_np[k]; }
            if (L->r_nl[p]) CUDA_TRY(cudaMemcpyAsync(L->recv_leaf.p + rl, S->send_leaf.p + sl, (size_t)L->r_nl[p] * sizeof(PackCell), cudaMemcpyDefault, L->st));
            if (L->r_nn[p]) CUDA_TRY(cudaMemcpyAsync(L->recv_node.p + rn, S->send_node.p + sn, (size_t)L->r_nn[p] * sizeof(PackCell), cudaMemcpyDefault, L->st));
            if (L->r_np[p]) CUDA_TRY(cudaMemcpyAsync(L->recv_part.p + (size_t)rp * L->psize, S->send_part.p + (size_t)sp * S->psize, (size_t)L->r_np[p] * L->psize, cudaMemcpyDefault, L->st));
            rl += L->r_nl[p]; rn += L->r_nn[p]; rp += L->r_np[p];
        }
        CUDA_TRY(cudaEventRecord(L->ev_done, L->st));
        L->done_recorded = true;
        CUDA_TRY(cudaStreamSynchronize(L->st));
    }
    return PN2_OK;
}

// Receiver side: append the received cells / ghost particles to the unified arrays (main stream, after the exchange
// on the LET stream has completed)
int pn2_let_unpack(pn2_ctx *h) {
    cudaStream_t st = h->stream;
    h->nrl = h->nrn = h->nrp = 0;
    h->peer_roots.clear();
    LetState *L = h->nranks > 1 ? let_state(h) : nullptr;
    const int np = L ? L->npeer : 0;
    if (L) {
        if (!L->done_recorded) { pn2_set_error("pn2_step_finish: the LET blocks were not exchanged (pn2_let_exchange_nccl / pn2_exchange_local)"); return PN2_ERR_STATE; }
        CUDA_TRY(cudaStreamWaitEvent(st, L->ev_done, 0));
    }
    long rl = 0, rn = 0, rp = 0;
    for (int p = 0; p < np; p++) { rl += L->r_nl[p]; rn += L->r_nn[p]; rp += L->r_np[p]; }
    if ((size_t)h->ncell + rl + rn >= (1u << PN2_IMG_SHIFT)) { pn2_set_error("pn2: more than 2^27 cells incl. the received LET"); return PN2_ERR_ARG; }
    h->nrl = (int)rl; h->nrn = (int)rn; h->nrp = (int)rp;
    h->info.n_let_nodes = rl + rn; h->info.n_let_bodies = rp;
    if (rl + rn > 0) {
        size_t nc = (size_t)h->ncell + rl + rn;
        PN2_TRY(h->geom.ensure(6 * nc + 6, true, st)); PN2_TRY(h->son.ensure(2 * nc + 2, true, st)); PN2_TRY(h->desc.ensure(nc + 1, true, st));
        PN2_TRY(h->M.ensure(NM * nc + NM, true, st));
        if (h->prm.precision != PN2_FP32) PN2_TRY(h->pos.ensure(3 * ((size_t)h->n + rp) + 3, true, st));
        else PN2_TRY(h->rel.ensure((size_t)h->n + rp + 1, true, st));
        long ol = 0, on = 0, op = 0;
        for (int p = 0; p < np; p++) {
            int nl = (int)L->r_nl[p], nn = (int)L->r_nn[p];
            if (nl + nn > 0) {
                let_unpack_kernel<<<(nl + nn + 127) / 128, 128, 0, st>>>(nl, nn, L->recv_leaf.p + ol, L->recv_node.p + on, h->ncell + (int)ol,
                                                                       h->ncell + (int)rl + (int)on, h->n + (int)op, h->geom.p, h->son.p,
                                                                       h->desc.p, h->M.p);
                h->launches++;
            }
            if (nn > 0) h->peer_roots.push_back(h->ncell + (int)rl + (int)on);      // the peer's root is its first packed node
            ol += nl; on += nn; op += L->r_np[p];
        }
        if (rp > 0) {
            if (h->prm.precision != PN2_FP32)
                CUDA_TRY(cudaMemcpyAsync(h->pos.p + 3 * (size_t)h->n, L->recv_part.p, (size_t)rp * 24, cudaMemcpyDeviceToDevice, st));
            else
                CUDA_TRY(cudaMemcpyAsync(h->rel.p + h->n, L->recv_part.p, (size_t)rp * 16, cudaMemcpyDeviceToDevice, st));
        }
    }
    KERNEL_CHECK();
    return PN2_OK;
}

// F(root) of one walk pass as the first span (unit 1 ..; the bump pointer starts behind it): which = 0: the local root
// and its periodic images (src/fmm.c:1028-1045 with the rank itself as sender); which = 1: every received root and its
// images (src/fmm.c:1021-1027, 1039-1045).  Enqueued on the main stream; the host copy lives in the context.
int pn2_walk_set_roots(pn2_ctx *h, int which) {
    const int nimg = h->prm.periodic ? 27 : 1;
    std::vector<unsigned> roots;
    if (which == 0) {
        if (h->nnode > 0) for (int s = 0; s < nimg; s++) roots.push_back((unsigned)h->nleaf | ((unsigned)s << PN2_IMG_SHIFT));
    } else {
        for (int r : h->peer_roots) for (int s = 0; s < nimg; s++) roots.push_back((unsigned)r | ((unsigned)s << PN2_IMG_SHIFT));
    }
    const size_t nroot = roots.size();
    const unsigned units = 1 + (unsigned)((nroot + 3) / 4);
    std::vector<unsigned> &span = h->root_span_host;
    span.assign(4 * (size_t)units, 0);
    span[0] = (unsigned)nroot;
    for (size_t k = 0; k < nroot; k++) span[4 + k] = roots[k];
    if (h->span_cap16 < 1 + (unsigned long long)units) { pn2_set_error("pn2: span buffer not allocated"); return PN2_ERR_STATE; }
    // a pageable source: the runtime stages it before returning, so the vector may be reused by the next pass
    CUDA_TRY(cudaMemcpyAsync(h->spans.p + 4, span.data(), span.size() * sizeof(unsigned), cudaMemcpyHostToDevice, h->stream));
    h->root_head = 1;
    h->root_units = units;
    h->root_count = (int)nroot;
    return PN2_OK;
}

// ---- communicator helpers ----
extern "C" int pn2_comm_unique_id(void *out128) {
    if (!out128) { pn2_set_error("pn2_comm_unique_id: null"); return PN2_ERR_ARG; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!nccl_load()) return PN2_ERR_NCCL;
    ncclUniqueId id;
    NCCL_TRY(ncclGetUniqueId(&id));
    memcpy(out128, &id, 128);
    return PN2_OK;
}

extern "C" int pn2_comm_init_rank(pn2_ctx *h, int rank, int nranks, const pn2_domain *all, const void *id128) {
    if (!h || !all || !id128 || nranks < 1 || rank < 0 || rank >= nranks) { pn2_set_error("pn2_comm_init_rank: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (!nccl_load()) return PN2_ERR_NCCL;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    NCCL_TRY(ncclCommInitRank(&comm, nranks, id, rank));
    h->rank = rank; h->nranks = nranks; h->nccl = comm; h->own_comm = true;
    h->all_dom.assign(all, all + nranks);
    return PN2_OK;
}

void pn2_comm_release(pn2_ctx *h) {
    if (h->own_comm && h->nccl && g_nccl.ok) ncclCommDestroy((ncclComm_t)h->nccl);
    h->nccl = nullptr; h->own_comm = false;
}
