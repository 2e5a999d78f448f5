// pn2_migrate.cu -- domain decomposition on the device (src/domains.c:163-375).
//
// The reference partitions each rank's Body array in place along the heap-numbered domain tree
// (prepare_body_inOrderOf_domain -> bksort_body_inplace: pos[D] > split goes right, D cycling x, y, z from the root),
// which leaves the array ordered by destination rank, then ships the blocks with an all-to-all-v
// (prepare_deliver_realloc_body).  Here every record finds its owner by descending the same tree (nranks - 1
// comparisons of the same doubles: the owner is bit-identical to the reference's), a stable radix sort by owner orders
// the records, and the blocks travel with grouped ncclSend / ncclRecv over NVLink (or device-to-device copies
// between the contexts of one process).  Records arrive in the reference's order of blocks (by source rank); inside
// a block the order is the sender's original order (the reference's is its partition order) -- the tree builder
// does not depend on it (Morton pre-sort).
#include <cub/cub.cuh>
#include "pn2_nccl.cuh"

struct MigState {
    DBuf<int> owner, owner2, idx, idx2, cnt_dev;
    DBuf<double> split, send, recv;
    DBuf<unsigned char> tmp;
    std::vector<int> sendcount, recvcount;
    int rec = 0, n = 0, n_recv = 0;
    bool packed = false, received = false;
};
static MigState *mig_state(pn2_ctx *h) {
    if (!h->mig) h->mig = new MigState();
    return h->mig;
}
void pn2_migrate_release(pn2_ctx *h) {
    if (!h->mig) return;
    MigState *M = h->mig;
    M->owner.release(); M->owner2.release(); M->idx.release(); M->idx2.release(); M->cnt_dev.release();
    M->split.release(); M->send.release(); M->recv.release(); M->tmp.release();
    delete M;
    h->mig = nullptr;
}

// src/initial.c:199-223: the heap index of the left-most domain node
static int mostleft_of(int P) {
    if (P == 1) return 0;
    int m = 1;
    while (m < 2 * P - 1) m *= 2;
    return m / 2 - 1;
}

__global__ void owner_kernel(int n, const double *__restrict__ rec, int rd, const double *__restrict__ split, int P,
                             int mostleft, int *__restrict__ owner, int *__restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = rec + (size_t)i * rd;
    int node = 0, D = 0;
    while (node < P - 1) {                                   // first_domain = P - 1 (src/initial.c:217)
        node = 2 * node + 1 + (p[D] > split[node] ? 1 : 0);  // src/domains.c:166-169: pos[D] > split counts right
        D = D == 2 ? 0 : D + 1;
    }
    owner[i] = (node - mostleft + P) % P;                    // src/domains.c:275
    if (idx) idx[i] = i;
}

__global__ void block_bounds_kernel(int n, const int *__restrict__ sorted_owner, int P, int *__restrict__ start) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;           // start[r] = first k with owner >= r; start[P] = n
    if (r > P) return;
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (sorted_owner[mid] < r) lo = mid + 1; else hi = mid;
    }
    start[r] = lo;
}

__global__ void gather_records_kernel(long total, int rd, const double *__restrict__ rec, const int *__restrict__ idx,
                                      double *__restrict__ out) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    long k = t / rd;
    int c = (int)(t - k * rd);
    out[t] = rec[(size_t)idx[k] * rd + c];
}

static int check_split(const char *who, pn2_ctx *h, const double *d_rec, int rd, int n, const double *split, int nranks) {
    if (!h || n < 0 || rd < 3 || !split || nranks < 1 || (n > 0 && !d_rec)) { pn2_set_error("%s: bad argument", who); return PN2_ERR_ARG; }
    return PN2_OK;
}

static int upload_split(pn2_ctx *h, MigState *M, const double *split, int P) {
    PN2_TRY(M->split.ensure(2 * (size_t)P));
    if (P > 1) CUDA_TRY(cudaMemcpyAsync(M->split.p, split, (size_t)(P - 1) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    return PN2_OK;
}

extern "C" int pn2_domain_owner_device(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, int nranks,
                                       int *d_owner) {
    PN2_TRY(check_split("pn2_domain_owner_device", h, d_rec, rec_doubles, n, split, nranks));
    if (n > 0 && !d_owner) { pn2_set_error("pn2_domain_owner_device: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    MigState *M = mig_state(h);
    PN2_TRY(upload_split(h, M, split, nranks));
    if (n > 0) {
        owner_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(n, d_rec, rec_doubles, M->split.p, nranks, mostleft_of(nranks), d_owner, nullptr);
        h->launches++;
    }
    KERNEL_CHECK();
    CUDA_TRY(cudaStreamSynchronize(h->stream));              // `split` is host memory of the caller
    return PN2_OK;
}

extern "C" int pn2_migrate_begin(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, int *sendcount) {
    PN2_TRY(check_split("pn2_migrate_begin", h, d_rec, rec_doubles, n, split, h ? h->nranks : 1));
    CUDA_TRY(cudaSetDevice(h->device));
    const int P = h->nranks;
    cudaStream_t st = h->stream;
    MigState *M = mig_state(h);
    M->packed = M->received = false;
    M->rec = rec_doubles; M->n = n;
    M->sendcount.assign(P, 0); M->recvcount.assign(P, 0);
    PN2_TRY(upload_split(h, M, split, P));
    PN2_TRY(M->owner.ensure(n + 1)); PN2_TRY(M->owner2.ensure(n + 1)); PN2_TRY(M->idx.ensure(n + 1)); PN2_TRY(M->idx2.ensure(n + 1));
    PN2_TRY(M->cnt_dev.ensure(4 * (size_t)P + 8));
    PN2_TRY(M->send.ensure((size_t)n * rec_doubles + 1));
    if (n > 0) {
        owner_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_rec, rec_doubles, M->split.p, P, mostleft_of(P), M->owner.p, M->idx.p);
        int bits = 1;
        while ((1 << bits) < P) bits++;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, M->owner.p, M->owner2.p, M->idx.p, M->idx2.p, n, 0, bits, st);
        PN2_TRY(M->tmp.ensure(tb + 16));
        cub::DeviceRadixSort::SortPairs(M->tmp.p, tb, M->owner.p, M->owner2.p, M->idx.p, M->idx2.p, n, 0, bits, st);   // stable
        block_bounds_kernel<<<(P + 256) / 256, 256, 0, st>>>(n, M->owner2.p, P, M->cnt_dev.p);
        const long total = (long)n * rec_doubles;
        gather_records_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, rec_doubles, d_rec, M->idx2.p, M->send.p);
        h->launches += 4;
        std::vector<int> start(P + 1);
        CUDA_TRY(cudaMemcpyAsync(start.data(), M->cnt_dev.p, (size_t)(P + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int r = 0; r < P; r++) M->sendcount[r] = start[r + 1] - start[r];
    } else {
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    KERNEL_CHECK();
    if (sendcount) memcpy(sendcount, M->sendcount.data(), (size_t)P * sizeof(int));
    M->packed = true;
    return PN2_OK;
}

static int finish_counts(pn2_ctx *h, MigState *M) {
    long tot = 0;
    for (int c : M->recvcount) tot += c;
    if (tot > 2147483647L) { pn2_set_error("pn2_migrate: more than 2^31 records on one rank"); return PN2_ERR_ARG; }
    M->n_recv = (int)tot;
    PN2_TRY(M->recv.ensure((size_t)tot * M->rec + 1));
    return PN2_OK;
}

extern "C" int pn2_migrate_exchange_nccl(pn2_ctx *h) {
    if (!h || !h->mig || !h->mig->packed) { pn2_set_error("pn2_migrate_exchange_nccl: no pn2_migrate_begin"); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    MigState *M = h->mig;
    const int P = h->nranks, me = h->rank;
    cudaStream_t st = h->stream;
    const size_t rb = (size_t)M->rec * sizeof(double);
    if (P == 1) {
        M->recvcount[0] = M->sendcount[0];
        PN2_TRY(finish_counts(h, M));
        if (M->n_recv) CUDA_TRY(cudaMemcpyAsync(M->recv.p, M->send.p, (size_t)M->n_recv * rb, cudaMemcpyDeviceToDevice, st));
        M->received = true;
        return PN2_OK;
    }
    if (!h->nccl) { pn2_set_error("pn2: no NCCL communicator (pn2_set_comm / pn2_comm_init_rank)"); return PN2_ERR_STATE; }
    if (!nccl_load()) return PN2_ERR_NCCL;
    ncclComm_t comm = (ncclComm_t)h->nccl;
    // counts (MPI_Alltoall, src/domains.c:313), record size checked on the way
    std::vector<int> sc(2 * (size_t)P), rc(2 * (size_t)P);
    for (int r = 0; r < P; r++) { sc[2 * r] = M->sendcount[r]; sc[2 * r + 1] = M->rec; }
    int *dsc = M->cnt_dev.p, *drc = M->cnt_dev.p + 2 * P;
    CUDA_TRY(cudaMemcpyAsync(dsc, sc.data(), sc.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    NCCL_TRY(ncclGroupStart());
    for (int r = 0; r < P; r++) {
        if (r == me) continue;
        NCCL_TRY(ncclSend(dsc + 2 * r, 2, ncclInt, r, comm, st));
        NCCL_TRY(ncclRecv(drc + 2 * r, 2, ncclInt, r, comm, st));
    }
    NCCL_TRY(ncclGroupEnd());
    CUDA_TRY(cudaMemcpyAsync(rc.data(), drc, rc.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int r = 0; r < P; r++) {
        if (r == me) { M->recvcount[r] = M->sendcount[r]; continue; }
        if (rc[2 * r + 1] != M->rec) { pn2_set_error("pn2_migrate: rank %d sends records of %d doubles, expected %d", r, rc[2 * r + 1], M->rec); return PN2_ERR_ARG; }
        M->recvcount[r] = rc[2 * r];
    }
    PN2_TRY(finish_counts(h, M));
    // payload (MPI_Isend / MPI_Recv per peer, src/domains.c:333-351)
    size_t so = 0, ro = 0;
    NCCL_TRY(ncclGroupStart());
    for (int r = 0; r < P; r++) {
        const size_t sb = (size_t)M->sendcount[r] * rb, rbytes = (size_t)M->recvcount[r] * rb;
        if (r == me) {
            if (sb) CUDA_TRY(cudaMemcpyAsync((char *)M->recv.p + ro, (const char *)M->send.p + so, sb, cudaMemcpyDeviceToDevice, st));
        } else {
            if (sb) NCCL_TRY(ncclSend((const char *)M->send.p + so, sb, ncclChar, r, comm, st));
            if (rbytes) NCCL_TRY(ncclRecv((char *)M->recv.p + ro, rbytes, ncclChar, r, comm, st));
        }
        so += sb; ro += rbytes;
    }
    NCCL_TRY(ncclGroupEnd());
    M->received = true;
    return PN2_OK;
}

extern "C" int pn2_migrate_exchange_local(pn2_ctx **hs, int nranks) {
    if (!hs || nranks < 1) { pn2_set_error("pn2_migrate_exchange_local: bad argument"); return PN2_ERR_ARG; }
    for (int r = 0; r < nranks; r++) {
        if (!hs[r] || hs[r]->nranks != nranks || hs[r]->rank != r || !hs[r]->mig || !hs[r]->mig->packed ||
            hs[r]->mig->rec != hs[0]->mig->rec) {
            pn2_set_error("pn2_migrate_exchange_local: context %d is not a packed rank %d of %d", r, r, nranks);
            return PN2_ERR_ARG;
        }
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaStreamSynchronize(hs[r]->stream));
    }
    for (int r = 0; r < nranks; r++) {
        pn2_ctx *h = hs[r];
        MigState *M = h->mig;
        const size_t rb = (size_t)M->rec * sizeof(double);
        for (int s = 0; s < nranks; s++) M->recvcount[s] = hs[s]->mig->sendcount[r];
        CUDA_TRY(cudaSetDevice(h->device));
        PN2_TRY(finish_counts(h, M));
        size_t ro = 0;
        for (int s = 0; s < nranks; s++) {
            MigState *S = hs[s]->mig;
            size_t so = 0;
            for (int k = 0; k < r; k++) so += (size_t)S->sendcount[k] * rb;
            const size_t nb = (size_t)M->recvcount[s] * rb;
            if (nb) CUDA_TRY(cudaMemcpyAsync((char *)M->recv.p + ro, (const char *)S->send.p + so, nb, cudaMemcpyDefault, h->stream));
            ro += nb;
        }
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        M->received = true;
    }
    return PN2_OK;
}

extern "C" int pn2_migrate_result(pn2_ctx *h, double **d_rec_out, int *n_out, int *recvcount) {
    if (!h || !h->mig || !h->mig->received) { pn2_set_error("pn2_migrate_result: no completed exchange"); return PN2_ERR_STATE; }
    if (d_rec_out) *d_rec_out = h->mig->recv.p;
    if (n_out) *n_out = h->mig->n_recv;
    if (recvcount) memcpy(recvcount, h->mig->recvcount.data(), h->mig->recvcount.size() * sizeof(int));
    return PN2_OK;
}

extern "C" int pn2_migrate_fetch(pn2_ctx *h, double *rec_host_out) {
    if (!h || !h->mig || !h->mig->received) { pn2_set_error("pn2_migrate_fetch: no completed exchange"); return PN2_ERR_STATE; }
    MigState *M = h->mig;
    if (M->n_recv > 0 && !rec_host_out) { pn2_set_error("pn2_migrate_fetch: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (M->n_recv > 0)
        CUDA_TRY(cudaMemcpyAsync(rec_host_out, M->recv.p, (size_t)M->n_recv * M->rec * sizeof(double), cudaMemcpyDefault, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PN2_OK;
}

extern "C" int pn2_migrate_device(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, double **d_rec_out,
                                  int *n_out) {
    PN2_TRY(pn2_migrate_begin(h, d_rec, rec_doubles, n, split, nullptr));
    PN2_TRY(pn2_migrate_exchange_nccl(h));
    return pn2_migrate_result(h, d_rec_out, n_out, nullptr);
}
