// pn2_pm.cu -- the particle-mesh long-range force on the device.
//
// Replaces partmesh_thread (src/partmesh.c:18-796: CIC deposit, mesh all-to-all into the FFT pencils, 4-point gradient,
// CIC gather) and subroutine convolution (src/conv.f90:128-247: 2DECOMP&FFT forward transform, Green function
// pref exp(-k^2 rs^2) sinc^-4 / k^2, backward transform).
//
// Layout: the whole NSIDE^3 periodic mesh lives in the HBM of every rank (1 GB at 512^3, 8.6 GB at 1024^3 -- against
// 180 GB), so the reference's pencil exchange (an MPI all-to-all-v of {x, y, z, value} keys, src/partmesh.c:188-352,
// and its inverse, :430-470) becomes ONE ncclAllReduce of the density mesh over NVLink / NVSwitch; every rank then
// transforms the same mesh with cuFFT (D2Z / Z2D, a library FFT like the reference's 2DECOMP) and gathers the force
// of its own particles.  No ghost layers: neighbours are found by periodic index arithmetic.
//
// Kernels (all HBM / L2-atomic bound; FP64 throughout like the reference):
//   pm_deposit_kernel : one thread per particle, the 8 CIC weights of src/partmesh.c:103-166 (same expressions, this
//                       file is compiled with -fmad=false), 8 atomicAdd(double) into the mesh
//   pm_green_kernel   : one thread per complex mode of the half spectrum; the per-axis factors exp(-l^2 a) / sinc^4 come
//                       from a host-built table (src/conv.f90:184-216 evaluated in double on the host), times
//                       pref / (l^2 + m^2 + n^2) and the deposit's renormalisation (NSIDE / BOX)^3 (:168-178)
//   pm_gather_kernel  : one thread per particle; 4-point differences f1 (u[+1] - u[-1]) - f2 (u[+2] - u[-2]) of the
//                       potential at the 8 CIC cells (src/partmesh.c:478-775), CIC-weighted
// Particles should be passed in tree order (pn2_pm_force_records after a force step does) so that the atomics and the
// 96 mesh reads per particle hit L2.
#include <cufft.h>
#include <dlfcn.h>
#include <math.h>
#include "pn2_nccl.cuh"

// cuFFT bound with dlopen (like NCCL: one copy per process, torch's when torch is loaded)
struct Pn2CufftApi {
    cufftResult (*Plan3d)(cufftHandle *, int, int, int, cufftType);
    cufftResult (*SetStream)(cufftHandle, cudaStream_t);
    cufftResult (*ExecD2Z)(cufftHandle, cufftDoubleReal *, cufftDoubleComplex *);
    cufftResult (*ExecZ2D)(cufftHandle, cufftDoubleComplex *, cufftDoubleReal *);
    cufftResult (*Destroy)(cufftHandle);
    bool ok = false;
};
static Pn2CufftApi g_fft;
static bool cufft_load() {
    if (g_fft.ok) return true;
    void *hd = dlopen("libcufft.so.11", RTLD_NOW | RTLD_NOLOAD);
    const char *env = getenv("PN2_CUFFT_LIB");
    if (!hd && env && *env) hd = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!hd) hd = dlopen("libcufft.so.11", RTLD_NOW | RTLD_GLOBAL);
    if (!hd) hd = dlopen("libcufft.so", RTLD_NOW | RTLD_GLOBAL);
    if (!hd) { pn2_set_error("pn2: cannot load libcufft.so.11: %s", dlerror()); return false; }
#define BIND(field, sym) *(void **)(&g_fft.field) = dlsym(hd, sym); if (!g_fft.field) { pn2_set_error("pn2: libcufft lacks %s", sym); return false; }
    BIND(Plan3d, "cufftPlan3d") BIND(SetStream, "cufftSetStream") BIND(ExecD2Z, "cufftExecD2Z") BIND(ExecZ2D, "cufftExecZ2D")
    BIND(Destroy, "cufftDestroy")
#undef BIND
    g_fft.ok = true;
    return true;
}
#define FFT_TRY(expr)                                                                                   \
    do {                                                                                                \
        cufftResult r_ = (expr);                                                                        \
        if (r_ != CUFFT_SUCCESS) { pn2_set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #expr, (int)r_); return PN2_ERR_CUDA; } \
    } while (0)

struct PmState {
    int nside = 0;
    cufftHandle plan_f = 0, plan_b = 0;
    bool have_plan = false;
    DBuf<double> mesh;                 // [N][N][N] density, then potential
    DBuf<double> spec;                 // [N][N][N/2+1] complex
    DBuf<double> axis;                 // [N] per-axis Green factor
    DBuf<double> pos, acc;             // packed staging of the records entry point
    double axis_rs = -1, axis_box = -1;
    bool open = false;                 // between pn2_pm_begin and pn2_pm_finish
    const double *d_pos = nullptr;
    int n = 0;
    double ms[4] = {0, 0, 0, 0};       // deposit, reduce, fft + green, gather
    cudaEvent_t ev[5] = {nullptr};
};

void pn2_pm_release(pn2_ctx *h) {
    if (!h->pm) return;
    PmState *P = h->pm;
    if (P->have_plan && g_fft.ok) { g_fft.Destroy(P->plan_f); g_fft.Destroy(P->plan_b); }
    P->mesh.release(); P->spec.release(); P->axis.release(); P->pos.release(); P->acc.release();
    for (int i = 0; i < 5; i++) if (P->ev[i]) cudaEventDestroy(P->ev[i]);
    delete P;
    h->pm = nullptr;
}

// the CIC cell / neighbour / weights of one coordinate: src/partmesh.c:103-116 (i = (int)(x norm), w = (x - (i + 1/2) delta) norm,
// neighbour on the side of the particle, weights w and 1 - w)
__device__ __forceinline__ void cic_axis(double x, double norm, double delta, int N, int &i, int &ii, double &w, double &wn) {
    i = (int)(x * norm);
    w = (x - (i + 0.5) * delta) * norm;
    if (w > 0) ii = i + 1;
    else { w = -w; ii = i - 1; }
    wn = 1.0 - w;
    // periodic wrap (the reference keeps ghost layers and wraps when it routes the cells: src/partmesh.c:262-275)
    i = i >= N ? i - N : (i < 0 ? i + N : i);
    ii = ii >= N ? ii - N : (ii < 0 ? ii + N : ii);
}

__global__ void pm_deposit_kernel(int n, const double *__restrict__ pos, int N, double norm, double delta, double mass,
                                  double *__restrict__ mesh) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int i, ii, j, jj, k, kk;
    double wi, win, wj, wjn, wk, wkn;
    cic_axis(pos[3 * (size_t)p], norm, delta, N, i, ii, wi, win);
    cic_axis(pos[3 * (size_t)p + 1], norm, delta, N, j, jj, wj, wjn);
    cic_axis(pos[3 * (size_t)p + 2], norm, delta, N, k, kk, wk, wkn);
    const size_t N1 = (size_t)N, N2 = N1 * N1;
    // src/partmesh.c:142-164, same products
    atomicAdd(&mesh[i * N2 + j * N1 + k], mass * win * wjn * wkn);
    atomicAdd(&mesh[ii * N2 + j * N1 + k], mass * wi * wjn * wkn);
    atomicAdd(&mesh[i * N2 + jj * N1 + k], mass * win * wj * wkn);
    atomicAdd(&mesh[i * N2 + j * N1 + kk], mass * win * wjn * wk);
    atomicAdd(&mesh[ii * N2 + jj * N1 + k], mass * wi * wj * wkn);
    atomicAdd(&mesh[ii * N2 + j * N1 + kk], mass * wi * wjn * wk);
    atomicAdd(&mesh[i * N2 + jj * N1 + kk], mass * win * wj * wk);
    atomicAdd(&mesh[ii * N2 + jj * N1 + kk], mass * wi * wj * wk);
}

// spectrum [N][N][N/2+1]: z *= scale T[l] T[m] T[n] / k2 (k = 0: scale alone, src/conv.f90:218-220)
__global__ void pm_green_kernel(int N, const double *__restrict__ axis, double scale, double2 *__restrict__ spec) {
    const int nh = N / 2 + 1;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)N * N * nh) return;
    const int c = (int)(t % nh);
    const int b = (int)((t / nh) % N);
    const int a = (int)(t / ((long)nh * N));
    const int half = N / 2;
    const int l = a > half ? a - N : a, m = b > half ? b - N : b, nn = c;      // c <= N / 2 is never wrapped (:180-183)
    const long k2i = (long)l * l + (long)m * m + (long)nn * nn;
    double gf = scale;
    if (k2i != 0) gf = scale * axis[a] * axis[b] * axis[c] / (double)k2i;
    double2 z = spec[t];
    z.x *= gf; z.y *= gf;
    spec[t] = z;
}

__device__ __forceinline__ int wrapN(int i, int N) { return i >= N ? i - N : (i < 0 ? i + N : i); }

// 4-point difference of the potential along axis `ax` at cell (a, b, c): src/partmesh.c:493-499 and the like
__device__ __forceinline__ double diff4(const double *__restrict__ u, int N, int a, int b, int c, int ax, double invx) {
    const size_t N1 = (size_t)N, N2 = N1 * N1;
    const double f1 = 4.0 / 3.0, f2 = 1.0 / 6.0;
    const int q = ax == 0 ? a : (ax == 1 ? b : c);                 // the coordinate that varies, and its stride
    const size_t sq = ax == 0 ? N2 : (ax == 1 ? N1 : 1);
    const size_t rest = a * N2 + b * N1 + c - q * sq;
    const double up1 = u[rest + wrapN(q + 1, N) * sq], um1 = u[rest + wrapN(q - 1, N) * sq];
    const double up2 = u[rest + wrapN(q + 2, N) * sq], um2 = u[rest + wrapN(q - 2, N) * sq];
    double d = f1 * invx * (up1 - um1);
    d -= f2 * invx * (up2 - um2);
    return d;
}

// one thread per (particle, component)
__global__ void __launch_bounds__(256) pm_gather_kernel(int n, const double *__restrict__ pos, int N, double norm, double delta, double invx,
                                                        const double *__restrict__ u, double *__restrict__ acc, int acc_stride) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3L * n) return;
    const int p = (int)(t / 3), ax = (int)(t - 3L * p);
    int i, ii, j, jj, k, kk;
    double wi, win, wj, wjn, wk, wkn;
    cic_axis(pos[3 * (size_t)p], norm, delta, N, i, ii, wi, win);
    cic_axis(pos[3 * (size_t)p + 1], norm, delta, N, j, jj, wj, wjn);
    cic_axis(pos[3 * (size_t)p + 2], norm, delta, N, k, kk, wk, wkn);
    // the 8 CIC cells in the reference's order dp[0..7]: (i,j,k) (ii,j,k) (i,jj,k) (ii,jj,k) (i,j,kk) (ii,j,kk) (i,jj,kk) (ii,jj,kk)
    const double d0 = diff4(u, N, i, j, k, ax, invx), d1 = diff4(u, N, ii, j, k, ax, invx);
    const double d2 = diff4(u, N, i, jj, k, ax, invx), d3 = diff4(u, N, ii, jj, k, ax, invx);
    const double d4 = diff4(u, N, i, j, kk, ax, invx), d5 = diff4(u, N, ii, j, kk, ax, invx);
    const double d6 = diff4(u, N, i, jj, kk, ax, invx), d7 = diff4(u, N, ii, jj, kk, ax, invx);
    // src/partmesh.c:575-582
    acc[(size_t)p * acc_stride + ax] = win * wjn * wkn * d0 + wi * wjn * wkn * d1 + win * wj * wkn * d2 + wi * wj * wkn * d3
                                       + win * wjn * wk * d4 + wi * wjn * wk * d5 + win * wj * wk * d6 + wi * wj * wk * d7;
}

static int pm_state(pn2_ctx *h, int nside, PmState **out) {
    if (!cufft_load()) return PN2_ERR_CUDA;
    if (!h->pm) h->pm = new PmState();
    PmState *P = h->pm;
    for (int i = 0; i < 5; i++) if (!P->ev[i]) CUDA_TRY(cudaEventCreate(&P->ev[i]));
    const size_t N = (size_t)nside;
    if (P->nside != nside) {
        if (P->have_plan) { g_fft.Destroy(P->plan_f); g_fft.Destroy(P->plan_b); P->have_plan = false; }
        PN2_TRY(P->mesh.ensure(N * N * N)); PN2_TRY(P->spec.ensure(2 * N * N * (N / 2 + 1))); PN2_TRY(P->axis.ensure(N));
        FFT_TRY(g_fft.Plan3d(&P->plan_f, nside, nside, nside, CUFFT_D2Z));
        FFT_TRY(g_fft.Plan3d(&P->plan_b, nside, nside, nside, CUFFT_Z2D));
        FFT_TRY(g_fft.SetStream(P->plan_f, h->stream));
        FFT_TRY(g_fft.SetStream(P->plan_b, h->stream));
        P->have_plan = true;
        P->nside = nside;
        P->axis_rs = -1;
    }
    if (P->axis_rs != h->prm.rs || P->axis_box != h->prm.box) {
        // per-axis factor of the Green function, src/conv.f90:172-216: exp(-l^2 (2 pi rs / box)^2) / sinc(pi l / N)^4 with the
        // reference's pi (a default-real literal widened to double, :143)
        const double PI_F = (double)3.1415926f;
        double ismth2 = 2 * PI_F * h->prm.rs / h->prm.box;
        ismth2 = ismth2 * ismth2;
        std::vector<double> ax(N);
        for (int a = 0; a < nside; a++) {
            int l = a > nside / 2 ? a - nside : a;
            double f = PI_F * l / nside;
            f = l == 0 ? 1.0 : sin(f) / f;
            double ff = 1.0 / f;
            ax[a] = exp(-((double)l * l) * ismth2) * ff * ff * ff * ff;
        }
        CUDA_TRY(cudaMemcpyAsync(P->axis.p, ax.data(), N * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        P->axis_rs = h->prm.rs; P->axis_box = h->prm.box;
    }
    *out = P;
    return PN2_OK;
}

// Phase 1: CIC deposit of this rank's particles into its copy of the mesh
extern "C" int pn2_pm_begin(pn2_ctx *h, const double *d_pos, int n, int nside) {
    if (!h || n < 0 || (n > 0 && !d_pos) || nside < 4 || nside > 4096) { pn2_set_error("pn2_pm_begin: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    PmState *P = nullptr;
    PN2_TRY(pm_state(h, nside, &P));
    cudaStream_t st = h->stream;
    const size_t N = (size_t)nside;
    CUDA_TRY(cudaEventRecord(P->ev[0], st));
    CUDA_TRY(cudaMemsetAsync(P->mesh.p, 0, N * N * N * sizeof(double), st));
    const double norm = nside / h->prm.box, delta = 1.0 / norm;                 // src/partmesh.c:98-99
    if (n > 0) {
        pm_deposit_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, d_pos, nside, norm, delta, h->prm.mass, P->mesh.p);
        h->launches++;
    }
    CUDA_TRY(cudaEventRecord(P->ev[1], st));
    KERNEL_CHECK();
    P->open = true; P->d_pos = d_pos; P->n = n;
    return PN2_OK;
}

// the sum of the ranks' meshes, in place on every rank (replaces the two all-to-all-v of src/partmesh.c:188-352, 430-470)
extern "C" int pn2_pm_reduce_nccl(pn2_ctx *h) {
    if (!h || !h->pm || !h->pm->open) { pn2_set_error("pn2_pm_reduce_nccl: no open PM step"); return PN2_ERR_STATE; }
    if (h->nranks <= 1) return PN2_OK;
    if (!h->nccl) { pn2_set_error("pn2: no NCCL communicator (pn2_set_comm / pn2_comm_init_rank)"); return PN2_ERR_STATE; }
    if (!nccl_load()) return PN2_ERR_NCCL;
    CUDA_TRY(cudaSetDevice(h->device));
    PmState *P = h->pm;
    const size_t N = (size_t)P->nside;
    NCCL_TRY(g_nccl.AllReduce(P->mesh.p, P->mesh.p, N * N * N, ncclDouble, ncclSum, (ncclComm_t)h->nccl, h->stream));
    return PN2_OK;
}

__global__ void pm_add_kernel(size_t n, double *__restrict__ a, const double *__restrict__ b) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] += b[i];
}

// all ranks are contexts of this process: rank 0's mesh accumulates the others' (ascending rank order), then is copied back
extern "C" int pn2_pm_reduce_local(pn2_ctx **hs, int nranks) {
    if (!hs || nranks < 1) { pn2_set_error("pn2_pm_reduce_local: bad argument"); return PN2_ERR_ARG; }
    for (int r = 0; r < nranks; r++) {
        if (!hs[r] || !hs[r]->pm || !hs[r]->pm->open || hs[r]->pm->nside != hs[0]->pm->nside) { pn2_set_error("pn2_pm_reduce_local: context %d has no matching open PM step", r); return PN2_ERR_STATE; }
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaStreamSynchronize(hs[r]->stream));
    }
    const size_t N = (size_t)hs[0]->pm->nside, N3 = N * N * N;
    pn2_ctx *h0 = hs[0];
    CUDA_TRY(cudaSetDevice(h0->device));
    DBuf<double> tmp;
    for (int r = 1; r < nranks; r++) {
        const double *src = hs[r]->pm->mesh.p;
        if (hs[r]->device != h0->device) {
            PN2_TRY(tmp.ensure(N3));
            CUDA_TRY(cudaMemcpyAsync(tmp.p, src, N3 * sizeof(double), cudaMemcpyDefault, h0->stream));
            src = tmp.p;
        }
        pm_add_kernel<<<(unsigned)((N3 + 255) / 256), 256, 0, h0->stream>>>(N3, h0->pm->mesh.p, src);
        h0->launches++;
    }
    CUDA_TRY(cudaStreamSynchronize(h0->stream));
    tmp.release();
    for (int r = 1; r < nranks; r++) {
        CUDA_TRY(cudaSetDevice(hs[r]->device));
        CUDA_TRY(cudaMemcpyAsync(hs[r]->pm->mesh.p, h0->pm->mesh.p, N3 * sizeof(double), cudaMemcpyDefault, hs[r]->stream));
        CUDA_TRY(cudaStreamSynchronize(hs[r]->stream));
    }
    return PN2_OK;
}

// Phase 2: convolution + gather.  d_acc_pm: acc_stride doubles between consecutive particles (3 = packed)
static int pm_finish(pn2_ctx *h, double *d_acc_pm, int acc_stride) {
    if (!h || !h->pm || !h->pm->open || (h->pm->n > 0 && !d_acc_pm)) { pn2_set_error("pn2_pm_finish: no open PM step / bad argument"); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    PmState *P = h->pm;
    cudaStream_t st = h->stream;
    const int nside = P->nside, n = P->n;
    const size_t N = (size_t)nside;
    P->open = false;
    CUDA_TRY(cudaEventRecord(P->ev[2], st));
    FFT_TRY(g_fft.ExecD2Z(P->plan_f, P->mesh.p, reinterpret_cast<cufftDoubleComplex *>(P->spec.p)));
    // pref of src/conv.f90:178 (with the 1 / N^3 of the unnormalised transforms) times the deposit's renormalisation
    // (NSIDE / BOX)^3 of src/partmesh.c:168-178, which is linear and therefore applied here
    const double PI_F = (double)3.1415926f;
    const double box = h->prm.box;
    const double pref = box * box / (PI_F * nside * nside * nside);
    double renormal = nside / box;
    renormal = renormal * renormal * renormal;
    const long nspec = (long)N * N * (N / 2 + 1);
    pm_green_kernel<<<(unsigned)((nspec + 255) / 256), 256, 0, st>>>(nside, P->axis.p, pref * renormal, reinterpret_cast<double2 *>(P->spec.p));
    FFT_TRY(g_fft.ExecZ2D(P->plan_b, reinterpret_cast<cufftDoubleComplex *>(P->spec.p), P->mesh.p));
    h->launches += 3;
    CUDA_TRY(cudaEventRecord(P->ev[3], st));
    const double norm = nside / box, delta = 1.0 / norm, invx = 0.5 * nside / box;   // src/partmesh.c:98-99, 474
    if (n > 0) {
        pm_gather_kernel<<<(unsigned)((3L * n + 255) / 256), 256, 0, st>>>(n, P->d_pos, nside, norm, delta, invx, P->mesh.p, d_acc_pm, acc_stride);
        h->launches++;
    }
    CUDA_TRY(cudaEventRecord(P->ev[4], st));
    KERNEL_CHECK();
    return PN2_OK;
}
extern "C" int pn2_pm_finish(pn2_ctx *h, double *d_acc_pm) { return pm_finish(h, d_acc_pm, 3); }

extern "C" int pn2_pm_force_device(pn2_ctx *h, const double *d_pos, int n, int nside, double *d_acc_pm) {
    PN2_TRY(pn2_pm_begin(h, d_pos, n, nside));
    if (h->nranks > 1) PN2_TRY(pn2_pm_reduce_nccl(h));
    return pn2_pm_finish(h, d_acc_pm);
}

__global__ void pm_gather_pos_kernel(long total, const double *__restrict__ rec, int rd, double *__restrict__ pos) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long i = t / 3;
    pos[t] = rec[(size_t)i * rd + (t - 3 * i)];
}

// on device-resident records (Body: acc_pm at doubles 9..11), so that PM + short-range + kick + drift need no host copy
extern "C" int pn2_pm_force_records(pn2_ctx *h, double *d_rec, int rec_doubles, int acc_pm_offset, int n, int nside) {
    if (!h || n < 0 || (n > 0 && !d_rec) || rec_doubles < 6 || acc_pm_offset < 3 || acc_pm_offset + 3 > rec_doubles) {
        pn2_set_error("pn2_pm_force_records: bad argument");
        return PN2_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    PmState *P = nullptr;
    PN2_TRY(pm_state(h, nside, &P));
    PN2_TRY(P->pos.ensure(3 * (size_t)n + 3));
    const long total = 3L * n;
    if (n > 0) {
        pm_gather_pos_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(total, d_rec, rec_doubles, P->pos.p);
        h->launches++;
    }
    PN2_TRY(pn2_pm_begin(h, P->pos.p, n, nside));
    if (h->nranks > 1) PN2_TRY(pn2_pm_reduce_nccl(h));
    return pm_finish(h, d_rec + acc_pm_offset, rec_doubles);
}

// inspection: the mesh as it is now (density between begin and finish, potential after finish), host array [N][N][N]
extern "C" int pn2_pm_get_mesh(pn2_ctx *h, double *mesh_host) {
    if (!h || !h->pm || !mesh_host || h->pm->nside == 0) { pn2_set_error("pn2_pm_get_mesh: no PM state"); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t N = (size_t)h->pm->nside;
    CUDA_TRY(cudaMemcpyAsync(mesh_host, h->pm->mesh.p, N * N * N * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PN2_OK;
}

// elapsed device time of the last PM evaluation (ms): deposit, mesh reduction, FFTs + Green function, gather
extern "C" int pn2_pm_get_timings(pn2_ctx *h, double ms[4]) {
    if (!h || !h->pm || !ms) { pn2_set_error("pn2_pm_get_timings: no PM state"); return PN2_ERR_STATE; }
    CUDA_TRY(cudaSetDevice(h->device));
    PmState *P = h->pm;
    CUDA_TRY(cudaEventSynchronize(P->ev[4]));
    for (int i = 0; i < 4; i++) {
        float t = 0;
        CUDA_TRY(cudaEventElapsedTime(&t, P->ev[i], P->ev[i + 1]));
        ms[i] = t;
    }
    return PN2_OK;
}
