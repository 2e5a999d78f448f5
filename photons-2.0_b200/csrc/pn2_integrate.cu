// pn2_integrate.cu -- the KDK update of the reference's driver loop (src/photoNs.c:150-196, 254-268) on device-resident
// Body records, so that a step needs no host round trip of positions and velocities.  HBM-bound
// streaming: 96 bytes read + 24 written per particle and kernel; explicit __dmul_rn / __dadd_rn keep the reference's
// rounding (mul, then add -- no FMA), so the results are bit-identical to the CPU loops.
#include "pn2_common.cuh"

#define BODY_DOUBLES 12        // inc/typesdef.h:25-31: pos[3], acc[3], vel[3], acc_pm[3]

__global__ void kick_kernel(long total, double *__restrict__ body, double dkh, int pm_first) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (particle, component)
    if (t >= total) return;
    const long i = t / 3;
    const int d = (int)(t - 3 * i);
    double *b = body + (size_t)i * BODY_DOUBLES;
    const double a1 = pm_first ? b[9 + d] : b[3 + d], a2 = pm_first ? b[3 + d] : b[9 + d];
    double v = b[6 + d];
    v = __dadd_rn(v, __dmul_rn(a1, dkh));
    v = __dadd_rn(v, __dmul_rn(a2, dkh));
    b[6 + d] = v;
}

__global__ void drift_kernel(long total, double *__restrict__ body, double dd, double box) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long i = t / 3;
    const int d = (int)(t - 3 * i);
    double *b = body + (size_t)i * BODY_DOUBLES;
    double x = __dadd_rn(b[d], __dmul_rn(b[6 + d], dd));
    while (x < 0.0) x = __dadd_rn(x, box);              // src/photoNs.c:177-195
    while (x >= box) x = __dsub_rn(x, box);
    b[d] = x;
}

extern "C" int pn2_kick_device(pn2_ctx *h, double *d_body, int n, double dkh, int pm_first) {
    if (!h || n < 0 || (n > 0 && !d_body)) { pn2_set_error("pn2_kick_device: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (n == 0) return PN2_OK;
    const long total = 3L * n;
    kick_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(total, d_body, dkh, pm_first);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

extern "C" int pn2_drift_device(pn2_ctx *h, double *d_body, int n, double dd, double box) {
    if (!h || n < 0 || (n > 0 && !d_body) || !(box > 0.0)) { pn2_set_error("pn2_drift_device: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (n == 0) return PN2_OK;
    const long total = 3L * n;
    drift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(total, d_body, dd, box);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

// ---- force step on records: positions gathered from / accelerations scattered to the strided records ----
__global__ void gather_pos_kernel(long total, const double *__restrict__ rec, int rd, double *__restrict__ pos) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long i = t / 3;
    pos[t] = rec[(size_t)i * rd + (t - 3 * i)];
}
__global__ void scatter_acc_records_kernel(long total, const double *__restrict__ acc, double *__restrict__ rec, int rd, int off) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long i = t / 3;
    rec[(size_t)i * rd + off + (t - 3 * i)] = acc[t];
}

extern "C" int pn2_force_step_records(pn2_ctx *h, double *d_rec, int rec_doubles, int acc_offset, int n, const pn2_domain *dom) {
    if (!h || n < 0 || (n > 0 && !d_rec) || rec_doubles < 6 || acc_offset < 3 || acc_offset + 3 > rec_doubles) {
        pn2_set_error("pn2_force_step_records: bad argument");
        return PN2_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(h->device));
    PN2_TRY(h->rec_pos.ensure(3 * (size_t)n + 3)); PN2_TRY(h->rec_acc.ensure(3 * (size_t)n + 3));
    const long total = 3L * n;
    if (n > 0) {
        gather_pos_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(total, d_rec, rec_doubles, h->rec_pos.p);
        h->launches++;
    }
    PN2_TRY(pn2_force_step_device(h, h->rec_pos.p, n, dom, h->rec_acc.p));
    if (n > 0) {
        scatter_acc_records_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(total, h->rec_acc.p, d_rec, rec_doubles, acc_offset);
        h->launches++;
    }
    KERNEL_CHECK();
    return PN2_OK;
}

// ---- Gadget-2 snapshot blocks <-> device Body records (src/snapshot.c:211-293, 397-503) ----
// The file holds float32 pos[N][3] and float32 vel[N][3] = v / a^1.5.  The reader widens the positions and multiplies the
// widened velocities by gdt2unit = a^1.5 (:261-276); the writer stores (float)pos and (float)((float)vel / gdt2unit)
// (:465-480).  The blocks are uploaded as they are in the file (half the bytes of the records) and converted on the device.
__global__ void snap_to_body_kernel(long n, const float *__restrict__ pos32, const float *__restrict__ vel32, double gdt2unit,
                                    double *__restrict__ body) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double *b = body + (size_t)i * BODY_DOUBLES;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        b[d] = (double)pos32[3 * i + d];
        b[3 + d] = 0.0;
        b[6 + d] = vel32 ? __dmul_rn((double)vel32[3 * i + d], gdt2unit) : 0.0;
        b[9 + d] = 0.0;
    }
}
__global__ void body_to_snap_kernel(long n, const double *__restrict__ body, double gdt2unit, float *__restrict__ pos32,
                                    float *__restrict__ vel32) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *b = body + (size_t)i * BODY_DOUBLES;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        pos32[3 * i + d] = (float)b[d];
        vel32[3 * i + d] = (float)__ddiv_rn((double)(float)b[6 + d], gdt2unit);
    }
}
extern "C" int pn2_snapshot_to_body_device(pn2_ctx *h, const float *d_pos32, const float *d_vel32, int n, double gdt2unit, double *d_body) {
    if (!h || n < 0 || (n > 0 && (!d_pos32 || !d_body))) { pn2_set_error("pn2_snapshot_to_body_device: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (n == 0) return PN2_OK;
    snap_to_body_kernel<<<(unsigned)((n + 255L) / 256), 256, 0, h->stream>>>(n, d_pos32, d_vel32, gdt2unit, d_body);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}
extern "C" int pn2_body_to_snapshot_device(pn2_ctx *h, const double *d_body, int n, double gdt2unit, float *d_pos32, float *d_vel32) {
    if (!h || n < 0 || (n > 0 && (!d_pos32 || !d_vel32 || !d_body)) || !(gdt2unit > 0.0)) { pn2_set_error("pn2_body_to_snapshot_device: bad argument"); return PN2_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(h->device));
    if (n == 0) return PN2_OK;
    body_to_snap_kernel<<<(unsigned)((n + 255L) / 256), 256, 0, h->stream>>>(n, d_body, gdt2unit, d_pos32, d_vel32);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}
