// ubench_ops.cu -- FMA-pipe throughput by operand form (register-read bandwidth of FFMA2 / FMUL2 / FADD2 / FFMA).
// Reported as FMA-pipe "ops" (2 per packed instruction) relative to 128 per clock and SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../photons-2.0_b200/csrc/pn2_p2p.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) op_kernel(float *out, int iters, float ua, float ub) {
    pn2_f2 v[8];
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = pk2(c[2 * i], c[2 * i + 1]);
    const float x = threadIdx.x * 0.5f;
    const pn2_f2 d = pk2(x, x + 1.f), e = pk2(x + 2.f, x + 3.f), xx = pk2(x, x), uu = pk2(ua, ua);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int n = (i + 1) & 7, m = (i + 2) & 7;
                if (MODE == 0) v[i] = fma2(v[i], d, e);
                if (MODE == 1) v[i] = fma2(v[i], v[n], uu);
                if (MODE == 2) v[i] = fma2(v[i], v[n], pk2(1.5f, 1.5f));
                if (MODE == 3) v[i] = fma2(v[i], v[n], xx);
                if (MODE == 4) v[i] = fma2(v[i], v[i], v[n]);
                if (MODE == 5) v[i] = mul2(v[i], v[n]);
                if (MODE == 6) v[i] = add2(v[i], xx);
                if (MODE == 7) v[i] = fma2(v[i], v[n], v[m]);
                if (MODE == 8) v[i] = fma2(v[i], xx, v[n]);
                if (MODE == 9) v[i] = fma2(v[i], uu, v[n]);
            }
            if (MODE == 10) {
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fmaf(c[i], c[(i + 1) & 15], c[(i + 2) & 15]);
            }
            if (MODE == 11) {
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fmaf(c[i], c[(i + 1) & 15], 1.5f);
            }
            if (MODE == 12) {
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fmaf(c[i], c[i], c[(i + 1) & 15]);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { float lo, hi; unpk2(v[i], lo, hi); s += lo + hi; }
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float *g_out; static int g_iters = 2000, g_grid;
template <int M> static float run() {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    op_kernel<M><<<g_grid, 256>>>(g_out, g_iters, 1.0001f, 1e-3f); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        CK(cudaEventRecord(e0)); op_kernel<M><<<g_grid, 256>>>(g_out, g_iters, 1.0001f, 1e-3f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    CK(cudaMalloc(&g_out, 64 << 20));
    const int sms = pr.multiProcessorCount;
    const double nominal = sms * 128.0 * clk * 1e3;
    g_grid = sms * 8;
    const double ops = (double)g_grid * 256 * g_iters * 6 * 16;
    const char *names[13] = {"FFMA2 v, d, e (3 pairs, 2 invariant)", "FFMA2 v, w, UR", "FFMA2 v, w, imm", "FFMA2 v, w, R.F32", "FFMA2 v, v, w",
                             "FMUL2 v, w", "FADD2 v, R.F32", "FFMA2 v, w, z (3 distinct pairs)", "FFMA2 v, R.F32, w", "FFMA2 v, UR, w",
                             "FFMA c, a, b (3 distinct regs)", "FFMA c, a, imm", "FFMA c, c, a"};
    float ms[13] = {run<0>(), run<1>(), run<2>(), run<3>(), run<4>(), run<5>(), run<6>(), run<7>(), run<8>(), run<9>(), run<10>(), run<11>(), run<12>()};
    for (int i = 0; i < 13; i++) printf("%-40s: %.3f of 128 ops/clk/SM  (%.2f cycles per warp instruction)\n", names[i], ops / ms[i] / 1e-3 / nominal,
                                        (i < 10 ? 2.0 : 1.0) / (ops / ms[i] / 1e-3 / nominal));
    return 0;
}
