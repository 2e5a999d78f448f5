// ubench_p2p.cu -- pipe microbenchmarks behind the P2P kernel design: what the FMA pipe, the packed
// FP32x2 forms and the MUFU unit deliver on this GPU, and the speed of light of the P2P inner loop alone (sources
// resident in shared memory, no list walk).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_p2p ubench_p2p.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../photons-2.0_b200/csrc/pn2_p2p.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>   // 0: FFMA, 1: FFMA2, 2: FFMA2 + MUFU (12:2), 3: FFMA + MUFU (24:2), 4: MUFU only, 5: FFMA2 3-reg operands
__global__ void __launch_bounds__(256) pipe_kernel(float *out, int iters, float a, float b) {
    float c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3f + i;
    float m0 = a + threadIdx.x, m1 = b + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fmaf(c[i], a, b);
        } else if (MODE == 1 || MODE == 5) {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    pn2_f2 v = pk2(c[i], c[i + 1]);
                    if (MODE == 1) v = fma2(v, pk2(a, a), pk2(b, b));
                    else v = fma2(v, pk2(c[(i + 2) & 15], c[(i + 3) & 15]), pk2(c[(i + 4) & 15], c[(i + 5) & 15]));
                    unpk2(v, c[i], c[i + 1]);
                }
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                pn2_f2 v = pk2(c[i], c[i + 1]);
                v = fma2(v, pk2(a, a), pk2(b, b));
                v = fma2(v, pk2(a, a), pk2(b, b));
                v = fma2(v, pk2(a, a), pk2(b, b));
                unpk2(v, c[i], c[i + 1]);
            }
            m0 = pn2_rsqrt(m0); m1 = pn2_ex2(m1); m0 = pn2_rsqrt(m0); m1 = pn2_ex2(m1);
        } else if (MODE == 3) {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) c[i] = fmaf(c[i], a, b);
            m0 = pn2_rsqrt(m0); m1 = pn2_ex2(m1); m0 = pn2_rsqrt(m0); m1 = pn2_ex2(m1);
        } else {
#pragma unroll
            for (int r = 0; r < 4; r++) { c[0] = pn2_rsqrt(c[0]); c[1] = pn2_ex2(c[1]); c[2] = pn2_rsqrt(c[2]); c[3] = pn2_ex2(c[3]); }
        }
    }
    float s = m0 + m1;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// P2P inner loop only: every warp keeps one stage (NSL leaves of 8 sources) in shared memory and replays it
template <int PACKED, int MINB>
__global__ void __launch_bounds__(128, MINB) p2p_sol_kernel(float *out, int stages, P2PConst pc) {
    constexpr int SW = 8;
    using ST = P2PStageF32<SW>;
    __shared__ float4 sm[4][2][ST::STAGE_F4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, q = lane / SW, j = lane % SW;
    const float xi = 0.01f * j, yi = 0.02f * j, zi = 0.03f * j;
    float x = 0.3f + 0.05f * lane, y = 0.1f * q, z = 0.07f * j;
    const float inv_eps = pc.inv_eps;
    float ax = 0, ay = 0, az = 0;
    if (PACKED) {
        P2PSinkPk sk;
        sk.nx = pk2(-xi, -xi); sk.ny = pk2(-yi, -yi); sk.nz = pk2(-zi, -zi);
        sk.ax = sk.ay = sk.az = pk2(0.f, 0.f);
        for (int b = 0; b < 2; b++) pk_store<SW, true>(reinterpret_cast<float *>(sm[wib][b]), q, j, x, y, z, 1.f);
        __syncwarp();
        for (int s = 0; s < stages; s++) {
            const float o = s * 1e-6f;               // a new centre offset per stage, as in the real kernel (and no hoisting)
            sk.nx = pk2(o - xi, o - xi); sk.ny = pk2(o - yi, o - yi); sk.nz = pk2(o - zi, o - zi);
            pk_row<SW, true>(reinterpret_cast<float *>(sm[wib][s & 1]), q, sk, inv_eps);
        }
        float lo, hi;
        unpk2(sk.ax, lo, hi); ax = lo + hi; unpk2(sk.ay, lo, hi); ay = lo + hi; unpk2(sk.az, lo, hi); az = lo + hi;
    } else {
        for (int b = 0; b < 2; b++) sm[wib][b][q * ST::ROW + j] = make_float4(x, y, z, 1.f);
        __syncwarp();
        for (int s = 0; s < stages; s++) {
            const float4 *row = &sm[wib][s & 1][q * ST::ROW];
            const float o = s * 1e-6f;
#pragma unroll
            for (int k = 0; k < SW; k++) p2p_interact_f32<true>(row[k], xi - o, yi - o, zi - o, ax, ay, az, inv_eps);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ax + ay + az;
}

static float timeit(void (*launch)(void), int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

static float *g_out; static int g_iters = 4000, g_grid, g_stages = 4000; static P2PConst g_pc;
template <int M> static void lp() { pipe_kernel<M><<<g_grid, 256>>>(g_out, g_iters, 1.0001f, 1e-3f); }
template <int P, int B> static void ls() { p2p_sol_kernel<P, B><<<g_grid, 128>>>(g_out, g_stages, g_pc); }

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("%s, %d SMs, max clock %d MHz\n", pr.name, pr.multiProcessorCount, clk / 1000);
    CK(cudaMalloc(&g_out, 64 << 20));
    const int sms = pr.multiProcessorCount;
    const double nominal = sms * 128.0 * clk * 1e3;
    g_grid = sms * 8;                    // 8 x 256 threads = 64 warps / SM
    double thr = (double)g_grid * 256;
    float ms;
    ms = timeit(lp<0>, 5); printf("FFMA  (reg,UR,UR)      : %.2f Tops/s (%.3f of 128/clk/SM)\n", thr * g_iters * 48 / ms / 1e9, thr * g_iters * 48 / ms / 1e-3 / nominal);
    ms = timeit(lp<1>, 5); printf("FFMA2 (pair,UR,UR)     : %.2f Tops/s (%.3f) [2 ops per instruction]\n", thr * g_iters * 48 / ms / 1e9, thr * g_iters * 48 / ms / 1e-3 / nominal);
    ms = timeit(lp<5>, 5); printf("FFMA2 (3 reg pairs)    : %.2f Tops/s (%.3f)\n", thr * g_iters * 48 / ms / 1e9, thr * g_iters * 48 / ms / 1e-3 / nominal);
    ms = timeit(lp<2>, 5); printf("FFMA2 + MUFU 24:4      : %.2f Tops/s FMA (%.3f)\n", thr * g_iters * 48 / ms / 1e9, thr * g_iters * 48 / ms / 1e-3 / nominal);
    ms = timeit(lp<3>, 5); printf("FFMA  + MUFU 48:4      : %.2f Tops/s FMA (%.3f)\n", thr * g_iters * 48 / ms / 1e9, thr * g_iters * 48 / ms / 1e-3 / nominal);
    ms = timeit(lp<4>, 5); printf("MUFU only              : %.2f Tmufu/s = %.2f per clk per SM\n", thr * g_iters * 16 / ms / 1e9, thr * g_iters * 16 / ms / 1e-3 / (sms * clk * 1e3));
    // P2P inner loop
    memset(&g_pc, 0, sizeof g_pc);
    g_pc.inv_eps = 40.f; g_pc.longshort = 1;
    auto report = [&](const char *name, float ms_, int blocks) {
        double inter = (double)sms * blocks * 128 * g_stages * 8;
        printf("%-34s: %.1f Gint/s, %.2f Tops/s = %.3f of nominal FMA peak\n", name, inter / ms_ / 1e6, inter * 24 / ms_ / 1e9, inter * 24 / ms_ / 1e-3 / nominal);
    };
    g_grid = sms * 8; ms = timeit(ls<0, 8>, 5); report("P2P loop scalar, 8 CTA/SM (32 warps)", ms, 8);
    g_grid = sms * 8; ms = timeit(ls<1, 8>, 5); report("P2P loop packed, 8 CTA/SM (32 warps)", ms, 8);
    g_grid = sms * 5; ms = timeit(ls<1, 5>, 5); report("P2P loop packed, 5 CTA/SM (20 warps)", ms, 5);
    g_grid = sms * 4; ms = timeit(ls<1, 4>, 5); report("P2P loop packed, 4 CTA/SM (16 warps)", ms, 4);
    g_grid = sms * 2; ms = timeit(ls<1, 2>, 5); report("P2P loop packed, 2 CTA/SM (8 warps)", ms, 2);
    g_grid = sms * 1; ms = timeit(ls<1, 1>, 5); report("P2P loop packed, 1 CTA/SM (4 warps)", ms, 1);
    g_grid = sms * 4; ms = timeit(ls<0, 4>, 5); report("P2P loop scalar, 4 CTA/SM (16 warps)", ms, 4);
    return 0;
}
