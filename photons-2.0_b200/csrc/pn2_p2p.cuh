// pn2_p2p.cuh -- the P2P leaf-pair direct sum with the PM/FMM Gaussian force split, as warp-level
// device code shared by the list-driven kernel (pn2_p2p.cu) and the fused walk+P2P kernel (pn2_walk.cu).
//
// Reference arithmetic: task_compute_p2p src/fmm.c:823-855, p2p_kernel_ex src/remotes.c:26-56:
//     a_i += sum_j (x_j - x_i) m S(r) g(r / 2rs),  S = 1/max(r, eps)^3,  g(u) = erfc(u) + 2u/sqrt(pi) exp(-u^2)
// Self pairs: the reference skips jp == ip by index (src/fmm.c:831); here dx = 0 contributes exactly 0
// with r^2 floored at a tiny constant, which also reproduces "distinct particles at r = 0 add 0".
//
// Mapping (one warp = one sink leaf): lane = slice * SW + sink, SW = sink slots (8, 16 or 32 >= MAXLEAF),
// NSL = 32 / SW slices.  A "stage" is NSL source leaves: lane (q, j) loads particle j of the q-th leaf,
// converts it to sink-leaf-relative coordinates and stores it to shared memory; after one __syncwarp
// every lane streams the SW particles of its slice's leaf with broadcast LDS.128.  Slices are reduced
// with __shfl_xor at the end, so every sink is owned by one warp: no atomics, deterministic sums.
//
// FP32 mode: coordinates are leaf-centre-relative and in units of 2 rs (u = r directly):
//     3 FADD (dx) + 3 FFMA (r^2) + MUFU.RSQ + FMUL (u) + FMUL + MUFU.EX2 (exp(-u^2)) + 9 FFMA (1 + u^2 R(u))
//     + FMNMX (softening) + 2 FMUL (1/r^3) + 2 FMUL (e, Q) + 3 FFMA (accumulate)  = 24 FMA-pipe + 2 MUFU + 1 ALU
#pragma once
#include "pn2_common.cuh"

#define PN2_RDEG 8                // g(u) = exp(-u^2) (1 + u^2 R(u)), deg R = 8: |err g| <= 3.3e-7 (tools/fit_g.py)
#define PN2_PAD_COORD 24.0f      // padding sources sit at u >= 24: exp(-u^2) underflows to exactly 0

template <int SW>
struct P2PStageF32 {
    static constexpr int NSL = 32 / SW;
    static constexpr int ROW = SW + 1;                 // float4 row stride: +1 (16 B) de-conflicts the NSL broadcast rows
    static constexpr int STAGE_F4 = NSL * ROW;
};

__device__ __forceinline__ float pn2_ex2(float x) {     // bare MUFU.EX2 (flushes denormals: exp(-u^2) < 1e-38 is 0)
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float pn2_rsqrt(float x) {   // bare MUFU.RSQ
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// interaction of one sink lane with one staged source particle (FP32, units of 2 rs)
template <bool LONGSHORT>
__device__ __forceinline__ void p2p_interact_f32(const float4 pj, float xi, float yi, float zi, float &ax, float &ay,
                                                 float &az, const float (&q)[PN2_RDEG + 1], float inv_eps) {
    float dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
    float r2 = fmaf(dx, dx, 1e-30f);
    r2 = fmaf(dy, dy, r2);
    r2 = fmaf(dz, dz, r2);
    float rinv = pn2_rsqrt(r2);
    float ri = fminf(rinv, inv_eps);
    float s = ri * ri * ri;
    if (LONGSHORT) {
        float u = r2 * rinv;
        float e = pn2_ex2(r2 * -1.4426950408889634f);
        float Q = q[PN2_RDEG];
#pragma unroll
        for (int k = PN2_RDEG - 1; k >= 0; k--) Q = fmaf(Q, u, q[k]);
        Q = fmaf(Q, r2, 1.0f);
        s = s * e * Q;
    } else {
        s *= pj.w;
    }
    ax = fmaf(dx, s, ax);
    ay = fmaf(dy, s, ay);
    az = fmaf(dz, s, az);
}

// FP64 parity arithmetic, absolute coordinates: the reference's own expression (src/fmm.c:834-852)
__device__ __forceinline__ void p2p_interact_f64(double xj, double yj, double zj, double wj, double xi, double yi, double zi,
                                                 double &ax, double &ay, double &az, double eps, double inv2rs,
                                                 int longshort) {
    double dx = xj - xi, dy = yj - yi, dz = zj - zi;
    double r2 = dx * dx + dy * dy + dz * dz;
    double dr = sqrt(r2);
    double ir3 = (dr < eps) ? wj / (eps * eps * eps) : wj / (dr * r2);
    if (r2 == 0.0) ir3 = 0.0;                          // self / coincident: dx = 0 anyway, avoid 0 * inf
    if (longshort) {
        double u = dr * inv2rs;
        ir3 *= erfc(u) + 1.1283791670955126 * u * exp(-u * u);
    }
    ax += dx * ir3; ay += dy * ir3; az += dz * ir3;
}
