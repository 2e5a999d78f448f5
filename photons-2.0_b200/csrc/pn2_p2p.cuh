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
// FP32 mode: coordinates are leaf-centre-relative and in units of lambda = 2 rs sqrt(ln 2) (see below):
//     3 FADD (dx) + 3 FFMA (r^2) + MUFU.RSQ + FMUL (u) + MUFU.EX2 (2^(-r^2)) + 8 FFMA (1 + u^2 R(u), deg R = 7)
//     + FMNMX (softening) + 2 FMUL (1/r^3) + 2 FMUL (e, Q) + 3 FFMA (accumulate)  = 22 FMA-pipe + 2 MUFU + 1 ALU
#pragma once
#include "pn2_common.cuh"

// g(u) = exp(-u^2) (1 + u^2 R(u)); R = weighted minimax polynomial from tools/fit_g.py, float32-checked.
// The coefficients are compile-time constants so that they become IMMEDIATE operands of FFMA / FFMA2: a third
// register operand costs FMA-pipe cycles (measured, tools/ubench/ubench_ops.cu: FFMA2 with two register pairs +
// immediate 2.04 cycles, + scalar register 2.21, three pairs 3.03).
#ifndef PN2_RDEG
#define PN2_RDEG 7                // deg R = 8: |err g| <= 3.3e-7;  7: <= 8.9e-7 (the default);  6: <= 3.3e-6
#endif
#if PN2_RDEG == 8
#define PN2_RCOEF {9.998987644e-01f, -7.508830079e-01f, 4.927910981e-01f, -2.807554342e-01f, 1.325016193e-01f, -4.790751029e-02f, 1.200598312e-02f, -1.810827398e-03f, 1.218777145e-04f}
#elif PN2_RDEG == 7
#define PN2_RCOEF {9.996877624e-01f, -7.485814399e-01f, 4.833320565e-01f, -2.610572656e-01f, 1.093096686e-01f, -3.186952180e-02f, 5.574667193e-03f, -4.318225181e-04f}
#elif PN2_RDEG == 6
#define PN2_RCOEF {9.990517726e-01f, -7.427107543e-01f, 4.632659763e-01f, -2.271847688e-01f, 7.821331668e-02f, -1.612388462e-02f, 1.459452176e-03f}
#else
#error "PN2_RDEG must be 6, 7 or 8"
#endif
// FP32 length unit: lambda = 2 rs sqrt(ln 2), so that exp(-u^2) = 2^(-r'^2) with r' = r / lambda and MUFU.EX2 takes
// -r'^2 directly (saves the multiplication by log2 e per interaction).  The polynomial is rescaled accordingly at
// compile time: 1 + u^2 R(u) = 1 + r'^2 R'(r'),  R'_k = R_k (sqrt(ln 2))^(k+2).
#define PN2_SQRT_LN2 0.83255461115769775635
struct Pn2RCoef {
    float v[PN2_RDEG + 1];
    constexpr Pn2RCoef() : v{} {
        constexpr double c[PN2_RDEG + 1] = PN2_RCOEF;
        double f = PN2_SQRT_LN2 * PN2_SQRT_LN2;
        for (int k = 0; k <= PN2_RDEG; k++) { v[k] = (float)(c[k] * f); f *= PN2_SQRT_LN2; }
    }
};
#define PN2_PAD_COORD 24.0f      // padding sources sit at u >= 24: exp(-u^2) underflows to exactly 0

#ifndef PN2_PACKED
#define PN2_PACKED 1              // 1: two sources per lane and step with packed FP32x2 arithmetic (FFMA2 / FADD2 / FMUL2)
#endif

// One stage in shared memory: NSL rows (one per slice), each the SW particles of one source leaf.
// Packed layout (PN2_PACKED): a row is SW/2 source PAIRS of 8 floats {x0 x1 y0 y1 | z0 z1 w0 w1}, so that one
// LDS.128 + one LDS.64 deliver the operands of the FP32x2 instructions already paired in even-aligned registers.
// Scalar layout: one float4 per source.  Row stride +16 B de-conflicts the NSL broadcast reads.
template <int SW>
struct P2PStageF32 {
    static constexpr int NSL = 32 / SW;
    static constexpr int ROW = SW + 1;                 // float4 units
    static constexpr int STAGE_F4 = NSL * ROW;
};

// ---- packed FP32x2 primitives (sm_100a: FADD2 / FMUL2 / FFMA2, one issue slot for two lanes' worth of FMA work) ----
typedef unsigned long long pn2_f2;                      // two floats in an even-aligned register pair {lo, hi}
__device__ __forceinline__ pn2_f2 pk2(float lo, float hi) {
    pn2_f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(pn2_f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pn2_f2 add2(pn2_f2 a, pn2_f2 b) {
    pn2_f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pn2_f2 mul2(pn2_f2 a, pn2_f2 b) {
    pn2_f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ pn2_f2 fma2(pn2_f2 a, pn2_f2 b, pn2_f2 c) {
    pn2_f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

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
                                                 float &az, float inv_eps) {
    constexpr Pn2RCoef qq{};
    float dx = pj.x - xi, dy = pj.y - yi, dz = pj.z - zi;
    float r2 = fmaf(dx, dx, 1e-30f);
    r2 = fmaf(dy, dy, r2);
    r2 = fmaf(dz, dz, r2);
    float rinv = pn2_rsqrt(r2);
    float ri = fminf(rinv, inv_eps);
    float s = ri * ri * ri;
    if (LONGSHORT) {
        float u = r2 * rinv;
        float e = pn2_ex2(-r2);
        float Q = qq.v[PN2_RDEG];
#pragma unroll
        for (int k = PN2_RDEG - 1; k >= 0; k--) Q = fmaf(Q, u, qq.v[k]);
        Q = fmaf(Q, r2, 1.0f);
        s = s * e * Q;
    } else {
        s *= pj.w;
    }
    ax = fmaf(dx, s, ax);
    ay = fmaf(dy, s, ay);
    az = fmaf(dz, s, az);
}

// Packed form: one sink lane against TWO staged sources.  Same operations as p2p_interact_f32, each FMA-pipe
// instruction doing both sources: 3 FADD2 + 3 FFMA2 + 2 FMUL2 (1/r^3) + FMUL2 (u) + 8 FFMA2
// + 2 FMUL2 + 3 FFMA2 = 22 FMA-pipe instructions, 4 MUFU, 2 FMNMX, 2 LDS per PAIR of interactions (15 issue
// slots per interaction instead of 28), so the FMA pipe (2 cycles per FP32x2 instruction), not the issue port, bounds it.
struct P2PSinkPk {
    pn2_f2 nx, ny, nz;      // (-xi, -xi) ...
    pn2_f2 ax, ay, az;      // two partial sums per component (even / odd sources)
};
template <bool LONGSHORT>
__device__ __forceinline__ void p2p_interact_pk(const float *pair /* 8 floats, 16-byte aligned */, P2PSinkPk &sk, float inv_eps) {
    constexpr Pn2RCoef qq{};
    const ulonglong2 xy = *reinterpret_cast<const ulonglong2 *>(pair);
    pn2_f2 dx = add2(xy.x, sk.nx), dy = add2(xy.y, sk.ny);
    pn2_f2 zz, ww = 0;
    if (LONGSHORT) zz = *reinterpret_cast<const pn2_f2 *>(pair + 4);
    else { const ulonglong2 zw = *reinterpret_cast<const ulonglong2 *>(pair + 4); zz = zw.x; ww = zw.y; }
    pn2_f2 dz = add2(zz, sk.nz);
    pn2_f2 r2 = fma2(dx, dx, pk2(1e-30f, 1e-30f));
    r2 = fma2(dy, dy, r2);
    r2 = fma2(dz, dz, r2);
    float r2a, r2b;
    unpk2(r2, r2a, r2b);
    const float rinva = pn2_rsqrt(r2a), rinvb = pn2_rsqrt(r2b);
    const pn2_f2 ri = pk2(fminf(rinva, inv_eps), fminf(rinvb, inv_eps));
    pn2_f2 s = mul2(mul2(ri, ri), ri);
    if (LONGSHORT) {
        const pn2_f2 u = mul2(r2, pk2(rinva, rinvb));
        const pn2_f2 e = pk2(pn2_ex2(-r2a), pn2_ex2(-r2b));
        pn2_f2 Q = pk2(qq.v[PN2_RDEG], qq.v[PN2_RDEG]);
#pragma unroll
        for (int k = PN2_RDEG - 1; k >= 0; k--) Q = fma2(Q, u, pk2(qq.v[k], qq.v[k]));
        Q = fma2(Q, r2, pk2(1.0f, 1.0f));
        s = mul2(mul2(s, e), Q);
    } else {
        s = mul2(s, ww);
    }
    sk.ax = fma2(dx, s, sk.ax);
    sk.ay = fma2(dy, s, sk.ay);
    sk.az = fma2(dz, s, sk.az);
}

// where lane (slice q, slot j) puts its staged source inside the packed stage (float index), and the row of slice q
template <int SW>
__device__ __forceinline__ int pk_store_index(int q, int j) { return q * (4 * P2PStageF32<SW>::ROW) + (j >> 1) * 8 + (j & 1); }
template <int SW, bool LONGSHORT>
__device__ __forceinline__ void pk_store(float *stage, int q, int j, float x, float y, float z, float w) {
    float *d = stage + pk_store_index<SW>(q, j);
    d[0] = x; d[2] = y; d[4] = z;
    if (!LONGSHORT) d[6] = w;
}
template <int SW, bool LONGSHORT>
__device__ __forceinline__ void pk_row(const float *stage, int q, P2PSinkPk &sk, float inv_eps) {
    const float *row = stage + q * (4 * P2PStageF32<SW>::ROW);
#pragma unroll
    for (int k = 0; k < SW / 2; k++) p2p_interact_pk<LONGSHORT>(row + 8 * k, sk, inv_eps);
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
