// pn2_operators.cuh -- Cartesian order-3 multipole operators (FP64) as device functions.
//
// Restates the operator maths of the reference (src/operator.c) with compile-time
// multi-index tables; every loop below has constant bounds and is fully unrolled, so the index
// arithmetic disappears at compile time.
//
// Storage order (inc/operator.h:24-67): for multi-index (a,b,c), o = a+b+c, t = b+c:
//     idx = {0,1,4,10}[o] + t(t+1)/2 + c
#pragma once
#include "pn2_common.cuh"
#define PN2_M2LTAB_QUAL static __device__
#include "pn2_m2ltab.h"

namespace pn2op {

__host__ __device__ constexpr int obase(int o) { return o == 0 ? 0 : (o == 1 ? 1 : (o == 2 ? 4 : 10)); }
__host__ __device__ constexpr int midx(int a, int b, int c) { return obase(a + b + c) + (b + c) * (b + c + 1) / 2 + c; }
// inverse table
__host__ __device__ constexpr int mi_x(int i) {
    constexpr int t[NM] = {0, 1, 0, 0, 2, 1, 1, 0, 0, 0, 3, 2, 2, 1, 1, 1, 0, 0, 0, 0};
    return t[i];
}
__host__ __device__ constexpr int mi_y(int i) {
    constexpr int t[NM] = {0, 0, 1, 0, 0, 1, 0, 2, 1, 0, 0, 1, 0, 2, 1, 0, 3, 2, 1, 0};
    return t[i];
}
__host__ __device__ constexpr int mi_z(int i) {
    constexpr int t[NM] = {0, 0, 0, 1, 0, 0, 1, 0, 1, 2, 0, 0, 1, 0, 1, 2, 0, 1, 2, 3};
    return t[i];
}
__host__ __device__ constexpr int mi_o(int i) { return mi_x(i) + mi_y(i) + mi_z(i); }
__host__ __device__ constexpr double inv_fact(int k) { return k <= 1 ? 1.0 : (k == 2 ? 0.5 : 1.0 / 6.0); }
__host__ __device__ constexpr double inv_mfact(int i) { return inv_fact(mi_x(i)) * inv_fact(mi_y(i)) * inv_fact(mi_z(i)); }

// t[i] = s^i / i!  for the 20 stored multi-indices
__device__ __forceinline__ void taylor_terms(double sx, double sy, double sz, double t[NM]) {
    double px[4] = {1.0, sx, sx * sx, sx * sx * sx};
    double py[4] = {1.0, sy, sy * sy, sy * sy * sy};
    double pz[4] = {1.0, sz, sz * sz, sz * sz * sz};
#pragma unroll
    for (int i = 0; i < NM; i++) t[i] = px[mi_x(i)] * py[mi_y(i)] * pz[mi_z(i)] * inv_mfact(i);
}

// P2M (src/operator.c:13-93): M_n += m (-d)^n / n!,  d = x_p - c
__device__ __forceinline__ void p2m_add(double dx, double dy, double dz, double m, double M[NM]) {
    double t[NM];
    taylor_terms(-dx, -dy, -dz, t);
#pragma unroll
    for (int i = 0; i < NM; i++) M[i] += m * t[i];
}

// M2M (src/operator.c:96-160): M'_n += sum_{k<=n} M_{n-k} s^k/k!,  s = c_parent - c_child
__device__ __forceinline__ void m2m_add(double sx, double sy, double sz, const double M[NM], double out[NM]) {
    double t[NM];
    taylor_terms(sx, sy, sz, t);
#pragma unroll
    for (int n = 0; n < NM; n++) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NM; k++) {
            if (mi_x(k) <= mi_x(n) && mi_y(k) <= mi_y(n) && mi_z(k) <= mi_z(n))
                a += M[midx(mi_x(n) - mi_x(k), mi_y(n) - mi_y(k), mi_z(n) - mi_z(k))] * t[k];
        }
        out[n] += a;
    }
}

// L2L (src/operator.c:395-494): L'_n += sum_{|k|<=3-|n|} L_{n+k} s^k/k!,  s = c_child - c_parent
__device__ __forceinline__ void l2l_add(double sx, double sy, double sz, const double L[NM], double out[NM]) {
    double t[NM];
    taylor_terms(sx, sy, sz, t);
#pragma unroll
    for (int n = 0; n < NM; n++) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NM; k++) {
            if (mi_o(k) + mi_o(n) <= 3) a += L[midx(mi_x(n) + mi_x(k), mi_y(n) + mi_y(k), mi_z(n) + mi_z(k))] * t[k];
        }
        out[n] += a;
    }
}

// L2P (src/operator.c:197-251): a_c += sum_{|n|<=2} L_{n+e_c} d^n/n!,  d = x_p - c_leaf
__device__ __forceinline__ void l2p_eval(double dx, double dy, double dz, const double L[NM], double a[3]) {
    double t[NM];
    taylor_terms(dx, dy, dz, t);
    double ax = 0.0, ay = 0.0, az = 0.0;
#pragma unroll
    for (int n = 0; n < 10; n++) {
        ax += L[midx(mi_x(n) + 1, mi_y(n), mi_z(n))] * t[n];
        ay += L[midx(mi_x(n), mi_y(n) + 1, mi_z(n))] * t[n];
        az += L[midx(mi_x(n), mi_y(n), mi_z(n) + 1)] * t[n];
    }
    a[0] = ax; a[1] = ay; a[2] = az;
}

// M2L (src/operator.c:255-392): L_n += sum_{|m|<=3-|n|} M_m D_{n+m}(R),  R = c_sink - c_source.
//   f_k = ((1/r) d/dr)^k G(r);  G = erfc(r/2rs)/r with LONGSHORT (:294-307), 1/r otherwise (:288-292)
//   D_0 = f0, D_i = f1 R_i, D_ij = f2 R_i R_j + f1 delta_ij,
//   D_ijk = f3 R_i R_j R_k + f2 (delta_ij R_k + delta_ik R_j + delta_jk R_i)
// E(u) = erfc(u) and X(u) = exp(-u^2)/sqrt(pi) from the piecewise degree-10 tables of pn2_m2ltab.h (tools/fit_m2l64.py:
// |error| <= 2e-16, i.e. the last bit of double; u >= 6.4: exactly 0), the tables in shared memory
struct M2LTab {
    const double (*E)[PN2_M2LTAB_KPAD];
    const double (*X)[PN2_M2LTAB_KPAD];
};
__device__ __forceinline__ void m2l_tab_eval(const M2LTab &tab, double u, double &E, double &X) {
    const double t = fma(u, PN2_M2LTAB_INVH, -0.5);
    const double tm = t + 6755399441055744.0;              // round to nearest by the 2^52 + 2^51 trick (no F2I / I2F)
    const int k = __double2loint(tm);
    const double d = t - (tm - 6755399441055744.0);
    const int kc = k < PN2_M2LTAB_K - 1 ? k : PN2_M2LTAB_K - 1;
    double e = tab.E[PN2_M2LTAB_DEG][kc], x = tab.X[PN2_M2LTAB_DEG][kc];
#pragma unroll
    for (int j = PN2_M2LTAB_DEG - 1; j >= 0; j--) { e = fma(e, d, tab.E[j][kc]); x = fma(x, d, tab.X[j][kc]); }
    E = e; X = x;
}

// tab != nullptr (long/short build): no libm, no division, no square root: 1/r from MUFU.RSQ64H + two Newton steps
__device__ __forceinline__ void m2l_add(double Rx, double Ry, double Rz, const double M[NM], double L[NM], double rs,
                                        int longshort, const M2LTab *tab = nullptr, double inv2rs = 0.0) {
    double r2 = Rx * Rx + Ry * Ry + Rz * Rz;
    double f0, f1, f2, f3;
    if (longshort && tab) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2));
        double e = fma(-(r2 * y), y, 1.0);
        y = fma(0.5 * y, e, y);                             // ~1e-12
        e = fma(-(r2 * y), y, 1.0);
        const double ir = fma(0.5 * y, e, y);               // 1/r to the last bits
        const double ir2 = ir * ir, r = r2 * ir;
        const double irs = 2.0 * inv2rs;                    // 1 / rs
        double E, X;
        m2l_tab_eval(*tab, r * inv2rs, E, X);
        const double irs2 = irs * irs;
        const double a = X * irs;                           // X / rs
        const double b = E * ir + a;
        f0 = E * ir;
        f1 = -b * ir2;
        f2 = (3.0 * b * ir2 + 0.5 * a * irs2) * ir2;
        f3 = -((15.0 * b * ir2 + 2.5 * a * irs2) * ir2 + 0.25 * a * irs2 * irs2) * ir2;
    } else {
    double r = sqrt(r2);
    double ir = 1.0 / r, ir2 = ir * ir;
    if (longshort) {
        double irs = 1.0 / rs;
        double u = 0.5 * r * irs;
        double X = exp(-u * u) * 0.56418958354775628695;   // 1/sqrt(pi)
        double E = erfc(u);
        double irs2 = irs * irs;
        double a = X * irs;                                // X / rs
        f0 = E * ir;
        f1 = -(E * ir + a) * ir2;                          // -(E + r X/rs)/r^3
        f2 = (3.0 * (E * ir + a) * ir2 + 0.5 * a * irs2) * ir2;
        f3 = -((15.0 * (E * ir + a) * ir2 + 2.5 * a * irs2) * ir2 + 0.25 * a * irs2 * irs2) * ir2;
    } else {
        f0 = ir; f1 = -ir * ir2; f2 = 3.0 * ir * ir2 * ir2; f3 = -15.0 * ir * ir2 * ir2 * ir2;
    }
    }
    double R[3] = {Rx, Ry, Rz};
    double D[NM];
    D[0] = f0;
#pragma unroll
    for (int i = 1; i < NM; i++) {
        const int a = mi_x(i), b = mi_y(i), c = mi_z(i), o = a + b + c;
        double px = a == 0 ? 1.0 : (a == 1 ? Rx : (a == 2 ? Rx * Rx : Rx * Rx * Rx));
        double py = b == 0 ? 1.0 : (b == 1 ? Ry : (b == 2 ? Ry * Ry : Ry * Ry * Ry));
        double pz = c == 0 ? 1.0 : (c == 1 ? Rz : (c == 2 ? Rz * Rz : Rz * Rz * Rz));
        double rr = px * py * pz;
        if (o == 1) D[i] = f1 * rr;
        else if (o == 2) D[i] = f2 * rr + ((a == 2 || b == 2 || c == 2) ? f1 : 0.0);
        else {
            // sum over the three index pairs of delta_(pair) * R_(remaining index)
            double lin = 0.0;
            const int e[3] = {a, b, c};
#pragma unroll
            for (int d = 0; d < 3; d++) {
                if (e[d] == 3) lin += 3.0 * R[d];
                else if (e[d] == 2) {
#pragma unroll
                    for (int g = 0; g < 3; g++) if (g != d && e[g] == 1) lin += R[g];
                }
            }
            D[i] = f3 * rr + f2 * lin;
        }
    }
#pragma unroll
    for (int n = 0; n < NM; n++) {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < NM; m++) {
            if (mi_o(m) + mi_o(n) <= 3) acc += M[m] * D[midx(mi_x(n) + mi_x(m), mi_y(n) + mi_y(m), mi_z(n) + mi_z(m))];
        }
        L[n] += acc;
    }
}

}  // namespace pn2op
