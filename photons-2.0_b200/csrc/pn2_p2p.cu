// pn2_p2p.cu -- list-driven P2P kernels: one warp per sink leaf over a CSR interaction list.
// Replaces task_compute_p2p (src/fmm.c:796-872) and task_compute_p2p_ext / p2p_kernel_ex
// (src/remotes.c:583-596, 14-57).  See pn2_p2p.cuh for the mapping and the arithmetic.
#include "pn2_p2p.cuh"

#define P2P_WARPS 4   // warps per CTA

// ------------------------------------------------------------------------------------------------
// FP32: leaf-relative float4 sources, software-pipelined staging
// ------------------------------------------------------------------------------------------------
template <int SW, bool LONGSHORT>
__global__ void __launch_bounds__(P2P_WARPS * 32)
p2p_csr_f32_kernel(long nseg, const int *__restrict__ seg_sink, const long *__restrict__ seg_off,
                   const unsigned *__restrict__ src, const LeafDesc *__restrict__ sink_desc,
                   const float4 *__restrict__ sink_rel, const LeafDesc *__restrict__ src_desc,
                   const float4 *__restrict__ src_rel, double *__restrict__ acc, P2PConst pc,
                   unsigned long long *__restrict__ counters, int local_src) {
    using ST = P2PStageF32<SW>;
    constexpr int NSL = ST::NSL;
    __shared__ float4 sm[P2P_WARPS][2][ST::STAGE_F4];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long seg = (long)blockIdx.x * P2P_WARPS + wib;
    if (seg >= nseg) return;
    const int q = lane / SW, j = lane % SW;

    const float inv_eps = pc.inv_eps;

    const int sink = seg_sink[seg];
    const LeafDesc sd = sink_desc[sink];
    float xi = 0.f, yi = 0.f, zi = 0.f;
    if (j < sd.npart) {
        float4 p = sink_rel[sd.first + j];
        xi = p.x; yi = p.y; zi = p.z;
    }
    float ax = 0.f, ay = 0.f, az = 0.f;
#if PN2_PACKED
    P2PSinkPk sk;
    sk.nx = pk2(-xi, -xi); sk.ny = pk2(-yi, -yi); sk.nz = pk2(-zi, -zi);
    sk.ax = sk.ay = sk.az = pk2(0.f, 0.f);
#endif
    const long o0 = seg_off[seg], o1 = seg_off[seg + 1];
    unsigned long long nint = 0;

    // loader for the stage starting at list position `base`: returns this lane's staged particle
    auto load_stage = [&](long base) -> float4 {
        float4 p = make_float4(PN2_PAD_COORD, PN2_PAD_COORD, PN2_PAD_COORD, 0.f);
        long idx = base + q;
        if (idx < o1) {
            unsigned e = src[idx];
            unsigned cell = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
            LeafDesc d = src_desc[cell];
            if (j < d.npart) {
                float4 r = src_rel[d.first + j];
                // displacement of the source leaf centre from the sink leaf centre, in units of lambda
                float Dx = (float)(((d.c[0] + pc.shift[img][0]) - sd.c[0]) * pc.inv_len);
                float Dy = (float)(((d.c[1] + pc.shift[img][1]) - sd.c[1]) * pc.inv_len);
                float Dz = (float)(((d.c[2] + pc.shift[img][2]) - sd.c[2]) * pc.inv_len);
                p = make_float4(r.x + Dx, r.y + Dy, r.z + Dz, 1.f);
            }
            if (j == 0) nint += (unsigned long long)(d.npart - ((local_src && e == (unsigned)sink) ? 1 : 0));
        }
        return p;
    };

    int buf = 0;
    float4 pnext = load_stage(o0);
    for (long base = o0; base < o1; base += NSL) {
#if PN2_PACKED
        float *stg = reinterpret_cast<float *>(sm[wib][buf]);
        pk_store<SW, LONGSHORT>(stg, q, j, pnext.x, pnext.y, pnext.z, pnext.w);
        __syncwarp();
        if (base + NSL < o1) pnext = load_stage(base + NSL);   // in flight during the compute below
        pk_row<SW, LONGSHORT>(stg, q, sk, inv_eps);
#else
        sm[wib][buf][q * ST::ROW + j] = pnext;
        __syncwarp();
        if (base + NSL < o1) pnext = load_stage(base + NSL);   // in flight during the compute below
        const float4 *row = &sm[wib][buf][q * ST::ROW];
#pragma unroll
        for (int k = 0; k < SW; k++) p2p_interact_f32<LONGSHORT>(row[k], xi, yi, zi, ax, ay, az, inv_eps);
#endif
        buf ^= 1;
    }
#if PN2_PACKED
    { float lo, hi; unpk2(sk.ax, lo, hi); ax = lo + hi; unpk2(sk.ay, lo, hi); ay = lo + hi; unpk2(sk.az, lo, hi); az = lo + hi; }
#endif
    // reduce the slices
#pragma unroll
    for (int m = SW; m < 32; m <<= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, m);
        ay += __shfl_xor_sync(0xffffffffu, ay, m);
        az += __shfl_xor_sync(0xffffffffu, az, m);
        nint += __shfl_xor_sync(0xffffffffu, nint, m);
    }
    if (q == 0 && j < sd.npart) {
        // back to length units: positions were scaled by 1/lambda, so dx/r^3 carries (1/lambda)^2
        const double sc = pc.mass * pc.inv_len * pc.inv_len;
        double *a = acc + 3 * (size_t)(sd.first + j);
        a[0] += (double)ax * sc; a[1] += (double)ay * sc; a[2] += (double)az * sc;
    }
    if (lane == 0 && counters) atomicAdd(&counters[0], nint * (unsigned long long)sd.npart);
}

// ------------------------------------------------------------------------------------------------
// FP64 parity mode: absolute double positions, erfc/exp; same mapping, sources read through L1/L2
// ------------------------------------------------------------------------------------------------
template <int SW>
__global__ void __launch_bounds__(P2P_WARPS * 32)
p2p_csr_f64_kernel(long nseg, const int *__restrict__ seg_sink, const long *__restrict__ seg_off,
                   const unsigned *__restrict__ src, const LeafDesc *__restrict__ sink_desc,
                   const double *__restrict__ sink_pos, const LeafDesc *__restrict__ src_desc,
                   const double *__restrict__ src_pos, double *__restrict__ acc, P2PConst pc,
                   unsigned long long *__restrict__ counters, int local_src) {
    constexpr int NSL = 32 / SW;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long seg = (long)blockIdx.x * P2P_WARPS + wib;
    if (seg >= nseg) return;
    const int q = lane / SW, j = lane % SW;
    const int sink = seg_sink[seg];
    const LeafDesc sd = sink_desc[sink];
    double xi = 0, yi = 0, zi = 0;
    if (j < sd.npart) {
        const double *p = sink_pos + 3 * (size_t)(sd.first + j);
        xi = p[0]; yi = p[1]; zi = p[2];
    }
    double ax = 0, ay = 0, az = 0;
    const long o0 = seg_off[seg], o1 = seg_off[seg + 1];
    unsigned long long nint = 0;
    for (long idx = o0 + q; idx < o1; idx += NSL) {
        unsigned e = src[idx];
        unsigned cell = e & PN2_CELL_MASK, img = e >> PN2_IMG_SHIFT;
        LeafDesc d = src_desc[cell];
        const double sx = pc.shift[img][0], sy = pc.shift[img][1], sz = pc.shift[img][2];
        for (int k = 0; k < d.npart; k++) {
            const double *p = src_pos + 3 * (size_t)(d.first + k);
            // the reference adds the displacement to the ghost copy first (src/remotes.c:85-90)
            p2p_interact_f64(p[0] + sx, p[1] + sy, p[2] + sz, pc.mass, xi, yi, zi, ax, ay, az, pc.soft, pc.inv2rs,
                             pc.longshort);
        }
        if (j == 0) nint += (unsigned long long)(d.npart - ((local_src && e == (unsigned)sink) ? 1 : 0));
    }
#pragma unroll
    for (int m = SW; m < 32; m <<= 1) {
        ax += __shfl_xor_sync(0xffffffffu, ax, m);
        ay += __shfl_xor_sync(0xffffffffu, ay, m);
        az += __shfl_xor_sync(0xffffffffu, az, m);
        nint += __shfl_xor_sync(0xffffffffu, nint, m);
    }
    if (q == 0 && j < sd.npart) {
        double *a = acc + 3 * (size_t)(sd.first + j);
        a[0] += ax; a[1] += ay; a[2] += az;
    }
    if (lane == 0 && counters) atomicAdd(&counters[0], nint * (unsigned long long)sd.npart);
}

// leaf-centre-relative scaled coordinates: one thread per (leaf, slot)
__global__ void relpos_kernel(int nleaf, const LeafDesc *__restrict__ desc, const double *__restrict__ pos,
                              float4 *__restrict__ rel, double inv_len, int maxleaf) {
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    int k = (int)(t / maxleaf), s = (int)(t % maxleaf);
    if (k >= nleaf) return;
    LeafDesc d = desc[k];
    if (s >= d.npart) return;
    const double *p = pos + 3 * (size_t)(d.first + s);
    rel[d.first + s] = make_float4((float)((p[0] - d.c[0]) * inv_len), (float)((p[1] - d.c[1]) * inv_len),
                                   (float)((p[2] - d.c[2]) * inv_len), 1.f);
}

int pn2_launch_relpos(pn2_ctx *h, const double *pos, const LeafDesc *desc, int nleaf, float4 *rel, int n) {
    (void)n;
    if (nleaf == 0) return PN2_OK;
    long nt = (long)nleaf * h->prm.maxleaf;
    relpos_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, h->stream>>>(nleaf, desc, pos, rel, h->pc.inv_len, h->prm.maxleaf);
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

template <int SW>
static int launch_sw(pn2_ctx *h, const CsrList &list, const SourceSet &src, int local) {
    unsigned grid = (unsigned)((list.nseg + P2P_WARPS - 1) / P2P_WARPS);
    if (h->prm.precision != PN2_FP32) {
        p2p_csr_f64_kernel<SW><<<grid, P2P_WARPS * 32, 0, h->stream>>>(list.nseg, list.seg_sink, list.seg_off, list.src,
                                                                       h->desc.p, h->pos.p, src.desc, src.pos, h->acc.p,
                                                                       h->pc, h->counters.p, local);
    } else if (h->prm.longshort) {
        p2p_csr_f32_kernel<SW, true><<<grid, P2P_WARPS * 32, 0, h->stream>>>(list.nseg, list.seg_sink, list.seg_off,
                                                                             list.src, h->desc.p, h->rel.p, src.desc,
                                                                             src.rel, h->acc.p, h->pc, h->counters.p, local);
    } else {
        p2p_csr_f32_kernel<SW, false><<<grid, P2P_WARPS * 32, 0, h->stream>>>(list.nseg, list.seg_sink, list.seg_off,
                                                                              list.src, h->desc.p, h->rel.p, src.desc,
                                                                              src.rel, h->acc.p, h->pc, h->counters.p, local);
    }
    h->launches++;
    KERNEL_CHECK();
    return PN2_OK;
}

int pn2_launch_p2p(pn2_ctx *h, const CsrList &list, const SourceSet &src, bool src_is_local) {
    int local = src_is_local ? 1 : 0;
    if (list.nseg == 0) return PN2_OK;
    int ml = h->prm.maxleaf;
    if (ml <= 8) return launch_sw<8>(h, list, src, local);
    if (ml <= 16) return launch_sw<16>(h, list, src, local);
    return launch_sw<32>(h, list, src, local);
}
