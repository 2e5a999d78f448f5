// pn2_tree.cu -- Mode B: the local k-d tree of src/fmm.c:30-264 built on the device.
//
// Same tree definition as the reference (build_kdtree / center_kdtree): binary splits at the coordinate
// MEAN of the node's particles, direction cycling x->y->z from direct_local_start, a child with
// <= MAXLEAF particles becomes a leaf, boxes are the domain box cut by the ancestors' splits.  The
// reference builds it by depth-first recursion with an in-place Hoare partition and a sequential
// FP64 sum (inherently serial); here it is built LEVEL BY LEVEL over all nodes of a depth at once, and
// entirely in INTEGER arithmetic so that a CPU restatement (oracle: pno_treeB_build) reproduces it bit
// for bit whatever the summation order:
//
//   0. q_d = trunc((x_d - lo_d) * 2^(32-e)) as uint32 per dimension, 2^e > largest box extent.
//      Morton pre-sort: 30-bit key from the top 10 bits of each q, LSD radix sort (stable: ties keep the caller's
//      order).  It fixes the particle order inside every leaf (the partitions below are stable) and gives the
//      deferred top levels their locality; 4 passes over 8-byte pairs instead of the 8 passes over 12-byte pairs of a
//      63-bit key (12 -> 3 ms at 512^3).
//   per level, direction dir = (direct0 + level) % 3, payload per particle = {qx, qy, qz, caller index}:
//   1. per-node sums of q_dir by segmented reduction + 64-bit integer atomics (nodesum_kernel) -- exact whatever the
//      order.  split = lo_dir + (sum / count) * 2^-(32-e) (only the boxes use the FP value).
//   2. flag_i = q_i * count > sum  ("> mean goes right", src/fmm.c:60-72; exact 64-bit products), exclusive
//      prefix sum of the flags gives every particle its slot in a STABLE partition of its node's range.
//   3. children: count <= MAXLEAF -> leaf, else node of the next level; ids are handed out by prefix sums over the
//      level (breadth-first numbering, deterministic).
//   4. scatter the 16-byte payload, the next node id and the next level's key to the other buffer.
//   ~70 B of HBM traffic per particle per level; the FP64 positions are gathered once at the end.
#include <cub/cub.cuh>
#include "pn2_common.cuh"

#define TB 256
static inline unsigned nb(long n) { return (unsigned)((n + TB - 1) / TB); }

// bit k of a 10-bit value -> bit 3k
__device__ __forceinline__ unsigned spread10(unsigned v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ unsigned quant32(double x, double lo, double S) {
    double f = __dmul_rn(__dsub_rn(x, lo), S);
    if (!(f > 0.0)) return 0u;
    if (f >= 4294967295.0) return 4294967295u;
    return (unsigned)f;                       // truncation
}

__global__ void quant_morton_kernel(int n, const double *__restrict__ pos, double lox, double loy, double loz, double S,
                                    uint4 *__restrict__ pay, unsigned *__restrict__ key, int *__restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned qx = quant32(pos[3 * (size_t)i], lox, S), qy = quant32(pos[3 * (size_t)i + 1], loy, S), qz = quant32(pos[3 * (size_t)i + 2], loz, S);
    pay[i] = make_uint4(qx, qy, qz, (unsigned)i);
    key[i] = (spread10(qx >> 22) << 2) | (spread10(qy >> 22) << 1) | spread10(qz >> 22);
    idx[i] = i;
}

__global__ void gather_pay_kernel(int n, const uint4 *__restrict__ pay_in, const int *__restrict__ idx, int dir,
                                  uint4 *__restrict__ pay_out, unsigned *__restrict__ qcur, int *__restrict__ seg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 p = pay_in[idx[i]];
    pay_out[i] = p;
    qcur[i] = dir == 0 ? p.x : (dir == 1 ? p.y : p.z);
    seg[i] = 0;                                          // everyone starts in the root
}

// Per-node coordinate sums without a prefix scan: the particles of a node are contiguous and `seg` (the node of a
// particle, -1 = already in a leaf) is constant along runs, so every thread sums its 8 consecutive items run by run and
// adds complete sums to n_sum with 64-bit integer atomics -- exact, order-independent, hence bit-reproducible.  Runs
// that cover a whole warp / a whole block are reduced with shuffles / shared memory first (top levels: one atomic
// per block instead of one per thread).  8 bytes of traffic per particle instead of the 16 of a 64-bit prefix scan.
#define NS_ITEMS 8
__global__ void __launch_bounds__(TB) nodesum_kernel(int n, const unsigned *__restrict__ q, const int *__restrict__ seg,
                                                     unsigned long long *__restrict__ n_sum) {
    __shared__ unsigned long long s_part[TB / 32];
    __shared__ int s_seg[TB / 32];
    const long base = ((long)blockIdx.x * TB + threadIdx.x) * NS_ITEMS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned qv[NS_ITEMS];
    int sv[NS_ITEMS];
    if (base + NS_ITEMS <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(q + base), b = *reinterpret_cast<const uint4 *>(q + base + 4);
        const int4 c = *reinterpret_cast<const int4 *>(seg + base), d = *reinterpret_cast<const int4 *>(seg + base + 4);
        qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = b.x; qv[5] = b.y; qv[6] = b.z; qv[7] = b.w;
        sv[0] = c.x; sv[1] = c.y; sv[2] = c.z; sv[3] = c.w; sv[4] = d.x; sv[5] = d.y; sv[6] = d.z; sv[7] = d.w;
    } else {
#pragma unroll
        for (int k = 0; k < NS_ITEMS; k++) {
            const bool in = base + k < n;
            qv[k] = in ? q[base + k] : 0u;
            sv[k] = in ? seg[base + k] : -1;
        }
    }
    // runs inside the thread: complete interior runs go out directly, the last run is kept
    int cur = sv[0];
    unsigned long long sum = qv[0];
    bool single = true;
#pragma unroll
    for (int k = 1; k < NS_ITEMS; k++) {
        if (sv[k] != cur) {
            if (cur >= 0) atomicAdd(&n_sum[cur], sum);
            cur = sv[k]; sum = 0; single = false;
        }
        sum += qv[k];
    }
    // (cur, sum) = the thread's last run; if the thread is one run and so is its warp, reduce the warp first
    const int s0 = __shfl_sync(0xffffffffu, cur, 0);
    const bool warp_uniform = __all_sync(0xffffffffu, single && cur == s0);
    if (!warp_uniform) {
        if (cur >= 0) atomicAdd(&n_sum[cur], sum);
        sum = 0; cur = -2;                                   // nothing left for the block stage
    } else {
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    }
    if (lane == 0) { s_part[wid] = sum; s_seg[wid] = cur; }
    __syncthreads();
    if (threadIdx.x == 0) {
        // consecutive warps of one run are merged: one atomic per run and block
        int rs = s_seg[0];
        unsigned long long rsum = s_part[0];
        for (int w = 1; w < TB / 32; w++) {
            if (s_seg[w] != rs) {
                if (rs >= 0) atomicAdd(&n_sum[rs], rsum);
                rs = s_seg[w]; rsum = 0;
            }
            rsum += s_part[w];
        }
        if (rs >= 0) atomicAdd(&n_sum[rs], rsum);
    }
}

// flag_i = "particle i goes right" = q_i * count > sum (exact 64-bit products; src/fmm.c:60-72), as a byte; the
// prefix sum of the flags (CUB, over the byte array) gives every particle its slot in the stable partition
__global__ void flag_kernel(int n, const unsigned *__restrict__ q, const int *__restrict__ seg,
                            const unsigned long long *__restrict__ n_sum, const int *__restrict__ n_count,
                            unsigned char *__restrict__ flag) {
    const long base = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;     // 4 items per thread: 16-byte loads, 4-byte store
    if (base > n) return;
    unsigned qv[4];
    int sv[4];
    if (base + 4 <= n) {
        const uint4 a = *reinterpret_cast<const uint4 *>(q + base);
        const int4 c = *reinterpret_cast<const int4 *>(seg + base);
        qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w;
        sv[0] = c.x; sv[1] = c.y; sv[2] = c.z; sv[3] = c.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const bool in = base + k < n;
            qv[k] = in ? q[base + k] : 0u;
            sv[k] = in ? seg[base + k] : -1;
        }
    }
    unsigned char f[4];
    int ls = -1, lc = 0;
    unsigned long long lsum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        f[k] = 0;
        const int s = sv[k];
        if (s >= 0) {
            if (s != ls) { ls = s; lc = n_count[s]; lsum = n_sum[s]; }       // consecutive particles share their node
            f[k] = (lc < 2) ? 1 : (((unsigned long long)qv[k] * (unsigned long long)lc > lsum) ? 1 : 0);    // len < 2: src/fmm.c:33-36
        }
    }
    // flag[] has n + 1 entries (the scan runs over n + 1 items); the buffer is padded to a multiple of 4
    *reinterpret_cast<uchar4 *>(flag + base) = make_uchar4(f[0], f[1], f[2], f[3]);
}
struct ByteToInt {
    __host__ __device__ int operator()(unsigned char v) const { return (int)v; }
};

// one thread per node of the level: split position from the node's coordinate sum
__global__ void split_kernel(int cnt, int node0, const int *__restrict__ n_count, const unsigned long long *__restrict__ n_sum,
                             double lo, double invS, double *__restrict__ n_split) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    int nd = node0 + k;
    double m = __ddiv_rn(__ull2double_rn(n_sum[nd]), (double)n_count[nd]);
    n_split[nd] = __dadd_rn(lo, __dmul_rn(m, invS));
}

// Deferred levels: the level's {node count, first node, leaves so far} live in DEVICE memory (lv[0..2], written by the
// previous level's children_top_kernel), so that the host can enqueue several levels without waiting for the counts:
// grids are sized by an upper bound (twice the previous level's) and threads beyond the count exit.
__global__ void split_top_kernel(const int *__restrict__ lv, const int *__restrict__ n_count, const unsigned long long *__restrict__ n_sum,
                                 double lo, double invS, double *__restrict__ n_split) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= lv[0]) return;
    int nd = lv[1] + k;
    double m = __ddiv_rn(__ull2double_rn(n_sum[nd]), (double)n_count[nd]);
    n_split[nd] = __dadd_rn(lo, __dmul_rn(m, invS));
}

// per node: how many leaf / node children (packed leaf | node << 32) for the numbering scan
__global__ void childcount_kernel(int cnt, int node0, const int *__restrict__ n_start, const int *__restrict__ n_count,
                                  const int *__restrict__ F, int maxleaf, unsigned long long *__restrict__ cc) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > cnt) return;
    unsigned long long v = 0;
    if (k < cnt) {
        int nd = node0 + k;
        int a = n_start[nd], c = n_count[nd];
        int R = F[a + c] - F[a];
        int c0 = c - R, c1 = R;
        unsigned nl = (c0 <= maxleaf) + (c1 <= maxleaf);
        v = (unsigned long long)nl | ((unsigned long long)(2 - nl) << 32);
    }
    cc[k] = v;
}

// per node: create the two children (records, boxes), ids from the scanned counts
__global__ void children_kernel(int cnt, int node0, int next_node0, int leaf0, int dir, int depth, int maxleaf,
                                int *__restrict__ n_start, int *__restrict__ n_count, int *__restrict__ n_son,
                                int *__restrict__ n_depth, double *__restrict__ n_box, const double *__restrict__ n_split,
                                int *__restrict__ l_start, int *__restrict__ l_count, double *__restrict__ l_box,
                                const int *__restrict__ F, const unsigned long long *__restrict__ ccs, int node_cap,
                                int leaf_cap, int *__restrict__ scal, unsigned long long *__restrict__ n_sum) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    int nd = node0 + k;
    int a = n_start[nd], c = n_count[nd];
    int R = F[a + c] - F[a];
    int cn[2] = {c - R, R};
    int st[2] = {a, a + (c - R)};
    unsigned long long off = ccs[k];
    int li = leaf0 + (int)(off & 0xffffffffULL), ni = next_node0 + (int)(off >> 32);
    double box[6];
#pragma unroll
    for (int d = 0; d < 6; d++) box[d] = n_box[6 * (size_t)nd + d];
    double sp = n_split[nd];
    for (int s = 0; s < 2; s++) {
        double cb[6];
#pragma unroll
        for (int d = 0; d < 6; d++) cb[d] = box[d];
        if (s == 0) cb[3 + dir] = sp; else cb[dir] = sp;          // center_kdtree, src/fmm.c:140-176
        if (cn[s] <= maxleaf) {
            if (li < leaf_cap) {
                l_start[li] = st[s]; l_count[li] = cn[s];
#pragma unroll
                for (int d = 0; d < 6; d++) l_box[6 * (size_t)li + d] = cb[d];
            } else atomicExch(&scal[3], 1);
            n_son[2 * (size_t)nd + s] = -(li + 2);
            li++;
        } else {
            if (ni < node_cap) {
                n_start[ni] = st[s]; n_count[ni] = cn[s]; n_depth[ni] = depth + 1; n_sum[ni] = 0ULL;
#pragma unroll
                for (int d = 0; d < 6; d++) n_box[6 * (size_t)ni + d] = cb[d];
            } else atomicExch(&scal[3], 1);
            n_son[2 * (size_t)nd + s] = ni;
            ni++;
        }
    }
    if (k == cnt - 1) {            // totals after this level
        unsigned long long tot = ccs[cnt];
        scal[0] = (int)(tot >> 32);                              // nodes of the next level
        scal[1] = leaf0 + (int)(tot & 0xffffffffULL);            // leaves so far
    }
}

__global__ void scatter_kernel(int n, const uint4 *__restrict__ pay, const int *__restrict__ seg, const int *__restrict__ F,
                               const int *__restrict__ n_start, const int *__restrict__ n_count, const int *__restrict__ n_son,
                               int next_dir, uint4 *__restrict__ pay_o, int *__restrict__ seg_o, unsigned *__restrict__ q_o) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = seg[i];
    int np = i, ns = -1;
    if (s >= 0) {
        int a = n_start[s], c = n_count[s];
        int Fa = F[a], R = F[a + c] - Fa, Fi = F[i];
        int fl = F[i + 1] - Fi, rank = Fi - Fa;
        np = fl ? a + (c - R) + rank : a + (i - a) - rank;
        int ch = n_son[2 * (size_t)s + fl];
        ns = ch >= 0 ? ch : -1;
    }
    uint4 p = pay[i];
    pay_o[np] = p;
    seg_o[np] = ns;
    q_o[np] = next_dir == 0 ? p.x : (next_dir == 1 ? p.y : p.z);
}

// ------------------------------------------------------------------------------------------------
// Top levels without moving the particles (deferred partition)
// ------------------------------------------------------------------------------------------------
// While the nodes are large the per-level stable partition is pure bookkeeping: a particle's final slot only depends on
// the node it ends up in and on its Morton rank.  The top levels therefore leave the payload where the Morton sort put
// it and only relabel: seg[i] = child slot 2 (node - node0) + side, the children's counts and coordinate sums (of the
// NEXT split direction: the next level needs no separate sum pass) accumulated per block in a small shared-memory
// table and flushed with integer atomics (exact, order-independent, as in nodesum_kernel).  24 bytes of traffic per
// particle and level instead of ~74.  A stable radix sort on the node start positions makes the nodes contiguous in
// Morton order -- exactly the state the level-by-level partitions reach -- once in the middle (so that the deep levels'
// nodes stay block-local) and once at the end.
#define TOP_ITEMS 4
#define TOP_TAB 256
__device__ __forceinline__ unsigned pay_dir(const uint4 &p, int dir) { return dir == 0 ? p.x : (dir == 1 ? p.y : p.z); }

__global__ void __launch_bounds__(TB) top_level_kernel(int n, const unsigned *__restrict__ qd, const unsigned *__restrict__ qn,
                                                       int *__restrict__ seg, const int *__restrict__ slot2id, const int *__restrict__ lv,
                                                       const int *__restrict__ n_count, const unsigned long long *__restrict__ n_sum,
                                                       unsigned long long *__restrict__ csum, unsigned *__restrict__ ccnt) {
    if (lv[0] == 0) return;                 // a level enqueued past the end of the tree: nothing to relabel
    const int node0 = lv[1];
    __shared__ int t_key[TOP_TAB];
    __shared__ unsigned long long t_sum[TOP_TAB];
    __shared__ unsigned t_cnt[TOP_TAB];
    for (int t = threadIdx.x; t < TOP_TAB; t += TB) { t_key[t] = -1; t_sum[t] = 0ULL; t_cnt[t] = 0u; }
    __syncthreads();
    auto table_add = [&](int key, unsigned long long sv, unsigned cv) {
        unsigned hh = ((unsigned)key * 2654435761u) >> 24;                  // 8 bits
        for (int probe = 0; probe < 8; probe++) {
            const int at = (int)((hh + probe) & (TOP_TAB - 1));
            const int old = atomicCAS(&t_key[at], -1, key);
            if (old == -1 || old == key) { atomicAdd(&t_sum[at], sv); atomicAdd(&t_cnt[at], cv); return; }
        }
        atomicAdd(&csum[key], sv); atomicAdd(&ccnt[key], cv);              // table full: straight to global memory
    };
    const long base = ((long)blockIdx.x * TB + threadIdx.x) * TOP_ITEMS;
    int cur = -1;
    unsigned long long cs = 0ULL;
    unsigned cc = 0u;
    // all loads of the thread's 4 items are issued before any is used: seg and the two coordinates (16-byte loads), the slot table
    int sv[TOP_ITEMS];
    unsigned dv[TOP_ITEMS], nv[TOP_ITEMS];
    const bool full = base + TOP_ITEMS <= n;
    if (full) {
        const int4 s4 = *reinterpret_cast<const int4 *>(seg + base);
        const uint4 d4 = *reinterpret_cast<const uint4 *>(qd + base), n4 = *reinterpret_cast<const uint4 *>(qn + base);
        sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w;
        dv[0] = d4.x; dv[1] = d4.y; dv[2] = d4.z; dv[3] = d4.w;
        nv[0] = n4.x; nv[1] = n4.y; nv[2] = n4.z; nv[3] = n4.w;
    } else {
#pragma unroll
        for (int k = 0; k < TOP_ITEMS; k++) {
            const bool in = base + k < n;
            sv[k] = in ? seg[base + k] : -1; dv[k] = in ? qd[base + k] : 0u; nv[k] = in ? qn[base + k] : 0u;
        }
    }
    if (slot2id) {
#pragma unroll
        for (int k = 0; k < TOP_ITEMS; k++) if (sv[k] >= 0) sv[k] = slot2id[sv[k]];       // the slot of the previous level -> node id / leaf code
    }
    int cv[TOP_ITEMS];
    unsigned long long smv[TOP_ITEMS];
#pragma unroll
    for (int k = 0; k < TOP_ITEMS; k++) {
        const bool live = sv[k] >= 0 && (k == 0 || sv[k] != sv[k - 1]);                  // consecutive particles mostly share their node
        cv[k] = live ? n_count[sv[k]] : (k > 0 ? cv[k - 1] : 0);
        smv[k] = live ? n_sum[sv[k]] : (k > 0 ? smv[k - 1] : 0ULL);
    }
#pragma unroll
    for (int k = 0; k < TOP_ITEMS; k++) {
        const int s = sv[k];
        if (s < 0) continue;                                                // in a leaf (or past the end): final
        const int fl = (cv[k] < 2) ? 1 : (((unsigned long long)dv[k] * (unsigned long long)cv[k] > smv[k]) ? 1 : 0);   // src/fmm.c:33-36, 60-72
        const int slot = 2 * (s - node0) + fl;
        sv[k] = slot;
        if (slot != cur) {
            if (cur >= 0) table_add(cur, cs, cc);
            cur = slot; cs = 0ULL; cc = 0u;
        }
        cs += nv[k];
        cc++;
    }
    if (full) *reinterpret_cast<int4 *>(seg + base) = make_int4(sv[0], sv[1], sv[2], sv[3]);
    else {
#pragma unroll
        for (int k = 0; k < TOP_ITEMS; k++) if (base + k < n) seg[base + k] = sv[k];
    }
    // the thread's last run: lanes with the same slot are merged first (the common case: the whole warp)
    const unsigned peers = __match_any_sync(0xffffffffu, cur);
    const unsigned lo = __reduce_add_sync(peers, (unsigned)(cs & 0xffffffULL));
    const unsigned hi = __reduce_add_sync(peers, (unsigned)(cs >> 24));
    const unsigned ct = __reduce_add_sync(peers, cc);
    if (cur >= 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) table_add(cur, (unsigned long long)lo + ((unsigned long long)hi << 24), ct);
    __syncthreads();
    for (int t = threadIdx.x; t < TOP_TAB; t += TB)
        if (t_key[t] >= 0) {
            atomicAdd(&csum[t_key[t]], t_sum[t]);
            atomicAdd(&ccnt[t_key[t]], t_cnt[t]);
        }
}

__global__ void childcount_top_kernel(const int *__restrict__ lv, int bound, const unsigned *__restrict__ ccnt, int maxleaf,
                                      unsigned long long *__restrict__ cc) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > bound) return;                  // the scan runs over bound + 1 entries: zeros beyond the level's count
    const int cnt = lv[0];
    unsigned long long v = 0;
    if (k < cnt) {
        unsigned nl = ((int)ccnt[2 * k] <= maxleaf) + ((int)ccnt[2 * k + 1] <= maxleaf);
        v = (unsigned long long)nl | ((unsigned long long)(2 - nl) << 32);
    }
    cc[k] = v;
}

// children_kernel for a deferred level: the counts come from ccnt, a node child starts with its coordinate sum, and
// slot2id tells the next level (and the final sort) where every slot went
__global__ void children_top_kernel(const int *__restrict__ lv, int *__restrict__ lv_next, int dir, int depth, int maxleaf,
                                    int *__restrict__ n_start, int *__restrict__ n_count, int *__restrict__ n_son,
                                    int *__restrict__ n_depth, double *__restrict__ n_box, const double *__restrict__ n_split,
                                    int *__restrict__ l_start, int *__restrict__ l_count, double *__restrict__ l_box,
                                    const unsigned long long *__restrict__ csum, const unsigned *__restrict__ ccnt,
                                    const unsigned long long *__restrict__ ccs, int node_cap, int leaf_cap, int *__restrict__ scal,
                                    unsigned long long *__restrict__ n_sum, int *__restrict__ slot2id) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = lv[0], node0 = lv[1], leaf0 = lv[2], next_node0 = node0 + cnt;
    if (cnt == 0 && k == 0) { lv_next[0] = 0; lv_next[1] = node0; lv_next[2] = leaf0; }
    if (k >= cnt) return;
    int nd = node0 + k;
    int a = n_start[nd];
    int cn[2] = {(int)ccnt[2 * k], (int)ccnt[2 * k + 1]};
    int st[2] = {a, a + cn[0]};
    unsigned long long off = ccs[k];
    int li = leaf0 + (int)(off & 0xffffffffULL), ni = next_node0 + (int)(off >> 32);
    double box[6];
#pragma unroll
    for (int d = 0; d < 6; d++) box[d] = n_box[6 * (size_t)nd + d];
    double sp = n_split[nd];
    for (int s = 0; s < 2; s++) {
        double cb[6];
#pragma unroll
        for (int d = 0; d < 6; d++) cb[d] = box[d];
        if (s == 0) cb[3 + dir] = sp; else cb[dir] = sp;          // center_kdtree, src/fmm.c:140-176
        if (cn[s] <= maxleaf) {
            if (li < leaf_cap) {
                l_start[li] = st[s]; l_count[li] = cn[s];
#pragma unroll
                for (int d = 0; d < 6; d++) l_box[6 * (size_t)li + d] = cb[d];
            } else atomicExch(&scal[3], 1);
            n_son[2 * (size_t)nd + s] = -(li + 2);
            slot2id[2 * k + s] = -(li + 2);
            li++;
        } else {
            if (ni < node_cap) {
                n_start[ni] = st[s]; n_count[ni] = cn[s]; n_depth[ni] = depth + 1; n_sum[ni] = csum[2 * k + s];
#pragma unroll
                for (int d = 0; d < 6; d++) n_box[6 * (size_t)ni + d] = cb[d];
            } else atomicExch(&scal[3], 1);
            n_son[2 * (size_t)nd + s] = ni;
            slot2id[2 * k + s] = ni;
            ni++;
        }
    }
    if (k == cnt - 1) {
        unsigned long long tot = ccs[cnt];
        lv_next[0] = (int)(tot >> 32);                            // nodes of the next level
        lv_next[1] = next_node0;
        lv_next[2] = leaf0 + (int)(tot & 0xffffffffULL);          // leaves so far
    }
}

// end of the deferred phase: the sort key of a particle = the start position of its node / leaf
__global__ void top_finish_kernel(int n, const int *__restrict__ seg, const int *__restrict__ slot2id, const int *__restrict__ n_start,
                                  const int *__restrict__ l_start, unsigned *__restrict__ key, int *__restrict__ seg_o, int *__restrict__ iota) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int s = seg[i];
    if (s >= 0 && slot2id) s = slot2id[s];
    key[i] = (unsigned)(s >= 0 ? n_start[s] : l_start[-(s + 2)]);
    seg_o[i] = s;                                  // node id, or the leaf code -(leaf + 2): needed again by the next sort
    iota[i] = i;
}
// Morton order -> coordinate arrays of the deferred builder (structure of arrays: a level reads two of the three)
__global__ void gather_soa_kernel(int n, const uint4 *__restrict__ pay_in, const int *__restrict__ idx, unsigned *__restrict__ qx,
                                  unsigned *__restrict__ qy, unsigned *__restrict__ qz, int *__restrict__ seg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 p = pay_in[idx[i]];
    qx[i] = p.x; qy[i] = p.y; qz[i] = p.z;
    seg[i] = 0;                                          // everyone starts in the root
}
// tree order: vals[j] = Morton rank of the j-th particle in tree order, idxm[] = caller index by Morton rank
__global__ void finalize_particles_soa_kernel(int n, const int *__restrict__ vals, const int *__restrict__ idxm,
                                              const double *__restrict__ pos_in, double *__restrict__ pos, int *__restrict__ order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t j = (size_t)idxm[vals[i]];
    order[i] = (int)j;
    pos[3 * (size_t)i] = pos_in[3 * j];
    pos[3 * (size_t)i + 1] = pos_in[3 * j + 1];
    pos[3 * (size_t)i + 2] = pos_in[3 * j + 2];
}

// tree-order positions and caller indices from the final payload
__global__ void finalize_particles_kernel(int n, const uint4 *__restrict__ pay, const double *__restrict__ pos_in,
                                          double *__restrict__ pos, int *__restrict__ order) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t j = (size_t)pay[i].w;
    order[i] = (int)j;
    pos[3 * (size_t)i] = pos_in[3 * j];
    pos[3 * (size_t)i + 1] = pos_in[3 * j + 1];
    pos[3 * (size_t)i + 2] = pos_in[3 * j + 2];
}

// cells: leaves 0..nleaf-1, nodes nleaf..; geometry as center_kdtree computes it (src/fmm.c:126-131)
__global__ void finalize_cells_kernel(int nleaf, int nnode, const int *__restrict__ l_start, const int *__restrict__ l_count,
                                      const double *__restrict__ l_box, const int *__restrict__ n_start,
                                      const int *__restrict__ n_count, const double *__restrict__ n_box,
                                      const int *__restrict__ n_son, const int *__restrict__ n_depth,
                                      double *__restrict__ geom, int *__restrict__ son, LeafDesc *__restrict__ desc,
                                      int *__restrict__ parent, int *__restrict__ depth, int *__restrict__ level_nodes) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nleaf + nnode) return;
    const double *b;
    int first, np;
    if (c < nleaf) {
        b = l_box + 6 * (size_t)c; first = l_start[c]; np = l_count[c];
        son[2 * (size_t)c] = -1; son[2 * (size_t)c + 1] = -1;
    } else {
        int k = c - nleaf;
        b = n_box + 6 * (size_t)k; first = n_start[k]; np = n_count[k];
        int dp = n_depth[k];
        depth[c] = dp;
        level_nodes[k] = c;
        if (k == 0) parent[c] = -1;
        for (int s = 0; s < 2; s++) {
            int ch = n_son[2 * (size_t)k + s];
            int cc = ch >= 0 ? nleaf + ch : -(ch + 2);
            son[2 * (size_t)c + s] = cc;
            parent[cc] = c;
            if (ch < 0) depth[cc] = dp + 1;
        }
    }
    LeafDesc d;
    for (int q = 0; q < 3; q++) {
        double ctr = __dmul_rn(0.5, __dadd_rn(b[3 + q], b[q]));
        geom[6 * (size_t)c + q] = ctr;
        geom[6 * (size_t)c + 3 + q] = __dsub_rn(b[3 + q], b[q]);
        d.c[q] = ctr;
    }
    d.first = first; d.npart = np;
    desc[c] = d;
}

// one attempt with room for `cap` leaves and `cap` nodes; *overflow is set when that was not enough
static int tree_build_once(pn2_ctx *h, const double *d_pos_in, int n, const pn2_domain *dom, int cap, bool *overflow) {
    cudaStream_t st = h->stream;
    *overflow = false;
    const int maxleaf = h->prm.maxleaf;
    h->n = n;
    h->dom = *dom;
    h->nleaf = h->nnode = h->ncell = h->nlevel = 0;
    h->nrl = h->nrn = h->nrp = 0;
    h->level_off.assign(1, 0);
    h->leaf_off.clear();
    PN2_TRY(h->pos.ensure(3 * (size_t)n + 3));
    PN2_TRY(h->acc.ensure(3 * (size_t)n + 3)); PN2_TRY(h->rel.ensure((size_t)n + 1));
    PN2_TRY(h->order.ensure(n + 1)); PN2_TRY(h->b_idx2.ensure(n + 1));
    PN2_TRY(h->b_seg.ensure(n + 1)); PN2_TRY(h->b_seg2.ensure(n + 1));
    PN2_TRY(h->b_q.ensure((size_t)(n > cap ? n : cap) + 2)); PN2_TRY(h->b_key2.ensure((size_t)(n > 2 * cap ? n : 2 * cap) + 2)); PN2_TRY(h->b_f.ensure((size_t)(n > 2 * cap ? n : 2 * cap) + 2)); PN2_TRY(h->b_flag.ensure((size_t)n + 16));
    PN2_TRY(h->b_pay.ensure((size_t)n + 4)); PN2_TRY(h->b_pay2.ensure((size_t)n + 1));      // b_pay also holds the three coordinate arrays
    PN2_TRY(h->b_qc.ensure((size_t)n + 1)); PN2_TRY(h->b_qc2.ensure((size_t)n + 1));
    PN2_TRY(h->b_scal.ensure(16));
    if (n == 0) return PN2_OK;
    PN2_TRY(h->n_start.ensure(cap)); PN2_TRY(h->n_count.ensure(cap)); PN2_TRY(h->n_son.ensure(2 * (size_t)cap));
    PN2_TRY(h->n_depth.ensure(cap)); PN2_TRY(h->n_box.ensure(6 * (size_t)cap)); PN2_TRY(h->n_split.ensure(cap));
    PN2_TRY(h->n_sum.ensure(cap));
    PN2_TRY(h->l_start.ensure(cap)); PN2_TRY(h->l_count.ensure(cap)); PN2_TRY(h->l_box.ensure(6 * (size_t)cap));

    // quantisation scale: extent * 2^(32-e) < 2^32
    double ext[3];
    for (int d = 0; d < 3; d++) ext[d] = dom->hi[d] - dom->lo[d];
    double emax = ext[0] > ext[1] ? ext[0] : ext[1];
    if (ext[2] > emax) emax = ext[2];
    int e2 = 0;
    frexp(emax, &e2);                    // emax = f * 2^e2, f in [0.5, 1)
    const double S = ldexp(1.0, 32 - e2), invS = ldexp(1.0, e2 - 32);

    // ---- 0. quantise + Morton pre-sort ----
    quant_morton_kernel<<<nb(n), TB, 0, st>>>(n, d_pos_in, dom->lo[0], dom->lo[1], dom->lo[2], S, h->b_pay2.p, h->b_qc.p, h->b_idx2.p);
    cub::TransformInputIterator<int, ByteToInt, const unsigned char *> flag_it(h->b_flag.p, ByteToInt());
    size_t tb = 0, tb3 = 0, tb4 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, h->b_qc.p, h->b_qc2.p, h->b_idx2.p, h->order.p, n, 0, 30, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, flag_it, h->b_f.p, n + 1, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb4, h->b_q.p, h->b_q.p, cap + 1, st);
    size_t tb5 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb5, h->b_qc.p, h->b_qc2.p, h->b_idx2.p, h->order.p, n, 0, 32, st);
    size_t need = tb;
    if (tb3 > need) need = tb3;
    if (tb4 > need) need = tb4;
    if (tb5 > need) need = tb5;
    PN2_TRY(h->tmp.ensure(need + 16));
    cub::DeviceRadixSort::SortPairs(h->tmp.p, tb, h->b_qc.p, h->b_qc2.p, h->b_idx2.p, h->order.p, n, 0, 30, st);
    const bool deferred = h->tree_top_target <= (long)n;        // PN2_TREE_TOP_TARGET > n: the level-by-level partitions instead
    unsigned *qsoa[3] = {reinterpret_cast<unsigned *>(h->b_pay.p), reinterpret_cast<unsigned *>(h->b_pay.p) + ((size_t)n + 3) / 4 * 4,
                         reinterpret_cast<unsigned *>(h->b_pay.p) + 2 * (((size_t)n + 3) / 4 * 4)};     // 16-byte aligned thirds of b_pay
    if (deferred) gather_soa_kernel<<<nb(n), TB, 0, st>>>(n, h->b_pay2.p, h->order.p, qsoa[0], qsoa[1], qsoa[2], h->b_seg.p);
    else gather_pay_kernel<<<nb(n), TB, 0, st>>>(n, h->b_pay2.p, h->order.p, dom->direct0 % 3, h->b_pay.p, h->b_qc.p, h->b_seg.p);
    h->launches += 3;

    // ---- root ----
    {
        int hs[2] = {0, n};
        double box[6] = {dom->lo[0], dom->lo[1], dom->lo[2], dom->hi[0], dom->hi[1], dom->hi[2]};
        int z = 0, m1[2] = {-1, -1};
        CUDA_TRY(cudaMemcpyAsync(h->n_start.p, &hs[0], sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(h->n_count.p, &hs[1], sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(h->n_depth.p, &z, sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(h->n_son.p, m1, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(h->n_box.p, box, 6 * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(h->b_scal.p, 0, 16 * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(h->n_sum.p, 0, sizeof(unsigned long long), st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }

    uint4 *pc = h->b_pay.p, *po = h->b_pay2.p;
    unsigned *qc = h->b_qc.p, *qo = h->b_qc2.p;
    int *sg = h->b_seg.p, *so = h->b_seg2.p;
    int node0 = 0, cnt = 1, nleaf = 0, level = 0;
    std::vector<int> level_off(1, 0), leaf_off(1, 0);

    // ---- deferred levels (see top_level_kernel): every level relabels the particles where the Morton sort put them; one
    //      stable sort on the start position of every particle's leaf then gives the tree order (Morton order inside a
    //      leaf, as the level-by-level stable partitions give it).  The caller index is all that moves.
    if (deferred) {
        nodesum_kernel<<<nb(((long)n + NS_ITEMS - 1) / NS_ITEMS), TB, 0, st>>>(n, qsoa[dom->direct0 % 3], sg, h->n_sum.p);      // the root's sum
        h->launches++;
        unsigned long long *csum = h->b_key2.p;
        unsigned *ccnt = reinterpret_cast<unsigned *>(h->b_f.p);
        int *s2i_buf[2] = {h->b_idx2.p, h->b_seg2.p};               // slot table of level l: s2i_buf[l & 1]
        // level descriptors {count, first node, leaves so far, -} in device memory: the levels are enqueued back to back
        // with grids sized by an upper bound of the count (it at most doubles per level); the host reads the counts
        // only where it has to -- every level on the first build, afterwards only over the last levels of the previous
        // step's tree (where the counts stop doubling and the tree ends)
        const int MAXLV = 208;
        PN2_TRY(h->b_lv.ensure(4 * MAXLV));
        std::vector<int> lvh(4 * MAXLV, 0);
        lvh[0] = 1;
        CUDA_TRY(cudaMemcpyAsync(h->b_lv.p, lvh.data(), 4 * sizeof(int), cudaMemcpyHostToDevice, st));
        const int sync_from = h->tree_levels_hint > 0 ? h->tree_levels_hint - 4 : 0;
        long bound = 1;
        bool ended = false;
        while (!ended) {
            if (level > 200) { pn2_set_error("pn2: tree deeper than 200 levels (more than MAXLEAF coincident particles?)"); return PN2_ERR_ARG; }
            const int dir = (dom->direct0 + level) % 3;
            const int *lv = h->b_lv.p + 4 * level;
            const int b = (int)bound;
            split_top_kernel<<<nb(b), TB, 0, st>>>(lv, h->n_count.p, h->n_sum.p, dom->lo[dir], invS, h->n_split.p);
            CUDA_TRY(cudaMemsetAsync(csum, 0, 2 * (size_t)b * sizeof(unsigned long long), st));
            CUDA_TRY(cudaMemsetAsync(ccnt, 0, 2 * (size_t)b * sizeof(unsigned), st));
            top_level_kernel<<<nb(((long)n + TOP_ITEMS - 1) / TOP_ITEMS), TB, 0, st>>>(n, qsoa[dir], qsoa[(dir + 1) % 3], sg, level > 0 ? s2i_buf[(level - 1) & 1] : nullptr, lv,
                                                                                   h->n_count.p, h->n_sum.p, csum, ccnt);
            childcount_top_kernel<<<nb(b + 1), TB, 0, st>>>(lv, b, ccnt, maxleaf, h->b_q.p);
            cub::DeviceScan::ExclusiveSum(h->tmp.p, tb4, h->b_q.p, h->b_q.p, b + 1, st);
            children_top_kernel<<<nb(b), TB, 0, st>>>(lv, h->b_lv.p + 4 * (level + 1), dir, level, maxleaf, h->n_start.p, h->n_count.p,
                                                      h->n_son.p, h->n_depth.p, h->n_box.p, h->n_split.p, h->l_start.p, h->l_count.p,
                                                      h->l_box.p, csum, ccnt, h->b_q.p, cap, cap, h->b_scal.p, h->n_sum.p, s2i_buf[level & 1]);
            h->launches += 5;
            level++;
            bound = 2 * bound > (long)cap ? (long)cap : 2 * bound;
            // (an unread level costs launches over `bound` nodes: past n / 16 nodes per level -- leaves are near -- the
            //  counts of an unbalanced tree fall far below the doubling bound, so there the host reads every level again)
            if (level >= sync_from || 2 * bound > (long)n / 16 + 1024) {
                int hs[4];
                CUDA_TRY(cudaMemcpyAsync(lvh.data(), h->b_lv.p, 4 * (size_t)(level + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaMemcpyAsync(hs, h->b_scal.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (hs[3]) { *overflow = true; return PN2_OK; }
                for (int l = 1; l <= level && !ended; l++)
                    if (lvh[4 * l] == 0) { level = l; ended = true; }          // level l has no nodes: the tree has l levels (levels enqueued past it did nothing)
                if (!ended) bound = lvh[4 * level];
            }
        }
        for (int l = 1; l <= level; l++) { level_off.push_back(lvh[4 * l + 1]); leaf_off.push_back(lvh[4 * l + 2]); }
        node0 = lvh[4 * level + 1]; nleaf = lvh[4 * level + 2]; cnt = 0;
        h->tree_levels_hint = level;
        int *s2i_prev = s2i_buf[(level - 1) & 1], *s2i = s2i_buf[level & 1];
        // tree order: sort the Morton ranks by the start position of their leaf (stable), then fetch positions / caller indices
        unsigned *key = h->b_qc.p, *key_s = h->b_qc2.p;
        int *iota = s2i, *vals = reinterpret_cast<int *>(h->b_pay2.p);      // the caller-order payload is consumed
        int bits = 1;
        while ((1L << bits) < (long)n && bits < 32) bits++;
        top_finish_kernel<<<nb(n), TB, 0, st>>>(n, sg, s2i_prev, h->n_start.p, h->l_start.p, key, reinterpret_cast<int *>(ccnt), iota);
        cub::DeviceRadixSort::SortPairs(h->tmp.p, tb5, key, key_s, iota, vals, n, 0, bits, st);
        PN2_TRY(h->order_alt.ensure(n + 1));
        finalize_particles_soa_kernel<<<nb(n), TB, 0, st>>>(n, vals, h->order.p, d_pos_in, h->pos.p, h->order_alt.p);
        h->launches += 3;
        std::swap(h->order, h->order_alt);
    }
    while (cnt > 0) {
        if (level > 200) { pn2_set_error("pn2: tree deeper than 200 levels (more than MAXLEAF coincident particles?)"); return PN2_ERR_ARG; }
        int dir = (dom->direct0 + level) % 3;
        double lo = dom->lo[dir];
        nodesum_kernel<<<nb(((long)n + NS_ITEMS - 1) / NS_ITEMS), TB, 0, st>>>(n, qc, sg, h->n_sum.p);
        split_kernel<<<nb(cnt), TB, 0, st>>>(cnt, node0, h->n_count.p, h->n_sum.p, lo, invS, h->n_split.p);
        flag_kernel<<<nb(((long)n + 4) / 4), TB, 0, st>>>(n, qc, sg, h->n_sum.p, h->n_count.p, h->b_flag.p);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tb3, flag_it, h->b_f.p, n + 1, st);
        childcount_kernel<<<nb(cnt + 1), TB, 0, st>>>(cnt, node0, h->n_start.p, h->n_count.p, h->b_f.p, maxleaf, h->b_q.p);
        cub::DeviceScan::ExclusiveSum(h->tmp.p, tb4, h->b_q.p, h->b_q.p, cnt + 1, st);
        children_kernel<<<nb(cnt), TB, 0, st>>>(cnt, node0, node0 + cnt, nleaf, dir, level, maxleaf, h->n_start.p, h->n_count.p,
                                                h->n_son.p, h->n_depth.p, h->n_box.p, h->n_split.p, h->l_start.p,
                                                h->l_count.p, h->l_box.p, h->b_f.p, h->b_q.p, cap, cap, h->b_scal.p, h->n_sum.p);
        scatter_kernel<<<nb(n), TB, 0, st>>>(n, pc, sg, h->b_f.p, h->n_start.p, h->n_count.p, h->n_son.p, (dir + 1) % 3, po, so, qo);
        h->launches += 8;
        int hs[4];
        CUDA_TRY(cudaMemcpyAsync(hs, h->b_scal.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (hs[3]) { *overflow = true; return PN2_OK; }
        std::swap(pc, po); std::swap(qc, qo); std::swap(sg, so);
        node0 += cnt;
        level_off.push_back(node0);
        cnt = hs[0];
        nleaf = hs[1];
        leaf_off.push_back(nleaf);
        level++;
    }
    if (!deferred) {
        finalize_particles_kernel<<<nb(n), TB, 0, st>>>(n, pc, d_pos_in, h->pos.p, h->order.p);
        h->launches++;
    }
    const int nnode = node0;
    if ((size_t)nleaf + nnode >= (1u << PN2_IMG_SHIFT)) { pn2_set_error("pn2: more than 2^27 cells on one device"); return PN2_ERR_ARG; }
    h->nleaf = nleaf; h->nnode = nnode; h->ncell = nleaf + nnode; h->nlevel = level;
    h->level_off = level_off;
    h->leaf_off = leaf_off;
    h->first_leaf = 0; h->last_leaf = nleaf; h->first_node = nleaf; h->last_node = nleaf + nnode - 1;
    size_t nc = (size_t)h->ncell;
    PN2_TRY(h->geom.ensure(6 * nc + 6)); PN2_TRY(h->son.ensure(2 * nc + 2)); PN2_TRY(h->desc.ensure(nc + 1));
    PN2_TRY(h->M.ensure(NM * nc + NM)); PN2_TRY(h->L.ensure(NM * nc + NM)); PN2_TRY(h->level_nodes.ensure(nnode + 1));
    PN2_TRY(h->parent.ensure(nc + 1)); PN2_TRY(h->depth.ensure(nc + 1));
    finalize_cells_kernel<<<nb(h->ncell), TB, 0, st>>>(nleaf, nnode, h->l_start.p, h->l_count.p, h->l_box.p, h->n_start.p,
                                                       h->n_count.p, h->n_box.p, h->n_son.p, h->n_depth.p, h->geom.p,
                                                       h->son.p, h->desc.p, h->parent.p, h->depth.p, h->level_nodes.p);
    h->launches++;
    PN2_TRY(h->has_l.ensure(nc + 1));
    CUDA_TRY(cudaMemsetAsync(h->has_l.p, 0, nc + 1, st));                // instead of clearing 160 bytes of L per cell
    h->use_lflags = true;
    CUDA_TRY(cudaMemsetAsync(h->acc.p, 0, 3 * (size_t)n * sizeof(double), st));
    KERNEL_CHECK();
    h->have_particles = true;
    h->have_tree = true;
    return PN2_OK;
}

// Capacity of the leaf / node records: the reference sizes NLEAF = NNODE = 2 NPART / MAXLEAF (src/fmm.c:203-204,
// unchecked); mean splits that peel off single particles (or MAXLEAF 1..2) need more, so an attempt that runs out of
// room is repeated with twice the capacity (a full binary tree over n distinct particles has < n leaves; empty leaves
// appear only where all coordinates of a node coincide in the split direction).
int pn2_tree_build_device(pn2_ctx *h, const double *d_pos_in, int n, const pn2_domain *dom) {
    long cap = 2L * n / h->prm.maxleaf + 1024;
    if (cap < (long)n / 2 + 1024) cap = (long)n / 2 + 1024;
    const long cap_max = 4L * n + 1024;
    for (;;) {
        if (cap > cap_max) cap = cap_max;
        bool overflow = false;
        PN2_TRY(tree_build_once(h, d_pos_in, n, dom, (int)cap, &overflow));
        if (!overflow) return PN2_OK;
        if (cap >= cap_max) { pn2_set_error("pn2: tree capacity %ld exceeded (degenerate particle distribution)", cap); return PN2_ERR_NOMEM; }
        cap *= 2;
    }
}
