// pn2_common.cuh -- shared declarations of libpn2gpu.so (sm_100a).  See include/pn2gpu.h for the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include <string>
#include "../../include/pn2gpu.h"

#define NM PN2_NMULTI
#define PN2_IMG_SHIFT 27                       // src entry = cell | image << 27 (27 images: 5 bits; 2^27 cells per rank incl. the LET)
#define PN2_CELL_MASK ((1u << PN2_IMG_SHIFT) - 1u)

void pn2_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            pn2_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return PN2_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)
#define PN2_TRY(expr)                                                                        \
    do {                                                                                     \
        int s_ = (expr);                                                                     \
        if (s_ != PN2_OK) return s_;                                                         \
    } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

// grow-only device buffer: the tree and the lists are rebuilt every step (src/fmm.c:216-217,1080-1083),
// the pools are not
template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n, bool keep = false, cudaStream_t st = 0) {
        if (n <= cap) return PN2_OK;
        size_t ncap = n + n / 8 + 64;
        T *q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) {
            pn2_set_error("cudaMalloc(%zu bytes): %s", ncap * sizeof(T), cudaGetErrorString(e));
            return PN2_ERR_NOMEM;
        }
        if (keep && p && cap) cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, st);
        if (p) { cudaStreamSynchronize(st); cudaFree(p); }
        p = q; cap = ncap;
        return PN2_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// packed descriptor of a leaf as a P2P source / sink: 32 bytes, one sector
struct __align__(32) LeafDesc {
    double c[3];       // box centre (displaced for LET leaves)
    int first;         // first particle in the particle array of its set
    int npart;
};

// a set of source leaves: local leaves, or the leaves of a received LET
struct SourceSet {
    const LeafDesc *desc;     // indexed by cell id (local) / flattened node id (LET)
    const float4 *rel;        // FP32 mode: (pos - leaf centre) / (2 rs), w = 1
    const double *pos;        // FP64 mode: absolute positions, double[n][3]
};

struct P2PConst {
    double inv2rs;            // 1 / (2 rs): the argument of the split factor, u = r / 2 rs (FP64 mode)
    double inv_len;           // 1 / lambda, lambda = 2 rs sqrt(ln 2): FP32 positions are stored in units of lambda (pn2_p2p.cuh)
    double rs, soft, mass;
    double shift[27][3];      // image displacements; index 0 = none, 1..26 = order of src/fmm.c:1028-1037
    float inv_eps;            // lambda / soft
    int longshort;
};

// CSR interaction list by sink
struct CsrList {
    long nseg;
    long npair = 0;           // total sources (0: unknown)
    const int *seg_sink;      // sink cell id
    const long *seg_off;      // [nseg + 1]
    const unsigned *src;      // source cell | image << 27
};

struct pn2_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    pn2_params prm{};
    P2PConst pc{};
    long launches = 0;

    // ---- particles (tree order) ----
    int n = 0;
    DBuf<double> pos;          // [n][3]
    DBuf<double> acc;          // [n][3]
    DBuf<float4> rel;          // [n]
    DBuf<float> tiles;         // Mode B FP32: leaf tiles, [nleaf + received leaves + 1][SW * 4]
    DBuf<double> tiles64;      // Mode B FP64: leaf tiles, [nleaf + received leaves + 1][SW][4]
    DBuf<double> gtab;         // FP64 mode: g(u) table (pn2_gtab.h)
    // ---- cells: leaves 0..nleaf-1, nodes nleaf..nleaf+nnode-1 ----
    int nleaf = 0, nnode = 0, ncell = 0, nlevel = 0;
    int first_leaf = 0, last_leaf = 0, first_node = 0, last_node = 0;   // Mode A id space
    DBuf<double> geom;         // [ncell][6] centre, width
    DBuf<int> son;             // [ncell][2]
    DBuf<LeafDesc> desc;       // [ncell] (first, npart valid for nodes too)
    DBuf<double> M, L;         // [ncell][20]
    DBuf<unsigned char> has_l; // Mode B: [ncell] 1 = L written this step (M2L sink or below one); L itself is never cleared
    bool use_lflags = false;
    DBuf<int> level_nodes;     // node cell ids grouped by depth
    std::vector<int> level_off;   // [nlevel + 1]
    std::vector<int> leaf_off;    // Mode B: [nlevel + 1]: the leaf sons of the nodes of level l are the leaves leaf_off[l] .. leaf_off[l + 1] - 1, in the order of their parents
    // ---- received LET (Mode A) ----
    int r_nnode = 0, r_nbody = 0;
    DBuf<LeafDesc> r_desc;
    DBuf<double> r_geom, r_M, r_pos;
    DBuf<float4> r_rel;
    // ---- scratch ----
    DBuf<int> ia, ib, ic, id_;
    DBuf<long> la;
    DBuf<unsigned> ua, ub;
    DBuf<unsigned char> tmp;
    DBuf<unsigned long long> counters;   // [8]
    double timings[8] = {0};
    bool have_particles = false, have_tree = false, have_remote = false;

    // ---- Mode B (device-built tree / lists) ----
    DBuf<int> order, order_alt;      // [n] caller index of the k-th particle in tree order (and its double buffer)
    DBuf<int> parent, depth;         // [ncell]
    DBuf<uint4> b_pay, b_pay2;       // build payload {qx, qy, qz, caller index}, ping-pong
    DBuf<unsigned> b_qc, b_qc2;      // the current level's key
    DBuf<unsigned long long> n_sum;  // per-node coordinate sums
    DBuf<int> b_idx2, b_seg, b_seg2;
    DBuf<unsigned long long> b_q;    // quantised coordinates / their scan, morton keys
    DBuf<unsigned long long> b_key2;
    DBuf<int> b_f;                   // scan of the flags
    DBuf<unsigned char> b_flag;      // flags
    DBuf<int> n_start, n_count, n_son, n_depth, l_start, l_count;   // build-time node / leaf records
    DBuf<double> n_box, n_split, l_box;                              // [cap][6] lo, hi
    DBuf<unsigned long long> b_cnt;  // per-level child counts (leaf | node << 32) and scan
    DBuf<int> b_scal;                // device scalars
    DBuf<int> b_lv;                  // per-level {node count, first node, leaves so far, -} of the deferred tree levels
    int tree_levels_hint = 0;        // levels of the previous step's tree: where the builder starts reading the counts back
    DBuf<unsigned> spans;            // O(im) span lists of the frontier pass (16-byte units)
    unsigned long long span_cap16 = 0, span_used16 = 0, walk_visits = 0;
    DBuf<unsigned> o_head;           // [ncell]
    DBuf<unsigned> m2l_pairs;        // [cap][2] (sink cell, src cell | image << 27), appended by the walk
    size_t m2l_cap = 0;
    DBuf<long> lst_off;              // dump mode: per-leaf offsets
    DBuf<unsigned> lst_src;
    DBuf<int> lst_sink;
    long lst_nsrc = 0;
    CsrList m2l_csr{};               // CSR view of the last step's M2L list (buffers ia/ic/la/ub)
    long tree_top_target = 1;        // PN2_TREE_TOP_TARGET > n selects the level-by-level partition builder (the checked alternative)
    pn2_domain dom{};
    pn2_step_info info{};
    bool have_step = false;
    // multi-rank: received LET cells are appended to the cell arrays (leaves then nodes), ghost particles to rel / pos
    int nrl = 0, nrn = 0, nrp = 0;
    unsigned root_head = 0;
    int rank = 0, nranks = 1;
    bool own_comm = false;
    struct LetState *let = nullptr;
    struct MigState *mig = nullptr;     // domain decomposition (pn2_migrate.cu)
    struct PmState *pm = nullptr;       // particle-mesh long-range force (pn2_pm.cu)
    DBuf<double> rec_pos, rec_acc;      // packed positions / accelerations of pn2_force_step_records
    DBuf<double> stage_in, stage_out;   // device staging of pn2_force_step's host positions / accelerations
    std::vector<pn2_domain> all_dom;
    void *nccl = nullptr;
    cudaEvent_t ev[16] = {nullptr};
    unsigned long long top0_host = 0;   // staging of the span bump pointer's start value
    std::vector<unsigned> root_span_host;   // F(root) of the current walk pass
    std::vector<int> peer_roots;            // root cell of every received LET (after pn2_let_unpack)
    int root_count = 0;
    bool let_unpacked = false;
    bool walk_active = false;               // the current walk pass visits the sink tree through active lists (pass 1)
    DBuf<int> act_nodes, act_leaf;
    DBuf<unsigned> act_count;
    long step_serial = 0, tiles_built_for = -1;   // the local leaf tiles are built once per step (first walk pass)
    bool step_open = false;
    unsigned root_units = 0;
    cudaEvent_t tev[4][2] = {{nullptr}};
};

// ---- kernels / launchers implemented across the .cu files ----
int pn2_launch_p2p(pn2_ctx *h, const CsrList &list, const SourceSet &src, bool src_is_local);
int pn2_launch_m2l(pn2_ctx *h, const CsrList &list, const double *src_geom, const double *src_M);
int pn2_launch_p2m(pn2_ctx *h);
int pn2_launch_m2m(pn2_ctx *h);
int pn2_launch_l2l_l2p(pn2_ctx *h);
int pn2_launch_relpos(pn2_ctx *h, const double *pos, const LeafDesc *desc, int ncell_leaf, float4 *rel, int n);
int pn2_build_csr(pn2_ctx *h, const int *h_s, const int *h_t, long n, int remote_src, int sinks_may_be_nodes, CsrList *out);
void pn2_modeb_release(pn2_ctx *h);
int pn2_csr_from_device_pairs(pn2_ctx *h, int *tcell, unsigned *scell, long n, CsrList *out);
int pn2_tree_build_device(pn2_ctx *h, const double *d_pos_in, int n, const pn2_domain *dom);
int pn2_walk_fused(pn2_ctx *h, int dump);
int pn2_walk_frontiers(pn2_ctx *h);
int pn2_let_pack_all(pn2_ctx *h);
int pn2_let_exchange_nccl(pn2_ctx *h);
int pn2_let_unpack(pn2_ctx *h);
int pn2_walk_set_roots(pn2_ctx *h, int which);
int pn2_let_tree_ready(pn2_ctx *h);
void pn2_let_release(pn2_ctx *h);
void pn2_migrate_release(pn2_ctx *h);
void pn2_pm_release(pn2_ctx *h);
void pn2_comm_release(pn2_ctx *h);
int pn2_step_begin(pn2_ctx *h, const double *d_pos, int n, const pn2_domain *dom);
int pn2_step_finish(pn2_ctx *h, double *d_acc);
void pn2_init_consts(pn2_ctx *h);
