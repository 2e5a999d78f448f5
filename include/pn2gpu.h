/*
 * pn2gpu.h -- C-ABI of libpn2gpu.so: the B200 (sm_100a) short-range FMM force path of photoNs-2.0.
 *
 * The reference (/root/reference, photoNs-2.0) has no FFI layer; its de-facto operator API is the
 * set of C prototypes in inc/operator.h:6-19, inc/kernels.h:4-8, inc/fmm.h:11-22 and
 * inc/remotes.h:5-16, called from src/fmm.c and src/remotes.c.  Every entry point below names the
 * reference function (file:line) it replaces.  Plain pointers and sizes only; all host pointers are
 * valid for the duration of the call only (the reference re-allocates part[], leaf[], btree[] every
 * step: src/fmm.c:216-217,1080-1083, src/domains.c:331-365).
 *
 * Two ways in:
 *   Mode A ("drop-in"): the host keeps building the k-d tree and the interaction lists
 *       (src/fmm.c:81-264, 406-712); each task batch the worker pthread used to evaluate
 *       (task_compute_p2p / task_compute_m2l / *_ext) is handed to the device instead.
 *   Mode B ("device step"): positions in, accelerations out; tree, lists, periodic images, LET
 *       exchange and all operators run on the device (pn2_force_step*).
 *
 * Every function returns PN2_OK (0) or a negative pn2_status; pn2_last_error() gives the text.
 * The reference itself reports errors with printf + exit(0) (src/fmm.c:412-414); the glue decides.
 * There is no CPU fallback: without a CUDA device pn2_create fails with PN2_ERR_NODEVICE.
 */
#ifndef PN2GPU_H
#define PN2GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN2_NMULTI 20           /* inc/typesdef.h:9-12: order-3 Cartesian, 1+3+6+10 */

typedef enum {
    PN2_OK = 0,
    PN2_ERR_NODEVICE = -1,      /* no CUDA device / wrong architecture */
    PN2_ERR_CUDA = -2,          /* a CUDA runtime call or kernel failed */
    PN2_ERR_ARG = -3,           /* bad argument (null pointer, id out of range, maxleaf > 32, ...) */
    PN2_ERR_STATE = -4,         /* call order violated (e.g. batch before set_tree) */
    PN2_ERR_NOMEM = -5,
    PN2_ERR_NCCL = -6
} pn2_status;

/* arithmetic modes of the P2P kernel (the multipole operators always run in FP64) */
#define PN2_FP64 0              /* double arithmetic throughout; Mode B: g(u) from a piecewise degree-8 table (|err| < 7e-11),
                                   rsqrt seed + Newton step, no libm: rms <= 1e-6 (measured ~1e-11) */
#define PN2_FP32 1              /* leaf-centre-relative float4 sources, rsqrt + ex2 + polynomial g(u): rms <= 1e-4 */
#define PN2_FP64_LIBM 2         /* the reference's own expression (sqrt, division, erfc, exp: src/fmm.c:834-852); slow,
                                   kept as the checker of the PN2_FP64 table kernel */

/* run parameters: globals of inc/photoNs.h:20-44,124-126, derived in src/initial.c:316-345 */
typedef struct {
    double box;                 /* BOXSIZE */
    double rs;                  /* splitRadius  = 1.25 BOX/NSIDE (or SPLITSCALE) */
    double cutoff;              /* cutoffRadius = 4.5 rs */
    double soft;                /* SoftenScale */
    double theta;               /* open_angle */
    double mass;                /* MASSPART */
    int32_t maxleaf;            /* MAXLEAF (MaxPackage), 1..32 */
    int32_t periodic;           /* built with -DPERIODIC_CONDITION */
    int32_t longshort;          /* built with -DLONGSHORT */
    int32_t precision;          /* PN2_FP64 | PN2_FP32 | PN2_FP64_LIBM */
} pn2_params;

/* byte-compatible views of the reference structs (sizes probed: 376, 392, 224, 32, 96 bytes) */
typedef struct {                /* Pack, inc/typesdef.h:34-44 */
    int32_t npart, ipart;
    double width[3], center[3];
    double M[PN2_NMULTI], L[PN2_NMULTI];
} pn2_pack;
typedef struct {                /* Node, inc/typesdef.h:46-57 */
    int32_t updated, npart;
    int32_t son[2];
    double split;
    double width[3], center[3];
    double M[PN2_NMULTI], L[PN2_NMULTI];
} pn2_node;
typedef struct {                /* RemoteNode, inc/photoNs.h:177-183 */
    int32_t npart;
    int32_t son[2];
    int32_t pad_;
    double width[3], center[3];
    double M[PN2_NMULTI];
} pn2_remote_node;
typedef struct {                /* RemoteBody, inc/photoNs.h:185-189 */
    double pos[3];
    double replenish;
} pn2_remote_body;

typedef struct pn2_ctx pn2_ctx;

/* ---- life cycle ------------------------------------------------------------------------------ */
int pn2_create(pn2_ctx **out, int device, const pn2_params *prm);
int pn2_destroy(pn2_ctx *h);
int pn2_set_params(pn2_ctx *h, const pn2_params *prm);
int pn2_sync(pn2_ctx *h);                       /* join point of the worker thread, src/fmm.c:386 */
const char *pn2_last_error(void);
int pn2_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *mem_bytes);

/* ---- Mode A: state upload -------------------------------------------------------------------- */
/* part[0..n): pos points at part[0].pos, stride_bytes = sizeof(Body) = 96 (inc/typesdef.h:25-32),
 * or 24 for a packed double[n][3].  Called after build_localtree has permuted part[] (src/fmm.c:180). */
int pn2_set_particles(pn2_ctx *h, const double *pos, size_t stride_bytes, int n);
/* leaf = &leaf[first_leaf] (Pack array), btree = &btree[first_node] (Node array); ids as in
 * src/fmm.c:203-212, 258-259: leaves first_leaf..last_leaf-1, nodes first_node..last_node inclusive. */
int pn2_set_tree(pn2_ctx *h, const pn2_pack *leaf, int first_leaf, int last_leaf,
                 const pn2_node *btree, int first_node, int last_node);
/* one received LET: exrtree[0..nnode), exrbody[0..nbody)  (src/remotes.c:740-746) */
int pn2_set_remote(pn2_ctx *h, const pn2_remote_node *rtree, int nnode, const pn2_remote_body *rbody, int nbody);

/* ---- Mode A: operators ----------------------------------------------------------------------- */
/* for every leaf p2m (src/operator.c:13), then walk_m2m(first_node) (src/operator.c:165): src/fmm.c:741-744 */
int pn2_p2m_m2m(pn2_ctx *h);
/* task_compute_p2p (src/fmm.c:796-872): n ordered leaf pairs, task_s = source ids, task_t = sink ids */
int pn2_p2p_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n);
/* task_compute_m2l (src/fmm.c:875-907) */
int pn2_m2l_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n);
/* task_compute_p2p_ext -> p2p_kernel_ex (src/remotes.c:583-596, 14-57): task_s indexes the LET set by pn2_set_remote */
int pn2_p2p_ext_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n);
/* task_compute_m2l_ext (src/remotes.c:598-628) */
int pn2_m2l_ext_batch(pn2_ctx *h, const int *task_s, const int *task_t, long n);
/* walk_l2l(first_node) then l2p on every leaf (src/fmm.c:1054-1057; src/operator.c:498, 197) */
int pn2_l2l_l2p(pn2_ctx *h);

/* ---- Mode A: results ------------------------------------------------------------------------- */
/* acc points at part[0].acc, stride as in pn2_set_particles; accumulate != 0 adds (the reference's
 * "+=" into part[].acc), 0 overwrites.  Synchronises. */
int pn2_get_acc(pn2_ctx *h, double *acc, size_t stride_bytes, int n, int accumulate);
int pn2_zero_acc(pn2_ctx *h);
/* copy M / L back into the host Pack / Node arrays (for the LET pack and the top tree, which stay on
 * the host in Mode A: src/remotes.c:60-169, src/toptree.c:11-50) */
int pn2_get_multipoles(pn2_ctx *h, pn2_pack *leaf, pn2_node *btree);
int pn2_get_locals(pn2_ctx *h, pn2_pack *leaf, pn2_node *btree);
/* interaction counter: ordered particle pairs evaluated since pn2_zero_acc (idxP2P-style load metric) */
int pn2_get_counters(pn2_ctx *h, double counters[8]);

/* ---- Mode B: device-built tree, lists, images, operators --------------------------------------- */
/* geometry of this rank's domain (src/toptree.c:150-181): box corners and the first split direction
 * (direct_local_start); for one rank: lo = 0, hi = BOX, direct0 = 0 */
typedef struct {
    double lo[3], hi[3];
    int32_t direct0;
    int32_t pad_;
} pn2_domain;

/* One whole short-range force evaluation of n local particles (fmm_construct + fmm_prepare + fmm_task +
 * fmm_ext of src/photoNs.c:97-116 without PM).  pos / acc are HOST arrays in the caller's particle
 * order (acc is overwritten). */
int pn2_force_step(pn2_ctx *h, const double *pos, size_t pos_stride, int n, const pn2_domain *dom,
                   double *acc, size_t acc_stride);
/* Same with DEVICE pointers (packed double[n][3]); stays asynchronous on the context's stream. */
int pn2_force_step_device(pn2_ctx *h, const double *d_pos, int n, const pn2_domain *dom, double *d_acc);

/* multi-rank Mode B: rank/nranks, the domain boxes of all ranks ([nranks] array), and an opaque
 * ncclComm_t (passed as void* so this header needs no nccl.h).  LET trees and ghost bodies are
 * exchanged with grouped ncclSend/ncclRecv (replaces src/remotes.c:684-751). */
int pn2_set_comm(pn2_ctx *h, int rank, int nranks, const pn2_domain *all_domains, void *nccl_comm);

/* Own communicator: rank 0 calls pn2_comm_unique_id and broadcasts the 128 bytes (any host transport); every rank then
 * calls pn2_comm_init_rank (ncclCommInitRank).  The context destroys the communicator it created. */
int pn2_comm_unique_id(void *out128);
int pn2_comm_init_rank(pn2_ctx *h, int rank, int nranks, const pn2_domain *all_domains, const void *id128);

/* The two halves of pn2_force_step_device, for drivers that own the exchange: begin = tree + upward pass + LET
 * pack for every peer; finish = LET unpack + lists + P2P + M2L + downward pass.  Between them the LET blocks must be
 * exchanged: pn2_force_step_device does it with NCCL; pn2_exchange_local does it with device-to-device copies
 * when all ranks are contexts of ONE process (hs[r] = the context of rank r). */
int pn2_step_begin(pn2_ctx *h, const double *d_pos, int n, const pn2_domain *dom);
int pn2_exchange_local(pn2_ctx **hs, int nranks);
int pn2_step_finish(pn2_ctx *h, double *d_acc);

/* ---- domain decomposition on the device (src/domains.c:268-375) ------------------------------------------------
 * Particles are RECORDS of rec_doubles doubles whose first three are the position (rec_doubles = 12 is the
 * reference's Body, inc/typesdef.h:25-31: pos, acc, vel, acc_pm; 3 = bare positions).  `split` is the heap-numbered
 * domain tree of inc/photoNs.h:285-291 / src/domains.c:399-428 as 2*nranks-1 doubles (host memory; only the
 * nranks-1 internal nodes are read): node n has sons 2n+1, 2n+2, the direction cycles x, y, z from the root, a
 * position with pos[D] > split goes right (bksort_body_inplace, src/domains.c:163-266), and domain node d belongs
 * to rank (d - mostleft + nranks) % nranks (prepare_body_inOrderOf_domain, src/domains.c:268-296).
 *
 * pn2_domain_owner_device : the owner rank of every record (no data movement).
 * pn2_migrate_begin       : classifies and orders the records by destination rank; sendcount[nranks] on return.
 * pn2_migrate_exchange_nccl / pn2_migrate_exchange_local : the all-to-all-v of prepare_deliver_realloc_body
 *     (src/domains.c:298-375) with grouped ncclSend/ncclRecv over the communicator of pn2_set_comm, or with
 *     device-to-device copies between the contexts of one process.
 * pn2_migrate_result      : the received records (device pointer owned by the context, valid until the next
 *     pn2_migrate_begin) in the reference's order: blocks by source rank, ascending.
 * pn2_migrate_fetch       : copies the received records to caller memory, host or device (n_out * rec_doubles
 *     doubles), e.g. for a host that keeps its Body array in host memory like the reference (part = part_b,
 *     src/domains.c:362-363).
 * pn2_migrate_device      : begin + exchange_nccl + result (one rank: a local reorder). */
int pn2_domain_owner_device(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, int nranks, int *d_owner);
int pn2_migrate_begin(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, int *sendcount);
int pn2_migrate_exchange_nccl(pn2_ctx *h);
int pn2_migrate_exchange_local(pn2_ctx **hs, int nranks);
int pn2_migrate_result(pn2_ctx *h, double **d_rec_out, int *n_out, int *recvcount);
int pn2_migrate_fetch(pn2_ctx *h, double *rec_host_out);
int pn2_migrate_device(pn2_ctx *h, const double *d_rec, int rec_doubles, int n, const double *split, double **d_rec_out, int *n_out);

/* ---- KDK integrator on the device (src/photoNs.c:150-196, 254-268) ---------------------------------
 * Operates on device arrays of the reference's Body records (12 doubles: pos 0-2, acc 3-5, vel 6-8, acc_pm 9-11,
 * inc/typesdef.h:25-31).  The factors come from the host (kick_loga / drift_loga, src/initial.c:639-683:
 * photons-2.0_b200/cosmology.py): dkh = 0.5 * kick * GravConst, dd = drift.
 *   pn2_kick_device : vel += a1 * dkh; vel += a2 * dkh -- two separately rounded updates in the reference's order:
 *                     pm_first != 0: a1 = acc_pm, a2 = acc (opening half kick, :158-169); else a1 = acc, a2 = acc_pm
 *                     (closing half kick, :257-268).
 *   pn2_drift_device: pos += vel * dd, then wrapped into [0, box) by repeated +/- box (:171-196).
 * Results are bit-identical to the reference's loops (no FMA contraction). */
int pn2_kick_device(pn2_ctx *h, double *d_body, int n, double dkh, int pm_first);
int pn2_drift_device(pn2_ctx *h, double *d_body, int n, double dd, double box);

/* The Mode B force step on device-resident RECORDS (first three doubles = position, e.g. the reference's Body): the
 * short-range accelerations are written to doubles acc_offset .. acc_offset + 2 of every record (Body.acc: 3), so that
 * force -> kick -> drift -> migrate runs without a host copy of the particles (src/photoNs.c:97-116, 150-196).
 * Same result as pn2_force_step_device on the packed positions. */
int pn2_force_step_records(pn2_ctx *h, double *d_rec, int rec_doubles, int acc_offset, int n, const pn2_domain *dom);

/* ---- Gadget-2 snapshot blocks <-> device Body records (src/snapshot.c:211-293, 397-503) --------------
 * d_pos32 / d_vel32: the file's float32 blocks [n][3] in device memory (vel32 = v / a^1.5; may be NULL on read: vel = 0);
 * gdt2unit = a^1.5 = (1 / (1 + redshift))^1.5 (:261, :469).  Same roundings as the reference's reader and writer. */
int pn2_snapshot_to_body_device(pn2_ctx *h, const float *d_pos32, const float *d_vel32, int n, double gdt2unit, double *d_body);
int pn2_body_to_snapshot_device(pn2_ctx *h, const double *d_body, int n, double gdt2unit, float *d_pos32, float *d_vel32);

/* ---- particle-mesh long-range force on the device (src/partmesh.c:18-796, src/conv.f90:128-247) ------
 * partmesh_thread: CIC deposit of the particles on the NSIDE^3 mesh (:98-178), the mesh all-to-all into the FFT pencils
 * (:188-352) and back (:430-470), convolution (conv.f90: FFT, Green function pref exp(-k^2 rs^2) sinc^-4 / k^2, inverse
 * FFT), 4-point gradient of the potential at the 8 CIC cells and CIC gather into Body.acc_pm (:472-775).
 * Here every rank holds the whole mesh in HBM: the exchange is one ncclAllReduce of the density, the transform is cuFFT
 * (a library FFT, as 2DECOMP&FFT is for the reference), every rank gathers the force of its own particles.  BOX, rs
 * (splitRadius) and MASSPART come from pn2_params; nside is the PM mesh side NSIDE.
 *   pn2_pm_force_device : packed device positions double[n][3] -> acc_pm double[n][3] (overwritten), caller order
 *   pn2_pm_force_records: on device records (Body: rec_doubles 12, acc_pm_offset 9)
 *   pn2_pm_begin / pn2_pm_reduce_nccl | pn2_pm_reduce_local / pn2_pm_finish: the phases, for drivers that own the
 *       reduction (pn2_pm_reduce_local: all ranks are contexts of one process)
 *   pn2_pm_get_mesh     : inspection (tests): the mesh as it is -- density after begin / reduce, potential after finish
 *   pn2_pm_get_timings  : ms of the last evaluation: deposit, mesh reduction, FFTs + Green function, gather */
int pn2_pm_force_device(pn2_ctx *h, const double *d_pos, int n, int nside, double *d_acc_pm);
int pn2_pm_force_records(pn2_ctx *h, double *d_rec, int rec_doubles, int acc_pm_offset, int n, int nside);
int pn2_pm_begin(pn2_ctx *h, const double *d_pos, int n, int nside);
int pn2_pm_reduce_nccl(pn2_ctx *h);
int pn2_pm_reduce_local(pn2_ctx **hs, int nranks);
int pn2_pm_finish(pn2_ctx *h, double *d_acc_pm);
int pn2_pm_get_mesh(pn2_ctx *h, double *mesh_host);
int pn2_pm_get_timings(pn2_ctx *h, double ms[4]);

/* ---- Mode B inspection (tests: bit-exact tree / list checks; not needed by the product path) --- */
typedef struct {
    int32_t n, nleaf, nnode, nlevel;
    int64_t n_p2p_pairs;        /* leaf pairs incl. images/remote */
    int64_t n_m2l_pairs;
    int64_t n_interactions;     /* ordered particle pairs (self terms excluded) */
    int64_t n_let_nodes, n_let_bodies;
    int64_t n_walk_visits;      /* (sink, source) cell pairs examined by the list builder */
    int64_t frontier_bytes;     /* size of the per-sink-node frontier lists */
} pn2_step_info;
int pn2_get_step_info(pn2_ctx *h, pn2_step_info *info);
/* particle permutation (order[k] = caller index of the k-th particle in tree order) */
int pn2_get_order(pn2_ctx *h, int *order, int n);
/* cells: leaves 0..nleaf-1 then nodes nleaf..nleaf+nnode-1 (node nleaf = root).  Any pointer may be NULL.
 * geom[c] = {center[3], width[3]}, son[c] = {son0, son1} (cell ids, -1 none; leaves: -1,-1),
 * range[c] = {first particle, npart} */
int pn2_get_cells(pn2_ctx *h, double *geom, int *son, int *range, double *M, double *L);
/* interaction lists of the last step in CSR form by sink.  Call with NULL arrays to get the sizes.
 * p2p: sink leaves; src entry = source leaf cell | (image index << 27), image 0 = unshifted,
 * 1..26 = the reference's shift order (src/fmm.c:1028-1037).  m2l: sinks are cells, src = cell | image << 27. */
int pn2_get_lists(pn2_ctx *h, int kind /* 0 p2p, 1 m2l */, long *nseg, long *nsrc,
                  int *seg_sink, long *seg_off, int *src);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* dependent-FFMA-free issue-rate microbenchmark: returns FMA-pipe ops/s (1 FFMA = 1 op) measured with
 * CUDA events on this device; the denominator of the FMA-pipe roofline. */
int pn2_fma_peak(pn2_ctx *h, int fp64, double *ops_per_s, double *ms);
/* elapsed device time of the kernels of the last pn2_force_step*, by phase (ms): 0 tree, 1 upward,
 * 2 leaf walk + P2P (fused kernel), 3 M2L, 4 downward, 5 LET pack+exchange, 6 total, 7 frontier pass (lists by sink node) */
int pn2_get_timings(pn2_ctx *h, double ms[8]);
/* the same 8 values followed by: 8 the M2L kernel alone (3 also holds the sort of the pair list into CSR form) */
int pn2_get_timings_ex(pn2_ctx *h, double *ms, int cap);
/* CUDA-event stopwatch on the context's stream (slot 0..3): bench.py brackets its timed region with it */
int pn2_timer_start(pn2_ctx *h, int slot);
int pn2_timer_stop(pn2_ctx *h, int slot, double *ms);   /* records the stop event, synchronises, returns elapsed ms */
/* number of kernel launches issued by this context since creation */
long pn2_launch_count(pn2_ctx *h);

#ifdef __cplusplus
}
#endif
#endif
