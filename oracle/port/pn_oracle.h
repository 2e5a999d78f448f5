/*
 * pn_oracle.h -- CPU restatement ("port") of photoNs-2.0's short-range FMM path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the checker the CUDA path is compared with; it is never
 * linked into, imported by, or executed from the product (only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * PARITY PINNED: every function below is checked in tests/test_oracle_golden.py against the
 * unmodified reference compiled from /root/reference (oracle/_ref) and against fixtures in
 * tests/golden/ generated from it (tests/golden/make_golden.py).
 *
 * Index spaces follow the reference (src/fmm.c:203-212): particles 0..n-1, leaves
 * first_leaf = n .. last_leaf-1, nodes first_node = n + leafcap .. last_node (inclusive).
 */
#ifndef PN_ORACLE_H
#define PN_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define PNO_NMULTI 20

typedef struct {
    double box;      /* BOXSIZE */
    double rs;       /* splitRadius   (src/initial.c:316-317) */
    double cutoff;   /* cutoffRadius = 4.5 rs (src/initial.c:337) */
    double soft;     /* SoftenScale   (src/initial.c:318) */
    double theta;    /* open_angle */
    double mass;     /* MASSPART */
    int maxleaf;     /* MAXLEAF */
    int periodic;    /* -DPERIODIC_CONDITION */
    int longshort;   /* -DLONGSHORT */
    int pad;
} pno_params;

typedef struct pno_tree pno_tree;

/* ---- operators (src/operator.c, src/fmm.c:796-872, src/remotes.c:14-57) ---- */
void pno_p2m(const double *pos, int ipart, int npart, const double center[3], double mass, double M[PNO_NMULTI]);
void pno_m2m(double dx, double dy, double dz, const double M[PNO_NMULTI], double tM[PNO_NMULTI]);
void pno_m2l(double dx, double dy, double dz, const double M[PNO_NMULTI], double toL[PNO_NMULTI], double rs, int longshort);
void pno_l2l(double dx, double dy, double dz, const double L[PNO_NMULTI], double toL[PNO_NMULTI]);
void pno_l2p(const double *pos, int ipart, int npart, const double center[3], const double L[PNO_NMULTI], double *acc);
int  pno_acceptance(const double wi[3], const double wj[3], const double dist[3], double cutoff, double theta, int longshort);
/* one (sink range, source range) pair; skip_same_index: local pairs skip jp == ip (src/fmm.c:831) */
void pno_p2p_pair(const double *sink_pos, int sink_i0, int sink_n, const double *src_pos, int src_i0, int src_n,
                  int skip_same_index, const pno_params *prm, double *acc /* [3*nsink_total], indexed like sink_pos */);

/* ---- local tree (src/fmm.c:30-264) ---- */
pno_tree *pno_tree_build(double *pos /* [3n], permuted in place */, long *ids /* [n] or NULL, permuted alongside */,
                         int n, int maxleaf, int direct0, const double bl[3], const double br[3]);
void pno_tree_free(pno_tree *t);
void pno_tree_sizes(const pno_tree *t, int *npart, int *first_leaf, int *last_leaf, int *first_node, int *last_node);
/* copy-out accessors; k = id - first_leaf or id - first_node */
void pno_tree_get_leaves(const pno_tree *t, int *npart, int *ipart, double *center, double *width, double *M, double *L);
void pno_tree_get_nodes(const pno_tree *t, int *npart, int *son, double *split, double *center, double *width, double *M, double *L);
void pno_tree_upward(pno_tree *t, const double *pos, double mass);             /* P2M on all leaves + walk_m2m */
void pno_tree_downward(pno_tree *t, const double *pos, double *acc);           /* walk_l2l + L2P */

/* ---- interaction lists (src/fmm.c:406-712).  Returns counts; lists are malloc'ed, free with pno_free ---- */
void pno_walk_local(const pno_tree *t, const pno_params *prm, int **p2p_s, int **p2p_t, long *np2p,
                    int **m2l_s, int **m2l_t, long *nm2l);
void pno_free(void *p);
void pno_eval_p2p(const pno_tree *t, const double *pos, const pno_params *prm, const int *s, const int *tt, long n, double *acc);
void pno_eval_m2l(pno_tree *t, const pno_params *prm, const int *s, const int *tt, long n);

/* ---- LET: prune+pack, remote walk, remote evaluation (src/remotes.c) ---- */
typedef struct {
    int nnode, nbody;
    int *npart;      /* [nnode] */
    int *son;        /* [2*nnode] */
    double *width;   /* [3*nnode] */
    double *center;  /* [3*nnode] */
    double *M;       /* [20*nnode] */
    double *body;    /* [3*nbody] */
    int *origin;     /* [nnode] local id of the packed cell (checker convenience, not sent by the reference) */
} pno_let;
pno_let *pno_let_pack(const pno_tree *t, const double *pos, const pno_params *prm, const double tcenter[3],
                      const double twidth[3], const double displace[3]);
void pno_let_free(pno_let *l);
void pno_walk_remote(const pno_tree *t, const pno_let *l, const pno_params *prm, int **p2p_s, int **p2p_t, long *np2p,
                     int **m2l_s, int **m2l_t, long *nm2l);
void pno_eval_p2p_remote(const pno_tree *t, const double *pos, const pno_let *l, const pno_params *prm,
                         const int *s, const int *tt, long n, double *acc);
void pno_eval_m2l_remote(pno_tree *t, const pno_let *l, const pno_params *prm, const int *s, const int *tt, long n);

/* ---- domain geometry (src/domains.c:42-84,399-472; src/toptree.c:150-181; src/initial.c:199-223) ---- */
/* boxes of the nranks domains for the initial equal-load decomposition; center/width are [3*nranks] by RANK */
void pno_domain_boxes(int nranks, double box, double *center, double *width, int *direct_start, double *splits /* [2*nranks-1] */);
int  pno_domain_of(const double x[3], int nranks, const double *splits);  /* rank owning position x */

/* ---- Mode B tree: restatement of the DEVICE builder (photons-2.0_b200/csrc/pn2_tree.cu); returns a pno_tree in the
 *      reference id space (leaves n.., nodes n+nleaf..; breadth-first numbering) so every walker / operator above works
 *      on it.  pos_out [3n]: positions in tree order; order [n]: input index of the k-th particle in tree order ---- */
pno_tree *pno_treeB_build(const double *pos_in, int n, int maxleaf, int direct0, const double bl[3], const double br[3],
                          double *pos_out, int *order);

/* ---- whole short-range force evaluation, nranks simulated sequentially in one process
 *      (src/photoNs.c:83-116 without PM).  acc is [3n] in INPUT order.  counters[8]:
 *      0 local leaf pairs, 1 local M2L pairs, 2 local interactions, 3 remote interactions,
 *      4 m2l calls, 5 leaves, 6 nodes, 7 remote leaf pairs ---- */
void pno_force(const double *pos, int n, int nranks, const pno_params *prm, double *acc, double *counters);

#ifdef __cplusplus
}
#endif
#endif
