/*
 * pn_oracle.c -- CPU restatement ("port") of photoNs-2.0's short-range FMM path.
 *
 * TEST INFRASTRUCTURE ONLY (see pn_oracle.h).  PARITY PINNED against the unmodified reference
 * (oracle/_ref) in tests/test_oracle_golden.py and against tests/golden/.
 *
 * Every function cites the reference lines it restates.  Geometry, tree construction and the
 * acceptance test reproduce the reference's floating-point expression ORDER (they decide tree
 * shape and list membership, which must be bit-exact); the multipole operators are written
 * from the operator maths with multi-index tables and agree with the
 * reference to rounding.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, like the reference's plain -O2 build).
 */
#include "pn_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NM PNO_NMULTI

/* ------------------------------------------------------------------------------------------
 * multi-index tables: storage order of inc/operator.h:24-67
 *   0:000 | 1:X 2:Y 3:Z | 4:XX 5:XY 6:XZ 7:YY 8:YZ 9:ZZ |
 *   10:XXX 11:XXY 12:XXZ 13:XYY 14:XYZ 15:XZZ 16:YYY 17:YYZ 18:YZZ 19:ZZZ
 * ------------------------------------------------------------------------------------------ */
static const int MI[NM][3] = {
    {0,0,0}, {1,0,0},{0,1,0},{0,0,1},
    {2,0,0},{1,1,0},{1,0,1},{0,2,0},{0,1,1},{0,0,2},
    {3,0,0},{2,1,0},{2,0,1},{1,2,0},{1,1,1},{1,0,2},{0,3,0},{0,2,1},{0,1,2},{0,0,3}};
static int MIDX[4][4][4];
static int tables_ready = 0;
static const double FACT[4] = {1.0, 1.0, 2.0, 6.0};

static void init_tables(void) {
    if (tables_ready) return;
    memset(MIDX, -1, sizeof MIDX);
    for (int i = 0; i < NM; i++) MIDX[MI[i][0]][MI[i][1]][MI[i][2]] = i;
    tables_ready = 1;
}
static inline int ord(int i) { return MI[i][0] + MI[i][1] + MI[i][2]; }
static inline double ipow(double x, int k) { double r = 1.0; while (k-- > 0) r *= x; return r; }

/* ------------------------------------------------------------------------------------------
 * P2M  (src/operator.c:13-93):  M_n = sum_p m (-d_p)^n / n!,  d_p = x_p - c
 * ------------------------------------------------------------------------------------------ */
void pno_p2m(const double *pos, int ipart, int npart, const double center[3], double mass, double M[NM]) {
    init_tables();
    for (int i = 0; i < NM; i++) M[i] = 0.0;
    for (int p = ipart; p < ipart + npart; p++) {
        double d[3] = {pos[3*p] - center[0], pos[3*p+1] - center[1], pos[3*p+2] - center[2]};
        for (int i = 0; i < NM; i++) {
            int o = ord(i);
            double sgn = (o & 1) ? -mass : mass;
            double t = sgn * ipow(d[0], MI[i][0]) * ipow(d[1], MI[i][1]) * ipow(d[2], MI[i][2]);
            M[i] += t / (FACT[MI[i][0]] * FACT[MI[i][1]] * FACT[MI[i][2]]);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * M2M  (src/operator.c:96-160):  M'_n += sum_{k<=n} M_{n-k} s^k / k!
 * ------------------------------------------------------------------------------------------ */
void pno_m2m(double dx, double dy, double dz, const double M[NM], double tM[NM]) {
    init_tables();
    double s[3] = {dx, dy, dz};
    for (int i = 0; i < NM; i++) {
        double acc = 0.0;
        for (int kx = 0; kx <= MI[i][0]; kx++)
            for (int ky = 0; ky <= MI[i][1]; ky++)
                for (int kz = 0; kz <= MI[i][2]; kz++) {
                    int j = MIDX[MI[i][0]-kx][MI[i][1]-ky][MI[i][2]-kz];
                    acc += M[j] * ipow(s[0], kx) * ipow(s[1], ky) * ipow(s[2], kz) / (FACT[kx] * FACT[ky] * FACT[kz]);
                }
        tM[i] += acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * M2L  (src/operator.c:255-392):  L_n += sum_{|m|<=3-|n|} M_m D_{n+m}(R),  R = c_sink - c_source
 *   f_k = ((1/r) d/dr)^k G(r),  G = erfc(r/2rs)/r with LONGSHORT (:294-307), else 1/r (:288-292)
 *   D_0 = f0, D_i = f1 R_i, D_ij = f2 R_i R_j + f1 d_ij,
 *   D_ijk = f3 R_i R_j R_k + f2 (d_ij R_k + d_ik R_j + d_jk R_i)
 * ------------------------------------------------------------------------------------------ */
void pno_m2l(double dx, double dy, double dz, const double M[NM], double toL[NM], double rs, int longshort) {
    init_tables();
    double R[3] = {dx, dy, dz};
    double r2 = dx*dx + dy*dy + dz*dz;
    double r = sqrt(r2);
    double ir = 1.0 / r, ir2 = ir*ir, ir3 = ir*ir2, ir4 = ir3*ir, ir5 = ir2*ir3, ir6 = ir2*ir4, ir7 = ir5*ir2;
    double f[4];
    f[0] = ir; f[1] = -ir3; f[2] = 3.0*ir5; f[3] = -15.0*ir7;
    if (longshort) {
        double irs = 1.0 / rs, irs2 = irs*irs, irs3 = irs2*irs, irs5 = irs3*irs2;
        double u = 0.5 * r / rs;
        double X = exp(-u*u) * (1.0 / sqrt(M_PI));
        double E = erfc(u);
        f[0] = ir * E;
        f[1] = -ir3 * (E + r * X * irs);
        f[2] = 3.0*ir5*E + (3.0*irs*ir4 + 0.5*ir2*irs3) * X;
        f[3] = -15.0*ir7*E - (15.0*ir6*irs + 2.5*ir4*irs3 + 0.25*ir2*irs5) * X;
    }
    /* derivative tensor for every stored multi-index */
    double D[NM];
    for (int i = 0; i < NM; i++) {
        int a[3] = {MI[i][0], MI[i][1], MI[i][2]};
        int o = a[0] + a[1] + a[2];
        double rr = ipow(R[0], a[0]) * ipow(R[1], a[1]) * ipow(R[2], a[2]);
        if (o == 0) D[i] = f[0];
        else if (o == 1) D[i] = f[1] * rr;
        else if (o == 2) {
            D[i] = f[2] * rr;
            if (a[0] == 2 || a[1] == 2 || a[2] == 2) D[i] += f[1];
        } else {
            /* third order: f3 RRR + f2 * (number of ways to pair two equal indices) * remaining R */
            D[i] = f[3] * rr;
            for (int d = 0; d < 3; d++) {
                if (a[d] == 3) D[i] += f[2] * R[d] * 3.0;
                else if (a[d] == 2) { for (int e = 0; e < 3; e++) if (e != d && a[e] == 1) D[i] += f[2] * R[e]; }
            }
        }
    }
    for (int n = 0; n < NM; n++) {
        int on = ord(n);
        double acc = 0.0;
        for (int m = 0; m < NM; m++) {
            if (ord(m) + on > 3) continue;
            int j = MIDX[MI[n][0]+MI[m][0]][MI[n][1]+MI[m][1]][MI[n][2]+MI[m][2]];
            acc += M[m] * D[j];
        }
        toL[n] += acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * L2L  (src/operator.c:395-494):  L'_n += sum_{|k|<=3-|n|} L_{n+k} s^k / k!,  s = c_child - c_parent
 * ------------------------------------------------------------------------------------------ */
void pno_l2l(double dx, double dy, double dz, const double L[NM], double toL[NM]) {
    init_tables();
    double s[3] = {dx, dy, dz};
    for (int n = 0; n < NM; n++) {
        int on = ord(n);
        double acc = 0.0;
        for (int k = 0; k < NM; k++) {
            if (ord(k) + on > 3) continue;
            int j = MIDX[MI[n][0]+MI[k][0]][MI[n][1]+MI[k][1]][MI[n][2]+MI[k][2]];
            acc += L[j] * ipow(s[0], MI[k][0]) * ipow(s[1], MI[k][1]) * ipow(s[2], MI[k][2])
                   / (FACT[MI[k][0]] * FACT[MI[k][1]] * FACT[MI[k][2]]);
        }
        toL[n] += acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * L2P  (src/operator.c:197-251):  a_i += sum_{|n|<=2} L_{n+e_i} d^n / n!,  d = x_p - c_leaf
 * (the potential is computed but never stored by the reference: :249 is commented out)
 * ------------------------------------------------------------------------------------------ */
void pno_l2p(const double *pos, int ipart, int npart, const double center[3], const double L[NM], double *acc) {
    init_tables();
    for (int p = ipart; p < ipart + npart; p++) {
        double d[3] = {pos[3*p] - center[0], pos[3*p+1] - center[1], pos[3*p+2] - center[2]};
        for (int c = 0; c < 3; c++) {
            double F = 0.0;
            for (int n = 0; n < 10; n++) { /* |n| <= 2 */
                int a[3] = {MI[n][0], MI[n][1], MI[n][2]};
                double t = ipow(d[0], a[0]) * ipow(d[1], a[1]) * ipow(d[2], a[2]) / (FACT[a[0]] * FACT[a[1]] * FACT[a[2]]);
                a[c] += 1;
                F += L[MIDX[a[0]][a[1]][a[2]]] * t;
            }
            acc[3*p + c] += F;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * acceptance  (src/fmm.c:267-326).  0 open, 1 accept (M2L), -1 drop.  Expression order kept.
 * ------------------------------------------------------------------------------------------ */
int pno_acceptance(const double wi[3], const double wj[3], const double dist[3], double cutoff, double theta, int longshort) {
    double w[3], gap[3];
    for (int d = 0; d < 3; d++) w[d] = (wi[d] + wj[d]) * 0.5;
    double dd2 = dist[0]*dist[0] + dist[1]*dist[1] + dist[2]*dist[2];
    for (int d = 0; d < 3; d++) {
        double a = dist[d];
        if (a < 0.0) a = -a;
        a -= w[d];
        if (a <= 0.0) a = 0.0;
        gap[d] = a;
    }
    if (gap[0] + gap[1] + gap[2] < 0.0001) return 0;
    double dm2 = gap[0]*gap[0] + gap[1]*gap[1] + gap[2]*gap[2];
    if (longshort) {
        double c2 = cutoff * cutoff;
        if (dm2 >= c2) return -1;
        if (dd2 > 1.0 * c2) return 0;
    }
    double wmax = w[0];
    if (w[1] > wmax) wmax = w[1];
    if (w[2] > wmax) wmax = w[2];
    wmax *= 2;
    return (wmax * wmax < theta * theta * dd2) ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------
 * P2P pair arithmetic  (src/fmm.c:823-855 local, src/remotes.c:26-56 remote)
 * ------------------------------------------------------------------------------------------ */
void pno_p2p_pair(const double *sink_pos, int sink_i0, int sink_n, const double *src_pos, int src_i0, int src_n,
                  int skip_same_index, const pno_params *prm, double *acc) {
    const double coeff = 2.0 / sqrt(M_PI);
    const double eps = prm->soft;
    for (int ip = sink_i0; ip < sink_i0 + sink_n; ip++) {
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int jp = src_i0; jp < src_i0 + src_n; jp++) {
            if (skip_same_index && jp == ip) continue;
            double dx = src_pos[3*jp] - sink_pos[3*ip];
            double dy = src_pos[3*jp+1] - sink_pos[3*ip+1];
            double dz = src_pos[3*jp+2] - sink_pos[3*ip+2];
            double dr = sqrt(dx*dx + dy*dy + dz*dz);
            double ir3 = (dr < eps) ? prm->mass / (eps*eps*eps) : prm->mass / (dr*dr*dr);
            if (prm->longshort) {
                double u = 0.5 * dr / prm->rs;
                ir3 *= (erfc(u) + coeff * u * exp(-u*u));
            }
            ax += dx * ir3; ay += dy * ir3; az += dz * ir3;
        }
        acc[3*ip] += ax; acc[3*ip+1] += ay; acc[3*ip+2] += az;
    }
}

/* ------------------------------------------------------------------------------------------
 * local k-d tree  (src/fmm.c:30-264)
 * ------------------------------------------------------------------------------------------ */
struct pno_tree {
    int n, maxleaf;
    int first_leaf, last_leaf, first_node, last_node;
    int leafcap, nodecap;
    /* leaves, indexed by id - first_leaf */
    int *lf_npart, *lf_ipart;
    double *lf_center, *lf_width, *lf_M, *lf_L;
    /* nodes, indexed by id - first_node */
    int *nd_npart, *nd_son;
    double *nd_split, *nd_center, *nd_width, *nd_M, *nd_L;
};

static inline int is_leaf(const pno_tree *t, int id) { return id < t->first_node; }
static inline const double *Wd(const pno_tree *t, int id) {
    return is_leaf(t, id) ? t->lf_width + 3*(id - t->first_leaf) : t->nd_width + 3*(id - t->first_node);
}
static inline const double *Cn(const pno_tree *t, int id) {
    return is_leaf(t, id) ? t->lf_center + 3*(id - t->first_leaf) : t->nd_center + 3*(id - t->first_node);
}
static inline double *Mp(const pno_tree *t, int id) {
    return is_leaf(t, id) ? t->lf_M + NM*(id - t->first_leaf) : t->nd_M + NM*(id - t->first_node);
}
static inline double *Lp(const pno_tree *t, int id) {
    return is_leaf(t, id) ? t->lf_L + NM*(id - t->first_leaf) : t->nd_L + NM*(id - t->first_node);
}
static inline int son(const pno_tree *t, int id, int k) { return t->nd_son[2*(id - t->first_node) + k]; }

static inline void swap_body(double *pos, long *ids, int a, int b) {
    for (int d = 0; d < 3; d++) { double x = pos[3*a+d]; pos[3*a+d] = pos[3*b+d]; pos[3*b+d] = x; }
    if (ids) { long x = ids[a]; ids[a] = ids[b]; ids[b] = x; }
}

/* mean split + in-place partition, src/fmm.c:30-78.  Returns the size of the low side. */
static int mean_partition(int D, double *pos, long *ids, int i0, int len, double *split) {
    if (len < 2) return 0;                         /* :33-36  -> npart = {0, len} */
    if (len == 2) {                                /* :38-48 */
        *split = 0.5 * (pos[3*i0 + D] + pos[3*(i0+1) + D]);
        if (pos[3*i0 + D] > pos[3*(i0+1) + D]) swap_body(pos, ids, i0, i0 + 1);
        return 1;
    }
    double mean = 0.0;
    for (int k = 0; k < len; k++) mean += pos[3*(i0+k) + D];   /* sequential sum, :53-56 */
    mean /= (double)len;
    int hi = len - 1;
    for (int k = 0; k < hi; k++) {                 /* :60-72 */
        if (pos[3*(i0+k) + D] > mean) {
            while (pos[3*(i0+hi) + D] > mean && hi > k) hi--;
            swap_body(pos, ids, i0 + k, i0 + hi);
        }
    }
    *split = mean;
    return hi;
}

static void build_rec(pno_tree *t, double *pos, long *ids, int D, int i0, int len, int node) {
    if (len == 0) return;                          /* src/fmm.c:83-84 */
    int k = node - t->first_node;
    if (k >= t->nodecap) { fprintf(stderr, "pn_oracle: node capacity exceeded\n"); exit(3); }
    t->nd_npart[k] = len;
    double split = 0.0;
    int nlo = mean_partition(D, pos, ids, i0, len, &split);
    int cnt[2] = {nlo, len - nlo};
    t->nd_split[k] = split;
    int ip = i0, nd = (D + 1) % 3;
    for (int s = 0; s < 2; s++) {                  /* :101-118 */
        if (cnt[s] <= t->maxleaf) {
            int lk = t->last_leaf - t->first_leaf;
            if (lk >= t->leafcap) { fprintf(stderr, "pn_oracle: leaf capacity exceeded\n"); exit(3); }
            t->lf_npart[lk] = cnt[s];
            t->lf_ipart[lk] = ip;
            t->nd_son[2*k + s] = t->last_leaf++;
        } else {
            int child = ++t->last_node;
            t->nd_son[2*k + s] = child;
            build_rec(t, pos, ids, nd, ip, cnt[s], child);
        }
        ip += cnt[s];
    }
}

/* boxes from splits, src/fmm.c:123-177.  left/right are modified and restored like the reference. */
static void center_rec(pno_tree *t, int D, int node, double left[3], double right[3]) {
    int k = node - t->first_node;
    for (int d = 0; d < 3; d++) {
        t->nd_width[3*k+d] = right[d] - left[d];
        t->nd_center[3*k+d] = 0.5 * (right[d] + left[d]);
    }
    int nd = (D + 1) % 3;
    for (int s = 0; s < 2; s++) {
        int c = t->nd_son[2*k + s];
        if (c < t->last_leaf) {                    /* leaf (also catches -1 like the reference's test) */
            if (c < t->first_leaf) continue;
            int lk = c - t->first_leaf;
            for (int d = 0; d < 3; d++) {
                t->lf_width[3*lk+d] = t->nd_width[3*k+d];
                t->lf_center[3*lk+d] = t->nd_center[3*k+d];
            }
            if (s == 0) {
                t->lf_width[3*lk+D] = t->nd_split[k] - left[D];
                t->lf_center[3*lk+D] = 0.5 * (left[D] + t->nd_split[k]);
            } else {
                t->lf_width[3*lk+D] = right[D] - t->nd_split[k];
                t->lf_center[3*lk+D] = 0.5 * (right[D] + t->nd_split[k]);
            }
        } else if (s == 0) {
            double tmp = right[D];
            right[D] = t->nd_split[k];
            center_rec(t, nd, c, left, right);
            right[D] = tmp;
        } else {
            double tmp = left[D];
            left[D] = t->nd_split[k];
            center_rec(t, nd, c, left, right);
            left[D] = tmp;
        }
    }
}

pno_tree *pno_tree_build(double *pos, long *ids, int n, int maxleaf, int direct0, const double bl[3], const double br[3]) {
    init_tables();
    pno_tree *t = (pno_tree *)calloc(1, sizeof *t);
    t->n = n; t->maxleaf = maxleaf;
    /* capacities and index spaces, src/fmm.c:203-212 */
    int cap = (int)(2.0 * ((double)n) / ((double)maxleaf));
    t->nodecap = cap; t->leafcap = cap;
    if (t->nodecap > n) t->nodecap = n + 1;
    if (t->leafcap > n) t->leafcap = n + 1;
    t->first_leaf = t->last_leaf = n;
    t->first_node = t->last_node = n + t->leafcap;
    int lc = t->leafcap > 0 ? t->leafcap : 1, nc = t->nodecap > 0 ? t->nodecap : 1;
    t->lf_npart = (int *)calloc(lc, sizeof(int)); t->lf_ipart = (int *)calloc(lc, sizeof(int));
    t->lf_center = (double *)calloc(3*lc, sizeof(double)); t->lf_width = (double *)calloc(3*lc, sizeof(double));
    t->lf_M = (double *)calloc(NM*lc, sizeof(double)); t->lf_L = (double *)calloc(NM*lc, sizeof(double));
    t->nd_npart = (int *)calloc(nc, sizeof(int)); t->nd_son = (int *)malloc(2*nc*sizeof(int));
    for (int i = 0; i < 2*nc; i++) t->nd_son[i] = -1;
    t->nd_split = (double *)calloc(nc, sizeof(double));
    t->nd_center = (double *)calloc(3*nc, sizeof(double)); t->nd_width = (double *)calloc(3*nc, sizeof(double));
    t->nd_M = (double *)calloc(NM*nc, sizeof(double)); t->nd_L = (double *)calloc(NM*nc, sizeof(double));
    build_rec(t, pos, ids, direct0, 0, n, t->first_node);
    double l[3] = {bl[0], bl[1], bl[2]}, r[3] = {br[0], br[1], br[2]};
    center_rec(t, direct0, t->first_node, l, r);
    return t;
}

void pno_tree_free(pno_tree *t) {
    if (!t) return;
    free(t->lf_npart); free(t->lf_ipart); free(t->lf_center); free(t->lf_width); free(t->lf_M); free(t->lf_L);
    free(t->nd_npart); free(t->nd_son); free(t->nd_split); free(t->nd_center); free(t->nd_width); free(t->nd_M); free(t->nd_L);
    free(t);
}

void pno_tree_sizes(const pno_tree *t, int *npart, int *first_leaf, int *last_leaf, int *first_node, int *last_node) {
    *npart = t->n; *first_leaf = t->first_leaf; *last_leaf = t->last_leaf; *first_node = t->first_node; *last_node = t->last_node;
}

void pno_tree_get_leaves(const pno_tree *t, int *npart, int *ipart, double *center, double *width, double *M, double *L) {
    int nl = t->last_leaf - t->first_leaf;
    if (npart) memcpy(npart, t->lf_npart, nl*sizeof(int));
    if (ipart) memcpy(ipart, t->lf_ipart, nl*sizeof(int));
    if (center) memcpy(center, t->lf_center, 3*nl*sizeof(double));
    if (width) memcpy(width, t->lf_width, 3*nl*sizeof(double));
    if (M) memcpy(M, t->lf_M, NM*nl*sizeof(double));
    if (L) memcpy(L, t->lf_L, NM*nl*sizeof(double));
}

void pno_tree_get_nodes(const pno_tree *t, int *npart, int *sons, double *split, double *center, double *width, double *M, double *L) {
    int nn = t->last_node - t->first_node + 1;
    if (npart) memcpy(npart, t->nd_npart, nn*sizeof(int));
    if (sons) memcpy(sons, t->nd_son, 2*nn*sizeof(int));
    if (split) memcpy(split, t->nd_split, nn*sizeof(double));
    if (center) memcpy(center, t->nd_center, 3*nn*sizeof(double));
    if (width) memcpy(width, t->nd_width, 3*nn*sizeof(double));
    if (M) memcpy(M, t->nd_M, NM*nn*sizeof(double));
    if (L) memcpy(L, t->nd_L, NM*nn*sizeof(double));
}

/* post-order M2M, src/operator.c:165-194 */
static void upward_rec(pno_tree *t, int node) {
    const double *c = Cn(t, node);
    for (int s = 0; s < 2; s++) {
        int ch = son(t, node, s);
        if (ch < 0) continue;
        if (!is_leaf(t, ch)) upward_rec(t, ch);
        const double *cc = Cn(t, ch);
        pno_m2m(c[0]-cc[0], c[1]-cc[1], c[2]-cc[2], Mp(t, ch), Mp(t, node));
    }
}

void pno_tree_upward(pno_tree *t, const double *pos, double mass) {
    int nl = t->last_leaf - t->first_leaf;
    for (int k = 0; k < nl; k++)                   /* src/fmm.c:741-742 */
        pno_p2m(pos, t->lf_ipart[k], t->lf_npart[k], t->lf_center + 3*k, mass, t->lf_M + NM*k);
    int nn = t->last_node - t->first_node + 1;
    memset(t->nd_M, 0, sizeof(double)*NM*nn);
    if (t->n > 0) upward_rec(t, t->first_node);    /* src/fmm.c:744 */
}

/* pre-order L2L, src/operator.c:498-528 */
static void downward_rec(pno_tree *t, int node) {
    if (is_leaf(t, node)) return;
    const double *c = Cn(t, node);
    for (int s = 0; s < 2; s++) {
        int ch = son(t, node, s);
        if (ch < t->first_leaf) return;            /* :524-525 */
        const double *cc = Cn(t, ch);
        pno_l2l(cc[0]-c[0], cc[1]-c[1], cc[2]-c[2], Lp(t, node), Lp(t, ch));
        if (!is_leaf(t, ch)) downward_rec(t, ch);
    }
}

void pno_tree_downward(pno_tree *t, const double *pos, double *acc) {
    if (t->n > 0) downward_rec(t, t->first_node);  /* src/fmm.c:1054 */
    int nl = t->last_leaf - t->first_leaf;
    for (int k = 0; k < nl; k++)                   /* src/fmm.c:1056-1057 */
        pno_l2p(pos, t->lf_ipart[k], t->lf_npart[k], t->lf_center + 3*k, t->lf_L + NM*k, acc);
}

/* ------------------------------------------------------------------------------------------
 * pair lists
 * ------------------------------------------------------------------------------------------ */
typedef struct { int *s, *t; long n, cap; } plist;
static void pl_push(plist *l, int s, int t) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2*l->cap : 1 << 16;
        l->s = (int *)realloc(l->s, l->cap*sizeof(int));
        l->t = (int *)realloc(l->t, l->cap*sizeof(int));
    }
    l->s[l->n] = s; l->t[l->n] = t; l->n++;
}

/* dual-tree walk, src/fmm.c:406-538 (P2P pass) and :569-712 (M2L pass).  The reference walks twice
 * with identical decisions; one traversal emitting into two lists gives the same two sequences. */
static void walk_rec(const pno_tree *t, const pno_params *prm, int im, int jm, plist *p2p, plist *m2l) {
    if (im == -1 || jm == -1) return;
    if (im == jm) {
        if (is_leaf(t, im)) { pl_push(p2p, jm, im); return; }
        for (int a = 0; a < 2; a++)
            for (int b = 0; b < 2; b++) walk_rec(t, prm, son(t, im, a), son(t, jm, b), p2p, m2l);
        return;
    }
    int li = is_leaf(t, im), lj = is_leaf(t, jm);
    if (li && lj) { pl_push(p2p, jm, im); return; }
    const double *ci = Cn(t, im), *cj = Cn(t, jm);
    double dist[3] = {ci[0]-cj[0], ci[1]-cj[1], ci[2]-cj[2]};
    int f = pno_acceptance(Wd(t, im), Wd(t, jm), dist, prm->cutoff, prm->theta, prm->longshort);
    if (f == 1) { pl_push(m2l, jm, im); return; }
    if (f == -1) return;
    if (li) { walk_rec(t, prm, im, son(t, jm, 0), p2p, m2l); walk_rec(t, prm, im, son(t, jm, 1), p2p, m2l); return; }
    if (lj) { walk_rec(t, prm, son(t, im, 0), jm, p2p, m2l); walk_rec(t, prm, son(t, im, 1), jm, p2p, m2l); return; }
    const double *wi = Wd(t, im), *wj = Wd(t, jm);
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2]) {   /* src/fmm.c:518-527 */
        walk_rec(t, prm, son(t, im, 0), jm, p2p, m2l); walk_rec(t, prm, son(t, im, 1), jm, p2p, m2l);
    } else {
        walk_rec(t, prm, im, son(t, jm, 0), p2p, m2l); walk_rec(t, prm, im, son(t, jm, 1), p2p, m2l);
    }
}

void pno_walk_local(const pno_tree *t, const pno_params *prm, int **p2p_s, int **p2p_t, long *np2p,
                    int **m2l_s, int **m2l_t, long *nm2l) {
    plist a = {0}, b = {0};
    if (t->n > 0) walk_rec(t, prm, t->first_node, t->first_node, &a, &b);
    *p2p_s = a.s; *p2p_t = a.t; *np2p = a.n; *m2l_s = b.s; *m2l_t = b.t; *nm2l = b.n;
}
void pno_free(void *p) { free(p); }

/* task_compute_p2p, src/fmm.c:796-872 */
void pno_eval_p2p(const pno_tree *t, const double *pos, const pno_params *prm, const int *s, const int *tt, long n, double *acc) {
    for (long k = 0; k < n; k++) {
        int li = tt[k] - t->first_leaf, lj = s[k] - t->first_leaf;
        pno_p2p_pair(pos, t->lf_ipart[li], t->lf_npart[li], pos, t->lf_ipart[lj], t->lf_npart[lj], 1, prm, acc);
    }
}
/* task_compute_m2l, src/fmm.c:875-907 */
void pno_eval_m2l(pno_tree *t, const pno_params *prm, const int *s, const int *tt, long n) {
    for (long k = 0; k < n; k++) {
        const double *ci = Cn(t, tt[k]), *cj = Cn(t, s[k]);
        pno_m2l(ci[0]-cj[0], ci[1]-cj[1], ci[2]-cj[2], Mp(t, s[k]), Lp(t, tt[k]), prm->rs, prm->longshort);
    }
}

/* ------------------------------------------------------------------------------------------
 * LET prune + pack, src/remotes.c:60-169
 * ------------------------------------------------------------------------------------------ */
typedef struct { pno_let *l; int ncap, bcap; } letbuf;
static void let_grow(letbuf *b, int need_nodes, int need_bodies) {
    pno_let *l = b->l;
    if (l->nnode + need_nodes > b->ncap) {
        while (l->nnode + need_nodes > b->ncap) b->ncap = b->ncap ? 2*b->ncap : 1024;
        l->npart = (int *)realloc(l->npart, b->ncap*sizeof(int));
        l->son = (int *)realloc(l->son, 2*b->ncap*sizeof(int));
        l->width = (double *)realloc(l->width, 3*b->ncap*sizeof(double));
        l->center = (double *)realloc(l->center, 3*b->ncap*sizeof(double));
        l->M = (double *)realloc(l->M, NM*b->ncap*sizeof(double));
        l->origin = (int *)realloc(l->origin, b->ncap*sizeof(int));
    }
    if (l->nbody + need_bodies > b->bcap) {
        while (l->nbody + need_bodies > b->bcap) b->bcap = b->bcap ? 2*b->bcap : 4096;
        l->body = (double *)realloc(l->body, 3*b->bcap*sizeof(double));
    }
}

static void pack_rec(const pno_tree *t, const double *pos, const pno_params *prm, const double tc[3], const double tw[3],
                     const double disp[3], letbuf *b, int isend, int ilocal) {
    pno_let *l = b->l;
    const double *c = Cn(t, ilocal), *w = Wd(t, ilocal);
    for (int d = 0; d < 3; d++) { l->center[3*isend+d] = c[d] + disp[d]; l->width[3*isend+d] = w[d]; }
    memcpy(l->M + NM*isend, Mp(t, ilocal), NM*sizeof(double));
    l->origin[isend] = ilocal;
    if (is_leaf(t, ilocal)) {                      /* :65-95 */
        int lk = ilocal - t->first_leaf;
        l->npart[isend] = t->lf_npart[lk];
        let_grow(b, 0, t->lf_npart[lk]);
        l->son[2*isend] = l->nbody;
        for (int p = t->lf_ipart[lk]; p < t->lf_ipart[lk] + t->lf_npart[lk]; p++) {
            for (int d = 0; d < 3; d++) l->body[3*l->nbody + d] = pos[3*p+d] + disp[d];
            l->nbody++;
        }
        l->son[2*isend+1] = l->nbody;
        return;
    }
    l->npart[isend] = t->nd_npart[ilocal - t->first_node];
    /* box-box gap between the target domain box and the displaced node box, :97-119 */
    double g[3], dr = 0.0;
    for (int d = 0; d < 3; d++) {
        g[d] = tc[d] - c[d] - disp[d];
        if (g[d] < 0.0) g[d] = -g[d];
        g[d] -= (tw[d] + w[d]) * 0.5;
    }
    for (int d = 0; d < 3; d++) if (g[d] > 0.0) dr += g[d]*g[d];
    dr = sqrt(dr);
    double wmax = w[0];
    if (wmax < w[1]) wmax = w[1];
    if (wmax < w[2]) wmax = w[2];
    l->son[2*isend] = -1; l->son[2*isend+1] = -1;
    if (prm->longshort && dr >= prm->cutoff) return;             /* :145-151 */
    if (wmax < 0.95 * prm->theta * dr) return;                  /* :154-158 */
    for (int s = 0; s < 2; s++) {                                /* :160-166 */
        int ch = son(t, ilocal, s);
        if (ch >= t->first_leaf) {
            let_grow(b, 1, 0);
            l = b->l;
            int slot = l->nnode++;
            l->son[2*isend + s] = slot;
            pack_rec(t, pos, prm, tc, tw, disp, b, slot, ch);
        }
    }
}

pno_let *pno_let_pack(const pno_tree *t, const double *pos, const pno_params *prm, const double tcenter[3],
                      const double twidth[3], const double displace[3]) {
    letbuf b = {0};
    b.l = (pno_let *)calloc(1, sizeof(pno_let));
    if (t->n == 0) return b.l;
    let_grow(&b, 1, 0);
    b.l->nnode = 1;
    pack_rec(t, pos, prm, tcenter, twidth, displace, &b, 0, t->first_node);
    return b.l;
}

void pno_let_free(pno_let *l) {
    if (!l) return;
    free(l->npart); free(l->son); free(l->width); free(l->center); free(l->M); free(l->body); free(l->origin); free(l);
}

/* remote walks, src/remotes.c:213-372 (P2P) and :405-552 (M2L), merged into one traversal */
static void rwalk_rec(const pno_tree *t, const pno_let *l, const pno_params *prm, int im, int jm, plist *p2p, plist *m2l) {
    int li = is_leaf(t, im);
    int rl = l->npart[jm] <= prm->maxleaf;                      /* "copy leaf", :228 */
    if (li && rl) { pl_push(p2p, jm, im); return; }
    const double *ci = Cn(t, im);
    double dist[3] = {ci[0]-l->center[3*jm], ci[1]-l->center[3*jm+1], ci[2]-l->center[3*jm+2]};
    int f = pno_acceptance(Wd(t, im), l->width + 3*jm, dist, prm->cutoff, prm->theta, prm->longshort);
    if (f == -1) return;
    int pruned = l->son[2*jm] < 0 || l->son[2*jm+1] < 0;
    if (li) {                                                   /* :243-281 / :427-464 */
        if (f == 1 || pruned) { pl_push(m2l, jm, im); return; }
        rwalk_rec(t, l, prm, im, l->son[2*jm], p2p, m2l); rwalk_rec(t, l, prm, im, l->son[2*jm+1], p2p, m2l);
        return;
    }
    if (f == 1) { pl_push(m2l, jm, im); return; }
    if (rl) {                                                   /* :283-319 / :466-500 */
        rwalk_rec(t, l, prm, son(t, im, 0), jm, p2p, m2l); rwalk_rec(t, l, prm, son(t, im, 1), jm, p2p, m2l);
        return;
    }
    const double *wi = Wd(t, im), *wj = l->width + 3*jm;        /* :351-361 / :531-541 */
    if (wi[0] + wi[1] + wi[2] > wj[0] + wj[1] + wj[2] || pruned) {
        rwalk_rec(t, l, prm, son(t, im, 0), jm, p2p, m2l); rwalk_rec(t, l, prm, son(t, im, 1), jm, p2p, m2l);
    } else {
        rwalk_rec(t, l, prm, im, l->son[2*jm], p2p, m2l); rwalk_rec(t, l, prm, im, l->son[2*jm+1], p2p, m2l);
    }
}

void pno_walk_remote(const pno_tree *t, const pno_let *l, const pno_params *prm, int **p2p_s, int **p2p_t, long *np2p,
                     int **m2l_s, int **m2l_t, long *nm2l) {
    plist a = {0}, b = {0};
    if (t->n > 0 && l->nnode > 0) rwalk_rec(t, l, prm, t->first_node, 0, &a, &b);
    *p2p_s = a.s; *p2p_t = a.t; *np2p = a.n; *m2l_s = b.s; *m2l_t = b.t; *nm2l = b.n;
}

/* task_compute_p2p_ext -> p2p_kernel_ex, src/remotes.c:583-596, :14-57 */
void pno_eval_p2p_remote(const pno_tree *t, const double *pos, const pno_let *l, const pno_params *prm,
                         const int *s, const int *tt, long n, double *acc) {
    for (long k = 0; k < n; k++) {
        int li = tt[k] - t->first_leaf, j = s[k];
        pno_p2p_pair(pos, t->lf_ipart[li], t->lf_npart[li], l->body, l->son[2*j], l->npart[j], 0, prm, acc);
    }
}
/* task_compute_m2l_ext, src/remotes.c:598-628 */
void pno_eval_m2l_remote(pno_tree *t, const pno_let *l, const pno_params *prm, const int *s, const int *tt, long n) {
    for (long k = 0; k < n; k++) {
        const double *ci = Cn(t, tt[k]);
        int j = s[k];
        pno_m2l(ci[0]-l->center[3*j], ci[1]-l->center[3*j+1], ci[2]-l->center[3*j+2], l->M + NM*j, Lp(t, tt[k]),
                prm->rs, prm->longshort);
    }
}

/* ------------------------------------------------------------------------------------------
 * domain geometry
 * ------------------------------------------------------------------------------------------ */
static int mostleft_of(int P) {                    /* src/initial.c:199-223 */
    int m = 1;
    while (m < 2*P - 1) m *= 2;
    m = m/2 - 1;
    if (P == 1) m = 0;
    return m;
}
static int domain_node_of_rank(int rank, int P) {
    int d = rank + mostleft_of(P);
    if (d > 2*P - 2) d -= P;
    return d;
}
/* leaf counts below each heap node: domain_initialize sets time_node = 1 on every domain and
 * fill_time_domtree sums them up (src/domains.c:7-19, 437-464) */
static double dom_time(int node, int P) {
    if (node >= P - 1) return 1.0;
    return dom_time(2*node+1, P) + dom_time(2*node+2, P);
}
/* domain_volume_part (src/domains.c:399-428) + center_toptree (src/toptree.c:150-181) */
static void dom_rec(int node, int P, int dim, double bl[3], double br[3], double *splits,
                    double *ncenter, double *nwidth, int *ndirect) {
    for (int d = 0; d < 3; d++) { nwidth[3*node+d] = br[d] - bl[d]; ncenter[3*node+d] = 0.5*(br[d] + bl[d]); }
    ndirect[node] = dim;
    if (node >= P - 1) return;
    double tl = dom_time(2*node+1, P), tr = dom_time(2*node+2, P);
    double norm = tl + tr;
    double frac = bl[dim] + (br[dim] - bl[dim]) * tl / norm;
    splits[node] = frac;
    double keep = br[dim];
    br[dim] = frac;
    dom_rec(2*node+1, P, (dim+1)%3, bl, br, splits, ncenter, nwidth, ndirect);
    br[dim] = keep;
    keep = bl[dim];
    bl[dim] = frac;
    dom_rec(2*node+2, P, (dim+1)%3, bl, br, splits, ncenter, nwidth, ndirect);
    bl[dim] = keep;
}

void pno_domain_boxes(int P, double box, double *center, double *width, int *direct_start, double *splits) {
    int len = 2*P - 1;
    double *nc = (double *)calloc(3*len, sizeof(double)), *nw = (double *)calloc(3*len, sizeof(double));
    int *nd = (int *)calloc(len, sizeof(int));
    double bl[3] = {0.0, 0.0, 0.0}, br[3] = {box, box, box};
    for (int i = 0; i < len; i++) splits[i] = 0.0;
    dom_rec(0, P, 0, bl, br, splits, nc, nw, nd);
    for (int r = 0; r < P; r++) {
        int dn = domain_node_of_rank(r, P);
        for (int d = 0; d < 3; d++) { center[3*r+d] = nc[3*dn+d]; width[3*r+d] = nw[3*dn+d]; }
        direct_start[r] = nd[dn];
    }
    free(nc); free(nw); free(nd);
}

/* prepare_body_inOrderOf_domain / bksort_body_inplace criterion (src/domains.c:163-296): pos > split -> right */
int pno_domain_of(const double x[3], int P, const double *splits) {
    int node = 0, dim = 0;
    while (node < P - 1) {
        node = (x[dim] > splits[node]) ? 2*node + 2 : 2*node + 1;
        dim = (dim + 1) % 3;
    }
    return (node - mostleft_of(P) + P) % P;
}

/* ------------------------------------------------------------------------------------------
 * whole force evaluation, ranks simulated one after another (src/photoNs.c:83-116 without PM;
 * src/fmm.c:719-770, 909-984, 987-1075)
 * ------------------------------------------------------------------------------------------ */
void pno_force(const double *pos_in, int n, int P, const pno_params *prm, double *acc_out, double *counters) {
    init_tables();
    double *dc = (double *)malloc(3*P*sizeof(double)), *dw = (double *)malloc(3*P*sizeof(double));
    double *splits = (double *)malloc((2*P-1)*sizeof(double));
    int *dstart = (int *)malloc(P*sizeof(int));
    pno_domain_boxes(P, prm->box, dc, dw, dstart, splits);
    /* connect_local_toptree (src/toptree.c:20-27) overwrites every domain's toptree box with its local
     * root box, i.e. with the box after one trip through left/right = center -/+ 0.5 width
     * (src/fmm.c:194-197, 126-131); prepare_sendtree2 prunes against THAT box (src/remotes.c:97-110). */
    double *pc = (double *)malloc(3*P*sizeof(double)), *pw = (double *)malloc(3*P*sizeof(double));
    for (int i = 0; i < 3*P; i++) {
        double l = dc[i] - 0.5*dw[i], r = dc[i] + 0.5*dw[i];
        pw[i] = r - l; pc[i] = 0.5*(r + l);
    }
    /* particle -> rank */
    int *cnt = (int *)calloc(P, sizeof(int)), *owner = (int *)malloc((n > 0 ? n : 1)*sizeof(int));
    for (int i = 0; i < n; i++) { owner[i] = pno_domain_of(pos_in + 3*i, P, splits); cnt[owner[i]]++; }
    double **pos = (double **)malloc(P*sizeof(double *)), **acc = (double **)malloc(P*sizeof(double *));
    long **ids = (long **)malloc(P*sizeof(long *));
    pno_tree **tree = (pno_tree **)malloc(P*sizeof(pno_tree *));
    for (int r = 0; r < P; r++) {
        pos[r] = (double *)malloc((3*cnt[r] + 3)*sizeof(double));
        acc[r] = (double *)calloc(3*cnt[r] + 3, sizeof(double));
        ids[r] = (long *)malloc((cnt[r] + 1)*sizeof(long));
        cnt[r] = 0;
    }
    for (int i = 0; i < n; i++) {
        int r = owner[i], k = cnt[r]++;
        pos[r][3*k] = pos_in[3*i]; pos[r][3*k+1] = pos_in[3*i+1]; pos[r][3*k+2] = pos_in[3*i+2];
        ids[r][k] = i;
    }
    for (int c = 0; c < 8; c++) counters[c] = 0.0;
    /* fmm_prepare + fmm_task per rank */
    for (int r = 0; r < P; r++) {
        double bl[3], br[3];
        for (int d = 0; d < 3; d++) { bl[d] = dc[3*r+d] - 0.5*dw[3*r+d]; br[d] = dc[3*r+d] + 0.5*dw[3*r+d]; }
        tree[r] = pno_tree_build(pos[r], ids[r], cnt[r], prm->maxleaf, dstart[r], bl, br);
        pno_tree_upward(tree[r], pos[r], prm->mass);
        int *ps, *pt, *ms, *mt; long np, nm;
        pno_walk_local(tree[r], prm, &ps, &pt, &np, &ms, &mt, &nm);
        pno_eval_p2p(tree[r], pos[r], prm, ps, pt, np, acc[r]);
        pno_eval_m2l(tree[r], prm, ms, mt, nm);
        counters[0] += (double)np; counters[1] += (double)nm; counters[4] += (double)nm;
        for (long k = 0; k < np; k++) {
            int a = tree[r]->lf_npart[ps[k] - tree[r]->first_leaf], b = tree[r]->lf_npart[pt[k] - tree[r]->first_leaf];
            counters[2] += (double)a*b - (ps[k] == pt[k] ? b : 0);
        }
        counters[5] += tree[r]->last_leaf - tree[r]->first_leaf;
        counters[6] += tree[r]->last_node - tree[r]->first_node + 1;
        free(ps); free(pt); free(ms); free(mt);
    }
    /* fmm_ext: for shift 0 peers 1..P-1, then 26 image shifts x all peers (src/fmm.c:1021-1045).
     * Receiver r gets, in call idx, the tree of sender (r - idx) mod P pruned against r's box
     * (src/remotes.c:690-720: sender packs for srank = rank + idx). */
    for (int pass = 0; pass < (prm->periodic ? 27 : 1); pass++) {
        double sh[3] = {0.0, 0.0, 0.0};
        if (pass > 0) {
            /* enumerate mi, mj, mk in {-1,0,1} skipping (0,0,0), order of src/fmm.c:1029-1037 */
            int q = pass - 1;
            if (q >= 13) q++;
            sh[0] = (double)(q / 9 - 1) * prm->box; sh[1] = (double)((q / 3) % 3 - 1) * prm->box; sh[2] = (double)(q % 3 - 1) * prm->box;
        }
        for (int idx = (pass == 0 ? 1 : 0); idx < P; idx++) {
            for (int r = 0; r < P; r++) {
                int sender = (r - idx + P) % P;
                pno_let *l = pno_let_pack(tree[sender], pos[sender], prm, pc + 3*r, pw + 3*r, sh);
                int *ps, *pt, *ms, *mt; long np, nm;
                pno_walk_remote(tree[r], l, prm, &ps, &pt, &np, &ms, &mt, &nm);
                pno_eval_p2p_remote(tree[r], pos[r], l, prm, ps, pt, np, acc[r]);
                pno_eval_m2l_remote(tree[r], l, prm, ms, mt, nm);
                counters[7] += (double)np; counters[4] += (double)nm;
                for (long k = 0; k < np; k++)
                    counters[3] += (double)l->npart[ps[k]] * tree[r]->lf_npart[pt[k] - tree[r]->first_leaf];
                free(ps); free(pt); free(ms); free(mt);
                pno_let_free(l);
            }
        }
    }
    for (int r = 0; r < P; r++) {
        pno_tree_downward(tree[r], pos[r], acc[r]);
        for (int k = 0; k < cnt[r]; k++) {
            long i = ids[r][k];
            acc_out[3*i] = acc[r][3*k]; acc_out[3*i+1] = acc[r][3*k+1]; acc_out[3*i+2] = acc[r][3*k+2];
        }
        pno_tree_free(tree[r]); free(pos[r]); free(acc[r]); free(ids[r]);
    }
    free(pos); free(acc); free(ids); free(tree); free(cnt); free(owner); free(dc); free(dw); free(pc); free(pw); free(splits); free(dstart);
}

/* ------------------------------------------------------------------------------------------
 * Mode B tree: CPU restatement of the DEVICE builder (photons-2.0_b200/csrc/pn2_tree.cu).
 * Same tree definition as src/fmm.c:30-264 (mean split, cycling direction, <= MAXLEAF -> leaf, boxes
 * cut by the ancestors' splits) but built level by level and in INTEGER arithmetic:
 *   q_d = trunc((x_d - lo_d) * 2^(32-e)) (uint32, 2^e > largest box extent); Morton pre-sort on the top
 *   10 bits of each q (30-bit key, stable: ties keep the caller's order); a particle goes right iff q * count > sum(q) over its node (exact 64-bit),
 *   all of them right if count < 2 (src/fmm.c:33-36); stable partition; split = lo + (sum/count) * 2^-(32-e);
 *   breadth-first ids (per level: nodes in range order, son 0 before son 1).
 * Everything here must match the device bit for bit.
 * ------------------------------------------------------------------------------------------ */
static unsigned long long spread10(unsigned long long v) {          /* bit k of v -> bit 3k */
    unsigned long long r = 0;
    for (int k = 0; k < 10; k++) r |= ((v >> k) & 1ULL) << (3 * k);
    return r;
}
static unsigned quant32(double x, double lo, double S) {
    double f = (x - lo) * S;
    if (!(f > 0.0)) return 0u;
    if (f >= 4294967295.0) return 4294967295u;
    return (unsigned)f;
}
typedef struct { unsigned long long key; int idx; } mkey;
static int mkey_cmp(const void *a, const void *b) {
    const mkey *x = (const mkey *)a, *y = (const mkey *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);          /* stable */
}

pno_tree *pno_treeB_build(const double *pos_in, int n, int maxleaf, int direct0, const double bl[3], const double br[3],
                          double *pos_out, int *order) {
    init_tables();
    pno_tree *t = (pno_tree *)calloc(1, sizeof *t);
    t->n = n; t->maxleaf = maxleaf;
    t->first_leaf = t->last_leaf = n;
    t->first_node = n; t->last_node = n - 1;
    if (n == 0) return t;
    double ext[3];
    for (int d = 0; d < 3; d++) ext[d] = br[d] - bl[d];
    double emax = ext[0] > ext[1] ? ext[0] : ext[1];
    if (ext[2] > emax) emax = ext[2];
    int e2 = 0;
    frexp(emax, &e2);
    const double S = ldexp(1.0, 32 - e2), invS = ldexp(1.0, e2 - 32);
    unsigned *q0 = (unsigned *)malloc(3*(size_t)n*sizeof(unsigned));
    mkey *mk = (mkey *)malloc(n*sizeof(mkey));
    for (int i = 0; i < n; i++) {
        for (int d = 0; d < 3; d++) q0[3*(size_t)i+d] = quant32(pos_in[3*(size_t)i+d], bl[d], S);
        mk[i].key = (spread10(q0[3*(size_t)i] >> 22) << 2) | (spread10(q0[3*(size_t)i+1] >> 22) << 1) | spread10(q0[3*(size_t)i+2] >> 22);
        mk[i].idx = i;
    }
    qsort(mk, n, sizeof(mkey), mkey_cmp);
    unsigned *qa = (unsigned *)malloc(3*(size_t)n*sizeof(unsigned)), *qb = (unsigned *)malloc(3*(size_t)n*sizeof(unsigned));
    int *ia = (int *)malloc(n*sizeof(int)), *ib = (int *)malloc(n*sizeof(int));
    for (int i = 0; i < n; i++) { ia[i] = mk[i].idx; for (int d = 0; d < 3; d++) qa[3*(size_t)i+d] = q0[3*(size_t)mk[i].idx+d]; }
    free(mk); free(q0);
    /* growable node / leaf records */
    int ncap = 1024, lcap = 1024, nn = 1, nl = 0;
    int *ns = (int *)malloc(ncap*sizeof(int)), *nc = (int *)malloc(ncap*sizeof(int)), *nson = (int *)malloc(2*ncap*sizeof(int));
    double *nbox = (double *)malloc(6*ncap*sizeof(double)), *nsp = (double *)malloc(ncap*sizeof(double));
    int *ls = (int *)malloc(lcap*sizeof(int)), *lc = (int *)malloc(lcap*sizeof(int));
    double *lbox = (double *)malloc(6*lcap*sizeof(double));
    ns[0] = 0; nc[0] = n;
    for (int d = 0; d < 3; d++) { nbox[d] = bl[d]; nbox[3+d] = br[d]; }
    int node0 = 0, cnt = 1, level = 0;
    while (cnt > 0) {
        int dir = (direct0 + level) % 3;
        double lo = bl[dir];
        int next0 = node0 + cnt, nnext = 0;
        for (int k = 0; k < cnt; k++) {
            int nd = node0 + k, a = ns[nd], c = nc[nd];
            unsigned long long sum = 0;
            for (int i = a; i < a + c; i++) sum += qa[3*(size_t)i+dir];
            double m = (double)sum / (double)c;
            double split = lo + m * invS;
            nsp[nd] = split;
            /* stable partition into qb / ib */
            int nleft = 0;
            for (int i = a; i < a + c; i++) {
                int fl = (c < 2) ? 1 : ((unsigned long long)qa[3*(size_t)i+dir] * (unsigned long long)c > sum);
                if (!fl) nleft++;
            }
            int pl = a, pr = a + nleft;
            for (int i = a; i < a + c; i++) {
                int fl = (c < 2) ? 1 : ((unsigned long long)qa[3*(size_t)i+dir] * (unsigned long long)c > sum);
                int dst = fl ? pr++ : pl++;
                qb[3*(size_t)dst] = qa[3*(size_t)i]; qb[3*(size_t)dst+1] = qa[3*(size_t)i+1]; qb[3*(size_t)dst+2] = qa[3*(size_t)i+2];
                ib[dst] = ia[i];
            }
            int cn[2] = {nleft, c - nleft}, st[2] = {a, a + nleft};
            for (int s = 0; s < 2; s++) {
                double cb[6];
                for (int d = 0; d < 6; d++) cb[d] = nbox[6*(size_t)nd+d];
                if (s == 0) cb[3+dir] = split; else cb[dir] = split;
                if (cn[s] <= maxleaf) {
                    if (nl == lcap) { lcap *= 2; ls = (int *)realloc(ls, lcap*sizeof(int)); lc = (int *)realloc(lc, lcap*sizeof(int)); lbox = (double *)realloc(lbox, 6*(size_t)lcap*sizeof(double)); }
                    ls[nl] = st[s]; lc[nl] = cn[s];
                    for (int d = 0; d < 6; d++) lbox[6*(size_t)nl+d] = cb[d];
                    nson[2*nd+s] = -(nl + 2);
                    nl++;
                } else {
                    int ni = next0 + nnext;
                    if (ni >= ncap) { ncap *= 2; ns = (int *)realloc(ns, ncap*sizeof(int)); nc = (int *)realloc(nc, ncap*sizeof(int)); nson = (int *)realloc(nson, 2*(size_t)ncap*sizeof(int)); nbox = (double *)realloc(nbox, 6*(size_t)ncap*sizeof(double)); nsp = (double *)realloc(nsp, ncap*sizeof(double)); }
                    ns[ni] = st[s]; nc[ni] = cn[s];
                    for (int d = 0; d < 6; d++) nbox[6*(size_t)ni+d] = cb[d];
                    nson[2*nd+s] = ni;
                    nnext++;
                }
            }
        }
        /* particles of finished leaves do not move: copy back only this level's node ranges */
        for (int k = 0; k < cnt; k++) {
            int nd = node0 + k, a = ns[nd], c = nc[nd];
            memcpy(qa + 3*(size_t)a, qb + 3*(size_t)a, 3*(size_t)c*sizeof(unsigned));
            memcpy(ia + a, ib + a, c*sizeof(int));
        }
        node0 = next0; cnt = nnext; nn = next0 + nnext; level++;
        if (level > 200) { fprintf(stderr, "pn_oracle: treeB deeper than 200 levels\n"); exit(3); }
    }
    nn = node0;
    /* fill the pno_tree in the reference id space: leaves n..n+nl-1, nodes n+nl..n+nl+nn-1 */
    t->leafcap = nl; t->nodecap = nn;
    t->first_leaf = n; t->last_leaf = n + nl; t->first_node = n + nl; t->last_node = n + nl + nn - 1;
    t->lf_npart = (int *)calloc(nl, sizeof(int)); t->lf_ipart = (int *)calloc(nl, sizeof(int));
    t->lf_center = (double *)calloc(3*(size_t)nl, sizeof(double)); t->lf_width = (double *)calloc(3*(size_t)nl, sizeof(double));
    t->lf_M = (double *)calloc(NM*(size_t)nl, sizeof(double)); t->lf_L = (double *)calloc(NM*(size_t)nl, sizeof(double));
    t->nd_npart = (int *)calloc(nn, sizeof(int)); t->nd_son = (int *)malloc(2*(size_t)nn*sizeof(int));
    t->nd_split = (double *)calloc(nn, sizeof(double));
    t->nd_center = (double *)calloc(3*(size_t)nn, sizeof(double)); t->nd_width = (double *)calloc(3*(size_t)nn, sizeof(double));
    t->nd_M = (double *)calloc(NM*(size_t)nn, sizeof(double)); t->nd_L = (double *)calloc(NM*(size_t)nn, sizeof(double));
    for (int k = 0; k < nl; k++) {
        t->lf_npart[k] = lc[k]; t->lf_ipart[k] = ls[k];
        for (int d = 0; d < 3; d++) {
            t->lf_center[3*k+d] = 0.5 * (lbox[6*(size_t)k+3+d] + lbox[6*(size_t)k+d]);
            t->lf_width[3*k+d] = lbox[6*(size_t)k+3+d] - lbox[6*(size_t)k+d];
        }
    }
    for (int k = 0; k < nn; k++) {
        t->nd_npart[k] = nc[k]; t->nd_split[k] = nsp[k];
        for (int d = 0; d < 3; d++) {
            t->nd_center[3*k+d] = 0.5 * (nbox[6*(size_t)k+3+d] + nbox[6*(size_t)k+d]);
            t->nd_width[3*k+d] = nbox[6*(size_t)k+3+d] - nbox[6*(size_t)k+d];
        }
        for (int s = 0; s < 2; s++) {
            int ch = nson[2*k+s];
            t->nd_son[2*k+s] = ch >= 0 ? t->first_node + ch : t->first_leaf + (-(ch + 2));
        }
    }
    if (pos_out) for (int i = 0; i < n; i++) for (int d = 0; d < 3; d++) pos_out[3*(size_t)i+d] = pos_in[3*(size_t)ia[i]+d];
    if (order) memcpy(order, ia, n*sizeof(int));
    free(qa); free(qb); free(ia); free(ib); free(ns); free(nc); free(nson); free(nbox); free(nsp); free(ls); free(lc); free(lbox);
    return t;
}
