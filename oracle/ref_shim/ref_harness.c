/*
 * ref_harness.c -- drives the UNMODIFIED reference short-range path and dumps
 * everything the parity tests need.  TEST INFRASTRUCTURE ONLY: never linked
 * into, imported by or executed from the product path.
 *
 * Sequence (mirrors /root/reference/src/photoNs.c:83-116, PM excluded):
 *   MPI_Init; setup_domain_index; domain_initialize; domain_decomposition;
 *   fmm_construct; fmm_prepare; fmm_task; fmm_ext;           -> dump
 *
 * Particle identity: Body has no id field (inc/typesdef.h:53-56), and both
 * domain_decomposition and build_kdtree permute part[] in place, so the global
 * input index is carried in Body.vel[0] (vel is never touched on this path).
 *
 * Usage:  PN_SHIM_NP=<ranks> ref_fmm <params.txt> <pos.f64> <outprefix>
 *   params.txt: "key value" lines: NPART_TOTAL BOXSIZE NSIDE MAXLEAF OPENANGLE
 *               SPLIT (<=0: 1.25*BOX/NSIDE) SOFT (<0: 0.03*BOX/N^(1/3)) MASS
 *               CAPTURE (0|1|2) REPEAT (force evaluations to time, default 1)
 *   pos.f64:    NPART_TOTAL x 3 doubles
 *   output:     <outprefix>.r<rank>.bin, a sequence of records
 *               { char name[24]; long nbytes; payload }
 */
#include "photoNs.h"
#include "fmm.h"
#include "domains.h"
#include "initial.h"
#include "pn_capture.h"
#ifdef PN2_GPU_GLUE
#include "pn2_fmm_glue.h"
#endif
#include <string.h>
#include <sys/time.h>

int pn_capture_level = 0;
PnPairList pn_cap_p2p, pn_cap_m2l;
PnRemoteCap *pn_rcap = NULL;
long pn_nrcap = 0, pn_rcap_cap = 0;

/* Fortran symbols of the (excluded) PM part: src/initial.c:226,468; inc/partmesh.h:25 */
void get_local_size_(int nside[3], int vp[2], int start[3], int end[3], int size[3]) {
    (void)vp;
    for (int d = 0; d < 3; d++) { start[d] = 0; end[d] = nside[d] - 1; size[d] = nside[d]; }
}
void convolution_(double *a, int *b, double *c) { (void)a; (void)b; (void)c; }

typedef struct {
    long npart_total;
    double box, split, soft, mass, theta;
    int nside, maxleaf, capture, repeat;
} PnRefParams;

static double wall(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void rec(FILE *f, const char *name, const void *data, long nbytes) {
    char nm[24];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 23);
    fwrite(nm, 1, 24, f);
    fwrite(&nbytes, sizeof(long), 1, f);
    if (nbytes > 0) fwrite(data, 1, nbytes, f);
}

static void reset_capture(void) {
    pn_cap_p2p.n = 0; pn_cap_m2l.n = 0;
    for (long i = 0; i < pn_nrcap; i++) {
        free(pn_rcap[i].tree); free(pn_rcap[i].body);
        free(pn_rcap[i].p2p.s); free(pn_rcap[i].p2p.t);
        free(pn_rcap[i].m2l.s); free(pn_rcap[i].m2l.t);
    }
    pn_nrcap = 0;
}

/* set the derived force parameters exactly as src/initial.c:316-345 does */
static void set_force_params(const PnRefParams *p) {
    BOXSIZE = p->box;
    NSIDE = p->nside;
    NPART_TOTAL = p->npart_total;
    MAXLEAF = p->maxleaf;
    open_angle = p->theta;
    MASSPART = p->mass;
    double invside = BOXSIZE / ((double)NSIDE);
    splitRadius = 1.25 * invside;
    SoftenScale = 0.03 * BOXSIZE / pow(((double)NPART_TOTAL), 0.3333333);
    if (p->split > 0.0) splitRadius = p->split;
    cutoffRadius = 4.5 * splitRadius;
    if (p->soft >= 0.0) SoftenScale = p->soft;
    BoxMinimum = 0.0;            /* src/initial.c:495-497 (used by the non-periodic build, src/fmm.c:350-351) */
    BoxMaximum = BOXSIZE;
}

int pn_ref_run(const double *pos, const PnRefParams *p, const char *outprefix) {
    MPI_Comm_rank(MPI_COMM_WORLD, &PROC_RANK);
    MPI_Comm_size(MPI_COMM_WORLD, &PROC_SIZE);
    MPI_Type_contiguous(sizeof(RemoteNode), MPI_CHAR, &strReNode);
    MPI_Type_contiguous(sizeof(RemoteBody), MPI_CHAR, &strReBody);
    MPI_Type_contiguous(sizeof(Body), MPI_CHAR, &strBody);
    set_force_params(p);
    pn_capture_level = p->capture;
    reset_capture();

    setup_domain_index();

    long lo = p->npart_total * PROC_RANK / PROC_SIZE;
    long hi = p->npart_total * (PROC_RANK + 1) / PROC_SIZE;
    NPART = (int)(hi - lo);
    NPART_MEAN = (int)(p->npart_total / PROC_SIZE);
    reset_mem();
    part = (Body *)pmalloc(sizeof(Body) * (NPART > 0 ? NPART : 1), 0);
    for (long n = lo; n < hi; n++) {
        Body *b = &part[n - lo];
        memset(b, 0, sizeof *b);
        b->pos[0] = pos[3 * n + 0];
        b->pos[1] = pos[3 * n + 1];
        b->pos[2] = pos[3 * n + 2];
        b->vel[0] = (double)n;
    }

    domain_initialize();
    DTIME_FRACTION = 1.0;
    domain_decomposition();

    double t_construct = 0, t_prepare = 0, t_task = 0, t_ext = 0, t_total = 0;
    int rep;
    for (rep = 0; rep < (p->repeat > 0 ? p->repeat : 1); rep++) {
        if (rep > 0) fmm_deconstruct();
        for (int n = 0; n < NPART; n++) part[n].acc[0] = part[n].acc[1] = part[n].acc[2] = 0.0;
        reset_capture();
        p2p_count_remote = 0; walk_m2l_count = 0;
        MPI_Barrier(MPI_COMM_WORLD);
        double t0 = wall();
        fmm_construct();
        double t1 = wall();
        fmm_prepare();
#ifdef PN2_GPU_GLUE
        pn2_glue_begin_step();          /* drop-in build: task batches go to libpn2gpu.so */
#endif
        double t2 = wall();
        fmm_task();
        double t3 = wall();
        fmm_ext();
#ifdef PN2_GPU_GLUE
        pn2_glue_end_step();
#endif
        MPI_Barrier(MPI_COMM_WORLD);
        double t4 = wall();
        t_construct += t1 - t0; t_prepare += t2 - t1; t_task += t3 - t2; t_ext += t4 - t3; t_total += t4 - t0;
    }

    char fname[512];
    snprintf(fname, sizeof fname, "%s.r%d.bin", outprefix, PROC_RANK);
    FILE *f = fopen(fname, "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", fname); return 1; }

    /* local interaction count, exactly as the lists define it */
    long nint_local = 0;
    for (long k = 0; k < pn_cap_p2p.n; k++) {
        int s = pn_cap_p2p.s[k], t = pn_cap_p2p.t[k];
        nint_local += (long)leaf[s].npart * leaf[t].npart - (s == t ? leaf[t].npart : 0);
    }

    double scal[24];
    memset(scal, 0, sizeof scal);
    scal[0] = BOXSIZE; scal[1] = splitRadius; scal[2] = cutoffRadius; scal[3] = SoftenScale;
    scal[4] = open_angle; scal[5] = MASSPART; scal[6] = (double)MAXLEAF; scal[7] = (double)NSIDE;
    scal[8] = (double)PROC_RANK; scal[9] = (double)PROC_SIZE; scal[10] = (double)NPART;
    scal[11] = (double)first_leaf; scal[12] = (double)last_leaf; scal[13] = (double)first_node;
    scal[14] = (double)last_node; scal[15] = (double)idxP2P; scal[16] = (double)idxM2L;
    scal[17] = (double)p2p_count_remote; scal[18] = (double)walk_m2l_count;
    scal[19] = (double)this_domain; scal[20] = (double)direct_local_start; scal[21] = (double)mostleft;
    scal[22] = (double)nint_local; scal[23] = (double)rep;
    rec(f, "scalars", scal, sizeof scal);
    double tim[5] = { t_construct, t_prepare, t_task, t_ext, t_total };
    rec(f, "timing", tim, sizeof tim);
    rec(f, "part", part, (long)sizeof(Body) * NPART);
    rec(f, "toptree", toptree, (long)sizeof(TopNode) * (2 * PROC_SIZE - 1));
    if (p->capture >= 1) {
        rec(f, "leaf", leaf + first_leaf, (long)sizeof(Pack) * (last_leaf - first_leaf));
        rec(f, "btree", btree + first_node, (long)sizeof(Node) * (last_node - first_node + 1));
        rec(f, "p2p_s", pn_cap_p2p.s, pn_cap_p2p.n * (long)sizeof(int));
        rec(f, "p2p_t", pn_cap_p2p.t, pn_cap_p2p.n * (long)sizeof(int));
        rec(f, "m2l_s", pn_cap_m2l.s, pn_cap_m2l.n * (long)sizeof(int));
        rec(f, "m2l_t", pn_cap_m2l.t, pn_cap_m2l.n * (long)sizeof(int));
    }
    if (p->capture >= 2) {
        long *hdr = (long *)malloc(sizeof(long) * 5 * (pn_nrcap + 1));
        for (long i = 0; i < pn_nrcap; i++) {
            hdr[5 * i + 0] = pn_rcap[i].seq; hdr[5 * i + 1] = pn_rcap[i].nnode; hdr[5 * i + 2] = pn_rcap[i].nbody;
            hdr[5 * i + 3] = pn_rcap[i].p2p.n; hdr[5 * i + 4] = pn_rcap[i].m2l.n;
        }
        rec(f, "rcap_hdr", hdr, (long)sizeof(long) * 5 * pn_nrcap);
        free(hdr);
        for (long i = 0; i < pn_nrcap; i++) {
            rec(f, "rcap_tree", pn_rcap[i].tree, (long)sizeof(RemoteNode) * pn_rcap[i].nnode);
            rec(f, "rcap_body", pn_rcap[i].body, (long)sizeof(RemoteBody) * pn_rcap[i].nbody);
            rec(f, "rcap_p2p_s", pn_rcap[i].p2p.s, pn_rcap[i].p2p.n * (long)sizeof(int));
            rec(f, "rcap_p2p_t", pn_rcap[i].p2p.t, pn_rcap[i].p2p.n * (long)sizeof(int));
            rec(f, "rcap_m2l_s", pn_rcap[i].m2l.s, pn_rcap[i].m2l.n * (long)sizeof(int));
            rec(f, "rcap_m2l_t", pn_rcap[i].m2l.t, pn_rcap[i].m2l.n * (long)sizeof(int));
        }
    }
    fclose(f);
    if (0 == PROC_RANK)
        printf("[ref_fmm] NP=%d N=%ld leaves=%d idxP2P=%lu idxM2L=%lu remoteint=%lu m2lcalls=%lu  t_total=%.4f s (prep %.4f task %.4f ext %.4f)\n",
               PROC_SIZE, p->npart_total, last_leaf - first_leaf, idxP2P, idxM2L, p2p_count_remote, walk_m2l_count,
               t_total / rep, t_prepare / rep, t_task / rep, t_ext / rep);
    fmm_deconstruct();
    domain_finalize();
    pfree(part, 0);
    return 0;
}

#ifdef PN_REF_MAIN
static int read_params(const char *fn, PnRefParams *p) {
    FILE *f = fopen(fn, "r");
    if (!f) return 1;
    char key[128];
    double v;
    memset(p, 0, sizeof *p);
    p->split = -1; p->soft = -1; p->theta = 0.4; p->maxleaf = 8; p->repeat = 1;
    while (fscanf(f, "%127s %lf", key, &v) == 2) {
        if (!strcmp(key, "NPART_TOTAL")) p->npart_total = (long)v;
        else if (!strcmp(key, "BOXSIZE")) p->box = v;
        else if (!strcmp(key, "NSIDE")) p->nside = (int)v;
        else if (!strcmp(key, "MAXLEAF")) p->maxleaf = (int)v;
        else if (!strcmp(key, "OPENANGLE")) p->theta = v;
        else if (!strcmp(key, "SPLIT")) p->split = v;
        else if (!strcmp(key, "SOFT")) p->soft = v;
        else if (!strcmp(key, "MASS")) p->mass = v;
        else if (!strcmp(key, "CAPTURE")) p->capture = (int)v;
        else if (!strcmp(key, "REPEAT")) p->repeat = (int)v;
    }
    fclose(f);
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_fmm params.txt pos.f64 outprefix\n"); return 2; }
    PnRefParams p;
    if (read_params(argv[1], &p)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    FILE *f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    double *pos = (double *)malloc(sizeof(double) * 3 * p.npart_total);
    if (fread(pos, sizeof(double), 3 * p.npart_total, f) != (size_t)(3 * p.npart_total)) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);
    MPI_Init(&argc, &argv);
    int rc = pn_ref_run(pos, &p, argv[3]);
    MPI_Finalize();
    return rc;
}
#endif
