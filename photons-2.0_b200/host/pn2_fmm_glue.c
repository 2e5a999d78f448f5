/*
 * pn2_fmm_glue.c -- see pn2_fmm_glue.h.  Compiled with the reference's include path (inc/photoNs.h declares the
 * globals part, leaf, btree, first_leaf ... splitRadius, cutoffRadius, SoftenScale, open_angle, MASSPART, MAXLEAF,
 * BOXSIZE) and -fcommon, like every reference source file.  Errors follow the reference's convention: message + exit.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "photoNs.h"
#include "pn2_fmm_glue.h"

static pn2_ctx *g_ctx = NULL;

static void die(const char *where) {
    fprintf(stderr, "pn2 glue: %s: %s\n", where, pn2_last_error());
    exit(1);
}
#define CK(call, where) do { if ((call) != PN2_OK) die(where); } while (0)

static void ensure_ctx(void) {
    pn2_params p;
    memset(&p, 0, sizeof p);
    p.box = BOXSIZE; p.rs = splitRadius; p.cutoff = cutoffRadius; p.soft = SoftenScale; p.theta = open_angle;
    p.mass = MASSPART; p.maxleaf = MAXLEAF;
#ifdef PERIODIC_CONDITION
    p.periodic = 1;
#endif
#ifdef LONGSHORT
    p.longshort = 1;
#endif
    const char *pm = getenv("PN2_PRECISION");
    p.precision = (pm && !strcmp(pm, "fp32")) ? PN2_FP32 : PN2_FP64;
    if (!g_ctx) {
        const char *dv = getenv("PN2_DEVICE");
        int ndev_rank = dv ? atoi(dv) : 0;
        CK(pn2_create(&g_ctx, ndev_rank, &p), "pn2_create");
    } else {
        CK(pn2_set_params(g_ctx, &p), "pn2_set_params");
    }
}

void pn2_glue_begin_step(void) {
    ensure_ctx();
    CK(pn2_set_particles(g_ctx, part[0].pos, sizeof(Body), NPART), "pn2_set_particles");
    CK(pn2_set_tree(g_ctx, (const pn2_pack *)&leaf[first_leaf], first_leaf, last_leaf, (const pn2_node *)&btree[first_node],
                    first_node, last_node), "pn2_set_tree");
    CK(pn2_p2m_m2m(g_ctx), "pn2_p2m_m2m");
}

void pn2_glue_local_batch(int kind, const int *s, const int *t, int nt) {
    if (kind == 0) CK(pn2_p2p_batch(g_ctx, s, t, nt), "pn2_p2p_batch");
    else CK(pn2_m2l_batch(g_ctx, s, t, nt), "pn2_m2l_batch");
}

void pn2_glue_set_remote(const void *rt, int nnode, const void *rb, int nbody) {
    CK(pn2_set_remote(g_ctx, (const pn2_remote_node *)rt, nnode, (const pn2_remote_body *)rb, nbody), "pn2_set_remote");
}

void pn2_glue_remote_batch(int kind, const int *s, const int *t, int nt) {
    if (kind == 0) CK(pn2_p2p_ext_batch(g_ctx, s, t, nt), "pn2_p2p_ext_batch");
    else CK(pn2_m2l_ext_batch(g_ctx, s, t, nt), "pn2_m2l_ext_batch");
}

void pn2_glue_end_step(void) {
    CK(pn2_l2l_l2p(g_ctx), "pn2_l2l_l2p");
    CK(pn2_get_acc(g_ctx, part[0].acc, sizeof(Body), NPART, 1), "pn2_get_acc");
}
