/*
 * pn_capture.h -- capture buffers shared by the reference wrappers and the harness.
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 */
#ifndef PN_CAPTURE_H
#define PN_CAPTURE_H
#include <stdlib.h>
#include <string.h>

typedef struct { int *s, *t; long n, cap; } PnPairList;

typedef struct {
    long seq;          /* 1-based index of the fmm_remote() call that received this tree */
    int nnode, nbody;
    void *tree;        /* RemoteNode[nnode] (224 B each) */
    void *body;        /* RemoteBody[nbody] (32 B each)  */
    PnPairList p2p, m2l;
} PnRemoteCap;

extern int pn_capture_level;          /* 0 none, 1 local lists, 2 + remote trees and lists */
extern PnPairList pn_cap_p2p, pn_cap_m2l;
extern PnRemoteCap *pn_rcap;
extern long pn_nrcap, pn_rcap_cap;

static inline void pn_pairs_append(PnPairList *L, const int *s, const int *t, long n) {
    if (L->n + n > L->cap) {
        long nc = L->cap ? L->cap : 1 << 16;
        while (nc < L->n + n) nc *= 2;
        L->s = (int *)realloc(L->s, nc * sizeof(int));
        L->t = (int *)realloc(L->t, nc * sizeof(int));
        L->cap = nc;
    }
    memcpy(L->s + L->n, s, n * sizeof(int));
    memcpy(L->t + L->n, t, n * sizeof(int));
    L->n += n;
}
#endif
