/*
 * ref_remotes_wrap.c -- compiles the reference's src/remotes.c UNMODIFIED, in
 * place, with pthread_create() redirected through a hook that records, per
 * fmm_remote() call (src/remotes.c:684-751), the received pruned tree
 * (exrtree/exrbody) and the (source, target) lists handed to
 * task_compute_p2p_ext / task_compute_m2l_ext (src/remotes.c:583-628).
 * TEST INFRASTRUCTURE ONLY.
 */
#include <pthread.h>
static int pn_hook_rem_create(pthread_t *tid, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
#define pthread_create(a, b, c, d) pn_hook_rem_create(a, b, c, d)
#include "src/remotes.c"
#undef pthread_create
#include "pn_capture.h"

static int pn_hook_rem_create(pthread_t *tid_, const pthread_attr_t *attr, void *(*fn)(void *), void *arg) {
    int *par = (int *)arg;
    int c = par[0], nt = par[1];
    if (pn_capture_level >= 2) {
        long seq = pn_shim_recv_seq[111];
        if (pn_nrcap == 0 || pn_rcap[pn_nrcap - 1].seq != seq) {
            if (pn_nrcap == pn_rcap_cap) {
                pn_rcap_cap = pn_rcap_cap ? 2 * pn_rcap_cap : 64;
                pn_rcap = (PnRemoteCap *)realloc(pn_rcap, pn_rcap_cap * sizeof(PnRemoteCap));
            }
            PnRemoteCap *r = &pn_rcap[pn_nrcap++];
            memset(r, 0, sizeof *r);
            r->seq = seq;
            r->nnode = (int)(pn_shim_recv_bytes[111] / (long)sizeof(RemoteNode));
            r->nbody = (int)(pn_shim_recv_bytes[112] / (long)sizeof(RemoteBody));
            r->tree = malloc(sizeof(RemoteNode) * (r->nnode + 1));
            r->body = malloc(sizeof(RemoteBody) * (r->nbody + 1));
            memcpy(r->tree, exrtree, sizeof(RemoteNode) * r->nnode);
            memcpy(r->body, exrbody, sizeof(RemoteBody) * r->nbody);
        }
        PnRemoteCap *r = &pn_rcap[pn_nrcap - 1];
        if (fn == task_compute_p2p_ext) pn_pairs_append(&r->p2p, task_s_ex[c], task_t_ex[c], nt);
        else if (fn == task_compute_m2l_ext) pn_pairs_append(&r->m2l, task_s_ex[c], task_t_ex[c], nt);
    }
    return pthread_create(tid_, attr, fn, arg);
}
