/*
 * ref_fmm_gpu_wrap.c -- the reference's src/fmm.c, UNMODIFIED and compiled in place, with the worker-thread hand-off
 * (pthread_create(task_compute_p2p / task_compute_m2l), src/fmm.c:394, 555, 945, 962) redirected to the device through
 * photons-2.0_b200/host/pn2_fmm_glue.h.  This is the drop-in demonstration: the reference's own tree build, dual-tree
 * walks and driver, the B200 kernels in place of its CPU task evaluators.  Built only where /root/reference exists.
 */
#include <pthread.h>
static int pn2_hook_fmm_create(pthread_t *tid, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
#define pthread_create(a, b, c, d) pn2_hook_fmm_create(a, b, c, d)
#include "src/fmm.c"
#undef pthread_create
#include "pn2_fmm_glue.h"

static void *pn2_noop(void *a) { return a; }

static int pn2_hook_fmm_create(pthread_t *tid_, const pthread_attr_t *attr, void *(*fn)(void *), void *arg) {
    int *par = (int *)arg;
    int c = par[0], nt = par[1];
    if (fn == task_compute_p2p) pn2_glue_local_batch(0, task_s[c], task_t[c], nt);
    else if (fn == task_compute_m2l) pn2_glue_local_batch(1, task_s[c], task_t[c], nt);
    else return pthread_create(tid_, attr, fn, arg);
    return pthread_create(tid_, attr, pn2_noop, NULL);      /* the caller joins tid (src/fmm.c:386) */
}
