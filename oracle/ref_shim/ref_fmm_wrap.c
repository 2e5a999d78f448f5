/*
 * ref_fmm_wrap.c -- compiles the reference's src/fmm.c UNMODIFIED, in place,
 * with pthread_create() redirected through a hook so the (file-static) task
 * lists handed to task_compute_p2p / task_compute_m2l (src/fmm.c:17-27,
 * 381-404, 543-565, 945, 962) can be recorded.  The reference code still runs
 * exactly as written; the hook only copies the batch and forwards the call.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <pthread.h>
static int pn_hook_fmm_create(pthread_t *tid, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
#define pthread_create(a, b, c, d) pn_hook_fmm_create(a, b, c, d)
#include "src/fmm.c"
#undef pthread_create
#include "pn_capture.h"

static int pn_hook_fmm_create(pthread_t *tid_, const pthread_attr_t *attr, void *(*fn)(void *), void *arg) {
    int *par = (int *)arg;
    int c = par[0], nt = par[1];
    if (pn_capture_level >= 1) {
        if (fn == task_compute_p2p) pn_pairs_append(&pn_cap_p2p, task_s[c], task_t[c], nt);
        else if (fn == task_compute_m2l) pn_pairs_append(&pn_cap_m2l, task_s[c], task_t[c], nt);
    }
    return pthread_create(tid_, attr, fn, arg);
}
