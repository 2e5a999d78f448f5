/*
 * ref_remotes_gpu_wrap.c -- the reference's src/remotes.c, UNMODIFIED and compiled in place, with
 * pthread_create(task_compute_p2p_ext / task_compute_m2l_ext) (src/remotes.c:201, 387) redirected to the device.
 * A newly received LET (src/remotes.c:740-746) is detected through the MPI shim's receive counter for tag 111.
 */
#include <pthread.h>
static int pn2_hook_rem_create(pthread_t *tid, const pthread_attr_t *attr, void *(*fn)(void *), void *arg);
#define pthread_create(a, b, c, d) pn2_hook_rem_create(a, b, c, d)
#include "src/remotes.c"
#undef pthread_create
#include "pn2_fmm_glue.h"

static void *pn2_noop_r(void *a) { return a; }
static long pn2_last_seq = -1;

static int pn2_hook_rem_create(pthread_t *tid_, const pthread_attr_t *attr, void *(*fn)(void *), void *arg) {
    int *par = (int *)arg;
    int c = par[0], nt = par[1];
    if (fn != task_compute_p2p_ext && fn != task_compute_m2l_ext) return pthread_create(tid_, attr, fn, arg);
    long seq = pn_shim_recv_seq[111];
    if (seq != pn2_last_seq) {
        pn2_last_seq = seq;
        pn2_glue_set_remote(exrtree, (int)(pn_shim_recv_bytes[111] / (long)sizeof(RemoteNode)), exrbody,
                            (int)(pn_shim_recv_bytes[112] / (long)sizeof(RemoteBody)));
    }
    pn2_glue_remote_batch(fn == task_compute_p2p_ext ? 0 : 1, task_s_ex[c], task_t_ex[c], nt);
    return pthread_create(tid_, attr, pn2_noop_r, NULL);
}
