/*
 * pn2_fmm_glue.h -- the drop-in glue a photoNs-2.0 maintainer adds to route the short-range task batches to
 * libpn2gpu.so (Mode A).  Host code in the reference's own language (C), above the C-ABI of include/pn2gpu.h.
 *
 * Call sites in the reference (see INTEGRATION.md for the patch):
 *   src/photoNs.c:97-116 / 231-252   fmm_prepare(); pn2_glue_begin_step(); fmm_task(); fmm_ext(); pn2_glue_end_step();
 *   src/fmm.c:394, 555               pthread_create(task_compute_p2p / task_compute_m2l)      -> pn2_glue_local_batch()
 *   src/remotes.c:201, 387           pthread_create(task_compute_p2p_ext / _m2l_ext)          -> pn2_glue_remote_batch()
 *   src/remotes.c:740-746            after the LET receive                                    -> pn2_glue_set_remote()
 */
#ifndef PN2_FMM_GLUE_H
#define PN2_FMM_GLUE_H
#include "pn2gpu.h"

#ifdef __cplusplus
extern "C" {
#endif
/* after fmm_prepare(): upload part[], leaf[], btree[]; P2M + M2M on the device (the host keeps its own M for the LET) */
void pn2_glue_begin_step(void);
/* one task batch of fmm_task(): kind 0 = task_compute_p2p (src/fmm.c:796), 1 = task_compute_m2l (src/fmm.c:875) */
void pn2_glue_local_batch(int kind, const int *task_s, const int *task_t, int nt);
/* a LET just received by fmm_remote(): exrtree[0..nnode), exrbody[0..nbody) */
void pn2_glue_set_remote(const void *exrtree, int nnode, const void *exrbody, int nbody);
/* one task batch of fmm_remote_task(): kind 0 = task_compute_p2p_ext, 1 = task_compute_m2l_ext (src/remotes.c:583, 598) */
void pn2_glue_remote_batch(int kind, const int *task_s, const int *task_t, int nt);
/* after fmm_ext(): L2L + L2P on the device, accelerations added into part[].acc */
void pn2_glue_end_step(void);
#ifdef __cplusplus
}
#endif
#endif
