// pn2_modeb.cu -- Mode B (device tree, device lists, images, LET): placeholder until the builder lands.
#include "pn2_common.cuh"
void pn2_modeb_release(pn2_ctx *h) { (void)h; }
#define NOTYET(name) { pn2_set_error(name ": Mode B is not built yet"); return PN2_ERR_STATE; }
extern "C" int pn2_force_step(pn2_ctx *, const double *, size_t, int, const pn2_domain *, double *, size_t) NOTYET("pn2_force_step")
extern "C" int pn2_force_step_device(pn2_ctx *, const double *, int, const pn2_domain *, double *) NOTYET("pn2_force_step_device")
extern "C" int pn2_set_comm(pn2_ctx *, int, int, const pn2_domain *, void *) NOTYET("pn2_set_comm")
extern "C" int pn2_get_step_info(pn2_ctx *, pn2_step_info *) NOTYET("pn2_get_step_info")
extern "C" int pn2_get_order(pn2_ctx *, int *, int) NOTYET("pn2_get_order")
extern "C" int pn2_get_cells(pn2_ctx *, double *, int *, int *, double *, double *) NOTYET("pn2_get_cells")
extern "C" int pn2_get_lists(pn2_ctx *, int, long *, long *, int *, long *, int *) NOTYET("pn2_get_lists")
extern "C" int pn2_get_timings(pn2_ctx *, double *) NOTYET("pn2_get_timings")
