/* force_step.c -- the C-ABI from plain C: one Mode B short-range force evaluation (device tree, device lists, all
 * operators) of positions read from a raw file of doubles, accelerations written to another.  This is the call the
 * reference's driver makes per step (src/photoNs.c:97-116: fmm_construct .. fmm_ext, PM excluded).
 *
 *   gcc -O2 -I../include force_step.c -L../photons-2.0_b200 -lpn2gpu -Wl,-rpath,'$ORIGIN/../photons-2.0_b200' -lm
 *   ./force_step pos.f64 acc.f64 <n> <box> <nside> <mass> <fp64|fp32>
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pn2gpu.h"

static void die(const char *what) {
    fprintf(stderr, "force_step: %s: %s\n", what, pn2_last_error());
    exit(1);
}

int main(int argc, char **argv) {
    if (argc != 8) { fprintf(stderr, "usage: %s pos.f64 acc.f64 n box nside mass fp64|fp32\n", argv[0]); return 2; }
    const int n = atoi(argv[3]);
    const double box = atof(argv[4]), nside = atof(argv[5]), mass = atof(argv[6]);
    double *pos = malloc(24 * (size_t)n), *acc = malloc(24 * (size_t)n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(pos, 24, (size_t)n, f) != (size_t)n) { fprintf(stderr, "force_step: cannot read %s\n", argv[1]); return 1; }
    fclose(f);

    pn2_params prm;                                       /* src/initial.c:316-345 */
    memset(&prm, 0, sizeof prm);
    prm.box = box;
    prm.rs = 1.25 * (box / nside);
    prm.cutoff = 4.5 * prm.rs;
    prm.soft = 0.03 * box / pow((double)n, 0.3333333);
    prm.theta = 0.4;
    prm.mass = mass;
    prm.maxleaf = 8;
    prm.periodic = 1;
    prm.longshort = 1;
    prm.precision = strcmp(argv[7], "fp64") == 0 ? PN2_FP64 : PN2_FP32;

    pn2_ctx *h = NULL;
    if (pn2_create(&h, 0, &prm) != PN2_OK) die("pn2_create");
    pn2_domain dom;
    memset(&dom, 0, sizeof dom);
    for (int d = 0; d < 3; d++) { dom.lo[d] = 0.0; dom.hi[d] = box; }
    dom.direct0 = 0;
    if (pn2_force_step(h, pos, 24, n, &dom, acc, 24) != PN2_OK) die("pn2_force_step");
    pn2_step_info info;
    if (pn2_get_step_info(h, &info) != PN2_OK) die("pn2_get_step_info");
    double s2 = 0.0;
    for (size_t i = 0; i < 3 * (size_t)n; i++) s2 += acc[i] * acc[i];
    printf("n %d leaves %d nodes %d interactions %lld m2l_pairs %lld rms_acc %.17g\n", n, info.nleaf, info.nnode,
           (long long)info.n_interactions, (long long)info.n_m2l_pairs, sqrt(s2 / n));
    f = fopen(argv[2], "wb");
    if (!f || fwrite(acc, 24, (size_t)n, f) != (size_t)n) { fprintf(stderr, "force_step: cannot write %s\n", argv[2]); return 1; }
    fclose(f);
    pn2_destroy(h);
    free(pos); free(acc);
    return 0;
}
