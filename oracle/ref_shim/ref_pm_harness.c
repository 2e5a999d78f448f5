/*
 * ref_pm_harness.c -- drives the UNMODIFIED reference particle-mesh force (src/partmesh.c:18-796,
 * partmesh_thread: CIC deposit, mesh exchange, 4-point gradient, CIC gather) on one rank and dumps what the
 * parity tests of the device PM path (photons-2.0_b200/csrc/pn2_pm.cu) need.
 * TEST INFRASTRUCTURE ONLY: never linked into, imported by or executed from the product path.
 *
 * The convolution itself (src/conv.f90:128-247, subroutine convolution) is Fortran on top of 2DECOMP&FFT, an
 * un-vendored third-party library that is absent from /root/reference (only inc/decomp_2d*.mod; mytype = 8, i.e.
 * double precision): it cannot be built here.  convolution_() below RESTATES it in C -- same Green function,
 * expression by expression, including the single-precision literal M_PI = 3.1415926 (conv.f90:143) -- around a
 * plain complex FFT.  PARITY UNPINNED for that one function: nothing in this container can run the original.
 * Everything else on the path (deposit, exchange, gradient, gather) IS the reference's own code.
 *
 * One rank only (NP = 1): the pencil exchange degenerates to a local copy (MSIZE0 = MSIZE1 = {NSIDE - 1}); the
 * PM force does not depend on how the particles are partitioned, so the multi-rank device path is compared with
 * this result as well.
 *
 * Usage: ref_pm <params.txt> <pos.f64> <out.bin>      params: NPART_TOTAL BOXSIZE NSIDE MASS SPLIT
 *   output records { char name[24]; long nbytes; payload }: "acc_pm" [n][3], "density" / "potential" [NSIDE^3]
 *   (the convolution's input and output arrays), "timing" [1] (seconds of partmesh_thread)
 */
#include "photoNs.h"
#include "partmesh.h"
#include <math.h>
#include <string.h>
#include <sys/time.h>

static double *cap_density = NULL, *cap_potential = NULL;

void get_local_size_(int nside[3], int vp[2], int start[3], int end[3], int size[3]) {
    (void)vp;
    for (int d = 0; d < 3; d++) { start[d] = 0; end[d] = nside[d] - 1; size[d] = nside[d]; }
}

/* ---- plain complex FFT: radix-2 splitting while the length is even, direct DFT for the odd remainder ---- */
typedef struct { double re, im; } cplx;
static void fft_rec(cplx *x, int n, int stride, cplx *out, int sign) {
    if (n == 1) { out[0] = x[0]; return; }
    if (n % 2) {
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0;
            for (int j = 0; j < n; j++) {
                double a = sign * 2.0 * M_PI * (double)(((long)j * k) % n) / n;
                double c = cos(a), s = sin(a);
                sr += x[j * stride].re * c - x[j * stride].im * s;
                si += x[j * stride].re * s + x[j * stride].im * c;
            }
            out[k].re = sr; out[k].im = si;
        }
        return;
    }
    int h = n / 2;
    fft_rec(x, h, 2 * stride, out, sign);
    fft_rec(x + stride, h, 2 * stride, out + h, sign);
    for (int k = 0; k < h; k++) {
        double a = sign * 2.0 * M_PI * k / n;
        double c = cos(a), s = sin(a);
        cplx e = out[k], o = out[k + h];
        double tr = o.re * c - o.im * s, ti = o.re * s + o.im * c;
        out[k].re = e.re + tr; out[k].im = e.im + ti;
        out[k + h].re = e.re - tr; out[k + h].im = e.im - ti;
    }
}
static void fft3d(cplx *a, int n, int sign) {
    cplx *line = (cplx *)malloc(sizeof(cplx) * n), *res = (cplx *)malloc(sizeof(cplx) * n);
    long s[3] = { (long)n * n, n, 1 };
    for (int ax = 0; ax < 3; ax++) {
        int b = (ax + 1) % 3, c = (ax + 2) % 3;
        for (int p = 0; p < n; p++)
            for (int q = 0; q < n; q++) {
                cplx *base = a + p * s[b] + q * s[c];
                for (int k = 0; k < n; k++) line[k] = base[k * s[ax]];
                fft_rec(line, n, 1, res, sign);
                for (int k = 0; k < n; k++) base[k * s[ax]] = res[k];
            }
    }
    free(line); free(res);
}

/* subroutine convolution(data, nside, param), src/conv.f90:128-247, restated: forward FFT (unnormalised, as
 * decomp_2d_fft_3d), multiply by gf, backward FFT (unnormalised; the 1/N^3 sits in pref).  data is indexed
 * (i * N + j) * N + k  <->  in(i, j, k) (:153-161), and (l, m, n) are the wave numbers of (i, j, k) (:178-206). */
void convolution_(double *data, int *nside, double *param) {
    const int N = nside[0];
    const long N3 = (long)N * N * N;
    if (!cap_density) cap_density = (double *)malloc(sizeof(double) * N3);
    memcpy(cap_density, data, sizeof(double) * N3);
    cplx *a = (cplx *)malloc(sizeof(cplx) * N3);
    for (long q = 0; q < N3; q++) { a[q].re = data[q]; a[q].im = 0.0; }
    fft3d(a, N, -1);
    const double PI_F = (double)3.1415926f;           /* real*8, parameter :: M_PI = 3.1415926 -- a default-real literal (:143) */
    const int nhalf = N / 2;                          /* :172 */
    const double smooth = param[0], box = param[1];
    double ismth2 = 2 * PI_F * smooth / box;          /* :176 */
    ismth2 = ismth2 * ismth2;
    const double pref = box * box / (PI_F * nside[0] * nside[1] * nside[2]);      /* :178 */
    for (int i = 0; i < N; i++) {
        int l = i; if (l > nhalf) l -= N;
        double fx = PI_F * l / nside[0]; fx = sin(fx) / fx; if (l == 0) fx = 1;    /* :207-211 */
        for (int j = 0; j < N; j++) {
            int m = j; if (m > nhalf) m -= N;
            double fy = PI_F * m / nside[1]; fy = sin(fy) / fy; if (m == 0) fy = 1;  /* :192-197 */
            for (int k = 0; k < N; k++) {
                int n = k; if (n > nhalf) n -= N;
                double fz = PI_F * n / nside[2]; fz = sin(fz) / fz; if (n == 0) fz = 1;  /* :184-189 */
                double k2 = (double)(float)(n * n + m * m + l * l);               /* REAL(...) :213 */
                double ff = 1.0 / (fx * fy * fz);                                  /* :215 */
                double gf = pref * exp(-k2 * ismth2) * ff * ff * ff * ff / k2;     /* :216 */
                if (l == 0 && m == 0 && n == 0) gf = pref;                         /* :218-220 */
                cplx *z = &a[((long)i * N + j) * N + k];
                z->re *= gf; z->im *= gf;
            }
        }
    }
    fft3d(a, N, +1);
    for (long q = 0; q < N3; q++) data[q] = a[q].re;
    free(a);
    if (!cap_potential) cap_potential = (double *)malloc(sizeof(double) * N3);
    memcpy(cap_potential, data, sizeof(double) * N3);
}

static void rec(FILE *f, const char *name, const void *data, long nbytes) {
    char nm[24];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 23);
    fwrite(nm, 1, 24, f);
    fwrite(&nbytes, sizeof(long), 1, f);
    if (nbytes > 0) fwrite(data, 1, nbytes, f);
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_pm params.txt pos.f64 out.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "r");
    if (!f) return 2;
    char key[128];
    double v, split = -1;
    long ntot = 0;
    while (fscanf(f, "%127s %lf", key, &v) == 2) {
        if (!strcmp(key, "NPART_TOTAL")) ntot = (long)v;
        else if (!strcmp(key, "BOXSIZE")) BOXSIZE = v;
        else if (!strcmp(key, "NSIDE")) NSIDE = (int)v;
        else if (!strcmp(key, "MASS")) MASSPART = v;
        else if (!strcmp(key, "SPLIT")) split = v;
    }
    fclose(f);
    splitRadius = 1.25 * (BOXSIZE / ((double)NSIDE));            /* src/initial.c:316-345 */
    if (split > 0.0) splitRadius = split;
    double *pos = (double *)malloc(sizeof(double) * 3 * ntot);
    f = fopen(argv[2], "rb");
    if (!f || fread(pos, sizeof(double), 3 * ntot, f) != (size_t)(3 * ntot)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 2; }
    fclose(f);

    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &PROC_RANK);
    MPI_Comm_size(MPI_COMM_WORLD, &PROC_SIZE);
    if (PROC_SIZE != 1) { fprintf(stderr, "ref_pm runs on one rank\n"); return 2; }
    PM_COMM_WORLD = MPI_COMM_WORLD;
    MPI_Type_contiguous(sizeof(MKey), MPI_CHAR, &strMKey);       /* src/initial.c:233 */
    /* the mesh decomposition of src/initial.c:245-246, 453-493 for one rank */
    vproc[0] = PROC_SIZE; vproc[1] = 1;
    pside[0] = NSIDE / vproc[0]; pside[1] = NSIDE / vproc[1];
    int nside[3] = { NSIDE, NSIDE, NSIDE };
    get_local_size_(nside, vproc, local_xstart, local_xend, local_xsize);
    MSIZE0 = (int *)malloc(sizeof(int)); MSIZE1 = (int *)malloc(sizeof(int));
    MSIZE0[0] = local_xend[1]; MSIZE1[0] = local_xend[2];
    data_length = (long)local_xsize[0] * local_xsize[1] * local_xsize[2];
    NPART = (int)ntot; NPART_TOTAL = ntot;
    reset_mem();
    part = (Body *)pmalloc(sizeof(Body) * (NPART > 0 ? NPART : 1), 0);
    for (long n = 0; n < ntot; n++) {
        part[n].pos[0] = pos[3 * n]; part[n].pos[1] = pos[3 * n + 1]; part[n].pos[2] = pos[3 * n + 2];
    }
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    partmesh_thread();
    gettimeofday(&t1, NULL);
    double sec = (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_usec - t0.tv_usec);

    double *acc = (double *)malloc(sizeof(double) * 3 * (ntot > 0 ? ntot : 1));
    for (long n = 0; n < ntot; n++) { acc[3 * n] = part[n].acc_pm[0]; acc[3 * n + 1] = part[n].acc_pm[1]; acc[3 * n + 2] = part[n].acc_pm[2]; }
    f = fopen(argv[3], "wb");
    if (!f) return 2;
    long N3 = (long)NSIDE * NSIDE * NSIDE;
    rec(f, "acc_pm", acc, (long)sizeof(double) * 3 * ntot);
    rec(f, "density", cap_density, (long)sizeof(double) * N3);
    rec(f, "potential", cap_potential, (long)sizeof(double) * N3);
    rec(f, "timing", &sec, sizeof sec);
    fclose(f);
    printf("[ref_pm] N=%ld NSIDE=%d  partmesh_thread %.3f s\n", ntot, NSIDE, sec);
    MPI_Finalize();
    return 0;
}
