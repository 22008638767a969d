/* shim_replay.c -- a C stand-in for the Fortran drop-in shim (fortran/adpres_b200_bind.f90 +
 * fortran/mod_cmfd_b200.f90), for hosts without a Fortran compiler.
 *
 * It performs THE SHIM'S call sequence with Fortran-layout arrays, statement for statement:
 *
 *   gpu_init            adp_create(-1) -> adp_comm_init_env -> adp_set_geometry
 *   outer(popt)         gpu_push_inputs (adp_set_xs_mask: 255 on the first call, 31 afterwards; adp_set_control)
 *                       adp_matrix_setup(1) ; [first: adp_init_flux] ; adp_outer_begin
 *                       do p = 1, nout: adp_outer_iter ; every nupd-th: nodal_upd = adp_nodal_upd + STOP codes ;
 *                                       exit test (ser < serc .and. fer < ferc .and. ndmax < 1.e-2)
 *                       STOP on p-1 == nout
 *                       gpu_pull_results (adp_get_state_mask(3) ; RODEJECT: adp_get_nod + AoS repack into
 *                                         nod(n,g)%df(6), %dn(6) exactly like TYPE NODE_DATA, mod_data.f90:58-62)
 *   PowDis(p)           adp_powdis
 *
 * so that the part of the boundary the tests cannot reach through ctypes -- the shim's upload masks, its
 * first-call logic, the repack loop, its error handling -- runs against the library and is checked against the
 * CPU oracle (tests/test_shim_replay.py).  Reference callers: forward / adjoint (mod_control.f90:21-44,61-80),
 * rod_eject (mod_trans.f90:50-95).
 *
 * usage: shim_replay <spec.bin> <out.bin> [ncalls]
 *   spec.bin  int32 header  nxx nyy nzz nnod ng nmat nout nin nac nupd kern bc[6] kind popt rodeject
 *             float64       serc ferc
 *             int32 arrays  ix iy iz (nnod) ysmin ysmax (nyy) xsmin xsmax (nxx) mat (nnod)
 *             float64 arrays xdel ydel zdel D sigr nuf sigf (nnod,ng) sigs (nnod,ng,ng) chi (nmat,ng)
 *                           dc (nnod,ng,6) exsrc (nnod,ng)            -- all column-major, as sdata holds them
 *   kind      0 outer, 1 outer_fs, 2 outer_ad          ncalls: the outer*() call is made this many times (default 2)
 *   out.bin   int32 niter[ncalls] status ; float64 Ke ser fer ndmax ; f0 (nnod,ng) fs0 (nnod) power (nnod)
 *             nod AoS (12 doubles per (n,g)) if rodeject
 * Build:  gcc -std=c99 -O2 -Wall -Iinclude examples/shim_replay.c -Ladpres_b200 -ladpres_b200 \
 *             -Wl,-rpath,$PWD/adpres_b200 -o /tmp/shim_replay                                          */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "adpres_b200.h"

typedef struct { double df[6], dn[6]; } node_data;   /* TYPE NODE_DATA (mod_data.f90:58-62) */

/* ---- what module sdata holds */
static int nxx, nyy, nzz, nnod, ng, nmat, nout, nin, nac, nupd, kern, bc[6], kind, popt, rodeject;
static double serc, ferc, Ke, ser, fer, ndmax;
static int im, jm, km;
static int *ix, *iy, *iz, *ysmin, *ysmax, *xsmin, *xsmax, *mat;
static double *xdel, *ydel, *zdel, *D, *sigr, *nuf, *sigf, *sigs, *chi, *dc, *exsrc;
static double *f0, *fs0, *s0;
static node_data *nod;                                /* nod(n,g) column-major */
static adp_ctx *ctx;

static void gpu_check(int ierr, const char *what)     /* subroutine gpu_check */
{
    if (ierr >= 0) return;
    fprintf(stderr, " adpres_b200: %s failed, code %d\n %s\n", what, ierr, adp_last_error(ctx));
    exit(10);
}

static void gpu_init(void)                            /* subroutine gpu_init */
{
    int ierr;
    if (ctx) return;
    ierr = adp_create(&ctx, -1);
    if (ierr != 0) { fprintf(stderr, "adpres_b200: no CUDA device (there is no CPU fallback)\n"); exit(11); }
    gpu_check(adp_comm_init_env(ctx), "adp_comm_init_env");
    gpu_check(adp_set_geometry(ctx, nxx, nyy, nzz, nnod, ng, nmat, ix, iy, iz, ysmin, ysmax, xsmin, xsmax, xdel, ydel,
                               zdel, bc, mat), "adp_set_geometry");
}

static void gpu_push_inputs(int with_exsrc)           /* subroutine gpu_push_inputs */
{
    static int first = 1;
    int mask;
    gpu_init();
    if (first) { mask = 255; first = 0; }
    else { mask = 31; if (with_exsrc) mask += 128; }   /* bxtab == 0 here: dc never changes */
    gpu_check(adp_set_xs_mask(ctx, mask, D, sigr, nuf, sigf, sigs, chi, dc, exsrc), "adp_set_xs_mask");
    gpu_check(adp_set_control(ctx, nout, nin, nac, nupd, serc, ferc, kern), "adp_set_control");
}

static void gpu_pull_results(void)                    /* subroutine gpu_pull_results */
{
    gpu_check(adp_get_state_mask(ctx, 3, f0, fs0, s0, &Ke), "adp_get_state_mask");
    if (rodeject) {
        double *df = (double *)malloc(sizeof(double) * 6 * (size_t)nnod * ng), *dn = (double *)malloc(sizeof(double) * 6 * (size_t)nnod * ng);
        gpu_check(adp_get_nod(ctx, df, dn), "adp_get_nod");
        for (int g = 0; g < ng; ++g)
            for (int n = 0; n < nnod; ++n) {
                memcpy(nod[(size_t)g * nnod + n].df, df + 6 * ((size_t)g * nnod + n), 6 * sizeof(double));   /* nod(n,g)%df = df(:,n,g) */
                memcpy(nod[(size_t)g * nnod + n].dn, dn + 6 * ((size_t)g * nnod + n), 6 * sizeof(double));
            }
        free(df); free(dn);
    }
}

static void nodal_upd(int nmode)                      /* subroutine nodal_upd (prints dropped: popt = 0 path) */
{
    int ierr = adp_nodal_upd(ctx, nmode, &ndmax, &im, &jm, &km);
    gpu_check(ierr, "adp_nodal_upd");
    if (ierr == 2) { printf("ERROR IN MATRIX DECOMP: DIAGONAL ELEMENTS CLOSE TO ZERO\n"); exit(2); }
    if (ierr == 3) { printf(" Error: Max. change in nodal coupling coefficient = %10.1f\n", ndmax); exit(3); }
}

/* subroutine outer_common(kind, popt, label, first); returns the iteration count, -1 = the reference's STOP */
static int outer_common(int *first)
{
    int p, mode = ADP_MODE_FORWARD, nmode = 1, init;
    gpu_push_inputs(kind == 1);
    gpu_check(adp_matrix_setup(ctx, 1), "adp_matrix_setup");
    if (kind == 1) mode = ADP_MODE_FIXEDSRC;
    if (kind == 2) { mode = ADP_MODE_ADJOINT; nmode = 0; }
    init = *first;
    if (kind == 2) init = *first && popt > 0;
    if (init) {
        f0 = (double *)calloc((size_t)nnod * ng, sizeof(double));
        fs0 = (double *)calloc((size_t)nnod, sizeof(double));
        s0 = (double *)calloc((size_t)nnod * ng, sizeof(double));
        gpu_check(adp_init_flux(ctx, kind == 2), "adp_init_flux");
        *first = 0;
    }
    gpu_check(adp_outer_begin(ctx, mode), "adp_outer_begin");
    for (p = 1; p <= nout; ++p) {
        gpu_check(adp_outer_iter(ctx, mode, p, &Ke, &ser, &fer), "adp_outer_iter");
        if (p % nupd == 0 && kern != ADP_KERN_FDM && !(kind == 2 && popt <= 0)) nodal_upd(nmode);
        if (ser < serc && fer < ferc && ndmax < (double)1.e-2f) break;
    }
    if (p - 1 == nout) return -1;
    gpu_pull_results();
    return p;
}

static void rd(FILE *f, void *p, size_t bytes)
{
    if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "shim_replay: short spec file\n"); exit(12); }
}
#define IALLOC(p, n) do { p = (int *)malloc(sizeof(int) * (size_t)(n)); rd(f, p, sizeof(int) * (size_t)(n)); } while (0)
#define DALLOC(p, n) do { p = (double *)malloc(sizeof(double) * (size_t)(n)); rd(f, p, sizeof(double) * (size_t)(n)); } while (0)

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: shim_replay <spec.bin> <out.bin> [ncalls]\n"); return 1; }
    const int ncalls = argc > 3 ? atoi(argv[3]) : 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int hdr[20];
    rd(f, hdr, sizeof(hdr));
    nxx = hdr[0]; nyy = hdr[1]; nzz = hdr[2]; nnod = hdr[3]; ng = hdr[4]; nmat = hdr[5]; nout = hdr[6]; nin = hdr[7];
    nac = hdr[8]; nupd = hdr[9]; kern = hdr[10];
    for (int i = 0; i < 6; ++i) bc[i] = hdr[11 + i];
    kind = hdr[17]; popt = hdr[18]; rodeject = hdr[19];
    double tol[2];
    rd(f, tol, sizeof(tol));
    serc = tol[0]; ferc = tol[1];
    IALLOC(ix, nnod); IALLOC(iy, nnod); IALLOC(iz, nnod); IALLOC(ysmin, nyy); IALLOC(ysmax, nyy); IALLOC(xsmin, nxx);
    IALLOC(xsmax, nxx); IALLOC(mat, nnod);
    DALLOC(xdel, nxx); DALLOC(ydel, nyy); DALLOC(zdel, nzz);
    const size_t NG = (size_t)nnod * ng;
    DALLOC(D, NG); DALLOC(sigr, NG); DALLOC(nuf, NG); DALLOC(sigf, NG); DALLOC(sigs, NG * ng); DALLOC(chi, (size_t)nmat * ng);
    DALLOC(dc, NG * 6); DALLOC(exsrc, NG);
    fclose(f);
    nod = (node_data *)calloc(NG, sizeof(node_data));

    int first = 1, status = 0;
    int *niter = (int *)calloc((size_t)ncalls, sizeof(int));
    for (int call = 0; call < ncalls && status == 0; ++call) {
        niter[call] = outer_common(&first);                   /* CALL outer(popt) */
        if (niter[call] < 0) {
            printf("  MAXIMUM NUMBER OF OUTER ITERATION IS REACHED\n");
            status = 1;
            gpu_pull_results();
        }
        printf("shim_replay: call %d: %d outer iterations, k-eff %.6f, ser %.3e, fer %.3e\n", call + 1, niter[call], Ke, ser, fer);
    }
    double *pw = (double *)calloc((size_t)nnod, sizeof(double));
    int ierr = adp_powdis(ctx, pw, kind == 1);                /* CALL PowDis(pow) */
    gpu_check(ierr, "adp_powdis");
    if (ierr == 4) { printf("   ERROR: TOTAL NODES POWER IS ZERO OR LESS\n"); status = 4; }

    f = fopen(argv[2], "wb");
    if (!f) { perror(argv[2]); return 1; }
    fwrite(niter, sizeof(int), (size_t)ncalls, f);
    fwrite(&status, sizeof(int), 1, f);
    double sc[4] = {Ke, ser, fer, ndmax};
    fwrite(sc, sizeof(double), 4, f);
    fwrite(f0, sizeof(double), NG, f);
    fwrite(fs0, sizeof(double), (size_t)nnod, f);
    fwrite(pw, sizeof(double), (size_t)nnod, f);
    if (rodeject) fwrite(nod, sizeof(node_data), NG, f);
    fclose(f);
    adp_destroy(ctx);
    return 0;
}
