/*
 * adpres_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A serial, fp64, plain-C restatement of the ADPRES 1.2 eigenvalue hot path:
 *   src/mod_cmfd.f90  (CMFD matrix, BiCGSTAB, sources, outer iterations, PowDis)
 *   src/mod_nodal.f90 (SANM / PNM two-node nodal coupling-coefficient update)
 * plus the few transient helpers of src/mod_trans.f90 that surround outer_tr (both bxtab branches).
 * Every function cites the reference file:line it follows; loop order, operation
 * order, data structures (ragged matrix rows A(n,g)%elmn, ind(n)%col, AoS nod{df,dn})
 * and the reference's quirks are kept on purpose.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The CUDA product (adpres_b200/csrc) never does.
 *
 * Parity pins (reference-produced numbers this oracle reproduces):
 *   1. the IAEA3Ds terminal trace of docs/quick-guides.md:161-191, every printed digit (k-eff
 *      1.029082, 129 outer iterations, first nodal update MAX. CHANGE = 3.16843E-01) --
 *      tests/test_oracle_golden.py;
 *   2. the six NEACRP critical boron concentrations on the %BCON cards of smpl/transient/NEACRP/*t
 *      (with oracle/th.py and the feedback XS update of the harness) -- tests/test_th.py;
 *   3. the two MOX/UO2 part-3 critical boron concentrations on the %BCON cards of
 *      smpl/transient/MOX/part4_* (through the %XTAB branch tables) -- tests/test_xtab.py.
 * Adjoint, fixed-source, PNM, multigroup > 2 and transient results (outer_tr, get_exsrc incl. its
 * bxtab = 1 branch, iPden, uPden) are NOT pinned by the reference ("parity unpinned" for those);
 * they are checked against published benchmark solutions for plausibility only.
 *
 * Build: gcc -O3 -ffp-contract=off (gfortran -O4 on x86-64 without -march does not
 * contract a*b+c into FMA either), see oracle/Makefile.
 *
 * Indexing: node ids, mesh indices, group and face numbers are 1-based like the Fortran;
 * helper macros map them to 0-based C storage with the Fortran column-major layout.
 */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NF 6 /* delayed neutron precursor families, mod_data.f90:120 */

typedef struct { double df[6], dn[6]; } node_data;      /* mod_data.f90:58-62 */
typedef struct { double *elmn; } fdm_matr;              /* mod_data.f90:46-49 */
typedef struct { int ncol; int *col; } fdm_ind;         /* mod_data.f90:50-54 */

/* status codes = the reference's STOP conditions */
enum {
    ORC_OK = 0,
    ORC_ERR_MAXOUTER = 1,   /* mod_cmfd.f90:498-505, 589-596, 688-695 */
    ORC_ERR_LU_DIAG = 2,    /* mod_nodal.f90:856-862 */
    ORC_ERR_NDMAX = 3,      /* mod_nodal.f90:131-142 */
    ORC_ERR_ZERO_POWER = 4, /* mod_cmfd.f90:1322-1326 */
    ORC_ERR_TH_NOUPD = 5    /* mod_cmfd.f90:788-794 */
};

#define MAXTRACE 200000

typedef struct orc {
    /* ---- sdata: sizes, geometry */
    int ng, nmat, nxx, nyy, nzz, nnod;
    int *ix, *iy, *iz, *mat, *xyz;
    int *ystag_smin, *ystag_smax, *xstag_smin, *xstag_smax;
    double *xdel, *ydel, *zdel, *vdel;
    int xeast, xwest, ynorth, ysouth, zbott, ztop;
    /* ---- sdata: node-wise XS */
    double *D, *sigr, *nuf, *sigf, *sigs, *chi, *dc, *exsrc;
    /* ---- sdata: matrix, state */
    fdm_matr *A;
    fdm_ind *ind;
    node_data *nod;
    double *f0, *fs0, *s0;
    double Ke, ser, fer, ndmax;
    int im, jm, km;
    /* ---- sdata: control */
    int nout, nin, nac, nupd, kern, nth;
    double serc, ferc;
    int fixedsrc_mode;          /* mode == 'FIXEDSRC' (PowDis) */
    /* ---- sdata: transient */
    double lamb[NF], ibeta[NF], *velo, *tbeta;
    /* %XTAB decks (bxtab == 1): kinetics data per material, m(mat)%lamb / %iBeta / %velo */
    int bxtab;
    double *mlamb, *mibeta, *mvelo;      /* (NF, nmat), (NF, nmat), (ng, nmat) column-major */
    double *c0, *ft, *fst, *omeg, *sigrp, *L, *dfis;
    double sth, bth, ht_cur;
    /* ---- SAVEd first-call flags */
    int coup_first, matrix_first, outer_first, have_state;
    /* ---- module nodal */
    int cmode;
    double *Bcn, *Bcp, *An, *Bn, *En, *Fn, *Gn, *Hn, *Ap, *Bp, *Ep, *Fp, *Gp, *Hp, *Lm2;
    double *S1, *S2, *S3;
    double *a1n, *a2n, *a3n, *a4n, *a1p, *a2p, *a3p, *a4p, *Ln1, *Lp1;
    int status;
    /* ---- timers (mod_data.f90:204-217) */
    double fdm_time, nod_time;
    /* ---- trace (what the reference prints per outer iteration) */
    int ntrace;
    double *tr_ke, *tr_ser, *tr_fer;
    double *tr_wall, *tr_fdm, *tr_nod; /* per outer iteration: wall clock and the two accumulators at its end (bench.py) */
    int nnodal;               /* nodal updates recorded */
    int *nu_p, *nu_im, *nu_jm, *nu_km;
    double *nu_ndmax;
    int nextrp, *ex_p;
    int cur_p;                /* outer iteration currently running (for the trace) */
} orc;

/* ---- Fortran-style accessors (1-based) */
#define N_ (o->nnod)
#define G_ (o->ng)
#define V2(a, n, g) ((a)[((size_t)(g) - 1) * N_ + ((n) - 1)])                      /* a(n,g)   */
#define SIGS(n, g, h) (o->sigs[(((size_t)(h) - 1) * G_ + ((g) - 1)) * N_ + ((n) - 1)]) /* sigs(n,g,h) */
#define DC(n, g, f) (o->dc[(((size_t)(f) - 1) * G_ + ((g) - 1)) * N_ + ((n) - 1)])  /* dc(n,g,f) */
#define CHI(m, g) (o->chi[((size_t)(g) - 1) * o->nmat + ((m) - 1)])                /* chi(mat,g) */
#define NOD(n, g) (o->nod[((size_t)(g) - 1) * N_ + ((n) - 1)])
#define AM(n, g) (o->A[((size_t)(g) - 1) * N_ + ((n) - 1)])
#define IND(n) (o->ind[(n) - 1])
#define XYZ(i, j, k) (o->xyz[(((size_t)(k) - 1) * o->nyy + ((j) - 1)) * o->nxx + ((i) - 1)])
#define IX(n) (o->ix[(n) - 1])
#define IY(n) (o->iy[(n) - 1])
#define IZ(n) (o->iz[(n) - 1])
#define MAT(n) (o->mat[(n) - 1])
#define XDEL(i) (o->xdel[(i) - 1])
#define YDEL(j) (o->ydel[(j) - 1])
#define ZDEL(k) (o->zdel[(k) - 1])
#define VDEL(n) (o->vdel[(n) - 1])
#define YSMIN(j) (o->ystag_smin[(j) - 1])
#define YSMAX(j) (o->ystag_smax[(j) - 1])
#define XSMIN(i) (o->xstag_smin[(i) - 1])
#define XSMAX(i) (o->xstag_smax[(i) - 1])
#define C0(n, i) (o->c0[((size_t)(i) - 1) * N_ + ((n) - 1)])
#define BC_(a, g, h) ((a)[((size_t)(h) - 1) * G_ + ((g) - 1)])                    /* B(g,h) */

static double get_time(void) /* cpu_time, mod_data.f90:209-217 */
{
    struct timespec ts;
    clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static double *dalloc(size_t n)
{
    double *p = (double *)calloc(n ? n : 1, sizeof(double));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static int *ialloc(size_t n)
{
    int *p = (int *)calloc(n ? n : 1, sizeof(int));
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}
static double *ddup(const double *s, size_t n) { double *p = dalloc(n); memcpy(p, s, n * sizeof(double)); return p; }
static int *idup(const int *s, size_t n) { int *p = ialloc(n); memcpy(p, s, n * sizeof(int)); return p; }

/* ======================================================================== lifecycle */
orc *orc_create(void)
{
    orc *o = (orc *)calloc(1, sizeof(orc));
    o->nout = 500; o->nin = 2; o->nac = 5; o->serc = 1e-5; o->ferc = 1e-5; /* mod_data.f90:78-83 */
    o->nth = 20; o->kern = 2;
    o->sth = 1.0; o->bth = 0.0;                                            /* mod_data.f90:121 */
    o->coup_first = o->matrix_first = o->outer_first = 1;
    o->ndmax = 0.0; /* never initialised by the reference; zero under static storage */
    o->tr_ke = dalloc(MAXTRACE); o->tr_ser = dalloc(MAXTRACE); o->tr_fer = dalloc(MAXTRACE);
    o->tr_wall = dalloc(MAXTRACE); o->tr_fdm = dalloc(MAXTRACE); o->tr_nod = dalloc(MAXTRACE);
    o->nu_p = ialloc(MAXTRACE); o->nu_im = ialloc(MAXTRACE); o->nu_jm = ialloc(MAXTRACE);
    o->nu_km = ialloc(MAXTRACE); o->nu_ndmax = dalloc(MAXTRACE); o->ex_p = ialloc(MAXTRACE);
    return o;
}

void orc_destroy(orc *o)
{
    if (!o) return;
    if (o->A) {
        for (size_t r = 0; r < (size_t)N_ * G_; ++r) free(o->A[r].elmn);
        free(o->A);
    }
    if (o->ind) { for (int n = 0; n < N_; ++n) free(o->ind[n].col); free(o->ind); }
    free(o->ix); free(o->iy); free(o->iz); free(o->mat); free(o->xyz);
    free(o->ystag_smin); free(o->ystag_smax); free(o->xstag_smin); free(o->xstag_smax);
    free(o->xdel); free(o->ydel); free(o->zdel); free(o->vdel);
    free(o->D); free(o->sigr); free(o->nuf); free(o->sigf); free(o->sigs); free(o->chi);
    free(o->dc); free(o->exsrc); free(o->nod); free(o->f0); free(o->fs0); free(o->s0);
    free(o->mlamb); free(o->mibeta); free(o->mvelo);
    free(o->velo); free(o->tbeta); free(o->c0); free(o->ft); free(o->fst); free(o->omeg);
    free(o->sigrp); free(o->L); free(o->dfis);
    free(o->Bcn); free(o->Bcp); free(o->An); free(o->Bn); free(o->En); free(o->Fn); free(o->Gn);
    free(o->Hn); free(o->Ap); free(o->Bp); free(o->Ep); free(o->Fp); free(o->Gp); free(o->Hp);
    free(o->Lm2); free(o->S1); free(o->S2); free(o->S3);
    free(o->a1n); free(o->a2n); free(o->a3n); free(o->a4n); free(o->a1p); free(o->a2p);
    free(o->a3p); free(o->a4p); free(o->Ln1); free(o->Lp1);
    free(o->tr_wall); free(o->tr_fdm); free(o->tr_nod);
    free(o->tr_ke); free(o->tr_ser); free(o->tr_fer); free(o->nu_p); free(o->nu_im);
    free(o->nu_jm); free(o->nu_km); free(o->nu_ndmax); free(o->ex_p);
    free(o);
}

/* geometry as sdata holds it after inp_geom1/2 + misc (mod_io.f90:809-1365) */
int orc_set_geometry(orc *o, int nxx, int nyy, int nzz, int nnod, int ng, int nmat,
                     const int *ix, const int *iy, const int *iz,
                     const int *ystag_smin, const int *ystag_smax,
                     const int *xstag_smin, const int *xstag_smax,
                     const double *xdel, const double *ydel, const double *zdel,
                     const int *bc, const int *mat)
{
    o->nxx = nxx; o->nyy = nyy; o->nzz = nzz; o->nnod = nnod; o->ng = ng; o->nmat = nmat;
    o->ix = idup(ix, nnod); o->iy = idup(iy, nnod); o->iz = idup(iz, nnod); o->mat = idup(mat, nnod);
    o->ystag_smin = idup(ystag_smin, nyy); o->ystag_smax = idup(ystag_smax, nyy);
    o->xstag_smin = idup(xstag_smin, nxx); o->xstag_smax = idup(xstag_smax, nxx);
    o->xdel = ddup(xdel, nxx); o->ydel = ddup(ydel, nyy); o->zdel = ddup(zdel, nzz);
    o->xeast = bc[0]; o->xwest = bc[1]; o->ynorth = bc[2]; o->ysouth = bc[3];
    o->zbott = bc[4]; o->ztop = bc[5];
    o->xyz = ialloc((size_t)nxx * nyy * nzz);
    o->vdel = dalloc(nnod);
    for (int n = 1; n <= nnod; ++n) {
        XYZ(IX(n), IY(n), IZ(n)) = n;
        VDEL(n) = XDEL(IX(n)) * YDEL(IY(n)) * ZDEL(IZ(n));      /* mod_io.f90:1341-1344 */
    }
    /* number of non-zero columns per row, mod_io.f90:1346-1360 */
    o->ind = (fdm_ind *)calloc(nnod, sizeof(fdm_ind));
    for (int n = 1; n <= nnod; ++n) {
        int noz = 0, i = IX(n), j = IY(n), k = IZ(n);
        if (k != 1) noz++;
        if (j != XSMIN(i)) noz++;
        if (i != YSMIN(j)) noz++;
        noz++;
        if (i != YSMAX(j)) noz++;
        if (j != XSMAX(i)) noz++;
        if (k != nzz) noz++;
        IND(n).ncol = noz;
        IND(n).col = ialloc(noz);
    }
    size_t NG = (size_t)nnod * ng;
    o->D = dalloc(NG); o->sigr = dalloc(NG); o->nuf = dalloc(NG); o->sigf = dalloc(NG);
    o->sigs = dalloc(NG * ng); o->chi = dalloc((size_t)nmat * ng); o->dc = dalloc(NG * 6);
    o->exsrc = dalloc(NG);
    o->velo = dalloc(ng); o->tbeta = dalloc(nmat);
    o->c0 = dalloc((size_t)nnod * NF); o->ft = dalloc(NG); o->fst = dalloc(nnod);
    o->omeg = dalloc(NG); o->sigrp = dalloc(NG); o->L = dalloc(NG); o->dfis = dalloc(nnod);
    o->f0 = dalloc(NG); o->fs0 = dalloc(nnod); o->s0 = dalloc(NG);
    /* module nodal work arrays */
    o->Bcn = dalloc((size_t)ng * ng); o->Bcp = dalloc((size_t)ng * ng);
    o->An = dalloc(ng); o->Bn = dalloc(ng); o->En = dalloc(ng); o->Fn = dalloc(ng);
    o->Gn = dalloc(ng); o->Hn = dalloc(ng); o->Ap = dalloc(ng); o->Bp = dalloc(ng);
    o->Ep = dalloc(ng); o->Fp = dalloc(ng); o->Gp = dalloc(ng); o->Hp = dalloc(ng);
    o->Lm2 = dalloc(ng);
    o->S1 = dalloc(NG); o->S2 = dalloc(NG); o->S3 = dalloc(NG);
    o->a1n = dalloc(ng); o->a2n = dalloc(ng); o->a3n = dalloc(ng); o->a4n = dalloc(ng);
    o->a1p = dalloc(ng); o->a2p = dalloc(ng); o->a3p = dalloc(ng); o->a4p = dalloc(ng);
    o->Ln1 = dalloc(ng); o->Lp1 = dalloc(ng);
    return ORC_OK;
}

/* node-wise cross sections as XS_updt leaves them (mod_xsec.f90:11-46,172-226);
 * all arrays Fortran column-major: D(nnod,ng) ... sigs(nnod,ng,ng) [g -> h], chi(nmat,ng),
 * dc(nnod,ng,6), exsrc(nnod,ng).  NULL leaves the current values. */
int orc_set_xs(orc *o, const double *D, const double *sigr, const double *nuf, const double *sigf,
               const double *sigs, const double *chi, const double *dc, const double *exsrc)
{
    size_t NG = (size_t)N_ * G_;
    if (D) memcpy(o->D, D, NG * sizeof(double));
    if (sigr) memcpy(o->sigr, sigr, NG * sizeof(double));
    if (nuf) memcpy(o->nuf, nuf, NG * sizeof(double));
    if (sigf) memcpy(o->sigf, sigf, NG * sizeof(double));
    if (sigs) memcpy(o->sigs, sigs, NG * G_ * sizeof(double));
    if (chi) memcpy(o->chi, chi, (size_t)o->nmat * G_ * sizeof(double));
    if (dc) memcpy(o->dc, dc, NG * 6 * sizeof(double));
    if (exsrc) memcpy(o->exsrc, exsrc, NG * sizeof(double));
    return ORC_OK;
}

/* %ITER / %KERN values (mod_io.f90:1522-1540, 1593-1621); kern: 0 FDM, 1 PNM, 2 SANM */
int orc_set_control(orc *o, int nout, int nin, int nac, int nupd, double serc, double ferc, int kern,
                    int fixedsrc_mode)
{
    o->nout = nout; o->nin = nin; o->nac = nac; o->nupd = nupd; o->serc = serc; o->ferc = ferc;
    o->kern = kern; o->fixedsrc_mode = fixedsrc_mode;
    return ORC_OK;
}

/* ======================================================================== mod_cmfd */

/* coup_coef, mod_cmfd.f90:11-137 */
static void coup_coef(orc *o)
{
    const double alb = 1.e30;
    double d1, d2;
    if (o->coup_first) {
        if (!o->nod) o->nod = (node_data *)calloc((size_t)N_ * G_, sizeof(node_data));
        for (int g = 1; g <= G_; ++g)
            for (int n = 1; n <= N_; ++n)
                for (int f = 0; f < 6; ++f) NOD(n, g).dn[f] = 0.0;
        o->coup_first = 0;
    }
    for (int g = 1; g <= G_; ++g) {
        for (int n = 1; n <= N_; ++n) {
            int i = IX(n), j = IY(n), k = IZ(n);
            double Dn = V2(o->D, n, g);
            /* x direction */
            if (i == YSMAX(j)) {
                if (o->xeast == 0) NOD(n, g).df[0] = 2.0 * alb * Dn / (2.0 * Dn + alb * XDEL(i));
                else if (o->xeast == 1) NOD(n, g).df[0] = Dn / (2.0 * Dn + 0.5 * XDEL(i));
                else NOD(n, g).df[0] = 0.0;
            } else {
                d2 = V2(o->D, XYZ(i + 1, j, k), g);
                NOD(n, g).df[0] = 2.0 * Dn * d2 / (Dn * XDEL(i + 1) + d2 * XDEL(i));
            }
            if (i == YSMIN(j)) {
                if (o->xwest == 0) NOD(n, g).df[1] = 2.0 * alb * Dn / (2.0 * Dn + alb * XDEL(i));
                else if (o->xwest == 1) NOD(n, g).df[1] = Dn / (2.0 * Dn + 0.5 * XDEL(i));
                else NOD(n, g).df[1] = 0.0;
            } else {
                d1 = V2(o->D, XYZ(i - 1, j, k), g);
                NOD(n, g).df[1] = 2.0 * Dn * d1 / (Dn * XDEL(i - 1) + d1 * XDEL(i));
            }
            /* y direction */
            if (j == XSMAX(i)) {
                if (o->ynorth == 0) NOD(n, g).df[2] = 2.0 * alb * Dn / (2.0 * Dn + alb * YDEL(j));
                else if (o->ynorth == 1) NOD(n, g).df[2] = Dn / (2.0 * Dn + 0.5 * YDEL(j));
                else NOD(n, g).df[2] = 0.0;
            } else {
                d2 = V2(o->D, XYZ(i, j + 1, k), g);
                NOD(n, g).df[2] = 2.0 * Dn * d2 / (Dn * YDEL(j + 1) + d2 * YDEL(j));
            }
            if (j == XSMIN(i)) {
                if (o->ysouth == 0) NOD(n, g).df[3] = 2.0 * alb * Dn / (2.0 * Dn + alb * YDEL(j));
                else if (o->ysouth == 1) NOD(n, g).df[3] = Dn / (2.0 * Dn + 0.5 * YDEL(j));
                else NOD(n, g).df[3] = 0.0;
            } else {
                d1 = V2(o->D, XYZ(i, j - 1, k), g);
                NOD(n, g).df[3] = 2.0 * Dn * d1 / (Dn * YDEL(j - 1) + d1 * YDEL(j));
            }
            /* z direction */
            if (k == o->nzz) {
                if (o->ztop == 0) NOD(n, g).df[4] = 2.0 * alb * Dn / (2.0 * Dn + alb * ZDEL(k));
                else if (o->ztop == 1) NOD(n, g).df[4] = Dn / (2.0 * Dn + 0.5 * ZDEL(k));
                else NOD(n, g).df[4] = 0.0;
            } else {
                d2 = V2(o->D, XYZ(i, j, k + 1), g);
                NOD(n, g).df[4] = 2.0 * Dn * d2 / (Dn * ZDEL(k + 1) + d2 * ZDEL(k));
            }
            if (k == 1) {
                if (o->zbott == 0) NOD(n, g).df[5] = 2.0 * alb * Dn / (2.0 * Dn + alb * ZDEL(k));
                else if (o->zbott == 1) NOD(n, g).df[5] = Dn / (2.0 * Dn + 0.5 * ZDEL(k));
                else NOD(n, g).df[5] = 0.0;
            } else {
                d1 = V2(o->D, XYZ(i, j, k - 1), g);
                NOD(n, g).df[5] = 2.0 * Dn * d1 / (Dn * ZDEL(k - 1) + d1 * ZDEL(k));
            }
        }
    }
}

/* set_ind, mod_cmfd.f90:141-213 */
static void set_ind(orc *o)
{
    int *nodp = ialloc((size_t)o->nxx * o->nyy);
#define NODP(i, j) nodp[((size_t)(j) - 1) * o->nxx + ((i) - 1)]
    int rec = 0;
    for (int j = 1; j <= o->nyy; ++j)
        for (int i = YSMIN(j); i <= YSMAX(j); ++i) { rec++; NODP(i, j) = rec; }
    int np = rec;
    for (int n = 1; n <= N_; ++n) {
        int i = IX(n), j = IY(n), k = IZ(n);
        rec = 0;
        if (k != 1) IND(n).col[rec++] = n - np;
        if (j != XSMIN(i)) IND(n).col[rec++] = n - (NODP(i, j) - NODP(i, j - 1));
        if (i != YSMIN(j)) IND(n).col[rec++] = n - 1;
        IND(n).col[rec++] = n;
        if (i != YSMAX(j)) IND(n).col[rec++] = n + 1;
        if (j != XSMAX(i)) IND(n).col[rec++] = n + (NODP(i, j + 1) - NODP(i, j));
        if (k != o->nzz) IND(n).col[rec++] = n + np;
    }
#undef NODP
    free(nodp);
}

/* matrix_setup, mod_cmfd.f90:217-304 */
void orc_matrix_setup(orc *o, int opt)
{
    if (o->matrix_first) {
        o->A = (fdm_matr *)calloc((size_t)N_ * G_, sizeof(fdm_matr));
        for (int n = 1; n <= N_; ++n)
            for (int g = 1; g <= G_; ++g) AM(n, g).elmn = dalloc(IND(n).ncol);
        set_ind(o);
        o->matrix_first = 0;
    }
    if (opt > 0) coup_coef(o);
    for (int g = 1; g <= G_; ++g) {
        for (int n = 1; n <= N_; ++n) {
            int i = IX(n), j = IY(n), k = IZ(n), rec = 0;
            const node_data *q = &NOD(n, g);
            double *e = AM(n, g).elmn;
            if (k != 1) e[rec++] = -(q->df[5] - q->dn[5]) / ZDEL(k);
            if (j != XSMIN(i)) e[rec++] = -(q->df[3] - q->dn[3]) / YDEL(j);
            if (i != YSMIN(j)) e[rec++] = -(q->df[1] - q->dn[1]) / XDEL(i);
            e[rec++] = (q->df[0] + q->df[1] - q->dn[0] + q->dn[1]) / XDEL(i) +
                       (q->df[2] + q->df[3] - q->dn[2] + q->dn[3]) / YDEL(j) +
                       (q->df[4] + q->df[5] - q->dn[4] + q->dn[5]) / ZDEL(k) + V2(o->sigr, n, g);
            if (i != YSMAX(j)) e[rec++] = -(q->df[0] + q->dn[0]) / XDEL(i);
            if (j != XSMAX(i)) e[rec++] = -(q->df[2] + q->dn[2]) / YDEL(j);
            if (k != o->nzz) e[rec++] = -(q->df[4] + q->dn[4]) / ZDEL(k);
        }
    }
}

/* sp_matvec, mod_cmfd.f90:1247-1268 (function result = fresh vector) */
void orc_sp_matvec(orc *o, int g, const double *x, double *v)
{
    for (int n = 1; n <= N_; ++n) v[n - 1] = 0.0;
    for (int n = 1; n <= N_; ++n) {
        const fdm_ind *id = &IND(n);
        const double *e = AM(n, g).elmn;
        for (int i = 0; i < id->ncol; ++i) v[n - 1] = v[n - 1] + e[i] * x[id->col[i] - 1];
    }
}

/* dproduct, mod_cmfd.f90:1272-1286 */
static double dproduct(int n, const double *a, const double *b)
{
    double x = 0.0;
    for (int i = 0; i < n; ++i) x = x + a[i] * b[i];
    return x;
}

/* l2norm, mod_cmfd.f90:1100-1116 */
static double l2norm(int n, const double *a)
{
    double x = 0.0;
    for (int i = 0; i < n; ++i) x = x + a[i] * a[i];
    return sqrt(x);
}

/* bicg, mod_cmfd.f90:1203-1243: exactly imax iterations, no tolerance, no guards */
void orc_bicg(orc *o, int imax, int g, const double *b, double *x)
{
    int N = N_;
    double *r = dalloc(N), *rs = dalloc(N), *v = dalloc(N), *p = dalloc(N), *s = dalloc(N),
           *t = dalloc(N), *tmp = dalloc(N);
    double rho, rho_prev, alpha, omega, beta, theta;
    orc_sp_matvec(o, g, x, tmp);
    for (int n = 0; n < N; ++n) r[n] = b[n] - tmp[n];
    for (int n = 0; n < N; ++n) rs[n] = r[n];
    rho = 1.0; alpha = 1.0; omega = 1.0;
    for (int n = 0; n < N; ++n) { v[n] = 0.0; p[n] = 0.0; }
    for (int i = 1; i <= imax; ++i) {
        rho_prev = rho;
        rho = dproduct(N, rs, r);
        beta = (rho / rho_prev) * (alpha / omega);
        for (int n = 0; n < N; ++n) p[n] = r[n] + beta * (p[n] - omega * v[n]);
        orc_sp_matvec(o, g, p, v);
        alpha = rho / dproduct(N, rs, v);
        for (int n = 0; n < N; ++n) s[n] = r[n] - alpha * v[n];
        orc_sp_matvec(o, g, s, t);
        theta = dproduct(N, t, t);
        omega = dproduct(N, t, s) / theta;
        for (int n = 0; n < N; ++n) x[n] = x[n] + alpha * p[n] + omega * s[n];
        for (int n = 0; n < N; ++n) r[n] = s[n] - omega * t[n];
    }
    free(r); free(rs); free(v); free(p); free(s); free(t); free(tmp);
}

/* FSrc, mod_cmfd.f90:956-977 */
void orc_fsrc(orc *o, double *fs)
{
    for (int n = 1; n <= N_; ++n) fs[n - 1] = 0.0;
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) fs[n - 1] = fs[n - 1] + V2(o->f0, n, g) * V2(o->nuf, n, g);
}

/* FSrcAd, mod_cmfd.f90:981-1002 */
void orc_fsrc_ad(orc *o, double *fs)
{
    for (int n = 1; n <= N_; ++n) fs[n - 1] = 0.0;
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) fs[n - 1] = fs[n - 1] + V2(o->f0, n, g) * CHI(MAT(n), g);
}

/* TSrc, mod_cmfd.f90:1006-1033 (zeroes ALL of s0 every call) */
void orc_tsrc(orc *o, int g, double Keff, double *bs)
{
    memset(o->s0, 0, (size_t)N_ * G_ * sizeof(double));
    for (int h = 1; h <= G_; ++h)
        for (int n = 1; n <= N_; ++n)
            if (g != h) V2(o->s0, n, g) = V2(o->s0, n, g) + SIGS(n, h, g) * V2(o->f0, n, h);
    for (int n = 1; n <= N_; ++n)
        bs[n - 1] = CHI(MAT(n), g) * o->fs0[n - 1] / Keff + V2(o->s0, n, g) + V2(o->exsrc, n, g);
}

/* TSrcAd, mod_cmfd.f90:1037-1064 */
void orc_tsrc_ad(orc *o, int g, double Keff, double *bs)
{
    memset(o->s0, 0, (size_t)N_ * G_ * sizeof(double));
    for (int h = 1; h <= G_; ++h)
        for (int n = 1; n <= N_; ++n)
            if (g != h) V2(o->s0, n, g) = V2(o->s0, n, g) + SIGS(n, g, h) * V2(o->f0, n, h);
    for (int n = 1; n <= N_; ++n)
        bs[n - 1] = V2(o->nuf, n, g) * o->fs0[n - 1] / Keff + V2(o->s0, n, g) + V2(o->exsrc, n, g);
}

/* TSrcTr, mod_cmfd.f90:1068-1096 */
void orc_tsrc_tr(orc *o, int g, double *bs)
{
    memset(o->s0, 0, (size_t)N_ * G_ * sizeof(double));
    for (int h = 1; h <= G_; ++h)
        for (int n = 1; n <= N_; ++n)
            if (g != h) V2(o->s0, n, g) = V2(o->s0, n, g) + SIGS(n, h, g) * V2(o->f0, n, h);
    for (int n = 1; n <= N_; ++n)
        bs[n - 1] = (1.0 - o->tbeta[MAT(n) - 1] + o->dfis[n - 1]) * CHI(MAT(n), g) * o->fs0[n - 1] +
                    V2(o->s0, n, g) + V2(o->exsrc, n, g);
}

/* Integrate, mod_cmfd.f90:1120-1139 */
double orc_integrate(orc *o, const double *s)
{
    double intg = 0.0;
    for (int n = 1; n <= N_; ++n) intg = intg + VDEL(n) * s[n - 1];
    return intg;
}

/* RelE, mod_cmfd.f90:1143-1168 */
static void RelE(orc *o, const double *newF, const double *oldF, double *rel)
{
    *rel = 0.0;
    for (int n = 0; n < N_; ++n)
        if (fabs(newF[n]) > 1.e-10) {
            double error = fabs(newF[n] - oldF[n]) / fabs(newF[n]);
            if (error > *rel) *rel = error;
        }
}

/* RelEg, mod_cmfd.f90:1172-1199 */
static void RelEg(orc *o, const double *newF, const double *oldF, double *rel)
{
    *rel = 0.0;
    for (int n = 1; n <= N_; ++n)
        for (int g = 1; g <= G_; ++g)
            if (fabs(V2(newF, n, g)) > 1.e-10) {
                double error = fabs(V2(newF, n, g) - V2(oldF, n, g)) / fabs(V2(newF, n, g));
                if (error > *rel) *rel = error;
            }
}

/* fiss_extrp, mod_cmfd.f90:308-335 */
static void fiss_extrp(orc *o, double e1, double e2, const double *erro, const double *errn, double *fs)
{
    double domiR = e2 / e1, mval = 0.0;
    for (int n = 0; n < N_; ++n) if (fabs(erro[n]) > mval) mval = fabs(erro[n]);
    if (mval * mval < 0.0) domiR = -domiR;
    double c = domiR / (1.0 - domiR);
    for (int n = 0; n < N_; ++n) fs[n] = fs[n] + c * errn[n];
    if (o->nextrp < MAXTRACE) o->ex_p[o->nextrp++] = o->cur_p;
}

int orc_nodal_update(orc *o, int cal_mode);
int orc_nodal_update_pnm(orc *o, int cal_mode);
void orc_get_exsrc(orc *o, double ht);

/* nodal_upd, mod_cmfd.f90:339-383 */
static int nodal_upd(orc *o, int nmode)
{
    double st = get_time();
    int rc;
    o->ndmax = 0.0;
    if (o->kern == 2) rc = orc_nodal_update(o, nmode);
    else rc = orc_nodal_update_pnm(o, nmode);
    if (rc) return rc;
    orc_matrix_setup(o, 0);
    if (o->nnodal < MAXTRACE) {
        int q = o->nnodal++;
        o->nu_p[q] = o->cur_p; o->nu_ndmax[q] = o->ndmax;
        o->nu_im[q] = o->im; o->nu_jm[q] = o->jm; o->nu_km[q] = o->km;
    }
    o->nod_time += get_time() - st;
    return ORC_OK;
}

static void trace_reset(orc *o) { o->ntrace = 0; o->nnodal = 0; o->nextrp = 0; }
static void trace_push(orc *o, double ke, double ser, double fer)
{
    if (o->ntrace < MAXTRACE) {
        int q = o->ntrace++;
        struct timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        o->tr_ke[q] = ke; o->tr_ser[q] = ser; o->tr_fer[q] = fer;
        o->tr_wall[q] = (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
        o->tr_fdm[q] = o->fdm_time; o->tr_nod[q] = o->nod_time;
    }
}

/* first-call initialisation shared by outer/outer_fs/outer_th, mod_cmfd.f90:448-454 */
static void init_flux(orc *o, int adjoint)
{
    o->Ke = 1.0;
    for (size_t r = 0; r < (size_t)N_ * G_; ++r) o->f0[r] = 1.0;
    if (adjoint) orc_fsrc_ad(o, o->fs0); else orc_fsrc(o, o->fs0);
    o->outer_first = 0;
}
void orc_init_flux(orc *o, int adjoint) { init_flux(o, adjoint); }

/* The common body of outer / outer_fs / outer_ad / outer_th / outer_tr.
 *   kind 0 = outer    (mod_cmfd.f90:415-509)
 *   kind 1 = outer_fs (mod_cmfd.f90:513-598)  no k-eff update
 *   kind 2 = outer_ad (mod_cmfd.f90:602-699)  groups swept G..1, nodal update only if popt>0
 *   kind 3 = outer_tr (mod_cmfd.f90:800-868)  TSrcTr, no k-eff, no STOP (returns maxi)
 *   kind 4 = outer_th (mod_cmfd.f90:703-796)  maxn iterations, no STOP on non-convergence
 */
static int outer_body(orc *o, int kind, int popt, int maxn, int *maxi, int *niter)
{
    int N = N_, G = G_;
    double *fs0c = dalloc(N), *bs = dalloc(N), *f0c = dalloc((size_t)N * G);
    double *errn = dalloc(N), *erro = dalloc(N);
    double Keo = 0.0, f = 0.0, fc = 0.0, e1, e2;
    int p, rc = ORC_OK, nloop = (kind == 4) ? maxn : o->nout;
    double st = get_time();

    orc_matrix_setup(o, 1);
    if (kind == 3) orc_get_exsrc(o, o->ht_cur);   /* mod_cmfd.f90:830 */
    if (kind != 1) f = orc_integrate(o, o->fs0);
    for (int n = 0; n < N; ++n) errn[n] = 1.0;
    e1 = orc_integrate(o, errn);   /* outer_tr leaves e1/errn uninitialised (harmless, :816-821) */
    o->fdm_time += get_time() - st;
    trace_reset(o);

    for (p = 1; p <= nloop; ++p) {
        st = get_time();
        o->cur_p = p;
        fc = f;
        memcpy(fs0c, o->fs0, (size_t)N * sizeof(double));
        memcpy(f0c, o->f0, (size_t)N * G * sizeof(double));
        Keo = o->Ke;
        memcpy(erro, errn, (size_t)N * sizeof(double));
        if (kind == 2) {
            for (int g = G; g >= 1; --g) {
                orc_tsrc_ad(o, g, o->Ke, bs);
                orc_bicg(o, o->nin, g, bs, &V2(o->f0, 1, g));
            }
            orc_fsrc_ad(o, o->fs0);
        } else {
            for (int g = 1; g <= G; ++g) {
                if (kind == 3) orc_tsrc_tr(o, g, bs); else orc_tsrc(o, g, o->Ke, bs);
                orc_bicg(o, o->nin, g, bs, &V2(o->f0, 1, g));
            }
            orc_fsrc(o, o->fs0);
        }
        for (int n = 0; n < N; ++n) errn[n] = o->fs0[n] - fs0c[n];
        e2 = l2norm(N, errn);
        if (p % o->nac == 0) fiss_extrp(o, e1, e2, erro, errn, o->fs0);
        e1 = e2;
        if (kind == 0 || kind == 2 || kind == 4) {
            f = orc_integrate(o, o->fs0);
            o->Ke = Keo * f / fc;
        }
        RelE(o, o->fs0, fs0c, &o->ser);
        RelEg(o, o->f0, f0c, &o->fer);
        o->fdm_time += get_time() - st;
        if (p % o->nupd == 0 && o->kern != 0 && !(kind == 2 && popt <= 0)) {
            int nmode = (kind == 2) ? 0 : (kind == 3 ? 2 : 1);
            rc = nodal_upd(o, nmode);
            if (rc) goto done;
        }
        trace_push(o, o->Ke, o->ser, o->fer);
        if ((o->ser < o->serc) && (o->fer < o->ferc) && (o->ndmax < (double)1.e-2f)) break; /* default-REAL literal, mod_cmfd.f90:495 */
    }
    if (niter) *niter = (p > nloop) ? nloop : p;
    if (kind == 3) { if (maxi) *maxi = (p == nloop + 1); }
    else if (kind != 4 && p - 1 == nloop) rc = ORC_ERR_MAXOUTER;
done:
    free(fs0c); free(bs); free(f0c); free(errn); free(erro);
    o->status = rc;
    return rc;
}

/* outer(popt), mod_cmfd.f90:415-509 */
int orc_outer(orc *o, int popt, int *niter)
{
    if (o->outer_first && !o->have_state) init_flux(o, 0);
    return outer_body(o, 0, popt, 0, NULL, niter);
}
/* outer_fs(popt), mod_cmfd.f90:513-598 */
int orc_outer_fs(orc *o, int popt, int *niter)
{
    if (o->outer_first && !o->have_state) init_flux(o, 0);
    return outer_body(o, 1, popt, 0, NULL, niter);
}
/* outer_ad(popt), mod_cmfd.f90:602-699 (initialises only when popt > 0, :635) */
int orc_outer_ad(orc *o, int popt, int *niter)
{
    if (o->outer_first && !o->have_state && popt > 0) init_flux(o, 1);
    return outer_body(o, 2, popt, 0, NULL, niter);
}
/* outer_th(maxn), mod_cmfd.f90:703-796 (nupd override :753 is left to the caller) */
int orc_outer_th(orc *o, int maxn, int *niter)
{
    if (o->outer_first && !o->have_state) init_flux(o, 0);
    return outer_body(o, 4, 0, maxn, NULL, niter);
}

/* get_exsrc, mod_cmfd.f90:872-952: bxtab == 1 branch :898-925, bxtab == 0 branch :926-949 */
#define MLAMB(m, i) o->mlamb[((m) - 1) * NF + (i) - 1]
#define MIBETA(m, i) o->mibeta[((m) - 1) * NF + (i) - 1]
#define MVELO(m, g) o->mvelo[((m) - 1) * G_ + (g) - 1]
void orc_get_exsrc(orc *o, double ht)
{
    for (int n = 1; n <= N_; ++n) o->dfis[n - 1] = 0.0;
    if (o->bxtab) {
        for (int n = 1; n <= N_; ++n) {
            double dt = 0.0, dtp = 0.0;
            int m = MAT(n);
            for (int i = 1; i <= NF; ++i) {
                double pxe = exp(-MLAMB(m, i) * ht), a1, a2;
                if (V2(o->nuf, n, G_) > 0.0) a1 = (1.0 - pxe) / (MLAMB(m, i) * ht);
                else a1 = 0.0;
                a2 = 1.0 - a1;
                a1 = a1 - pxe;
                o->dfis[n - 1] = o->dfis[n - 1] + MIBETA(m, i) * a2;
                dt = dt + MLAMB(m, i) * C0(n, i) * pxe + MIBETA(m, i) * a1 * o->fst[n - 1];
                dtp = dtp + MLAMB(m, i) * C0(n, i);
            }
            for (int g = 1; g <= G_; ++g) {
                double pthet = -V2(o->L, n, g) - V2(o->sigrp, n, g) * V2(o->ft, n, g) + V2(o->s0, n, g) +
                               (1.0 - o->tbeta[m - 1]) * CHI(m, g) * o->fst[n - 1] + CHI(m, g) * dtp;
                V2(o->exsrc, n, g) = CHI(m, g) * dt +
                                     exp(V2(o->omeg, n, g) * ht) * V2(o->ft, n, g) / (o->sth * MVELO(m, g) * ht) +
                                     o->bth * pthet;
            }
        }
        return;
    }
    for (int n = 1; n <= N_; ++n) {
        double dt = 0.0, dtp = 0.0;
        for (int i = 1; i <= NF; ++i) {
            double pxe = exp(-o->lamb[i - 1] * ht);
            double a1 = (1.0 - pxe) / (o->lamb[i - 1] * ht);
            double a2 = 1.0 - a1;
            a1 = a1 - pxe;
            o->dfis[n - 1] = o->dfis[n - 1] + o->ibeta[i - 1] * a2;
            dt = dt + o->lamb[i - 1] * C0(n, i) * pxe + o->ibeta[i - 1] * a1 * o->fst[n - 1];
            dtp = dtp + o->lamb[i - 1] * C0(n, i);
        }
        for (int g = 1; g <= G_; ++g) {
            double pthet = -V2(o->L, n, g) - V2(o->sigrp, n, g) * V2(o->ft, n, g) + V2(o->s0, n, g) +
                           (1.0 - o->tbeta[MAT(n) - 1]) * CHI(MAT(n), g) * o->fst[n - 1] +
                           CHI(MAT(n), g) * dtp;
            V2(o->exsrc, n, g) = CHI(MAT(n), g) * dt +
                                 exp(V2(o->omeg, n, g) * ht) * V2(o->ft, n, g) / (o->sth * o->velo[g - 1] * ht) +
                                 o->bth * pthet;
        }
    }
}

/* outer_tr(ht, maxi), mod_cmfd.f90:800-868 */
int orc_outer_tr(orc *o, double ht, int *maxi, int *niter)
{
    o->ht_cur = ht;
    return outer_body(o, 3, 0, 0, maxi, niter);
}

/* PowDis, mod_cmfd.f90:1290-1333 */
int orc_powdis(orc *o, double *p)
{
    double tpow, pw;
    for (int n = 0; n < N_; ++n) p[n] = 0.0;
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) {
            pw = V2(o->f0, n, g) * V2(o->sigf, n, g) * VDEL(n);
            if (pw < 0.0) pw = 0.0;
            p[n - 1] = p[n - 1] + pw;
        }
    tpow = 0.0;
    for (int n = 0; n < N_; ++n) tpow = tpow + p[n];
    if (tpow <= 0.0 && !o->fixedsrc_mode) return ORC_ERR_ZERO_POWER;
    for (int n = 0; n < N_; ++n) p[n] = p[n] / tpow;
    return ORC_OK;
}

/* ======================================================================== mod_nodal */

/* Lxyz, mod_nodal.f90:901-1005 */
void orc_lxyz(orc *o, int n, int g, double *L1, double *L2, double *L3)
{
    double jp, jm;
    int p = 0, m = 0;
    int i = IX(n), j = IY(n), k = IZ(n);
    const node_data *q = &NOD(n, g);
    double fn = V2(o->f0, n, g);

    if (i != YSMAX(j)) p = XYZ(i + 1, j, k);
    if (i != YSMIN(j)) m = XYZ(i - 1, j, k);
    if (i == YSMAX(j)) {
        if (o->xeast == 2) jp = 0.0; else jp = q->df[0] * fn - q->dn[0] * fn;
    } else jp = -q->df[0] * (V2(o->f0, p, g) - fn) - q->dn[0] * (V2(o->f0, p, g) + fn);
    if (i == YSMIN(j)) {
        if (o->xwest == 2) jm = 0.0; else jm = -q->df[1] * fn - q->dn[1] * fn;
    } else jm = -q->df[1] * (fn - V2(o->f0, m, g)) - q->dn[1] * (fn + V2(o->f0, m, g));
    *L1 = (jp - jm) / XDEL(i);

    if (j != XSMAX(i)) p = XYZ(i, j + 1, k);
    if (j != XSMIN(i)) m = XYZ(i, j - 1, k);
    if (j == XSMAX(i)) {
        if (o->ynorth == 2) jp = 0.0; else jp = q->df[2] * fn - q->dn[2] * fn;
    } else jp = -q->df[2] * (V2(o->f0, p, g) - fn) - q->dn[2] * (V2(o->f0, p, g) + fn);
    if (j == XSMIN(i)) {
        if (o->ysouth == 2) jm = 0.0; else jm = -q->df[3] * fn - q->dn[3] * fn;
    } else jm = -q->df[3] * (fn - V2(o->f0, m, g)) - q->dn[3] * (fn + V2(o->f0, m, g));
    *L2 = (jp - jm) / YDEL(j);

    if (k != o->nzz) p = XYZ(i, j, k + 1);
    if (k != 1) m = XYZ(i, j, k - 1);
    if (k == o->nzz) {
        if (o->ztop == 2) jp = 0.0; else jp = q->df[4] * fn - q->dn[4] * fn;
    } else jp = -q->df[4] * (V2(o->f0, p, g) - fn) - q->dn[4] * (V2(o->f0, p, g) + fn);
    if (k == 1) {
        if (o->zbott == 2) jm = 0.0; else jm = -q->df[5] * fn - q->dn[5] * fn;
    } else jm = -q->df[5] * (fn - V2(o->f0, m, g)) - q->dn[5] * (fn + V2(o->f0, m, g));
    *L3 = (jp - jm) / ZDEL(k);
}

/* get_source, mod_nodal.f90:1009-1043 */
static void get_source(orc *o)
{
    double L1, L2, L3;
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) {
            orc_lxyz(o, n, g, &L1, &L2, &L3);
            if (o->cmode == 2) {
                V2(o->S1, n, g) = L2 + L3 - V2(o->exsrc, n, g);
                V2(o->S2, n, g) = L1 + L3 - V2(o->exsrc, n, g);
                V2(o->S3, n, g) = L1 + L2 - V2(o->exsrc, n, g);
            } else {
                V2(o->S1, n, g) = L2 + L3;
                V2(o->S2, n, g) = L1 + L3;
                V2(o->S3, n, g) = L1 + L2;
            }
        }
}

/* geometry of the u-line through node n: own index c, line ends lo/hi, neighbours p/m,
 * mesh sizes, boundary codes -- the common preamble of TLUpd1/TLUpd2 */
typedef struct { int c, lo, hi, p, m, bcm, bcp; double hc, hm, hp; const double *S; } uline;
static void get_uline(orc *o, int u, int n, uline *q)
{
    int i = IX(n), j = IY(n), k = IZ(n);
    q->p = q->m = 0; q->hm = q->hp = 0.0;
    if (u == 1) {
        q->c = i; q->lo = YSMIN(j); q->hi = YSMAX(j); q->bcm = o->xwest; q->bcp = o->xeast; q->S = o->S1;
        q->hc = XDEL(i);
        if (i != q->hi) { q->p = XYZ(i + 1, j, k); q->hp = XDEL(i + 1); }
        if (i != q->lo) { q->m = XYZ(i - 1, j, k); q->hm = XDEL(i - 1); }
    } else if (u == 2) {
        q->c = j; q->lo = XSMIN(i); q->hi = XSMAX(i); q->bcm = o->ysouth; q->bcp = o->ynorth; q->S = o->S2;
        q->hc = YDEL(j);
        if (j != q->hi) { q->p = XYZ(i, j + 1, k); q->hp = YDEL(j + 1); }
        if (j != q->lo) { q->m = XYZ(i, j - 1, k); q->hm = YDEL(j - 1); }
    } else {
        q->c = k; q->lo = 1; q->hi = o->nzz; q->bcm = o->zbott; q->bcp = o->ztop; q->S = o->S3;
        q->hc = ZDEL(k);
        if (k != q->hi) { q->p = XYZ(i, j, k + 1); q->hp = ZDEL(k + 1); }
        if (k != q->lo) { q->m = XYZ(i, j, k - 1); q->hm = ZDEL(k - 1); }
    }
}

/* TLUpd1, mod_nodal.f90:1047-1199 (the three direction branches are textually identical
 * up to the arrays used; one body serves all three) */
static void TLUpd1(orc *o, int u, int n, int g, double *Lmom1)
{
    uline q; get_uline(o, u, n, &q);
    const double *S = q.S;
    double tm, tp, p1m, p2m, p1p, p2p, hp, L;
    if (q.c == q.lo) {
        if (q.bcm == 2) {
            tm = 1.0; tp = q.hp / q.hc;
            p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0;
            hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
            L = (p1m * p2m * (V2(S, q.p, g) - V2(S, n, g))) / hp;
        } else {
            tp = q.hp / q.hc; p1p = tp + 1.0;
            L = (V2(S, q.p, g) - V2(S, n, g)) / p1p;
        }
    } else if (q.c == q.hi) {
        if (q.bcp == 2) {
            tm = q.hm / q.hc; tp = 1.0;
            p1m = tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
            hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
            L = (p1p * p2p * (V2(S, n, g) - V2(S, q.m, g))) / hp;
        } else {
            tm = q.hm / q.hc; p1m = tm + 1.0;
            L = (V2(S, n, g) - V2(S, q.m, g)) / p1m;
        }
    } else {
        tm = q.hm / q.hc; tp = q.hp / q.hc;
        p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
        hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
        L = (p1m * p2m * (V2(S, q.p, g) - V2(S, n, g)) + p1p * p2p * (V2(S, n, g) - V2(S, q.m, g))) / hp;
    }
    *Lmom1 = 0.25 * (q.hc * q.hc) / V2(o->D, n, g) * L;
}

/* TLUpd2, mod_nodal.f90:1203-1341 */
static void TLUpd2(orc *o, int u, int n, int g, double *Lmom2)
{
    uline q; get_uline(o, u, n, &q);
    const double *S = q.S;
    double tm, tp, p1m, p1p, hp, L;
    if (q.c == q.lo) {
        if (q.bcm == 2) {
            tm = 1.0; tp = q.hp / q.hc; p1m = tm + 1.0; p1p = tp + 1.0;
            hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
            L = (p1m * (V2(S, q.p, g) - V2(S, n, g))) / hp;
        } else L = 0.0;
    } else if (q.c == q.hi) {
        if (q.bcp == 2) {
            tm = q.hm / q.hc; tp = 1.0; p1m = tm + 1.0; p1p = tp + 1.0;
            hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
            L = (p1p * (V2(S, q.m, g) - V2(S, n, g))) / hp;
        } else L = 0.0;
    } else {
        tm = q.hm / q.hc; tp = q.hp / q.hc; p1m = tm + 1.0; p1p = tp + 1.0;
        hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
        L = (p1m * (V2(S, q.p, g) - V2(S, n, g)) + p1p * (V2(S, q.m, g) - V2(S, n, g))) / hp;
    }
    *Lmom2 = 0.25 * (q.hc * q.hc) / V2(o->D, n, g) * L;
}

static double udel(orc *o, int u, int n)
{
    return (u == 1) ? XDEL(IX(n)) : (u == 2) ? YDEL(IY(n)) : ZDEL(IZ(n));
}

/* get_B, mod_nodal.f90:1345-1405 (cmode 2 keeps the chi*chi of :1383-1384) */
static void get_B(orc *o, int u, int n, double *B)
{
    double dum, dn = udel(o, u, n);
    int m = MAT(n);
    for (int g = 1; g <= G_; ++g)
        for (int h = 1; h <= G_; ++h) {
            if (o->cmode == 1) {
                if (g == h) dum = V2(o->sigr, n, g) - CHI(m, g) * V2(o->nuf, n, h) / o->Ke;
                else dum = -SIGS(n, h, g) - CHI(m, g) * V2(o->nuf, n, h) / o->Ke;
            } else if (o->cmode == 2) {
                if (g == h)
                    dum = V2(o->sigr, n, g) - (1.0 - o->tbeta[m - 1] + o->dfis[n - 1]) * CHI(m, g) * CHI(m, g) * V2(o->nuf, n, h);
                else
                    dum = -SIGS(n, h, g) - (1.0 - o->tbeta[m - 1] + o->dfis[n - 1]) * CHI(m, g) * V2(o->nuf, n, h);
            } else {
                if (g == h) dum = V2(o->sigr, n, g) - CHI(m, g) * V2(o->nuf, n, h) / o->Ke;
                else dum = -SIGS(n, g, h) - CHI(m, h) * V2(o->nuf, n, g) / o->Ke;
            }
            BC_(B, g, h) = 0.25 * (dn * dn) / V2(o->D, n, g) * dum;
        }
}

/* get_ABEFGH, mod_nodal.f90:1409-1451 */
static void get_ABEFGH(orc *o, int n, int u, double *Aa, double *Bb, double *Ee, double *Ff, double *Gg, double *Hh)
{
    double dn = udel(o, u, n);
    for (int g = 1; g <= G_; ++g) {
        double alp = 0.5 * sqrt(V2(o->sigr, n, g) / V2(o->D, n, g)) * dn;
        double alp2 = alp * alp;
        double sh = sinh(alp), ch = cosh(alp);
        double m0c = sh / alp;
        double m1s = 3.0 * (ch / alp - sh / alp2);
        double m2c = 5.0 * (sh / alp - 3.0 * ch / alp2 + 3.0 * sh / (alp * alp * alp));
        Aa[g - 1] = (sh - m1s) / (alp2 * m1s);
        Bb[g - 1] = (ch - m0c - m2c) / (alp2 * m2c);
        Ee[g - 1] = (m0c / m2c - 3.0 / alp2);
        Ff[g - 1] = (alp * ch - m1s) / (alp2 * m1s);
        Gg[g - 1] = (alp * sh - 3.0 * m2c) / (ch - m0c - m2c);
        Hh[g - 1] = (alp * ch - m1s) / (sh - m1s);
    }
}

/* LU_solve, mod_nodal.f90:829-897: Doolittle, no pivoting, abort on |mat(i,i)| < 1e-4.
 * mat is msize x msize column-major: mat(i,j) = mat[(j-1)*msize + (i-1)]. */
static int LU_solve(orc *o, int nt, int msize, const double *mat, const double *b, double *x)
{
    (void)nt;
    double Lm[32 * 32], U[32 * 32], y[32];
    double *Lp = Lm, *Up = U, *yp = y;
    if (msize > 32) {
        Lp = dalloc((size_t)msize * msize); Up = dalloc((size_t)msize * msize); yp = dalloc(msize);
    }
#define M_(a, i, j) ((a)[((size_t)(j) - 1) * msize + ((i) - 1)])
    for (int q = 0; q < msize * msize; ++q) { Up[q] = mat[q]; Lp[q] = 0.0; }
    for (int i = 1; i <= msize; ++i) {
        if (fabs(M_(mat, i, i)) < (double)10e-5f) { /* default-REAL literal in the reference */
            if (msize > 32) { free(Lp); free(Up); free(yp); }
            o->status = ORC_ERR_LU_DIAG;
            return ORC_ERR_LU_DIAG;
        }
        M_(Lp, i, i) = 1.0;
        for (int j = i + 1; j <= msize; ++j) {
            double piv = M_(Up, j, i) / M_(Up, i, i);
            M_(Lp, j, i) = piv;
            for (int k = i; k <= msize; ++k) M_(Up, j, k) = M_(Up, j, k) - piv * M_(Up, i, k);
            M_(Up, j, i) = 0.0;
        }
    }
    yp[0] = b[0];
    for (int i = 2; i <= msize; ++i) {
        double isum = 0.0;
        for (int k = 1; k <= i - 1; ++k) isum = isum + M_(Lp, i, k) * yp[k - 1];
        yp[i - 1] = b[i - 1] - isum;
    }
    x[msize - 1] = yp[msize - 1] / M_(Up, msize, msize);
    for (int i = msize - 1; i >= 1; --i) {
        double isum = 0.0;
        for (int k = i + 1; k <= msize; ++k) isum = isum + M_(Up, i, k) * x[k - 1];
        x[i - 1] = (yp[i - 1] - isum) / M_(Up, i, i);
    }
#undef M_
    if (msize > 32) { free(Lp); free(Up); free(yp); }
    return ORC_OK;
}

/* get_a2matvec, mod_nodal.f90:769-825 (uses Bcp, Ep of the current "p" node; sets Lm2) */
static void get_a2matvec(orc *o, int u, int n, double *A, double *b)
{
    int G = G_;
    double S, dn = udel(o, u, n);
    const double *Su = (u == 1) ? o->S1 : (u == 2) ? o->S2 : o->S3;
    for (int g = 1; g <= G; ++g) {
        double Bf = 0.0;
        if (o->cmode == 2) S = 0.25 * (dn * dn) / V2(o->D, n, g) * V2(Su, n, g);
        else S = 0.25 * (dn * dn) / V2(o->D, n, g) * (V2(Su, n, g) - V2(o->exsrc, n, g));
        for (int h = 1; h <= G; ++h) {
            if (h == g) BC_(A, g, g) = BC_(o->Bcp, g, h) * o->Ep[g - 1] + 3.0;
            else BC_(A, g, h) = BC_(o->Bcp, g, h) * o->Ep[g - 1];
            Bf = Bf + BC_(o->Bcp, g, h) * V2(o->f0, n, h);
        }
        TLUpd2(o, u, n, g, &o->Lm2[g - 1]);
        b[g - 1] = Bf - o->Ep[g - 1] * o->Lm2[g - 1] + S;
    }
}

/* get_a4, mod_nodal.f90:740-765 */
static void get_a4(orc *o, const double *a2, double *a4)
{
    for (int g = 1; g <= G_; ++g) {
        double Bf = 0.0;
        for (int h = 1; h <= G_; ++h) Bf = Bf + BC_(o->Bcp, g, h) * a2[h - 1];
        a4[g - 1] = o->Bp[g - 1] * (Bf + o->Lm2[g - 1]);
    }
}

/* get_a3, mod_nodal.f90:702-736 */
static void get_a3(orc *o, int cp, const double *a1, const double *Lmn1, double *a3)
{
    const double *Bc = (cp == 1) ? o->Bcn : o->Bcp;
    const double *Ac = (cp == 1) ? o->An : o->Ap;
    for (int g = 1; g <= G_; ++g) {
        double Bf = 0.0;
        for (int h = 1; h <= G_; ++h) Bf = Bf + BC_(Bc, g, h) * a1[h - 1];
        a3[g - 1] = Ac[g - 1] * (Bf + Lmn1[g - 1]);
    }
}

/* get_a1matvec_first, mod_nodal.f90:483-552 */
static void get_a1matvec_first(orc *o, int bc, int u, int p, const double *a2p, const double *a4p, double *A, double *b)
{
    int G = G_, sf = (u == 1) ? 2 : (u == 2) ? 4 : 6;
    double dn = udel(o, u, p);
    for (int g = 1; g <= G; ++g) {
        TLUpd1(o, u, p, g, &o->Lp1[g - 1]);
        double Pp = 2.0 * V2(o->D, p, g) / dn;
        double Apg = o->Ap[g - 1], Fpg = o->Fp[g - 1], Gpg = o->Gp[g - 1], Hpg = o->Hp[g - 1];
        double L1 = o->Lp1[g - 1], dcp = DC(p, g, sf);
        if (bc == 2) {
            for (int h = 1; h <= G; ++h) {
                if (h == g) BC_(A, g, g) = Pp * (BC_(o->Bcp, g, h) * Fpg + 1.0);
                else BC_(A, g, h) = Pp * BC_(o->Bcp, g, h) * Fpg;
            }
            b[g - 1] = Pp * (3.0 * a2p[g - 1] + Gpg * a4p[g - 1] - Fpg * L1);
        } else if (bc == 1) {
            for (int h = 1; h <= G; ++h) {
                if (h == g)
                    BC_(A, g, g) = -dcp * (1.0 + Apg * BC_(o->Bcp, g, h)) - 2.0 * Pp * (Apg * BC_(o->Bcp, g, h) * Hpg + 1.0);
                else
                    BC_(A, g, h) = -dcp * Apg * BC_(o->Bcp, g, h) - 2.0 * Pp * Apg * BC_(o->Bcp, g, h) * Hpg;
            }
            b[g - 1] = 2.0 * Pp * (Apg * Hpg * L1 - 3.0 * a2p[g - 1] - Gpg * a4p[g - 1]) -
                       dcp * (a2p[g - 1] + a4p[g - 1] + V2(o->f0, p, g) - Apg * L1);
        } else {
            for (int h = 1; h <= G; ++h) {
                if (h == g) BC_(A, g, g) = dcp * (1.0 + Apg * BC_(o->Bcp, g, h));
                else BC_(A, g, h) = dcp * Apg * BC_(o->Bcp, g, h);
            }
            b[g - 1] = dcp * (a2p[g - 1] + a4p[g - 1] + V2(o->f0, p, g) - Apg * L1);
        }
    }
}

/* get_a1matvec_last, mod_nodal.f90:556-626 */
static void get_a1matvec_last(orc *o, int bc, int u, int n, const double *a2n, const double *a4n, double *A, double *b)
{
    int G = G_, sf = (u == 1) ? 1 : (u == 2) ? 3 : 5;
    double dn = udel(o, u, n);
    for (int g = 1; g <= G; ++g) {
        o->Ln1[g - 1] = o->Lp1[g - 1];
        double Pn = 2.0 * V2(o->D, n, g) / dn;
        double Ang = o->An[g - 1], Fng = o->Fn[g - 1], Gng = o->Gn[g - 1], Hng = o->Hn[g - 1];
        double L1 = o->Ln1[g - 1], dcn = DC(n, g, sf);
        if (bc == 2) {
            for (int h = 1; h <= G; ++h) {
                if (h == g) BC_(A, g, g) = -Pn * (BC_(o->Bcn, g, h) * Fng + 1.0);
                else BC_(A, g, h) = -Pn * BC_(o->Bcn, g, h) * Fng;
            }
            b[g - 1] = Pn * (3.0 * a2n[g - 1] + Gng * a4n[g - 1] + Fng * L1);
        } else if (bc == 1) {
            for (int h = 1; h <= G; ++h) {
                if (h == g)
                    BC_(A, g, g) = dcn * (1.0 + Ang * BC_(o->Bcn, g, h)) + 2.0 * Pn * (Ang * BC_(o->Bcn, g, h) * Hng + 1.0);
                else
                    BC_(A, g, h) = dcn * Ang * BC_(o->Bcn, g, h) + 2.0 * Pn * Ang * BC_(o->Bcn, g, h) * Hng;
            }
            b[g - 1] = -2.0 * Pn * (Ang * Hng * L1 + 3.0 * a2n[g - 1] + Gng * a4n[g - 1]) -
                       dcn * (a2n[g - 1] + a4n[g - 1] + V2(o->f0, n, g) + Ang * L1);
        } else {
            for (int h = 1; h <= G; ++h) {
                if (h == g) BC_(A, g, g) = dcn * (1.0 + Ang * BC_(o->Bcn, g, h));
                else BC_(A, g, h) = dcn * Ang * BC_(o->Bcn, g, h);
            }
            b[g - 1] = -dcn * (a2n[g - 1] + a4n[g - 1] + V2(o->f0, n, g) + Ang * L1);
        }
    }
}

/* get_a1matvec, mod_nodal.f90:630-698 (2G x 2G; the ADF cross terms of :693-694 are kept) */
static void get_a1matvec(orc *o, int u, int n, int p, const double *a2n, const double *a4n,
                         const double *a2p, const double *a4p, double *A, double *b)
{
    int G = G_, G2 = 2 * G_, sf = (u == 1) ? 1 : (u == 2) ? 3 : 5;
    double hn = udel(o, u, n), hp = udel(o, u, p);
#define R_(i, j) A[((size_t)(j) - 1) * G2 + ((i) - 1)]
    for (int g = 1; g <= G; ++g) {
        o->Ln1[g - 1] = o->Lp1[g - 1];
        TLUpd1(o, u, p, g, &o->Lp1[g - 1]);
        double Pn = 2.0 * V2(o->D, n, g) / hn, Pp = 2.0 * V2(o->D, p, g) / hp;
        for (int h = 1; h <= G; ++h) {
            if (h == g) {
                R_(g, g) = -Pn * (BC_(o->Bcn, g, h) * o->Fn[g - 1] + 1.0);
                R_(g, g + G) = Pp * (BC_(o->Bcp, g, h) * o->Fp[g - 1] + 1.0);
            } else {
                R_(g, h) = -Pn * BC_(o->Bcn, g, h) * o->Fn[g - 1];
                R_(g, h + G) = Pp * BC_(o->Bcp, g, h) * o->Fp[g - 1];
            }
        }
        b[g - 1] = Pn * (3.0 * a2n[g - 1] + o->Gn[g - 1] * a4n[g - 1] + o->Fn[g - 1] * o->Ln1[g - 1]) +
                   Pp * (3.0 * a2p[g - 1] + o->Gp[g - 1] * a4p[g - 1] - o->Fp[g - 1] * o->Lp1[g - 1]);
    }
    for (int g = 1; g <= G; ++g) {
        double dcn = DC(n, g, sf), dcp = DC(p, g, sf + 1);
        for (int h = 1; h <= G; ++h) {
            if (h == g) {
                R_(g + G, g) = dcn * (BC_(o->Bcn, g, h) * o->An[g - 1] + 1.0);
                R_(g + G, g + G) = dcp * (BC_(o->Bcp, g, h) * o->Ap[g - 1] + 1.0);
            } else {
                R_(g + G, h) = dcn * BC_(o->Bcn, g, h) * o->An[g - 1];
                R_(g + G, h + G) = dcp * BC_(o->Bcp, g, h) * o->Ap[g - 1];
            }
        }
        b[g + G - 1] = dcp * (a2p[g - 1] + a4p[g - 1] + V2(o->f0, p, g) - o->An[g - 1] * o->Ln1[g - 1]) -
                       dcn * (a2n[g - 1] + a4n[g - 1] + V2(o->f0, n, g) + o->Ap[g - 1] * o->Lp1[g - 1]);
    }
#undef R_
}

/* get_coefs_first, mod_nodal.f90:375-406 */
static int get_coefs_first(orc *o, int bc, int u, int p)
{
    int G = G_, rc;
    double A[32 * 32], b[32];
    if (G > 32) return -1;
    get_a2matvec(o, u, p, A, b);
    if ((rc = LU_solve(o, p, G, A, b, o->a2p))) return rc;
    get_a4(o, o->a2p, o->a4p);
    get_a1matvec_first(o, bc, u, p, o->a2p, o->a4p, A, b);
    if ((rc = LU_solve(o, p, G, A, b, o->a1p))) return rc;
    get_a3(o, 2, o->a1p, o->Lp1, o->a3p);
    return ORC_OK;
}

/* get_coefs_last, mod_nodal.f90:410-438 */
static int get_coefs_last(orc *o, int bc, int u, int n)
{
    int G = G_, rc;
    double A[32 * 32], b[32];
    for (int g = 0; g < G; ++g) { o->a2n[g] = o->a2p[g]; o->a4n[g] = o->a4p[g]; }
    get_a1matvec_last(o, bc, u, n, o->a2n, o->a4n, A, b);
    if ((rc = LU_solve(o, n, G, A, b, o->a1n))) return rc;
    get_a3(o, 1, o->a1n, o->Ln1, o->a3n);
    return ORC_OK;
}

/* get_coefs, mod_nodal.f90:442-479 */
static int get_coefs(orc *o, int u, int n, int p)
{
    int G = G_, rc;
    static double R[64 * 64], s[64], sx[64];
    double A[32 * 32], b[32];
    for (int g = 0; g < G; ++g) { o->a2n[g] = o->a2p[g]; o->a4n[g] = o->a4p[g]; }
    get_a2matvec(o, u, p, A, b);
    if ((rc = LU_solve(o, p, G, A, b, o->a2p))) return rc;
    get_a4(o, o->a2p, o->a4p);
    get_a1matvec(o, u, n, p, o->a2n, o->a4n, o->a2p, o->a4p, R, s);
    if ((rc = LU_solve(o, n, 2 * G, R, s, sx))) return rc;
    for (int g = 0; g < G; ++g) o->a1n[g] = sx[g];
    get_a3(o, 1, o->a1n, o->Ln1, o->a3n);
    return ORC_OK;
}

/* nodal_coup_upd, mod_nodal.f90:282-371; n or p == 0 means "not present" */
static void nodal_coup_upd(orc *o, int u, const double *a1, const double *a2, const double *a3, const double *a4, int n, int p)
{
    double dh, jp, nder, ndpr;
    int sf = (u == 1) ? 1 : (u == 2) ? 3 : 5;
    dh = udel(o, u, n ? n : p);
    if (n && p) {
        for (int g = 1; g <= G_; ++g) {
            jp = -2.0 * V2(o->D, n, g) / dh * (a1[g - 1] + 3.0 * a2[g - 1] + o->Hn[g - 1] * a3[g - 1] + o->Gn[g - 1] * a4[g - 1]);
            ndpr = NOD(n, g).dn[sf - 1];
            NOD(n, g).dn[sf - 1] = (NOD(n, g).df[sf - 1] * (V2(o->f0, n, g) - V2(o->f0, p, g)) - jp) / (V2(o->f0, n, g) + V2(o->f0, p, g));
            NOD(p, g).dn[sf] = NOD(n, g).dn[sf - 1];
            nder = fabs(NOD(n, g).dn[sf - 1] - ndpr);
            if (nder > o->ndmax) { o->ndmax = nder; o->im = IX(n); o->jm = IY(n); o->km = IZ(n); }
        }
    } else if (p) {
        for (int g = 1; g <= G_; ++g) {
            jp = -2.0 * V2(o->D, p, g) / dh * (a1[g - 1] - 3.0 * a2[g - 1] + o->Hp[g - 1] * a3[g - 1] - o->Gp[g - 1] * a4[g - 1]);
            ndpr = NOD(p, g).dn[sf];
            NOD(p, g).dn[sf] = -(jp / V2(o->f0, p, g) + NOD(p, g).df[sf]);
            nder = fabs(NOD(p, g).dn[sf] - ndpr);
            if (nder > o->ndmax) { o->ndmax = nder; o->im = IX(p); o->jm = IY(p); o->km = IZ(p); }
        }
    } else {
        for (int g = 1; g <= G_; ++g) {
            jp = -2.0 * V2(o->D, n, g) / dh * (a1[g - 1] + 3.0 * a2[g - 1] + o->Hn[g - 1] * a3[g - 1] + o->Gn[g - 1] * a4[g - 1]);
            ndpr = NOD(n, g).dn[sf - 1];
            NOD(n, g).dn[sf - 1] = -(jp / V2(o->f0, n, g) - NOD(n, g).df[sf - 1]);
            nder = fabs(NOD(n, g).dn[sf - 1] - ndpr);
            if (nder > o->ndmax) { o->ndmax = nder; o->im = IX(n); o->jm = IY(n); o->km = IZ(n); }
        }
    }
}

static void shift_p_to_n(orc *o)
{
    size_t G = G_;
    memcpy(o->Bcn, o->Bcp, G * G * sizeof(double));
    memcpy(o->An, o->Ap, G * sizeof(double)); memcpy(o->Bn, o->Bp, G * sizeof(double));
    memcpy(o->En, o->Ep, G * sizeof(double)); memcpy(o->Fn, o->Fp, G * sizeof(double));
    memcpy(o->Gn, o->Gp, G * sizeof(double)); memcpy(o->Hn, o->Hp, G * sizeof(double));
}

/* one line of the sweep: nodes line[0..len-1] along direction u (mod_nodal.f90:55-73) */
static int sweep_line(orc *o, int u, const int *line, int len, int bc_first, int bc_last, int sanm)
{
    int rc, n, p;
    p = line[0];
    get_B(o, u, p, o->Bcp);
    if (sanm) get_ABEFGH(o, p, u, o->Ap, o->Bp, o->Ep, o->Fp, o->Gp, o->Hp);
    if ((rc = get_coefs_first(o, bc_first, u, p))) return rc;
    nodal_coup_upd(o, u, o->a1p, o->a2p, o->a3p, o->a4p, 0, p);
    for (int q = 0; q + 1 < len; ++q) {
        n = line[q]; p = line[q + 1];
        shift_p_to_n(o);
        get_B(o, u, p, o->Bcp);
        if (sanm) get_ABEFGH(o, p, u, o->Ap, o->Bp, o->Ep, o->Fp, o->Gp, o->Hp);
        if ((rc = get_coefs(o, u, n, p))) return rc;
        nodal_coup_upd(o, u, o->a1n, o->a2n, o->a3n, o->a4n, n, p);
    }
    n = line[len - 1];
    shift_p_to_n(o);
    if ((rc = get_coefs_last(o, bc_last, u, n))) return rc;
    nodal_coup_upd(o, u, o->a1n, o->a2n, o->a3n, o->a4n, n, 0);
    return ORC_OK;
}

/* nodal_update (SANM), mod_nodal.f90:18-147; nodal_update_pnm, :151-278 */
static int nodal_update_common(orc *o, int cal_mode, int sanm)
{
    int rc = ORC_OK;
    int maxlen = o->nxx > o->nyy ? o->nxx : o->nyy;
    if (o->nzz > maxlen) maxlen = o->nzz;
    int *line = ialloc(maxlen);
    if (!sanm) { /* mod_nodal.f90:180-182 */
        for (int g = 0; g < G_; ++g) {
            o->An[g] = 1.0 / 15.0; o->Bn[g] = 1.0 / 35.0; o->En[g] = 2.0 / 7.0;
            o->Fn[g] = 2.0 / 5.0; o->Gn[g] = 10.0; o->Hn[g] = 6.0;
            o->Ap[g] = o->An[g]; o->Bp[g] = o->Bn[g]; o->Ep[g] = o->En[g];
            o->Fp[g] = o->Fn[g]; o->Gp[g] = o->Gn[g]; o->Hp[g] = o->Hn[g];
        }
    }
    o->cmode = cal_mode;
    get_source(o);
    /* x sweeps */
    for (int k = 1; k <= o->nzz && !rc; ++k)
        for (int j = 1; j <= o->nyy && !rc; ++j) {
            int len = 0;
            for (int i = YSMIN(j); i <= YSMAX(j); ++i) line[len++] = XYZ(i, j, k);
            rc = sweep_line(o, 1, line, len, o->xwest, o->xeast, sanm);
        }
    /* y sweeps */
    for (int k = 1; k <= o->nzz && !rc; ++k)
        for (int i = 1; i <= o->nxx && !rc; ++i) {
            int len = 0;
            for (int j = XSMIN(i); j <= XSMAX(i); ++j) line[len++] = XYZ(i, j, k);
            rc = sweep_line(o, 2, line, len, o->ysouth, o->ynorth, sanm);
        }
    /* z sweeps */
    for (int j = 1; j <= o->nyy && !rc; ++j)
        for (int i = YSMIN(j); i <= YSMAX(j) && !rc; ++i) {
            int len = 0;
            for (int k = 1; k <= o->nzz; ++k) line[len++] = XYZ(i, j, k);
            rc = sweep_line(o, 3, line, len, o->zbott, o->ztop, sanm);
        }
    free(line);
    if (rc) return rc;
    if (o->ndmax > 1.e3) { o->status = ORC_ERR_NDMAX; return ORC_ERR_NDMAX; }
    return ORC_OK;
}
int orc_nodal_update(orc *o, int cal_mode) { return nodal_update_common(o, cal_mode, 1); }
int orc_nodal_update_pnm(orc *o, int cal_mode) { return nodal_update_common(o, cal_mode, 0); }

/* nodal_upd as a public entry (ndmax reset + update + matrix_setup(0)), mod_cmfd.f90:339-383 */
int orc_nodal_upd(orc *o, int nmode) { o->cur_p = 0; return nodal_upd(o, nmode); }

/* ======================================================================== mod_trans helpers */

/* PowTot, mod_trans.f90:523-557 */
double orc_powtot(orc *o, const double *fx)
{
    double tpow = 0.0, pw;
    double *p = dalloc(N_);
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) {
            pw = V2(fx, n, g) * V2(o->sigf, n, g) * VDEL(n);
            if (pw < 0.0) pw = 0.0;
            p[n - 1] = p[n - 1] + pw;
        }
    for (int n = 0; n < N_; ++n) tpow = tpow + p[n];
    free(p);
    return tpow;
}

/* iPden, mod_trans.f90:561-597 */
void orc_ipden(orc *o)
{
    if (o->bxtab) {
        for (int n = 1; n <= N_; ++n)
            for (int j = 1; j <= NF; ++j) {
                if (V2(o->nuf, n, G_) > 0.0) {      /* if it is fuel */
                    double blamb = MIBETA(MAT(n), j) / MLAMB(MAT(n), j);
                    C0(n, j) = blamb * o->fs0[n - 1];
                } else C0(n, j) = 0.0;
            }
        return;
    }
    for (int n = 1; n <= N_; ++n)
        for (int j = 1; j <= NF; ++j) {
            double blamb = o->ibeta[j - 1] / o->lamb[j - 1];
            C0(n, j) = blamb * o->fs0[n - 1];
        }
}

/* uPden, mod_trans.f90:601-644 */
void orc_upden(orc *o, double ht)
{
    if (o->bxtab) {
        for (int i = 1; i <= NF; ++i)
            for (int n = 1; n <= N_; ++n)
                if (V2(o->nuf, n, G_) > 0.0) {
                    int m = MAT(n);
                    double pxe = exp(-MLAMB(m, i) * ht);
                    double a1 = (1.0 - pxe) / (MLAMB(m, i) * ht);
                    double a2 = 1.0 - a1;
                    a1 = a1 - pxe;
                    C0(n, i) = C0(n, i) * pxe + MIBETA(m, i) / MLAMB(m, i) * (a1 * o->fst[n - 1] + a2 * o->fs0[n - 1]);
                }
        return;
    }
    for (int i = 1; i <= NF; ++i) {
        double pxe = exp(-o->lamb[i - 1] * ht);
        double a1 = (1.0 - pxe) / (o->lamb[i - 1] * ht);
        double a2 = 1.0 - a1;
        a1 = a1 - pxe;
        for (int n = 1; n <= N_; ++n)
            C0(n, i) = C0(n, i) * pxe + o->ibeta[i - 1] / o->lamb[i - 1] * (a1 * o->fst[n - 1] + a2 * o->fs0[n - 1]);
    }
}

/* reactivity, mod_trans.f90:648-688 (also fills L(n,g)) */
double orc_reactivity(orc *o, const double *af, const double *sigrp)
{
    double src = 0.0, rem = 0.0, lea = 0.0, fde = 0.0, L1, L2, L3;
    double *scg = dalloc(N_);
    for (int g = 1; g <= G_; ++g) {
        for (int n = 0; n < N_; ++n) scg[n] = 0.0;
        for (int h = 1; h <= G_; ++h)
            for (int n = 1; n <= N_; ++n)
                if (g != h) scg[n - 1] = scg[n - 1] + SIGS(n, h, g) * V2(o->f0, n, h);
        for (int n = 1; n <= N_; ++n) {
            orc_lxyz(o, n, g, &L1, &L2, &L3);
            V2(o->L, n, g) = L1 + L2 + L3;
            src = src + V2(af, n, g) * (scg[n - 1] + CHI(MAT(n), g) * o->fs0[n - 1]) * VDEL(n);
            rem = rem + V2(af, n, g) * V2(sigrp, n, g) * V2(o->f0, n, g) * VDEL(n);
            lea = lea + V2(af, n, g) * V2(o->L, n, g) * VDEL(n);
            fde = fde + V2(af, n, g) * CHI(MAT(n), g) * o->fs0[n - 1] * VDEL(n);
        }
    }
    free(scg);
    return (src - lea - rem) / fde;
}

/* ======================================================================== state access */
int orc_set_kinetics(orc *o, const double *ibeta, const double *lamb, const double *velo,
                     const double *tbeta, double sth, double bth)
{
    memcpy(o->ibeta, ibeta, NF * sizeof(double)); memcpy(o->lamb, lamb, NF * sizeof(double));
    memcpy(o->velo, velo, G_ * sizeof(double)); memcpy(o->tbeta, tbeta, o->nmat * sizeof(double));
    o->sth = sth; o->bth = bth; o->bxtab = 0;
    return ORC_OK;
}
/* %XTAB decks: iBeta, lamb (NF per material) and velo (ng per material) of m(1:nmat), tbeta(nmat) */
int orc_set_kinetics_xtab(orc *o, const double *mibeta, const double *mlamb, const double *mvelo,
                          const double *tbeta, double sth, double bth)
{
    free(o->mlamb); free(o->mibeta); free(o->mvelo);
    o->mlamb = dalloc((size_t)NF * o->nmat); o->mibeta = dalloc((size_t)NF * o->nmat); o->mvelo = dalloc((size_t)G_ * o->nmat);
    memcpy(o->mibeta, mibeta, (size_t)NF * o->nmat * sizeof(double));
    memcpy(o->mlamb, mlamb, (size_t)NF * o->nmat * sizeof(double));
    memcpy(o->mvelo, mvelo, (size_t)G_ * o->nmat * sizeof(double));
    memcpy(o->tbeta, tbeta, o->nmat * sizeof(double));
    o->sth = sth; o->bth = bth; o->bxtab = 1;
    return ORC_OK;
}
/* transient state set by trans_calc before outer_tr (mod_trans.f90:398-416); NULL = keep */
int orc_set_transient(orc *o, const double *c0, const double *ft, const double *fst, const double *omeg,
                      const double *sigrp, const double *L)
{
    size_t NG = (size_t)N_ * G_;
    if (c0) memcpy(o->c0, c0, (size_t)N_ * NF * sizeof(double));
    if (ft) memcpy(o->ft, ft, NG * sizeof(double));
    if (fst) memcpy(o->fst, fst, N_ * sizeof(double));
    if (omeg) memcpy(o->omeg, omeg, NG * sizeof(double));
    if (sigrp) memcpy(o->sigrp, sigrp, NG * sizeof(double));
    if (L) memcpy(o->L, L, NG * sizeof(double));
    return ORC_OK;
}
int orc_get_transient(orc *o, double *c0, double *exsrc, double *dfis, double *L)
{
    size_t NG = (size_t)N_ * G_;
    if (c0) memcpy(c0, o->c0, (size_t)N_ * NF * sizeof(double));
    if (exsrc) memcpy(exsrc, o->exsrc, NG * sizeof(double));
    if (dfis) memcpy(dfis, o->dfis, N_ * sizeof(double));
    if (L) memcpy(L, o->L, NG * sizeof(double));
    return ORC_OK;
}
int orc_set_state(orc *o, const double *f0, const double *fs0, double Ke)
{
    size_t NG = (size_t)N_ * G_;
    if (f0) memcpy(o->f0, f0, NG * sizeof(double));
    if (fs0) memcpy(o->fs0, fs0, N_ * sizeof(double));
    o->Ke = Ke; o->have_state = 1;
    return ORC_OK;
}
int orc_set_s0(orc *o, const double *s0) { memcpy(o->s0, s0, (size_t)N_ * G_ * sizeof(double)); return ORC_OK; }
int orc_set_dfis(orc *o, const double *dfis) { memcpy(o->dfis, dfis, (size_t)N_ * sizeof(double)); return ORC_OK; }
int orc_get_state(orc *o, double *f0, double *fs0, double *s0, double *Ke, double *ser, double *fer)
{
    size_t NG = (size_t)N_ * G_;
    if (f0) memcpy(f0, o->f0, NG * sizeof(double));
    if (fs0) memcpy(fs0, o->fs0, N_ * sizeof(double));
    if (s0) memcpy(s0, o->s0, NG * sizeof(double));
    if (Ke) *Ke = o->Ke;
    if (ser) *ser = o->ser;
    if (fer) *fer = o->fer;
    return ORC_OK;
}
/* nod(n,g)%df / %dn repacked as df(6,nnod,ng), dn(6,nnod,ng) column-major (face fastest) */
int orc_get_nod(orc *o, double *df, double *dn)
{
    if (!o->nod) return -1;
    for (size_t r = 0; r < (size_t)N_ * G_; ++r)
        for (int f = 0; f < 6; ++f) {
            if (df) df[r * 6 + f] = o->nod[r].df[f];
            if (dn) dn[r * 6 + f] = o->nod[r].dn[f];
        }
    return ORC_OK;
}
int orc_set_nod_dn(orc *o, const double *dn)
{
    if (!o->nod) { o->nod = (node_data *)calloc((size_t)N_ * G_, sizeof(node_data)); }
    o->coup_first = 0;
    for (size_t r = 0; r < (size_t)N_ * G_; ++r)
        for (int f = 0; f < 6; ++f) o->nod[r].dn[f] = dn[r * 6 + f];
    return ORC_OK;
}
/* the matrix in 7-diagonal form for kernel-level comparisons: a(7,nnod,ng) column-major,
 * order z-,y-,x-,diag,x+,y+,z+ (set_ind order); absent neighbours give 0 */
int orc_get_matrix_dia(orc *o, double *a)
{
    for (int g = 1; g <= G_; ++g)
        for (int n = 1; n <= N_; ++n) {
            int i = IX(n), j = IY(n), k = IZ(n), rec = 0;
            double *dst = a + (((size_t)(g - 1)) * N_ + (n - 1)) * 7;
            const double *e = AM(n, g).elmn;
            dst[0] = (k != 1) ? e[rec++] : 0.0;
            dst[1] = (j != XSMIN(i)) ? e[rec++] : 0.0;
            dst[2] = (i != YSMIN(j)) ? e[rec++] : 0.0;
            dst[3] = e[rec++];
            dst[4] = (i != YSMAX(j)) ? e[rec++] : 0.0;
            dst[5] = (j != XSMAX(i)) ? e[rec++] : 0.0;
            dst[6] = (k != o->nzz) ? e[rec++] : 0.0;
        }
    return ORC_OK;
}
int orc_get_sources(orc *o, double *S1, double *S2, double *S3)
{
    size_t NG = (size_t)N_ * G_;
    memcpy(S1, o->S1, NG * sizeof(double)); memcpy(S2, o->S2, NG * sizeof(double)); memcpy(S3, o->S3, NG * sizeof(double));
    return ORC_OK;
}
/* get_source alone (for the kernel-level test of the Lxyz kernel) */
int orc_get_source(orc *o, int cmode) { o->cmode = cmode; get_source(o); return ORC_OK; }
/* get_ABEFGH for one node/direction: out[6*ng] = A,B,E,F,G,H */
int orc_abefgh(orc *o, int n, int u, double *out)
{
    int G = G_;
    get_ABEFGH(o, n, u, out, out + G, out + 2 * G, out + 3 * G, out + 4 * G, out + 5 * G);
    return ORC_OK;
}

double orc_ndmax(orc *o) { return o->ndmax; }
void orc_set_ndmax(orc *o, double v) { o->ndmax = v; }
void orc_get_ndloc(orc *o, int *ijk) { ijk[0] = o->im; ijk[1] = o->jm; ijk[2] = o->km; }
void orc_times(orc *o, double *fdm, double *nod) { *fdm = o->fdm_time; *nod = o->nod_time; }
void orc_reset_times(orc *o) { o->fdm_time = 0.0; o->nod_time = 0.0; }
int orc_trace(orc *o, int maxn, double *ke, double *ser, double *fer)
{
    int n = o->ntrace < maxn ? o->ntrace : maxn;
    memcpy(ke, o->tr_ke, n * sizeof(double)); memcpy(ser, o->tr_ser, n * sizeof(double));
    memcpy(fer, o->tr_fer, n * sizeof(double));
    return o->ntrace;
}
/* timing of the last outer*() call per iteration (test / bench infrastructure, no counterpart in the reference):
 * wall clock at the end of iteration q and the accumulated "CMFD" / "nodal update" CPU times the reference prints
 * in its breakdown (ADPRES.f90:55-81), so that a window of iterations of ONE outer() call can be timed */
int orc_trace_times(orc *o, int maxn, double *wall, double *fdm, double *nod)
{
    int n = o->ntrace < maxn ? o->ntrace : maxn;
    memcpy(wall, o->tr_wall, n * sizeof(double)); memcpy(fdm, o->tr_fdm, n * sizeof(double));
    memcpy(nod, o->tr_nod, n * sizeof(double));
    return o->ntrace;
}
int orc_nodal_trace(orc *o, int maxn, int *p, double *ndmax, int *im, int *jm, int *km)
{
    int n = o->nnodal < maxn ? o->nnodal : maxn;
    memcpy(p, o->nu_p, n * sizeof(int)); memcpy(ndmax, o->nu_ndmax, n * sizeof(double));
    memcpy(im, o->nu_im, n * sizeof(int)); memcpy(jm, o->nu_jm, n * sizeof(int)); memcpy(km, o->nu_km, n * sizeof(int));
    return o->nnodal;
}
int orc_extrp_trace(orc *o, int maxn, int *p)
{
    int n = o->nextrp < maxn ? o->nextrp : maxn;
    memcpy(p, o->ex_p, n * sizeof(int));
    return o->nextrp;
}
