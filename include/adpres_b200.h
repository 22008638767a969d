/*
 * adpres_b200.h -- C ABI of the B200-native ADPRES eigenvalue hot path.
 *
 * What it replaces.  ADPRES (Fortran 90) has no plugin / FFI interface: its modules USE
 * each other and share state through the global module `sdata` (src/mod_data.f90).  The
 * boundary therefore consists of the *bodies* of the hot-path procedures of
 *     src/mod_cmfd.f90   (outer, outer_ad, outer_fs, outer_th, outer_tr, PowDis, ...)
 *     src/mod_nodal.f90  (nodal_update, nodal_update_pnm, Lxyz)
 * The Fortran procedure names and argument lists stay; their bodies pack the `sdata`
 * arrays and call the functions below through ISO_C_BINDING (fortran/adpres_b200_bind.f90,
 * INTEGRATION.md).  Each entry point cites the reference code it stands for.
 *
 * Conventions
 *  - plain C: pointers + sizes only; every pointer is HOST memory owned by the caller.
 *  - arrays are Fortran column-major exactly as `sdata` stores them, node index fastest:
 *      f0(nnod,ng), sigs(nnod,ng,ng) [from g to h], dc(nnod,ng,6), chi(nmat,ng) ...
 *    node / mesh / material indices are 1-based as in the Fortran.
 *  - face index: 1=x+ (east) 2=x- 3=y+ (north) 4=y- 5=z+ (top) 6=z-  (mod_data.f90:59-60)
 *  - boundary codes: 0 zero flux, 1 zero incoming current, 2 reflective.
 *  - all arithmetic is fp64 (dp = selected_real_kind(10,15), mod_data.f90:5).
 *  - return value: 0 ok; >0 one of the reference's STOP conditions (ADP_STOP_*);
 *    <0 a CUDA / NCCL / usage error (adp_last_error() has the text).
 *  - one context per process and GPU; calls are synchronous and not re-entrant, like the
 *    SAVE'd single-threaded reference.  With adp_comm_init() the core is split in z-slabs
 *    over the ranks (one process per GPU); every rank passes the same GLOBAL arrays.
 *  - there is NO CPU fallback: without a CUDA device adp_create() fails.
 */
#ifndef ADPRES_B200_H
#define ADPRES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct adp_ctx adp_ctx;

/* ---- status codes ------------------------------------------------------------------- */
enum {
    ADP_OK = 0,
    ADP_STOP_MAXOUTER = 1,   /* "MAXIMUM NUMBER OF OUTER ITERATION IS REACHED" mod_cmfd.f90:498-505,589-596,688-695 */
    ADP_STOP_LU_DIAG = 2,    /* "ERROR IN MATRIX DECOMP: DIAGONAL ELEMENTS CLOSE TO ZERO" mod_nodal.f90:856-862 */
    ADP_STOP_NDMAX = 3,      /* "Max. change in nodal coupling coefficient" > 1e3, mod_nodal.f90:131-142 */
    ADP_STOP_ZERO_POWER = 4, /* "TOTAL NODES POWER IS ZERO OR LESS" mod_cmfd.f90:1322-1326 */
    ADP_STOP_STEAM_TABLE = 5, /* "ENTHALPY / MODERATOR TEMP. IS OUT OF THE RANGE ... IN THE STEAM TABLE" mod_th.f90:236-244,293-301 */
    ADP_STOP_XTAB_RANGE = 6, /* "... IS OUT OF THE RANGE OF THE BRANCH PARAMETER" mod_xsec.f90:569-574,594-599,619-624,644-649 */
    ADP_STOP_XS_CHECK = 8,   /* "Negative diffusion coefficient encountered" (Dsigr_updt, mod_xsec.f90:217) or one of check_xs's
                                "ERROR IN THE ... CROSS SECTION" STOPs (:104-151): sigtr < 1.e-5, D < 1.e-20, sigr / nuf / sigs < 0
                                after a device-side XS update */
    ADP_STOP_XTAB_NOROD = 7, /* "CONTROL ROD BANK ... COINCIDES WITH MATERIAL ... THAT DOES NOT HAVE CONTROL ROD DATA IN XTAB FILE" mod_xsec.f90:336-343 */
    ADP_ERR_CUDA = -1,
    ADP_ERR_USAGE = -2,
    ADP_ERR_NCCL = -3,
    ADP_ERR_UNSUPPORTED = -4
};

/* kern: the %KERN card (mod_io.f90:1593-1621) */
enum { ADP_KERN_FDM = 0, ADP_KERN_PNM = 1, ADP_KERN_SANM = 2 };

/* which outer iteration: selects TSrc/FSrc flavour, group sweep order and k-eff update */
enum {
    ADP_MODE_FORWARD = 0,  /* outer     mod_cmfd.f90:415-509  TSrc  + FSrc,   g = 1..G, Ke update   */
    ADP_MODE_ADJOINT = 1,  /* outer_ad  mod_cmfd.f90:602-699  TSrcAd+ FSrcAd, g = G..1, Ke update   */
    ADP_MODE_FIXEDSRC = 2, /* outer_fs  mod_cmfd.f90:513-598  TSrc  + FSrc,   no Ke update          */
    ADP_MODE_TRANSIENT = 3 /* outer_tr  mod_cmfd.f90:800-868  TSrcTr+ FSrc,   no Ke update          */
};

/* ---- lifecycle ---------------------------------------------------------------------- */
int adp_create(adp_ctx **ctx, int device /* -1: from ADP_LOCAL_RANK / LOCAL_RANK, else 0 */);
int adp_destroy(adp_ctx *ctx);
const char *adp_last_error(const adp_ctx *ctx);
const char *adp_version(void);

/* z-slab decomposition over ranks (one process per GPU).  The 128-byte NCCL unique id is
 * produced on rank 0 by adp_comm_unique_id() and distributed by the launcher (torchrun /
 * MPI / a file).  Must be called before adp_set_geometry().  Not calling it = 1 rank. */
int adp_comm_unique_id(void *uid128);
int adp_comm_init(adp_ctx *ctx, int nranks, int rank, const void *uid128);
/* the same from the environment (ADP_NRANKS|WORLD_SIZE, ADP_RANK|RANK, and ADP_UID_FILE = a path private
 * to this job, or ADP_JOB_ID|MASTER_PORT from which /tmp/adpres_b200.<job>.uid is derived; with neither the
 * call fails with ADP_ERR_USAGE): what the Fortran driver calls when it is started once per GPU */
int adp_comm_init_env(adp_ctx *ctx);
/* planes [k0, k1) (0-based) owned by this rank, valid after adp_set_geometry() */
int adp_slab(const adp_ctx *ctx, int *k0, int *k1);

/* ---- problem definition (what inp_geom1/2 + misc leave in sdata, mod_io.f90:809-1365) -- */
int adp_set_geometry(adp_ctx *ctx, int nxx, int nyy, int nzz, int nnod, int ng, int nmat,
                     const int *ix, const int *iy, const int *iz,          /* (nnod) */
                     const int *ystag_smin, const int *ystag_smax,        /* (nyy) ystag(j)%smin/smax */
                     const int *xstag_smin, const int *xstag_smax,        /* (nxx) xstag(i)%smin/smax */
                     const double *xdel, const double *ydel, const double *zdel,
                     const int bc[6] /* xeast, xwest, ynorth, ysouth, zbott, ztop */,
                     const int *mat /* (nnod) */);

/* node-wise cross sections as XS_updt leaves them (mod_xsec.f90:11-46).  A NULL pointer
 * keeps the values already on the device (e.g. only sigr changes in trans_calc,
 * mod_trans.f90:398-412). */
int adp_set_xs(adp_ctx *ctx, const double *D, const double *sigr, const double *nuf,
               const double *sigf, const double *sigs /* (nnod,ng,ng) */,
               const double *chi /* (nmat,ng) */, const double *dc /* (nnod,ng,6) */,
               const double *exsrc /* (nnod,ng) */);

/* The same for callers that cannot pass NULL (Fortran assumed-size dummies): only the arrays whose bit is set in
 * `mask` are read, the others keep the values on the device.  bit 0 D, 1 sigr, 2 nuf, 3 sigf, 4 sigs, 5 chi, 6 dc,
 * 7 exsrc.  The shim uses it to upload per outer*() call only what its caller changed: after the first call of a
 * run dc changes only through XStab_updt (mod_xsec.f90:50-86), exsrc only in outer_fs / on the device in outer_tr
 * (get_exsrc, mod_cmfd.f90:830), chi never. */
int adp_set_xs_mask(adp_ctx *ctx, int mask, const double *D, const double *sigr, const double *nuf,
                    const double *sigf, const double *sigs, const double *chi, const double *dc,
                    const double *exsrc);

/* %ITER and %KERN values (mod_io.f90:1522-1540; defaults mod_data.f90:78-86) */
int adp_set_control(adp_ctx *ctx, int nout, int nin, int nac, int nupd, double serc, double ferc,
                    int kern);

/* ---- the hot path, fine grained ------------------------------------------------------- */
/* matrix_setup(opt): opt > 0 recomputes the FDM coupling coefficients (coup_coef) first.
 * mod_cmfd.f90:217-304, 11-137.  The first call with opt > 0 zeroes nod%dn (:25-33). */
int adp_matrix_setup(adp_ctx *ctx, int opt);

/* first-call initialisation of outer*: Ke = 1, f0 = 1, fs0 = FSrc / FSrcAd.
 * mod_cmfd.f90:448-454, 635-641 */
int adp_init_flux(adp_ctx *ctx, int adjoint);

/* the statements before the `do p` loop: f = Integrate(fs0); errn = 1; e1 = Integrate(errn)
 * mod_cmfd.f90:457-459 (and 553-554, 644-646).  Call once per outer*() after matrix_setup. */
int adp_outer_begin(adp_ctx *ctx, int mode);

/* one pass of the `do p = 1, nout` loop body up to and including RelE/RelEg
 * (mod_cmfd.f90:467-487; 562-578; 654-674; 838-854): G x (TSrc* + bicg), FSrc*, errn, l2norm,
 * fission-source extrapolation when mod(p,nac)==0, Integrate, k-eff update, RelE, RelEg.
 * Returns the scalars the Fortran loop prints and tests. */
int adp_outer_iter(adp_ctx *ctx, int mode, int p, double *Ke, double *ser, double *fer);

/* nodal_upd(popt, nmode): ndmax = 0; nodal_update | nodal_update_pnm (cmode = nmode:
 * 0 adjoint, 1 forward, 2 transient); matrix_setup(0).  mod_cmfd.f90:339-383,
 * mod_nodal.f90:18-278.  im/jm/km = location of the maximum (ties: lowest node number). */
int adp_nodal_upd(adp_ctx *ctx, int nmode, double *ndmax, int *im, int *jm, int *km);

/* PowDis(p): nodal power, normalised to sum 1 (mod_cmfd.f90:1290-1333).  p(nnod). */
int adp_powdis(adp_ctx *ctx, double *p, int fixedsrc_mode);

/* Integrate(s) = sum vdel*s (mod_cmfd.f90:1120-1139); s(nnod) host array */
int adp_integrate(adp_ctx *ctx, const double *s, double *result);

/* ---- transient terms (outer_tr / get_exsrc, mod_cmfd.f90:800-952) ---------------------- */
/* kinetics data: iBeta(6), lamb(6), velo(ng), tbeta(nmat), sth, bth (mod_data.f90:120-127) */
int adp_set_kinetics(adp_ctx *ctx, const double *ibeta, const double *lamb, const double *velo,
                     const double *tbeta, double sth, double bth);
/* %XTAB decks (bxtab = 1): kinetics data per material -- m(mat)%iBeta(6), %lamb(6), %velo(ng) (mod_data.f90:187-188,
 * read by inp_xtab mod_io.f90:3800-3812) as (6,nmat), (6,nmat), (ng,nmat) column-major -- and tbeta(nmat).  Selects the
 * bxtab = 1 branches of get_exsrc (mod_cmfd.f90:898-925), iPden / uPden (mod_trans.f90:574-585,617-629) and of the
 * time-absorption term in adp_begin_time_step (:405-411): precursors only where nuf(n,ng) > 0. */
int adp_set_kinetics_xtab(adp_ctx *ctx, const double *mibeta, const double *mlamb, const double *mvelo,
                          const double *tbeta, double sth, double bth);
/* state of the previous time level saved by trans_calc (mod_trans.f90:398-416) and the
 * precursors / frequencies / leakage: c0(nnod,6), ft(nnod,ng), fst(nnod), omeg(nnod,ng),
 * sigrp(nnod,ng), L(nnod,ng).  NULL keeps the device copy. */
int adp_set_transient(adp_ctx *ctx, const double *c0, const double *ft, const double *fst,
                      const double *omeg, const double *sigrp, const double *L);
/* get_exsrc(ht, exsrc): fills exsrc and dfis on the device (mod_cmfd.f90:872-952, both bxtab branches) */
int adp_get_exsrc(adp_ctx *ctx, double ht);

/* ---- optional: XS_updt on the device for %XSEC (+ %CROD) decks --------------- */
/* material tables xsigtr..xsigf (nmat,ng), xsigs (nmat,ng,ng) [g -> h]           mod_io.f90:683-762 */
int adp_set_material_xs(adp_ctx *ctx, const double *xsigtr, const double *xsiga, const double *xnuf,
                        const double *xsigf, const double *xsigs);
/* %CROD: node-wise bank map fbmap(nxx,nyy), rod increments per material          mod_io.f90:2148-2330 */
int adp_set_crod(adp_ctx *ctx, int nb, double pos0, double ssize, const int *fbmap, const double *dsigtr,
                 const double *dsiga, const double *dnuf, const double *dsigf, const double *dsigs);
/* base_updt + crod_updt(bpos) + Dsigr_updt -> D, sigr, nuf, sigf, sigs on the device
 * (mod_xsec.f90:172-296); chi, dc, exsrc stay as adp_set_xs left them */
int adp_xs_update(adp_ctx *ctx, const double *bpos /* (nb) or NULL without rods */);
/* feedback cards %BCON / %CBCS (which 0), %FTEM (1), %MTEM (2), %CDEN (3): reference value and
 * d(sigtr, siga, nuf, sigf)(nmat,ng), dsigs(nmat,ng,ng) per unit change               mod_io.f90:2486-2957 */
int adp_set_feedback(adp_ctx *ctx, int which, double ref, const double *dsigtr, const double *dsiga, const double *dnuf,
                     const double *dsigf, const double *dsigs);
/* XS_updt(bcon, ftem, mtem, cden, bpos): base_updt, bcon_updt, ftem_updt (SQRT), mtem_updt, cden_updt,
 * crod_updt, Dsigr_updt (mod_xsec.f90:11-46,172-516); ftem / mtem / cden host (nnod) or NULL = the
 * thermal-hydraulic state on the device (adp_th_upd / adp_set_th_state) */
int adp_xs_update_th(adp_ctx *ctx, double bcon, const double *ftem, const double *mtem, const double *cden,
                     const double *bpos);
int adp_get_xs(adp_ctx *ctx, double *D, double *sigr, double *nuf, double *sigf, double *sigs);

/* ---- optional: XStab_updt on the device for %XTAB decks (branch-table cross sections) ---------- */
/* The tables inp_xtab reads into m(1:nmat) (MBRANCH, mod_data.f90:176-192; mod_io.f90:3648-4061):
 *   dims (4,nmat) column-major: nd, nb, nf, nm (coolant density, boron, fuel / moderator temperature)
 *   trod (nmat): 1 = the material has a rodded set
 *   par: pd(nd), pb(nb), pf(nf), pm(nm) of material 1, then of material 2, ... (a dimension of 1
 *        contributes one unused value, like branchPar)
 *   xs, rxs: per material one (nd,nb,nf,nm,nval) block, LAST index fastest, nval = 4 ng + ng*ng + 6 ng
 *        values per branch point packed [sigtr(ng), siga(ng), nuf(ng), sigf(ng), sigs(g,h) g slow,
 *        dc(g,face) g slow]; rxs blocks of materials without a rodded set are ignored (rxs may be
 *        NULL if no material has one).  chi goes through adp_set_xs as usual. */
int adp_set_xtab(adp_ctx *ctx, const int *dims, const int *trod, const double *par, const double *xs,
                 const double *rxs);
/* %CROD of an %XTAB deck: bank map only, the rodded cross sections are in the tables (mod_io.f90:2225-2232) */
int adp_set_crod_map(adp_ctx *ctx, int nb, double pos0, double ssize, const int *fbmap);
/* XStab_updt(bcon, ftem, mtem, cden, bpos): brInterp for every node, crod_tab_updt, Dsigr_updt
 * (mod_xsec.f90:50-86,300-390,520-788) -> D, sigr, nuf, sigf, sigs AND dc on the device.  ftem / mtem /
 * cden host (nnod) or NULL = the thermal-hydraulic state on the device.  Returns ADP_STOP_XTAB_RANGE /
 * ADP_STOP_XTAB_NOROD for the reference's two STOPs. */
int adp_xs_update_xtab(adp_ctx *ctx, double bcon, const double *ftem, const double *mtem, const double *cden,
                       const double *bpos);
int adp_get_dc(adp_ctx *ctx, double *dc /* (nnod,ng,6) */);

/* ---- optional: the time-step glue of mod_trans.f90 on the device ------------- */
/* With these a time step uploads only the new cross sections and reads back scalars.
 *   adp_save_adjoint      af = f0 after outer_ad                       mod_trans.f90:65-66
 *   adp_ipden             iPden                                        mod_trans.f90:561-597
 *   adp_update_omeg       omeg = bextr ? LOG(f0/ft)/ht : 0 (%EXTR)        mod_trans.f90:128-134,150-154
 *   adp_begin_time_step   sigrp = sigr; sigr += 1/(sth v ht) + omeg/v; ft = f0; fst = fs0   :398-416
 *   adp_upden             uPden(ht)                                    mod_trans.f90:601-644
 *   adp_powtot            PowTot(f0, tpow)                             mod_trans.f90:523-557
 *   adp_reactivity        reactivity(af, sigr|sigrp, rho), fills L     mod_trans.f90:648-688   */
int adp_save_adjoint(adp_ctx *ctx);
int adp_ipden(adp_ctx *ctx);
int adp_update_omeg(adp_ctx *ctx, double ht, int bextr);
int adp_begin_time_step(adp_ctx *ctx, double ht);
int adp_upden(adp_ctx *ctx, double ht);
int adp_powtot(adp_ctx *ctx, double *tpow);
int adp_reactivity(adp_ctx *ctx, int use_sigrp, double *rho);

/* ---- optional: result reductions on the device ------------------------------- */
/* The reference prints assembly / axial averages of PowDis and of the flux through dense
 * fx(nxx,nyy,nzz[,ng]) boxes on the host.  Here the O(nnod) stage (column sums over z, plane
 * sums) runs on the device and only np or nzz values come back; the assembly loops and the
 * normalisation follow in the reference's order.  Outputs are Fortran column-major.
 *   adp_asm_pow   CALL PowDis(pow); CALL AsmPow(pow): fasm(nx,ny), location of the maximum   mod_io.f90:3267-3405
 *   adp_axi_pow   CALL PowDis(pow); CALL AxiPow(pow): faxi(nz), plane of the maximum         mod_io.f90:3409-3494
 *   adp_asm_flux  CALL AsmFlux(f0 [, norm]): fasm(nx,ny,ng), negf = 1 if a value is negative  mod_io.f90:3498-3644 */
int adp_asm_pow(adp_ctx *ctx, int nx, int ny, const int *xdiv, const int *ydiv, double *fasm, int *xmax, int *ymax);
int adp_axi_pow(adp_ctx *ctx, int nz, const int *zdiv, double *faxi, int *amax);
int adp_asm_flux(adp_ctx *ctx, int nx, int ny, const int *xdiv, const int *ydiv, int use_norm, double norm,
                 double *fasm, int *negf);

/* ---- optional: thermal-hydraulic channel solve on the device ---------------- */
/* th_upd / th_trans of mod_th.f90: per radial channel an axial enthalpy march, per node a 13-point
 * tridiagonal solve of the radial pin conduction.  The XS feedback (XS_updt with ftem, mtem, cden,
 * bcon) and the th_iter / cbsearch loops stay with the caller.
 *   adp_set_th        what inp_ther leaves in sdata: pi (the reference's default-REAL 3.14159265),
 *                     rf, rg, rc, dia, dh, farea, cflow, cf, tin, rpos(12), rdel(12), stab(ntem,6)   mod_io.f90:2958-3174
 *   adp_set_th_state  tfm(nnod,13), heatf, ent, ftem, mtem, cden, frate (NULL = keep; frate NULL before
 *                     the first transient step = cflow)                                              mod_io.f90:3101-3113
 *   adp_th_pline      CALL PowDis(npow); pline = npow*pow*ppow*0.01/(node_nf*zdel) (form 0, th_iter,
 *                     mod_th.f90:57-64) or npow*pow*xppow/(node_nf*zdel) (form 1, trans_calc,
 *                     mod_trans.f90:430-441), kept on the device
 *   adp_th_upd        CALL th_upd(pline); th_err = AbsE(ftem, otem) if requested                     mod_th.f90:594-699,94-119
 *   adp_th_trans      CALL th_trans(pline, h)                                                        mod_th.f90:440-591
 * xpline: host (nnod) linear power density in W/cm, or NULL = the one adp_th_pline left on the device. */
int adp_set_th(adp_ctx *ctx, double pi, double rf, double rg, double rc, double dia, double dh, double farea, double cflow,
               double cf, double tin, const double *rpos, const double *rdel, int ntem, const double *stab);
int adp_set_th_state(adp_ctx *ctx, const double *tfm, const double *heatf, const double *ent, const double *ftem,
                     const double *mtem, const double *cden, const double *frate);
int adp_get_th_state(adp_ctx *ctx, double *tfm, double *heatf, double *ent, double *ftem, double *mtem, double *cden,
                     double *frate);
int adp_th_pline(adp_ctx *ctx, double pow, double ppow, int form, const double *node_nf);
int adp_th_upd(adp_ctx *ctx, const double *xpline, double *th_err);
int adp_th_trans(adp_ctx *ctx, const double *xpline, double h);

/* ---- state exchange with the Fortran side ----------------------------------------------- */
/* f0(nnod,ng), fs0(nnod), s0(nnod,ng); NULL = skip.  (drivers read them after outer*)
 * On several ranks every node array that comes back to the host (this call, adp_powdis, adp_get_nod, ...) is
 * completed with the other ranks' slabs, because the unchanged Fortran drivers consume whole sdata arrays on every
 * rank (th_upd's axial march, reactivity, the printers); adp_set_option(ctx, "gather_results", 0) returns only the
 * owned slab.  All ranks must make these calls together. */
int adp_get_state(adp_ctx *ctx, double *f0, double *fs0, double *s0, double *Ke);
/* the same with a mask instead of NULL pointers: bit 0 f0, 1 fs0, 2 s0 (Fortran callers).  No caller of outer*()
 * reads s0 -- only get_exsrc does, which runs on the device -- so the shim leaves bit 2 clear. */
int adp_get_state_mask(adp_ctx *ctx, int mask, double *f0, double *fs0, double *s0, double *Ke);
/* restart / KNE1 paths: load flux, fission source and k-eff (NULL keeps) */
int adp_set_state(adp_ctx *ctx, const double *f0, const double *fs0, double Ke);
/* restart path for s0(nnod,ng): column g (1-based) is loaded, the others are zero, which is the
 * only shape TSrc* ever leave behind (mod_cmfd.f90:1022); g = 0 clears s0 */
int adp_set_s0(adp_ctx *ctx, const double *s0, int g);
/* nod(n,g)%df / %dn as df(6,nnod,ng), dn(6,nnod,ng) (Lxyz in reactivity, mod_trans.f90:677) */
int adp_get_nod(adp_ctx *ctx, double *df, double *dn);
int adp_set_nod_dn(adp_ctx *ctx, const double *dn);
/* L(nnod,ng) = L1+L2+L3 of Lxyz for every node, evaluated on the device and left there as the L
 * that get_exsrc uses (reactivity, mod_trans.f90:677-678; saves copying nod%df/dn back) */
int adp_lxyz_total(adp_ctx *ctx, double *L);
/* exsrc(nnod,ng), dfis(nnod) after adp_get_exsrc */
int adp_get_exsrc_arrays(adp_ctx *ctx, double *exsrc, double *dfis);
/* ndmax persists across outer*() calls and starts at 0 (mod_data.f90:199) */
int adp_get_ndmax(adp_ctx *ctx, double *ndmax);
/* sdata's ser / fer after the last outer iteration (th_iter, cbsearch read them: mod_th.f90:77,789) */
int adp_get_errors(adp_ctx *ctx, double *ser, double *fer);

/* ---- whole procedures (host loop in C++, adpres_b200/csrc/host_cmfd.cpp) ---------------- */
/* Trace callback: called once per outer iteration with what the reference prints
 * (mod_cmfd.f90:491-494); event: 0 iteration line, 1 "FISSION SOURCE EXTRAPOLATED",
 * 2 "NODAL COUPLING UPDATED" (then a = ndmax, i,j,k = location). */
typedef void (*adp_trace_fn)(void *user, int event, int p, double a, double b, double c, int i,
                             int j, int k);
int adp_set_trace(adp_ctx *ctx, adp_trace_fn fn, void *user);

/* outer(popt) / outer_ad(popt) / outer_fs(popt): mod_cmfd.f90:415-509 / 602-699 / 513-598.
 * niter receives the number of outer iterations done. */
int adp_outer(adp_ctx *ctx, int popt, int *niter);
int adp_outer_ad(adp_ctx *ctx, int popt, int *niter);
int adp_outer_fs(adp_ctx *ctx, int popt, int *niter);
/* outer_th(maxn): mod_cmfd.f90:703-796 (at most maxn iterations, no STOP) */
int adp_outer_th(adp_ctx *ctx, int maxn, int *niter);
/* outer_tr(ht, maxi): mod_cmfd.f90:800-868 */
int adp_outer_tr(adp_ctx *ctx, double ht, int *maxi, int *niter);

/* ---- kernel-level entry points (used by the parity tests and the micro-benchmarks) ------ */
/* v = A_g x  (sp_matvec, mod_cmfd.f90:1247-1268); g 1-based; x, v host (nnod) */
int adp_sp_matvec(adp_ctx *ctx, int g, const double *x, double *v);
/* bicg(imax, g, b, x): mod_cmfd.f90:1203-1243; x in/out host (nnod) */
int adp_bicg(adp_ctx *ctx, int imax, int g, const double *b, double *x);
/* the assembled matrix as a(7,nnod,ng): z-,y-,x-,diag,x+,y+,z+ (set_ind order) */
int adp_get_matrix(adp_ctx *ctx, double *a);
/* get_source (mod_nodal.f90:1009-1043): S1,S2,S3 (nnod,ng) for cmode */
int adp_get_source(adp_ctx *ctx, int cmode, double *S1, double *S2, double *S3);

/* runtime options (no reference counterpart; every default is the measured best):
 *   "graphs" 0/1            CUDA-graph replay of an outer iteration (default 1)
 *   "grid_blocks" n         persistent grid size (default: SM count x resident CTAs of each kernel); also another grouping
 *                           of the partial sums of the global reductions (tools/order_probe.py)
 *   "gather_results" 0/1    several ranks: complete node arrays returned to the host with the other ranks' slabs (default 1)
 *   "peer_push", "peer_allreduce", "fuse_mail", "mail_ll", "mail_timeout_s"   multi-rank data plane (csrc/comm.cu, mail.cuh)
 *   "st_var", "st_m_var", "spmv_var", "st_tma", "fuse_st", "balance_rounds"   formulations of the BiCGSTAB kernels kept for A/B
 *   "nodal_coop" -1/0/1/2   nodal kernels: automatic / one thread per item / 16 lanes per surface / quad kernels
 *   "nodal_fused" 0/1       fused per-direction nodal kernels for G <= 2 (experiment, slower)
 *   "lazy_adf" 0/1          adp_set_xs defers the upload of dc and sigf -- the arrays the CMFD iteration never reads -- to
 *                           adp_outer_begin, on a second stream; the first consumer (nodal update, PowDis, XS update, getters)
 *                           waits for it.  Both host arrays must stay unchanged until such a consumer has returned
 *                           (default 0: every array is on the device when adp_set_xs returns)
 *   "reset_nodal" 1         back to the state before the first coup_coef call: the next adp_matrix_setup(1) zeroes dn
 *   "profile" 0/1           per-launch-site events, read with adp_profile_report
 *   "bench_warmup" n        untimed launches in front of adp_bench_kernel's timed ones */
int adp_set_option(adp_ctx *ctx, const char *name, int value);

/* ---- measurement ------------------------------------------------------------------------ */
/* number of kernel launches issued by this context so far (graph launches count the kernels
 * inside) and device time accumulated per class, in ms, measured with CUDA events */
int adp_launch_count(const adp_ctx *ctx, long long *launches);
/* device-resident micro-benchmark of one kernel class on the current problem, no host copies:
 * what: 0 B SpMV + (rs,v); 1 C fused s/t; 2 D x,r update; 3 A p update; 4 P source+residual;
 * 5 F fission source + norms; 6 nodal source; 7 whole nodal update; 8 plain SpMV; 9 matrix_setup(0);
 * 10..14 the multi-rank forms of B, C, D, A, P launched on this rank alone (no mailbox wait).
 * Runs `reps` launches (alternating over groups), returns the average device ms per launch. */
int adp_bench_kernel(adp_ctx *ctx, int what, int reps, double *avg_ms);
/* Measurement aid (no reference counterpart): after adp_set_option(ctx, "profile", 1) every kernel launch of the CMFD path
 * (mod_cmfd.f90:415-509 outer loop body) is followed by an event; the report sums the time between consecutive events per
 * launch site (source line of csrc/cmfd_kernels.cu), i.e. kernel + the gap before it -- inside a multi-rank step this
 * includes the time a kernel waits for its peers.  Use with option "graphs" = 0.  Returns the number of sites (<= max). */
int adp_profile_report(adp_ctx *ctx, int max, int *lines, int *counts, double *ms);
/* `nsteps` passes of the outer loop body starting at p_first, nodal update + matrix_setup(0)
 * whenever mod(p,nupd)==0, enqueued back to back with ONE synchronisation at the end (no
 * per-iteration exit test): the device-resident throughput path. */
int adp_outer_steps(adp_ctx *ctx, int mode, int p_first, int nsteps, double *Ke, double *ser, double *fer);
/* CUDA-event stopwatch on the library's stream */
int adp_timer_start(adp_ctx *ctx);
int adp_timer_stop(adp_ctx *ctx, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* ADPRES_B200_H */
