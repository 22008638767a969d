// Thermal-hydraulic channel solve: th_upd and th_trans of mod_th.f90:440-699.
//
// Every radial channel (i,j) is independent: its coolant enthalpy is marched from the inlet
// plane upwards (entm(i,j), in transients also the flow rate bfrate(i,j), are handed from plane
// to plane) and every node solves a 13-point tridiagonal system for the radial temperature
// profile of its average fuel pin.  One thread per plane position r walks up the planes of this
// rank's z-slab; loads and stores are coalesced across r because all state is [col][NV].
// With several ranks the march is a chain: a slab receives entm / bfrate from the slab below
// (ncclRecv), marches, and passes them on (ncclSend).
//
// Arithmetic follows the reference statement by statement (operand order, repeated products for
// t**2 / t**3, the default-REAL pi, geths called with its arguments in the reference's order).
// Only pow() for Pr**0.4, Re**0.8 comes from the device math library (<= 2 ulp).
#include "adp_internal.cuh"

#include <cmath>

namespace {

constexpr int TH_NT = 12;            // nm + 2 radial meshes (mod_data.f90:159-160), nt + 1 unknowns

struct ThArgs {
    adp_ctx::ThPar P;
    const double *stab;              // (ntem, 6) column-major: T, rho, h, Pr, kv, k
    const double *xpline;            // [NV] linear power density (W/cm)
    double *tfm;                     // [nt+1][NV]
    double *heatf, *ent, *ftem, *mtem, *cden, *frate;
    double *chain;                   // [2][np]: entm, bfrate from the slab below (in) / for the slab above (out)
    double h;                        // time step (transient)
    int *errflag;
};

__device__ __forceinline__ long long node_idx(const Geo &G, int kl, int r)
{
    return (long long)(kl + ADP_GH) * G.np + r;
}

// gettd (mod_th.f90:261-317); false = enthalpy outside the table by more than 10 %
__device__ __forceinline__ bool gettd(const double *__restrict__ stab, int ntem, double ent, double &t, double &rho,
                                      double &prx, double &kvx, double &tcx, double &Rx)
{
    const double *h = stab + 2 * ntem;
    int i1 = -1, i2 = -1;
    bool ok = true;
    if (ent >= h[0] && ent <= h[ntem - 1]) {
        for (int i = 1; i < ntem; ++i)
            if (ent >= h[i - 1] && ent <= h[i]) { i1 = i - 1; i2 = i; break; }
    } else if (ent < h[0] && (h[0] - ent) / h[0] < 0.1) {
        i1 = 0; i2 = 1;
    } else if (ent > h[ntem - 1] && (ent - h[ntem - 1]) / h[ntem - 1] < 0.1) {
        i1 = ntem - 2; i2 = ntem - 1;
    } else {
        ok = false; i1 = 0; i2 = 1;          // keep going with finite numbers; the caller reports the STOP
    }
    const double ratx = (ent - h[i1]) / (h[i2] - h[i1]);
    t = stab[i1] + ratx * (stab[i2] - stab[i1]);
    rho = stab[ntem + i1] + ratx * (stab[ntem + i2] - stab[ntem + i1]);
    prx = stab[3 * ntem + i1] + ratx * (stab[3 * ntem + i2] - stab[3 * ntem + i1]);
    kvx = stab[4 * ntem + i1] + ratx * (stab[4 * ntem + i2] - stab[4 * ntem + i1]);
    tcx = stab[5 * ntem + i1] + ratx * (stab[5 * ntem + i2] - stab[5 * ntem + i1]);
    Rx = 1000.0 * (stab[ntem + i2] - stab[ntem + i1]) / (h[i2] - h[i1]);
    return ok;
}

__device__ __forceinline__ double getkc(double t) { return 7.51 + 2.09e-2 * t - 1.45e-5 * (t * t) + 7.67e-9 * (t * t * t); }
__device__ __forceinline__ double getkf(double t) { return 1.05 + 2150.0 / (t - 73.15); }
__device__ __forceinline__ double getcpc(double t) { return 252.54 + 0.11474 * t; }
__device__ __forceinline__ double getcpf(double t) { return 162.3 + 0.3038 * t - 2.391e-4 * (t * t) + 6.404e-8 * (t * t * t); }

// geths (mod_th.f90:413-436) as the reference CALLS it: geths(cden, Pr, kv, tcon) against the
// dummy arguments (xden, tc, kv, Pr) -- inside, the Nusselt number is built from the value
// passed last (the conductivity) and the result scaled by the value passed second (the Prandtl
// number).  Replicated, not corrected.
__device__ __forceinline__ double geths(const adp_ctx::ThPar &P, double xden, double second, double kv, double last)
{
    const double cvelo = P.cflow / (P.farea * xden * 1000.0);
    const double Re = cvelo * P.dh / (kv * 1.e-6);
    const double Nu = 0.023 * pow(last, 0.4) * pow(Re, 0.8);
    return (second / P.dh) * Nu;
}

template <bool TRANS>
__global__ void __launch_bounds__(ADP_TILE) k_th_march(Geo G, ThArgs A)
{
    const int r = blockIdx.x * ADP_TILE + threadIdx.x;
    if (r >= G.np) return;
    const adp_ctx::ThPar &P = A.P;
    const long long NV = G.NV;
    const double Hg = 1.e4, fdens = 10.412e3, cdens = 6.6e3, alp = 0.7;
    bool ok = true;
    // what the channel carries across planes: enthalpy (and flow rate) at the lower node boundary
    double entm = (G.k0 == 0) ? P.enti : A.chain[r];
    double bfr = (G.k0 == 0) ? P.cflow : A.chain[G.np + r];
    for (int kl = 0; kl < G.nzl; ++kl) {
        const long long idx = node_idx(G, kl, r);
        const double zdel = G.hz[1 + G.k0 + kl];
        const double xpl = A.xpline[idx];
        const double cpline = A.heatf[idx] * P.pi * P.dia + P.cf * xpl * 100.0;
        double ent, mt, rho, Pr, kv, tcon, R;
        if (!TRANS) {
            const double zd = zdel * 0.01;
            ent = entm + 0.5 * cpline * zd / P.cflow;
            ok = gettd(A.stab, P.ntem, ent, mt, rho, Pr, kv, tcon, R) && ok;
            entm = 2.0 * ent - entm;
        } else {
            const double mdens = A.cden[idx] * 1000.0;
            const double vol = P.farea * zdel * 0.01;
            const double fr = A.frate[idx], entp = A.ent[idx];
            const double eps = mdens * vol / A.h;
            ent = (cpline * zdel * 0.01 + 2.0 * fr * entm + eps * entp) / (eps + 2.0 * fr);
            ok = gettd(A.stab, P.ntem, ent, mt, rho, Pr, kv, tcon, R) && ok;
            entm = 2.0 * ent - entm;
            const double frn = bfr - 0.5 * vol / A.h * R * (ent - entp);
            A.frate[idx] = frn;
            bfr = 2.0 * frn - bfr;
        }
        A.ent[idx] = ent; A.mtem[idx] = mt; A.cden[idx] = rho;
        const double hs = geths(P, rho, Pr, kv, tcon);
        const double pdens = TRANS ? 100.0 * xpl / (P.pi * (P.rf * P.rf))
                                   : (1.0 - P.cf) * 100.0 * xpl / (P.pi * (P.rf * P.rf));
        // ---- tridiagonal system of the radial pin conduction (mod_th.f90:641-686 / 511-574)
        double tf[TH_NT + 1], a[TH_NT + 1], b[TH_NT + 1], c[TH_NT + 1], d[TH_NT + 1];
#pragma unroll
        for (int i = 0; i <= TH_NT; ++i) { tf[i] = A.tfm[(size_t)i * NV + idx]; a[i] = 0.0; c[i] = 0.0; }
        double kt1 = getkf(tf[0]), kt2 = getkf(tf[1]);
        double kt = 2.0 * kt1 * kt2 / (kt1 + kt2);
        double xc = kt * P.rpos[0] / P.rdel[0], xa, eta = 0.0;
        if (TRANS) eta = fdens * getcpf(tf[0]) * (P.rpos[0] * P.rpos[0]) / (2.0 * A.h);
        b[0] = TRANS ? xc + eta : xc;
        c[0] = -xc;
        d[0] = TRANS ? pdens * 0.5 * (P.rpos[0] * P.rpos[0]) + eta * tf[0] : pdens * 0.5 * (P.rpos[0] * P.rpos[0]);
#pragma unroll
        for (int i = 1; i <= TH_NT - 3; ++i) {                      // Fortran rows 2 .. nt-2
            kt1 = kt2;
            kt2 = getkf(tf[i + 1]);
            kt = 2.0 * kt1 * kt2 / (kt1 + kt2);
            xa = xc;
            xc = kt * P.rpos[i] / P.rdel[i];
            const double ring = P.rpos[i] * P.rpos[i] - P.rpos[i - 1] * P.rpos[i - 1];
            a[i] = -xa;
            c[i] = -xc;
            if (TRANS) {
                eta = fdens * getcpf(tf[i]) * ring / (2.0 * A.h);
                b[i] = xa + xc + eta;
                d[i] = pdens * 0.5 * ring + eta * tf[i];
            } else {
                b[i] = xa + xc;
                d[i] = pdens * 0.5 * ring;
            }
        }
        {   // fuel-gap interface, row nt-1
            constexpr int i = TH_NT - 2;
            xa = xc;
            xc = P.rg * Hg;
            const double ring = P.rf * P.rf - P.rpos[i - 1] * P.rpos[i - 1];
            a[i] = -xa;
            c[i] = -xc;
            if (TRANS) {
                eta = fdens * getcpf(tf[i]) * ring / (2.0 * A.h);
                b[i] = xa + xc + eta;
                d[i] = pdens * 0.5 * ring + eta * tf[i];
            } else {
                b[i] = xa + xc;
                d[i] = pdens * 0.5 * ring;
            }
        }
        {   // gap-cladding interface, row nt
            constexpr int i = TH_NT - 1;
            kt1 = getkc(tf[i]); kt2 = getkc(tf[i + 1]);
            kt = 2.0 * kt1 * kt2 / (kt1 + kt2);
            xa = xc;
            xc = kt * P.rpos[i] / P.rdel[i];
            a[i] = -xa;
            c[i] = -xc;
            if (TRANS) {
                eta = cdens * getcpc(tf[i]) * (P.rpos[i] * P.rpos[i] - P.rg * P.rg) / (2.0 * A.h);
                b[i] = xa + xc + eta;
                d[i] = eta * tf[i];
            } else {
                b[i] = xa + xc;
                d[i] = 0.0;
            }
        }
        {   // cladding-coolant interface, row nt+1
            constexpr int i = TH_NT;
            xa = xc;
            a[i] = -xa;
            if (TRANS) {
                eta = cdens * getcpc(tf[i]) * (P.rc * P.rc - P.rpos[i - 1] * P.rpos[i - 1]) / (2.0 * A.h);
                xc = P.rc * hs;
                b[i] = xa + xc + eta;
                d[i] = P.rc * hs * mt + eta * tf[i];
            } else {
                b[i] = xa + hs * P.rc;
                d[i] = P.rc * hs * mt;
            }
        }
        // TridiaSolve (mod_th.f90:380-409)
        c[0] = c[0] / b[0];
        d[0] = d[0] / b[0];
#pragma unroll
        for (int i = 1; i <= TH_NT; ++i) {
            const double den = b[i] - a[i] * c[i - 1];
            c[i] = c[i] / den;
            d[i] = (d[i] - a[i] * d[i - 1]) / den;
        }
        tf[TH_NT] = d[TH_NT];
#pragma unroll
        for (int i = TH_NT - 1; i >= 0; --i) tf[i] = d[i] - c[i] * tf[i + 1];
#pragma unroll
        for (int i = 0; i <= TH_NT; ++i) A.tfm[(size_t)i * NV + idx] = tf[i];
        A.ftem[idx] = (1.0 - alp) * tf[0] + alp * tf[TH_NT - 2];
        A.heatf[idx] = hs * (tf[TH_NT] - mt);
    }
    A.chain[r] = entm;
    A.chain[G.np + r] = bfr;
    if (!ok) atomicExch(A.errflag, ADP_STOP_STEAM_TABLE);
}

// th_iter (mod_th.f90:61-64): pline = npow * pow * ppow * 0.01 / (node_nf * zdel)
// trans_calc (mod_trans.f90:439-440): pline = npow * pow * xppow / (node_nf * zdel)
__global__ void __launch_bounds__(ADP_TILE) k_th_pline(Geo G, const double *__restrict__ npow, double pw, double ppow, int form,
                                                        const double *__restrict__ nodenf, double *__restrict__ pline)
{
    const int r = blockIdx.x * ADP_TILE + threadIdx.x;
    if (r >= G.np) return;
    const double nf = nodenf[r];
    for (int kl = 0; kl < G.nzl; ++kl) {
        const long long idx = node_idx(G, kl, r);
        const double zd = G.hz[1 + G.k0 + kl];
        pline[idx] = form ? npow[idx] * pw * ppow / (nf * zd) : npow[idx] * pw * ppow * 0.01 / (nf * zd);
    }
}

// AbsE (mod_th.f90:94-119): max |new - old| over entries with |new| > 1e-10; one CTA, fixed order
__global__ void __launch_bounds__(ADP_TILE) k_th_abse(Geo G, const double *__restrict__ fnew, const double *__restrict__ fold,
                                                       double *__restrict__ out)
{
    __shared__ double sm[ADP_TILE];
    double m = 0.0;
    for (long long q = threadIdx.x; q < (long long)G.nzl * G.np; q += ADP_TILE) {
        const long long idx = q + (long long)ADP_GH * G.np;
        if (fabs(fnew[idx]) > 1.e-10) m = fmax(m, fabs(fnew[idx] - fold[idx]));
    }
    sm[threadIdx.x] = m;
    __syncthreads();
    for (int o = ADP_TILE / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

}  // namespace

#define TRY(x)              \
    do {                    \
        int rc__ = (x);     \
        if (rc__) return rc__; \
    } while (0)

int adp_th_alloc(adp_ctx *c)
{
    if (c->d_tfm) return ADP_OK;
    const size_t NV = (size_t)c->NV;
    double **one[] = {&c->d_heatf, &c->d_ent, &c->d_ftem, &c->d_mtem, &c->d_cden, &c->d_frate, &c->d_pline};
    CUDA_TRY(c, cudaMalloc((void **)&c->d_tfm, (TH_NT + 1) * NV * sizeof(double)));
    CUDA_TRY(c, cudaMemsetAsync(c->d_tfm, 0, (TH_NT + 1) * NV * sizeof(double), c->stream));
    for (double **q : one) {
        CUDA_TRY(c, cudaMalloc((void **)q, NV * sizeof(double)));
        CUDA_TRY(c, cudaMemsetAsync(*q, 0, NV * sizeof(double), c->stream));
    }
    CUDA_TRY(c, cudaMalloc((void **)&c->d_nodenf, (size_t)c->np * sizeof(double)));
    CUDA_TRY(c, cudaMalloc((void **)&c->d_chain, (size_t)2 * c->np * sizeof(double)));
    return ADP_OK;
}

extern "C" int adp_set_th(adp_ctx *c, double pi, double rf, double rg, double rc, double dia, double dh, double farea,
                          double cflow, double cf, double tin, const double *rpos, const double *rdel, int ntem,
                          const double *stab)
{   // what inp_ther leaves in sdata (mod_io.f90:2958-3174); nt = 12 radial meshes
    if (!c || !rpos || !rdel || !stab) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_th: geometry not set");
    ADP_REQUIRE(c, ntem >= 2 && ntem <= 64, "adp_set_th: steam table needs 2..64 rows");
    CUDA_TRY(c, cudaSetDevice(c->device));
    adp_ctx::ThPar &P = c->th;
    P.pi = pi; P.rf = rf; P.rg = rg; P.rc = rc; P.dia = dia; P.dh = dh; P.farea = farea; P.cflow = cflow; P.cf = cf;
    P.tin = tin; P.ntem = ntem;
    for (int i = 0; i < TH_NT; ++i) { P.rpos[i] = rpos[i]; P.rdel[i] = rdel[i]; }
    // getent(tin, enti) (mod_th.f90:219-258): the reference STOPs outside the table
    const double *T = stab, *H = stab + 2 * ntem;
    if (tin < T[0] || tin > T[ntem - 1]) {
        c->err = "ERROR : MODERATOR TEMP. IS OUT OF THE RANGE OF DATA IN THE STEAM TABLE";
        return ADP_STOP_STEAM_TABLE;
    }
    {
        double t2 = T[0], e2 = H[0];
        for (int i = 1; i < ntem; ++i) {
            const double t1 = t2, e1 = e2;
            t2 = T[i]; e2 = H[i];
            if (tin >= t1 && tin <= t2) { P.enti = e1 + (tin - t1) / (t2 - t1) * (e2 - e1); break; }
        }
    }
    if (c->d_stab) { cudaFree(c->d_stab); c->d_stab = nullptr; }
    CUDA_TRY(c, cudaMalloc((void **)&c->d_stab, (size_t)ntem * 6 * sizeof(double)));
    CUDA_TRY(c, cudaMemcpy(c->d_stab, stab, (size_t)ntem * 6 * sizeof(double), cudaMemcpyHostToDevice));
    TRY(adp_th_alloc(c));
    c->th_set = true;
    return ADP_OK;
}

extern "C" int adp_set_th_state(adp_ctx *c, const double *tfm, const double *heatf, const double *ent, const double *ftem,
                                const double *mtem, const double *cden, const double *frate)
{   // tfm(nnod, nt+1), the others (nnod); NULL keeps the device copy.  frate = NULL before the first
    // adp_th_trans means frate = cflow (mod_th.f90:482-486).
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->th_set, "adp_set_th_state: call adp_set_th first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (tfm) TRY(adp_upload_nodes(c, c->d_tfm, tfm, TH_NT + 1));
    if (heatf) TRY(adp_upload_nodes(c, c->d_heatf, heatf, 1));
    if (ent) TRY(adp_upload_nodes(c, c->d_ent, ent, 1));
    if (ftem) TRY(adp_upload_nodes(c, c->d_ftem, ftem, 1));
    if (mtem) TRY(adp_upload_nodes(c, c->d_mtem, mtem, 1));
    if (cden) TRY(adp_upload_nodes(c, c->d_cden, cden, 1));
    if (frate) TRY(adp_upload_nodes(c, c->d_frate, frate, 1));
    else if (!c->th_state_set) {
        std::vector<double> fr((size_t)c->nnod, c->th.cflow);
        TRY(adp_upload_nodes(c, c->d_frate, fr.data(), 1));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->th_state_set = true;
    return ADP_OK;
}

extern "C" int adp_get_th_state(adp_ctx *c, double *tfm, double *heatf, double *ent, double *ftem, double *mtem,
                                double *cden, double *frate)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->th_state_set, "adp_get_th_state: no thermal-hydraulic state");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (tfm) TRY(adp_download_nodes(c, tfm, c->d_tfm, TH_NT + 1));
    if (heatf) TRY(adp_download_nodes(c, heatf, c->d_heatf, 1));
    if (ent) TRY(adp_download_nodes(c, ent, c->d_ent, 1));
    if (ftem) TRY(adp_download_nodes(c, ftem, c->d_ftem, 1));
    if (mtem) TRY(adp_download_nodes(c, mtem, c->d_mtem, 1));
    if (cden) TRY(adp_download_nodes(c, cden, c->d_cden, 1));
    if (frate) TRY(adp_download_nodes(c, frate, c->d_frate, 1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_th_pline(adp_ctx *c, double pw, double ppow, int form, const double *node_nf)
{   // CALL PowDis(npow) + the pline loop of th_iter (form 0, mod_th.f90:57-64) or trans_calc (form 1,
    // ppow = xppow, mod_trans.f90:430-441); node_nf(nxx, nyy) = fuel pins per node (mod_io.f90:3064-3082)
    if (!c || !node_nf) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->th_set && c->have_flux, "adp_th_pline: needs adp_set_th and a flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    std::vector<double> nf(c->np);
    for (int r = 0; r < c->np; ++r) nf[r] = node_nf[(size_t)(c->h_iy[r] - 1) * c->nxx + (c->h_ix[r] - 1)];
    CUDA_TRY(c, cudaMemcpy(c->d_nodenf, nf.data(), (size_t)c->np * sizeof(double), cudaMemcpyHostToDevice));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_powdis(c, c->d_stage));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (c->h_scal[S_POW] <= 0.0) { c->err = "ERROR: TOTAL NODES POWER IS ZERO OR LESS"; return ADP_STOP_ZERO_POWER; }
    TRY(adp_k_scale_by_slot(c, c->d_stage, S_POW));
    k_th_pline<<<(c->np + ADP_TILE - 1) / ADP_TILE, ADP_TILE, 0, c->stream>>>(c->geo, c->d_stage, pw, ppow, form, c->d_nodenf,
                                                                            c->d_pline);
    c->launches++;
    CUDA_TRY(c, cudaPeekAtLastError());
    c->th_pline_set = true;
    return ADP_OK;
}

static int th_march(adp_ctx *c, const double *xpline, bool trans, double h, double *th_err)
{
    ADP_REQUIRE(c, c->th_set && c->th_state_set, "thermal-hydraulic update: needs adp_set_th and adp_set_th_state");
    ADP_REQUIRE(c, xpline || c->th_pline_set, "thermal-hydraulic update: no linear power density (pass it or call adp_th_pline)");
    ADP_REQUIRE(c, !trans || h > 0.0, "adp_th_trans: time step must be positive");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (xpline) { TRY(adp_upload_nodes(c, c->d_pline, xpline, 1)); c->th_pline_set = true; }
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    if (th_err)   // otem = ftem
        CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, c->d_ftem, (size_t)c->NV * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    ThArgs A{};
    A.P = c->th; A.stab = c->d_stab; A.xpline = c->d_pline; A.tfm = c->d_tfm; A.heatf = c->d_heatf; A.ent = c->d_ent;
    A.ftem = c->d_ftem; A.mtem = c->d_mtem; A.cden = c->d_cden; A.frate = c->d_frate; A.chain = c->d_chain; A.h = h;
    A.errflag = c->d_errflag;
    TRY(adp_comm_chain_recv(c, c->d_chain, 2 * c->np));
    const int grid = (c->np + ADP_TILE - 1) / ADP_TILE;
    if (trans) k_th_march<true><<<grid, ADP_TILE, 0, c->stream>>>(c->geo, A);
    else k_th_march<false><<<grid, ADP_TILE, 0, c->stream>>>(c->geo, A);
    c->launches++;
    CUDA_TRY(c, cudaPeekAtLastError());
    TRY(adp_comm_chain_send(c, c->d_chain, 2 * c->np));
    if (th_err) {
        k_th_abse<<<1, ADP_TILE, 0, c->stream>>>(c->geo, c->d_ftem, c->d_stage, c->d_scal + S_TMP0);
        c->launches++;
        TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_TMP0, 1));
    }
    if (c->nranks > 1) {      // any rank's steam-table violation stops all of them
        CUDA_TRY(c, cudaMemcpyAsync(c->h_flags, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        double f = (double)c->h_flags[0];
        CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP1, &f, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_TMP1, 1));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        ADP_CHECK_FAULT(c);
        c->h_flags[0] = (int)c->h_scal[S_TMP1];
    } else {
        CUDA_TRY(c, cudaMemcpyAsync(c->h_flags, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        ADP_CHECK_FAULT(c);
    }
    if (th_err) *th_err = c->h_scal[S_TMP0];
    if (c->h_flags[0] == ADP_STOP_STEAM_TABLE) {
        c->err = "ERROR: ENTHALPY IS OUT OF THE RANGE IN THE STEAM TABLE. CHECK INPUT COOLANT MASS FLOW RATE OR CORE POWER";
        return ADP_STOP_STEAM_TABLE;
    }
    return ADP_OK;
}

extern "C" int adp_th_upd(adp_ctx *c, const double *xpline, double *th_err)
{   // CALL th_upd(pline) (mod_th.f90:594-699) [+ otem = ftem ... CALL AbsE(ftem, otem, th_err), :43,70]
    if (!c) return ADP_ERR_USAGE;
    return th_march(c, xpline, false, 0.0, th_err);
}

extern "C" int adp_th_trans(adp_ctx *c, const double *xpline, double h)
{   // CALL th_trans(pline, h) (mod_th.f90:440-591)
    if (!c) return ADP_ERR_USAGE;
    return th_march(c, xpline, true, h, nullptr);
}
