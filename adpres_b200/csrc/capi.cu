// capi.cu -- implementation of the C ABI declared in include/adpres_b200.h
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "adp_internal.cuh"
#include "xtab_node.cuh"

void adp_k_preload_cmfd(adp_ctx *c);
void adp_k_preload_nodal(adp_ctx *c);

static thread_local std::string g_create_err;

extern "C" const char *adp_version(void) { return "adpres_b200 0.1 (sm_100a)"; }

extern "C" const char *adp_last_error(const adp_ctx *c) { return c ? c->err.c_str() : g_create_err.c_str(); }

template <typename T>
static int dev_alloc(adp_ctx *c, T **p, size_t n, bool zero = true)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) n = 1;
    CUDA_TRY(c, cudaMalloc((void **)p, n * sizeof(T)));
    if (zero) CUDA_TRY(c, cudaMemsetAsync(*p, 0, n * sizeof(T), c->stream));
    return ADP_OK;
}
#define TRY(x)              \
    do {                    \
        int rc__ = (x);     \
        if (rc__) return rc__; \
    } while (0)

extern "C" int adp_create(adp_ctx **out, int device)
{
    if (!out) return ADP_ERR_USAGE;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("adp_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return ADP_ERR_CUDA;
    }
    if (device < 0) {   // -1: take the device from the launcher's environment (one process per GPU)
        const char *v = getenv("ADP_LOCAL_RANK");
        if (!v) v = getenv("LOCAL_RANK");
        device = v ? atoi(v) % ndev : 0;
    }
    if (device >= ndev) { g_create_err = "adp_create: bad device index"; return ADP_ERR_USAGE; }
    adp_ctx *c = new adp_ctx();
    c->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_err = std::string("adp_create: ") + cudaGetErrorString(e);
        delete c;
        return ADP_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    c->sm_count = prop.multiProcessorCount;
    if (const char *v = getenv("ADP_MAIL_TIMEOUT_S")) { const double t = atof(v); if (t > 0.0) c->mail_timeout_s = t; }
    if (getenv("ADP_NO_FUSE_MAIL")) c->fuse_mail = false;
    if (getenv("ADP_NO_MAIL_LL")) c->mail_ll = false;
    // persistent grids: a multiple of the SM count (148 on B200) x resident CTAs per SM
    c->grid_blocks = 0;
    if (dev_alloc(c, &c->d_scal, S_COUNT) || dev_alloc(c, &c->d_part, 4 * ADP_MAXPART) || dev_alloc(c, &c->d_ticket, 1) ||
        dev_alloc(c, &c->d_argidx, 1) || dev_alloc(c, &c->d_errflag, 1) ||
        cudaMallocHost((void **)&c->h_scal, (S_COUNT + 8) * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void **)&c->h_flags, 8 * sizeof(long long)) != cudaSuccess) {
        g_create_err = "adp_create: allocation failed: " + c->err;
        delete c;
        return ADP_ERR_CUDA;
    }
    *out = c;
    return ADP_OK;
}

static void free_xtab(adp_ctx *c)
{
    if (c->d_brmeta) cudaFree(c->d_brmeta);
    if (c->d_brtoff) cudaFree(c->d_brtoff);
    if (c->d_brpar) cudaFree(c->d_brpar);
    if (c->d_brtab) cudaFree(c->d_brtab);
    if (c->d_brrtab) cudaFree(c->d_brrtab);
    c->d_brmeta = nullptr; c->d_brtoff = nullptr; c->d_brpar = c->d_brtab = c->d_brrtab = nullptr;
}

static void free_graphs(adp_ctx *c)
{
    for (auto &kv : c->graphs) cudaGraphExecDestroy(kv.second);
    c->graphs.clear();
}

extern "C" int adp_destroy(adp_ctx *c)
{
    if (!c) return ADP_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); c->stream2 = nullptr; }
    if (c->ev_lazy) { cudaEventDestroy(c->ev_lazy); c->ev_lazy = nullptr; }
    free_graphs(c);
    adp_comm_destroy(c);
    void *ptrs[] = {c->d_nodp, c->d_ypm, c->d_ypp, c->d_ixr, c->d_iyr, c->d_mat, c->d_flag, c->d_hx, c->d_hy, c->d_hz, c->d_area,
                    c->d_f0[0], c->d_f0[1], c->d_fs[0], c->d_fs[1], c->d_r, c->d_rs, c->d_p, c->d_v, c->d_v2, c->d_s, c->d_t,
                    c->d_s0, c->d_a, c->d_df, c->d_dn, c->d_D, c->d_sigr, c->d_nuf, c->d_sigf, c->d_exsrc, c->d_sigs,
                    c->d_dc, c->d_chi, c->d_S, c->d_c0, c->d_ft, c->d_fst, c->d_omeg, c->d_sigrp, c->d_L, c->d_dfis,
                    c->d_tbeta, c->d_velo, c->d_af, c->d_xtab, c->d_dtab, c->d_fb, c->d_bpos, c->d_dumtop, c->d_nd, c->d_abefgh, c->d_mail, c->d_arseq, c->d_mail_table, c->d_scal, c->d_part, c->d_ticket, c->d_argidx, c->d_errflag, c->d_stage, c->d_gather};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->d_res) cudaFree(c->d_res);
    {
        double *th[] = {c->d_stab, c->d_tfm, c->d_heatf, c->d_ent, c->d_ftem, c->d_mtem, c->d_cden, c->d_frate, c->d_pline, c->d_nodenf, c->d_chain};
        for (double *q : th) if (q) cudaFree(q);
        for (int f = 0; f < 4; ++f) if (c->d_ftab[f]) cudaFree(c->d_ftab[f]);
        void *br[] = {c->d_brmeta, c->d_brtoff, c->d_brpar, c->d_brtab, c->d_brrtab, c->d_mkin};
        for (void *q : br) if (q) cudaFree(q);
    }
    cudaStreamDestroy(c->stream);
    delete c;
    return ADP_OK;
}

extern "C" int adp_slab(const adp_ctx *c, int *k0, int *k1)
{
    if (!c || !c->geometry_set) return ADP_ERR_USAGE;
    if (k0) *k0 = c->k0;
    if (k1) *k1 = c->k1;
    return ADP_OK;
}

// ---------------------------------------------------------------------------------------------
// host <-> device transfer of node arrays.  Host: Fortran (nnod, ncol) column-major, global.
// Device: [col][NV] with ghost planes.  `ghost` selects whether the ghost planes inside the
// core are filled from the (global) host array as well -- static data needs no communication
// because every rank holds the global arrays.
// ---------------------------------------------------------------------------------------------
static int upload_nodes(adp_ctx *c, double *d, const double *h, int ncol, bool ghost)
{
    const int np = c->np;
    const int ka = ghost ? std::max(0, c->k0 - ADP_GH) : c->k0;
    const int kb = ghost ? std::min(c->nzz, c->k1 + ADP_GH) : c->k1;
    const size_t cnt = (size_t)(kb - ka) * np;
    for (int col = 0; col < ncol; ++col) {
        const double *src = h + (size_t)col * c->nnod + (size_t)ka * np;
        double *dst = d + (size_t)col * c->NV + (size_t)(ka - (c->k0 - ADP_GH)) * np;
        CUDA_TRY(c, cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    return ADP_OK;
}
static int upload_nodes_int(adp_ctx *c, int *d, const int *h)
{
    const int np = c->np;
    const int ka = std::max(0, c->k0 - ADP_GH), kb = std::min(c->nzz, c->k1 + ADP_GH);
    CUDA_TRY(c, cudaMemcpyAsync(d + (size_t)(ka - (c->k0 - ADP_GH)) * np, h + (size_t)ka * np,
                                (size_t)(kb - ka) * np * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    return ADP_OK;
}
// own planes of this rank -> the matching part of the global host array
static int download_nodes(adp_ctx *c, double *h, const double *d, int ncol)
{
    const size_t cnt = (size_t)c->nzl * c->np;
    for (int col = 0; col < ncol; ++col) {
        const double *own = d + (size_t)col * c->NV + (size_t)ADP_GH * c->np;
        CUDA_TRY(c, cudaMemcpyAsync(h + (size_t)col * c->nnod + (size_t)c->k0 * c->np, own, cnt * sizeof(double),
                                    cudaMemcpyDeviceToHost, c->stream));
        // several ranks: the rows of the other slabs too (option "gather_results", default on), so that the host
        // code that consumes whole sdata arrays stays correct on every rank
        if (c->nranks > 1 && c->gather_results) TRY(adp_comm_gather_column(c, h + (size_t)col * c->nnod, own));
    }
    return ADP_OK;
}

// the same transfers for the other translation units (results.cu, th.cu)
int adp_upload_nodes(adp_ctx *c, double *d, const double *h, int ncol) { return upload_nodes(c, d, h, ncol, false); }
int adp_download_nodes(adp_ctx *c, double *h, const double *d, int ncol) { return download_nodes(c, h, d, ncol); }

// ---------------------------------------------------------------------------------------------
extern "C" int adp_set_geometry(adp_ctx *c, int nxx, int nyy, int nzz, int nnod, int ng, int nmat, const int *ix,
                                const int *iy, const int *iz, const int *ysmin, const int *ysmax, const int *xsmin,
                                const int *xsmax, const double *xdel, const double *ydel, const double *zdel,
                                const int bc[6], const int *mat)
{
    if (!c) return ADP_ERR_USAGE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    adp_comm_unmap_peers(c);
    // a deferred upload (option "lazy_adf") targets buffers this call re-allocates
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    c->lazy_dc = c->lazy_sigf = nullptr; c->lazy_pending = false;
    ADP_REQUIRE(c, ng >= 1 && ng <= ADP_MAXG, "adp_set_geometry: ng must be 1..16");
    ADP_REQUIRE(c, nnod > 0 && nzz > 0 && nnod % nzz == 0, "adp_set_geometry: nnod must be np*nzz (plane-invariant core outline)");
    free_graphs(c);
    c->nxx = nxx; c->nyy = nyy; c->nzz = nzz; c->nnod = nnod; c->ng = ng; c->nmat = nmat;
    const int np = nnod / nzz;
    c->np = np;
    for (int i = 0; i < 6; ++i) c->bc[i] = bc[i];
    c->h_ix.assign(ix, ix + nnod); c->h_iy.assign(iy, iy + nnod); c->h_iz.assign(iz, iz + nnod);
    // numbering must be the reference's: k-major, then j, then i = smin..smax (mod_io.f90:1319-1330)
    {
        int n = 0;
        bool okn = true;
        for (int j = 1; j <= nyy && okn; ++j)
            for (int i = ysmin[j - 1]; i <= ysmax[j - 1]; ++i, ++n)
                if (n >= np || ix[n] != i || iy[n] != j || iz[n] != 1) { okn = false; break; }
        ADP_REQUIRE(c, okn && n == np, "adp_set_geometry: node numbering is not the reference's k,j,i order");
        ADP_REQUIRE(c, iz[nnod - 1] == nzz && ix[nnod - 1] == ix[np - 1], "adp_set_geometry: planes differ");
    }
    // z-slab of this rank
    {
        const int base = nzz / c->nranks, rem = nzz % c->nranks;
        c->k0 = c->rank * base + std::min(c->rank, rem);
        c->k1 = c->k0 + base + (c->rank < rem ? 1 : 0);
        c->nzl = c->k1 - c->k0;
        ADP_REQUIRE(c, c->nzl >= ((c->nranks > 1) ? 2 : 1), "adp_set_geometry: fewer than 2 planes per rank");
    }
    c->NL = (long long)np * c->nzl;
    c->NV = (long long)np * (c->nzl + 2 * ADP_GH);
    // plane tables
    std::vector<int> nodp((size_t)nxx * nyy, 0), ypm(np), ypp(np), ixr(np), iyr(np);
    std::vector<unsigned char> flag(np);
    std::vector<double> hx(np), hy(np), area(np), hz(nzz + 2, 0.0);
    for (int r = 0; r < np; ++r) nodp[(size_t)(iy[r] - 1) * nxx + (ix[r] - 1)] = r + 1;
    c->h_nodp = nodp;
    c->h_xdel.assign(xdel, xdel + nxx); c->h_ydel.assign(ydel, ydel + nyy); c->h_zdel.assign(zdel, zdel + nzz);
    for (int r = 0; r < np; ++r) {
        const int i = ix[r], j = iy[r];
        unsigned f = 0;
        if (i == ysmin[j - 1]) f |= FLAG_XM;
        if (i == ysmax[j - 1]) f |= FLAG_XP;
        if (j == xsmin[i - 1]) f |= FLAG_YM;
        if (j == xsmax[i - 1]) f |= FLAG_YP;
        flag[r] = (unsigned char)f;
        // set_ind: n -+ (nodp(i,j) - nodp(i,j-+1))   (mod_cmfd.f90:180,202)
        ypm[r] = (f & FLAG_YM) ? 0 : (r + 1) - nodp[(size_t)(j - 2) * nxx + (i - 1)];
        ypp[r] = (f & FLAG_YP) ? 0 : nodp[(size_t)j * nxx + (i - 1)] - (r + 1);
        ADP_REQUIRE(c, ypm[r] >= 0 && ypp[r] >= 0 && ((f & FLAG_YM) || ypm[r] > 0) && ((f & FLAG_YP) || ypp[r] > 0),
                    "adp_set_geometry: inconsistent staggering (ystag/xstag)");
        ixr[r] = i; iyr[r] = j;
        hx[r] = xdel[i - 1]; hy[r] = ydel[j - 1];
        area[r] = xdel[i - 1] * ydel[j - 1];
    }
    for (int k = 0; k < nzz; ++k) hz[1 + k] = zdel[k];
    hz[0] = zdel[0]; hz[nzz + 1] = zdel[nzz - 1];
    TRY(dev_alloc(c, &c->d_ypm, np)); TRY(dev_alloc(c, &c->d_ypp, np)); TRY(dev_alloc(c, &c->d_ixr, np));
    TRY(dev_alloc(c, &c->d_iyr, np)); TRY(dev_alloc(c, &c->d_flag, np)); TRY(dev_alloc(c, &c->d_hx, np + 2));
    TRY(dev_alloc(c, &c->d_hy, np)); TRY(dev_alloc(c, &c->d_area, np)); TRY(dev_alloc(c, &c->d_hz, nzz + 2));
    TRY(dev_alloc(c, &c->d_nodp, (size_t)nxx * nyy));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_nodp, nodp.data(), (size_t)nxx * nyy * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ypm, ypm.data(), np * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ypp, ypp.data(), np * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_ixr, ixr.data(), np * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_iyr, iyr.data(), np * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_flag, flag.data(), np, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_hx + 1, hx.data(), np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_hy, hy.data(), np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_area, area.data(), np * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_hz, hz.data(), (nzz + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // the std::vectors die at return

    Geo &G = c->geo;
    G.np = np; G.nzl = c->nzl; G.nzz = nzz; G.k0 = c->k0;
    G.tpp = (np + ADP_TILE - 1) / ADP_TILE;
    G.ntiles = G.tpp * c->nzl;
    G.NV = c->NV;
    for (int i = 0; i < 6; ++i) G.bc[i] = bc[i];
    G.ypm = c->d_ypm; G.ypp = c->d_ypp; G.flag = c->d_flag; G.hx = c->d_hx + 1; G.hy = c->d_hy; G.hz = c->d_hz;
    G.area = c->d_area; G.ixr = c->d_ixr; G.iyr = c->d_iyr;
    c->geoxy.nxx = nxx; c->geoxy.nyy = nyy; c->geoxy.nodp = c->d_nodp;

    const size_t NV = (size_t)c->NV, Gn = (size_t)ng;
    TRY(dev_alloc(c, &c->d_mat, NV));
    {   // ghost planes outside the core keep material 1 so that table look-ups stay in range
        std::vector<int> ones(NV, 1);
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // dev_alloc's zero fill runs on c->stream, the blocking copy does not
        CUDA_TRY(c, cudaMemcpy(c->d_mat, ones.data(), NV * sizeof(int), cudaMemcpyHostToDevice));
    }
    TRY(upload_nodes_int(c, c->d_mat, mat));
    for (int w = 0; w < 2; ++w) { TRY(dev_alloc(c, &c->d_f0[w], Gn * NV)); TRY(dev_alloc(c, &c->d_fs[w], NV)); }
    TRY(dev_alloc(c, &c->d_r, NV)); TRY(dev_alloc(c, &c->d_rs, NV)); TRY(dev_alloc(c, &c->d_p, NV));
    TRY(dev_alloc(c, &c->d_v, NV)); TRY(dev_alloc(c, &c->d_s, NV)); TRY(dev_alloc(c, &c->d_t, NV));
    if (c->nranks > 1) TRY(dev_alloc(c, &c->d_v2, NV));
    TRY(dev_alloc(c, &c->d_s0, NV));
    TRY(dev_alloc(c, &c->d_a, Gn * 7 * NV));
    TRY(dev_alloc(c, &c->d_df, Gn * 6 * NV)); TRY(dev_alloc(c, &c->d_dn, Gn * 6 * NV));
    TRY(dev_alloc(c, &c->d_D, Gn * NV)); TRY(dev_alloc(c, &c->d_sigr, Gn * NV)); TRY(dev_alloc(c, &c->d_nuf, Gn * NV));
    TRY(dev_alloc(c, &c->d_sigf, Gn * NV)); TRY(dev_alloc(c, &c->d_exsrc, Gn * NV));
    TRY(dev_alloc(c, &c->d_sigs, Gn * Gn * NV)); TRY(dev_alloc(c, &c->d_dc, 6 * Gn * NV));
    TRY(dev_alloc(c, &c->d_chi, Gn * nmat)); TRY(dev_alloc(c, &c->d_S, 3 * Gn * NV));
    TRY(dev_alloc(c, &c->d_tbeta, nmat)); TRY(dev_alloc(c, &c->d_dfis, NV)); TRY(dev_alloc(c, &c->d_velo, ng));
    TRY(dev_alloc(c, &c->d_stage, NV));
    // ghost planes outside the core stay zero (dev_alloc clears): coup_coef never divides by their D
    // (boundary branches, mod_cmfd.f90:45-130) and their matrix coefficients are zero
    // buffers allocated on first use are sized by the geometry: drop them, they come back on demand
    {
        double **lazy[] = {&c->d_nd, &c->d_abefgh, &c->d_c0, &c->d_ft, &c->d_fst, &c->d_omeg, &c->d_sigrp, &c->d_L,
                           &c->d_af, &c->d_xtab, &c->d_dtab, &c->d_bpos, &c->d_dumtop, &c->d_res, &c->d_stab, &c->d_tfm,
                           &c->d_heatf, &c->d_ent, &c->d_ftem, &c->d_mtem, &c->d_cden, &c->d_frate, &c->d_pline, &c->d_nodenf,
                           &c->d_chain, &c->d_gather};
        for (double **q : lazy)
            if (*q) { cudaFree(*q); *q = nullptr; }
        if (c->d_fb) { cudaFree(c->d_fb); c->d_fb = nullptr; }
        if (c->h_res) { cudaFreeHost(c->h_res); c->h_res = nullptr; }
        c->res_elems = 0; c->nb = 0; c->kinetics_set = false;
        c->th_set = c->th_state_set = c->th_pline_set = false;
        for (int f = 0; f < 4; ++f)
            if (c->d_ftab[f]) { cudaFree(c->d_ftab[f]); c->d_ftab[f] = nullptr; }
        free_xtab(c);
        if (c->d_mkin) { cudaFree(c->d_mkin); c->d_mkin = nullptr; }
        c->kin_xtab = false;
    }
    c->abefgh_valid = false;
    c->geometry_set = true;
    c->xs_set = false; c->matrix_ready = false; c->have_flux = false; c->coup_first = true;
    c->outer_first = c->outer_ad_first = true;
    c->ndmax = 0.0; c->s0_group = 0;
    for (int g = 0; g < ADP_MAXG; ++g) c->cur[g] = 0;
    c->fcur = 0;
    memset(c->xghost_valid, 0, sizeof(c->xghost_valid));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    TRY(adp_comm_map_peers(c));
    return ADP_OK;
}

static int ensure_transient(adp_ctx *c)
{
    if (c->d_c0) return ADP_OK;
    const size_t NV = (size_t)c->NV, Gn = (size_t)c->ng;
    TRY(dev_alloc(c, &c->d_c0, ADP_NF * NV)); TRY(dev_alloc(c, &c->d_ft, Gn * NV)); TRY(dev_alloc(c, &c->d_fst, NV));
    TRY(dev_alloc(c, &c->d_omeg, Gn * NV)); TRY(dev_alloc(c, &c->d_sigrp, Gn * NV)); TRY(dev_alloc(c, &c->d_L, Gn * NV));
    return ADP_OK;
}

// ---- option "lazy_adf": deferred upload of dc and sigf (see adp_ctx) ------------------------------------------------
static int upload_nodes_on(adp_ctx *c, cudaStream_t st, double *d, const double *h, int ncol)
{
    const int np = c->np;
    const int ka = std::max(0, c->k0 - ADP_GH), kb = std::min(c->nzz, c->k1 + ADP_GH);
    const size_t cnt = (size_t)(kb - ka) * np;
    for (int col = 0; col < ncol; ++col)
        CUDA_TRY(c, cudaMemcpyAsync(d + (size_t)col * c->NV + (size_t)(ka - (c->k0 - ADP_GH)) * np,
                                    h + (size_t)col * c->nnod + (size_t)ka * np, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
    return ADP_OK;
}
int adp_lazy_enqueue(adp_ctx *c)
{
    if (!c->lazy_dc && !c->lazy_sigf) return ADP_OK;
    if (!c->stream2) CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    if (!c->ev_lazy) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_lazy, cudaEventDisableTiming));
    if (c->lazy_dc) TRY(upload_nodes_on(c, c->stream2, c->d_dc, c->lazy_dc, c->ng * 6));
    if (c->lazy_sigf) TRY(upload_nodes_on(c, c->stream2, c->d_sigf, c->lazy_sigf, c->ng));
    CUDA_TRY(c, cudaEventRecord(c->ev_lazy, c->stream2));
    c->lazy_dc = c->lazy_sigf = nullptr;
    c->lazy_pending = true;
    return ADP_OK;
}
int adp_lazy_sync(adp_ctx *c)
{
    TRY(adp_lazy_enqueue(c));            // uploads that were never started (no adp_outer_begin in between)
    if (c->lazy_pending) {
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_lazy, 0));
        c->lazy_pending = false;
    }
    return ADP_OK;
}

extern "C" int adp_set_xs(adp_ctx *c, const double *D, const double *sigr, const double *nuf, const double *sigf,
                          const double *sigs, const double *chi, const double *dc, const double *exsrc)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_xs: call adp_set_geometry first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int G = c->ng;
    if (D) TRY(upload_nodes(c, c->d_D, D, G, true));
    if (sigr) TRY(upload_nodes(c, c->d_sigr, sigr, G, true));
    if (D || sigr) c->abefgh_valid = false;   // the SANM constants A..H depend on sigr/D only
    if (nuf) TRY(upload_nodes(c, c->d_nuf, nuf, G, true));
    if (c->lazy_adf) {
        // a deferred upload still in flight must not be overtaken by this one; the new arrays go up from adp_outer_begin
        TRY(adp_lazy_sync(c));
        if (sigf) c->lazy_sigf = sigf;
        if (dc) c->lazy_dc = dc;
    } else {
        if (sigf) TRY(upload_nodes(c, c->d_sigf, sigf, G, true));
        if (dc) TRY(upload_nodes(c, c->d_dc, dc, G * 6, true));      // host dc(n,g,f): column g + G*f -> device [f][g]
    }
    if (sigs) TRY(upload_nodes(c, c->d_sigs, sigs, G * G, true));   // host (n,g,h): column g + G*h = device [h][g]
    if (exsrc) TRY(upload_nodes(c, c->d_exsrc, exsrc, G, true));
    if (chi) {
        // host chi(nmat, ng) column-major = [g][nmat]
        CUDA_TRY(c, cudaMemcpyAsync(c->d_chi, chi, (size_t)G * c->nmat * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (D && sigr && nuf && sigf && sigs && chi && dc && exsrc) c->xs_set = true;
    return ADP_OK;
}

extern "C" int adp_set_xs_mask(adp_ctx *c, int mask, const double *D, const double *sigr, const double *nuf, const double *sigf,
                               const double *sigs, const double *chi, const double *dc, const double *exsrc)
{
    auto sel = [mask](int bit, const double *p) { return (mask >> bit) & 1 ? p : nullptr; };
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->xs_set || (mask & 0xff) == 0xff, "adp_set_xs_mask: the first upload must carry every array");
    return adp_set_xs(c, sel(0, D), sel(1, sigr), sel(2, nuf), sel(3, sigf), sel(4, sigs), sel(5, chi), sel(6, dc), sel(7, exsrc));
}

extern "C" int adp_set_control(adp_ctx *c, int nout, int nin, int nac, int nupd, double serc, double ferc, int kern)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, nin >= 0 && nac >= 1 && nupd >= 1 && nout >= 1, "adp_set_control: bad iteration control");
    ADP_REQUIRE(c, kern >= ADP_KERN_FDM && kern <= ADP_KERN_SANM, "adp_set_control: kern must be 0 FDM, 1 PNM, 2 SANM");
    if (nin != c->nin) free_graphs(c);
    c->nout = nout; c->nin = nin; c->nac = nac; c->nupd = nupd; c->serc = serc; c->ferc = ferc; c->kern = kern;
    if (c->geometry_set) {
        cudaSetDevice(c->device);
        adp_k_preload_cmfd(c);
        if (kern != ADP_KERN_FDM) adp_k_preload_nodal(c);
    }
    return ADP_OK;
}

extern "C" int adp_matrix_setup(adp_ctx *c, int opt)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->xs_set, "adp_matrix_setup: cross sections not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (opt > 0) {
        if (c->coup_first) {   // nod%dn = 0 on the first call (mod_cmfd.f90:25-33)
            CUDA_TRY(c, cudaMemsetAsync(c->d_dn, 0, (size_t)c->ng * 6 * c->NV * sizeof(double), c->stream));
            c->coup_first = false;
        }
        TRY(adp_k_coup_coef(c));
    }
    TRY(adp_k_matrix_setup(c));
    c->matrix_ready = true;
    return ADP_OK;
}

extern "C" int adp_init_flux(adp_ctx *c, int adjoint)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->xs_set, "adp_init_flux: cross sections not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return adp_k_init_flux(c, adjoint);
}

extern "C" int adp_outer_begin(adp_ctx *c, int mode)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_outer_begin: needs adp_matrix_setup and a flux (adp_init_flux/adp_set_state)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_lazy_enqueue(c));            // option "lazy_adf": dc and sigf travel while the outer iterations run
    return adp_k_outer_begin(c, mode);
}

// ---- one outer iteration -----------------------------------------------------------------
static int issue_outer_iter(adp_ctx *c, int mode, bool extrap, bool readback = true)
{
    const int G = c->ng;
    if (mode == ADP_MODE_ADJOINT) {
        for (int g = G - 1; g >= 0; --g) TRY(adp_k_bicg_group(c, mode, g, c->nin, g == 0));
    } else {
        for (int g = 0; g < G; ++g) TRY(adp_k_bicg_group(c, mode, g, c->nin, g == G - 1));
    }
    TRY(adp_k_outer_tail(c, mode, extrap));
    if (readback) CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return ADP_OK;
}

// Launch one outer iteration: CUDA-graph replay when possible, direct launches otherwise.
static int launch_outer_iter(adp_ctx *c, int mode, bool extrap)
{
    // multi-rank: graphs only once no NCCL call is left inside the iteration (halos pushed by the
    // kernels, reductions through the mailboxes, ghost flux planes current)
    bool use_graph = c->use_graphs;
    if (c->nranks > 1) {
        use_graph = use_graph && c->peer_ok && c->peer_ar;
        for (int g = 0; g < c->ng && use_graph; ++g) use_graph = c->xghost_valid[c->cur[g]][g];
    }
    if (!use_graph) return issue_outer_iter(c, mode, extrap);
    unsigned long long key = (unsigned long long)mode | ((unsigned long long)(extrap ? 1 : 0) << 4) |
                             ((unsigned long long)c->fcur << 5);
    for (int g = 0; g < c->ng; ++g) key |= (unsigned long long)c->cur[g] << (8 + g);
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        cudaGraph_t graph = nullptr;
        const long long l0 = c->launches;
        const int s0g = c->s0_group, fcur = c->fcur;
        int cur[ADP_MAXG];
        memcpy(cur, c->cur, sizeof(cur));
        CUDA_TRY(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = issue_outer_iter(c, mode, extrap);
        cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
        // capture advanced the host-side bookkeeping once although nothing ran; undo it on every path
        // (success: the launch below redoes it; failure: the device buffers have not moved)
        const long long captured = c->launches - l0;
        c->launches = l0; c->s0_group = s0g; c->fcur = fcur;
        memcpy(c->cur, cur, sizeof(cur));
        if (rc || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (rc) return rc;
            CUDA_TRY(c, e);
        }
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CUDA_TRY(c, e);
        c->graph_launches[key] = captured;
        it = c->graphs.emplace(key, exec).first;
    }
    CUDA_TRY(c, cudaGraphLaunch(it->second, c->stream));
    // host bookkeeping of what the graph did
    for (int g = 0; g < c->ng; ++g) c->cur[g] ^= 1;
    c->fcur ^= 1;
    c->s0_group = (mode == ADP_MODE_ADJOINT) ? 1 : c->ng;
    if (c->nranks > 1) for (int g = 0; g < c->ng; ++g) c->xghost_valid[c->cur[g]][g] = true;
    c->launches += c->graph_launches[key];
    return ADP_OK;
}

extern "C" int adp_outer_iter(adp_ctx *c, int mode, int p, double *Ke, double *ser, double *fer)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_outer_iter: needs adp_matrix_setup and a flux");
    ADP_REQUIRE(c, mode >= ADP_MODE_FORWARD && mode <= ADP_MODE_TRANSIENT, "adp_outer_iter: bad mode");
    ADP_REQUIRE(c, mode != ADP_MODE_TRANSIENT || c->kinetics_set, "adp_outer_iter: transient mode needs adp_set_kinetics");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(launch_outer_iter(c, mode, (p % c->nac) == 0));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (Ke) *Ke = c->h_scal[S_KE];
    c->last_ser = c->h_scal[S_SER]; c->last_fer = c->h_scal[S_FER];
    if (ser) *ser = c->last_ser;
    if (fer) *fer = c->last_fer;
    if (c->nranks > 1 && !std::isfinite(c->h_scal[S_KE])) {
        int flag = 0;
        cudaMemcpy(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost);
        char buf[512];
        snprintf(buf, sizeof(buf), "adp_outer_iter: non-finite k-eff at p=%d rank %d (errflag %d%s): Ke %g F %g FC %g FINT %g E2SQ %g RSV %g TT %g TS %g RHO %g %g SER %g FER %g",
                 p, c->rank, flag, flag == ADP_ERR_NCCL ? " = peer all-reduce timeout" : "", c->h_scal[S_KE], c->h_scal[S_F], c->h_scal[S_FC],
                 c->h_scal[S_FINT], c->h_scal[S_E2SQ], c->h_scal[S_RSV], c->h_scal[S_TT], c->h_scal[S_TS], c->h_scal[S_RHO0], c->h_scal[S_RHO1],
                 c->h_scal[S_SER], c->h_scal[S_FER]);
        c->err = buf;
        return ADP_ERR_NCCL;
    }
    return ADP_OK;
}

extern "C" int adp_nodal_upd(adp_ctx *c, int nmode, double *ndmax, int *im, int *jm, int *km)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_nodal_upd: needs adp_matrix_setup and a flux");
    ADP_REQUIRE(c, c->kern != ADP_KERN_FDM, "adp_nodal_upd: kern is FDM");
    ADP_REQUIRE(c, nmode != 2 || c->kinetics_set, "adp_nodal_upd: cmode 2 needs adp_set_kinetics");
    CUDA_TRY(c, cudaSetDevice(c->device));
    // ndmax = 0 (mod_cmfd.f90:358); the location keeps its previous value if nothing changes
    CUDA_TRY(c, cudaMemsetAsync(c->d_scal + S_NDMAX, 0, sizeof(double), c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    TRY(adp_k_nodal_update(c, nmode));
    if (c->nranks > 1) {
        // global maximum and, among the ranks holding it, the lowest node number
        CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP0, c->d_scal + S_NDMAX, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_NDMAX, 1));
        CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        ADP_CHECK_FAULT(c);
        if (c->h_scal[S_TMP0] != c->h_scal[S_NDMAX]) {
            long long big = 0x7fffffffffffffffLL;
            CUDA_TRY(c, cudaMemcpyAsync(c->d_argidx, &big, sizeof(big), cudaMemcpyHostToDevice, c->stream));
        }
        TRY(adp_comm_allreduce_min_ll(c, c->d_argidx, 1));
        int flag = 0;
        CUDA_TRY(c, cudaMemcpyAsync(&flag, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        double f = flag;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP1, &f, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_TMP1, 1));
        CUDA_TRY(c, cudaMemcpyAsync(&f, c->d_scal + S_TMP1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        flag = (int)f;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_errflag, &flag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    TRY(adp_k_matrix_setup(c));   // matrix_setup(0), mod_cmfd.f90:366
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_flags, c->d_argidx, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_flags + 2, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    c->ndmax = c->h_scal[S_NDMAX];
    if (c->ndmax > 0.0) {
        const long long loc = *(long long *)c->h_flags;
        if (loc >= 0 && loc < c->nnod) { c->im = c->h_ix[loc]; c->jm = c->h_iy[loc]; c->km = c->h_iz[loc]; }
    }
    if (ndmax) *ndmax = c->ndmax;
    if (im) *im = c->im;
    if (jm) *jm = c->jm;
    if (km) *km = c->km;
    if (c->h_flags[2] != 0) { c->err = "ERROR IN MATRIX DECOMP: DIAGONAL ELEMENTS CLOSE TO ZERO"; return ADP_STOP_LU_DIAG; }
    if (c->ndmax > 1.e3) { c->err = "Max. change in nodal coupling coefficient > 1e3: the two-node nonlinear iteration seems not stable"; return ADP_STOP_NDMAX; }
    return ADP_OK;
}

extern "C" int adp_get_ndmax(adp_ctx *c, double *ndmax)
{
    if (!c || !ndmax) return ADP_ERR_USAGE;
    *ndmax = c->ndmax;
    return ADP_OK;
}

extern "C" int adp_get_errors(adp_ctx *c, double *ser, double *fer)
{   // sdata's `ser`, `fer` after the last outer iteration (read by th_iter, cbsearch: mod_th.f90:77,789)
    if (!c) return ADP_ERR_USAGE;
    if (ser) *ser = c->last_ser;
    if (fer) *fer = c->last_fer;
    return ADP_OK;
}

extern "C" int adp_powdis(adp_ctx *c, double *p, int fixedsrc_mode)
{
    if (!c || !p) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_powdis: no flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_powdis(c, c->d_stage));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (c->h_scal[S_POW] <= 0.0 && !fixedsrc_mode) { c->err = "ERROR: TOTAL NODES POWER IS ZERO OR LESS"; return ADP_STOP_ZERO_POWER; }
    TRY(adp_k_scale_by_slot(c, c->d_stage, S_POW));
    TRY(download_nodes(c, p, c->d_stage, 1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_integrate(adp_ctx *c, const double *s, double *result)
{
    if (!c || !s || !result) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_integrate: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(upload_nodes(c, c->d_stage, s, 1, false));
    TRY(adp_k_integrate(c, c->d_stage, S_TMP0));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    *result = c->h_scal[S_TMP0];
    return ADP_OK;
}

// ---- transient ------------------------------------------------------------------------------
extern "C" int adp_set_kinetics(adp_ctx *c, const double *ibeta, const double *lamb, const double *velo,
                                const double *tbeta, double sth, double bth)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_kinetics: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    for (int i = 0; i < ADP_NF; ++i) { c->ibeta[i] = ibeta[i]; c->lamb[i] = lamb[i]; }
    c->sth = sth; c->bth = bth;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_velo, velo, c->ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tbeta, tbeta, c->nmat * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    TRY(ensure_transient(c));
    c->kinetics_set = true;
    c->kin_xtab = false;
    return ADP_OK;
}

extern "C" int adp_set_kinetics_xtab(adp_ctx *c, const double *mibeta, const double *mlamb, const double *mvelo,
                                     const double *tbeta, double sth, double bth)
{   // %XTAB decks (bxtab = 1): m(mat)%iBeta(6), %lamb(6), %velo(ng) per material; switches get_exsrc, iPden, uPden
    // and the time-absorption term of adp_begin_time_step to their bxtab = 1 branches
    if (!c || !mibeta || !mlamb || !mvelo || !tbeta) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_kinetics_xtab: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t nk = (size_t)ADP_NF * c->nmat, nv = (size_t)c->ng * c->nmat;
    TRY(dev_alloc(c, &c->d_mkin, 2 * nk + nv, false));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_mkin, mlamb, nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_mkin + nk, mibeta, nk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_mkin + 2 * nk, mvelo, nv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_tbeta, tbeta, c->nmat * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->sth = sth; c->bth = bth;
    TRY(ensure_transient(c));
    c->kinetics_set = true;
    c->kin_xtab = true;
    return ADP_OK;
}

extern "C" int adp_set_transient(adp_ctx *c, const double *c0, const double *ft, const double *fst, const double *omeg,
                                 const double *sigrp, const double *L)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_transient: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(ensure_transient(c));
    if (c0) TRY(upload_nodes(c, c->d_c0, c0, ADP_NF, false));
    if (ft) TRY(upload_nodes(c, c->d_ft, ft, c->ng, false));
    if (fst) TRY(upload_nodes(c, c->d_fst, fst, 1, false));
    if (omeg) TRY(upload_nodes(c, c->d_omeg, omeg, c->ng, false));
    if (sigrp) TRY(upload_nodes(c, c->d_sigrp, sigrp, c->ng, false));
    if (L) TRY(upload_nodes(c, c->d_L, L, c->ng, false));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_get_exsrc(adp_ctx *c, double ht)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->kinetics_set, "adp_get_exsrc: needs adp_set_kinetics");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return adp_k_get_exsrc(c, ht);
}

extern "C" int adp_get_exsrc_arrays(adp_ctx *c, double *exsrc, double *dfis)
{
    if (!c) return ADP_ERR_USAGE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (exsrc) TRY(download_nodes(c, exsrc, c->d_exsrc, c->ng));
    if (dfis) TRY(download_nodes(c, dfis, c->d_dfis, 1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

// ---- XS update on the device: %XSEC material tables + %CROD -------------
static int upload_tables(adp_ctx *c, double **dst, const double *a, const double *b, const double *d, const double *e,
                         const double *s5)
{
    const size_t mg = (size_t)c->nmat * c->ng, tot = 4 * mg + mg * c->ng;
    if (!*dst) TRY(dev_alloc(c, dst, tot));
    const double *src[4] = {a, b, d, e};
    for (int i = 0; i < 4; ++i)
        CUDA_TRY(c, cudaMemcpyAsync(*dst + i * mg, src[i], mg * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(*dst + 4 * mg, s5, mg * c->ng * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_set_material_xs(adp_ctx *c, const double *xsigtr, const double *xsiga, const double *xnuf,
                                   const double *xsigf, const double *xsigs)
{
    if (!c || !xsigtr || !xsiga || !xnuf || !xsigf || !xsigs) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_material_xs: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return upload_tables(c, &c->d_xtab, xsigtr, xsiga, xnuf, xsigf, xsigs);
}

// bank of every plane position, rod length above every plane, core height -- accumulated in the
// reference's order (mod_io.f90:1274-1277, mod_xsec.f90:252-277)
static int set_crod_geometry(adp_ctx *c, int nb, double pos0, double ssize, const int *fbmap)
{
    c->nb = nb; c->pos0 = pos0; c->ssize = ssize;
    std::vector<int> fb(c->np);
    for (int r = 0; r < c->np; ++r) fb[r] = fbmap[(size_t)(c->h_iy[r] - 1) * c->nxx + (c->h_ix[r] - 1)];
    std::vector<double> hz(c->nzz + 2), dumtop(c->nzz);
    CUDA_TRY(c, cudaMemcpy(hz.data(), c->d_hz, (c->nzz + 2) * sizeof(double), cudaMemcpyDeviceToHost));
    double coreh = 0.0;
    for (int k = 0; k < c->nzz; ++k) coreh = coreh + hz[1 + k];
    c->coreh = coreh;
    double dum = 0.0;
    for (int k = c->nzz - 1; k >= 0; --k) { dumtop[k] = dum; dum = dum + hz[1 + k]; }
    if (!c->d_fb) { TRY(dev_alloc(c, &c->d_fb, c->np)); TRY(dev_alloc(c, &c->d_dumtop, c->nzz)); }
    if (c->d_bpos) { cudaFree(c->d_bpos); c->d_bpos = nullptr; }
    TRY(dev_alloc(c, &c->d_bpos, nb));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));       // dev_alloc's zero fill runs on c->stream, the blocking copies do not
    CUDA_TRY(c, cudaMemcpy(c->d_fb, fb.data(), c->np * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_dumtop, dumtop.data(), c->nzz * sizeof(double), cudaMemcpyHostToDevice));
    return ADP_OK;
}

extern "C" int adp_set_crod(adp_ctx *c, int nb, double pos0, double ssize, const int *fbmap, const double *dsigtr,
                            const double *dsiga, const double *dnuf, const double *dsigf, const double *dsigs)
{
    if (!c || !fbmap || nb < 1 || !dsigtr || !dsiga || !dnuf || !dsigf || !dsigs) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_crod: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(upload_tables(c, &c->d_dtab, dsigtr, dsiga, dnuf, dsigf, dsigs));
    return set_crod_geometry(c, nb, pos0, ssize, fbmap);
}

extern "C" int adp_set_crod_map(adp_ctx *c, int nb, double pos0, double ssize, const int *fbmap)
{
    if (!c || !fbmap || nb < 1) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_crod_map: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return set_crod_geometry(c, nb, pos0, ssize, fbmap);
}

// After a device-side XS update: the STOP flag the kernel may have raised, agreed on by all ranks.
static int xs_update_stop(adp_ctx *c)
{
    CUDA_TRY(c, cudaMemcpyAsync(c->h_flags, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (c->nranks > 1) {      // any rank's STOP stops all of them
        double f = (double)c->h_flags[0];
        CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_TMP1, &f, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_TMP1, 1));
        CUDA_TRY(c, cudaMemcpyAsync(&f, c->d_scal + S_TMP1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        c->h_flags[0] = (int)f;
    }
    switch (c->h_flags[0]) {
    case 0: return ADP_OK;
    case ADP_STOP_XTAB_RANGE:
        c->err = "ERROR: A TH PARAMETER OR THE BORON CONCENTRATION IS OUT OF THE RANGE OF THE BRANCH PARAMETER";
        return ADP_STOP_XTAB_RANGE;
    case ADP_STOP_XTAB_NOROD:
        c->err = "CONTROL ROD BANK COINCIDES WITH A MATERIAL THAT DOES NOT HAVE CONTROL ROD DATA IN XTAB FILE";
        return ADP_STOP_XTAB_NOROD;
    default:
        c->err = "Negative diffusion coefficient encountered / ERROR IN THE CROSS SECTIONS (removal, nu*fission or scattering XS is negative)";
        return ADP_STOP_XS_CHECK;
    }
}

extern "C" int adp_xs_update(adp_ctx *c, const double *bpos)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->d_xtab != nullptr, "adp_xs_update: call adp_set_material_xs first");
    ADP_REQUIRE(c, c->xs_set, "adp_xs_update: chi / dc / exsrc come from adp_set_xs (call it once)");
    ADP_REQUIRE(c, !c->d_fb || bpos, "adp_xs_update: bank positions missing");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (c->d_fb) CUDA_TRY(c, cudaMemcpyAsync(c->d_bpos, bpos, c->nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_xs_update(c));
    return xs_update_stop(c);
}

extern "C" int adp_set_feedback(adp_ctx *c, int which, double ref, const double *dsigtr, const double *dsiga,
                                const double *dnuf, const double *dsigf, const double *dsigs)
{   // %BCON / %CBCS (which 0), %FTEM (1), %MTEM (2), %CDEN (3): reference value and the cross-section
    // changes per unit change of the parameter (mod_io.f90:2486-2957)
    if (!c || !dsigtr || !dsiga || !dnuf || !dsigf || !dsigs) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_feedback: geometry not set");
    ADP_REQUIRE(c, which >= 0 && which < 4, "adp_set_feedback: which must be 0 (bcon), 1 (ftem), 2 (mtem) or 3 (cden)");
    CUDA_TRY(c, cudaSetDevice(c->device));
    c->fref[which] = ref;
    return upload_tables(c, &c->d_ftab[which], dsigtr, dsiga, dnuf, dsigf, dsigs);
}

extern "C" int adp_xs_update_th(adp_ctx *c, double bcon, const double *ftem, const double *mtem, const double *cden,
                                const double *bpos)
{   // XS_updt(bcon, ftem, mtem, cden, bpos) (mod_xsec.f90:11-46) with the feedback cards given by
    // adp_set_feedback; ftem / mtem / cden: host (nnod) or NULL = the state adp_th_upd left on the device
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->d_xtab != nullptr, "adp_xs_update_th: call adp_set_material_xs first");
    ADP_REQUIRE(c, c->xs_set, "adp_xs_update_th: chi / dc / exsrc come from adp_set_xs (call it once)");
    ADP_REQUIRE(c, !c->d_fb || bpos, "adp_xs_update_th: bank positions missing");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_th_alloc(c));
    const double *host[3] = {ftem, mtem, cden};
    double *dev[3] = {c->d_ftem, c->d_mtem, c->d_cden};
    for (int i = 0; i < 3; ++i) {
        if (!c->d_ftab[1 + i]) continue;
        ADP_REQUIRE(c, host[i] || c->th_state_set, "adp_xs_update_th: a feedback parameter is neither passed nor on the device");
        if (host[i]) TRY(adp_upload_nodes(c, dev[i], host[i], 1));
    }
    c->bcon = bcon;
    if (c->d_fb) CUDA_TRY(c, cudaMemcpyAsync(c->d_bpos, bpos, c->nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    c->xs_feedback = true;
    TRY(adp_lazy_sync(c));
    const int rc = adp_k_xs_update(c);
    c->xs_feedback = false;
    if (rc) return rc;
    return xs_update_stop(c);
}

// ---- %XTAB branch tables: XStab_updt on the device ----------------------------------------------
extern "C" int adp_set_xtab(adp_ctx *c, const int *dims, const int *trod, const double *par, const double *xs,
                            const double *rxs)
{
    if (!c || !dims || !trod || !par || !xs) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_xtab: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const int nmat = c->nmat;
    std::vector<int> meta((size_t)nmat * 6);
    std::vector<long long> toff(nmat);
    long long npar = 0, ntab = 0;
    bool any_rod = false;
    ADP_REQUIRE(c, xtab_pack_meta(nmat, c->ng, dims, trod, meta.data(), toff.data(), &npar, &ntab, &any_rod),
                "adp_set_xtab: ERROR: MINIMUM NUMBER OF BRANCH IS 1");
    ADP_REQUIRE(c, !any_rod || rxs, "adp_set_xtab: a material has a rodded set (trod = 1) but rxs is NULL");
    free_xtab(c);
    // (no zero-fill: every byte is overwritten by the blocking copies below)
    TRY(dev_alloc(c, &c->d_brmeta, meta.size(), false)); TRY(dev_alloc(c, &c->d_brtoff, (size_t)nmat, false));
    TRY(dev_alloc(c, &c->d_brpar, (size_t)npar, false)); TRY(dev_alloc(c, &c->d_brtab, (size_t)ntab, false));
    if (any_rod) TRY(dev_alloc(c, &c->d_brrtab, (size_t)ntab, false));
    CUDA_TRY(c, cudaMemcpy(c->d_brmeta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_brtoff, toff.data(), (size_t)nmat * sizeof(long long), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_brpar, par, (size_t)npar * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->d_brtab, xs, (size_t)ntab * sizeof(double), cudaMemcpyHostToDevice));
    if (any_rod) CUDA_TRY(c, cudaMemcpy(c->d_brrtab, rxs, (size_t)ntab * sizeof(double), cudaMemcpyHostToDevice));
    return ADP_OK;
}

extern "C" int adp_xs_update_xtab(adp_ctx *c, double bcon, const double *ftem, const double *mtem, const double *cden,
                                  const double *bpos)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->d_brtab != nullptr, "adp_xs_update_xtab: call adp_set_xtab first");
    ADP_REQUIRE(c, c->xs_set, "adp_xs_update_xtab: chi / exsrc come from adp_set_xs (call it once)");
    ADP_REQUIRE(c, !c->d_fb || bpos, "adp_xs_update_xtab: bank positions missing");
    ADP_REQUIRE(c, !c->d_fb || c->d_brrtab, "adp_xs_update_xtab: control rods but no rodded tables");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_th_alloc(c));
    const double *host[3] = {ftem, mtem, cden};
    double *dev[3] = {c->d_ftem, c->d_mtem, c->d_cden};
    for (int i = 0; i < 3; ++i) {
        ADP_REQUIRE(c, host[i] || c->th_state_set, "adp_xs_update_xtab: a feedback parameter is neither passed nor on the device");
        if (host[i]) TRY(adp_upload_nodes(c, dev[i], host[i], 1));
    }
    c->bcon = bcon;
    if (c->d_fb) CUDA_TRY(c, cudaMemcpyAsync(c->d_bpos, bpos, c->nb * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_xs_update_xtab(c));
    return xs_update_stop(c);
}

extern "C" int adp_get_dc(adp_ctx *c, double *dc)
{
    if (!c || !dc) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->xs_set, "adp_get_dc: cross sections not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_lazy_sync(c));
    TRY(download_nodes(c, dc, c->d_dc, c->ng * 6));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_get_xs(adp_ctx *c, double *D, double *sigr, double *nuf, double *sigf, double *sigs)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->xs_set, "adp_get_xs: cross sections not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (D) TRY(download_nodes(c, D, c->d_D, c->ng));
    if (sigr) TRY(download_nodes(c, sigr, c->d_sigr, c->ng));
    if (nuf) TRY(download_nodes(c, nuf, c->d_nuf, c->ng));
    if (sigf) { TRY(adp_lazy_sync(c)); TRY(download_nodes(c, sigf, c->d_sigf, c->ng)); }
    if (sigs) TRY(download_nodes(c, sigs, c->d_sigs, c->ng * c->ng));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

// ---- transient time-step glue on the device -----------------------------
// These replace, optionally, host loops of mod_trans.f90 so that a time step does not move
// flux-sized arrays over PCIe: only the new cross sections go up and a few scalars come back.
extern "C" int adp_save_adjoint(adp_ctx *c)
{   // `af = f0` after outer_ad (mod_trans.f90:65-66), kept on the device
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_save_adjoint: no flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->d_af) TRY(dev_alloc(c, &c->d_af, (size_t)c->ng * c->NV));
    for (int g = 0; g < c->ng; ++g)
        CUDA_TRY(c, cudaMemcpyAsync(c->d_af + (size_t)g * c->NV, c->d_f0[c->cur[g]] + (size_t)g * c->NV, c->NV * sizeof(double),
                                    cudaMemcpyDeviceToDevice, c->stream));
    return ADP_OK;
}
extern "C" int adp_ipden(adp_ctx *c)
{   // iPden (mod_trans.f90:561-597)
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->kinetics_set && c->have_flux, "adp_ipden: needs adp_set_kinetics and a flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return adp_k_ipden(c);
}
extern "C" int adp_upden(adp_ctx *c, double ht)
{   // uPden (mod_trans.f90:601-644); fst is the fission source saved by adp_begin_time_step
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->kinetics_set && c->have_flux, "adp_upden: needs adp_set_kinetics and a flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return adp_k_upden(c, ht);
}
extern "C" int adp_begin_time_step(adp_ctx *c, double ht)
{   // trans_calc (mod_trans.f90:398-416) after XS_updt: sigrp = sigr; sigr += 1/(sth v ht) + omeg/v; ft = f0; fst = fs0
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->kinetics_set && c->have_flux, "adp_begin_time_step: needs adp_set_kinetics and a flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    return adp_k_begin_step(c, ht);
}
extern "C" int adp_update_omeg(adp_ctx *c, double ht, int bextr)
{   // rod_eject (mod_trans.f90:128-134,150-154): omeg = LOG(f0 / ft) / ht (%EXTR) or 0; ft is the flux
    // saved by the previous adp_begin_time_step, so call it before the next one
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_update_omeg: no flux");
    ADP_REQUIRE(c, ht > 0.0, "adp_update_omeg: time step must be positive");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(ensure_transient(c));
    return adp_k_omeg(c, ht, bextr);
}
extern "C" int adp_powtot(adp_ctx *c, double *tpow)
{   // PowTot (mod_trans.f90:523-557) of the current flux
    if (!c || !tpow) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_powtot: no flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_powdis(c, c->d_stage));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    *tpow = c->h_scal[S_POW];
    return ADP_OK;
}
extern "C" int adp_reactivity(adp_ctx *c, int use_sigrp, double *rho)
{   // reactivity(af, sigr | sigrp, rho) (mod_trans.f90:648-688): fills L (Lxyz) and returns rho;
    // use_sigrp = 0: the removal term uses the current sigr (before the first step, :95), 1: sigrp
    if (!c || !rho) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->d_af != nullptr, "adp_reactivity: call adp_save_adjoint after outer_ad first");
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_reactivity: needs matrix and flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(ensure_transient(c));
    TRY(adp_k_lxyz_total(c, c->d_L));
    TRY(adp_k_reactivity(c, c->d_af, use_sigrp ? c->d_sigrp : c->d_sigr));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    const double src = c->h_scal[S_TMP0], rem = c->h_scal[S_TMP1], lea = c->h_scal[S_E2SQ], fde = c->h_scal[S_FINT];
    *rho = (src - lea - rem) / fde;
    return ADP_OK;
}

// ---- state ------------------------------------------------------------------------------------
extern "C" int adp_get_state(adp_ctx *c, double *f0, double *fs0, double *s0, double *Ke)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_get_state: no flux yet");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (f0)
        for (int g = 0; g < c->ng; ++g)
            TRY(download_nodes(c, f0 + (size_t)g * c->nnod, c->d_f0[c->cur[g]] + (size_t)g * c->NV, 1));
    if (fs0) TRY(download_nodes(c, fs0, c->d_fs[c->fcur], 1));
    if (s0) {
        // TSrc* zero all of s0 and fill only the column of the group being solved
        // (mod_cmfd.f90:1022,1053,1084): after an outer iteration only the last group's column is non-zero
        const bool whole = c->nranks == 1 || c->gather_results;
        for (int g = 0; g < c->ng; ++g) {
            double *dst = s0 + (size_t)g * c->nnod + (whole ? 0 : (size_t)c->k0 * c->np);
            if (g + 1 == c->s0_group) TRY(download_nodes(c, s0 + (size_t)g * c->nnod, c->d_s0, 1));
            else memset(dst, 0, (whole ? (size_t)c->nnod : (size_t)c->nzl * c->np) * sizeof(double));
        }
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (Ke) *Ke = c->h_scal[S_KE];
    return ADP_OK;
}

extern "C" int adp_get_state_mask(adp_ctx *c, int mask, double *f0, double *fs0, double *s0, double *Ke)
{
    return adp_get_state(c, (mask & 1) ? f0 : nullptr, (mask & 2) ? fs0 : nullptr, (mask & 4) ? s0 : nullptr, Ke);
}

extern "C" int adp_set_state(adp_ctx *c, const double *f0, const double *fs0, double Ke)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_state: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (f0)
        for (int g = 0; g < c->ng; ++g)
            TRY(upload_nodes(c, c->d_f0[c->cur[g]] + (size_t)g * c->NV, f0 + (size_t)g * c->nnod, 1, false));
    if (fs0) TRY(upload_nodes(c, c->d_fs[c->fcur], fs0, 1, false));
    CUDA_TRY(c, cudaMemcpyAsync(c->d_scal + S_KE, &Ke, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (f0) memset(c->xghost_valid, 0, sizeof(c->xghost_valid));
    if (f0 && fs0) c->have_flux = true;
    c->outer_first = false;
    return ADP_OK;
}

extern "C" int adp_set_s0(adp_ctx *c, const double *s0, int g)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set && g >= 0 && g <= c->ng, "adp_set_s0: geometry not set or bad group");
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (g > 0) {
        ADP_REQUIRE(c, s0 != nullptr, "adp_set_s0: s0 missing");
        TRY(upload_nodes(c, c->d_s0, s0 + (size_t)(g - 1) * c->nnod, 1, false));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    c->s0_group = g;
    return ADP_OK;
}

// device [g][6][NV]  <->  host df(6, nnod, ng) (face fastest) via the pinned staging buffer
static int ensure_host_stage(adp_ctx *c, size_t elems)
{
    if (c->stage_elems >= elems) return ADP_OK;
    if (c->h_stage) cudaFreeHost(c->h_stage);
    c->h_stage = nullptr;
    CUDA_TRY(c, cudaMallocHost((void **)&c->h_stage, elems * sizeof(double)));
    c->stage_elems = elems;
    return ADP_OK;
}

extern "C" int adp_get_nod(adp_ctx *c, double *df, double *dn)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->matrix_ready, "adp_get_nod: call adp_matrix_setup first");
    CUDA_TRY(c, cudaSetDevice(c->device));
    // one global host column at a time through the pinned stage (own rows; with gather_results the other ranks' too),
    // scattered into the reference's AoS nod(n,g)%df(6) / %dn(6)
    const bool whole = c->nranks == 1 || c->gather_results;
    const size_t lo = whole ? 0 : (size_t)c->k0 * c->np, hi = whole ? (size_t)c->nnod : (size_t)c->k1 * c->np;
    TRY(ensure_host_stage(c, (size_t)c->nnod));
    for (int which = 0; which < 2; ++which) {
        double *dst = which ? dn : df;
        const double *src = which ? c->d_dn : c->d_df;
        if (!dst) continue;
        for (int g = 0; g < c->ng; ++g)
            for (int f = 0; f < 6; ++f) {
                TRY(download_nodes(c, c->h_stage, src + ((size_t)g * 6 + f) * c->NV, 1));
                CUDA_TRY(c, cudaStreamSynchronize(c->stream));
                double *o = dst + (size_t)g * c->nnod * 6 + f;
                for (size_t n = lo; n < hi; ++n) o[n * 6] = c->h_stage[n];
            }
    }
    return ADP_OK;
}

extern "C" int adp_set_nod_dn(adp_ctx *c, const double *dn)
{
    if (!c || !dn) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->geometry_set, "adp_set_nod_dn: geometry not set");
    CUDA_TRY(c, cudaSetDevice(c->device));
    // include one ghost plane on interior slab boundaries
    const int ka = std::max(0, c->k0 - 1), kb = std::min(c->nzz, c->k1 + 1);
    const size_t cnt = (size_t)(kb - ka) * c->np, off = (size_t)ka * c->np;
    TRY(ensure_host_stage(c, cnt));
    for (int g = 0; g < c->ng; ++g)
        for (int f = 0; f < 6; ++f) {
            const double *in = dn + ((size_t)g * c->nnod + off) * 6 + f;
            for (size_t n = 0; n < cnt; ++n) c->h_stage[n] = in[n * 6];
            CUDA_TRY(c, cudaMemcpyAsync(c->d_dn + ((size_t)g * 6 + f) * c->NV + (size_t)(ka - (c->k0 - ADP_GH)) * c->np,
                                        c->h_stage, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        }
    c->coup_first = false;
    return ADP_OK;
}

// ---- kernel-level entry points --------------------------------------------------------------
extern "C" int adp_sp_matvec(adp_ctx *c, int g, const double *x, double *v)
{
    if (!c || !x || !v) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->matrix_ready && g >= 1 && g <= c->ng, "adp_sp_matvec: matrix not set up or bad group");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(upload_nodes(c, c->d_p, x, 1, false));
    TRY(adp_k_spmv(c, g - 1, c->d_p, c->d_v));
    TRY(download_nodes(c, v, c->d_v, 1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_bicg(adp_ctx *c, int imax, int g, const double *b, double *x)
{
    if (!c || !b || !x) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->matrix_ready && g >= 1 && g <= c->ng, "adp_bicg: matrix not set up or bad group");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(upload_nodes(c, c->d_stage, b, 1, false));
    // x lives in the spare flux buffer of group g
    double *dx = c->d_f0[c->cur[g - 1] ^ 1] + (size_t)(g - 1) * c->NV;
    TRY(upload_nodes(c, dx, x, 1, false));
    TRY(adp_k_bicg_raw(c, g - 1, imax, c->d_stage, dx));
    TRY(download_nodes(c, x, dx, 1));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_get_matrix(adp_ctx *c, double *a)
{
    if (!c || !a) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->matrix_ready, "adp_get_matrix: matrix not set up");
    CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t NL = (size_t)c->NL, off = (size_t)c->k0 * c->np;
    TRY(ensure_host_stage(c, NL));
    for (int g = 0; g < c->ng; ++g)
        for (int d = 0; d < 7; ++d) {
            CUDA_TRY(c, cudaMemcpyAsync(c->h_stage, c->d_a + ((size_t)g * 7 + d) * c->NV + (size_t)ADP_GH * c->np,
                                        NL * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            double *o = a + ((size_t)g * c->nnod + off) * 7 + d;
            for (size_t n = 0; n < NL; ++n) o[n * 7] = c->h_stage[n];
        }
    return ADP_OK;
}

extern "C" int adp_get_source(adp_ctx *c, int cmode, double *S1, double *S2, double *S3)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_get_source: needs matrix and flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_k_nodal_source(c, cmode));
    double *out[3] = {S1, S2, S3};
    for (int u = 0; u < 3; ++u)
        if (out[u]) TRY(download_nodes(c, out[u], c->d_S + (size_t)u * c->ng * c->NV, c->ng));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

// L(n,g) = L1 + L2 + L3 of Lxyz (mod_nodal.f90:901-1005) for all nodes, as `reactivity`
// (mod_trans.f90:677-678) forms it -- on the device, so that the time-step driver reads back
// nnod*ng doubles instead of nod%df/dn (24x as much) and loops over Lxyz on the host.
extern "C" int adp_lxyz_total(adp_ctx *c, double *L)
{
    if (!c || !L) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_lxyz_total: needs matrix and flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(ensure_transient(c));
    TRY(adp_k_lxyz_total(c, c->d_L));
    TRY(download_nodes(c, L, c->d_L, c->ng));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

extern "C" int adp_set_trace(adp_ctx *c, adp_trace_fn fn, void *user)
{
    if (!c) return ADP_ERR_USAGE;
    c->trace = fn; c->trace_user = user;
    return ADP_OK;
}

extern "C" int adp_launch_count(const adp_ctx *c, long long *launches)
{
    if (!c || !launches) return ADP_ERR_USAGE;
    *launches = c->launches;
    return ADP_OK;
}

extern "C" int adp_set_option(adp_ctx *c, const char *name, int value)
{
    if (!c || !name) return ADP_ERR_USAGE;
    if (!strcmp(name, "profile")) {
        // 1: start a per-launch profile of the CMFD kernels (run without CUDA graphs); read it with adp_profile_report
        for (auto &pe : c->prof_ev) cudaEventDestroy(pe.second);
        c->prof_ev.clear();
        if (c->prof_start) { cudaEventDestroy(c->prof_start); c->prof_start = nullptr; }
        c->prof = value != 0;
        if (c->prof) {
            CUDA_TRY(c, cudaSetDevice(c->device));
            CUDA_TRY(c, cudaEventCreate(&c->prof_start));
            CUDA_TRY(c, cudaEventRecord(c->prof_start, c->stream));
        }
        return ADP_OK;
    }
    if (!strcmp(name, "lazy_adf")) {
        if (!value) { const int rc = adp_lazy_sync(c); if (rc) return rc; }
        c->lazy_adf = value != 0;
        return ADP_OK;
    }
    if (!strcmp(name, "reset_nodal")) {
        // back to the state before the first coup_coef call of a run: the next adp_matrix_setup(1) zeroes dn
        // (mod_cmfd.f90:28-36, first call) and ndmax starts at 0 again (mod_data.f90:199) -- repeated timing passes
        c->coup_first = true; c->ndmax = 0.0; c->im = c->jm = c->km = 0;
        return ADP_OK;
    }
    if (!strcmp(name, "graphs")) { c->use_graphs = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "peer_push")) { if (!value) c->peer_ok = false; return ADP_OK; }
    if (!strcmp(name, "peer_allreduce")) { if (!value) c->peer_ar = false; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "gather_results")) { c->gather_results = value != 0; return ADP_OK; }
    if (!strcmp(name, "fuse_mail")) { c->fuse_mail = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "mail_ll")) { c->mail_ll = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "mail_timeout_s")) { c->mail_timeout_s = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "balance_rounds")) { c->balance_rounds = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "bench_warmup")) { c->bench_warmup = value; return ADP_OK; }
    if (!strcmp(name, "nodal_coop")) { c->nodal_coop = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "nodal_fused")) { c->nodal_fused = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "spmv_var")) { c->spmv_var = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "st_var")) { c->st_var = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "st_m_var")) { c->st_m_var = value; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "st_tma")) { c->st_tma = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "fuse_st")) { c->fuse_st = value != 0; free_graphs(c); return ADP_OK; }
    if (!strcmp(name, "grid_blocks")) {
        ADP_REQUIRE(c, value >= 1 && value <= ADP_MAXPART, "grid_blocks out of range");
        c->grid_blocks = value; c->grid_override = true; free_graphs(c); return ADP_OK;
    }
    c->err = std::string("unknown option ") + name;
    return ADP_ERR_USAGE;
}

// ---- nodal update without host round trip (used by adp_outer_steps) -------------------------
static int enqueue_nodal_upd(adp_ctx *c, int nmode)
{
    CUDA_TRY(c, cudaMemsetAsync(c->d_scal + S_NDMAX, 0, sizeof(double), c->stream));
    TRY(adp_k_nodal_update(c, nmode));
    if (c->nranks > 1) TRY(adp_comm_allreduce_max_nccl(c, c->d_scal + S_NDMAX, 1));
    TRY(adp_k_matrix_setup(c));
    return ADP_OK;
}

// Enqueue `nsteps` consecutive passes of the outer loop body (p = p_first ...), including the
// nodal update + matrix_setup(0) whenever mod(p, nupd) == 0, WITHOUT the per-iteration host
// read-back and exit test; one synchronisation at the end.  This is the device-resident
// throughput path (bench `value`); adp_outer() with its per-iteration test is the product path.
extern "C" int adp_outer_steps(adp_ctx *c, int mode, int p_first, int nsteps, double *Ke, double *ser, double *fer)
{
    if (!c) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_outer_steps: needs adp_matrix_setup and a flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    CUDA_TRY(c, cudaMemsetAsync(c->d_errflag, 0, sizeof(int), c->stream));
    for (int p = p_first; p < p_first + nsteps; ++p) {
        const bool extrap = (p % c->nac) == 0;
        TRY(launch_outer_iter(c, mode, extrap));
        if (p % c->nupd == 0 && c->kern != ADP_KERN_FDM) {
            const int nmode = (mode == ADP_MODE_ADJOINT) ? 0 : (mode == ADP_MODE_TRANSIENT) ? 2 : 1;
            TRY(enqueue_nodal_upd(c, nmode));
        }
    }
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_flags + 2, c->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (Ke) *Ke = c->h_scal[S_KE];
    if (ser) *ser = c->h_scal[S_SER];
    if (fer) *fer = c->h_scal[S_FER];
    if (c->h_flags[2] != 0) { c->err = "ERROR IN MATRIX DECOMP: DIAGONAL ELEMENTS CLOSE TO ZERO"; return ADP_STOP_LU_DIAG; }
    return ADP_OK;
}

// CUDA-event timer on the library's stream (torch.cuda.Event only sees torch's stream)
extern "C" int adp_timer_start(adp_ctx *c)
{
    if (!c) return ADP_ERR_USAGE;
    CUDA_TRY(c, cudaSetDevice(c->device));
    if (!c->ev0) { CUDA_TRY(c, cudaEventCreate(&c->ev0)); CUDA_TRY(c, cudaEventCreate(&c->ev1)); }
    CUDA_TRY(c, cudaEventRecord(c->ev0, c->stream));
    return ADP_OK;
}
extern "C" int adp_timer_stop(adp_ctx *c, double *ms)
{
    if (!c || !ms || !c->ev0) return ADP_ERR_USAGE;
    CUDA_TRY(c, cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(c->ev1));
    float f = 0.f;
    CUDA_TRY(c, cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    return ADP_OK;
}

int adp_k_bench_one(adp_ctx *c, int what, int g);

// device-resident micro-benchmarks of one kernel class on the current problem (no host copies
// in the timed region).  what: 0 B SpMV+dot, 1 C fused s/t, 2 D, 3 A, 4 P, 5 F, 8 plain SpMV,
// 6 nodal source, 7 whole nodal update (source + 3 surface launches), 9 matrix_setup(0).
// Launches alternate over the energy groups so that consecutive launches stream different
// matrices (working set >> L2).  Returns the average device time per launch in ms.
extern "C" int adp_profile_report(adp_ctx *c, int max, int *lines, int *counts, double *ms)
{
    if (!c || !lines || !counts || !ms || max < 1) return -1;
    if (!c->prof_start || c->prof_ev.empty()) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    std::map<int, std::pair<int, double>> acc;
    cudaEvent_t prev = c->prof_start;
    for (auto &pe : c->prof_ev) {
        float dt = 0.f;
        cudaEventElapsedTime(&dt, prev, pe.second);
        auto &a = acc[pe.first];
        a.first++; a.second += dt;
        prev = pe.second;
    }
    int n = 0;
    for (auto &kv : acc) {
        if (n >= max) break;
        lines[n] = kv.first; counts[n] = kv.second.first; ms[n] = kv.second.second;
        ++n;
    }
    return n;
}

extern "C" int adp_bench_kernel(adp_ctx *c, int what, int reps, double *avg_ms)
{
    if (!c || !avg_ms || reps < 1) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux && c->matrix_ready, "adp_bench_kernel: needs matrix and flux");
    CUDA_TRY(c, cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    CUDA_TRY(c, cudaEventCreate(&e0));
    CUDA_TRY(c, cudaEventCreate(&e1));
    int rc = ADP_OK;
    auto body = [&](int i) -> int {
        switch (what) {
        case 6: return adp_k_nodal_source(c, 1);
        case 7: return adp_k_nodal_update(c, 1);
        case 9: return adp_k_matrix_setup(c);
        default: return adp_k_bench_one(c, what, i % c->ng);
        }
    };
    for (int i = 0; i < c->bench_warmup && !rc; ++i) rc = body(i);
    CUDA_TRY(c, cudaEventRecord(e0, c->stream));
    for (int i = 0; i < reps && !rc; ++i) rc = body(i);
    CUDA_TRY(c, cudaEventRecord(e1, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *avg_ms = (double)ms / reps;
    return rc;
}
