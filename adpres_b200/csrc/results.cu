// Result reductions: AsmPow, AxiPow and AsmFlux of mod_io.f90:3267-3644 without the
// fx(nxx,nyy,nzz[,ng]) automatic arrays and without copying the node arrays to the host.
//
// The reference scatters the node vector into a dense box, averages every (i,j) column over z,
// then averages the columns of an assembly, then normalises.  Only the first stage touches
// O(nnod) data; it runs on the device (one thread per plane position marching through the planes
// in the reference's k order, so a single-rank column sum is bit-identical to the reference's
// serial loop).  What is left -- np column values per group, or nzz plane sums -- is finished on
// the host in exactly the reference's loop order.
#include "adp_internal.cuh"

#include <cmath>

#define RES_TILE 64   // np is only ~1e4..1e5: small CTAs spread the columns over all SMs

// column sums: AsmPow  summ = summ + fx(i,j,k)*zdel(k)                     (mod_io.f90:3305-3311)
//              AsmFlux summ = summ + fx(i,j,k,g)*xdel(i)*ydel(j)*zdel(k)   (mod_io.f90:3545-3551)
__global__ void __launch_bounds__(RES_TILE) k_column_sum(Geo G, const double *__restrict__ vec, int flux_weights,
                                                          double *__restrict__ out)
{
    const int r = blockIdx.x * RES_TILE + threadIdx.x;
    if (r >= G.np) return;
    const double hx = G.hx[r], hy = G.hy[r];
    double summ = 0.0;
    for (int kl = 0; kl < G.nzl; ++kl) {
        const double v = vec[(long long)(kl + ADP_GH) * G.np + r];
        const double hz = G.hz[1 + G.k0 + kl];
        summ = flux_weights ? summ + v * hx * hy * hz : summ + v * hz;
    }
    out[r] = summ;
}

// plane sums for AxiPow: summ = summ + fx(i,j,ztot) over one plane (mod_io.f90:3443-3448); one CTA
// per owned plane, fixed-shape tree (deterministic; differs from the serial sum by rounding only)
__global__ void __launch_bounds__(ADP_TILE) k_plane_sum(Geo G, const double *__restrict__ vec, double *__restrict__ out)
{
    __shared__ double sm[ADP_TILE];
    const int kl = blockIdx.x;
    double acc = 0.0;
    for (int r = threadIdx.x; r < G.np; r += ADP_TILE) acc = acc + vec[(long long)(kl + ADP_GH) * G.np + r];
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = ADP_TILE / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[G.k0 + kl] = sm[0];
}

#define TRY(x)              \
    do {                    \
        int rc__ = (x);     \
        if (rc__) return rc__; \
    } while (0)

static int launch_check(adp_ctx *c)
{
    c->launches++;
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { c->err = std::string("kernel launch: ") + cudaGetErrorString(e); return ADP_ERR_CUDA; }
    return ADP_OK;
}

static int ensure_result_buffers(adp_ctx *c)
{
    const size_t need = (size_t)std::max(c->np, c->nzz);
    if (c->res_elems >= need) return ADP_OK;
    if (c->d_res) cudaFree(c->d_res);
    if (c->h_res) cudaFreeHost(c->h_res);
    c->d_res = nullptr; c->h_res = nullptr; c->res_elems = 0;
    CUDA_TRY(c, cudaMalloc((void **)&c->d_res, need * sizeof(double)));
    CUDA_TRY(c, cudaMallocHost((void **)&c->h_res, need * sizeof(double)));
    c->res_elems = need;
    return ADP_OK;
}

// column sums of one device vector -> c->h_res[0..np)
static int column_sums(adp_ctx *c, const double *d_vec, int flux_weights)
{
    TRY(ensure_result_buffers(c));
    k_column_sum<<<(c->np + RES_TILE - 1) / RES_TILE, RES_TILE, 0, c->stream>>>(c->geo, d_vec, flux_weights, c->d_res);
    TRY(launch_check(c));
    TRY(adp_comm_allreduce_sum_nccl(c, c->d_res, c->np));     // slabs hold partial columns
    CUDA_TRY(c, cudaMemcpyAsync(c->h_res, c->d_res, (size_t)c->np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return ADP_OK;
}

static int check_divisions(adp_ctx *c, int n, const int *div, int total, const char *what)
{
    long long s = 0;
    bool ok = n > 0 && div != nullptr;
    for (int i = 0; ok && i < n; ++i) { ok = div[i] > 0; s += div[i]; }
    if (!ok || s != total) { c->err = std::string(what) + ": divisions do not add up to the fine mesh"; return ADP_ERR_USAGE; }
    return ADP_OK;
}

// fasm(i,j) = sum_{ly,lx} fnode(lx,ly) xdel(lx) ydel(ly) / sum xdel ydel over the assembly's
// rectangle, columns outside the core outline counting as zero power but full area
// (mod_io.f90:3317-3341, 3557-3589)
static void assembly_average(const adp_ctx *c, const double *fnode_r, int nx, int ny, const int *xdiv, const int *ydiv,
                             double *fasm)
{
    const int nxx = c->nxx;
    int ys = 1, yf = 0;
    for (int j = 1; j <= ny; ++j) {
        yf += ydiv[j - 1];
        int xf = 0, xs = 1;
        for (int i = 1; i <= nx; ++i) {
            xf += xdiv[i - 1];
            double summ = 0.0, vsumm = 0.0;
            for (int ly = ys; ly <= yf; ++ly)
                for (int lx = xs; lx <= xf; ++lx) {
                    const int r = c->h_nodp[(size_t)(ly - 1) * nxx + (lx - 1)];
                    const double f = r ? fnode_r[r - 1] : 0.0;
                    summ = summ + f * c->h_xdel[lx - 1] * c->h_ydel[ly - 1];
                    vsumm = vsumm + c->h_xdel[lx - 1] * c->h_ydel[ly - 1];
                }
            fasm[(size_t)(j - 1) * nx + (i - 1)] = summ / vsumm;
            xs += xdiv[i - 1];
        }
        ys += ydiv[j - 1];
    }
}

extern "C" int adp_asm_pow(adp_ctx *c, int nx, int ny, const int *xdiv, const int *ydiv, double *fasm, int *xmax,
                           int *ymax)
{   // CALL PowDis(pow); CALL AsmPow(pow)   (mod_control.f90:35-44, mod_io.f90:3267-3405)
    if (!c || !fasm) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_asm_pow: no flux");
    TRY(check_divisions(c, nx, xdiv, c->nxx, "adp_asm_pow (x)"));
    TRY(check_divisions(c, ny, ydiv, c->nyy, "adp_asm_pow (y)"));
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_powdis(c, c->d_stage));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (c->h_scal[S_POW] <= 0.0) { c->err = "ERROR: TOTAL NODES POWER IS ZERO OR LESS"; return ADP_STOP_ZERO_POWER; }
    TRY(adp_k_scale_by_slot(c, c->d_stage, S_POW));
    TRY(column_sums(c, c->d_stage, 0));
    double vsumm = 0.0;
    for (int k = 0; k < c->nzz; ++k) vsumm = vsumm + c->h_zdel[k];
    std::vector<double> fnode(c->np);
    for (int r = 0; r < c->np; ++r) fnode[r] = c->h_res[r] / vsumm;
    assembly_average(c, fnode.data(), nx, ny, xdiv, ydiv, fasm);
    // normalise to a mean of 1 over the assemblies with power, find the maximum (mod_io.f90:3343-3363)
    int nfuel = 0;
    double totp = 0.0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
            if (fasm[(size_t)j * nx + i] > 0.0) { ++nfuel; totp = totp + fasm[(size_t)j * nx + i]; }
    int im = 1, jm = 1;
    double fmax = 0.0;
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            double &f = fasm[(size_t)j * nx + i];
            if (totp > 0.0) f = (double)(float)nfuel / totp * f;      // REAL(nfuel): default (single) real
            if (f > fmax) { im = i + 1; jm = j + 1; fmax = f; }
        }
    if (xmax) *xmax = im;
    if (ymax) *ymax = jm;
    return ADP_OK;
}

extern "C" int adp_axi_pow(adp_ctx *c, int nz, const int *zdiv, double *faxi, int *amax)
{   // CALL PowDis(pow); CALL AxiPow(pow)   (mod_io.f90:3409-3494)
    if (!c || !faxi) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_axi_pow: no flux");
    TRY(check_divisions(c, nz, zdiv, c->nzz, "adp_axi_pow (z)"));
    CUDA_TRY(c, cudaSetDevice(c->device));
    TRY(ensure_result_buffers(c));
    TRY(adp_lazy_sync(c));
    TRY(adp_k_powdis(c, c->d_stage));
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, c->d_scal, S_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    ADP_CHECK_FAULT(c);
    if (c->h_scal[S_POW] <= 0.0) { c->err = "ERROR: TOTAL NODES POWER IS ZERO OR LESS"; return ADP_STOP_ZERO_POWER; }
    TRY(adp_k_scale_by_slot(c, c->d_stage, S_POW));
    CUDA_TRY(c, cudaMemsetAsync(c->d_res, 0, (size_t)c->nzz * sizeof(double), c->stream));
    k_plane_sum<<<c->nzl, ADP_TILE, 0, c->stream>>>(c->geo, c->d_stage, c->d_res);
    TRY(launch_check(c));
    TRY(adp_comm_allreduce_sum_nccl(c, c->d_res, c->nzz));    // every rank filled only its own planes
    CUDA_TRY(c, cudaMemcpyAsync(c->h_res, c->d_res, (size_t)c->nzz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    // plane volume in the reference's order: sum over (j, i) of vdel = xdel*ydel*zdel
    int nfuel = 0, ztot = 0;
    double totp = 0.0;
    for (int k = 0; k < nz; ++k) {
        double summ = 0.0, vsumm = 0.0;
        for (int lz = 0; lz < zdiv[k]; ++lz, ++ztot) {
            summ = summ + c->h_res[ztot];
            for (int r = 0; r < c->np; ++r)
                vsumm = vsumm + c->h_xdel[c->h_ix[r] - 1] * c->h_ydel[c->h_iy[r] - 1] * c->h_zdel[ztot];
        }
        faxi[k] = summ / vsumm;
        if (faxi[k] > 0.0) { ++nfuel; totp = totp + faxi[k]; }
    }
    double fmax = 0.0;
    int am = 1;
    for (int k = 0; k < nz; ++k) {
        faxi[k] = (double)(float)nfuel / totp * faxi[k];
        if (faxi[k] > fmax) { am = k + 1; fmax = faxi[k]; }
    }
    if (amax) *amax = am;
    return ADP_OK;
}

extern "C" int adp_asm_flux(adp_ctx *c, int nx, int ny, const int *xdiv, const int *ydiv, int use_norm, double norm,
                            double *fasm, int *negf)
{   // CALL AsmFlux(f0 [, norm])   (mod_io.f90:3498-3644); fasm(nx,ny,ng)
    if (!c || !fasm) return ADP_ERR_USAGE;
    ADP_REQUIRE(c, c->have_flux, "adp_asm_flux: no flux");
    TRY(check_divisions(c, nx, xdiv, c->nxx, "adp_asm_flux (x)"));
    TRY(check_divisions(c, ny, ydiv, c->nyy, "adp_asm_flux (y)"));
    CUDA_TRY(c, cudaSetDevice(c->device));
    int neg = 0;
    std::vector<double> fnode(c->np), vcol(c->np);
    for (int r = 0; r < c->np; ++r) {      // vsumm = sum_k xdel(i)*ydel(j)*zdel(k)
        double v = 0.0;
        const double hx = c->h_xdel[c->h_ix[r] - 1], hy = c->h_ydel[c->h_iy[r] - 1];
        for (int k = 0; k < c->nzz; ++k) v = v + hx * hy * c->h_zdel[k];
        vcol[r] = v;
    }
    for (int g = 0; g < c->ng; ++g) {
        TRY(column_sums(c, c->d_f0[c->cur[g]] + (size_t)g * c->NV, 1));
        for (int r = 0; r < c->np; ++r) fnode[r] = c->h_res[r] / vcol[r];
        double *fg = fasm + (size_t)g * nx * ny;
        assembly_average(c, fnode.data(), nx, ny, xdiv, ydiv, fg);
        double totp = 0.0;
        for (int q = 0; q < nx * ny; ++q) {       // (j outer, i inner) = storage order
            if (fg[q] > 0.0) totp = totp + fg[q];
            if (fg[q] < 0.0) neg = 1;
        }
        if (use_norm)
            for (int q = 0; q < nx * ny; ++q) fg[q] = norm / totp * fg[q] * norm;     // sic (mod_io.f90:3598)
    }
    if (negf) *negf = neg;
    return ADP_OK;
}
