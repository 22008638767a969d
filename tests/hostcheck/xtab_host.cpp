// Test infrastructure (NOT part of libadpres_b200.so): the per-node %XTAB cross-section update of the
// CUDA library -- adpres_b200/csrc/xtab_node.cuh, the code k_xs_update_xtab runs per thread -- compiled
// for the host with g++, so that its arithmetic, table indexing and rod logic can be checked against the
// numpy restatement on machines without a GPU (tests/test_xtab.py).  The loop below mirrors the kernel
// body of k_xs_update_xtab (cmfd_kernels.cu) with idx = kg * np + r and no ghost planes.
#include <vector>
#include "../../adpres_b200/csrc/xtab_node.cuh"

extern "C" int xtab_host_update(int ng, int nmat, const int *dims, const int *trod, const double *par, const double *xs,
                                const double *rxs, int np, int nzz, const int *mat, const int *fb /* [np] or NULL */,
                                const double *bpos, const double *hz /* [nzz] */, double pos0, double ssize, double bcon,
                                const double *ftem, const double *mtem, const double *cden, double *D, double *sigr,
                                double *nuf, double *sigf, double *sigs, double *dc, double *w_out)
{
    std::vector<int> meta((size_t)nmat * 6);
    std::vector<long long> toff(nmat);
    long long npar = 0, ntab = 0;
    bool any_rod = false;
    if (!xtab_pack_meta(nmat, ng, dims, trod, meta.data(), toff.data(), &npar, &ntab, &any_rod)) return -2;
    XtabTables T;
    T.ng = ng; T.nval = 4 * ng + ng * ng + 6 * ng;
    T.meta = meta.data(); T.toff = toff.data(); T.par = par; T.xs = xs; T.rxs = rxs;
    XtabOut O{D, sigr, nuf, sigf, sigs, dc};
    // set_crod_geometry (capi.cu): core height and rod length above every plane, in the reference's order
    std::vector<double> dumtop(nzz);
    double coreh = 0.0;
    for (int k = 0; k < nzz; ++k) coreh = coreh + hz[k];
    double dum = 0.0;
    for (int k = nzz - 1; k >= 0; --k) { dumtop[k] = dum; dum = dum + hz[k]; }
    const long long NV = (long long)np * nzz;
    int flag = 0;
    for (int kg = 0; kg < nzz; ++kg)
        for (int r = 0; r < np; ++r) {
            const long long idx = (long long)kg * np + r;
            double w = -1.0;
            const int b = fb ? fb[r] : 0;
            if (b > 0) w = xt_rod_fraction(coreh - pos0 - bpos[b - 1] * ssize, dumtop[kg], hz[kg], kg == nzz - 1);
            if (w_out) w_out[idx] = w;
            const int rc = xtab_node(T, mat[idx] - 1, w, b > 0, cden[idx], bcon, ftem[idx], mtem[idx], O, NV, idx);
            if (rc) flag = rc;
        }
    return flag;
}
