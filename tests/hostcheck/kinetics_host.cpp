// Test infrastructure (NOT part of libadpres_b200.so): the per-node bxtab = 1 kinetics code of the CUDA
// library -- adpres_b200/csrc/kinetics_node.cuh, what k_ipden_xtab / k_upden_xtab / k_get_exsrc_xtab /
// k_begin_step_xtab run per thread -- compiled for the host with g++ and looped over all nodes, so that it
// can be compared with the C oracle on machines without a GPU (tests/test_xtab.py).  Arrays are laid out
// like the device's: [column][NV] with NV = nnod (no ghost planes).
#include "../../adpres_b200/csrc/kinetics_node.cuh"

static KinTab tab(int ng, int nmat, const double *mlamb, const double *mibeta, const double *mvelo)
{
    KinTab K;
    K.ng = ng; K.nmat = nmat; K.lamb = mlamb; K.ibeta = mibeta; K.velo = mvelo;
    return K;
}

extern "C" void kin_host_ipden(int ng, int nmat, const double *mlamb, const double *mibeta, const double *mvelo, long long nnod,
                               const int *mat, const double *nuf, const double *fs, double *c0)
{
    const KinTab K = tab(ng, nmat, mlamb, mibeta, mvelo);
    for (long long n = 0; n < nnod; ++n) kx_ipden(K, mat[n] - 1, nuf[(size_t)(ng - 1) * nnod + n] > 0.0, fs[n], c0, nnod, n);
}

extern "C" void kin_host_upden(int ng, int nmat, const double *mlamb, const double *mibeta, const double *mvelo, long long nnod,
                               const int *mat, const double *nuf, double ht, const double *fst, const double *fs, double *c0)
{
    const KinTab K = tab(ng, nmat, mlamb, mibeta, mvelo);
    for (long long n = 0; n < nnod; ++n)
        kx_upden(K, mat[n] - 1, nuf[(size_t)(ng - 1) * nnod + n] > 0.0, ht, fst[n], fs[n], c0, nnod, n);
}

extern "C" void kin_host_exsrc(int ng, int nmat, const double *mlamb, const double *mibeta, const double *mvelo, long long nnod,
                               const int *mat, const double *nuf, double ht, double sth, double bth, const double *c0,
                               const double *fst, const double *tbeta, const double *chi, const double *L, const double *sigrp,
                               const double *ft, const double *s0col, int s0_group, const double *omeg, double *exsrc, double *dfis)
{
    const KinTab K = tab(ng, nmat, mlamb, mibeta, mvelo);
    KxExsrc A;
    A.ht = ht; A.sth = sth; A.bth = bth; A.c0 = c0; A.fst = fst; A.tbeta = tbeta; A.chi = chi; A.L = L; A.sigrp = sigrp;
    A.ft = ft; A.s0 = s0col; A.s0_group = s0_group; A.omeg = omeg; A.exsrc = exsrc; A.dfis = dfis;
    for (long long n = 0; n < nnod; ++n) kx_exsrc(K, A, mat[n] - 1, nuf[(size_t)(ng - 1) * nnod + n] > 0.0, nnod, n);
}

extern "C" void kin_host_time_absorption(int ng, int nmat, const double *mlamb, const double *mibeta, const double *mvelo,
                                         long long nnod, const int *mat, double sth, double ht, const double *omeg, double *sigr,
                                         double *sigrp)
{
    const KinTab K = tab(ng, nmat, mlamb, mibeta, mvelo);
    for (long long n = 0; n < nnod; ++n) kx_time_absorption(K, mat[n] - 1, sth, ht, omeg, sigr, sigrp, nnod, n);
}
