// kinetics_node.cuh -- the bxtab = 1 branches of the transient glue: kinetics data per material
// (m(mat(n))%iBeta / %lamb / %velo of an %XTAB library) instead of one core-wide set, precursors only
// where nuf(n, ng) > 0.  Reference: get_exsrc mod_cmfd.f90:898-925, iPden mod_trans.f90:574-585,
// uPden :617-629, trans_calc :405-411.  One node per call; __host__ __device__ like xtab_node.cuh so that
// tests/hostcheck can run the same source on the CPU (test infrastructure only).
#pragma once
#include <math.h>
#include "xtab_node.cuh"

#define KX_NF 6

struct KinTab {
    int ng, nmat;
    const double *lamb, *ibeta;         // [nmat][6]
    const double *velo;                 // [nmat][ng]
};

// iPden: c0(n,j) = iBeta(j) / lamb(j) * fs0(n) in fuel, 0 elsewhere
ADP_HD inline void kx_ipden(const KinTab &K, int m, bool fuel, double fs, double *c0, long long NV, long long idx)
{
    for (int j = 0; j < KX_NF; ++j) {
        if (fuel) {
            const double blamb = K.ibeta[m * KX_NF + j] / K.lamb[m * KX_NF + j];
            c0[(size_t)j * NV + idx] = blamb * fs;
        } else
            c0[(size_t)j * NV + idx] = 0.0;
    }
}

// uPden: fuel nodes only
ADP_HD inline void kx_upden(const KinTab &K, int m, bool fuel, double ht, double fst, double fs, double *c0, long long NV,
                            long long idx)
{
    if (!fuel) return;
    for (int i = 0; i < KX_NF; ++i) {
        const double lam = K.lamb[m * KX_NF + i], bet = K.ibeta[m * KX_NF + i];
        const double pxe = exp(-lam * ht);
        double a1 = (1.0 - pxe) / (lam * ht);
        const double a2 = 1.0 - a1;
        a1 = a1 - pxe;
        double *c = c0 + (size_t)i * NV + idx;
        *c = *c * pxe + bet / lam * (a1 * fst + a2 * fs);
    }
}

struct KxExsrc {
    double ht, sth, bth;
    const double *c0;                   // [6][NV]
    const double *fst, *tbeta, *chi;    // [NV], [nmat], [g][nmat]
    const double *L, *sigrp, *ft, *s0, *omeg;   // [G][NV]; s0: one column, that of group s0_group
    int s0_group;                       // 0-based group whose s0 column is non-zero (-1 none)
    double *exsrc, *dfis;               // [G][NV], [NV]
};

// get_exsrc
ADP_HD inline void kx_exsrc(const KinTab &K, const KxExsrc &A, int m, bool fuel, long long NV, long long idx)
{
    double dt = 0.0, dtp = 0.0, dfis = 0.0;
    const double fst = A.fst[idx];
    for (int i = 0; i < KX_NF; ++i) {
        const double lam = K.lamb[m * KX_NF + i], bet = K.ibeta[m * KX_NF + i];
        const double pxe = exp(-lam * A.ht);
        double a1 = fuel ? (1.0 - pxe) / (lam * A.ht) : 0.0;
        const double a2 = 1.0 - a1;
        a1 = a1 - pxe;
        const double c0 = A.c0[(size_t)i * NV + idx];
        dfis = dfis + bet * a2;
        dt = dt + lam * c0 * pxe + bet * a1 * fst;
        dtp = dtp + lam * c0;
    }
    A.dfis[idx] = dfis;
    for (int g = 0; g < K.ng; ++g) {
        const double chi = A.chi[g * K.nmat + m];
        const double ft = A.ft[(size_t)g * NV + idx];
        const double s0 = (g == A.s0_group) ? A.s0[idx] : 0.0;
        const double pthet = -A.L[(size_t)g * NV + idx] - A.sigrp[(size_t)g * NV + idx] * ft + s0 + (1.0 - A.tbeta[m]) * chi * fst + chi * dtp;
        A.exsrc[(size_t)g * NV + idx] = chi * dt + exp(A.omeg[(size_t)g * NV + idx] * A.ht) * ft / (A.sth * K.velo[m * K.ng + g] * A.ht) + A.bth * pthet;
    }
}

// trans_calc: sigrp = sigr; sigr = sigr + 1/(sth v ht) + omeg / v with v = m(mat(n))%velo(g); ft = f0 is left to the caller
ADP_HD inline void kx_time_absorption(const KinTab &K, int m, double sth, double ht, const double *omeg, double *sigr,
                                      double *sigrp, long long NV, long long idx)
{
    for (int g = 0; g < K.ng; ++g) {
        const double sr = sigr[(size_t)g * NV + idx], v = K.velo[m * K.ng + g];
        sigrp[(size_t)g * NV + idx] = sr;
        sigr[(size_t)g * NV + idx] = sr + 1.0 / (sth * v * ht) + omeg[(size_t)g * NV + idx] / v;
    }
}
