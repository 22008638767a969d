// nodal_kernels.cu -- the two-node nonlinear nodal update (reference: src/mod_nodal.f90).
//
// The reference walks every line of nodes sequentially and carries a2p, a4p, Lp1, Bcp,
// Ap..Hp from one interface to the next (mod_nodal.f90:53-125).  Every carried quantity is a
// pure function of one node and one direction, and the transverse-leakage sources S1..S3 are
// frozen by get_source before any dn is overwritten (mod_nodal.f90:49), so all surfaces are
// independent.  Here:
//   k_nodal_source    one thread per node: Lxyz -> S1,S2,S3              mod_nodal.f90:901-1043
//   k_nodal_abefgh    one thread per node: the SANM constants A,B,E,F,G,H per (direction, group)
//                     (sinh/cosh); they depend only on sigr, D and the mesh, so they are cached
//                     until the cross sections change                      :1409-1451
//   k_nodal_nodedir   one thread per (node, direction): everything the sweep carries for that
//                     node -- B matrix, transverse-leakage moments, a2 (GxG LU), a4 -- computed
//                     ONCE and stored ([3][G*G+7G][NV] doubles)            :1047-1405,740-825
//   k_nodal_surfaces  one thread per (node, direction): the surface on the node's "+" side
//                     (two-node 2Gx2G problem, or the one-node boundary problem), plus the
//                     "-" boundary surface if the node is the first of its line; a1, a3, the
//                     surface current and the new dn live in registers.    :282-698
// Each surface is evaluated with the reference's operation order (same Doolittle LU without
// pivoting, same accumulation order), so results differ from the sweep only through the libm
// (sinh/cosh) and not at all for PNM.
#include "adp_internal.cuh"

namespace {

// resident CTAs per SM asked of ptxas for the G <= 2 node-direction / surfaces kernels (A/B: tools/build_variant.sh -DNODAL_LB_ND=..)
#ifndef NODAL_LB_ND
#define NODAL_LB_ND 3
#endif
#ifndef NODAL_LB_SF
#define NODAL_LB_SF 2
#endif
#define FOR_EACH_ROW(G_, KLO, NPL)                                                                  \
    for (int tile__ = blockIdx.x; tile__ < (G_).tpp * (NPL); tile__ += gridDim.x)                   \
        for (int kl = (KLO) + tile__ / (G_).tpp, r = (tile__ % (G_).tpp) * ADP_TILE + threadIdx.x, \
                 once__ = 1;                                                                        \
             once__ && r < (G_).np; once__ = 0)

__device__ __forceinline__ long long node_idx(const Geo &G, int kl, int r)
{
    return (long long)(kl + ADP_GH) * G.np + r;
}

struct NodalArgs {
    int ng, nmat, cmode, kern;
    long long nnod_total;
    const double *f0[ADP_MAXG];          // current flux per group
    const double *f0a, *f0b;             // the two flux buffers [G][NV]; bit g of curmask: group g lives in f0b
    unsigned curmask;                    // (a pointer array indexed by a runtime g would be copied to a stack frame)
    const double *D, *sigr, *nuf, *exsrc; // [G][NV]
    const double *sigs;                  // [h][g][NV] = sigs(n,g,h)
    const double *chi;                   // [g][nmat]
    const double *dc;                    // [f][g][NV]
    const int *mat;
    const double *tbeta, *dfis;
    const double *df;                    // [g][6][NV]
    double *dn;                          // [g][6][NV]
    double *S;                           // [3][G][NV]
    double *nd;                          // [3][G*G+7G][NV]  per (direction, node): Bc, A, F, G, H, a2, a4, L1
    double *abefgh;                      // [3][6][G][NV]    SANM constants cache
    const double *scal;
    int *errflag;
};

// ---------------------------------------------------------------------------------------
// Lxyz + get_source (mod_nodal.f90:901-1043)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ADP_TILE, 4) k_nodal_source(Geo G, NodalArgs A, double *__restrict__ Ltot)
{
    const long long NV = G.NV;
    const int np = G.np;
    FOR_EACH_ROW(G, 0, G.nzl)
    {
        const long long idx = node_idx(G, kl, r);
        const int kg = G.k0 + kl;
        const unsigned fl = G.flag[r];
        const int ym = G.ypm[r], yp = G.ypp[r];
        const double hx = G.hx[r], hy = G.hy[r], hz = G.hz[1 + kg];
        for (int g = 0; g < A.ng; ++g) {
            const double *f0 = A.f0[g];
            const double *df = A.df + (size_t)g * 6 * NV, *dn = A.dn + (size_t)g * 6 * NV;
            const double fn = f0[idx];
            double jp, jm;
            // x
            if (fl & FLAG_XP) jp = (G.bc[0] == 2) ? 0.0 : df[idx] * fn - dn[idx] * fn;
            else { const double fp = f0[idx + 1]; jp = -df[idx] * (fp - fn) - dn[idx] * (fp + fn); }
            if (fl & FLAG_XM) jm = (G.bc[1] == 2) ? 0.0 : -df[NV + idx] * fn - dn[NV + idx] * fn;
            else { const double fm = f0[idx - 1]; jm = -df[NV + idx] * (fn - fm) - dn[NV + idx] * (fn + fm); }
            const double L1 = (jp - jm) / hx;
            // y
            if (fl & FLAG_YP) jp = (G.bc[2] == 2) ? 0.0 : df[2 * NV + idx] * fn - dn[2 * NV + idx] * fn;
            else { const double fp = f0[idx + yp]; jp = -df[2 * NV + idx] * (fp - fn) - dn[2 * NV + idx] * (fp + fn); }
            if (fl & FLAG_YM) jm = (G.bc[3] == 2) ? 0.0 : -df[3 * NV + idx] * fn - dn[3 * NV + idx] * fn;
            else { const double fm = f0[idx - ym]; jm = -df[3 * NV + idx] * (fn - fm) - dn[3 * NV + idx] * (fn + fm); }
            const double L2 = (jp - jm) / hy;
            // z
            if (kg == G.nzz - 1) jp = (G.bc[5] == 2) ? 0.0 : df[4 * NV + idx] * fn - dn[4 * NV + idx] * fn;
            else { const double fp = f0[idx + np]; jp = -df[4 * NV + idx] * (fp - fn) - dn[4 * NV + idx] * (fp + fn); }
            if (kg == 0) jm = (G.bc[4] == 2) ? 0.0 : -df[5 * NV + idx] * fn - dn[5 * NV + idx] * fn;
            else { const double fm = f0[idx - np]; jm = -df[5 * NV + idx] * (fn - fm) - dn[5 * NV + idx] * (fn + fm); }
            const double L3 = (jp - jm) / hz;
            if (Ltot) {            // reactivity (mod_trans.f90:677-678): L(n,g) = L1 + L2 + L3
                Ltot[(size_t)g * NV + idx] = L1 + L2 + L3;
                continue;
            }
            double *S1 = A.S + ((size_t)0 * A.ng + g) * NV, *S2 = A.S + ((size_t)1 * A.ng + g) * NV,
                   *S3 = A.S + ((size_t)2 * A.ng + g) * NV;
            if (A.cmode == 2) {
                const double ex = A.exsrc[(size_t)g * NV + idx];
                S1[idx] = L2 + L3 - ex; S2[idx] = L1 + L3 - ex; S3[idx] = L1 + L2 - ex;
            } else {
                S1[idx] = L2 + L3; S2[idx] = L1 + L3; S3[idx] = L1 + L2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// geometry of the line through a node in direction u
// ---------------------------------------------------------------------------------------
struct Line {
    bool has_m, has_p;
    long long off_m, off_p;  // index distance to the m / p neighbour
    double h, hm, hp;        // node size and the neighbours' sizes in direction u
    int bcm, bcp;            // boundary codes at the line's ends
};

__device__ __forceinline__ Line line_of(const Geo &G, int u, int kl, int r)
{
    Line q;
    if (u == 0) {
        const unsigned f = G.flag[r];
        q.has_m = !(f & FLAG_XM); q.has_p = !(f & FLAG_XP);
        q.off_m = 1; q.off_p = 1;
        q.h = G.hx[r]; q.hm = q.has_m ? G.hx[r - 1] : 0.0; q.hp = q.has_p ? G.hx[r + 1] : 0.0;
        q.bcm = G.bc[1]; q.bcp = G.bc[0];
    } else if (u == 1) {
        const unsigned f = G.flag[r];
        q.has_m = !(f & FLAG_YM); q.has_p = !(f & FLAG_YP);
        q.off_m = G.ypm[r]; q.off_p = G.ypp[r];
        q.h = G.hy[r]; q.hm = q.has_m ? G.hy[r - q.off_m] : 0.0; q.hp = q.has_p ? G.hy[r + q.off_p] : 0.0;
        q.bcm = G.bc[3]; q.bcp = G.bc[2];
    } else {
        const int kg = G.k0 + kl;
        q.has_m = kg > 0; q.has_p = kg < G.nzz - 1;
        q.off_m = G.np; q.off_p = G.np;
        q.h = G.hz[1 + kg]; q.hm = G.hz[kg]; q.hp = G.hz[2 + kg];
        q.bcm = G.bc[4]; q.bcp = G.bc[5];
    }
    return q;
}

// LU_solve (mod_nodal.f90:829-897): Doolittle without pivoting, operation order kept.
// Returns false if |mat(i,i)| < 1e-4 for an original diagonal element.
template <int M>
__device__ __forceinline__ bool lu_solve(double (&U)[M][M], const double (&b)[M], double (&x)[M])
{
    bool ok = true;
#pragma unroll
    for (int i = 0; i < M; ++i) ok = ok && !(fabs(U[i][i]) < (double)10e-5f);
    // decomposition; multipliers are kept in the (otherwise zeroed) lower triangle
#pragma unroll
    for (int i = 0; i < M; ++i) {
#pragma unroll
        for (int j = i + 1; j < M; ++j) {
            const double piv = U[j][i] / U[i][i];
#pragma unroll
            for (int k = i + 1; k < M; ++k) U[j][k] = U[j][k] - piv * U[i][k];
            U[j][i] = piv;
        }
    }
    double y[M];
    y[0] = b[0];
#pragma unroll
    for (int i = 1; i < M; ++i) {
        double isum = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) isum = isum + U[i][k] * y[k];
        y[i] = b[i] - isum;
    }
    x[M - 1] = y[M - 1] / U[M - 1][M - 1];
#pragma unroll
    for (int i = M - 2; i >= 0; --i) {
        double isum = 0.0;
#pragma unroll
        for (int k = i + 1; k < M; ++k) isum = isum + U[i][k] * x[k];
        x[i] = (y[i] - isum) / U[i][i];
    }
    return ok;
}

// everything the sweep carries for one (node, direction)
template <int NG>
struct NodeDir {
    double Bc[NG][NG];
    double A[NG], B[NG], E[NG], F[NG], Gc[NG], H[NG];
    double a2[NG], a4[NG], L1[NG];
    double f0[NG], D[NG];
};

// get_B + get_ABEFGH + TLUpd1/2 + get_a2matvec + LU + get_a4 for node idx, direction u
// (mod_nodal.f90:1345-1451, 1047-1341, 769-825, 740-765)
template <int NG, int KERN>
__device__ __forceinline__ bool node_dir(const Geo &G, const NodalArgs &A, int u, long long idx, const Line &q,
                                         NodeDir<NG> &nd)
{
    const long long NV = G.NV;
    const int m = A.mat[idx] - 1;
    const double Ke = A.scal[S_KE];
    const double hh = q.h * q.h;
    double nuf[NG], chi[NG], sigr[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        nd.f0[g] = A.f0[g][idx];
        nd.D[g] = A.D[(size_t)g * NV + idx];
        nuf[g] = A.nuf[(size_t)g * NV + idx];
        sigr[g] = A.sigr[(size_t)g * NV + idx];
        chi[g] = A.chi[g * A.nmat + m];
    }
    double tfac = 0.0;
    if (A.cmode == 2) tfac = 1.0 - A.tbeta[m] + A.dfis[idx];
    // ---- get_B
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            double dum;
            if (A.cmode == 1) {
                if (g == h) dum = sigr[g] - chi[g] * nuf[h] / Ke;
                else dum = -A.sigs[((size_t)g * NG + h) * NV + idx] - chi[g] * nuf[h] / Ke;   // sigs(n,h,g)
            } else if (A.cmode == 2) {
                if (g == h) dum = sigr[g] - tfac * chi[g] * chi[g] * nuf[h];                    // sic, :1383-1384
                else dum = -A.sigs[((size_t)g * NG + h) * NV + idx] - tfac * chi[g] * nuf[h];
            } else {
                if (g == h) dum = sigr[g] - chi[g] * nuf[h] / Ke;
                else dum = -A.sigs[((size_t)h * NG + g) * NV + idx] - chi[h] * nuf[g] / Ke;   // sigs(n,g,h)
            }
            nd.Bc[g][h] = 0.25 * hh / nd.D[g] * dum;
        }
    }
    // ---- get_ABEFGH (SANM) or the PNM constants (mod_nodal.f90:180-182): B and E here, A, F, G, H where the
    //      surface is solved (load_afgh) -- they are not part of the stored node-direction record
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (KERN == ADP_KERN_SANM) {
            const double *cc = A.abefgh + ((size_t)u * 6 * NG + g) * NV + idx;   // [u][c][g][NV]
            nd.B[g] = cc[(size_t)1 * NG * NV];
            nd.E[g] = cc[(size_t)2 * NG * NV];
        } else {
            nd.B[g] = 1.0 / 35.0; nd.E[g] = 2.0 / 7.0;
        }
    }
    // ---- transverse leakage moments (TLUpd1 / TLUpd2) and the a2 system
    double M2[NG][NG], b[NG], Lm2[NG];
    const double *Su = A.S + (size_t)u * NG * NV;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double Sn = Su[(size_t)g * NV + idx];
        const double Sp = q.has_p ? Su[(size_t)g * NV + idx + q.off_p] : 0.0;
        const double Sm = q.has_m ? Su[(size_t)g * NV + idx - q.off_m] : 0.0;
        double l1, l2, tm, tp, p1m, p2m, p1p, p2p, hp;
        if (!q.has_m) {
            if (q.bcm == 2) {
                tm = 1.0; tp = q.hp / q.h;
                p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0;
                hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
                l1 = (p1m * p2m * (Sp - Sn)) / hp;
                l2 = (p1m * (Sp - Sn)) / hp;
            } else {
                tp = q.hp / q.h; p1p = tp + 1.0;
                l1 = (Sp - Sn) / p1p;
                l2 = 0.0;
            }
        } else if (!q.has_p) {
            if (q.bcp == 2) {
                tm = q.hm / q.h; tp = 1.0;
                p1m = tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
                hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
                l1 = (p1p * p2p * (Sn - Sm)) / hp;
                l2 = (p1p * (Sm - Sn)) / hp;
            } else {
                tm = q.hm / q.h; p1m = tm + 1.0;
                l1 = (Sn - Sm) / p1m;
                l2 = 0.0;
            }
        } else {
            tm = q.hm / q.h; tp = q.hp / q.h;
            p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
            hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
            l1 = (p1m * p2m * (Sp - Sn) + p1p * p2p * (Sn - Sm)) / hp;
            l2 = (p1m * (Sp - Sn) + p1p * (Sm - Sn)) / hp;
        }
        nd.L1[g] = 0.25 * hh / nd.D[g] * l1;
        Lm2[g] = 0.25 * hh / nd.D[g] * l2;
        // get_a2matvec
        double S;
        if (A.cmode == 2) S = 0.25 * hh / nd.D[g] * Sn;
        else S = 0.25 * hh / nd.D[g] * (Sn - A.exsrc[(size_t)g * NV + idx]);
        double Bf = 0.0;
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            M2[g][h] = (h == g) ? nd.Bc[g][h] * nd.E[g] + 3.0 : nd.Bc[g][h] * nd.E[g];
            Bf = Bf + nd.Bc[g][h] * nd.f0[h];
        }
        b[g] = Bf - nd.E[g] * Lm2[g] + S;
    }
    const bool ok = lu_solve<NG>(M2, b, nd.a2);
    // get_a4
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        double Bf = 0.0;
#pragma unroll
        for (int h = 0; h < NG; ++h) Bf = Bf + nd.Bc[g][h] * nd.a2[h];
        nd.a4[g] = nd.B[g] * (Bf + Lm2[g]);
    }
    return ok;
}

// get_a3 (mod_nodal.f90:702-736)
template <int NG>
__device__ __forceinline__ void get_a3(const NodeDir<NG> &nd, const double (&a1)[NG], double (&a3)[NG])
{
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        double Bf = 0.0;
#pragma unroll
        for (int h = 0; h < NG; ++h) Bf = Bf + nd.Bc[g][h] * a1[h];
        a3[g] = nd.A[g] * (Bf + nd.L1[g]);
    }
}


// get_ABEFGH (mod_nodal.f90:1409-1451) for every direction and group of one node
template <int NG>
__global__ void __launch_bounds__(ADP_TILE, 3) k_nodal_abefgh(Geo G, const double *__restrict__ D, const double *__restrict__ sigr,
                                                              double *__restrict__ out, int klo, int npl)
{
    const long long NV = G.NV;
    FOR_EACH_ROW(G, klo, npl)
    {
        const long long idx = node_idx(G, kl, r);
        const double hu[3] = {G.hx[r], G.hy[r], G.hz[1 + G.k0 + kl]};
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const double ratio = sigr[(size_t)g * NV + idx] / D[(size_t)g * NV + idx];
            double c[3][6];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                bool reuse = false;
#pragma unroll
                for (int w = 0; w < u; ++w)
                    if (!reuse && hu[w] == hu[u]) {   // same inputs -> same bits; skip the transcendental work
#pragma unroll
                        for (int k = 0; k < 6; ++k) c[u][k] = c[w][k];
                        reuse = true;
                    }
                if (reuse) continue;
                const double alp = 0.5 * sqrt(ratio) * hu[u];
                const double alp2 = alp * alp;
                const double sh = sinh(alp), ch = cosh(alp);
                const double m0c = sh / alp;
                const double m1s = 3.0 * (ch / alp - sh / alp2);
                const double m2c = 5.0 * (sh / alp - 3.0 * ch / alp2 + 3.0 * sh / (alp * alp * alp));
                c[u][0] = (sh - m1s) / (alp2 * m1s);
                c[u][1] = (ch - m0c - m2c) / (alp2 * m2c);
                c[u][2] = (m0c / m2c - 3.0 / alp2);
                c[u][3] = (alp * ch - m1s) / (alp2 * m1s);
                c[u][4] = (alp * sh - 3.0 * m2c) / (ch - m0c - m2c);
                c[u][5] = (alp * ch - m1s) / (sh - m1s);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int k = 0; k < 6; ++k) out[(((size_t)u * 6 + k) * NG + g) * NV + idx] = c[u][k];
        }
    }
}

// A, F, G, H of get_ABEFGH for (node idx, direction u): from the SANM constants cache, or the PNM values
template <int NG, int KERN>
__device__ __forceinline__ void load_afgh(const NodalArgs &A, int u, long long NV, long long idx, NodeDir<NG> &nd)
{
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (KERN == ADP_KERN_SANM) {
            const double *cc = A.abefgh + ((size_t)u * 6 * NG + g) * NV + idx;   // [u][c][g][NV]
            nd.A[g] = cc[0];
            nd.F[g] = cc[(size_t)3 * NG * NV];
            nd.Gc[g] = cc[(size_t)4 * NG * NV];
            nd.H[g] = cc[(size_t)5 * NG * NV];
        } else {
            nd.A[g] = 1.0 / 15.0; nd.F[g] = 2.0 / 5.0; nd.Gc[g] = 10.0; nd.H[g] = 6.0;
        }
    }
}

// store layout of one (direction, node): slots [0,NG*NG) Bc(g,h); then a2, a4, L1 (NG each).  Round 1 also kept
// A, F, G, H in the record: 4G doubles written and read back per (node, direction) although the constants cache
// (SANM) or four literals (PNM) already hold them -- 16 of the ~79 doubles per node-direction the update moved at G = 2.
#define ND_SLOTS(NG) ((NG) * (NG) + 3 * (NG))
template <int NG>
__device__ __forceinline__ void nd_store(const NodeDir<NG> &nd, double *__restrict__ base, long long NV, long long idx)
{
    int sl = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int h = 0; h < NG; ++h) base[(size_t)(sl++) * NV + idx] = nd.Bc[g][h];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        base[(size_t)(sl + 0 * NG + g) * NV + idx] = nd.a2[g];
        base[(size_t)(sl + 1 * NG + g) * NV + idx] = nd.a4[g];
        base[(size_t)(sl + 2 * NG + g) * NV + idx] = nd.L1[g];
    }
}
template <int NG, int KERN>
__device__ __forceinline__ void nd_load(NodeDir<NG> &nd, const double *__restrict__ base, long long NV, long long idx,
                                        const NodalArgs &A, int u)
{
    int sl = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int h = 0; h < NG; ++h) nd.Bc[g][h] = base[(size_t)(sl++) * NV + idx];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        nd.a2[g] = base[(size_t)(sl + 0 * NG + g) * NV + idx];
        nd.a4[g] = base[(size_t)(sl + 1 * NG + g) * NV + idx];
        nd.L1[g] = base[(size_t)(sl + 2 * NG + g) * NV + idx];
        nd.f0[g] = A.f0[g][idx];
        nd.D[g] = A.D[(size_t)g * NV + idx];
    }
    load_afgh<NG, KERN>(A, u, NV, idx, nd);
}

// one thread per (node, direction): compute and store what the sweep carries for that node
template <int NG, int KERN>
__global__ void __launch_bounds__(ADP_TILE, (NG <= 2) ? NODAL_LB_ND : 1) k_nodal_nodedir(Geo G, NodalArgs A, int u, int klo, int npl)
{
    bool ok = true;
    FOR_EACH_ROW(G, klo, npl)
    {
        const long long idx = node_idx(G, kl, r);
        const Line q = line_of(G, u, kl, r);
        NodeDir<NG> nd;
        ok = node_dir<NG, KERN>(G, A, u, idx, q, nd) && ok;
        nd_store<NG>(nd, A.nd + (size_t)u * ND_SLOTS(NG) * G.NV, G.NV, idx);
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
}

struct ArgMax {
    double *scal;
    double *part;              // [ADP_MAXPART]
    long long *part_loc;       // [ADP_MAXPART]
    unsigned int *ticket;
    long long *loc_out;
};

__device__ __forceinline__ void argmax_combine(double &v, long long &l, double v2, long long l2)
{
    if (v2 > v || (v2 == v && l2 < l)) { v = v2; l = l2; }
}

template <int TPB = ADP_TILE>
__device__ __forceinline__ void grid_argmax(double v, long long l, const ArgMax &ro)
{
    __shared__ double smv[TPB / 32];
    __shared__ long long sml[TPB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_down_sync(0xffffffffu, v, o);
        const long long l2 = __shfl_down_sync(0xffffffffu, l, o);
        argmax_combine(v, l, v2, l2);
    }
    if (lane == 0) { smv[wid] = v; sml[wid] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TPB / 32; ++w) argmax_combine(v, l, smv[w], sml[w]);
        ro.part[blockIdx.x] = v;
        ro.part_loc[blockIdx.x] = l;
        __threadfence();
        const unsigned int t = atomicAdd(ro.ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double bv = -1.0;
    long long bl = 0x7fffffffffffffffLL;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
        argmax_combine(bv, bl, __ldcg(&ro.part[b]), __ldcg(&ro.part_loc[b]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_down_sync(0xffffffffu, bv, o);
        const long long l2 = __shfl_down_sync(0xffffffffu, bl, o);
        argmax_combine(bv, bl, v2, l2);
    }
    __syncthreads();
    if (lane == 0) { smv[wid] = bv; sml[wid] = bl; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TPB / 32; ++w) argmax_combine(bv, bl, smv[w], sml[w]);
        // strict ">" against the running maximum of the directions already swept
        // (x before y before z, as in the reference's sweep order)
        if (bv > ro.scal[S_NDMAX]) { ro.scal[S_NDMAX] = bv; *ro.loc_out = bl; }
        *ro.ticket = 0u;
    }
}

// ---------------------------------------------------------------------------------------
// one thread per (node, direction u): surface on the "+" side (+ the "-" boundary surface
// of the first node of a line).  get_coefs / get_coefs_first / get_coefs_last +
// nodal_coup_upd (mod_nodal.f90:282-698).
// ---------------------------------------------------------------------------------------
template <int NG, int KERN>
__global__ void __launch_bounds__(ADP_TILE, (NG <= 2) ? NODAL_LB_SF : 1) k_nodal_surfaces(Geo G, NodalArgs A, int u, int klo, int npl, ArgMax am)
{
    const long long NV = G.NV;
    const double *ndbase = A.nd + (size_t)u * ND_SLOTS(NG) * NV;
    double best = -1.0;
    long long best_loc = 0x7fffffffffffffffLL;
    bool ok = true;
    FOR_EACH_ROW(G, klo, npl)
    {
        const long long idx = node_idx(G, kl, r);
        const bool owned = (kl >= 0 && kl < G.nzl);
        const long long gnode = (long long)(G.k0 + kl) * G.np + r;   // global node number - 1
        const Line qn = line_of(G, u, kl, r);
        const int sf = 2 * u;                                         // 0-based "+" face; "-" face = sf + 1
        NodeDir<NG> n;
        nd_load<NG, KERN>(n, ndbase, NV, idx, A, u);
        double a1[NG], a3[NG];

        if (!qn.has_m && owned) {
            // ---- first node of the line: one-node problem on its "-" face (get_a1matvec_first)
            double M1[NG][NG], b[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double Pp = 2.0 * n.D[g] / qn.h;
                const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idx];
                if (qn.bcm == 2) {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? Pp * (n.Bc[g][h] * n.F[g] + 1.0) : Pp * n.Bc[g][h] * n.F[g];
                    b[g] = Pp * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] - n.F[g] * n.L1[g]);
                } else if (qn.bcm == 1) {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? -dcp * (1.0 + n.A[g] * n.Bc[g][h]) - 2.0 * Pp * (n.A[g] * n.Bc[g][h] * n.H[g] + 1.0)
                                            : -dcp * n.A[g] * n.Bc[g][h] - 2.0 * Pp * n.A[g] * n.Bc[g][h] * n.H[g];
                    b[g] = 2.0 * Pp * (n.A[g] * n.H[g] * n.L1[g] - 3.0 * n.a2[g] - n.Gc[g] * n.a4[g]) -
                           dcp * (n.a2[g] + n.a4[g] + n.f0[g] - n.A[g] * n.L1[g]);
                } else {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? dcp * (1.0 + n.A[g] * n.Bc[g][h]) : dcp * n.A[g] * n.Bc[g][h];
                    b[g] = dcp * (n.a2[g] + n.a4[g] + n.f0[g] - n.A[g] * n.L1[g]);
                }
            }
            ok = lu_solve<NG>(M1, b, a1) && ok;
            get_a3<NG>(n, a1, a3);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double jp = -2.0 * n.D[g] / qn.h * (a1[g] - 3.0 * n.a2[g] + n.H[g] * a3[g] - n.Gc[g] * n.a4[g]);
                double *dn = A.dn + ((size_t)g * 6 + sf + 1) * NV;
                const double dfm = A.df[((size_t)g * 6 + sf + 1) * NV + idx];
                const double ndpr = dn[idx];
                const double nw = -(jp / n.f0[g] + dfm);
                dn[idx] = nw;
                const double nder = fabs(nw - ndpr);
                if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
            }
        }

        if (qn.has_p) {
            // ---- interior surface between n = this node and p = its "+" neighbour (get_coefs)
            const long long idp = idx + qn.off_p;
            const int klp = (u == 2) ? kl + 1 : kl;
            const int rp = (u == 0) ? r + 1 : (u == 1) ? r + (int)qn.off_p : r;
            const Line qp = line_of(G, u, klp, rp);
            NodeDir<NG> p;
            nd_load<NG, KERN>(p, ndbase, NV, idp, A, u);
            double R[2 * NG][2 * NG], s[2 * NG], sx[2 * NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double Pn = 2.0 * n.D[g] / qn.h, Pp = 2.0 * p.D[g] / qp.h;
#pragma unroll
                for (int h = 0; h < NG; ++h) {
                    if (h == g) {
                        R[g][g] = -Pn * (n.Bc[g][h] * n.F[g] + 1.0);
                        R[g][g + NG] = Pp * (p.Bc[g][h] * p.F[g] + 1.0);
                    } else {
                        R[g][h] = -Pn * n.Bc[g][h] * n.F[g];
                        R[g][h + NG] = Pp * p.Bc[g][h] * p.F[g];
                    }
                }
                s[g] = Pn * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] + n.F[g] * n.L1[g]) +
                       Pp * (3.0 * p.a2[g] + p.Gc[g] * p.a4[g] - p.F[g] * p.L1[g]);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
                const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idp];
#pragma unroll
                for (int h = 0; h < NG; ++h) {
                    if (h == g) {
                        R[g + NG][g] = dcn * (n.Bc[g][h] * n.A[g] + 1.0);
                        R[g + NG][g + NG] = dcp * (p.Bc[g][h] * p.A[g] + 1.0);
                    } else {
                        R[g + NG][h] = dcn * n.Bc[g][h] * n.A[g];
                        R[g + NG][h + NG] = dcp * p.Bc[g][h] * p.A[g];
                    }
                }
                // ADF cross terms exactly as mod_nodal.f90:693-694 (An*Ln1 with dc_p, Ap*Lp1 with dc_n)
                s[g + NG] = dcp * (p.a2[g] + p.a4[g] + p.f0[g] - n.A[g] * n.L1[g]) -
                            dcn * (n.a2[g] + n.a4[g] + n.f0[g] + p.A[g] * p.L1[g]);
            }
            ok = lu_solve<2 * NG>(R, s, sx) && ok;
#pragma unroll
            for (int g = 0; g < NG; ++g) a1[g] = sx[g];
            get_a3<NG>(n, a1, a3);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double jp = -2.0 * n.D[g] / qn.h * (a1[g] + 3.0 * n.a2[g] + n.H[g] * a3[g] + n.Gc[g] * n.a4[g]);
                double *dn = A.dn + ((size_t)g * 6 + sf) * NV;
                const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
                const double ndpr = dn[idx];
                const double nw = (dfp * (n.f0[g] - p.f0[g]) - jp) / (n.f0[g] + p.f0[g]);
                dn[idx] = nw;
                A.dn[((size_t)g * 6 + sf + 1) * NV + idp] = nw;
                if (owned) {
                    const double nder = fabs(nw - ndpr);
                    if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
                }
            }
        } else if (owned) {
            // ---- last node of the line: one-node problem on its "+" face (get_a1matvec_last)
            double M1[NG][NG], b[NG];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double Pn = 2.0 * n.D[g] / qn.h;
                const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
                if (qn.bcp == 2) {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? -Pn * (n.Bc[g][h] * n.F[g] + 1.0) : -Pn * n.Bc[g][h] * n.F[g];
                    b[g] = Pn * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] + n.F[g] * n.L1[g]);
                } else if (qn.bcp == 1) {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? dcn * (1.0 + n.A[g] * n.Bc[g][h]) + 2.0 * Pn * (n.A[g] * n.Bc[g][h] * n.H[g] + 1.0)
                                            : dcn * n.A[g] * n.Bc[g][h] + 2.0 * Pn * n.A[g] * n.Bc[g][h] * n.H[g];
                    b[g] = -2.0 * Pn * (n.A[g] * n.H[g] * n.L1[g] + 3.0 * n.a2[g] + n.Gc[g] * n.a4[g]) -
                           dcn * (n.a2[g] + n.a4[g] + n.f0[g] + n.A[g] * n.L1[g]);
                } else {
#pragma unroll
                    for (int h = 0; h < NG; ++h)
                        M1[g][h] = (h == g) ? dcn * (1.0 + n.A[g] * n.Bc[g][h]) : dcn * n.A[g] * n.Bc[g][h];
                    b[g] = -dcn * (n.a2[g] + n.a4[g] + n.f0[g] + n.A[g] * n.L1[g]);
                }
            }
            ok = lu_solve<NG>(M1, b, a1) && ok;
            get_a3<NG>(n, a1, a3);
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const double jp = -2.0 * n.D[g] / qn.h * (a1[g] + 3.0 * n.a2[g] + n.H[g] * a3[g] + n.Gc[g] * n.a4[g]);
                double *dn = A.dn + ((size_t)g * 6 + sf) * NV;
                const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
                const double ndpr = dn[idx];
                const double nw = -(jp / n.f0[g] - dfp);
                dn[idx] = nw;
                const double nder = fabs(nw - ndpr);
                if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
            }
        }
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax(best, best_loc, am);
}


// =======================================================================================
// Fused per-direction kernels (G <= 4): node-direction record and surface solve in ONE kernel, the record
// carried in registers the way the reference's sweep carries a2p, a4p, Lp1, Bcp, Ap..Hp from one
// interface to the next (mod_nodal.f90:53-125) -- the [3][G*G+7G][NV] node-direction store of the two-kernel
// form (written once, read twice: 3.1x the algorithmic traffic at G = 2, ncu round 1) is gone.
//   z  one thread per (plane position, chunk of planes) marches up k: at every step it builds the record of the
//      next node p and solves the surface (n, p); loads are coalesced across plane positions.
//   y  one thread per (plane, i, chunk of rows) marches along j; the lanes of a warp are consecutive i of the
//      same row, so the loads are coalesced here too; a lane outside the jagged outline idles for that row.
//   x  the neighbour is the next thread: every thread builds the record of its own node, the "+" neighbour's
//      record comes through shared memory (tiles overlap by one node).
// A chunk recomputes the record of its first node (1 / chunk length of extra work).  Every surface sees
// exactly the operations of the two-kernel form, so the results are bit-identical to it.
// =======================================================================================
#define NODAL_TPB 128
#define LLMAX 0x7fffffffffffffffLL

// one-node problem on the "-" face of the first node of a line (get_a1matvec_first + get_a3 + nodal_coup_upd):
// nw = new dn of face sf+1 of node idx
template <int NG>
__device__ __forceinline__ bool solve_first(const NodalArgs &A, const NodeDir<NG> &n, double h, int bcm, int sf, long long NV,
                                            long long idx, double (&nw)[NG])
{
    double M1[NG][NG], b[NG], a1[NG], a3[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double Pp = 2.0 * n.D[g] / h;
        const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idx];
        if (bcm == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? Pp * (n.Bc[g][hh] * n.F[g] + 1.0) : Pp * n.Bc[g][hh] * n.F[g];
            b[g] = Pp * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] - n.F[g] * n.L1[g]);
        } else if (bcm == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? -dcp * (1.0 + n.A[g] * n.Bc[g][hh]) - 2.0 * Pp * (n.A[g] * n.Bc[g][hh] * n.H[g] + 1.0)
                                      : -dcp * n.A[g] * n.Bc[g][hh] - 2.0 * Pp * n.A[g] * n.Bc[g][hh] * n.H[g];
            b[g] = 2.0 * Pp * (n.A[g] * n.H[g] * n.L1[g] - 3.0 * n.a2[g] - n.Gc[g] * n.a4[g]) -
                   dcp * (n.a2[g] + n.a4[g] + n.f0[g] - n.A[g] * n.L1[g]);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? dcp * (1.0 + n.A[g] * n.Bc[g][hh]) : dcp * n.A[g] * n.Bc[g][hh];
            b[g] = dcp * (n.a2[g] + n.a4[g] + n.f0[g] - n.A[g] * n.L1[g]);
        }
    }
    const bool ok = lu_solve<NG>(M1, b, a1);
    get_a3<NG>(n, a1, a3);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double jp = -2.0 * n.D[g] / h * (a1[g] - 3.0 * n.a2[g] + n.H[g] * a3[g] - n.Gc[g] * n.a4[g]);
        const double dfm = A.df[((size_t)g * 6 + sf + 1) * NV + idx];
        nw[g] = -(jp / n.f0[g] + dfm);
    }
    return ok;
}

// one-node problem on the "+" face of the last node of a line (get_a1matvec_last): nw = new dn of face sf of node idx
template <int NG>
__device__ __forceinline__ bool solve_last(const NodalArgs &A, const NodeDir<NG> &n, double h, int bcp, int sf, long long NV,
                                           long long idx, double (&nw)[NG])
{
    double M1[NG][NG], b[NG], a1[NG], a3[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double Pn = 2.0 * n.D[g] / h;
        const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
        if (bcp == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? -Pn * (n.Bc[g][hh] * n.F[g] + 1.0) : -Pn * n.Bc[g][hh] * n.F[g];
            b[g] = Pn * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] + n.F[g] * n.L1[g]);
        } else if (bcp == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? dcn * (1.0 + n.A[g] * n.Bc[g][hh]) + 2.0 * Pn * (n.A[g] * n.Bc[g][hh] * n.H[g] + 1.0)
                                      : dcn * n.A[g] * n.Bc[g][hh] + 2.0 * Pn * n.A[g] * n.Bc[g][hh] * n.H[g];
            b[g] = -2.0 * Pn * (n.A[g] * n.H[g] * n.L1[g] + 3.0 * n.a2[g] + n.Gc[g] * n.a4[g]) -
                   dcn * (n.a2[g] + n.a4[g] + n.f0[g] + n.A[g] * n.L1[g]);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                M1[g][hh] = (hh == g) ? dcn * (1.0 + n.A[g] * n.Bc[g][hh]) : dcn * n.A[g] * n.Bc[g][hh];
            b[g] = -dcn * (n.a2[g] + n.a4[g] + n.f0[g] + n.A[g] * n.L1[g]);
        }
    }
    const bool ok = lu_solve<NG>(M1, b, a1);
    get_a3<NG>(n, a1, a3);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double jp = -2.0 * n.D[g] / h * (a1[g] + 3.0 * n.a2[g] + n.H[g] * a3[g] + n.Gc[g] * n.a4[g]);
        const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
        nw[g] = -(jp / n.f0[g] - dfp);
    }
    return ok;
}

// two-node problem between n (node idx, size hn) and its "+" neighbour p (node idp, size hp): get_a1matvec + LU 2G x 2G +
// get_a3 + nodal_coup_upd; nw = new dn of face sf of n = face sf+1 of p
template <int NG>
__device__ __forceinline__ bool solve_inner(const NodalArgs &A, const NodeDir<NG> &n, const NodeDir<NG> &p, double hn, double hp,
                                            int sf, long long NV, long long idx, long long idp, double (&nw)[NG])
{
    double R[2 * NG][2 * NG], s[2 * NG], sx[2 * NG], a1[NG], a3[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double Pn = 2.0 * n.D[g] / hn, Pp = 2.0 * p.D[g] / hp;
#pragma unroll
        for (int hh = 0; hh < NG; ++hh) {
            if (hh == g) {
                R[g][g] = -Pn * (n.Bc[g][hh] * n.F[g] + 1.0);
                R[g][g + NG] = Pp * (p.Bc[g][hh] * p.F[g] + 1.0);
            } else {
                R[g][hh] = -Pn * n.Bc[g][hh] * n.F[g];
                R[g][hh + NG] = Pp * p.Bc[g][hh] * p.F[g];
            }
        }
        s[g] = Pn * (3.0 * n.a2[g] + n.Gc[g] * n.a4[g] + n.F[g] * n.L1[g]) +
               Pp * (3.0 * p.a2[g] + p.Gc[g] * p.a4[g] - p.F[g] * p.L1[g]);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
        const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idp];
#pragma unroll
        for (int hh = 0; hh < NG; ++hh) {
            if (hh == g) {
                R[g + NG][g] = dcn * (n.Bc[g][hh] * n.A[g] + 1.0);
                R[g + NG][g + NG] = dcp * (p.Bc[g][hh] * p.A[g] + 1.0);
            } else {
                R[g + NG][hh] = dcn * n.Bc[g][hh] * n.A[g];
                R[g + NG][hh + NG] = dcp * p.Bc[g][hh] * p.A[g];
            }
        }
        // ADF cross terms exactly as mod_nodal.f90:693-694 (An*Ln1 with dc_p, Ap*Lp1 with dc_n)
        s[g + NG] = dcp * (p.a2[g] + p.a4[g] + p.f0[g] - n.A[g] * n.L1[g]) -
                    dcn * (n.a2[g] + n.a4[g] + n.f0[g] + p.A[g] * p.L1[g]);
    }
    const bool ok = lu_solve<2 * NG>(R, s, sx);
#pragma unroll
    for (int g = 0; g < NG; ++g) a1[g] = sx[g];
    get_a3<NG>(n, a1, a3);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double jp = -2.0 * n.D[g] / hn * (a1[g] + 3.0 * n.a2[g] + n.H[g] * a3[g] + n.Gc[g] * n.a4[g]);
        const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
        nw[g] = (dfp * (n.f0[g] - p.f0[g]) - jp) / (n.f0[g] + p.f0[g]);
    }
    return ok;
}

// store the new dn of one face (and of the matching face of the "+" neighbour), track max |delta dn| (nodal_coup_upd)
template <int NG>
__device__ __forceinline__ void apply_dn(const NodalArgs &A, long long NV, long long idx, int face, long long idp, int facep,
                                         const double (&nw)[NG], bool count, long long gnode, double &best, long long &best_loc)
{
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        double *dn = A.dn + ((size_t)g * 6 + face) * NV;
        const double ndpr = dn[idx];
        dn[idx] = nw[g];
        if (idp >= 0) A.dn[((size_t)g * 6 + facep) * NV + idp] = nw[g];
        if (count) {
            const double nder = fabs(nw[g] - ndpr);
            if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
        }
    }
}

// ---- z: march up the planes -----------------------------------------------------------------
template <int NG, int KERN>
__global__ void __launch_bounds__(NODAL_TPB) k_nodal_march_z(Geo G, NodalArgs A, int kfirst, int klast, int nchunk, ArgMax am)
{
    const long long NV = G.NV;
    const int tiles = (G.np + NODAL_TPB - 1) / NODAL_TPB;
    const int ns = klast - kfirst;                               // interior surfaces along a line
    double best = -1.0;
    long long best_loc = LLMAX;
    bool ok = true;
    for (int item = blockIdx.x; item < tiles * nchunk; item += gridDim.x) {
        const int c = item / tiles, r = (item % tiles) * NODAL_TPB + threadIdx.x;
        if (r >= G.np) continue;
        const int s_lo = (int)((long long)ns * c / nchunk), s_hi = (int)((long long)ns * (c + 1) / nchunk);
        int kl = kfirst + s_lo;
        long long idx = node_idx(G, kl, r);
        Line qn = line_of(G, 2, kl, r);
        NodeDir<NG> n;
        double nw[NG];
        ok = node_dir<NG, KERN>(G, A, 2, idx, qn, n) && ok;
        load_afgh<NG, KERN>(A, 2, NV, idx, n);
        if (c == 0 && !qn.has_m && kl >= 0 && kl < G.nzl) {      // bottom face of the core
            ok = solve_first<NG>(A, n, qn.h, qn.bcm, 4, NV, idx, nw) && ok;
            apply_dn<NG>(A, NV, idx, 5, -1, 0, nw, true, (long long)(G.k0 + kl) * G.np + r, best, best_loc);
        }
        for (int s = s_lo; s < s_hi; ++s) {
            const long long idp = idx + G.np;
            const Line qp = line_of(G, 2, kl + 1, r);
            NodeDir<NG> p;
            ok = node_dir<NG, KERN>(G, A, 2, idp, qp, p) && ok;
            load_afgh<NG, KERN>(A, 2, NV, idp, p);
            ok = solve_inner<NG>(A, n, p, qn.h, qp.h, 4, NV, idx, idp, nw) && ok;
            apply_dn<NG>(A, NV, idx, 4, idp, 5, nw, kl >= 0 && kl < G.nzl, (long long)(G.k0 + kl) * G.np + r, best, best_loc);
            n = p; qn = qp; idx = idp; ++kl;
        }
        if (c == nchunk - 1 && !qn.has_p && kl >= 0 && kl < G.nzl) {   // top face of the core
            ok = solve_last<NG>(A, n, qn.h, qn.bcp, 4, NV, idx, nw) && ok;
            apply_dn<NG>(A, NV, idx, 4, -1, 0, nw, true, (long long)(G.k0 + kl) * G.np + r, best, best_loc);
        }
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax<NODAL_TPB>(best, best_loc, am);
}

// ---- y: march along the rows ----------------------------------------------------------------
template <int NG, int KERN>
__global__ void __launch_bounds__(NODAL_TPB) k_nodal_march_y(Geo G, GeoXY X, NodalArgs A, int nchunk, ArgMax am)
{
    const long long NV = G.NV;
    const long long total = (long long)G.nzl * X.nxx;            // threads: (plane, i)
    const int tiles = (int)((total + NODAL_TPB - 1) / NODAL_TPB);
    double best = -1.0;
    long long best_loc = LLMAX;
    bool ok = true;
    for (int item = blockIdx.x; item < tiles * nchunk; item += gridDim.x) {
        const int c = item / tiles;
        const long long t = (long long)(item % tiles) * NODAL_TPB + threadIdx.x;
        if (t >= total) continue;
        const int kl = (int)(t / X.nxx), i = (int)(t % X.nxx);
        // this chunk owns the surfaces whose lower node lies in rows [jc0, jc1); row jc1 is visited for its record only
        const int jc0 = (int)((long long)X.nyy * c / nchunk), jc1 = (int)((long long)X.nyy * (c + 1) / nchunk);
        bool have = false;
        NodeDir<NG> n;
        Line qn;
        long long idx = 0;
        int rn = 0;
        double nw[NG];
        for (int j = jc0; j <= jc1 && j < X.nyy; ++j) {
            const int rp1 = X.nodp[(size_t)j * X.nxx + i];
            if (!rp1) { have = false; continue; }                // outside the outline (the rows of a column are contiguous)
            if (j == jc1 && !have) break;                        // nothing to pair the extra row with
            const int r = rp1 - 1;
            const long long idp = node_idx(G, kl, r);
            const Line qp = line_of(G, 1, kl, r);
            const long long gp = (long long)(G.k0 + kl) * G.np + r;
            NodeDir<NG> p;
            ok = node_dir<NG, KERN>(G, A, 1, idp, qp, p) && ok;
            load_afgh<NG, KERN>(A, 1, NV, idp, p);
            if (!qp.has_m && j < jc1) {
                ok = solve_first<NG>(A, p, qp.h, qp.bcm, 2, NV, idp, nw) && ok;
                apply_dn<NG>(A, NV, idp, 3, -1, 0, nw, true, gp, best, best_loc);
            }
            if (have) {
                ok = solve_inner<NG>(A, n, p, qn.h, qp.h, 2, NV, idx, idp, nw) && ok;
                apply_dn<NG>(A, NV, idx, 2, idp, 3, nw, true, (long long)(G.k0 + kl) * G.np + rn, best, best_loc);
            }
            if (!qp.has_p && j < jc1) {
                ok = solve_last<NG>(A, p, qp.h, qp.bcp, 2, NV, idp, nw) && ok;
                apply_dn<NG>(A, NV, idp, 2, -1, 0, nw, true, gp, best, best_loc);
            }
            n = p; qn = qp; idx = idp; rn = r; have = true;
        }
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax<NODAL_TPB>(best, best_loc, am);
}

// ---- x: the "+" neighbour is the next thread ------------------------------------------------------
// what solve_inner reads of the "+" neighbour: Bc, A, F, Gc, a2, a4, L1, f0, D
template <int NG>
__device__ __forceinline__ void rec_to_smem(const NodeDir<NG> &n, double *sm, int tid)
{
    int sl = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int h = 0; h < NG; ++h) sm[(sl++) * NODAL_TPB + tid] = n.Bc[g][h];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        sm[(sl + 0 * NG + g) * NODAL_TPB + tid] = n.A[g];
        sm[(sl + 1 * NG + g) * NODAL_TPB + tid] = n.F[g];
        sm[(sl + 2 * NG + g) * NODAL_TPB + tid] = n.Gc[g];
        sm[(sl + 3 * NG + g) * NODAL_TPB + tid] = n.a2[g];
        sm[(sl + 4 * NG + g) * NODAL_TPB + tid] = n.a4[g];
        sm[(sl + 5 * NG + g) * NODAL_TPB + tid] = n.L1[g];
        sm[(sl + 6 * NG + g) * NODAL_TPB + tid] = n.f0[g];
        sm[(sl + 7 * NG + g) * NODAL_TPB + tid] = n.D[g];
    }
}
template <int NG>
__device__ __forceinline__ void rec_from_smem(NodeDir<NG> &n, const double *sm, int tid)
{
    int sl = 0;
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
        for (int h = 0; h < NG; ++h) n.Bc[g][h] = sm[(sl++) * NODAL_TPB + tid];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        n.A[g] = sm[(sl + 0 * NG + g) * NODAL_TPB + tid];
        n.F[g] = sm[(sl + 1 * NG + g) * NODAL_TPB + tid];
        n.Gc[g] = sm[(sl + 2 * NG + g) * NODAL_TPB + tid];
        n.a2[g] = sm[(sl + 3 * NG + g) * NODAL_TPB + tid];
        n.a4[g] = sm[(sl + 4 * NG + g) * NODAL_TPB + tid];
        n.L1[g] = sm[(sl + 5 * NG + g) * NODAL_TPB + tid];
        n.f0[g] = sm[(sl + 6 * NG + g) * NODAL_TPB + tid];
        n.D[g] = sm[(sl + 7 * NG + g) * NODAL_TPB + tid];
        n.B[g] = n.E[g] = n.H[g] = 0.0;                           // not used of the "+" neighbour
    }
}

template <int NG, int KERN>
__global__ void __launch_bounds__(NODAL_TPB) k_nodal_pair_x(Geo G, NodalArgs A, ArgMax am)
{
    extern __shared__ double sm_rec[];                           // [NG*NG + 8 NG][NODAL_TPB]
    const long long NV = G.NV;
    const long long total = (long long)G.nzl * G.np;             // flattened (plane, plane position)
    const long long tiles = (total + (NODAL_TPB - 2)) / (NODAL_TPB - 1);   // tiles overlap by one node
    const int tid = threadIdx.x;
    double best = -1.0;
    long long best_loc = LLMAX;
    bool ok = true;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long q = tile * (NODAL_TPB - 1) + tid;
        const bool valid = q < total;
        const int kl = valid ? (int)(q / G.np) : 0, r = valid ? (int)(q % G.np) : 0;
        const long long idx = node_idx(G, kl, r);
        const Line qn = line_of(G, 0, kl, r);
        NodeDir<NG> n;
        if (valid) {
            ok = node_dir<NG, KERN>(G, A, 0, idx, qn, n) && ok;
            load_afgh<NG, KERN>(A, 0, NV, idx, n);
            rec_to_smem<NG>(n, sm_rec, tid);
        }
        __syncthreads();
        if (valid && tid < NODAL_TPB - 1) {                      // the last thread of a tile only lends its record
            const long long gnode = (long long)(G.k0 + kl) * G.np + r;
            double nw[NG];
            if (!qn.has_m) {
                ok = solve_first<NG>(A, n, qn.h, qn.bcm, 0, NV, idx, nw) && ok;
                apply_dn<NG>(A, NV, idx, 1, -1, 0, nw, true, gnode, best, best_loc);
            }
            if (qn.has_p) {
                NodeDir<NG> p;
                rec_from_smem<NG>(p, sm_rec, tid + 1);
                ok = solve_inner<NG>(A, n, p, qn.h, qn.hp, 0, NV, idx, idx + 1, nw) && ok;
                apply_dn<NG>(A, NV, idx, 0, idx + 1, 1, nw, true, gnode, best, best_loc);
            } else {
                ok = solve_last<NG>(A, n, qn.h, qn.bcp, 0, NV, idx, nw) && ok;
                apply_dn<NG>(A, NV, idx, 0, -1, 0, nw, true, gnode, best, best_loc);
            }
        }
        __syncthreads();                                         // the next tile overwrites the records
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax<NODAL_TPB>(best, best_loc, am);
}

// ---------------------------------------------------------------------------------------
// Cooperative surfaces kernel for many groups.  One thread per surface keeps two node-direction
// records and a 2G x 2G system in registers; at G = 8 that is 3.2 KB of local memory per thread
// (ptxas: 17 KB spill stores, 29 KB spill loads) and the three launches took 88 ms on the C2 mesh
// (45 ms here, at 2 CTAs per SM; the kernel is bound by the latency of the dependent shuffle /
// division chain, so occupancy matters: 1 CTA per SM 63 ms, 3 CTAs per SM with spills 63 ms).  A group
// of 16 lanes owns one surface: lane l < 2G holds ROW l of the system (G rows of current
// continuity, G rows of flux continuity), the Doolittle elimination broadcasts the pivot row
// with shuffles, and forward / back substitution pass y(k), x(k) from lane to lane.  Every matrix
// element sees exactly the operations of LU_solve (mod_nodal.f90:829-897) in the same order, so
// the result is bit-identical to the one-thread-per-surface kernel.
// ---------------------------------------------------------------------------------------
#define COOP_W 16
#define COOP_FULL 0xffffffffu
// Both 16-lane groups of a warp always run the same code (a group without work computes on a
// valid dummy item and discards the result), so every shuffle uses the constant full mask; a
// per-group mask costs a MATCH/VOTE validity check per shuffle.
__device__ __forceinline__ double shfl16(double v, int src)
{
    return __shfl_sync(COOP_FULL, v, src, COOP_W);
}

// compile-time loop: the body receives std::integral_constant<int, I>, so every array index is a
// constant and the rows stay in registers ("#pragma unroll" alone left the 120-body elimination
// nest of the 16 x 16 system partially rolled, with the row in local memory)
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F &&f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// rows 0..M-1 of U in lanes 0..M-1 of the 16-lane group (lanes >= M carry dummy rows and only
// take part in the shuffles); b, x: one element per lane.  Returns false on the 1e-4 diagonal abort.
template <int M>
__device__ __forceinline__ bool coop_lu_solve(int sl, double (&row)[M], double b, double &x)
{
    // NOTE: never write "if (sl == i) ... row[i]": the compiler turns that chain into row[sl] and
    // moves the row to local memory.  A lane gets its own diagonal from the broadcast instead.
    double diag = 1.0;
    static_for<0, M>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const double d = shfl16(row[i], i);
        if (sl == i) diag = d;
    });
    const bool ok = !(fabs(diag) < (double)10e-5f);          // original diagonal (mod_nodal.f90:856)
    // decomposition: U(j,k) = U(j,k) - piv U(i,k), piv = U(j,i) / U(i,i) kept in U(j,i)
    static_for<0, M>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const double uii = shfl16(row[i], i);
        if (sl == i) diag = uii;                             // U(i,i) is final from step i on
        const double piv = row[i] / uii;
        static_for<i + 1, M>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const double uik = shfl16(row[k], i);
            if (sl > i) row[k] = row[k] - piv * uik;
        });
        if (sl > i) row[i] = piv;
    });
    // forward substitution: y(i) = b(i) - sum_{k<i} L(i,k) y(k), k ascending
    double isum = 0.0, y = b;
    static_for<0, M>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        if (sl == k && k > 0) y = b - isum;
        const double yk = shfl16(y, k);
        if (sl > k) isum = isum + row[k] * yk;
    });
    // back substitution: x(i) = (y(i) - sum_{k>i} U(i,k) x(k)) / U(i,i), k ascending: lane i adds
    // up only once all x(k > i) have arrived
    double xs[M] = {};
    x = 0.0;
    static_for<0, M>([&](auto tc) {
        constexpr int i = M - 1 - decltype(tc)::value;
        double bs = 0.0;
        static_for<i + 1, M>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            bs = bs + row[k] * xs[k];
        });
        if (sl == i) x = (i == M - 1) ? y / diag : (y - bs) / diag;
        xs[i] = shfl16(x, i);
    });
    return ok;
}

// what one lane keeps of a node-direction record: row g of Bc and the group-g scalars
template <int NG>
struct NodeDirRow {
    double Bc[NG];
    double A, F, Gc, H, a2, a4, L1, f0, D;
};
template <int NG, int KERN>
__device__ __forceinline__ void nd_load_row(NodeDirRow<NG> &nd, int g, const double *__restrict__ base, long long NV,
                                            long long idx, const NodalArgs &A, int u)
{
#pragma unroll
    for (int h = 0; h < NG; ++h) nd.Bc[h] = base[(size_t)(g * NG + h) * NV + idx];
    const double *sc = base + (size_t)(NG * NG + g) * NV + idx;
    nd.a2 = sc[0];
    nd.a4 = sc[(size_t)1 * NG * NV];
    nd.L1 = sc[(size_t)2 * NG * NV];
    if (KERN == ADP_KERN_SANM) {
        const double *cc = A.abefgh + ((size_t)u * 6 * NG + g) * NV + idx;   // [u][c][g][NV]
        nd.A = cc[0];
        nd.F = cc[(size_t)3 * NG * NV];
        nd.Gc = cc[(size_t)4 * NG * NV];
        nd.H = cc[(size_t)5 * NG * NV];
    } else {
        nd.A = 1.0 / 15.0; nd.F = 2.0 / 5.0; nd.Gc = 10.0; nd.H = 6.0;
    }
    nd.f0 = (((A.curmask >> g) & 1u) ? A.f0b : A.f0a)[(size_t)g * NV + idx];
    nd.D = A.D[(size_t)g * NV + idx];
}

// one-node boundary problem (get_a1matvec_first / _last), G rows in lanes 0..G-1, then get_a3,
// the boundary current and the dn update of nodal_coup_upd.  first: "-" face of the first node.
template <int NG>
__device__ __forceinline__ bool coop_boundary(const NodalArgs &A, bool active, int sl, int g, bool first, int bc,
                                              const NodeDirRow<NG> &n, double h, int sf, long long NV, long long idx,
                                              double &nder_out)
{
    const double P = 2.0 * n.D / h;
    const int face = first ? sf + 1 : sf;
    const double dcf = A.dc[((size_t)face * NG + g) * NV + idx];
    double row[NG], b;
    if (first) {
        if (bc == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? P * (n.Bc[hh] * n.F + 1.0) : P * n.Bc[hh] * n.F;
            b = P * (3.0 * n.a2 + n.Gc * n.a4 - n.F * n.L1);
        } else if (bc == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                row[hh] = (hh == g) ? -dcf * (1.0 + n.A * n.Bc[hh]) - 2.0 * P * (n.A * n.Bc[hh] * n.H + 1.0)
                                    : -dcf * n.A * n.Bc[hh] - 2.0 * P * n.A * n.Bc[hh] * n.H;
            b = 2.0 * P * (n.A * n.H * n.L1 - 3.0 * n.a2 - n.Gc * n.a4) - dcf * (n.a2 + n.a4 + n.f0 - n.A * n.L1);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) : dcf * n.A * n.Bc[hh];
            b = dcf * (n.a2 + n.a4 + n.f0 - n.A * n.L1);
        }
    } else {
        if (bc == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? -P * (n.Bc[hh] * n.F + 1.0) : -P * n.Bc[hh] * n.F;
            b = P * (3.0 * n.a2 + n.Gc * n.a4 + n.F * n.L1);
        } else if (bc == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) + 2.0 * P * (n.A * n.Bc[hh] * n.H + 1.0)
                                    : dcf * n.A * n.Bc[hh] + 2.0 * P * n.A * n.Bc[hh] * n.H;
            b = -2.0 * P * (n.A * n.H * n.L1 + 3.0 * n.a2 + n.Gc * n.a4) - dcf * (n.a2 + n.a4 + n.f0 + n.A * n.L1);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) : dcf * n.A * n.Bc[hh];
            b = -dcf * (n.a2 + n.a4 + n.f0 + n.A * n.L1);
        }
    }
    double a1;
    bool ok = coop_lu_solve<NG>(sl, row, b, a1);
    if (sl >= NG || !active) ok = true;            // dummy rows / dummy item
    // get_a3: a3(g) = A(g) (sum_h B(g,h) a1(h) + L1(g))
    double Bf = 0.0;
#pragma unroll
    for (int hh = 0; hh < NG; ++hh) Bf = Bf + n.Bc[hh] * shfl16(a1, hh);
    const double a3 = n.A * (Bf + n.L1);
    nder_out = -1.0;
    if (sl < NG && active) {
        double *dn = A.dn + ((size_t)g * 6 + face) * NV;
        const double dff = A.df[((size_t)g * 6 + face) * NV + idx];
        const double ndpr = dn[idx];
        double nw;
        if (first) {
            const double jp = -2.0 * n.D / h * (a1 - 3.0 * n.a2 + n.H * a3 - n.Gc * n.a4);
            nw = -(jp / n.f0 + dff);
        } else {
            const double jp = -2.0 * n.D / h * (a1 + 3.0 * n.a2 + n.H * a3 + n.Gc * n.a4);
            nw = -(jp / n.f0 - dff);
        }
        dn[idx] = nw;
        nder_out = fabs(nw - ndpr);
    }
    return ok;
}

template <int NG, int KERN>
__global__ void __launch_bounds__(ADP_TILE, 2) k_nodal_surfaces_coop(Geo G, NodalArgs A, int u, int klo, int npl, ArgMax am)
{
    static_assert(2 * NG <= COOP_W, "a 16-lane group holds at most 16 rows");
    const long long NV = G.NV;
    const double *ndbase = A.nd + (size_t)u * ND_SLOTS(NG) * NV;
    const int sl = threadIdx.x & (COOP_W - 1);
    const int grp = threadIdx.x / COOP_W;                      // group within the CTA
    constexpr int GPB = ADP_TILE / COOP_W;                     // groups per CTA
    const int g = sl % NG;                                     // lanes >= 2 NG repeat rows, results unused
    const bool lower = sl < NG;                                // current-continuity rows / the lanes that own group g
    const int sf = 2 * u;
    double best = -1.0;
    long long best_loc = 0x7fffffffffffffffLL;
    bool ok = true;
    const long long total = (long long)npl * G.np;
    // the trip count is the same for every thread of the CTA; a group past the end repeats the last item as a dummy
    for (long long base = (long long)blockIdx.x * GPB; base < total; base += (long long)gridDim.x * GPB) {
        const bool live = base + grp < total;
        const long long item = live ? base + grp : total - 1;
        const int kl = klo + (int)(item / G.np), r = (int)(item % G.np);
        const long long idx = node_idx(G, kl, r);
        const bool owned = (kl >= 0 && kl < G.nzl);
        const long long gnode = (long long)(G.k0 + kl) * G.np + r;
        const Line qn = line_of(G, u, kl, r);
        NodeDirRow<NG> n;
        nd_load_row<NG, KERN>(n, g, ndbase, NV, idx, A, u);
        const bool do_first = live && !qn.has_m && owned;
        const bool do_inner = live && qn.has_p;
        const bool do_last = live && !qn.has_p && owned;
        double nder;
        if (__any_sync(COOP_FULL, do_first)) {
            ok = coop_boundary<NG>(A, do_first, sl, g, true, qn.bcm, n, qn.h, sf, NV, idx, nder) && ok;
            if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
        }
        if (__any_sync(COOP_FULL, do_inner)) {
            // a group without a "+" neighbour pairs the node with itself and discards the result
            const long long idp = do_inner ? idx + qn.off_p : idx;
            const int klp = (do_inner && u == 2) ? kl + 1 : kl;
            const int rp = !do_inner ? r : (u == 0) ? r + 1 : (u == 1) ? r + (int)qn.off_p : r;
            const Line qp = line_of(G, u, klp, rp);
            NodeDirRow<NG> p;
            nd_load_row<NG, KERN>(p, g, ndbase, NV, idp, A, u);
            const double Pn = 2.0 * n.D / qn.h, Pp = 2.0 * p.D / qp.h;
            const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
            const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idp];
            double row[2 * NG], b;
            if (lower) {
#pragma unroll
                for (int h = 0; h < NG; ++h) {
                    row[h] = (h == g) ? -Pn * (n.Bc[h] * n.F + 1.0) : -Pn * n.Bc[h] * n.F;
                    row[h + NG] = (h == g) ? Pp * (p.Bc[h] * p.F + 1.0) : Pp * p.Bc[h] * p.F;
                }
                b = Pn * (3.0 * n.a2 + n.Gc * n.a4 + n.F * n.L1) + Pp * (3.0 * p.a2 + p.Gc * p.a4 - p.F * p.L1);
            } else {
#pragma unroll
                for (int h = 0; h < NG; ++h) {
                    row[h] = (h == g) ? dcn * (n.Bc[h] * n.A + 1.0) : dcn * n.Bc[h] * n.A;
                    row[h + NG] = (h == g) ? dcp * (p.Bc[h] * p.A + 1.0) : dcp * p.Bc[h] * p.A;
                }
                // ADF cross terms exactly as mod_nodal.f90:693-694 (An*Ln1 with dc_p, Ap*Lp1 with dc_n)
                b = dcp * (p.a2 + p.a4 + p.f0 - n.A * n.L1) - dcn * (n.a2 + n.a4 + n.f0 + p.A * p.L1);
            }
            double sx;
            const bool okl = coop_lu_solve<2 * NG>(sl, row, b, sx);
            if (sl < 2 * NG && do_inner) ok = okl && ok;
            double Bf = 0.0;
#pragma unroll
            for (int h = 0; h < NG; ++h) Bf = Bf + n.Bc[h] * shfl16(sx, h);      // a1(h) = sx(h)
            const double a3 = n.A * (Bf + n.L1);
            if (lower && do_inner) {
                const double jp = -2.0 * n.D / qn.h * (sx + 3.0 * n.a2 + n.H * a3 + n.Gc * n.a4);
                double *dn = A.dn + ((size_t)g * 6 + sf) * NV;
                const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
                const double ndpr = dn[idx];
                const double nw = (dfp * (n.f0 - p.f0) - jp) / (n.f0 + p.f0);
                dn[idx] = nw;
                A.dn[((size_t)g * 6 + sf + 1) * NV + idp] = nw;
                if (owned) {
                    const double nd2 = fabs(nw - ndpr);
                    if (nd2 > best || (nd2 == best && gnode < best_loc)) { best = nd2; best_loc = gnode; }
                }
            }
        }
        if (__any_sync(COOP_FULL, do_last)) {
            ok = coop_boundary<NG>(A, do_last, sl, g, false, qn.bcp, n, qn.h, sf, NV, idx, nder) && ok;
            if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
        }
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax(best, best_loc, am);
}

// =======================================================================================
// Quad kernels for many groups (G >= 5, round 2).  FOUR lanes own one (node, direction) / one surface; the rows of
// the G x G (a2) and 2G x 2G (a1) systems are dealt cyclically -- row i lives in lane i mod 4, slot i div 4 -- so at
// G = 8 a lane holds 2 / 4 rows in registers (compile-time indices, no local memory), the Doolittle elimination
// broadcasts the pivot row with 4-lane shuffles and every lane updates its rows: 8 surfaces per warp instead of the 2
// of the 16-lane form, all lanes busy until the last pivots, one division per row and pivot in flight per lane.
// The one-thread-per-item forms spill kilobytes (8 x 8: 10 ms per direction on the C2 mesh, 16 x 16: 29 ms), the
// 16-lane form idles half its lanes (15 ms).  Every element sees LU_solve's operations (mod_nodal.f90:829-897) in
// the same order: bit-identical results.
// =======================================================================================
#ifndef QW
#define QW 4          // lanes per (node, direction) / surface in the quad kernels; -DQW=8 for A/B
#endif
__device__ __forceinline__ double shflq(double v, int src) { return __shfl_sync(COOP_FULL, v, src, QW); }

// rows i = sl + QW * s (s = 0 .. RPL-1) of the M x M system in lane sl; rows >= M are padding (all zero on entry)
template <int M>
__device__ __forceinline__ bool quad_lu_solve(int sl, double (&row)[(M + QW - 1) / QW][M], const double (&b)[(M + QW - 1) / QW],
                                              double (&x)[(M + QW - 1) / QW])
{
    constexpr int RPL = (M + QW - 1) / QW;
    bool ok = true;
    static_for<0, M>([&](auto ic) {                              // original diagonal (mod_nodal.f90:856)
        constexpr int i = decltype(ic)::value;
        const double d = shflq(row[i / QW][i], i % QW);
        ok = ok && !(fabs(d) < (double)10e-5f);
    });
    // decomposition: U(j,k) = U(j,k) - piv U(i,k), piv = U(j,i) / U(i,i) kept in U(j,i)
    static_for<0, M>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int oi = i % QW, si = i / QW;
        const double uii = shflq(row[si][i], oi);
        double piv[RPL];
        static_for<0, RPL>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if constexpr (QW * s + QW - 1 > i) piv[s] = row[s][i] / uii;
        });
        static_for<i + 1, M>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const double uik = shflq(row[si][k], oi);
            static_for<0, RPL>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                if constexpr (QW * s + QW - 1 > i)
                    if (sl + QW * s > i) row[s][k] = row[s][k] - piv[s] * uik;
            });
        });
        static_for<0, RPL>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if constexpr (QW * s + QW - 1 > i)
                if (sl + QW * s > i) row[s][i] = piv[s];
        });
    });
    // forward substitution: y(i) = b(i) - sum_{k<i} L(i,k) y(k), k ascending
    double isum[RPL], y[RPL];
    static_for<0, RPL>([&](auto sc) { isum[decltype(sc)::value] = 0.0; y[decltype(sc)::value] = b[decltype(sc)::value]; });
    static_for<0, M>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        constexpr int ok_ = k % QW, sk = k / QW;
        if (k > 0 && sl == ok_) y[sk] = b[sk] - isum[sk];
        const double yk = shflq(y[sk], ok_);
        static_for<0, RPL>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            if constexpr (QW * s + QW - 1 > k)
                if (sl + QW * s > k) isum[s] = isum[s] + row[s][k] * yk;
        });
    });
    // back substitution: x(i) = (y(i) - sum_{k>i} U(i,k) x(k)) / U(i,i), k ascending
    double xs[M];
    static_for<0, M>([&](auto tc) {
        constexpr int i = M - 1 - decltype(tc)::value;
        constexpr int oi = i % QW, si = i / QW;
        double bs = 0.0;
        static_for<i + 1, M>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            bs = bs + row[si][k] * xs[k];
        });
        const double xi = (i == M - 1) ? y[si] / row[si][i] : (y[si] - bs) / row[si][i];   // meaningful in the owner lane only
        if (sl == oi) x[si] = xi;
        xs[i] = shflq(xi, oi);
    });
    return ok;
}

// ---- node-direction records, four lanes per (node, direction): lane sl computes the groups g = sl + 4 s
template <int NG, int KERN>
__global__ void __launch_bounds__(ADP_TILE, 2) k_nodal_nodedir_q(Geo G, NodalArgs A, int u, int klo, int npl)
{
    constexpr int RG = (NG + QW - 1) / QW;
    constexpr int QPB = ADP_TILE / QW;
    const long long NV = G.NV;
    const int sl = threadIdx.x & (QW - 1), quad = threadIdx.x / QW;
    const long long total = (long long)npl * G.np;
    const double Ke = A.scal[S_KE];
    double *ndb = A.nd + (size_t)u * ND_SLOTS(NG) * NV;
    const double *Su = A.S + (size_t)u * NG * NV;
    bool ok = true;
    for (long long base = (long long)blockIdx.x * QPB; base < total; base += (long long)gridDim.x * QPB) {
        const bool live = base + quad < total;
        const long long item = live ? base + quad : total - 1;
        const int kl = klo + (int)(item / G.np), r = (int)(item % G.np);
        const long long idx = node_idx(G, kl, r);
        const Line q = line_of(G, u, kl, r);
        const int m = A.mat[idx] - 1;
        const double hh = q.h * q.h;
        double tfac = 0.0;
        if (A.cmode == 2) tfac = 1.0 - A.tbeta[m] + A.dfis[idx];
        double f0[NG], nuf[NG], chi[NG];
#pragma unroll
        for (int h = 0; h < NG; ++h) {
            f0[h] = (((A.curmask >> h) & 1u) ? A.f0b : A.f0a)[(size_t)h * NV + idx];
            nuf[h] = A.nuf[(size_t)h * NV + idx];
            chi[h] = A.chi[h * A.nmat + m];
        }
        double Bc[RG][NG], M2[RG][NG], bb[RG], a2[RG], L1[RG], Lm2[RG], Bcst[RG];
        static_for<0, RG>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int g = sl + QW * s;
            const bool gv = g < NG;
            const int gg = gv ? g : 0;                           // padding rows compute on group 0 and are discarded
            const double Dg = A.D[(size_t)gg * NV + idx], sigr = A.sigr[(size_t)gg * NV + idx];
            const double nufg = A.nuf[(size_t)gg * NV + idx], chig = A.chi[gg * A.nmat + m];
            // get_B (mod_nodal.f90:1345-1405)
#pragma unroll
            for (int h = 0; h < NG; ++h) {
                double dum;
                if (A.cmode == 1) {
                    if (gg == h) dum = sigr - chig * nuf[h] / Ke;
                    else dum = -A.sigs[((size_t)gg * NG + h) * NV + idx] - chig * nuf[h] / Ke;       // sigs(n,h,g)
                } else if (A.cmode == 2) {
                    if (gg == h) dum = sigr - tfac * chig * chig * nuf[h];                             // sic, :1383-1384
                    else dum = -A.sigs[((size_t)gg * NG + h) * NV + idx] - tfac * chig * nuf[h];
                } else {
                    if (gg == h) dum = sigr - chig * nuf[h] / Ke;
                    else dum = -A.sigs[((size_t)h * NG + gg) * NV + idx] - chi[h] * nufg / Ke;        // sigs(n,g,h)
                }
                Bc[s][h] = 0.25 * hh / Dg * dum;
            }
            double Bq, Eq;
            if (KERN == ADP_KERN_SANM) {
                const double *cc = A.abefgh + ((size_t)u * 6 * NG + gg) * NV + idx;
                Bq = cc[(size_t)1 * NG * NV];
                Eq = cc[(size_t)2 * NG * NV];
            } else { Bq = 1.0 / 35.0; Eq = 2.0 / 7.0; }
            Bcst[s] = Bq;
            // TLUpd1 / TLUpd2 (mod_nodal.f90:1047-1341)
            const double Sn = Su[(size_t)gg * NV + idx];
            const double Sp = q.has_p ? Su[(size_t)gg * NV + idx + q.off_p] : 0.0;
            const double Sm = q.has_m ? Su[(size_t)gg * NV + idx - q.off_m] : 0.0;
            double l1, l2, tm, tp, p1m, p2m, p1p, p2p, hp;
            if (!q.has_m) {
                if (q.bcm == 2) {
                    tm = 1.0; tp = q.hp / q.h;
                    p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0;
                    hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
                    l1 = (p1m * p2m * (Sp - Sn)) / hp;
                    l2 = (p1m * (Sp - Sn)) / hp;
                } else {
                    tp = q.hp / q.h; p1p = tp + 1.0;
                    l1 = (Sp - Sn) / p1p;
                    l2 = 0.0;
                }
            } else if (!q.has_p) {
                if (q.bcp == 2) {
                    tm = q.hm / q.h; tp = 1.0;
                    p1m = tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
                    hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
                    l1 = (p1p * p2p * (Sn - Sm)) / hp;
                    l2 = (p1p * (Sm - Sn)) / hp;
                } else {
                    tm = q.hm / q.h; p1m = tm + 1.0;
                    l1 = (Sn - Sm) / p1m;
                    l2 = 0.0;
                }
            } else {
                tm = q.hm / q.h; tp = q.hp / q.h;
                p1m = tm + 1.0; p2m = 2.0 * tm + 1.0; p1p = tp + 1.0; p2p = 2.0 * tp + 1.0;
                hp = 2.0 * p1m * p1p * (tm + tp + 1.0);
                l1 = (p1m * p2m * (Sp - Sn) + p1p * p2p * (Sn - Sm)) / hp;
                l2 = (p1m * (Sp - Sn) + p1p * (Sm - Sn)) / hp;
            }
            L1[s] = 0.25 * hh / Dg * l1;
            Lm2[s] = 0.25 * hh / Dg * l2;
            // get_a2matvec (mod_nodal.f90:769-825)
            double S;
            if (A.cmode == 2) S = 0.25 * hh / Dg * Sn;
            else S = 0.25 * hh / Dg * (Sn - A.exsrc[(size_t)gg * NV + idx]);
            double Bf = 0.0;
#pragma unroll
            for (int h = 0; h < NG; ++h) {
                const double be = Bc[s][h] * Eq;
                M2[s][h] = gv ? ((h == gg) ? be + 3.0 : be) : 0.0;
                Bf = Bf + Bc[s][h] * f0[h];
            }
            bb[s] = gv ? Bf - Eq * Lm2[s] + S : 0.0;
        });
        const bool okl = quad_lu_solve<NG>(sl, M2, bb, a2);
        if (live) ok = okl && ok;
        double a2all[NG];
        static_for<0, NG>([&](auto hc) {
            constexpr int h = decltype(hc)::value;
            a2all[h] = shflq(a2[h / QW], h % QW);
        });
        static_for<0, RG>([&](auto sc) {
            constexpr int s = decltype(sc)::value;
            const int g = sl + QW * s;
            if (live && g < NG) {
                // get_a4 (mod_nodal.f90:740-765)
                double Bf = 0.0;
#pragma unroll
                for (int h = 0; h < NG; ++h) Bf = Bf + Bc[s][h] * a2all[h];
                const double a4 = Bcst[s] * (Bf + Lm2[s]);
#pragma unroll
                for (int h = 0; h < NG; ++h) ndb[(size_t)(g * NG + h) * NV + idx] = Bc[s][h];
                ndb[(size_t)(NG * NG + 0 * NG + g) * NV + idx] = a2[s];
                ndb[(size_t)(NG * NG + 1 * NG + g) * NV + idx] = a4;
                ndb[(size_t)(NG * NG + 2 * NG + g) * NV + idx] = L1[s];
            }
        });
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
}

// row g of a one-node boundary problem (get_a1matvec_first / _last): matrix row and right-hand side
template <int NG>
__device__ __forceinline__ void boundary_row(const NodeDirRow<NG> &n, int g, bool first, int bc, double P, double dcf,
                                             double (&row)[NG], double &b)
{
    if (first) {
        if (bc == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? P * (n.Bc[hh] * n.F + 1.0) : P * n.Bc[hh] * n.F;
            b = P * (3.0 * n.a2 + n.Gc * n.a4 - n.F * n.L1);
        } else if (bc == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                row[hh] = (hh == g) ? -dcf * (1.0 + n.A * n.Bc[hh]) - 2.0 * P * (n.A * n.Bc[hh] * n.H + 1.0)
                                    : -dcf * n.A * n.Bc[hh] - 2.0 * P * n.A * n.Bc[hh] * n.H;
            b = 2.0 * P * (n.A * n.H * n.L1 - 3.0 * n.a2 - n.Gc * n.a4) - dcf * (n.a2 + n.a4 + n.f0 - n.A * n.L1);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) : dcf * n.A * n.Bc[hh];
            b = dcf * (n.a2 + n.a4 + n.f0 - n.A * n.L1);
        }
    } else {
        if (bc == 2) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? -P * (n.Bc[hh] * n.F + 1.0) : -P * n.Bc[hh] * n.F;
            b = P * (3.0 * n.a2 + n.Gc * n.a4 + n.F * n.L1);
        } else if (bc == 1) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh)
                row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) + 2.0 * P * (n.A * n.Bc[hh] * n.H + 1.0)
                                    : dcf * n.A * n.Bc[hh] + 2.0 * P * n.A * n.Bc[hh] * n.H;
            b = -2.0 * P * (n.A * n.H * n.L1 + 3.0 * n.a2 + n.Gc * n.a4) - dcf * (n.a2 + n.a4 + n.f0 + n.A * n.L1);
        } else {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[hh] = (hh == g) ? dcf * (1.0 + n.A * n.Bc[hh]) : dcf * n.A * n.Bc[hh];
            b = -dcf * (n.a2 + n.a4 + n.f0 + n.A * n.L1);
        }
    }
}

// one-node boundary problem over a quad: G rows, lane sl holds the groups g = sl + 4 s; then get_a3, the boundary
// current and the dn update (nodal_coup_upd)
template <int NG, int KERN>
__device__ __forceinline__ bool quad_boundary(const NodalArgs &A, bool active, int sl, bool first, int bc, double h, int u,
                                              const double *__restrict__ ndbase, long long NV, long long idx, long long gnode,
                                              double &best, long long &best_loc)
{
    constexpr int RG = (NG + QW - 1) / QW;
    const int sf = 2 * u, face = first ? sf + 1 : sf;
    double row[RG][NG], b[RG], a1[RG];
    static_for<0, RG>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        const int g = sl + QW * s, gg = g < NG ? g : 0;
        NodeDirRow<NG> n;
        nd_load_row<NG, KERN>(n, gg, ndbase, NV, idx, A, u);
        const double P = 2.0 * n.D / h;
        const double dcf = A.dc[((size_t)face * NG + gg) * NV + idx];
        boundary_row<NG>(n, gg, first, bc, P, dcf, row[s], b[s]);
        if (g >= NG) {
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) row[s][hh] = 0.0;
            b[s] = 0.0;
        }
    });
    bool ok = quad_lu_solve<NG>(sl, row, b, a1);
    if (!active) ok = true;
    double a1all[NG];
    static_for<0, NG>([&](auto hc) {
        constexpr int hq = decltype(hc)::value;
        a1all[hq] = shflq(a1[hq / QW], hq % QW);
    });
    static_for<0, RG>([&](auto sc) {
        constexpr int s = decltype(sc)::value;
        const int g = sl + QW * s;
        if (active && g < NG) {
            NodeDirRow<NG> n;
            nd_load_row<NG, KERN>(n, g, ndbase, NV, idx, A, u);
            double Bf = 0.0;
#pragma unroll
            for (int hh = 0; hh < NG; ++hh) Bf = Bf + n.Bc[hh] * a1all[hh];
            const double a3 = n.A * (Bf + n.L1);                 // get_a3
            double *dn = A.dn + ((size_t)g * 6 + face) * NV;
            const double dff = A.df[((size_t)g * 6 + face) * NV + idx];
            const double ndpr = dn[idx];
            double nw;
            if (first) {
                const double jp = -2.0 * n.D / h * (a1[s] - 3.0 * n.a2 + n.H * a3 - n.Gc * n.a4);
                nw = -(jp / n.f0 + dff);
            } else {
                const double jp = -2.0 * n.D / h * (a1[s] + 3.0 * n.a2 + n.H * a3 + n.Gc * n.a4);
                nw = -(jp / n.f0 - dff);
            }
            dn[idx] = nw;
            const double nder = fabs(nw - ndpr);
            if (nder > best || (nder == best && gnode < best_loc)) { best = nder; best_loc = gnode; }
        }
    });
    return ok;
}

// ---- surfaces, four lanes per surface: lane sl holds the rows i = sl + 4 s of the 2G x 2G two-node system
// (rows 0 .. G-1 current continuity of group i, rows G .. 2G-1 flux continuity of group i - G)
template <int NG, int KERN>
__global__ void __launch_bounds__(ADP_TILE, 1) k_nodal_surfaces_q(Geo G, NodalArgs A, int u, int klo, int npl, ArgMax am)
{
    constexpr int M = 2 * NG, RPL = (M + QW - 1) / QW;
    constexpr int QPB = ADP_TILE / QW;
    const long long NV = G.NV;
    const double *ndbase = A.nd + (size_t)u * ND_SLOTS(NG) * NV;
    const int sl = threadIdx.x & (QW - 1), quad = threadIdx.x / QW;
    const int sf = 2 * u;
    double best = -1.0;
    long long best_loc = 0x7fffffffffffffffLL;
    bool ok = true;
    const long long total = (long long)npl * G.np;
    for (long long base = (long long)blockIdx.x * QPB; base < total; base += (long long)gridDim.x * QPB) {
        const bool live = base + quad < total;
        const long long item = live ? base + quad : total - 1;
        const int kl = klo + (int)(item / G.np), r = (int)(item % G.np);
        const long long idx = node_idx(G, kl, r);
        const bool owned = (kl >= 0 && kl < G.nzl);
        const long long gnode = (long long)(G.k0 + kl) * G.np + r;
        const Line qn = line_of(G, u, kl, r);
        const bool do_first = live && !qn.has_m && owned;
        const bool do_inner = live && qn.has_p;
        const bool do_last = live && !qn.has_p && owned;
        if (__any_sync(COOP_FULL, do_first))
            ok = quad_boundary<NG, KERN>(A, do_first, sl, true, qn.bcm, qn.h, u, ndbase, NV, idx, gnode, best, best_loc) && ok;
        if (__any_sync(COOP_FULL, do_inner)) {
            // a quad without a "+" neighbour pairs the node with itself and discards the result
            const long long idp = do_inner ? idx + qn.off_p : idx;
            const int klp = (do_inner && u == 2) ? kl + 1 : kl;
            const int rp = !do_inner ? r : (u == 0) ? r + 1 : (u == 1) ? r + (int)qn.off_p : r;
            const Line qp = line_of(G, u, klp, rp);
            double row[RPL][M], b[RPL], sx[RPL];
            static_for<0, RPL>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                const int i = sl + QW * s;
                const bool iv = i < M, lower = i < NG;
                const int g = iv ? (lower ? i : i - NG) : 0;
                NodeDirRow<NG> n, p;
                nd_load_row<NG, KERN>(n, g, ndbase, NV, idx, A, u);
                nd_load_row<NG, KERN>(p, g, ndbase, NV, idp, A, u);
                if (lower) {
                    const double Pn = 2.0 * n.D / qn.h, Pp = 2.0 * p.D / qp.h;
#pragma unroll
                    for (int h = 0; h < NG; ++h) {
                        row[s][h] = (h == g) ? -Pn * (n.Bc[h] * n.F + 1.0) : -Pn * n.Bc[h] * n.F;
                        row[s][h + NG] = (h == g) ? Pp * (p.Bc[h] * p.F + 1.0) : Pp * p.Bc[h] * p.F;
                    }
                    b[s] = Pn * (3.0 * n.a2 + n.Gc * n.a4 + n.F * n.L1) + Pp * (3.0 * p.a2 + p.Gc * p.a4 - p.F * p.L1);
                } else {
                    const double dcn = A.dc[((size_t)sf * NG + g) * NV + idx];
                    const double dcp = A.dc[((size_t)(sf + 1) * NG + g) * NV + idp];
#pragma unroll
                    for (int h = 0; h < NG; ++h) {
                        row[s][h] = (h == g) ? dcn * (n.Bc[h] * n.A + 1.0) : dcn * n.Bc[h] * n.A;
                        row[s][h + NG] = (h == g) ? dcp * (p.Bc[h] * p.A + 1.0) : dcp * p.Bc[h] * p.A;
                    }
                    // ADF cross terms exactly as mod_nodal.f90:693-694 (An*Ln1 with dc_p, Ap*Lp1 with dc_n)
                    b[s] = dcp * (p.a2 + p.a4 + p.f0 - n.A * n.L1) - dcn * (n.a2 + n.a4 + n.f0 + p.A * p.L1);
                }
                if (!iv) {
#pragma unroll
                    for (int h = 0; h < M; ++h) row[s][h] = 0.0;
                    b[s] = 0.0;
                }
            });
            const bool okl = quad_lu_solve<M>(sl, row, b, sx);
            if (do_inner) ok = okl && ok;
            double a1all[NG];
            static_for<0, NG>([&](auto hc) {
                constexpr int h = decltype(hc)::value;
                a1all[h] = shflq(sx[h / QW], h % QW);             // a1(h) = sx(h)
            });
            static_for<0, RPL>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                if constexpr (QW * s < NG) {                     // slots that can hold a current-continuity row
                    const int g = sl + QW * s;
                    if (do_inner && g < NG) {
                        NodeDirRow<NG> n;
                        nd_load_row<NG, KERN>(n, g, ndbase, NV, idx, A, u);
                        const double f0p = (((A.curmask >> g) & 1u) ? A.f0b : A.f0a)[(size_t)g * NV + idp];
                        double Bf = 0.0;
#pragma unroll
                        for (int h = 0; h < NG; ++h) Bf = Bf + n.Bc[h] * a1all[h];
                        const double a3 = n.A * (Bf + n.L1);
                        const double jp = -2.0 * n.D / qn.h * (sx[s] + 3.0 * n.a2 + n.H * a3 + n.Gc * n.a4);
                        double *dn = A.dn + ((size_t)g * 6 + sf) * NV;
                        const double dfp = A.df[((size_t)g * 6 + sf) * NV + idx];
                        const double ndpr = dn[idx];
                        const double nw = (dfp * (n.f0 - f0p) - jp) / (n.f0 + f0p);
                        dn[idx] = nw;
                        A.dn[((size_t)g * 6 + sf + 1) * NV + idp] = nw;
                        if (owned) {
                            const double nd2 = fabs(nw - ndpr);
                            if (nd2 > best || (nd2 == best && gnode < best_loc)) { best = nd2; best_loc = gnode; }
                        }
                    }
                }
            });
        }
        if (__any_sync(COOP_FULL, do_last))
            ok = quad_boundary<NG, KERN>(A, do_last, sl, false, qn.bcp, qn.h, u, ndbase, NV, idx, gnode, best, best_loc) && ok;
    }
    if (!ok) atomicExch(A.errflag, ADP_STOP_LU_DIAG);
    grid_argmax(best, best_loc, am);
}

// grid of a fused kernel: about two waves of its resident CTAs, at most one CTA per work item
template <typename K>
static int fused_ctas(adp_ctx *c, K kernel, size_t smem)
{
    static std::map<const void *, int> cache;
    auto it = cache.find((const void *)kernel);
    int per_sm = 1;
    if (it == cache.end()) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NODAL_TPB, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cache[(const void *)kernel] = per_sm;
    } else per_sm = it->second;
    return c->sm_count * per_sm;
}

template <int NG, int KERN>
void launch_nodal_fused(adp_ctx *c, const NodalArgs &A, const ArgMax &am)
{
    if constexpr (NG <= 2) {
        const int zlo = (c->k0 > 0) ? -1 : 0, zhi = (c->k1 < c->nzz) ? c->nzl + 1 : c->nzl;
        auto clampi = [](long long v, long long lo, long long hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); };
        // x
        {
            const size_t smem = (size_t)(NG * NG + 8 * NG) * NODAL_TPB * sizeof(double);
            const long long total = (long long)c->nzl * c->np, tiles = (total + (NODAL_TPB - 2)) / (NODAL_TPB - 1);
            const int grid = clampi(tiles, 1, std::min<long long>(ADP_MAXPART, 4LL * fused_ctas(c, k_nodal_pair_x<NG, KERN>, smem)));
            k_nodal_pair_x<NG, KERN><<<grid, NODAL_TPB, smem, c->stream>>>(c->geo, A, am);
            c->launches++;
        }
        // y: chunks of rows
        {
            const long long total = (long long)c->nzl * c->nxx, tiles = (total + NODAL_TPB - 1) / NODAL_TPB;
            const int want = 2 * fused_ctas(c, k_nodal_march_y<NG, KERN>, 0);
            const int nchunk = clampi((want + tiles - 1) / tiles, 1, std::max(1, c->nyy / 8));
            const int grid = clampi(tiles * nchunk, 1, ADP_MAXPART);
            k_nodal_march_y<NG, KERN><<<grid, NODAL_TPB, 0, c->stream>>>(c->geo, c->geoxy, A, nchunk, am);
            c->launches++;
        }
        // z: chunks of planes; the neighbours' boundary planes join the line on interior slab boundaries
        {
            const int kfirst = zlo, klast = zhi - 1, ns = klast - kfirst;
            const long long tiles = (c->np + NODAL_TPB - 1) / NODAL_TPB;
            const int want = 2 * fused_ctas(c, k_nodal_march_z<NG, KERN>, 0);
            const int nchunk = clampi((want + tiles - 1) / tiles, 1, std::max(1, ns / 8));
            const int grid = clampi(tiles * nchunk, 1, ADP_MAXPART);
            k_nodal_march_z<NG, KERN><<<grid, NODAL_TPB, 0, c->stream>>>(c->geo, A, kfirst, klast, nchunk, am);
            c->launches++;
        }
    }
}

template <int NG>
void launch_nodal(adp_ctx *c, const NodalArgs &A, const ArgMax &am, bool refresh_abefgh)
{
    // planes that need node-direction data: own planes, plus the neighbour's boundary plane for
    // the z surfaces shared with the slabs below / above
    const int zlo = (c->k0 > 0) ? -1 : 0, zhi = (c->k1 < c->nzz) ? c->nzl + 1 : c->nzl;
    const int tpp = c->geo.tpp;
    if (A.kern == ADP_KERN_SANM && refresh_abefgh) {
        k_nodal_abefgh<NG><<<adp_grid(c, k_nodal_abefgh<NG>, tpp * (zhi - zlo)), ADP_TILE, 0, c->stream>>>(
            c->geo, c->d_D, c->d_sigr, A.abefgh, zlo, zhi - zlo);
        c->launches++;
    }
    // experiment (option nodal_fused = 1, G <= 2): one fused kernel per direction, the node-direction record stays in registers
    const bool fused = NG <= 2 && c->nodal_fused > 0;
    if (fused) {
        if (A.kern == ADP_KERN_SANM) launch_nodal_fused<NG, ADP_KERN_SANM>(c, A, am);
        else launch_nodal_fused<NG, ADP_KERN_PNM>(c, A, am);
        return;
    }
    // which form: nodal_coop -1 = automatic (quad kernels from G = 5), 0 one thread per item, 1 sixteen lanes per surface
    // (round 1's form for G >= 7), 2 quad kernels
    const int form = (c->nodal_coop < 0) ? (NG >= 5 ? 2 : 0) : c->nodal_coop;
    const bool sanm = A.kern == ADP_KERN_SANM;
    for (int u = 0; u < 3; ++u) {
        const int klo = (u == 2) ? zlo : 0, npl = ((u == 2) ? zhi : c->nzl) - klo;
        if constexpr (NG >= 3) {
            if (form == 2) {
                const int tiles = (int)(((long long)npl * c->np + ADP_TILE / QW - 1) / (ADP_TILE / QW));
                if (sanm) k_nodal_nodedir_q<NG, ADP_KERN_SANM><<<adp_grid(c, k_nodal_nodedir_q<NG, ADP_KERN_SANM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl);
                else k_nodal_nodedir_q<NG, ADP_KERN_PNM><<<adp_grid(c, k_nodal_nodedir_q<NG, ADP_KERN_PNM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl);
                c->launches++;
                continue;
            }
        }
        if (sanm)
            k_nodal_nodedir<NG, ADP_KERN_SANM><<<adp_grid(c, k_nodal_nodedir<NG, ADP_KERN_SANM>, tpp * npl), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl);
        else
            k_nodal_nodedir<NG, ADP_KERN_PNM><<<adp_grid(c, k_nodal_nodedir<NG, ADP_KERN_PNM>, tpp * npl), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl);
        c->launches++;
    }
    for (int u = 0; u < 3; ++u) {
        const int klo = (u == 2) ? zlo : 0, npl = c->nzl - klo;     // kl = -1: the surface shared with the slab below
        // measured on the C2 mesh (whole update, ms), one thread per surface -> 16 lanes: G = 8: 116 -> 77; G = 6: 46 -> 45;
        // G = 4: 12 -> 44; quad kernels: see DESIGN.md
        const long long items = (long long)npl * c->np;
        if constexpr (NG >= 3) {
            if (form == 2) {
                const int tiles = (int)((items + ADP_TILE / QW - 1) / (ADP_TILE / QW));
                if (sanm) k_nodal_surfaces_q<NG, ADP_KERN_SANM><<<adp_grid(c, k_nodal_surfaces_q<NG, ADP_KERN_SANM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
                else k_nodal_surfaces_q<NG, ADP_KERN_PNM><<<adp_grid(c, k_nodal_surfaces_q<NG, ADP_KERN_PNM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
                c->launches++;
                continue;
            }
        }
        if (form == 1) {
            const int tiles = (int)((items + ADP_TILE / COOP_W - 1) / (ADP_TILE / COOP_W));
            if (sanm) k_nodal_surfaces_coop<NG, ADP_KERN_SANM><<<adp_grid(c, k_nodal_surfaces_coop<NG, ADP_KERN_SANM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
            else k_nodal_surfaces_coop<NG, ADP_KERN_PNM><<<adp_grid(c, k_nodal_surfaces_coop<NG, ADP_KERN_PNM>, tiles), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
        } else {
            if (sanm) k_nodal_surfaces<NG, ADP_KERN_SANM><<<adp_grid(c, k_nodal_surfaces<NG, ADP_KERN_SANM>, tpp * npl), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
            else k_nodal_surfaces<NG, ADP_KERN_PNM><<<adp_grid(c, k_nodal_surfaces<NG, ADP_KERN_PNM>, tpp * npl), ADP_TILE, 0, c->stream>>>(c->geo, A, u, klo, npl, am);
        }
        c->launches++;
    }
}

template <int NG>
void preload_nodal(adp_ctx *c)
{
    adp_grid(c, k_nodal_abefgh<NG>, 1); adp_grid(c, k_nodal_nodedir<NG, ADP_KERN_SANM>, 1);
    adp_grid(c, k_nodal_nodedir<NG, ADP_KERN_PNM>, 1);
    adp_grid(c, k_nodal_surfaces<NG, ADP_KERN_SANM>, 1); adp_grid(c, k_nodal_surfaces<NG, ADP_KERN_PNM>, 1);
    adp_grid(c, k_nodal_surfaces_coop<NG, ADP_KERN_SANM>, 1); adp_grid(c, k_nodal_surfaces_coop<NG, ADP_KERN_PNM>, 1);
    if constexpr (NG >= 3) {
        adp_grid(c, k_nodal_nodedir_q<NG, ADP_KERN_SANM>, 1); adp_grid(c, k_nodal_nodedir_q<NG, ADP_KERN_PNM>, 1);
        adp_grid(c, k_nodal_surfaces_q<NG, ADP_KERN_SANM>, 1); adp_grid(c, k_nodal_surfaces_q<NG, ADP_KERN_PNM>, 1);
    }
    if constexpr (NG <= 2) {
        const size_t smx = (size_t)(NG * NG + 8 * NG) * NODAL_TPB * sizeof(double);
        fused_ctas(c, k_nodal_pair_x<NG, ADP_KERN_SANM>, smx); fused_ctas(c, k_nodal_pair_x<NG, ADP_KERN_PNM>, smx);
        fused_ctas(c, k_nodal_march_y<NG, ADP_KERN_SANM>, 0); fused_ctas(c, k_nodal_march_y<NG, ADP_KERN_PNM>, 0);
        fused_ctas(c, k_nodal_march_z<NG, ADP_KERN_SANM>, 0); fused_ctas(c, k_nodal_march_z<NG, ADP_KERN_PNM>, 0);
    }
}

}  // namespace

static NodalArgs make_args(adp_ctx *c, int cmode)
{
    NodalArgs A{};
    A.ng = c->ng; A.nmat = c->nmat; A.cmode = cmode; A.kern = c->kern;
    A.nnod_total = c->nnod;
    for (int g = 0; g < c->ng; ++g) A.f0[g] = c->d_f0[c->cur[g]] + (size_t)g * c->NV;
    A.f0a = c->d_f0[0]; A.f0b = c->d_f0[1]; A.curmask = 0;
    for (int g = 0; g < c->ng; ++g) if (c->cur[g]) A.curmask |= 1u << g;
    A.D = c->d_D; A.sigr = c->d_sigr; A.nuf = c->d_nuf; A.exsrc = c->d_exsrc; A.sigs = c->d_sigs;
    A.chi = c->d_chi; A.dc = c->d_dc; A.mat = c->d_mat; A.tbeta = c->d_tbeta; A.dfis = c->d_dfis;
    A.df = c->d_df; A.dn = c->d_dn; A.S = c->d_S; A.scal = c->d_scal; A.errflag = c->d_errflag;
    A.nd = c->d_nd; A.abefgh = c->d_abefgh;
    return A;
}


int adp_k_nodal_source(adp_ctx *c, int cmode)
{
    NodalArgs A = make_args(c, cmode);
    int rc;
    if (c->nranks > 1)
        for (int g = 0; g < c->ng; ++g)
            if ((rc = adp_comm_halo(c, c->d_f0[c->cur[g]] + (size_t)g * c->NV, 1))) return rc;
    k_nodal_source<<<adp_grid(c, k_nodal_source, c->geo.ntiles), ADP_TILE, 0, c->stream>>>(c->geo, A, nullptr);
    c->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) { c->err = "k_nodal_source launch failed"; return ADP_ERR_CUDA; }
    return ADP_OK;
}

// L(n,g) = L1 + L2 + L3 of Lxyz for every node (what `reactivity` needs), into d_L
int adp_k_lxyz_total(adp_ctx *c, double *d_L)
{
    NodalArgs A = make_args(c, 1);
    int rc;
    if (c->nranks > 1)
        for (int g = 0; g < c->ng; ++g)
            if ((rc = adp_comm_halo(c, c->d_f0[c->cur[g]] + (size_t)g * c->NV, 1))) return rc;
    k_nodal_source<<<adp_grid(c, k_nodal_source, c->geo.ntiles), ADP_TILE, 0, c->stream>>>(c->geo, A, d_L);
    c->launches++;
    if (cudaPeekAtLastError() != cudaSuccess) { c->err = "k_nodal_source (Lxyz) launch failed"; return ADP_ERR_CUDA; }
    return ADP_OK;
}

int adp_k_nodal_update(adp_ctx *c, int cmode)
{
    if (c->ng > 8) { c->err = "nodal update supports 1..8 energy groups"; return ADP_ERR_UNSUPPORTED; }
    // scratch of the update, allocated on first use: SANM constants cache and -- only for the two-kernel form
    // (G > 4, or option nodal_fused = 0) -- the node-direction store
    const size_t NS = (size_t)c->ng * c->ng + 3 * c->ng;
    const bool fused = c->ng <= 2 && c->nodal_fused > 0;
    if (!c->d_abefgh) {
        if (cudaMalloc((void **)&c->d_abefgh, (size_t)3 * 6 * c->ng * c->NV * sizeof(double)) != cudaSuccess) {
            c->err = "adp_k_nodal_update: out of device memory for the SANM constants";
            return ADP_ERR_CUDA;
        }
        c->abefgh_valid = false;
    }
    if (!fused && !c->d_nd) {
        if (cudaMalloc((void **)&c->d_nd, 3 * NS * c->NV * sizeof(double)) != cudaSuccess) {
            c->err = "adp_k_nodal_update: out of device memory for the node-direction store";
            return ADP_ERR_CUDA;
        }
    }
    int rc = adp_lazy_sync(c);            // the ADFs (option "lazy_adf": uploaded behind the outer iterations)
    if (rc) return rc;
    rc = adp_k_nodal_source(c, cmode);
    if (rc) return rc;
    NodalArgs A = make_args(c, cmode);
    if (c->nranks > 1) {
        // S3 is needed two planes deep across a slab boundary (quadratic transverse-leakage fit)
        for (int g = 0; g < c->ng; ++g)
            if ((rc = adp_comm_halo(c, c->d_S + ((size_t)2 * c->ng + g) * c->NV, 2))) return rc;
    }
    ArgMax am;
    am.scal = c->d_scal; am.part = c->d_part; am.part_loc = (long long *)(c->d_part + ADP_MAXPART);
    am.ticket = c->d_ticket; am.loc_out = c->d_argidx;
    const bool refresh = !c->abefgh_valid;
    switch (c->ng) {
    case 1: launch_nodal<1>(c, A, am, refresh); break;
    case 2: launch_nodal<2>(c, A, am, refresh); break;
    case 3: launch_nodal<3>(c, A, am, refresh); break;
    case 4: launch_nodal<4>(c, A, am, refresh); break;
    case 5: launch_nodal<5>(c, A, am, refresh); break;
    case 6: launch_nodal<6>(c, A, am, refresh); break;
    case 7: launch_nodal<7>(c, A, am, refresh); break;
    case 8: launch_nodal<8>(c, A, am, refresh); break;
    default: break;
    }
    if (c->kern == ADP_KERN_SANM) c->abefgh_valid = true;
    if (cudaPeekAtLastError() != cudaSuccess) {
        c->err = std::string("nodal update launch failed: ") + cudaGetErrorString(cudaGetLastError());
        return ADP_ERR_CUDA;
    }
    return ADP_OK;
}

void adp_k_preload_nodal(adp_ctx *c)
{
    adp_grid(c, k_nodal_source, 1);
    switch (c->ng) {
    case 1: preload_nodal<1>(c); break; case 2: preload_nodal<2>(c); break; case 3: preload_nodal<3>(c); break;
    case 4: preload_nodal<4>(c); break; case 5: preload_nodal<5>(c); break; case 6: preload_nodal<6>(c); break;
    case 7: preload_nodal<7>(c); break; case 8: preload_nodal<8>(c); break; default: break;
    }
}
