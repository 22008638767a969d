// xtab_node.cuh -- XStab_updt for one node (%XTAB branch-table cross sections).
//
// brInterp (mod_xsec.f90:520-788) for the unrodded and, under a control rod, the rodded branch
// tables, the volume-weighted mix of crod_tab_updt (:300-390) and Dsigr_updt (:199-226).  The tables
// of a material are one (nd, nb, nf, nm, nval) block, last index fastest, nval = 4G + G*G + 6G
// values packed [sigtr(G), siga(G), nuf(G), sigf(G), sigs(g -> h, g slow), dc(g, face)] -- a few KB
// that stay in L1/L2.  Every value sees the reference's operations in the reference's order
// (a + radx * (b - a); moderator temperature first, then fuel temperature, boron, coolant density),
// so the result is bit-exact against the reference arithmetic.
//
// The functions are __host__ __device__ and free of CUDA-only constructs: k_xs_update_xtab
// (cmfd_kernels.cu) calls them per thread, and tests/hostcheck compiles the very same source with
// g++ to check it against the numpy restatement on machines without a GPU (test infrastructure
// only -- nothing in the product library executes this on the host).
#pragma once
#include "../../include/adpres_b200.h"

#ifdef __CUDACC__
#define ADP_HD __host__ __device__
#define ADP_NOINLINE __noinline__
#else
#define ADP_HD
#define ADP_NOINLINE __attribute__((noinline))
#endif

struct XtabTables {
    int ng, nval;
    const int *meta;                    // [nmat][6]: nd, nb, nf, nm, trod, offset of pd|pb|pf|pm in par
    const long long *toff;              // [nmat] offset of the material's block in xs / rxs
    const double *par, *xs, *rxs;
};
struct XtabOut {
    double *D, *sigr, *nuf, *sigf;      // [g][NV]
    double *sigs;                       // [h][g][NV] = sigs(n, g, h)
    double *dc;                         // [f][g][NV]
};

// Host side of adp_set_xtab: per-material meta records and block offsets from dims (4 per material: nd, nb,
// nf, nm) and trod.  meta: 6 ints per material, toff: one offset per material.  Returns false on a dimension < 1.
inline bool xtab_pack_meta(int nmat, int ng, const int *dims, const int *trod, int *meta, long long *toff,
                           long long *npar_out, long long *ntab_out, bool *any_rod_out)
{
    const long long nval = 4LL * ng + (long long)ng * ng + 6LL * ng;
    long long npar = 0, ntab = 0;
    bool any_rod = false;
    for (int m = 0; m < nmat; ++m) {
        const int *d = dims + 4 * m;
        if (d[0] < 1 || d[1] < 1 || d[2] < 1 || d[3] < 1) return false;
        for (int k = 0; k < 4; ++k) meta[(long long)m * 6 + k] = d[k];
        meta[(long long)m * 6 + 4] = trod[m];
        meta[(long long)m * 6 + 5] = (int)npar;
        toff[m] = ntab;
        npar += (long long)d[0] + d[1] + d[2] + d[3];
        ntab += (long long)d[0] * d[1] * d[2] * d[3] * nval;
        any_rod = any_rod || trod[m] == 1;
    }
    *npar_out = npar; *ntab_out = ntab; *any_rod_out = any_rod;
    return true;
}

// Outcome of the top-down sweep of crod_updt / crod_tab_updt (mod_xsec.f90:253-279,322-372) for the node
// whose upper face lies `dum` below the core top: -1 = the sweep never reaches it (unrodded), 1 = fully
// rodded, else the rodded fraction of the node holding the rod tip.  rodh = distance of the tip from the
// core top; a tip above the core (rodh < 0) never meets the `partial` test, so the whole column counts as
// rodded -- the reference's behaviour.  rodh == dum below the top plane: the node above took it as a
// partial node with vfrac = 1 and the sweep EXITed.
ADP_HD inline double xt_rod_fraction(double rodh, double dum, double hz, bool top_plane)
{
    if (rodh < 0.0) return 1.0;
    if (rodh > dum + hz) return 1.0;
    if (rodh > dum || (rodh == dum && top_plane)) return (rodh - dum) / hz;
    return -1.0;
}

// The STOPs that follow every XS update in the reference: Dsigr_updt's "Negative diffusion coefficient encountered"
// (sigtr < 1.e-5, mod_xsec.f90:217) and check_xs (:104-127; its scattering test is done where sigs is formed).  The
// literals are default REAL in the reference.
ADP_HD inline bool xs_check_fails(double sigtr, double D, double sigr, double nuf)
{
    return sigtr < (double)1.e-5f || D < (double)1.e-20f || sigr < 0.0 || nuf < 0.0;
}

// the two closest branch points (0-based i1, i2); up to 20 % (boron: 100 ppm) outside the table the end
// interval extrapolates, beyond that the reference STOPs (mod_xsec.f90:556-650)
ADP_HD inline bool xt_bracket(double x, const double *par, int dim, bool absolute, int &i1, int &i2)
{
    i1 = 0; i2 = 0;
    if (dim <= 1) return true;
    const double lo = par[0], hi = par[dim - 1];
    if (x >= lo && x <= hi) {
        for (int s = 1; s < dim; ++s)
            if (x >= par[s - 1] && x <= par[s]) { i1 = s - 1; i2 = s; break; }
        return true;
    }
    if (x < lo && (absolute ? (lo - x) < 100.0 : (lo - x) / lo < (double)0.2f)) { i1 = 0; i2 = 1; return true; }
    if (x > hi && (absolute ? (x - hi) < 100.0 : (x - hi) / hi < (double)0.2f)) { i1 = dim - 2; i2 = dim - 1; return true; }
    return false;
}

struct XtPoint {
    long long o[8][2];                  // offsets of the 8 (s,t,u) corners at v1 / v2
    double rm, rf, rb, rd;              // interpolation weights (radx) per parameter
    bool im, iff, ib, id;               // parameter has more than one branch
};

// one packed value c at the point q: xs(1..8) of brInterp folded down to xs(1)
static ADP_HD ADP_NOINLINE double xt_value(const double *tab, const XtPoint &q, int c)
{
    double x[8];
    for (int i = 0; i < 8; ++i) {
        const double a = tab[q.o[i][0] + c];
        if (q.im) { const double b = tab[q.o[i][1] + c]; x[i] = a + q.rm * (b - a); }
        else x[i] = a;
    }
    if (q.iff) {
        x[0] = x[0] + q.rf * (x[1] - x[0]); x[2] = x[2] + q.rf * (x[3] - x[2]);
        x[4] = x[4] + q.rf * (x[5] - x[4]); x[6] = x[6] + q.rf * (x[7] - x[6]);
    }
    if (q.ib) { x[0] = x[0] + q.rb * (x[2] - x[0]); x[4] = x[4] + q.rb * (x[6] - x[4]); }
    if (q.id) x[0] = x[0] + q.rd * (x[4] - x[0]);
    return x[0];
}

// XStab_updt for node idx of material m (0-based).  w: xt_rod_fraction of the node (< 0: unrodded);
// rodded_column: the node lies under a control rod bank (negative values are then suppressed for
// everything but sigtr, mod_xsec.f90:374-387).  Returns 0 or the reference's STOP as ADP_STOP_XTAB_*.
ADP_HD inline int xtab_node(const XtabTables &T, int m, double w, bool rodded_column, double xc, double xb, double xf,
                            double xm, const XtabOut &O, long long NV, long long idx)
{
    const int ng = T.ng;
    const int *mt = T.meta + 6 * m;
    const int nd = mt[0], nb = mt[1], nf = mt[2], nm = mt[3];
    const double *pd = T.par + mt[5], *pb = pd + nd, *pf = pb + nb, *pm = pf + nf;
    if (w >= 0.0 && mt[4] != 1) return ADP_STOP_XTAB_NOROD;
    int s[2], t[2], u[2], v[2];
    const bool ok = xt_bracket(xc, pd, nd, false, s[0], s[1]) & xt_bracket(xb, pb, nb, true, t[0], t[1]) &
                    xt_bracket(xf, pf, nf, false, u[0], u[1]) & xt_bracket(xm, pm, nm, false, v[0], v[1]);
    if (!ok) return ADP_STOP_XTAB_RANGE;
    XtPoint q;
    q.im = nm > 1; q.iff = nf > 1; q.ib = nb > 1; q.id = nd > 1;
    q.rm = q.im ? (xm - pm[v[0]]) / (pm[v[1]] - pm[v[0]]) : 0.0;
    q.rf = q.iff ? (xf - pf[u[0]]) / (pf[u[1]] - pf[u[0]]) : 0.0;
    q.rb = q.ib ? (xb - pb[t[0]]) / (pb[t[1]] - pb[t[0]]) : 0.0;
    q.rd = q.id ? (xc - pd[s[0]]) / (pd[s[1]] - pd[s[0]]) : 0.0;
    for (int i = 0; i < 8; ++i) {        // xs(1..8) of brInterp: corner i = (s, t, u) from bits 4, 2, 1 of i
        const long long base = (((long long)s[(i >> 2) & 1] * nb + t[(i >> 1) & 1]) * nf + u[i & 1]) * nm;
        q.o[i][0] = T.toff[m] + (base + v[0]) * T.nval;
        q.o[i][1] = T.toff[m] + (base + v[1]) * T.nval;
    }
    // one packed value: unrodded, rodded or the volume-weighted mix
    auto val = [&](int c) -> double {
        double x = xt_value(T.xs, q, c);
        if (w >= 0.0) {
            const double xr = xt_value(T.rxs, q, c);
            x = (w == 1.0) ? xr : (1.0 - w) * x + w * xr;
        }
        if (rodded_column && c >= ng && x < 0.0) x = 0.0;
        return x;
    };
    int rc = ADP_OK;
    for (int g = 0; g < ng; ++g) {
        const double sigtr = val(g), siga = val(ng + g), nuf = val(2 * ng + g), sigf = val(3 * ng + g);
        double dum = 0.0;
        for (int h = 0; h < ng; ++h) {
            const double ss = val(4 * ng + g * ng + h);                     // sigs(n, g, h): g -> h
            O.sigs[((size_t)h * ng + g) * NV + idx] = ss;
            if (h != g) dum = dum + ss;
            if (ss < 0.0) rc = ADP_STOP_XS_CHECK;
        }
        const double D = 1.0 / (3.0 * sigtr), sigr = siga + dum;
        O.D[(size_t)g * NV + idx] = D;
        O.sigr[(size_t)g * NV + idx] = sigr;
        O.nuf[(size_t)g * NV + idx] = nuf;
        O.sigf[(size_t)g * NV + idx] = sigf;
        for (int f = 0; f < 6; ++f)
            O.dc[((size_t)f * ng + g) * NV + idx] = val(4 * ng + ng * ng + g * 6 + f);
        if (xs_check_fails(sigtr, D, sigr, nuf)) rc = ADP_STOP_XS_CHECK;
    }
    return rc;
}
