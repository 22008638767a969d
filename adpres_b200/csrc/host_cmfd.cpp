// host_cmfd.cpp -- the host-side mirror of the reference's outer-iteration procedures.
//
// In the drop-in deployment these loops stay in Fortran (src/mod_cmfd.f90 keeps `do p = 1,
// nout`, the prints and the STOPs, and calls adp_outer_iter / adp_nodal_upd once per pass --
// see fortran/adpres_b200_cmfd.f90).  No Fortran compiler exists in this image, so the same
// loops are restated here in C++ with the reference's names, argument meaning and error
// behaviour; the parity tests drive these.  Nothing in here computes on the CPU: every
// numerical statement of the loop body is a CUDA kernel behind the C ABI.
#include <cstdio>

#include "adp_internal.cuh"

namespace adpres {
namespace cmfd {

enum Kind { OUTER, OUTER_FS, OUTER_AD, OUTER_TR, OUTER_TH };

// Common body of outer / outer_fs / outer_ad / outer_th / outer_tr
// (mod_cmfd.f90:415-509 / 513-598 / 602-699 / 703-796 / 800-868).
static int outer_body(adp_ctx *c, Kind kind, int popt, int maxn, double ht, int *maxi, int *niter)
{
    const int mode = (kind == OUTER_AD) ? ADP_MODE_ADJOINT
                   : (kind == OUTER_FS) ? ADP_MODE_FIXEDSRC
                   : (kind == OUTER_TR) ? ADP_MODE_TRANSIENT : ADP_MODE_FORWARD;
    int rc;
    // CALL matrix_setup(1)
    if ((rc = adp_matrix_setup(c, 1))) return rc;
    // first-call allocation / initialisation (:448-454, :544-550, :635-641, :737-743)
    if (!c->have_flux) {
        if (kind == OUTER_AD) {
            if (popt > 0 && (rc = adp_init_flux(c, 1))) return rc;
        } else if (kind != OUTER_TR) {
            if ((rc = adp_init_flux(c, 0))) return rc;
        }
    }
    if (!c->have_flux) { c->err = "outer*: no flux (outer_ad(0) / outer_tr need a previous forward solution)"; return ADP_ERR_USAGE; }
    // get_exsrc(ht, exsrc)  (:830)
    if (kind == OUTER_TR && (rc = adp_get_exsrc(c, ht))) return rc;
    // f = Integrate(fs0); errn = 1; e1 = Integrate(errn)
    if ((rc = adp_outer_begin(c, mode))) return rc;

    const int nloop = (kind == OUTER_TH) ? maxn : c->nout;
    int p;
    double Ke = 0.0, ser = 0.0, fer = 0.0;
    for (p = 1; p <= nloop; ++p) {
        if ((rc = adp_outer_iter(c, mode, p, &Ke, &ser, &fer))) return rc;
        if (p % c->nac == 0 && popt > 0 && c->trace) c->trace(c->trace_user, 1, p, 0, 0, 0, 0, 0, 0);
        // Nodal coefficients update (:490, :581, :679-680, :781-784, :857)
        if (p % c->nupd == 0 && c->kern != ADP_KERN_FDM && !(kind == OUTER_AD && popt <= 0)) {
            const int nmode = (kind == OUTER_AD) ? 0 : (kind == OUTER_TR) ? 2 : 1;
            double ndmax; int im, jm, km;
            if ((rc = adp_nodal_upd(c, nmode, &ndmax, &im, &jm, &km))) return rc;
            if (popt > 0 && c->trace) c->trace(c->trace_user, 2, p, ndmax, 0, 0, im, jm, km);
        }
        if (c->trace) c->trace(c->trace_user, 0, p, Ke, ser, fer, 0, 0, 0);
        // exit test (:495); 1.e-2 is a default-REAL literal in the reference
        if ((ser < c->serc) && (fer < c->ferc) && (c->ndmax < (double)1.e-2f)) break;
    }
    if (niter) *niter = (p > nloop) ? nloop : p;
    if (kind == OUTER_TR) {
        if (maxi) *maxi = (p == nloop + 1) ? 1 : 0;          // :861-865
    } else if (kind != OUTER_TH && p - 1 == nloop) {
        c->err = "MAXIMUM NUMBER OF OUTER ITERATION IS REACHED. CHECK PROBLEM SPECIFICATION OR CHANGE ITERATION CONTROL (%ITER).";
        return ADP_STOP_MAXOUTER;                             // :498-505
    }
    return ADP_OK;
}

int outer(adp_ctx *c, int popt, int *niter) { return outer_body(c, OUTER, popt, 0, 0.0, nullptr, niter); }
int outer_fs(adp_ctx *c, int popt, int *niter) { return outer_body(c, OUTER_FS, popt, 0, 0.0, nullptr, niter); }
int outer_ad(adp_ctx *c, int popt, int *niter) { return outer_body(c, OUTER_AD, popt, 0, 0.0, nullptr, niter); }
int outer_th(adp_ctx *c, int maxn, int *niter) { return outer_body(c, OUTER_TH, 0, maxn, 0.0, nullptr, niter); }
int outer_tr(adp_ctx *c, double ht, int *maxi, int *niter) { return outer_body(c, OUTER_TR, 0, 0, ht, maxi, niter); }

}  // namespace cmfd
}  // namespace adpres

extern "C" int adp_outer(adp_ctx *c, int popt, int *niter) { return c ? adpres::cmfd::outer(c, popt, niter) : ADP_ERR_USAGE; }
extern "C" int adp_outer_fs(adp_ctx *c, int popt, int *niter) { return c ? adpres::cmfd::outer_fs(c, popt, niter) : ADP_ERR_USAGE; }
extern "C" int adp_outer_ad(adp_ctx *c, int popt, int *niter) { return c ? adpres::cmfd::outer_ad(c, popt, niter) : ADP_ERR_USAGE; }
extern "C" int adp_outer_th(adp_ctx *c, int maxn, int *niter) { return c ? adpres::cmfd::outer_th(c, maxn, niter) : ADP_ERR_USAGE; }
extern "C" int adp_outer_tr(adp_ctx *c, double ht, int *maxi, int *niter)
{
    return c ? adpres::cmfd::outer_tr(c, ht, maxi, niter) : ADP_ERR_USAGE;
}
