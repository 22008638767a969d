"""ctypes binding of the C ABI (include/adpres_b200.h) -- the same entry points the Fortran
ISO_C_BINDING shim binds (fortran/adpres_b200_bind.f90).  There is no CPU fallback: if the
CUDA library is missing or no device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADPRES_B200_LIB", os.path.join(_HERE, "libadpres_b200.so"))   # override: A/B builds

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

MODE_FORWARD, MODE_ADJOINT, MODE_FIXEDSRC, MODE_TRANSIENT = 0, 1, 2, 3
STOP_MAXOUTER, STOP_LU_DIAG, STOP_NDMAX, STOP_ZERO_POWER, STOP_STEAM_TABLE = 1, 2, 3, 4, 5
STOP_XTAB_RANGE, STOP_XTAB_NOROD, STOP_XS_CHECK = 6, 7, 8

TRACE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int)

# every symbol include/adpres_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "adp_create", "adp_destroy", "adp_last_error", "adp_version", "adp_comm_unique_id", "adp_comm_init", "adp_comm_init_env", "adp_slab", "adp_set_xs_mask", "adp_get_state_mask",
    "adp_set_geometry", "adp_set_xs", "adp_set_control", "adp_matrix_setup", "adp_init_flux", "adp_outer_begin",
    "adp_outer_iter", "adp_nodal_upd", "adp_powdis", "adp_integrate", "adp_set_kinetics", "adp_set_kinetics_xtab", "adp_set_transient",
    "adp_get_exsrc", "adp_set_material_xs", "adp_set_crod", "adp_xs_update", "adp_set_feedback", "adp_xs_update_th", "adp_get_xs",
    "adp_set_xtab", "adp_set_crod_map", "adp_xs_update_xtab", "adp_get_dc", "adp_save_adjoint", "adp_ipden", "adp_update_omeg", "adp_begin_time_step", "adp_upden", "adp_powtot", "adp_asm_pow", "adp_axi_pow", "adp_asm_flux", "adp_set_th", "adp_set_th_state", "adp_get_th_state", "adp_th_pline",
    "adp_th_upd", "adp_th_trans",
    "adp_reactivity", "adp_get_state", "adp_set_state", "adp_set_s0", "adp_get_nod", "adp_set_nod_dn", "adp_lxyz_total", "adp_get_exsrc_arrays",
    "adp_get_ndmax", "adp_get_errors", "adp_set_trace", "adp_outer", "adp_outer_ad", "adp_outer_fs", "adp_outer_th", "adp_outer_tr",
    "adp_sp_matvec", "adp_bicg", "adp_get_matrix", "adp_get_source", "adp_set_option", "adp_launch_count",
    "adp_bench_kernel", "adp_profile_report", "adp_outer_steps", "adp_timer_start", "adp_timer_stop",
]

_lib = None


class AdpresError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"adpres_b200 error {code}: {msg}")
        self.code = code


def load():
    """Load libadpres_b200.so.  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m adpres_b200.build` "
                               "(there is no CPU fallback for the CUDA hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.adp_last_error.restype = C.c_char_p
        _lib.adp_last_error.argtypes = [C.c_void_p]
        _lib.adp_version.restype = C.c_char_p
    return _lib


def _d(a):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous), \
        "expected a contiguous float64 array"
    return a.ctypes.data_as(_dp)


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def pack_xtab(p):
    """The arguments of adp_set_xtab from the branch tables of a deck (p.xtab, deck.read_xtab_composition):
    dims (nmat, 4) int32 C-order [= (4, nmat) column-major], trod (nmat), par, xs, rxs (None without rodded sets)."""
    dims = np.array([[t["nd"], t["nb"], t["nf"], t["nm"]] for t in p.xtab], dtype=np.int32)
    trod = np.array([t["trod"] for t in p.xtab], dtype=np.int32)
    par = np.concatenate([np.concatenate([t["pd"], t["pb"], t["pf"], t["pm"]]) for t in p.xtab]).astype(np.float64)
    xs = np.concatenate([np.ascontiguousarray(t["xs"]).ravel() for t in p.xtab])
    rxs = None
    if (trod == 1).any():
        rxs = np.concatenate([np.ascontiguousarray(t["rxs"] if t["rxs"] is not None else np.zeros_like(t["xs"])).ravel()
                              for t in p.xtab])
    return dims, trod, par, xs, rxs


class Solver:
    """One device context fed with the arrays of a ``deck.Problem`` (what ``sdata`` holds).
    Method names are those of the reference procedures."""

    def __init__(self, p, device=0, nranks=1, rank=0, uid=None, nupd=None, nout=None, nin=None, nac=None,
                 serc=None, ferc=None, kern=None):
        self.L = load()
        self.p = p
        self.N, self.G = p.nnod, p.ng
        self.h = C.c_void_p()
        rc = self.L.adp_create(C.byref(self.h), int(device))
        if rc:
            raise AdpresError(rc, self.L.adp_last_error(None).decode())
        if nranks > 1:
            self._chk(self.L.adp_comm_init(self.h, nranks, rank, uid))
        self.nranks, self.rank = nranks, rank
        self._trace_cb = None
        self.load_problem(p, nupd=nupd, nout=nout, nin=nin, nac=nac, serc=serc, ferc=ferc, kern=kern)

    def load_problem(self, p, **control):
        """Geometry + cross sections + iteration control of `p` into this context (a context can be
        re-used for another deck: adp_set_geometry re-sizes everything)."""
        self.p = p
        self.N, self.G = p.nnod, p.ng
        self._chk(self.L.adp_set_geometry(self.h, p.nxx, p.nyy, p.nzz, p.nnod, p.ng, p.nmat, _i(p.ix), _i(p.iy),
                                          _i(p.iz), _i(p.ystag_smin), _i(p.ystag_smax), _i(p.xstag_smin),
                                          _i(p.xstag_smax), _d(p.xdel), _d(p.ydel), _d(p.zdel), _i(p.bc), _i(p.mat)))
        k0, k1 = C.c_int(), C.c_int()
        self.L.adp_slab(self.h, C.byref(k0), C.byref(k1))
        self.k0, self.k1 = k0.value, k1.value
        self.own = slice(self.k0 * p.npl, self.k1 * p.npl)      # node range this rank owns
        self.set_xs()
        self.set_control(**control)
        self.trace_rows, self.trace_nodal, self.trace_extrp = [], [], []

    # ---- plumbing
    def _chk(self, rc):
        if rc < 0:
            raise AdpresError(rc, self.L.adp_last_error(self.h).decode())
        return rc

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.adp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self):
        return self.L.adp_last_error(self.h).decode()

    def set_option(self, name, value):
        self._chk(self.L.adp_set_option(self.h, name.encode(), int(value)))

    def reset_nodal(self):
        """the next matrix_setup(1) zeroes dn again, ndmax = 0 (state before the first coup_coef call of a run)"""
        self.set_option("reset_nodal", 1)

    def profile_report(self):
        """[(launch site, launches, total ms)] since set_option("profile", 1); the site is the source line of
        csrc/cmfd_kernels.cu with the launch (kernel name looked up in the source)"""
        n = 256
        lines, counts, ms = (C.c_int * n)(), (C.c_int * n)(), (C.c_double * n)()
        m = self.L.adp_profile_report(self.h, n, lines, counts, ms)
        src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "cmfd_kernels.cu")).read().split("\n")
        out = []
        for q in range(max(m, 0)):
            ln = lines[q]
            text = " ".join(x.strip() for x in src[max(0, ln - 4):ln])
            import re as _re
            names = _re.findall(r"(k_[a-z_]+|launch_[a-z_]+)", text)
            out.append(("%s:%d" % (names[-1] if names else "?", ln), counts[q], ms[q]))
        return out

    # ---- inputs
    def set_xs(self, **kw):
        p = self.p
        keep = {}

        def g(k):
            if kw and k not in kw:
                return None                      # partial update: NULL keeps the device copy
            a = np.asfortranarray(kw.get(k, getattr(p, k)), dtype=np.float64)
            keep[k] = a
            return _d(a)
        self._chk(self.L.adp_set_xs(self.h, g("D"), g("sigr"), g("nuf"), g("sigf"), g("sigs"), g("chi"), g("dc"),
                                    g("exsrc")))

    def set_control(self, nout=None, nin=None, nac=None, nupd=None, serc=None, ferc=None, kern=None):
        p = self.p
        v = lambda a, b: b if a is None else a
        self.ctl = dict(nout=v(nout, p.nout), nin=v(nin, p.nin), nac=v(nac, p.nac), nupd=v(nupd, p.nupd),
                        serc=v(serc, p.serc), ferc=v(ferc, p.ferc), kern=v(kern, p.kern))
        c = self.ctl
        self._chk(self.L.adp_set_control(self.h, c["nout"], c["nin"], c["nac"], c["nupd"], C.c_double(c["serc"]),
                                         C.c_double(c["ferc"]), c["kern"]))

    def set_state(self, f0=None, fs0=None, Ke=1.0):
        f0 = None if f0 is None else np.asfortranarray(f0)
        self._chk(self.L.adp_set_state(self.h, _d(f0), _d(fs0), C.c_double(Ke)))

    def set_s0(self, s0, g):
        self._chk(self.L.adp_set_s0(self.h, _d(np.asfortranarray(s0)), int(g)))

    def set_kinetics(self, ibeta, lamb, velo, tbeta, sth, bth):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (ibeta, lamb, velo, tbeta)]
        self._chk(self.L.adp_set_kinetics(self.h, _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]), C.c_double(sth), C.c_double(bth)))

    def set_kinetics_xtab(self, mibeta, mlamb, mvelo, tbeta, sth, bth):
        """%XTAB decks: per-material iBeta / lamb (nmat, 6) and velo (nmat, ng), rows = materials"""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (mibeta, mlamb, mvelo, tbeta)]
        self._chk(self.L.adp_set_kinetics_xtab(self.h, _d(a[0]), _d(a[1]), _d(a[2]), _d(a[3]), C.c_double(sth), C.c_double(bth)))

    def set_transient(self, c0=None, ft=None, fst=None, omeg=None, sigrp=None, L=None):
        f = lambda a: None if a is None else np.asfortranarray(a, dtype=np.float64)
        arrs = [f(x) for x in (c0, ft, fst, omeg, sigrp, L)]
        self._chk(self.L.adp_set_transient(self.h, *[_d(a) for a in arrs]))

    def set_nod_dn(self, dn):
        self._chk(self.L.adp_set_nod_dn(self.h, _d(np.asfortranarray(dn))))

    # ---- trace (what the reference prints)
    def enable_trace(self):
        self.trace_rows, self.trace_nodal, self.trace_extrp = [], [], []

        def cb(user, event, p, a, b, c, i, j, k):
            if event == 0:
                self.trace_rows.append((p, a, b, c))
            elif event == 1:
                self.trace_extrp.append(p)
            else:
                self.trace_nodal.append((p, a, i, j, k))
        self._trace_cb = TRACE_FN(cb)
        self._chk(self.L.adp_set_trace(self.h, self._trace_cb, None))

    # ---- hot path
    def matrix_setup(self, opt):
        return self._chk(self.L.adp_matrix_setup(self.h, opt))

    def init_flux(self, adjoint=False):
        return self._chk(self.L.adp_init_flux(self.h, int(adjoint)))

    def outer_begin(self, mode=MODE_FORWARD):
        return self._chk(self.L.adp_outer_begin(self.h, mode))

    def outer_iter(self, mode, p):
        ke, ser, fer = C.c_double(), C.c_double(), C.c_double()
        self._chk(self.L.adp_outer_iter(self.h, mode, p, C.byref(ke), C.byref(ser), C.byref(fer)))
        return ke.value, ser.value, fer.value

    def nodal_upd(self, nmode):
        nd, i, j, k = C.c_double(), C.c_int(), C.c_int(), C.c_int()
        rc = self._chk(self.L.adp_nodal_upd(self.h, nmode, C.byref(nd), C.byref(i), C.byref(j), C.byref(k)))
        return rc, nd.value, (i.value, j.value, k.value)

    def _run(self, fn, *args):
        n = C.c_int(0)
        rc = self._chk(fn(self.h, *args, C.byref(n)))
        return rc, n.value

    def outer(self, popt=1):
        return self._run(self.L.adp_outer, popt)

    def outer_fs(self, popt=1):
        return self._run(self.L.adp_outer_fs, popt)

    def outer_ad(self, popt=1):
        return self._run(self.L.adp_outer_ad, popt)

    def outer_th(self, maxn):
        return self._run(self.L.adp_outer_th, maxn)

    def outer_tr(self, ht):
        maxi, n = C.c_int(0), C.c_int(0)
        rc = self._chk(self.L.adp_outer_tr(self.h, C.c_double(ht), C.byref(maxi), C.byref(n)))
        return rc, bool(maxi.value), n.value

    def get_exsrc(self, ht):
        return self._chk(self.L.adp_get_exsrc(self.h, C.c_double(ht)))

    def powdis(self, fixedsrc=False):
        pw = np.zeros(self.N)
        rc = self._chk(self.L.adp_powdis(self.h, _d(pw), int(fixedsrc)))
        return rc, pw

    # ---- result reductions (mod_io.f90 AsmPow / AxiPow / AsmFlux)
    def asm_pow(self):
        p = self.p
        fasm = np.zeros((p.nx, p.ny), order="F")
        im, jm = C.c_int(), C.c_int()
        self._chk(self.L.adp_asm_pow(self.h, p.nx, p.ny, _i(np.ascontiguousarray(p.xdiv, dtype=np.int32)),
                                     _i(np.ascontiguousarray(p.ydiv, dtype=np.int32)), _d(fasm), C.byref(im), C.byref(jm)))
        return fasm, im.value, jm.value

    def axi_pow(self):
        p = self.p
        faxi = np.zeros(p.nz)
        am = C.c_int()
        self._chk(self.L.adp_axi_pow(self.h, p.nz, _i(np.ascontiguousarray(p.zdiv, dtype=np.int32)), _d(faxi), C.byref(am)))
        return faxi, am.value

    def asm_flux(self, norm=None):
        p = self.p
        fasm = np.zeros((p.nx, p.ny, p.ng), order="F")
        neg = C.c_int()
        self._chk(self.L.adp_asm_flux(self.h, p.nx, p.ny, _i(np.ascontiguousarray(p.xdiv, dtype=np.int32)),
                                      _i(np.ascontiguousarray(p.ydiv, dtype=np.int32)), int(norm is not None),
                                      C.c_double(norm if norm is not None else 0.0), _d(fasm), C.byref(neg)))
        return fasm, neg.value

    # ---- thermal-hydraulic channel solve (mod_th.f90 th_upd / th_trans)
    def set_th(self, th):
        """th: the dict of deck.Problem.th_setup()"""
        self._th = th
        stab = np.asfortranarray(th["stab"], dtype=np.float64)
        rc = self.L.adp_set_th(self.h, C.c_double(th["pi"]), C.c_double(th["rf"]), C.c_double(th["rg"]), C.c_double(th["rc"]),
                               C.c_double(th["dia"]), C.c_double(th["dh"]), C.c_double(th["farea"]), C.c_double(th["cflow"]),
                               C.c_double(th["cf"]), C.c_double(th["tin"]), _d(np.ascontiguousarray(th["rpos"])),
                               _d(np.ascontiguousarray(th["rdel"])), int(th["ntem"]), _d(stab))
        return self._chk(rc)

    def set_th_state(self, st):
        keep = {k: np.asfortranarray(st[k], dtype=np.float64) for k in st if st.get(k) is not None}
        g = lambda k: _d(keep[k]) if k in keep else None
        self._chk(self.L.adp_set_th_state(self.h, g("tfm"), g("heatf"), g("ent"), g("ftem"), g("mtem"), g("cden"), g("frate")))

    def th_state(self):
        N = self.N
        st = dict(tfm=np.zeros((N, 13), order="F"), heatf=np.zeros(N), ent=np.zeros(N), ftem=np.zeros(N), mtem=np.zeros(N),
                  cden=np.zeros(N), frate=np.zeros(N))
        self._chk(self.L.adp_get_th_state(self.h, _d(st["tfm"]), _d(st["heatf"]), _d(st["ent"]), _d(st["ftem"]), _d(st["mtem"]),
                                          _d(st["cden"]), _d(st["frate"])))
        return st

    def th_pline(self, pow_, ppow, form=0):
        nf = np.asfortranarray(self._th["node_nf"], dtype=np.float64)
        return self._chk(self.L.adp_th_pline(self.h, C.c_double(pow_), C.c_double(ppow), int(form), _d(nf)))

    def th_upd(self, xpline=None, want_err=True):
        e = C.c_double()
        x = None if xpline is None else np.ascontiguousarray(xpline, dtype=np.float64)
        rc = self._chk(self.L.adp_th_upd(self.h, _d(x), C.byref(e) if want_err else None))
        return rc, e.value

    def th_trans(self, xpline, h):
        x = None if xpline is None else np.ascontiguousarray(xpline, dtype=np.float64)
        return self._chk(self.L.adp_th_trans(self.h, _d(x), C.c_double(h)))

    def integrate(self, s):
        r = C.c_double()
        self._chk(self.L.adp_integrate(self.h, _d(np.ascontiguousarray(s, dtype=np.float64)), C.byref(r)))
        return r.value

    # ---- kernel level
    def sp_matvec(self, g, x):
        v = np.zeros(self.N)
        self._chk(self.L.adp_sp_matvec(self.h, g, _d(np.ascontiguousarray(x)), _d(v)))
        return v

    def bicg(self, imax, g, b, x):
        x = np.array(x, dtype=np.float64)
        self._chk(self.L.adp_bicg(self.h, imax, g, _d(np.ascontiguousarray(b)), _d(x)))
        return x

    def matrix_dia(self):
        a = np.zeros((7, self.N, self.G), order="F")
        self._chk(self.L.adp_get_matrix(self.h, _d(a)))
        return a

    def get_source(self, cmode):
        S = [np.zeros((self.N, self.G), order="F") for _ in range(3)]
        self._chk(self.L.adp_get_source(self.h, cmode, _d(S[0]), _d(S[1]), _d(S[2])))
        return S

    # ---- outputs
    def state(self):
        f0 = np.zeros((self.N, self.G), order="F")
        fs0 = np.zeros(self.N)
        s0 = np.zeros((self.N, self.G), order="F")
        ke = C.c_double()
        self._chk(self.L.adp_get_state(self.h, _d(f0), _d(fs0), _d(s0), C.byref(ke)))
        ser, fer = C.c_double(), C.c_double()
        self.L.adp_get_errors(self.h, C.byref(ser), C.byref(fer))
        return dict(f0=f0, fs0=fs0, s0=s0, Ke=ke.value, ser=ser.value, fer=fer.value)

    def nod(self):
        df = np.zeros((6, self.N, self.G), order="F")
        dn = np.zeros((6, self.N, self.G), order="F")
        self._chk(self.L.adp_get_nod(self.h, _d(df), _d(dn)))
        return df, dn

    def exsrc_arrays(self):
        ex = np.zeros((self.N, self.G), order="F")
        dfis = np.zeros(self.N)
        self._chk(self.L.adp_get_exsrc_arrays(self.h, _d(ex), _d(dfis)))
        return ex, dfis

    @property
    def ndmax(self):
        r = C.c_double()
        self.L.adp_get_ndmax(self.h, C.byref(r))
        return r.value

    def launch_count(self):
        n = C.c_longlong()
        self.L.adp_launch_count(self.h, C.byref(n))
        return n.value

    def bench_kernel(self, what, reps):
        ms = C.c_double()
        self._chk(self.L.adp_bench_kernel(self.h, what, reps, C.byref(ms)))
        return ms.value

    def outer_steps(self, mode, p_first, nsteps):
        ke, ser, fer = C.c_double(), C.c_double(), C.c_double()
        rc = self._chk(self.L.adp_outer_steps(self.h, mode, p_first, nsteps, C.byref(ke), C.byref(ser), C.byref(fer)))
        return rc, ke.value, ser.value, fer.value

    def timer_start(self):
        self._chk(self.L.adp_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._chk(self.L.adp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def lxyz_total(self):
        L = np.zeros((self.N, self.G), order="F")
        self._chk(self.L.adp_lxyz_total(self.h, _d(L)))
        return L

    # ---- time-step glue on the device (mod_trans.f90 callers)
    def save_adjoint(self):
        self._chk(self.L.adp_save_adjoint(self.h))

    def ipden(self):
        self._chk(self.L.adp_ipden(self.h))

    def update_omeg(self, ht, bextr):
        self._chk(self.L.adp_update_omeg(self.h, C.c_double(ht), int(bextr)))

    def begin_time_step(self, ht):
        self._chk(self.L.adp_begin_time_step(self.h, C.c_double(ht)))

    def upden(self, ht):
        self._chk(self.L.adp_upden(self.h, C.c_double(ht)))

    def powtot(self):
        r = C.c_double()
        self._chk(self.L.adp_powtot(self.h, C.byref(r)))
        return r.value

    def reactivity(self, use_sigrp):
        r = C.c_double()
        self._chk(self.L.adp_reactivity(self.h, int(use_sigrp), C.byref(r)))
        return r.value

    # ---- XS update on the device (%XSEC + %CROD decks)
    def set_material_xs(self, p=None):
        p = p or self.p
        a = [np.asfortranarray(x, dtype=np.float64) for x in (p.xsigtr, p.xsiga, p.xnuf, p.xsigf, p.xsigs)]
        self._chk(self.L.adp_set_material_xs(self.h, *[_d(x) for x in a]))

    def set_crod(self, p=None):
        p = p or self.p
        c = p.crod
        ia, ja, _ = p._node_assembly_maps()
        fbmap = np.asfortranarray(c["bmap"][np.ix_(ia, ja)].astype(np.int32))          # (nxx, nyy) column-major
        a = [np.asfortranarray(c[k], dtype=np.float64) for k in ("dsigtr", "dsiga", "dnuf", "dsigf", "dsigs")]
        self._chk(self.L.adp_set_crod(self.h, int(c["nb"]), C.c_double(c["pos0"]), C.c_double(c["ssize"]),
                                      fbmap.ctypes.data_as(_ip), *[_d(x) for x in a]))

    def xs_update(self, bpos=None):
        b = None if bpos is None else np.ascontiguousarray(bpos, dtype=np.float64)
        return self._chk(self.L.adp_xs_update(self.h, _d(b)))           # 0 or STOP_XS_CHECK

    def set_feedback(self, p=None):
        """feedback cards of the deck (p.fbk) -> device tables"""
        p = p or self.p
        for which, key in enumerate(("bcon", "ftem", "mtem", "cden")):
            t = (p.fbk or {}).get(key)
            if t is None:
                continue
            a = [np.asfortranarray(t[k], dtype=np.float64) for k in ("sigtr", "siga", "nuf", "sigf", "sigs")]
            self._chk(self.L.adp_set_feedback(self.h, which, C.c_double(t["ref"]), *[_d(x) for x in a]))

    def xs_update_th(self, bcon, ftem=None, mtem=None, cden=None, bpos=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        ft, mt, cd, bp = f(ftem), f(mtem), f(cden), f(bpos)
        return self._chk(self.L.adp_xs_update_th(self.h, C.c_double(bcon), _d(ft), _d(mt), _d(cd), _d(bp)))   # 0 or STOP_XS_CHECK

    # ---- XS update on the device (%XTAB branch tables)
    def set_xtab(self, p=None):
        """branch tables of the deck (p.xtab, deck.read_xtab_composition) -> device"""
        p = p or self.p
        dims, trod, par, xs, rxs = pack_xtab(p)
        self._chk(self.L.adp_set_xtab(self.h, dims.ctypes.data_as(_ip), trod.ctypes.data_as(_ip), _d(par), _d(xs), _d(rxs)))

    def set_crod_map(self, p=None):
        p = p or self.p
        c = p.crod
        ia, ja, _ = p._node_assembly_maps()
        fbmap = np.asfortranarray(c["bmap"][np.ix_(ia, ja)].astype(np.int32))          # (nxx, nyy) column-major
        self._chk(self.L.adp_set_crod_map(self.h, int(c["nb"]), C.c_double(c["pos0"]), C.c_double(c["ssize"]),
                                          fbmap.ctypes.data_as(_ip)))

    def xs_update_xtab(self, bcon, ftem=None, mtem=None, cden=None, bpos=None):
        """XStab_updt on the device; returns 0 or ADP_STOP_XTAB_RANGE / ADP_STOP_XTAB_NOROD / ADP_STOP_XS_CHECK"""
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        ft, mt, cd, bp = f(ftem), f(mtem), f(cden), f(bpos)
        return self._chk(self.L.adp_xs_update_xtab(self.h, C.c_double(bcon), _d(ft), _d(mt), _d(cd), _d(bp)))

    def get_dc(self):
        dc = np.zeros((self.N, self.G, 6), order="F")
        self._chk(self.L.adp_get_dc(self.h, _d(dc)))
        return dc

    def get_xs(self):
        N, G = self.N, self.G
        out = dict(D=np.zeros((N, G), order="F"), sigr=np.zeros((N, G), order="F"), nuf=np.zeros((N, G), order="F"),
                   sigf=np.zeros((N, G), order="F"), sigs=np.zeros((N, G, G), order="F"))
        self._chk(self.L.adp_get_xs(self.h, _d(out["D"]), _d(out["sigr"]), _d(out["nuf"]), _d(out["sigf"]), _d(out["sigs"])))
        return out
