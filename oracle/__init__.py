"""ctypes binding of the CPU oracle (oracle/adpres_oracle.c).

TEST INFRASTRUCTURE.  Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; the product
(``adpres_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "adpres_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_integrate.restype = C.c_double
        _lib.orc_powtot.restype = C.c_double
        _lib.orc_reactivity.restype = C.c_double
        _lib.orc_ndmax.restype = C.c_double
    return _lib


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and (a.flags.f_contiguous or a.flags.c_contiguous)
    return a.ctypes.data_as(_dp)


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


class Oracle:
    """The reference algorithm on the CPU, driven with the arrays of a ``deck.Problem``."""

    def __init__(self, p, nupd=None, nout=None, nin=None, nac=None, serc=None, ferc=None, kern=None):
        self.L = lib()
        self.p = p
        self.h = C.c_void_p(self.L.orc_create())
        self.N, self.G = p.nnod, p.ng
        self.L.orc_set_geometry(self.h, p.nxx, p.nyy, p.nzz, p.nnod, p.ng, p.nmat, _i(p.ix), _i(p.iy), _i(p.iz),
                                _i(p.ystag_smin), _i(p.ystag_smax), _i(p.xstag_smin), _i(p.xstag_smax),
                                _d(p.xdel), _d(p.ydel), _d(p.zdel), _i(p.bc), _i(p.mat))
        self.set_xs()
        self.set_control(nout=nout, nin=nin, nac=nac, nupd=nupd, serc=serc, ferc=ferc, kern=kern)

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    # ---- inputs
    def set_xs(self, **kw):
        p = self.p
        g = lambda k: _d(np.asfortranarray(kw.get(k, getattr(p, k))))
        self.L.orc_set_xs(self.h, g("D"), g("sigr"), g("nuf"), g("sigf"), g("sigs"), g("chi"), g("dc"), g("exsrc"))

    def set_control(self, nout=None, nin=None, nac=None, nupd=None, serc=None, ferc=None, kern=None):
        p = self.p
        v = lambda a, b: b if a is None else a
        self.ctl = dict(nout=v(nout, p.nout), nin=v(nin, p.nin), nac=v(nac, p.nac), nupd=v(nupd, p.nupd),
                        serc=v(serc, p.serc), ferc=v(ferc, p.ferc), kern=v(kern, p.kern))
        c = self.ctl
        self.L.orc_set_control(self.h, c["nout"], c["nin"], c["nac"], c["nupd"], C.c_double(c["serc"]),
                               C.c_double(c["ferc"]), c["kern"], int(p.mode == "FIXEDSRC"))

    def set_state(self, f0=None, fs0=None, Ke=1.0):
        self.L.orc_set_state(self.h, _d(f0), _d(fs0), C.c_double(Ke))

    def set_kinetics(self, ibeta, lamb, velo, tbeta, sth, bth):
        self.L.orc_set_kinetics(self.h, _d(np.ascontiguousarray(ibeta)), _d(np.ascontiguousarray(lamb)),
                                _d(np.ascontiguousarray(velo)), _d(np.ascontiguousarray(tbeta)),
                                C.c_double(sth), C.c_double(bth))

    def set_kinetics_xtab(self, mibeta, mlamb, mvelo, tbeta, sth, bth):
        """%XTAB decks: per-material iBeta / lamb (nmat, 6) and velo (nmat, ng) -- rows = materials"""
        self.L.orc_set_kinetics_xtab(self.h, _d(np.ascontiguousarray(mibeta)), _d(np.ascontiguousarray(mlamb)),
                                     _d(np.ascontiguousarray(mvelo)), _d(np.ascontiguousarray(tbeta)),
                                     C.c_double(sth), C.c_double(bth))

    def set_transient(self, c0=None, ft=None, fst=None, omeg=None, sigrp=None, L=None):
        self.L.orc_set_transient(self.h, _d(c0), _d(ft), _d(fst), _d(omeg), _d(sigrp), _d(L))

    # ---- hot-path entry points (names of the reference procedures)
    def _run(self, fn, *args):
        n = C.c_int(0)
        rc = fn(self.h, *args, C.byref(n))
        return rc, n.value

    def init_flux(self, adjoint=False):
        self.L.orc_init_flux(self.h, int(adjoint))

    def outer(self, popt=1):
        return self._run(self.L.orc_outer, popt)

    def outer_fs(self, popt=1):
        return self._run(self.L.orc_outer_fs, popt)

    def outer_ad(self, popt=1):
        return self._run(self.L.orc_outer_ad, popt)

    def outer_th(self, maxn):
        return self._run(self.L.orc_outer_th, maxn)

    def outer_tr(self, ht):
        maxi, n = C.c_int(0), C.c_int(0)
        rc = self.L.orc_outer_tr(self.h, C.c_double(ht), C.byref(maxi), C.byref(n))
        return rc, bool(maxi.value), n.value

    def matrix_setup(self, opt):
        self.L.orc_matrix_setup(self.h, opt)

    def nodal_upd(self, nmode):
        return self.L.orc_nodal_upd(self.h, nmode)

    def sp_matvec(self, g, x):
        v = np.empty(self.N)
        self.L.orc_sp_matvec(self.h, g, _d(np.ascontiguousarray(x)), _d(v))
        return v

    def bicg(self, imax, g, b, x):
        x = np.array(x, dtype=np.float64)
        self.L.orc_bicg(self.h, imax, g, _d(np.ascontiguousarray(b)), _d(x))
        return x

    def tsrc(self, g, keff, kind="fwd"):
        bs = np.empty(self.N)
        if kind == "fwd":
            self.L.orc_tsrc(self.h, g, C.c_double(keff), _d(bs))
        elif kind == "adj":
            self.L.orc_tsrc_ad(self.h, g, C.c_double(keff), _d(bs))
        else:
            self.L.orc_tsrc_tr(self.h, g, _d(bs))
        return bs

    def fsrc(self, adjoint=False):
        fs = np.empty(self.N)
        (self.L.orc_fsrc_ad if adjoint else self.L.orc_fsrc)(self.h, _d(fs))
        return fs

    def integrate(self, s):
        return self.L.orc_integrate(self.h, _d(np.ascontiguousarray(s)))

    def powdis(self, fixedsrc=None):
        """PowDis; `fixedsrc` is accepted for interface parity with capi.Solver.powdis (the oracle knows the mode)"""
        pw = np.empty(self.N)
        rc = self.L.orc_powdis(self.h, _d(pw))
        return rc, pw

    def get_exsrc(self, ht):
        self.L.orc_get_exsrc(self.h, C.c_double(ht))

    def get_source(self, cmode):
        self.L.orc_get_source(self.h, cmode)
        S = [np.empty((self.N, self.G), order="F") for _ in range(3)]
        self.L.orc_get_sources(self.h, _d(S[0]), _d(S[1]), _d(S[2]))
        return S

    def powtot(self, fx):
        return self.L.orc_powtot(self.h, _d(np.asfortranarray(fx)))

    def ipden(self):
        self.L.orc_ipden(self.h)

    def upden(self, ht):
        self.L.orc_upden(self.h, C.c_double(ht))

    def reactivity(self, af, sigrp):
        return self.L.orc_reactivity(self.h, _d(np.asfortranarray(af)), _d(np.asfortranarray(sigrp)))

    # ---- outputs
    def state(self):
        f0 = np.empty((self.N, self.G), order="F")
        fs0 = np.empty(self.N)
        s0 = np.empty((self.N, self.G), order="F")
        ke, ser, fer = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_get_state(self.h, _d(f0), _d(fs0), _d(s0), C.byref(ke), C.byref(ser), C.byref(fer))
        return dict(f0=f0, fs0=fs0, s0=s0, Ke=ke.value, ser=ser.value, fer=fer.value)

    def nod(self):
        df = np.empty((6, self.N, self.G), order="F")
        dn = np.empty((6, self.N, self.G), order="F")
        self.L.orc_get_nod(self.h, _d(df), _d(dn))
        return df, dn

    def set_nod_dn(self, dn):
        self.L.orc_set_nod_dn(self.h, _d(np.asfortranarray(dn)))

    def matrix_dia(self):
        a = np.empty((7, self.N, self.G), order="F")
        self.L.orc_get_matrix_dia(self.h, _d(a))
        return a

    def transient(self):
        c0 = np.empty((self.N, 6), order="F")
        ex = np.empty((self.N, self.G), order="F")
        dfis = np.empty(self.N)
        L = np.empty((self.N, self.G), order="F")
        self.L.orc_get_transient(self.h, _d(c0), _d(ex), _d(dfis), _d(L))
        return dict(c0=c0, exsrc=ex, dfis=dfis, L=L)

    def abefgh(self, n, u):
        out = np.empty(6 * self.G)
        self.L.orc_abefgh(self.h, n, u, _d(out))
        return out.reshape(6, self.G)

    @property
    def ndmax(self):
        return self.L.orc_ndmax(self.h)

    @ndmax.setter
    def ndmax(self, v):
        self.L.orc_set_ndmax(self.h, C.c_double(v))

    def ndloc(self):
        a = (C.c_int * 3)()
        self.L.orc_get_ndloc(self.h, a)
        return tuple(a)

    def times(self):
        a, b = C.c_double(), C.c_double()
        self.L.orc_times(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def reset_times(self):
        self.L.orc_reset_times(self.h)

    def trace(self):
        n = 200000
        ke, ser, fer = np.empty(n), np.empty(n), np.empty(n)
        m = self.L.orc_trace(self.h, n, _d(ke), _d(ser), _d(fer))
        return ke[:m].copy(), ser[:m].copy(), fer[:m].copy()

    def trace_times(self):
        """(wall clock, cumulative CMFD cpu time, cumulative nodal cpu time) at the end of every iteration of the last outer*()"""
        n = 200000
        w, f, d = np.empty(n), np.empty(n), np.empty(n)
        m = self.L.orc_trace_times(self.h, n, _d(w), _d(f), _d(d))
        return w[:m].copy(), f[:m].copy(), d[:m].copy()

    def nodal_trace(self):
        n = 4096
        p, im, jm, km = (np.empty(n, dtype=np.int32) for _ in range(4))
        nd = np.empty(n)
        m = min(self.L.orc_nodal_trace(self.h, n, _i(p), _d(nd), _i(im), _i(jm), _i(km)), n)
        return [(int(p[q]), float(nd[q]), int(im[q]), int(jm[q]), int(km[q])) for q in range(m)]

    def extrp_trace(self):
        n = 65536
        p = np.empty(n, dtype=np.int32)
        m = min(self.L.orc_extrp_trace(self.h, n, _i(p)), n)
        return p[:m].tolist()
