"""examples/shim_replay.c -- the C stand-in that performs the Fortran shim's call sequence (gpu_init, gpu_push_inputs with
its upload masks, the outer loop with nodal_upd and the exit test, gpu_pull_results with the AoS repack of nod) -- builds
against include/adpres_b200.h alone, and on a GPU its results equal the CPU oracle's for the same calls
(reference callers: forward / adjoint, mod_control.f90:21-44,61-80; rod_eject, mod_trans.f90:50-95)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_problem


def _build(tmp_path):
    from adpres_b200 import capi
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    capi.load()
    libdir = os.path.join(ROOT, "adpres_b200")
    exe = str(tmp_path / "shim_replay")
    cmd = [gcc, "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "shim_replay.c"), "-L", libdir, "-ladpres_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _write_spec(path, p, kind, popt, rodeject, nout=None):
    kern = {" FDM": 0, " PNM": 1, "SANM": 2}.get(p.kern, p.kern) if isinstance(p.kern, str) else int(p.kern)
    hdr = [p.nxx, p.nyy, p.nzz, p.nnod, p.ng, p.nmat, nout or p.nout, p.nin, p.nac, p.nupd, kern] + [int(b) for b in p.bc] + \
          [kind, popt, rodeject]
    with open(path, "wb") as fh:
        fh.write(struct.pack("<20i", *hdr))
        fh.write(struct.pack("<2d", p.serc, p.ferc))
        for a in (p.ix, p.iy, p.iz, p.ystag_smin, p.ystag_smax, p.xstag_smin, p.xstag_smax, p.mat):
            fh.write(np.ascontiguousarray(a, dtype=np.int32).tobytes())
        for a in (p.xdel, p.ydel, p.zdel, p.D, p.sigr, p.nuf, p.sigf, p.sigs, p.chi, p.dc, p.exsrc):
            fh.write(np.asfortranarray(a, dtype=np.float64).tobytes(order="F"))


def _read_out(path, p, ncalls, rodeject):
    raw = open(path, "rb").read()
    N, G = p.nnod, p.ng
    off = 0
    niter = struct.unpack_from("<%di" % ncalls, raw, off); off += 4 * ncalls
    status, = struct.unpack_from("<i", raw, off); off += 4
    ke, ser, fer, ndmax = struct.unpack_from("<4d", raw, off); off += 32
    f0 = np.frombuffer(raw, dtype=np.float64, count=N * G, offset=off).reshape((N, G), order="F"); off += 8 * N * G
    fs0 = np.frombuffer(raw, dtype=np.float64, count=N, offset=off); off += 8 * N
    pw = np.frombuffer(raw, dtype=np.float64, count=N, offset=off); off += 8 * N
    nod = None
    if rodeject:
        nod = np.frombuffer(raw, dtype=np.float64, count=12 * N * G, offset=off).reshape((G, N, 12))   # nod(n,g): [g][n]{df(6),dn(6)}
    return dict(niter=niter, status=status, Ke=ke, ser=ser, fer=fer, ndmax=ndmax, f0=f0, fs0=fs0, pw=pw, nod=nod)


def test_shim_replay_builds_as_c99_against_the_header(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "usage" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("deck,kind", [("IAEA3Ds", 0), ("adjoint", 2), ("fixed_source", 1)])
def test_shim_replay_matches_oracle(tmp_path, deck, kind):
    """two consecutive outer*() calls through the shim's sequence (the second one exercises the reduced upload mask and
    the not-first path) against the oracle making the same two calls"""
    from oracle import Oracle
    exe = _build(tmp_path)
    p = load_problem(deck)
    spec, out = str(tmp_path / "spec.bin"), str(tmp_path / "out.bin")
    _write_spec(spec, p, kind, 1, 1)
    r = subprocess.run([exe, spec, out, "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    res = _read_out(out, p, 2, True)
    o = Oracle(p)
    fn = (o.outer, o.outer_fs, o.outer_ad)[kind]
    its = []
    for _ in range(2):
        rc, n = fn(1)
        assert rc == 0
        its.append(n)
    assert res["status"] == 0 and list(res["niter"]) == its, (res["niter"], its)
    st = o.state()
    if kind != 1:
        assert abs(res["Ke"] - st["Ke"]) < 1e-8
    assert np.abs(res["f0"] - st["f0"]).max() / np.abs(st["f0"]).max() < 1e-7
    assert np.abs(res["fs0"] - st["fs0"]).max() / max(np.abs(st["fs0"]).max(), 1e-300) < 1e-7
    rc, pw = o.powdis()
    assert np.abs(res["pw"] - pw).max() < 1e-9
    df, dn = o.nod()                                   # (6, N, G)
    nod = res["nod"]
    for g in range(p.ng):
        assert np.allclose(nod[g, :, :6], df[:, :, g].T, rtol=1e-12, atol=0)
        assert np.abs(nod[g, :, 6:] - dn[:, :, g].T).max() < 1e-8
