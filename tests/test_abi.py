"""CPU-only: the C-ABI shared library loads and exports every symbol include/adpres_b200.h
declares; without a GPU the product fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from adpres_b200 import capi

HEADER = os.path.join(ROOT, "include", "adpres_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(adp_[a-z0-9_]+)\s*\(", text)) - {"adp_trace_fn"})


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/adpres_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == names


def test_version_string():
    assert b"sm_100a" in capi.load().adp_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = capi.load().adp_create(ctypes.byref(h), 0)
    assert rc < 0 and not h.value
    assert b"no CPU fallback" in capi.load().adp_last_error(None)


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (only tests, smoke and the bench baseline may)."""
    pkg = os.path.join(ROOT, "adpres_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_header_is_plain_c_and_a_c_client_links(tmp_path):
    """The boundary is a C ABI: include/adpres_b200.h compiles as C99 (no C++ or torch types in
    the signatures) and a C program links against the library through it alone.  Without a GPU
    the client reports the loud adp_create failure (exit code 3); with one it runs a usage-error
    round trip (exit code 0)."""
    import shutil
    import subprocess
    import torch
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    capi.load()                                   # builds the library if needed
    libdir = os.path.join(ROOT, "adpres_b200")
    exe = str(tmp_path / "abi_probe")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "abi_probe.c"), "-L", libdir, "-ladpres_b200", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert "sm_100a" in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0 and "cross sections not set" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


def test_missing_library_fails_loudly(tmp_path):
    """no silent fallback when the CUDA library has not been built: capi.load() raises"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from adpres_b200 import capi\n"
            "try:\n    capi.load()\nexcept RuntimeError as e:\n    print('RAISED', e)\n" % ROOT)
    env = dict(os.environ, ADPRES_B200_LIB=str(tmp_path / "nowhere" / "libadpres_b200.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert "RAISED" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr
