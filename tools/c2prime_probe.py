#!/usr/bin/env python
"""Robustness of the iteration control on the C2' mesh (170 x 170 x 418, 0.91 cm planes): full eigenvalue solves with
several (nin, nupd) cards and two summation orders (library builds), reporting STOP code, outer count, k-eff, the largest
ndmax of the nodal updates and the largest source error after the tenth outer iteration.
usage: python tools/c2prime_probe.py name=lib.so [name=lib.so ...] -- nin,nupd [nin,nupd ...]"""
import json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

def child(lib, cards):
    sys.path.insert(0, ROOT)
    from adpres_b200 import capi
    capi.LIB_PATH = os.path.abspath(lib)
    import bench
    from adpres_b200.deck import Problem
    with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
        p = Problem.from_spec(json.load(fh)).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[22] * 19)
    for nin, nupd in cards:
        s = capi.Solver(p, **dict(bench.CTL, nin=nin, nupd=nupd, nout=3000))
        s.enable_trace()
        try:
            rc, n = s.outer(1)
            ke = s.state()["Ke"]
        except Exception as e:
            rc, n, ke = -99, 0, float("nan")
            print("   error:", e)
        nd = [x[1] for x in s.trace_nodal]
        ser = [r[2] for r in s.trace_rows[10:]]
        print("   nin=%2d nupd=%3d : rc=%d outers=%4d keff=%.10f  max ndmax %.3e (first %.3e, %d updates)  max ser after p=10 %.3e" %
              (nin, nupd, rc, n, ke, max(nd) if nd else 0, nd[0] if nd else 0, len(nd), max(ser) if ser else 0), flush=True)
        s.close()

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], [tuple(int(x) for x in c.split(",")) for c in sys.argv[3:]])
    else:
        k = sys.argv.index("--")
        for spec in sys.argv[1:k]:
            name, lib = spec.split("=")
            print(name, flush=True)
            subprocess.run([sys.executable, __file__, "--child", lib] + sys.argv[k + 1:])
