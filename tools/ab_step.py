#!/usr/bin/env python
"""A/B of library builds on ONE box: ms per outer iteration (BASELINE configs[1], nin = 10, no nodal update inside the
timed steps) and the isolated kernel times, alternating the builds so that box-to-box and thermal drift cancel.
Each build runs in its own process (the library is loaded once per process).
usage: python tools/ab_step.py name=path/to/lib.so [name=path ...] [--reps 3] [--steps 100]"""
import json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

def child(lib, steps):
    sys.path.insert(0, ROOT)
    from adpres_b200 import capi
    capi.LIB_PATH = lib
    import bench
    p = bench.load_c2()
    s = capi.Solver(p, **bench.CTL)
    s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
    s.outer_steps(capi.MODE_FORWARD, 1, 5)
    out = {"steps": []}
    for rep in range(3):
        s.timer_start()
        s.outer_steps(capi.MODE_FORWARD, 51, 49)          # p = 51..99: no nodal update (nupd = 50)
        out["steps"].append(s.timer_stop() / 49)
    out["kern"] = {str(w): s.bench_kernel(w, 20) for w in (4, 0, 1, 2, 3, 5)}
    try:
        out["kern"].update({str(w): s.bench_kernel(w, 20) for w in (14, 10, 11, 12, 13)})     # the multi-rank variants (round 2 builds)
    except Exception:
        pass
    print(json.dumps(out))

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    libs = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    res = {n: [] for n, _ in libs}
    for r in range(reps):
        for n, path in libs:
            o = subprocess.run([sys.executable, __file__, "--child", os.path.abspath(path), "100"], capture_output=True, text=True)
            try:
                res[n].append(json.loads(o.stdout.strip().split("\n")[-1]))
            except Exception:
                print(n, "failed:", o.stderr[-2000:])
    names = {"4": "P", "0": "B", "1": "C", "2": "D", "3": "A", "5": "F", "14": "Pm", "10": "Bm", "11": "Cm", "12": "Dm", "13": "Am"}
    for n, rs in res.items():
        st = sorted(x for r in rs for x in r["steps"])
        print("%-12s ms/step min %.4f med %.4f max %.4f | " % (n, st[0], st[len(st) // 2], st[-1]) +
              "  ".join("%s %.2f" % (names[k], 1e3 * min(r["kern"][k] for r in rs)) for k in names if all(k in r["kern"] for r in rs)))
