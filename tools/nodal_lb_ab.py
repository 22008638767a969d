#!/usr/bin/env python
"""G = 2 nodal update on BASELINE configs[1] with library builds that differ in the launch bounds of the node-direction /
surfaces kernels (tools/build_variant.sh name -DNODAL_LB_ND=n -DNODAL_LB_SF=m).  usage: python tools/nodal_lb_ab.py name=lib.so ..."""
import json, os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

def child(lib):
    sys.path.insert(0, ROOT)
    from adpres_b200 import capi
    capi.LIB_PATH = os.path.abspath(lib)
    import bench
    p = bench.load_c2()
    s = capi.Solver(p, **bench.CTL)
    s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
    s.outer_steps(capi.MODE_FORWARD, 1, 3)
    print(json.dumps([s.bench_kernel(7, 5) for _ in range(3)] + [s.bench_kernel(6, 5)]))

if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for spec in sys.argv[1:]:
            name, lib = spec.split("=")
            o = subprocess.run([sys.executable, __file__, "--child", lib], capture_output=True, text=True)
            try:
                r = json.loads(o.stdout.strip().split("\n")[-1])
                print("%-10s nodal update %s ms   (source kernel %.3f ms)" % (name, " ".join("%.3f" % x for x in r[:3]), r[3]), flush=True)
            except Exception:
                print(name, "failed", o.stderr[-1500:])
