#!/usr/bin/env python
"""Run an unchanged ADPRES input deck end to end on the B200 path, without a Fortran compiler:

    python examples/run_deck.py /path/to/ADPRES/smpl/static/IAEA3Ds
    python examples/run_deck.py smpl/static/NEACRP/A1            # critical boron search with TH feedback
    python examples/run_deck.py smpl/transient/LMW --steps 40    # rod ejection, first 40 time steps

The deck is read by the harness's reader (adpres_b200/deck.py), the mode driver of the reference
(mod_control.f90 forward / adjoint / fixedsrc, mod_th.f90 cbsearch / cbsearcht, mod_trans.f90 rod_eject /
rod_eject_th) is the harness restatement, and everything below it -- outer iterations, nodal updates,
cross-section update, thermal-hydraulic channels, time-step glue -- runs on the GPU through the C ABI.
In a deployment the unchanged Fortran drivers make the same calls.

`run()` is written against the solver interface shared by capi.Solver and the test oracle, so the tests
drive the same function on the CPU (tests/test_run_deck.py)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def power_map(p, npow):
    """the radial assembly power map AsmPow prints (north at the top)"""
    a = p.asm_power(npow)
    rows = []
    for j in range(p.ny - 1, -1, -1):
        rows.append(" ".join("%7.4f" % a[i, j] if a[i, j] > 0 else "       " for i in range(p.nx)))
    return "\n".join(rows)


def run(p, solver, glue=None, steps=None, log=print, device_resident=False):
    """Dispatch on %MODE like ADPRES.f90.  `glue`: thermal.HostGlue / DeviceGlue for decks with feedback cards
    (BCSEARCH, RODEJECT with %THER).  Returns a dict of the headline results."""
    from adpres_b200 import thermal, transient
    out = {"mode": p.mode, "nnod": p.nnod, "ng": p.ng}
    t0 = time.perf_counter()
    if p.mode in ("FORWARD", "ADJOINT", "FIXEDSRC"):
        call = {"FORWARD": solver.outer, "ADJOINT": solver.outer_ad, "FIXEDSRC": solver.outer_fs}[p.mode]
        rc, n = call(1)                  # forward / adjoint / fixedsrc call outer*(1) (mod_control.f90:33,73,112)
        st = solver.state()
        out.update(status=rc, outers=n, keff=st["Ke"], ser=st["ser"], fer=st["fer"])
        log(f"  {p.mode}: {n} outer iterations, status {rc}" + ("" if p.mode == "FIXEDSRC" else f", K-EFF = {st['Ke']:.6f}"))
        if p.mode != "ADJOINT":
            rc2, npow = solver.powdis(p.mode == "FIXEDSRC")
            if rc2 == 0:
                out["asm_power"] = p.asm_power(npow)
                log("  Radial Power Distribution\n" + power_map(p, npow))
    elif p.mode == "BCSEARCH":
        search = thermal.cbsearcht if p.ther is not None else thermal.cbsearch
        log("  Itr  Boron Concentration          K-EFF    FLUX REL. ERROR   FISS. SOURCE REL. ERROR")
        bc, rows = search(glue, log=log)
        for r in rows[:2]:
            log(f"{r[0]:3d} {r[1]:10.2f} {r[2]:14.5f}")
        st = solver.state()
        out.update(bcon=bc, keff=st["Ke"], guesses=len(rows))
        log(f"  CRITICAL BORON CONCENTRATION = {bc:.2f} ppm after {len(rows)} guesses (K-EFF {st['Ke']:.6f})")
        if p.ther is not None:
            f = glue.th_fields()
            fuel = p.nuf[:, p.ng - 1] > 0
            out.update(tf_avg=float(f["ftem"][fuel].mean()), tm_max=float(f["mtem"].max()))
            log(f"  AVERAGE DOPPLER TEMPERATURE (unweighted) : {out['tf_avg']:.1f} K;  MAXIMUM MODERATOR TEMPERATURE : {out['tm_max']:.1f} K")
    elif p.mode == "RODEJECT":
        log("  Step  Time(s)  React.($)   Rel. Power")
        if p.ther is not None:
            fn = transient.rod_eject_th_device if device_resident else transient.rod_eject_th
            tr = fn(p, glue, max_steps=steps, log=log)
        elif device_resident:
            tr = transient.rod_eject_device_glue(p, solver, max_steps=steps, log=log, device_xs=True)
        else:
            tr = transient.rod_eject(p, solver, max_steps=steps, log=log)
        for r in tr:
            if p.ther is None:
                log(f"{r[0]:4d} {r[1]:9.3f} {r[2]:10.4f} {r[3]:14.5E}")
        pk = max(tr, key=lambda r: r[3])
        out.update(trace=tr, peak_power=pk[3], peak_time=pk[1], max_reactivity=max(r[2] for r in tr))
        log(f"  PEAK RELATIVE POWER {pk[3]:.5E} AT {pk[1]:.4f} s; MAX REACTIVITY {out['max_reactivity']:.4f} $")
    else:
        raise ValueError(f"MODE {p.mode} IS UNIDENTIFIED")
    out["seconds"] = time.perf_counter() - t0
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("deck")
    ap.add_argument("--steps", type=int, default=None, help="RODEJECT: stop after this many time steps")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args()
    from adpres_b200 import capi, thermal
    from adpres_b200.deck import read_deck
    p = read_deck(args.deck)
    print(f"  deck {args.deck}: MODE {p.mode}, {p.nxx} x {p.nyy} x {p.nzz} mesh, {p.nnod} nodes, {p.ng} groups, {p.nmat} materials")
    s = capi.Solver(p, device=args.device)            # fails loudly without a GPU: there is no CPU fallback
    needs_glue = p.mode == "BCSEARCH" or (p.mode == "RODEJECT" and p.ther is not None)
    glue = thermal.DeviceGlue(p, s) if needs_glue else None
    res = run(p, s, glue, steps=args.steps, device_resident=True)
    print(f"  done in {res['seconds']:.2f} s, {s.launch_count()} kernel launches")


if __name__ == "__main__":
    main()
