#!/usr/bin/env python
"""Generate the fixtures under tests/golden/ from the reference tree (run in the container
that has /root/reference; the GPU box does not).

* ``<deck>.spec.json``   -- the parsed content of a reference sample deck (cross sections,
  geometry, control cards) as ``adpres_b200.deck.Problem.to_spec()`` emits it.  These are
  the *inputs*; they let GPU-side tests run the reference's decks without the tree.
* ``iaea3ds_trace.json`` -- the only ADPRES-produced golden numbers in the reference:
  the terminal trace printed in docs/quick-guides.md:161-191 and the k-eff in
  smpl/static/IAEA3Ds:3-4.  Parsed from the docs file, not typed by hand.
* ``neacrp_bcon.json``, ``mox_bcon.json`` -- critical boron concentrations ADPRES itself found for
  the static NEACRP decks / MOX part 3, read from the %BCON card of the matching transient decks.
* ``header_keff.json``   -- the external-reference k-eff values quoted in the deck headers
  (IAEA2D, BIBLIS, KOEBERG): +-few-pcm sanity values, not ADPRES outputs.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from adpres_b200.deck import read_deck  # noqa: E402

REF = os.environ.get("ADPRES_REFERENCE", "/root/reference")
DECKS = {
    "IAEA3Ds": "smpl/static/IAEA3Ds", "IAEA2D": "smpl/static/IAEA2D", "BIBLIS": "smpl/static/BIBLIS",
    "KOEBERG": "smpl/static/KOEBERG", "DVP": "smpl/static/DVP", "PNM": "smpl/static/PNM",
    "FDM": "smpl/static/FDM", "adjoint": "smpl/static/adjoint", "fixed_source": "smpl/static/fixed_source",
    "LMW": "smpl/transient/LMW",
    "NEACRP_A1": "smpl/static/NEACRP/A1", "NEACRP_A2": "smpl/static/NEACRP/A2", "NEACRP_B1": "smpl/static/NEACRP/B1",
    "NEACRP_B2": "smpl/static/NEACRP/B2", "NEACRP_C1": "smpl/static/NEACRP/C1", "NEACRP_C2": "smpl/static/NEACRP/C2",
    "NEACRP_A1t": "smpl/transient/NEACRP/A1t",
    "MOX_ARO": "smpl/static/MOX/part1_aro_helios", "MOX_ARI": "smpl/static/MOX/part1_ari_helios",
    "FDM_1D": "smpl/static/FDM-1D", "CBCsearch": "smpl/static/CBCsearch", "MOX_1B_A1": "smpl/static/MOX/Part1b/aroA1",
    "MOX_1D_E7": "smpl/static/MOX/Part1d/ariE7", "MOX_ARO_SERPENT": "smpl/static/MOX/part1_aro_serpent",
    "MOX_ARI_SERPENT": "smpl/static/MOX/part1_ari_serpent",
    # %XTAB decks: the spec carries the branch tables of the compositions the deck selects
    "MOX_P2_HELIOS": "smpl/static/MOX/part2_helios", "MOX_P3_HELIOS": "smpl/static/MOX/part3_helios",
    "MOX_P3_SERPENT": "smpl/static/MOX/part3_serpent", "MOX_P4_HELIOS": "smpl/transient/MOX/part4_helios",
}


def main():
    for name, rel in DECKS.items():
        p = read_deck(os.path.join(REF, rel))
        with open(os.path.join(HERE, name + ".spec.json"), "w") as fh:
            json.dump(p.to_spec(), fh, separators=(",", ":"))
    # ---- NEACRP: the boron concentration each TRANSIENT deck starts from is the critical boron the
    # reference found for the matching static deck (%BCON first number, smpl/transient/NEACRP/*t)
    bcon = {}
    for case in ("A1", "A2", "B1", "B2", "C1", "C2"):
        lines = open(os.path.join(REF, "smpl/transient/NEACRP", case + "t")).read().splitlines()
        i = next(k for k, ln in enumerate(lines) if ln.strip().upper().startswith("%BCON"))
        bcon[case] = float(lines[i + 1].split()[0])
    with open(os.path.join(HERE, "neacrp_bcon.json"), "w") as fh:
        json.dump({"source": "first number of the %BCON card of smpl/transient/NEACRP/<case>t", "ppm": bcon}, fh)
    # ---- MOX/UO2 benchmark: part 4 (rod ejection from hot zero power) starts from the critical boron the
    # reference found for part 3 -- same rule as NEACRP
    mox = {}
    for lib in ("helios", "serpent"):
        lines = open(os.path.join(REF, "smpl/transient/MOX", "part4_" + lib)).read().splitlines()
        i = next(k for k, ln in enumerate(lines) if ln.strip().upper().startswith("%BCON"))
        mox["P3_" + lib.upper()] = float(lines[i + 1].split()[0])
    with open(os.path.join(HERE, "mox_bcon.json"), "w") as fh:
        json.dump({"source": "first number of the %BCON card of smpl/transient/MOX/part4_<library>", "ppm": mox}, fh)
    # ---- docs trace
    text = open(os.path.join(REF, "docs/quick-guides.md")).read()
    rows = []
    for m in re.finditer(r"^\s*(\d+)\s+(\d\.\d{6})\s+(\d\.\d{5}E[+-]\d\d)\s+(\d\.\d{5}E[+-]\d\d)\s*$", text, re.M):
        rows.append([int(m.group(1)), m.group(2), m.group(3), m.group(4)])
    m = re.search(r"MAX\. CHANGE IN NODAL COUPLING COEF\.=\s+(\S+) AT NODE I =\s*(\d+), J =\s*(\d+), K =\s*(\d+)", text)
    nodal = {"before_iter": 22, "ndmax": m.group(1), "i": int(m.group(2)), "j": int(m.group(3)), "k": int(m.group(4))}
    keff = re.search(r"MULTIPLICATION EFFECTIVE \(K-EFF\) =\s+(\d\.\d+)", text).group(1)
    head = open(os.path.join(REF, "smpl/static/IAEA3Ds")).read()
    deck_keff = re.search(r"ADPRES K-EFF\s*:\s*(\d\.\d+)", head).group(1)
    timing = {k: float(v) for k, v in re.findall(r"(Input reading time|XSEC processing time|CMFD time|Nodal update time|Total time)\s*:\s*([\d.]+)", text)}
    with open(os.path.join(HERE, "iaea3ds_trace.json"), "w") as fh:
        json.dump({"source": "docs/quick-guides.md:161-201, smpl/static/IAEA3Ds:3-4", "rows": rows, "nodal_update": nodal,
                   "keff": keff, "deck_header_keff": deck_keff, "outers": rows[-1][0],
                   "extrapolated_before": [5, 20], "cpu_seconds": timing}, fh, indent=1)
    hk = {}
    for name in ("IAEA2D", "BIBLIS", "KOEBERG"):
        h = open(os.path.join(REF, "smpl/static", name)).read()
        hk[name] = float(re.search(r"K-EFF REF\s*:\s*(\d\.\d+)", h).group(1))
    with open(os.path.join(HERE, "header_keff.json"), "w") as fh:
        json.dump(hk, fh, indent=1)
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
