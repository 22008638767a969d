"""examples/run_deck.py dispatches a deck on its %MODE like ADPRES.f90 and prints the reference-style results.
Its `run()` is written against the solver interface capi.Solver and the oracle share, so the dispatch logic is
exercised here on the CPU with the oracle as back end (the GPU back end: `python examples/run_deck.py <deck>`)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, load_problem


@pytest.fixture(scope="module")
def run_deck():
    spec = importlib.util.spec_from_file_location("run_deck", os.path.join(ROOT, "examples", "run_deck.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _oracle_run(run_deck, name, steps=None):
    from adpres_b200 import thermal
    from oracle import Oracle, th as oth
    p = load_problem(name)
    o = Oracle(p)
    glue = thermal.HostGlue(p, o, oth) if (p.mode == "BCSEARCH" or (p.mode == "RODEJECT" and p.ther is not None)) else None
    lines = []
    return p, run_deck.run(p, o, glue, steps=steps, log=lines.append), lines


def test_forward_adjoint_fixed_source(run_deck):
    p, r, lines = _oracle_run(run_deck, "IAEA3Ds")
    assert r["status"] == 0 and r["outers"] == 129 and "%.6f" % r["keff"] == "1.029082"
    assert abs(r["asm_power"][r["asm_power"] > 0].mean() - 1.0) < 1e-12            # AsmPow normalisation
    assert any("K-EFF = 1.029082" in ln for ln in lines) and any("Radial Power Distribution" in ln for ln in lines)
    p, r, _ = _oracle_run(run_deck, "adjoint")
    assert r["mode"] == "ADJOINT" and abs(r["keff"] - 1.029082) < 2e-5 and "asm_power" not in r
    p, r, _ = _oracle_run(run_deck, "fixed_source")
    assert r["mode"] == "FIXEDSRC" and r["status"] == 0


def test_boron_search_modes(run_deck):
    _, r, lines = _oracle_run(run_deck, "CBCsearch")                                   # cbsearch (no %THER)
    assert abs(r["bcon"] - 1257.32) < 0.05 and "tf_avg" not in r
    _, r, lines = _oracle_run(run_deck, "MOX_P3_HELIOS")                               # cbsearcht, %XTAB
    assert abs(r["bcon"] - 1341.99) < 0.02 and abs(r["tf_avg"] - 560.0) < 0.01
    assert any("CRITICAL BORON CONCENTRATION = 1341.99 ppm" in ln for ln in lines)


def test_rod_ejection_modes(run_deck):
    _, r, _ = _oracle_run(run_deck, "LMW", steps=3)                                    # rod_eject (no %THER)
    assert len(r["trace"]) == 4 and r["trace"][0][3] == 1.0 and r["peak_power"] > 1.0
    _, r, _ = _oracle_run(run_deck, "NEACRP_A1t", steps=3)                             # rod_eject_th
    assert len(r["trace"]) == 4 and abs(r["trace"][0][3] - 1.0e-6) < 1e-18 and r["max_reactivity"] > 0.0
