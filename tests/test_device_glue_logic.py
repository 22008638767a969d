"""The device-glue drivers of the harness (thermal.DeviceGlue, transient.rod_eject_th_device) on the CPU: a
stand-in for capi.Solver (tests/fake_device.py: C oracle + numpy XS update + oracle/th.py behind the Solver
interface, arrays kept resident like on the device) runs the same driver code the GPU tests run.  This covers the
Python side of the device-resident paths without a GPU, and -- the stand-in's precursor / delayed-source /
reactivity arithmetic being the C oracle's, the host-glue drivers' being numpy -- checks the two restatements of
iPden / uPden / get_exsrc (both bxtab branches) against each other at driver level."""
import numpy as np
import pytest

from conftest import load_problem


def _both(name, steps, tight=True):
    from adpres_b200 import thermal, transient
    from oracle import Oracle, th as oth
    from fake_device import FakeDeviceSolver
    ps = []
    for _ in range(2):
        p = load_problem(name)
        if tight:
            p.serc = p.ferc = 1e-9
            p.nout = 5000
        ps.append(p)
    host = transient.rod_eject_th(ps[0], thermal.HostGlue(ps[0], Oracle(ps[0]), oth), max_steps=steps)
    fake = FakeDeviceSolver(ps[1])
    dev = transient.rod_eject_th_device(ps[1], thermal.DeviceGlue(ps[1], fake), max_steps=steps)
    return host, dev, fake


@pytest.mark.parametrize("name,steps", [("NEACRP_A1t", 12), ("MOX_P4_HELIOS", 4)])
def test_device_glue_rod_ejection_equals_host_glue(name, steps):
    host, dev, fake = _both(name, steps)
    assert len(host) == len(dev) == steps + 1
    for a, b in zip(host, dev):
        assert a[0] == b[0] and a[1] == b[1]
        # both run the time steps to ser, fer < 1e-9; the exit iteration may differ by a few (round-off of the two glue
        # arithmetics), so the traces agree at the level of that tolerance
        assert abs(a[2] - b[2]) < 1e-6, (a, b)                       # reactivity [$]
        assert abs(a[3] / b[3] - 1.0) < 1e-6, (a, b)                 # relative power
        assert a[5] == b[5]                                          # (the exit iteration itself depends on round-off: DESIGN.md 2)
        assert abs(a[6] / b[6] - 1.0) < 1e-8                         # max fuel centreline temperature
    assert fake.kin_xtab == (name == "MOX_P4_HELIOS")
    assert ("set_xtab" in fake.calls) == (name == "MOX_P4_HELIOS")


@pytest.mark.parametrize("name,gold", [("NEACRP_A1", 560.53), ("MOX_P3_HELIOS", 1341.99)])
def test_device_glue_boron_search(name, gold):
    from adpres_b200 import thermal
    from fake_device import FakeDeviceSolver
    p = load_problem(name)
    g = thermal.DeviceGlue(p, FakeDeviceSolver(p))
    bc, rows = thermal.cbsearcht(g)
    assert abs(bc - gold) < 0.05
    f = g.th_fields()
    assert set(f) == {"ftem", "mtem", "cden"} and f["ftem"].shape == (p.nnod,)


def test_device_glue_boron_search_without_th():
    """CBCsearch: no %THER card -> DeviceGlue passes the card values of the TH fields with every update"""
    from adpres_b200 import thermal
    from fake_device import FakeDeviceSolver
    p = load_problem("CBCsearch")
    g = thermal.DeviceGlue(p, FakeDeviceSolver(p))
    assert g.th is None and g._first is not None
    bc, rows = thermal.cbsearch(g)
    assert abs(bc - 1257.32) < 0.05


def test_device_glue_raises_the_reference_stops():
    from adpres_b200 import thermal
    from fake_device import FakeDeviceSolver
    p = load_problem("MOX_P3_HELIOS")
    g = thermal.DeviceGlue(p, FakeDeviceSolver(p))
    with pytest.raises(thermal.StopError, match="OUT OF THE RANGE"):
        g.xs_update(2500.0)
