"""The harness's deck reader (adpres_b200/deck.py) restates the data preparation of mod_io.f90; these
tests cover its input conventions and the reference's input checks on small hand-written decks (CPU)."""
import os

import numpy as np
import pytest

from adpres_b200.deck import parse_deck, read_xtab_composition, _strip_comments, _tokens

MINI = """
! a 2 x 2 x 2 bare cube, one material      (comment lines and trailing comments are dropped)
%MODE
FORWARD
%XSEC
1 1              ! ng, nmat
0.3 0.02 0.03 0.03 1.0 0.0
%GEOM
2 2 2
2*10.0           ! n*v repeats a value, as in Fortran list-directed input
2*1
10.0, 10.0       ! commas separate too
1 1
2*10.0
2*1
1
1 1
1 1
1 1
{bc}
"""


def test_comment_stripping_and_repeat_counts():
    assert _strip_comments("a b ! c\n\n   \n! only a comment\n x ") == ["a b", "x"]
    assert _strip_comments("* INPUT\n 1 1 * ADF\n2", mark="*") == ["1 1", "2"]
    assert _tokens("3*1.5, 2 4*7") == ["1.5", "1.5", "1.5", "2", "7", "7", "7", "7"]


def test_minimal_deck_and_defaults():
    p = parse_deck(MINI.format(bc="0 0 0 0 0 0"))
    assert (p.nnod, p.npl, p.ng, p.nmat, p.kern) == (8, 4, 1, 1, 2)
    assert (p.nout, p.nin, p.serc, p.ferc, p.nac) == (500, 2, 1e-5, 1e-5, 5)         # mod_data.f90:78-86
    assert p.nupd == 3                                                                # ceiling((2 + 2 + 2) / 2.5)
    assert np.array_equal(p.ix, [1, 2, 1, 2, 1, 2, 1, 2]) and np.array_equal(p.iz, [1, 1, 1, 1, 2, 2, 2, 2])
    assert np.allclose(p.D, 1.0 / 0.9) and np.allclose(p.sigr, 0.02) and np.all(p.dc == 1.0) and np.all(p.vdel == 1000.0)


def test_reference_input_checks():
    with pytest.raises(ValueError, match="MODE"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0").replace("%MODE\nFORWARD", ""))
    with pytest.raises(ValueError, match="XSEC OR %XTAB"):
        parse_deck("%MODE\nFORWARD\n%GEOM\n1 1 1\n")
    with pytest.raises(ValueError, match="GEOM"):
        parse_deck("%MODE\nFORWARD\n%XSEC\n1 1\n0.3 0.02 0.03 0.03 1.0 0.0\n")
    with pytest.raises(ValueError, match="boundary condition"):
        parse_deck(MINI.format(bc="0 0 3 0 0 0"))
    with pytest.raises(ValueError, match="sigtr"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0").replace("0.3 0.02", "0.0 0.02"))
    with pytest.raises(ValueError, match="greater than number of materials"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0").replace("1\n1 1\n1 1\n1 1\n", "1\n1 1\n1 2\n1 1\n"))
    with pytest.raises(ValueError, match="GREATER THAN NUMBER OF PLANAR"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0").replace("1\n1 1\n1 1\n1 1\n", "1\n1 2\n1 1\n1 1\n"))
    with pytest.raises(ValueError, match="Zero material"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0").replace("2 2 2\n", "2 2 2\n").replace("1\n1 1\n1 1\n1 1\n", "2\n1 2\n1 1\n1 1\n1 1\n1 0\n"))
    with pytest.raises(ValueError, match="THETA"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0") + "%THET\n1.5\n")
    with pytest.raises(ValueError, match="NODAL KERNEL"):
        parse_deck(MINI.format(bc="0 0 0 0 0 0") + "%KERN\nNEM\n")
    with pytest.raises(ValueError, match="expected 6 values"):
        parse_deck(MINI.format(bc="0 0 0 0 0"))


def test_cards_iter_kern_thet():
    p = parse_deck(MINI.format(bc="2 2 2 2 1 1") + "%ITER\n300 4 1.e-6 1.d-6 10 7 25 30\n%KERN\n pnm\n%THET\n0.5\n")
    assert (p.nout, p.nin, p.serc, p.ferc, p.nac, p.nupd, p.th_niter, p.nth, p.biter) == (300, 4, 1e-6, 1e-6, 10, 7, 25, 30, 1)
    assert p.kern == 1 and p.sth == 0.5 and p.bth == 1.0
    assert list(p.bc) == [2, 2, 2, 2, 1, 1]


def test_file_indirection(tmp_path):
    (tmp_path / "xs.inc").write_text("1 1   ! ng nmat\n0.3 0.02 0.03 0.03 1.0 0.0\n")
    deck = MINI.format(bc="0 0 0 0 0 0").replace("1 1              ! ng, nmat\n0.3 0.02 0.03 0.03 1.0 0.0", "FILE /nowhere/xs.inc")
    p = parse_deck(deck, base_dir=str(tmp_path))          # absolute path of another machine -> looked up next to the deck
    assert p.nnod == 8 and np.allclose(p.sigr, 0.02)


XTAB_LIB = """* Input Control
1  1    * ADF, CROD
*  dens boron ftem mtem
2 1 1 2
0.6 0.8
500. 600.
* composition 1 : transport
{c1}
* composition 2
{c2}
"""


def _composition(base):
    """ng = 1: per (un)rodded set 5 tables (sigtr, siga, nuf, sigf, adf) + 1 scattering table, each ng*nb*nf*nm = 2 records
    of nd = 2 values; then chi, 1/v, lambda, beta"""
    rec = []
    for rod in (0.0, 100.0):
        for table in range(6):
            for v in range(2):
                rec.append("%g %g" % (base + rod + 10 * table + v, base + rod + 10 * table + v + 0.5))
    rec += ["1.0", "2.0e-6", "0.01 0.02 0.03 0.04 0.05 0.06", "1e-4 2e-4 3e-4 4e-4 5e-4 6e-4"]
    return "\n".join(rec)


def test_xtab_library_layout_and_composition_skipping():
    lines = _strip_comments(XTAB_LIB.format(c1=_composition(1.0), c2=_composition(1000.0)), mark="*")
    t1, t2 = read_xtab_composition(lines, 1, 1), read_xtab_composition(lines, 2, 1)
    assert (t1["nd"], t1["nb"], t1["nf"], t1["nm"], t1["tadf"], t1["trod"]) == (2, 1, 1, 2, 1, 1)
    assert t1["pb"].tolist() == [0.0] and t1["pm"].tolist() == [500.0, 600.0]
    # packing [sigtr, siga, nuf, sigf, sigs, dc x 6]; the file order is sigtr, siga, nuf, sigf, SIGS, ADF
    assert t1["xs"][:, 0, 0, 0, :5].tolist() == [[1.0, 11.0, 21.0, 31.0, 41.0], [1.5, 11.5, 21.5, 31.5, 41.5]]
    assert t1["xs"][0, 0, 0, 1, :5].tolist() == [2.0, 12.0, 22.0, 32.0, 42.0]                 # second moderator temperature
    assert np.all(t1["xs"][0, 0, 0, 0, 5:] == 51.0)                                           # one ADF copied to six faces
    assert t1["rxs"][0, 0, 0, 0, 0] == 101.0 and t2["xs"][0, 0, 0, 0, 0] == 1000.0 and t2["rxs"][1, 0, 0, 1, 4] == 1141.5
    assert t2["velo"][0] == 1.0 / 2.0e-6 and t2["ibeta"][5] == 6e-4
    with pytest.raises(ValueError, match="END OF FILE"):
        read_xtab_composition(lines, 3, 1)
    bad = _strip_comments(XTAB_LIB.format(c1=_composition(1.0), c2="").replace("0.6 0.8", "0.8 0.6"), mark="*")
    with pytest.raises(ValueError, match="SMALL to BIG"):
        read_xtab_composition(bad, 1, 1)
