"""Input-deck reader and node numbering for the test / bench harness.

In a drop-in deployment the unchanged Fortran parser (``src/mod_io.f90``) fills the ``sdata``
arrays and the patched ``mod_cmfd`` / ``mod_nodal`` bodies hand them to the C ABI
(``include/adpres_b200.h``).  No Fortran compiler exists in this image, so the harness needs
its own way to turn the reference's decks into exactly the arrays the Fortran side would
pass.  This module restates only that data preparation (nothing of the hot path):

* comment stripping / ``%CARD`` splitting        -- ``mod_io.f90:452-576``
* ``%XSEC``                                       -- ``mod_io.f90:683-762``
* ``%GEOM`` (sizes, divisions, planars, stagger)  -- ``mod_io.f90:809-1143``
* node numbering, ``vdel``, default ``nupd``      -- ``mod_io.f90:1293-1365``
* ``%ITER`` / ``%KERN`` / ``%THET`` / ``%ESRC`` / ``%ADF`` -- ``mod_io.f90:1522-1689,1369-1517,1732-2097``
* ``base_updt`` + ``Dsigr_updt``                  -- ``mod_xsec.f90:172-226``
* ``%XTAB`` branch tables (``inp_xtab``, ``readXS``)  -- ``mod_io.f90:3648-4087``; ``XStab_updt``,
  ``brInterp``, ``crod_tab_updt``                 -- ``mod_xsec.f90:50-86,300-390,520-788``

All arrays are produced in Fortran (column-major) memory order with the reference's
1-based node / mesh indices, i.e. bit-for-bit what ``sdata`` would hold.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional

import numpy as np

KERN_FDM, KERN_PNM, KERN_SANM = 0, 1, 2
TH_PI = float(np.float32(3.14159265))     # REAL(DP), PARAMETER :: pi = 3.14159265 (single-precision literal)
TH_NM = 10                                 # fuel meat meshes (mod_data.f90:159); nt = nm + 2
# steam table at 15.5 MPa (mod_io.f90:3124-3142): T [K], rho [g/cm3], h [J/kg], Pr, kin. viscosity [1e-6 m2/s], k [W/mK]
TH_STAB = np.array([
    [543.15, 0.78106745, 1182595.0, 0.820773, 0.128988, 0.60720],
    [553.15, 0.76428125, 1232620.0, 0.829727, 0.126380, 0.59285],
    [563.15, 0.74619716, 1284170.0, 0.846662, 0.124118, 0.57710],
    [573.15, 0.72650785, 1337630.0, 0.863597, 0.121856, 0.55970],
    [583.15, 0.70475081, 1393570.0, 0.915035, 0.120105, 0.54045],
    [593.15, 0.68018488, 1452895.0, 0.966472, 0.118354, 0.51880],
    [603.15, 0.65150307, 1517175.0, 1.166745, 0.143630, 0.49420],
    [613.15, 0.61590149, 1589770.0, 1.515852, 0.195931, 0.46550],
    [617.91, 0.59896404, 1624307.1, 1.681940, 0.220813, 0.45185]])
_KERN_CODE = {"FDM": KERN_FDM, "PNM": KERN_PNM, "SANM": KERN_SANM}


# --------------------------------------------------------------------------- low-level text
def _strip_comments(text: str, mark: str = "!") -> List[str]:
    """``inp_comments`` (mod_io.f90:452-492): drop blank lines and everything after the comment
    mark ('!' in decks, '*' in %XTAB library files)."""
    out = []
    for raw in text.splitlines():
        line = raw.strip()
        pos = line.find(mark)
        if pos < 0:
            if line:
                out.append(line)
        elif pos > 0:
            line = line[:pos].rstrip()
            if line:
                out.append(line)
    return out


def _split_cards(lines: List[str], base_dir: str) -> Dict[str, List[str]]:
    """``inp_rewrite`` (mod_io.f90:496-576): lines after ``%NAME`` belong to card NAME."""
    cards: Dict[str, List[str]] = {}
    cur: Optional[str] = None
    for line in lines:
        if "%" in line:
            cur = line[line.index("%") + 1:].strip().upper()
            cards.setdefault(cur, [])
            continue
        if cur is None:
            continue
        if "FILE" in line.upper().split()[:1]:
            # ``FILE <path>`` indirection (mod_io.f90:519-535)
            fname = line.split(None, 1)[1].strip()
            if not os.path.isabs(fname) or not os.path.exists(fname):
                fname = os.path.join(base_dir, os.path.basename(fname))
            with open(fname) as fh:
                cards[cur].extend(_strip_comments(fh.read()))
            continue
        cards[cur].append(line)
    return cards


def _tokens(line: str) -> List[str]:
    """Fortran list-directed record: blanks/commas separate, ``n*v`` repeats."""
    out: List[str] = []
    for tok in line.replace(",", " ").split():
        if "*" in tok:
            rep, val = tok.split("*", 1)
            out.extend([val] * int(rep))
        else:
            out.append(tok)
    return out


def _f(tok: str) -> float:
    return float(tok.lower().replace("d", "e"))


class _Reader:
    """One READ per record, like the reference ('x' sentinel prevents record run-on)."""

    def __init__(self, lines: List[str], card: str):
        self.lines, self.pos, self.card = lines, 0, card

    def more(self) -> bool:
        return self.pos < len(self.lines)

    def rec(self, n: Optional[int] = None) -> List[str]:
        if self.pos >= len(self.lines):
            raise ValueError(f"%{self.card}: unexpected end of card")
        toks = _tokens(self.lines[self.pos])
        self.pos += 1
        if n is not None:
            if len(toks) < n:
                raise ValueError(f"%{self.card}: expected {n} values, got {toks}")
            toks = toks[:n]
        return toks

    def ints(self, n: int) -> List[int]:
        return [int(t) for t in self.rec(n)]

    def floats(self, n: int) -> List[float]:
        return [_f(t) for t in self.rec(n)]


# --------------------------------------------------------------------------- the problem
@dataclasses.dataclass
class Problem:
    """Everything ``sdata`` holds for the hot path, Fortran layout, 1-based indices."""
    mode: str
    ng: int
    nmat: int
    # assemblies
    nx: int
    ny: int
    nz: int
    xsize: np.ndarray
    ysize: np.ndarray
    zsize: np.ndarray
    xdiv: np.ndarray
    ydiv: np.ndarray
    zdiv: np.ndarray
    zpln: np.ndarray            # (nz) planar id, 1-based
    planars: np.ndarray         # (npl, nx, ny) material per assembly, index [pl, i, j] 0-based
    bc: np.ndarray              # xeast, xwest, ynorth, ysouth, zbott, ztop
    # material-wise XS (nmat, ng[, ng]) Fortran order
    xsigtr: np.ndarray
    xsiga: np.ndarray
    xnuf: np.ndarray
    xsigf: np.ndarray
    xsigs: np.ndarray
    chi: np.ndarray
    # iteration control
    nout: int = 500
    nin: int = 2
    serc: float = 1.0e-5
    ferc: float = 1.0e-5
    nac: int = 5
    nupd: int = 0
    th_niter: int = 30
    nth: int = 20
    kern: int = KERN_SANM
    biter: int = 0
    sth: float = 1.0
    bth: float = 0.0
    mdc: Optional[np.ndarray] = None     # (nmat, ng, 6) ADF per material
    adf_rot: Optional[list] = None       # [(rot, x1,x2,y1,y2,z1,z2), ...]
    esrc: Optional[list] = None          # [(sden, spec[ng], [(zpos, [(xpos, ypos), ...]), ...]), ...]
    crod: Optional[dict] = None          # %CROD: nb, nstep, pos0, ssize, bpos[nb], bmap(nx,ny), dsigtr/dsiga/dnuf/dsigf (nmat,ng), dsigs (nmat,ng,ng)
    ejct: Optional[dict] = None          # %EJCT: fbpos, tmove, bspeed, ttot, tstep1, tdiv, tstep2, ibeta, lamb, velo
    bextr: int = 0
    fbk: Optional[dict] = None            # %BCON / %CBCS / %FTEM / %MTEM / %CDEN: {name: dict(val, ref, sigtr, siga, nuf, sigf, sigs)}
    ther: Optional[dict] = None           # %THER raw inputs (ppow, pow, tin, cmflow, rf, tg, tc, ppitch, nfpin, ngt, cf)
    xtab: Optional[list] = None           # %XTAB: per material the branch tables of read_xtab_composition()
    cards: Optional[Dict[str, List[str]]] = None
    # ---- node-wise (filled by build())
    nxx: int = 0
    nyy: int = 0
    nzz: int = 0
    nnod: int = 0
    npl: int = 0                # nodes per z-plane (np in set_ind)
    xdel: np.ndarray = None
    ydel: np.ndarray = None
    zdel: np.ndarray = None
    ystag_smin: np.ndarray = None
    ystag_smax: np.ndarray = None
    xstag_smin: np.ndarray = None
    xstag_smax: np.ndarray = None
    ix: np.ndarray = None
    iy: np.ndarray = None
    iz: np.ndarray = None
    mat: np.ndarray = None
    vdel: np.ndarray = None
    sigtr: np.ndarray = None
    siga: np.ndarray = None
    nuf: np.ndarray = None
    sigf: np.ndarray = None
    sigs: np.ndarray = None     # (nnod, ng, ng) F-order: sigs[n, g, h] = g -> h
    D: np.ndarray = None
    sigr: np.ndarray = None
    dc: np.ndarray = None       # (nnod, ng, 6) F-order
    exsrc: np.ndarray = None    # (nnod, ng) F-order

    # ------------------------------------------------------------------ fixtures
    _SPEC_FIELDS = ("mode", "ng", "nmat", "nx", "ny", "nz", "xsize", "ysize", "zsize", "xdiv", "ydiv",
                    "zdiv", "zpln", "planars", "bc", "xsigtr", "xsiga", "xnuf", "xsigf", "xsigs", "chi",
                    "nout", "nin", "serc", "ferc", "nac", "nupd", "th_niter", "nth", "kern", "biter",
                    "sth", "bth", "mdc", "adf_rot", "esrc", "crod", "ejct", "bextr", "ther", "fbk", "xtab")

    def to_spec(self) -> dict:
        """JSON-able problem specification (what the deck says, before node expansion).
        Used for the fixtures under tests/golden/ that travel to the GPU box, where
        /root/reference (and so the decks themselves) does not exist."""
        d = {}
        for k in self._SPEC_FIELDS:
            v = getattr(self, k)
            if k == "nupd" and not self.biter:
                v = 0
            if k == "xtab" and v is not None:
                v = [{kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in t.items()} for t in v]
            elif k == "fbk" and v is not None:
                v = {name: {kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in t.items()} for name, t in v.items()}
            elif isinstance(v, dict):
                v = {kk: (vv.tolist() if isinstance(vv, np.ndarray) else vv) for kk, vv in v.items()}
            d[k] = v.tolist() if isinstance(v, np.ndarray) else v
        return d

    @staticmethod
    def from_spec(d: dict) -> "Problem":
        kw = dict(d)
        for k in ("xsize", "ysize", "zsize"):
            kw[k] = np.array(kw[k], dtype=np.float64)
        for k in ("xdiv", "ydiv", "zdiv", "zpln", "planars", "bc"):
            kw[k] = np.array(kw[k], dtype=np.int32)
        for k in ("xsigtr", "xsiga", "xnuf", "xsigf", "xsigs", "chi"):
            kw[k] = np.asfortranarray(np.array(kw[k], dtype=np.float64))
        if kw.get("mdc") is not None:
            kw["mdc"] = np.array(kw["mdc"], dtype=np.float64)
        if kw.get("adf_rot") is not None:
            kw["adf_rot"] = [tuple(r) for r in kw["adf_rot"]]
        for card in ("crod", "ejct"):
            if kw.get(card) is not None:
                kw[card] = {kk: (np.array(vv) if isinstance(vv, list) else vv) for kk, vv in kw[card].items()}
                if card == "crod":
                    kw[card]["bmap"] = kw[card]["bmap"].astype(np.int32)
        if kw.get("fbk") is not None:
            kw["fbk"] = {name: {kk: (np.array(vv, dtype=np.float64) if isinstance(vv, list) else vv) for kk, vv in t.items()}
                         for name, t in kw["fbk"].items()}
        if kw.get("xtab") is not None:
            kw["xtab"] = [{kk: (np.array(vv, dtype=np.float64) if isinstance(vv, list) else vv) for kk, vv in t.items()}
                          for t in kw["xtab"]]
        return Problem(**kw).build()

    # ------------------------------------------------------------------ geometry
    def refine(self, xdiv=None, ydiv=None, zdiv=None) -> "Problem":
        """Return a copy with new assembly divisions (the synthetic refined meshes of
        BASELINE.json configs 2-5 change only lines 3/5/7 of %GEOM)."""
        p = dataclasses.replace(self)
        if xdiv is not None:
            p.xdiv = np.broadcast_to(np.asarray(xdiv, dtype=np.int32), (self.nx,)).copy()
        if ydiv is not None:
            p.ydiv = np.broadcast_to(np.asarray(ydiv, dtype=np.int32), (self.ny,)).copy()
        if zdiv is not None:
            p.zdiv = np.broadcast_to(np.asarray(zdiv, dtype=np.int32), (self.nz,)).copy()
        if not p.biter:
            p.nupd = 0
        return p.build()

    def build(self) -> "Problem":
        """inp_geom1/2 + misc + XS_updt(base) -> node-wise arrays."""
        nx, ny, nz = self.nx, self.ny, self.nz
        self.nxx, self.nyy, self.nzz = int(self.xdiv.sum()), int(self.ydiv.sum()), int(self.zdiv.sum())
        # node sizes: div = size / REAL(div)  (mod_io.f90:925-953)
        self.xdel = np.repeat(self.xsize / self.xdiv.astype(np.float64), self.xdiv)
        self.ydel = np.repeat(self.ysize / self.ydiv.astype(np.float64), self.ydiv)
        self.zdel = np.repeat(self.zsize / self.zdiv.astype(np.float64), self.zdiv)
        ia = np.repeat(np.arange(nx), self.xdiv)      # node -> assembly index
        ja = np.repeat(np.arange(ny), self.ydiv)
        ka = np.repeat(np.arange(nz), self.zdiv)
        # mnum(i,j,1): staggering is taken from plane 1 only (mod_io.f90:1071-1110)
        pl1 = self.planars[self.zpln[0] - 1]
        m1 = pl1[np.ix_(ia, ja)]                       # (nxx, nyy)
        nzmask = m1 != 0
        if not nzmask.any(axis=0).all() or not nzmask.any(axis=1).all():
            raise ValueError("empty row/column in planar 1 (unsupported by the reference too)")
        i1 = np.arange(1, self.nxx + 1)
        j1 = np.arange(1, self.nyy + 1)
        self.ystag_smin = np.where(nzmask, i1[:, None], self.nxx + 1).min(axis=0).astype(np.int32)
        self.ystag_smax = np.where(nzmask, i1[:, None], 0).max(axis=0).astype(np.int32)
        self.xstag_smin = np.where(nzmask, j1[None, :], self.nyy + 1).min(axis=1).astype(np.int32)
        self.xstag_smax = np.where(nzmask, j1[None, :], 0).max(axis=1).astype(np.int32)
        # numbering k, j, i (mod_io.f90:1319-1330)
        ii, jj = [], []
        for j in range(self.nyy):
            lo, hi = self.ystag_smin[j], self.ystag_smax[j]
            ii.append(np.arange(lo, hi + 1, dtype=np.int32))
            jj.append(np.full(hi - lo + 1, j + 1, dtype=np.int32))
        ip, jp = np.concatenate(ii), np.concatenate(jj)
        self.npl = int(ip.size)
        self.nnod = self.npl * self.nzz
        self.ix = np.tile(ip, self.nzz)
        self.iy = np.tile(jp, self.nzz)
        self.iz = np.repeat(np.arange(1, self.nzz + 1, dtype=np.int32), self.npl)
        # material per node
        matp = np.empty((self.nzz, self.npl), dtype=np.int32)
        for k in range(self.nzz):
            pl = self.planars[self.zpln[ka[k]] - 1]
            matp[k] = pl[ia[ip - 1], ja[jp - 1]]
        self.mat = matp.reshape(-1)
        if (self.mat == 0).any():
            raise ValueError("Zero material found inside core. Check material assignment")
        if (self.mat > self.nmat).any():
            raise ValueError("material id greater than number of materials")
        self.vdel = self.xdel[self.ix - 1] * self.ydel[self.iy - 1] * self.zdel[self.iz - 1]
        if self.nupd == 0:
            # nupd = ceiling((nxx+nyy+nzz)/2.5), default REAL arithmetic (mod_io.f90:1363)
            self.nupd = int(math.ceil(float(np.float32(self.nxx + self.nyy + self.nzz) / np.float32(2.5))))
        if self.xtab is not None:
            self.chi = np.asfortranarray(np.array([t["chi"] for t in self.xtab], dtype=np.float64))
        self.update_xs()
        if self.xtab is None:
            self._build_adf()          # %XTAB decks: the ADFs come out of the tables (XStab_updt)
        self._build_esrc()
        return self

    def update_xs(self, bpos=None, bcon=None, ftem=None, mtem=None, cden=None) -> None:
        """XS_updt (mod_xsec.f90:11-46): base_updt, then bcon_updt / ftem_updt / mtem_updt / cden_updt
        (:396-516) for the feedback cards of the deck, crod_updt(bpos), Dsigr_updt.  A parameter
        left at None takes the value of its card.  %XTAB decks go through XStab_updt instead."""
        if self.xtab is not None:
            self._xstab_updt(bpos, bcon, ftem, mtem, cden)
            return
        m = self.mat - 1
        N, G = self.nnod, self.ng
        self.sigtr = np.asfortranarray(self.xsigtr[m, :])
        self.siga = np.asfortranarray(self.xsiga[m, :])
        self.nuf = np.asfortranarray(self.xnuf[m, :])
        self.sigf = np.asfortranarray(self.xsigf[m, :])
        self.sigs = np.asfortranarray(self.xsigs[m, :, :])
        given = dict(bcon=bcon, ftem=ftem, mtem=mtem, cden=cden)
        for key in ("bcon", "ftem", "mtem", "cden"):                     # the reference's order
            t = (self.fbk or {}).get(key)
            if t is None:
                continue
            x = t["val"] if given[key] is None else given[key]
            if key == "ftem":
                delta = np.sqrt(x) - np.sqrt(t["ref"])
            else:
                delta = x - t["ref"]
            delta = np.broadcast_to(np.asarray(delta, dtype=np.float64), (N,))
            self.sigtr = self.sigtr + t["sigtr"][m, :] * delta[:, None]
            self.siga = self.siga + t["siga"][m, :] * delta[:, None]
            self.nuf = self.nuf + t["nuf"][m, :] * delta[:, None]
            self.sigf = self.sigf + t["sigf"][m, :] * delta[:, None]
            self.sigs = self.sigs + t["sigs"][m, :, :] * delta[:, None, None]
        if self.crod is not None:
            self.crod_updt(self.crod["bpos"] if bpos is None else bpos)
        self.finish_xs()

    @property
    def coreh(self) -> float:
        h = 0.0
        for z in self.zdel:          # mod_io.f90:1274-1277
            h = h + z
        return h

    def crod_updt(self, bpos) -> None:
        """crod_updt (mod_xsec.f90:230-296): volume-weighted rodded cross sections, rods enter from
        the top; the partially rodded node gets the fraction vfrac of the increment."""
        c = self.crod
        ia, ja, _ = self._node_assembly_maps()
        fbmap = c["bmap"][np.ix_(ia, ja)]                 # (nxx, nyy) node-wise bank map
        coreh = self.coreh
        # node number of (i, j, k): plane-invariant position + k * npl
        pos = np.full((self.nxx + 1, self.nyy + 1), -1, dtype=np.int64)
        pos[self.ix[:self.npl], self.iy[:self.npl]] = np.arange(self.npl)
        for j in range(1, self.nyy + 1):
            for i in range(1, self.nxx + 1):
                b = fbmap[i - 1, j - 1]
                if b <= 0 or pos[i, j] < 0:
                    continue
                rodh = coreh - c["pos0"] - bpos[b - 1] * c["ssize"]
                dum = 0.0
                for k in range(self.nzz, 0, -1):
                    n = pos[i, j] + (k - 1) * self.npl
                    mm = self.mat[n] - 1
                    if rodh >= dum and rodh <= dum + self.zdel[k - 1]:
                        vfrac = (rodh - dum) / self.zdel[k - 1]
                    else:
                        vfrac = None
                    w = 1.0 if vfrac is None else vfrac
                    if vfrac is None:
                        self.sigtr[n, :] = self.sigtr[n, :] + c["dsigtr"][mm, :]
                        self.siga[n, :] = self.siga[n, :] + c["dsiga"][mm, :]
                        self.nuf[n, :] = self.nuf[n, :] + c["dnuf"][mm, :]
                        self.sigf[n, :] = self.sigf[n, :] + c["dsigf"][mm, :]
                        self.sigs[n, :, :] = self.sigs[n, :, :] + c["dsigs"][mm, :, :]
                        dum = dum + self.zdel[k - 1]
                    else:
                        self.sigtr[n, :] = self.sigtr[n, :] + w * c["dsigtr"][mm, :]
                        self.siga[n, :] = self.siga[n, :] + w * c["dsiga"][mm, :]
                        self.nuf[n, :] = self.nuf[n, :] + w * c["dnuf"][mm, :]
                        self.sigf[n, :] = self.sigf[n, :] + w * c["dsigf"][mm, :]
                        self.sigs[n, :, :] = self.sigs[n, :, :] + w * c["dsigs"][mm, :, :]
                        break
                col = pos[i, j] + np.arange(self.nzz) * self.npl
                for a in (self.siga, self.nuf, self.sigf, self.sigs):
                    blk = a[col]
                    blk[blk < 0.0] = 0.0
                    a[col] = blk

    # ------------------------------------------------------------------ %XTAB branch tables
    def _br_interp(self, t: dict, rod: int, cden, bcon, ftem, mtem) -> np.ndarray:
        """brInterp (mod_xsec.f90:520-788) for all nodes of one material at once: the two closest
        branch points per parameter (up to 20 % -- boron: 100 ppm -- outside the table the end
        interval extrapolates, further out the reference STOPs), then linear interpolation in the
        order moderator temperature, fuel temperature, boron, coolant density with the
        reference's operation order `a + radx * (b - a)`.  Returns (n, nval) packed like t["xs"]."""
        tab = t["rxs"] if rod else t["xs"]
        n = len(cden)

        def bracket(x, par, dim, absolute, what):
            i1 = np.zeros(n, dtype=np.int64)
            i2 = np.zeros(n, dtype=np.int64)
            if dim <= 1:
                return i1, i2
            x = np.broadcast_to(np.asarray(x, dtype=np.float64), (n,))
            inside = (x >= par[0]) & (x <= par[dim - 1])
            found = np.zeros(n, dtype=bool)
            for s in range(1, dim):                       # DO s = 2, mx ... EXIT at the first hit
                c = inside & ~found & (x >= par[s - 1]) & (x <= par[s])
                i1[c], i2[c] = s - 1, s
                found |= c
            if absolute:
                lo = ~inside & (x < par[0]) & ((par[0] - x) < 100.0)
                hi = ~inside & ~lo & (x > par[dim - 1]) & ((x - par[dim - 1]) < 100.0)
            else:
                with np.errstate(divide="ignore", invalid="ignore"):
                    lo = ~inside & (x < par[0]) & ((par[0] - x) / par[0] < float(np.float32(0.2)))
                    hi = ~inside & ~lo & (x > par[dim - 1]) & ((x - par[dim - 1]) / par[dim - 1] < float(np.float32(0.2)))
            i1[lo], i2[lo] = 0, 1
            i1[hi], i2[hi] = dim - 2, dim - 1
            bad = ~(inside | lo | hi)
            if bad.any():
                raise ValueError(f"ERROR: {what} {x[bad][0]:.3f} IS OUT OF THE RANGE OF THE BRANCH PARAMETER")
            return i1, i2

        s1, s2 = bracket(cden, t["pd"], t["nd"], False, "COOLANT DENSITY")
        t1, t2 = bracket(bcon, t["pb"], t["nb"], True, "BORON CONCENTRATION")
        u1, u2 = bracket(ftem, t["pf"], t["nf"], False, "FUEL TEMPERATURE")
        v1, v2 = bracket(mtem, t["pm"], t["nm"], False, "MODERATOR TEMPERATURE")
        xs = [None] * 9
        corners = ((s1, t1, u1), (s1, t1, u2), (s1, t2, u1), (s1, t2, u2), (s2, t1, u1), (s2, t1, u2), (s2, t2, u1), (s2, t2, u2))
        if t["nm"] > 1:
            pm = t["pm"]
            radx = ((np.asarray(mtem, dtype=np.float64) - pm[v1]) / (pm[v2] - pm[v1]))[:, None]
            for i, (a, b, c) in enumerate(corners, start=1):
                xs[i] = tab[a, b, c, v1] + radx * (tab[a, b, c, v2] - tab[a, b, c, v1])
        else:
            for i, (a, b, c) in enumerate(corners, start=1):
                xs[i] = tab[a, b, c, v1]
        if t["nf"] > 1:
            pf = t["pf"]
            radx = ((np.asarray(ftem, dtype=np.float64) - pf[u1]) / (pf[u2] - pf[u1]))[:, None]
            for i in (1, 3, 5, 7):
                xs[i] = xs[i] + radx * (xs[i + 1] - xs[i])
        if t["nb"] > 1:
            pb = t["pb"]
            radx = ((np.broadcast_to(np.asarray(bcon, dtype=np.float64), (n,)) - pb[t1]) / (pb[t2] - pb[t1]))[:, None]
            xs[1] = xs[1] + radx * (xs[3] - xs[1])
            xs[5] = xs[5] + radx * (xs[7] - xs[5])
        if t["nd"] > 1:
            pd = t["pd"]
            radx = ((np.asarray(cden, dtype=np.float64) - pd[s1]) / (pd[s2] - pd[s1]))[:, None]
            xs[1] = xs[1] + radx * (xs[5] - xs[1])
        return xs[1]

    def rod_fractions(self, bpos) -> np.ndarray:
        """The sweep of crod_updt / crod_tab_updt (mod_xsec.f90:253-279,322-372) reduced to its outcome per
        node: -1 = not visited (unrodded), 1 = fully rodded, 0 <= w <= 1 = the partially rodded node where
        the sweep EXITs.  (A rod tip above the core top, rodh < 0, never meets the `partial` test, so the
        whole column counts as fully rodded -- the reference's behaviour.)"""
        c = self.crod
        w = np.full(self.nnod, -1.0)
        ia, ja, _ = self._node_assembly_maps()
        fbmap = c["bmap"][np.ix_(ia, ja)]
        coreh = self.coreh
        pos = np.full((self.nxx + 1, self.nyy + 1), -1, dtype=np.int64)
        pos[self.ix[:self.npl], self.iy[:self.npl]] = np.arange(self.npl)
        for j in range(1, self.nyy + 1):
            for i in range(1, self.nxx + 1):
                b = fbmap[i - 1, j - 1]
                if b <= 0 or pos[i, j] < 0:
                    continue
                rodh = coreh - c["pos0"] - bpos[b - 1] * c["ssize"]
                dum = 0.0
                for k in range(self.nzz, 0, -1):
                    n = pos[i, j] + (k - 1) * self.npl
                    if rodh >= dum and rodh <= dum + self.zdel[k - 1]:
                        w[n] = (rodh - dum) / self.zdel[k - 1]
                        break
                    w[n] = 1.0
                    dum = dum + self.zdel[k - 1]
        return w

    def rodded_columns(self) -> np.ndarray:
        """(nnod) bool: node lies in a column under a control rod bank (fbmap > 0)."""
        ia, ja, _ = self._node_assembly_maps()
        fbmap = self.crod["bmap"][np.ix_(ia, ja)]
        return fbmap[self.ix - 1, self.iy - 1] > 0

    def xtab_defaults(self):
        """What the XS update sees before the first TH solve: inp_ther sets ftem = 900., cden = 0.711,
        mtem = 500. (default-REAL literals, mod_io.f90:3103-3105); bcon is the %BCON value if the card
        is read (RODEJECT only, :294), else rbcon, which XTAB decks never set (0)."""
        b = (self.fbk or {}).get("bcon")
        return dict(bcon=0.0 if b is None else b["val"], ftem=900.0, mtem=500.0, cden=float(np.float32(0.711)))

    def _xstab_updt(self, bpos, bcon, ftem, mtem, cden) -> None:
        """XStab_updt (mod_xsec.f90:50-86): brInterp(unrodded) for every node, crod_tab_updt (:300-390)
        -- rodded nodes take the rodded table, the partially rodded node the volume-weighted mix, negative
        values in rodded columns are suppressed --, Dsigr_updt.  The ADFs come out of the tables too."""
        N, G = self.nnod, self.ng
        dflt = self.xtab_defaults()
        bcon = dflt["bcon"] if bcon is None else float(bcon)
        ftem = np.broadcast_to(np.asarray(dflt["ftem"] if ftem is None else ftem, dtype=np.float64), (N,))
        mtem = np.broadcast_to(np.asarray(dflt["mtem"] if mtem is None else mtem, dtype=np.float64), (N,))
        cden = np.broadcast_to(np.asarray(dflt["cden"] if cden is None else cden, dtype=np.float64), (N,))
        nval = 4 * G + G * G + 6 * G
        val = np.empty((N, nval))
        for mn, t in enumerate(self.xtab):
            sel = np.nonzero(self.mat == mn + 1)[0]
            if sel.size:
                val[sel] = self._br_interp(t, 0, cden[sel], bcon, ftem[sel], mtem[sel])
        if self.crod is not None:
            w = self.rod_fractions(self.crod["bpos"] if bpos is None else bpos)
            hit = np.nonzero(w >= 0.0)[0]
            for mn in np.unique(self.mat[hit]):
                t = self.xtab[mn - 1]
                if t["trod"] != 1:
                    raise ValueError(f"CONTROL ROD BANK COINCIDES WITH MATERIAL NUMBER {mn} THAT DOES NOT HAVE CONTROL ROD DATA IN XTAB FILE")
                sel = hit[self.mat[hit] == mn]
                rod = self._br_interp(t, 1, cden[sel], bcon, ftem[sel], mtem[sel])
                # fully rodded nodes take the rodded set; the partially rodded node the volume-weighted mix
                # (vfrac == 1 there gives 0 * unrodded + rodded, the same value)
                vf = w[sel][:, None]
                val[sel] = np.where(vf == 1.0, rod, (1.0 - vf) * val[sel] + vf * rod)
            cols = self.rodded_columns()
            blk = val[cols, G:]                     # siga, nuf, sigf, sigs, dc (not sigtr)
            blk[blk < 0.0] = 0.0
            val[cols, G:] = blk
        self.sigtr = np.asfortranarray(val[:, 0:G])
        self.siga = np.asfortranarray(val[:, G:2 * G])
        self.nuf = np.asfortranarray(val[:, 2 * G:3 * G])
        self.sigf = np.asfortranarray(val[:, 3 * G:4 * G])
        self.sigs = np.asfortranarray(val[:, 4 * G:4 * G + G * G].reshape(N, G, G))
        self.dc = np.asfortranarray(val[:, 4 * G + G * G:].reshape(N, G, 6))
        self.finish_xs()

    def finish_xs(self) -> None:
        """Dsigr_updt: D = 1/(3 sigtr); sigr = siga + sum_{h != g} sigs(g -> h), h ascending; then check_xs
        (mod_xsec.f90:90-168) -- the reference STOPs on a vanishing diffusion coefficient or a negative
        removal / nu-fission / scattering cross section or fission spectrum."""
        N, G = self.nnod, self.ng
        if (self.sigtr < float(np.float32(1.0e-5))).any():
            raise ValueError("Negative diffusion coefficient encountered")
        self.D = np.asfortranarray(1.0 / (3.0 * self.sigtr))
        sigr = np.zeros((N, G), order="F")
        for g in range(G):
            dum = np.zeros(N)
            for h in range(G):
                if h != g:
                    dum = dum + self.sigs[:, g, h]
            sigr[:, g] = self.siga[:, g] + dum
        self.sigr = sigr
        for arr, what in ((self.D < float(np.float32(1.0e-20)), "DIFFUSION COEF. IS CLOSE TO ZERO OR NEGATIVE"),
                          (self.sigr < 0.0, "REMOVAL XS IS NEGATIVE"), (self.nuf < 0.0, "NU*FISSION XS IS NEGATIVE"),
                          (self.chi < 0.0, "FISSION SPECTRUM IS NEGATIVE"), (self.sigs < 0.0, "SCATTERING XS IS NEGATIVE")):
            if arr.any():
                raise ValueError("ERROR IN THE CROSS SECTIONS: " + what)

    # ------------------------------------------------------------------ ADF / ESRC
    def _node_assembly_maps(self):
        ia = np.repeat(np.arange(self.nx), self.xdiv)
        ja = np.repeat(np.arange(self.ny), self.ydiv)
        ka = np.repeat(np.arange(self.nz), self.zdiv)
        return ia, ja, ka

    def _build_adf(self) -> None:
        N, G = self.nnod, self.ng
        self.dc = np.ones((N, G, 6), order="F")
        if self.mdc is None:
            return
        nx, ny, nz = self.nx, self.ny, self.nz
        ia, ja, ka = self._node_assembly_maps()
        # xdc(i,j,k,g) = mdc(asm material)   (mod_io.f90:1784-1792)
        asm_mat = np.stack([self.planars[self.zpln[k] - 1] for k in range(nz)], axis=2)  # (nx,ny,nz)
        xdc = np.zeros((nx, ny, nz, G, 6))
        nzm = asm_mat != 0
        xdc[nzm] = self.mdc[asm_mat[nzm] - 1]
        for (rot, x1, x2, y1, y2, z1, z2) in (self.adf_rot or []):
            blk = xdc[x1 - 1:x2, y1 - 1:y2, z1 - 1:z2]
            a = blk[..., :4].copy()
            if rot == 1:
                blk[..., 0], blk[..., 1], blk[..., 2], blk[..., 3] = a[..., 3], a[..., 2], a[..., 0], a[..., 1]
            elif rot == 2:
                blk[..., 0], blk[..., 1], blk[..., 2], blk[..., 3] = a[..., 1], a[..., 0], a[..., 3], a[..., 2]
            elif rot == 3:
                blk[..., 0], blk[..., 1], blk[..., 2], blk[..., 3] = a[..., 2], a[..., 3], a[..., 1], a[..., 0]
        tx = np.concatenate([[1], 1 + np.cumsum(self.xdiv)])
        ty = np.concatenate([[1], 1 + np.cumsum(self.ydiv)])
        tz = np.concatenate([[1], 1 + np.cumsum(self.zdiv)])
        i, j, k = self.ix, self.iy, self.iz
        ai, aj, ak = ia[i - 1], ja[j - 1], ka[k - 1]
        src = xdc[ai, aj, ak]                       # (N, G, 6)
        # only faces lying on an assembly boundary take the ADF (mod_io.f90:2071-2078);
        # note faces 5/6: value 5 is applied at the assembly *bottom*, 6 at its top.
        sel = [(i == tx[ai + 1] - 1), (i == tx[ai]), (j == ty[aj + 1] - 1), (j == ty[aj]),
               (k == tz[ak]), (k == tz[ak + 1] - 1)]
        for u in range(6):
            self.dc[sel[u], :, u] = src[sel[u], :, u]

    def _build_esrc(self) -> None:
        N, G = self.nnod, self.ng
        self.exsrc = np.zeros((N, G), order="F")
        if not self.esrc or self.mode != "FIXEDSRC":
            return
        ia, ja, ka = self._node_assembly_maps()
        ai, aj, ak = ia[self.ix - 1], ja[self.iy - 1], ka[self.iz - 1]
        for sden, spec, zlist in self.esrc:
            for zpos, xy in zlist:
                for xpos, ypos in xy:
                    sel = (ai == xpos - 1) & (aj == ypos - 1) & (ak == zpos - 1)
                    for g in range(G):
                        self.exsrc[sel, g] += sden * spec[g]

    # ------------------------------------------------------------------ results
    def th_setup(self) -> dict:
        """Derived thermal-hydraulic data of inp_ther (mod_io.f90:3036-3146): pin geometry, sub-channel
        flow, fuel pins per node, radial pin mesh, steam table at 15.5 MPa.  `pi` is the reference's
        default-REAL literal 3.14159265 (mod_data.f90:169)."""
        t = self.ther
        if t is None:
            raise ValueError("deck has no %THER card")
        if t["tg"] > 0.25 * t["rf"] or t["tc"] > 0.25 * t["rf"]:
            raise ValueError("ERROR: GAP / CLADDING THICKNESS IS TO LARGE (> 0.25*rf)")
        pi = TH_PI
        nm, nt = TH_NM, TH_NM + 2
        rf, tg, tc, ppitch = t["rf"], t["tg"], t["tc"], t["ppitch"]
        rg = rf + tg
        rc = rg + tc
        dia = 2.0 * rc
        dh = dia * ((4.0 / pi) * (ppitch / dia) ** 2 - 1.0)
        farea = ppitch ** 2 - 0.25 * pi * dia ** 2
        cflow = t["cmflow"] / float(np.float32(t["nfpin"]))
        area = self.xsize[:, None] * self.ysize[None, :]
        barea = area.max()
        ia = np.repeat(np.arange(self.nx), self.xdiv)
        ja = np.repeat(np.arange(self.ny), self.ydiv)
        div = (self.xdiv[:, None] * self.ydiv[None, :]).astype(np.float32).astype(np.float64)
        node_nf = np.zeros((self.nxx, self.nyy), order="F")
        for j in range(self.nyy):
            for i in range(self.ystag_smin[j] - 1, self.ystag_smax[j]):
                node_nf[i, j] = area[ia[i], ja[j]] * float(np.float32(t["nfpin"])) / (barea * div[ia[i], ja[j]])
        rdel = np.zeros(nt)
        rdel[:nm] = rf / float(np.float32(nm))
        rdel[nm] = tg
        rdel[nm + 1] = tc
        rpos = np.zeros(nt)
        rpos[0] = 0.5 * rdel[0]
        for i in range(1, nt):
            rpos[i] = rpos[i - 1] + 0.5 * (rdel[i - 1] + rdel[i])
        return dict(pi=pi, nm=nm, nt=nt, rf=rf, rg=rg, rc=rc, dia=dia, dh=dh, farea=farea, cflow=cflow, cf=t["cf"],
                    tin=t["tin"], pow=t["pow"], ppow=t["ppow"], node_nf=node_nf, rdel=rdel, rpos=rpos,
                    stab=TH_STAB.copy(order="F"), ntem=TH_STAB.shape[0])

    def asm_power(self, pow_n: np.ndarray) -> np.ndarray:
        """AsmPow normalisation (mod_io.f90:3267-3363): axial average weighted by zdel,
        area-weighted per assembly, scaled so the mean over assemblies with power>0 is 1.
        Returns (nx, ny)."""
        fx = np.zeros((self.nxx, self.nyy, self.nzz))
        fx[self.ix - 1, self.iy - 1, self.iz - 1] = pow_n
        fnode = (fx * self.zdel[None, None, :]).sum(axis=2) / self.zdel.sum()
        ia, ja, _ = self._node_assembly_maps()
        area = self.xdel[:, None] * self.ydel[None, :]
        fasm = np.zeros((self.nx, self.ny))
        for a in range(self.nx):
            for b in range(self.ny):
                sel = np.ix_(ia == a, ja == b)
                fasm[a, b] = (fnode[sel] * area[sel]).sum() / area[sel].sum()
        pos = fasm > 0
        totp = fasm[pos].sum()
        if totp > 0:
            fasm = float(pos.sum()) / totp * fasm
        return fasm


# --------------------------------------------------------------------------- %XTAB library files
def _find_xtab_file(name: str, base_dir: str) -> str:
    """The sample decks carry the absolute paths of the author's machine
    (/home/imronuke/ADPRES/smpl/xsec/...): if the path does not exist, look for its `smpl/...` tail
    under the tree the deck itself lives in."""
    if os.path.exists(name):
        return os.path.abspath(name)
    parts = name.replace("\\", "/").split("/")
    if "smpl" in parts:
        tail = parts[parts.index("smpl"):]
        d = os.path.abspath(base_dir)
        while True:
            cand = os.path.join(d, *tail)
            if os.path.exists(cand):
                return cand
            if os.path.dirname(d) == d:
                break
            d = os.path.dirname(d)
    cand = os.path.join(base_dir, os.path.basename(name))
    if os.path.exists(cand):
        return cand
    raise FileNotFoundError(f"XTAB File Open Failed: {name}")


def read_xtab_composition(lines: List[str], cnum: int, ng: int, fname: str = "") -> dict:
    """One composition of a %XTAB library (inp_xtab mod_io.f90:3719-3812, readXS :3916-4061).
    `lines` is the comment-stripped file.  Per branch point (s = coolant density, t = boron,
    u = fuel temperature, v = moderator temperature) the values are packed as
    [sigtr(G), siga(G), nuf(G), sigf(G), sigs(G,G) row g -> column h, dc(G,6)]."""
    r = _Reader(lines, "XTAB " + fname)
    tadf, trod = r.ints(2)
    nd, nb, nf, nm = r.ints(4)
    if min(nd, nb, nf, nm) < 1:
        raise ValueError("ERROR: MINIMUM NUMBER OF BRANCH IS 1")
    pars = []
    for dim, what in ((nd, "COOLANT DENSITY"), (nb, "BORON CONCENTRATION"), (nf, "FUEL TEMPERATURE"), (nm, "MODERATOR TEMPERATURE")):
        if dim > 1:                                         # branchPar (:3879-3912)
            par = np.array(r.floats(dim))
            if (par[:-1] > par[1:]).any():
                raise ValueError(f"{what} PARAMETER SHALL BE IN ORDER, SMALL to BIG")
        else:
            par = np.zeros(1)
        pars.append(par)
    if tadf not in (1, 2):
        raise ValueError("XTAB libraries without ADFs leave dc undefined in the reference; unsupported")
    nskip = ng * nb * nf * nm
    per_set = (5 if tadf == 1 else 10) * nskip + ng * nskip          # records of one (un)rodded set
    r.pos += (cnum - 1) * ((2 if trod == 1 else 1) * per_set + 4)    # skipRead (:3780-3798)
    if r.pos >= len(lines):
        raise ValueError(f"END OF FILE REACHED FOR XTAB FILE {fname}")
    nval = 4 * ng + ng * ng + 6 * ng

    def read_set():
        a = np.zeros((nd, nb, nf, nm, nval))

        def block(col):
            for v in range(nm):
                for u in range(nf):
                    for t in range(nb):
                        a[:, t, u, v, col] = r.floats(nd)
        for kind in range(4):                               # sigtr, siga, nuf, sigf
            for g in range(ng):
                block(kind * ng + g)
        for g in range(ng):
            for h in range(ng):
                block(4 * ng + g * ng + h)
        o = 4 * ng + ng * ng
        if tadf == 1:
            for g in range(ng):
                block(o + g * 6)
                for k in range(1, 6):
                    a[..., o + g * 6 + k] = a[..., o + g * 6]
        else:
            for g in range(ng):
                for k in range(6):
                    block(o + g * 6 + k)
        return a
    xs = read_set()
    rxs = read_set() if trod == 1 else None
    chi = np.array(r.floats(ng))
    velo = 1.0 / np.array(r.floats(ng))                     # the library holds inverse velocities
    lamb = np.array(r.floats(6))
    ibeta = np.array(r.floats(6))
    return dict(tadf=tadf, trod=trod, nd=nd, nb=nb, nf=nf, nm=nm, pd=pars[0], pb=pars[1], pf=pars[2], pm=pars[3],
                xs=xs, rxs=rxs, chi=chi, velo=velo, lamb=lamb, ibeta=ibeta)


# --------------------------------------------------------------------------- parsing
def parse_deck(text: str, base_dir: str = ".") -> Problem:
    cards = _split_cards(_strip_comments(text), base_dir)
    if "MODE" not in cards:
        raise ValueError("CARD %MODE DOES NOT PRESENT")
    mode = cards["MODE"][0].split()[0].upper()
    if "XSEC" not in cards and "XTAB" not in cards:
        raise ValueError("CARD %XSEC OR %XTAB DOES NOT PRESENT")
    if "GEOM" not in cards:
        raise ValueError("CARD %GEOM DOES NOT PRESENT")

    # ---- %XSEC (mod_io.f90:683-762); xsigs(mat, g, h) = scattering g -> h
    xtab = None
    if "XSEC" in cards:
        r = _Reader(cards["XSEC"], "XSEC")
        ng, nmat = r.ints(2)
    else:
        # ---- %XTAB (inp_xtab, mod_io.f90:3648-3875): ng, nmat, then per material a library file and
        # the number of the composition inside it
        r = _Reader(cards["XTAB"], "XTAB")
        ng, nmat = r.ints(2)
        xtab, files = [], {}
        for i in range(nmat):
            v = cards["XTAB"][r.pos].split()
            r.pos += 1
            path = _find_xtab_file(v[0], base_dir)
            if path not in files:
                with open(path) as fh:
                    files[path] = _strip_comments(fh.read(), mark="*")
            xtab.append(read_xtab_composition(files[path], int(v[1]), ng, os.path.basename(path)))
    xsigtr = np.zeros((nmat, ng), order="F")
    xsiga = np.zeros((nmat, ng), order="F")
    xnuf = np.zeros((nmat, ng), order="F")
    xsigf = np.zeros((nmat, ng), order="F")
    chi = np.zeros((nmat, ng), order="F")
    xsigs = np.zeros((nmat, ng, ng), order="F")
    for i in range(nmat if xtab is None else 0):
        for g in range(ng):
            v = r.floats(5 + ng)
            xsigtr[i, g], xsiga[i, g], xnuf[i, g], xsigf[i, g], chi[i, g] = v[:5]
            xsigs[i, g, :] = v[5:]
            if xsigtr[i, g] <= 0.0:
                raise ValueError("Transport cross section (sigtr) is zero or negative")

    # ---- %GEOM part 1 (mod_io.f90:809-953)
    r = _Reader(cards["GEOM"], "GEOM")
    nx, ny, nz = r.ints(3)
    xsize = np.array(r.floats(nx)); xdiv = np.array(r.ints(nx), dtype=np.int32)
    ysize = np.array(r.floats(ny)); ydiv = np.array(r.ints(ny), dtype=np.int32)
    zsize = np.array(r.floats(nz)); zdiv = np.array(r.ints(nz), dtype=np.int32)
    # ---- %GEOM part 2 (mod_io.f90:958-1143): planars are listed north (j = ny) first
    npl = r.ints(1)[0]
    zpln = np.array(r.ints(nz), dtype=np.int32)
    if (zpln > npl).any():               # mod_io.f90:990-1001
        raise ValueError(f"ERROR: PLANAR {int(zpln.max())} IS GREATER THAN NUMBER OF PLANAR")
    if (zpln < 1).any():
        raise ValueError("ERROR: PLANAR SHOULD BE AT LEAST 1")
    planars = np.zeros((npl, nx, ny), dtype=np.int32)
    for k in range(npl):
        for j in range(ny - 1, -1, -1):
            planars[k, :, j] = r.ints(nx)
    bc = np.array(r.ints(6), dtype=np.int32)
    if (bc > 2).any() or (bc < 0).any():
        raise ValueError("wrong boundary condition")

    p = Problem(mode=mode, ng=ng, nmat=nmat, nx=nx, ny=ny, nz=nz, xsize=xsize, ysize=ysize,
                zsize=zsize, xdiv=xdiv, ydiv=ydiv, zdiv=zdiv, zpln=zpln, planars=planars, bc=bc,
                xsigtr=xsigtr, xsiga=xsiga, xnuf=xnuf, xsigf=xsigf, xsigs=xsigs, chi=chi,
                cards=cards, xtab=xtab)

    if "KERN" in cards:                 # mod_io.f90:1593-1621
        name = cards["KERN"][0].split()[0].upper()
        if name not in _KERN_CODE:
            raise ValueError(f"COULD NOT RECOGNIZE NODAL KERNEL: {name}")
        p.kern = _KERN_CODE[name]
    if "ITER" in cards:                 # mod_io.f90:1522-1540
        v = _tokens(cards["ITER"][0])
        p.nout, p.nin = int(v[0]), int(v[1])
        p.serc, p.ferc = _f(v[2]), _f(v[3])
        p.nac, p.nupd, p.th_niter, p.nth = int(v[4]), int(v[5]), int(v[6]), int(v[7])
        p.biter = 1
    if "THET" in cards:                 # mod_io.f90:1650-1689
        p.sth = _f(_tokens(cards["THET"][0])[0])
        if p.sth < 0.001 or p.sth > 1.0:
            raise ValueError("THETA VALUE out of range")
        p.bth = (1.0 - p.sth) / p.sth
    if "ADF" in cards:                  # mod_io.f90:1774-1831
        r = _Reader(cards["ADF"], "ADF")
        mdc = np.zeros((nmat, ng, 6))
        for i in range(nmat):
            for g in range(ng):
                mdc[i, g, :] = r.floats(6)
        rots = []
        while r.more():
            rot = r.ints(1)[0]
            if rot < 1:
                break
            while True:
                v = r.ints(6)
                if min(v) < 1:
                    break
                rots.append((rot, *v))
        p.mdc, p.adf_rot = mdc, rots
    if "CROD" in cards:                 # mod_io.f90:2148-2330 (%XSEC decks: material-wise increments)
        r = _Reader(cards["CROD"], "CROD")
        v = r.rec(2)
        nb, nstep = int(v[0]), _f(v[1])
        pos0, ssize = r.floats(2)
        bpos = np.array(r.floats(nb))
        bmap = np.zeros((nx, ny), dtype=np.int32)
        for j in range(ny - 1, -1, -1):
            bmap[:, j] = r.ints(nx)
        dsigtr = np.zeros((nmat, ng)); dsiga = np.zeros((nmat, ng)); dnuf = np.zeros((nmat, ng))
        dsigf = np.zeros((nmat, ng)); dsigs = np.zeros((nmat, ng, ng))
        for i in range(nmat if xtab is None else 0):    # %XTAB decks: the rodded sets are in the library (:2225-2232)
            for g in range(ng):
                v = r.floats(4 + ng)
                dsigtr[i, g], dsiga[i, g], dnuf[i, g], dsigf[i, g] = v[:4]
                dsigs[i, g, :] = v[4:]
        p.crod = dict(nb=nb, nstep=nstep, pos0=pos0, ssize=ssize, bpos=bpos, bmap=bmap, dsigtr=dsigtr, dsiga=dsiga,
                      dnuf=dnuf, dsigf=dsigf, dsigs=dsigs)
    if "EJCT" in cards and mode == "RODEJECT":   # mod_io.f90:2332-2440
        r = _Reader(cards["EJCT"], "EJCT")
        nb = p.crod["nb"]
        fb = np.array([r.floats(3) for _ in range(nb)])
        ttot, tstep1, tdiv, tstep2 = r.floats(4)
        if xtab is None:
            ibeta = np.array(r.floats(6)); lamb = np.array(r.floats(6)); velo = np.array(r.floats(ng))
        else:                            # %XTAB decks: per-material kinetics data come from the library (:2389)
            ibeta = lamb = velo = None
        p.ejct = dict(fbpos=fb[:, 0].copy(), tmove=fb[:, 1].copy(), bspeed=fb[:, 2].copy(), ttot=ttot, tstep1=tstep1,
                      tdiv=tdiv, tstep2=tstep2, ibeta=ibeta, lamb=lamb, velo=velo)
    if "EXTR" in cards:
        p.bextr = 1
    # feedback cards (mod_io.f90:2486-2957): reference value(s), then nmat x ng records of
    # d(sigtr, siga, nuf, sigf, sigs(1..ng)) per unit change of the parameter
    fbk = {}
    for card, key, nval in (("CBCS", "bcon", 1), ("BCON", "bcon", 2), ("FTEM", "ftem", 2), ("MTEM", "mtem", 2), ("CDEN", "cden", 2)):
        if card not in cards:
            continue
        if xtab is not None and not (card == "BCON" and mode == "RODEJECT"):
            continue                      # %XTAB decks: only %BCON of a RODEJECT deck is read (mod_io.f90:293-306)
        r = _Reader(cards[card], card)
        v = r.floats(nval)
        tab = dict(val=v[0], ref=v[-1], sigtr=np.zeros((nmat, ng)), siga=np.zeros((nmat, ng)), nuf=np.zeros((nmat, ng)),
                   sigf=np.zeros((nmat, ng)), sigs=np.zeros((nmat, ng, ng)))
        for i in range(nmat if xtab is None else 0):      # ... and only its two numbers (:2605)
            for g in range(ng):
                w = r.floats(4 + ng)
                tab["sigtr"][i, g], tab["siga"][i, g], tab["nuf"][i, g], tab["sigf"][i, g] = w[:4]
                tab["sigs"][i, g, :] = w[4:]
        fbk[key] = tab
    if fbk:
        p.fbk = fbk
    if "THER" in cards:                          # mod_io.f90:2996-3034 (derived data: Problem.th_setup)
        r = _Reader(cards["THER"], "THER")
        ppow = r.floats(1)[0]
        pow_ = r.floats(1)[0]
        tin, cmflow = r.floats(2)
        rf, tg, tc, ppitch = r.floats(4)
        nfpin, ngt = r.ints(2)
        cf = r.floats(1)[0]
        p.ther = dict(ppow=ppow, pow=pow_, tin=tin, cmflow=cmflow, rf=rf, tg=tg, tc=tc, ppitch=ppitch, nfpin=nfpin,
                      ngt=ngt, cf=cf)
    if "ESRC" in cards and mode == "FIXEDSRC":   # mod_io.f90:1369-1517
        r = _Reader(cards["ESRC"], "ESRC")
        nsrc = r.ints(1)[0]
        srcs = []
        for _ in range(nsrc):
            sden = r.floats(1)[0]
            spec = r.floats(ng)
            zlist = []
            while True:
                zpos = r.ints(1)[0]
                if zpos < 1:
                    break
                xy = []
                while True:
                    xpos, ypos = r.ints(2)
                    if xpos < 1 or ypos < 1:
                        break
                    xy.append((xpos, ypos))
                zlist.append((zpos, xy))
            srcs.append((sden, spec, zlist))
        p.esrc = srcs
    return p.build()


def read_deck(path: str) -> Problem:
    with open(path) as fh:
        return parse_deck(fh.read(), os.path.dirname(os.path.abspath(path)))
