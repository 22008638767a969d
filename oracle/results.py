"""CPU oracle for the result reductions AsmPow / AxiPow / AsmFlux (reference: src/mod_io.f90).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product (adp_asm_pow, adp_axi_pow,
adp_asm_flux in adpres_b200/csrc/results.cu) never imports this.  Parity unpinned: the reference
repository holds no printed power map; the functions below follow the cited loops statement by
statement -- same accumulation order, same operand order -- in plain Python / numpy (the loops
over k are vectorised across columns only, which keeps every column's serial order).
"""
import numpy as np


def _dense(p, fn):
    """fx = 0; fx(ix(n), iy(n), iz(n)) = fn(n)      (mod_io.f90:3297-3300)"""
    fx = np.zeros((p.nxx, p.nyy, p.nzz))
    fx[p.ix - 1, p.iy - 1, p.iz - 1] = fn
    return fx


def _assembly_average(p, fnode):
    """mod_io.f90:3317-3341 (AsmPow) = 3557-3589 (AsmFlux): rectangle sums, ly outer, lx inner;
    vsumm counts the whole rectangle, also outside the jagged core outline."""
    fasm = np.zeros((p.nx, p.ny))
    ys, yf = 1, 0
    for j in range(1, p.ny + 1):
        yf += int(p.ydiv[j - 1])
        xf, xs = 0, 1
        for i in range(1, p.nx + 1):
            xf += int(p.xdiv[i - 1])
            summ = vsumm = 0.0
            for ly in range(ys, yf + 1):
                for lx in range(xs, xf + 1):
                    summ = summ + fnode[lx - 1, ly - 1] * p.xdel[lx - 1] * p.ydel[ly - 1]
                    vsumm = vsumm + p.xdel[lx - 1] * p.ydel[ly - 1]
            fasm[i - 1, j - 1] = summ / vsumm
            xs += int(p.xdiv[i - 1])
        ys += int(p.ydiv[j - 1])
    return fasm


def _inside(p):
    """(nxx, nyy) mask of DO i = ystag(j)%smin, ystag(j)%smax"""
    m = np.zeros((p.nxx, p.nyy), dtype=bool)
    for j in range(p.nyy):
        m[p.ystag_smin[j] - 1:p.ystag_smax[j], j] = True
    return m


def asm_pow(p, fn):
    """AsmPow (mod_io.f90:3267-3405) -> (fasm(nx,ny), xmax, ymax)"""
    fx = _dense(p, fn)
    summ = np.zeros((p.nxx, p.nyy))
    vsumm = 0.0
    for k in range(p.nzz):                                   # :3305-3311, serial in k per column
        summ = summ + fx[:, :, k] * p.zdel[k]
        vsumm = vsumm + p.zdel[k]
    fnode = np.where(_inside(p), summ / vsumm, 0.0)
    fasm = _assembly_average(p, fnode)
    nfuel, totp = 0, 0.0
    for j in range(p.ny):
        for i in range(p.nx):
            if fasm[i, j] > 0.0:
                nfuel += 1
                totp = totp + fasm[i, j]
    xmax = ymax = 1
    fmax = 0.0
    for j in range(p.ny):
        for i in range(p.nx):
            if totp > 0.0:
                fasm[i, j] = float(np.float32(nfuel)) / totp * fasm[i, j]      # REAL(nfuel)
            if fasm[i, j] > fmax:
                xmax, ymax, fmax = i + 1, j + 1, fasm[i, j]
    return fasm, xmax, ymax


def axi_pow(p, fn):
    """AxiPow (mod_io.f90:3409-3494) -> (faxi(nz), amax); the node loop is serial (lz, j, i),
    i.e. node-number order inside every plane."""
    npl = p.npl
    faxi = np.zeros(p.nz)
    nfuel, totp, ztot = 0, 0.0, 0
    for k in range(p.nz):
        summ = vsumm = 0.0
        for _ in range(int(p.zdiv[k])):
            for n in range(ztot * npl, (ztot + 1) * npl):
                summ = summ + fn[n]
                vsumm = vsumm + p.vdel[n]
            ztot += 1
        faxi[k] = summ / vsumm
        if faxi[k] > 0.0:
            nfuel += 1
            totp = totp + faxi[k]
    fmax, amax = 0.0, 1
    for k in range(p.nz):
        faxi[k] = float(np.float32(nfuel)) / totp * faxi[k]
        if faxi[k] > fmax:
            amax, fmax = k + 1, faxi[k]
    return faxi, amax


def asm_flux(p, f, norm=None):
    """AsmFlux (mod_io.f90:3498-3644) -> (fasm(nx,ny,ng), negf)"""
    G = f.shape[1]
    out = np.zeros((p.nx, p.ny, G))
    negf = 0
    inside = _inside(p)
    for g in range(G):
        fx = _dense(p, f[:, g])
        summ = np.zeros((p.nxx, p.nyy))
        vsumm = np.zeros((p.nxx, p.nyy))
        xd, yd = p.xdel[:, None], p.ydel[None, :]
        for k in range(p.nzz):                               # :3545-3551
            summ = summ + fx[:, :, k] * xd * yd * p.zdel[k]
            vsumm = vsumm + xd * yd * p.zdel[k]
        fnode = np.where(inside, summ / vsumm, 0.0)
        fasm = _assembly_average(p, fnode)
        totp = 0.0
        for j in range(p.ny):
            for i in range(p.nx):
                if fasm[i, j] > 0.0:
                    totp = totp + fasm[i, j]
                if fasm[i, j] < 0.0:
                    negf = 1
        if norm is not None:
            fasm = norm / totp * fasm * norm                 # sic (:3598)
        out[:, :, g] = fasm
    return out, negf
