"""z-slab decomposition used by the multi-GPU path (host-side logic, no CUDA needed).

The core is numbered k-major with a constant number of nodes per plane (mod_io.f90:1319-1330),
so rank r owns the contiguous planes [k0, k1) = node range [k0*np, k1*np).  The same formula is
used inside the library (adp_set_geometry in csrc/capi.cu); tests check that they agree.
"""
from __future__ import annotations

import numpy as np


def slab_planes(nzz: int, nranks: int, rank: int):
    """Planes [k0, k1) (0-based) of `rank`: as even as possible, the first nzz % nranks ranks get one more."""
    base, rem = divmod(nzz, nranks)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def exchange_halo(dist, local: np.ndarray, npl: int, nplanes: int, rank: int, nranks: int, gh: int = 2):
    """Reference implementation of the ghost-plane exchange on the CPU (torch.distributed, gloo).
    `local` holds [gh ghost planes | owned planes | gh ghost planes]; the `nplanes` (<= gh) planes
    adjacent to the owned block are filled from the neighbours.  Mirrors adp_comm_halo (csrc/comm.cu)."""
    import torch
    n = nplanes * npl
    t = torch.from_numpy(local)
    lo, hi = gh * npl, local.size - gh * npl            # owned block [lo, hi)
    reqs = []
    if rank + 1 < nranks:
        reqs.append(dist.isend(t[hi - n:hi].clone(), rank + 1))
        reqs.append(dist.irecv(t[hi:hi + n], rank + 1))
    if rank > 0:
        reqs.append(dist.isend(t[lo:lo + n].clone(), rank - 1))
        reqs.append(dist.irecv(t[lo - n:lo], rank - 1))
    for r in reqs:
        r.wait()
    return local


def slab_spmv(a_dia: np.ndarray, x_local: np.ndarray, ypm: np.ndarray, ypp: np.ndarray, npl: int, nplanes_own: int, gh: int):
    """y = A x on the owned rows from the 7 diagonals (set_ind order z-,y-,x-,diag,x+,y+,z+) with the
    device layout of csrc/adp_internal.cuh: x_local has gh ghost planes per side, neighbours are
    idx-+npl (z), idx-+ypm/ypp[r] (y), idx-+1 (x); absent neighbours have zero coefficients."""
    nown = npl * nplanes_own
    idx = np.arange(nown) + gh * npl
    r = np.arange(nown) % npl
    y = np.zeros(nown)
    for d, off in enumerate((-npl, -ypm[r], -1, 0, 1, ypp[r], npl)):
        y = y + a_dia[d] * x_local[idx + off]
    return y
