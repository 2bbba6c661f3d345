import numpy as np


def active(pmb):
    """slices of the active zone (k, j, i)"""
    return (slice(pmb.ks, pmb.ke + 1), slice(pmb.js, pmb.je + 1), slice(pmb.is_, pmb.ie + 1))


def empty_state(pmb, mhd, out=None):
    """AthenaArray storage is zero-initialised (athena_arrays.hpp:527-538).  `out` lets the
    caller provide the buffers (e.g. pinned host memory) instead of allocating new ones."""
    names = ("u", "b1", "b2", "b3") if mhd else ("u",)
    if out is not None:
        for n in names:
            assert out[n].shape == pmb.shape(n), (n, out[n].shape)
            out[n][...] = 0.0
        return out
    return {n: np.zeros(pmb.shape(n)) for n in names}


def coords(pmb):
    return {n: pmb.coord(n) for n in ("x1f", "x2f", "x3f", "x1v", "x2v", "x3v", "dx1f", "dx2f",
                                      "dx3f")}


def add_magnetic_energy(pmb, out):
    """u(IEN) += 0.5*(bcc^2) from face averages (pgen/blast.cpp:197-208 pattern)"""
    k, j, i = active(pmb)
    b1, b2, b3 = out["b1"], out["b2"], out["b3"]
    k1 = slice(pmb.ks + 1, pmb.ke + 2) if pmb.ncells3 > 1 else slice(1, 2)
    j1 = slice(pmb.js + 1, pmb.je + 2) if pmb.ncells2 > 1 else slice(1, 2)
    i1 = slice(pmb.is_ + 1, pmb.ie + 2)
    bc1 = 0.5 * (b1[k, j, i] + b1[k, j, i1])
    bc2 = 0.5 * (b2[k, j, i] + b2[k, j1, i])
    bc3 = 0.5 * (b3[k, j, i] + b3[k1, j, i])
    out["u"][4][k, j, i] += 0.5 * (bc1 ** 2 + bc2 ** 2 + bc3 ** 2)
