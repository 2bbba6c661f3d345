"""Orszag-Tang vortex (src/pgen/orszag_tang.cpp:40-130): B from the vector potential
Az = B0/(4 pi) (cos(4 pi x1) - 2 cos(2 pi x2)); d = 25/(36 pi), p = 5/(12 pi),
v = (v0 sin(2 pi x2), -v0 sin(2 pi x1), 0), B0 = 1/sqrt(4 pi), v0 = 1."""
import numpy as np

from ._util import active, add_magnetic_energy, coords, empty_state


def orszag_tang(pmb, pin, out=None):
    gm1 = pin.get_real("hydro", "gamma") - 1.0
    B0 = 1.0 / np.sqrt(4.0 * np.pi)
    d0 = 25.0 / (36.0 * np.pi)
    v0 = 1.0
    p0 = 5.0 / (12.0 * np.pi)
    c = coords(pmb)
    out = empty_state(pmb, True, out)
    k, j, i = active(pmb)
    x1f, x2f = c["x1f"], c["x2f"]
    az = B0 / (4.0 * np.pi) * (np.cos(4.0 * np.pi * x1f)[None, :]
                               - 2.0 * np.cos(2.0 * np.pi * x2f)[:, None])   # (nc2+1, nc1+1)
    dx1, dx2 = c["dx1f"], c["dx2f"]
    js, je, is_, ie = pmb.js, pmb.je, pmb.is_, pmb.ie
    b1 = (az[js + 1:je + 2, is_:ie + 2] - az[js:je + 1, is_:ie + 2]) / dx2[js:je + 1, None]
    b2 = (az[js:je + 2, is_:ie + 1] - az[js:je + 2, is_ + 1:ie + 2]) / dx1[None, is_:ie + 1]
    out["b1"][k, j, slice(is_, ie + 2)] = b1[None]
    out["b2"][k, slice(js, je + 2), i] = b2[None]
    X = c["x1v"][i][None, None, :]
    Y = c["x2v"][j][None, :, None]
    u = out["u"]
    u[0][k, j, i] = d0
    u[1][k, j, i] = d0 * v0 * np.sin(2.0 * np.pi * Y) + 0 * X
    u[2][k, j, i] = -d0 * v0 * np.sin(2.0 * np.pi * X) + 0 * Y
    u[4][k, j, i] = p0 / gm1 + 0.5 * (u[1][k, j, i] ** 2 + u[2][k, j, i] ** 2) / d0
    add_magnetic_energy(pmb, out)
    return out
