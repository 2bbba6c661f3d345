"""Spherical blast wave (src/pgen/blast.cpp:94-212), Cartesian branch: ambient (damb, pamb),
over-pressured sphere of radius `radius` (prat), uniform field b0 at `angle` degrees in x1-x2."""
import numpy as np

from ._util import active, coords, empty_state


def blast(pmb, pin, out=None):
    m = pmb.pmy_mesh
    mhd = m.mhd
    rout = pin.get_real("problem", "radius")
    rin = rout - pin.get_or_add_real("problem", "ramp", 0.0)
    pa = pin.get_or_add_real("problem", "pamb", 1.0)
    da = pin.get_or_add_real("problem", "damb", 1.0)
    prat = pin.get_real("problem", "prat")
    drat = pin.get_or_add_real("problem", "drat", 1.0)
    b0 = pin.get_or_add_real("problem", "b0", 0.0) if mhd else 0.0
    angle = (np.pi / 180.0) * pin.get_or_add_real("problem", "angle", 0.0) if mhd else 0.0
    gm1 = pin.get_real("hydro", "gamma") - 1.0
    x0 = [pin.get_or_add_real("problem", "x%d_0" % d, 0.0) for d in (1, 2, 3)]
    c = coords(pmb)
    out = empty_state(pmb, mhd, out)
    k, j, i = active(pmb)
    X = c["x1v"][i][None, None, :]
    Y = c["x2v"][j][None, :, None]
    Z = c["x3v"][k][:, None, None]
    rad = np.sqrt((X - x0[0]) ** 2 + (Y - x0[1]) ** 2 + (Z - x0[2]) ** 2)
    den = np.full(rad.shape, da)
    pres = np.full(rad.shape, pa)
    inner = rad < rout
    if rin < rout:  # ramp region (blast.cpp:150-163)
        ramp = inner & (rad >= rin)
        f = (rad - rin) / (rout - rin)
        den = np.where(ramp, np.exp(np.log(drat * da) * (1 - f) + np.log(da) * f), den)
        pres = np.where(ramp, np.exp(np.log(prat * pa) * (1 - f) + np.log(pa) * f), pres)
        core = rad < rin
    else:
        core = inner
    den = np.where(core, drat * da, den)
    pres = np.where(core, prat * pa, pres)
    u = out["u"]
    u[0][k, j, i] = den
    u[4][k, j, i] = pres / gm1
    if mhd:
        out["b1"][k, j, slice(pmb.is_, pmb.ie + 2)] = b0 * np.cos(angle)
        out["b2"][k, slice(pmb.js, pmb.je + 2), i] = b0 * np.sin(angle)
        out["b3"][slice(pmb.ks, pmb.ke + 2), j, i] = 0.0
        u[4][k, j, i] += 0.5 * b0 * b0
    return out
