"""Kelvin-Helmholtz, iprob=1 (src/pgen/kh.cpp:60-110): slip surfaces at |x2| = 0.25, density
ratio drat, shear vflow, random velocity perturbations of amplitude amp.  The reference draws
the perturbations from ran2 seeded with -1-gid (kh.cpp:71); here a counter-based numpy
generator seeded with the same gid is used (synthetic data: not bit-identical to ran2)."""
import numpy as np

from ._util import active, coords, empty_state


def kh(pmb, pin, out=None):
    gm1 = pin.get_real("hydro", "gamma") - 1.0
    vflow = pin.get_real("problem", "vflow")
    drat = pin.get_real("problem", "drat")
    amp = pin.get_real("problem", "amp")
    c = coords(pmb)
    out = empty_state(pmb, False, out)
    k, j, i = active(pmb)
    shape = (pmb.ke - pmb.ks + 1, pmb.je - pmb.js + 1, pmb.ie - pmb.is_ + 1)
    rng = np.random.default_rng(1 + pmb.gid)
    Y = np.broadcast_to(c["x2v"][j][None, :, None], shape)
    inner = np.abs(Y) < 0.25
    d = np.where(inner, drat, 1.0)
    u = out["u"]
    u[0][k, j, i] = d
    m1 = np.where(inner, -drat * (vflow + amp * (rng.random(shape) - 0.5)),
                  vflow + amp * (rng.random(shape) - 0.5))
    m2 = np.where(inner, drat * amp * (rng.random(shape) - 0.5), amp * (rng.random(shape) - 0.5))
    u[1][k, j, i] = m1
    u[2][k, j, i] = m2
    u[4][k, j, i] = 2.5 / gm1 + 0.5 * (m1 ** 2 + m2 ** 2) / d
    return out
