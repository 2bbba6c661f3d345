"""Shock tube (src/pgen/shock_tube.cpp:306-...): constant L/R states split at xshock along
shock_dir; hydro or MHD (bxl, byl, bzl / bxr, byr, bzr)."""
import numpy as np

from ._util import active, coords, empty_state


def shock_tube(pmb, pin, out=None):
    m = pmb.pmy_mesh
    mhd = m.mhd
    sd = pin.get_integer("problem", "shock_dir")
    xs = pin.get_real("problem", "xshock")
    gm1 = pin.get_real("hydro", "gamma") - 1.0

    def side(s):
        w = {k: pin.get_real("problem", k + s) for k in ("d", "u", "v", "w", "p")}
        if mhd:
            for k in ("bx", "by", "bz"):
                w[k] = pin.get_real("problem", k + s)
        return w

    wl, wr = side("l"), side("r")
    c = coords(pmb)
    out = empty_state(pmb, mhd, out)
    k, j, i = active(pmb)
    xv = [c["x1v"][i][None, None, :], c["x2v"][j][None, :, None], c["x3v"][k][:, None, None]]
    shape = (pmb.ke - pmb.ks + 1, pmb.je - pmb.js + 1, pmb.ie - pmb.is_ + 1)
    left = np.broadcast_to(xv[sd - 1] < xs, shape)

    def pick(key):
        return np.where(left, wl[key], wr[key])

    d = pick("d")
    vel = [pick("u"), pick("v"), pick("w")]          # along shock_dir, then cyclic
    u = out["u"]
    u[0][k, j, i] = d
    for n in range(3):
        u[1 + (sd - 1 + n) % 3][k, j, i] = vel[n] * d
    e = pick("p") / gm1 + 0.5 * d * (vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2)
    if mhd:
        bb = [pick("bx"), pick("by"), pick("bz")]
        e = e + 0.5 * (bb[0] ** 2 + bb[1] ** 2 + bb[2] ** 2)
        names = ("b1", "b2", "b3")
        for n in range(3):
            comp = (sd - 1 + n) % 3
            arr = out[names[comp]]
            # face arrays: extend the cell-centred pattern by one face in their own direction
            sl = [k, j, i]
            rng = [(pmb.ks, pmb.ke), (pmb.js, pmb.je), (pmb.is_, pmb.ie)]
            ax = 2 - comp
            lo, hi = rng[ax]
            sl[ax] = slice(lo, hi + 2)
            src = bb[n]
            pad = [(0, 0)] * 3
            pad[ax] = (0, 1)
            arr[tuple(sl)] = np.pad(src, pad, mode="edge")
    u[4][k, j, i] = e
    return out
