"""Helpers for the GPU parity tests: build a product Mesh (through the C ABI) from a golden
fixture / athinput parameter blocks, mirror state between the oracle and the device."""
import numpy as np

import athena_gamma_b200 as ab


def pin_from_par(par):
    pin = ab.ParameterInput()
    for b, kv in par.items():
        for k, v in kv.items():
            pin.set(b, k, v)
    return pin


def mesh_from_golden(g, **kw):
    pin = pin_from_par(g.par)
    m = ab.Mesh(pin, mhd=g.mhd, flux=g.solver, nghost=g.ng, nscalars=g.nscalars, eos=g.eos, **kw)
    import util
    for face, fn in util.user_bcs_for(g).items():
        m.enroll_user_boundary_function(face, fn)
    for n, loc in enumerate(g.locs):
        pmb = m.block_of(*loc)
        if pmb is None:
            continue
        for f in g.fields:
            pmb.set(f, g.init[n][f])
    return m


def copy_oracle_to_device(om, m, names):
    """Upload the oracle's current arrays `names` into the matching device registers."""
    for pmb in m.my_blocks:
        b = om.block_of(pmb.lx1, pmb.lx2, pmb.lx3)
        for nm in names:
            a = om.array(b, nm)
            if a is not None:
                pmb.set(nm, np.array(a))
