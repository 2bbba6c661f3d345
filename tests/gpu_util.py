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
    if util.user_source_for(g):
        m.enroll_user_explicit_source_function(util.user_source_for(g))
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


def central_gravity_source_torch(par):
    """Device variant of util.central_gravity_source: the same operation sequence as separate
    torch kernels (one IEEE rounding each, no FMA contraction) on the library's stream."""
    import torch
    pr = par.get("problem", {})
    gm = float(pr.get("gm", 0.5))
    soft2 = float(pr.get("soft2", 0.01))
    sdecay = float(pr.get("sdecay", 0.3))

    def fn(pmb, time, dt, prim, prim_scalar, bcc, cons, cons_scalar, stream):
        st = torch.cuda.ExternalStream(stream)
        with torch.cuda.stream(st):
            w = torch.as_tensor(prim, device="cuda")
            u = torch.as_tensor(cons, device="cuda")
            amp = gm*(1.0 + 0.5*time)
            K = slice(pmb.ks, pmb.ke + 1)
            J = slice(pmb.js, pmb.je + 1)
            I = slice(pmb.is_, pmb.ie + 1)
            x = torch.as_tensor(pmb.coord("x1v")[I], device="cuda")[None, None, :]
            y = torch.as_tensor(pmb.coord("x2v")[J], device="cuda")[None, :, None]
            z = torch.as_tensor(pmb.coord("x3v")[K], device="cuda")[:, None, None]
            rsq = (x*x + y*y) + (z*z + soft2)
            r = torch.sqrt(rsq)
            # tensor / tensor: torch evaluates (python scalar)/tensor as scalar*reciprocal(tensor),
            # which rounds differently from the IEEE division the C++ function performs
            fac = torch.full_like(rsq, amp)/(rsq*r)
            den = w[0, K, J, I]
            s1 = (dt*den)*(fac*x)
            s2 = (dt*den)*(fac*y)
            s3 = (dt*den)*(fac*z)
            u[1, K, J, I] -= s1
            u[2, K, J, I] -= s2
            u[3, K, J, I] -= s3
            if u.shape[0] > 4:
                u[4, K, J, I] -= (s1*w[1, K, J, I] + s2*w[2, K, J, I]) + s3*w[3, K, J, I]
            if cons_scalar is not None:
                s = torch.as_tensor(cons_scalar, device="cuda")
                rr = torch.as_tensor(prim_scalar, device="cuda")
                for n in range(s.shape[0]):
                    s[n, K, J, I] -= (dt*sdecay)*(den*rr[n, K, J, I])
    return fn
