"""Shared helpers for the parity tests (golden fixtures <-> oracle <-> CUDA path)."""
import glob
import json
import os

import numpy as np

import oracle

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(include_smr=False):
    """fixtures; the static-mesh-refinement ones (smr_*) are oracle-only so far"""
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if include_smr or not n.startswith("smr_")]


def gpu_params(names):
    """pytest params of golden names (every fixture is a hard requirement on the GPU)"""
    return list(names)


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.meta = json.loads(str(z["meta"]))
        self.par = self.meta["par"]
        self.mhd = bool(self.meta["mhd"])
        self.solver = self.meta["solver"]
        self.ng = int(self.meta["nghost"])
        self.ncycles = int(self.meta["ncycles"])
        self.nscalars = int(self.meta.get("nscalars", 0))
        self.eos = self.meta.get("eos", "adiabatic")
        self.dts = z["dts"]
        # (lx1, lx2, lx3) -- plus the level on a refined mesh, where lx alone is ambiguous
        self.locs = [tuple(int(v) for v in l) for l in z["locs"]]
        self.final_time = float(z["final_time"])
        self.final_dt = float(z["final_dt"])
        self.hst = z["hst"] if "hst" in z.files else None
        self.fields = ("u", "b1", "b2", "b3") if self.mhd else ("u",)
        if self.nscalars:
            self.fields += ("s",)
        self.init = [{f: z["init_%s_%d" % (f, n)] for f in self.fields}
                     for n in range(len(self.locs))]
        self.final = [{f: z["final_%s_%d" % (f, n)] for f in self.fields}
                      for n in range(len(self.locs))]

    def as_rst(self, which="init"):
        blocks = getattr(self, which)
        return {"blocks": [dict(loc=(self.locs[n] + (0,))[:4], **blocks[n])
                           for n in range(len(self.locs))]}


def shock_cloud_inner_x1(par):
    """The user boundary function of pgen/shk_cloud.cpp:196-210 (ShockCloudInnerX1): holds the
    inner-x1 ghost zones at the post-shock state; constants as in ProblemGenerator :71-93."""
    gmma = float(par["hydro"]["gamma"])
    gmma1 = gmma - 1.0
    mach = float(par["problem"]["Mach"])
    dr, pr, ur = 1.0, 1.0/gmma, 0.0
    jump1 = (gmma + 1.0)/(gmma1 + 2.0/(mach*mach))
    jump2 = (2.0*gmma*mach*mach - gmma1)/(gmma + 1.0)
    jump3 = 2.0*(1.0 - 1.0/(mach*mach))/(gmma + 1.0)
    dl = dr*jump1
    pl = pr*jump2
    ul = ur + jump3*mach*float(np.sqrt(gmma*pr/dr))

    def fn(pmb, pco, prim, b, time, dt, il, iu, jl, ju, kl, ku, ngh):
        for i in range(1, ngh + 1):
            prim[0, kl:ku+1, jl:ju+1, il-i] = dl
            prim[1, kl:ku+1, jl:ju+1, il-i] = ul
            prim[2, kl:ku+1, jl:ju+1, il-i] = 0.0
            prim[3, kl:ku+1, jl:ju+1, il-i] = 0.0
            prim[4, kl:ku+1, jl:ju+1, il-i] = pl
    return fn


def central_gravity_source(par):
    """The SrcTermFunc of oracle/pgen/usersrc.cpp (CentralGravity), restated with the same
    operation order (only + - * / sqrt, so numpy reproduces it bit for bit)."""
    pr = par.get("problem", {})
    gm = float(pr.get("gm", 0.5))
    soft2 = float(pr.get("soft2", 0.01))
    sdecay = float(pr.get("sdecay", 0.3))

    def fn(pmb, time, dt, prim, prim_scalar, bcc, cons, cons_scalar):
        amp = gm*(1.0 + 0.5*time)
        K = slice(pmb.ks, pmb.ke + 1)
        J = slice(pmb.js, pmb.je + 1)
        I = slice(pmb.is_, pmb.ie + 1)
        x = pmb.coord("x1v")[I][None, None, :]
        y = pmb.coord("x2v")[J][None, :, None]
        z = pmb.coord("x3v")[K][:, None, None]
        rsq = (x*x + y*y) + (z*z + soft2)
        r = np.sqrt(rsq)
        fac = amp/(rsq*r)
        den = prim[0, K, J, I]
        s1 = (dt*den)*(fac*x)
        s2 = (dt*den)*(fac*y)
        s3 = (dt*den)*(fac*z)
        cons[1, K, J, I] -= s1
        cons[2, K, J, I] -= s2
        cons[3, K, J, I] -= s3
        if cons.shape[0] > 4:
            cons[4, K, J, I] -= (s1*prim[1, K, J, I] + s2*prim[2, K, J, I]) + s3*prim[3, K, J, I]
        if cons_scalar is not None:
            for n in range(cons_scalar.shape[0]):
                cons_scalar[n, K, J, I] -= (dt*sdecay)*(den*prim_scalar[n, K, J, I])
    return fn


def user_source_for(g):
    if g.name.startswith("usersrc"):
        return central_gravity_source(g.par)
    return None


def user_bcs_for(g):
    """{face: boundary function} a fixture needs (BoundaryFace numbering ix1=0 ... ox3=5)"""
    if g.name.startswith("shkcloud"):
        return {0: shock_cloud_inner_x1(g.par)}
    return {}


def oracle_from_golden(g):
    p = oracle.params_from_athinput(g.par, g.mhd, g.solver, ng=g.ng, nscalars=g.nscalars,
                                    eos=g.eos)
    m = oracle.OracleMesh(p)
    for face, fn in user_bcs_for(g).items():
        m.enroll_user_boundary_function(face, fn)
    if user_source_for(g):
        m.enroll_user_explicit_source_function(user_source_for(g))
    m.load_rst(g.as_rst("init"))
    m.initialize()
    return m


def assert_bitwise(a, b, what=""):
    """numeric equality of every element (treats +0 == -0, NaN != NaN fails)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b):
        bad = np.argwhere(a != b)
        i = tuple(bad[0])
        raise AssertionError("%s: %d/%d elements differ; first at %s: %.17e vs %.17e (max |d| %.3e)"
                             % (what, len(bad), a.size, i, a[i], b[i],
                                np.nanmax(np.abs(a - b))))
