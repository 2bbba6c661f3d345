"""ctypes wrapper of the CPU oracle (oracle/libathena_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg -- never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BC = {"periodic": 0, "outflow": 1, "reflecting": 2, "user": 3}
SOLVER = {"hlle": 0, "hllc": 1, "hlld": 2, "roe": 3, "lhllc": 4, "lhlld": 5, "llf": 6}
INTEGRATOR = {"vl2": 0, "rk2": 1, "rk1": 2, "rk3": 3}
DEFAULT_FLOOR = float(np.sqrt(1024 * float(np.finfo(np.float32).tiny)))  # eos ctor


class AoParams(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int),
                ("bx1", C.c_int), ("bx2", C.c_int), ("bx3", C.c_int),
                ("x1min", C.c_double), ("x1max", C.c_double),
                ("x2min", C.c_double), ("x2max", C.c_double),
                ("x3min", C.c_double), ("x3max", C.c_double),
                ("bc", C.c_int * 6), ("ng", C.c_int), ("mhd", C.c_int),
                ("solver", C.c_int), ("xorder", C.c_int), ("integrator", C.c_int),
                ("gamma", C.c_double), ("dfloor", C.c_double), ("pfloor", C.c_double),
                ("cfl", C.c_double), ("tlim", C.c_double), ("start_time", C.c_double),
                ("nscalars", C.c_int), ("eos", C.c_int), ("sfloor", C.c_double),
                ("iso_cs", C.c_double), ("grav_acc", C.c_double * 3), ("char_proj", C.c_int),
                ("xrat", C.c_double * 3), ("nref", C.c_int), ("ref", (C.c_double * 6) * 8),
                ("ref_level", C.c_int * 8)]


# AoBValFunc (athena_oracle.h): user-enrolled boundary function with plain arrays
_DP = C.POINTER(C.c_double)
BVALFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_int, _DP, _DP, _DP, _DP, C.c_double, C.c_double,
                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


SRCTERMFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, C.c_double, _DP, _DP, _DP, _DP,
                          _DP)


class FaceFieldView:
    """FaceField (src/athena.hpp:95-105) as numpy views"""
    def __init__(self, x1f, x2f, x3f):
        self.x1f, self.x2f, self.x3f = x1f, x2f, x3f


def build():
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "libathena_oracle.so")
        srcs = [os.path.join(HERE, f) for f in
                ("oracle_physics.c", "oracle_mesh.c", "oracle_smr.c", "oracle_internal.h",
                 "athena_oracle.h")]
        if (not os.path.exists(so)
                or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)):
            build()
        L = C.CDLL(so)
        L.ao_create.restype = C.c_void_p
        L.ao_create.argtypes = [C.POINTER(AoParams)]
        L.ao_destroy.argtypes = [C.c_void_p]
        L.ao_nblocks.argtypes = [C.c_void_p]
        L.ao_block_level.argtypes = [C.c_void_p, C.c_int]
        L.ao_block_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long)]
        L.ao_array.restype = C.POINTER(C.c_double)
        L.ao_array.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_long)]
        for f in ("ao_initialize", "ao_emf_exchange", "ao_exchange_cc", "ao_exchange_fc",
                  "ao_exchange_scalars"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.ao_cycle.restype = C.c_double
        L.ao_cycle.argtypes = [C.c_void_p]
        L.ao_time.restype = C.c_double
        L.ao_time.argtypes = [C.c_void_p]
        L.ao_dt.restype = C.c_double
        L.ao_dt.argtypes = [C.c_void_p]
        L.ao_ncycle.argtypes = [C.c_void_p]
        L.ao_set_time_dt.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ao_calc_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ao_calc_scalar_fluxes.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ao_integrate_scalars.argtypes = [C.c_void_p, C.c_int, C.c_int]
        for f in ("ao_corner_e", "ao_swap_cc", "ao_swap_fc", "ao_zero_reg1", "ao_primitives",
                  "ao_physical_bcs"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int]
        for f in ("ao_weighted_ave_cc", "ao_weighted_ave_fc"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_double)]
        L.ao_add_flux_div.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.ao_add_source_terms.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double]
        L.ao_ct.argtypes = [C.c_void_p, C.c_int, C.c_double]
        for f in ("ao_cons2prim", "ao_prim2cons", "ao_scalar_cons2prim", "ao_scalar_prim2cons"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 6
        L.ao_enroll_user_source.argtypes = [C.c_void_p, SRCTERMFUNC, C.c_void_p]
        L.ao_enroll_user_bc.argtypes = [C.c_void_p, C.c_int, BVALFUNC, C.c_void_p]
        L.ao_history.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ao_new_block_dt.restype = C.c_double
        L.ao_new_block_dt.argtypes = [C.c_void_p, C.c_int]
        dp = C.POINTER(C.c_double)
        L.ao_riemann.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, dp, C.c_double,
                                 C.c_double, C.c_double, dp, dp]
        L.ao_riemann_dv.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, dp, dp, dp, C.c_double,
                                    C.c_double, C.c_double, dp, dp]
        L.ao_riemann_iso.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, dp, C.c_double,
                                     C.c_double, dp]
        L.ao_recon_char.argtypes = [C.c_int, C.c_int, C.c_long, dp, dp, C.c_double, C.c_double,
                                    C.c_double, C.c_double, C.c_double, dp, dp]
        L.ao_plm.argtypes = [C.c_long, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, dp]
        L.ao_ppm.argtypes = [C.c_long, C.c_int, dp, dp, dp, dp, dp, C.c_double, C.c_double,
                             dp, dp]
        ip = C.c_int
        L.ao_recon_line.argtypes = [ip]*7 + [dp, dp, dp, ip, dp, ip, ip, dp, dp]
        L.ao_recon_line_char.argtypes = [ip]*8 + [dp, dp, dp, dp, dp, C.c_double, C.c_double,
                                                  C.c_double, ip, ip, dp, dp]
        L.ao_bcc_weights.argtypes = [ip]*6 + [dp]*5
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def params_from_athinput(par, mhd, solver, ng=None, nscalars=0, eos="adiabatic"):
    """par: dict of blocks (oracle/ref_run.parse_athinput or the product's ParameterInput)."""
    mesh, t = par["mesh"], par["time"]
    mb = par.get("meshblock", {})
    xorder = int(str(t.get("xorder", "2")).rstrip("c"))
    p = AoParams()
    p.nx1, p.nx2, p.nx3 = int(mesh["nx1"]), int(mesh.get("nx2", 1)), int(mesh.get("nx3", 1))
    p.bx1 = int(mb.get("nx1", p.nx1))
    p.bx2 = int(mb.get("nx2", p.nx2))
    p.bx3 = int(mb.get("nx3", p.nx3))
    p.x1min, p.x1max = float(mesh["x1min"]), float(mesh["x1max"])
    p.x2min, p.x2max = float(mesh.get("x2min", -0.5)), float(mesh.get("x2max", 0.5))
    p.x3min, p.x3max = float(mesh.get("x3min", -0.5)), float(mesh.get("x3max", 0.5))
    for i, k in enumerate(("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc")):
        p.bc[i] = BC[mesh.get(k, "periodic")]
    p.ng = ng if ng is not None else (3 if xorder == 3 else 2)
    p.mhd = int(bool(mhd))
    p.solver = SOLVER[solver]
    p.xorder = xorder
    p.char_proj = int(str(t.get("xorder", "2")).endswith("c"))
    p.integrator = INTEGRATOR[t.get("integrator", "vl2")]
    h = par.get("hydro", {})
    p.eos = 1 if eos == "isothermal" else 0
    p.gamma = float(h["gamma"]) if p.eos == 0 else 0.0
    p.iso_cs = float(h.get("iso_sound_speed", 0.0))
    p.nscalars = int(nscalars)
    p.sfloor = float(h.get("sfloor", DEFAULT_FLOOR))
    for d in range(3):
        p.grav_acc[d] = float(h.get("grav_acc%d" % (d + 1), 0.0))
    p.dfloor = float(h.get("dfloor", DEFAULT_FLOOR))
    p.pfloor = float(h.get("pfloor", DEFAULT_FLOOR))
    p.cfl = float(t["cfl_number"])
    p.tlim = float(t["tlim"])
    p.start_time = float(t.get("start_time", 0.0))
    for d in range(3):
        p.xrat[d] = float(mesh.get("x%drat" % (d + 1), 1.0))
    # <refinementN> blocks in input order (Mesh ctor, src/mesh/mesh.cpp:330-465)
    p.nref = 0
    if mesh.get("refinement", "none") == "static":
        lim = {"x1min": p.x1min, "x1max": p.x1max, "x2min": p.x2min, "x2max": p.x2max,
               "x3min": p.x3min, "x3max": p.x3max}
        for name, blk in par.items():
            if not name.startswith("refinement"):
                continue
            for c, k in enumerate(("x1min", "x1max", "x2min", "x2max", "x3min", "x3max")):
                p.ref[p.nref][c] = float(blk.get(k, lim[k]))
            p.ref_level[p.nref] = int(blk["level"])
            p.nref += 1
    return p


class OracleMesh:
    def __init__(self, params):
        self.L = lib()
        self.p = params
        self.h = self.L.ao_create(C.byref(params))
        if not self.h:
            raise ValueError("oracle: static refinement is restated for hydro without scalars on "
                             "uniformly spaced levels, even MeshBlock sizes and even NGHOST only")
        self.nb = self.L.ao_nblocks(self.h)
        self.info = []
        for b in range(self.nb):
            out = (C.c_long * 12)()
            self.L.ao_block_info(self.h, b, out)
            self.info.append(dict(zip(("lx1", "lx2", "lx3", "nc1", "nc2", "nc3", "is", "ie",
                                       "js", "je", "ks", "ke"), list(out))))
            self.info[-1]["level"] = self.L.ao_block_level(self.h, b)

    def __del__(self):
        try:
            self.L.ao_destroy(self.h)
        except Exception:
            pass

    def shape(self, b, name):
        i = self.info[b]
        n1, n2, n3 = i["nc1"], i["nc2"], i["nc3"]
        nh = 4 if self.p.eos == 1 else 5      # NHYDRO (configure.py:374-377)
        if name in ("u", "u1", "w"):
            return (nh, n3, n2, n1)
        ns = self.p.nscalars
        if name in ("s", "s1", "r"):
            return (ns, n3, n2, n1)
        if name in ("sflux1", "sflux2", "sflux3"):
            d = int(name[-1]) - 1
            sh = [n3, n2, n1]
            sh[2 - d] += 1
            return (ns,) + tuple(sh)
        if name in ("bcc", "cc_e"):
            return (3, n3, n2, n1)
        if name in ("b1", "b1_1", "wght1", "e2_x1f", "e3_x1f"):
            return (n3, n2, n1 + 1)
        if name in ("b2", "b1_2", "wght2", "e1_x2f", "e3_x2f"):
            return (n3, n2 + 1, n1)
        if name in ("b3", "b1_3", "wght3", "e1_x3f", "e2_x3f"):
            return (n3 + 1, n2, n1)
        if name == "flux1":
            return (nh, n3, n2, n1 + 1)
        if name == "flux2":
            return (nh, n3, n2 + 1, n1)
        if name == "flux3":
            return (nh, n3 + 1, n2, n1)
        if name == "e1":
            return (n3 + 1, n2 + 1, n1)
        if name == "e2":
            return (n3 + 1, n2, n1 + 1)
        if name == "e3":
            return (n3, n2 + 1, n1 + 1)
        return None

    def array(self, b, name):
        """numpy VIEW of the oracle's array (re-fetch after swaps: pointers move)."""
        n = C.c_long()
        ptr = self.L.ao_array(self.h, b, name.encode(), C.byref(n))
        if not ptr or n.value == 0:
            return None
        a = np.ctypeslib.as_array(ptr, shape=(n.value,))
        shp = self.shape(b, name)
        return a.reshape(shp) if shp else a

    def block_of(self, lx1, lx2, lx3, level=None):
        """level: only needed (and only checked) on a refined mesh"""
        for b, i in enumerate(self.info):
            if (i["lx1"], i["lx2"], i["lx3"]) == (lx1, lx2, lx3) and (
                    level is None or self.p.nref == 0 or i["level"] == level):
                return b
        raise KeyError((lx1, lx2, lx3, level))

    def load_rst(self, rst):
        """Fill u / b of every block from a parsed reference restart dump (ref_run.read_rst)."""
        for blk in rst["blocks"]:
            b = self.block_of(*blk["loc"][:4])
            self.array(b, "u")[...] = blk["u"]
            if self.p.nscalars > 0:
                self.array(b, "s")[...] = blk["s"]
            if self.p.mhd:
                for nm in ("b1", "b2", "b3"):
                    self.array(b, nm)[...] = blk[nm]

    def enroll_user_boundary_function(self, face, fn):
        """Mesh::EnrollUserBoundaryFunction: fn(pmb, pco, prim, b, time, dt, il, iu, jl, ju,
        kl, ku, ngh) with prim / b.x?f writable numpy views; pmb = pco = OracleBlockView."""
        mesh = self

        def tramp(_user, blk, prim, b1, b2, b3, time, dt, il, iu, jl, ju, kl, ku, ngh):
            view = OracleBlockView(mesh, blk)
            w = np.ctypeslib.as_array(prim, shape=mesh.shape(blk, "w"))
            bf = None
            if mesh.p.mhd:
                bf = FaceFieldView(*[np.ctypeslib.as_array(p, shape=mesh.shape(blk, nm))
                                     for p, nm in ((b1, "b1"), (b2, "b2"), (b3, "b3"))])
            fn(view, view, w, bf, time, dt, il, iu, jl, ju, kl, ku, ngh)
        cb = BVALFUNC(tramp)
        self._keep = getattr(self, "_keep", []) + [cb]
        self.L.ao_enroll_user_bc(self.h, face, cb, None)

    def enroll_user_explicit_source_function(self, fn):
        """Mesh::EnrollUserExplicitSourceFunction: fn(pmb, time, dt, prim, prim_scalar, bcc,
        cons, cons_scalar) with numpy views (None where the array does not exist)."""
        mesh = self

        def view(ptr, blk, name):
            shp = mesh.shape(blk, name)
            if not ptr or shp is None or 0 in shp:
                return None
            return np.ctypeslib.as_array(ptr, shape=shp)

        def tramp(_user, blk, time, dt, prim, prs, bcc, cons, cs):
            fn(OracleBlockView(mesh, blk), time, dt, view(prim, blk, "w"),
               view(prs, blk, "r") if mesh.p.nscalars else None,
               view(bcc, blk, "bcc") if mesh.p.mhd else None, view(cons, blk, "u"),
               view(cs, blk, "s") if mesh.p.nscalars else None)
        cb = SRCTERMFUNC(tramp)
        self._keep = getattr(self, "_keep", []) + [cb]
        self.L.ao_enroll_user_source(self.h, cb, None)

    def history(self):
        """the sums HistoryOutput writes (mass, momenta, KE, tot-E, [ME], [scalars])"""
        out = np.zeros(32)
        n = self.L.ao_history(self.h, _dp(out))
        return out[:n]

    def initialize(self):
        self.L.ao_initialize(self.h)

    def cycle(self):
        return self.L.ao_cycle(self.h)

    @property
    def time(self):
        return self.L.ao_time(self.h)

    @property
    def dt(self):
        return self.L.ao_dt(self.h)

    def set_time_dt(self, t, dt):
        self.L.ao_set_time_dt(self.h, t, dt)


class OracleBlockView:
    """what a boundary function may ask of pmb / pco: index ranges and coordinates"""
    def __init__(self, mesh, b):
        self.mesh, self.b = mesh, b
        i = mesh.info[b]
        self.lx1, self.lx2, self.lx3 = i["lx1"], i["lx2"], i["lx3"]
        self.is_, self.ie, self.js, self.je, self.ks, self.ke = (i["is"], i["ie"], i["js"],
                                                                 i["je"], i["ks"], i["ke"])

    def coord(self, name):
        return np.array(self.mesh.array(self.b, name))


def riemann(solver, mhd, wl, wr, bx, gamma, dt=0.0, dx=1.0, dvn=None, dvt=None):
    """wl, wr: (nwave, n) sweep-ordered primitives. Returns flux (nwave, n), wct (n)."""
    L = lib()
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    n = wl.shape[1]
    bx = np.ascontiguousarray(bx if bx is not None else np.zeros(n), dtype=np.float64)
    flx = np.zeros_like(wl)
    wct = np.zeros(n)
    dvn = np.ascontiguousarray(dvn if dvn is not None else np.zeros(n), dtype=np.float64)
    dvt = np.ascontiguousarray(dvt if dvt is not None else np.zeros(n), dtype=np.float64)
    L.ao_riemann_dv(SOLVER[solver], int(mhd), n, _dp(wl), _dp(wr), _dp(bx), _dp(dvn), _dp(dvt),
                    gamma, dt, dx, _dp(flx), _dp(wct))
    return flx, wct


def riemann_iso(solver, mhd, wl, wr, bx, iso_cs, dfloor=DEFAULT_FLOOR):
    L = lib()
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    wr = np.ascontiguousarray(wr, dtype=np.float64)
    n = wl.shape[1]
    bx = np.ascontiguousarray(bx if bx is not None else np.zeros(n), dtype=np.float64)
    flx = np.zeros_like(wl)
    L.ao_riemann_iso(SOLVER[solver], int(mhd), n, _dp(wl), _dp(wr), _dp(bx), iso_cs, dfloor,
                     _dp(flx))
    return flx


def plm(qm1, q, qp1, wp=0.5, wm=0.5):
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.float64)
    nv, n = q.shape
    ql, qr = np.zeros_like(q), np.zeros_like(q)
    L.ao_plm(n, nv, _dp(np.ascontiguousarray(qm1)), _dp(q), _dp(np.ascontiguousarray(qp1)),
             wp, wm, _dp(ql), _dp(qr))
    return ql, qr


def ppm(qm2, qm1, q, qp1, qp2, dfloor=DEFAULT_FLOOR, pfloor=DEFAULT_FLOOR):
    L = lib()
    q = np.ascontiguousarray(q, dtype=np.float64)
    nv, n = q.shape
    ql, qr = np.zeros_like(q), np.zeros_like(q)
    L.ao_ppm(n, nv, _dp(np.ascontiguousarray(qm2)), _dp(np.ascontiguousarray(qm1)), _dp(q),
             _dp(np.ascontiguousarray(qp1)), _dp(np.ascontiguousarray(qp2)), dfloor, pfloor,
             _dp(ql), _dp(qr))
    return ql, qr
