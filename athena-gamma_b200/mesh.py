"""Mesh: host mirror of the reference's Mesh / MeshBlock / TimeIntegratorTaskList surface for
the accelerated path.  It owns no physics: every operation is a call through the C ABI
(include/athena_b200.h) into the CUDA library; there is no CPU fallback.

Reference surface mirrored: Mesh(pin) construction from <mesh>/<meshblock>/<time>/<hydro>
(src/mesh/mesh.cpp:63-548), ProblemGenerator hook (src/pgen/default_pgen.cpp), Mesh::Initialize
(mesh.cpp:1367-1651), the main loop (src/main.cpp:430-515) and the zone-cycle accounting
(main.cpp:480,585-595).
"""
import ctypes as C

import numpy as np

from . import lib
from .athinput import ParameterInput

# EquationOfState ctor default floors: sqrt(1024*float_min) (src/eos/adiabatic_mhd.cpp:29-31)
DEFAULT_FLOOR = float(np.sqrt(1024 * float(np.finfo(np.float32).tiny)))


class MeshBlock:
    """View of one local MeshBlock (index ranges as in src/mesh/meshblock.cpp:55-80)."""

    def __init__(self, mesh, lid):
        self.pmy_mesh, self.lid = mesh, lid
        info = (C.c_long * 13)()
        lib.check(mesh.L.ab_block_info(mesh.h, lid, info))
        (self.gid, self.lx1, self.lx2, self.lx3, self.ncells1, self.ncells2, self.ncells3,
         self.is_, self.ie, self.js, self.je, self.ks, self.ke) = [int(v) for v in info]
        self.level = int(mesh.L.ab_block_level(mesh.h, lid)) if hasattr(mesh.L, "ab_block_level") \
            else 0

    def shape(self, name):
        n1, n2, n3 = self.ncells1, self.ncells2, self.ncells3
        # NHYDRO: 5 adiabatic, 4 isothermal (configure.py:374-377)
        nh = 4 if getattr(getattr(self.pmy_mesh, "params", None), "eos", 0) == 1 else 5
        if name in ("u", "u1", "w"):
            return (nh, n3, n2, n1)
        if name in ("s", "s1", "r"):
            return (self.pmy_mesh.params.nscalars, n3, n2, n1)
        if name in ("sflux1", "sflux2", "sflux3"):
            ns = self.pmy_mesh.params.nscalars
            d = int(name[-1]) - 1
            sh = [n3, n2, n1]
            sh[2 - d] += 1
            return (ns,) + tuple(sh)
        if name == "bcc":
            return (3, n3, n2, n1)
        if name in ("b1", "b1_1", "wght1", "e2_x1f", "e3_x1f"):
            return (n3, n2, n1 + 1)
        if name in ("b2", "b1_2", "wght2", "e1_x2f", "e3_x2f"):
            return (n3, n2 + 1, n1)
        if name in ("b3", "b1_3", "wght3", "e1_x3f", "e2_x3f"):
            return (n3 + 1, n2, n1)
        if name in ("flux1", "flux2", "flux3"):
            d = int(name[-1]) - 1
            s = [n3, n2, n1]
            s[2 - d] += 1
            return (nh,) + tuple(s)
        if name == "e1":
            return (n3 + 1, n2 + 1, n1)
        if name == "e2":
            return (n3 + 1, n2, n1 + 1)
        if name == "e3":
            return (n3, n2 + 1, n1 + 1)
        raise KeyError(name)

    def get(self, name):
        """Download a register (AthenaArray layout) into a new numpy array."""
        m = self.pmy_mesh
        a = np.empty(self.shape(name), dtype=np.float64)
        assert a.size == m.L.ab_reg_size(m.h, self.lid, lib.REG[name]), name
        lib.check(m.L.ab_download(m.h, self.lid, lib.REG[name],
                                  a.ctypes.data_as(C.POINTER(C.c_double))))
        return a

    def set(self, name, a):
        m = self.pmy_mesh
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.shape(name), (name, a.shape, self.shape(name))
        lib.check(m.L.ab_upload(m.h, self.lid, lib.REG[name],
                                a.ctypes.data_as(C.POINTER(C.c_double))))

    def coord(self, name):
        m = self.pmy_mesh
        n = {"x1f": self.ncells1 + 1, "x2f": self.ncells2 + 1, "x3f": self.ncells3 + 1,
             "x1v": self.ncells1, "x2v": self.ncells2, "x3v": self.ncells3,
             "dx1f": self.ncells1, "dx2f": self.ncells2, "dx3f": self.ncells3}[name]
        a = np.empty(n)
        lib.check(m.L.ab_download_coord(m.h, self.lid, lib.COORD[name],
                                        a.ctypes.data_as(C.POINTER(C.c_double))))
        return a


class Mesh:
    @staticmethod
    def make_params(pin, mhd, flux, nghost=None, rank=0, nranks=1, device=0, nscalars=0,
                    eos="adiabatic"):
        """AbMeshParams from the athinput blocks + the configure-time choices."""
        p = lib.AbMeshParams()
        p.nx1 = pin.get_integer("mesh", "nx1")
        p.nx2 = pin.get_or_add_integer("mesh", "nx2", 1)
        p.nx3 = pin.get_or_add_integer("mesh", "nx3", 1)
        has_mb = "meshblock" in pin.blocks
        p.bx1 = pin.get_or_add_integer("meshblock", "nx1", p.nx1) if has_mb else p.nx1
        p.bx2 = pin.get_or_add_integer("meshblock", "nx2", p.nx2) if has_mb else p.nx2
        p.bx3 = pin.get_or_add_integer("meshblock", "nx3", p.nx3) if has_mb else p.nx3
        p.x1min, p.x1max = pin.get_real("mesh", "x1min"), pin.get_real("mesh", "x1max")
        p.x2min = pin.get_or_add_real("mesh", "x2min", -0.5)
        p.x2max = pin.get_or_add_real("mesh", "x2max", 0.5)
        p.x3min = pin.get_or_add_real("mesh", "x3min", -0.5)
        p.x3max = pin.get_or_add_real("mesh", "x3max", 0.5)
        for i, k in enumerate(("ix1_bc", "ox1_bc", "ix2_bc", "ox2_bc", "ix3_bc", "ox3_bc")):
            flag = pin.get_or_add_string("mesh", k, "periodic")
            if flag not in lib.BC:
                raise ValueError("### FATAL ERROR in Mesh: boundary flag '%s' not supported on "
                                 "the B200 path" % flag)
            p.bc[i] = lib.BC[flag]
        # Mesh ctor (mesh.cpp:69-75): geometric cell-size ratios; != 1 selects the mesh-generator
        # coordinates and the nonuniform reconstruction branches
        for d in range(3):
            p.xrat[d] = pin.get_or_add_real("mesh", "x%drat" % (d + 1), 1.0)
        xo = pin.get_or_add_string("time", "xorder", "2")
        p.char_proj = int(xo.endswith("c"))        # reconstruction.cpp:60-80: "2c", "3c"
        p.xorder = int(xo.rstrip("c"))
        p.nghost = nghost if nghost is not None else (3 if p.xorder == 3 else 2)
        p.mhd = int(bool(mhd))
        if flux == "default":
            flux = "hlld" if mhd else "hllc"    # configure.py:299-308
        p.solver = lib.SOLVER[flux]
        p.integrator = lib.INTEGRATOR[pin.get_or_add_string("time", "integrator", "vl2")]
        # the isothermal EquationOfState reads hydro/iso_sound_speed instead of hydro/gamma
        p.gamma = pin.get_real("hydro", "gamma") if eos == "adiabatic" else 0.0
        p.dfloor = pin.get_or_add_real("hydro", "dfloor", DEFAULT_FLOOR)
        p.pfloor = pin.get_or_add_real("hydro", "pfloor", DEFAULT_FLOOR)
        p.cfl_number = pin.get_real("time", "cfl_number")
        p.tlim = pin.get_real("time", "tlim")
        p.start_time = pin.get_or_add_real("time", "start_time", 0.0)
        p.rank, p.nranks, p.device = rank, nranks, device
        # configure.py --nscalars / --eos ; hydro/sfloor (eos ctor default like dfloor)
        p.nscalars = int(nscalars)
        p.eos = lib.EOS[eos]
        p.sfloor = pin.get_or_add_real("hydro", "sfloor", DEFAULT_FLOOR)
        p.iso_sound_speed = (pin.get_real("hydro", "iso_sound_speed") if eos == "isothermal"
                             else pin.get_or_add_real("hydro", "iso_sound_speed", 0.0))
        # HydroSourceTerms ctor (hydro/srcterms/hydro_srcterms.cpp:68-75)
        for d in range(3):
            p.grav_acc[d] = pin.get_or_add_real("hydro", "grav_acc%d" % (d + 1), 0.0)
        return p, flux

    def __init__(self, pin, mhd, flux, nghost=None, rank=0, nranks=1, device=0, nscalars=0,
                 eos="adiabatic"):
        """pin: ParameterInput (or athinput text); mhd / flux / nghost / nscalars / eos: what
        configure.py's -b / --flux / --nghost / --nscalars / --eos fix at compile time in the
        reference."""
        if not isinstance(pin, ParameterInput):
            pin = ParameterInput(text=pin)
        self.pin = pin
        self.L = lib.load()
        p, flux = self.make_params(pin, mhd, flux, nghost, rank, nranks, device, nscalars, eos)
        self.params = p
        self.mhd, self.flux = bool(mhd), flux
        self.nlim = pin.get_or_add_integer("time", "nlim", -1)
        h = C.c_void_p()
        # mesh/refinement = static: <refinementN> blocks in input order (mesh.cpp:323-465)
        self.refinement = self.refinement_regions(pin, p)
        if self.refinement is not None:
            lib.check(self.L.ab_mesh_create_refined(C.byref(p), self.refinement,
                                                    len(self.refinement), C.byref(h)))
        else:
            lib.check(self.L.ab_mesh_create(C.byref(p), C.byref(h)))
        self.h = h
        self.nbtotal = self.L.ab_mesh_nblocks_total(h)
        self.nblocal = self.L.ab_mesh_nblocks_local(h)
        self.my_blocks = [MeshBlock(self, l) for l in range(self.nblocal)]
        self.time, self.dt, self.ncycle = p.start_time, float("inf"), 0
        self.zones_per_block = p.bx1 * p.bx2 * p.bx3

    @staticmethod
    def refinement_regions(pin, p):
        """AbRefinementRegion array of the <refinementN> blocks, or None without
        mesh/refinement = static (adaptive refinement is not supported)."""
        mode = pin.get_or_add_string("mesh", "refinement", "none")
        if mode == "none":
            return None
        if mode != "static":
            raise ValueError("### FATAL ERROR in Mesh: mesh/refinement = %s is not supported on "
                             "the B200 path (static only)" % mode)
        names = [b for b in pin.blocks if b.startswith("refinement")]
        regs = (lib.AbRefinementRegion * len(names))()
        lim = {"x1min": p.x1min, "x1max": p.x1max, "x2min": p.x2min, "x2max": p.x2max,
               "x3min": p.x3min, "x3max": p.x3max}
        for r, b in zip(regs, names):
            for k in lim:
                setattr(r, k, pin.get_real(b, k) if pin.does_parameter_exist(b, k) else lim[k])
            r.level = pin.get_integer(b, "level")
        return regs

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.ab_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- multi-process plumbing -------------------------------------------------------------
    def init_comm(self, broadcast_bytes):
        """broadcast_bytes(bytes|None) -> bytes: rank 0 passes the NCCL id, all ranks get it
        (torch.distributed / MPI_Bcast in the host application)."""
        if self.params.nranks <= 1:
            return
        idb = (C.c_ubyte * 128)()
        if self.params.rank == 0:
            lib.check(self.L.ab_comm_unique_id(idb))
            data = broadcast_bytes(bytes(idb))
        else:
            data = broadcast_bytes(None)
        idb = (C.c_ubyte * 128).from_buffer_copy(data)
        lib.check(self.L.ab_comm_init(self.h, idb))

    # ---- problem generator hook + initialisation -----------------------------------------------
    def block_of(self, lx1, lx2, lx3, level=None):
        """level: needed (and checked) on a refined mesh only"""
        for b in self.my_blocks:
            if (b.lx1, b.lx2, b.lx3) == (lx1, lx2, lx3) and (
                    level is None or self.refinement is None or b.level == level):
                return b
        return None

    def enroll_user_boundary_function(self, face, fn):
        """Mesh::EnrollUserBoundaryFunction (src/mesh/mesh.cpp): fn has the BValFunc signature
        (src/athena.hpp:179-182) fn(pmb, pco, prim, b, time, dt, il, iu, jl, ju, kl, ku, ngh);
        prim and b.x1f/x2f/x3f are writable numpy views of the staged host arrays, pmb and pco
        are the MeshBlock (index ranges, coord())."""
        mesh = self

        class _FaceField:
            pass

        def tramp(_user, lid, prim, b1, b2, b3, time, dt, il, iu, jl, ju, kl, ku, ngh):
            pmb = mesh.my_blocks[lid]
            w = np.ctypeslib.as_array(prim, shape=pmb.shape("w"))
            bf = None
            if mesh.mhd:
                bf = _FaceField()
                bf.x1f = np.ctypeslib.as_array(b1, shape=pmb.shape("b1"))
                bf.x2f = np.ctypeslib.as_array(b2, shape=pmb.shape("b2"))
                bf.x3f = np.ctypeslib.as_array(b3, shape=pmb.shape("b3"))
            fn(pmb, pmb, w, bf, time, dt, il, iu, jl, ju, kl, ku, ngh)
        cb = lib.BVALFUNC(tramp)
        self._bval_keepalive = getattr(self, "_bval_keepalive", []) + [cb]
        lib.check(self.L.ab_enroll_user_boundary_function(self.h, face, cb, None))

    def enroll_user_explicit_source_function(self, fn, device=False):
        """Mesh::EnrollUserExplicitSourceFunction: fn(pmb, time, dt, prim, prim_scalar, bcc,
        cons, cons_scalar) (SrcTermFunc, src/athena.hpp:185-189).  device=False: numpy views of
        host staging arrays.  device=True: fn(pmb, time, dt, prim, prim_scalar, bcc, cons,
        cons_scalar, stream) gets objects exposing __cuda_array_interface__ over the library's
        device registers and the raw cudaStream_t; it must only enqueue work on that stream."""
        mesh = self

        class _Dev:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8",
                                                 "data": (int(ptr), False), "version": 2}

        def wrap(ptr, pmb, name, dev):
            if not ptr:
                return None
            shp = pmb.shape(name)
            if dev:
                return _Dev(ptr, shp)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=shp)

        if device:
            def tramp(_user, lid, time, dt, prim, prs, bcc, cons, cs, stream):
                pmb = mesh.my_blocks[lid]
                fn(pmb, time, dt, wrap(prim, pmb, "w", True), wrap(prs, pmb, "r", True),
                   wrap(bcc, pmb, "bcc", True), wrap(cons, pmb, "u", True),
                   wrap(cs, pmb, "s", True), stream)
            cb = lib.SRCTERMFUNC_DEVICE(tramp)
            call = self.L.ab_enroll_user_explicit_source_function_device
        else:
            def tramp(_user, lid, time, dt, prim, prs, bcc, cons, cs):
                pmb = mesh.my_blocks[lid]
                fn(pmb, time, dt, wrap(prim, pmb, "w", False), wrap(prs, pmb, "r", False),
                   wrap(bcc, pmb, "bcc", False), wrap(cons, pmb, "u", False),
                   wrap(cs, pmb, "s", False))
            cb = lib.SRCTERMFUNC(tramp)
            call = self.L.ab_enroll_user_explicit_source_function
        self._src_keepalive = getattr(self, "_src_keepalive", []) + [cb]
        lib.check(call(self.h, cb, None))

    def problem_generator(self, pgen_fn):
        """Calls pgen_fn(pmb, pin) -> dict(u=..., b1=..., b2=..., b3=...) per local MeshBlock
        (the MeshBlock::ProblemGenerator hook) and uploads the AthenaArrays."""
        for pmb in self.my_blocks:
            out = pgen_fn(pmb, self.pin)
            for k, v in out.items():
                pmb.set(k, v)

    def initialize(self):
        lib.check(self.L.ab_mesh_initialize(self.h))
        self._refresh()

    def _refresh(self):
        t, dt, n = C.c_double(), C.c_double(), C.c_long()
        lib.check(self.L.ab_mesh_state(self.h, C.byref(t), C.byref(dt), C.byref(n)))
        self.time, self.dt, self.ncycle = t.value, dt.value, n.value

    # ---- main loop -----------------------------------------------------------------------------
    def cycles(self, n, async_=False):
        """Advance n cycles (or until tlim).  Returns the dt used by each executed cycle."""
        lib.check(self.L.ab_mesh_set_async(self.h, int(async_)))
        lib.check(self.L.ab_mesh_cycles(self.h, n))
        out = np.empty(max(n, 1))
        got = lib.check(self.L.ab_mesh_dt_history(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), n))
        self._refresh()
        return out[:got]

    def run(self):
        """main.cpp:430 loop: while (time < tlim && (nlim < 0 || ncycle < nlim))."""
        dts = []
        while self.time < self.params.tlim and (self.nlim < 0 or self.ncycle < self.nlim):
            n = 64 if self.nlim < 0 else min(64, self.nlim - self.ncycle)
            dts.extend(self.cycles(n))
        return np.array(dts)

    def history(self):
        """HistoryOutput sums (outputs/history.cpp:69-169): mass, 1..3-mom, 1..3-KE, tot-E,
        [1..3-ME], [scalars], reduced on the device over all MeshBlocks of all ranks."""
        out = np.zeros(32)
        n = lib.check(self.L.ab_history(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), 32))
        return out[:n]

    def sync(self):
        lib.check(self.L.ab_mesh_sync(self.h))

    @property
    def launch_count(self):
        return self.L.ab_mesh_launch_count(self.h)

    @property
    def cuda_stream(self):
        return self.L.ab_mesh_stream(self.h)


class MeshPlan:
    """Host-only twin of Mesh (ab_plan_create): MeshBlock list, load balance and the
    cross-rank message plan of one rank.  Needs no GPU; owns no device memory."""

    def __init__(self, pin, mhd, flux, nghost=None, rank=0, nranks=1, nscalars=0):
        if not isinstance(pin, ParameterInput):
            pin = ParameterInput(text=pin)
        self.L = lib.load()
        p, _ = Mesh.make_params(pin, mhd, flux, nghost, rank, nranks, 0, nscalars)
        self.params = p
        h = C.c_void_p()
        lib.check(self.L.ab_plan_create(C.byref(p), C.byref(h)))
        self.h = h
        self.nbtotal = self.L.ab_mesh_nblocks_total(h)
        self.nblocal = self.L.ab_mesh_nblocks_local(h)
        self.my_blocks = [MeshBlock(self, l) for l in range(self.nblocal)]

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.ab_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ranklist(self):
        out = (C.c_int * self.nbtotal)()
        lib.check(self.L.ab_plan_ranklist(self.h, out, self.nbtotal))
        return list(out)

    def messages(self, kind):
        """list of dict(dir, peer, key, count, lid) in buffer order; kind 0 ghost, 1 EMF"""
        n = lib.check(self.L.ab_plan_messages(self.h, kind, None, 0))
        buf = (C.c_long * (5 * max(n, 1)))()
        lib.check(self.L.ab_plan_messages(self.h, kind, buf, n))
        return [dict(zip(("dir", "peer", "key", "count", "lid"), buf[5 * i:5 * i + 5]))
                for i in range(n)]
