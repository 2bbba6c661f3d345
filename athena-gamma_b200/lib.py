"""ctypes binding of include/athena_b200.h (the C ABI of libathena_b200.so)."""
import ctypes as C
import os

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

AB_OK, AB_ERR_ARG, AB_ERR_NO_DEVICE, AB_ERR_CUDA, AB_ERR_NCCL, AB_ERR_STATE = 0, -1, -2, -3, -4, -5
BC = {"periodic": 0, "outflow": 1, "reflecting": 2, "user": 3}
SOLVER = {"hlle": 0, "hllc": 1, "hlld": 2, "roe": 3, "lhllc": 4, "lhlld": 5, "llf": 6}
INTEGRATOR = {"vl2": 0, "rk2": 1, "rk1": 2, "rk3": 3}
REG = {"u": 0, "u1": 1, "w": 2, "bcc": 3, "b1": 4, "b2": 5, "b3": 6, "b1_1": 7, "b1_2": 8,
       "b1_3": 9, "flux1": 10, "flux2": 11, "flux3": 12, "e1": 13, "e2": 14, "e3": 15,
       "wght1": 16, "wght2": 17, "wght3": 18, "e3_x1f": 19, "e2_x1f": 20, "e1_x2f": 21,
       "e3_x2f": 22, "e2_x3f": 23, "e1_x3f": 24,
       "s": 25, "s1": 26, "r": 27, "sflux1": 28, "sflux2": 29, "sflux3": 30}
EOS = {"adiabatic": 0, "isothermal": 1}
COORD = {"x1f": 0, "x2f": 1, "x3f": 2, "x1v": 3, "x2v": 4, "x3v": 5, "dx1f": 6, "dx2f": 7,
         "dx3f": 8}

# every symbol include/athena_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "ab_last_error", "ab_device_count", "ab_mesh_create", "ab_mesh_destroy",
    "ab_mesh_nblocks_total", "ab_mesh_nblocks_local", "ab_block_info", "ab_reg_size",
    "ab_plan_create", "ab_plan_messages", "ab_plan_ranklist", "ab_plan_geometry",
    "ab_enroll_user_boundary_function", "ab_enroll_user_explicit_source_function",
    "ab_enroll_user_explicit_source_function_device", "ab_upload", "ab_download", "ab_download_coord", "ab_comm_unique_id", "ab_comm_init",
    "ab_cons2prim", "ab_prim2cons", "ab_primitives", "ab_calc_fluxes", "ab_corner_e",
    "ab_weighted_ave", "ab_swap", "ab_zero", "ab_add_flux_div", "ab_add_source_terms", "ab_ct", "ab_physical_bcs",
    "ab_calc_scalar_fluxes", "ab_add_scalar_flux_div", "ab_scalar_cons2prim",
    "ab_scalar_prim2cons", "ab_new_block_dt", "ab_emf_exchange", "ab_bvals_exchange", "ab_mesh_initialize",
    "ab_bvals_send", "ab_bvals_recv_try", "ab_bvals_set", "ab_emf_send", "ab_emf_recv_try",
    "ab_clear_boundary", "ab_physical_bcs_at",
    "ab_mesh_cycles", "ab_mesh_set_async", "ab_mesh_state", "ab_mesh_set_time_dt",
    "ab_history", "ab_mesh_dt_history", "ab_mesh_profile", "ab_mesh_profile_read",
    "ab_mesh_launch_count", "ab_mesh_stream", "ab_mesh_sync",
    "ab_stage_begin", "ab_stage_upload_all", "ab_stage_commit", "ab_stage_download_all",
    "ab_stage_sync",
    "ab_smr_last_error", "ab_smr_plan_create", "ab_smr_plan_destroy", "ab_smr_plan_nblocks",
    "ab_smr_plan_blocks", "ab_smr_plan_neighbors", "ab_smr_plan_transfers",
    "ab_mesh_create_refined", "ab_block_level", "ab_plan_create_refined",
]


class AbRefinementRegion(C.Structure):
    _fields_ = [("x1min", C.c_double), ("x1max", C.c_double), ("x2min", C.c_double),
                ("x2max", C.c_double), ("x3min", C.c_double), ("x3max", C.c_double),
                ("level", C.c_int)]


class AbMeshParams(C.Structure):
    _fields_ = [("nx1", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int),
                ("bx1", C.c_int), ("bx2", C.c_int), ("bx3", C.c_int),
                ("x1min", C.c_double), ("x1max", C.c_double),
                ("x2min", C.c_double), ("x2max", C.c_double),
                ("x3min", C.c_double), ("x3max", C.c_double),
                ("bc", C.c_int * 6), ("nghost", C.c_int), ("mhd", C.c_int),
                ("solver", C.c_int), ("xorder", C.c_int), ("integrator", C.c_int),
                ("gamma", C.c_double), ("dfloor", C.c_double), ("pfloor", C.c_double),
                ("cfl_number", C.c_double), ("tlim", C.c_double), ("start_time", C.c_double),
                ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int),
                ("nscalars", C.c_int), ("eos", C.c_int), ("sfloor", C.c_double),
                ("iso_sound_speed", C.c_double), ("grav_acc", C.c_double * 3),
                ("char_proj", C.c_int), ("xrat", C.c_double * 3)]


# AbBValFunc (include/athena_b200.h): user-enrolled boundary function on host arrays
_DP = C.POINTER(C.c_double)
BVALFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_int, _DP, _DP, _DP, _DP, C.c_double, C.c_double,
                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)


SRCTERMFUNC = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, C.c_double, _DP, _DP, _DP, _DP,
                          _DP)
SRCTERMFUNC_DEVICE = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class AbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libathena_b200 error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def load():
    """Load the CUDA extension; raises (never falls back) when it is missing or unbuildable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    # AB_LIB: kernel-tuning builds and the emulated test library only.  Otherwise the library is
    # (re)built whenever it does not match the sources in the tree, so that tests and bench can
    # never run a stale binary.
    so = os.environ.get("AB_LIB") or _build.build()
    L = C.CDLL(so)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.c_int
    L.ab_last_error.restype = C.c_char_p
    L.ab_mesh_create.argtypes = [C.POINTER(AbMeshParams), C.POINTER(vp)]
    L.ab_mesh_destroy.argtypes = [vp]
    L.ab_mesh_nblocks_total.argtypes = [vp]
    L.ab_mesh_nblocks_local.argtypes = [vp]
    L.ab_block_info.argtypes = [vp, ip, C.POINTER(C.c_long)]
    L.ab_reg_size.restype = C.c_long
    L.ab_reg_size.argtypes = [vp, ip, ip]
    L.ab_plan_create.argtypes = [C.POINTER(AbMeshParams), C.POINTER(vp)]
    L.ab_plan_messages.argtypes = [vp, ip, C.POINTER(C.c_long), ip]
    L.ab_plan_ranklist.argtypes = [vp, C.POINTER(C.c_int), ip]
    L.ab_plan_geometry.argtypes = [vp, ip, ip, ip, dp, ip]
    L.ab_stage_begin.argtypes = [vp, C.POINTER(C.c_int), ip]
    L.ab_stage_upload_all.argtypes = [vp, C.POINTER(dp)]
    L.ab_stage_commit.argtypes = [vp]
    L.ab_stage_download_all.argtypes = [vp, C.POINTER(dp)]
    L.ab_stage_sync.argtypes = [vp]
    L.ab_smr_last_error.restype = C.c_char_p
    L.ab_smr_plan_create.argtypes = [C.POINTER(AbMeshParams), C.POINTER(AbRefinementRegion), ip,
                                     C.POINTER(vp)]
    L.ab_smr_plan_destroy.argtypes = [vp]
    L.ab_smr_plan_nblocks.argtypes = [vp]
    L.ab_smr_plan_blocks.argtypes = [vp, C.POINTER(C.c_long), ip]
    L.ab_smr_plan_neighbors.argtypes = [vp, ip, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ab_smr_plan_transfers.restype = C.c_long
    L.ab_smr_plan_transfers.argtypes = [vp, C.POINTER(C.c_long), C.c_long]
    L.ab_mesh_create_refined.argtypes = [C.POINTER(AbMeshParams), C.POINTER(AbRefinementRegion),
                                         ip, C.POINTER(vp)]
    L.ab_block_level.argtypes = [vp, ip]
    L.ab_plan_create_refined.argtypes = [C.POINTER(AbMeshParams), C.POINTER(AbRefinementRegion),
                                         ip, C.POINTER(vp)]
    L.ab_history.argtypes = [vp, dp, ip]
    L.ab_enroll_user_explicit_source_function.argtypes = [vp, SRCTERMFUNC, vp]
    L.ab_enroll_user_explicit_source_function_device.argtypes = [vp, SRCTERMFUNC_DEVICE, vp]
    L.ab_enroll_user_boundary_function.argtypes = [vp, ip, BVALFUNC, vp]
    L.ab_upload.argtypes = [vp, ip, ip, dp]
    L.ab_download.argtypes = [vp, ip, ip, dp]
    L.ab_download_coord.argtypes = [vp, ip, ip, dp]
    L.ab_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.ab_comm_init.argtypes = [vp, C.POINTER(C.c_ubyte)]
    L.ab_cons2prim.argtypes = [vp, ip] + [ip] * 6
    L.ab_prim2cons.argtypes = [vp, ip] + [ip] * 6
    for f in ("ab_primitives", "ab_corner_e", "ab_physical_bcs"):
        getattr(L, f).argtypes = [vp, ip]
    L.ab_calc_fluxes.argtypes = [vp, ip, ip, C.c_double]
    L.ab_weighted_ave.argtypes = [vp, ip, ip, ip, dp]
    L.ab_swap.argtypes = [vp, ip, ip]
    L.ab_zero.argtypes = [vp, ip, ip]
    L.ab_add_flux_div.argtypes = [vp, ip, C.c_double]
    L.ab_add_source_terms.argtypes = [vp, ip, C.c_double, C.c_double]
    L.ab_calc_scalar_fluxes.argtypes = [vp, ip, ip]
    L.ab_add_scalar_flux_div.argtypes = [vp, ip, C.c_double]
    L.ab_scalar_cons2prim.argtypes = [vp, ip] + [ip] * 6
    L.ab_scalar_prim2cons.argtypes = [vp, ip] + [ip] * 6
    L.ab_ct.argtypes = [vp, ip, C.c_double]
    L.ab_new_block_dt.argtypes = [vp, ip, dp]
    for f in ("ab_emf_exchange", "ab_bvals_exchange", "ab_mesh_initialize", "ab_mesh_sync"):
        getattr(L, f).argtypes = [vp]
    for f in ("ab_bvals_send", "ab_bvals_recv_try", "ab_bvals_set"):
        getattr(L, f).argtypes = [vp, ip, ip]
    L.ab_physical_bcs_at.argtypes = [vp, ip, C.c_double, C.c_double]
    for f in ("ab_emf_send", "ab_emf_recv_try", "ab_clear_boundary"):
        getattr(L, f).argtypes = [vp, ip]
    L.ab_mesh_cycles.argtypes = [vp, ip]
    L.ab_mesh_set_async.argtypes = [vp, ip]
    L.ab_mesh_state.argtypes = [vp, dp, dp, C.POINTER(C.c_long)]
    L.ab_mesh_set_time_dt.argtypes = [vp, C.c_double, C.c_double]
    L.ab_mesh_dt_history.argtypes = [vp, dp, ip]
    L.ab_mesh_profile.argtypes = [vp, ip]
    L.ab_mesh_profile_read.argtypes = [vp, dp]
    L.ab_mesh_launch_count.restype = C.c_long
    L.ab_mesh_launch_count.argtypes = [vp]
    L.ab_mesh_stream.restype = vp
    L.ab_mesh_stream.argtypes = [vp]
    _LIB = L
    return L


def check(rc):
    if rc < 0:
        raise AbError(rc, load().ab_last_error().decode())
    return rc
