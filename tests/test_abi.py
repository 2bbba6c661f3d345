"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/athena_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import athena_gamma_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "athena_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ab_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = ab.lib.load()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert sorted(ab.lib.SYMBOLS) == syms, "lib.py binding list out of date with the header"


def test_struct_layout_matches_header():
    # 6 ints, 6 doubles, 6 ints (bc), 5 ints (+pad), 6 doubles, 3 + 2 ints (+pad), 2 doubles,
    # grav_acc[3], char_proj (+pad), xrat[3]; checked against gcc's sizeof / offsetof of
    # include/athena_b200.h
    P = ab.lib.AbMeshParams
    assert C.sizeof(P) == 6 * 4 + 6 * 8 + 6 * 4 + 5 * 4 + 4 + 6 * 8 + 5 * 4 + 4 + 5 * 8 + 4 + 4 \
        + 3 * 8 == 264
    assert P.nscalars.offset == 180 and P.sfloor.offset == 192
    assert P.char_proj.offset == 232 and P.xrat.offset == 240


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pin = ab.ParameterInput(path=os.path.join(ROOT, "inputs", "athinput.blast"))
    pin.modify_from_cmdline(["mesh/nx1=16", "mesh/nx2=16", "mesh/nx3=16", "meshblock/nx1=16",
                             "meshblock/nx2=16", "meshblock/nx3=16"])
    with pytest.raises(ab.lib.AbError) as e:
        ab.Mesh(pin, mhd=True, flux="hlld")
    assert e.value.code == ab.lib.AB_ERR_NO_DEVICE


def test_bad_arguments_are_rejected():
    L = ab.lib.load()
    p = ab.lib.AbMeshParams()
    p.nx1, p.nx2, p.nx3, p.bx1, p.bx2, p.bx3 = 10, 1, 1, 3, 1, 1
    h = C.c_void_p()
    assert L.ab_mesh_create(C.byref(p), C.byref(h)) == ab.lib.AB_ERR_ARG
    assert b"divisible" in L.ab_last_error()


def test_parameter_input_surface():
    pin = ab.ParameterInput(path=os.path.join(ROOT, "inputs", "athinput.linear_wave3d"))
    assert pin.get_integer("mesh", "nx1") == 128
    assert pin.get_real("time", "cfl_number") == 0.3
    assert pin.get_boolean("problem", "compute_error") is True
    with pytest.raises(KeyError):
        pin.modify_from_cmdline(["nosuch/key=1"])        # parameter_input.cpp:351
    with pytest.raises(KeyError):
        pin.modify_from_cmdline(["mesh/nosuchkey=1"])
    pin.modify_from_cmdline(["time/nlim=7"])
    assert pin.get_integer("time", "nlim") == 7
    assert pin.get_or_add_real("hydro", "dfloor", 1.5) == 1.5
    assert pin.does_parameter_exist("hydro", "dfloor")
