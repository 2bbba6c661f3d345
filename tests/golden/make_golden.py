#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run where /root/reference is mounted and oracle/_ref has been built
(`python oracle/build_ref.py`).  For each case the reference binary is run with a restart
dump every cycle; the fixture keeps the parameter blocks, the full dt sequence (the 17-digit
`dt=` values of the stdout cycle lines, bit-exact as doubles), the initial dump (cycle 0,
after Mesh::Initialize) and the final dump (u and face b of every MeshBlock, ghosts included).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_run  # noqa: E402

I = os.path.join(ROOT, "inputs")
LW = {"mesh/nx1": 16, "mesh/nx2": 8, "mesh/nx3": 8}
BL = {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16, "problem/radius": 0.3}
OT = {"mesh/nx1": 32, "mesh/nx2": 32}
KH = {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16}


def mb(a, b, c=1):
    return {"meshblock/nx1": a, "meshblock/nx2": b, "meshblock/nx3": c}


KS = {"mesh/nx1": 16, "mesh/nx2": 32, "mesh/nx3": 1}

# name -> (cfg, pgen, athinput, overrides, solver, mhd, ncycles[, nscalars[, eos]])
# fixtures named shkcloud* need the user boundary function tests/util.py:shock_cloud_inner_x1
CASES = {
    "c2_linwave_hlld_plm_vl2_1blk": ("mhd_hlld_ng2", "linear_wave", "athinput.linear_wave3d",
                                     dict(LW, **mb(16, 8, 8)), "hlld", True, 4),
    "c2_linwave_hlld_plm_vl2_8blk": ("mhd_hlld_ng2", "linear_wave", "athinput.linear_wave3d",
                                     dict(LW, **mb(8, 4, 4)), "hlld", True, 4),
    "c5_blast_hlld_plm_vl2_8blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                   dict(BL, **mb(8, 8, 8)), "hlld", True, 6),
    "c5_blast_hlld_plm_vl2_1blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                   dict(BL, **mb(16, 16, 16)), "hlld", True, 6),
    "c3_ot_hlld_ppm_vl2_4blk": ("mhd_hlld_ng3", "orszag_tang", "athinput.orszag_tang",
                                dict(OT, **mb(16, 16)), "hlld", True, 5),
    "c4_kh_hllc_ppm_rk2_8blk": ("hydro_hllc_ng3", "kh", "athinput.kh",
                                dict(KH, **mb(8, 8, 8)), "hllc", False, 5),
    "c1_sod_hllc_plm_vl2_2blk": ("hydro_hllc_ng2", "shock_tube", "athinput.sod",
                                 {"mesh/nx1": 64, "meshblock/nx1": 32}, "hllc", False, 8),
    "sod_hlle_plm_vl2": ("hydro_hlle_ng2", "shock_tube", "athinput.sod",
                         {"mesh/nx1": 64, "meshblock/nx1": 64}, "hlle", False, 8),
    "sod_roe_plm_vl2": ("hydro_roe_ng2", "shock_tube", "athinput.sod",
                        {"mesh/nx1": 64, "meshblock/nx1": 64}, "roe", False, 8),
    "linwave_mhd_hlle_plm_vl2": ("mhd_hlle_ng2", "linear_wave", "athinput.linear_wave3d",
                                 dict(LW, **mb(16, 8, 8), **{"problem/amp": 0.1}),
                                 "hlle", True, 4),
    # the fork's production solvers (confignotes)
    "blast_lhlld_plm_vl2_8blk": ("mhd_lhlld_ng2", "blast", "athinput.blast",
                                 dict(BL, **mb(8, 8, 8)), "lhlld", True, 6),
    "ot_lhlld_plm_vl2_4blk": ("mhd_lhlld_ng2", "orszag_tang", "athinput.orszag_tang",
                              dict(OT, **mb(16, 16), **{"time/xorder": 2}), "lhlld", True, 5),
    "blast_lhllc_plm_vl2_8blk": ("hydro_lhllc_ng2", "blast", "athinput.blast",
                                 dict(BL, **mb(8, 8, 8)), "lhllc", False, 6),
    "sod_lhllc_plm_vl2_2blk": ("hydro_lhllc_ng2", "shock_tube", "athinput.sod",
                               {"mesh/nx1": 64, "meshblock/nx1": 32}, "lhllc", False, 8),
    # physical boundaries: reflecting box, and a reflect / outflow / periodic mix
    "blast_refl_hlld_plm_vl2_8blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                     dict(BL, **mb(8, 8, 8), **{"problem/radius": 0.6},
                                          **{"mesh/%s%d_bc" % (s, d): "reflecting"
                                             for s in "io" for d in (1, 2, 3)}),
                                     "hlld", True, 8),
    "blast_mixedbc_hllc_plm_vl2_8blk": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                        dict(BL, **mb(8, 8, 8), **{
                                            "problem/radius": 0.6,
                                            "mesh/ix1_bc": "reflecting", "mesh/ox1_bc": "outflow",
                                            "mesh/ix3_bc": "outflow", "mesh/ox3_bc": "reflecting"}),
                                        "hllc", False, 8),
    # other integrators / orders of the same path
    "kh2d_hllc_plm_rk3_4blk": ("hydro_hllc_ng2", "kh", "athinput.kh",
                               {"mesh/nx1": 32, "mesh/nx2": 32, "mesh/nx3": 1, "time/xorder": 2,
                                "time/integrator": "rk3", **mb(16, 16, 1)}, "hllc", False, 4),
    "ot_hlld_dc_rk1_4blk": ("mhd_hlld_ng2", "orszag_tang", "athinput.orszag_tang",
                            dict(OT, **mb(16, 16), **{"time/xorder": 1, "time/integrator": "rk1",
                                                      "time/cfl_number": 0.3}),
                            "hlld", True, 5),
    "blast_hlld_ppm_rk3_8blk": ("mhd_hlld_ng3", "blast", "athinput.blast",
                                dict(BL, **mb(8, 8, 8), **{"time/xorder": 3,
                                                           "time/integrator": "rk3"}),
                                "hlld", True, 3),
    "linwave_mhd_roe_plm_vl2_2blk": ("mhd_roe_ng2", "linear_wave", "athinput.linear_wave3d",
                                     dict(LW, **mb(8, 8, 8), **{"problem/amp": 0.1}),
                                     "roe", True, 4),
    # user-enrolled boundary function (pgen/shk_cloud.cpp:ShockCloudInnerX1 on inner x1)
    "shkcloud2d_hllc_plm_vl2_4blk": ("hydro_hllc_ng2", "shk_cloud", "athinput.shk_cloud",
                                     {"mesh/nx1": 32, "mesh/nx2": 16, **mb(16, 8, 1)},
                                     "hllc", False, 8),
    "shkcloud3d_hlld_plm_vl2_8blk": ("mhd_hlld_ng2", "shk_cloud", "athinput.shk_cloud",
                                     {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16,
                                      **mb(8, 8, 8)}, "hlld", True, 5),
    # user-enrolled explicit source function (oracle/pgen/usersrc.cpp, our own test pgen built
    # against the reference; tests/util.py:central_gravity_source restates it)
    "usersrc_lhllc_plm_vl2_8blk_s1": ("hydro_lhllc_ng2_s1", "usersrc", "athinput.blast",
                                      dict(BL, **mb(8, 8, 8)), "lhllc", False, 5, 1),
    "usersrc_hlld_plm_rk3_8blk": ("mhd_hlld_ng2", "usersrc", "athinput.blast",
                                  dict(BL, **mb(8, 8, 8), **{"time/integrator": "rk3"}),
                                  "hlld", True, 4),
    "usersrc_iso_hlle_plm_rk2_8blk": ("hydro_hlle_iso_ng2", "usersrc", "athinput.blast",
                                      dict(BL, **mb(8, 8, 8), **{"time/integrator": "rk2",
                                           "hydro/iso_sound_speed": 0.8}),
                                      "hlle", False, 5, 0, "isothermal"),
    # dimensional / shape edge cases of the MHD path: 1-D (corner E and CT special cases, the
    # duplicated x2/x3 faces), 2-D with mixed physical boundaries, non-cubic 3-D blocks
    "bw1d_hlld_plm_vl2_2blk": ("mhd_hlld_ng2", "shock_tube", "athinput.bw",
                               {"mesh/nx1": 64, "meshblock/nx1": 32}, "hlld", True, 8),
    "bw2d_x2_hlld_plm_rk2_4blk": ("mhd_hlld_ng2", "shock_tube", "athinput.bw",
                                  {"mesh/nx1": 8, "mesh/nx2": 32, "meshblock/nx1": 4,
                                   "meshblock/nx2": 16, "problem/shock_dir": 2,
                                   "mesh/ix1_bc": "periodic", "mesh/ox1_bc": "periodic",
                                   "mesh/ix2_bc": "outflow", "mesh/ox2_bc": "reflecting",
                                   "time/integrator": "rk2"}, "hlld", True, 6),
    "blast_noncubic_hlld_ppm_rk2_6blk": ("mhd_hlld_ng3", "blast", "athinput.blast",
                                         {"mesh/nx1": 24, "mesh/nx2": 12, "mesh/nx3": 8,
                                          "problem/radius": 0.3, "time/xorder": 3,
                                          "time/integrator": "rk2", **mb(12, 6, 4)},
                                         "hlld", True, 4),
    # characteristic reconstruction (time/xorder = 2c, 3c; reconstruct/characteristic.cpp)
    "blast_hlld_plmc_vl2_8blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                 dict(BL, **mb(8, 8, 8), **{"time/xorder": "2c"}),
                                 "hlld", True, 6),
    "blast_hllc_plmc_vl2_8blk": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                 dict(BL, **mb(8, 8, 8), **{"time/xorder": "2c"}),
                                 "hllc", False, 6),
    "blast_hlld_ppmc_rk3_8blk": ("mhd_hlld_ng3", "blast", "athinput.blast",
                                 dict(BL, **mb(8, 8, 8), **{"time/xorder": "3c",
                                                            "time/integrator": "rk3"}),
                                 "hlld", True, 3),
    "ot_hlld_ppmc_vl2_4blk": ("mhd_hlld_ng3", "orszag_tang", "athinput.orszag_tang",
                              dict(OT, **mb(16, 16), **{"time/xorder": "3c"}), "hlld", True, 5),
    "kh_hllc_ppmc_rk2_8blk": ("hydro_hllc_ng3", "kh", "athinput.kh",
                              dict(KH, **mb(8, 8, 8), **{"time/xorder": "3c"}), "hllc", False, 5),
    # LLF
    "blast_llf_plm_vl2_8blk": ("hydro_llf_ng2", "blast", "athinput.blast", dict(BL, **mb(8, 8, 8)),
                               "llf", False, 5),
    "blast_mhd_llf_plm_vl2_8blk": ("mhd_llf_ng2", "blast", "athinput.blast",
                                   dict(BL, **mb(8, 8, 8)), "llf", True, 5),
    "iso_blast_mhd_llf_plm_vl2_8blk": ("mhd_llf_iso_ng2", "blast", "athinput.blast",
                                       dict(BL, **mb(8, 8, 8),
                                            **{"hydro/iso_sound_speed": 0.4082482905,
                                               "problem/drat": 5.0}),
                                       "llf", True, 5, 0, "isothermal"),
    # constant acceleration source term (hydro/srcterms/constant_acc.cpp) in a closed box
    "blast_grav_hllc_plm_vl2_8blk": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                     dict(BL, **mb(8, 8, 8), **{
                                         "hydro/grav_acc1": 0.3, "hydro/grav_acc3": -1.0,
                                         "mesh/ix3_bc": "reflecting", "mesh/ox3_bc": "reflecting"}),
                                     "hllc", False, 6),
    "blast_grav_hlld_plm_rk2_8blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                     dict(BL, **mb(8, 8, 8), **{
                                         "hydro/grav_acc2": -0.7, "time/integrator": "rk2",
                                         "mesh/ix2_bc": "reflecting", "mesh/ox2_bc": "reflecting"}),
                                     "hlld", True, 5),
    "iso_blast_grav_hlle_plm_vl2_8blk": ("hydro_hlle_iso_ng2", "blast", "athinput.blast",
                                         dict(BL, **mb(8, 8, 8),
                                              **{"hydro/iso_sound_speed": 0.4082482905,
                                                 "problem/drat": 5.0, "hydro/grav_acc1": -0.5}),
                                         "hlle", False, 5, 0, "isothermal"),
    # isothermal EOS (eos/isothermal_{hydro,mhd}.cpp; hlle.cpp / hlle_mhd.cpp / hlld_iso.cpp)
    "iso_khs_hlle_plm_vl2_4blk_s1": ("hydro_hlle_iso_ng2_s1", "kh", "athinput.kh_scalar",
                                     dict(KS, **mb(8, 16, 1)), "hlle", False, 6, 1,
                                     "isothermal"),
    "iso_blast_hlle_plm_vl2_8blk": ("hydro_hlle_iso_ng2", "blast", "athinput.blast",
                                    dict(BL, **mb(8, 8, 8), **{"hydro/iso_sound_speed": 0.4082482905,
                                                               "problem/drat": 5.0}),
                                    "hlle", False, 6, 0, "isothermal"),
    "iso_blast_mhd_hlld_plm_vl2_8blk": ("mhd_hlld_iso_ng2", "blast", "athinput.blast",
                                        dict(BL, **mb(8, 8, 8),
                                             **{"hydro/iso_sound_speed": 0.4082482905,
                                                "problem/drat": 5.0}),
                                        "hlld", True, 6, 0, "isothermal"),
    "iso_ot_hlld_plm_rk2_4blk": ("mhd_hlld_iso_ng2", "orszag_tang", "athinput.orszag_tang",
                                 dict(OT, **mb(16, 16), **{"time/xorder": 2,
                                                           "time/integrator": "rk2",
                                                           "hydro/iso_sound_speed": 0.7}),
                                 "hlld", True, 5, 0, "isothermal"),
    "iso_ot_mhd_hlle_plm_vl2_4blk": ("mhd_hlle_iso_ng2", "orszag_tang", "athinput.orszag_tang",
                                     dict(OT, **mb(16, 16), **{"time/xorder": 2,
                                                               "hydro/iso_sound_speed": 0.7}),
                                     "hlle", True, 5, 0, "isothermal"),
    # Roe's solver with the isothermal EOS (hydro and MHD)
    "iso_blast_roe_plm_vl2_8blk": ("hydro_roe_iso_ng2", "blast", "athinput.blast",
                                   dict(BL, **mb(8, 8, 8), **{"hydro/iso_sound_speed": 0.4082482905,
                                                              "problem/drat": 5.0}),
                                   "roe", False, 6, 0, "isothermal"),
    "iso_kh2d_roe_plm_rk2_4blk": ("hydro_roe_iso_ng2", "kh", "athinput.kh",
                                  {"mesh/nx1": 32, "mesh/nx2": 32, "mesh/nx3": 1, "time/xorder": 2,
                                   "time/integrator": "rk2", "hydro/iso_sound_speed": 0.9,
                                   **mb(16, 16, 1)}, "roe", False, 5, 0, "isothermal"),
    "iso_blast_mhd_roe_plm_vl2_8blk": ("mhd_roe_iso_ng2", "blast", "athinput.blast",
                                       dict(BL, **mb(8, 8, 8),
                                            **{"hydro/iso_sound_speed": 0.4082482905,
                                               "problem/drat": 5.0}),
                                       "roe", True, 6, 0, "isothermal"),
    "iso_ot_mhd_roe_plm_rk2_4blk": ("mhd_roe_iso_ng2", "orszag_tang", "athinput.orszag_tang",
                                    dict(OT, **mb(16, 16), **{"time/xorder": 2,
                                                              "time/integrator": "rk2",
                                                              "hydro/iso_sound_speed": 0.7}),
                                    "roe", True, 5, 0, "isothermal"),
    # passive scalars (src/scalars): the fork's production build carries one (confignotes)
    "khs_lhllc_plm_vl2_4blk_s1": ("hydro_lhllc_ng2_s1", "kh", "athinput.kh_scalar",
                                  dict(KS, **mb(8, 16, 1)), "lhllc", False, 6, 1),
    "sods_lhllc_plm_vl2_2blk_s1": ("hydro_lhllc_ng2_s1", "shock_tube", "athinput.sod",
                                   {"mesh/nx1": 64, "meshblock/nx1": 32}, "lhllc", False, 8, 1),
    "khs3d_hllc_ppm_rk3_8blk_s2": ("hydro_hllc_ng3_s2", "kh", "athinput.kh_scalar",
                                   dict({"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16},
                                        **mb(8, 8, 8), **{"time/xorder": 3,
                                                          "time/integrator": "rk3"}),
                                   "hllc", False, 3, 2),
    "khs3d_mhd_hlld_plm_vl2_8blk_s1": ("mhd_hlld_ng2_s1", "kh", "athinput.kh_scalar",
                                       dict({"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16},
                                            **mb(8, 8, 8)), "hlld", True, 4, 1),
    # static mesh refinement, hydro (oracle only so far; the device path rejects it): tree and
    # Z-ordered block list, level-aware neighbours, restriction / prolongation, flux correction
    "smr_blast2d_hllc_plm_vl2": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                 {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 1,
                                  "problem/radius": 0.3, **mb(4, 4, 1), "mesh/refinement": "static",
                                  "refinement1/x1min": -0.1, "refinement1/x1max": 0.1,
                                  "refinement1/x2min": -0.1, "refinement1/x2max": 0.1,
                                  "refinement1/level": 1}, "hllc", False, 6),
    "smr_blast3d_hllc_plm_vl2": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                 {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16,
                                  "problem/radius": 0.3, **mb(4, 4, 4), "mesh/refinement": "static",
                                  "refinement1/x1min": -0.1, "refinement1/x1max": 0.1,
                                  "refinement1/x2min": -0.1, "refinement1/x2max": 0.1,
                                  "refinement1/x3min": -0.1, "refinement1/x3max": 0.1,
                                  "refinement1/level": 1}, "hllc", False, 4),
    "smr_khs2d_lhllc_plm_vl2_s1": ("hydro_lhllc_ng2_s1", "kh", "athinput.kh_scalar",
                                   dict(KS, **mb(4, 8, 1), **{"mesh/refinement": "static",
                                        "refinement1/x1min": -0.1, "refinement1/x1max": 0.1,
                                        "refinement1/x2min": -0.3, "refinement1/x2max": 0.0,
                                        "refinement1/level": 2}), "lhllc", False, 6, 1),
    "smr_khs3d_hllc_ppm_rk3_ng4_s2": ("hydro_hllc_ng4_s2", "kh", "athinput.kh_scalar",
                                      dict({"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16},
                                           **mb(8, 8, 8), **{"time/xorder": 3,
                                           "time/integrator": "rk3", "mesh/refinement": "static",
                                           "mesh/ix2_bc": "reflecting", "mesh/ox2_bc": "reflecting",
                                           "refinement1/x1min": -0.5, "refinement1/x1max": -0.3,
                                           "refinement1/x2min": -0.5, "refinement1/x2max": -0.3,
                                           "refinement1/x3min": 0.3, "refinement1/x3max": 0.5,
                                           "refinement1/level": 1}), "hllc", False, 3, 2),
    "smr_sod1d_hllc_plm_vl2": ("hydro_hllc_ng2", "shock_tube", "athinput.sod",
                               {"mesh/nx1": 64, "meshblock/nx1": 8, "mesh/refinement": "static",
                                "refinement1/x1min": -0.1, "refinement1/x1max": 0.15,
                                "refinement1/level": 2}, "hllc", False, 8),
    "smr_blast2d_lvl2_bcs_hllc_plm_rk2": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                          {"mesh/nx1": 16, "mesh/nx2": 24, "mesh/nx3": 1,
                                           "problem/radius": 0.3, **mb(4, 4, 1),
                                           "mesh/refinement": "static",
                                           "time/integrator": "rk2",
                                           "mesh/ix1_bc": "reflecting", "mesh/ox1_bc": "outflow",
                                           "mesh/ix2_bc": "outflow", "mesh/ox2_bc": "reflecting",
                                           "refinement1/x1min": -0.5, "refinement1/x1max": -0.2,
                                           "refinement1/x2min": -0.1, "refinement1/x2max": 0.1,
                                           "refinement1/level": 2,
                                           "refinement2/x1min": 0.3, "refinement2/x1max": 0.5,
                                           "refinement2/x2min": 0.5, "refinement2/x2max": 0.75,
                                           "refinement2/level": 1}, "hllc", False, 6),
    "smr_blast3d_refl_lhllc_plm_rk3": ("hydro_lhllc_ng2", "blast", "athinput.blast",
                                       {"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 8,
                                        "problem/radius": 0.3, **mb(4, 4, 4),
                                        "mesh/refinement": "static", "time/integrator": "rk3",
                                        **{"mesh/%s%d_bc" % (s, d): "reflecting"
                                           for s in "io" for d in (1, 2, 3)},
                                        "refinement1/x1min": -0.5, "refinement1/x1max": -0.3,
                                        "refinement1/x2min": 0.0, "refinement1/x2max": 0.2,
                                        "refinement1/x3min": -0.5, "refinement1/x3max": -0.3,
                                        "refinement1/level": 1}, "lhllc", False, 3),
    "smr_kh3d_hllc_ppm_rk2_ng4": ("hydro_hllc_ng4", "kh", "athinput.kh",
                                  dict({"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16},
                                       **mb(4, 4, 4), **{"mesh/refinement": "static",
                                       "refinement1/x1min": -0.1, "refinement1/x1max": 0.1,
                                       "refinement1/x2min": -0.1, "refinement1/x2max": 0.1,
                                       "refinement1/x3min": -0.1, "refinement1/x3max": 0.1,
                                       "refinement1/level": 1}), "hllc", False, 3),
    # nonuniform (geometric) mesh spacing, mesh/x?rat != 1: mesh generator, dx?v, nonuniform
    # PLM / PPM branches (plm.cpp:81-105, ppm.cpp:196-207,282-300, reconstruction.cpp:434-461)
    # and the weighted cell-centred field (field.cpp:139-172)
    "blast_nonuni_hlld_plm_vl2_8blk": ("mhd_hlld_ng2", "blast", "athinput.blast",
                                       dict(BL, **mb(8, 8, 8), **{
                                           "mesh/x1rat": 1.05, "mesh/x2rat": 0.96,
                                           "mesh/x3rat": 1.03}), "hlld", True, 6),
    "blast_nonuni_hlld_ppm_rk3_8blk": ("mhd_hlld_ng3", "blast", "athinput.blast",
                                       dict(BL, **mb(8, 8, 8), **{
                                           "mesh/x1rat": 0.95, "mesh/x2rat": 1.04,
                                           "mesh/x3rat": 1.06, "time/xorder": 3,
                                           "time/integrator": "rk3",
                                           "mesh/ix1_bc": "reflecting", "mesh/ox1_bc": "outflow",
                                           "mesh/ix3_bc": "outflow", "mesh/ox3_bc": "reflecting"}),
                                       "hlld", True, 3),
    "blast_nonuni_x2_hllc_plmc_vl2_8blk": ("hydro_hllc_ng2", "blast", "athinput.blast",
                                           dict(BL, **mb(8, 8, 8), **{
                                               "mesh/x2rat": 1.07, "time/xorder": "2c",
                                               "mesh/ix2_bc": "outflow", "mesh/ox2_bc": "outflow"}),
                                           "hllc", False, 6),
    "ot_nonuni_hlld_ppmc_vl2_4blk": ("mhd_hlld_ng3", "orszag_tang", "athinput.orszag_tang",
                                     dict(OT, **mb(16, 16), **{"time/xorder": "3c",
                                                               "mesh/x1rat": 1.02,
                                                               "mesh/x2rat": 0.97}),
                                     "hlld", True, 5),
    "sod_nonuni_hllc_plm_vl2_2blk": ("hydro_hllc_ng2", "shock_tube", "athinput.sod",
                                     {"mesh/nx1": 64, "meshblock/nx1": 32, "mesh/x1rat": 1.03},
                                     "hllc", False, 8),
    "khs_nonuni_lhllc_plm_vl2_4blk_s1": ("hydro_lhllc_ng2_s1", "kh", "athinput.kh_scalar",
                                         dict(KS, **mb(8, 16, 1), **{"mesh/x1rat": 1.04,
                                                                     "mesh/x2rat": 0.98}),
                                         "lhllc", False, 6, 1),
    "khs3d_nonuni_hllc_ppm_rk3_8blk_s2": ("hydro_hllc_ng3_s2", "kh", "athinput.kh_scalar",
                                          dict({"mesh/nx1": 16, "mesh/nx2": 16, "mesh/nx3": 16},
                                               **mb(8, 8, 8), **{"time/xorder": 3,
                                                                 "time/integrator": "rk3",
                                                                 "mesh/x1rat": 1.03,
                                                                 "mesh/x2rat": 1.05,
                                                                 "mesh/x3rat": 0.95}),
                                          "hllc", False, 3, 2),
    "iso_blast_nonuni_mhd_hlld_plm_vl2_8blk": ("mhd_hlld_iso_ng2", "blast", "athinput.blast",
                                               dict(BL, **mb(8, 8, 8),
                                                    **{"hydro/iso_sound_speed": 0.4082482905,
                                                       "problem/drat": 5.0, "mesh/x1rat": 1.05,
                                                       "mesh/x3rat": 0.94}),
                                               "hlld", True, 6, 0, "isothermal"),
}


def make(name):
    cfg, pgen, inp, ov, solver, mhd, ncyc = CASES[name][:7]
    nscalars = CASES[name][7] if len(CASES[name]) > 7 else 0
    eos = CASES[name][8] if len(CASES[name]) > 8 else "adiabatic"
    nhydro = 4 if eos == "isothermal" else 5
    ov = dict(ov)
    ov["time/nlim"] = ncyc
    res = ref_run.run_reference(cfg, pgen, os.path.join(I, inp), ov, rst_every_cycle=True,
                                hst_every_cycle=True)
    first = ref_run.read_rst(res["rst"][0], nhydro=nhydro, mhd=mhd, nscalars=nscalars)
    last = ref_run.read_rst(res["rst"][ncyc], nhydro=nhydro, mhd=mhd, nscalars=nscalars)
    out = {"meta": json.dumps({"cfg": cfg, "pgen": pgen, "solver": solver, "mhd": mhd,
                               "ncycles": ncyc, "nghost": first["nghost"],
                               "nscalars": nscalars, "eos": eos, "par": first["par"]}),
           "dts": np.array(res["dts"][:ncyc + 1]),
           # refined meshes keep the level too (lx alone is ambiguous across levels)
           "locs": np.array([b["loc"][:4] if name.startswith("smr_") else b["loc"][:3]
                             for b in first["blocks"]], dtype=np.int64),
           "final_time": np.array(last["time"]), "final_dt": np.array(last["dt"])}
    if res["hst"] is not None:
        # one row per cycle 0..ncyc: time, dt, then the history sums (outputs/history.cpp)
        out["hst"] = res["hst"][:ncyc + 1]
    for tag, r in (("init", first), ("final", last)):
        for n, b in enumerate(r["blocks"]):
            out["%s_u_%d" % (tag, n)] = b["u"]
            if nscalars:
                out["%s_s_%d" % (tag, n)] = b["s"]
            if mhd:
                for f in ("b1", "b2", "b3"):
                    out["%s_%s_%d" % (tag, f, n)] = b[f]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    ref_run.cleanup(res)
    print("wrote", name, "dts", len(res["dts"]))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        make(nm)
