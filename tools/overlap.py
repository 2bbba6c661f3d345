import ctypes as C, os, sys, json
import numpy as np
sys.path.insert(0, "/root/repo")
import torch
import athena_gamma_b200 as ab
import bench
blk = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "512,512,512").split(","))
pin, pgen_name, mhd, flux, ng, blk = bench.make_pin(ab, "c5", 1, blk)
mesh = ab.Mesh(pin, mhd=mhd, flux=flux, nghost=ng, rank=0, nranks=1, device=0)
for pmb in mesh.my_blocks:
    st = ab.pgen.BY_NAME[pgen_name](pmb, pin)
    for k, v in st.items():
        pmb.set(k, v)
mesh.initialize()
mesh.cycles(3, async_=True); mesh.sync()
L = mesh.L
L.ab_debug_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
out = (C.c_double*8)()
for grid in (148, 296, 444, 592, 1184):
    for nmem in (3,):
        rc = L.ab_debug_overlap(mesh.h, grid, nmem, out)
        print("grid_mem=%d nmem=%d rc=%d flux_alone=%.2f mem_full=%.2f mem_capped=%.2f | concurrent: flux_end=%.2f mem_end=%.2f" % (grid, nmem, rc, out[0], out[1], out[2], out[3], out[4]), flush=True)
