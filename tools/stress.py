import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import util, gpu_util
names = sys.argv[1].split(",")
reps = int(sys.argv[2])
fails = 0
for r in range(reps):
    for name in names:
        g = util.Golden(name)
        try:
            m = gpu_util.mesh_from_golden(g)
            m.initialize()
            dts = m.cycles(g.ncycles)
            ok = list(dts) == list(g.dts[:g.ncycles])
            for n, loc in enumerate(g.locs):
                pmb = m.block_of(*loc)
                for f in g.fields:
                    ok &= np.array_equal(pmb.get(f), g.final[n][f])
            if not ok:
                print("rep", r, name, "MISMATCH", flush=True); fails += 1
            del m
        except Exception as e:
            print("rep", r, name, "EXC", str(e)[:300], flush=True)
            sys.exit(1)
print("done reps", reps, "fails", fails)
