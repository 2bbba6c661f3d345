#!/usr/bin/env python3
"""bench.py -- zone-cycles/s of the per-MeshBlock hydro/MHD update on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c2|c3|c4] [--impl ours|reference]

Default workload (N=1): BASELINE.json configs[4] / the north_star target: 3-D MHD blast wave,
512^3 zones per GPU, HLLD + PLM + VL2 + CT, periodic, one 512^3 MeshBlock per GPU; weak-scaled
(N x 512^3: 1024x512x512, 1024x1024x512, 1024^3) with MeshBlocks sharded over the ranks and the
ghost / EMF exchange over NCCL.  A "step" is one cycle (all integrator stages + new dt).
One JSON line is printed by rank 0.

  value      zone-cycles/s with the state resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the C ABI with HOST buffers: every step uploads u and the face
             fields from pinned host memory, runs Mesh::Initialize-style ghost fill + one cycle,
             and downloads u and b again; the copies of neighbouring steps overlap the kernels
             (ab_stage_*).  e2e_plain: the same sequence strictly serial
  e2e_resident  (extra) the drop-in's normal mode: state resident, one ab_mesh_cycles(1) +
             ab_history per step with their host synchronisation and read-backs
  roofline   the reconstruct+Riemann kernel (dominant): algorithmic bytes / CUDA-event duration;
             traffic / fp64_pipe_util / dram_fraction / local_mem_bytes / bound from the ncu
             summary of that kernel (profiles/flux_ncu.json, with the hash of the device sources it
             was taken on)
  cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref) on the host cores, a
             bounded sample of the same problem (c5: 256^3 in 64 MeshBlocks)
  same_config   (N=1) the GPU on exactly that <mesh>/<meshblock>
  other_workloads  (N=1) short device-resident runs of BASELINE configs C2, C3, C4 and of the c5
             kernels on a developed flow
  parity     (N>1) the reference goldens sharded over the ranks of this run, bit-exact, after
             the timed regions (tests/multirank_check.py)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_ZONE_CYCLE = {"mhd": 320.0, "hydro": 200.0}    # B_alg
FLUX_BYTES_PER_FACE = {"mhd": 128.0, "hydro": 80.0}      # (8 in + 8 out) / (5+5) doubles

WORKLOADS = {
    # name: (athinput, pgen, mhd, flux, nghost, per-GPU block (bx1,bx2,bx3), overrides)
    "c5": ("athinput.blast", "blast", True, "hlld", 2, (512, 512, 512), {}),
    "c2": ("athinput.linear_wave3d", "linear_wave", True, "hlld", 2, (128, 64, 64), {}),
    "c4": ("athinput.kh", "kh", False, "hllc", 3, (512, 512, 512), {}),
    "c3": ("athinput.orszag_tang", "orszag_tang", True, "hlld", 3, (2048, 2048, 1), {}),
}
REF_CFG = {"c5": ("mhd_hlld_ng2", "blast"), "c2": ("mhd_hlld_ng2", "linear_wave"),
           "c4": ("hydro_hllc_ng3", "kh"), "c3": ("mhd_hlld_ng3", "orszag_tang")}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def weak_mesh(nranks, block):
    """replicate the per-GPU block along x1, then x2, then x3 (Z-order gives one block/rank)"""
    reps = [1, 1, 1]
    d = 0
    n = nranks
    while n > 1:
        if block[d] > 1:
            reps[d] *= 2
            n //= 2
        d = (d + 1) % 3
    return reps


def make_pin(ab, wl, nranks, block=None, per_gpu=None):
    """block: MeshBlock size; per_gpu: zones per GPU (default = one MeshBlock per GPU)"""
    inp, pgen, mhd, flux, ng, blk, ov = WORKLOADS[wl]
    if block:
        blk = block
    gpu = per_gpu or blk
    pin = ab.ParameterInput(path=os.path.join(ROOT, "inputs", inp))
    reps = weak_mesh(nranks, gpu)
    for d in range(3):
        n = d + 1
        pin.set("mesh", "nx%d" % n, gpu[d]*reps[d])
        pin.set("meshblock", "nx%d" % n, blk[d])
        if reps[d] > 1:   # keep the zone size: stretch the domain
            lo, hi = pin.get_real("mesh", "x%dmin" % n), pin.get_real("mesh", "x%dmax" % n)
            pin.set("mesh", "x%dmax" % n, repr(lo + (hi - lo)*reps[d]))
    pin.set("time", "nlim", -1)
    pin.set("time", "tlim", 1.0e30)      # fixed cycle count; never clamp dt to tlim
    for k, v in ov.items():
        b, key = k.split("/")
        pin.set(b, key, v)
    return pin, pgen, mhd, flux, ng, blk


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_sample(wl, threads):
    """mesh / MeshBlock of the bounded CPU sample of workload `wl`: the same problem at the same
    integrator / solver / order, cut into at least `threads` MeshBlocks (the reference's OpenMP
    parallelises over MeshBlocks only, task_list.cpp:71-88).  The GPU arm runs the SAME mesh and
    MeshBlocks in its `same_config` leg."""
    blk = WORKLOADS[wl][5]
    ndim = sum(1 for b in blk if b > 1)
    if wl == "c2":
        mesh, bs = [128, 64, 64], [64, 32, 32]
    elif ndim == 3:
        mesh, bs = [256, 256, 256], [64, 64, 64]
    else:
        mesh, bs = [2048, 2048, 1], [256, 256, 1]
    nb = lambda: int(np.prod([m//b for m, b in zip(mesh, bs)]))   # noqa: E731
    while nb() < threads and min(b for b in bs if b > 1) > 16:
        bs = [b//2 if b > 1 else 1 for b in bs]
    return mesh, bs


def run_reference_cpu(wl, budget_s=25.0, steps=None, warmup=0):
    """Times the UNMODIFIED reference (oracle/_ref) on the host cores with OpenMP over
    MeshBlocks, on a bounded sample of the workload (reference_sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_run
    cfg, pgen = REF_CFG[wl]
    if not ref_run.have_ref(cfg, pgen):
        return None
    inp, _, mhd, flux, ng, blk, ov = WORKLOADS[wl]
    threads = os.cpu_count() or 1
    mesh, bs = reference_sample(wl, threads)
    nblocks = int(np.prod([m//b for m, b in zip(mesh, bs)]))
    threads = min(threads, nblocks)
    over = {"time/tlim": 1e30, "time/ncycle_out": 0}
    for d in range(3):
        over["mesh/nx%d" % (d+1)] = mesh[d]
        over["meshblock/nx%d" % (d+1)] = bs[d]
    over.update(ov)
    zones = int(np.prod(mesh))
    inp_path = os.path.join(ROOT, "inputs", inp)
    if steps is None:
        # calibrate with 2 cycles, then fill the budget
        over["time/nlim"] = 2
        r = ref_run.run_reference(cfg, pgen, inp_path, over, threads=threads)
        ref_run.cleanup(r)
        rate = r["zcps_omp"] or r["zcps"]
        ncyc = int(max(2, min(200, budget_s*rate/zones)))
    else:
        # a bounded sample: at most ~2 minutes of CPU work whatever --steps asks for
        over["time/nlim"] = 2
        r = ref_run.run_reference(cfg, pgen, inp_path, over, threads=threads)
        ref_run.cleanup(r)
        rate = r["zcps_omp"] or r["zcps"]
        ncyc = int(max(2, min(steps + warmup, 120.0*rate/zones)))
    over["time/nlim"] = ncyc
    r = ref_run.run_reference(cfg, pgen, inp_path, over, threads=threads)
    ref_run.cleanup(r)
    rate = r["zcps_omp"] or r["zcps"]
    return {"value": rate, "unit": "zone-cycles/s", "cores": threads, "kind": "reference",
            "sample": "%s %s mesh, %d MeshBlocks of %s, %d cycles, OpenMP %d threads, "
                      "g++ -O3 (reference default flags)" %
                      (pgen, "x".join(str(v) for v in mesh), nblocks,
                       "x".join(str(v) for v in bs), ncyc, threads),
            "ms_per_step": 1e3*zones/rate, "zones": zones, "cycles": ncyc,
            "mesh": mesh, "meshblock": bs}


def build_mesh(ab, wl, world, rank, device, dist, block=None, per_gpu=None, pinned=False):
    """Mesh of workload `wl` filled by the reference's problem-generator formula (evaluated on
    the host, as in the reference) and initialised.  pinned: keep the host arrays in pinned
    memory (they are the e2e legs' host-side state)."""
    import torch
    pin, pgen_name, mhd, flux, ng, blk = make_pin(ab, wl, world, block, per_gpu)
    mesh = ab.Mesh(pin, mhd=mhd, flux=flux, nghost=ng, rank=rank, nranks=world, device=device)
    if world > 1:
        def bcast(data):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.tensor(list(data), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        mesh.init_comm(bcast)
    names = ["u"] + (["b1", "b2", "b3"] if mhd else [])
    host = []
    for pmb in mesh.my_blocks:
        d = {nm: torch.zeros(pmb.shape(nm), dtype=torch.float64, pin_memory=pinned)
             for nm in names}
        st = ab.pgen.BY_NAME[pgen_name](pmb, pin, out={nm: t.numpy() for nm, t in d.items()})
        for k, v in st.items():
            pmb.set(k, v)
        host.append(d if pinned else None)
        if not pinned:
            del d, st
    mesh.initialize()
    return mesh, pin, names, host, blk


def timed_cycles(mesh, steps, warmup, device, dist, profile=False):
    """W untimed cycles, then exactly K cycles between CUDA events on the library's stream,
    barrier + synchronize on both sides, max over ranks.  Returns (ms, launches, prof)."""
    import torch

    def barrier():
        mesh.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        mesh.sync()
    mesh.cycles(warmup, async_=True)
    barrier()
    L = mesh.L
    if profile:
        L.ab_mesh_profile(mesh.h, 1)
    stream = torch.cuda.ExternalStream(mesh.cuda_stream, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = mesh.launch_count
    barrier()
    e0.record(stream)
    mesh.cycles(steps, async_=True)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = mesh.launch_count - launches0
    prof = (C.c_double*18)()
    if profile:
        L.ab_mesh_profile_read(mesh.h, prof)
        L.ab_mesh_profile(mesh.h, 0)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, launches, prof


def flux_profile(prof, nd, xo):
    return {"x%d_o%d" % (d+1, o+1): prof[d*3+o]/prof[9+d*3+o]
            for d in range(nd) for o in range(3) if prof[9+d*3+o] > 0}


def side_workload(ab, wl, device, steps, warmup, block=None, per_gpu=None):
    """a short device-resident run of another workload on one GPU (N=1 only): its throughput,
    ms per cycle and the per-direction averages of its reconstruct+Riemann kernels"""
    import gc
    mesh, pin, names, host, blk = build_mesh(ab, wl, 1, 0, device, None, block, per_gpu)
    ms, launches, prof = timed_cycles(mesh, steps, warmup, device, None, profile=True)
    zones = mesh.nbtotal*mesh.zones_per_block
    nd = 1 + (mesh.params.nx2 > 1) + (mesh.params.nx3 > 1)
    fp = flux_profile(prof, nd, mesh.params.xorder)
    out = {"value": zones*steps/(ms*1e-3), "unit": "zone-cycles/s", "ms_per_step": ms/steps,
           "steps": steps, "mesh": [mesh.params.nx1, mesh.params.nx2, mesh.params.nx3],
           "meshblock": list(blk), "meshblocks": mesh.nbtotal, "gpu_launches": int(launches),
           "flux_avg_ms_by_dir_order": fp,
           "flux_kernels_share_of_step": sum(prof[i] for i in range(9))/ms if ms > 0 else None}
    del mesh
    gc.collect()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--block", default=None, help="MeshBlock size, e.g. 256,256,256")
    ap.add_argument("--per-gpu", default=None,
                    help="zones per GPU, e.g. 512,512,512 with --block 128,128,128 = 64 "
                         "MeshBlocks per GPU (default: one MeshBlock per GPU)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-pipelined", action="store_true", help=argparse.SUPPRESS)  # default now
    ap.add_argument("--no-e2e-pipelined", action="store_true",
                    help="skip the e2e leg that overlaps the copies with the kernels (ab_stage_*)")
    ap.add_argument("--no-side", action="store_true",
                    help="skip the same_config and other_workloads legs (N=1 only)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-first-runs", action="store_true", help=argparse.SUPPRESS)  # accepted, no-op
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = a.workload
    mhd = WORKLOADS[wl][2]
    block = tuple(int(x) for x in a.block.split(",")) if a.block else None
    per_gpu = tuple(int(x) for x in a.per_gpu.split(",")) if a.per_gpu else None
    cfg_desc = {"workload": "%s: %s" % (wl, {
        "c5": "3D MHD blast wave, HLLD+PLM+VL2+CT, periodic (BASELINE configs[4])",
        "c2": "3D MHD linear wave 128x64x64, HLLD+PLM+VL2, 1 MeshBlock (BASELINE configs[1])",
        "c4": "3D Kelvin-Helmholtz hydro, HLLC+PPM+RK2 (BASELINE configs[3])",
        "c3": "2D Orszag-Tang MHD, HLLD+PPM+VL2 (BASELINE configs[2])"}[wl])}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_cpu(wl, steps=a.steps, warmup=a.warmup)
        if r is None:
            print(json.dumps({"impl": "reference",
                              "unavailable": "oracle/_ref binary missing (build needs /root/reference)"}))
            return 0
        cfg_desc["sample"] = r["sample"]
        cfg_desc["mesh"], cfg_desc["meshblock"] = r["mesh"], r["meshblock"]
        cfg_desc["note"] = ("bounded sample of the workload on the host cores; the GPU arm's "
                            "`same_config` leg runs this exact <mesh>/<meshblock>")
        out = {"impl": "reference", "metric": "zone-cycles/sec", "value": r["value"],
               "unit": "zone-cycles/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg_desc,
               "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
               "e2e": {"value": r["value"], "unit": "zone-cycles/s", "h2d_bytes_per_step": 0,
                       "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return 0

    import torch
    import athena_gamma_b200 as ab
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0
    torch.cuda.set_device(device)
    pin, pgen_name, mhd, flux, ng, blk = make_pin(ab, wl, world, block, per_gpu)
    mesh = ab.Mesh(pin, mhd=mhd, flux=flux, nghost=ng, rank=rank, nranks=world, device=device)
    if world > 1:
        def bcast(data):
            t = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                t.copy_(torch.tensor(list(data), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        mesh.init_comm(bcast)
    # synthetic input: the reference's own problem generator formula, evaluated on the host,
    # written straight into pinned host buffers (they are the e2e leg's host-side state)
    names = ["u"] + (["b1", "b2", "b3"] if mhd else [])
    need = sum(int(np.prod(mesh.my_blocks[0].shape(nm)))*8 for nm in names)*mesh.nblocal
    do_e2e = not a.no_e2e
    e2e_skip = None
    try:
        import psutil
        avail = psutil.virtual_memory().available
        # every rank of the node needs its buffers + ~0.5x for pgen temporaries
        if do_e2e and avail < 1.6*need*max(world, 1) + (8 << 30):
            do_e2e = False
            e2e_skip = "host memory: %.0f GB available < %.0f GB needed for pinned state" % (
                avail/2**30, 1.6*need*world/2**30)
    except Exception:
        pass
    if dist:   # all ranks must take the same decision
        flag = torch.tensor([1 if do_e2e else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if do_e2e and int(flag.item()) == 0:
            do_e2e, e2e_skip = False, "host memory short on another rank"
    pinned = []
    for pmb in mesh.my_blocks:
        d = {nm: torch.zeros(pmb.shape(nm), dtype=torch.float64, pin_memory=do_e2e)
             for nm in names}
        st = ab.pgen.BY_NAME[pgen_name](pmb, pin, out={nm: t.numpy() for nm, t in d.items()})
        for k, v in st.items():
            pmb.set(k, v)
        pinned.append(d if do_e2e else None)
        if not do_e2e:
            del d, st
    mesh.initialize()
    zones = mesh.nbtotal*mesh.zones_per_block
    zones_local = mesh.nblocal*mesh.zones_per_block

    def barrier():
        mesh.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        mesh.sync()

    # ---- device-resident throughput -----------------------------------------------------------
    mesh.cycles(a.warmup, async_=True)
    barrier()
    L = mesh.L
    L.ab_mesh_profile(mesh.h, 1)
    stream = torch.cuda.ExternalStream(mesh.cuda_stream, device=device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(device)
    sampler.start()
    launches0 = mesh.launch_count
    barrier()
    e0.record(stream)
    mesh.cycles(a.steps, async_=True)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = mesh.launch_count - launches0
    prof = (C.c_double*18)()
    L.ab_mesh_profile_read(mesh.h, prof)
    L.ab_mesh_profile(mesh.h, 0)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = zones*a.steps/(ms*1e-3)

    # ---- roofline of the dominant kernel (reconstruct + Riemann, full-order stage) ------------
    peak, peak_src = load_peaks()
    xo = mesh.params.xorder
    nd = 1 + (mesh.params.nx2 > 1) + (mesh.params.nx3 > 1)
    tot_ms = sum(prof[d*3 + (xo-1)] for d in range(nd))
    tot_n = sum(prof[9 + d*3 + (xo-1)] for d in range(nd))
    kname = "k_flux<dir,order=%d,%s,%s>" % (xo, flux, "mhd" if mhd else "hydro")
    faces = mesh.zones_per_block      # ~ one interface per zone per direction
    fb = FLUX_BYTES_PER_FACE["mhd" if mhd else "hydro"]
    roof = None
    if tot_n > 0:
        avg_ms = tot_ms/tot_n
        ach = fb*faces/(avg_ms*1e-3)/1e9
        share = sum(prof[i] for i in range(9))/ms if ms > 0 else None
        # ncu evidence of this kernel (one --set full capture per build, summarised by
        # tools/ncu_summary.py into profiles/flux_ncu.json with the source hash of the build it
        # was taken on): DRAM bytes per launch, FP64-pipe utilisation, DRAM fraction, local
        # memory traffic.  Nothing here is hard-coded; without the file the keys are null.
        ncu = None
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "flux_ncu.json")))
            key = "%s:%s" % (wl, "x".join(str(v) for v in blk))
            if key in tj.get("kernels", {}):
                ncu = dict(tj["kernels"][key])
                ncu["source"] = "profiles/flux_ncu.json"
                ncu["captured_on_kernel_hash"] = ncu.get("kernel_hash") or tj.get("kernel_hash")
                ncu["kernel_hash_now"] = ab.build.kernel_hash()   # device sources + flags
                traffic = ncu.get("dram_bytes_per_launch")
        except Exception:
            pass
        bound = "hbm"
        if ncu and ncu.get("fp64_pipe_util") is not None and ncu.get("dram_fraction") is not None:
            bound = "fp64" if ncu["fp64_pipe_util"] > ncu["dram_fraction"] else "hbm"
        roof = {"bound": bound, "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach/peak, "traffic": traffic, "peak_source": peak_src,
                "fp64_pipe_util": ncu.get("fp64_pipe_util") if ncu else None,
                "dram_fraction": ncu.get("dram_fraction") if ncu else None,
                "local_mem_bytes": ncu.get("local_mem_bytes_per_launch") if ncu else None,
                "ncu": ncu,
                "avg_launch_ms": avg_ms, "launches_timed": int(tot_n),
                "algorithmic_bytes_per_launch": fb*faces,
                "flux_kernels_share_of_step": share,
                "flux_avg_ms_by_dir_order": flux_profile(prof, nd, xo)}
    b_alg = BYTES_PER_ZONE_CYCLE["mhd" if mhd else "hydro"]
    # measured DRAM traffic of one whole cycle (ncu --set full over its 23 launches, summarised by
    # tools/ncu_cycle.py into profiles/cycle_ncu.json with the source hash of that build)
    cyc = None
    try:
        cj = json.load(open(os.path.join(ROOT, "profiles", "cycle_ncu.json")))
        if cj.get("workload") == "%s:%s" % (wl, "x".join(str(v) for v in blk)):
            cyc = cj
    except Exception:
        pass
    cycle_roof = {"B_alg_bytes_per_zone_cycle": b_alg,
                  "achieved_GBs_per_gpu": b_alg*value/world/1e9,
                  "frac_of_measured_peak": b_alg*value/world/1e9/peak,
                  "frac_of_8TBs": b_alg*value/world/8.0e12}
    if cyc:
        moved = float(cyc["dram_bytes_per_zone_cycle"])
        cycle_roof.update({
            "measured_dram_bytes_per_zone_cycle": moved,
            "measured_dram_GBs_per_gpu": moved*value/world/1e9,
            "measured_dram_frac_of_measured_peak": moved*value/world/1e9/peak,
            "source": "profiles/cycle_ncu.json",
            "captured_on_kernel_hash": cyc.get("kernel_hash"),
            "kernel_hash_now": ab.build.kernel_hash()})

    # ---- end-to-end through the C ABI with host buffers ---------------------------------------
    e2e = None if e2e_skip is None else {"value": None, "skipped": e2e_skip}
    if do_e2e:
        nbytes = sum(t.numel()*8 for d in pinned for t in d.values())
        dp = C.POINTER(C.c_double)

        def e2e_step():
            for pmb, d in zip(mesh.my_blocks, pinned):
                for nm in names:
                    ab.lib.check(L.ab_upload(mesh.h, pmb.lid, ab.lib.REG[nm],
                                             C.cast(d[nm].data_ptr(), dp)))
            ab.lib.check(L.ab_mesh_initialize(mesh.h))
            ab.lib.check(L.ab_mesh_cycles(mesh.h, 1))
            for pmb, d in zip(mesh.my_blocks, pinned):
                for nm in names:
                    ab.lib.check(L.ab_download(mesh.h, pmb.lid, ab.lib.REG[nm],
                                               C.cast(d[nm].data_ptr(), dp)))
        nsteps = max(1, min(a.steps, 3))
        L.ab_mesh_set_async(mesh.h, 0)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(nsteps):
            e2e_step()
        barrier()
        dt_wall = time.perf_counter() - t0
        if dist:
            t = torch.tensor([dt_wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_wall = float(t.item())
        e2e = {"value": zones*nsteps/dt_wall, "unit": "zone-cycles/s",
               "h2d_bytes_per_step": nbytes*world if world > 1 else nbytes,
               "d2h_bytes_per_step": nbytes*world if world > 1 else nbytes,
               "steps": nsteps,
               "what": "per step: upload u,b (pinned host) -> ghost fill + cons2prim + dt -> "
                       "1 cycle -> download u,b"}

    # ---- the same end-to-end step, pipelined (opt-in): the upload of step n+1 and the download
    # of step n-1 run on copy streams while step n computes (ab_stage_*, include/athena_b200.h).
    # Same bytes, same work per step; the result lands in its own pinned buffers because the
    # input buffers are being read by the next upload at that time.
    e2e_pipe = None
    pipe_ok = do_e2e and not a.no_e2e_pipelined
    if pipe_ok:
        # the leg pins a second copy of the state (its results land in their own buffers): skip it
        # when the ranks of this node together would take more than half of the free host memory
        try:
            import psutil
            need = sum(t.numel()*8 for d in pinned for t in d.values())*max(world, 1)
            if need > 0.5*psutil.virtual_memory().available:
                pipe_ok = False
                e2e_pipe = {"value": None, "error": "skipped: another %.0f GB of pinned host memory needed"
                                                    % (need/1e9)}
        except Exception:
            pass
    if pipe_ok:
        try:
            dp = C.POINTER(C.c_double)
            outbuf = [{nm: torch.zeros_like(d[nm]).pin_memory() for nm in names} for d in pinned]
            nreg = len(names)
            regs = (C.c_int*nreg)(*[ab.lib.REG[nm] for nm in names])
            inp = (dp*(nreg*len(pinned)))(*[C.cast(d[nm].data_ptr(), dp)
                                             for d in pinned for nm in names])
            outp = (dp*(nreg*len(pinned)))(*[C.cast(d[nm].data_ptr(), dp)
                                              for d in outbuf for nm in names])
            ab.lib.check(L.ab_stage_begin(mesh.h, regs, nreg))
            L.ab_mesh_set_async(mesh.h, 0)

            def pipe_run(n):
                ab.lib.check(L.ab_stage_upload_all(mesh.h, inp))
                for s_ in range(n):
                    ab.lib.check(L.ab_stage_commit(mesh.h))
                    if s_ + 1 < n:
                        ab.lib.check(L.ab_stage_upload_all(mesh.h, inp))
                    ab.lib.check(L.ab_mesh_initialize(mesh.h))
                    ab.lib.check(L.ab_mesh_cycles(mesh.h, 1))
                    ab.lib.check(L.ab_stage_download_all(mesh.h, outp))
                ab.lib.check(L.ab_stage_sync(mesh.h))
            npipe = max(2, a.steps)      # fill + drain of the pipeline amortise over the steps
            pipe_run(1)
            barrier()
            t0 = time.perf_counter()
            pipe_run(npipe)
            barrier()
            dtp = time.perf_counter() - t0
            if dist:
                t = torch.tensor([dtp], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtp = float(t.item())
            nb_ = sum(t.numel()*8 for d in pinned for t in d.values())*max(world, 1)
            e2e_pipe = {"value": zones*npipe/dtp, "unit": "zone-cycles/s", "steps": npipe,
                        "h2d_bytes_per_step": nb_, "d2h_bytes_per_step": nb_,
                        "what": "same step as e2e; uploads / downloads of neighbouring steps "
                                "overlap the kernels (copy streams + device staging buffers)"}
            del outbuf
        except Exception as ex:
            e2e_pipe = {"value": None, "error": str(ex)[:200]}

    # headline e2e: the pipelined sequence when it ran (same bytes, same work per step, copies
    # inside the timed region); the strictly serial sequence stays on the line as e2e_plain
    e2e_plain = None
    if e2e_pipe and e2e_pipe.get("value"):
        e2e_plain, e2e = e2e, e2e_pipe
    elif e2e_pipe:
        e2e_plain = e2e_pipe       # carries the error text

    # ---- the drop-in's normal mode: state stays resident, the host loop calls one cycle at a
    # time and reads back what Mesh::NewTimeStep / HistoryOutput need (dt, time, history sums)
    e2e_res = None
    if not a.no_e2e:
        try:
            nres = max(1, min(a.steps, 5))
            L.ab_mesh_set_async(mesh.h, 0)
            hist = (C.c_double*32)()
            barrier()
            t0 = time.perf_counter()
            d2h = 0
            for _ in range(nres):
                ab.lib.check(L.ab_mesh_cycles(mesh.h, 1))       # syncs: reads time, dt, ncycle
                nq = ab.lib.check(L.ab_history(mesh.h, hist, 32))
                d2h = 6*8 + nq*8
            barrier()
            dtw = time.perf_counter() - t0
            if dist:
                t = torch.tensor([dtw], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dtw = float(t.item())
            e2e_res = {"value": zones*nres/dtw, "unit": "zone-cycles/s", "steps": nres,
                       "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                       "what": "state resident on the GPU; per step the host calls "
                               "ab_mesh_cycles(1) (synchronises; time, dt, ncycle come back) and "
                               "ab_history (device reduction + NCCL sum, 11 doubles back)"}
        except Exception as ex:
            e2e_res = {"value": None, "error": str(ex)[:200]}

    cpu_full = None
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        try:
            cpu_full = run_reference_cpu(wl)
            if cpu_full:
                cpu = {k: cpu_full[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:   # the baseline must never break the product's bench line
            cpu = {"value": None, "unit": "zone-cycles/s", "cores": os.cpu_count(),
                   "kind": "reference", "sample": "failed: %s" % ex}

    # facts of the main mesh the line needs, then free its ~59 GB before the side legs
    mesh_dims = [mesh.params.nx1, mesh.params.nx2, mesh.params.nx3]
    nbtotal = mesh.nbtotal
    import gc
    del mesh, pinned
    gc.collect()

    # ---- same_config (N=1): the GPU on exactly the <mesh>/<meshblock> the CPU baseline and the
    # `--impl reference` arm run (a bounded sample of the workload with enough MeshBlocks for the
    # host's OpenMP threads), so that a like-for-like ratio exists beside the headline
    same_cfg = None
    others = None
    if rank == 0 and world == 1 and not a.no_side:
        try:
            smesh, sblk = reference_sample(wl, os.cpu_count() or 1)
            if cpu_full:
                smesh, sblk = cpu_full["mesh"], cpu_full["meshblock"]
            r = side_workload(ab, wl, device, max(4, min(a.steps, 10)), 3, block=tuple(sblk),
                              per_gpu=tuple(smesh))
            same_cfg = {"mesh": smesh, "meshblock": sblk, "meshblocks": r["meshblocks"],
                        "value": r["value"], "unit": "zone-cycles/s",
                        "ms_per_step": r["ms_per_step"], "gpu_launches": r["gpu_launches"],
                        "cpu_value": cpu["value"] if cpu else None,
                        "cpu_cores": cpu["cores"] if cpu else None}
        except Exception as ex:
            same_cfg = {"value": None, "error": str(ex)[:300]}
        # ---- the other BASELINE.json configurations, short device-resident runs
        others = {}
        side = {"c2": (dict(), 100, 10),
                "c3": (dict(block=(512, 512, 1), per_gpu=(2048, 2048, 1)), 20, 5),
                "c4": (dict(block=(128, 128, 128), per_gpu=(512, 512, 512)), 4, 2)}
        if wl != "c5":
            side["c5"] = (dict(), 4, 2)
        for name, (kw, st, wu) in side.items():
            if name == wl:
                continue
            try:
                others[name] = side_workload(ab, name, device, st, wu, **kw)
            except Exception as ex:
                others[name] = {"value": None, "error": str(ex)[:300]}
        # developed-flow figure for the c5 kernels (HLLD+PLM+VL2, 3-D MHD): the linear-wave
        # problem of c2 on a 512x256x256 MeshBlock -- every interface has non-zero jumps in every
        # variable, none of the zero-dividend shortcuts the mostly static blast state takes fires
        try:
            r = side_workload(ab, "c2", device, 6, 3, block=(512, 256, 256), per_gpu=(512, 256, 256))
            r["what"] = ("c2's linear wave on one 512x256x256 MeshBlock: same kernels as c5 on a "
                         "state with gradients everywhere; compare ms per interface with c5")
            r["flux_ns_per_interface"] = {k: 1e6*v/(513.0*256*256) for k, v in
                                          r["flux_avg_ms_by_dir_order"].items()}
            others["c5_kernels_on_developed_flow"] = r
        except Exception as ex:
            others["c5_kernels_on_developed_flow"] = {"value": None, "error": str(ex)[:300]}

    # ---- multi-rank correctness on the scaling record: the golden fixtures of the reference,
    # MeshBlocks sharded over these ranks, NCCL ghost / EMF exchange and dt reduction, bit for
    # bit (tests/multirank_check.py; checker only, after every timed region)
    parity = None
    if dist:
        try:
            for pth in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
                if pth not in sys.path:
                    sys.path.insert(0, pth)
            import multirank_check
            nrun, nfail = multirank_check.check_goldens(rank, world, local_rank, verbose=False)
            parity = {"ranks": world, "fixtures": nrun, "failed": nfail,
                      "what": "reference goldens sharded over the ranks, state and dt bit-exact"}
        except Exception as ex:
            parity = {"ranks": world, "fixtures": 0, "failed": None, "error": str(ex)[:300]}

    if rank == 0:
        cfg_desc.update({"mesh": mesh_dims,
                         "meshblock": list(blk), "meshblocks_total": nbtotal,
                         "parallelism": "MeshBlocks sharded over %d GPU(s), NCCL halo" % world,
                         "l2_policy": "working set (%.1f GB/GPU) >> 126 MB L2; no flush needed"
                                      % (zones_local*8*55/1e9 if mhd else zones_local*8*35/1e9)})
        out = {"metric": "zone-cycles/sec", "value": value, "unit": "zone-cycles/s",
               "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms/a.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": cfg_desc, "roofline": roof,
               "roofline_cycle": cycle_roof, "cpu_baseline": cpu, "e2e": e2e,
               "e2e_resident": e2e_res, "e2e_plain": e2e_plain,
               "same_config": same_cfg, "other_workloads": others, "parity": parity,
               "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out))
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
