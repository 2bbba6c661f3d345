#!/usr/bin/env python
"""Host<->device copy rates from pinned memory with and without binding the process to the
GPU's NUMA node before the pinned allocation (the e2e legs of bench.py are PCIe-bound)."""
import os
import subprocess
import sys
import time

import torch


def gpu_numa(dev):
    p = torch.cuda.get_device_properties(dev)
    bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    base = "/sys/bus/pci/devices/" + bdf
    node = open(base + "/numa_node").read().strip()
    cpus = open(base + "/local_cpulist").read().strip()
    return bdf, int(node), cpus


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def rates(tag, n=1 << 30):
    dev = torch.device("cuda", 0)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def t(fn, reps=4):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0)/reps

    def up():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def down():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        up(); down()
    print("%s: H2D %.1f GB/s  D2H %.1f GB/s  both %.1f + %.1f GB/s" %
          (tag, n/t(up)/1e9, n/t(down)/1e9, n/t(both)/1e9, n/t(both)/1e9), flush=True)


if __name__ == "__main__":
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:])
    print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    try:
        print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout[:1500])
    except Exception as ex:
        print("lscpu:", ex)
    bdf, node, cpus = gpu_numa(0)
    print("gpu0", bdf, "numa", node, "cpus", cpus)
    rates("unbound")
    if cpus:
        os.sched_setaffinity(0, parse_cpulist(cpus))
        rates("bound to the GPU's cpus")
