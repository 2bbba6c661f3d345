#!/usr/bin/env python
"""Can one process per GPU map another process's device buffer (cudaIpc*) on this box, and how
fast are stores into it?  torchrun --nproc-per-node 2 tools/ipc_probe.py"""
import os
import time

import torch
import torch.distributed as dist
from cuda import cudart


def ck(r):
    if isinstance(r, tuple):
        err, rest = r[0], r[1:]
    else:
        err, rest = r, ()
    if int(err) != 0:
        raise RuntimeError("CUDA error %s" % err)
    return rest[0] if len(rest) == 1 else rest


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 256 << 20
    ptr = ck(cudart.cudaMalloc(n))
    ck(cudart.cudaMemset(ptr, rank + 1, n))
    h = ck(cudart.cudaIpcGetMemHandle(ptr))
    handles = [None]*world
    dist.all_gather_object(handles, bytes(h.reserved))
    peer = (rank + 1) % world
    ph = cudart.cudaIpcMemHandle_t()
    ph.reserved = handles[peer]
    try:
        pptr = ck(cudart.cudaIpcOpenMemHandle(ph, cudart.cudaIpcMemLazyEnablePeerAccess))
    except Exception as ex:
        print("rank %d: cudaIpcOpenMemHandle FAILED: %s" % (rank, ex), flush=True)
        dist.barrier()
        return
    src = ck(cudart.cudaMalloc(n))
    ck(cudart.cudaMemset(src, 0x40 + rank, n))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        ck(cudart.cudaMemcpy(pptr, src, n, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice))
    ck(cudart.cudaDeviceSynchronize())
    dt = (time.perf_counter() - t0)/5
    dist.barrier()
    back = torch.empty(16, dtype=torch.uint8)
    ck(cudart.cudaMemcpy(back.data_ptr(), ptr, 16, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost))
    writer = (rank - 1) % world
    print("rank %d: mapped peer %d ok; my buffer now holds 0x%x (expected 0x%x from rank %d); "
          "copy into the peer %.1f GB/s" % (rank, peer, int(back[0]), 0x40 + writer, writer, n/dt/1e9),
          flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
