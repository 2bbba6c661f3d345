// ab_batch.cuh -- one launch per task over ALL MeshBlocks of a rank.
//
// The reference runs a task for every MeshBlock of a rank (task_list.cpp:71-88).  All MeshBlocks
// have the same shape, every array of a block lives at the same offset of the block's slab and the
// slabs of a rank's blocks are BlkDev::bstride bytes apart in one allocation (ab_mesh.cu,
// alloc_blocks).  A batched launch passes block 0's view and sets gridDim.y (gridDim.z for the
// kernels whose y counts slabs) to the number of blocks: block n's view is block 0's with every
// pointer moved by n*bstride.  A single-block launch has gridDim.y = 1, i.e. a zero offset.
#ifndef AB_BATCH_CUH_
#define AB_BATCH_CUH_
#include "ab_kernels.h"

namespace ab {

__device__ __forceinline__ int fast_div(int t, FastDiv f) {     // ab_types.h
  return f.m ? (int)(__umulhi((unsigned)t, f.m) >> f.s) : t;
}

template <class T>
__device__ __forceinline__ T *blk_mv(T *p, long off) {
  return (T *)((char *)p + off);
}
template <class T>
__device__ __forceinline__ T *blk_mv_opt(T *p, long off) {     // pointers that may be null
  return p ? (T *)((char *)p + off) : p;
}

// the view of local block n (only the members a kernel uses cost anything)
__device__ __forceinline__ BlkDev blk_view(const BlkDev &b0, unsigned n) {
  BlkDev b = b0;
  const long off = (long)n*b0.bstride;
  b.u = blk_mv(b.u, off); b.u1 = blk_mv(b.u1, off); b.w = blk_mv(b.w, off);
  b.bcc = blk_mv(b.bcc, off); b.cc_e = blk_mv(b.cc_e, off);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    b.b[d] = blk_mv(b.b[d], off); b.b1[d] = blk_mv(b.b1[d], off);
    b.flux[d] = blk_mv(b.flux[d], off);
    b.ef[d][0] = blk_mv(b.ef[d][0], off); b.ef[d][1] = blk_mv(b.ef[d][1], off);
    b.wght[d] = blk_mv(b.wght[d], off); b.e[d] = blk_mv(b.e[d], off);
    b.sflux[d] = blk_mv(b.sflux[d], off);
  }
  b.s = blk_mv(b.s, off); b.s1 = blk_mv(b.s1, off); b.r = blk_mv(b.r, off);
  b.x1f = blk_mv(b.x1f, off); b.x2f = blk_mv(b.x2f, off); b.x3f = blk_mv(b.x3f, off);
  b.x1v = blk_mv(b.x1v, off); b.x2v = blk_mv(b.x2v, off); b.x3v = blk_mv(b.x3v, off);
  b.dx1f = blk_mv(b.dx1f, off); b.dx2f = blk_mv(b.dx2f, off); b.dx3f = blk_mv(b.dx3f, off);
  b.bcw = blk_mv_opt(b.bcw, off);
  return b;
}

// blockIdx.y read where it is called (not hoisted to the top of the kernel): the Riemann-sweep
// kernels shift their OUTPUT pointers with it right before the stores, so that no shifted
// pointer stays live in registers across the solver (they run at the register limit)
__device__ __forceinline__ unsigned late_block_y() {
#if defined(__CUDA_ARCH__)
  unsigned y;
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(y));
  return y;
#else
  return blockIdx.y;
#endif
}

__device__ __forceinline__ ReconGeom geom_view(const ReconGeom &g0, const BlkDev &b0, unsigned n) {
  ReconGeom g = g0;
  const long off = (long)n*b0.bstride;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    g.wp[d] = blk_mv(g.wp[d], off); g.wm[d] = blk_mv(g.wm[d], off);
    g.nu[d] = blk_mv_opt(g.nu[d], off);
  }
  return g;
}

}  // namespace ab
#endif
