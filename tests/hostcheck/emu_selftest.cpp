// TEST-ONLY: positive controls for the CUDA stand-in (emu/cuda_runtime.h) -- that the checks
// built on it can fail.  Launches are written in the rewritten form directly.
//   emu_selftest shift   : a kernel with a dependence between threads of one launch prints a
//                          checksum that differs between AB_EMU_ORDER=forward and =reverse
//   emu_selftest reduce  : shuffles + __syncthreads give the block sum (fibers)
//   emu_selftest oob     : an out-of-range store (AddressSanitizer build aborts)
#include "cuda_runtime.h"

__global__ void k_shift(double *a, int n) {
  int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t + 1 < n) a[t] = a[t + 1] + 1.0;       // reads what thread t+1 writes: a race on a GPU
}
__global__ void k_block_sum(const double *a, double *out) {
  __shared__ double sm[4];
  double v = a[blockIdx.x*blockDim.x + threadIdx.x];
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0] + sm[1] + sm[2] + sm[3];
}
__global__ void k_store(double *a, int i) { a[i] = 1.0; }

int main(int argc, char **argv) {
  const int n = 512;
  double *a, *out;
  cudaMalloc(&a, n*sizeof(double));
  cudaMalloc(&out, 4*sizeof(double));
  for (int i = 0; i < n; ++i) a[i] = (double)i;
  const char *mode = argc > 1 ? argv[1] : "";
  if (!strcmp(mode, "shift")) {
    ab_emu::launch(4, 128, false, [&]() { k_shift(a, n); });
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += a[i]*(i + 1);
    printf("%.17g\n", s);
  } else if (!strcmp(mode, "reduce")) {
    ab_emu::launch(4, 128, true, [&]() { k_block_sum(a, out); });
    printf("%.17g %.17g %.17g %.17g\n", out[0], out[1], out[2], out[3]);
  } else if (!strcmp(mode, "oob")) {
    ab_emu::launch(1, 1, false, [&]() { k_store(a, n + 2); });
    printf("stored\n");
  }
  cudaFree(a);
  cudaFree(out);
  return 0;
}
