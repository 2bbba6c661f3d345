// TEST-ONLY.  A minimal CUDA-on-the-host emulation so that the CPU test-suite can run the
// product's device sources (csrc/*.cu: kernels, launch glue, C ABI) without a GPU:
//   * __global__ functions become plain functions; a launch (rewritten from <<< >>> by
//     tests/hostcheck/build_mesh_host.py) runs the grid block by block, thread by thread, in
//     order (or last to first with AB_EMU_ORDER=reverse) -- kernels that use __syncthreads / warp shuffles are run with one fiber per thread
//     and every such primitive is a block-wide phase boundary;
//   * the runtime API is synchronous and in-order: streams and events are inert handles,
//     "device" memory is host memory (filled with NaN bit patterns on allocation so that a read
//     of never-written device memory shows up).
// It checks indexing, launch geometry, work lists and the order of operations -- everything that
// is independent of the GPU's parallel execution.  It is not part of the product: the shipped
// libathena_b200.so has no host execution path and fails loudly without a CUDA device.
#ifndef AB_EMU_CUDA_RUNTIME_H_
#define AB_EMU_CUDA_RUNTIME_H_
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
typedef struct ab_emu_stream *cudaStream_t;
typedef struct ab_emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1,
                      cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3,
                      cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };

inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulation error" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) {
  *p = (T *)malloc(n ? n : 1);
  if (!*p) return cudaErrorEmu;
  memset((void *)*p, 0xFF, n);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind,
                                   cudaStream_t = 0) {
  memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) {
  memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = (cudaStream_t)malloc(8);
  return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }

inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a*b) >> 32); }
inline unsigned long long atomicMin(unsigned long long *a, unsigned long long v) {
  unsigned long long old = *a;
  if (v < old) *a = v;
  return old;
}

namespace ab_emu {
// one block at a time; cooperative kernels: one fiber per thread, every barrier ends a phase
struct Fiber { ucontext_t ctx; char *stack; bool done; };
struct State {
  bool coop = false;            // set by launch_coop for the block being run
  int cur = -1;                 // running fiber
  ucontext_t sched;
  std::vector<Fiber> fib;
  unsigned long long xch[1024];
  void (*entry)(void *) = nullptr;
  void *arg = nullptr;
};
inline State g;
static const size_t STACK = 256*1024;

// AB_EMU_ORDER=reverse runs blocks and threads last to first: a kernel whose result depends on
// the order in which its threads run (one thread reading what another writes) shows up as a
// difference between the two orders
inline bool reversed() {
  static const int r = [] { const char *e = getenv("AB_EMU_ORDER"); return (e && e[0] == 'r') ? 1 : 0; }();
  return r != 0;
}

inline void barrier() {
  if (!g.coop) {
    fprintf(stderr, "ab_emu: __syncthreads / shuffle in a kernel not launched cooperatively "
                    "(add it to COOP in tests/hostcheck/build_mesh_host.py)\n");
    abort();
  }
  swapcontext(&g.fib[g.cur].ctx, &g.sched);
}
inline void trampoline() {
  g.entry(g.arg);
  g.fib[g.cur].done = true;
  swapcontext(&g.fib[g.cur].ctx, &g.sched);
}
template <class F> inline void call_fn(void *p) { (*(F *)p)(); }

template <class F> inline void run_block(dim3 b, F &fn, bool coop) {
  const int nt = (int)(b.x*b.y*b.z);
  if (!coop) {
    for (int i = 0; i < nt; ++i) {
      const unsigned t = (unsigned)(reversed() ? nt - 1 - i : i);
      threadIdx = {t % b.x, (t / b.x) % b.y, t / (b.x*b.y)};
      fn();
    }
    return;
  }
  if ((int)g.fib.size() < nt) {
    size_t old = g.fib.size();
    g.fib.resize(nt);
    for (size_t i = old; i < g.fib.size(); ++i) g.fib[i].stack = (char *)malloc(STACK);
  }
  g.coop = true;
  g.entry = &call_fn<F>;
  g.arg = (void *)&fn;
  for (int t = 0; t < nt; ++t) {
    Fiber &f = g.fib[t];
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = STACK;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())trampoline, 0);
  }
  int live = nt;
  while (live > 0) {            // one pass = one phase between two barriers
    live = 0;
    for (int i = 0; i < nt; ++i) {
      const int t = reversed() ? nt - 1 - i : i;
      if (g.fib[t].done) continue;
      g.cur = t;
      threadIdx = {(unsigned)t % b.x, ((unsigned)t / b.x) % b.y, (unsigned)t / (b.x*b.y)};
      swapcontext(&g.sched, &g.fib[t].ctx);
      if (!g.fib[t].done) ++live;
    }
  }
  g.coop = false;
  g.cur = -1;
}

template <class F> inline void launch(dim3 gr, dim3 b, bool coop, F fn) {
  gridDim = gr;
  blockDim = b;
  const long nb = (long)gr.x*gr.y*gr.z;
  for (long i = 0; i < nb; ++i) {
    const long c = reversed() ? nb - 1 - i : i;
    blockIdx = {(unsigned)(c % gr.x), (unsigned)((c / gr.x) % gr.y), (unsigned)(c / ((long)gr.x*gr.y))};
    run_block(b, fn, coop);
  }
}
}  // namespace ab_emu

inline void __syncthreads() { ab_emu::barrier(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int s) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned t = threadIdx.x + blockDim.x*(threadIdx.y + blockDim.y*threadIdx.z);
  memcpy(&ab_emu::g.xch[t], &v, sizeof(T));
  ab_emu::barrier();
  T r;
  memcpy(&r, &ab_emu::g.xch[t ^ (unsigned)s], sizeof(T));
  ab_emu::barrier();
  return r;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned t = threadIdx.x + blockDim.x*(threadIdx.y + blockDim.y*threadIdx.z);
  memcpy(&ab_emu::g.xch[t], &v, sizeof(T));
  ab_emu::barrier();
  T r = v;                                   // lanes below d keep their own value
  if ((t & 31u) >= d) memcpy(&r, &ab_emu::g.xch[t - d], sizeof(T));
  ab_emu::barrier();
  return r;
}
#endif  // AB_EMU_CUDA_RUNTIME_H_
