// fluxbench.cu -- kernel-tuning harness for the reconstruct+Riemann sweep (k_flux), stand-alone.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
//        -I athena-gamma_b200/csrc [-DAB_...] -o fluxbench tools/fluxbench.cu
//   ./fluxbench [n=256] [mode=0|1] [reps=5] [ng=2] [hydro=0|1]   (hydro=1: HLLC instead of HLLD)
//
// Fills one n^3 MHD MeshBlock (w, bcc, face b) with either a mostly static state (mode 0: the
// 512^3 blast after a few cycles is uniform outside a small sphere) or a developed flow (mode
// 1: every cell has velocity and field gradients), runs every (dir, order) HLLD sweep `reps`
// times, prints the CUDA-event average per launch and a 64-bit checksum of all outputs.  Two
// builds with different -D flags must print identical checksums: the tuning knobs may not
// change a single bit.  Not part of the product; nothing here is linked into the library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "ab_flux.cuh"

namespace ab { std::atomic<long> g_launches{0}; }
using namespace ab;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_checksum(const double *p, long n, unsigned long long *out) {
  unsigned long long acc = 0;
  for (long i = blockIdx.x*(long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x*blockDim.x) {
    unsigned long long v = (unsigned long long)__double_as_longlong(p[i]);
    acc += v*(unsigned long long)(2*i + 1);
  }
  atomicAdd(out, acc);
}

static double *dalloc(long n, int fill_nan = 1) {
  double *p; CK(cudaMalloc(&p, n*sizeof(double)));
  CK(cudaMemset(p, fill_nan ? 0xFF : 0, n*sizeof(double)));
  return p;
}

int main(int argc, char **argv) {
  int n = argc > 1 ? atoi(argv[1]) : 256;
  int mode = argc > 2 ? atoi(argv[2]) : 0;
  int reps = argc > 3 ? atoi(argv[3]) : 5;
  int ng = argc > 4 ? atoi(argv[4]) : 2;
  int hydro = argc > 5 ? atoi(argv[5]) : 0;
  BlkDev b; memset(&b, 0, sizeof(b));
  b.ng = ng; b.nc1 = b.nc2 = b.nc3 = n + 2*ng;
  b.is = b.js = b.ks = ng; b.ie = b.je = b.ke = ng + n - 1;
  b.f2 = b.f3 = 1; b.nh = 5;
  const long nc = b.nc1, ncell = nc*nc*nc;
  const long nf1 = nc*nc*(nc+1);
  // host state
  std::vector<double> w(5*ncell), bcc(3*ncell), bf[3];
  for (int d = 0; d < 3; ++d) bf[d].assign(nf1, 0.0);
  const double dx = 1.0/n, PI = 3.14159265358979323846;
  auto X = [&](long i) { return (i - ng + 0.5)*dx; };
  auto bfield = [&](int comp, double x, double y, double z) -> double {
    if (mode == 0) return comp == 0 ? 0.8660254037844387 : (comp == 1 ? 0.5 : 0.0);
    // divergence-free enough for a flux benchmark: each component independent of its own axis
    if (comp == 0) return 0.7 + 0.3*sin(2*PI*y)*cos(2*PI*z);
    if (comp == 1) return 0.4 + 0.3*sin(2*PI*z + 1.0)*cos(2*PI*x);
    return 0.2 + 0.3*sin(2*PI*x + 2.0)*cos(2*PI*y);
  };
  for (long k = 0; k < nc; ++k) for (long j = 0; j < nc; ++j) for (long i = 0; i < nc; ++i) {
    long o = (k*nc + j)*nc + i;
    double x = X(i), y = X(j), z = X(k);
    double r = sqrt((x-.5)*(x-.5) + (y-.5)*(y-.5) + (z-.5)*(z-.5));
    double d, vx, vy, vz, p;
    if (mode == 0) {
      const bool in = r < 0.12;     // small disturbed sphere, everything else exactly uniform
      d = in ? 1.0 + 0.3*cos(9*r/0.12) : 1.0;
      p = in ? 0.1 + 5.0*(0.12 - r)/0.12 : 0.1;
      vx = in ? 2.0*(x-.5) : 0.0; vy = in ? 2.0*(y-.5) : 0.0; vz = in ? 2.0*(z-.5) : 0.0;
    } else {
      d = 1.0 + 0.4*sin(2*PI*x)*sin(4*PI*y + 0.3)*cos(2*PI*z);
      p = 0.6 + 0.3*cos(4*PI*x + 0.2)*sin(2*PI*y)*sin(2*PI*z + 0.7);
      vx = 0.8*sin(2*PI*y + 0.1)*cos(2*PI*z); vy = 0.8*sin(2*PI*z + 0.5)*cos(4*PI*x);
      vz = 0.8*sin(4*PI*x + 0.9)*cos(2*PI*y);
    }
    w[o] = d; w[ncell + o] = vx; w[2*ncell + o] = vy; w[3*ncell + o] = vz; w[4*ncell + o] = p;
    for (int c = 0; c < 3; ++c) bcc[c*ncell + o] = bfield(c, x, y, z);
  }
  for (long k = 0; k < nc; ++k) for (long j = 0; j < nc; ++j) for (long i = 0; i <= nc; ++i)
    bf[0][(k*nc + j)*(nc+1) + i] = bfield(0, X(i) - 0.5*dx, X(j), X(k));
  for (long k = 0; k < nc; ++k) for (long j = 0; j <= nc; ++j) for (long i = 0; i < nc; ++i)
    bf[1][(k*(nc+1) + j)*nc + i] = bfield(1, X(i), X(j) - 0.5*dx, X(k));
  for (long k = 0; k <= nc; ++k) for (long j = 0; j < nc; ++j) for (long i = 0; i < nc; ++i)
    bf[2][(k*nc + j)*nc + i] = bfield(2, X(i), X(j), X(k) - 0.5*dx);
  b.w = dalloc(5*ncell); b.bcc = dalloc(3*ncell);
  CK(cudaMemcpy(b.w, w.data(), 5*ncell*8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b.bcc, bcc.data(), 3*ncell*8, cudaMemcpyHostToDevice));
  for (int d = 0; d < 3; ++d) {
    b.b[d] = dalloc(nf1);
    CK(cudaMemcpy(b.b[d], bf[d].data(), nf1*8, cudaMemcpyHostToDevice));
    b.flux[d] = dalloc(5*nf1); b.ef[d][0] = dalloc(nf1); b.ef[d][1] = dalloc(nf1);
    b.wght[d] = dalloc(nf1);
  }
  std::vector<double> dxf(nc + 1, dx), half(nc + 1, 0.5);
  double *ddx = dalloc(nc + 1), *dhalf = dalloc(nc + 1);
  CK(cudaMemcpy(ddx, dxf.data(), (nc+1)*8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dhalf, half.data(), (nc+1)*8, cudaMemcpyHostToDevice));
  b.dx1f = b.dx2f = b.dx3f = ddx;
  ReconGeom g; memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) { g.wp[d] = dhalf; g.wm[d] = dhalf; }
  Params p; memset(&p, 0, sizeof(p));
  p.gamma = 5.0/3.0; p.dfloor = 3.4e-18; p.pfloor = 3.4e-18; p.mhd = !hydro;
  p.solver = hydro ? SOLVER_HLLC : SOLVER_HLLD;
  p.xorder = 2;
  const double dt = 0.3*dx/2.0;
  unsigned long long *dsum; CK(cudaMalloc(&dsum, 8));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int maxorder = ng >= 3 ? 3 : 2;
  double total = 0.0;
  unsigned long long all = 0;
  for (int order = 1; order <= maxorder; ++order) for (int dir = 0; dir < 3; ++dir) {
    auto run = [&]() {
      if (hydro) flux_order<SOLVER_HLLC,false,false>(b, g, p, order, dir, dt, nullptr, 0, 1);
      else flux_order<SOLVER_HLLD,true,false>(b, g, p, order, dir, dt, nullptr, 0, 1);
    };
    for (int r = 0; r < 2; ++r) run();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemset(dsum, 0, 8));
    // only the faces the sweep writes are defined; the rest stays at the NaN fill (0xFF..),
    // identical in every build, so the checksum covers exactly the written values
    k_checksum<<<1024, 256>>>(b.flux[dir], 5*nf1, dsum);
    if (!hydro) {
      k_checksum<<<1024, 256>>>(b.ef[dir][0], nf1, dsum);
      k_checksum<<<1024, 256>>>(b.ef[dir][1], nf1, dsum);
      k_checksum<<<1024, 256>>>(b.wght[dir], nf1, dsum);
    }
    unsigned long long h; CK(cudaMemcpy(&h, dsum, 8, cudaMemcpyDeviceToHost));
    printf("x%d_o%d %8.4f ms  sum %016llx\n", dir + 1, order, ms/reps, h);
    total += ms/reps; all ^= h*(unsigned long long)(dir*7 + order);
  }
  printf("TOTAL n=%d mode=%d ng=%d hydro=%d: %.4f ms  checksum %016llx\n", n, mode, ng, hydro, total, all);
  return 0;
}
