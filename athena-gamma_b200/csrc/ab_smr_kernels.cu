// ab_smr_kernels.cu -- device kernels of static mesh refinement for cell-centred variables
// (hydro + passive scalars): restriction, prolongation, the coarse-buffer EOS and boundary
// functions of ProlongateBoundaries, and the fine -> coarse flux correction.
//
// Reference: src/mesh/mesh_refinement.cpp:106-176,386-540; src/bvals/bvals_refine.cpp:96-570;
// src/bvals/cc/flux_correction_cc.cpp:69-290.  The per-cell bodies are in ab_smr_cells.cuh
// (__host__ __device__: tests/hostcheck runs the same code on the CPU against the oracle); the
// index boxes come from the host planner (ab_smr.cpp) through ab_smr_exec.h.
// All kernels are small (ghost-zone volumes): one thread per coarse cell, x1 fastest.
// Compile with -fmad=false like the rest.
#include "ab_kernels.h"
#include "ab_smr_cells.cuh"

namespace ab {

extern std::atomic<long> g_launches;

namespace {

constexpr int SB = 128;

__device__ __forceinline__ bool box_cell(const SmrBox &bx, int &i, int &j, int &k) {
  const int ni = bx.ei - bx.si + 1, nj = bx.ej - bx.sj + 1, nk = bx.ek - bx.sk + 1;
  const long t = (long)blockIdx.x*SB + threadIdx.x;
  if (t >= (long)ni*nj*nk) return false;
  const long r = t / ni;
  i = bx.si + (int)(t - r*ni);
  k = bx.sk + (int)(r / nj);
  j = bx.sj + (int)(r - (long)(k - bx.sk)*nj);
  return true;
}

inline unsigned box_grid(const SmrBox &bx) {
  const long n = (long)(bx.ei - bx.si + 1)*(bx.ej - bx.sj + 1)*(bx.ek - bx.sk + 1);
  return (unsigned)((n + SB - 1)/SB);
}

__global__ void __launch_bounds__(SB) k_smr_restrict(SmrGeom g, const double *__restrict__ fine,
                                                    double *__restrict__ coarse, int nvar,
                                                    SmrBox bx) {
  int ci, cj, ck;
  if (box_cell(bx, ci, cj, ck)) smr_restrict_cell(g, fine, coarse, nvar, ci, cj, ck);
}
__global__ void __launch_bounds__(SB) k_smr_prolong(SmrGeom g, const double *__restrict__ coarse,
                                                   double *__restrict__ fine, int nvar,
                                                   SmrBox bx) {
  int i, j, k;
  if (box_cell(bx, i, j, k)) smr_prolong_cell(g, coarse, fine, nvar, i, j, k);
}
__global__ void __launch_bounds__(SB) k_smr_c2p(SmrGeom g, Params p, double *cu, double *cw, int ns,
                                               double *cs, double *cr, SmrBox bx) {
  int i, j, k;
  if (box_cell(bx, i, j, k)) smr_c2p_cell(g, p, cu, cw, ns, cs, cr, i, j, k);
}
__global__ void __launch_bounds__(SB) k_smr_bc(SmrGeom g, double *cw, int nh, double *cr, int ns,
                                              int face, int refl, int lo, int hi, SmrBox bx) {
  int i, j, k;
  if (box_cell(bx, i, j, k)) smr_bc_cell(g, cw, nh, cr, ns, face, refl, lo, hi, i, j, k);
}
__global__ void __launch_bounds__(SB) k_smr_flux(SmrGeom gf, const double *__restrict__ ffl,
                                                double *__restrict__ cfl, int nvar, int dir,
                                                int fpos, int cpos, int a0, int b0, int na,
                                                int nb, int compact) {
  const int t = blockIdx.x*SB + threadIdx.x;
  if (t >= na*nb) return;
  const int ib = t / na, ia = t - ib*na;
  smr_flux_cell(gf, ffl, cfl, nvar, dir, fpos, cpos, a0, b0, ia, ib, compact ? na : 0, nb);
}

}  // namespace

void launch_smr_restrict(const SmrGeom &g, const double *fine, double *coarse, int nvar,
                         const SmrBox &bx, cudaStream_t s) {
  if (nvar <= 0) return;
  k_smr_restrict<<<box_grid(bx), SB, 0, s>>>(g, fine, coarse, nvar, bx); ++g_launches;
}
void launch_smr_prolong(const SmrGeom &g, const double *coarse, double *fine, int nvar,
                        const SmrBox &bx, cudaStream_t s) {
  if (nvar <= 0) return;
  k_smr_prolong<<<box_grid(bx), SB, 0, s>>>(g, coarse, fine, nvar, bx); ++g_launches;
}
void launch_smr_c2p(const SmrGeom &g, const Params &p, double *cu, double *cw, int ns, double *cs,
                    double *cr, const SmrBox &bx, cudaStream_t s) {
  k_smr_c2p<<<box_grid(bx), SB, 0, s>>>(g, p, cu, cw, ns, cs, cr, bx); ++g_launches;
}
void launch_smr_bc(const SmrGeom &g, double *cw, int nh, double *cr, int ns, int face, int refl,
                   int lo, int hi, const SmrBox &bx, cudaStream_t s) {
  k_smr_bc<<<box_grid(bx), SB, 0, s>>>(g, cw, nh, cr, ns, face, refl, lo, hi, bx); ++g_launches;
}
void launch_smr_flux(const SmrGeom &gf, const double *fine_flux, double *coarse_flux, int nvar,
                     int dir, int fpos, int cpos, int a0, int b0, int na, int nb,
                     cudaStream_t s, int compact) {
  if (nvar <= 0 || na*nb <= 0) return;
  k_smr_flux<<<(na*nb + SB - 1)/SB, SB, 0, s>>>(gf, fine_flux, coarse_flux, nvar, dir, fpos, cpos,
                                                a0, b0, na, nb, compact); ++g_launches;
}

}  // namespace ab
