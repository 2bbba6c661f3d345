// ab_kernels.cu -- hand-written sm_100a kernels of the per-MeshBlock hydro/MHD update.
//
// All arithmetic is FP64 on the CUDA cores (nothing on this path is a dense contraction, so
// no tensor cores); the file must be compiled with -fmad=false so that products and sums
// round exactly like the reference's SSE2 build.  Threads map to the fastest (x1) index so
// that every global access is coalesced; stencil neighbours are served by L1/L2.
#include <float.h>
#include <stdint.h>
#include <string.h>
#include "ab_kernels.h"
#include "ab_physics.cuh"
#include "ab_batch.cuh"

namespace ab {

// ---- AthenaArray-style indexing -------------------------------------------------------------
#define CCI(b, n, k, j, i) ((((long)(n)*(b).nc3 + (k))*(b).nc2 + (j))*(b).nc1 + (i))
#define F1I(b, k, j, i) (((long)(k)*(b).nc2 + (j))*((b).nc1 + 1) + (i))
#define F2I(b, k, j, i) (((long)(k)*((b).nc2 + 1) + (j))*(b).nc1 + (i))
#define F3I(b, k, j, i) (((long)(k)*(b).nc2 + (j))*(b).nc1 + (i))
#define E1I(b, k, j, i) (((long)(k)*((b).nc2 + 1) + (j))*(b).nc1 + (i))
#define E2I(b, k, j, i) (((long)(k)*(b).nc2 + (j))*((b).nc1 + 1) + (i))
#define E3I(b, k, j, i) (((long)(k)*((b).nc2 + 1) + (j))*((b).nc1 + 1) + (i))

std::atomic<long> g_launches{0};      // kernel launches issued (bench accounting)
constexpr int BX = 128;  // threads along x1 per CTA

static inline dim3 grid3(int ni, int nj, int nk) {
  return dim3((unsigned)((ni + BX - 1)/BX), (unsigned)nj, (unsigned)nk);
}

// =============================================================================================
// EquationOfState::ConservedToPrimitive (+ Field::CalculateCellCenteredField)
// eos/adiabatic_mhd.cpp:41-90, eos/adiabatic_hydro.cpp:39-80, field/field.cpp:112-180
// =============================================================================================
// CFL min-reduction tail: warp shuffles, one shared-memory step per CTA, then ONE atomicMin per
// CTA spread over DT_SLOTS addresses (positive doubles order like their bit patterns).  A
// single address would serialise ~10^5-10^6 atomics per launch in L2.
__device__ __forceinline__ void block_min_to_slots(double m, unsigned long long *slots) {
  __shared__ double sm[BX/32];
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    double o2 = __shfl_xor_sync(0xffffffffu, m, s);
    m = dmin(m, o2);
  }
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < BX/32; ++w) m = dmin(m, sm[w]);
    if (m < DBL_MAX) {
      unsigned slot = (blockIdx.x + 7u*blockIdx.y + 13u*blockIdx.z) & (DT_SLOTS - 1);
      slot = (slot ^ (blockIdx.x >> 6)) & (DT_SLOTS - 1);     // spreads flattened 1-D grids too
      atomicMin(slots + slot, (unsigned long long)__double_as_longlong(m));
    }
  }
}

// FLAGS bit0: also store the cell-centred EMF cc_e = -(v x B) used by ComputeCornerE
// (field/calculate_corner_e.cpp:131-180) -- saves a pass over w and bcc;
// bit1: also reduce Hydro::NewBlockTimeStep (hydro/new_blockdt.cpp:64-134) over the ACTIVE
// cells of the launch (last integrator stage) -- w, bcc, b are already in registers.
// Resident CTAs per SM asked of the compiler for the HBM-bound kernels (more loads in flight;
// measured together on B200 at 512^3: 76.9 -> 75.1 ms per cycle, tune4.log; 16/10/16/12 spills
// and is slower).
#ifndef AB_C2P_MINB
#define AB_C2P_MINB 6      // k_cons2prim with the fused CFL reduction
#endif
#ifndef AB_CE_MINB
#define AB_CE_MINB 10      // k_corner_e3d
#endif
#ifndef AB_FC_MINB
#define AB_FC_MINB 1       // k_integrate_fc: a hint of 10 cost 0.13 ms per launch (ncu v6 vs v9)
#endif
#ifndef AB_CC_MINB
#define AB_CC_MINB 8       // k_integrate_cc
#endif
// one cell of ConservedToPrimitive (+ cc_e, + the cell's CFL limit folded into m)
template <bool MHD, int FLAGS>
__device__ __forceinline__ void c2p_cell(const BlkDev &b, const Params &p, int i, int j, int k,
                                         double &m) {
  double gm1 = p.gamma - 1.0;
  double pb = 0.0;
  double bcc1 = 0.0, bcc2 = 0.0, bcc3 = 0.0, bf1 = 0.0, bf2 = 0.0, bf3 = 0.0;
  long o = CCI(b,0,k,j,i);
  long sv = (long)b.nc3*b.nc2*b.nc1;
  if (MHD) {
    bf1 = b.b[0][F1I(b,k,j,i)]; bf2 = b.b[1][F2I(b,k,j,i)]; bf3 = b.b[2][F3I(b,k,j,i)];
    if (b.bcw) {   // nonuniform spacing: interpolate to the volume centre (field.cpp:139-172)
      const double *__restrict__ q1 = b.bcw, *q2 = q1 + 2*b.nc1, *q3 = q2 + 2*b.nc2;
      bcc1 = q1[i]*bf1 + q1[b.nc1+i]*b.b[0][F1I(b,k,j,i+1)];
      bcc2 = q2[j]*bf2 + q2[b.nc2+j]*b.b[1][F2I(b,k,j+1,i)];
      bcc3 = q3[k]*bf3 + q3[b.nc3+k]*b.b[2][F3I(b,k+1,j,i)];
    } else {
      bcc1 = 0.5*bf1 + 0.5*b.b[0][F1I(b,k,j,i+1)];
      bcc2 = 0.5*bf2 + 0.5*b.b[1][F2I(b,k,j+1,i)];
      bcc3 = 0.5*bf3 + 0.5*b.b[2][F3I(b,k+1,j,i)];
    }
    b.bcc[o] = bcc1;
    b.bcc[o+sv] = bcc2;
    b.bcc[o+2*sv] = bcc3;
    pb = 0.5*(sqr(bcc1) + sqr(bcc2) + sqr(bcc3));
  }
  // isothermal EOS (eos/isothermal_{hydro,mhd}.cpp:39-75): density floor and velocities only
  const bool iso = (p.eos != 0);
  double u_d = b.u[o], u_m1 = b.u[o+sv], u_m2 = b.u[o+2*sv], u_m3 = b.u[o+3*sv];
  double u_e = iso ? 0.0 : b.u[o+4*sv];
  double u_d0 = u_d, u_e0 = u_e;
  u_d = (u_d > p.dfloor) ? u_d : p.dfloor;
  double di = 1.0/u_d;
  double w_p = 0.0;
  if (!iso) {
    double e_k = 0.5*di*(sqr(u_m1) + sqr(u_m2) + sqr(u_m3));
    if (MHD) {
      w_p = gm1*(u_e - e_k - pb);
      u_e = (w_p > p.pfloor) ? u_e : ((p.pfloor/gm1) + e_k + pb);
    } else {
      w_p = gm1*(u_e - e_k);
      u_e = (w_p > p.pfloor) ? u_e : ((p.pfloor/gm1) + e_k);
    }
    w_p = (w_p > p.pfloor) ? w_p : p.pfloor;
  }
  // floors write back into cons (the reference always stores; storing only changes is
  // value-identical and saves two stores per cell)
  if (u_d != u_d0) b.u[o] = u_d;
  if (u_e != u_e0) b.u[o+4*sv] = u_e;
  double vx = u_m1*di, vy = u_m2*di, vz = u_m3*di;
  b.w[o] = u_d;
  b.w[o+sv] = vx;
  b.w[o+2*sv] = vy;
  b.w[o+3*sv] = vz;
  if (!iso) b.w[o+4*sv] = w_p;
  if (MHD && (FLAGS & 1)) {
    if (b.f3) {
      b.cc_e[o] = vz*bcc2 - vy*bcc3;
      b.cc_e[o+sv] = vx*bcc3 - vz*bcc1;
      b.cc_e[o+2*sv] = vy*bcc1 - vx*bcc2;
    } else if (b.f2) {
      b.cc_e[o] = vy*bcc1 - vx*bcc2;
    }
  }
  if ((FLAGS & 2) && i >= b.is && i <= b.ie && j >= b.js && j <= b.je && k >= b.ks &&
      k <= b.ke) {
    double dt1 = b.dx1f[i], dt2 = b.dx2f[j], dt3 = b.dx3f[k];
    if (MHD) {
      double bx = bcc1 + fabs(bf1 - bcc1);
      double cf = iso ? fast_speed_iso(p.iso_cs, u_d, bcc2, bcc3, bx)
                      : fast_speed(p.gamma, u_d, w_p, bcc2, bcc3, bx);
      dt1 /= (fabs(vx) + cf);
      bx = bcc2 + fabs(bf2 - bcc2);
      cf = iso ? fast_speed_iso(p.iso_cs, u_d, bcc3, bcc1, bx)
               : fast_speed(p.gamma, u_d, w_p, bcc3, bcc1, bx);
      dt2 /= (fabs(vy) + cf);
      bx = bcc3 + fabs(bf3 - bcc3);
      cf = iso ? fast_speed_iso(p.iso_cs, u_d, bcc1, bcc2, bx)
               : fast_speed(p.gamma, u_d, w_p, bcc1, bcc2, bx);
      dt3 /= (fabs(vz) + cf);
    } else {
      double cs = iso ? p.iso_cs : sound_speed(p.gamma, u_d, w_p);
      dt1 /= (fabs(vx) + cs);
      dt2 /= (fabs(vy) + cs);
      dt3 /= (fabs(vz) + cs);
    }
    m = dmin(m, dt1);
    if (b.f2) m = dmin(m, dt2);
    if (b.f3) m = dmin(m, dt3);
  }
}

template <bool MHD, int FLAGS>
__global__ void __launch_bounds__(BX, ((FLAGS & 2) ? AB_C2P_MINB : 1)) k_cons2prim(BlkDev b0, Params p, int il, int jl,
                                                  int kl, int ni, int nj, int ntot,
                                                  unsigned long long *dtmin, FastDiv di, FastDiv dj) {
  const BlkDev b = blk_view(b0, blockIdx.y);      // blockIdx.y = local MeshBlock (ab_batch.cuh)
  // the cell range is flattened: rows of nx1+2*NGHOST cells do not pad to the CTA width (a
  // 132-cell row used to occupy two 128-thread CTAs)
  const int t = blockIdx.x*BX + threadIdx.x;
  int r = fast_div(t, di);                    // t / ni (a run-time division is ~20 instructions)
  const int i = il + (t - r*ni);
  const int kk = fast_div(r, dj);
  const int j = jl + (r - kk*nj);
  const int k = kl + kk;
  double m = DBL_MAX;
  if (t < ntot) c2p_cell<MHD,FLAGS>(b, p, i, j, k, m);
  if (FLAGS & 2) block_min_to_slots(m, dtmin + blockIdx.y*DT_SLOTS);
}

// The same over up to six boxes in ONE launch (the ghost shell of a block as six slabs, for the
// schedule that converts the active cells while the ghost zones are still travelling):
// end[q] = inclusive prefix of the boxes' cell counts, a thread finds its box by a linear scan.
struct C2PBoxes { int n; int il[6], jl[6], kl[6], ni[6], nj[6], end[6]; };
template <bool MHD, int FLAGS>
__global__ void __launch_bounds__(BX) k_cons2prim_boxes(BlkDev b0, Params p, C2PBoxes bx) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  int t = blockIdx.x*BX + threadIdx.x;
  if (t >= bx.end[bx.n - 1]) return;
  int q = 0;
  while (t >= bx.end[q]) ++q;
  if (q > 0) t -= bx.end[q - 1];
  const int ni = bx.ni[q], nj = bx.nj[q];
  int r = t / ni;
  const int i = bx.il[q] + (t - r*ni);
  const int kk = r / nj;
  const int j = bx.jl[q] + (r - kk*nj);
  const int k = bx.kl[q] + kk;
  double m = DBL_MAX;
  c2p_cell<MHD,(FLAGS & 1)>(b, p, i, j, k, m);
}

void launch_cons2prim(const BlkDev &b, const Params &p, int il, int iu, int jl, int ju, int kl,
                      int ku, cudaStream_t s, int flags, unsigned long long *dtmin, int nb) {
  const int ni = iu-il+1, nj = ju-jl+1, ntot = ni*nj*(ku-kl+1);
  const dim3 g((unsigned)((ntot + BX - 1)/BX), (unsigned)nb);
  const FastDiv di = make_fastdiv(ni), dj = make_fastdiv(nj);
  if (!p.mhd) flags &= ~1;
  if (p.mhd) {
    switch (flags & 3) {
      case 0: k_cons2prim<true,0><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj); break;
      case 1: k_cons2prim<true,1><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj); break;
      case 2: k_cons2prim<true,2><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj); break;
      default: k_cons2prim<true,3><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj); break;
    }
  } else {
    if (flags & 2) k_cons2prim<false,2><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj);
    else k_cons2prim<false,0><<<g, BX, 0, s>>>(b, p, il, jl, kl, ni, nj, ntot, dtmin, di, dj);
  }
  ++g_launches;
}

void launch_cons2prim_boxes(const BlkDev &b, const Params &p, int nbox, const int (*box)[6],
                            cudaStream_t s, int flags, int nb) {
  C2PBoxes bx;
  memset(&bx, 0, sizeof(bx));
  int tot = 0;
  for (int q = 0; q < nbox; ++q) {
    const int ni = box[q][1] - box[q][0] + 1, nj = box[q][3] - box[q][2] + 1,
              nk = box[q][5] - box[q][4] + 1;
    if (ni <= 0 || nj <= 0 || nk <= 0) continue;
    const int n = bx.n++;
    bx.il[n] = box[q][0]; bx.jl[n] = box[q][2]; bx.kl[n] = box[q][4];
    bx.ni[n] = ni; bx.nj[n] = nj;
    tot += ni*nj*nk;
    bx.end[n] = tot;
  }
  if (bx.n == 0) return;
  const dim3 g((unsigned)((tot + BX - 1)/BX), (unsigned)nb);
  if (p.mhd) {
    if (flags & 1) k_cons2prim_boxes<true,1><<<g, BX, 0, s>>>(b, p, bx);
    else k_cons2prim_boxes<true,0><<<g, BX, 0, s>>>(b, p, bx);
  } else {
    k_cons2prim_boxes<false,0><<<g, BX, 0, s>>>(b, p, bx);
  }
  ++g_launches;
}

// EquationOfState::PrimitiveToConserved (adiabatic_hydro.cpp:89-123, adiabatic_mhd.cpp:99-136)
template <bool MHD>
__global__ void __launch_bounds__(BX) k_prim2cons(BlkDev b, Params p, int il, int iu, int jl,
                                                  int kl) {
  int i = il + blockIdx.x*BX + threadIdx.x;
  if (i > iu) return;
  int j = jl + blockIdx.y, k = kl + blockIdx.z;
  long o = CCI(b,0,k,j,i);
  long sv = (long)b.nc3*b.nc2*b.nc1;
  double w_d = b.w[o], w_vx = b.w[o+sv], w_vy = b.w[o+2*sv], w_vz = b.w[o+3*sv];
  b.u[o] = w_d;
  b.u[o+sv] = w_vx*w_d;
  b.u[o+2*sv] = w_vy*w_d;
  b.u[o+3*sv] = w_vz*w_d;
  if (p.eos != 0) return;        // isothermal: no energy equation
  double igm1 = 1.0/(p.gamma - 1.0);
  double w_p = b.w[o+4*sv];
  if (MHD) {
    double bcc1 = b.bcc[o], bcc2 = b.bcc[o+sv], bcc3 = b.bcc[o+2*sv];
    b.u[o+4*sv] = w_p*igm1 + 0.5*(w_d*(sqr(w_vx) + sqr(w_vy) + sqr(w_vz))
                                  + (sqr(bcc1) + sqr(bcc2) + sqr(bcc3)));
  } else {
    b.u[o+4*sv] = w_p*igm1 + 0.5*w_d*(sqr(w_vx) + sqr(w_vy) + sqr(w_vz));
  }
}

void launch_prim2cons(const BlkDev &b, const Params &p, int il, int iu, int jl, int ju, int kl,
                      int ku, cudaStream_t s) {
  dim3 g = grid3(iu-il+1, ju-jl+1, ku-kl+1);
  if (p.mhd) { k_prim2cons<true><<<g, BX, 0, s>>>(b, p, il, iu, jl, kl); } else { k_prim2cons<false><<<g, BX, 0, s>>>(b, p, il, iu, jl, kl); }
  ++g_launches;
}

__global__ void __launch_bounds__(BX) k_calc_bcc(BlkDev b, int il, int iu, int jl, int kl) {
  int i = il + blockIdx.x*BX + threadIdx.x;
  if (i > iu) return;
  int j = jl + blockIdx.y, k = kl + blockIdx.z;
  double l1 = 0.5, r1 = 0.5, l2 = 0.5, r2 = 0.5, l3 = 0.5, r3 = 0.5;
  if (b.bcw) {   // nonuniform spacing (field.cpp:139-172)
    const double *__restrict__ q1 = b.bcw, *q2 = q1 + 2*b.nc1, *q3 = q2 + 2*b.nc2;
    l1 = q1[i]; r1 = q1[b.nc1+i]; l2 = q2[j]; r2 = q2[b.nc2+j]; l3 = q3[k]; r3 = q3[b.nc3+k];
  }
  b.bcc[CCI(b,0,k,j,i)] = l1*b.b[0][F1I(b,k,j,i)] + r1*b.b[0][F1I(b,k,j,i+1)];
  b.bcc[CCI(b,1,k,j,i)] = l2*b.b[1][F2I(b,k,j,i)] + r2*b.b[1][F2I(b,k,j+1,i)];
  b.bcc[CCI(b,2,k,j,i)] = l3*b.b[2][F3I(b,k,j,i)] + r3*b.b[2][F3I(b,k+1,j,i)];
}

void launch_calc_bcc(const BlkDev &b, int il, int iu, int jl, int ju, int kl, int ku,
                     cudaStream_t s) {
  k_calc_bcc<<<grid3(iu-il+1, ju-jl+1, ku-kl+1), BX, 0, s>>>(b, il, iu, jl, kl); ++g_launches;
}

}  // namespace ab
#include "ab_flux.cuh"
namespace ab {

// uniform spacing: instantiated here; nonuniform (mesh/x?rat != 1) in ab_flux_nu.cu
void launch_flux_dir(const BlkDev &b, const ReconGeom &g, const Params &p, int order, int dir,
                     double dt_val, const double *dt_ptr, cudaStream_t s, int nb) {
  if (g.nu[dir] && order > 1) launch_flux_dir_nu(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
  else launch_flux_dir_t<false>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
}

void launch_fluxes(const BlkDev &b, const ReconGeom &g, const Params &p, int order,
                   double dt_val, const double *dt_ptr, cudaStream_t s) {
  launch_flux_dir(b, g, p, order, 0, dt_val, dt_ptr, s);
  if (b.f2) launch_flux_dir(b, g, p, order, 1, dt_val, dt_ptr, s);
  if (b.f3) launch_flux_dir(b, g, p, order, 2, dt_val, dt_ptr, s);
}

// =============================================================================================
// Passive scalars: PassiveScalars::CalculateFluxes + ComputeUpwindFlux
// (scalars/calculate_scalar_fluxes.cpp:41-382), EquationOfState::PassiveScalar* (eos_scalars.cpp)
// =============================================================================================
// One thread per face and species: reconstruct the concentration r on both sides with the same
// limiter as the hydro variables (dc_simple / plm_simple / ppm_simple.cpp) and upwind it with
// the Riemann solver's mass flux.  Only the faces AddFluxDivergence reads are evaluated (the
// reference also sweeps one more transverse row whose result is never used).  HBM-bound:
// r (2S+2 values, stencil from L1/L2), mass flux in, s_flux out.
template <int DIR, int ORDER>
__global__ void __launch_bounds__(BX) k_scalar_flux(BlkDev b0, ReconGeom g0, double sfloor,
                                                    int ni, int nj, int ntot) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  const ReconGeom g = geom_view(g0, b0, blockIdx.y);
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  const int st = (DIR == 0) ? 1 : ((DIR == 1) ? n1 : n1*n2);
  const int sf = (DIR == 0) ? b.nc3*n2*(n1+1) : ((DIR == 1) ? b.nc3*(n2+1)*n1 : (b.nc3+1)*n2*n1);
  const double *__restrict__ mflx = b.flux[DIR];       // IDN component = first sf entries
  double *__restrict__ out = b.sflux[DIR];
  for (int t = blockIdx.x*BX + threadIdx.x; t < ntot; t += gridDim.x*BX) {
    int r = t / ni;
    const int i = b.is + (t - r*ni);
    const int kk = r / nj;
    const int j = b.js + (r - kk*nj);
    const int k = b.ks + kk;
    const int oc = (k*n2 + j)*n1 + i;
    const int c = (DIR == 0) ? i : ((DIR == 1) ? j : k);
    int of;
    if (DIR == 0) of = (k*n2 + j)*(n1+1) + i;
    else if (DIR == 1) of = (k*(n2+1) + j)*n1 + i;
    else of = oc;
    const double fluid_flx = mflx[of];
    // nonuniform spacing along the sweep (plm_simple.cpp:68-95,148-168,227-243;
    // ppm_simple.cpp:172-183,258-276): geometry rows of the two cells, else nullptr
    const double *tl = (ORDER > 1 && g.nu[DIR]) ? g.nu[DIR] + (c-1)*NUG : nullptr;
    for (int n = 0; n < b.ns; ++n) {
      const double *__restrict__ q = b.r + n*sv + oc;
      double rl, rr, dummy;
      if (ORDER == 1) {
        rl = q[-st]; rr = q[0];
      } else if (ORDER == 2) {
        if (tl) {
          plm_nu<DIR+1>(q[-2*st], q[-st], q[0], g.wp[DIR][c-1], g.wm[DIR][c-1], tl, rl, dummy);
          plm_nu<DIR+1>(q[-st], q[0], q[st], g.wp[DIR][c], g.wm[DIR][c], tl + NUG, dummy, rr);
        } else {
          plm(q[-2*st], q[-st], q[0], g.wp[DIR][c-1], g.wm[DIR][c-1], rl, dummy);
          plm(q[-st], q[0], q[st], g.wp[DIR][c], g.wm[DIR][c], dummy, rr);
        }
      } else {
        if (tl) {
          ppm_nu(q[-3*st], q[-2*st], q[-st], q[0], q[st], tl, rl, dummy);
          ppm_nu(q[-2*st], q[-st], q[0], q[st], q[2*st], tl + NUG, dummy, rr);
        } else {
          ppm(q[-3*st], q[-2*st], q[-st], q[0], q[st], rl, dummy);
          ppm(q[-2*st], q[-st], q[0], q[st], q[2*st], dummy, rr);
        }
        // EquationOfState::ApplyPassiveScalarFloors (eos_scalars.cpp:161-175)
        rl = (rl > sfloor) ? rl : sfloor;
        rr = (rr > sfloor) ? rr : sfloor;
      }
      out[of + n*sf] = (fluid_flx >= 0.0) ? fluid_flx*rl : fluid_flx*rr;
    }
  }
}

void launch_scalar_fluxes(const BlkDev &b, const ReconGeom &g, const Params &p, int order,
                          cudaStream_t s, int nb) {
  if (b.ns <= 0) return;
  const int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1, nx3 = b.ke-b.ks+1;
  for (int dir = 0; dir < 3; ++dir) {
    if ((dir == 1 && !b.f2) || (dir == 2 && !b.f3)) continue;
    const int ni = nx1 + (dir == 0), nj = nx2 + (dir == 1), nk = nx3 + (dir == 2);
    const int ntot = ni*nj*nk;
    const dim3 grid((unsigned)((ntot + BX - 1)/BX), (unsigned)nb);
#define AB_SF(D, O) k_scalar_flux<D,O><<<grid, BX, 0, s>>>(b, g, p.sfloor, ni, nj, ntot)
    if (dir == 0) { if (order == 1) AB_SF(0,1); else if (order == 2) AB_SF(0,2); else AB_SF(0,3); }
    else if (dir == 1) { if (order == 1) AB_SF(1,1); else if (order == 2) AB_SF(1,2); else AB_SF(1,3); }
    else { if (order == 1) AB_SF(2,1); else if (order == 2) AB_SF(2,2); else AB_SF(2,3); }
#undef AB_SF
    ++g_launches;
  }
}

// EquationOfState::PassiveScalarConservedToPrimitive (eos_scalars.cpp:31-60): floor s at
// sfloor*rho (rho = the already floored u(IDN)), r = s/rho; and PassiveScalarPrimitiveToConserved
// (eos_scalars.cpp:133-152) when TO_CONS.
template <bool TO_CONS>
__global__ void __launch_bounds__(BX) k_scalar_eos(BlkDev b0, double sfloor, int il, int jl, int kl,
                                                   int ni, int nj, int ntot) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  for (int t = blockIdx.x*BX + threadIdx.x; t < ntot; t += gridDim.x*BX) {
    int r = t / ni;
    const int i = il + (t - r*ni);
    const int kk = r / nj;
    const int j = jl + (r - kk*nj);
    const int o = ((kl + kk)*n2 + j)*n1 + i;
    const double d = b.u[o];
    for (int n = 0; n < b.ns; ++n) {
      if (TO_CONS) {
        b.s[o + n*sv] = b.r[o + n*sv]*d;
      } else {
        const double s0 = b.s[o + n*sv];
        const double s_n = (s0 < sfloor*d) ? sfloor*d : s0;
        if (s_n != s0) b.s[o + n*sv] = s_n;
        b.r[o + n*sv] = s_n/d;
      }
    }
  }
}

void launch_scalar_eos(const BlkDev &b, const Params &p, int to_cons, int il, int iu, int jl,
                       int ju, int kl, int ku, cudaStream_t s, int nb) {
  if (b.ns <= 0) return;
  const int ni = iu-il+1, nj = ju-jl+1, nk = ku-kl+1;
  const int ntot = ni*nj*nk;
  if (ntot <= 0) return;
  const dim3 grid((unsigned)((ntot + BX - 1)/BX), (unsigned)nb);
  if (to_cons) k_scalar_eos<true><<<grid, BX, 0, s>>>(b, p.sfloor, il, jl, kl, ni, nj, ntot);
  else k_scalar_eos<false><<<grid, BX, 0, s>>>(b, p.sfloor, il, jl, kl, ni, nj, ntot);
  ++g_launches;
}

// =============================================================================================
// Field::ComputeCornerE (field/calculate_corner_e.cpp:28-236)
// =============================================================================================
__global__ void __launch_bounds__(BX) k_cc_e(BlkDev b, int i0, int i1, int j0, int k0) {
  int i = i0 + blockIdx.x*BX + threadIdx.x;
  if (i > i1) return;
  int j = j0 + blockIdx.y, k = k0 + blockIdx.z;
  long o = CCI(b,0,k,j,i);
  long sv = (long)b.nc3*b.nc2*b.nc1;
  double vx = b.w[o+sv], vy = b.w[o+2*sv], vz = b.w[o+3*sv];
  double b1 = b.bcc[o], b2 = b.bcc[o+sv], b3 = b.bcc[o+2*sv];
  if (b.f3) {
    b.cc_e[o] = vz*b2 - vy*b3;
    b.cc_e[o+sv] = vx*b3 - vz*b1;
    b.cc_e[o+2*sv] = vy*b1 - vx*b2;
  } else {
    b.cc_e[o] = vy*b1 - vx*b2;   // 2-D: single component (E3), stored in slot 0
  }
}

// upwinded gradient term of the corner EMF (GS05/SG07), calculate_corner_e.cpp:190-196
__device__ __forceinline__ double de_term(double wt, double ef_a, double cc_a, double ef_b,
                                          double cc_b) {
  return (1.0-wt)*(ef_a - cc_a) + (wt)*(ef_b - cc_b);
}

__global__ void __launch_bounds__(BX, AB_CE_MINB) k_corner_e3d(BlkDev b0, int ni, int nj, int ntot, FastDiv di, FastDiv dj) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  int t = blockIdx.x*BX + threadIdx.x;
  if (t >= ntot) return;
  int r = fast_div(t, di);
  const int i = b.is + (t - r*ni);
  const int kk = fast_div(r, dj);
  const int j = b.js + (r - kk*nj);
  const int k = b.ks + kk;
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  // element offsets of (k,j,i) in the cell-, x1f-, x2f-, x3f-shaped arrays and their strides
  const int oc = (k*n2 + j)*n1 + i,      cj = n1,   ck = n1*n2;        // cell / x3f shape
  const int o1 = (k*n2 + j)*(n1+1) + i,  j1 = n1+1, k1 = (n1+1)*n2;    // x1f shape
  const int o2 = (k*(n2+1) + j)*n1 + i,  j2 = n1,   k2 = n1*(n2+1);    // x2f shape
  const double *__restrict__ cc1 = b.cc_e;
  const double *__restrict__ cc2 = b.cc_e + sv;
  const double *__restrict__ cc3 = b.cc_e + 2*sv;
  const double *__restrict__ w_x1f = b.wght[0];
  const double *__restrict__ w_x2f = b.wght[1];
  const double *__restrict__ w_x3f = b.wght[2];
  const double *__restrict__ e3_x1f = b.ef[0][0];
  const double *__restrict__ e2_x1f = b.ef[0][1];
  const double *__restrict__ e1_x2f = b.ef[1][0];
  const double *__restrict__ e3_x2f = b.ef[1][1];
  const double *__restrict__ e2_x3f = b.ef[2][0];
  const double *__restrict__ e1_x3f = b.ef[2][1];
  {
    const double f3c = e1_x3f[oc], f3m = e1_x3f[oc-cj];       // e1_x3f(k,j,i), (k,j-1,i)
    const double f2c = e1_x2f[o2], f2m = e1_x2f[o2-k2];       // e1_x2f(k,j,i), (k-1,j,i)
    double de1_l3 = de_term(w_x2f[o2-k2], f3c, cc1[oc-ck], f3m, cc1[oc-ck-cj]);
    double de1_r3 = de_term(w_x2f[o2], f3c, cc1[oc], f3m, cc1[oc-cj]);
    double de1_l2 = de_term(w_x3f[oc-cj], f2c, cc1[oc-cj], f2m, cc1[oc-ck-cj]);
    double de1_r2 = de_term(w_x3f[oc], f2c, cc1[oc], f2m, cc1[oc-ck]);
    b.e[0][(k*(n2+1) + j)*n1 + i] =
        0.25*(de1_l3 + de1_r3 + de1_l2 + de1_r2 + f2m + f2c + f3m + f3c);
  }
  {
    const double f3c = e2_x3f[oc], f3m = e2_x3f[oc-1];        // e2_x3f(k,j,i), (k,j,i-1)
    const double f1c = e2_x1f[o1], f1m = e2_x1f[o1-k1];       // e2_x1f(k,j,i), (k-1,j,i)
    double de2_l3 = de_term(w_x1f[o1-k1], f3c, cc2[oc-ck], f3m, cc2[oc-ck-1]);
    double de2_r3 = de_term(w_x1f[o1], f3c, cc2[oc], f3m, cc2[oc-1]);
    double de2_l1 = de_term(w_x3f[oc-1], f1c, cc2[oc-1], f1m, cc2[oc-ck-1]);
    double de2_r1 = de_term(w_x3f[oc], f1c, cc2[oc], f1m, cc2[oc-ck]);
    b.e[1][(k*n2 + j)*(n1+1) + i] =
        0.25*(de2_l3 + de2_r3 + de2_l1 + de2_r1 + f3m + f3c + f1m + f1c);
  }
  {
    const double f2c = e3_x2f[o2], f2m = e3_x2f[o2-1];        // e3_x2f(k,j,i), (k,j,i-1)
    const double f1c = e3_x1f[o1], f1m = e3_x1f[o1-j1];       // e3_x1f(k,j,i), (k,j-1,i)
    double de3_l2 = de_term(w_x1f[o1-j1], f2c, cc3[oc-cj], f2m, cc3[oc-cj-1]);
    double de3_r2 = de_term(w_x1f[o1], f2c, cc3[oc], f2m, cc3[oc-1]);
    double de3_l1 = de_term(w_x2f[o2-1], f1c, cc3[oc-1], f1m, cc3[oc-cj-1]);
    double de3_r1 = de_term(w_x2f[o2], f1c, cc3[oc], f1m, cc3[oc-cj]);
    b.e[2][(k*(n2+1) + j)*(n1+1) + i] =
        0.25*(de3_l1 + de3_r1 + de3_l2 + de3_r2 + f2m + f2c + f1m + f1c);
  }
  (void)j2;
}

// 2-D (calculate_corner_e.cpp:50-128): grid.y covers j in [js, je+1]; k = ks
__global__ void __launch_bounds__(BX) k_corner_e2d(BlkDev b0) {
  const BlkDev b = blk_view(b0, blockIdx.z);      // grid.y counts rows here: blocks along z
  int i = b.is + blockIdx.x*BX + threadIdx.x;
  if (i > b.ie+1) return;
  int j = b.js + blockIdx.y, k = b.ks;
  const double *cc = b.cc_e;
  const double *w_x1f = b.wght[0], *w_x2f = b.wght[1];
  const double *e3_x1f = b.ef[0][0], *e2_x1f = b.ef[0][1];
  const double *e1_x2f = b.ef[1][0], *e3_x2f = b.ef[1][1];
  if (j <= b.je) {
    double v = e2_x1f[F1I(b,k,j,i)];
    b.e[1][E2I(b,b.ke+1,j,i)] = v;
    b.e[1][E2I(b,k,j,i)] = v;
  }
  if (i <= b.ie) {
    double v = e1_x2f[F2I(b,k,j,i)];
    b.e[0][E1I(b,b.ke+1,j,i)] = v;
    b.e[0][E1I(b,k,j,i)] = v;
  }
#define C3(k,j,i) CCI(b,0,k,j,i)
  double de3_l2 = de_term(w_x1f[F1I(b,k,j-1,i)], e3_x2f[F2I(b,k,j,i)], cc[C3(k,j-1,i)],
                          e3_x2f[F2I(b,k,j,i-1)], cc[C3(k,j-1,i-1)]);
  double de3_r2 = de_term(w_x1f[F1I(b,k,j,i)], e3_x2f[F2I(b,k,j,i)], cc[C3(k,j,i)],
                          e3_x2f[F2I(b,k,j,i-1)], cc[C3(k,j,i-1)]);
  double de3_l1 = de_term(w_x2f[F2I(b,k,j,i-1)], e3_x1f[F1I(b,k,j,i)], cc[C3(k,j,i-1)],
                          e3_x1f[F1I(b,k,j-1,i)], cc[C3(k,j-1,i-1)]);
  double de3_r1 = de_term(w_x2f[F2I(b,k,j,i)], e3_x1f[F1I(b,k,j,i)], cc[C3(k,j,i)],
                          e3_x1f[F1I(b,k,j-1,i)], cc[C3(k,j-1,i)]);
  b.e[2][E3I(b,k,j,i)] = 0.25*(de3_l1 + de3_r1 + de3_l2 + de3_r2 + e3_x2f[F2I(b,k,j,i-1)] +
                               e3_x2f[F2I(b,k,j,i)] + e3_x1f[F1I(b,k,j-1,i)] + e3_x1f[F1I(b,k,j,i)]);
#undef C3
}

// 1-D (calculate_corner_e.cpp:38-48)
__global__ void __launch_bounds__(BX) k_corner_e1d(BlkDev b0) {
  const BlkDev b = blk_view(b0, blockIdx.z);
  int i = b.is + blockIdx.x*BX + threadIdx.x;
  if (i > b.ie+1) return;
  int ks = b.ks, js = b.js;
  double v2 = b.ef[0][1][F1I(b,ks,js,i)], v3 = b.ef[0][0][F1I(b,ks,js,i)];
  b.e[1][E2I(b,ks,js,i)] = v2;
  b.e[1][E2I(b,b.ke+1,js,i)] = v2;
  b.e[2][E3I(b,ks,js,i)] = v3;
  b.e[2][E3I(b,ks,b.je+1,i)] = v3;
}

// nb > 1 (one launch over nb MeshBlocks) requires have_cc_e: k_cc_e has no free grid dimension
void launch_corner_e(const BlkDev &b, cudaStream_t s, int have_cc_e, int nb) {
  int nx1 = b.ie - b.is + 1, nx2 = b.je - b.js + 1, nx3 = b.ke - b.ks + 1;
  if (!b.f2) {
    k_corner_e1d<<<grid3(nx1+1, 1, nb), BX, 0, s>>>(b); ++g_launches;
  } else if (!b.f3) {
    if (!have_cc_e) { k_cc_e<<<grid3(nx1+2, nx2+2, 1), BX, 0, s>>>(b, b.is-1, b.ie+1, b.js-1, b.ks); ++g_launches; }
    k_corner_e2d<<<grid3(nx1+1, nx2+1, nb), BX, 0, s>>>(b); ++g_launches;
  } else {
    if (!have_cc_e) { k_cc_e<<<grid3(nx1+2, nx2+2, nx3+2), BX, 0, s>>>(b, b.is-1, b.ie+1, b.js-1, b.ks-1); ++g_launches; }
    const int ntot = (nx1+1)*(nx2+1)*(nx3+1);
    k_corner_e3d<<<dim3((unsigned)((ntot + BX - 1)/BX), (unsigned)nb), BX, 0, s>>>(b, nx1+1, nx2+1, ntot, make_fastdiv(nx1+1), make_fastdiv(nx2+1)); ++g_launches;
  }
}

// =============================================================================================
// EMF boundary consistency (bvals/fc/flux_correction_fc.cpp): pack what every face / edge
// neighbour needs (LoadFluxBoundaryBufferSameLevel :51-307), then add the neighbours'
// contributions in neighbour-list order and scale (SetFluxBoundarySameLevel :689-905,
// AverageFluxBoundary :1355-1541).
// =============================================================================================

// One "slab" = the set of edge-EMF elements with one index pinned to a block face.
// comp 0/1/2 = e1/e2/e3.  (a, b) run over the two free indices.
struct SlabIdx { int k, j, i; };

// number of elements and (k,j,i) of element t of the buffer section `sec` (0 or 1) that the
// reference packs for face neighbour `fid` (3-D / 2-D / 1-D layouts)
__device__ __forceinline__ long face_sec_count(const BlkDev &b, int fid, int sec) {
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1, nx3 = b.ke-b.ks+1;
  if (b.f3) {
    if (fid < 2) return sec == 0 ? (long)(nx3+1)*nx2 : (long)nx3*(nx2+1);   // e2 | e3
    if (fid < 4) return sec == 0 ? (long)(nx3+1)*nx1 : (long)nx3*(nx1+1);   // e1 | e3
    return sec == 0 ? (long)(nx2+1)*nx1 : (long)nx2*(nx1+1);                // e1 | e2
  } else if (b.f2) {
    if (fid < 2) return sec == 0 ? nx2 : nx2+1;                             // e2 | e3
    return sec == 0 ? nx1 : nx1+1;                                          // e1 | e3
  }
  return 1;                                                                  // e2 | e3
}
__device__ __forceinline__ int face_sec_comp(const BlkDev &b, int fid, int sec) {
  if (fid < 2) return sec == 0 ? 1 : 2;
  if (fid < 4) return sec == 0 ? 0 : 2;
  return sec == 0 ? 0 : 1;
}
__device__ __forceinline__ SlabIdx face_sec_index(const BlkDev &b, int fid, int sec, long t) {
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1;
  SlabIdx x;
  if (fid < 2) {
    x.i = (fid == 0) ? b.is : b.ie+1;
    int nj = (sec == 0) ? nx2 : nx2+1;
    if (!b.f2) nj = 1;
    x.k = b.ks + (int)(t / nj); x.j = b.js + (int)(t % nj);
  } else if (fid < 4) {
    x.j = (fid == 2) ? b.js : b.je+1;
    int ni = (sec == 0) ? nx1 : nx1+1;
    x.k = b.ks + (int)(t / ni); x.i = b.is + (int)(t % ni);
  } else {
    x.k = (fid == 4) ? b.ks : b.ke+1;
    int ni = (sec == 0) ? nx1 : nx1+1;
    x.j = b.js + (int)(t / ni); x.i = b.is + (int)(t % ni);
  }
  return x;
}
__device__ __forceinline__ long e_index(const BlkDev &b, int comp, int k, int j, int i) {
  return comp == 0 ? E1I(b,k,j,i) : (comp == 1 ? E2I(b,k,j,i) : E3I(b,k,j,i));
}
__device__ __forceinline__ long edge_count(const BlkDev &b, int eid) {
  if (eid < 4) return b.ke-b.ks+1;
  if (eid < 8) return b.je-b.js+1;
  return b.ie-b.is+1;
}
__device__ __forceinline__ SlabIdx edge_index(const BlkDev &b, int eid, long t) {
  SlabIdx x;
  if (eid < 4) {
    x.i = ((eid & 1) == 0) ? b.is : b.ie+1; x.j = ((eid & 2) == 0) ? b.js : b.je+1;
    x.k = b.ks + (int)t;
  } else if (eid < 8) {
    x.i = ((eid & 1) == 0) ? b.is : b.ie+1; x.k = ((eid & 2) == 0) ? b.ks : b.ke+1;
    x.j = b.js + (int)t;
  } else {
    x.j = ((eid & 1) == 0) ? b.js : b.je+1; x.k = ((eid & 2) == 0) ? b.ks : b.ke+1;
    x.i = b.is + (int)t;
  }
  return x;
}

// grid.y = 0..5 faces, 6..17 edges
// grid.z = local MeshBlocks; plans = device array of their plans (a plan passed by value would
// have to be copied to local memory by every thread to be indexed with a run-time face / edge id)
// The edge-EMF arrays are picked with a run-time component index: `b` stays the kernel parameter
// (indexable in the constant bank; a shifted copy would live in local memory) and the block
// offset `off` is applied to the one pointer that is dereferenced.
__global__ void __launch_bounds__(256) k_emf_pack(BlkDev b, const EmfPlan *__restrict__ plans) {
  const long off = (long)blockIdx.z*b.bstride;
  const EmfPlan &pl = plans[blockIdx.z];
  int id = blockIdx.y;
  long t = (long)blockIdx.x*256 + threadIdx.x;
  if (id < 6) {
    double *dst = pl.face_dst[id];
    if (!dst) return;
    long n0 = face_sec_count(b, id, 0), n1 = face_sec_count(b, id, 1);
    if (t >= n0 + n1) return;
    int sec = t < n0 ? 0 : 1;
    long tt = sec ? t - n0 : t;
    SlabIdx x = face_sec_index(b, id, sec, tt);
    int comp = face_sec_comp(b, id, sec);
    dst[t] = blk_mv(b.e[comp], off)[e_index(b, comp, x.k, x.j, x.i)];
  } else {
    int eid = id - 6;
    double *dst = pl.edge_dst[eid];
    if (!dst) return;
    if (t >= edge_count(b, eid)) return;
    SlabIdx x = edge_index(b, eid, t);
    int comp = eid < 4 ? 2 : (eid < 8 ? 1 : 0);
    dst[t] = blk_mv(b.e[comp], off)[e_index(b, comp, x.k, x.j, x.i)];
  }
}

void launch_emf_pack(const BlkDev &b, const EmfPlan *plans_dev, cudaStream_t s, int nb) {
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1, nx3 = b.ke-b.ks+1;
  long m = 2L*(nx1+1)*(nx2+1);
  if (2L*(nx1+1)*(nx3+1) > m) m = 2L*(nx1+1)*(nx3+1);
  if (2L*(nx2+1)*(nx3+1) > m) m = 2L*(nx2+1)*(nx3+1);
  int nid = b.f3 ? 18 : (b.f2 ? 10 : 2);
  k_emf_pack<<<dim3((unsigned)((m + 255)/256), nid, nb), 256, 0, s>>>(b, plans_dev); ++g_launches;
}

// offset of element (k,j,i) of component comp inside the buffer packed by the neighbour
// across face `fid` (the neighbour packed its OPPOSITE face with the same (free-index) layout)
__device__ __forceinline__ long face_buf_offset(const BlkDev &b, int fid, int comp, int k,
                                                int j, int i) {
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1;
  int sec = (face_sec_comp(b, fid, 0) == comp) ? 0 : 1;
  long base = sec ? face_sec_count(b, fid, 0) : 0;
  if (fid < 2) {
    int nj = (sec == 0) ? nx2 : nx2+1;
    if (!b.f2) nj = 1;
    return base + (long)(k - b.ks)*nj + (j - b.js);
  } else if (fid < 4) {
    int ni = (sec == 0) ? nx1 : nx1+1;
    return base + (long)(k - b.ks)*ni + (i - b.is);
  }
  int ni = (sec == 0) ? nx1 : nx1+1;
  return base + (long)(j - b.js)*ni + (i - b.is);
}

// corrected value of one boundary edge-EMF element: own value + neighbours' (faces in
// neighbour order x1,x2,x3, then the edge neighbour) then the averaging factor.
// lo/hi flags: -1 / +1 when the element sits on the inner / outer block face of a direction.
__device__ __forceinline__ double emf_corrected(const BlkDev &b, long off, const EmfPlan &pl,
                                                int comp, int k, int j, int i) {
  double v = blk_mv(b.e[comp], off)[e_index(b, comp, k, j, i)];
  // which block faces does this element touch?  (e1 lives on x2/x3 faces, e2 on x1/x3, e3 on x1/x2)
  int s1 = 0, s2 = 0, s3 = 0;
  if (comp != 0) s1 = (i == b.is) ? -1 : ((i == b.ie+1) ? 1 : 0);
  if (comp != 1 && b.f2) s2 = (j == b.js) ? -1 : ((j == b.je+1) ? 1 : 0);
  if (comp != 2 && b.f3) s3 = (k == b.ks) ? -1 : ((k == b.ke+1) ? 1 : 0);
  // 2-D / 1-D: e1(k+1), e2(k+1), e3(j+1) duplicates receive the k (j) = start value's buffer
  int kk = k, jj = j;
  if (!b.f3 && comp != 2) kk = b.ks;
  if (!b.f2 && comp == 2) jj = b.js;
  int nface = 0;
  int f1 = -1, f2 = -1, f3 = -1;
  if (s1) { f1 = (s1 < 0) ? 0 : 1; nface++; }
  if (s2) { f2 = (s2 < 0) ? 2 : 3; nface++; }
  if (s3) { f3 = (s3 < 0) ? 4 : 5; nface++; }
  if (nface == 0) return v;
  if (f1 >= 0 && pl.face_src[f1]) v += pl.face_src[f1][face_buf_offset(b, f1, comp, kk, jj, i)];
  if (f2 >= 0 && pl.face_src[f2]) v += pl.face_src[f2][face_buf_offset(b, f2, comp, kk, jj, i)];
  if (f3 >= 0 && pl.face_src[f3]) v += pl.face_src[f3][face_buf_offset(b, f3, comp, kk, jj, i)];
  if (nface == 2) {
    int eid;
    long t;
    if (comp == 2) { eid = ((s1 > 0) ? 1 : 0) | ((s2 > 0) ? 2 : 0); t = k - b.ks; }
    else if (comp == 1) { eid = 4 + (((s1 > 0) ? 1 : 0) | ((s3 > 0) ? 2 : 0)); t = j - b.js; }
    else { eid = 8 + (((s2 > 0) ? 1 : 0) | ((s3 > 0) ? 2 : 0)); t = i - b.is; }
    if (pl.edge_src[eid]) v += pl.edge_src[eid][t];
    if (pl.nedge_fine[eid] != 1) v *= 1.0/(double)pl.nedge_fine[eid];
  } else {
    int f = f1 >= 0 ? f1 : (f2 >= 0 ? f2 : f3);
    if (pl.face_avg[f]) v *= 0.5;
  }
  return v;
}

// One thread per boundary element; slabs enumerated without duplicates:
// for each component, the two faces of its first bounding direction take their full extent,
// the faces of its second bounding direction exclude the lines already covered.
__global__ void __launch_bounds__(256) k_emf_apply(BlkDev b, const EmfPlan *__restrict__ plans) {
  const long off = (long)blockIdx.z*b.bstride;
  const EmfPlan &pl = plans[blockIdx.z];
  int slab = blockIdx.y;            // comp*4 + {0,1: first dir lo/hi ; 2,3: second dir lo/hi}
  int comp = slab >> 2, which = slab & 3;
  long t = (long)blockIdx.x*256 + threadIdx.x;
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1, nx3 = b.ke-b.ks+1;
  int k, j, i;
  // extents of the component's index space (calculate_corner_e ranges that CT reads)
  // e1: k in [ks,ke+1], j in [js,je+1], i in [is,ie]
  // e2: k in [ks,ke+1], j in [js,je],   i in [is,ie+1]
  // e3: k in [ks,ke],   j in [js,je+1], i in [is,ie+1]
  int nk = (comp == 2) ? nx3 : nx3+1, nj = (comp == 1) ? nx2 : nx2+1, ni = (comp == 0) ? nx1 : nx1+1;
  if (comp == 0) {            // e1: first dir x2 (j pinned), second dir x3 (k pinned)
    if (which < 2) {
      if (!b.f2) return;
      if (t >= (long)nk*ni) return;
      j = which == 0 ? b.js : b.je+1; k = b.ks + (int)(t/ni); i = b.is + (int)(t%ni);
    } else {
      if (!b.f3) return;
      if (t >= (long)(nj-2)*ni) return;
      k = which == 2 ? b.ks : b.ke+1; j = b.js+1 + (int)(t/ni); i = b.is + (int)(t%ni);
    }
  } else if (comp == 1) {     // e2: first dir x1 (i pinned), second dir x3 (k pinned)
    if (which < 2) {
      if (t >= (long)nk*nj) return;
      i = which == 0 ? b.is : b.ie+1; k = b.ks + (int)(t/nj); j = b.js + (int)(t%nj);
    } else {
      if (!b.f3) return;
      if (t >= (long)nj*(ni-2)) return;
      k = which == 2 ? b.ks : b.ke+1; j = b.js + (int)(t/(ni-2)); i = b.is+1 + (int)(t%(ni-2));
    }
  } else {                    // e3: first dir x1 (i pinned), second dir x2 (j pinned)
    if (which < 2) {
      if (t >= (long)nk*nj) return;
      i = which == 0 ? b.is : b.ie+1; k = b.ks + (int)(t/nj); j = b.js + (int)(t%nj);
    } else {
      if (!b.f2) return;
      if (t >= (long)nk*(ni-2)) return;
      j = which == 2 ? b.js : b.je+1; k = b.ks + (int)(t/(ni-2)); i = b.is+1 + (int)(t%(ni-2));
    }
  }
  double v = emf_corrected(b, off, pl, comp, k, j, i);
  blk_mv(b.e[comp], off)[e_index(b, comp, k, j, i)] = v;
}

void launch_emf_apply(const BlkDev &b, const EmfPlan *plans_dev, cudaStream_t s, int nb) {
  int nx1 = b.ie-b.is+1, nx2 = b.je-b.js+1, nx3 = b.ke-b.ks+1;
  long m = (long)(nx1+1)*(nx2+1);
  if ((long)(nx1+1)*(nx3+1) > m) m = (long)(nx1+1)*(nx3+1);
  if ((long)(nx2+1)*(nx3+1) > m) m = (long)(nx2+1)*(nx3+1);
  k_emf_apply<<<dim3((unsigned)((m + 255)/256), 12, nb), 256, 0, s>>>(b, plans_dev); ++g_launches;
}

// =============================================================================================
// MeshBlock::WeightedAve (mesh/weighted_ave.cpp:33-236 / 238-...) restricted to two registers
// =============================================================================================
__device__ __forceinline__ double wave2(double out, double in, double w0, double w1) {
  if (w0 == 1.0) {
    if (w1 != 0.0) out += w1*in;
  } else if (w0 == 0.0) {
    if (w1 == 1.0) out = in; else out = w1*in;
  } else {
    if (w1 != 0.0) out = w0*out + w1*in; else out *= w0;
  }
  return out;
}

__global__ void __launch_bounds__(BX) k_wave_cc(BlkDev b, double *out, const double *in,
                                                double w0, double w1, int nvar) {
  int i = b.is + blockIdx.x*BX + threadIdx.x;
  if (i > b.ie) return;
  int j = b.js + blockIdx.y, k = b.ks + blockIdx.z;
  long sv = (long)b.nc3*b.nc2*b.nc1;
  long o = CCI(b,0,k,j,i);
  for (int n = 0; n < nvar; ++n) out[o+n*sv] = wave2(out[o+n*sv], in[o+n*sv], w0, w1);
}

void launch_weighted_ave_cc(const BlkDev &b, double *out, const double *in, double w0,
                            double w1, cudaStream_t s, int nvar) {
  k_wave_cc<<<grid3(b.ie-b.is+1, b.je-b.js+1, b.ke-b.ks+1), BX, 0, s>>>(b, out, in, w0, w1,
                                                                        nvar); ++g_launches;
}

struct FcPtrs { double *o[3]; const double *i[3]; };

__global__ void __launch_bounds__(BX) k_wave_fc(BlkDev b, FcPtrs f, double w0, double w1) {
  int i = b.is + blockIdx.x*BX + threadIdx.x;
  if (i > b.ie+1) return;
  int j = b.js + blockIdx.y, k = b.ks + blockIdx.z;
  if (j <= b.je && k <= b.ke) {
    long o = F1I(b,k,j,i); f.o[0][o] = wave2(f.o[0][o], f.i[0][o], w0, w1);
  }
  if (i <= b.ie && k <= b.ke) {
    long o = F2I(b,k,j,i); f.o[1][o] = wave2(f.o[1][o], f.i[1][o], w0, w1);
  }
  if (i <= b.ie && j <= b.je) {
    long o = F3I(b,k,j,i); f.o[2][o] = wave2(f.o[2][o], f.i[2][o], w0, w1);
  }
}

void launch_weighted_ave_fc(const BlkDev &b, double *const out[3], double *const in[3],
                            double w0, double w1, cudaStream_t s) {
  FcPtrs f;
  for (int d = 0; d < 3; ++d) { f.o[d] = out[d]; f.i[d] = in[d]; }
  k_wave_fc<<<grid3(b.ie-b.is+2, b.je-b.js+2, b.ke-b.ks+2), BX, 0, s>>>(b, f, w0, w1); ++g_launches;
}

// =============================================================================================
// IntegrateHydro: register average + Hydro::AddFluxDivergence, fused
// =============================================================================================
// Flattened over the active cells of the planes [k0, k0+nk).
// `c` selects the variable set: (u, u1, flux) with NVAR = NHYDRO for IntegrateHydro, or
// (s, s1, s_flux) with NVAR = 0 -> c.nvar scalars for IntegrateScalars
// (time_integrator.cpp:2141-2185, scalars/add_scalar_flux_divergence.cpp:43-97).
// SRC: HydroSourceTerms::ConstantAcceleration (hydro/srcterms/constant_acc.cpp:25-77) rides in
// the same pass (SRC_TERM follows INT_HYD with dt = beta*dt and the stage-start primitives,
// time_integrator.cpp:1655-1678): u(IM1+d) += src_d, u(IEN) += src_1*vx, then src_2*vy, src_3*vz.
struct CcSet { double *u, *u1; const double *f[3]; int nvar; double g[3]; };

// (Resolving mode / zero_init / dimensionality at compile time was measured: the specialised 3-D
// kernels take 64 instead of 56 registers and run 5.55 instead of 5.28 ms at 512^3 -- reverted.)
template <int NVAR, bool SRC>
__global__ void __launch_bounds__(BX, AB_CC_MINB) k_integrate_cc(BlkDev b0, CcSet c0, int mode, int zero_init,
                                                     double delta, double g1, double g2,
                                                     double beta, double dt_val,
                                                     const double *dt_ptr, int k0, int ni,
                                                     int nj, int ntot, FastDiv di, FastDiv dj) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  CcSet c = c0;
  {
    const long off = (long)blockIdx.y*b0.bstride;
    c.u = blk_mv(c.u, off); c.u1 = blk_mv(c.u1, off);
#pragma unroll
    for (int d = 0; d < 3; ++d) c.f[d] = blk_mv(c.f[d], off);
  }
  const double wght = beta*(dt_ptr ? *dt_ptr : dt_val);
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  const int s1 = b.nc3*n2*(n1+1), s2 = b.nc3*(n2+1)*n1, s3 = (b.nc3+1)*n2*n1;
  {
    const int t = blockIdx.x*BX + threadIdx.x;
    if (t >= ntot) return;
    int r = fast_div(t, di);
    const int i = b.is + (t - r*ni);
    const int kk = fast_div(r, dj);
    const int j = b.js + (r - kk*nj);
    const int k = k0 + kk;
    const int o = (k*n2 + j)*n1 + i;
    const int o1 = (k*n2 + j)*(n1+1) + i, o2 = (k*(n2+1) + j)*n1 + i, o3 = o;
    // Cartesian face areas / volume (coordinates/coordinates.cpp:436-533)
    const double dx1 = b.dx1f[i], dx2 = b.dx2f[j], dx3 = b.dx3f[k];
    const double a1 = dx2*dx3, a2 = dx1*dx3, a3 = dx1*dx2;
    const double vol = dx1*dx2*dx3;
    const int nvar = NVAR > 0 ? NVAR : c.nvar;
#pragma unroll
    for (int n = 0; n < nvar; ++n) {
      double uo;
      if (mode == 0) {
        uo = c.u[o+n*sv];
      } else if (mode == 1) {
        uo = zero_init ? 0.0 : c.u[o+n*sv];
        if (delta != 0.0) uo += delta*c.u1[o+n*sv];
      } else {
        double u1v = zero_init ? 0.0 : c.u1[o+n*sv];
        double un = c.u[o+n*sv];
        if (delta != 0.0 || zero_init) {
          if (delta != 0.0) u1v += delta*un;
          c.u1[o+n*sv] = u1v;
        }
        uo = wave2(un, u1v, g1, g2);
      }
      double dflx = (a1*c.f[0][o1+1+n*s1] - a1*c.f[0][o1+n*s1]);
      if (b.f2) dflx += (a2*c.f[1][o2+n1+n*s2] - a2*c.f[1][o2+n*s2]);
      if (b.f3) dflx += (a3*c.f[2][o3+n1*n2+n*s3] - a3*c.f[2][o3+n*s3]);
      double v = uo - wght*dflx/vol;
      if (SRC) {
        const double w_d = b.w[o];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (c.g[d] == 0.0) continue;
          const double src = wght*w_d*c.g[d];
          if (n == IM1 + d) v += src;
          if (n == IEN) v += src*b.w[o + (IVX + d)*sv];
        }
      }
      c.u[o+n*sv] = v;
    }
  }
}

void launch_integrate_cc(const BlkDev &b, int mode, int zero_init, double delta, double g1,
                         double g2, double beta, double dt_val, const double *dt_ptr,
                         cudaStream_t s, int kl, int ku, int grid, int scalars,
                         const double *gacc, int nb) {
  if (kl < 0) { kl = b.ks; ku = b.ke; }
  const int ni = b.ie-b.is+1, nj = b.je-b.js+1, nk = ku-kl+1;
  const int ntot = ni*nj*nk;
  const dim3 g((unsigned)((ntot + BX - 1)/BX), (unsigned)nb);
  const FastDiv di = make_fastdiv(ni), dj = make_fastdiv(nj);
  (void)grid;
  CcSet c;
  if (scalars) {
    c.u = b.s; c.u1 = b.s1; c.nvar = b.ns;
    for (int d = 0; d < 3; ++d) c.f[d] = b.sflux[d];
    for (int d = 0; d < 3; ++d) c.g[d] = 0.0;
    k_integrate_cc<0,false><<<g, BX, 0, s>>>(b, c, mode, zero_init, delta, g1, g2, beta, dt_val,
                                               dt_ptr, kl, ni, nj, ntot, di, dj);
  } else {
    c.u = b.u; c.u1 = b.u1; c.nvar = b.nh;
    bool src = false;
    for (int d = 0; d < 3; ++d) {
      c.f[d] = b.flux[d];
      c.g[d] = gacc ? gacc[d] : 0.0;
      src = src || (c.g[d] != 0.0);
    }
#define AB_ICC(NV, SR) k_integrate_cc<NV,SR><<<g, BX, 0, s>>>(b, c, mode, zero_init, delta, g1, g2, \
                                                             beta, dt_val, dt_ptr, kl, ni, nj, ntot, di, dj)
    if (b.nh == NHYDRO) { if (src) AB_ICC(NHYDRO, true); else AB_ICC(NHYDRO, false); }
    else { if (src) AB_ICC(4, true); else AB_ICC(4, false); }
#undef AB_ICC
  }
  ++g_launches;
}

// HydroSourceTerms::ConstantAcceleration as its own pass (task-level entry point)
struct Acc3 { double g[3]; };
__global__ void __launch_bounds__(BX) k_const_accel(BlkDev b, Acc3 a, double dt, int ni, int nj,
                                                    int ntot) {
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  for (int t = blockIdx.x*BX + threadIdx.x; t < ntot; t += gridDim.x*BX) {
    int r = t / ni;
    const int i = b.is + (t - r*ni);
    const int kk = r / nj;
    const int j = b.js + (r - kk*nj);
    const int o = ((b.ks + kk)*n2 + j)*n1 + i;
    const double w_d = b.w[o];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (a.g[d] == 0.0) continue;
      const double src = dt*w_d*a.g[d];
      b.u[o + (IM1 + d)*sv] += src;
      if (b.nh == NHYDRO) b.u[o + IEN*sv] += src*b.w[o + (IVX + d)*sv];
    }
  }
}

void launch_const_accel(const BlkDev &b, const double *g, double dt, cudaStream_t s) {
  const int ni = b.ie-b.is+1, nj = b.je-b.js+1, nk = b.ke-b.ks+1;
  const int ntot = ni*nj*nk;
  Acc3 a;
  for (int d = 0; d < 3; ++d) a.g[d] = g[d];
  k_const_accel<<<(ntot + BX - 1)/BX, BX, 0, s>>>(b, a, dt, ni, nj, ntot); ++g_launches;
}

// register average of one face value (same modes as k_integrate_cc)
template <int MODE>
__device__ __forceinline__ double fc_avg(double *__restrict__ bo, double *__restrict__ b1,
                                         int o, int zero_init, double delta, double g1,
                                         double g2) {
  if (MODE == 0) return bo[o];
  if (MODE == 1) {
    double v = zero_init ? 0.0 : bo[o];
    if (delta != 0.0) v += delta*b1[o];
    return v;
  }
  double b1v = zero_init ? 0.0 : b1[o];
  double bn = bo[o];
  if (delta != 0.0 || zero_init) {
    if (delta != 0.0) b1v += delta*bn;
    b1[o] = b1v;
  }
  return wave2(bn, b1v, g1, g2);
}

// IntegrateField: face-register average + Field::CT (field/ct.cpp:31-116), fused.
// Thread (k,j,i) of the flattened box [ks,ke+1]x[js,je+1]x[is,ie+1] updates x1f(k,j,i)
// [j<=je,k<=ke], x2f(k,j,i) [i<=ie,k<=ke] and x3f(k,j,i) [i<=ie,j<=je]; the nine edge EMFs it
// needs are loaded once (e?(k,j,i) is shared by two of the three updates).
template <int MODE>
__global__ void __launch_bounds__(BX, AB_FC_MINB) k_integrate_fc(BlkDev b0, int ni, int nj, int ntot,
                                                     int zero_init, double delta, double g1,
                                                     double g2, double beta, double dt_val,
                                                     const double *dt_ptr, FastDiv di, FastDiv dj) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  int t = blockIdx.x*BX + threadIdx.x;
  if (t >= ntot) return;
  int r = fast_div(t, di);
  const int i = b.is + (t - r*ni);
  const int kk = fast_div(r, dj);
  const int j = b.js + (r - kk*nj);
  const int k = b.ks + kk;
  const double wght = beta*(dt_ptr ? *dt_ptr : dt_val);
  const double *__restrict__ e1 = b.e[0];
  const double *__restrict__ e2 = b.e[1];
  const double *__restrict__ e3 = b.e[2];
  const bool in_i = (i <= b.ie), in_j = (j <= b.je), in_k = (k <= b.ke);
  const double dx1 = b.dx1f[in_i ? i : b.ie], dx2 = b.dx2f[in_j ? j : b.je],
               dx3 = b.dx3f[in_k ? k : b.ke];
  // edge-array offsets (EdgeField shapes, athena.hpp:107-115)
  const int n1 = b.nc1, n2 = b.nc2;
  const int o1 = (k*(n2+1) + j)*n1 + i;          // x1e(k,j,i)
  const int o2 = (k*n2 + j)*(n1+1) + i;          // x2e(k,j,i)
  const int o3 = (k*(n2+1) + j)*(n1+1) + i;      // x3e(k,j,i)
  const bool f2 = b.f2, f3 = b.f3;
  const bool do1 = in_j && in_k, do2 = in_i && in_k, do3 = in_i && in_j;
  // All loads are issued up front (memory-level parallelism): an element that a thread does
  // not need is replaced by a harmless in-bounds re-read of x?e(k,j,i) of a valid thread.
  const double e3c = in_k ? e3[o3] : 0.0;
  const double e2c = in_j ? e2[o2] : 0.0;
  const double e1c = (in_i && f2) ? e1[o1] : 0.0;
  const double e3_jp = (do1 && f2) ? e3[o3 + (n1+1)] : 0.0;          // x3e(k,j+1,i)
  const double e2_kp = (do1 && f3) ? e2[o2 + n2*(n1+1)] : 0.0;       // x2e(k+1,j,i)
  const double e3_ip = do2 ? e3[o3 + 1] : 0.0;                       // x3e(k,j,i+1)
  const double e1_kp = (do2 && f3) ? e1[o1 + (n2+1)*n1] : 0.0;       // x1e(k+1,j,i)
  const double e2_ip = do3 ? e2[o2 + 1] : 0.0;                       // x2e(k,j,i+1)
  const double e1_jp = (do3 && f2) ? e1[o1 + n1] : 0.0;              // x1e(k,j+1,i)
  const int ob1 = (k*n2 + j)*(n1+1) + i, ob2 = (k*(n2+1) + j)*n1 + i, ob3 = (k*n2 + j)*n1 + i;
  double v1 = 0.0, v2 = 0.0, v3 = 0.0;
  if (do1) v1 = fc_avg<MODE>(b.b[0], b.b1[0], ob1, zero_init, delta, g1, g2);
  if (do2) v2 = fc_avg<MODE>(b.b[1], b.b1[1], ob2, zero_init, delta, g1, g2);
  if (do3) v3 = fc_avg<MODE>(b.b[2], b.b1[2], ob3, zero_init, delta, g1, g2);
  if (do1) {                           // B1 (ct.cpp:40-62)
    if (f2) {
      const double c = wght/(dx2*dx3);        // (wght/area(i)), identical in both terms
      v1 -= c*(dx3*e3_jp - dx3*e3c);
      if (f3) v1 += c*(dx2*e2_kp - dx2*e2c);
    }
    b.b[0][ob1] = v1;
  }
  if (do2) {                           // B2 (ct.cpp:64-92)
    const double c = wght/(dx1*dx3);
    v2 += c*(dx3*e3_ip - dx3*e3c);
    if (f3) v2 -= c*(dx1*e1_kp - dx1*e1c);
    b.b[1][ob2] = v2;
  }
  if (do3) {                           // B3 (ct.cpp:94-114)
    const double c = wght/(dx1*dx2);
    v3 -= c*(dx2*e2_ip - dx2*e2c);
    if (f2) v3 += c*(dx1*e1_jp - dx1*e1c);
    b.b[2][ob3] = v3;
  }
}

void launch_integrate_fc(const BlkDev &b, int mode, int zero_init, double delta, double g1,
                         double g2, double beta, double dt_val, const double *dt_ptr,
                         cudaStream_t s, int nb) {
  const int ni = b.ie-b.is+2, nj = b.je-b.js+2, nk = b.ke-b.ks+2;
  const int ntot = ni*nj*nk;
  const dim3 g((unsigned)((ntot + BX - 1)/BX), (unsigned)nb);
  const FastDiv di = make_fastdiv(ni), dj = make_fastdiv(nj);
  if (mode == 0) k_integrate_fc<0><<<g, BX, 0, s>>>(b, ni, nj, ntot, zero_init, delta, g1, g2, beta, dt_val, dt_ptr, di, dj);
  else if (mode == 1) k_integrate_fc<1><<<g, BX, 0, s>>>(b, ni, nj, ntot, zero_init, delta, g1, g2, beta, dt_val, dt_ptr, di, dj);
  else k_integrate_fc<2><<<g, BX, 0, s>>>(b, ni, nj, ntot, zero_init, delta, g1, g2, beta, dt_val, dt_ptr, di, dj);
  ++g_launches;
}

// =============================================================================================
// Ghost-zone exchange as box copies (bvals/cc/bvals_cc.cpp:201-216,300-336,
// bvals/fc/bvals_fc.cpp:344-397,583-684; utils/buffer_utils.cpp ordering n,k,j,i)
// =============================================================================================
// All boxes of one exchange phase form ONE flat element list (CopyBox::offset = exclusive prefix
// of the element counts): a thread finds its box by binary search.  A (boxes x chunks) grid
// launched millions of empty CTAs on many-block meshes (6656 boxes of very different sizes at
// 64 MeshBlocks: 3.3 ms per phase).
// `chunked`: a table of one int per 256 consecutive elements follows the boxes in the same
// allocation -- the box that holds the chunk's first element; a thread scans forward from there
// (0-2 steps) instead of a 13-step binary search over boxes 120 bytes apart.  The element index
// inside a box is decoded with 32-bit multiply-high divisions (a box has < 2^31 elements): the
// five 64-bit divisions it took before made this copy instruction-bound (0.5 ms for 27 M
// elements at 64 MeshBlocks).
__global__ void __launch_bounds__(256) k_copy_boxes(const CopyBox *__restrict__ boxes, int n,
                                                    long total, int chunked) {
  const int *__restrict__ first = reinterpret_cast<const int *>(boxes + n);
  for (long t0 = (long)blockIdx.x*256; t0 < total; t0 += (long)gridDim.x*256) {
    const long t = t0 + threadIdx.x;
    if (t >= total) break;
    int lo;
    if (chunked & 1) {
      lo = first[t0 >> 8];
      while (lo + 1 < n && boxes[lo + 1].offset <= t) ++lo;
    } else {
      lo = 0;
      int hi = n - 1;
      while (lo < hi) {                       // last box with offset <= t
        const int mid = (lo + hi + 1) >> 1;
        if (boxes[mid].offset <= t) lo = mid; else hi = mid - 1;
      }
    }
    const CopyBox &bx = boxes[lo];
    const int e = (int)(t - bx.offset);
    const int v = fast_div(e, bx.d_per);
    const int r = e - v*(bx.ni*bx.nj*bx.nk);
    const int r2 = fast_div(r, bx.d_ni);
    const int i = r - r2*bx.ni;
    const int k = fast_div(r2, bx.d_nj);
    const int j = r2 - k*bx.nj;
    bx.dst[v*bx.dst_sv + (long)(bx.dk0+k)*bx.dst_s3 + (long)(bx.dj0+j)*bx.dst_s2 + (bx.di0+i)] =
        bx.src[v*bx.src_sv + (long)(bx.sk0+k)*bx.src_s3 + (long)(bx.sj0+j)*bx.src_s2 + (bx.si0+i)];
  }
  // bit 1 of `chunked`: the destinations are another GPU's memory (direct ghost-zone exchange):
  // make the stores visible system-wide before the kernel ends
#ifndef AB_HOST_EMU
  if (chunked & 2) __threadfence_system();
#endif
}

void launch_copy_boxes(const CopyBox *boxes_dev, int n, long total_elems, cudaStream_t s,
                       int chunked) {
  if (n <= 0 || total_elems <= 0) return;
  long gx = (total_elems + 255)/256;
  if (gx > 148L*64) gx = 148L*64;
  k_copy_boxes<<<(unsigned)gx, 256, 0, s>>>(boxes_dev, n, total_elems, chunked); ++g_launches;
}

// =============================================================================================
// Outflow / reflecting physical boundaries on primitives and face fields
// (bvals/cc/outflow_cc.cpp, cc/hydro/reflect_hydro.cpp, fc/outflow_fc.cpp, fc/reflect_fc.cpp).
// One thread per pair of transverse indices; loops over the ng ghost layers.  Ghost layer g
// copies the last active cell/face (outflow) or the mirrored one with the normal velocity and
// the normal field negated (reflect).
// =============================================================================================
__global__ void __launch_bounds__(BX) k_phys_bc(BlkDev b, int mhd, int face, int refl, int il,
                                                int iu, int jl, int ju, int kl, int ku) {
  const int d = face >> 1, upper = face & 1, ng = b.ng;
  const int a = blockIdx.x*BX + threadIdx.x, c = blockIdx.y;
  const int lo = d == 0 ? il : (d == 1 ? jl : kl), hi = d == 0 ? iu : (d == 1 ? ju : ku);
  for (int g = 1; g <= ng; ++g) {
    const int gc = upper ? hi + g : lo - g;
    const int sc = refl ? (upper ? hi - g + 1 : lo + g - 1) : (upper ? hi : lo);
    const int gn = upper ? hi + g + 1 : lo - g;
    const int sn = refl ? (upper ? hi - g + 1 : lo + g) : (upper ? hi + 1 : lo);
    if (d == 0) {
      const int j = jl + a, k = kl + c;
      if (j > ju+1 || k > ku+1) return;
      if (j <= ju && k <= ku) {
        for (int n = 0; n < b.nh; ++n) {
          double v = b.w[CCI(b,n,k,j,sc)];
          b.w[CCI(b,n,k,j,gc)] = (refl && n == IVX) ? -v : v;
        }
        if (mhd) { double v = b.b[0][F1I(b,k,j,sn)]; b.b[0][F1I(b,k,j,gn)] = refl ? -v : v; }
        // passive scalars: outflow_cc.cpp / reflect_cc.cpp on r (copy / mirror, no sign)
        for (int n = 0; n < b.ns; ++n) b.r[CCI(b,n,k,j,gc)] = b.r[CCI(b,n,k,j,sc)];
      }
      if (mhd && k <= ku) b.b[1][F2I(b,k,j,gc)] = b.b[1][F2I(b,k,j,sc)];
      if (mhd && j <= ju) b.b[2][F3I(b,k,j,gc)] = b.b[2][F3I(b,k,j,sc)];
    } else if (d == 1) {
      const int i = il + a, k = kl + c;
      if (i > iu+1 || k > ku+1) return;
      if (i <= iu && k <= ku) {
        for (int n = 0; n < b.nh; ++n) {
          double v = b.w[CCI(b,n,k,sc,i)];
          b.w[CCI(b,n,k,gc,i)] = (refl && n == IVY) ? -v : v;
        }
        if (mhd) { double v = b.b[1][F2I(b,k,sn,i)]; b.b[1][F2I(b,k,gn,i)] = refl ? -v : v; }
        for (int n = 0; n < b.ns; ++n) b.r[CCI(b,n,k,gc,i)] = b.r[CCI(b,n,k,sc,i)];
      }
      if (mhd && k <= ku) b.b[0][F1I(b,k,gc,i)] = b.b[0][F1I(b,k,sc,i)];
      if (mhd && i <= iu) b.b[2][F3I(b,k,gc,i)] = b.b[2][F3I(b,k,sc,i)];
    } else {
      const int i = il + a, j = jl + c;
      if (i > iu+1 || j > ju+1) return;
      if (i <= iu && j <= ju) {
        for (int n = 0; n < b.nh; ++n) {
          double v = b.w[CCI(b,n,sc,j,i)];
          b.w[CCI(b,n,gc,j,i)] = (refl && n == IVZ) ? -v : v;
        }
        if (mhd) { double v = b.b[2][F3I(b,sn,j,i)]; b.b[2][F3I(b,gn,j,i)] = refl ? -v : v; }
        for (int n = 0; n < b.ns; ++n) b.r[CCI(b,n,gc,j,i)] = b.r[CCI(b,n,sc,j,i)];
      }
      if (mhd && j <= ju) b.b[0][F1I(b,gc,j,i)] = b.b[0][F1I(b,sc,j,i)];
      if (mhd && i <= iu) b.b[1][F2I(b,gc,j,i)] = b.b[1][F2I(b,sc,j,i)];
    }
  }
}

void launch_phys_bc(const BlkDev &b, int mhd, int face, int refl, int il, int iu, int jl,
                    int ju, int kl, int ku, cudaStream_t s) {
  int d = face >> 1;
  int na, nc;
  if (d == 0) { na = ju-jl+2; nc = ku-kl+2; }
  else if (d == 1) { na = iu-il+2; nc = ku-kl+2; }
  else { na = iu-il+2; nc = ju-jl+2; }
  k_phys_bc<<<dim3((unsigned)((na + BX - 1)/BX), (unsigned)nc), BX, 0, s>>>(b, mhd, face, refl, il,
                                                                          iu, jl, ju, kl, ku); ++g_launches;
}

// =============================================================================================
// Hydro::NewBlockTimeStep (hydro/new_blockdt.cpp:42-190): CFL min-reduction
// =============================================================================================
template <bool MHD>
__global__ void __launch_bounds__(BX) k_new_dt(BlkDev b, Params p, unsigned long long *out) {
  int i = b.is + blockIdx.x*BX + threadIdx.x;
  int j = b.js + blockIdx.y, k = b.ks + blockIdx.z;
  double m = DBL_MAX;
  if (i <= b.ie) {
    long sv = (long)b.nc3*b.nc2*b.nc1;
    long o = CCI(b,0,k,j,i);
    const bool iso = (p.eos != 0);
    double d = b.w[o], vx = b.w[o+sv], vy = b.w[o+2*sv], vz = b.w[o+3*sv];
    double pr = iso ? 0.0 : b.w[o+4*sv];
    double dt1 = b.dx1f[i], dt2 = b.dx2f[j], dt3 = b.dx3f[k];
    if (MHD) {
      double b1c = b.bcc[o], b2c = b.bcc[o+sv], b3c = b.bcc[o+2*sv];
      double bx = b1c + fabs(b.b[0][F1I(b,k,j,i)] - b1c);
      double cf = iso ? fast_speed_iso(p.iso_cs, d, b2c, b3c, bx)
                      : fast_speed(p.gamma, d, pr, b2c, b3c, bx);
      dt1 /= (fabs(vx) + cf);
      bx = b2c + fabs(b.b[1][F2I(b,k,j,i)] - b2c);
      cf = iso ? fast_speed_iso(p.iso_cs, d, b3c, b1c, bx)
               : fast_speed(p.gamma, d, pr, b3c, b1c, bx);
      dt2 /= (fabs(vy) + cf);
      bx = b3c + fabs(b.b[2][F3I(b,k,j,i)] - b3c);
      cf = iso ? fast_speed_iso(p.iso_cs, d, b1c, b2c, bx)
               : fast_speed(p.gamma, d, pr, b1c, b2c, bx);
      dt3 /= (fabs(vz) + cf);
    } else {
      double cs = iso ? p.iso_cs : sound_speed(p.gamma, d, pr);
      dt1 /= (fabs(vx) + cs);
      dt2 /= (fabs(vy) + cs);
      dt3 /= (fabs(vz) + cs);
    }
    m = dmin(m, dt1);
    if (b.f2) m = dmin(m, dt2);
    if (b.f3) m = dmin(m, dt3);
  }
  block_min_to_slots(m, out);
}

void launch_new_block_dt(const BlkDev &b, const Params &p, unsigned long long *out_bits,
                         cudaStream_t s) {
  dim3 g = grid3(b.ie-b.is+1, b.je-b.js+1, b.ke-b.ks+1);
  if (p.mhd) { k_new_dt<true><<<g, BX, 0, s>>>(b, p, out_bits); } else { k_new_dt<false><<<g, BX, 0, s>>>(b, p, out_bits); }
  ++g_launches;
}

// =============================================================================================
// HistoryOutput sums (outputs/history.cpp:69-169): volume-weighted sums over the active cells.
// Two deterministic levels (fixed grid, fixed order): per-CTA partials, then one thread per
// quantity adds the partials in CTA order onto the running total of the mesh.
// =============================================================================================
constexpr int HIST_T = 256;
constexpr int HIST_MAXQ = 32;

template <bool ISO>
__global__ void __launch_bounds__(HIST_T) k_history(BlkDev b, int mhd, int nq, int ni, int nj,
                                                    int ntot, double *partial) {
  constexpr int NB = ISO ? 7 : 8;    // quantities ahead of the magnetic energies (NHYDRO + 3)
  const int n1 = b.nc1, n2 = b.nc2;
  const int sv = b.nc3*n2*n1;
  double acc[HIST_MAXQ];
#pragma unroll
  for (int q = 0; q < HIST_MAXQ; ++q) acc[q] = 0.0;
  for (int t = blockIdx.x*HIST_T + threadIdx.x; t < ntot; t += gridDim.x*HIST_T) {
    int r = t / ni;
    const int i = b.is + (t - r*ni);
    const int kk = r / nj;
    const int j = b.js + (r - kk*nj);
    const int k = b.ks + kk;
    const int o = (k*n2 + j)*n1 + i;
    const double vol = b.dx1f[i]*b.dx2f[j]*b.dx3f[k];
    const double u_d = b.u[o], u_mx = b.u[o+sv], u_my = b.u[o+2*sv], u_mz = b.u[o+3*sv];
    acc[0] += vol*u_d;
    acc[1] += vol*u_mx;
    acc[2] += vol*u_my;
    acc[3] += vol*u_mz;
    acc[4] += vol*0.5*sqr(u_mx)/u_d;
    acc[5] += vol*0.5*sqr(u_my)/u_d;
    acc[6] += vol*0.5*sqr(u_mz)/u_d;
    if (!ISO) acc[7] += vol*b.u[o+4*sv];
    if (mhd) {
      const double bcc1 = b.bcc[o], bcc2 = b.bcc[o+sv], bcc3 = b.bcc[o+2*sv];
      acc[NB] += vol*0.5*bcc1*bcc1;
      acc[NB + 1] += vol*0.5*bcc2*bcc2;
      acc[NB + 2] += vol*0.5*bcc3*bcc3;
    }
    // scalars follow the magnetic energies (history.cpp:158-162)
    if (mhd) {
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < b.ns) acc[NB + 3 + n] += vol*b.s[o + n*sv];
    } else {
#pragma unroll
      for (int n = 0; n < 16; ++n)
        if (n < b.ns) acc[NB + n] += vol*b.s[o + n*sv];
    }
  }
  __shared__ double sm[HIST_T/32][HIST_MAXQ];
#pragma unroll
  for (int q = 0; q < HIST_MAXQ; ++q) {
    double v = acc[q];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < nq) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < HIST_T/32; ++w) v += sm[w][threadIdx.x];
    partial[blockIdx.x*HIST_MAXQ + threadIdx.x] = v;
  }
}

__global__ void k_history_final(const double *partial, int ncta, int nq, int first,
                                double *out) {
  const int q = threadIdx.x;
  if (q >= nq) return;
  double v = first ? 0.0 : out[q];
  for (int c = 0; c < ncta; ++c) v += partial[c*HIST_MAXQ + q];
  out[q] = v;
}

int history_grid() { return 148*2; }

// partial: device scratch of history_grid()*32 doubles; out: device totals (nq doubles)
void launch_history(const BlkDev &b, int mhd, int nq, int first, double *partial, double *out,
                    cudaStream_t s) {
  const int ni = b.ie-b.is+1, nj = b.je-b.js+1, nk = b.ke-b.ks+1;
  const int ntot = ni*nj*nk;
  int g = (ntot + HIST_T - 1)/HIST_T;
  if (g > history_grid()) g = history_grid();
  if (b.nh == NHYDRO) k_history<false><<<g, HIST_T, 0, s>>>(b, mhd, nq, ni, nj, ntot, partial);
  else k_history<true><<<g, HIST_T, 0, s>>>(b, mhd, nq, ni, nj, ntot, partial);
  ++g_launches;
  k_history_final<<<1, 32, 0, s>>>(partial, g, nq, first, out); ++g_launches;
}

__global__ void k_fill_u64(unsigned long long *p, int n, unsigned long long v) {
  int t = blockIdx.x*blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}
void launch_fill_u64(unsigned long long *p, int n, unsigned long long v, cudaStream_t s) {
  k_fill_u64<<<(n + 127)/128, 128, 0, s>>>(p, n, v); ++g_launches;
}

// state: [0]=time [1]=dt [2]=tlim [3]=cfl [4]=min over (local or all) blocks of new_block_dt
// [5]=ncycle.  Mesh::NewTimeStep (mesh/mesh.cpp:1078-1119) + main.cpp:478-481 bookkeeping.
// phase 0: reduce local blocks into state[4] (times cfl, per block like new_blockdt.cpp:164)
// phase 1: apply NewTimeStep with state[4] (after the optional cross-rank MIN all-reduce)
__global__ void k_mesh_new_dt(double *st, const unsigned long long *blk_min, int nb, int phase,
                              int advance_time) {
  if (blockIdx.x != 0) return;
  if (phase == 0) {
    // one warp over the nb*DT_SLOTS partial minima (a single thread took 90 us at 64 blocks).
    // min_n(cfl * min_q x) == cfl * min_{n,q} x bit for bit: rounding is monotonic and cfl > 0.
    double m = DBL_MAX;
    for (int q = threadIdx.x; q < nb*DT_SLOTS; q += 32)
      m = dmin(m, __longlong_as_double((long long)blk_min[q]));
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) m = dmin(m, __shfl_xor_sync(0xffffffffu, m, sft));
    if (threadIdx.x == 0) st[4] = m*st[3];
  } else if (threadIdx.x == 0) {
    // st[1] = Mesh::dt as the reference keeps it; st[6] = the dt the next cycle integrates
    // with: 0 once time has reached tlim, so that cycles launched asynchronously past tlim
    // change nothing and are not counted (the reference's loop stops there, main.cpp:430)
    const bool ran = st[0] < st[2];
    if (advance_time && ran) { st[5] += 1.0; st[0] += st[1]; }
    if (!advance_time || ran) {
      double dt = 2.0*st[1];
      dt = dmin(dt, st[4]);
      if (st[0] < st[2] && (st[2] - st[0]) < dt) dt = st[2] - st[0];
      st[1] = dt;
    }
    st[6] = (st[0] < st[2]) ? st[1] : 0.0;
  }
}
void launch_mesh_new_dt(double *state, const unsigned long long *blk_min, int nb,
                        int advance_time, cudaStream_t s) {
  // advance_time: bit0 = advance, bit1 = phase
  k_mesh_new_dt<<<1, 32, 0, s>>>(state, blk_min, nb, (advance_time >> 1) & 1, advance_time & 1); ++g_launches;
}

}  // namespace ab
