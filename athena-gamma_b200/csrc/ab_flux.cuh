// ab_flux.cuh -- the Riemann-sweep kernel (Hydro::CalculateFluxes) as a template, shared by the
// two translation units that instantiate it: ab_kernels.cu (uniform spacing, the production
// path) and ab_flux_nu.cu (nonuniform spacing, mesh/x?rat != 1), compiled in parallel.
#ifndef AB_FLUX_CUH_
#define AB_FLUX_CUH_
#include "ab_kernels.h"
#include "ab_physics.cuh"
#include "ab_batch.cuh"

namespace ab {

extern std::atomic<long> g_launches;

// =============================================================================================
// Hydro::CalculateFluxes: reconstruction + Riemann solver, one thread per interface.
// hydro/calculate_fluxes.cpp:36-378; reconstruct/{dc,plm,ppm}.cpp; hydro/rsolvers/*
// =============================================================================================

// sweep-ordered primitives of one cell: (rho, v_dir, v_dir+1, v_dir+2, p [, B_dir+1, B_dir+2])
// (32-bit element offsets: the host checks that every register has < 2^31 elements)
template <int DIR, bool MHD, bool ISO = false>
__device__ __forceinline__ void load_cell(const double *__restrict__ w,
                                          const double *__restrict__ bcc, int o, int sv,
                                          double *q) {
  q[IDN] = w[o];
  q[IVX] = w[o + (1 + DIR)*sv];
  q[IVY] = w[o + (1 + (DIR+1)%3)*sv];
  q[IVZ] = w[o + (1 + (DIR+2)%3)*sv];
  q[IPR] = ISO ? 0.0 : w[o + 4*sv];       // isothermal: w has 4 variables, slot unused
  if (MHD) {
    q[IBY] = bcc[o + ((DIR+1)%3)*sv];
    q[IBZ] = bcc[o + ((DIR+2)%3)*sv];
  }
}

// 96 registers per thread (about 180 B of spills) with 18 single-warp CTAs per SM measured best
// on B200 (profiles/r1_tuning_log.md): at 128-thread CTAs 3/4/5/6/8 CTAs per SM give
// 1.27/1.10/1.05/1.06/1.37 ms per 256^3 HLLD+PLM sweep and the spill-free 144-register build
// is 20 % slower; at equal registers, smaller CTAs (the warps of an SM start and stall on
// their loads less in lockstep) and 18 instead of 20 warps take another 5 % off the sweeps.
#ifndef AB_FLUX_BX
#define AB_FLUX_BX 32
#endif
#ifndef AB_FLUX_MINB
#define AB_FLUX_MINB 18
#endif
// x3 sweep: faces are visited strip by strip (AB_X3_STRIP rows of j, all k) so that the four
// k-planes of the stencil stay in L2 between consecutive k (plane-major order re-read w/bcc
// from DRAM: 19 GB instead of 8.8 GB per 512^3 sweep).
#ifdef AB_FLUX_MAXREG
#define AB_FLUX_BOUNDS __maxnreg__(AB_FLUX_MAXREG)
#else
#define AB_FLUX_BOUNDS __launch_bounds__(AB_FLUX_BX, AB_FLUX_MINB)
#endif
#ifndef AB_X3_STRIP
#define AB_X3_STRIP 32
#endif
// 1: xorder = 3 sweeps reconstruct every cell once (k_flux_ppm_*); 0: per interface (k_flux)
#ifndef AB_PPM_SHARED
#define AB_PPM_SHARED 1
#endif

// divisors of the flattened face index: ni, nj, and for the strip-major x3 order ni*STRIP*nk,
// ni*STRIP (full strips) and ni*(rows of the last, partial strip)
struct FluxIdx { FastDiv ni, nj, per_full, per_k, per_k_last; int last_strip; };

// The face range [i0,i0+ni) x [j0,j0+nj) x [k0,k0+nk) is flattened so that every thread of a
// CTA has work (rows of nx1+1 faces do not pad to a multiple of the CTA width).
template <int DIR, int ORDER, int SOLVER, bool MHD, bool NU>
__global__ void AB_FLUX_BOUNDS
k_flux(BlkDev b0, ReconGeom g0, Params p, int i0, int ni, int j0, int nj, int k0, int nk,
       int ntot, double dt_val, const double *dt_ptr, FluxIdx fx) {
  const BlkDev b = blk_view(b0, blockIdx.y);      // blockIdx.y = local MeshBlock (ab_batch.cuh)
  const ReconGeom g = geom_view(g0, b0, blockIdx.y);
  constexpr int NW = MHD ? 7 : 5;
  constexpr bool ISO = solver_is_iso<SOLVER>;
  int t = blockIdx.x*AB_FLUX_BX + threadIdx.x;
  if (t >= ntot) return;
  int i, j, k;
  if (DIR == 2 && AB_X3_STRIP > 0) {
    // strip-major: strip s of j-rows, then k, then j within the strip, then i
    const int per_full = ni*AB_X3_STRIP*nk;
    const int s = fast_div(t, fx.per_full);
    int r = t - s*per_full;
    int rows = nj - s*AB_X3_STRIP;
    rows = rows < AB_X3_STRIP ? rows : AB_X3_STRIP;
    const int per_k = ni*rows;
    const int kk = fast_div(r, s == fx.last_strip ? fx.per_k_last : fx.per_k);
    r -= kk*per_k;
    const int jj = fast_div(r, fx.ni);
    i = i0 + (r - jj*ni);
    j = j0 + s*AB_X3_STRIP + jj;
    k = k0 + kk;
  } else {
    int r = fast_div(t, fx.ni);
    i = i0 + (t - r*ni);
    const int kk = fast_div(r, fx.nj);
    j = j0 + (r - kk*nj);
    k = k0 + kk;
  }
  const int sv = b.nc3*b.nc2*b.nc1;
  const int st = (DIR == 0) ? 1 : ((DIR == 1) ? b.nc1 : b.nc1*b.nc2);
  const int oc = (k*b.nc2 + j)*b.nc1 + i;      // cell on the upper side of the face
  const int c = (DIR == 0) ? i : ((DIR == 1) ? j : k);
  const double *__restrict__ w = b.w;
  const double *__restrict__ bcc = b.bcc;
  // face-array offset and variable stride
  int of, sf;
  if (DIR == 0) { of = (k*b.nc2 + j)*(b.nc1+1) + i; sf = b.nc3*b.nc2*(b.nc1+1); }
  else if (DIR == 1) { of = (k*(b.nc2+1) + j)*b.nc1 + i; sf = b.nc3*(b.nc2+1)*b.nc1; }
  else { of = (k*b.nc2 + j)*b.nc1 + i; sf = (b.nc3+1)*b.nc2*b.nc1; }
  // Every global load of the thread is issued here, ahead of the reconstruction, so that their
  // latencies overlap (ncu: the face field / dt / dx loads used to sit right in front of their
  // first use inside the Riemann solver and cost 10 % of the kernel in long-scoreboard stalls).
  double bxi = 0.0, dt = 0.0, dxw = 0.0;
  if (MHD) {
    bxi = b.b[DIR][of];
    dt = dt_ptr ? *dt_ptr : dt_val;
    dxw = (DIR == 0) ? b.dx1f[i] : ((DIR == 1) ? b.dx2f[j] : b.dx3f[k]);
    dt = (1024.0)*dt;               // GetWeightForCT: (1024*dt*dflx)/(dx*(rhol + rhor))
  }

  double wl[NW], wr[NW];
  if (ORDER == 1) {
    // DonorCell (reconstruct/dc.cpp)
    load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, wl);
    load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, wr);
  } else if (ORDER == 2) {
    double qm2[NW], qm1[NW], q0[NW], qp1[NW];
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 2*st, sv, qm2);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, qm1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, q0);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + st, sv, qp1);
    const double wp_l = g.wp[DIR][c-1], wm_l = g.wm[DIR][c-1];
    const double wp_r = g.wp[DIR][c], wm_r = g.wm[DIR][c];
    const double *tl = NU ? g.nu[DIR] + (c-1)*NUG : nullptr;   // geometry rows of the two cells
#pragma unroll
    for (int n = 0; n < NW; ++n) {
      double dummy;
      if (NU) {
        plm_nu<DIR+1>(qm2[n], qm1[n], q0[n], wp_l, wm_l, tl, wl[n], dummy);
        plm_nu<DIR+1>(qm1[n], q0[n], qp1[n], wp_r, wm_r, tl + NUG, dummy, wr[n]);
      } else {
        plm(qm2[n], qm1[n], q0[n], wp_l, wm_l, wl[n], dummy);
        plm(qm1[n], q0[n], qp1[n], wp_r, wm_r, dummy, wr[n]);
      }
    }
  } else if (ORDER == 4) {
    // xorder = 2c: PLM on characteristic variables (reconstruct/characteristic.cpp); the
    // eigenvectors need the cell-centred field along the sweep of the two cells
    double qm2[NW], qm1[NW], q0[NW], qp1[NW], dummy[NW];
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 2*st, sv, qm2);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, qm1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, q0);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + st, sv, qp1);
    const double bxl = MHD ? bcc[oc - st + DIR*sv] : 0.0, bxr = MHD ? bcc[oc + DIR*sv] : 0.0;
    const double *tl = NU ? g.nu[DIR] + (c-1)*NUG : nullptr;
    plm_char<MHD,(NU ? DIR+1 : 0)>(qm2, qm1, q0, bxl, p.gamma, g.wp[DIR][c-1], g.wm[DIR][c-1],
                                   p.dfloor, p.pfloor, wl, dummy, tl);
    plm_char<MHD,(NU ? DIR+1 : 0)>(qm1, q0, qp1, bxr, p.gamma, g.wp[DIR][c], g.wm[DIR][c],
                                   p.dfloor, p.pfloor, dummy, wr, tl + NUG);
  } else if (ORDER == 5) {
    // xorder = 3c: PPM on characteristic variables
    double qm3[NW], qm2[NW], qm1[NW], q0[NW], qp1[NW], qp2[NW], dummy[NW];
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 3*st, sv, qm3);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 2*st, sv, qm2);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, qm1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, q0);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + st, sv, qp1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + 2*st, sv, qp2);
    const double bxl = MHD ? bcc[oc - st + DIR*sv] : 0.0, bxr = MHD ? bcc[oc + DIR*sv] : 0.0;
    const double *tl = NU ? g.nu[DIR] + (c-1)*NUG : nullptr;
    ppm_char<MHD,NU>(qm3, qm2, qm1, q0, qp1, bxl, p.gamma, p.dfloor, p.pfloor, wl, dummy, tl);
    ppm_char<MHD,NU>(qm2, qm1, q0, qp1, qp2, bxr, p.gamma, p.dfloor, p.pfloor, dummy, wr,
                     tl + NUG);
  } else {
    double qm3[NW], qm2[NW], qm1[NW], q0[NW], qp1[NW], qp2[NW];
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 3*st, sv, qm3);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - 2*st, sv, qm2);
    load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, qm1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, q0);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + st, sv, qp1);
    load_cell<DIR,MHD,ISO>(w, bcc, oc + 2*st, sv, qp2);
    const double *tl = NU ? g.nu[DIR] + (c-1)*NUG : nullptr;
#pragma unroll
    for (int n = 0; n < NW; ++n) {
      double dummy;
      if (NU) {
        ppm_nu(qm3[n], qm2[n], qm1[n], q0[n], qp1[n], tl, wl[n], dummy);
        ppm_nu(qm2[n], qm1[n], q0[n], qp1[n], qp2[n], tl + NUG, dummy, wr[n]);
      } else {
        ppm(qm3[n], qm2[n], qm1[n], q0[n], qp1[n], wl[n], dummy);
        ppm(qm2[n], qm1[n], q0[n], qp1[n], qp2[n], dummy, wr[n]);
      }
    }
    // ApplyPrimitiveFloors (ppm.cpp:326-332)
    wl[IDN] = (wl[IDN] > p.dfloor) ? wl[IDN] : p.dfloor;
    wr[IDN] = (wr[IDN] > p.dfloor) ? wr[IDN] : p.dfloor;
    if (!ISO) {
      wl[IPR] = (wl[IPR] > p.pfloor) ? wl[IPR] : p.pfloor;
      wr[IPR] = (wr[IPR] > p.pfloor) ? wr[IPR] : p.pfloor;
    }
  }

  // LHLLC / LHLLD shock detector inputs: Hydro::CalculateVelocityDifferences
  // (hydro/calculate_velocity_differences.cpp:20-90); dvt stays 0 in 1-D
  double dvn = 0.0, dvt = 0.0;
  if (SOLVER == SOLVER_LHLLC || SOLVER == SOLVER_LHLLD) {
    const int ol = oc - st;                       // lower cell of the interface
    dvn = w[oc + (1 + DIR)*sv] - w[ol + (1 + DIR)*sv];
    if (b.f2) {
      bool first = true;
#pragma unroll
      for (int t = 1; t <= 2; ++t) {
        const int td = (DIR + t) % 3;             // transverse direction, reference order
        if (td == 2 && !b.f3) continue;
        const int ts = (td == 0) ? 1 : ((td == 1) ? b.nc1 : b.nc1*b.nc2);
        const double *__restrict__ wt = w + (1 + td)*sv;
        double dl = dmin(wt[ol + ts] - wt[ol], wt[ol] - wt[ol - ts]);
        double dr = dmin(wt[oc + ts] - wt[oc], wt[oc] - wt[oc - ts]);
        double v = dmin(dl, dr);
        dvt = first ? v : dmin(dvt, v);
        first = false;
      }
    }
  }
  if (MHD) dxw = dxw*(wl[IDN] + wr[IDN]);   // the two values the CT weight needs stay live, not four
  double f[NW];
  riemann<SOLVER,MHD>(wl, wr, bxi, ISO ? p.iso_cs : p.gamma, dvn, dvt, f, p.dfloor);

  const long offo = (long)late_block_y()*b0.bstride;      // b0 = the launch's first block
  double *__restrict__ flx = blk_mv(b0.flux[DIR], offo);
  flx[of] = f[IDN];
  flx[of + (1 + DIR)*sf] = f[IVX];
  flx[of + (1 + (DIR+1)%3)*sf] = f[IVY];
  flx[of + (1 + (DIR+2)%3)*sf] = f[IVZ];
  if (!ISO) flx[of + 4*sf] = f[IEN];
  if (MHD) {
    blk_mv(b0.ef[DIR][0], offo)[of] = -f[IBY];
    blk_mv(b0.ef[DIR][1], offo)[of] = f[IBZ];
    blk_mv(b0.wght[DIR], offo)[of] = weight_for_ct_pre(f[IDN], dxw, dt);
  }
}

// =============================================================================================
// PPM sweeps with the reconstruction shared between the two interfaces of a cell.
//
// k_flux reconstructs per interface: the thread of face c runs PPM on cell c-1 (keeping only the
// state at its upper face) and on cell c (keeping only the lower one).  PPM is ~100 FP64
// instructions per variable, more than the Riemann solver itself (hydro: 10 PPM evaluations per
// interface against one HLLC), so here every cell is reconstructed ONCE and hands its upper-face
// state to the thread of the next interface:
//   * x1 sweep: a warp holds 32 consecutive cells of a row, the state travels one lane up by
//     warp shuffle; lanes 1..31 own the 31 faces between them (no shared memory, no barrier);
//   * x2 / x3 sweeps: a CTA of (ROWS+1) warps holds that many consecutive cells along the
//     sweep for 32 positions of the flattened transverse plane; the state goes through shared
//     memory, one barrier; warps 1..ROWS own the faces.
// The values are those of ppm() / ppm_nu() exactly as k_flux calls them, so the fluxes are the
// same bits (ppm.cpp:111-332 computes both face states of a cell in one pass as well).
// =============================================================================================
// Measured on B200 (256^3, developed flow, ms per x2 / x3 sweep; profiles/r2_tuning_log.md):
// hydro (PPM+HLLC, 90 registers unconstrained): 8 rows + 3 CTAs/SM (72 registers, 27 warps)
// 1.03 against 1.20 with 2 CTAs/SM; MHD (PPM+HLLD, at the 96-register limit either way): 7 rows
// (8 warps) + 2 CTAs/SM 2.06 against 2.15 with 8 rows, 3 CTAs/SM spill 390 B and lose.
#ifndef AB_PPM_ROWS_HYDRO
#define AB_PPM_ROWS_HYDRO 8
#endif
#ifndef AB_PPM_MINB_HYDRO
#define AB_PPM_MINB_HYDRO 3
#endif
#ifndef AB_PPM_ROWS_MHD
#define AB_PPM_ROWS_MHD 7
#endif
#ifndef AB_PPM_MINB_MHD
#define AB_PPM_MINB_MHD 2
#endif
#ifndef AB_PPM_X1_MINB_HYDRO
#define AB_PPM_X1_MINB_HYDRO 28   // 72 registers: 1.16 against 1.22 ms per 256^3 sweep at 18
#endif
#ifndef AB_PPM_X1_MINB_MHD
#define AB_PPM_X1_MINB_MHD 18
#endif
template <bool MHD> struct PpmTune {
  static constexpr int ROWS = MHD ? AB_PPM_ROWS_MHD : AB_PPM_ROWS_HYDRO;
  static constexpr int MINB = MHD ? AB_PPM_MINB_MHD : AB_PPM_MINB_HYDRO;
  static constexpr int X1_MINB = MHD ? AB_PPM_X1_MINB_MHD : AB_PPM_X1_MINB_HYDRO;
};

// both face states of cell (k,j,i) along DIR, floors applied (ppm.cpp:326-332)
template <int DIR, bool MHD, bool ISO, bool NU>
__device__ __forceinline__ void ppm_cell(const BlkDev &b, const ReconGeom &g, const Params &p,
                                         int oc, int c, double *plus, double *minus) {
  constexpr int NW = MHD ? 7 : 5;
  const int sv = b.nc3*b.nc2*b.nc1;
  const int st = (DIR == 0) ? 1 : ((DIR == 1) ? b.nc1 : b.nc1*b.nc2);
  const double *__restrict__ w = b.w;
  const double *__restrict__ bcc = b.bcc;
  double qm2[NW], qm1[NW], q0[NW], qp1[NW], qp2[NW];
  load_cell<DIR,MHD,ISO>(w, bcc, oc - 2*st, sv, qm2);
  load_cell<DIR,MHD,ISO>(w, bcc, oc - st, sv, qm1);
  load_cell<DIR,MHD,ISO>(w, bcc, oc, sv, q0);
  load_cell<DIR,MHD,ISO>(w, bcc, oc + st, sv, qp1);
  load_cell<DIR,MHD,ISO>(w, bcc, oc + 2*st, sv, qp2);
  const double *tl = NU ? g.nu[DIR] + c*NUG : nullptr;
#pragma unroll
  for (int n = 0; n < NW; ++n) {
    if (NU) ppm_nu(qm2[n], qm1[n], q0[n], qp1[n], qp2[n], tl, plus[n], minus[n]);
    else ppm(qm2[n], qm1[n], q0[n], qp1[n], qp2[n], plus[n], minus[n]);
  }
  plus[IDN] = (plus[IDN] > p.dfloor) ? plus[IDN] : p.dfloor;
  minus[IDN] = (minus[IDN] > p.dfloor) ? minus[IDN] : p.dfloor;
  if (!ISO) {
    plus[IPR] = (plus[IPR] > p.pfloor) ? plus[IPR] : p.pfloor;
    minus[IPR] = (minus[IPR] > p.pfloor) ? minus[IPR] : p.pfloor;
  }
}

// Riemann solve and stores of the face (k,j,i) of direction DIR from its two states -- the second
// half of k_flux
template <int DIR, int SOLVER, bool MHD>
__device__ __forceinline__ void flux_face(const BlkDev &b, const BlkDev &b0, const Params &p,
                                          int i, int j, int k, double *wl, double *wr,
                                          double dt_val, const double *dt_ptr) {
  constexpr int NW = MHD ? 7 : 5;
  constexpr bool ISO = solver_is_iso<SOLVER>;
  const int sv = b.nc3*b.nc2*b.nc1;
  const int st = (DIR == 0) ? 1 : ((DIR == 1) ? b.nc1 : b.nc1*b.nc2);
  const int oc = (k*b.nc2 + j)*b.nc1 + i;
  const double *__restrict__ w = b.w;
  int of, sf;
  if (DIR == 0) { of = (k*b.nc2 + j)*(b.nc1+1) + i; sf = b.nc3*b.nc2*(b.nc1+1); }
  else if (DIR == 1) { of = (k*(b.nc2+1) + j)*b.nc1 + i; sf = b.nc3*(b.nc2+1)*b.nc1; }
  else { of = (k*b.nc2 + j)*b.nc1 + i; sf = (b.nc3+1)*b.nc2*b.nc1; }
  double bxi = 0.0, dt = 0.0, dxw = 0.0;
  if (MHD) {
    bxi = b.b[DIR][of];
    dt = dt_ptr ? *dt_ptr : dt_val;
    dxw = (DIR == 0) ? b.dx1f[i] : ((DIR == 1) ? b.dx2f[j] : b.dx3f[k]);
    dt = (1024.0)*dt;
  }
  double dvn = 0.0, dvt = 0.0;
  if (SOLVER == SOLVER_LHLLC || SOLVER == SOLVER_LHLLD) {
    const int ol = oc - st;
    dvn = w[oc + (1 + DIR)*sv] - w[ol + (1 + DIR)*sv];
    if (b.f2) {
      bool first = true;
#pragma unroll
      for (int t = 1; t <= 2; ++t) {
        const int td = (DIR + t) % 3;
        if (td == 2 && !b.f3) continue;
        const int ts = (td == 0) ? 1 : ((td == 1) ? b.nc1 : b.nc1*b.nc2);
        const double *__restrict__ wt = w + (1 + td)*sv;
        double dl = dmin(wt[ol + ts] - wt[ol], wt[ol] - wt[ol - ts]);
        double dr = dmin(wt[oc + ts] - wt[oc], wt[oc] - wt[oc - ts]);
        double v = dmin(dl, dr);
        dvt = first ? v : dmin(dvt, v);
        first = false;
      }
    }
  }
  if (MHD) dxw = dxw*(wl[IDN] + wr[IDN]);
  double f[NW];
  riemann<SOLVER,MHD>(wl, wr, bxi, ISO ? p.iso_cs : p.gamma, dvn, dvt, f, p.dfloor);
  const long offo = (long)late_block_y()*b0.bstride;      // b0 = the launch's first block
  double *__restrict__ flx = blk_mv(b0.flux[DIR], offo);
  flx[of] = f[IDN];
  flx[of + (1 + DIR)*sf] = f[IVX];
  flx[of + (1 + (DIR+1)%3)*sf] = f[IVY];
  flx[of + (1 + (DIR+2)%3)*sf] = f[IVZ];
  if (!ISO) flx[of + 4*sf] = f[IEN];
  if (MHD) {
    blk_mv(b0.ef[DIR][0], offo)[of] = -f[IBY];
    blk_mv(b0.ef[DIR][1], offo)[of] = f[IBZ];
    blk_mv(b0.wght[DIR], offo)[of] = weight_for_ct_pre(f[IDN], dxw, dt);
  }
}

// x1 sweep: one warp per 31 faces
struct PpmIdx { FastDiv nseg, nj, ni; int nsegs, gp, nst, np; };

template <int SOLVER, bool MHD, bool NU>
__global__ void __launch_bounds__(32, PpmTune<MHD>::X1_MINB)
k_flux_ppm_x1(BlkDev b0, ReconGeom g0, Params p, int i0, int ni, int j0, int nj, int k0, int nk,
              double dt_val, const double *dt_ptr, PpmIdx fx) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  const ReconGeom g = geom_view(g0, b0, blockIdx.y);
  constexpr int NW = MHD ? 7 : 5;
  constexpr bool ISO = solver_is_iso<SOLVER>;
  // The cells the sweep reconstructs -- per row i0-1 (halo) .. i0+ni-1 -- form one stream over all
  // rows; warp w holds the 32 cells from position 31*w on (consecutive warps overlap by one cell).
  // A lane other than lane 0 whose cell is not a row's halo cell owns that cell's lower face and
  // finds the cell below in the lane below.  Rows need not be a multiple of 31 faces long: 129
  // faces per row (128^3 MeshBlocks) fill 96 % of the lanes instead of 83 %.
  const int lane = threadIdx.x;
  const int sp = blockIdx.x*31 + lane;           // position in the stream
  const int row = fast_div(sp, fx.nseg);         // nseg = ni + 1 cells per row
  const int c = sp - row*fx.nsegs;
  const int kk = fast_div(row, fx.nj);
  const int j = j0 + (row - kk*nj), k = k0 + kk;
  const int i = i0 - 1 + c;                      // this lane's cell; its lower face is face i
  const bool have = (sp < fx.np);
  double plus[NW], minus[NW];
#pragma unroll
  for (int n = 0; n < NW; ++n) { plus[n] = 0.0; minus[n] = 0.0; }
  if (have) ppm_cell<0,MHD,ISO,NU>(b, g, p, (k*b.nc2 + j)*b.nc1 + i, i, plus, minus);
  double wl[NW];
#pragma unroll
  for (int n = 0; n < NW; ++n) wl[n] = __shfl_up_sync(0xffffffffu, plus[n], 1);
  if (!have || lane == 0 || c == 0) return;
  flux_face<0,SOLVER,MHD>(b, b0, p, i, j, k, wl, minus, dt_val, dt_ptr);
}

// x2 / x3 sweeps.  The transverse plane (x2 sweep: (k,i); x3 sweep: (j,i)) is flattened into np
// positions, 32 per CTA; along the sweep a CTA covers PpmTune::ROWS faces.  CTA order: groups of gp
// transverse tiles, inside a group the sweep tiles one after the other (consecutive CTAs re-read
// the 5 stencil rows they share from L2), inside a sweep tile the gp transverse tiles.
template <int DIR, int SOLVER, bool MHD, bool NU>
__global__ void __launch_bounds__(32*(PpmTune<MHD>::ROWS + 1), PpmTune<MHD>::MINB)
k_flux_ppm_t(BlkDev b0, ReconGeom g0, Params p, int i0, int ni, int a0, int c0, int nc,
             double dt_val, const double *dt_ptr, PpmIdx fx) {
  const BlkDev b = blk_view(b0, blockIdx.y);
  const ReconGeom g = geom_view(g0, b0, blockIdx.y);
  constexpr int NW = MHD ? 7 : 5;
  constexpr bool ISO = solver_is_iso<SOLVER>;
  constexpr int ROWS = PpmTune<MHD>::ROWS;
  __shared__ double sh[ROWS][NW][32];
  const int t = blockIdx.x;
  const int per_group = fx.gp*fx.nst;
  const int grp = t/per_group;                    // uniform per CTA: a plain division is fine
  const int rem = t - grp*per_group;
  const int stile = rem/fx.gp;
  const int ptile = grp*fx.gp + (rem - stile*fx.gp);
  const int lane = threadIdx.x, r = threadIdx.y;
  const int pf = ptile*32 + lane;
  const int a = fast_div(pf, fx.ni);
  const int i = i0 + (pf - a*ni);
  const int c = c0 + stile*ROWS - 1 + r;          // this warp's cell along the sweep
  const bool have = (pf < fx.np) && (c <= c0 + nc - 1);
  const int j = (DIR == 1) ? c : a0 + a, k = (DIR == 1) ? a0 + a : c;
  double plus[NW], minus[NW];
  if (have) {
    ppm_cell<DIR,MHD,ISO,NU>(b, g, p, (k*b.nc2 + j)*b.nc1 + i, c, plus, minus);
    if (r < ROWS) {
#pragma unroll
      for (int n = 0; n < NW; ++n) sh[r][n][lane] = plus[n];
    }
  }
  __syncthreads();
  if (!have || r == 0) return;
  double wl[NW];
#pragma unroll
  for (int n = 0; n < NW; ++n) wl[n] = sh[r-1][n][lane];
  flux_face<DIR,SOLVER,MHD>(b, b0, p, i, j, k, wl, minus, dt_val, dt_ptr);
}

template <int DIR, int SOLVER, bool MHD, bool NU>
static void flux_dir_ppm(const BlkDev &b, const ReconGeom &g, const Params &p, int i0, int ni,
                         int j0, int nj, int k0, int nk, double dt_val, const double *dt_ptr,
                         cudaStream_t s, int nb) {
  PpmIdx fx;
  fx.ni = make_fastdiv(ni); fx.nj = make_fastdiv(nj);
  if constexpr (DIR == 0) {
    fx.nsegs = ni + 1; fx.nseg = make_fastdiv(fx.nsegs);      // cells per row of the stream
    fx.gp = fx.nst = 0;
    fx.np = (ni + 1)*nj*nk;                                   // cells in the stream
    k_flux_ppm_x1<SOLVER,MHD,NU><<<dim3((unsigned)((fx.np - 1 + 30)/31), (unsigned)nb), 32, 0, s>>>(b, g, p, i0, ni, j0, nj, k0, nk,
                                                              dt_val, dt_ptr, fx);
  } else {
    const int na = (DIR == 1) ? nk : nj;           // rows of the transverse plane
    const int nc = (DIR == 1) ? nj : nk;           // faces along the sweep
    fx.np = ni*na;
    const int ntile = (fx.np + 31)/32;
    // x2: a group = the tiles of one k-plane; x3: the tiles of a strip of 32 j-rows
    fx.gp = (DIR == 1) ? (ni + 31)/32 : ni;
    if (fx.gp > ntile) fx.gp = ntile;
    const int ngrp = (ntile + fx.gp - 1)/fx.gp;
    constexpr int ROWS = PpmTune<MHD>::ROWS;
    fx.nst = (nc + ROWS - 1)/ROWS;
    fx.nsegs = 0; fx.nseg = make_fastdiv(1);
    k_flux_ppm_t<DIR,SOLVER,MHD,NU><<<dim3((unsigned)(ngrp*fx.gp*fx.nst), (unsigned)nb), dim3(32, ROWS + 1), 0, s>>>(
        b, g, p, i0, ni, (DIR == 1) ? k0 : j0, (DIR == 1) ? j0 : k0, nc, dt_val, dt_ptr, fx);
  }
  ++g_launches;
}

template <int DIR, int ORDER, int SOLVER, bool MHD, bool NU>
static void flux_dir(const BlkDev &b, const ReconGeom &g, const Params &p, double dt_val,
                     const double *dt_ptr, cudaStream_t s, int nb) {
  int is = b.is, ie = b.ie, js = b.js, je = b.je, ks = b.ks, ke = b.ke;
  int i0, i1, j0, j1, k0, k1;
  // loop limits of calculate_fluxes.cpp:62-74,164-173,273-279
  if (DIR == 0) {
    i0 = is; i1 = ie+1; j0 = js; j1 = je; k0 = ks; k1 = ke;
    if (MHD && b.f2) { j0 = js-1; j1 = je+1; if (b.f3) { k0 = ks-1; k1 = ke+1; } }
  } else if (DIR == 1) {
    i0 = is-1; i1 = ie+1; j0 = js; j1 = je+1; k0 = ks; k1 = ke;
    if (MHD && b.f3) { k0 = ks-1; k1 = ke+1; }
  } else {
    i0 = is; i1 = ie; j0 = js; j1 = je; k0 = ks; k1 = ke+1;
    if (MHD) { i0 = is-1; i1 = ie+1; j0 = js-1; j1 = je+1; }
  }
  int ni = i1-i0+1, nj = j1-j0+1, nk = k1-k0+1;
  if constexpr (ORDER == 3 && AB_PPM_SHARED) {
    flux_dir_ppm<DIR,SOLVER,MHD,NU>(b, g, p, i0, ni, j0, nj, k0, nk, dt_val, dt_ptr, s, nb);
  } else {
    int ntot = ni*nj*nk;
    FluxIdx fx;
    fx.ni = make_fastdiv(ni); fx.nj = make_fastdiv(nj);
    fx.per_full = make_fastdiv(ni*AB_X3_STRIP*nk); fx.per_k = make_fastdiv(ni*AB_X3_STRIP);
    fx.last_strip = (AB_X3_STRIP > 0 && nj % (AB_X3_STRIP > 0 ? AB_X3_STRIP : 1)) ? nj/(AB_X3_STRIP > 0 ? AB_X3_STRIP : 1) : -1;
    fx.per_k_last = make_fastdiv(ni*(AB_X3_STRIP > 0 ? nj % AB_X3_STRIP : 1));
    k_flux<DIR,ORDER,SOLVER,MHD,NU><<<dim3((unsigned)((ntot + AB_FLUX_BX - 1)/AB_FLUX_BX), (unsigned)nb), AB_FLUX_BX, 0, s>>>(
        b, g, p, i0, ni, j0, nj, k0, nk, ntot, dt_val, dt_ptr, fx); ++g_launches;
  }
}

template <int ORDER, int SOLVER, bool MHD, bool NU>
static void flux_all(const BlkDev &b, const ReconGeom &g, const Params &p, int dir,
                     double dt_val, const double *dt_ptr, cudaStream_t s, int nb) {
  if (dir == 0) flux_dir<0,ORDER,SOLVER,MHD,NU>(b, g, p, dt_val, dt_ptr, s, nb);
  else if (dir == 1) flux_dir<1,ORDER,SOLVER,MHD,NU>(b, g, p, dt_val, dt_ptr, s, nb);
  else flux_dir<2,ORDER,SOLVER,MHD,NU>(b, g, p, dt_val, dt_ptr, s, nb);
}

template <int SOLVER, bool MHD, bool NU>
static void flux_order(const BlkDev &b, const ReconGeom &g, const Params &p, int order, int dir,
                       double dt_val, const double *dt_ptr, cudaStream_t s, int nb = 1) {
  // order 4 / 5 = xorder 2c / 3c (characteristic variables; adiabatic EOS only)
  constexpr bool ISO = solver_is_iso<SOLVER>;
  if (order > 1 && p.char_proj && !ISO) order += 2;
  if (order == 1) {   // donor cell has no geometry: uniform instantiation only
    if constexpr (!NU) flux_all<1,SOLVER,MHD,false>(b, g, p, dir, dt_val, dt_ptr, s, nb);
  } else if (order == 2) flux_all<2,SOLVER,MHD,NU>(b, g, p, dir, dt_val, dt_ptr, s, nb);
  else if (order == 3) flux_all<3,SOLVER,MHD,NU>(b, g, p, dir, dt_val, dt_ptr, s, nb);
  else if (order == 4) flux_all<(ISO ? 2 : 4),SOLVER,MHD,NU>(b, g, p, dir, dt_val, dt_ptr, s, nb);
  else flux_all<(ISO ? 3 : 5),SOLVER,MHD,NU>(b, g, p, dir, dt_val, dt_ptr, s, nb);
}

template <bool NU>
static void launch_flux_dir_t(const BlkDev &b, const ReconGeom &g, const Params &p, int order,
                              int dir, double dt_val, const double *dt_ptr, cudaStream_t s,
                              int nb) {
  if (p.solver == SOLVER_LLF) {   // --flux=llf, either EOS
    if (p.eos != 0) {
      if (p.mhd) flux_order<SOLVER_LLF_ISO,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
      else flux_order<SOLVER_LLF_ISO,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    } else {
      if (p.mhd) flux_order<SOLVER_LLF,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
      else flux_order<SOLVER_LLF,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    }
  } else if (p.eos != 0) {   // isothermal: hlle / roe (hydro), hlle / hlld / roe (MHD) -- configure.py:299-325
    if (p.solver == SOLVER_ROE) {
      if (p.mhd) flux_order<SOLVER_ROE_ISO,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
      else flux_order<SOLVER_ROE_ISO,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    } else if (!p.mhd) flux_order<SOLVER_HLLE_ISO,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else if (p.solver == SOLVER_HLLD) flux_order<SOLVER_HLLD_ISO,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else flux_order<SOLVER_HLLE_ISO,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
  } else if (p.mhd) {
    if (p.solver == SOLVER_HLLD) flux_order<SOLVER_HLLD,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else if (p.solver == SOLVER_LHLLD) flux_order<SOLVER_LHLLD,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else if (p.solver == SOLVER_HLLE) flux_order<SOLVER_HLLE,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else flux_order<SOLVER_ROE,true,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
  } else {
    if (p.solver == SOLVER_HLLC) flux_order<SOLVER_HLLC,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else if (p.solver == SOLVER_LHLLC) flux_order<SOLVER_LHLLC,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else if (p.solver == SOLVER_HLLE) flux_order<SOLVER_HLLE,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
    else flux_order<SOLVER_ROE,false,NU>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
  }
}

}  // namespace ab
#endif
