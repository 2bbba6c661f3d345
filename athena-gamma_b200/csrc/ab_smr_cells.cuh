// ab_smr_cells.cuh -- per-cell bodies of the static-mesh-refinement kernels, __host__ __device__
// so that the very same code also runs on the CPU in tests/hostcheck (the product only calls
// them from the kernels in ab_smr_kernels.cu).  Reference: src/mesh/mesh_refinement.cpp:106-176,
// 386-540; src/bvals/bvals_refine.cpp:367-440; src/bvals/cc/flux_correction_cc.cpp:69-290.
#ifndef AB_SMR_CELLS_CUH_
#define AB_SMR_CELLS_CUH_
#include "ab_physics.cuh"

namespace ab {

// MeshRefinement::RestrictCellCenteredValues, one coarse cell, all variables
AB_HD void smr_restrict_cell(const SmrGeom &g, const double *fine, double *coarse, int nvar,
                             int ci, int cj, int ck) {
  const int nd = g.ndim;
  const int i = (ci - g.cis)*2 + g.is, j = nd > 1 ? (cj - g.cjs)*2 + g.js : 0,
            k = nd > 2 ? (ck - g.cks)*2 + g.ks : 0;
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int dk = 0; dk < (nd > 2 ? 2 : 1); ++dk) for (int dj = 0; dj < (nd > 1 ? 2 : 1); ++dj)
    for (int di = 0; di < 2; ++di)
      v[dk*4 + di*2 + dj] = g.dx1f[i+di]*g.dx2f[j+dj]*g.dx3f[k+dk];
  const long svf = (long)g.nc3*g.nc2*g.nc1, svc = (long)g.cnc3*g.cnc2*g.cnc1;
  for (int n = 0; n < nvar; ++n) {
    double f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int dk = 0; dk < (nd > 2 ? 2 : 1); ++dk) for (int dj = 0; dj < (nd > 1 ? 2 : 1); ++dj)
      for (int di = 0; di < 2; ++di)
        f[dk*4 + di*2 + dj] = fine[n*svf + ((long)(k+dk)*g.nc2 + (j+dj))*g.nc1 + (i+di)];
    coarse[n*svc + ((long)ck*g.cnc2 + cj)*g.cnc1 + ci] = restrict_cc(nd, f, v);
  }
}

// MeshRefinement::ProlongateCellCenteredValues, one coarse cell -> its 2^ndim fine cells
AB_HD void smr_prolong_cell(const SmrGeom &g, const double *coarse, double *fine, int nvar,
                            int i, int j, int k) {
  const int nd = g.ndim;
  const int fi = (i - g.cis)*2 + g.is, fj = nd > 1 ? (j - g.cjs)*2 + g.js : 0,
            fk = nd > 2 ? (k - g.cks)*2 + g.ks : 0;
  const long svf = (long)g.nc3*g.nc2*g.nc1, svc = (long)g.cnc3*g.cnc2*g.cnc1;
  const long s1 = 1, s2 = g.cnc1, s3 = (long)g.cnc2*g.cnc1;
  const double dx1m = g.cx1v[i] - g.cx1v[i-1], dx1p = g.cx1v[i+1] - g.cx1v[i];
  const double d1m = g.cx1v[i] - g.x1v[fi], d1p = g.x1v[fi+1] - g.cx1v[i];
  double dx2m = 0, dx2p = 0, d2m = 0, d2p = 0, dx3m = 0, dx3p = 0, d3m = 0, d3p = 0;
  if (nd > 1) {
    dx2m = g.cx2v[j] - g.cx2v[j-1]; dx2p = g.cx2v[j+1] - g.cx2v[j];
    d2m = g.cx2v[j] - g.x2v[fj]; d2p = g.x2v[fj+1] - g.cx2v[j];
  }
  if (nd > 2) {
    dx3m = g.cx3v[k] - g.cx3v[k-1]; dx3p = g.cx3v[k+1] - g.cx3v[k];
    d3m = g.cx3v[k] - g.x3v[fk]; d3p = g.x3v[fk+1] - g.cx3v[k];
  }
  for (int n = 0; n < nvar; ++n) {
    const double *c = coarse + n*svc + ((long)k*g.cnc2 + j)*g.cnc1 + i;
    const double cc = c[0];
    const double g1 = prolong_grad(c[-s1], cc, c[s1], dx1m, dx1p);
    const double g2 = nd > 1 ? prolong_grad(c[-s2], cc, c[s2], dx2m, dx2p) : 0.0;
    const double g3 = nd > 2 ? prolong_grad(c[-s3], cc, c[s3], dx3m, dx3p) : 0.0;
    double out[8];
    prolong_cc(nd, cc, g1, g2, g3, d1m, d1p, d2m, d2p, d3m, d3p, out);
    for (int dk = 0; dk < (nd > 2 ? 2 : 1); ++dk) for (int dj = 0; dj < (nd > 1 ? 2 : 1); ++dj)
      for (int di = 0; di < 2; ++di)
        fine[n*svf + ((long)(fk+dk)*g.nc2 + (fj+dj))*g.nc1 + (fi+di)] = out[dk*4 + di*2 + dj];
  }
}

// ConservedToPrimitive (+ PassiveScalarConservedToPrimitive) on the coarse buffers, floors
// written back (bvals_refine.cpp:367-383; eos/adiabatic_hydro.cpp:39-80, isothermal_hydro.cpp,
// eos_scalars.cpp:31-60)
AB_HD void smr_c2p_cell(const SmrGeom &g, const Params &p, double *cu, double *cw, int ns,
                        double *cs, double *cr, int i, int j, int k) {
  const long sv = (long)g.cnc3*g.cnc2*g.cnc1, o = ((long)k*g.cnc2 + j)*g.cnc1 + i;
  const bool iso = (p.eos != 0);
  double u_d = cu[o];
  const double u_m1 = cu[o+sv], u_m2 = cu[o+2*sv], u_m3 = cu[o+3*sv];
  u_d = (u_d > p.dfloor) ? u_d : p.dfloor;
  cu[o] = u_d;
  const double di = 1.0/u_d;
  cw[o] = u_d; cw[o+sv] = u_m1*di; cw[o+2*sv] = u_m2*di; cw[o+3*sv] = u_m3*di;
  if (!iso) {
    const double gm1 = p.gamma - 1.0;
    double u_e = cu[o+4*sv];
    const double e_k = 0.5*di*(sqr(u_m1) + sqr(u_m2) + sqr(u_m3));
    double w_p = gm1*(u_e - e_k);
    u_e = (w_p > p.pfloor) ? u_e : ((p.pfloor/gm1) + e_k);
    w_p = (w_p > p.pfloor) ? w_p : p.pfloor;
    cu[o+4*sv] = u_e;
    cw[o+4*sv] = w_p;
  }
  for (int n = 0; n < ns; ++n) {
    double s_n = cs[o + n*sv];
    s_n = (s_n < p.sfloor*u_d) ? p.sfloor*u_d : s_n;
    cs[o + n*sv] = s_n;
    cr[o + n*sv] = s_n/u_d;
  }
}

// outflow / reflecting boundary function on the coarse primitives, one ghost layer
// (ApplyPhysicalBoundariesOnCoarseLevel: DispatchBoundaryFunctions with ngh = 1); (i,j,k) runs
// over the transverse range, lo / hi are the last active coarse indices along the normal
AB_HD void smr_bc_cell(const SmrGeom &g, double *cw, int nh, double *cr, int ns, int face,
                       int refl, int lo, int hi, int i, int j, int k) {
  const int d = face >> 1, upper = face & 1;
  const long sv = (long)g.cnc3*g.cnc2*g.cnc1;
  const long st = d == 0 ? 1 : (d == 1 ? g.cnc1 : (long)g.cnc2*g.cnc1);
  int ijk[3] = {i, j, k};
  ijk[d] = upper ? hi : lo;
  const long src = ((long)ijk[2]*g.cnc2 + ijk[1])*g.cnc1 + ijk[0];
  const long dst = upper ? src + st : src - st;
  for (int n = 0; n < nh; ++n) {
    const double sign = (refl && n == 1 + d) ? -1.0 : 1.0;
    cw[n*sv + dst] = sign*cw[n*sv + src];
  }
  for (int n = 0; n < ns; ++n) cr[n*sv + dst] = cr[n*sv + src];
}

// LoadFluxBoundaryBufferToCoarser + SetFluxBoundaryFromFiner for one coarse face cell (ia, ib =
// coarse offsets along the two transverse directions, a fastest: dir 0 -> (j,k), 1 -> (i,k),
// 2 -> (i,j)): the area-weighted mean of the fine fluxes replaces the coarse flux
// compact_na > 0: the coarse values go into a message buffer instead (the fine block's rank sends
// them): element (n, ib, ia) at (n*compact_nb + ib)*compact_na + ia
AB_HD void smr_flux_cell(const SmrGeom &gf, const double *ffl, double *cfl, int nvar, int dir,
                         int fpos, int cpos, int a0, int b0, int ia, int ib, int compact_na = 0,
                         int compact_nb = 0) {
  const int ndim = gf.ndim;
  const int nc1 = gf.nc1, nc2 = gf.nc2, nc3 = gf.nc3;
  long sf_f, o00, sa, sb, o_c;
  double a00, a01, a10, a11;
  if (dir == 0) {
    const int j = (ndim > 1) ? gf.js + 2*ia : 0, k = (ndim > 2) ? gf.ks + 2*ib : 0;
    sf_f = (long)nc3*nc2*(nc1+1); sa = nc1 + 1; sb = (long)nc2*(nc1+1);
    o00 = ((long)k*nc2 + j)*(nc1+1) + fpos;
    o_c = ((long)(b0 + ib)*nc2 + (a0 + ia))*(nc1+1) + cpos;
    const double dj0 = gf.dx2f[j], dj1 = (ndim > 1) ? gf.dx2f[j+1] : 0.0;
    const double dk0 = gf.dx3f[k], dk1 = (ndim > 2) ? gf.dx3f[k+1] : 0.0;
    a00 = dj0*dk0; a01 = dj1*dk0; a10 = dj0*dk1; a11 = dj1*dk1;
  } else if (dir == 1) {
    const int i = gf.is + 2*ia, k = (ndim > 2) ? gf.ks + 2*ib : 0;
    sf_f = (long)nc3*(nc2+1)*nc1; sa = 1; sb = (long)(nc2+1)*nc1;
    o00 = ((long)k*(nc2+1) + fpos)*nc1 + i;
    o_c = ((long)(b0 + ib)*(nc2+1) + cpos)*nc1 + (a0 + ia);
    const double di0 = gf.dx1f[i], di1 = gf.dx1f[i+1];
    const double dk0 = gf.dx3f[k], dk1 = (ndim > 2) ? gf.dx3f[k+1] : 0.0;
    a00 = di0*dk0; a01 = di1*dk0; a10 = di0*dk1; a11 = di1*dk1;
  } else {
    const int i = gf.is + 2*ia, j = gf.js + 2*ib;
    sf_f = (long)(nc3+1)*nc2*nc1; sa = 1; sb = nc1;
    o00 = ((long)fpos*nc2 + j)*nc1 + i;
    o_c = ((long)cpos*nc2 + (b0 + ib))*nc1 + (a0 + ia);
    const double di0 = gf.dx1f[i], di1 = gf.dx1f[i+1];
    const double dj0 = gf.dx2f[j], dj1 = gf.dx2f[j+1];
    a00 = di0*dj0; a01 = di1*dj0; a10 = di0*dj1; a11 = di1*dj1;
  }
  // fine faces per coarse face: 4 in 3-D, 2 in 2-D, 1 in 1-D
  const int nfine = (ndim == 3) ? 4 : (ndim == 2 ? 2 : 1);
  for (int n = 0; n < nvar; ++n) {
    const double *f = ffl + n*sf_f + o00;
    double val;
    if (nfine == 4) {
      const double tarea = a00 + a01 + a10 + a11;
      val = (f[0]*a00 + f[sa]*a01 + f[sb]*a10 + f[sa+sb]*a11)/tarea;
    } else if (nfine == 2) {
      const double tarea = a00 + a01;
      val = (f[0]*a00 + f[sa]*a01)/tarea;
    } else {
      val = f[0];
    }
    if (compact_na > 0) cfl[((long)n*compact_nb + ib)*compact_na + ia] = val;
    else cfl[n*sf_f + o_c] = val;
  }
}

}  // namespace ab
#endif
