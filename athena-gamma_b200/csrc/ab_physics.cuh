// ab_physics.cuh -- point-wise device physics of the per-MeshBlock hydro/MHD update.
//
// Double precision, written for sm_100a and compiled with -fmad=false: every expression keeps
// the reference's operation order and parenthesisation so that results are bit-identical to
// the reference's default x86-64 (SSE2, no FMA) build.  The functions are __host__ __device__
// so that tests can also exercise the very same arithmetic on the CPU (tests/hostcheck); the
// product only ever calls them from CUDA kernels.
//
// Reference: src/eos/adiabatic_{hydro,mhd}.cpp, src/reconstruct/{dc,plm,ppm}.cpp,
// src/hydro/rsolvers/{hydro/{hllc,hlle,roe},mhd/{hlld,hlle_mhd,roe_mhd}}.cpp,
// src/hydro/hydro.cpp:154-158.
#ifndef AB_PHYSICS_CUH_
#define AB_PHYSICS_CUH_

#include <math.h>
#include <cmath>
#include "ab_types.h"

#if defined(__CUDACC__)
#define AB_HD __host__ __device__ __forceinline__
#else
#define AB_HD static inline
#endif

namespace ab {

// variable indices (src/athena.hpp:136-144)
enum : int { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4, IVX = 1, IVY = 2, IVZ = 3, IPR = 4,
             IBY = 5, IBZ = 6 };
enum : int { SOLVER_HLLE = 0, SOLVER_HLLC = 1, SOLVER_HLLD = 2, SOLVER_ROE = 3,
             SOLVER_LHLLC = 4, SOLVER_LHLLD = 5, SOLVER_LLF = 6 };

// std::min / std::max semantics of the reference
AB_HD double dmin(double a, double b) { return (b < a) ? b : a; }
AB_HD double dmax(double a, double b) { return (a < b) ? b : a; }
AB_HD double sqr(double x) { return x*x; }
AB_HD double sgn(double x) { return (x < 0.0) ? -1.0 : 1.0; }
AB_HD bool same_sgn(double a, double b) { return (a < 0.0) == (b < 0.0); }

// a/b, bit-identical to IEEE division.  The GPU's FP64 division falls into a ~100-instruction
// slow path when the dividend is zero or tiny; exactly-zero dividends are common on this path
// (static, uniform regions: zero mass flux, zero pressure jump), so they are answered
// directly with the correctly signed zero.  0/0, 0/NaN still go through the division.
AB_HD double fdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  // the common case (a != 0) costs one integer test on the ALU pipe, not FP64 compares
  if ((__double_as_longlong(a) << 1) == 0) {
    if (b == b && b != 0.0)
      return __longlong_as_double((__double_as_longlong(a) ^ __double_as_longlong(b)) &
                                  (long long)0x8000000000000000ull);
  }
#else   // the same shortcut when the tests compile this header for the host
  if (a == 0.0 && b == b && b != 0.0) return (std::signbit(a) != std::signbit(b)) ? -0.0 : 0.0;
#endif
  return a/b;
}

// EquationOfState::SoundSpeed (eos/adiabatic_hydro.cpp:125)
AB_HD double sound_speed(double gamma, double d, double p) { return sqrt(gamma*p/d); }

// EquationOfState::FastMagnetosonicSpeed (eos/adiabatic_mhd.cpp:149-156)
AB_HD double fast_speed(double gamma, double d, double p, double by, double bz, double bx) {
  double asq = gamma*p;
  double vaxsq = bx*bx;
  double ct2 = (by*by + bz*bz);
  double qsq = vaxsq + ct2 + asq;
  double tmp = vaxsq + ct2 - asq;
  return sqrt(0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2))/d);
}

// Hydro::GetWeightForCT (hydro/hydro.cpp:154-158)
AB_HD double weight_for_ct(double dflx, double rhol, double rhor, double dx, double dt) {
  double v_over_c = fdiv((1024.0)*dt*dflx, (dx*(rhol + rhor)));
  double tmp_min = dmin(0.5, v_over_c);
  return 0.5 + dmax(-0.5, tmp_min);
}

// the same with c1024dt = 1024*dt and dxrho = dx*(rhol + rhor) formed by the caller
AB_HD double weight_for_ct_pre(double dflx, double dxrho, double c1024dt) {
  double v_over_c = fdiv(c1024dt*dflx, dxrho);
  double tmp_min = dmin(0.5, v_over_c);
  return 0.5 + dmax(-0.5, tmp_min);
}

// ------------------------------------------------------------------------ reconstruction

// PLM, uniform Cartesian limiter (reconstruct/plm.cpp:69-77) with the coordinate face weights
// of plm.cpp:114-119: plus = state at the cell's upper face, minus = at its lower face.
AB_HD void plm(double qm1, double q, double qp1, double wp, double wm,
               double &plus, double &minus) {
  double dwl = (q - qm1);
  double dwr = (qp1 - q);
  double dw2 = dwl*dwr;
  double dwm = 0.0;                       // reference: computed, then zeroed when dw2 <= 0
  if (dw2 > 0.0) dwm = 2.0*dw2/(dwl + dwr);
  plus = q + wp*dwm;
  minus = q - wm*dwm;
}

// a/6.0, bit-identical to the IEEE quotient, in three FP64 instructions instead of the ~15 of
// a division: q0 = a*RN(1/6), exact residual r = a - 6*q0 by FMA, q = q0 + r*RN(1/6) by FMA
// (Markstein's correction; RN(1/6) is within 2^-54 of 1/6, so q0 is close enough for the one
// correction to round correctly -- checked against a/6.0 on 4e8 doubles, random and with
// significands at both ends of the binade, |a| in [1e-280, 1e300]).  Outside that range (and for
// zeros) the residual could underflow: true division there.
AB_HD double div6(double a) {
  const double fa = fabs(a);
  if (!(fa > 1.0e-280 && fa < 1.0e300)) return fdiv(a, 6.0);
  const double y = 1.0/6.0;
  const double q0 = a*y;
  const double r = fma(-6.0, q0, a);
  return fma(r, y, q0);
}

// PPM limiter pieces (reconstruct/ppm.cpp:143-190): CD eq 84-85 extremum fix of one interface.
// The reference evaluates the limited value everywhere and keeps it only at a local extremum
// (qa_tmp*qb_tmp < 0); here it is evaluated only where it is kept (same value, no division in
// monotone regions).
AB_HD double ppm_face_fix(double dph, double qlo, double qhi, double d2lo, double d2hi) {
  const double C2 = 1.25;
  double qa_tmp = dph - qlo;
  double qb_tmp = qhi - dph;
  if (qa_tmp*qb_tmp < 0.0) {
    double qa = 3.0*(qlo + qhi - 2.0*dph);
    double qb = d2lo;
    double qc = d2hi;
    double qd = 0.0;
    if (same_sgn(qa, qb) && same_sgn(qa, qc)) {
      qd = sgn(qa)*dmin(C2*fabs(qb), dmin(C2*fabs(qc), fabs(qa)));
    }
    dph = 0.5*(qlo + qhi) - div6(qd);
  }
  return dph;
}

// PPM for one variable of one cell, uniform Cartesian mesh (reconstruct/ppm.cpp:111-309;
// coefficients c1..c4=1/2, c5=1/6, c6=-1/6 from reconstruction.cpp:422-433)
AB_HD void ppm(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
               double &plus, double &minus) {
  const double C2 = 1.25;
  const double c1 = 0.5, c2 = 0.5, c3 = 0.5, c4 = 0.5, c5 = 1.0/6.0, c6 = -1.0/6.0;
  double qa = (q - q_im1);
  double qb = (q_ip1 - q);
  double dd_im1 = c1*qa + c2*(q_im1 - q_im2);
  double dd     = c1*qb + c2*qa;
  double dd_ip1 = c1*(q_ip2 - q_ip1) + c2*qb;
  double dph = (c3*q_im1 + c4*q) + (c5*dd_im1 + c6*dd);
  double dph_ip1 = (c3*q + c4*q_ip1) + (c5*dd + c6*dd_ip1);
  double d2qc_im1 = q_im2 + q     - 2.0*q_im1;
  double d2qc     = q_im1 + q_ip1 - 2.0*q;
  double d2qc_ip1 = q     + q_ip2 - 2.0*q_ip1;
  dph = ppm_face_fix(dph, q_im1, q, d2qc_im1, d2qc);
  dph_ip1 = ppm_face_fix(dph_ip1, q, q_ip1, d2qc, d2qc_ip1);
  double d2qf = 6.0*(dph + dph_ip1 - 2.0*q);
  double qminus = dph;
  double qplus = dph_ip1;
  double dqf_minus = q - qminus;
  double dqf_plus = qplus - q;
  double qa_tmp = dqf_minus*dqf_plus;
  double qb_tmp = (q_ip1 - q)*(q - q_im1);
  double qa2 = d2qc_im1, qb2 = d2qc, qc2 = d2qc_ip1, qd = d2qf;
  double qe = 0.0;
  if (same_sgn(qa2, qb2) && same_sgn(qa2, qc2) && same_sgn(qa2, qd)) {
    qe = sgn(qd)*dmin(dmin(C2*fabs(qa2), C2*fabs(qb2)), dmin(C2*fabs(qc2), fabs(qd)));
  }
  qa2 = dmax(fabs(q_im1), fabs(q_im2));
  qb2 = dmax(dmax(fabs(q), fabs(q_ip1)), fabs(q_ip2));
  double rho = 0.0;
  if (fabs(qd) > (1.0e-12)*dmax(qa2, qb2)) rho = qe/qd;
  double tmp_m = q - rho*dqf_minus;
  double tmp_p = q + rho*dqf_plus;
  double tmp2_m = q - 2.0*dqf_plus;
  double tmp2_p = q + 2.0*dqf_minus;
  if ((qa_tmp <= 0.0 || qb_tmp <= 0.0)) {
    if (rho <= (1.0 - (1.0e-12))) {
      qminus = tmp_m;
      qplus = tmp_p;
    }
  } else {
    if (fabs(dqf_minus) >= 2.0*fabs(dqf_plus)) qminus = tmp2_m;
    if (fabs(dqf_plus) >= 2.0*fabs(dqf_minus)) qplus = tmp2_p;
  }
  plus = qplus;
  minus = qminus;
}

// ------------------------------------------------------------------------ characteristic projection
// reconstruct/characteristic.cpp:36-278 (LeftEigenmatrixDotVector) and :280-520
// (RightEigenmatrixDotVector), adiabatic hydro and adiabatic MHD, in sweep order: w = the cell's
// primitives (IDN, vx, vy, vz, IPR[, By, Bz]) with vx along the sweep, bx = the cell-centred
// field along the sweep; vect is transformed in place.
struct MhdEig { double id, sqrtd, isqrtd, cf, cs, asq, a, bet2, bet3, alpha_f, alpha_s, s; };

AB_HD void mhd_eig(double gamma, const double *w, double bx, MhdEig *e) {
  e->id = 1.0/w[IDN];
  e->sqrtd = sqrt(w[IDN]);
  e->isqrtd = 1.0/e->sqrtd;
  double btsq = sqr(w[IBY]) + sqr(w[IBZ]);
  double bxsq = bx*bx;
  double gamp = gamma*w[IPR];
  double tdif = bxsq + btsq - gamp;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*gamp*btsq);
  double cfsq = 0.5*(bxsq + btsq + gamp + cf2_cs2);
  double cssq = gamp*bxsq/cfsq;
  cfsq *= e->id;
  e->cf = sqrt(cfsq);
  cssq *= e->id;
  e->cs = sqrt(cssq);
  e->asq = gamp*e->id;
  e->a = sqrt(e->asq);
  double bt = sqrt(btsq);
  e->bet2 = 0.0; e->bet3 = 0.0;
  if (bt != 0.0) { e->bet2 = w[IBY]/bt; e->bet3 = w[IBZ]/bt; }
  if ((cfsq - cssq) <= 0.0) { e->alpha_f = 1.0; e->alpha_s = 0.0; }
  else if ((e->asq - cssq) <= 0.0) { e->alpha_f = 0.0; e->alpha_s = 1.0; }
  else if ((cfsq - e->asq) <= 0.0) { e->alpha_f = 1.0; e->alpha_s = 0.0; }
  else {
    e->alpha_f = sqrt((e->asq - cssq)/(cfsq - cssq));
    e->alpha_s = sqrt((cfsq - e->asq)/(cfsq - cssq));
  }
  e->s = sgn(bx);
}

template <bool MHD>
AB_HD void char_left(double gamma, const double *w, double bx, double *vect) {
  if (MHD) {
    MhdEig e;
    mhd_eig(gamma, w, bx, &e);
    double id = e.id, cf = e.cf, cs = e.cs, asq = e.asq, a = e.a, bet2 = e.bet2, bet3 = e.bet3;
    double alpha_f = e.alpha_f, alpha_s = e.alpha_s, s = e.s, isqrtd = e.isqrtd, sqrtd = e.sqrtd;
    double nf = 0.5/asq;
    double qf = nf*cf*alpha_f*s;
    double qs = nf*cs*alpha_s*s;
    double af_prime = 0.5*alpha_f/(a*sqrtd);
    double as_prime = 0.5*alpha_s/(a*sqrtd);
    double v_0 = nf*alpha_f*(vect[IPR]*id - cf*vect[IVX]) +
                 qs*(bet2*vect[IVY] + bet3*vect[IVZ]) +
                 as_prime*(bet2*vect[IBY] + bet3*vect[IBZ]);
    double v_1 = 0.5*(bet2*(vect[IBZ]*s*isqrtd + vect[IVZ]) -
                      bet3*(vect[IBY]*s*isqrtd + vect[IVY]));
    double v_2 = nf*alpha_s*(vect[IPR]*id - cs*vect[IVX]) -
                 qf*(bet2*vect[IVY] + bet3*vect[IVZ]) -
                 af_prime*(bet2*vect[IBY] + bet3*vect[IBZ]);
    double v_3 = vect[IDN] - vect[IPR]/asq;
    double v_4 = nf*alpha_s*(vect[IPR]*id + cs*vect[IVX]) +
                 qf*(bet2*vect[IVY] + bet3*vect[IVZ]) -
                 af_prime*(bet2*vect[IBY] + bet3*vect[IBZ]);
    double v_5 = 0.5*(bet2*(vect[IBZ]*s*isqrtd - vect[IVZ]) -
                      bet3*(vect[IBY]*s*isqrtd - vect[IVY]));
    double v_6 = nf*alpha_f*(vect[IPR]*id + cf*vect[IVX]) -
                 qs*(bet2*vect[IVY] + bet3*vect[IVZ]) +
                 as_prime*(bet2*vect[IBY] + bet3*vect[IBZ]);
    vect[0] = v_0; vect[1] = v_1; vect[2] = v_2; vect[3] = v_3; vect[4] = v_4; vect[5] = v_5;
    vect[6] = v_6;
  } else {
    double asq = gamma*w[IPR]/w[IDN];
    double a = sqrt(asq);
    double v_0 = 0.5*(vect[IPR]/asq - w[IDN]*vect[IVX]/a);
    double v_1 = vect[IDN] - vect[IPR]/asq;
    double v_2 = vect[IVY];
    double v_3 = vect[IVZ];
    double v_4 = 0.5*(vect[IPR]/asq + w[IDN]*vect[IVX]/a);
    vect[0] = v_0; vect[1] = v_1; vect[2] = v_2; vect[3] = v_3; vect[4] = v_4;
  }
}

template <bool MHD>
AB_HD void char_right(double gamma, const double *w, double bx, double *vect) {
  if (MHD) {
    MhdEig e;
    mhd_eig(gamma, w, bx, &e);
    double cf = e.cf, cs = e.cs, asq = e.asq, a = e.a, bet2 = e.bet2, bet3 = e.bet3;
    double alpha_f = e.alpha_f, alpha_s = e.alpha_s, s = e.s, sqrtd = e.sqrtd;
    double qf = cf*alpha_f*s;
    double qs = cs*alpha_s*s;
    double af = a*alpha_f*sqrtd;
    double as = a*alpha_s*sqrtd;
    double v_0 = w[IDN]*(alpha_f*(vect[0] + vect[6]) +
                         alpha_s*(vect[2] + vect[4])) + vect[3];
    double v_1 = cf*alpha_f*(vect[6]-vect[0]) + cs*alpha_s*(vect[4]-vect[2]);
    double v_2 = bet2*(qs*(vect[0] - vect[6]) + qf*(vect[4] - vect[2]))
                 + bet3*(vect[5] - vect[1]);
    double v_3 = bet3*(qs*(vect[0] - vect[6]) + qf*(vect[4] - vect[2]))
                 + bet2*(vect[1] - vect[5]);
    double v_4 = w[IDN]*asq*(alpha_f*(vect[0] + vect[6]) +
                             alpha_s*(vect[2] + vect[4]));
    double v_5 = bet2*(as*(vect[0] + vect[6]) - af*(vect[2] + vect[4]))
                 - bet3*s*sqrtd*(vect[5] + vect[1]);
    double v_6 = bet3*(as*(vect[0] + vect[6]) - af*(vect[2] + vect[4]))
                 + bet2*s*sqrtd*(vect[5] + vect[1]);
    vect[IDN] = v_0; vect[IVX] = v_1; vect[IVY] = v_2; vect[IVZ] = v_3; vect[IPR] = v_4;
    vect[IBY] = v_5; vect[IBZ] = v_6;
  } else {
    double asq = gamma*w[IPR]/w[IDN];
    double a = sqrt(asq);
    double v_0 = vect[0] + vect[1] + vect[4];
    double v_1 = a*(vect[4] - vect[0])/w[IDN];
    double v_2 = vect[2];
    double v_3 = vect[3];
    double v_4 = asq*(vect[0] + vect[4]);
    vect[IDN] = v_0; vect[IVX] = v_1; vect[IVY] = v_2; vect[IVZ] = v_3; vect[IPR] = v_4;
  }
}

// ---- static mesh refinement: restriction and prolongation of cell-centred variables ----------
// MeshRefinement::RestrictCellCenteredValues (mesh/mesh_refinement.cpp:106-176): volume-weighted
// mean of the 2^ndim fine cells, with the reference's pairing of the sums.  f / v: fine values
// and volumes in the order (k,j,i), (k,j+1,i), (k,j,i+1), (k,j+1,i+1), then the same at k+1.
AB_HD double restrict_cc(int ndim, const double *f, const double *v) {
  if (ndim == 3) {
    const double tvol = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    return (((f[0]*v[0] + f[1]*v[1]) + (f[2]*v[2] + f[3]*v[3]))
            + ((f[4]*v[4] + f[5]*v[5]) + (f[6]*v[6] + f[7]*v[7])))/tvol;
  } else if (ndim == 2) {
    const double tvol = (v[0] + v[1]) + (v[2] + v[3]);
    return ((f[0]*v[0] + f[1]*v[1]) + (f[2]*v[2] + f[3]*v[3]))/tvol;
  }
  const double tvol = v[0] + v[2];
  return (f[0]*v[0] + f[2]*v[2])/tvol;
}

// minmod-limited gradient of MeshRefinement::ProlongateCellCenteredValues
// (mesh_refinement.cpp:430-445)
AB_HD double prolong_grad(double cm, double cc, double cp, double dxm, double dxp) {
  const double gm = (cc - cm)/dxm;
  const double gp = (cp - cc)/dxp;
  return 0.5*(sgn(gm) + sgn(gp))*dmin(fabs(gm), fabs(gp));
}

// the 2^ndim fine values of one coarse cell (mesh_refinement.cpp:447-456,497-501,533-534); out in
// the order of restrict_cc; g? limited gradients, d?m / d?p distances from the coarse centre to
// the lower / upper fine centres
AB_HD void prolong_cc(int ndim, double cc, double g1, double g2, double g3, double d1m, double d1p,
                      double d2m, double d2p, double d3m, double d3p, double *out) {
  if (ndim == 3) {
    out[0] = cc - (g1*d1m + g2*d2m + g3*d3m);
    out[2] = cc + (g1*d1p - g2*d2m - g3*d3m);
    out[1] = cc - (g1*d1m - g2*d2p + g3*d3m);
    out[3] = cc + (g1*d1p + g2*d2p - g3*d3m);
    out[4] = cc - (g1*d1m + g2*d2m - g3*d3p);
    out[6] = cc + (g1*d1p - g2*d2m + g3*d3p);
    out[5] = cc - (g1*d1m - g2*d2p - g3*d3p);
    out[7] = cc + (g1*d1p + g2*d2p + g3*d3p);
  } else if (ndim == 2) {
    out[0] = cc - (g1*d1m + g2*d2m);
    out[2] = cc + (g1*d1p - g2*d2m);
    out[1] = cc - (g1*d1m - g2*d2p);
    out[3] = cc + (g1*d1p + g2*d2p);
  } else {
    out[0] = cc - g1*d1m;
    out[2] = cc + g1*d1p;
  }
}

// ---- nonuniform (geometric, mesh/x?rat != 1) spacing -----------------------------------------
// Row `t` of the per-index geometry table the host builds (ab_mesh.cu: make_recon_table), NUG
// doubles per cell index along the sweep: dxf, dxv(i), dxv(i-1), cf, cb, dxf/dxv(i),
// dxf/dxv(i-1), c1..c6.  MODE = 1 + sweep direction: the three reference routines differ in
// rounding order (x1 multiplies by dx1f, then divides by dx1v: plm.cpp:85-86; x2 / x3 use the
// pre-divided ratios: plm.cpp:198-204,308-313) and x3 keeps the original van Leer expression
// (plm.cpp:314-318) where x1 / x2 use Mignone's (plm.cpp:94-96,207-209).  MODE 0 = uniform.
template <int MODE>
AB_HD double plm_slope(double dwl, double dwr, const double *t) {
  double dwm = 0.0;
  if (MODE == 0) {
    double dw2 = dwl*dwr;
    if (dw2 > 0.0) dwm = 2.0*dw2/(dwl + dwr);
    return dwm;
  }
  double dqF, dqB;
  if (MODE == 1) { dqF = dwr*t[0]/t[1]; dqB = dwl*t[0]/t[2]; }
  else { dqF = dwr*t[5]; dqB = dwl*t[6]; }
  double dq2 = dqF*dqB;
  if (dq2 > 0.0) {
    if (MODE == 3) dwm = 2.0*dq2/(dqF + dqB);
    else dwm = (dq2*(t[3]*dqB + t[4]*dqF)/(sqr(dqB) + sqr(dqF) + dq2*(t[3] + t[4] - 2.0)));
  }
  return dwm;
}

template <int MODE>
AB_HD void plm_nu(double qm1, double q, double qp1, double wp, double wm, const double *t,
                  double &plus, double &minus) {
  double dwm = plm_slope<MODE>(q - qm1, qp1 - q, t);
  plus = q + wp*dwm;
  minus = q - wm*dwm;
}

// PPM with nonuniform Cartesian spacing (ppm.cpp:111-129,196-207,282-300 and the x2 / x3 twins):
// CW84 interface values with the per-cell weights of reconstruction.cpp:434-461, strict
// monotonicity (Mignone eq 45), Mignone's parabola limiter with h ratios 2
// (reconstruction.cpp:412-419).  t = table row of the cell; rows of i-1 / i+1 are adjacent.
AB_HD void ppm_nu(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                  const double *t, double &plus, double &minus) {
  const double *tm = t - NUG, *tp = t + NUG;
  double qa = (q - q_im1);
  double qb = (q_ip1 - q);
  double dd_im1 = tm[7]*qa + tm[8]*(q_im1 - q_im2);
  double dd     = t[7]*qb + t[8]*qa;
  double dd_ip1 = tp[7]*(q_ip2 - q_ip1) + tp[8]*qb;
  double dph = (t[9]*q_im1 + t[10]*q) + (t[11]*dd_im1 + t[12]*dd);
  double dph_ip1 = (tp[9]*q + tp[10]*q_ip1) + (tp[11]*dd + tp[12]*dd_ip1);
  dph     = dmin(dph, dmax(q, q_im1));
  dph_ip1 = dmin(dph_ip1, dmax(q, q_ip1));
  dph     = dmax(dph, dmin(q, q_im1));
  dph_ip1 = dmax(dph_ip1, dmin(q, q_ip1));
  double dqf_minus = q - dph;
  double dqf_plus = dph_ip1 - q;
  double qminus = dph, qplus = dph_ip1;
  if (dqf_minus*dqf_plus <= 0.0) {
    qminus = q;
    qplus = q;
  } else {
    if (fabs(dqf_minus) >= 2.0*fabs(dqf_plus)) qminus = q - 2.0*dqf_plus;
    if (fabs(dqf_plus) >= 2.0*fabs(dqf_minus)) qplus = q + 2.0*dqf_minus;
  }
  plus = qplus;
  minus = qminus;
}

// xorder = 2c: PLM on characteristic variables for one cell (plm.cpp:62-66,107-130): both face
// states of the cell, floors re-applied.  q* sweep-ordered, NW = 5 / 7.
template <bool MHD, int MODE = 0>
AB_HD void plm_char(const double *qm1, const double *q, const double *qp1, double bx,
                    double gamma, double wp, double wm, double dfloor, double pfloor,
                    double *plus, double *minus, const double *t = nullptr) {
  constexpr int NW = MHD ? 7 : 5;
  double dwl[7], dwr[7], dwm[7];
  for (int n = 0; n < NW; ++n) { dwl[n] = (q[n] - qm1[n]); dwr[n] = (qp1[n] - q[n]); }
  char_left<MHD>(gamma, q, bx, dwl);
  char_left<MHD>(gamma, q, bx, dwr);
  for (int n = 0; n < NW; ++n) dwm[n] = plm_slope<MODE>(dwl[n], dwr[n], t);
  char_right<MHD>(gamma, q, bx, dwm);
  for (int n = 0; n < NW; ++n) { plus[n] = q[n] + wp*dwm[n]; minus[n] = q[n] - wm*dwm[n]; }
  plus[IDN] = (plus[IDN] > dfloor) ? plus[IDN] : dfloor;
  minus[IDN] = (minus[IDN] > dfloor) ? minus[IDN] : dfloor;
  plus[IPR] = (plus[IPR] > pfloor) ? plus[IPR] : pfloor;
  minus[IPR] = (minus[IPR] > pfloor) ? minus[IPR] : pfloor;
}

// xorder = 3c: PPM on characteristic variables for one cell (ppm.cpp:66-75,311-332): the five
// stencil states are projected with the cell's own eigenvectors, reconstructed, and both face
// states projected back; floors re-applied.
template <bool MHD, bool NU = false>
AB_HD void ppm_char(const double *qm2, const double *qm1, const double *q, const double *qp1,
                    const double *qp2, double bx, double gamma, double dfloor, double pfloor,
                    double *plus, double *minus, const double *t = nullptr) {
  constexpr int NW = MHD ? 7 : 5;
  double c0[7], c1[7], c2[7], c3[7], c4[7];
  for (int n = 0; n < NW; ++n) {
    c0[n] = qm2[n]; c1[n] = qm1[n]; c2[n] = q[n]; c3[n] = qp1[n]; c4[n] = qp2[n];
  }
  char_left<MHD>(gamma, q, bx, c0);
  char_left<MHD>(gamma, q, bx, c1);
  char_left<MHD>(gamma, q, bx, c2);
  char_left<MHD>(gamma, q, bx, c3);
  char_left<MHD>(gamma, q, bx, c4);
  for (int n = 0; n < NW; ++n) {
    if (NU) ppm_nu(c0[n], c1[n], c2[n], c3[n], c4[n], t, plus[n], minus[n]);
    else ppm(c0[n], c1[n], c2[n], c3[n], c4[n], plus[n], minus[n]);
  }
  char_right<MHD>(gamma, q, bx, plus);
  char_right<MHD>(gamma, q, bx, minus);
  plus[IDN] = (plus[IDN] > dfloor) ? plus[IDN] : dfloor;
  minus[IDN] = (minus[IDN] > dfloor) ? minus[IDN] : dfloor;
  plus[IPR] = (plus[IPR] > pfloor) ? plus[IPR] : pfloor;
  minus[IPR] = (minus[IPR] > pfloor) ? minus[IPR] : pfloor;
}

// ------------------------------------------------------------------------ hydro solvers
// wl/wr: sweep-ordered primitives (IDN, vx, vy, vz, IPR); f: (IDN, mx, my, mz, IEN).

// HLLC (hydro/rsolvers/hydro/hllc.cpp:32-179) and, with LOW, the low-dissipation LHLLC of
// Minoshima et al. 2021 (hydro/rsolvers/hydro/lhllc.cpp:26-175): shock detector th from the
// velocity differences dvn, dvt (hydro/calculate_velocity_differences.cpp) and the chi/phi
// pressure fix.
template <bool LOW>
AB_HD void hllc_t(const double *wli, const double *wri, double gamma, double dvn, double dvt,
                  double *flxi) {
  double gm1 = gamma - 1.0;
  double igm1 = 1.0/gm1;
  double cl = sound_speed(gamma, wli[IDN], wli[IPR]);
  double cr = sound_speed(gamma, wri[IDN], wri[IPR]);
  double vsql = sqr(wli[IVX]) + sqr(wli[IVY]) + sqr(wli[IVZ]);
  double vsqr = sqr(wri[IVX]) + sqr(wri[IVY]) + sqr(wri[IVZ]);
  double el = wli[IPR]*igm1 + 0.5*wli[IDN]*vsql;
  double er = wri[IPR]*igm1 + 0.5*wri[IDN]*vsqr;
  double rhoa = .5*(wli[IDN] + wri[IDN]);
  double ca = .5*(cl + cr);
  double pmid = .5*(wli[IPR] + wri[IPR] + (wli[IVX]-wri[IVX])*rhoa*ca);
  double ql = (pmid <= wli[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wli[IPR]-1.0));
  double qr = (pmid <= wri[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wri[IPR]-1.0));
  double al = wli[IVX] - cl*ql;
  double ar = wri[IVX] + cr*qr;
  double bp = ar > 0.0 ? ar : (1.0e-20);
  double bm = al < 0.0 ? al : -(1.0e-20);
  double vxl, vxr, am, cp;
  if (!LOW) {
    vxl = wli[IVX] - al;
    vxr = wri[IVX] - ar;
    double tl = wli[IPR] + vxl*wli[IDN]*wli[IVX];
    double tr = wri[IPR] + vxr*wri[IDN]*wri[IVX];
    double ml = wli[IDN]*vxl;
    double mr = -(wri[IDN]*vxr);
    am = fdiv((tl - tr), (ml + mr));
    cp = (ml*tr + mr*tl)/(ml + mr);
  } else {
    vxl = al - wli[IVX];
    vxr = ar - wri[IVX];
    double ml = wli[IDN]*vxl;
    double mr = wri[IDN]*vxr;
    double cmax = dmax(cl, cr);
    double th1 = dmin(1.0, (cmax-dmin(dvn,0.0))/(cmax-dmin(dvt,0.0)));
    double th = th1*th1*th1*th1;
    am = fdiv((mr*wri[IVX] - ml*wli[IVX] - th*(wri[IPR]-wli[IPR])), (mr - ml));
    double chi = dmin(1.0, sqrt(dmax(vsql, vsqr))/cmax);
    double phi = chi*(2.0 - chi);
    cp = (mr*wli[IPR] - ml*wri[IPR] + phi*mr*ml*(wri[IVX]-wli[IVX]))/(mr - ml);
  }
  cp = cp > 0.0 ? cp : 0.0;
  vxl = wli[IVX] - bm;
  vxr = wri[IVX] - bp;
  double fl0 = wli[IDN]*vxl, fr0 = wri[IDN]*vxr;
  double fl1 = wli[IDN]*wli[IVX]*vxl + wli[IPR], fr1 = wri[IDN]*wri[IVX]*vxr + wri[IPR];
  double fl2 = wli[IDN]*wli[IVY]*vxl, fr2 = wri[IDN]*wri[IVY]*vxr;
  double fl3 = wli[IDN]*wli[IVZ]*vxl, fr3 = wri[IDN]*wri[IVZ]*vxr;
  double fl4 = el*vxl + wli[IPR]*wli[IVX], fr4 = er*vxr + wri[IPR]*wri[IVX];
  double sl, sr, sm;
  if (am >= 0.0) {
    sl = fdiv(am, (am - bm));
    sr = 0.0;
    sm = -bm/(am - bm);
  } else {
    sl = 0.0;
    sr = fdiv(-am, (bp - am));
    sm = bp/(bp - am);
  }
  flxi[IDN] = sl*fl0 + sr*fr0;
  flxi[IVX] = sl*fl1 + sr*fr1 + sm*cp;
  flxi[IVY] = sl*fl2 + sr*fr2;
  flxi[IVZ] = sl*fl3 + sr*fr3;
  flxi[IEN] = sl*fl4 + sr*fr4 + sm*cp*am;
}

// HLLE, adiabatic hydro (hydro/rsolvers/hydro/hlle.cpp:38-162)
AB_HD void hlle_hydro(const double *wli, const double *wri, double gamma, double *flxi) {
  double gm1 = gamma - 1.0;
  double igm1 = 1.0/gm1;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  double rvx = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  double rvy = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  double rvz = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double el = wli[IPR]*igm1 + 0.5*wli[IDN]*(sqr(wli[IVX]) + sqr(wli[IVY]) + sqr(wli[IVZ]));
  double er = wri[IPR]*igm1 + 0.5*wri[IDN]*(sqr(wri[IVX]) + sqr(wri[IVY]) + sqr(wri[IVZ]));
  double hroe = ((el + wli[IPR])/sqrtdl + (er + wri[IPR])/sqrtdr)*isdlpdr;
  double cl = sound_speed(gamma, wli[IDN], wli[IPR]);
  double cr = sound_speed(gamma, wri[IDN], wri[IPR]);
  double q = hroe - 0.5*(sqr(rvx) + sqr(rvy) + sqr(rvz));
  double a = (q < 0.0) ? 0.0 : sqrt(gm1*q);
  double al = dmin((rvx - a), (wli[IVX] - cl));
  double ar = dmax((rvx + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  double fl[5], fr[5];
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl;
  fr[IVX] = wri[IDN]*wri[IVX]*vxr;
  fl[IVY] = wli[IDN]*wli[IVY]*vxl;
  fr[IVY] = wri[IDN]*wri[IVY]*vxr;
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl;
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr;
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = el*vxl + wli[IPR]*wli[IVX];
  fr[IEN] = er*vxr + wri[IPR]*wri[IVX];
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
#pragma unroll
  for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
}

// Roe, adiabatic hydro (hydro/rsolvers/hydro/roe.cpp:42-352)
AB_HD void roe_hydro(const double *wli, const double *wri, double gamma, double *flxi) {
  double gm1 = gamma - 1.0;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  double v1 = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  double v2 = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  double v3 = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double el = wli[IPR]/gm1 + 0.5*wli[IDN]*(sqr(wli[IVX]) + sqr(wli[IVY]) + sqr(wli[IVZ]));
  double er = wri[IPR]/gm1 + 0.5*wri[IDN]*(sqr(wri[IVX]) + sqr(wri[IVY]) + sqr(wri[IVZ]));
  double h = ((el + wli[IPR])/sqrtdl + (er + wri[IPR])/sqrtdr)*isdlpdr;
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  double fl[5], fr[5], du[5], ev[5];
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl*wli[IVX];
  fr[IVX] = mxr*wri[IVX];
  fl[IVY] = mxl*wli[IVY];
  fr[IVY] = mxr*wri[IVY];
  fl[IVZ] = mxl*wli[IVZ];
  fr[IVZ] = mxr*wri[IVZ];
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = (el + wli[IPR])*wli[IVX];
  fr[IEN] = (er + wri[IPR])*wri[IVX];
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  du[IEN] = er - el;
#pragma unroll
  for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  {  // RoeFlux (roe.cpp:206-290)
    double vsq = v1*v1 + v2*v2 + v3*v3;
    double q = h - 0.5*vsq;
    double cs_sq = (q < 0.0) ? (1.0e-20) : gm1*q;
    double cs = sqrt(cs_sq);
    ev[0] = v1 - cs;
    ev[1] = v1;
    ev[2] = v1;
    ev[3] = v1;
    ev[4] = v1 + cs;
    double a[5];
    double na = 0.5/cs_sq;
    a[0]  = du[0]*(0.5*gm1*vsq + v1*cs);
    a[0] -= du[1]*(gm1*v1 + cs);
    a[0] -= du[2]*gm1*v2;
    a[0] -= du[3]*gm1*v3;
    a[0] += du[4]*gm1;
    a[0] *= na;
    a[1]  = du[0]*(-v2);
    a[1] += du[2];
    a[2]  = du[0]*(-v3);
    a[2] += du[3];
    double qa = gm1/cs_sq;
    a[3]  = du[0]*(1.0 - na*gm1*vsq);
    a[3] += du[1]*qa*v1;
    a[3] += du[2]*qa*v2;
    a[3] += du[3]*qa*v3;
    a[3] -= du[4]*qa;
    a[4]  = du[0]*(0.5*gm1*vsq - v1*cs);
    a[4] -= du[1]*(gm1*v1 - cs);
    a[4] -= du[2]*gm1*v2;
    a[4] -= du[3]*gm1*v3;
    a[4] += du[4]*gm1;
    a[4] *= na;
    double coeff[5];
#pragma unroll
    for (int n = 0; n < 5; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];
    double dens = wli[IDN] + a[0];
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    flxi[0] += coeff[0];
    flxi[0] += coeff[3];
    flxi[0] += coeff[4];
    flxi[1] += coeff[0]*(v1 - cs);
    flxi[1] += coeff[3]*v1;
    flxi[1] += coeff[4]*(v1 + cs);
    flxi[2] += coeff[0]*v2;
    flxi[2] += coeff[1];
    flxi[2] += coeff[3]*v2;
    flxi[2] += coeff[4]*v2;
    flxi[3] += coeff[0]*v3;
    flxi[3] += coeff[2];
    flxi[3] += coeff[3]*v3;
    flxi[3] += coeff[4]*v3;
    flxi[4] += coeff[0]*(h - v1*cs);
    flxi[4] += coeff[1]*v2;
    flxi[4] += coeff[2]*v3;
    flxi[4] += coeff[3]*0.5*vsq;
    flxi[4] += coeff[4]*(h + v1*cs);
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) flxi[n] = fl[n];
  }
  if (ev[4] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 5; ++n) flxi[n] = fr[n];
  }
  if (llf_flag != 0) {
    double cl = sound_speed(gamma, wli[IDN], wli[IPR]);
    double cr = sound_speed(gamma, wri[IDN], wri[IPR]);
    double a = 0.5*dmax((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
#pragma unroll
    for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
}

// ------------------------------------------------------------------------ MHD solvers
// wl/wr: (IDN, vx, vy, vz, IPR, By, Bz); f: (IDN, mx, my, mz, IEN, F(By), F(Bz)).

struct Cons1D { double d, mx, my, mz, e, by, bz; };

// c ? a : b on doubles (two 32-bit selects on the ALU pipe; never a branch)
AB_HD double dsel(bool c, double a, double b) { return c ? a : b; }

// HLLD (hydro/rsolvers/mhd/hlld.cpp:38-382).
// The reference evaluates both outer states, both star and both double-star states and both
// L/R fluxes, then selects one of six results (hlld.cpp:315-369).  Here the side of the
// contact the interface lies on is decided as soon as the contact speed spd2 is known, the
// states of that side ("own") and the few transverse values of the other side ("oth") the
// double-star state needs are picked with selects, and ONE copy of the star / double-star
// algebra runs on them.  Every expression is the reference's, with the same operands in the
// same order (the L and R formulas of hlld.cpp are textually identical up to the L/R suffix;
// the few that are not -- the double-star averages, which always read "left, then right" --
// are evaluated on re-selected L/R views), so the result is bit-identical, a warp whose
// threads fall on both sides of the contact does not run two code paths, and the live register
// set after the selection is one state, not two.
// Side selection: the reference takes the left family when spd1 >= 0 or spd2 >= 0
// (hlld.cpp:327-343); spd1 = spd2 - |Bx|/sqrt(rho*_L) <= spd2, so that is spd2 >= 0 (it could
// differ only if rho*_L were exactly -0, where the reference's result is NaN anyway).
// With LOW: the low-dissipation LHLLD (hydro/rsolvers/mhd/lhlld.cpp:36-390).
template <bool LOW>
AB_HD void hlld_t(const double *wli, const double *wri, double bxi, double gamma, double dvn,
                  double dvt, double *flxi) {
  const double SMALL_NUMBER = 1.0e-8;
  const double igm1 = 1.0/(gamma - 1.0);
  const double bxsq = bxi*bxi;
  const double pbl = 0.5*(bxsq + (sqr(wli[IBY]) + sqr(wli[IBZ])));
  const double pbr = 0.5*(bxsq + (sqr(wri[IBY]) + sqr(wri[IBZ])));
  const double cfl = fast_speed(gamma, wli[IDN], wli[IPR], wli[IBY], wli[IBZ], bxi);
  const double cfr = fast_speed(gamma, wri[IDN], wri[IPR], wri[IBY], wri[IBZ], bxi);
  const double spd0 = dmin(wli[IVX]-cfl, wri[IVX]-cfr);
  const double spd4 = dmax(wli[IVX]+cfl, wri[IVX]+cfr);
  const double ptl = wli[IPR] + pbl;
  const double ptr = wri[IPR] + pbr;
  const double sdl = spd0 - wli[IVX];
  const double sdr = spd4 - wri[IVX];
  const double sdld = sdl*wli[IDN], sdrd = sdr*wri[IDN];
  const double mxl = wli[IVX]*wli[IDN], mxr = wri[IVX]*wri[IDN];
  double spd2, cfmax = 0.0;
  if (!LOW) {
    spd2 = fdiv((sdr*mxr - sdl*mxl + (ptl - ptr)), (sdrd - sdld));
  } else {
    cfmax = dmax(cfl, cfr);
    double th1 = dmin(1.0, (cfmax-dmin(dvn,0.0))/(cfmax-dmin(dvt,0.0)));
    double th = th1*th1*th1*th1;
    spd2 = fdiv((sdr*mxr - sdl*mxl + th*(ptl - ptr)), (sdrd - sdld));
  }
  const bool sup_l = (spd0 >= 0.0);
  const bool sup_r = !sup_l && (spd4 <= 0.0);
  const bool left = sup_l || (!sup_r && (spd2 >= 0.0));

  // total pressure of the star region (both outer states enter; hlld.cpp:175-178)
  double ptst;
  if (!LOW) {
    double ptstl = ptl + sdld*(spd2-wli[IVX]);
    double ptstr = ptr + sdrd*(spd2-wri[IVX]);
    ptst = 0.5*(ptstr + ptstl);
  } else {
    double kel = 0.5*wli[IDN]*(sqr(wli[IVX]) + (sqr(wli[IVY]) + sqr(wli[IVZ])));
    double ker = 0.5*wri[IDN]*(sqr(wri[IVX]) + (sqr(wri[IVY]) + sqr(wri[IVZ])));
    double clsq = ((pbl + kel) + sqrt(sqr(pbl + kel) - 2.0*kel*bxsq))/wli[IDN];
    double crsq = ((pbr + ker) + sqrt(sqr(pbr + ker) - 2.0*ker*bxsq))/wri[IDN];
    double chi = dmin(1.0, sqrt(dmax(clsq, crsq))/cfmax);
    double phi = chi*(2.0 - chi);
    ptst = (sdrd*ptl - sdld*ptr + phi*sdrd*sdld*(wri[IVX]-wli[IVX]))/(sdrd - sdld);
  }

  // ---- own side: outer state, its flux
  const double od = dsel(left, wli[IDN], wri[IDN]);
  const double ovx = dsel(left, wli[IVX], wri[IVX]);
  const double ovy = dsel(left, wli[IVY], wri[IVY]);
  const double ovz = dsel(left, wli[IVZ], wri[IVZ]);
  const double opr = dsel(left, wli[IPR], wri[IPR]);
  const double oby = dsel(left, wli[IBY], wri[IBY]);
  const double obz = dsel(left, wli[IBZ], wri[IBZ]);
  const double pb = dsel(left, pbl, pbr);
  const double pt = dsel(left, ptl, ptr);
  const double sd = dsel(left, sdl, sdr);
  const double sdd = dsel(left, sdld, sdrd);
  const double so = dsel(left, spd0, spd4);       // outer wave speed of this side
  const double umx = dsel(left, mxl, mxr);
  const double umy = ovy*od;
  const double umz = ovz*od;
  const double ke = 0.5*od*(sqr(ovx) + (sqr(ovy) + sqr(ovz)));
  const double ue = opr*igm1 + ke + pb;
  const double f_d  = umx;
  const double f_mx = umx*ovx + pt - bxsq;
  const double f_my = umy*ovx - bxi*oby;
  const double f_mz = umz*ovx - bxi*obz;
  const double f_e  = ovx*(ue + pt - bxsq) - bxi*(ovy*oby + ovz*obz);
  const double f_by = oby*ovx - bxi*ovy;
  const double f_bz = obz*ovx - bxi*ovz;
  if (sup_l || sup_r) {
    flxi[IDN] = f_d; flxi[IVX] = f_mx; flxi[IVY] = f_my; flxi[IVZ] = f_mz;
    flxi[IEN] = f_e; flxi[IBY] = f_by; flxi[IBZ] = f_bz;
    return;
  }

  // ---- own side: star state (hlld.cpp:147-233)
  const double sdm = so - spd2;
  const double sdm_inv = 1.0/sdm;
  const double st_d = sdd*sdm_inv;
  const double st_d_inv = 1.0/st_d;
  const double sqrtd = sqrt(st_d);
  const double xa = fdiv(fabs(bxi), sqrtd);
  const double spdi = spd2 - dsel(left, xa, -xa);     // spd1 = spd2 - xa | spd3 = spd2 + xa
  // hlld.cpp:327-369: left family: U*_L if spd1 >= 0 else U**_L; right: U**_R if spd3 > 0 else U*_R
  const bool star = left ? (spdi >= 0.0) : !(spdi > 0.0);
  const double st_mx = st_d*spd2;
  double st_my, st_mz, st_by, st_bz;
  {
    const double den = sdd*sdm - bxsq;
    if (fabs(den) < (SMALL_NUMBER)*ptst) {
      st_my = st_d*ovy;
      st_mz = st_d*ovz;
      st_by = oby;
      st_bz = obz;
    } else {
      double tmp = fdiv(bxi*(sd - sdm), den);
      st_my = st_d*(ovy - oby*tmp);
      st_mz = st_d*(ovz - obz*tmp);
      tmp = ((LOW ? sdd*sd : od*sqr(sd)) - bxsq)/den;
      st_by = oby*tmp;
      st_bz = obz*tmp;
    }
  }
  const double vbst = (st_mx*bxi+(st_my*st_by+st_mz*st_bz))*st_d_inv;
  const double st_e = (sd*ue - pt*ovx + ptst*spd2 +
                       bxi*(ovx*bxi + (ovy*oby + ovz*obz) - vbst))*sdm_inv;
  if (star) {
    flxi[IDN] = f_d  + so*(st_d - od);
    flxi[IVX] = f_mx + so*(st_mx - umx);
    flxi[IVY] = f_my + so*(st_my - umy);
    flxi[IVZ] = f_mz + so*(st_mz - umz);
    flxi[IEN] = f_e  + so*(st_e - ue);
    flxi[IBY] = f_by + so*(st_by - oby);
    flxi[IBZ] = f_bz + so*(st_bz - obz);
    return;
  }

  // ---- double-star state of this side (hlld.cpp:239-281)
  double ds_my = st_my, ds_mz = st_mz, ds_by = st_by, ds_bz = st_bz, ds_e = st_e;
  if (!(0.5*bxsq < (SMALL_NUMBER)*ptst)) {
    // transverse star state of the other side
    const double td = dsel(left, wri[IDN], wli[IDN]);
    const double tvy = dsel(left, wri[IVY], wli[IVY]);
    const double tvz = dsel(left, wri[IVZ], wli[IVZ]);
    const double tby = dsel(left, wri[IBY], wli[IBY]);
    const double tbz = dsel(left, wri[IBZ], wli[IBZ]);
    const double tsd = dsel(left, sdr, sdl);
    const double tsdd = dsel(left, sdrd, sdld);
    const double tsdm = dsel(left, spd4, spd0) - spd2;
    const double tsdm_inv = 1.0/tsdm;
    const double tst_d = tsdd*tsdm_inv;
    const double tst_d_inv = 1.0/tst_d;
    const double tsqrtd = sqrt(tst_d);
    double tst_my, tst_mz, tst_by, tst_bz;
    {
      const double den = tsdd*tsdm - bxsq;
      if (fabs(den) < (SMALL_NUMBER)*ptst) {
        tst_my = tst_d*tvy;
        tst_mz = tst_d*tvz;
        tst_by = tby;
        tst_bz = tbz;
      } else {
        double tmp = fdiv(bxi*(tsd - tsdm), den);
        tst_my = tst_d*(tvy - tby*tmp);
        tst_mz = tst_d*(tvz - tbz*tmp);
        tmp = ((LOW ? tsdd*tsd : td*sqr(tsd)) - bxsq)/den;
        tst_by = tby*tmp;
        tst_bz = tbz*tmp;
      }
    }
    // L / R views: the averages below read "left, then right" whichever side is ours
    const double l_d = dsel(left, st_d, tst_d), r_d = dsel(left, tst_d, st_d);
    const double l_di = dsel(left, st_d_inv, tst_d_inv), r_di = dsel(left, tst_d_inv, st_d_inv);
    const double sqrtdl = dsel(left, sqrtd, tsqrtd), sqrtdr = dsel(left, tsqrtd, sqrtd);
    const double l_my = dsel(left, st_my, tst_my), r_my = dsel(left, tst_my, st_my);
    const double l_mz = dsel(left, st_mz, tst_mz), r_mz = dsel(left, tst_mz, st_mz);
    const double l_by = dsel(left, st_by, tst_by), r_by = dsel(left, tst_by, st_by);
    const double l_bz = dsel(left, st_bz, tst_bz), r_bz = dsel(left, tst_bz, st_bz);
    (void)r_d;
    const double invsumd = 1.0/(sqrtdl + sqrtdr);
    const double bxsig = (bxi > 0.0 ? 1.0 : -1.0);
    double tmp = invsumd*(sqrtdl*(l_my*l_di) + sqrtdr*(r_my*r_di) + bxsig*(r_by - l_by));
    const double uldst_my = l_d*tmp;
    ds_my = st_d*tmp;
    tmp = invsumd*(sqrtdl*(l_mz*l_di) + sqrtdr*(r_mz*r_di) + bxsig*(r_bz - l_bz));
    const double uldst_mz = l_d*tmp;
    ds_mz = st_d*tmp;
    ds_by = invsumd*(sqrtdl*r_by + sqrtdr*l_by +
                     bxsig*sqrtdl*sqrtdr*((r_my*r_di) - (l_my*l_di)));
    ds_bz = invsumd*(sqrtdl*r_bz + sqrtdr*l_bz +
                     bxsig*sqrtdl*sqrtdr*((r_mz*r_di) - (l_mz*l_di)));
    // hlld.cpp:276-279: both energies use the LEFT double-star momenta
    tmp = spd2*bxi + fdiv((uldst_my*ds_by + uldst_mz*ds_bz), l_d);
    const double pe = sqrtd*bxsig*(vbst - tmp);
    ds_e = st_e - dsel(left, pe, -pe);   // uldst.e = ulst.e - ..., urdst.e = urst.e + ...
  }
  flxi[IDN] = f_d  + so*(st_d - od) + spdi*(st_d - st_d);
  flxi[IVX] = f_mx + so*(st_mx - umx) + spdi*(st_mx - st_mx);
  flxi[IVY] = f_my + so*(st_my - umy) + spdi*(ds_my - st_my);
  flxi[IVZ] = f_mz + so*(st_mz - umz) + spdi*(ds_mz - st_mz);
  flxi[IEN] = f_e  + so*(st_e - ue) + spdi*(ds_e - st_e);
  flxi[IBY] = f_by + so*(st_by - oby) + spdi*(ds_by - st_by);
  flxi[IBZ] = f_bz + so*(st_bz - obz) + spdi*(ds_bz - st_bz);
}

// Roe averages shared by hlle_mhd.cpp:64-85 and roe_mhd.cpp:92-113
struct RoeAvgMHD { double d, v1, v2, v3, b2, b3, x, y, pbl, pbr, el, er, h; };

AB_HD void roe_avg_mhd(const double *wli, const double *wri, double bxi, double gm1,
                       RoeAvgMHD &r) {
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  r.d = sqrtdl*sqrtdr;
  r.v1 = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  r.v2 = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  r.v3 = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  r.b2 = (sqrtdr*wli[IBY] + sqrtdl*wri[IBY])*isdlpdr;
  r.b3 = (sqrtdr*wli[IBZ] + sqrtdl*wri[IBZ])*isdlpdr;
  r.x = 0.5*(sqr(wli[IBY]-wri[IBY]) + sqr(wli[IBZ]-wri[IBZ]))/(sqr(sqrtdl+sqrtdr));
  r.y = 0.5*(wli[IDN] + wri[IDN])/r.d;
  r.pbl = 0.5*(bxi*bxi + sqr(wli[IBY]) + sqr(wli[IBZ]));
  r.pbr = 0.5*(bxi*bxi + sqr(wri[IBY]) + sqr(wri[IBZ]));
  r.el = wli[IPR]/gm1 + 0.5*wli[IDN]*(sqr(wli[IVX])+sqr(wli[IVY])+sqr(wli[IVZ])) + r.pbl;
  r.er = wri[IPR]/gm1 + 0.5*wri[IDN]*(sqr(wri[IVX])+sqr(wri[IVY])+sqr(wri[IVZ])) + r.pbr;
  r.h = ((r.el + wli[IPR] + r.pbl)/sqrtdl + (r.er + wri[IPR] + r.pbr)/sqrtdr)*isdlpdr;
}

// HLLE, adiabatic MHD (hydro/rsolvers/mhd/hlle_mhd.cpp:25-182)
AB_HD void hlle_mhd(const double *wli, const double *wri, double bxi, double gamma,
                    double *flxi) {
  double gm1 = gamma - 1.0;
  RoeAvgMHD r;
  roe_avg_mhd(wli, wri, bxi, gm1, r);
  double cl = fast_speed(gamma, wli[IDN], wli[IPR], wli[IBY], wli[IBZ], bxi);
  double cr = fast_speed(gamma, wri[IDN], wri[IPR], wri[IBY], wri[IBZ], bxi);
  double btsq = sqr(r.b2) + sqr(r.b3);
  double vaxsq = bxi*bxi/r.d;
  double bt_starsq = (gm1 - (gm1 - 1.0)*r.y)*btsq;
  double hp = r.h - (vaxsq + btsq/r.d);
  double vsq = sqr(r.v1) + sqr(r.v2) + sqr(r.v3);
  double twid_asq = dmax((gm1*(hp-0.5*vsq)-(gm1-1.0)*r.x), 0.0);
  double ct2 = bt_starsq/r.d;
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_asq*ct2);
  double cfsq = 0.5*(tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = dmin((r.v1 - a), (wli[IVX] - cl));
  double ar = dmax((r.v1 + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  double fl[7], fr[7];
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + r.pbl - sqr(bxi);
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + r.pbr - sqr(bxi);
  fl[IVY] = wli[IDN]*wli[IVY]*vxl - bxi*wli[IBY];
  fr[IVY] = wri[IDN]*wri[IVY]*vxr - bxi*wri[IBY];
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl - bxi*wli[IBZ];
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr - bxi*wri[IBZ];
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = r.el*vxl + wli[IVX]*(wli[IPR] + r.pbl - bxi*bxi);
  fr[IEN] = r.er*vxr + wri[IVX]*(wri[IPR] + r.pbr - bxi*bxi);
  fl[IEN] -= bxi*(wli[IBY]*wli[IVY] + wli[IBZ]*wli[IVZ]);
  fr[IEN] -= bxi*(wri[IBY]*wri[IVY] + wri[IBZ]*wri[IVZ]);
  fl[IBY] = wli[IBY]*vxl - bxi*wli[IVY];
  fr[IBY] = wri[IBY]*vxr - bxi*wri[IVY];
  fl[IBZ] = wli[IBZ]*vxl - bxi*wli[IVZ];
  fr[IBZ] = wri[IBZ]*vxr - bxi*wri[IVZ];
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
#pragma unroll
  for (int n = 0; n < 7; ++n) flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
}

// RoeFlux, adiabatic MHD (hydro/rsolvers/mhd/roe_mhd.cpp:245-470)
AB_HD void roe_flux_mhd(const RoeAvgMHD &r, double b1, const double *du, double wl_d,
                        double gm1, double *flx, double &ev0, double &ev6, int &llf_flag) {
  double d = r.d, v1 = r.v1, v2 = r.v2, v3 = r.v3, b2 = r.b2, b3 = r.b3, x = r.x, y = r.y;
  double di = 1.0/d;
  double btsq = b2*b2 + b3*b3;
  double vaxsq = b1*b1*di;
  double vsq = v1*v1 + v2*v2 + v3*v3;
  double hp = r.h - (vaxsq + btsq*di);
  double bt_starsq = (gm1 - (gm1 - 1.0)*y)*btsq;
  double twid_csq = dmax((gm1*(hp-0.5*vsq)-(gm1-1.0)*x), 1.0e-20);
  double ct2 = bt_starsq*di;
  double tsum = vaxsq + ct2 + twid_csq;
  double tdif = vaxsq + ct2 - twid_csq;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_csq*ct2);
  double cfsq = 0.5*(tsum + cf2_cs2);
  double cf = sqrt(cfsq);
  double cssq = twid_csq*vaxsq/cfsq;
  double cs = sqrt(cssq);
  double bt = sqrt(btsq);
  double bt_star = sqrt(bt_starsq);
  double bet2 = 0.0;
  double bet3 = 0.0;
  if (bt != 0.0) {
    bet2 = b2/bt;
    bet3 = b3/bt;
  }
  double bet2_star = bet2/sqrt(gm1 - (gm1-1.0)*y);
  double bet3_star = bet3/sqrt(gm1 - (gm1-1.0)*y);
  double bet_starsq = bet2_star*bet2_star + bet3_star*bet3_star;
  double vbet = v2*bet2_star + v3*bet3_star;
  double q2_star = 0.0;
  double q3_star = 0.0;
  if (bet_starsq != 0.0) {
    q2_star = bet2_star/bet_starsq;
    q3_star = bet3_star/bet_starsq;
  }
  double alpha_f, alpha_s;
  if ((cfsq - cssq) <= 0.0) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else if ((twid_csq - cssq) <= 0.0) {
    alpha_f = 0.0;
    alpha_s = 1.0;
  } else if ((cfsq - twid_csq) <= 0.0) {
    alpha_f = 1.0;
    alpha_s = 0.0;
  } else {
    alpha_f = sqrt((twid_csq - cssq)/(cfsq - cssq));
    alpha_s = sqrt((cfsq - twid_csq)/(cfsq - cssq));
  }
  double sqrtd = sqrt(d);
  double isqrtd = 1.0/sqrtd;
  double s = sgn(b1);
  double twid_c = sqrt(twid_csq);
  double qf = cf*alpha_f*s;
  double qs = cs*alpha_s*s;
  double af_prime = twid_c*alpha_f*isqrtd;
  double as_prime = twid_c*alpha_s*isqrtd;
  double afpbb = af_prime*bt_star*bet_starsq;
  double aspbb = as_prime*bt_star*bet_starsq;
  double vqstr = (v2*q2_star + v3*q3_star);
  double vax = sqrt(vaxsq);
  double norm = 0.5/twid_csq;
  double cff = norm*alpha_f*cf;
  double css = norm*alpha_s*cs;
  double qf_hat = qf*norm;
  double qs_hat = qs*norm;
  double af = norm*af_prime*d;
  double as = norm*as_prime*d;
  double afpb = norm*af_prime*bt_star;
  double aspb = norm*as_prime*bt_star;

  double ev[7];
  ev[0] = v1 - cf;
  ev[1] = v1 - vax;
  ev[2] = v1 - cs;
  ev[3] = v1;
  ev[4] = v1 + cs;
  ev[5] = v1 + vax;
  ev[6] = v1 + cf;
  ev0 = ev[0];
  ev6 = ev[6];

  double a[7];
  double alpha_f_bar = alpha_f*gm1*norm;
  double alpha_s_bar = alpha_s*gm1*norm;
  double gm1a = gm1/twid_csq;

  a[0]  = du[0]*(alpha_f_bar*(vsq-hp) + cff*(cf+v1) - qs_hat*vqstr - aspb);
  a[0] -= du[1]*(alpha_f_bar*v1 + cff);
  a[0] -= du[2]*(alpha_f_bar*v2 - qs_hat*q2_star);
  a[0] -= du[3]*(alpha_f_bar*v3 - qs_hat*q3_star);
  a[0] += du[4]*alpha_f_bar;
  a[0] += du[5]*(as*q2_star - alpha_f_bar*b2);
  a[0] += du[6]*(as*q3_star - alpha_f_bar*b3);

  a[1]  = du[0]*(v2*bet3 - v3*bet2);
  a[1] -= du[2]*bet3;
  a[1] += du[3]*bet2;
  a[1] -= du[5]*sqrtd*bet3*s;
  a[1] += du[6]*sqrtd*bet2*s;
  a[1] *= 0.5;

  a[2]  = du[0]*(alpha_s_bar*(vsq-hp) + css*(cs+v1) + qf_hat*vqstr + afpb);
  a[2] -= du[1]*(alpha_s_bar*v1 + css);
  a[2] -= du[2]*(alpha_s_bar*v2 + qf_hat*q2_star);
  a[2] -= du[3]*(alpha_s_bar*v3 + qf_hat*q3_star);
  a[2] += du[4]*alpha_s_bar;
  a[2] -= du[5]*(af*q2_star + alpha_s_bar*b2);
  a[2] -= du[6]*(af*q3_star + alpha_s_bar*b3);

  a[3]  = du[0]*(1.0 - gm1a*(0.5*vsq - (gm1-1.0)*x/gm1));
  a[3] += du[1]*gm1a*v1;
  a[3] += du[2]*gm1a*v2;
  a[3] += du[3]*gm1a*v3;
  a[3] -= du[4]*gm1a;
  a[3] += du[5]*gm1a*b2;
  a[3] += du[6]*gm1a*b3;

  a[4]  = du[0]*(alpha_s_bar*(vsq-hp) + css*(cs-v1) - qf_hat*vqstr + afpb);
  a[4] -= du[1]*(alpha_s_bar*v1 - css);
  a[4] -= du[2]*(alpha_s_bar*v2 - qf_hat*q2_star);
  a[4] -= du[3]*(alpha_s_bar*v3 - qf_hat*q3_star);
  a[4] += du[4]*alpha_s_bar;
  a[4] -= du[5]*(af*q2_star + alpha_s_bar*b2);
  a[4] -= du[6]*(af*q3_star + alpha_s_bar*b3);

  a[5]  = du[0]*(v3*bet2 - v2*bet3);
  a[5] += du[2]*bet3;
  a[5] -= du[3]*bet2;
  a[5] -= du[5]*sqrtd*bet3*s;
  a[5] += du[6]*sqrtd*bet2*s;
  a[5] *= 0.5;

  a[6]  = du[0]*(alpha_f_bar*(vsq-hp) + cff*(cf-v1) + qs_hat*vqstr - aspb);
  a[6] -= du[1]*(alpha_f_bar*v1 - cff);
  a[6] -= du[2]*(alpha_f_bar*v2 + qs_hat*q2_star);
  a[6] -= du[3]*(alpha_f_bar*v3 + qs_hat*q3_star);
  a[6] += du[4]*alpha_f_bar;
  a[6] += du[5]*(as*q2_star - alpha_f_bar*b2);
  a[6] += du[6]*(as*q3_star - alpha_f_bar*b3);

  double coeff[7];
#pragma unroll
  for (int n = 0; n < 7; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];

  double dens = wl_d + a[0]*alpha_f;
  if (dens < 0.0) llf_flag = 1;
  dens += a[2]*alpha_s;
  if (dens < 0.0) llf_flag = 1;
  dens += a[3];
  if (dens < 0.0) llf_flag = 1;
  dens += a[4]*alpha_s;
  if (dens < 0.0) llf_flag = 1;

  flx[0] += coeff[0]*alpha_f;
  flx[0] += coeff[2]*alpha_s;
  flx[0] += coeff[3];
  flx[0] += coeff[4]*alpha_s;
  flx[0] += coeff[6]*alpha_f;

  flx[1] += coeff[0]*(alpha_f*(v1 - cf));
  flx[1] += coeff[2]*(alpha_s*(v1 - cs));
  flx[1] += coeff[3]*v1;
  flx[1] += coeff[4]*(alpha_s*(v1 + cs));
  flx[1] += coeff[6]*(alpha_f*(v1 + cf));

  flx[2] += coeff[0]*(alpha_f*v2 + qs*bet2_star);
  flx[2] -= coeff[1]*bet3;
  flx[2] += coeff[2]*(alpha_s*v2 - qf*bet2_star);
  flx[2] += coeff[3]*v2;
  flx[2] += coeff[4]*(alpha_s*v2 + qf*bet2_star);
  flx[2] += coeff[5]*bet3;
  flx[2] += coeff[6]*(alpha_f*v2 - qs*bet2_star);

  flx[3] += coeff[0]*(alpha_f*v3 + qs*bet3_star);
  flx[3] += coeff[1]*bet2;
  flx[3] += coeff[2]*(alpha_s*v3 - qf*bet3_star);
  flx[3] += coeff[3]*v3;
  flx[3] += coeff[4]*(alpha_s*v3 + qf*bet3_star);
  flx[3] -= coeff[5]*bet2;
  flx[3] += coeff[6]*(alpha_f*v3 - qs*bet3_star);

  flx[4] += coeff[0]*(alpha_f*(hp - v1*cf) + qs*vbet + aspbb);
  flx[4] -= coeff[1]*(v2*bet3 - v3*bet2);
  flx[4] += coeff[2]*(alpha_s*(hp - v1*cs) - qf*vbet - afpbb);
  flx[4] += coeff[3]*(0.5*vsq + (gm1-1.0)*x/gm1);
  flx[4] += coeff[4]*(alpha_s*(hp + v1*cs) + qf*vbet - afpbb);
  flx[4] += coeff[5]*(v1*bet3 - v3*bet2);
  flx[4] += coeff[6]*(alpha_f*(hp + v1*cf) - qs*vbet + aspbb);

  flx[5] += coeff[0]*as_prime*bet2_star;
  flx[5] -= coeff[1]*bet3*s*isqrtd;
  flx[5] -= coeff[2]*af_prime*bet2_star;
  flx[5] -= coeff[4]*af_prime*bet2_star;
  flx[5] -= coeff[5]*bet3*s*isqrtd;
  flx[5] += coeff[6]*as_prime*bet2_star;

  flx[6] += coeff[0]*as_prime*bet3_star;
  flx[6] += coeff[1]*bet2*s*isqrtd;
  flx[6] -= coeff[2]*af_prime*bet3_star;
  flx[6] -= coeff[4]*af_prime*bet3_star;
  flx[6] += coeff[5]*bet2*s*isqrtd;
  flx[6] += coeff[6]*as_prime*bet3_star;
}

// Roe, adiabatic MHD (hydro/rsolvers/mhd/roe_mhd.cpp:43-238)
AB_HD void roe_mhd(const double *wli, const double *wri, double bxi, double gamma,
                   double *flxi) {
  double gm1 = gamma - 1.0;
  RoeAvgMHD r;
  roe_avg_mhd(wli, wri, bxi, gm1, r);
  double fl[7], fr[7], du[7];
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl*wli[IVX] + r.pbl - sqr(bxi);
  fr[IVX] = mxr*wri[IVX] + r.pbr - sqr(bxi);
  fl[IVY] = mxl*wli[IVY] - bxi*wli[IBY];
  fr[IVY] = mxr*wri[IVY] - bxi*wri[IBY];
  fl[IVZ] = mxl*wli[IVZ] - bxi*wli[IBZ];
  fr[IVZ] = mxr*wri[IVZ] - bxi*wri[IBZ];
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = (r.el + wli[IPR] + r.pbl - bxi*bxi)*wli[IVX];
  fr[IEN] = (r.er + wri[IPR] + r.pbr - bxi*bxi)*wri[IVX];
  fl[IEN] -= bxi*(wli[IBY]*wli[IVY] + wli[IBZ]*wli[IVZ]);
  fr[IEN] -= bxi*(wri[IBY]*wri[IVY] + wri[IBZ]*wri[IVZ]);
  fl[IBY] = wli[IBY]*wli[IVX] - bxi*wli[IVY];
  fr[IBY] = wri[IBY]*wri[IVX] - bxi*wri[IVY];
  fl[IBZ] = wli[IBZ]*wli[IVX] - bxi*wli[IVZ];
  fr[IBZ] = wri[IBZ]*wri[IVX] - bxi*wri[IVZ];
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  du[IEN] = r.er - r.el;
  du[IBY] = wri[IBY] - wli[IBY];
  du[IBZ] = wri[IBZ] - wli[IBZ];
#pragma unroll
  for (int n = 0; n < 7; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  double ev0, ev6;
  roe_flux_mhd(r, bxi, du, wli[IDN], gm1, flxi, ev0, ev6, llf_flag);
  if (ev0 >= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) flxi[n] = fl[n];
  }
  if (ev6 <= 0.0) {
#pragma unroll
    for (int n = 0; n < 7; ++n) flxi[n] = fr[n];
  }
  if (llf_flag != 0) {
    double cfl = fast_speed(gamma, wli[IDN], wli[IPR], wli[IBY], wli[IBZ], bxi);
    double cfr = fast_speed(gamma, wri[IDN], wri[IPR], wri[IBY], wri[IBZ], bxi);
    double a = 0.5*dmax((fabs(wli[IVX]) + cfl), (fabs(wri[IVX]) + cfr));
    // the reference's LLF fallback does not touch the IBY/IBZ fluxes (roe_mhd.cpp:222-232)
#pragma unroll
    for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
}

// ------------------------------------------------------------------------ isothermal EOS
// NON_BAROTROPIC_EOS == 0 (configure.py --eos=isothermal): NHYDRO = 4, no IPR / IEN.  The
// sweep-ordered work arrays keep the 7-slot layout of the adiabatic code; slot 4 is unused.

// src/eos/isothermal_mhd.cpp:122-129
AB_HD double fast_speed_iso(double cs, double d, double by, double bz, double bx) {
  double asq = (cs*cs)*d;
  double vaxsq = bx*bx;
  double ct2 = by*by + bz*bz;
  double qsq = vaxsq + ct2 + asq;
  double tmp = vaxsq + ct2 - asq;
  return sqrt(0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2))/d);
}
AB_HD double fast_speed_iso(double cs, const double *prim, double bx) {
  return fast_speed_iso(cs, prim[IDN], prim[IBY], prim[IBZ], bx);
}

// src/hydro/rsolvers/hydro/hlle.cpp:38-162 (isothermal branch)
AB_HD void hlle_hydro_iso(const double *wli, const double *wri, double iso_cs, double *flxi) {
  double wroe[5], fl[5], fr[5];
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double cl = iso_cs, cr = iso_cs, a = iso_cs;
  double al = dmin((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = dmax((wroe[IVX] + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl;
  fr[IVX] = wri[IDN]*wri[IVX]*vxr;
  fl[IVY] = wli[IDN]*wli[IVY]*vxl;
  fr[IVY] = wri[IDN]*wri[IVY]*vxr;
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl;
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr;
  fl[IVX] += (iso_cs*iso_cs)*wli[IDN];
  fr[IVX] += (iso_cs*iso_cs)*wri[IDN];
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
  for (int n = 0; n < 4; ++n)
    flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
  flxi[IEN] = 0.0;
}

// src/hydro/rsolvers/mhd/hlle_mhd.cpp:25-182 (isothermal branch)
AB_HD void hlle_mhd_iso(const double *wli, const double *wri, double bxi, double iso_cs,
                         double *flxi) {
  double wroe[7], fl[7], fr[7];
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  wroe[IBY] = (sqrtdr*wli[IBY] + sqrtdl*wri[IBY])*isdlpdr;
  wroe[IBZ] = (sqrtdr*wli[IBZ] + sqrtdl*wri[IBZ])*isdlpdr;
  double x = 0.5*(sqr(wli[IBY]-wri[IBY]) + sqr(wli[IBZ]-wri[IBZ]))/(sqr(sqrtdl+sqrtdr));
  double y = 0.5*(wli[IDN] + wri[IDN])/wroe[IDN];
  double pbl = 0.5*(bxi*bxi + sqr(wli[IBY]) + sqr(wli[IBZ]));
  double pbr = 0.5*(bxi*bxi + sqr(wri[IBY]) + sqr(wri[IBZ]));
  double cl = fast_speed_iso(iso_cs, wli, bxi);
  double cr = fast_speed_iso(iso_cs, wri, bxi);
  double btsq = sqr(wroe[IBY]) + sqr(wroe[IBZ]);
  double vaxsq = bxi*bxi/wroe[IDN];
  double bt_starsq = btsq*y;
  double twid_asq = iso_cs*iso_cs + x;
  double ct2 = bt_starsq/wroe[IDN];
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_asq*ct2);
  double cfsq = 0.5*(tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = dmin((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = dmax((wroe[IVX] + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + pbl - sqr(bxi);
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + pbr - sqr(bxi);
  fl[IVY] = wli[IDN]*wli[IVY]*vxl - bxi*wli[IBY];
  fr[IVY] = wri[IDN]*wri[IVY]*vxr - bxi*wri[IBY];
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl - bxi*wli[IBZ];
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr - bxi*wri[IBZ];
  fl[IVX] += (iso_cs*iso_cs)*wli[IDN];
  fr[IVX] += (iso_cs*iso_cs)*wri[IDN];
  fl[IBY] = wli[IBY]*vxl - bxi*wli[IVY];
  fr[IBY] = wri[IBY]*vxr - bxi*wri[IVY];
  fl[IBZ] = wli[IBZ]*vxl - bxi*wli[IVZ];
  fr[IBZ] = wri[IBZ]*vxr - bxi*wri[IVZ];
  fl[IEN] = fr[IEN] = 0.0;
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
  for (int n = 0; n < 7; ++n)
    flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
}

// src/hydro/rsolvers/mhd/hlld_iso.cpp:36-288 (Mignone 2007)
AB_HD void hlld_iso(const double *wli, const double *wri, double bxi, double cs, double dfloor,
                     double *flxi) {
  Cons1D ul, ur, ulst, urst, ucst, fl, fr;
  double spd[5];
  ul.d  = wli[IDN];
  ul.mx = wli[IVX]*ul.d;
  ul.my = wli[IVY]*ul.d;
  ul.mz = wli[IVZ]*ul.d;
  ul.by = wli[IBY];
  ul.bz = wli[IBZ];
  ur.d  = wri[IDN];
  ur.mx = wri[IVX]*ur.d;
  ur.my = wri[IVY]*ur.d;
  ur.mz = wri[IVZ]*ur.d;
  ur.by = wri[IBY];
  ur.bz = wri[IBZ];
  double cfl = fast_speed_iso(cs, wli, bxi);
  double cfr = fast_speed_iso(cs, wri, bxi);
  spd[0] = dmin(wli[IVX]-cfl, wri[IVX]-cfr);
  spd[4] = dmax(wli[IVX]+cfl, wri[IVX]+cfr);
  double bxsq = bxi*bxi;
  double ptl = sqr(cs)*wli[IDN] + 0.5*(bxsq + sqr(wli[IBY]) + sqr(wli[IBZ]));
  double ptr = sqr(cs)*wri[IDN] + 0.5*(bxsq + sqr(wri[IBY]) + sqr(wri[IBZ]));
  fl.d  = ul.mx;
  fl.mx = ul.mx*wli[IVX] + ptl - bxsq;
  fl.my = ul.my*wli[IVX] - bxi*ul.by;
  fl.mz = ul.mz*wli[IVX] - bxi*ul.bz;
  fl.by = ul.by*wli[IVX] - bxi*wli[IVY];
  fl.bz = ul.bz*wli[IVX] - bxi*wli[IVZ];
  fr.d  = ur.mx;
  fr.mx = ur.mx*wri[IVX] + ptr - bxsq;
  fr.my = ur.my*wri[IVX] - bxi*ur.by;
  fr.mz = ur.mz*wri[IVX] - bxi*ur.bz;
  fr.by = ur.by*wri[IVX] - bxi*wri[IVY];
  fr.bz = ur.bz*wri[IVX] - bxi*wri[IVZ];
  double idspd = 1.0/(spd[4]-spd[0]);
  double dhll = (spd[4]*ur.d - spd[0]*ul.d - fr.d + fl.d)*idspd;
  dhll = dmax(dhll, dfloor);
  double sqrtdhll = sqrt(dhll);
  double fdhll  = (spd[4]*fl.d  - spd[0]*fr.d  + spd[4]*spd[0]*(ur.d -ul.d ))*idspd;
  double fmxhll = (spd[4]*fl.mx - spd[0]*fr.mx + spd[4]*spd[0]*(ur.mx-ul.mx))*idspd;
  double ustar = fdhll/dhll;
  double mxhll = (spd[4]*ur.mx - spd[0]*ul.mx - fr.mx + fl.mx)*idspd;
  spd[1] = ustar - fabs(bxi)/sqrtdhll;
  spd[3] = ustar + fabs(bxi)/sqrtdhll;
  ulst.d  = dhll;
  ulst.mx = mxhll;
  double tmp = (spd[0]-spd[1])*(spd[0]-spd[3]);
  if (fabs(spd[0]-spd[1]) < (1.0e-8)*cs) {
    ulst.my = ul.my;
    ulst.mz = ul.mz;
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double mfact = bxi*(ustar-wli[IVX])/tmp;
    double bfact = (ul.d*sqr(spd[0]-wli[IVX]) - bxsq)/(dhll*tmp);
    ulst.my = dhll*wli[IVY] - ul.by*mfact;
    ulst.mz = dhll*wli[IVZ] - ul.bz*mfact;
    ulst.by = ul.by*bfact;
    ulst.bz = ul.bz*bfact;
  }
  urst.d  = dhll;
  urst.mx = mxhll;
  tmp = (spd[4]-spd[1])*(spd[4]-spd[3]);
  if (fabs(spd[4]-spd[3]) < (1.0e-8)*cs) {
    urst.my = ur.my;
    urst.mz = ur.mz;
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double mfact = bxi*(ustar-wri[IVX])/tmp;
    double bfact = (ur.d*sqr(spd[4]-wri[IVX]) - bxsq)/(dhll*tmp);
    urst.my = dhll*wri[IVY] - ur.by*mfact;
    urst.mz = dhll*wri[IVZ] - ur.bz*mfact;
    urst.by = ur.by*bfact;
    urst.bz = ur.bz*bfact;
  }
  double x = sqrtdhll*(bxi > 0.0 ? 1.0 : -1.0);
  ucst.d  = dhll;
  ucst.mx = mxhll;
  ucst.my = 0.5*(ulst.my + urst.my + (urst.by-ulst.by)*x);
  ucst.mz = 0.5*(ulst.mz + urst.mz + (urst.bz-ulst.bz)*x);
  ucst.by = 0.5*(ulst.by + urst.by + (urst.my-ulst.my)/x);
  ucst.bz = 0.5*(ulst.bz + urst.bz + (urst.mz-ulst.mz)/x);
  if (spd[0] >= 0.0) {
    flxi[IDN] = fl.d; flxi[IVX] = fl.mx; flxi[IVY] = fl.my; flxi[IVZ] = fl.mz;
    flxi[IBY] = fl.by; flxi[IBZ] = fl.bz;
  } else if (spd[4] <= 0.0) {
    flxi[IDN] = fr.d; flxi[IVX] = fr.mx; flxi[IVY] = fr.my; flxi[IVZ] = fr.mz;
    flxi[IBY] = fr.by; flxi[IBZ] = fr.bz;
  } else if (spd[1] >= 0.0) {
    flxi[IDN] = fl.d  + spd[0]*(ulst.d  - ul.d);
    flxi[IVX] = fl.mx + spd[0]*(ulst.mx - ul.mx);
    flxi[IVY] = fl.my + spd[0]*(ulst.my - ul.my);
    flxi[IVZ] = fl.mz + spd[0]*(ulst.mz - ul.mz);
    flxi[IBY] = fl.by + spd[0]*(ulst.by - ul.by);
    flxi[IBZ] = fl.bz + spd[0]*(ulst.bz - ul.bz);
  } else if (spd[3] <= 0.0) {
    flxi[IDN] = fr.d  + spd[4]*(urst.d  - ur.d);
    flxi[IVX] = fr.mx + spd[4]*(urst.mx - ur.mx);
    flxi[IVY] = fr.my + spd[4]*(urst.my - ur.my);
    flxi[IVZ] = fr.mz + spd[4]*(urst.mz - ur.mz);
    flxi[IBY] = fr.by + spd[4]*(urst.by - ur.by);
    flxi[IBZ] = fr.bz + spd[4]*(urst.bz - ur.bz);
  } else {
    flxi[IDN] = dhll*ustar;
    flxi[IVX] = fmxhll;
    flxi[IVY] = ucst.my*ustar - bxi*ucst.by;
    flxi[IVZ] = ucst.mz*ustar - bxi*ucst.bz;
    flxi[IBY] = ucst.by*ustar - bxi*ucst.my/ucst.d;
    flxi[IBZ] = ucst.bz*ustar - bxi*ucst.mz/ucst.d;
  }
  flxi[IEN] = 0.0;
}

// Roe, isothermal hydro (hydro/rsolvers/hydro/roe.cpp:42-352, the NON_BAROTROPIC_EOS == 0
// branches; RoeFlux :300-349): four waves, no energy equation
AB_HD void roe_hydro_iso(const double *wli, const double *wri, double iso_cs, double *flxi) {
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  double v1 = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  double v2 = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  double v3 = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  double fl[4], fr[4], du[4], ev[4];
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl*wli[IVX];
  fr[IVX] = mxr*wri[IVX];
  fl[IVY] = mxl*wli[IVY];
  fr[IVY] = mxr*wri[IVY];
  fl[IVZ] = mxl*wli[IVZ];
  fr[IVZ] = mxr*wri[IVZ];
  fl[IVX] += (iso_cs*iso_cs)*wli[IDN];
  fr[IVX] += (iso_cs*iso_cs)*wri[IDN];
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
#pragma unroll
  for (int n = 0; n < 4; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  {
    ev[0] = v1 - iso_cs;
    ev[1] = v1;
    ev[2] = v1;
    ev[3] = v1 + iso_cs;
    double a[4];
    a[0]  = du[0]*(0.5 + 0.5*v1/iso_cs);
    a[0] -= du[1]*0.5/iso_cs;
    a[1]  = du[0]*(-v2);
    a[1] += du[2];
    a[2]  = du[0]*(-v3);
    a[2] += du[3];
    a[3]  = du[0]*(0.5 - 0.5*v1/iso_cs);
    a[3] += du[1]*0.5/iso_cs;
    double coeff[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];
    double dens = wli[IDN] + a[0];
    if (dens < 0.0) llf_flag = 1;
    dens += a[3];
    if (dens < 0.0) llf_flag = 1;
    flxi[0] += coeff[0];
    flxi[0] += coeff[3];
    flxi[1] += coeff[0]*(v1 - iso_cs);
    flxi[1] += coeff[3]*(v1 + iso_cs);
    flxi[2] += coeff[0]*v2;
    flxi[2] += coeff[1];
    flxi[2] += coeff[3]*v2;
    flxi[3] += coeff[0]*v3;
    flxi[3] += coeff[2];
    flxi[3] += coeff[3]*v3;
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) flxi[n] = fl[n];
  }
  if (ev[3] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 4; ++n) flxi[n] = fr[n];
  }
  if (llf_flag != 0) {   // SoundSpeed of the isothermal EOS is the constant iso_cs
    double a = 0.5*dmax((fabs(wli[IVX]) + iso_cs), (fabs(wri[IVX]) + iso_cs));
#pragma unroll
    for (int n = 0; n < 4; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
  flxi[IEN] = 0.0;
}

// Roe, isothermal MHD (hydro/rsolvers/mhd/roe_mhd.cpp:43-238 with NON_BAROTROPIC_EOS == 0;
// RoeFlux :245-345,498-612): six waves.  wli / wri / flxi keep the 7-slot layout (slot 4 unused).
AB_HD void roe_mhd_iso(const double *wli, const double *wri, double bxi, double iso_cs,
                       double *flxi) {
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  double d  = sqrtdl*sqrtdr;
  double v1 = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  double v2 = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  double v3 = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double b2 = (sqrtdr*wli[IBY] + sqrtdl*wri[IBY])*isdlpdr;
  double b3 = (sqrtdr*wli[IBZ] + sqrtdl*wri[IBZ])*isdlpdr;
  double x = 0.5*(sqr(wli[IBY] - wri[IBY]) + sqr(wli[IBZ] - wri[IBZ]))/(sqr(sqrtdl + sqrtdr));
  double y = 0.5*(wli[IDN] + wri[IDN])/d;
  double pbl = 0.5*(bxi*bxi + sqr(wli[IBY]) + sqr(wli[IBZ]));
  double pbr = 0.5*(bxi*bxi + sqr(wri[IBY]) + sqr(wri[IBZ]));
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  double fl[6], fr[6], du[6], ev[6], flx[6];
  fl[0] = mxl;
  fr[0] = mxr;
  fl[1] = mxl*wli[IVX] + pbl - sqr(bxi);
  fr[1] = mxr*wri[IVX] + pbr - sqr(bxi);
  fl[2] = mxl*wli[IVY] - bxi*wli[IBY];
  fr[2] = mxr*wri[IVY] - bxi*wri[IBY];
  fl[3] = mxl*wli[IVZ] - bxi*wli[IBZ];
  fr[3] = mxr*wri[IVZ] - bxi*wri[IBZ];
  fl[1] += (iso_cs*iso_cs)*wli[IDN];
  fr[1] += (iso_cs*iso_cs)*wri[IDN];
  fl[4] = wli[IBY]*wli[IVX] - bxi*wli[IVY];
  fr[4] = wri[IBY]*wri[IVX] - bxi*wri[IVY];
  fl[5] = wli[IBZ]*wli[IVX] - bxi*wli[IVZ];
  fr[5] = wri[IBZ]*wri[IVX] - bxi*wri[IVZ];
  du[0] = wri[IDN]          - wli[IDN];
  du[1] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[2] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[3] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  du[4] = wri[IBY] - wli[IBY];
  du[5] = wri[IBZ] - wli[IBZ];
#pragma unroll
  for (int n = 0; n < 6; ++n) flx[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  {
    double b1 = bxi;
    double di = 1.0/d;
    double btsq = b2*b2 + b3*b3;
    double vaxsq = b1*b1*di;
    double bt_starsq = btsq*y;
    double twid_csq = (iso_cs*iso_cs) + x;
    double ct2 = bt_starsq*di;
    double tsum = vaxsq + ct2 + twid_csq;
    double tdif = vaxsq + ct2 - twid_csq;
    double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_csq*ct2);
    double cfsq = 0.5*(tsum + cf2_cs2);
    double cf = sqrt(cfsq);
    double cssq = twid_csq*vaxsq/cfsq;
    double cs = sqrt(cssq);
    double bt = sqrt(btsq);
    double bt_star = sqrt(bt_starsq);
    double bet2 = 0.0, bet3 = 0.0;
    if (bt != 0.0) {
      bet2 = b2/bt;
      bet3 = b3/bt;
    }
    double bet2_star = bet2/sqrt(y);
    double bet3_star = bet3/sqrt(y);
    double bet_starsq = bet2_star*bet2_star + bet3_star*bet3_star;
    double q2_star = 0.0, q3_star = 0.0;
    if (bet_starsq != 0.0) {
      q2_star = bet2_star/bet_starsq;
      q3_star = bet3_star/bet_starsq;
    }
    double alpha_f, alpha_s;
    if ((cfsq - cssq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else if ((twid_csq - cssq) <= 0.0) {
      alpha_f = 0.0;
      alpha_s = 1.0;
    } else if ((cfsq - twid_csq) <= 0.0) {
      alpha_f = 1.0;
      alpha_s = 0.0;
    } else {
      alpha_f = sqrt((twid_csq - cssq)/(cfsq - cssq));
      alpha_s = sqrt((cfsq - twid_csq)/(cfsq - cssq));
    }
    double sqrtd = sqrt(d);
    double isqrtd = 1.0/sqrtd;
    double s = (b1 < 0.0) ? -1.0 : 1.0;
    double twid_c = sqrt(twid_csq);
    double qf = cf*alpha_f*s;
    double qs = cs*alpha_s*s;
    double af_prime = twid_c*alpha_f*isqrtd;
    double as_prime = twid_c*alpha_s*isqrtd;
    double vqstr = (v2*q2_star + v3*q3_star);
    double vax = sqrt(vaxsq);
    double norm = 0.5/twid_csq;
    double cff = norm*alpha_f*cf;
    double css = norm*alpha_s*cs;
    double qf_hat = qf*norm;
    double qs_hat = qs*norm;
    double af = norm*af_prime*d;
    double as = norm*as_prime*d;
    double afpb = norm*af_prime*bt_star;
    double aspb = norm*as_prime*bt_star;
    ev[0] = v1 - cf;
    ev[1] = v1 - vax;
    ev[2] = v1 - cs;
    ev[3] = v1 + cs;
    ev[4] = v1 + vax;
    ev[5] = v1 + cf;
    double a[6];
    a[0]  = du[0]*(cff*(cf+v1) - qs_hat*vqstr - aspb);
    a[0] -= du[1]*cff;
    a[0] += du[2]*qs_hat*q2_star;
    a[0] += du[3]*qs_hat*q3_star;
    a[0] += du[4]*as*q2_star;
    a[0] += du[5]*as*q3_star;
    a[1]  = du[0]*(v2*bet3 - v3*bet2);
    a[1] -= du[2]*bet3;
    a[1] += du[3]*bet2;
    a[1] -= du[4]*sqrtd*bet3*s;
    a[1] += du[5]*sqrtd*bet2*s;
    a[1] *= 0.5;
    a[2]  = du[0]*(css*(cs+v1) + qf_hat*vqstr + afpb);
    a[2] -= du[1]*css;
    a[2] -= du[2]*qf_hat*q2_star;
    a[2] -= du[3]*qf_hat*q3_star;
    a[2] -= du[4]*af*q2_star;
    a[2] -= du[5]*af*q3_star;
    a[3]  = du[0]*(css*(cs-v1) - qf_hat*vqstr + afpb);
    a[3] += du[1]*css;
    a[3] += du[2]*qf_hat*q2_star;
    a[3] += du[3]*qf_hat*q3_star;
    a[3] -= du[4]*af*q2_star;
    a[3] -= du[5]*af*q3_star;
    a[4]  = du[0]*(v3*bet2 - v2*bet3);
    a[4] += du[2]*bet3;
    a[4] -= du[3]*bet2;
    a[4] -= du[4]*sqrtd*bet3*s;
    a[4] += du[5]*sqrtd*bet2*s;
    a[4] *= 0.5;
    a[5]  = du[0]*(cff*(cf-v1) + qs_hat*vqstr - aspb);
    a[5] += du[1]*cff;
    a[5] -= du[2]*qs_hat*q2_star;
    a[5] -= du[3]*qs_hat*q3_star;
    a[5] += du[4]*as*q2_star;
    a[5] += du[5]*as*q3_star;
    double coeff[6];
#pragma unroll
    for (int n = 0; n < 6; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];
    double dens = wli[IDN] + a[0]*alpha_f;
    if (dens < 0.0) llf_flag = 1;
    dens += a[2]*alpha_s;
    if (dens < 0.0) llf_flag = 1;
    dens += a[3]*alpha_s;
    if (dens < 0.0) llf_flag = 1;
    flx[0] += coeff[0]*alpha_f;
    flx[0] += coeff[2]*alpha_s;
    flx[0] += coeff[3]*alpha_s;
    flx[0] += coeff[5]*alpha_f;
    flx[1] += coeff[0]*alpha_f*(v1 - cf);
    flx[1] += coeff[2]*alpha_s*(v1 - cs);
    flx[1] += coeff[3]*alpha_s*(v1 + cs);
    flx[1] += coeff[5]*alpha_f*(v1 + cf);
    flx[2] += coeff[0]*(alpha_f*v2 + qs*bet2_star);
    flx[2] -= coeff[1]*bet3;
    flx[2] += coeff[2]*(alpha_s*v2 - qf*bet2_star);
    flx[2] += coeff[3]*(alpha_s*v2 + qf*bet2_star);
    flx[2] += coeff[4]*bet3;
    flx[2] += coeff[5]*(alpha_f*v2 - qs*bet2_star);
    flx[3] += coeff[0]*(alpha_f*v3 + qs*bet3_star);
    flx[3] += coeff[1]*bet2;
    flx[3] += coeff[2]*(alpha_s*v3 - qf*bet3_star);
    flx[3] += coeff[3]*(alpha_s*v3 + qf*bet3_star);
    flx[3] -= coeff[4]*bet2;
    flx[3] += coeff[5]*(alpha_f*v3 - qs*bet3_star);
    flx[4] += coeff[0]*as_prime*bet2_star;
    flx[4] -= coeff[1]*bet3*s/sqrtd;
    flx[4] -= coeff[2]*af_prime*bet2_star;
    flx[4] -= coeff[3]*af_prime*bet2_star;
    flx[4] -= coeff[4]*bet3*s/sqrtd;
    flx[4] += coeff[5]*as_prime*bet2_star;
    flx[5] += coeff[0]*as_prime*bet3_star;
    flx[5] += coeff[1]*bet2*s/sqrtd;
    flx[5] -= coeff[2]*af_prime*bet3_star;
    flx[5] -= coeff[3]*af_prime*bet3_star;
    flx[5] += coeff[4]*bet2*s/sqrtd;
    flx[5] += coeff[5]*as_prime*bet3_star;
  }
  if (ev[0] >= 0.0) {
#pragma unroll
    for (int n = 0; n < 6; ++n) flx[n] = fl[n];
  }
  if (ev[5] <= 0.0) {
#pragma unroll
    for (int n = 0; n < 6; ++n) flx[n] = fr[n];
  }
  if (llf_flag != 0) {
    double cfl = fast_speed_iso(iso_cs, wli, bxi);
    double cfr = fast_speed_iso(iso_cs, wri, bxi);
    double a = 0.5*dmax((fabs(wli[IVX]) + cfl), (fabs(wri[IVX]) + cfr));
    // the fallback leaves the field fluxes untouched (roe_mhd.cpp:222-232)
#pragma unroll
    for (int n = 0; n < 4; ++n) flx[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
  flxi[IDN] = flx[0];
  flxi[IVX] = flx[1];
  flxi[IVY] = flx[2];
  flxi[IVZ] = flx[3];
  flxi[IEN] = 0.0;
  flxi[IBY] = flx[4];
  flxi[IBZ] = flx[5];
}

// internal solver ids of the isothermal variants (the ABI keeps AB_SOLVER_* + AB_EOS_*)
enum : int { SOLVER_HLLE_ISO = 8, SOLVER_HLLD_ISO = 9, SOLVER_LLF_ISO = 10, SOLVER_ROE_ISO = 11 };
template <int SOLVER> constexpr bool solver_is_iso = (SOLVER >= SOLVER_HLLE_ISO);

// LLF (hydro/rsolvers/hydro/llf.cpp:34-125, mhd/llf_mhd.cpp:34-170), both EOS; `ga` is gamma
// (adiabatic) or the isothermal sound speed
template <bool MHD, bool ISO>
AB_HD void llf_t(const double *wli, const double *wri, double bxi, double ga, double *flxi) {
  double fl[7], fr[7], du[7];
  const double gm1 = ga - 1.0;
  double cl, cr;
  if (MHD) {
    cl = ISO ? fast_speed_iso(ga, wli, bxi)
             : fast_speed(ga, wli[IDN], wli[IPR], wli[IBY], wli[IBZ], bxi);
    cr = ISO ? fast_speed_iso(ga, wri, bxi)
             : fast_speed(ga, wri[IDN], wri[IPR], wri[IBY], wri[IBZ], bxi);
  } else {
    cl = ISO ? ga : sound_speed(ga, wli[IDN], wli[IPR]);
    cr = ISO ? ga : sound_speed(ga, wri[IDN], wri[IPR]);
  }
  const double a = 0.5*dmax((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
  const double mxl = wli[IDN]*wli[IVX];
  const double mxr = wri[IDN]*wri[IVX];
  double pbl = 0.0, pbr = 0.0;
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  if (MHD) {
    pbl = 0.5*(bxi*bxi + sqr(wli[IBY]) + sqr(wli[IBZ]));
    pbr = 0.5*(bxi*bxi + sqr(wri[IBY]) + sqr(wri[IBZ]));
    fl[IVX] = mxl*wli[IVX] + pbl - sqr(bxi);
    fr[IVX] = mxr*wri[IVX] + pbr - sqr(bxi);
    fl[IVY] = mxl*wli[IVY] - bxi*wli[IBY];
    fr[IVY] = mxr*wri[IVY] - bxi*wri[IBY];
    fl[IVZ] = mxl*wli[IVZ] - bxi*wli[IBZ];
    fr[IVZ] = mxr*wri[IVZ] - bxi*wri[IBZ];
  } else {
    fl[IVX] = mxl*wli[IVX];
    fr[IVX] = mxr*wri[IVX];
    fl[IVY] = mxl*wli[IVY];
    fr[IVY] = mxr*wri[IVY];
    fl[IVZ] = mxl*wli[IVZ];
    fr[IVZ] = mxr*wri[IVZ];
  }
  double el = 0.0, er = 0.0;
  fl[IEN] = fr[IEN] = 0.0;
  if (!ISO) {
    if (MHD) {
      el = wli[IPR]/gm1 + 0.5*wli[IDN]*(sqr(wli[IVX])+sqr(wli[IVY])+sqr(wli[IVZ])) + pbl;
      er = wri[IPR]/gm1 + 0.5*wri[IDN]*(sqr(wri[IVX])+sqr(wri[IVY])+sqr(wri[IVZ])) + pbr;
    } else {
      el = wli[IPR]/gm1 + 0.5*wli[IDN]*(sqr(wli[IVX]) + sqr(wli[IVY]) + sqr(wli[IVZ]));
      er = wri[IPR]/gm1 + 0.5*wri[IDN]*(sqr(wri[IVX]) + sqr(wri[IVY]) + sqr(wri[IVZ]));
    }
    fl[IVX] += wli[IPR];
    fr[IVX] += wri[IPR];
    if (MHD) {
      fl[IEN] = (el + wli[IPR] + pbl - bxi*bxi)*wli[IVX];
      fr[IEN] = (er + wri[IPR] + pbr - bxi*bxi)*wri[IVX];
      fl[IEN] -= bxi*(wli[IBY]*wli[IVY] + wli[IBZ]*wli[IVZ]);
      fr[IEN] -= bxi*(wri[IBY]*wri[IVY] + wri[IBZ]*wri[IVZ]);
    } else {
      fl[IEN] = (el + wli[IPR])*wli[IVX];
      fr[IEN] = (er + wri[IPR])*wri[IVX];
    }
  } else {
    fl[IVX] += (ga*ga)*wli[IDN];
    fr[IVX] += (ga*ga)*wri[IDN];
  }
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  du[IEN] = ISO ? 0.0 : (er - el);
  if (MHD) {
    fl[IBY] = wli[IBY]*wli[IVX] - bxi*wli[IVY];
    fr[IBY] = wri[IBY]*wri[IVX] - bxi*wri[IVY];
    fl[IBZ] = wli[IBZ]*wli[IVX] - bxi*wli[IVZ];
    fr[IBZ] = wri[IBZ]*wri[IVX] - bxi*wri[IVZ];
    du[IBY] = wri[IBY] - wli[IBY];
    du[IBZ] = wri[IBZ] - wli[IBZ];
  }
#pragma unroll
  for (int n = 0; n < (MHD ? 7 : 5); ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  if (ISO) flxi[IEN] = 0.0;
}

// compile-time dispatch
// `gamma` carries the isothermal sound speed and `dfloor` the density floor for the *_ISO ids
template <int SOLVER, bool MHD>
AB_HD void riemann(const double *wli, const double *wri, double bxi, double gamma, double dvn,
                   double dvt, double *flxi, double dfloor = 0.0) {
  if (SOLVER == SOLVER_LLF) {
    llf_t<MHD,false>(wli, wri, bxi, gamma, flxi);
  } else if (SOLVER == SOLVER_LLF_ISO) {
    llf_t<MHD,true>(wli, wri, bxi, gamma, flxi);
  } else if (SOLVER == SOLVER_HLLE_ISO) {
    if (MHD) hlle_mhd_iso(wli, wri, bxi, gamma, flxi);
    else hlle_hydro_iso(wli, wri, gamma, flxi);
  } else if (SOLVER == SOLVER_HLLD_ISO) {
    hlld_iso(wli, wri, bxi, gamma, dfloor, flxi);
  } else if (SOLVER == SOLVER_ROE_ISO) {
    if (MHD) roe_mhd_iso(wli, wri, bxi, gamma, flxi);
    else roe_hydro_iso(wli, wri, gamma, flxi);
  } else if (!MHD) {
    if (SOLVER == SOLVER_HLLC) hllc_t<false>(wli, wri, gamma, 0.0, 0.0, flxi);
    else if (SOLVER == SOLVER_LHLLC) hllc_t<true>(wli, wri, gamma, dvn, dvt, flxi);
    else if (SOLVER == SOLVER_HLLE) hlle_hydro(wli, wri, gamma, flxi);
    else roe_hydro(wli, wri, gamma, flxi);
  } else {
    if (SOLVER == SOLVER_HLLD) hlld_t<false>(wli, wri, bxi, gamma, 0.0, 0.0, flxi);
    else if (SOLVER == SOLVER_LHLLD) hlld_t<true>(wli, wri, bxi, gamma, dvn, dvt, flxi);
    else if (SOLVER == SOLVER_HLLE) hlle_mhd(wli, wri, bxi, gamma, flxi);
    else roe_mhd(wli, wri, bxi, gamma, flxi);
  }
}

}  // namespace ab
#endif  // AB_PHYSICS_CUH_
