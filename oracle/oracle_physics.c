/* oracle_physics.c -- point-wise physics of the reference path, restated in plain C.
 *
 * TEST INFRASTRUCTURE ONLY (see athena_oracle.h).  Every function cites the reference
 * file:line it restates.  Expression order and parenthesisation follow the reference
 * exactly (the parity bar is bit-for-bit); compile with -ffp-contract=off, no -ffast-math.
 */
#include <math.h>
#include <float.h>
#include <string.h>
#include "oracle_internal.h"

/* std::min / std::max semantics: (b<a)?b:a and (a<b)?b:a */
static inline double mn(double a, double b) { return (b < a) ? b : a; }
static inline double mx(double a, double b) { return (a < b) ? b : a; }
#define SQR(x) ((x)*(x))
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
#define TINY_NUMBER 1.0e-20

/* src/eos/adiabatic_hydro.cpp:125 */
double ao_sound_speed(double gamma, const double *prim) {
  return sqrt(gamma*prim[IPR]/prim[IDN]);
}

/* src/eos/adiabatic_mhd.cpp:149-156 */
double ao_fast_speed(double gamma, const double *prim, double bx) {
  double asq = gamma*prim[IPR];
  double vaxsq = bx*bx;
  double ct2 = (prim[IBY]*prim[IBY] + prim[IBZ]*prim[IBZ]);
  double qsq = vaxsq + ct2 + asq;
  double tmp = vaxsq + ct2 - asq;
  return sqrt(0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2))/prim[IDN]);
}

/* src/hydro/hydro.cpp:154-158 */
double ao_weight_for_ct(double dflx, double rhol, double rhor, double dx, double dt) {
  double v_over_c = (1024.0)*dt*dflx/(dx*(rhol + rhor));
  double tmp_min = mn(0.5, v_over_c);
  return 0.5 + mx(-0.5, tmp_min);
}

/* ---------------------------------------------------------------- hydro solvers */

/* src/hydro/rsolvers/hydro/hllc.cpp:32-179 */
static void hllc(const double *wli, const double *wri, double gamma, double *flxi) {
  double fl[5], fr[5];
  double gm1 = gamma - 1.0;
  double igm1 = 1.0/gm1;
  double cl = ao_sound_speed(gamma, wli);
  double cr = ao_sound_speed(gamma, wri);
  double el = wli[IPR]*igm1 + 0.5*wli[IDN]*(SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]));
  double er = wri[IPR]*igm1 + 0.5*wri[IDN]*(SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]));
  double rhoa = .5*(wli[IDN] + wri[IDN]);
  double ca = .5*(cl + cr);
  double pmid = .5*(wli[IPR] + wri[IPR] + (wli[IVX]-wri[IVX])*rhoa*ca);
  double umid = .5*(wli[IVX] + wri[IVX] + (wli[IPR]-wri[IPR])/(rhoa*ca));
  double rhol = wli[IDN] + (wli[IVX] - umid)*rhoa/ca;
  double rhor = wri[IDN] + (umid - wri[IVX])*rhoa/ca;
  (void)rhol; (void)rhor;
  double ql = (pmid <= wli[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wli[IPR]-1.0));
  double qr = (pmid <= wri[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wri[IPR]-1.0));
  double al = wli[IVX] - cl*ql;
  double ar = wri[IVX] + cr*qr;
  double bp = ar > 0.0 ? ar : (TINY_NUMBER);
  double bm = al < 0.0 ? al : -(TINY_NUMBER);
  double vxl = wli[IVX] - al;
  double vxr = wri[IVX] - ar;
  double tl = wli[IPR] + vxl*wli[IDN]*wli[IVX];
  double tr = wri[IPR] + vxr*wri[IDN]*wri[IVX];
  double ml = wli[IDN]*vxl;
  double mr = -(wri[IDN]*vxr);
  double am = (tl - tr)/(ml + mr);
  double cp = (ml*tr + mr*tl)/(ml + mr);
  cp = cp > 0.0 ? cp : 0.0;
  vxl = wli[IVX] - bm;
  vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + wli[IPR];
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + wri[IPR];
  fl[IVY] = wli[IDN]*wli[IVY]*vxl;
  fr[IVY] = wri[IDN]*wri[IVY]*vxr;
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl;
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr;
  fl[IEN] = el*vxl + wli[IPR]*wli[IVX];
  fr[IEN] = er*vxr + wri[IPR]*wri[IVX];
  double sl, sr, sm;
  if (am >= 0.0) {
    sl = am/(am - bm);
    sr = 0.0;
    sm = -bm/(am - bm);
  } else {
    sl = 0.0;
    sr = -am/(bp - am);
    sm = bp/(bp - am);
  }
  flxi[IDN] = sl*fl[IDN] + sr*fr[IDN];
  flxi[IVX] = sl*fl[IVX] + sr*fr[IVX] + sm*cp;
  flxi[IVY] = sl*fl[IVY] + sr*fr[IVY];
  flxi[IVZ] = sl*fl[IVZ] + sr*fr[IVZ];
  flxi[IEN] = sl*fl[IEN] + sr*fr[IEN] + sm*cp*am;
}

/* src/hydro/rsolvers/hydro/lhllc.cpp:26-175 (Minoshima et al. 2021 low-dissipation HLLC);
 * dvn, dvt from Hydro::CalculateVelocityDifferences */
static void lhllc(const double *wli, const double *wri, double gamma, double dvn, double dvt,
                  double *flxi) {
  double fl[5], fr[5];
  double gm1 = gamma - 1.0;
  double igm1 = 1.0/gm1;
  double cl = ao_sound_speed(gamma, wli);
  double cr = ao_sound_speed(gamma, wri);
  double vsql = SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]);
  double vsqr = SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]);
  double el = wli[IPR]*igm1 + 0.5*wli[IDN]*vsql;
  double er = wri[IPR]*igm1 + 0.5*wri[IDN]*vsqr;
  double rhoa = .5*(wli[IDN] + wri[IDN]);
  double ca = .5*(cl + cr);
  double pmid = .5*(wli[IPR] + wri[IPR] + (wli[IVX]-wri[IVX])*rhoa*ca);
  double umid = .5*(wli[IVX] + wri[IVX] + (wli[IPR]-wri[IPR])/(rhoa*ca));
  double rhol = wli[IDN] + (wli[IVX] - umid)*rhoa/ca;
  double rhor = wri[IDN] + (umid - wri[IVX])*rhoa/ca;
  (void)rhol; (void)rhor;
  double ql = (pmid <= wli[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wli[IPR]-1.0));
  double qr = (pmid <= wri[IPR]) ? 1.0 :
      sqrt(1.0 + (gamma + 1)/(2*gamma)*(pmid/wri[IPR]-1.0));
  double al = wli[IVX] - cl*ql;
  double ar = wri[IVX] + cr*qr;
  double bp = ar > 0.0 ? ar : (TINY_NUMBER);
  double bm = al < 0.0 ? al : -(TINY_NUMBER);
  double vxl = al - wli[IVX];
  double vxr = ar - wri[IVX];
  double ml = wli[IDN]*vxl;
  double mr = wri[IDN]*vxr;
  double cmax = mx(cl, cr);
  double th1 = mn(1.0, (cmax-mn(dvn,0.0))/(cmax-mn(dvt,0.0)));
  double th = th1*th1*th1*th1;
  double am = (mr*wri[IVX] - ml*wli[IVX] - th*(wri[IPR]-wli[IPR]))/(mr - ml);
  double chi = mn(1.0, sqrt(mx(vsql, vsqr))/cmax);
  double phi = chi*(2.0 - chi);
  double cp = (mr*wli[IPR] - ml*wri[IPR] + phi*mr*ml*(wri[IVX]-wli[IVX]))/(mr - ml);
  cp = cp > 0.0 ? cp : 0.0;
  vxl = wli[IVX] - bm;
  vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + wli[IPR];
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + wri[IPR];
  fl[IVY] = wli[IDN]*wli[IVY]*vxl;
  fr[IVY] = wri[IDN]*wri[IVY]*vxr;
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl;
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr;
  fl[IEN] = el*vxl + wli[IPR]*wli[IVX];
  fr[IEN] = er*vxr + wri[IPR]*wri[IVX];
  double sl, sr, sm;
  if (am >= 0.0) {
    sl = am/(am - bm);
    sr = 0.0;
    sm = -bm/(am - bm);
  } else {
    sl = 0.0;
    sr = -am/(bp - am);
    sm = bp/(bp - am);
  }
  flxi[IDN] = sl*fl[IDN] + sr*fr[IDN];
  flxi[IVX] = sl*fl[IVX] + sr*fr[IVX] + sm*cp;
  flxi[IVY] = sl*fl[IVY] + sr*fr[IVY];
  flxi[IVZ] = sl*fl[IVZ] + sr*fr[IVZ];
  flxi[IEN] = sl*fl[IEN] + sr*fr[IEN] + sm*cp*am;
}

/* src/hydro/rsolvers/hydro/hlle.cpp:38-162 (adiabatic branch) */
static void hlle_hydro(const double *wli, const double *wri, double gamma, double *flxi) {
  double wroe[5], fl[5], fr[5];
  double gm1 = gamma - 1.0;
  double igm1 = 1.0/gm1;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double el = wli[IPR]*igm1 + 0.5*wli[IDN]*(SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]));
  double er = wri[IPR]*igm1 + 0.5*wri[IDN]*(SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]));
  double hroe = ((el + wli[IPR])/sqrtdl + (er + wri[IPR])/sqrtdr)*isdlpdr;
  double cl = ao_sound_speed(gamma, wli);
  double cr = ao_sound_speed(gamma, wri);
  double q = hroe - 0.5*(SQR(wroe[IVX]) + SQR(wroe[IVY]) + SQR(wroe[IVZ]));
  double a = (q < 0.0) ? 0.0 : sqrt(gm1*q);
  double al = mn((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = mx((wroe[IVX] + a), (wri[IVX] + cr));
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
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = el*vxl + wli[IPR]*wli[IVX];
  fr[IEN] = er*vxr + wri[IPR]*wri[IVX];
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
  for (int n = 0; n < 5; ++n)
    flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
}

/* src/hydro/rsolvers/hydro/roe.cpp:206-352 (adiabatic branch of RoeFlux) */
static void roe_flux_hydro(const double *wroe, const double *du, const double *wli,
                           double gm1, double *flx, double *ev, int *llf_flag) {
  double v1 = wroe[IVX];
  double v2 = wroe[IVY];
  double v3 = wroe[IVZ];
  double h = wroe[IPR];
  double vsq = v1*v1 + v2*v2 + v3*v3;
  double q = h - 0.5*vsq;
  double cs_sq = (q < 0.0) ? (TINY_NUMBER) : gm1*q;
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
  for (int n = 0; n < 5; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];
  double dens = wli[IDN] + a[0];
  if (dens < 0.0) *llf_flag = 1;
  dens += a[3];
  if (dens < 0.0) *llf_flag = 1;
  flx[0] += coeff[0];
  flx[0] += coeff[3];
  flx[0] += coeff[4];
  flx[1] += coeff[0]*(v1 - cs);
  flx[1] += coeff[3]*v1;
  flx[1] += coeff[4]*(v1 + cs);
  flx[2] += coeff[0]*v2;
  flx[2] += coeff[1];
  flx[2] += coeff[3]*v2;
  flx[2] += coeff[4]*v2;
  flx[3] += coeff[0]*v3;
  flx[3] += coeff[2];
  flx[3] += coeff[3]*v3;
  flx[3] += coeff[4]*v3;
  flx[4] += coeff[0]*(h - v1*cs);
  flx[4] += coeff[1]*v2;
  flx[4] += coeff[2]*v3;
  flx[4] += coeff[3]*0.5*vsq;
  flx[4] += coeff[4]*(h + v1*cs);
}

/* src/hydro/rsolvers/hydro/roe.cpp:42-200 */
static void roe_hydro(const double *wli, const double *wri, double gamma, double *flxi) {
  double wroe[5], fl[5], fr[5], ev[5], du[5];
  double gm1 = gamma - 1.0;
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double el = wli[IPR]/gm1 + 0.5*wli[IDN]*(SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]));
  double er = wri[IPR]/gm1 + 0.5*wri[IDN]*(SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]));
  wroe[IPR] = ((el + wli[IPR])/sqrtdl + (er + wri[IPR])/sqrtdr)*isdlpdr;
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
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
  for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  roe_flux_hydro(wroe, du, wli, gm1, flxi, ev, &llf_flag);
  if (ev[0] >= 0.0) for (int n = 0; n < 5; ++n) flxi[n] = fl[n];
  if (ev[4] <= 0.0) for (int n = 0; n < 5; ++n) flxi[n] = fr[n];
  if (llf_flag != 0) {
    double cl = ao_sound_speed(gamma, wli);
    double cr = ao_sound_speed(gamma, wri);
    double a = 0.5*mx((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
    for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
}

/* ---------------------------------------------------------------- MHD solvers */

typedef struct { double d, mx, my, mz, e, by, bz; } Cons1D;
#define SMALL_NUMBER 1.0e-8

/* src/hydro/rsolvers/mhd/hlld.cpp:38-382 */
static void hlld(const double *wli, const double *wri, double bxi, double gamma,
                 double *flxi) {
  double spd[5];
  Cons1D ul, ur, ulst, uldst, urdst, urst, fl, fr;
  double igm1 = 1.0/(gamma - 1.0);
  double bxsq = bxi*bxi;
  double pbl = 0.5*(bxsq + (SQR(wli[IBY]) + SQR(wli[IBZ])));
  double pbr = 0.5*(bxsq + (SQR(wri[IBY]) + SQR(wri[IBZ])));
  double kel = 0.5*wli[IDN]*(SQR(wli[IVX]) + (SQR(wli[IVY]) + SQR(wli[IVZ])));
  double ker = 0.5*wri[IDN]*(SQR(wri[IVX]) + (SQR(wri[IVY]) + SQR(wri[IVZ])));
  ul.d  = wli[IDN];
  ul.mx = wli[IVX]*ul.d;
  ul.my = wli[IVY]*ul.d;
  ul.mz = wli[IVZ]*ul.d;
  ul.e  = wli[IPR]*igm1 + kel + pbl;
  ul.by = wli[IBY];
  ul.bz = wli[IBZ];
  ur.d  = wri[IDN];
  ur.mx = wri[IVX]*ur.d;
  ur.my = wri[IVY]*ur.d;
  ur.mz = wri[IVZ]*ur.d;
  ur.e  = wri[IPR]*igm1 + ker + pbr;
  ur.by = wri[IBY];
  ur.bz = wri[IBZ];
  double cfl = ao_fast_speed(gamma, wli, bxi);
  double cfr = ao_fast_speed(gamma, wri, bxi);
  spd[0] = mn(wli[IVX]-cfl, wri[IVX]-cfr);
  spd[4] = mx(wli[IVX]+cfl, wri[IVX]+cfr);
  double ptl = wli[IPR] + pbl;
  double ptr = wri[IPR] + pbr;
  fl.d  = ul.mx;
  fl.mx = ul.mx*wli[IVX] + ptl - bxsq;
  fl.my = ul.my*wli[IVX] - bxi*ul.by;
  fl.mz = ul.mz*wli[IVX] - bxi*ul.bz;
  fl.e  = wli[IVX]*(ul.e + ptl - bxsq) - bxi*(wli[IVY]*ul.by + wli[IVZ]*ul.bz);
  fl.by = ul.by*wli[IVX] - bxi*wli[IVY];
  fl.bz = ul.bz*wli[IVX] - bxi*wli[IVZ];
  fr.d  = ur.mx;
  fr.mx = ur.mx*wri[IVX] + ptr - bxsq;
  fr.my = ur.my*wri[IVX] - bxi*ur.by;
  fr.mz = ur.mz*wri[IVX] - bxi*ur.bz;
  fr.e  = wri[IVX]*(ur.e + ptr - bxsq) - bxi*(wri[IVY]*ur.by + wri[IVZ]*ur.bz);
  fr.by = ur.by*wri[IVX] - bxi*wri[IVY];
  fr.bz = ur.bz*wri[IVX] - bxi*wri[IVZ];
  double sdl = spd[0] - wli[IVX];
  double sdr = spd[4] - wri[IVX];
  spd[2] = (sdr*ur.mx - sdl*ul.mx + (ptl - ptr))/(sdr*ur.d - sdl*ul.d);
  double sdml = spd[0] - spd[2];
  double sdmr = spd[4] - spd[2];
  double sdml_inv = 1.0/sdml;
  double sdmr_inv = 1.0/sdmr;
  ulst.d = ul.d*sdl*sdml_inv;
  urst.d = ur.d*sdr*sdmr_inv;
  double ulst_d_inv = 1.0/ulst.d;
  double urst_d_inv = 1.0/urst.d;
  double sqrtdl = sqrt(ulst.d);
  double sqrtdr = sqrt(urst.d);
  spd[1] = spd[2] - fabs(bxi)/sqrtdl;
  spd[3] = spd[2] + fabs(bxi)/sqrtdr;
  double ptstl = ptl + ul.d*sdl*(spd[2]-wli[IVX]);
  double ptstr = ptr + ur.d*sdr*(spd[2]-wri[IVX]);
  double ptst = 0.5*(ptstr + ptstl);
  ulst.mx = ulst.d*spd[2];
  if (fabs(ul.d*sdl*sdml-bxsq) < (SMALL_NUMBER)*ptst) {
    ulst.my = ulst.d*wli[IVY];
    ulst.mz = ulst.d*wli[IVZ];
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double tmp = bxi*(sdl - sdml)/(ul.d*sdl*sdml - bxsq);
    ulst.my = ulst.d*(wli[IVY] - ul.by*tmp);
    ulst.mz = ulst.d*(wli[IVZ] - ul.bz*tmp);
    tmp = (ul.d*SQR(sdl) - bxsq)/(ul.d*sdl*sdml - bxsq);
    ulst.by = ul.by*tmp;
    ulst.bz = ul.bz*tmp;
  }
  double vbstl = (ulst.mx*bxi+(ulst.my*ulst.by+ulst.mz*ulst.bz))*ulst_d_inv;
  ulst.e = (sdl*ul.e - ptl*wli[IVX] + ptst*spd[2] +
            bxi*(wli[IVX]*bxi + (wli[IVY]*ul.by + wli[IVZ]*ul.bz) - vbstl))*sdml_inv;
  urst.mx = urst.d*spd[2];
  if (fabs(ur.d*sdr*sdmr - bxsq) < (SMALL_NUMBER)*ptst) {
    urst.my = urst.d*wri[IVY];
    urst.mz = urst.d*wri[IVZ];
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double tmp = bxi*(sdr - sdmr)/(ur.d*sdr*sdmr - bxsq);
    urst.my = urst.d*(wri[IVY] - ur.by*tmp);
    urst.mz = urst.d*(wri[IVZ] - ur.bz*tmp);
    tmp = (ur.d*SQR(sdr) - bxsq)/(ur.d*sdr*sdmr - bxsq);
    urst.by = ur.by*tmp;
    urst.bz = ur.bz*tmp;
  }
  double vbstr = (urst.mx*bxi+(urst.my*urst.by+urst.mz*urst.bz))*urst_d_inv;
  urst.e = (sdr*ur.e - ptr*wri[IVX] + ptst*spd[2] +
            bxi*(wri[IVX]*bxi + (wri[IVY]*ur.by + wri[IVZ]*ur.bz) - vbstr))*sdmr_inv;
  if (0.5*bxsq < (SMALL_NUMBER)*ptst) {
    uldst = ulst;
    urdst = urst;
  } else {
    double invsumd = 1.0/(sqrtdl + sqrtdr);
    double bxsig = (bxi > 0.0 ? 1.0 : -1.0);
    uldst.d = ulst.d;
    urdst.d = urst.d;
    uldst.mx = ulst.mx;
    urdst.mx = urst.mx;
    double tmp = invsumd*(sqrtdl*(ulst.my*ulst_d_inv) + sqrtdr*(urst.my*urst_d_inv) +
                          bxsig*(urst.by - ulst.by));
    uldst.my = uldst.d*tmp;
    urdst.my = urdst.d*tmp;
    tmp = invsumd*(sqrtdl*(ulst.mz*ulst_d_inv) + sqrtdr*(urst.mz*urst_d_inv) +
                   bxsig*(urst.bz - ulst.bz));
    uldst.mz = uldst.d*tmp;
    urdst.mz = urdst.d*tmp;
    tmp = invsumd*(sqrtdl*urst.by + sqrtdr*ulst.by +
                   bxsig*sqrtdl*sqrtdr*((urst.my*urst_d_inv) - (ulst.my*ulst_d_inv)));
    uldst.by = urdst.by = tmp;
    tmp = invsumd*(sqrtdl*urst.bz + sqrtdr*ulst.bz +
                   bxsig*sqrtdl*sqrtdr*((urst.mz*urst_d_inv) - (ulst.mz*ulst_d_inv)));
    uldst.bz = urdst.bz = tmp;
    tmp = spd[2]*bxi + (uldst.my*uldst.by + uldst.mz*uldst.bz)/uldst.d;
    uldst.e = ulst.e - sqrtdl*bxsig*(vbstl - tmp);
    urdst.e = urst.e + sqrtdr*bxsig*(vbstr - tmp);
  }
  uldst.d = spd[1]*(uldst.d - ulst.d);
  uldst.mx = spd[1]*(uldst.mx - ulst.mx);
  uldst.my = spd[1]*(uldst.my - ulst.my);
  uldst.mz = spd[1]*(uldst.mz - ulst.mz);
  uldst.e = spd[1]*(uldst.e - ulst.e);
  uldst.by = spd[1]*(uldst.by - ulst.by);
  uldst.bz = spd[1]*(uldst.bz - ulst.bz);
  ulst.d = spd[0]*(ulst.d - ul.d);
  ulst.mx = spd[0]*(ulst.mx - ul.mx);
  ulst.my = spd[0]*(ulst.my - ul.my);
  ulst.mz = spd[0]*(ulst.mz - ul.mz);
  ulst.e = spd[0]*(ulst.e - ul.e);
  ulst.by = spd[0]*(ulst.by - ul.by);
  ulst.bz = spd[0]*(ulst.bz - ul.bz);
  urdst.d = spd[3]*(urdst.d - urst.d);
  urdst.mx = spd[3]*(urdst.mx - urst.mx);
  urdst.my = spd[3]*(urdst.my - urst.my);
  urdst.mz = spd[3]*(urdst.mz - urst.mz);
  urdst.e = spd[3]*(urdst.e - urst.e);
  urdst.by = spd[3]*(urdst.by - urst.by);
  urdst.bz = spd[3]*(urdst.bz - urst.bz);
  urst.d = spd[4]*(urst.d  - ur.d);
  urst.mx = spd[4]*(urst.mx - ur.mx);
  urst.my = spd[4]*(urst.my - ur.my);
  urst.mz = spd[4]*(urst.mz - ur.mz);
  urst.e = spd[4]*(urst.e - ur.e);
  urst.by = spd[4]*(urst.by - ur.by);
  urst.bz = spd[4]*(urst.bz - ur.bz);
  if (spd[0] >= 0.0) {
    flxi[IDN] = fl.d; flxi[IVX] = fl.mx; flxi[IVY] = fl.my; flxi[IVZ] = fl.mz;
    flxi[IEN] = fl.e; flxi[IBY] = fl.by; flxi[IBZ] = fl.bz;
  } else if (spd[4] <= 0.0) {
    flxi[IDN] = fr.d; flxi[IVX] = fr.mx; flxi[IVY] = fr.my; flxi[IVZ] = fr.mz;
    flxi[IEN] = fr.e; flxi[IBY] = fr.by; flxi[IBZ] = fr.bz;
  } else if (spd[1] >= 0.0) {
    flxi[IDN] = fl.d  + ulst.d;
    flxi[IVX] = fl.mx + ulst.mx;
    flxi[IVY] = fl.my + ulst.my;
    flxi[IVZ] = fl.mz + ulst.mz;
    flxi[IEN] = fl.e  + ulst.e;
    flxi[IBY] = fl.by + ulst.by;
    flxi[IBZ] = fl.bz + ulst.bz;
  } else if (spd[2] >= 0.0) {
    flxi[IDN] = fl.d  + ulst.d + uldst.d;
    flxi[IVX] = fl.mx + ulst.mx + uldst.mx;
    flxi[IVY] = fl.my + ulst.my + uldst.my;
    flxi[IVZ] = fl.mz + ulst.mz + uldst.mz;
    flxi[IEN] = fl.e  + ulst.e + uldst.e;
    flxi[IBY] = fl.by + ulst.by + uldst.by;
    flxi[IBZ] = fl.bz + ulst.bz + uldst.bz;
  } else if (spd[3] > 0.0) {
    flxi[IDN] = fr.d + urst.d + urdst.d;
    flxi[IVX] = fr.mx + urst.mx + urdst.mx;
    flxi[IVY] = fr.my + urst.my + urdst.my;
    flxi[IVZ] = fr.mz + urst.mz + urdst.mz;
    flxi[IEN] = fr.e + urst.e + urdst.e;
    flxi[IBY] = fr.by + urst.by + urdst.by;
    flxi[IBZ] = fr.bz + urst.bz + urdst.bz;
  } else {
    flxi[IDN] = fr.d  + urst.d;
    flxi[IVX] = fr.mx + urst.mx;
    flxi[IVY] = fr.my + urst.my;
    flxi[IVZ] = fr.mz + urst.mz;
    flxi[IEN] = fr.e  + urst.e;
    flxi[IBY] = fr.by + urst.by;
    flxi[IBZ] = fr.bz + urst.bz;
  }
}

/* src/hydro/rsolvers/mhd/lhlld.cpp:36-390 (low-dissipation HLLD) */
static void lhlld(const double *wli, const double *wri, double bxi, double gamma,
                  double dvn, double dvt, double *flxi) {
  double spd[5];
  Cons1D ul, ur, ulst, uldst, urdst, urst, fl, fr;
  double igm1 = 1.0/(gamma - 1.0);
  double bxsq = bxi*bxi;
  double pbl = 0.5*(bxsq + (SQR(wli[IBY]) + SQR(wli[IBZ])));
  double pbr = 0.5*(bxsq + (SQR(wri[IBY]) + SQR(wri[IBZ])));
  double kel = 0.5*wli[IDN]*(SQR(wli[IVX]) + (SQR(wli[IVY]) + SQR(wli[IVZ])));
  double ker = 0.5*wri[IDN]*(SQR(wri[IVX]) + (SQR(wri[IVY]) + SQR(wri[IVZ])));
  ul.d  = wli[IDN];
  ul.mx = wli[IVX]*ul.d;
  ul.my = wli[IVY]*ul.d;
  ul.mz = wli[IVZ]*ul.d;
  ul.e  = wli[IPR]*igm1 + kel + pbl;
  ul.by = wli[IBY];
  ul.bz = wli[IBZ];
  ur.d  = wri[IDN];
  ur.mx = wri[IVX]*ur.d;
  ur.my = wri[IVY]*ur.d;
  ur.mz = wri[IVZ]*ur.d;
  ur.e  = wri[IPR]*igm1 + ker + pbr;
  ur.by = wri[IBY];
  ur.bz = wri[IBZ];
  double cfl = ao_fast_speed(gamma, wli, bxi);
  double cfr = ao_fast_speed(gamma, wri, bxi);
  spd[0] = mn(wli[IVX]-cfl, wri[IVX]-cfr);
  spd[4] = mx(wli[IVX]+cfl, wri[IVX]+cfr);
  double cfmax = mx(cfl, cfr);
  double ptl = wli[IPR] + pbl;
  double ptr = wri[IPR] + pbr;
  fl.d  = ul.mx;
  fl.mx = ul.mx*wli[IVX] + ptl - bxsq;
  fl.my = ul.my*wli[IVX] - bxi*ul.by;
  fl.mz = ul.mz*wli[IVX] - bxi*ul.bz;
  fl.e  = wli[IVX]*(ul.e + ptl - bxsq) - bxi*(wli[IVY]*ul.by + wli[IVZ]*ul.bz);
  fl.by = ul.by*wli[IVX] - bxi*wli[IVY];
  fl.bz = ul.bz*wli[IVX] - bxi*wli[IVZ];
  fr.d  = ur.mx;
  fr.mx = ur.mx*wri[IVX] + ptr - bxsq;
  fr.my = ur.my*wri[IVX] - bxi*ur.by;
  fr.mz = ur.mz*wri[IVX] - bxi*ur.bz;
  fr.e  = wri[IVX]*(ur.e + ptr - bxsq) - bxi*(wri[IVY]*ur.by + wri[IVZ]*ur.bz);
  fr.by = ur.by*wri[IVX] - bxi*wri[IVY];
  fr.bz = ur.bz*wri[IVX] - bxi*wri[IVZ];
  double sdl = spd[0] - wli[IVX];
  double sdr = spd[4] - wri[IVX];
  double sdld = sdl*ul.d;
  double sdrd = sdr*ur.d;
  double th1 = mn(1.0, (cfmax-mn(dvn,0.0))/(cfmax-mn(dvt,0.0)));
  double th = th1*th1*th1*th1;
  spd[2] = (sdr*ur.mx - sdl*ul.mx + th*(ptl - ptr))/(sdrd - sdld);
  double sdml = spd[0] - spd[2];
  double sdmr = spd[4] - spd[2];
  double sdml_inv = 1.0/sdml;
  double sdmr_inv = 1.0/sdmr;
  ulst.d = sdld*sdml_inv;
  urst.d = sdrd*sdmr_inv;
  double ulst_d_inv = 1.0/ulst.d;
  double urst_d_inv = 1.0/urst.d;
  double sqrtdl = sqrt(ulst.d);
  double sqrtdr = sqrt(urst.d);
  spd[1] = spd[2] - fabs(bxi)/sqrtdl;
  spd[3] = spd[2] + fabs(bxi)/sqrtdr;
  double clsq = ((pbl + kel) + sqrt(SQR(pbl + kel) - 2.0*kel*bxsq))/ul.d;
  double crsq = ((pbr + ker) + sqrt(SQR(pbr + ker) - 2.0*ker*bxsq))/ur.d;
  double chi = mn(1.0, sqrt(mx(clsq, crsq))/cfmax);
  double phi = chi*(2.0 - chi);
  double ptst = (sdrd*ptl - sdld*ptr + phi*sdrd*sdld*(wri[IVX]-wli[IVX]))/(sdrd - sdld);
  ulst.mx = ulst.d*spd[2];
  if (fabs(sdld*sdml-bxsq) < (SMALL_NUMBER)*ptst) {
    ulst.my = ulst.d*wli[IVY];
    ulst.mz = ulst.d*wli[IVZ];
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double tmp = bxi*(sdl - sdml)/(sdld*sdml - bxsq);
    ulst.my = ulst.d*(wli[IVY] - ul.by*tmp);
    ulst.mz = ulst.d*(wli[IVZ] - ul.bz*tmp);
    tmp = (sdld*sdl - bxsq)/(sdld*sdml - bxsq);
    ulst.by = ul.by*tmp;
    ulst.bz = ul.bz*tmp;
  }
  double vbstl = (ulst.mx*bxi+(ulst.my*ulst.by+ulst.mz*ulst.bz))*ulst_d_inv;
  ulst.e = (sdl*ul.e - ptl*wli[IVX] + ptst*spd[2] +
            bxi*(wli[IVX]*bxi + (wli[IVY]*ul.by + wli[IVZ]*ul.bz) - vbstl))*sdml_inv;
  urst.mx = urst.d*spd[2];
  if (fabs(sdrd*sdmr - bxsq) < (SMALL_NUMBER)*ptst) {
    urst.my = urst.d*wri[IVY];
    urst.mz = urst.d*wri[IVZ];
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double tmp = bxi*(sdr - sdmr)/(sdrd*sdmr - bxsq);
    urst.my = urst.d*(wri[IVY] - ur.by*tmp);
    urst.mz = urst.d*(wri[IVZ] - ur.bz*tmp);
    tmp = (sdrd*sdr - bxsq)/(sdrd*sdmr - bxsq);
    urst.by = ur.by*tmp;
    urst.bz = ur.bz*tmp;
  }
  double vbstr = (urst.mx*bxi+(urst.my*urst.by+urst.mz*urst.bz))*urst_d_inv;
  urst.e = (sdr*ur.e - ptr*wri[IVX] + ptst*spd[2] +
            bxi*(wri[IVX]*bxi + (wri[IVY]*ur.by + wri[IVZ]*ur.bz) - vbstr))*sdmr_inv;
  if (0.5*bxsq < (SMALL_NUMBER)*ptst) {
    uldst = ulst;
    urdst = urst;
  } else {
    double invsumd = 1.0/(sqrtdl + sqrtdr);
    double bxsig = (bxi > 0.0 ? 1.0 : -1.0);
    uldst.d = ulst.d;
    urdst.d = urst.d;
    uldst.mx = ulst.mx;
    urdst.mx = urst.mx;
    double tmp = invsumd*(sqrtdl*(ulst.my*ulst_d_inv) + sqrtdr*(urst.my*urst_d_inv) +
                          bxsig*(urst.by - ulst.by));
    uldst.my = uldst.d*tmp;
    urdst.my = urdst.d*tmp;
    tmp = invsumd*(sqrtdl*(ulst.mz*ulst_d_inv) + sqrtdr*(urst.mz*urst_d_inv) +
                   bxsig*(urst.bz - ulst.bz));
    uldst.mz = uldst.d*tmp;
    urdst.mz = urdst.d*tmp;
    tmp = invsumd*(sqrtdl*urst.by + sqrtdr*ulst.by +
                   bxsig*sqrtdl*sqrtdr*((urst.my*urst_d_inv) - (ulst.my*ulst_d_inv)));
    uldst.by = urdst.by = tmp;
    tmp = invsumd*(sqrtdl*urst.bz + sqrtdr*ulst.bz +
                   bxsig*sqrtdl*sqrtdr*((urst.mz*urst_d_inv) - (ulst.mz*ulst_d_inv)));
    uldst.bz = urdst.bz = tmp;
    tmp = spd[2]*bxi + (uldst.my*uldst.by + uldst.mz*uldst.bz)/uldst.d;
    uldst.e = ulst.e - sqrtdl*bxsig*(vbstl - tmp);
    urdst.e = urst.e + sqrtdr*bxsig*(vbstr - tmp);
  }
  uldst.d = spd[1]*(uldst.d - ulst.d);
  uldst.mx = spd[1]*(uldst.mx - ulst.mx);
  uldst.my = spd[1]*(uldst.my - ulst.my);
  uldst.mz = spd[1]*(uldst.mz - ulst.mz);
  uldst.e = spd[1]*(uldst.e - ulst.e);
  uldst.by = spd[1]*(uldst.by - ulst.by);
  uldst.bz = spd[1]*(uldst.bz - ulst.bz);
  ulst.d = spd[0]*(ulst.d - ul.d);
  ulst.mx = spd[0]*(ulst.mx - ul.mx);
  ulst.my = spd[0]*(ulst.my - ul.my);
  ulst.mz = spd[0]*(ulst.mz - ul.mz);
  ulst.e = spd[0]*(ulst.e - ul.e);
  ulst.by = spd[0]*(ulst.by - ul.by);
  ulst.bz = spd[0]*(ulst.bz - ul.bz);
  urdst.d = spd[3]*(urdst.d - urst.d);
  urdst.mx = spd[3]*(urdst.mx - urst.mx);
  urdst.my = spd[3]*(urdst.my - urst.my);
  urdst.mz = spd[3]*(urdst.mz - urst.mz);
  urdst.e = spd[3]*(urdst.e - urst.e);
  urdst.by = spd[3]*(urdst.by - urst.by);
  urdst.bz = spd[3]*(urdst.bz - urst.bz);
  urst.d = spd[4]*(urst.d  - ur.d);
  urst.mx = spd[4]*(urst.mx - ur.mx);
  urst.my = spd[4]*(urst.my - ur.my);
  urst.mz = spd[4]*(urst.mz - ur.mz);
  urst.e = spd[4]*(urst.e - ur.e);
  urst.by = spd[4]*(urst.by - ur.by);
  urst.bz = spd[4]*(urst.bz - ur.bz);
  if (spd[0] >= 0.0) {
    flxi[IDN] = fl.d; flxi[IVX] = fl.mx; flxi[IVY] = fl.my; flxi[IVZ] = fl.mz;
    flxi[IEN] = fl.e; flxi[IBY] = fl.by; flxi[IBZ] = fl.bz;
  } else if (spd[4] <= 0.0) {
    flxi[IDN] = fr.d; flxi[IVX] = fr.mx; flxi[IVY] = fr.my; flxi[IVZ] = fr.mz;
    flxi[IEN] = fr.e; flxi[IBY] = fr.by; flxi[IBZ] = fr.bz;
  } else if (spd[1] >= 0.0) {
    flxi[IDN] = fl.d  + ulst.d;
    flxi[IVX] = fl.mx + ulst.mx;
    flxi[IVY] = fl.my + ulst.my;
    flxi[IVZ] = fl.mz + ulst.mz;
    flxi[IEN] = fl.e  + ulst.e;
    flxi[IBY] = fl.by + ulst.by;
    flxi[IBZ] = fl.bz + ulst.bz;
  } else if (spd[2] >= 0.0) {
    flxi[IDN] = fl.d  + ulst.d + uldst.d;
    flxi[IVX] = fl.mx + ulst.mx + uldst.mx;
    flxi[IVY] = fl.my + ulst.my + uldst.my;
    flxi[IVZ] = fl.mz + ulst.mz + uldst.mz;
    flxi[IEN] = fl.e  + ulst.e + uldst.e;
    flxi[IBY] = fl.by + ulst.by + uldst.by;
    flxi[IBZ] = fl.bz + ulst.bz + uldst.bz;
  } else if (spd[3] > 0.0) {
    flxi[IDN] = fr.d + urst.d + urdst.d;
    flxi[IVX] = fr.mx + urst.mx + urdst.mx;
    flxi[IVY] = fr.my + urst.my + urdst.my;
    flxi[IVZ] = fr.mz + urst.mz + urdst.mz;
    flxi[IEN] = fr.e + urst.e + urdst.e;
    flxi[IBY] = fr.by + urst.by + urdst.by;
    flxi[IBZ] = fr.bz + urst.bz + urdst.bz;
  } else {
    flxi[IDN] = fr.d  + urst.d;
    flxi[IVX] = fr.mx + urst.mx;
    flxi[IVY] = fr.my + urst.my;
    flxi[IVZ] = fr.mz + urst.mz;
    flxi[IEN] = fr.e  + urst.e;
    flxi[IBY] = fr.by + urst.by;
    flxi[IBZ] = fr.bz + urst.bz;
  }
}

/* Roe averages shared by hlle_mhd.cpp:64-85 and roe_mhd.cpp:92-113 */
static void roe_avg_mhd(const double *wli, const double *wri, double bxi, double gm1,
                        double *wroe, double *x, double *y, double *pbl, double *pbr,
                        double *el, double *er, double *hroe) {
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  wroe[IBY] = (sqrtdr*wli[IBY] + sqrtdl*wri[IBY])*isdlpdr;
  wroe[IBZ] = (sqrtdr*wli[IBZ] + sqrtdl*wri[IBZ])*isdlpdr;
  *x = 0.5*(SQR(wli[IBY]-wri[IBY]) + SQR(wli[IBZ]-wri[IBZ]))/(SQR(sqrtdl+sqrtdr));
  *y = 0.5*(wli[IDN] + wri[IDN])/wroe[IDN];
  *pbl = 0.5*(bxi*bxi + SQR(wli[IBY]) + SQR(wli[IBZ]));
  *pbr = 0.5*(bxi*bxi + SQR(wri[IBY]) + SQR(wri[IBZ]));
  *el = wli[IPR]/gm1 + 0.5*wli[IDN]*(SQR(wli[IVX])+SQR(wli[IVY])+SQR(wli[IVZ])) + *pbl;
  *er = wri[IPR]/gm1 + 0.5*wri[IDN]*(SQR(wri[IVX])+SQR(wri[IVY])+SQR(wri[IVZ])) + *pbr;
  *hroe = ((*el + wli[IPR] + *pbl)/sqrtdl + (*er + wri[IPR] + *pbr)/sqrtdr)*isdlpdr;
}

/* src/hydro/rsolvers/mhd/hlle_mhd.cpp:25-182 (adiabatic branch) */
static void hlle_mhd(const double *wli, const double *wri, double bxi, double gamma,
                     double *flxi) {
  double wroe[7], fl[7], fr[7];
  double gm1 = gamma - 1.0;
  double x, y, pbl, pbr, el, er, hroe;
  roe_avg_mhd(wli, wri, bxi, gm1, wroe, &x, &y, &pbl, &pbr, &el, &er, &hroe);
  double cl = ao_fast_speed(gamma, wli, bxi);
  double cr = ao_fast_speed(gamma, wri, bxi);
  double btsq = SQR(wroe[IBY]) + SQR(wroe[IBZ]);
  double vaxsq = bxi*bxi/wroe[IDN];
  double bt_starsq = (gm1 - (gm1 - 1.0)*y)*btsq;
  double hp = hroe - (vaxsq + btsq/wroe[IDN]);
  double vsq = SQR(wroe[IVX]) + SQR(wroe[IVY]) + SQR(wroe[IVZ]);
  double twid_asq = mx((gm1*(hp-0.5*vsq)-(gm1-1.0)*x), 0.0);
  double ct2 = bt_starsq/wroe[IDN];
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_asq*ct2);
  double cfsq = 0.5*(tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = mn((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = mx((wroe[IVX] + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + pbl - SQR(bxi);
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + pbr - SQR(bxi);
  fl[IVY] = wli[IDN]*wli[IVY]*vxl - bxi*wli[IBY];
  fr[IVY] = wri[IDN]*wri[IVY]*vxr - bxi*wri[IBY];
  fl[IVZ] = wli[IDN]*wli[IVZ]*vxl - bxi*wli[IBZ];
  fr[IVZ] = wri[IDN]*wri[IVZ]*vxr - bxi*wri[IBZ];
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = el*vxl + wli[IVX]*(wli[IPR] + pbl - bxi*bxi);
  fr[IEN] = er*vxr + wri[IVX]*(wri[IPR] + pbr - bxi*bxi);
  fl[IEN] -= bxi*(wli[IBY]*wli[IVY] + wli[IBZ]*wli[IVZ]);
  fr[IEN] -= bxi*(wri[IBY]*wri[IVY] + wri[IBZ]*wri[IVZ]);
  fl[IBY] = wli[IBY]*vxl - bxi*wli[IVY];
  fr[IBY] = wri[IBY]*vxr - bxi*wri[IVY];
  fl[IBZ] = wli[IBZ]*vxl - bxi*wli[IVZ];
  fr[IBZ] = wri[IBZ]*vxr - bxi*wri[IVZ];
  double tmp = 0.0;
  if (bp != bm) tmp = 0.5*(bp + bm)/(bp - bm);
  for (int n = 0; n < 7; ++n)
    flxi[n] = 0.5*(fl[n]+fr[n]) + (fl[n]-fr[n])*tmp;
}

/* src/hydro/rsolvers/mhd/roe_mhd.cpp:245-612 (adiabatic branch of RoeFlux) */
static void roe_flux_mhd(const double *wroe, double b1, double x, double y,
                         const double *du, const double *wli, double gm1,
                         double *flx, double *ev, int *llf_flag) {
  double d  = wroe[IDN];
  double v1 = wroe[IVX];
  double v2 = wroe[IVY];
  double v3 = wroe[IVZ];
  double b2 = wroe[IBY];
  double b3 = wroe[IBZ];
  double di = 1.0/d;
  double btsq = b2*b2 + b3*b3;
  double vaxsq = b1*b1*di;
  double vsq = v1*v1 + v2*v2 + v3*v3;
  double hp = wroe[IPR] - (vaxsq + btsq*di);
  double bt_starsq = (gm1 - (gm1 - 1.0)*y)*btsq;
  double twid_csq = mx((gm1*(hp-0.5*vsq)-(gm1-1.0)*x), TINY_NUMBER);
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
  double s = SIGN(b1);
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

  ev[0] = v1 - cf;
  ev[1] = v1 - vax;
  ev[2] = v1 - cs;
  ev[3] = v1;
  ev[4] = v1 + cs;
  ev[5] = v1 + vax;
  ev[6] = v1 + cf;

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
  for (int n = 0; n < 7; ++n) coeff[n] = -0.5*fabs(ev[n])*a[n];

  double dens = wli[IDN] + a[0]*alpha_f;
  if (dens < 0.0) *llf_flag = 1;
  dens += a[2]*alpha_s;
  if (dens < 0.0) *llf_flag = 1;
  dens += a[3];
  if (dens < 0.0) *llf_flag = 1;
  dens += a[4]*alpha_s;
  if (dens < 0.0) *llf_flag = 1;

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

/* src/hydro/rsolvers/mhd/roe_mhd.cpp:43-238 */
static void roe_mhd(const double *wli, const double *wri, double bxi, double gamma,
                    double *flxi) {
  double wroe[7], fl[7], fr[7], ev[7], du[7];
  double gm1 = gamma - 1.0;
  double x, y, pbl, pbr, el, er, hroe;
  roe_avg_mhd(wli, wri, bxi, gm1, wroe, &x, &y, &pbl, &pbr, &el, &er, &hroe);
  wroe[IPR] = hroe;
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  fl[IVX] = mxl*wli[IVX] + pbl - SQR(bxi);
  fr[IVX] = mxr*wri[IVX] + pbr - SQR(bxi);
  fl[IVY] = mxl*wli[IVY] - bxi*wli[IBY];
  fr[IVY] = mxr*wri[IVY] - bxi*wri[IBY];
  fl[IVZ] = mxl*wli[IVZ] - bxi*wli[IBZ];
  fr[IVZ] = mxr*wri[IVZ] - bxi*wri[IBZ];
  fl[IVX] += wli[IPR];
  fr[IVX] += wri[IPR];
  fl[IEN] = (el + wli[IPR] + pbl - bxi*bxi)*wli[IVX];
  fr[IEN] = (er + wri[IPR] + pbr - bxi*bxi)*wri[IVX];
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
  du[IEN] = er - el;
  du[IBY] = wri[IBY] - wli[IBY];
  du[IBZ] = wri[IBZ] - wli[IBZ];
  for (int n = 0; n < 7; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  roe_flux_mhd(wroe, bxi, x, y, du, wli, gm1, flxi, ev, &llf_flag);
  if (ev[0] >= 0.0) for (int n = 0; n < 7; ++n) flxi[n] = fl[n];
  if (ev[6] <= 0.0) for (int n = 0; n < 7; ++n) flxi[n] = fr[n];
  if (llf_flag != 0) {
    double cfl = ao_fast_speed(gamma, wli, bxi);
    double cfr = ao_fast_speed(gamma, wri, bxi);
    double a = 0.5*mx((fabs(wli[IVX]) + cfl), (fabs(wri[IVX]) + cfr));
    /* the reference's LLF fallback leaves the IBY/IBZ fluxes untouched
       (roe_mhd.cpp:222-232) */
    for (int n = 0; n < 5; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
}

/* ---------------------------------------------------------------- isothermal EOS
 * (NON_BAROTROPIC_EOS == 0: NHYDRO = 4, no IPR/IEN; the sweep-ordered work arrays below keep
 * the 7-slot layout of the adiabatic code with slot 4 unused) */

/* src/eos/isothermal_mhd.cpp:122-129 */
double ao_fast_speed_iso(double cs, const double *prim, double bx) {
  double asq = (cs*cs)*prim[IDN];
  double vaxsq = bx*bx;
  double ct2 = prim[IBY]*prim[IBY] + prim[IBZ]*prim[IBZ];
  double qsq = vaxsq + ct2 + asq;
  double tmp = vaxsq + ct2 - asq;
  return sqrt(0.5*(qsq + sqrt(tmp*tmp + 4.0*asq*ct2))/prim[IDN]);
}

/* src/hydro/rsolvers/hydro/hlle.cpp:38-162 (isothermal branch) */
static void hlle_hydro_iso(const double *wli, const double *wri, double iso_cs, double *flxi) {
  double wroe[5], fl[5], fr[5];
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double cl = iso_cs, cr = iso_cs, a = iso_cs;
  double al = mn((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = mx((wroe[IVX] + a), (wri[IVX] + cr));
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

/* src/hydro/rsolvers/mhd/hlle_mhd.cpp:25-182 (isothermal branch) */
static void hlle_mhd_iso(const double *wli, const double *wri, double bxi, double iso_cs,
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
  double x = 0.5*(SQR(wli[IBY]-wri[IBY]) + SQR(wli[IBZ]-wri[IBZ]))/(SQR(sqrtdl+sqrtdr));
  double y = 0.5*(wli[IDN] + wri[IDN])/wroe[IDN];
  double pbl = 0.5*(bxi*bxi + SQR(wli[IBY]) + SQR(wli[IBZ]));
  double pbr = 0.5*(bxi*bxi + SQR(wri[IBY]) + SQR(wri[IBZ]));
  double cl = ao_fast_speed_iso(iso_cs, wli, bxi);
  double cr = ao_fast_speed_iso(iso_cs, wri, bxi);
  double btsq = SQR(wroe[IBY]) + SQR(wroe[IBZ]);
  double vaxsq = bxi*bxi/wroe[IDN];
  double bt_starsq = btsq*y;
  double twid_asq = iso_cs*iso_cs + x;
  double ct2 = bt_starsq/wroe[IDN];
  double tsum = vaxsq + ct2 + twid_asq;
  double tdif = vaxsq + ct2 - twid_asq;
  double cf2_cs2 = sqrt(tdif*tdif + 4.0*twid_asq*ct2);
  double cfsq = 0.5*(tsum + cf2_cs2);
  double a = sqrt(cfsq);
  double al = mn((wroe[IVX] - a), (wli[IVX] - cl));
  double ar = mx((wroe[IVX] + a), (wri[IVX] + cr));
  double bp = ar > 0.0 ? ar : 0.0;
  double bm = al < 0.0 ? al : 0.0;
  double vxl = wli[IVX] - bm;
  double vxr = wri[IVX] - bp;
  fl[IDN] = wli[IDN]*vxl;
  fr[IDN] = wri[IDN]*vxr;
  fl[IVX] = wli[IDN]*wli[IVX]*vxl + pbl - SQR(bxi);
  fr[IVX] = wri[IDN]*wri[IVX]*vxr + pbr - SQR(bxi);
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

/* src/hydro/rsolvers/mhd/hlld_iso.cpp:36-288 (Mignone 2007) */
static void hlld_iso(const double *wli, const double *wri, double bxi, double cs, double dfloor,
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
  double cfl = ao_fast_speed_iso(cs, wli, bxi);
  double cfr = ao_fast_speed_iso(cs, wri, bxi);
  spd[0] = mn(wli[IVX]-cfl, wri[IVX]-cfr);
  spd[4] = mx(wli[IVX]+cfl, wri[IVX]+cfr);
  double bxsq = bxi*bxi;
  double ptl = SQR(cs)*wli[IDN] + 0.5*(bxsq + SQR(wli[IBY]) + SQR(wli[IBZ]));
  double ptr = SQR(cs)*wri[IDN] + 0.5*(bxsq + SQR(wri[IBY]) + SQR(wri[IBZ]));
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
  dhll = mx(dhll, dfloor);
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
  if (fabs(spd[0]-spd[1]) < (SMALL_NUMBER)*cs) {
    ulst.my = ul.my;
    ulst.mz = ul.mz;
    ulst.by = ul.by;
    ulst.bz = ul.bz;
  } else {
    double mfact = bxi*(ustar-wli[IVX])/tmp;
    double bfact = (ul.d*SQR(spd[0]-wli[IVX]) - bxsq)/(dhll*tmp);
    ulst.my = dhll*wli[IVY] - ul.by*mfact;
    ulst.mz = dhll*wli[IVZ] - ul.bz*mfact;
    ulst.by = ul.by*bfact;
    ulst.bz = ul.bz*bfact;
  }
  urst.d  = dhll;
  urst.mx = mxhll;
  tmp = (spd[4]-spd[1])*(spd[4]-spd[3]);
  if (fabs(spd[4]-spd[3]) < (SMALL_NUMBER)*cs) {
    urst.my = ur.my;
    urst.mz = ur.mz;
    urst.by = ur.by;
    urst.bz = ur.bz;
  } else {
    double mfact = bxi*(ustar-wri[IVX])/tmp;
    double bfact = (ur.d*SQR(spd[4]-wri[IVX]) - bxsq)/(dhll*tmp);
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

/* LLF: src/hydro/rsolvers/hydro/llf.cpp:34-125 and mhd/llf_mhd.cpp:34-170 (both EOS) */
static void llf(int mhd, int iso, const double *wli, const double *wri, double bxi,
                double gamma, double iso_cs, double *flxi) {
  double fl[7], fr[7], du[7];
  double gm1 = gamma - 1.0;
  double cl, cr;
  if (mhd) {
    cl = iso ? ao_fast_speed_iso(iso_cs, wli, bxi) : ao_fast_speed(gamma, wli, bxi);
    cr = iso ? ao_fast_speed_iso(iso_cs, wri, bxi) : ao_fast_speed(gamma, wri, bxi);
  } else {
    cl = iso ? iso_cs : ao_sound_speed(gamma, wli);
    cr = iso ? iso_cs : ao_sound_speed(gamma, wri);
  }
  double a = 0.5*mx((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  double pbl = 0.0, pbr = 0.0;
  fl[IDN] = mxl;
  fr[IDN] = mxr;
  if (mhd) {
    pbl = 0.5*(bxi*bxi + SQR(wli[IBY]) + SQR(wli[IBZ]));
    pbr = 0.5*(bxi*bxi + SQR(wri[IBY]) + SQR(wri[IBZ]));
    fl[IVX] = mxl*wli[IVX] + pbl - SQR(bxi);
    fr[IVX] = mxr*wri[IVX] + pbr - SQR(bxi);
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
  if (!iso) {
    if (mhd) {
      el = wli[IPR]/gm1 + 0.5*wli[IDN]*(SQR(wli[IVX])+SQR(wli[IVY])+SQR(wli[IVZ])) + pbl;
      er = wri[IPR]/gm1 + 0.5*wri[IDN]*(SQR(wri[IVX])+SQR(wri[IVY])+SQR(wri[IVZ])) + pbr;
    } else {
      el = wli[IPR]/gm1 + 0.5*wli[IDN]*(SQR(wli[IVX]) + SQR(wli[IVY]) + SQR(wli[IVZ]));
      er = wri[IPR]/gm1 + 0.5*wri[IDN]*(SQR(wri[IVX]) + SQR(wri[IVY]) + SQR(wri[IVZ]));
    }
    fl[IVX] += wli[IPR];
    fr[IVX] += wri[IPR];
    if (mhd) {
      fl[IEN] = (el + wli[IPR] + pbl - bxi*bxi)*wli[IVX];
      fr[IEN] = (er + wri[IPR] + pbr - bxi*bxi)*wri[IVX];
      fl[IEN] -= bxi*(wli[IBY]*wli[IVY] + wli[IBZ]*wli[IVZ]);
      fr[IEN] -= bxi*(wri[IBY]*wri[IVY] + wri[IBZ]*wri[IVZ]);
    } else {
      fl[IEN] = (el + wli[IPR])*wli[IVX];
      fr[IEN] = (er + wri[IPR])*wri[IVX];
    }
  } else {
    fl[IVX] += (iso_cs*iso_cs)*wli[IDN];
    fr[IVX] += (iso_cs*iso_cs)*wri[IDN];
  }
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  du[IEN] = iso ? 0.0 : (er - el);
  int nw = 5;
  if (mhd) {
    fl[IBY] = wli[IBY]*wli[IVX] - bxi*wli[IVY];
    fr[IBY] = wri[IBY]*wri[IVX] - bxi*wri[IVY];
    fl[IBZ] = wli[IBZ]*wli[IVX] - bxi*wli[IVZ];
    fr[IBZ] = wri[IBZ]*wri[IVX] - bxi*wri[IVZ];
    du[IBY] = wri[IBY] - wli[IBY];
    du[IBZ] = wri[IBZ] - wli[IBZ];
    nw = 7;
  }
  for (int n = 0; n < nw; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  if (iso) flxi[IEN] = 0.0;
}

/* ---- Roe's solver, isothermal EOS (the NON_BAROTROPIC_EOS == 0 branches of
 * src/hydro/rsolvers/hydro/roe.cpp:42-352 and src/hydro/rsolvers/mhd/roe_mhd.cpp:43-613) */
static void roe_hydro_iso(const double *wli, const double *wri, double iso_cs, double *flxi) {
  double wroe[5], fl[5], fr[5], ev[4], du[5];
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  fl[IDN] = mxl; fr[IDN] = mxr;
  fl[IVX] = mxl*wli[IVX]; fr[IVX] = mxr*wri[IVX];
  fl[IVY] = mxl*wli[IVY]; fr[IVY] = mxr*wri[IVY];
  fl[IVZ] = mxl*wli[IVZ]; fr[IVZ] = mxr*wri[IVZ];
  fl[IVX] += (iso_cs*iso_cs)*wli[IDN];
  fr[IVX] += (iso_cs*iso_cs)*wri[IDN];
  du[IDN] = wri[IDN]          - wli[IDN];
  du[IVX] = wri[IDN]*wri[IVX] - wli[IDN]*wli[IVX];
  du[IVY] = wri[IDN]*wri[IVY] - wli[IDN]*wli[IVY];
  du[IVZ] = wri[IDN]*wri[IVZ] - wli[IDN]*wli[IVZ];
  for (int n = 0; n < 4; ++n) flxi[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  {  /* RoeFlux, isothermal hydrodynamics (roe.cpp:300-349) */
    double v1 = wroe[IVX], v2 = wroe[IVY], v3 = wroe[IVZ];
    ev[0] = v1 - iso_cs; ev[1] = v1; ev[2] = v1; ev[3] = v1 + iso_cs;
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
  if (ev[0] >= 0.0) for (int n = 0; n < 4; ++n) flxi[n] = fl[n];
  if (ev[3] <= 0.0) for (int n = 0; n < 4; ++n) flxi[n] = fr[n];
  if (llf_flag != 0) {
    double cl = iso_cs, cr = iso_cs;      /* EquationOfState::SoundSpeed, isothermal */
    double a = 0.5*mx((fabs(wli[IVX]) + cl), (fabs(wri[IVX]) + cr));
    for (int n = 0; n < 4; ++n) flxi[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
  flxi[IEN] = 0.0;
}

/* wli / wri / flxi keep the 7-slot layout (IBY = 5, IBZ = 6, slot 4 unused); inside, the six
 * waves are numbered as in the reference's isothermal build (IBY = 4, IBZ = 5) */
static void roe_mhd_iso(const double *wli, const double *wri, double bxi, double iso_cs,
                        double *flxi) {
  double wroe[7], fl[6], fr[6], ev[6], du[6], flx[6];
  double sqrtdl = sqrt(wli[IDN]);
  double sqrtdr = sqrt(wri[IDN]);
  double isdlpdr = 1.0/(sqrtdl + sqrtdr);
  wroe[IDN] = sqrtdl*sqrtdr;
  wroe[IVX] = (sqrtdl*wli[IVX] + sqrtdr*wri[IVX])*isdlpdr;
  wroe[IVY] = (sqrtdl*wli[IVY] + sqrtdr*wri[IVY])*isdlpdr;
  wroe[IVZ] = (sqrtdl*wli[IVZ] + sqrtdr*wri[IVZ])*isdlpdr;
  wroe[IBY] = (sqrtdr*wli[IBY] + sqrtdl*wri[IBY])*isdlpdr;
  wroe[IBZ] = (sqrtdr*wli[IBZ] + sqrtdl*wri[IBZ])*isdlpdr;
  double x = 0.5*(SQR(wli[IBY]-wri[IBY]) + SQR(wli[IBZ]-wri[IBZ]))/(SQR(sqrtdl+sqrtdr));
  double y = 0.5*(wli[IDN] + wri[IDN])/wroe[IDN];
  double pbl = 0.5*(bxi*bxi + SQR(wli[IBY]) + SQR(wli[IBZ]));
  double pbr = 0.5*(bxi*bxi + SQR(wri[IBY]) + SQR(wri[IBZ]));
  double mxl = wli[IDN]*wli[IVX];
  double mxr = wri[IDN]*wri[IVX];
  fl[0] = mxl; fr[0] = mxr;
  fl[1] = mxl*wli[IVX] + pbl - SQR(bxi);
  fr[1] = mxr*wri[IVX] + pbr - SQR(bxi);
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
  for (int n = 0; n < 6; ++n) flx[n] = 0.5*(fl[n] + fr[n]);
  int llf_flag = 0;
  {  /* RoeFlux, isothermal MHD (roe_mhd.cpp:245-345,498-612) */
    double d = wroe[IDN], v1 = wroe[IVX], v2 = wroe[IVY], v3 = wroe[IVZ];
    double b1 = bxi, b2 = wroe[IBY], b3 = wroe[IBZ];
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
    if (bt != 0.0) { bet2 = b2/bt; bet3 = b3/bt; }
    double bet2_star = bet2/sqrt(y);
    double bet3_star = bet3/sqrt(y);
    double bet_starsq = bet2_star*bet2_star + bet3_star*bet3_star;
    double q2_star = 0.0, q3_star = 0.0;
    if (bet_starsq != 0.0) { q2_star = bet2_star/bet_starsq; q3_star = bet3_star/bet_starsq; }
    double alpha_f, alpha_s;
    if ((cfsq - cssq) <= 0.0) { alpha_f = 1.0; alpha_s = 0.0; }
    else if ((twid_csq - cssq) <= 0.0) { alpha_f = 0.0; alpha_s = 1.0; }
    else if ((cfsq - twid_csq) <= 0.0) { alpha_f = 1.0; alpha_s = 0.0; }
    else {
      alpha_f = sqrt((twid_csq - cssq)/(cfsq - cssq));
      alpha_s = sqrt((cfsq - twid_csq)/(cfsq - cssq));
    }
    double sqrtd = sqrt(d);
    double isqrtd = 1.0/sqrtd;
    double s = SIGN(b1);
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
    ev[0] = v1 - cf; ev[1] = v1 - vax; ev[2] = v1 - cs;
    ev[3] = v1 + cs; ev[4] = v1 + vax; ev[5] = v1 + cf;
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
  if (ev[0] >= 0.0) for (int n = 0; n < 6; ++n) flx[n] = fl[n];
  if (ev[5] <= 0.0) for (int n = 0; n < 6; ++n) flx[n] = fr[n];
  if (llf_flag != 0) {
    double cfl = ao_fast_speed_iso(iso_cs, wli, bxi);
    double cfr = ao_fast_speed_iso(iso_cs, wri, bxi);
    double a = 0.5*mx((fabs(wli[IVX]) + cfl), (fabs(wri[IVX]) + cfr));
    /* the fallback leaves the field fluxes untouched (roe_mhd.cpp:222-232) */
    for (int n = 0; n < 4; ++n) flx[n] = 0.5*(fl[n] + fr[n]) - a*du[n];
  }
  flxi[IDN] = flx[0]; flxi[IVX] = flx[1]; flxi[IVY] = flx[2]; flxi[IVZ] = flx[3];
  flxi[IEN] = 0.0; flxi[IBY] = flx[4]; flxi[IBZ] = flx[5];
}

/* One interface, isothermal EOS (configure.py:299-325 allows hlle [hydro], hlle / hlld [MHD];
 * plus roe and llf with either) */
void ao_riemann_point_iso(int solver, int mhd, const double *wli, const double *wri,
                          double bxi, double iso_cs, double dfloor, double *flxi) {
  if (solver == AO_SOLVER_LLF) llf(mhd, 1, wli, wri, bxi, 0.0, iso_cs, flxi);
  else if (solver == AO_SOLVER_ROE) { if (mhd) roe_mhd_iso(wli, wri, bxi, iso_cs, flxi); else roe_hydro_iso(wli, wri, iso_cs, flxi); }
  else if (!mhd) hlle_hydro_iso(wli, wri, iso_cs, flxi);
  else if (solver == AO_SOLVER_HLLD) hlld_iso(wli, wri, bxi, iso_cs, dfloor, flxi);
  else hlle_mhd_iso(wli, wri, bxi, iso_cs, flxi);
}

/* One interface.  wli/wri in sweep-rotated order (IDN,ivx,ivy,ivz,IPR[,IBY,IBZ]). */
void ao_riemann_point(int solver, int mhd, const double *wli, const double *wri,
                      double bxi, double gamma, double dvn, double dvt, double *flxi) {
  if (solver == AO_SOLVER_LLF) { llf(mhd, 0, wli, wri, bxi, gamma, 0.0, flxi); return; }
  if (!mhd) {
    if (solver == AO_SOLVER_LHLLC) lhllc(wli, wri, gamma, dvn, dvt, flxi);
    else if (solver == AO_SOLVER_HLLC) hllc(wli, wri, gamma, flxi);
    else if (solver == AO_SOLVER_HLLE) hlle_hydro(wli, wri, gamma, flxi);
    else roe_hydro(wli, wri, gamma, flxi);
  } else {
    if (solver == AO_SOLVER_LHLLD) lhlld(wli, wri, bxi, gamma, dvn, dvt, flxi);
    else if (solver == AO_SOLVER_HLLD) hlld(wli, wri, bxi, gamma, flxi);
    else if (solver == AO_SOLVER_HLLE) hlle_mhd(wli, wri, bxi, gamma, flxi);
    else roe_mhd(wli, wri, bxi, gamma, flxi);
  }
}

void ao_riemann_iso(int solver, int mhd, long n, const double *wl, const double *wr,
                    const double *bx, double iso_cs, double dfloor, double *flx) {
  int nw = mhd ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double wli[7], wri[7], f[7];
    for (int v = 0; v < nw; ++v) { wli[v] = wl[v*n+i]; wri[v] = wr[v*n+i]; }
    ao_riemann_point_iso(solver, mhd, wli, wri, mhd ? bx[i] : 0.0, iso_cs, dfloor, f);
    for (int v = 0; v < nw; ++v) flx[v*n+i] = f[v];
  }
}

void ao_riemann(int solver, int mhd, long n, const double *wl, const double *wr,
                const double *bx, double gamma, double dt, double dx,
                double *flx, double *wct) {
  ao_riemann_dv(solver, mhd, n, wl, wr, bx, 0, 0, gamma, dt, dx, flx, wct);
}

void ao_riemann_dv(int solver, int mhd, long n, const double *wl, const double *wr,
                   const double *bx, const double *dvn, const double *dvt, double gamma,
                   double dt, double dx, double *flx, double *wct) {
  int nw = mhd ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double wli[7], wri[7], f[7];
    for (int v = 0; v < nw; ++v) { wli[v] = wl[v*n+i]; wri[v] = wr[v*n+i]; }
    ao_riemann_point(solver, mhd, wli, wri, mhd ? bx[i] : 0.0, gamma, dvn ? dvn[i] : 0.0,
                     dvt ? dvt[i] : 0.0, f);
    for (int v = 0; v < nw; ++v) flx[v*n+i] = f[v];
    if (mhd && wct) wct[i] = ao_weight_for_ct(f[IDN], wli[IDN], wri[IDN], dx, dt);
  }
}

/* ---------------------------------------------------------------- characteristic projection
 * src/reconstruct/characteristic.cpp:36-278 (LeftEigenmatrixDotVector) and :280-520
 * (RightEigenmatrixDotVector), adiabatic hydro and adiabatic MHD, in sweep order: w = the
 * cell's primitives (IDN, vx, vy, vz, IPR[, By, Bz]) with vx along the sweep, bx = the
 * cell-centred field along the sweep; vect is transformed in place. */
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)

typedef struct { double id, sqrtd, isqrtd, cf, cs, asq, a, bet2, bet3, alpha_f, alpha_s, s; } MhdEig;

static void mhd_eig(double gamma, const double *w, double bx, MhdEig *e) {
  e->id = 1.0/w[IDN];
  e->sqrtd = sqrt(w[IDN]);
  e->isqrtd = 1.0/e->sqrtd;
  double btsq = SQR(w[IBY]) + SQR(w[IBZ]);
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
  e->s = SIGN(bx);
}

void ao_char_left(int mhd, double gamma, const double *w, double bx, double *vect) {
  if (mhd) {
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

void ao_char_right(int mhd, double gamma, const double *w, double bx, double *vect) {
  if (mhd) {
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

/* ---------------------------------------------------------------- reconstruction */

/* src/reconstruct/plm.cpp:69-77,114-119 (uniform Cartesian branch), one variable of one cell */
void ao_plm_point(double qm1, double q, double qp1, double wp, double wm,
                  double *plus, double *minus) {
  double dwl = (q - qm1);
  double dwr = (qp1 - q);
  double dw2 = dwl*dwr;
  double dwm = 2.0*dw2/(dwl + dwr);
  if (dw2 <= 0.0) dwm = 0.0;
  *plus = q + wp*dwm;
  *minus = q - wm*dwm;
}

/* src/reconstruct/ppm.cpp:111-309 (uniform Cartesian branch), one variable of one cell;
 * uniform coefficients c1..c4=1/2, c5=1/6, c6=-1/6 (reconstruction.cpp:422-433) */
void ao_ppm_point(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                  double *plus, double *minus) {
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
  {
    double qa_tmp = dph - q_im1;
    double qb_tmp = q - dph;
    double qa2 = 3.0*(q_im1 + q - 2.0*dph);
    double qb2 = d2qc_im1;
    double qc2 = d2qc;
    double qd = 0.0;
    if (SIGN(qa2) == SIGN(qb2) && SIGN(qa2) == SIGN(qc2)) {
      qd = SIGN(qa2)*mn(C2*fabs(qb2), mn(C2*fabs(qc2), fabs(qa2)));
    }
    double dph_tmp = 0.5*(q_im1 + q) - qd/6.0;
    if (qa_tmp*qb_tmp < 0.0) dph = dph_tmp;
  }
  {
    double qa_tmp = dph_ip1 - q;
    double qb_tmp = q_ip1 - dph_ip1;
    double qa2 = 3.0*(q + q_ip1 - 2.0*dph_ip1);
    double qb2 = d2qc;
    double qc2 = d2qc_ip1;
    double qd = 0.0;
    if (SIGN(qa2) == SIGN(qb2) && SIGN(qa2) == SIGN(qc2)) {
      qd = SIGN(qa2)*mn(C2*fabs(qb2), mn(C2*fabs(qc2), fabs(qa2)));
    }
    double dphip1_tmp = 0.5*(q + q_ip1) - qd/6.0;
    if (qa_tmp*qb_tmp < 0.0) dph_ip1 = dphip1_tmp;
  }
  double d2qf = 6.0*(dph + dph_ip1 - 2.0*q);
  double qminus = dph;
  double qplus = dph_ip1;
  double dqf_minus = q - qminus;
  double dqf_plus = qplus - q;
  {
    double qa_tmp = dqf_minus*dqf_plus;
    double qb_tmp = (q_ip1 - q)*(q - q_im1);
    double qa2 = d2qc_im1;
    double qb2 = d2qc;
    double qc2 = d2qc_ip1;
    double qd = d2qf;
    double qe = 0.0;
    if (SIGN(qa2) == SIGN(qb2) && SIGN(qa2) == SIGN(qc2) && SIGN(qa2) == SIGN(qd)) {
      qe = SIGN(qd)*mn(mn(C2*fabs(qa2), C2*fabs(qb2)), mn(C2*fabs(qc2), fabs(qd)));
    }
    qa2 = mx(fabs(q_im1), fabs(q_im2));
    qb2 = mx(mx(fabs(q), fabs(q_ip1)), fabs(q_ip2));
    double rho = 0.0;
    if (fabs(qd) > (1.0e-12)*mx(qa2, qb2)) rho = qe/qd;
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
  }
  *plus = qplus;
  *minus = qminus;
}

/* limited slope of one variable: uniform van Leer (plm.cpp:69-77) or the nonuniform branches.
 * x1 (mode 1) multiplies by dx1f before dividing by dx1v (plm.cpp:85-86), x2 / x3 use the
 * pre-divided ratios (plm.cpp:198-204,308-313); x3 keeps the original VL expression
 * (plm.cpp:314-318), x1 / x2 the Mignone-corrected one (plm.cpp:94-96,207-209). */
static double plm_slope_g(double dwl, double dwr, const AoReconGeom *g) {
  double dwm;
  if (g->mode == 0) {
    double dw2 = dwl*dwr;
    dwm = 2.0*dw2/(dwl + dwr);
    if (dw2 <= 0.0) dwm = 0.0;
    return dwm;
  }
  double dqF, dqB;
  if (g->mode == 1) { dqF = dwr*g->dxf/g->dxv; dqB = dwl*g->dxf/g->dxvm; }
  else { dqF = dwr*g->dxF; dqB = dwl*g->dxB; }
  double dq2 = dqF*dqB;
  if (g->mode == 3) dwm = 2.0*dq2/(dqF + dqB);
  else dwm = (dq2*(g->cf*dqB + g->cb*dqF)/(SQR(dqB) + SQR(dqF) + dq2*(g->cf + g->cb - 2.0)));
  if (dq2 <= 0.0) dwm = 0.0;
  return dwm;
}

void ao_plm_point_g(double qm1, double q, double qp1, const AoReconGeom *g,
                    double *plus, double *minus) {
  double dwm = plm_slope_g(q - qm1, qp1 - q, g);
  *plus = q + g->wp*dwm;
  *minus = q - g->wm*dwm;
}

/* PPM, nonuniform Cartesian spacing: CW interface values with the per-cell weights, strict
 * monotonicity (Mignone eq 45) and the Mignone limiter with h ratios = 2
 * (ppm.cpp:111-129,196-207,282-300 and the x2 / x3 twins; reconstruction.cpp:412-419) */
void ao_ppm_point_g(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                    const AoReconGeom *g, double *plus, double *minus) {
  if (g->mode == 0) { ao_ppm_point(q_im2, q_im1, q, q_ip1, q_ip2, plus, minus); return; }
  double qa = (q - q_im1);
  double qb = (q_ip1 - q);
  double dd_im1 = g->c1m*qa + g->c2m*(q_im1 - q_im2);
  double dd     = g->c1*qb + g->c2*qa;
  double dd_ip1 = g->c1p*(q_ip2 - q_ip1) + g->c2p*qb;
  double dph = (g->c3*q_im1 + g->c4*q) + (g->c5*dd_im1 + g->c6*dd);
  double dph_ip1 = (g->c3p*q + g->c4p*q_ip1) + (g->c5p*dd + g->c6p*dd_ip1);
  dph     = mn(dph, mx(q, q_im1));
  dph_ip1 = mn(dph_ip1, mx(q, q_ip1));
  dph     = mx(dph, mn(q, q_im1));
  dph_ip1 = mx(dph_ip1, mn(q, q_ip1));
  double qminus = dph, qplus = dph_ip1;
  double dqf_minus = q - qminus;
  double dqf_plus = qplus - q;
  double e = dqf_minus*dqf_plus;
  if (e <= 0.0) {
    qminus = q;
    qplus = q;
  } else {
    if (fabs(dqf_minus) >= 2.0*fabs(dqf_plus)) qminus = q - 2.0*dqf_plus;
    if (fabs(dqf_plus) >= 2.0*fabs(dqf_minus)) qplus = q + 2.0*dqf_minus;
  }
  *plus = qplus;
  *minus = qminus;
}

void ao_plm(long n, int nvar, const double *qm1, const double *q, const double *qp1,
            double wp, double wm, double *ql_plus, double *qr_minus) {
  for (long i = 0; i < (long)nvar*n; ++i)
    ao_plm_point(qm1[i], q[i], qp1[i], wp, wm, &ql_plus[i], &qr_minus[i]);
}

void ao_ppm(long n, int nvar, const double *qm2, const double *qm1, const double *q,
            const double *qp1, const double *qp2, double dfloor, double pfloor,
            double *ql_plus, double *qr_minus) {
  for (long i = 0; i < (long)nvar*n; ++i)
    ao_ppm_point(qm2[i], qm1[i], q[i], qp1[i], qp2[i], &ql_plus[i], &qr_minus[i]);
  /* ApplyPrimitiveFloors on both states (ppm.cpp:326-332; eos/adiabatic_*.cpp) */
  for (long i = 0; i < n; ++i) {
    double *a[2] = {ql_plus, qr_minus};
    for (int s = 0; s < 2; ++s) {
      double *d = &a[s][IDN*n+i], *p = &a[s][IPR*n+i];
      *d = (*d > dfloor) ? *d : dfloor;
      *p = (*p > pfloor) ? *p : pfloor;
    }
  }
}

/* ---------------------------------------------------------------- characteristic reconstruction
 * One cell: stencil st[5][7] = cells -2..+2 in sweep order (rows 0 and 4 unused for order 2).
 * xorder = 2c: plm.cpp:62-66,107-130; xorder = 3c: ppm.cpp:66-75,311-332.  Eigenvectors of the
 * cell itself; floors re-applied to both face states. */
void ao_recon_char_point(int order, int mhd, double st[5][7], double bx, double gamma,
                         const AoReconGeom *g, double dfloor, double pfloor, double *pl,
                         double *mi) {
  int nw = mhd ? 7 : 5;
  const double *q = st[2];
  if (order == 2) {
    double dwl[7], dwr[7], dwm[7];
    for (int n = 0; n < nw; ++n) { dwl[n] = (q[n] - st[1][n]); dwr[n] = (st[3][n] - q[n]); }
    ao_char_left(mhd, gamma, q, bx, dwl);
    ao_char_left(mhd, gamma, q, bx, dwr);
    for (int n = 0; n < nw; ++n) dwm[n] = plm_slope_g(dwl[n], dwr[n], g);
    ao_char_right(mhd, gamma, q, bx, dwm);
    for (int n = 0; n < nw; ++n) { pl[n] = q[n] + g->wp*dwm[n]; mi[n] = q[n] - g->wm*dwm[n]; }
  } else {
    double c[5][7], w0[7];
    for (int n = 0; n < 7; ++n) w0[n] = q[n];
    for (int o = 0; o < 5; ++o) {
      for (int n = 0; n < 7; ++n) c[o][n] = st[o][n];
      ao_char_left(mhd, gamma, w0, bx, c[o]);
    }
    for (int n = 0; n < nw; ++n)
      ao_ppm_point_g(c[0][n], c[1][n], c[2][n], c[3][n], c[4][n], g, &pl[n], &mi[n]);
    ao_char_right(mhd, gamma, w0, bx, pl);
    ao_char_right(mhd, gamma, w0, bx, mi);
  }
  pl[IDN] = (pl[IDN] > dfloor) ? pl[IDN] : dfloor;
  mi[IDN] = (mi[IDN] > dfloor) ? mi[IDN] : dfloor;
  pl[IPR] = (pl[IPR] > pfloor) ? pl[IPR] : pfloor;
  mi[IPR] = (mi[IPR] > pfloor) ? mi[IPR] : pfloor;
}

/* batch version for tests: q[(o*7 + v)*n + i] */
void ao_recon_char(int order, int mhd, long n, const double *q, const double *bx, double gamma,
                   double wp, double wm, double dfloor, double pfloor, double *plus,
                   double *minus) {
  int nw = mhd ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double st[5][7], pl[7], mi[7];
    for (int o = 0; o < 5; ++o) for (int v = 0; v < 7; ++v) st[o][v] = q[(o*7 + v)*n + i];
    AoReconGeom g;
    memset(&g, 0, sizeof(g));
    g.wp = wp; g.wm = wm;
    ao_recon_char_point(order, mhd, st, mhd ? bx[i] : 0.0, gamma, &g, dfloor, pfloor, pl, mi);
    for (int v = 0; v < nw; ++v) { plus[v*n + i] = pl[v]; minus[v*n + i] = mi[v]; }
  }
}
