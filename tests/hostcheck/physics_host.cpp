// TEST-ONLY: compiles the product's device physics header (ab_physics.cuh) for the host so the
// CPU test-suite can compare its arithmetic with the oracle without a GPU.  Not part of the
// product; the product library contains no host execution path for these functions.
#include "../../athena-gamma_b200/csrc/ab_physics.cuh"
#include "../../athena-gamma_b200/csrc/ab_types.h"

template <int S, bool M>
static void run(long n, const double *wl, const double *wr, const double *bx, const double *dvn,
                const double *dvt, double gamma, double dt, double dx, double *flx,
                double *wct) {
  const int nw = M ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double a[7], b[7], f[7];
    for (int v = 0; v < nw; ++v) { a[v] = wl[v*n+i]; b[v] = wr[v*n+i]; }
    ab::riemann<S, M>(a, b, M ? bx[i] : 0.0, gamma, dvn ? dvn[i] : 0.0, dvt ? dvt[i] : 0.0, f);
    for (int v = 0; v < nw; ++v) flx[v*n+i] = f[v];
    if (M && wct) wct[i] = ab::weight_for_ct(f[0], a[0], b[0], dx, dt);
  }
}

extern "C" {
void hc_riemann(int solver, int mhd, long n, const double *wl, const double *wr,
                const double *bx, const double *dvn, const double *dvt, double gamma, double dt,
                double dx, double *flx, double *wct) {
#define RUN(S, M) run<S, M>(n, wl, wr, bx, dvn, dvt, gamma, dt, dx, flx, wct)
  if (solver == 6) { if (mhd) RUN(6, true); else RUN(6, false); return; }
  if (!mhd) {
    if (solver == 1) RUN(1, false);
    else if (solver == 0) RUN(0, false);
    else if (solver == 4) RUN(4, false);
    else RUN(3, false);
  } else {
    if (solver == 2) RUN(2, true);
    else if (solver == 0) RUN(0, true);
    else if (solver == 5) RUN(5, true);
    else RUN(3, true);
  }
#undef RUN
}
// isothermal solvers: solver 0 = hlle (hydro / MHD), 2 = hlld (MHD), 3 = roe, 6 = llf
void hc_riemann_iso(int solver, int mhd, long n, const double *wl, const double *wr,
                    const double *bx, double iso_cs, double dfloor, double *flx) {
  const int nw = mhd ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double a[7], b[7], f[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int v = 0; v < nw; ++v) { a[v] = wl[v*n+i]; b[v] = wr[v*n+i]; }
    if (solver == 6 && mhd) ab::riemann<ab::SOLVER_LLF_ISO, true>(a, b, bx[i], iso_cs, 0.0, 0.0, f, dfloor);
    else if (solver == 6) ab::riemann<ab::SOLVER_LLF_ISO, false>(a, b, 0.0, iso_cs, 0.0, 0.0, f, dfloor);
    else if (solver == 3 && mhd) ab::riemann<ab::SOLVER_ROE_ISO, true>(a, b, bx[i], iso_cs, 0.0, 0.0, f, dfloor);
    else if (solver == 3) ab::riemann<ab::SOLVER_ROE_ISO, false>(a, b, 0.0, iso_cs, 0.0, 0.0, f, dfloor);
    else if (!mhd) ab::riemann<ab::SOLVER_HLLE_ISO, false>(a, b, 0.0, iso_cs, 0.0, 0.0, f, dfloor);
    else if (solver == 2) ab::riemann<ab::SOLVER_HLLD_ISO, true>(a, b, bx[i], iso_cs, 0.0, 0.0, f, dfloor);
    else ab::riemann<ab::SOLVER_HLLE_ISO, true>(a, b, bx[i], iso_cs, 0.0, 0.0, f, dfloor);
    for (int v = 0; v < nw; ++v) flx[v*n+i] = f[v];
  }
}
// characteristic reconstruction of one cell per column: q[5][7][n] (stencil -2..+2, sweep order),
// bx[n]; order 2 -> plm_char, 3 -> ppm_char; out plus/minus [7][n]
void hc_recon_char(int order, int mhd, long n, const double *q, const double *bx, double gamma,
                   double wp, double wm, double dfloor, double pfloor, double *plus,
                   double *minus) {
  const int nw = mhd ? 7 : 5;
  for (long i = 0; i < n; ++i) {
    double st[5][7], pl[7], mi[7];
    for (int o = 0; o < 5; ++o) for (int v = 0; v < 7; ++v) st[o][v] = q[(o*7 + v)*n + i];
    const double b = mhd ? bx[i] : 0.0;
    if (order == 2) {
      if (mhd) ab::plm_char<true>(st[1], st[2], st[3], b, gamma, wp, wm, dfloor, pfloor, pl, mi);
      else ab::plm_char<false>(st[1], st[2], st[3], b, gamma, wp, wm, dfloor, pfloor, pl, mi);
    } else {
      if (mhd) ab::ppm_char<true>(st[0], st[1], st[2], st[3], st[4], b, gamma, dfloor, pfloor, pl, mi);
      else ab::ppm_char<false>(st[0], st[1], st[2], st[3], st[4], b, gamma, dfloor, pfloor, pl, mi);
    }
    for (int v = 0; v < nw; ++v) { plus[v*n + i] = pl[v]; minus[v*n + i] = mi[v]; }
  }
}
void hc_plm(long n, int nvar, const double *qm1, const double *q, const double *qp1,
            double wp, double wm, double *ql, double *qr) {
  for (long i = 0; i < (long)nvar*n; ++i) ab::plm(qm1[i], q[i], qp1[i], wp, wm, ql[i], qr[i]);
}
void hc_ppm(long n, int nvar, const double *qm2, const double *qm1, const double *q,
            const double *qp1, const double *qp2, double *ql, double *qr) {
  for (long i = 0; i < (long)nvar*n; ++i)
    ab::ppm(qm2[i], qm1[i], q[i], qp1[i], qp2[i], ql[i], qr[i]);
}
}  // extern "C"

// nonuniform spacing: reconstruct cells lo..hi of a 1-D line with the product's geometry
// table (ab_plan_geometry); mode = 1 + direction, 0 = uniform.  q[v*nc + i].
template <int MODE>
static void line_t(int order, int nc, const double *wp, const double *wm, const double *tab,
                   int nvar, const double *q, int lo, int hi, double *plus, double *minus) {
  for (int v = 0; v < nvar; ++v) for (int i = lo; i <= hi; ++i) {
    const double *c = q + (long)v*nc + i;
    const double *t = tab ? tab + (long)i*ab::NUG : nullptr;
    double &pl = plus[(long)v*nc + i], &mi = minus[(long)v*nc + i];
    if (order == 2) {
      if (MODE) ab::plm_nu<MODE>(c[-1], c[0], c[1], wp[i], wm[i], t, pl, mi);
      else ab::plm(c[-1], c[0], c[1], wp[i], wm[i], pl, mi);
    } else {
      if (MODE) ab::ppm_nu(c[-2], c[-1], c[0], c[1], c[2], t, pl, mi);
      else ab::ppm(c[-2], c[-1], c[0], c[1], c[2], pl, mi);
    }
  }
}
extern "C" void hc_recon_line(int mode, int order, int nc, const double *wp, const double *wm,
                   const double *tab, int nvar, const double *q, int lo, int hi, double *plus,
                   double *minus) {
  if (mode == 0) line_t<0>(order, nc, wp, wm, tab, nvar, q, lo, hi, plus, minus);
  else if (mode == 1) line_t<1>(order, nc, wp, wm, tab, nvar, q, lo, hi, plus, minus);
  else if (mode == 2) line_t<2>(order, nc, wp, wm, tab, nvar, q, lo, hi, plus, minus);
  else line_t<3>(order, nc, wp, wm, tab, nvar, q, lo, hi, plus, minus);
}
// characteristic variant: q[v*nc + i] with 7 sweep-ordered slots, bx[i]
template <bool MHD, int MODE>
static void line_char_t(int order, int nc, const double *wp, const double *wm, const double *tab,
                        const double *q, const double *bx, double gamma, double dfloor,
                        double pfloor, int lo, int hi, double *plus, double *minus) {
  const int nw = MHD ? 7 : 5;
  for (int i = lo; i <= hi; ++i) {
    double st[5][7], pl[7], mi[7];
    for (int o = -2; o <= 2; ++o) for (int v = 0; v < 7; ++v) st[o+2][v] = q[(long)v*nc + i + o];
    const double b = MHD ? bx[i] : 0.0;
    const double *t = tab ? tab + (long)i*ab::NUG : nullptr;
    if (order == 2)
      ab::plm_char<MHD,MODE>(st[1], st[2], st[3], b, gamma, wp[i], wm[i], dfloor, pfloor, pl, mi, t);
    else
      ab::ppm_char<MHD,(MODE != 0)>(st[0], st[1], st[2], st[3], st[4], b, gamma, dfloor, pfloor,
                                    pl, mi, t);
    for (int v = 0; v < nw; ++v) { plus[(long)v*nc + i] = pl[v]; minus[(long)v*nc + i] = mi[v]; }
  }
}
extern "C" void hc_recon_line_char(int mode, int order, int mhd, int nc, const double *wp, const double *wm,
                        const double *tab, const double *q, const double *bx, double gamma,
                        double dfloor, double pfloor, int lo, int hi, double *plus,
                        double *minus) {
#define LC(M, D) line_char_t<M,D>(order, nc, wp, wm, tab, q, bx, gamma, dfloor, pfloor, lo, hi, plus, minus)
  if (mhd) { if (mode == 0) LC(true,0); else if (mode == 1) LC(true,1); else if (mode == 2) LC(true,2); else LC(true,3); }
  else { if (mode == 0) LC(false,0); else if (mode == 1) LC(false,1); else if (mode == 2) LC(false,2); else LC(false,3); }
#undef LC
}

// static mesh refinement: restriction of a fine box / prolongation of a coarse box on plain arrays
// laid out like the MeshBlock's (fine nc3 x nc2 x nc1, coarse cnc3 x cnc2 x cnc1), with the
// product's point functions.  dims: {nc1,nc2,nc3, cnc1,cnc2,cnc3, is,js,ks, cis,cjs,cks, ndim}
extern "C" void hc_smr_restrict(const int *dims, const double *dx1, const double *dx2,
                                const double *dx3, const double *fine, double *coarse,
                                const int *box) {
  const int nc1 = dims[0], nc2 = dims[1], c1 = dims[3], c2 = dims[4];
  const int is = dims[6], js = dims[7], ks = dims[8], cis = dims[9], cjs = dims[10], cks = dims[11];
  const int nd = dims[12];
  for (int ck = box[4]; ck <= box[5]; ++ck) for (int cj = box[2]; cj <= box[3]; ++cj)
    for (int ci = box[0]; ci <= box[1]; ++ci) {
      const int i = (ci - cis)*2 + is, j = nd > 1 ? (cj - cjs)*2 + js : 0,
                k = nd > 2 ? (ck - cks)*2 + ks : 0;
      double f[8] = {0}, v[8] = {0};
      for (int dk = 0; dk < (nd > 2 ? 2 : 1); ++dk) for (int dj = 0; dj < (nd > 1 ? 2 : 1); ++dj)
        for (int di = 0; di < 2; ++di) {
          const int q = dk*4 + di*2 + dj;
          f[q] = fine[((long)(k+dk)*nc2 + (j+dj))*nc1 + (i+di)];
          v[q] = dx1[i+di]*dx2[j+dj]*dx3[k+dk];
        }
      coarse[((long)ck*c2 + cj)*c1 + ci] = ab::restrict_cc(nd, f, v);
    }
}
extern "C" void hc_smr_prolong(const int *dims, const double *x1v, const double *x2v,
                               const double *x3v, const double *cx1v, const double *cx2v,
                               const double *cx3v, const double *coarse, double *fine,
                               const int *box) {
  const int nc1 = dims[0], nc2 = dims[1], c1 = dims[3], c2 = dims[4];
  const int is = dims[6], js = dims[7], ks = dims[8], cis = dims[9], cjs = dims[10], cks = dims[11];
  const int nd = dims[12];
  for (int k = box[4]; k <= box[5]; ++k) for (int j = box[2]; j <= box[3]; ++j)
    for (int i = box[0]; i <= box[1]; ++i) {
      const int fi = (i - cis)*2 + is, fj = nd > 1 ? (j - cjs)*2 + js : 0,
                fk = nd > 2 ? (k - cks)*2 + ks : 0;
      auto C = [&](int kk, int jj, int ii) { return coarse[((long)kk*c2 + jj)*c1 + ii]; };
      const double cc = C(k, j, i);
      double g1, g2 = 0, g3 = 0, d2m = 0, d2p = 0, d3m = 0, d3p = 0;
      g1 = ab::prolong_grad(C(k,j,i-1), cc, C(k,j,i+1), cx1v[i] - cx1v[i-1], cx1v[i+1] - cx1v[i]);
      const double d1m = cx1v[i] - x1v[fi], d1p = x1v[fi+1] - cx1v[i];
      if (nd > 1) {
        g2 = ab::prolong_grad(C(k,j-1,i), cc, C(k,j+1,i), cx2v[j] - cx2v[j-1], cx2v[j+1] - cx2v[j]);
        d2m = cx2v[j] - x2v[fj]; d2p = x2v[fj+1] - cx2v[j];
      }
      if (nd > 2) {
        g3 = ab::prolong_grad(C(k-1,j,i), cc, C(k+1,j,i), cx3v[k] - cx3v[k-1], cx3v[k+1] - cx3v[k]);
        d3m = cx3v[k] - x3v[fk]; d3p = x3v[fk+1] - cx3v[k];
      }
      double out[8];
      ab::prolong_cc(nd, cc, g1, g2, g3, d1m, d1p, d2m, d2p, d3m, d3p, out);
      for (int dk = 0; dk < (nd > 2 ? 2 : 1); ++dk) for (int dj = 0; dj < (nd > 1 ? 2 : 1); ++dj)
        for (int di = 0; di < 2; ++di)
          fine[((long)(fk+dk)*nc2 + (fj+dj))*nc1 + (fi+di)] = out[dk*4 + di*2 + dj];
    }
}

// make_fastdiv (ab_types.h) with the device's fast_div formula (ab_batch.cuh: __umulhi(t, m) >> s,
// m = 0 meaning d = 1): number of t in [t0, t1] (stride `step`) where it differs from t / d
extern "C" long hc_fastdiv_mismatches(int d, long t0, long t1, long step) {
  const ab::FastDiv f = ab::make_fastdiv(d);
  long bad = 0;
  for (long t = t0; t <= t1; t += step) {
    const unsigned hi = (unsigned)(((unsigned long long)(unsigned)t*f.m) >> 32);
    const int q = f.m ? (int)(hi >> f.s) : (int)t;
    if (q != (int)(t / d)) ++bad;
  }
  return bad;
}
