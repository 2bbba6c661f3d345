/* oracle_internal.h -- shared declarations of the CPU oracle (TEST INFRASTRUCTURE ONLY). */
#ifndef ORACLE_INTERNAL_H_
#define ORACLE_INTERNAL_H_
#include "athena_oracle.h"

/* variable indices: src/athena.hpp:136-144 */
enum { IDN = 0, IM1 = 1, IM2 = 2, IM3 = 3, IEN = 4 };
enum { IVX = 1, IVY = 2, IVZ = 3, IPR = 4, IBY = 5, IBZ = 6 };
enum { IB1 = 0, IB2 = 1, IB3 = 2 };
#define NHYDRO 5   /* adiabatic; oracle_mesh.c redefines it as the mesh's run-time count */

double ao_sound_speed(double gamma, const double *prim);
double ao_fast_speed(double gamma, const double *prim, double bx);
double ao_weight_for_ct(double dflx, double rhol, double rhor, double dx, double dt);
void ao_riemann_point(int solver, int mhd, const double *wli, const double *wri,
                      double bxi, double gamma, double dvn, double dvt, double *flxi);
double ao_fast_speed_iso(double cs, const double *prim, double bx);
void ao_riemann_point_iso(int solver, int mhd, const double *wli, const double *wri,
                          double bxi, double iso_cs, double dfloor, double *flxi);
void ao_char_left(int mhd, double gamma, const double *w, double bx, double *vect);
void ao_char_right(int mhd, double gamma, const double *w, double bx, double *vect);
/* reconstruction geometry of one cell along one direction.  mode 0: uniform spacing (only
 * wp, wm are used); 1 / 2 / 3: the nonuniform branches of the x1 / x2 / x3 routines, which
 * differ in rounding order and (x3) in the limiter (plm.cpp:81-105,194-214,306-322);
 * c*: PPM weights of cells i-1 (m), i, i+1 (p) (reconstruction.cpp:434-461,559-582,609-631) */
typedef struct {
  double wp, wm;
  int mode;
  double dxf, dxv, dxvm, cf, cb, dxF, dxB;
  double c1m, c2m, c1, c2, c1p, c2p, c3, c4, c5, c6, c3p, c4p, c5p, c6p;
} AoReconGeom;
void ao_recon_char_point(int order, int mhd, double st[5][7], double bx, double gamma,
                         const AoReconGeom *g, double dfloor, double pfloor, double *pl,
                         double *mi);
void ao_plm_point_g(double qm1, double q, double qp1, const AoReconGeom *g,
                    double *plus, double *minus);
void ao_ppm_point_g(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                    const AoReconGeom *g, double *plus, double *minus);
void ao_plm_point(double qm1, double q, double qp1, double wp, double wm,
                  double *plus, double *minus);
void ao_ppm_point(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                  double *plus, double *minus);
#endif
