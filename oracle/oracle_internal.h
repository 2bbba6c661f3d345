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
void ao_recon_char_point(int order, int mhd, double st[5][7], double bx, double gamma,
                         double wp, double wm, double dfloor, double pfloor, double *pl,
                         double *mi);
void ao_plm_point(double qm1, double q, double qp1, double wp, double wm,
                  double *plus, double *minus);
void ao_ppm_point(double q_im2, double q_im1, double q, double q_ip1, double q_ip2,
                  double *plus, double *minus);
#endif
