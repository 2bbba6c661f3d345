/* athena_oracle.h -- CPU restatement (plain C) of the reference's per-MeshBlock hydro/MHD update.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg.  The product (athena-gamma_b200/) never includes or links it.
 *
 * Covers one-level meshes (hydro / MHD, all solvers, scalars, both EOS, uniform and geometric
 * spacing) and statically refined hydro meshes (oracle_smr.c).
 *
 * Pinned bit-for-bit against the unmodified reference (oracle/_ref, built by
 * oracle/build_ref.py) by tests/test_oracle_vs_reference.py and the committed fixtures in
 * tests/golden/.
 */
#ifndef ATHENA_ORACLE_H_
#define ATHENA_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { AO_BC_PERIODIC = 0, AO_BC_OUTFLOW = 1, AO_BC_REFLECT = 2, AO_BC_USER = 3 };
enum { AO_SOLVER_HLLE = 0, AO_SOLVER_HLLC = 1, AO_SOLVER_HLLD = 2, AO_SOLVER_ROE = 3,
       AO_SOLVER_LHLLC = 4, AO_SOLVER_LHLLD = 5, AO_SOLVER_LLF = 6 };
enum { AO_INT_VL2 = 0, AO_INT_RK2 = 1, AO_INT_RK1 = 2, AO_INT_RK3 = 3 };

typedef struct {
  int nx1, nx2, nx3;    /* <mesh> size */
  int bx1, bx2, bx3;    /* <meshblock> size */
  double x1min, x1max, x2min, x2max, x3min, x3max;
  int bc[6];            /* ix1,ox1,ix2,ox2,ix3,ox3 */
  int ng;               /* NGHOST */
  int mhd;              /* MAGNETIC_FIELDS_ENABLED */
  int solver;           /* AO_SOLVER_* */
  int xorder;           /* 1,2,3 */
  int integrator;       /* AO_INT_* */
  double gamma, dfloor, pfloor, cfl, tlim, start_time;
  int nscalars;         /* NSCALARS (configure.py --nscalars) */
  int eos;              /* 0 adiabatic ; 1 isothermal (configure.py --eos) */
  double sfloor;        /* hydro/sfloor (eos ctor default sqrt(1024*FLT_MIN)) */
  double iso_cs;        /* hydro/iso_sound_speed */
  double grav_acc[3];   /* hydro/grav_acc1..3 (hydro/srcterms/hydro_srcterms.cpp:68-75) */
  int char_proj;        /* time/xorder = 2c / 3c: reconstruct characteristic variables */
  double xrat[3];       /* mesh/x1rat..x3rat (0 or 1 = uniform): geometric cell-size ratio */
  /* mesh/refinement = static: <refinementN> blocks (src/mesh/mesh.cpp:323-465); hydro only */
  int nref;
  double ref[8][6];     /* x1min,x1max,x2min,x2max,x3min,x3max of each region */
  int ref_level[8];     /* its level (root = 0) */
} AoParams;

typedef struct AoMesh AoMesh;

AoMesh *ao_create(const AoParams *p);
void ao_destroy(AoMesh *m);
int ao_nblocks(const AoMesh *m);
int ao_block_level(const AoMesh *m, int b);   /* LogicalLocation::level (0 on a one-level mesh) */
/* out[0..2]=lx1..3, out[3..5]=nc1..3, out[6..11]=is,ie,js,je,ks,ke */
void ao_block_info(const AoMesh *m, int b, long *out);
/* array access by name: u u1 w bcc b1 b2 b3 b1_1 b1_2 b1_3 flux1 flux2 flux3 e1 e2 e3
 * wght1 wght2 wght3 e2_x1f e3_x1f e1_x2f e3_x2f e1_x3f e2_x3f x1f x2f x3f x1v x2v x3v
 * dx1f dx2f dx3f ; passive scalars: s s1 r sflux1 sflux2 sflux3 ;
 * returns pointer, *n = number of doubles */
double *ao_array(AoMesh *m, int b, const char *name, long *n);

/* Mesh::Initialize after the problem generator filled u (and b): ghost exchange,
 * ConservedToPrimitive, physical boundaries, NewBlockTimeStep + NewTimeStep */
void ao_initialize(AoMesh *m);
/* one full cycle (all stages) + time/dt update; returns the dt that was used */
double ao_cycle(AoMesh *m);
double ao_time(const AoMesh *m);
double ao_dt(const AoMesh *m);
void ao_set_time_dt(AoMesh *m, double time, double dt);
int ao_ncycle(const AoMesh *m);

/* task-level entry points (same granularity as the reference's task bodies) */
void ao_calc_fluxes(AoMesh *m, int b, int order);
void ao_corner_e(AoMesh *m, int b);
void ao_emf_exchange(AoMesh *m);           /* SendFluxCorrection + ReceiveFluxCorrection */
void ao_weighted_ave_cc(AoMesh *m, int b, int out_reg, int in1_reg, const double w[5]);
void ao_weighted_ave_fc(AoMesh *m, int b, int out_reg, int in1_reg, const double w[5]);
void ao_swap_cc(AoMesh *m, int b);          /* u <-> u1 */
void ao_swap_fc(AoMesh *m, int b);          /* b <-> b1 */
void ao_zero_reg1(AoMesh *m, int b);        /* u1.ZeroClear, b1.ZeroClear */
void ao_add_flux_div(AoMesh *m, int b, double wght);
void ao_ct(AoMesh *m, int b, double wght);
void ao_exchange_cc(AoMesh *m);             /* Send/Receive/SetBoundaries for u */
void ao_exchange_fc(AoMesh *m);             /* same for b */
void ao_cons2prim(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku);
void ao_prim2cons(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku);
void ao_primitives(AoMesh *m, int b);       /* Primitives task: range from neighbours */
void ao_physical_bcs(AoMesh *m, int b);
double ao_new_block_dt(AoMesh *m, int b);
/* Mesh::EnrollUserBoundaryFunction (src/mesh/mesh.cpp): BValFunc of athena.hpp:179-182 with
 * plain arrays -- prim = w(NHYDRO,k,j,i), bf = {b.x1f, b.x2f, b.x3f} (NULL without MHD) */
typedef void (*AoBValFunc)(void *user, int block, double *prim, double *b1f, double *b2f,
                           double *b3f, double time, double dt, int il, int iu, int jl, int ju,
                           int kl, int ku, int ngh);
void ao_enroll_user_bc(AoMesh *m, int face, AoBValFunc fn, void *user);
/* Mesh::EnrollUserExplicitSourceFunction: SrcTermFunc (src/athena.hpp:185-189) on plain arrays;
 * called last in AddSourceTerms with time = start-of-stage time, dt = beta*dt */
typedef void (*AoSrcTermFunc)(void *user, int block, double time, double dt, const double *prim,
                              const double *prim_scalar, const double *bcc, double *cons,
                              double *cons_scalar);
void ao_enroll_user_source(AoMesh *m, AoSrcTermFunc fn, void *user);
/* HydroSourceTerms::AddSourceTerms: constant acceleration (hydro/srcterms/constant_acc.cpp),
 * then the user-enrolled source function */
void ao_add_source_terms(AoMesh *m, int b, double time, double dt);
/* HistoryOutput::WriteOutputFile sums (src/outputs/history.cpp:69-169): mass, 1..3-mom,
 * 1..3-KE, tot-E, [1..3-ME], [scalars]; returns the number of values written to out */
int ao_history(AoMesh *m, double *out);
/* passive scalars (src/scalars, src/eos/eos_scalars.cpp) */
void ao_calc_scalar_fluxes(AoMesh *m, int b, int order);   /* PassiveScalars::CalculateFluxes */
void ao_integrate_scalars(AoMesh *m, int b, int stage);    /* IntegrateScalars task */
void ao_exchange_scalars(AoMesh *m);                       /* Send/Receive/SetBoundaries for s */
void ao_scalar_cons2prim(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku);
void ao_scalar_prim2cons(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku);

/* point-wise kernels on n independent interfaces / cells (structure-of-arrays, stride n) */
void ao_riemann(int solver, int mhd, long n, const double *wl, const double *wr,
                const double *bx, double gamma, double dt, double dx,
                double *flx, double *wct);
/* same with the LHLLC/LHLLD shock-detector inputs (NULL = 0) */
void ao_riemann_dv(int solver, int mhd, long n, const double *wl, const double *wr,
                   const double *bx, const double *dvn, const double *dvt, double gamma,
                   double dt, double dx, double *flx, double *wct);
/* isothermal EOS: hlle (hydro), hlle / hlld (MHD); slot 4 of the 5/7-slot vectors is unused */
void ao_riemann_iso(int solver, int mhd, long n, const double *wl, const double *wr,
                    const double *bx, double iso_cs, double dfloor, double *flx);
/* characteristic reconstruction (xorder 2c / 3c) of n independent cells, sweep order:
 * q[(o*7 + v)*n + i] = variable v of stencil cell o-2; plus/minus [v*n + i] */
void ao_recon_char(int order, int mhd, long n, const double *q, const double *bx, double gamma,
                   double wp, double wm, double dfloor, double pfloor, double *plus,
                   double *minus);
void ao_plm(long n, int nvar, const double *qm1, const double *q, const double *qp1,
            double wp, double wm, double *ql_plus, double *qr_minus);
void ao_ppm(long n, int nvar, const double *qm2, const double *qm1, const double *q,
            const double *qp1, const double *qp2, double dfloor, double pfloor,
            double *ql_plus, double *qr_minus);

/* test hook (refined meshes): transfer list of one exchange + ProlongateBoundaries + flux
 * correction, 12 longs per row (layout: oracle_smr.c); returns the number of rows */
long ao_smr_transfers(AoMesh *m, long *rows, long max_rows);
/* test hooks (refined meshes): restrict u -> coarse_u / prolongate coarse_w -> w over the coarse
 * box {si,ei,sj,ej,sk,ek} of block b; arrays by name: coarse_u coarse_w cx1v cx2v cx3v */
void ao_smr_restrict_box(AoMesh *m, int b, const int *box);
void ao_smr_prolong_box(AoMesh *m, int b, const int *box);
void ao_smr_step(AoMesh *m, int what);   /* 0 exchange, 1 ProlongateBoundaries, 2 flux correction */
/* neighbour list of block b: rows of 8 ints {ox1, ox2, ox3, type, gid, level, fi1, fi2};
 * nblevel (27 ints, [k][j][i]) when not NULL; returns the number of neighbours */
int ao_neighbors(const AoMesh *m, int b, int *rows, int *nblevel);

/* test entries for the reconstruction geometry of one direction (nonuni: x?rat != 1) */
void ao_recon_line(int dir, int nonuni, int order, int nc, int s, int e, int ng, const double *xf,
                   const double *xv, const double *dxf, int nvar, const double *q, int lo, int hi,
                   double *plus, double *minus);
void ao_recon_line_char(int dir, int nonuni, int order, int mhd, int nc, int s, int e, int ng,
                        const double *xf, const double *xv, const double *dxf, const double *q,
                        const double *bx, double gamma, double dfloor, double pfloor, int lo,
                        int hi, double *plus, double *minus);
void ao_bcc_weights(int dir, int nonuni, int nc, int s, int e, int ng, const double *xf,
                    const double *xv, const double *dxf, double *lw, double *rw);

#ifdef __cplusplus
}
#endif
#endif
