/* oracle_mesh.c -- MeshBlocks, ghost exchange, EMF correction, integrator glue of the
 * reference path, restated in plain C for uniform Cartesian meshes (periodic / outflow).
 *
 * TEST INFRASTRUCTURE ONLY (see athena_oracle.h).  Each function cites the reference
 * file:line it restates.  Loops are deliberately simple (one cell / face / edge at a time).
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "oracle_internal.h"

static inline double mn(double a, double b) { return (b < a) ? b : a; }
/* NHYDRO is 5 (adiabatic) or 4 (isothermal, configure.py:374-377): every function below that
 * uses it has the mesh `m` in scope */
#undef NHYDRO
#define NHYDRO (m->nh)
#define ISO(m) ((m)->p.eos == 1)
#define SQR(x) ((x)*(x))

typedef struct {
  int ox1, ox2, ox3, type; /* type: 0 face 1 edge 2 corner */
  int gid, bufid, targetid, fid, eid;
  int level, fi1, fi2;     /* multilevel meshes: neighbour's level, finer-leaf indices */
} Nb;
typedef struct TNode TNode;

typedef struct AoBlock {
  int gid;
  long lx1, lx2, lx3;
  int nc1, nc2, nc3, is, ie, js, je, ks, ke;
  double bx1min, bx1max, bx2min, bx2max, bx3min, bx3max;
  int bcs[6];            /* -1: neighbour block (block/periodic), else AO_BC_* physical */
  int nblevel[3][3][3];
  int nnb; Nb nb[56];      /* 26 on one level; up to 56 with finer neighbours */
  /* static mesh refinement (oracle_smr.c): level, the MeshRefinement's coarse buffers */
  int level, cng, cis, cie, cjs, cje, cks, cke, cnc1, cnc2, cnc3;
  double *cx1f, *cx2f, *cx3f, *cx1v, *cx2v, *cx3v, *coarse_u, *coarse_w, *coarse_s,
         *coarse_r;
  int nedge_fine[12];
  double *x1f, *x2f, *x3f, *x1v, *x2v, *x3v, *dx1f, *dx2f, *dx3f;
  AoReconGeom *rg[3];    /* per-index reconstruction geometry along x1, x2, x3 */
  double *bw[3][2];      /* CalculateCellCenteredField weights (lw, rw) per index */
  double *u, *u1, *w, *bcc, *flux[3];
  double *b[3], *b1[3], *e[3], *wght[3];
  double *e2_x1f, *e3_x1f, *e1_x2f, *e3_x2f, *e1_x3f, *e2_x3f, *cc_e;
  double *s, *s1, *r, *sflux[3];   /* PassiveScalars::s, s1, r, s_flux (scalars.hpp:40-56) */
  double *recv[26]; long recvn[26];
  double new_dt;
} AoBlock;

struct AoMesh {
  AoParams p;
  int nh;                /* NHYDRO */
  int ndim, f2, f3;
  int nrbx1, nrbx2, nrbx3, nb;
  AoBlock *blk;
  int *gid_of;           /* [lx3][lx2][lx1] -> gid */
  double time, dt;
  int ncycle;
  int nstages;
  double beta[4], delta[4], g1[4], g2[4], g3[4], ebeta[4];
  double cfl;
  double xrat[3];        /* mesh/x?rat with 0 read as 1 (uniform) */
  int multilevel, root_level; TNode *root;
  AoBValFunc user_bc[6]; void *user_bc_arg[6];
  AoSrcTermFunc user_src; void *user_src_arg;
  double sbeta[4];
  double bc_time, bc_dt;   /* (time, dt) handed to boundary functions: end of stage, beta*dt */
  /* canonical buffer-id table: src/bvals/bvals_base.cpp:153-256 */
  int nni; int ni[26][3];
};

/* ---- indexing (src/athena_arrays.hpp:140-143: last index fastest) ---- */
#define CC(B,n,k,j,i) ((((long)(n)*(B)->nc3 + (k))*(B)->nc2 + (j))*(B)->nc1 + (i))
#define F1(B,k,j,i) (((long)(k)*(B)->nc2 + (j))*((B)->nc1+1) + (i))
#define F2(B,k,j,i) (((long)(k)*((B)->nc2+1) + (j))*(B)->nc1 + (i))
#define F3(B,k,j,i) (((long)(k)*(B)->nc2 + (j))*(B)->nc1 + (i))
#define FL1(B,n,k,j,i) ((((long)(n)*(B)->nc3 + (k))*(B)->nc2 + (j))*((B)->nc1+1) + (i))
#define FL2(B,n,k,j,i) ((((long)(n)*(B)->nc3 + (k))*((B)->nc2+1) + (j))*(B)->nc1 + (i))
#define FL3(B,n,k,j,i) ((((long)(n)*((B)->nc3+1) + (k))*(B)->nc2 + (j))*(B)->nc1 + (i))
/* EdgeField (src/athena.hpp:107-115) */
#define E1(B,k,j,i) (((long)(k)*((B)->nc2+1) + (j))*(B)->nc1 + (i))
#define E2(B,k,j,i) (((long)(k)*(B)->nc2 + (j))*((B)->nc1+1) + (i))
#define E3(B,k,j,i) (((long)(k)*((B)->nc2+1) + (j))*((B)->nc1+1) + (i))

static double *dalloc(long n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

/* src/mesh/mesh.hpp:389-405 */
static double mesh_gen_x(long index, long nrange) {
  long noffset = index - (nrange)/2;
  long noffset_ceil = index - (nrange+1)/2;
  return (double)(noffset + noffset_ceil)/(2.0*nrange);
}
/* src/mesh/mesh.hpp:467-488 */
static double uniform_gen(double x, double xmin, double xmax) {
  return 0.5*(xmin + xmax) + (x*xmax - x*xmin);
}

/* DefaultMeshGeneratorX? (src/mesh/mesh.hpp:411-455): geometric spacing, x in [0,1] */
static double default_gen(double x, double xmin, double xmax, double rat, int nx) {
  double lw, rw;
  if (rat == 1.0) {
    rw = x; lw = 1.0 - x;
  } else {
    double ratn = pow(rat, nx);
    double rnx = pow(rat, x*nx);
    lw = (rnx - ratn)/(1.0 - ratn);
    rw = 1.0 - lw;
  }
  return xmin*lw + xmax*rw;
}

/* Coordinates ctor (src/coordinates/coordinates.cpp:92-160 and twins): uniform branch, or the
 * mesh-generator branch when x?rat != 1 (Mesh::use_uniform_meshgen_fn_, mesh.cpp:278-289);
 * Cartesian x?v (src/coordinates/cartesian.cpp:25-75) */
static void make_coords(int nx_mesh, int bx, int ng, long lx, double mmin, double mmax,
                        double bmin, double bmax, int nc, int refl_in, int refl_out, double rat,
                        double **xf, double **xv, double **dxf) {
  *xf = dalloc(nc + 1); *xv = dalloc(nc); *dxf = dalloc(nc);
  if (nc == 1) {
    (*dxf)[0] = bmax - bmin;
    (*xf)[0] = bmin;
    (*xf)[1] = bmax;
    (*xv)[0] = 0.5*((*xf)[1] + (*xf)[0]);
    return;
  }
  int il = ng, iu = ng + bx - 1;
  if (rat != 1.0) {
    for (int i = il - ng; i <= iu + ng + 1; ++i) {
      long noffset = (long)(i - il) + lx*bx;
      double rx = (double)noffset/(double)nx_mesh;   /* ComputeMeshGeneratorX, [0,1] */
      (*xf)[i] = default_gen(rx, mmin, mmax, rat, nx_mesh);
    }
    (*xf)[il] = bmin;
    (*xf)[iu+1] = bmax;
    for (int i = il - ng; i <= iu + ng; ++i) (*dxf)[i] = (*xf)[i+1] - (*xf)[i];
  } else {
    double dx = (bmax - bmin)/(iu - il + 1);
    for (int i = il - ng; i <= iu + ng + 1; ++i) {
      long noffset = (long)(i - il) + lx*bx;
      double rx = mesh_gen_x(noffset, nx_mesh);
      (*xf)[i] = uniform_gen(rx, mmin, mmax);
    }
    (*xf)[il] = bmin;
    (*xf)[iu+1] = bmax;
    for (int i = il - ng; i <= iu + ng; ++i) (*dxf)[i] = dx;
  }
  /* reflecting boundaries mirror the ghost spacing (coordinates.cpp:147-160) */
  if (refl_in) for (int i = 1; i <= ng; ++i) {
    (*dxf)[il-i] = (*dxf)[il+i-1];
    (*xf)[il-i] = (*xf)[il-i+1] - (*dxf)[il-i];
  }
  if (refl_out) for (int i = 1; i <= ng; ++i) {
    (*dxf)[iu+i] = (*dxf)[iu-i+1];
    (*xf)[iu+i+1] = (*xf)[iu+i] + (*dxf)[iu+i];
  }
  for (int i = il - ng; i <= iu + ng; ++i) (*xv)[i] = 0.5*((*xf)[i+1] + (*xf)[i]);
}

/* Reconstruction ctor (src/reconstruct/reconstruction.cpp:196-213,422-461,549-582,599-640) and
 * the per-index factors the PLM / CalculateCellCenteredField loops evaluate on the fly
 * (plm.cpp:85-93,114-119,198-204,308-313; field.cpp:139-172); dx?v as cartesian.cpp:31-75.
 * Index ranges as in the reference: the x2 / x3 weight loops start one cell later than x1's,
 * anything outside keeps the zero of NewAthenaArray. */
static AoReconGeom *make_recon_geom(int dir, int nonuni, int nc, int s, int e, int ng,
                                    const double *xf, const double *xv, const double *dxf,
                                    double *bw[2]) {
  AoReconGeom *g = (AoReconGeom *)calloc((size_t)nc, sizeof(AoReconGeom));
  bw[0] = dalloc(nc); bw[1] = dalloc(nc);
  for (int i = 0; i < nc; ++i) {
    g[i].wp = (xf[i+1] - xv[i])/dxf[i];
    g[i].wm = (xv[i] - xf[i])/dxf[i];
    bw[0][i] = 0.5; bw[1][i] = 0.5;
  }
  if (!nonuni || nc == 1) return g;
  double *dxv = dalloc(nc);
  for (int i = s - ng; i <= e + ng - 1; ++i) dxv[i] = xv[i+1] - xv[i];
  double *c[6];
  for (int n = 0; n < 6; ++n) c[n] = dalloc(nc);
  int first = (dir == 0) ? s - ng + 1 : s - ng + 2;
  for (int i = first; i <= e + ng - 1; ++i) {
    double dx_im1 = dxf[i-1], dx_i = dxf[i], dx_ip1 = dxf[i+1];
    double qe = dx_i/(dx_im1 + dx_i + dx_ip1);
    c[0][i] = qe*(2.0*dx_im1+dx_i)/(dx_ip1 + dx_i);
    c[1][i] = qe*(2.0*dx_ip1+dx_i)/(dx_im1 + dx_i);
    if (i > s - ng + 1) {
      double dx_im2 = dxf[i-2];
      double qa = dx_im2 + dx_im1 + dx_i + dx_ip1;
      double qb = dx_im1/(dx_im1 + dx_i);
      double qc = (dx_im2 + dx_im1)/(2.0*dx_im1 + dx_i);
      double qd = (dx_ip1 + dx_i)/(2.0*dx_i + dx_im1);
      qb = qb + 2.0*dx_i*qb/qa*(qc-qd);
      c[2][i] = 1.0 - qb;
      c[3][i] = qb;
      c[4][i] = dx_i/qa*qd;
      c[5][i] = -dx_im1/qa*qc;
    }
  }
  for (int i = 0; i < nc; ++i) {
    AoReconGeom *q = &g[i];
    q->mode = dir + 1;
    q->dxf = dxf[i]; q->dxv = dxv[i]; q->dxvm = (i > 0) ? dxv[i-1] : 0.0;
    q->cf = q->dxv/(xf[i+1] - xv[i]);
    q->cb = q->dxvm/(xv[i] - xf[i]);
    q->dxF = q->dxf/q->dxv;
    q->dxB = q->dxf/q->dxvm;
    q->c1 = c[0][i]; q->c2 = c[1][i]; q->c3 = c[2][i]; q->c4 = c[3][i]; q->c5 = c[4][i];
    q->c6 = c[5][i];
    if (i > 0) { q->c1m = c[0][i-1]; q->c2m = c[1][i-1]; }
    if (i < nc - 1) {
      q->c1p = c[0][i+1]; q->c2p = c[1][i+1]; q->c3p = c[2][i+1]; q->c4p = c[3][i+1];
      q->c5p = c[4][i+1]; q->c6p = c[5][i+1];
    }
    bw[0][i] = (xf[i+1] - xv[i])/dxf[i];
    bw[1][i] = (xv[i] - xf[i])/dxf[i];
  }
  for (int n = 0; n < 6; ++n) free(c[n]);
  free(dxv);
  return g;
}

/* test entries: reconstruct cells lo..hi of a 1-D line along direction dir with the geometry of
 * (xf, xv, dxf); q[v*nc + i]; order 2 (PLM) / 3 (PPM).  Floors are not applied here. */
void ao_recon_line(int dir, int nonuni, int order, int nc, int s, int e, int ng, const double *xf,
                   const double *xv, const double *dxf, int nvar, const double *q, int lo, int hi,
                   double *plus, double *minus) {
  double *bw[2];
  AoReconGeom *g = make_recon_geom(dir, nonuni, nc, s, e, ng, xf, xv, dxf, bw);
  for (int v = 0; v < nvar; ++v) for (int i = lo; i <= hi; ++i) {
    const double *c = q + (long)v*nc + i;
    if (order == 2) ao_plm_point_g(c[-1], c[0], c[1], &g[i], &plus[(long)v*nc+i], &minus[(long)v*nc+i]);
    else ao_ppm_point_g(c[-2], c[-1], c[0], c[1], c[2], &g[i], &plus[(long)v*nc+i], &minus[(long)v*nc+i]);
  }
  free(g); free(bw[0]); free(bw[1]);
}
/* characteristic variant: 7 sweep-ordered variable slots, bx[i]; floors applied */
void ao_recon_line_char(int dir, int nonuni, int order, int mhd, int nc, int s, int e, int ng,
                        const double *xf, const double *xv, const double *dxf, const double *q,
                        const double *bx, double gamma, double dfloor, double pfloor, int lo,
                        int hi, double *plus, double *minus) {
  double *bw[2];
  AoReconGeom *g = make_recon_geom(dir, nonuni, nc, s, e, ng, xf, xv, dxf, bw);
  int nw = mhd ? 7 : 5;
  for (int i = lo; i <= hi; ++i) {
    double st[5][7], pl[7], mi[7];
    for (int o = -2; o <= 2; ++o) for (int v = 0; v < 7; ++v) st[o+2][v] = q[(long)v*nc + i + o];
    ao_recon_char_point(order, mhd, st, mhd ? bx[i] : 0.0, gamma, &g[i], dfloor, pfloor, pl, mi);
    for (int v = 0; v < nw; ++v) { plus[(long)v*nc+i] = pl[v]; minus[(long)v*nc+i] = mi[v]; }
  }
  free(g); free(bw[0]); free(bw[1]);
}
/* CalculateCellCenteredField weights of a line (field.cpp:139-172) */
void ao_bcc_weights(int dir, int nonuni, int nc, int s, int e, int ng, const double *xf,
                    const double *xv, const double *dxf, double *lw, double *rw) {
  double *bw[2];
  AoReconGeom *g = make_recon_geom(dir, nonuni, nc, s, e, ng, xf, xv, dxf, bw);
  for (int i = 0; i < nc; ++i) { lw[i] = bw[0][i]; rw[i] = bw[1][i]; }
  free(g); free(bw[0]); free(bw[1]);
}

/* Mesh::SetBlockSizeAndBoundaries (src/mesh/mesh.cpp:1668-1751) for one direction */
static double block_edge(long lx, int nrbx, double mmin, double mmax, double rat, int nx_mesh) {
  if (rat != 1.0) return default_gen((double)lx/(double)nrbx, mmin, mmax, rat, nx_mesh);
  return uniform_gen(mesh_gen_x(lx, nrbx), mmin, mmax);
}

static void block_extent(long lx, int nrbx, double mmin, double mmax, int bc_in, int bc_out,
                         int nx_mesh, double rat, double *bmin, double *bmax, int *bcs_in,
                         int *bcs_out) {
  if (nx_mesh == 1) {
    *bmin = mmin; *bmax = mmax; *bcs_in = bc_in; *bcs_out = bc_out;
    return;
  }
  if (lx == 0) { *bmin = mmin; *bcs_in = bc_in; }
  else { *bmin = block_edge(lx, nrbx, mmin, mmax, rat, nx_mesh); *bcs_in = -1; }
  if (lx == nrbx - 1) { *bmax = mmax; *bcs_out = bc_out; }
  else { *bmax = block_edge(lx + 1, nrbx, mmin, mmax, rat, nx_mesh); *bcs_out = -1; }
}

#include "oracle_smr.c"

static int find_ni(const AoMesh *m, int o1, int o2, int o3) {
  for (int n = 0; n < m->nni; ++n)
    if (m->ni[n][0] == o1 && m->ni[n][1] == o2 && m->ni[n][2] == o3) return n;
  return -1;
}

static int cmp_morton(const void *a, const void *b) {
  const unsigned long long *x = (const unsigned long long *)a, *y = (const unsigned long long *)b;
  return (x[0] > y[0]) - (x[0] < y[0]);
}

static void set_integrator(AoMesh *m) {
  /* src/task_list/time_integrator.cpp:104-604 */
  double cfl_limit = 1.0;
  for (int s = 0; s < 4; ++s) { m->g1[s] = 0; m->g2[s] = 1; m->g3[s] = 0; m->delta[s] = 0; }
  m->delta[0] = 1.0;
  for (int s = 0; s < 4; ++s) { m->ebeta[s] = 1.0; m->sbeta[s] = 0.0; }
  switch (m->p.integrator) {
    case AO_INT_VL2:
      m->nstages = 2; m->beta[0] = 0.5; m->beta[1] = 1.0; m->ebeta[0] = 0.5; m->sbeta[1] = 0.5;
      if (m->ndim >= 2) cfl_limit = 0.5;
      break;
    case AO_INT_RK1:
      m->nstages = 1; m->beta[0] = 1.0;
      break;
    case AO_INT_RK2:
      m->nstages = 2; m->beta[0] = 1.0; m->beta[1] = 0.5; m->sbeta[1] = 1.0;
      m->g1[1] = 0.5; m->g2[1] = 0.5;
      break;
    default: /* rk3 */
      m->nstages = 3; m->beta[0] = 1.0; m->beta[1] = 0.25; m->beta[2] = 0.66666666666666667;
      m->ebeta[1] = 0.5; m->sbeta[1] = 1.0; m->sbeta[2] = 0.5;
      m->g1[1] = 0.25; m->g2[1] = 0.75;
      m->g1[2] = 0.66666666666666667; m->g2[2] = 0.33333333333333333;
      break;
  }
  m->cfl = m->p.cfl;
  if (m->cfl > cfl_limit) m->cfl = cfl_limit;  /* time_integrator.cpp:886-894 */
}

AoMesh *ao_create(const AoParams *p) {
  /* static refinement is restated for hydro on uniformly spaced levels only (oracle_smr.c) */
  if (p->nref > 0 && (p->mhd || (p->xrat[0] != 0.0 && p->xrat[0] != 1.0)
                      || (p->xrat[1] != 0.0 && p->xrat[1] != 1.0)
                      || (p->xrat[2] != 0.0 && p->xrat[2] != 1.0) || p->nref > 8
                      || p->bx1 % 2 || (p->nx2 > 1 && p->bx2 % 2) || (p->nx3 > 1 && p->bx3 % 2)
                      || p->ng % 2))
    return NULL;
  AoMesh *m = (AoMesh *)calloc(1, sizeof(AoMesh));
  m->p = *p;
  for (int d = 0; d < 3; ++d) m->xrat[d] = (p->xrat[d] == 0.0) ? 1.0 : p->xrat[d];
  if (p->nx2 == 1) m->xrat[1] = 1.0;
  if (p->nx3 == 1) m->xrat[2] = 1.0;
  m->nh = (p->eos == 1) ? 4 : 5;
  m->f2 = p->nx2 > 1; m->f3 = p->nx3 > 1;
  m->ndim = m->f3 ? 3 : (m->f2 ? 2 : 1);
  m->nrbx1 = p->nx1/p->bx1; m->nrbx2 = p->nx2/p->bx2; m->nrbx3 = p->nx3/p->bx3;
  m->nb = m->nrbx1*m->nrbx2*m->nrbx3;
  m->time = p->start_time; m->dt = DBL_MAX; m->ncycle = 0;
  set_integrator(m);
  /* canonical neighbour enumeration, uniform mesh (bvals_base.cpp:153-256) */
  int b = 0;
  for (int n = -1; n <= 1; n += 2) { m->ni[b][0] = n; m->ni[b][1] = 0; m->ni[b][2] = 0; b++; }
  if (m->ndim >= 2)
    for (int n = -1; n <= 1; n += 2) { m->ni[b][0] = 0; m->ni[b][1] = n; m->ni[b][2] = 0; b++; }
  if (m->ndim == 3)
    for (int n = -1; n <= 1; n += 2) { m->ni[b][0] = 0; m->ni[b][1] = 0; m->ni[b][2] = n; b++; }
  if (m->ndim >= 2)
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) {
      m->ni[b][0] = n; m->ni[b][1] = mm; m->ni[b][2] = 0; b++; }
  if (m->ndim == 3) {
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) {
      m->ni[b][0] = n; m->ni[b][1] = 0; m->ni[b][2] = mm; b++; }
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) {
      m->ni[b][0] = 0; m->ni[b][1] = n; m->ni[b][2] = mm; b++; }
    for (int l = -1; l <= 1; l += 2) for (int mm = -1; mm <= 1; mm += 2)
      for (int n = -1; n <= 1; n += 2) {
        m->ni[b][0] = n; m->ni[b][1] = mm; m->ni[b][2] = l; b++; }
  }
  m->nni = b;

  /* Z-ordered block list (src/mesh/meshblock_tree.cpp:87-110,336-352) */
  unsigned long long (*keys)[4] = malloc(sizeof(unsigned long long[4])*(size_t)m->nb);
  int c = 0;
  for (int k = 0; k < m->nrbx3; ++k) for (int j = 0; j < m->nrbx2; ++j)
    for (int i = 0; i < m->nrbx1; ++i) {
      unsigned long long key = 0;
      for (int bit = 0; bit < 20; ++bit)
        key |= ((unsigned long long)((i >> bit) & 1) << (3*bit))
             | ((unsigned long long)((j >> bit) & 1) << (3*bit + 1))
             | ((unsigned long long)((k >> bit) & 1) << (3*bit + 2));
      keys[c][0] = key; keys[c][1] = i; keys[c][2] = j; keys[c][3] = k; c++;
    }
  qsort(keys, (size_t)m->nb, sizeof(keys[0]), cmp_morton);
  m->gid_of = (int *)malloc(sizeof(int)*(size_t)m->nb);
  TNode **leaves = NULL;
  m->multilevel = (p->nref > 0);
  if (m->multilevel) {   /* static refinement: block list from the tree (oracle_smr.c) */
    smr_build_tree(m);
    int cnt = 0;
    tree_list(m->root, NULL, &cnt);
    leaves = (TNode **)malloc(sizeof(TNode *)*(size_t)cnt);
    cnt = 0;
    tree_list(m->root, leaves, &cnt);
    m->nb = cnt;
  }
  m->blk = (AoBlock *)calloc((size_t)m->nb, sizeof(AoBlock));
  int ng = p->ng;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    int dl = 0;          /* level above the root grid */
    if (m->multilevel) {
      B->gid = g; B->lx1 = leaves[g]->lx1; B->lx2 = leaves[g]->lx2; B->lx3 = leaves[g]->lx3;
      B->level = leaves[g]->level;
      dl = B->level - m->root_level;
    } else {
    B->gid = g; B->lx1 = (long)keys[g][1]; B->lx2 = (long)keys[g][2]; B->lx3 = (long)keys[g][3];
    m->gid_of[(B->lx3*m->nrbx2 + B->lx2)*m->nrbx1 + B->lx1] = g;
    }
    const int nr1 = m->nrbx1 << dl, nr2 = m->nrbx2 << dl, nr3 = m->nrbx3 << dl;
    const int nm1 = p->nx1 << dl, nm2 = m->f2 ? p->nx2 << dl : 1, nm3 = m->f3 ? p->nx3 << dl : 1;
    /* MeshBlock ctor index ranges (src/mesh/meshblock.cpp:55-80) */
    B->is = ng; B->ie = ng + p->bx1 - 1; B->nc1 = p->bx1 + 2*ng;
    if (m->f2) { B->js = ng; B->je = ng + p->bx2 - 1; B->nc2 = p->bx2 + 2*ng; }
    else { B->js = B->je = 0; B->nc2 = 1; }
    if (m->f3) { B->ks = ng; B->ke = ng + p->bx3 - 1; B->nc3 = p->bx3 + 2*ng; }
    else { B->ks = B->ke = 0; B->nc3 = 1; }
    /* nx1 > 1 always, so the x1 extent goes through the generator even for one block */
    block_extent(B->lx1, nr1, p->x1min, p->x1max, p->bc[0], p->bc[1], p->nx1 > 1 ? p->nx1 : 2,
                 m->xrat[0], &B->bx1min, &B->bx1max, &B->bcs[0], &B->bcs[1]);
    block_extent(B->lx2, nr2, p->x2min, p->x2max, p->bc[2], p->bc[3], p->nx2,
                 m->xrat[1], &B->bx2min, &B->bx2max, &B->bcs[2], &B->bcs[3]);
    block_extent(B->lx3, nr3, p->x3min, p->x3max, p->bc[4], p->bc[5], p->nx3,
                 m->xrat[2], &B->bx3min, &B->bx3max, &B->bcs[4], &B->bcs[5]);
    make_coords(nm1, p->bx1, ng, B->lx1, p->x1min, p->x1max, B->bx1min, B->bx1max,
                B->nc1, B->bcs[0] == AO_BC_REFLECT, B->bcs[1] == AO_BC_REFLECT, m->xrat[0],
                &B->x1f, &B->x1v, &B->dx1f);
    make_coords(nm2, p->bx2, ng, B->lx2, p->x2min, p->x2max, B->bx2min, B->bx2max,
                B->nc2, B->bcs[2] == AO_BC_REFLECT, B->bcs[3] == AO_BC_REFLECT, m->xrat[1],
                &B->x2f, &B->x2v, &B->dx2f);
    make_coords(nm3, p->bx3, ng, B->lx3, p->x3min, p->x3max, B->bx3min, B->bx3max,
                B->nc3, B->bcs[4] == AO_BC_REFLECT, B->bcs[5] == AO_BC_REFLECT, m->xrat[2],
                &B->x3f, &B->x3v, &B->dx3f);
    if (m->multilevel) {
      /* MeshBlock ctor, multilevel branch (meshblock.cpp:82-100): cnghost = (NGHOST+1)/2 + 1 */
      int cng = (ng + 1)/2 + 1;
      B->cng = cng;
      B->cis = cng; B->cie = cng + p->bx1/2 - 1; B->cnc1 = p->bx1/2 + 2*cng;
      if (m->f2) { B->cjs = cng; B->cje = cng + p->bx2/2 - 1; B->cnc2 = p->bx2/2 + 2*cng; }
      else { B->cjs = B->cje = 0; B->cnc2 = 1; }
      if (m->f3) { B->cks = cng; B->cke = cng + p->bx3/2 - 1; B->cnc3 = p->bx3/2 + 2*cng; }
      else { B->cks = B->cke = 0; B->cnc3 = 1; }
      make_coarse_coords(nm1, p->bx1, cng, B->lx1, p->x1min, p->x1max, B->bx1min, B->bx1max,
                         B->cnc1, B->bcs[0] == AO_BC_REFLECT, B->bcs[1] == AO_BC_REFLECT,
                         &B->cx1f, &B->cx1v);
      make_coarse_coords(nm2, p->bx2, cng, B->lx2, p->x2min, p->x2max, B->bx2min, B->bx2max,
                         B->cnc2, B->bcs[2] == AO_BC_REFLECT, B->bcs[3] == AO_BC_REFLECT,
                         &B->cx2f, &B->cx2v);
      make_coarse_coords(nm3, p->bx3, cng, B->lx3, p->x3min, p->x3max, B->bx3min, B->bx3max,
                         B->cnc3, B->bcs[4] == AO_BC_REFLECT, B->bcs[5] == AO_BC_REFLECT,
                         &B->cx3f, &B->cx3v);
      long cncc = (long)B->cnc1*B->cnc2*B->cnc3;
      B->coarse_u = dalloc(NHYDRO*cncc); B->coarse_w = dalloc(NHYDRO*cncc);
      if (p->nscalars > 0) { B->coarse_s = dalloc(p->nscalars*cncc); B->coarse_r = dalloc(p->nscalars*cncc); }
    }
    B->rg[0] = make_recon_geom(0, m->xrat[0] != 1.0, B->nc1, B->is, B->ie, ng, B->x1f, B->x1v,
                               B->dx1f, B->bw[0]);
    B->rg[1] = make_recon_geom(1, m->xrat[1] != 1.0, B->nc2, B->js, B->je, ng, B->x2f, B->x2v,
                               B->dx2f, B->bw[1]);
    B->rg[2] = make_recon_geom(2, m->xrat[2] != 1.0, B->nc3, B->ks, B->ke, ng, B->x3f, B->x3v,
                               B->dx3f, B->bw[2]);
    long ncc = (long)B->nc1*B->nc2*B->nc3;
    B->u = dalloc(NHYDRO*ncc); B->u1 = dalloc(NHYDRO*ncc); B->w = dalloc(NHYDRO*ncc);
    B->flux[0] = dalloc(NHYDRO*(long)B->nc3*B->nc2*(B->nc1+1));
    B->flux[1] = dalloc(NHYDRO*(long)B->nc3*(B->nc2+1)*B->nc1);
    B->flux[2] = dalloc(NHYDRO*(long)(B->nc3+1)*B->nc2*B->nc1);
    if (p->nscalars > 0) {   /* PassiveScalars ctor (src/scalars/scalars.cpp:31-75) */
      int ns = p->nscalars;
      B->s = dalloc(ns*ncc); B->s1 = dalloc(ns*ncc); B->r = dalloc(ns*ncc);
      B->sflux[0] = dalloc(ns*(long)B->nc3*B->nc2*(B->nc1+1));
      B->sflux[1] = dalloc(ns*(long)B->nc3*(B->nc2+1)*B->nc1);
      B->sflux[2] = dalloc(ns*(long)(B->nc3+1)*B->nc2*B->nc1);
    }
    if (p->mhd) {
      long n1 = (long)B->nc3*B->nc2*(B->nc1+1), n2 = (long)B->nc3*(B->nc2+1)*B->nc1,
           n3 = (long)(B->nc3+1)*B->nc2*B->nc1;
      B->b[0] = dalloc(n1); B->b[1] = dalloc(n2); B->b[2] = dalloc(n3);
      B->b1[0] = dalloc(n1); B->b1[1] = dalloc(n2); B->b1[2] = dalloc(n3);
      B->wght[0] = dalloc(n1); B->wght[1] = dalloc(n2); B->wght[2] = dalloc(n3);
      B->e2_x1f = dalloc(n1); B->e3_x1f = dalloc(n1);
      B->e1_x2f = dalloc(n2); B->e3_x2f = dalloc(n2);
      B->e1_x3f = dalloc(n3); B->e2_x3f = dalloc(n3);
      B->bcc = dalloc(3*ncc);
      B->e[0] = dalloc((long)(B->nc3+1)*(B->nc2+1)*B->nc1);
      B->e[1] = dalloc((long)(B->nc3+1)*B->nc2*(B->nc1+1));
      B->e[2] = dalloc((long)B->nc3*(B->nc2+1)*(B->nc1+1));
      B->cc_e = dalloc(3*ncc);
    }
  }
  free(keys);
  free(leaves);
  if (m->multilevel) {
    for (int g = 0; g < m->nb; ++g) smr_search_neighbors(m, &m->blk[g]);
    return m;
  }
  /* neighbours (src/bvals/bvals_base.cpp:299-480), same level only */
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
      B->nblevel[k][j][i] = -1;
    B->nblevel[1][1][1] = 0;
    B->nnb = 0;
    for (int n = 0; n < m->nni; ++n) {
      int o1 = m->ni[n][0], o2 = m->ni[n][1], o3 = m->ni[n][2];
      long l1 = B->lx1 + o1, l2 = B->lx2 + o2, l3 = B->lx3 + o3;
      int ok = 1;
      if (l1 < 0) { if (p->bc[0] == AO_BC_PERIODIC) l1 = m->nrbx1 - 1; else ok = 0; }
      if (l1 >= m->nrbx1) { if (p->bc[1] == AO_BC_PERIODIC) l1 = 0; else ok = 0; }
      if (l2 < 0) { if (p->bc[2] == AO_BC_PERIODIC) l2 = m->nrbx2 - 1; else ok = 0; }
      if (l2 >= m->nrbx2) { if (p->bc[3] == AO_BC_PERIODIC) l2 = 0; else ok = 0; }
      if (l3 < 0) { if (p->bc[4] == AO_BC_PERIODIC) l3 = m->nrbx3 - 1; else ok = 0; }
      if (l3 >= m->nrbx3) { if (p->bc[5] == AO_BC_PERIODIC) l3 = 0; else ok = 0; }
      if (!ok) continue;
      Nb *nb = &B->nb[B->nnb++];
      nb->ox1 = o1; nb->ox2 = o2; nb->ox3 = o3;
      int nz = (o1 != 0) + (o2 != 0) + (o3 != 0);
      nb->type = nz - 1;
      nb->gid = m->gid_of[(l3*m->nrbx2 + l2)*m->nrbx1 + l1];
      nb->bufid = n;
      nb->targetid = find_ni(m, -o1, -o2, -o3);
      nb->fid = -1; nb->eid = -1;
      /* NeighborBlock::SetNeighbor (bvals_base.cpp:54-66) */
      if (nb->type == 0) {
        if (o1 == -1) nb->fid = 0; else if (o1 == 1) nb->fid = 1;
        else if (o2 == -1) nb->fid = 2; else if (o2 == 1) nb->fid = 3;
        else if (o3 == -1) nb->fid = 4; else nb->fid = 5;
      } else if (nb->type == 1) {
        if (o3 == 0) nb->eid = (((o1 + 1) >> 1) | ((o2 + 1) & 2));
        else if (o2 == 0) nb->eid = (4 + (((o1 + 1) >> 1) | ((o3 + 1) & 2)));
        else nb->eid = (8 + (((o2 + 1) >> 1) | ((o3 + 1) & 2)));
      }
      B->nblevel[o3+1][o2+1][o1+1] = 0;
    }
    /* CountFineEdges (src/bvals/fc/bvals_fc.cpp:1129-1193), single level */
    int eid = 0;
    for (int e = 0; e < 12; ++e) B->nedge_fine[e] = 1;
    if (m->f2) {
      for (int o2 = -1; o2 <= 1; o2 += 2) for (int o1 = -1; o1 <= 1; o1 += 2) {
        int nis = (o1-1 > -1) ? o1-1 : -1, nie = (o1+1 < 1) ? o1+1 : 1;
        int njs = (o2-1 > -1) ? o2-1 : -1, nje = (o2+1 < 1) ? o2+1 : 1;
        int nf = 0;
        for (int nj = njs; nj <= nje; nj++) for (int ni = nis; ni <= nie; ni++)
          if (B->nblevel[1][nj+1][ni+1] == 0) nf++;
        B->nedge_fine[eid++] = nf;
      }
    }
    if (m->f3) {
      for (int o3 = -1; o3 <= 1; o3 += 2) for (int o1 = -1; o1 <= 1; o1 += 2) {
        int nis = (o1-1 > -1) ? o1-1 : -1, nie = (o1+1 < 1) ? o1+1 : 1;
        int nks = (o3-1 > -1) ? o3-1 : -1, nke = (o3+1 < 1) ? o3+1 : 1;
        int nf = 0;
        for (int nk = nks; nk <= nke; nk++) for (int ni = nis; ni <= nie; ni++)
          if (B->nblevel[nk+1][1][ni+1] == 0) nf++;
        B->nedge_fine[eid++] = nf;
      }
      for (int o3 = -1; o3 <= 1; o3 += 2) for (int o2 = -1; o2 <= 1; o2 += 2) {
        int njs = (o2-1 > -1) ? o2-1 : -1, nje = (o2+1 < 1) ? o2+1 : 1;
        int nks = (o3-1 > -1) ? o3-1 : -1, nke = (o3+1 < 1) ? o3+1 : 1;
        int nf = 0;
        for (int nk = nks; nk <= nke; nk++) for (int nj = njs; nj <= nje; nj++)
          if (B->nblevel[nk+1][nj+1][1] == 0) nf++;
        B->nedge_fine[eid++] = nf;
      }
    }
  }
  return m;
}

void ao_destroy(AoMesh *m) {
  if (!m) return;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    double *ptrs[] = {B->x1f, B->x2f, B->x3f, B->x1v, B->x2v, B->x3v, B->dx1f, B->dx2f,
      B->dx3f, B->u, B->u1, B->w, B->bcc, B->flux[0], B->flux[1], B->flux[2], B->b[0],
      B->b[1], B->b[2], B->b1[0], B->b1[1], B->b1[2], B->e[0], B->e[1], B->e[2],
      B->wght[0], B->wght[1], B->wght[2], B->e2_x1f, B->e3_x1f, B->e1_x2f, B->e3_x2f,
      B->e1_x3f, B->e2_x3f, B->cc_e, B->s, B->s1, B->r, B->sflux[0], B->sflux[1],
      B->sflux[2]};
    for (size_t i = 0; i < sizeof(ptrs)/sizeof(ptrs[0]); ++i) free(ptrs[i]);
    for (int d = 0; d < 3; ++d) { free(B->rg[d]); free(B->bw[d][0]); free(B->bw[d][1]); }
    free(B->cx1f); free(B->cx2f); free(B->cx3f); free(B->cx1v); free(B->cx2v); free(B->cx3v);
    free(B->coarse_u); free(B->coarse_w); free(B->coarse_s); free(B->coarse_r);
  }
  tree_free(m->root);
  free(m->blk); free(m->gid_of); free(m);
}

int ao_nblocks(const AoMesh *m) { return m->nb; }
int ao_block_level(const AoMesh *m, int b) { return m->blk[b].level; }
int ao_neighbors(const AoMesh *m, int b, int *rows, int *nblevel) {
  const AoBlock *B = &m->blk[b];
  for (int n = 0; n < B->nnb && rows; ++n) {
    const Nb *nb = &B->nb[n];
    int *r = rows + 8*n;
    r[0] = nb->ox1; r[1] = nb->ox2; r[2] = nb->ox3; r[3] = nb->type; r[4] = nb->gid;
    r[5] = m->multilevel ? nb->level : 0; r[6] = nb->fi1; r[7] = nb->fi2;
  }
  if (nblevel) for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
    nblevel[(k*3 + j)*3 + i] = B->nblevel[k][j][i];
  return B->nnb;
}
double ao_time(const AoMesh *m) { return m->time; }
double ao_dt(const AoMesh *m) { return m->dt; }
int ao_ncycle(const AoMesh *m) { return m->ncycle; }
void ao_set_time_dt(AoMesh *m, double time, double dt) { m->time = time; m->dt = dt; }

void ao_block_info(const AoMesh *m, int b, long *out) {
  const AoBlock *B = &m->blk[b];
  out[0] = B->lx1; out[1] = B->lx2; out[2] = B->lx3;
  out[3] = B->nc1; out[4] = B->nc2; out[5] = B->nc3;
  out[6] = B->is; out[7] = B->ie; out[8] = B->js; out[9] = B->je; out[10] = B->ks;
  out[11] = B->ke;
}

double *ao_array(AoMesh *m, int b, const char *name, long *n) {
  AoBlock *B = &m->blk[b];
  long ncc = (long)B->nc1*B->nc2*B->nc3;
  long n1 = (long)B->nc3*B->nc2*(B->nc1+1), n2 = (long)B->nc3*(B->nc2+1)*B->nc1,
       n3 = (long)(B->nc3+1)*B->nc2*B->nc1;
  long ne1 = (long)(B->nc3+1)*(B->nc2+1)*B->nc1, ne2 = (long)(B->nc3+1)*B->nc2*(B->nc1+1),
       ne3 = (long)B->nc3*(B->nc2+1)*(B->nc1+1);
  struct { const char *nm; double *p; long n; } t[] = {
    {"u", B->u, NHYDRO*ncc}, {"u1", B->u1, NHYDRO*ncc}, {"w", B->w, NHYDRO*ncc},
    {"bcc", B->bcc, 3*ncc}, {"b1", B->b[0], n1}, {"b2", B->b[1], n2}, {"b3", B->b[2], n3},
    {"b1_1", B->b1[0], n1}, {"b1_2", B->b1[1], n2}, {"b1_3", B->b1[2], n3},
    {"flux1", B->flux[0], NHYDRO*n1}, {"flux2", B->flux[1], NHYDRO*n2},
    {"flux3", B->flux[2], NHYDRO*n3}, {"e1", B->e[0], ne1}, {"e2", B->e[1], ne2},
    {"e3", B->e[2], ne3}, {"wght1", B->wght[0], n1}, {"wght2", B->wght[1], n2},
    {"wght3", B->wght[2], n3}, {"e2_x1f", B->e2_x1f, n1}, {"e3_x1f", B->e3_x1f, n1},
    {"e1_x2f", B->e1_x2f, n2}, {"e3_x2f", B->e3_x2f, n2}, {"e1_x3f", B->e1_x3f, n3},
    {"e2_x3f", B->e2_x3f, n3}, {"x1f", B->x1f, B->nc1+1}, {"x2f", B->x2f, B->nc2+1},
    {"x3f", B->x3f, B->nc3+1}, {"x1v", B->x1v, B->nc1}, {"x2v", B->x2v, B->nc2},
    {"x3v", B->x3v, B->nc3}, {"dx1f", B->dx1f, B->nc1}, {"dx2f", B->dx2f, B->nc2},
    {"dx3f", B->dx3f, B->nc3}, {"cc_e", B->cc_e, 3*ncc},
    {"coarse_u", B->coarse_u, NHYDRO*(long)B->cnc1*B->cnc2*B->cnc3},
    {"coarse_w", B->coarse_w, NHYDRO*(long)B->cnc1*B->cnc2*B->cnc3},
    {"cx1v", B->cx1v, B->cnc1}, {"cx2v", B->cx2v, B->cnc2}, {"cx3v", B->cx3v, B->cnc3},
    {"coarse_s", B->coarse_s, m->p.nscalars*(long)B->cnc1*B->cnc2*B->cnc3},
    {"coarse_r", B->coarse_r, m->p.nscalars*(long)B->cnc1*B->cnc2*B->cnc3},
    {"s", B->s, m->p.nscalars*ncc}, {"s1", B->s1, m->p.nscalars*ncc},
    {"r", B->r, m->p.nscalars*ncc}, {"sflux1", B->sflux[0], m->p.nscalars*n1},
    {"sflux2", B->sflux[1], m->p.nscalars*n2}, {"sflux3", B->sflux[2], m->p.nscalars*n3}};
  for (size_t i = 0; i < sizeof(t)/sizeof(t[0]); ++i)
    if (strcmp(t[i].nm, name) == 0) { *n = t[i].p ? t[i].n : 0; return t[i].p; }
  *n = 0;
  return NULL;
}

/* ------------------------------------------------------------------ EOS */

/* Field::CalculateCellCenteredField (src/field/field.cpp:112-180, uniform weights) +
 * EquationOfState::ConservedToPrimitive (src/eos/adiabatic_mhd.cpp:41-90 /
 * adiabatic_hydro.cpp:39-80) */
void ao_cons2prim(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku) {
  AoBlock *B = &m->blk[b];
  double gm1 = m->p.gamma - 1.0;
  double dfl = m->p.dfloor, pfl = m->p.pfloor;
  for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
    double pb = 0.0;
    if (m->p.mhd) {
      double bcc1 = B->bw[0][0][i]*B->b[0][F1(B,k,j,i)] + B->bw[0][1][i]*B->b[0][F1(B,k,j,i+1)];
      double bcc2 = B->bw[1][0][j]*B->b[1][F2(B,k,j,i)] + B->bw[1][1][j]*B->b[1][F2(B,k,j+1,i)];
      double bcc3 = B->bw[2][0][k]*B->b[2][F3(B,k,j,i)] + B->bw[2][1][k]*B->b[2][F3(B,k+1,j,i)];
      B->bcc[CC(B,IB1,k,j,i)] = bcc1;
      B->bcc[CC(B,IB2,k,j,i)] = bcc2;
      B->bcc[CC(B,IB3,k,j,i)] = bcc3;
      pb = 0.5*(SQR(bcc1) + SQR(bcc2) + SQR(bcc3));
    }
    double *u_d = &B->u[CC(B,IDN,k,j,i)];
    double u_m1 = B->u[CC(B,IM1,k,j,i)], u_m2 = B->u[CC(B,IM2,k,j,i)],
           u_m3 = B->u[CC(B,IM3,k,j,i)];
    *u_d = (*u_d > dfl) ? *u_d : dfl;
    B->w[CC(B,IDN,k,j,i)] = *u_d;
    double di = 1.0/(*u_d);
    B->w[CC(B,IVX,k,j,i)] = u_m1*di;
    B->w[CC(B,IVY,k,j,i)] = u_m2*di;
    B->w[CC(B,IVZ,k,j,i)] = u_m3*di;
    if (ISO(m)) continue;   /* isothermal_{hydro,mhd}.cpp: density floor and velocities only */
    double *u_e = &B->u[CC(B,IEN,k,j,i)];
    double e_k = 0.5*di*(SQR(u_m1) + SQR(u_m2) + SQR(u_m3));
    double w_p;
    if (m->p.mhd) {
      w_p = gm1*(*u_e - e_k - pb);
      *u_e = (w_p > pfl) ? *u_e : ((pfl/gm1) + e_k + pb);
    } else {
      w_p = gm1*(*u_e - e_k);
      *u_e = (w_p > pfl) ? *u_e : ((pfl/gm1) + e_k);
    }
    w_p = (w_p > pfl) ? w_p : pfl;
    B->w[CC(B,IPR,k,j,i)] = w_p;
  }
}

/* cell-centred field only (used by physical BCs, bvals.cpp:466-470) */
static void calc_bcc(AoBlock *B, int il, int iu, int jl, int ju, int kl, int ku) {
  for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
    B->bcc[CC(B,IB1,k,j,i)] = B->bw[0][0][i]*B->b[0][F1(B,k,j,i)] + B->bw[0][1][i]*B->b[0][F1(B,k,j,i+1)];
    B->bcc[CC(B,IB2,k,j,i)] = B->bw[1][0][j]*B->b[1][F2(B,k,j,i)] + B->bw[1][1][j]*B->b[1][F2(B,k,j+1,i)];
    B->bcc[CC(B,IB3,k,j,i)] = B->bw[2][0][k]*B->b[2][F3(B,k,j,i)] + B->bw[2][1][k]*B->b[2][F3(B,k+1,j,i)];
  }
}

/* EquationOfState::PrimitiveToConserved (adiabatic_hydro.cpp:89-123, adiabatic_mhd.cpp:99-136) */
void ao_prim2cons(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku) {
  AoBlock *B = &m->blk[b];
  double igm1 = ISO(m) ? 0.0 : 1.0/(m->p.gamma - 1.0);
  for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
    double w_d = B->w[CC(B,IDN,k,j,i)], w_vx = B->w[CC(B,IVX,k,j,i)],
           w_vy = B->w[CC(B,IVY,k,j,i)], w_vz = B->w[CC(B,IVZ,k,j,i)],
           w_p = ISO(m) ? 0.0 : B->w[CC(B,IPR,k,j,i)];
    B->u[CC(B,IDN,k,j,i)] = w_d;
    B->u[CC(B,IM1,k,j,i)] = w_vx*w_d;
    B->u[CC(B,IM2,k,j,i)] = w_vy*w_d;
    B->u[CC(B,IM3,k,j,i)] = w_vz*w_d;
    if (ISO(m)) continue;
    if (m->p.mhd) {
      double bcc1 = B->bcc[CC(B,IB1,k,j,i)], bcc2 = B->bcc[CC(B,IB2,k,j,i)],
             bcc3 = B->bcc[CC(B,IB3,k,j,i)];
      B->u[CC(B,IEN,k,j,i)] = w_p*igm1 + 0.5*(w_d*(SQR(w_vx) + SQR(w_vy) + SQR(w_vz))
                                              + (SQR(bcc1) + SQR(bcc2) + SQR(bcc3)));
    } else {
      B->u[CC(B,IEN,k,j,i)] = w_p*igm1 + 0.5*w_d*(SQR(w_vx) + SQR(w_vy) + SQR(w_vz));
    }
  }
}

/* TimeIntegratorTaskList::Primitives range (src/task_list/time_integrator.cpp:1965-1983) */
void ao_primitives(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  int ng = m->p.ng;
  int il = B->is, iu = B->ie, jl = B->js, ju = B->je, kl = B->ks, ku = B->ke;
  if (B->nblevel[1][1][0] != -1) il -= ng;
  if (B->nblevel[1][1][2] != -1) iu += ng;
  if (B->nblevel[1][0][1] != -1) jl -= ng;
  if (B->nblevel[1][2][1] != -1) ju += ng;
  if (B->nblevel[0][1][1] != -1) kl -= ng;
  if (B->nblevel[2][1][1] != -1) ku += ng;
  ao_cons2prim(m, b, il, iu, jl, ju, kl, ku);
  if (m->p.nscalars > 0) ao_scalar_cons2prim(m, b, il, iu, jl, ju, kl, ku);
}

/* EquationOfState::PassiveScalarConservedToPrimitive (src/eos/eos_scalars.cpp:31-60): floor
 * the conserved scalar at sfloor*rho (rho = the already floored u(IDN)), then r = s/rho */
void ao_scalar_cons2prim(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku) {
  AoBlock *B = &m->blk[b];
  double sfl = m->p.sfloor;
  for (int n = 0; n < m->p.nscalars; ++n) for (int k = kl; k <= ku; ++k)
    for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
      double d = B->u[CC(B,IDN,k,j,i)];
      double *s_n = &B->s[CC(B,n,k,j,i)];
      *s_n = (*s_n < sfl*d) ? sfl*d : *s_n;
      B->r[CC(B,n,k,j,i)] = *s_n/d;
    }
}

/* EquationOfState::PassiveScalarPrimitiveToConserved (src/eos/eos_scalars.cpp:133-152) */
void ao_scalar_prim2cons(AoMesh *m, int b, int il, int iu, int jl, int ju, int kl, int ku) {
  AoBlock *B = &m->blk[b];
  for (int n = 0; n < m->p.nscalars; ++n) for (int k = kl; k <= ku; ++k)
    for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
      B->s[CC(B,n,k,j,i)] = B->r[CC(B,n,k,j,i)]*B->u[CC(B,IDN,k,j,i)];
}

/* ------------------------------------------------------------------ fluxes */

/* gather the NWAVE sweep-ordered primitives of one cell (plm.cpp:45-58,159-172,272-285) */
static void cell_state(const AoMesh *m, const AoBlock *B, int dir, int k, int j, int i,
                       double *q) {
  for (int n = 0; n < NHYDRO; ++n) q[n] = B->w[CC(B,n,k,j,i)];
  if (ISO(m)) q[IPR] = 0.0;    /* slot unused: the isothermal wave vector has no pressure */
  if (m->p.mhd) {
    int by = (dir + 1) % 3, bz = (dir + 2) % 3;
    q[IBY] = B->bcc[CC(B,by,k,j,i)];
    q[IBZ] = B->bcc[CC(B,bz,k,j,i)];
  }
}

/* L/R states of cell (k,j,i) along dir: `plus` is the state at its upper face (becomes wl
 * of face+1), `minus` at its lower face (wr of its own face).
 * dc.cpp:24-110, plm.cpp:27-370, ppm.cpp:44-940 */
/* sweep-ordered copy: velocities rotated so that slot IVX is the sweep direction */
static void to_sweep(int dir, const double *q, double *p) {
  int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
  p[IDN] = q[IDN]; p[IVX] = q[ivx]; p[IVY] = q[ivy]; p[IVZ] = q[ivz]; p[IPR] = q[IPR];
  p[IBY] = q[IBY]; p[IBZ] = q[IBZ];
}
static void from_sweep(int dir, const double *p, double *q) {
  int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
  q[IDN] = p[IDN]; q[ivx] = p[IVX]; q[ivy] = p[IVY]; q[ivz] = p[IVZ]; q[IPR] = p[IPR];
  q[IBY] = p[IBY]; q[IBZ] = p[IBZ];
}

/* xorder = 2c / 3c: the same cell reconstruction on characteristic variables
 * (plm.cpp:62-66,107-110,122-130; ppm.cpp:66-75,311-332; characteristic.cpp).  The
 * eigenvectors are those of the cell itself; floors are re-applied to both face states. */
static void recon_cell_char(const AoMesh *m, const AoBlock *B, int dir, int order, int k, int j,
                            int i, double *plus, double *minus) {
  int nw = m->p.mhd ? 7 : 5, mhd = m->p.mhd;
  int dk = (dir == 2), dj = (dir == 1), di = (dir == 0);
  double gamma = m->p.gamma;
  double t[7], q[7], pl[7], mi[7];
  cell_state(m, B, dir, k, j, i, t); to_sweep(dir, t, q);
  double bx = mhd ? B->bcc[CC(B,dir,k,j,i)] : 0.0;
  double st[5][7];     /* stencil -2..+2 in sweep order */
  for (int o = -2; o <= 2; ++o) {
    if (order == 2 && (o == -2 || o == 2)) continue;
    cell_state(m, B, dir, k+o*dk, j+o*dj, i+o*di, t); to_sweep(dir, t, st[o+2]);
  }
  int c = dir == 0 ? i : (dir == 1 ? j : k);
  if (order == 2) for (int n = 0; n < 7; ++n) { st[0][n] = 0.0; st[4][n] = 0.0; }
  (void)q; (void)nw;
  ao_recon_char_point(order, mhd, st, bx, gamma, &B->rg[dir][c], m->p.dfloor, m->p.pfloor, pl, mi);
  from_sweep(dir, pl, plus);
  from_sweep(dir, mi, minus);
}

static void recon_cell(const AoMesh *m, const AoBlock *B, int dir, int order, int k, int j,
                       int i, double *plus, double *minus) {
  int nw = m->p.mhd ? 7 : 5;
  int dk = (dir == 2), dj = (dir == 1), di = (dir == 0);
  double q[7];
  if (m->p.char_proj && order > 1) { recon_cell_char(m, B, dir, order, k, j, i, plus, minus); return; }
  cell_state(m, B, dir, k, j, i, q);
  if (order == 1) {
    for (int n = 0; n < nw; ++n) plus[n] = minus[n] = q[n];
    return;
  }
  double qm1[7], qp1[7];
  cell_state(m, B, dir, k-dk, j-dj, i-di, qm1);
  cell_state(m, B, dir, k+dk, j+dj, i+di, qp1);
  const AoReconGeom *rg = &B->rg[dir][dir == 0 ? i : (dir == 1 ? j : k)];
  if (order == 2) {
    for (int n = 0; n < nw; ++n) ao_plm_point_g(qm1[n], q[n], qp1[n], rg, &plus[n], &minus[n]);
    return;
  }
  double qm2[7], qp2[7];
  cell_state(m, B, dir, k-2*dk, j-2*dj, i-2*di, qm2);
  cell_state(m, B, dir, k+2*dk, j+2*dj, i+2*di, qp2);
  for (int n = 0; n < nw; ++n)
    ao_ppm_point_g(qm2[n], qm1[n], q[n], qp1[n], qp2[n], rg, &plus[n], &minus[n]);
  /* ApplyPrimitiveFloors on both (ppm.cpp:326-332) */
  plus[IDN] = (plus[IDN] > m->p.dfloor) ? plus[IDN] : m->p.dfloor;
  minus[IDN] = (minus[IDN] > m->p.dfloor) ? minus[IDN] : m->p.dfloor;
  if (!ISO(m)) {
    plus[IPR] = (plus[IPR] > m->p.pfloor) ? plus[IPR] : m->p.pfloor;
    minus[IPR] = (minus[IPR] > m->p.pfloor) ? minus[IPR] : m->p.pfloor;
  }
}

/* Hydro::CalculateVelocityDifferences (src/hydro/calculate_velocity_differences.cpp:20-90):
 * normal velocity jump dvn and the minimum transverse velocity difference dvt around the
 * interface (LHLLC / LHLLD shock detector).  1-D leaves dvt at its zero initial value. */
static void velocity_differences(const AoMesh *m, const AoBlock *B, int dir, int k, int j, int i,
                                 double *dvn, double *dvt) {
  const double *w = B->w;
  int d[3][3] = {{0,0,1},{0,1,0},{1,0,0}};   /* (dk,dj,di) of x1,x2,x3 */
  int ivx = IVX + dir;
  int lk = k - d[dir][0], lj = j - d[dir][1], li = i - d[dir][2];   /* lower cell */
  *dvn = w[CC(B,ivx,k,j,i)] - w[CC(B,ivx,lk,lj,li)];
  *dvt = 0.0;
  if (!m->f2) return;
  /* transverse directions in the reference's order: (dir+1)%3 then (dir+2)%3, skipping x3 in 2-D */
  double res = 0.0;
  int first = 1;
  for (int t = 1; t <= 2; ++t) {
    int td = (dir + t) % 3;
    if (!m->f3 && td == 2) continue;
    int tv = IVX + td;
    int tk = d[td][0], tj = d[td][1], ti = d[td][2];
    double dl = mn(w[CC(B,tv,lk+tk,lj+tj,li+ti)] - w[CC(B,tv,lk,lj,li)],
                   w[CC(B,tv,lk,lj,li)] - w[CC(B,tv,lk-tk,lj-tj,li-ti)]);
    double dr = mn(w[CC(B,tv,k+tk,j+tj,i+ti)] - w[CC(B,tv,k,j,i)],
                   w[CC(B,tv,k,j,i)] - w[CC(B,tv,k-tk,j-tj,i-ti)]);
    double v = mn(dl, dr);
    if (first) { res = v; first = 0; } else { res = mn(res, v); }
  }
  *dvt = res;
}

/* one interface: face (k,j,i) of direction dir lies between cell-1 and cell */
static void face_flux(AoMesh *m, AoBlock *B, int dir, int order, int k, int j, int i) {
  int dk = (dir == 2), dj = (dir == 1), di = (dir == 0);
  double wl[7], wr[7], tmp[7], wli[7], wri[7], f[7];
  recon_cell(m, B, dir, order, k-dk, j-dj, i-di, wl, tmp);
  recon_cell(m, B, dir, order, k, j, i, tmp, wr);
  /* rotate velocities: ivx = IVX+dir (hlld.cpp:44-45,66-82) */
  int ivx = IVX + dir, ivy = IVX + (dir + 1) % 3, ivz = IVX + (dir + 2) % 3;
  wli[IDN] = wl[IDN]; wli[IVX] = wl[ivx]; wli[IVY] = wl[ivy]; wli[IVZ] = wl[ivz];
  wli[IPR] = wl[IPR]; wli[IBY] = wl[IBY]; wli[IBZ] = wl[IBZ];
  wri[IDN] = wr[IDN]; wri[IVX] = wr[ivx]; wri[IVY] = wr[ivy]; wri[IVZ] = wr[ivz];
  wri[IPR] = wr[IPR]; wri[IBY] = wr[IBY]; wri[IBZ] = wr[IBZ];
  double bxi = 0.0;
  if (m->p.mhd)
    bxi = dir == 0 ? B->b[0][F1(B,k,j,i)] : (dir == 1 ? B->b[1][F2(B,k,j,i)]
                                                       : B->b[2][F3(B,k,j,i)]);
  double dvn = 0.0, dvt = 0.0;
  if (m->p.solver == AO_SOLVER_LHLLC || m->p.solver == AO_SOLVER_LHLLD)
    velocity_differences(m, B, dir, k, j, i, &dvn, &dvt);
  if (ISO(m))
    ao_riemann_point_iso(m->p.solver, m->p.mhd, wli, wri, bxi, m->p.iso_cs, m->p.dfloor, f);
  else
    ao_riemann_point(m->p.solver, m->p.mhd, wli, wri, bxi, m->p.gamma, dvn, dvt, f);
  double *flx = B->flux[dir];
  long o[5];
  for (int n = 0; n < NHYDRO; ++n)
    o[n] = dir == 0 ? FL1(B,n,k,j,i) : (dir == 1 ? FL2(B,n,k,j,i) : FL3(B,n,k,j,i));
  flx[o[IDN]] = f[IDN];
  flx[o[ivx]] = f[IVX];
  flx[o[ivy]] = f[IVY];
  flx[o[ivz]] = f[IVZ];
  if (!ISO(m)) flx[o[IEN]] = f[IEN];
  if (m->p.mhd) {
    /* ey = -F(By), ez = F(Bz); CT weight (hlld.cpp:371-379); dxw = CenterWidth = dx?f */
    long fo = dir == 0 ? F1(B,k,j,i) : (dir == 1 ? F2(B,k,j,i) : F3(B,k,j,i));
    double dxw = dir == 0 ? B->dx1f[i] : (dir == 1 ? B->dx2f[j] : B->dx3f[k]);
    double *ey = dir == 0 ? B->e3_x1f : (dir == 1 ? B->e1_x2f : B->e2_x3f);
    double *ez = dir == 0 ? B->e2_x1f : (dir == 1 ? B->e3_x2f : B->e1_x3f);
    ey[fo] = -f[IBY];
    ez[fo] = f[IBZ];
    B->wght[dir][fo] = ao_weight_for_ct(f[IDN], wli[IDN], wri[IDN], dxw, m->dt);
  }
}

/* Hydro::CalculateFluxes loop limits (src/hydro/calculate_fluxes.cpp:62-74,164-173,273-279) */
void ao_calc_fluxes(AoMesh *m, int b, int order) {
  AoBlock *B = &m->blk[b];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  int il, iu, jl, ju, kl, ku;
  jl = js; ju = je; kl = ks; ku = ke;
  if (m->p.mhd) {
    if (m->f2) {
      if (!m->f3) { jl = js-1; ju = je+1; kl = ks; ku = ke; }
      else { jl = js-1; ju = je+1; kl = ks-1; ku = ke+1; }
    }
  }
  for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j)
    for (int i = is; i <= ie+1; ++i) face_flux(m, B, 0, order, k, j, i);
  if (m->f2) {
    il = is-1; iu = ie+1; kl = ks; ku = ke;
    if (m->p.mhd) { if (!m->f3) { kl = ks; ku = ke; } else { kl = ks-1; ku = ke+1; } }
    for (int k = kl; k <= ku; ++k) for (int j = js; j <= je+1; ++j)
      for (int i = il; i <= iu; ++i) face_flux(m, B, 1, order, k, j, i);
  }
  if (m->f3) {
    il = is; iu = ie; jl = js; ju = je;
    if (m->p.mhd) { il = is-1; iu = ie+1; jl = js-1; ju = je+1; }
    for (int k = ks; k <= ke+1; ++k) for (int j = jl; j <= ju; ++j)
      for (int i = il; i <= iu; ++i) face_flux(m, B, 2, order, k, j, i);
  }
}

/* Field::ComputeCornerE (src/field/calculate_corner_e.cpp:28-236) */
void ao_corner_e(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  double *e1 = B->e[0], *e2 = B->e[1], *e3 = B->e[2];
  double *w_x1f = B->wght[0], *w_x2f = B->wght[1], *w_x3f = B->wght[2];
  double *w = B->w, *bcc = B->bcc, *cc = B->cc_e;
  if (!m->f2) {
    for (int i = is; i <= ie+1; ++i) {
      e2[E2(B,ks,js,i)] = B->e2_x1f[F1(B,ks,js,i)];
      e2[E2(B,ke+1,js,i)] = B->e2_x1f[F1(B,ks,js,i)];
      e3[E3(B,ks,js,i)] = B->e3_x1f[F1(B,ks,js,i)];
      e3[E3(B,ks,je+1,i)] = B->e3_x1f[F1(B,ks,js,i)];
    }
    return;
  }
  if (!m->f3) {
    /* 2-D: cc_e_ is a 3-D array (k,j,i) */
    for (int k = ks; k <= ke; ++k) for (int j = js-1; j <= je+1; ++j)
      for (int i = is-1; i <= ie+1; ++i)
        cc[CC(B,0,k,j,i)] = w[CC(B,IVY,k,j,i)]*bcc[CC(B,IB1,k,j,i)]
                          - w[CC(B,IVX,k,j,i)]*bcc[CC(B,IB2,k,j,i)];
    for (int j = js; j <= je; ++j) for (int i = is; i <= ie+1; ++i)
      e2[E2(B,ke+1,j,i)] = e2[E2(B,ks,j,i)] = B->e2_x1f[F1(B,ks,j,i)];
    for (int j = js; j <= je+1; ++j) for (int i = is; i <= ie; ++i)
      e1[E1(B,ke+1,j,i)] = e1[E1(B,ks,j,i)] = B->e1_x2f[F2(B,ks,j,i)];
    for (int k = ks; k <= ke; ++k) for (int j = js; j <= je+1; ++j)
      for (int i = is; i <= ie+1; ++i) {
        const double *e3_x2f = B->e3_x2f, *e3_x1f = B->e3_x1f;
        double de3_l2 = (1.0-w_x1f[F1(B,k,j-1,i)])*(e3_x2f[F2(B,k,j,i)] - cc[CC(B,0,k,j-1,i)]) +
                        (    w_x1f[F1(B,k,j-1,i)])*(e3_x2f[F2(B,k,j,i-1)] - cc[CC(B,0,k,j-1,i-1)]);
        double de3_r2 = (1.0-w_x1f[F1(B,k,j,i)])*(e3_x2f[F2(B,k,j,i)] - cc[CC(B,0,k,j,i)]) +
                        (    w_x1f[F1(B,k,j,i)])*(e3_x2f[F2(B,k,j,i-1)] - cc[CC(B,0,k,j,i-1)]);
        double de3_l1 = (1.0-w_x2f[F2(B,k,j,i-1)])*(e3_x1f[F1(B,k,j,i)] - cc[CC(B,0,k,j,i-1)]) +
                        (    w_x2f[F2(B,k,j,i-1)])*(e3_x1f[F1(B,k,j-1,i)] - cc[CC(B,0,k,j-1,i-1)]);
        double de3_r1 = (1.0-w_x2f[F2(B,k,j,i)])*(e3_x1f[F1(B,k,j,i)] - cc[CC(B,0,k,j,i)]) +
                        (    w_x2f[F2(B,k,j,i)])*(e3_x1f[F1(B,k,j-1,i)] - cc[CC(B,0,k,j-1,i)]);
        e3[E3(B,k,j,i)] = 0.25*(de3_l1 + de3_r1 + de3_l2 + de3_r2 + e3_x2f[F2(B,k,j,i-1)] +
                                e3_x2f[F2(B,k,j,i)] + e3_x1f[F1(B,k,j-1,i)] + e3_x1f[F1(B,k,j,i)]);
      }
    return;
  }
  for (int k = ks-1; k <= ke+1; ++k) for (int j = js-1; j <= je+1; ++j)
    for (int i = is-1; i <= ie+1; ++i) {
      cc[CC(B,IB1,k,j,i)] = w[CC(B,IVZ,k,j,i)]*bcc[CC(B,IB2,k,j,i)] - w[CC(B,IVY,k,j,i)]*bcc[CC(B,IB3,k,j,i)];
      cc[CC(B,IB2,k,j,i)] = w[CC(B,IVX,k,j,i)]*bcc[CC(B,IB3,k,j,i)] - w[CC(B,IVZ,k,j,i)]*bcc[CC(B,IB1,k,j,i)];
      cc[CC(B,IB3,k,j,i)] = w[CC(B,IVY,k,j,i)]*bcc[CC(B,IB1,k,j,i)] - w[CC(B,IVX,k,j,i)]*bcc[CC(B,IB2,k,j,i)];
    }
  const double *e1_x2f = B->e1_x2f, *e1_x3f = B->e1_x3f, *e2_x1f = B->e2_x1f,
               *e2_x3f = B->e2_x3f, *e3_x1f = B->e3_x1f, *e3_x2f = B->e3_x2f;
  for (int k = ks; k <= ke+1; ++k) for (int j = js; j <= je+1; ++j)
    for (int i = is; i <= ie+1; ++i) {
      double de1_l3 = (1.0-w_x2f[F2(B,k-1,j,i)])*(e1_x3f[F3(B,k,j,i)] - cc[CC(B,IB1,k-1,j,i)]) +
                      (    w_x2f[F2(B,k-1,j,i)])*(e1_x3f[F3(B,k,j-1,i)] - cc[CC(B,IB1,k-1,j-1,i)]);
      double de1_r3 = (1.0-w_x2f[F2(B,k,j,i)])*(e1_x3f[F3(B,k,j,i)] - cc[CC(B,IB1,k,j,i)]) +
                      (    w_x2f[F2(B,k,j,i)])*(e1_x3f[F3(B,k,j-1,i)] - cc[CC(B,IB1,k,j-1,i)]);
      double de1_l2 = (1.0-w_x3f[F3(B,k,j-1,i)])*(e1_x2f[F2(B,k,j,i)] - cc[CC(B,IB1,k,j-1,i)]) +
                      (    w_x3f[F3(B,k,j-1,i)])*(e1_x2f[F2(B,k-1,j,i)] - cc[CC(B,IB1,k-1,j-1,i)]);
      double de1_r2 = (1.0-w_x3f[F3(B,k,j,i)])*(e1_x2f[F2(B,k,j,i)] - cc[CC(B,IB1,k,j,i)]) +
                      (    w_x3f[F3(B,k,j,i)])*(e1_x2f[F2(B,k-1,j,i)] - cc[CC(B,IB1,k-1,j,i)]);
      e1[E1(B,k,j,i)] = 0.25*(de1_l3 + de1_r3 + de1_l2 + de1_r2 + e1_x2f[F2(B,k-1,j,i)] +
                              e1_x2f[F2(B,k,j,i)] + e1_x3f[F3(B,k,j-1,i)] + e1_x3f[F3(B,k,j,i)]);

      double de2_l3 = (1.0-w_x1f[F1(B,k-1,j,i)])*(e2_x3f[F3(B,k,j,i)] - cc[CC(B,IB2,k-1,j,i)]) +
                      (    w_x1f[F1(B,k-1,j,i)])*(e2_x3f[F3(B,k,j,i-1)] - cc[CC(B,IB2,k-1,j,i-1)]);
      double de2_r3 = (1.0-w_x1f[F1(B,k,j,i)])*(e2_x3f[F3(B,k,j,i)] - cc[CC(B,IB2,k,j,i)]) +
                      (    w_x1f[F1(B,k,j,i)])*(e2_x3f[F3(B,k,j,i-1)] - cc[CC(B,IB2,k,j,i-1)]);
      double de2_l1 = (1.0-w_x3f[F3(B,k,j,i-1)])*(e2_x1f[F1(B,k,j,i)] - cc[CC(B,IB2,k,j,i-1)]) +
                      (    w_x3f[F3(B,k,j,i-1)])*(e2_x1f[F1(B,k-1,j,i)] - cc[CC(B,IB2,k-1,j,i-1)]);
      double de2_r1 = (1.0-w_x3f[F3(B,k,j,i)])*(e2_x1f[F1(B,k,j,i)] - cc[CC(B,IB2,k,j,i)]) +
                      (    w_x3f[F3(B,k,j,i)])*(e2_x1f[F1(B,k-1,j,i)] - cc[CC(B,IB2,k-1,j,i)]);
      e2[E2(B,k,j,i)] = 0.25*(de2_l3 + de2_r3 + de2_l1 + de2_r1 + e2_x3f[F3(B,k,j,i-1)] +
                              e2_x3f[F3(B,k,j,i)] + e2_x1f[F1(B,k-1,j,i)] + e2_x1f[F1(B,k,j,i)]);

      double de3_l2 = (1.0-w_x1f[F1(B,k,j-1,i)])*(e3_x2f[F2(B,k,j,i)] - cc[CC(B,IB3,k,j-1,i)]) +
                      (    w_x1f[F1(B,k,j-1,i)])*(e3_x2f[F2(B,k,j,i-1)] - cc[CC(B,IB3,k,j-1,i-1)]);
      double de3_r2 = (1.0-w_x1f[F1(B,k,j,i)])*(e3_x2f[F2(B,k,j,i)] - cc[CC(B,IB3,k,j,i)]) +
                      (    w_x1f[F1(B,k,j,i)])*(e3_x2f[F2(B,k,j,i-1)] - cc[CC(B,IB3,k,j,i-1)]);
      double de3_l1 = (1.0-w_x2f[F2(B,k,j,i-1)])*(e3_x1f[F1(B,k,j,i)] - cc[CC(B,IB3,k,j,i-1)]) +
                      (    w_x2f[F2(B,k,j,i-1)])*(e3_x1f[F1(B,k,j-1,i)] - cc[CC(B,IB3,k,j-1,i-1)]);
      double de3_r1 = (1.0-w_x2f[F2(B,k,j,i)])*(e3_x1f[F1(B,k,j,i)] - cc[CC(B,IB3,k,j,i)]) +
                      (    w_x2f[F2(B,k,j,i)])*(e3_x1f[F1(B,k,j-1,i)] - cc[CC(B,IB3,k,j-1,i)]);
      e3[E3(B,k,j,i)] = 0.25*(de3_l1 + de3_r1 + de3_l2 + de3_r2 + e3_x2f[F2(B,k,j,i-1)] +
                              e3_x2f[F2(B,k,j,i)] + e3_x1f[F1(B,k,j-1,i)] + e3_x1f[F1(B,k,j,i)]);
    }
}

/* ------------------------------------------------------------------ EMF correction */

/* LoadFluxBoundaryBufferSameLevel (src/bvals/fc/flux_correction_fc.cpp:51-307) */
static long emf_load(const AoMesh *m, const AoBlock *B, const Nb *nb, double *buf) {
  long p = 0;
  const double *e1 = B->e[0], *e2 = B->e[1], *e3 = B->e[2];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  if (nb->type == 0) {
    if (m->f3) {
      if (nb->fid == 0 || nb->fid == 1) {
        int i = nb->fid == 0 ? is : ie+1;
        for (int k = ks; k <= ke+1; k++) for (int j = js; j <= je; j++) buf[p++] = e2[E2(B,k,j,i)];
        for (int k = ks; k <= ke; k++) for (int j = js; j <= je+1; j++) buf[p++] = e3[E3(B,k,j,i)];
      } else if (nb->fid == 2 || nb->fid == 3) {
        int j = nb->fid == 2 ? js : je+1;
        for (int k = ks; k <= ke+1; k++) for (int i = is; i <= ie; i++) buf[p++] = e1[E1(B,k,j,i)];
        for (int k = ks; k <= ke; k++) for (int i = is; i <= ie+1; i++) buf[p++] = e3[E3(B,k,j,i)];
      } else {
        int k = nb->fid == 4 ? ks : ke+1;
        for (int j = js; j <= je+1; j++) for (int i = is; i <= ie; i++) buf[p++] = e1[E1(B,k,j,i)];
        for (int j = js; j <= je; j++) for (int i = is; i <= ie+1; i++) buf[p++] = e2[E2(B,k,j,i)];
      }
    } else if (m->f2) {
      int k = ks;
      if (nb->fid == 0 || nb->fid == 1) {
        int i = nb->fid == 0 ? is : ie+1;
        for (int j = js; j <= je; j++) buf[p++] = e2[E2(B,k,j,i)];
        for (int j = js; j <= je+1; j++) buf[p++] = e3[E3(B,k,j,i)];
      } else {
        int j = nb->fid == 2 ? js : je+1;
        for (int i = is; i <= ie; i++) buf[p++] = e1[E1(B,k,j,i)];
        for (int i = is; i <= ie+1; i++) buf[p++] = e3[E3(B,k,j,i)];
      }
    } else {
      int i = nb->fid == 0 ? is : ie+1;
      buf[p++] = e2[E2(B,ks,js,i)];
      buf[p++] = e3[E3(B,ks,js,i)];
    }
  } else if (nb->type == 1) {
    if (nb->eid >= 0 && nb->eid < 4) {
      int i = ((nb->eid & 1) == 0) ? is : ie+1;
      int j = ((nb->eid & 2) == 0) ? js : je+1;
      for (int k = ks; k <= ke; k++) buf[p++] = e3[E3(B,k,j,i)];
    } else if (nb->eid >= 4 && nb->eid < 8) {
      int i = ((nb->eid & 1) == 0) ? is : ie+1;
      int k = ((nb->eid & 2) == 0) ? ks : ke+1;
      for (int j = js; j <= je; j++) buf[p++] = e2[E2(B,k,j,i)];
    } else {
      int j = ((nb->eid & 1) == 0) ? js : je+1;
      int k = ((nb->eid & 2) == 0) ? ks : ke+1;
      for (int i = is; i <= ie; i++) buf[p++] = e1[E1(B,k,j,i)];
    }
  }
  return p;
}

/* SetFluxBoundarySameLevel (flux_correction_fc.cpp:689-905): ADDS the neighbour's EMFs */
static void emf_set(const AoMesh *m, AoBlock *B, const Nb *nb, const double *buf) {
  long p = 0;
  double *e1 = B->e[0], *e2 = B->e[1], *e3 = B->e[2];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  if (nb->type == 0) {
    if (m->f3) {
      if (nb->fid == 0 || nb->fid == 1) {
        int i = nb->fid == 0 ? is : ie+1;
        for (int k = ks; k <= ke+1; k++) for (int j = js; j <= je; j++) e2[E2(B,k,j,i)] += buf[p++];
        for (int k = ks; k <= ke; k++) for (int j = js; j <= je+1; j++) e3[E3(B,k,j,i)] += buf[p++];
      } else if (nb->fid == 2 || nb->fid == 3) {
        int j = nb->fid == 2 ? js : je+1;
        for (int k = ks; k <= ke+1; k++) for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] += buf[p++];
        for (int k = ks; k <= ke; k++) for (int i = is; i <= ie+1; i++) e3[E3(B,k,j,i)] += buf[p++];
      } else {
        int k = nb->fid == 4 ? ks : ke+1;
        for (int j = js; j <= je+1; j++) for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] += buf[p++];
        for (int j = js; j <= je; j++) for (int i = is; i <= ie+1; i++) e2[E2(B,k,j,i)] += buf[p++];
      }
    } else if (m->f2) {
      int k = ks;
      if (nb->fid == 0 || nb->fid == 1) {
        int i = nb->fid == 0 ? is : ie+1;
        for (int j = js; j <= je; j++) { e2[E2(B,k+1,j,i)] += buf[p]; e2[E2(B,k,j,i)] += buf[p++]; }
        for (int j = js; j <= je+1; j++) e3[E3(B,k,j,i)] += buf[p++];
      } else {
        int j = nb->fid == 2 ? js : je+1;
        for (int i = is; i <= ie; i++) { e1[E1(B,k+1,j,i)] += buf[p]; e1[E1(B,k,j,i)] += buf[p++]; }
        for (int i = is; i <= ie+1; i++) e3[E3(B,k,j,i)] += buf[p++];
      }
    } else {
      int i = nb->fid == 0 ? is : ie+1, j = js, k = ks;
      e2[E2(B,k+1,j,i)] += buf[p];
      e2[E2(B,k,j,i)] += buf[p++];
      e3[E3(B,k,j+1,i)] += buf[p];
      e3[E3(B,k,j,i)] += buf[p++];
    }
  } else if (nb->type == 1) {
    if (nb->eid >= 0 && nb->eid < 4) {
      int i = ((nb->eid & 1) == 0) ? is : ie+1;
      int j = ((nb->eid & 2) == 0) ? js : je+1;
      for (int k = ks; k <= ke; k++) e3[E3(B,k,j,i)] += buf[p++];
    } else if (nb->eid >= 4 && nb->eid < 8) {
      int i = ((nb->eid & 1) == 0) ? is : ie+1;
      int k = ((nb->eid & 2) == 0) ? ks : ke+1;
      for (int j = js; j <= je; j++) e2[E2(B,k,j,i)] += buf[p++];
    } else {
      int j = ((nb->eid & 1) == 0) ? js : je+1;
      int k = ((nb->eid & 2) == 0) ? ks : ke+1;
      for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] += buf[p++];
    }
  }
}

/* AverageFluxBoundary (flux_correction_fc.cpp:1355-1541), same-level branches */
static void emf_average(const AoMesh *m, AoBlock *B) {
  double *e1 = B->e[0], *e2 = B->e[1], *e3 = B->e[2];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  int nface = 2*m->ndim;
  for (int n = 0; n < nface; n++) {
    if (B->bcs[n] != -1 && B->bcs[n] != AO_BC_PERIODIC) continue;
    if (n == 0 || n == 1) {
      int i = n == 0 ? is : ie+1;
      double div = 0.5;
      if (m->f3) {
        for (int k = ks+1; k <= ke; k++) for (int j = js; j <= je; j++) e2[E2(B,k,j,i)] *= div;
        for (int k = ks; k <= ke; k++) for (int j = js+1; j <= je; j++) e3[E3(B,k,j,i)] *= div;
      } else if (m->f2) {
        for (int j = js; j <= je; j++) { e2[E2(B,ks,j,i)] *= div; e2[E2(B,ks+1,j,i)] *= div; }
        for (int j = js+1; j <= je; j++) e3[E3(B,ks,j,i)] *= div;
      } else {
        e2[E2(B,ks,js,i)] *= 0.5; e2[E2(B,ks+1,js,i)] *= 0.5;
        e3[E3(B,ks,js,i)] *= 0.5; e3[E3(B,ks,js+1,i)] *= 0.5;
      }
    }
    if (n == 2 || n == 3) {
      int j = n == 2 ? js : je+1;
      if (m->f3) {
        for (int k = ks+1; k <= ke; k++) for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] *= 0.5;
        for (int k = ks; k <= ke; k++) for (int i = is+1; i <= ie; i++) e3[E3(B,k,j,i)] *= 0.5;
      } else if (m->f2) {
        for (int i = is; i <= ie; i++) { e1[E1(B,ks,j,i)] *= 0.5; e1[E1(B,ks+1,j,i)] *= 0.5; }
        for (int i = is+1; i <= ie; i++) e3[E3(B,ks,j,i)] *= 0.5;
      }
    }
    if (n == 4 || n == 5) {
      int k = n == 4 ? ks : ke+1;
      for (int j = js+1; j <= je; j++) for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] *= 0.5;
      for (int j = js; j <= je; j++) for (int i = is+1; i <= ie; i++) e2[E2(B,k,j,i)] *= 0.5;
    }
  }
  int nedge = m->ndim == 3 ? 12 : (m->ndim == 2 ? 4 : 0);
  for (int n = 0; n < nedge; n++) {
    if (B->nedge_fine[n] == 1) continue;
    double div = 1.0/(double)B->nedge_fine[n];
    if (n < 4) {
      int i = ((n & 1) == 0) ? is : ie+1;
      int j = ((n & 2) == 0) ? js : je+1;
      for (int k = ks; k <= ke; k++) e3[E3(B,k,j,i)] *= div;
    } else if (n < 8) {
      int i = ((n & 1) == 0) ? is : ie+1;
      int k = ((n & 2) == 0) ? ks : ke+1;
      for (int j = js; j <= je; j++) e2[E2(B,k,j,i)] *= div;
    } else {
      int j = ((n & 1) == 0) ? js : je+1;
      int k = ((n & 2) == 0) ? ks : ke+1;
      for (int i = is; i <= ie; i++) e1[E1(B,k,j,i)] *= div;
    }
  }
}

/* SendFluxCorrection (flux_correction_fc.cpp:623-680) on every block, then
 * ReceiveFluxCorrection (:1610-1749): add in neighbour-list order, then average */
void ao_emf_exchange(AoMesh *m) {
  if (!m->p.mhd) return;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    long maxn = 2L*((long)B->nc1+1)*((long)B->nc2+1) + 2L*((long)B->nc1+1)*((long)B->nc3+1)
              + 2L*((long)B->nc2+1)*((long)B->nc3+1);
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      if (nb->type > 1) break;
      double *buf = dalloc(maxn);
      long cnt = emf_load(m, B, nb, buf);
      AoBlock *T = &m->blk[nb->gid];
      T->recv[nb->targetid] = buf; T->recvn[nb->targetid] = cnt;
    }
  }
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      if (nb->type > 1) break;
      emf_set(m, B, nb, B->recv[nb->bufid]);
      free(B->recv[nb->bufid]); B->recv[nb->bufid] = NULL;
    }
    emf_average(m, B);
  }
}

/* ------------------------------------------------------------------ integration */

static double *cc_reg(AoBlock *B, int r) { return r == 0 ? B->u : B->u1; }

/* MeshBlock::WeightedAve, cell-centred (src/mesh/weighted_ave.cpp:33-236), registers
 * u (0) / u1 (1); only wght[0..2] can be non-zero for the integrators restated here */
void ao_weighted_ave_cc(AoMesh *m, int b, int out_reg, int in1_reg, const double w[5]) {
  AoBlock *B = &m->blk[b];
  double *uo = cc_reg(B, out_reg), *ui = cc_reg(B, in1_reg);
  for (int n = 0; n < NHYDRO; ++n) for (int k = B->ks; k <= B->ke; ++k)
    for (int j = B->js; j <= B->je; ++j) for (int i = B->is; i <= B->ie; ++i) {
      long o = CC(B,n,k,j,i);
      if (w[0] == 1.0) {
        if (w[1] != 0.0) uo[o] += w[1]*ui[o];
      } else if (w[0] == 0.0) {
        if (w[1] == 1.0) uo[o] = ui[o];
        else uo[o] = w[1]*ui[o];
      } else {
        if (w[1] != 0.0) uo[o] = w[0]*uo[o] + w[1]*ui[o];
        else uo[o] *= w[0];
      }
    }
}

static void wave_fc_one(double *bo, const double *bi, long o, const double w[5]) {
  if (w[0] == 1.0) {
    if (w[1] != 0.0) bo[o] += w[1]*bi[o];
  } else if (w[0] == 0.0) {
    if (w[1] == 1.0) bo[o] = bi[o];
    else bo[o] = w[1]*bi[o];
  } else {
    if (w[1] != 0.0) bo[o] = w[0]*bo[o] + w[1]*bi[o];
    else bo[o] *= w[0];
  }
}

/* MeshBlock::WeightedAve, face-centred (weighted_ave.cpp:238-...) */
void ao_weighted_ave_fc(AoMesh *m, int b, int out_reg, int in1_reg, const double w[5]) {
  AoBlock *B = &m->blk[b];
  double **bo = out_reg == 0 ? B->b : B->b1, **bi = in1_reg == 0 ? B->b : B->b1;
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie+1; ++i)
    wave_fc_one(bo[0], bi[0], F1(B,k,j,i), w);
  for (int k = ks; k <= ke; ++k) for (int j = js; j <= je+1; ++j) for (int i = is; i <= ie; ++i)
    wave_fc_one(bo[1], bi[1], F2(B,k,j,i), w);
  for (int k = ks; k <= ke+1; ++k) for (int j = js; j <= je; ++j) for (int i = is; i <= ie; ++i)
    wave_fc_one(bo[2], bi[2], F3(B,k,j,i), w);
  (void)m;
}

void ao_swap_cc(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b]; double *t = B->u; B->u = B->u1; B->u1 = t;
}
void ao_swap_fc(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  for (int d = 0; d < 3; ++d) { double *t = B->b[d]; B->b[d] = B->b1[d]; B->b1[d] = t; }
}
/* StartupTaskList stage 1 (time_integrator.cpp:1386-1397) */
void ao_zero_reg1(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  memset(B->u1, 0, sizeof(double)*NHYDRO*(size_t)B->nc1*B->nc2*B->nc3);
  if (m->p.nscalars > 0)
    memset(B->s1, 0, sizeof(double)*(size_t)m->p.nscalars*B->nc1*B->nc2*B->nc3);
  if (m->p.mhd) {
    memset(B->b1[0], 0, sizeof(double)*(size_t)B->nc3*B->nc2*(B->nc1+1));
    memset(B->b1[1], 0, sizeof(double)*(size_t)B->nc3*(B->nc2+1)*B->nc1);
    memset(B->b1[2], 0, sizeof(double)*(size_t)(B->nc3+1)*B->nc2*B->nc1);
  }
}

/* Hydro::AddFluxDivergence (src/hydro/add_flux_divergence.cpp:39-96); Cartesian areas and
 * volume (src/coordinates/coordinates.cpp:436-533) */
void ao_add_flux_div(AoMesh *m, int b, double wght) {
  AoBlock *B = &m->blk[b];
  for (int k = B->ks; k <= B->ke; ++k) for (int j = B->js; j <= B->je; ++j)
    for (int n = 0; n < NHYDRO; ++n) for (int i = B->is; i <= B->ie; ++i) {
      double x1area = B->dx2f[j]*B->dx3f[k];
      double dflx = (x1area*B->flux[0][FL1(B,n,k,j,i+1)] - x1area*B->flux[0][FL1(B,n,k,j,i)]);
      if (m->f2) {
        double x2area = B->dx1f[i]*B->dx3f[k];
        dflx += (x2area*B->flux[1][FL2(B,n,k,j+1,i)] - x2area*B->flux[1][FL2(B,n,k,j,i)]);
      }
      if (m->f3) {
        double x3area = B->dx1f[i]*B->dx2f[j];
        dflx += (x3area*B->flux[2][FL3(B,n,k+1,j,i)] - x3area*B->flux[2][FL3(B,n,k,j,i)]);
      }
      double vol = B->dx1f[i]*B->dx2f[j]*B->dx3f[k];
      B->u[CC(B,n,k,j,i)] -= wght*dflx/vol;
    }
}

/* Field::CT (src/field/ct.cpp:31-116) */
void ao_ct(AoMesh *m, int b, double wght) {
  AoBlock *B = &m->blk[b];
  const double *e1 = B->e[0], *e2 = B->e[1], *e3 = B->e[2];
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  for (int k = ks; k <= ke; ++k) for (int j = js; j <= je; ++j) {
    if (m->f2) {
      for (int i = is; i <= ie+1; ++i) {
        double area = B->dx2f[j]*B->dx3f[k];
        double len = B->dx3f[k], len_p1 = B->dx3f[k];
        B->b[0][F1(B,k,j,i)] -= (wght/area)*(len_p1*e3[E3(B,k,j+1,i)] - len*e3[E3(B,k,j,i)]);
      }
      if (m->f3) {
        for (int i = is; i <= ie+1; ++i) {
          double area = B->dx2f[j]*B->dx3f[k];
          double len = B->dx2f[j], len_p1 = B->dx2f[j];
          B->b[0][F1(B,k,j,i)] += (wght/area)*(len_p1*e2[E2(B,k+1,j,i)] - len*e2[E2(B,k,j,i)]);
        }
      }
    }
  }
  for (int k = ks; k <= ke; ++k) for (int j = js; j <= je+1; ++j) {
    for (int i = is; i <= ie; ++i) {
      double area = B->dx1f[i]*B->dx3f[k];
      double len = B->dx3f[k];
      B->b[1][F2(B,k,j,i)] += (wght/area)*(len*e3[E3(B,k,j,i+1)] - len*e3[E3(B,k,j,i)]);
    }
    if (m->f3) {
      for (int i = is; i <= ie; ++i) {
        double area = B->dx1f[i]*B->dx3f[k];
        double len = B->dx1f[i], len_p1 = B->dx1f[i];
        B->b[1][F2(B,k,j,i)] -= (wght/area)*(len_p1*e1[E1(B,k+1,j,i)] - len*e1[E1(B,k,j,i)]);
      }
    }
  }
  for (int k = ks; k <= ke+1; ++k) for (int j = js; j <= je; ++j) {
    for (int i = is; i <= ie; ++i) {
      double area = B->dx1f[i]*B->dx2f[j];
      double len = B->dx2f[j];
      B->b[2][F3(B,k,j,i)] -= (wght/area)*(len*e2[E2(B,k,j,i+1)] - len*e2[E2(B,k,j,i)]);
    }
    if (m->f2) {
      for (int i = is; i <= ie; ++i) {
        double area = B->dx1f[i]*B->dx2f[j];
        double len = B->dx1f[i], len_p1 = B->dx1f[i];
        B->b[2][F3(B,k,j,i)] += (wght/area)*(len_p1*e1[E1(B,k,j+1,i)] - len*e1[E1(B,k,j,i)]);
      }
    }
  }
}

/* ------------------------------------------------------------------ ghost exchange */

/* BufferUtility::PackData / UnpackData order n,k,j,i (src/utils/buffer_utils.cpp) */
static long pack3(const double *a, long s2, long s1, int si, int ei, int sj, int ej, int sk,
                  int ek, double *buf, long p) {
  for (int k = sk; k <= ek; ++k) for (int j = sj; j <= ej; ++j) for (int i = si; i <= ei; ++i)
    buf[p++] = a[((long)k*s2 + j)*s1 + i];
  return p;
}
static long unpack3(double *a, long s2, long s1, int si, int ei, int sj, int ej, int sk,
                    int ek, const double *buf, long p) {
  for (int k = sk; k <= ek; ++k) for (int j = sj; j <= ej; ++j) for (int i = si; i <= ei; ++i)
    a[((long)k*s2 + j)*s1 + i] = buf[p++];
  return p;
}

/* CellCenteredBoundaryVariable::LoadBoundaryBufferSameLevel / SetBoundarySameLevel
 * (src/bvals/cc/bvals_cc.cpp:201-216,300-336) */
static void exchange_cc_var(AoMesh *m, int nvar, int scalars) {
  int ng = m->p.ng;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      int si = (nb->ox1 > 0) ? (B->ie - ng + 1) : B->is, ei = (nb->ox1 < 0) ? (B->is + ng - 1) : B->ie;
      int sj = (nb->ox2 > 0) ? (B->je - ng + 1) : B->js, ej = (nb->ox2 < 0) ? (B->js + ng - 1) : B->je;
      int sk = (nb->ox3 > 0) ? (B->ke - ng + 1) : B->ks, ek = (nb->ox3 < 0) ? (B->ks + ng - 1) : B->ke;
      long cnt = (long)nvar*(ei-si+1)*(ej-sj+1)*(ek-sk+1);
      double *buf = dalloc(cnt);
      long p = 0;
      const double *src = scalars ? B->s : B->u;
      for (int v = 0; v < nvar; ++v)
        p = pack3(src + (long)v*B->nc3*B->nc2*B->nc1, B->nc2, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      AoBlock *T = &m->blk[nb->gid];
      T->recv[nb->targetid] = buf; T->recvn[nb->targetid] = cnt;
    }
  }
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      int si, ei, sj, ej, sk, ek;
      if (nb->ox1 == 0) { si = B->is; ei = B->ie; }
      else if (nb->ox1 > 0) { si = B->ie + 1; ei = B->ie + ng; }
      else { si = B->is - ng; ei = B->is - 1; }
      if (nb->ox2 == 0) { sj = B->js; ej = B->je; }
      else if (nb->ox2 > 0) { sj = B->je + 1; ej = B->je + ng; }
      else { sj = B->js - ng; ej = B->js - 1; }
      if (nb->ox3 == 0) { sk = B->ks; ek = B->ke; }
      else if (nb->ox3 > 0) { sk = B->ke + 1; ek = B->ke + ng; }
      else { sk = B->ks - ng; ek = B->ks - 1; }
      const double *buf = B->recv[nb->bufid];
      long p = 0;
      double *dst = scalars ? B->s : B->u;
      for (int v = 0; v < nvar; ++v)
        p = unpack3(dst + (long)v*B->nc3*B->nc2*B->nc1, B->nc2, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      free(B->recv[nb->bufid]); B->recv[nb->bufid] = NULL;
    }
  }
}

void ao_exchange_cc(AoMesh *m) { exchange_cc_var(m, NHYDRO, 0); }
/* PassiveScalars::sbvar is a CellCenteredBoundaryVariable on s (scalars.cpp:59-72): same boxes */
void ao_exchange_scalars(AoMesh *m) { if (m->p.nscalars > 0) exchange_cc_var(m, m->p.nscalars, 1); }

/* FaceCenteredBoundaryVariable::LoadBoundaryBufferSameLevel / SetBoundarySameLevel
 * (src/bvals/fc/bvals_fc.cpp:344-397,583-684), uniform (non-multilevel) mesh */
void ao_exchange_fc(AoMesh *m) {
  if (!m->p.mhd) return;
  int ng = m->p.ng;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      int si, ei, sj, ej, sk, ek;
      double *buf = dalloc(3L*(B->nc1+1)*(B->nc2+1)*(B->nc3+1));
      long p = 0;
      /* bx1 */
      if (nb->ox1 == 0) { si = is; ei = ie + 1; }
      else if (nb->ox1 > 0) { si = ie - ng + 1; ei = ie; }
      else { si = is + 1; ei = is + ng; }
      if (nb->ox2 == 0) { sj = js; ej = je; }
      else if (nb->ox2 > 0) { sj = je - ng + 1; ej = je; }
      else { sj = js; ej = js + ng - 1; }
      if (nb->ox3 == 0) { sk = ks; ek = ke; }
      else if (nb->ox3 > 0) { sk = ke - ng + 1; ek = ke; }
      else { sk = ks; ek = ks + ng - 1; }
      p = pack3(B->b[0], B->nc2, B->nc1+1, si, ei, sj, ej, sk, ek, buf, p);
      /* bx2 */
      if (nb->ox1 == 0) { si = is; ei = ie; }
      else if (nb->ox1 > 0) { si = ie - ng + 1; ei = ie; }
      else { si = is; ei = is + ng - 1; }
      if (!m->f2) { sj = js; ej = je; }
      else if (nb->ox2 == 0) { sj = js; ej = je + 1; }
      else if (nb->ox2 > 0) { sj = je - ng + 1; ej = je; }
      else { sj = js + 1; ej = js + ng; }
      p = pack3(B->b[1], B->nc2+1, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      /* bx3 */
      if (nb->ox2 == 0) { sj = js; ej = je; }
      else if (nb->ox2 > 0) { sj = je - ng + 1; ej = je; }
      else { sj = js; ej = js + ng - 1; }
      if (!m->f3) { sk = ks; ek = ke; }
      else if (nb->ox3 == 0) { sk = ks; ek = ke + 1; }
      else if (nb->ox3 > 0) { sk = ke - ng + 1; ek = ke; }
      else { sk = ks + 1; ek = ks + ng; }
      p = pack3(B->b[2], B->nc2, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      AoBlock *T = &m->blk[nb->gid];
      T->recv[nb->targetid] = buf; T->recvn[nb->targetid] = p;
    }
  }
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      int si, ei, sj, ej, sk, ek;
      const double *buf = B->recv[nb->bufid];
      long p = 0;
      /* bx1: the shared face itself is never overwritten (si = ie+2) */
      if (nb->ox1 == 0) { si = is; ei = ie + 1; }
      else if (nb->ox1 > 0) { si = ie + 2; ei = ie + ng + 1; }
      else { si = is - ng; ei = is - 1; }
      if (nb->ox2 == 0) { sj = js; ej = je; }
      else if (nb->ox2 > 0) { sj = je + 1; ej = je + ng; }
      else { sj = js - ng; ej = js - 1; }
      if (nb->ox3 == 0) { sk = ks; ek = ke; }
      else if (nb->ox3 > 0) { sk = ke + 1; ek = ke + ng; }
      else { sk = ks - ng; ek = ks - 1; }
      p = unpack3(B->b[0], B->nc2, B->nc1+1, si, ei, sj, ej, sk, ek, buf, p);
      /* bx2 */
      if (nb->ox1 == 0) { si = is; ei = ie; }
      else if (nb->ox1 > 0) { si = ie + 1; ei = ie + ng; }
      else { si = is - ng; ei = is - 1; }
      if (!m->f2) { sj = js; ej = je; }
      else if (nb->ox2 == 0) { sj = js; ej = je + 1; }
      else if (nb->ox2 > 0) { sj = je + 2; ej = je + ng + 1; }
      else { sj = js - ng; ej = js - 1; }
      p = unpack3(B->b[1], B->nc2+1, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      if (!m->f2) {
        for (int i = si; i <= ei; ++i) B->b[1][F2(B,sk,sj+1,i)] = B->b[1][F2(B,sk,sj,i)];
      }
      /* bx3 */
      if (nb->ox2 == 0) { sj = js; ej = je; }
      else if (nb->ox2 > 0) { sj = je + 1; ej = je + ng; }
      else { sj = js - ng; ej = js - 1; }
      if (!m->f3) { sk = ks; ek = ke; }
      else if (nb->ox3 == 0) { sk = ks; ek = ke + 1; }
      else if (nb->ox3 > 0) { sk = ke + 2; ek = ke + ng + 1; }
      else { sk = ks - ng; ek = ks - 1; }
      p = unpack3(B->b[2], B->nc2, B->nc1, si, ei, sj, ej, sk, ek, buf, p);
      if (!m->f3) {
        for (int j = sj; j <= ej; ++j) for (int i = si; i <= ei; ++i)
          B->b[2][F3(B,sk+1,j,i)] = B->b[2][F3(B,sk,j,i)];
      }
      free(B->recv[nb->bufid]); B->recv[nb->bufid] = NULL;
    }
  }
}

/* ------------------------------------------------------------------ physical BCs */

/* outflow / reflecting physical boundary on primitives + face fields
 * (src/bvals/cc/outflow_cc.cpp, cc/hydro/reflect_hydro.cpp, fc/outflow_fc.cpp, fc/reflect_fc.cpp).
 * lo/hi = active range along the boundary's own direction; ghost layer g copies from
 * outflow: the last active cell / face; reflect: the mirrored one, with the normal velocity
 * and the normal field negated. */
static void phys_bc(AoMesh *m, AoBlock *B, int face, int refl, int il, int iu, int jl, int ju,
                    int kl, int ku) {
  int ng = m->p.ng, mhd = m->p.mhd;
  int d = face/2, upper = face & 1;
  int lo = d == 0 ? il : (d == 1 ? jl : kl), hi = d == 0 ? iu : (d == 1 ? ju : ku);
  /* loop bounds of the two transverse directions (a: slower, c: faster index) */
  for (int g = 1; g <= ng; ++g) {
    int gc = upper ? hi + g : lo - g;                       /* ghost cell / transverse face */
    int sc = refl ? (upper ? hi - g + 1 : lo + g - 1) : (upper ? hi : lo);
    int gn = upper ? hi + g + 1 : lo - g;                   /* ghost normal face */
    int sn = refl ? (upper ? hi - g + 1 : lo + g) : (upper ? hi + 1 : lo);
    double sgn_n = refl ? -1.0 : 1.0;
    if (d == 0) {
      for (int n = 0; n < NHYDRO; ++n) for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) {
        double v = B->w[CC(B,n,k,j,sc)];
        B->w[CC(B,n,k,j,gc)] = (refl && n == IVX) ? -v : v;
      }
      if (!mhd) continue;
      for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j)
        B->b[0][F1(B,k,j,gn)] = sgn_n*B->b[0][F1(B,k,j,sn)];
      for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju+1; ++j)
        B->b[1][F2(B,k,j,gc)] = B->b[1][F2(B,k,j,sc)];
      for (int k = kl; k <= ku+1; ++k) for (int j = jl; j <= ju; ++j)
        B->b[2][F3(B,k,j,gc)] = B->b[2][F3(B,k,j,sc)];
    } else if (d == 1) {
      for (int n = 0; n < NHYDRO; ++n) for (int k = kl; k <= ku; ++k) for (int i = il; i <= iu; ++i) {
        double v = B->w[CC(B,n,k,sc,i)];
        B->w[CC(B,n,k,gc,i)] = (refl && n == IVY) ? -v : v;
      }
      if (!mhd) continue;
      for (int k = kl; k <= ku; ++k) for (int i = il; i <= iu+1; ++i)
        B->b[0][F1(B,k,gc,i)] = B->b[0][F1(B,k,sc,i)];
      for (int k = kl; k <= ku; ++k) for (int i = il; i <= iu; ++i)
        B->b[1][F2(B,k,gn,i)] = sgn_n*B->b[1][F2(B,k,sn,i)];
      for (int k = kl; k <= ku+1; ++k) for (int i = il; i <= iu; ++i)
        B->b[2][F3(B,k,gc,i)] = B->b[2][F3(B,k,sc,i)];
    } else {
      for (int n = 0; n < NHYDRO; ++n) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
        double v = B->w[CC(B,n,sc,j,i)];
        B->w[CC(B,n,gc,j,i)] = (refl && n == IVZ) ? -v : v;
      }
      if (!mhd) continue;
      for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu+1; ++i)
        B->b[0][F1(B,gc,j,i)] = B->b[0][F1(B,sc,j,i)];
      for (int j = jl; j <= ju+1; ++j) for (int i = il; i <= iu; ++i)
        B->b[1][F2(B,gc,j,i)] = B->b[1][F2(B,sc,j,i)];
      for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
        B->b[2][F3(B,gn,j,i)] = sgn_n*B->b[2][F3(B,sn,j,i)];
    }
  }
}

/* CellCenteredBoundaryVariable outflow / reflect on the primitive scalars r
 * (src/bvals/cc/outflow_cc.cpp, reflect_cc.cpp: a plain copy / mirror, no sign change) */
static void phys_bc_scalars(AoMesh *m, AoBlock *B, int face, int refl, int il, int iu, int jl,
                            int ju, int kl, int ku) {
  int ng = m->p.ng;
  int d = face/2, upper = face & 1;
  int lo = d == 0 ? il : (d == 1 ? jl : kl), hi = d == 0 ? iu : (d == 1 ? ju : ku);
  for (int g = 1; g <= ng; ++g) {
    int gc = upper ? hi + g : lo - g;
    int sc = refl ? (upper ? hi - g + 1 : lo + g - 1) : (upper ? hi : lo);
    for (int n = 0; n < m->p.nscalars; ++n) {
      if (d == 0) {
        for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j)
          B->r[CC(B,n,k,j,gc)] = B->r[CC(B,n,k,j,sc)];
      } else if (d == 1) {
        for (int k = kl; k <= ku; ++k) for (int i = il; i <= iu; ++i)
          B->r[CC(B,n,k,gc,i)] = B->r[CC(B,n,k,sc,i)];
      } else {
        for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i)
          B->r[CC(B,n,gc,j,i)] = B->r[CC(B,n,sc,j,i)];
      }
    }
  }
}

/* one face of ApplyPhysicalBoundaries: boundary function on w, b (and r), then bcc and the
 * conserved variables of the ghost slab [g*] (bvals.cpp:476-497) */
static void bc_face(AoMesh *m, int b, int face, int il, int iu, int jl, int ju, int kl, int ku,
                    int gil, int giu, int gjl, int gju, int gkl, int gku) {
  AoBlock *B = &m->blk[b];
  int refl = B->bcs[face] == AO_BC_REFLECT;
  if (B->bcs[face] == AO_BC_USER) {   /* DispatchBoundaryFunctions (bvals.cpp:617-620) */
    if (m->user_bc[face])
      m->user_bc[face](m->user_bc_arg[face], b, B->w, B->b[0], B->b[1], B->b[2], m->bc_time,
                       m->bc_dt, il, iu, jl, ju, kl, ku, m->p.ng);
  } else {
    phys_bc(m, B, face, refl, il, iu, jl, ju, kl, ku);
    if (m->p.nscalars > 0) phys_bc_scalars(m, B, face, refl, il, iu, jl, ju, kl, ku);
  }
  if (m->p.mhd) calc_bcc(B, gil, giu, gjl, gju, gkl, gku);
  ao_prim2cons(m, b, gil, giu, gjl, gju, gkl, gku);
  if (m->p.nscalars > 0) ao_scalar_prim2cons(m, b, gil, giu, gjl, gju, gkl, gku);
}

/* BoundaryValues::ApplyPhysicalBoundaries (src/bvals/bvals.cpp:436-620) */
void ao_physical_bcs(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  int ng = m->p.ng;
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  int bis = is - ng, bie = ie + ng, bjs = js, bje = je, bks = ks, bke = ke;
  int app[6];
  for (int f = 0; f < 6; ++f) app[f] = (B->bcs[f] != -1 && B->bcs[f] != AO_BC_PERIODIC);
  if (!app[2] && m->f2) bjs = js - ng;
  if (!app[3] && m->f2) bje = je + ng;
  if (!app[4] && m->f3) bks = ks - ng;
  if (!app[5] && m->f3) bke = ke + ng;
  if (app[0]) bc_face(m, b, 0, is, ie, bjs, bje, bks, bke, is-ng, is-1, bjs, bje, bks, bke);
  if (app[1]) bc_face(m, b, 1, is, ie, bjs, bje, bks, bke, ie+1, ie+ng, bjs, bje, bks, bke);
  if (m->f2) {
    if (app[2]) bc_face(m, b, 2, bis, bie, js, je, bks, bke, bis, bie, js-ng, js-1, bks, bke);
    if (app[3]) bc_face(m, b, 3, bis, bie, js, je, bks, bke, bis, bie, je+1, je+ng, bks, bke);
  }
  if (m->f3) {
    bjs = js - ng; bje = je + ng;
    if (app[4]) bc_face(m, b, 4, bis, bie, bjs, bje, ks, ke, bis, bie, bjs, bje, ks-ng, ks-1);
    if (app[5]) bc_face(m, b, 5, bis, bie, bjs, bje, ks, ke, bis, bie, bjs, bje, ke+1, ke+ng);
  }
}

/* ------------------------------------------------------------------ passive scalars */

/* both face states of r(n) in cell (k,j,i) along dir (dc_simple.cpp, plm_simple.cpp,
 * ppm_simple.cpp: the same limiters as the hydro variables, no characteristic projection);
 * xorder 3 re-applies the concentration floor (calculate_scalar_fluxes.cpp:83-92) */
static void recon_scalar(const AoMesh *m, const AoBlock *B, int dir, int order, int n, int k,
                         int j, int i, double *plus, double *minus) {
  int dk = (dir == 2), dj = (dir == 1), di = (dir == 0);
  const double *r = B->r;
  double q = r[CC(B,n,k,j,i)];
  if (order == 1) { *plus = *minus = q; return; }
  double qm1 = r[CC(B,n,k-dk,j-dj,i-di)], qp1 = r[CC(B,n,k+dk,j+dj,i+di)];
  const AoReconGeom *rg = &B->rg[dir][dir == 0 ? i : (dir == 1 ? j : k)];
  if (order == 2) {
    ao_plm_point_g(qm1, q, qp1, rg, plus, minus);
    return;
  }
  double qm2 = r[CC(B,n,k-2*dk,j-2*dj,i-2*di)], qp2 = r[CC(B,n,k+2*dk,j+2*dj,i+2*di)];
  ao_ppm_point_g(qm2, qm1, q, qp1, qp2, rg, plus, minus);
  /* EquationOfState::ApplyPassiveScalarFloors (eos_scalars.cpp:161-175) */
  *plus = (*plus > m->p.sfloor) ? *plus : m->p.sfloor;
  *minus = (*minus > m->p.sfloor) ? *minus : m->p.sfloor;
}

/* PassiveScalars::CalculateFluxes + ComputeUpwindFlux
 * (src/scalars/calculate_scalar_fluxes.cpp:41-382): upwind the reconstructed concentration
 * with the Riemann solver's mass flux.  The reference sweeps a widened transverse range
 * (:64-71,159-166,261); only the faces AddFluxDivergence reads are evaluated here (the extra
 * rows use mass fluxes that hydro runs never compute and are never read). */
void ao_calc_scalar_fluxes(AoMesh *m, int b, int order) {
  AoBlock *B = &m->blk[b];
  if (m->p.nscalars <= 0) return;
  int is = B->is, ie = B->ie, js = B->js, je = B->je, ks = B->ks, ke = B->ke;
  for (int dir = 0; dir < m->ndim; ++dir) {
    int dk = (dir == 2), dj = (dir == 1), di = (dir == 0);
    for (int n = 0; n < m->p.nscalars; ++n)
      for (int k = ks; k <= ke + dk; ++k) for (int j = js; j <= je + dj; ++j)
        for (int i = is; i <= ie + di; ++i) {
          double rl, rr, tmp;
          recon_scalar(m, B, dir, order, n, k-dk, j-dj, i-di, &rl, &tmp);
          recon_scalar(m, B, dir, order, n, k, j, i, &tmp, &rr);
          long fo = dir == 0 ? FL1(B,IDN,k,j,i) : (dir == 1 ? FL2(B,IDN,k,j,i) : FL3(B,IDN,k,j,i));
          long so = dir == 0 ? FL1(B,n,k,j,i) : (dir == 1 ? FL2(B,n,k,j,i) : FL3(B,n,k,j,i));
          double fluid_flx = B->flux[dir][fo];
          if (fluid_flx >= 0.0) B->sflux[dir][so] = fluid_flx*rl;
          else B->sflux[dir][so] = fluid_flx*rr;
        }
  }
}

/* TimeIntegratorTaskList::IntegrateScalars (src/task_list/time_integrator.cpp:2141-2185):
 * WeightedAve on (s1, s), swap or second average, PassiveScalars::AddFluxDivergence
 * (src/scalars/add_scalar_flux_divergence.cpp:43-97) */
void ao_integrate_scalars(AoMesh *m, int b, int stage) {
  AoBlock *B = &m->blk[b];
  int ns = m->p.nscalars;
  if (ns <= 0) return;
  int s = stage - 1;
  double w[5] = {1.0, m->delta[s], 0.0, 0.0, 0.0};
  double w2[5] = {m->g1[s], m->g2[s], m->g3[s], 0.0, 0.0};
  for (int pass = 0; pass < 2; ++pass) {
    double *uo = pass == 0 ? B->s1 : B->s, *ui = pass == 0 ? B->s : B->s1;
    const double *ww = pass == 0 ? w : w2;
    if (pass == 1 && w2[0] == 0.0 && w2[1] == 1.0 && w2[2] == 0.0) {
      double *t = B->s; B->s = B->s1; B->s1 = t;
      break;
    }
    for (int n = 0; n < ns; ++n) for (int k = B->ks; k <= B->ke; ++k)
      for (int j = B->js; j <= B->je; ++j) for (int i = B->is; i <= B->ie; ++i) {
        long o = CC(B,n,k,j,i);
        if (ww[0] == 1.0) {
          if (ww[1] != 0.0) uo[o] += ww[1]*ui[o];
        } else if (ww[0] == 0.0) {
          if (ww[1] == 1.0) uo[o] = ui[o];
          else uo[o] = ww[1]*ui[o];
        } else {
          if (ww[1] != 0.0) uo[o] = ww[0]*uo[o] + ww[1]*ui[o];
          else uo[o] *= ww[0];
        }
      }
  }
  double wght = m->beta[s]*m->dt;
  for (int k = B->ks; k <= B->ke; ++k) for (int j = B->js; j <= B->je; ++j)
    for (int n = 0; n < ns; ++n) for (int i = B->is; i <= B->ie; ++i) {
      double x1area = B->dx2f[j]*B->dx3f[k];
      double dflx = (x1area*B->sflux[0][FL1(B,n,k,j,i+1)] - x1area*B->sflux[0][FL1(B,n,k,j,i)]);
      if (m->f2) {
        double x2area = B->dx1f[i]*B->dx3f[k];
        dflx += (x2area*B->sflux[1][FL2(B,n,k,j+1,i)] - x2area*B->sflux[1][FL2(B,n,k,j,i)]);
      }
      if (m->f3) {
        double x3area = B->dx1f[i]*B->dx2f[j];
        dflx += (x3area*B->sflux[2][FL3(B,n,k+1,j,i)] - x3area*B->sflux[2][FL3(B,n,k,j,i)]);
      }
      double vol = B->dx1f[i]*B->dx2f[j]*B->dx3f[k];
      B->s[CC(B,n,k,j,i)] -= wght*dflx/vol;
    }
}

/* HydroSourceTerms::ConstantAcceleration (src/hydro/srcterms/constant_acc.cpp:25-77), called
 * from the SRC_TERM task with dt = beta*dt on the stage's u, using the stage-start primitives
 * (time_integrator.cpp:1655-1678) */
static void const_accel(AoMesh *m, int b, double dt) {
  AoBlock *B = &m->blk[b];
  for (int d = 0; d < 3; ++d) {
    double g = m->p.grav_acc[d];
    if (g == 0.0) continue;
    for (int k = B->ks; k <= B->ke; ++k) for (int j = B->js; j <= B->je; ++j)
      for (int i = B->is; i <= B->ie; ++i) {
        double src = dt*B->w[CC(B,IDN,k,j,i)]*g;
        B->u[CC(B,IM1+d,k,j,i)] += src;
        if (!ISO(m)) B->u[CC(B,IEN,k,j,i)] += src*B->w[CC(B,IVX+d,k,j,i)];
      }
  }
}

/* HydroSourceTerms::AddSourceTerms (hydro/srcterms/hydro_srcterms.cpp:117-156) */
void ao_add_source_terms(AoMesh *m, int b, double time, double dt) {
  const_accel(m, b, dt);
  if (m->user_src)
    m->user_src(m->user_src_arg, b, time, dt, m->blk[b].w, m->blk[b].r, m->blk[b].bcc,
                m->blk[b].u, m->blk[b].s);
}

/* ------------------------------------------------------------------ history */

/* HistoryOutput::WriteOutputFile (src/outputs/history.cpp:69-169): volume-weighted sums over
 * the active cells, MeshBlocks in list order, one running accumulator per quantity */
int ao_history(AoMesh *m, double *out) {
  int nme = m->p.mhd ? 3 : 0;
  int nout = NHYDRO + 3 + nme + m->p.nscalars;
  for (int n = 0; n < nout; ++n) out[n] = 0.0;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    for (int k = B->ks; k <= B->ke; ++k) for (int j = B->js; j <= B->je; ++j)
      for (int i = B->is; i <= B->ie; ++i) {
        double vol = B->dx1f[i]*B->dx2f[j]*B->dx3f[k];
        double u_d = B->u[CC(B,IDN,k,j,i)], u_mx = B->u[CC(B,IM1,k,j,i)],
               u_my = B->u[CC(B,IM2,k,j,i)], u_mz = B->u[CC(B,IM3,k,j,i)];
        out[0] += vol*u_d;
        out[1] += vol*u_mx;
        out[2] += vol*u_my;
        out[3] += vol*u_mz;
        out[4] += vol*0.5*SQR(u_mx)/u_d;
        out[5] += vol*0.5*SQR(u_my)/u_d;
        out[6] += vol*0.5*SQR(u_mz)/u_d;
        if (!ISO(m)) out[7] += vol*B->u[CC(B,IEN,k,j,i)];
        if (m->p.mhd) {
          double bcc1 = B->bcc[CC(B,IB1,k,j,i)], bcc2 = B->bcc[CC(B,IB2,k,j,i)],
                 bcc3 = B->bcc[CC(B,IB3,k,j,i)];
          out[NHYDRO + 3] += vol*0.5*bcc1*bcc1;
          out[NHYDRO + 4] += vol*0.5*bcc2*bcc2;
          out[NHYDRO + 5] += vol*0.5*bcc3*bcc3;
        }
        for (int n = 0; n < m->p.nscalars; ++n)
          out[NHYDRO + 3 + nme + n] += vol*B->s[CC(B,n,k,j,i)];
      }
  }
  return nout;
}

/* ------------------------------------------------------------------ time step */

/* Hydro::NewBlockTimeStep (src/hydro/new_blockdt.cpp:42-190) */
double ao_new_block_dt(AoMesh *m, int b) {
  AoBlock *B = &m->blk[b];
  double min_dt = DBL_MAX;
  double gamma = m->p.gamma;
  for (int k = B->ks; k <= B->ke; ++k) for (int j = B->js; j <= B->je; ++j)
    for (int i = B->is; i <= B->ie; ++i) {
      double wi[7];
      for (int n = 0; n < NHYDRO; ++n) wi[n] = B->w[CC(B,n,k,j,i)];
      double dt1 = B->dx1f[i], dt2 = B->dx2f[j], dt3 = B->dx3f[k];
      if (m->p.mhd) {
        double b1c = B->bcc[CC(B,IB1,k,j,i)], b2c = B->bcc[CC(B,IB2,k,j,i)],
               b3c = B->bcc[CC(B,IB3,k,j,i)];
        double bx = b1c + fabs(B->b[0][F1(B,k,j,i)] - b1c);
        wi[IBY] = b2c; wi[IBZ] = b3c;
        double cf = ISO(m) ? ao_fast_speed_iso(m->p.iso_cs, wi, bx) : ao_fast_speed(gamma, wi, bx);
        dt1 /= (fabs(wi[IVX]) + cf);
        wi[IBY] = b3c; wi[IBZ] = b1c;
        bx = b2c + fabs(B->b[1][F2(B,k,j,i)] - b2c);
        cf = ISO(m) ? ao_fast_speed_iso(m->p.iso_cs, wi, bx) : ao_fast_speed(gamma, wi, bx);
        dt2 /= (fabs(wi[IVY]) + cf);
        wi[IBY] = b1c; wi[IBZ] = b2c;
        bx = b3c + fabs(B->b[2][F3(B,k,j,i)] - b3c);
        cf = ISO(m) ? ao_fast_speed_iso(m->p.iso_cs, wi, bx) : ao_fast_speed(gamma, wi, bx);
        dt3 /= (fabs(wi[IVZ]) + cf);
      } else {
        double cs = ISO(m) ? m->p.iso_cs : ao_sound_speed(gamma, wi);
        dt1 /= (fabs(wi[IVX]) + cs);
        dt2 /= (fabs(wi[IVY]) + cs);
        dt3 /= (fabs(wi[IVZ]) + cs);
      }
      min_dt = mn(min_dt, dt1);
      if (m->f2) min_dt = mn(min_dt, dt2);
      if (m->f3) min_dt = mn(min_dt, dt3);
    }
  min_dt *= m->cfl;
  B->new_dt = min_dt;
  return min_dt;
}

/* Mesh::NewTimeStep (src/mesh/mesh.cpp:1078-1119) */
static void new_time_step(AoMesh *m) {
  m->dt = 2.0*m->dt;
  for (int g = 0; g < m->nb; ++g) m->dt = mn(m->dt, m->blk[g].new_dt);
  if (m->time < m->p.tlim && (m->p.tlim - m->time) < m->dt) m->dt = m->p.tlim - m->time;
}

void ao_enroll_user_source(AoMesh *m, AoSrcTermFunc fn, void *user) {
  m->user_src = fn; m->user_src_arg = user;
}

void ao_enroll_user_bc(AoMesh *m, int face, AoBValFunc fn, void *user) {
  m->user_bc[face] = fn; m->user_bc_arg[face] = user;
}

/* Mesh::Initialize after ProblemGenerator (src/mesh/mesh.cpp:1416-1649) */
void ao_initialize(AoMesh *m) {
  m->bc_time = m->time; m->bc_dt = 0.0;   /* ApplyPhysicalBoundaries(time, 0.0, ...) mesh.cpp:1515 */
  if (m->multilevel) { smr_exchange_cc(m, 0); smr_exchange_cc(m, 1); } else {
  ao_exchange_cc(m);
  ao_exchange_fc(m);
  ao_exchange_scalars(m);
  }
  for (int g = 0; g < m->nb; ++g) {
    if (m->multilevel) smr_prolongate_boundaries(m, g);   /* mesh.cpp:1527-1528 */
    ao_primitives(m, g);
    ao_physical_bcs(m, g);
  }
  for (int g = 0; g < m->nb; ++g) ao_new_block_dt(m, g);
  new_time_step(m);
}

/* one cycle = all stages of TimeIntegratorTaskList (task order of
 * src/task_list/time_integrator.cpp:899-1098; bodies :1442-2083) then the main-loop
 * bookkeeping of src/main.cpp:478-485 */
double ao_cycle(AoMesh *m) {
  double dt = m->dt;
  for (int stage = 1; stage <= m->nstages; ++stage) {
    int s = stage - 1;
    for (int g = 0; g < m->nb; ++g) if (stage == 1) ao_zero_reg1(m, g);
    int order = (m->p.integrator == AO_INT_VL2 && stage == 1) ? 1 : m->p.xorder;
    for (int g = 0; g < m->nb; ++g) {
      ao_calc_fluxes(m, g, order);
      if (m->p.mhd) ao_corner_e(m, g);
      ao_calc_scalar_fluxes(m, g, order);
    }
    ao_emf_exchange(m);
    if (m->multilevel) { smr_flux_correction(m, 0); smr_flux_correction(m, 1); }   /* SEND_HYDFLX / RECV_HYDFLX before INT_HYD */
    for (int g = 0; g < m->nb; ++g) {
      double w[5] = {1.0, m->delta[s], 0.0, 0.0, 0.0};
      ao_weighted_ave_cc(m, g, 1, 0, w);
      double w2[5] = {m->g1[s], m->g2[s], m->g3[s], 0.0, 0.0};
      if (w2[0] == 0.0 && w2[1] == 1.0 && w2[2] == 0.0) ao_swap_cc(m, g);
      else ao_weighted_ave_cc(m, g, 0, 1, w2);
      ao_add_flux_div(m, g, m->beta[s]*dt);
      const_accel(m, g, m->beta[s]*dt);            /* SRC_TERM after INT_HYD */
      if (m->p.mhd) {
        ao_weighted_ave_fc(m, g, 1, 0, w);
        if (w2[0] == 0.0 && w2[1] == 1.0 && w2[2] == 0.0) ao_swap_fc(m, g);
        else ao_weighted_ave_fc(m, g, 0, 1, w2);
        ao_ct(m, g, m->beta[s]*dt);
      }
      ao_integrate_scalars(m, g, stage);
      /* user-defined source terms, last in AddSourceTerms (hydro_srcterms.cpp:150-153); the
       * SRC_TERM task follows INT_HYD | INT_SCLR; time = start of stage */
      if (m->user_src)
        m->user_src(m->user_src_arg, g, m->time + m->sbeta[s]*dt, m->beta[s]*dt, m->blk[g].w,
                    m->blk[g].r, m->blk[g].bcc, m->blk[g].u, m->blk[g].s);
    }
    if (m->multilevel) { smr_exchange_cc(m, 0); smr_exchange_cc(m, 1); } else {
    ao_exchange_cc(m);
    ao_exchange_fc(m);
    ao_exchange_scalars(m);
    }
    /* PhysicalBoundary task: t_end_stage, beta*dt (time_integrator.cpp:2045-2062) */
    m->bc_time = m->time + m->ebeta[s]*dt; m->bc_dt = m->beta[s]*dt;
    for (int g = 0; g < m->nb; ++g) {
      if (m->multilevel) smr_prolongate_boundaries(m, g);   /* PROLONG: after SETB, before CONS2PRIM */
      ao_primitives(m, g);
      ao_physical_bcs(m, g);
    }
    if (stage == m->nstages)
      for (int g = 0; g < m->nb; ++g) ao_new_block_dt(m, g);
  }
  m->ncycle++;
  m->time += dt;
  new_time_step(m);
  return dt;
}
