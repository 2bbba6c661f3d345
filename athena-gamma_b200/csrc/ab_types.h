// ab_types.h -- plain structs shared by the host runtime and the CUDA kernels.
#ifndef AB_TYPES_H_
#define AB_TYPES_H_

namespace ab {

constexpr int NHYDRO = 5;    // adiabatic; BlkDev::nh holds the run-time count (4 when isothermal)
constexpr int MAX_NB = 26;
constexpr int NUG = 13;        // doubles per cell index in the nonuniform-reconstruction table (ab_physics.cuh)
constexpr int DT_SLOTS = 64;   // atomicMin targets per MeshBlock for the CFL reduction

// Device view of one MeshBlock.  Array layouts are the reference's AthenaArray layouts
// (src/athena_arrays.hpp:140-143: last index fastest; sizes src/hydro/hydro.cpp:31-47,
// src/field/field.cpp:27-38, src/athena.hpp:95-115).
struct BlkDev {
  int nc1, nc2, nc3;              // cells incl. ghosts (1 in a degenerate dimension)
  int is, ie, js, je, ks, ke;     // active range
  int ng;
  int f2, f3;                     // mesh is >=2-D / 3-D
  int nh;                         // NHYDRO of this build: 5 adiabatic, 4 isothermal
  // conserved / primitive registers: NHYDRO x nc3 x nc2 x nc1
  double *u, *u1, *w;
  // face fields: x1f nc3 x nc2 x (nc1+1); x2f nc3 x (nc2+1) x nc1; x3f (nc3+1) x nc2 x nc1
  double *b[3], *b1[3];
  double *bcc;                    // 3 x nc3 x nc2 x nc1
  double *flux[3];                // NHYDRO x (face-array shape)
  double *ef[3][2];               // face EMFs: [dir][0]=ey (e3_x1f,e1_x2f,e2_x3f), [dir][1]=ez
  double *wght[3];                // CT upwind weights (face-array shapes)
  double *e[3];                   // edge EMFs: x1e (nc3+1)(nc2+1)nc1, x2e (nc3+1)nc2(nc1+1), x3e nc3(nc2+1)(nc1+1)
  double *cc_e;                   // 3 x nc3 x nc2 x nc1 cell-centred EMF
  // passive scalars (src/scalars/scalars.hpp:40-56): ns x nc3 x nc2 x nc1; fluxes on the faces
  int ns;
  double *s, *s1, *r, *sflux[3];
  // 1-D geometry (src/coordinates/coordinates.cpp:125-145, cartesian.cpp:25-75)
  const double *x1f, *x2f, *x3f, *x1v, *x2v, *x3v, *dx1f, *dx2f, *dx3f;
  // nullptr when every direction is uniformly spaced; else the CalculateCellCenteredField
  // weights (field/field.cpp:139-172): lw[nc1], rw[nc1] of x1, then x2, then x3
  const double *bcw;
  // bytes between the slabs of consecutive local MeshBlocks (one launch over all blocks of a
  // rank, ab_batch.cuh); 0 when the blocks were allocated separately (AB_DEBUG_ALLOC)
  long bstride;
};

// EMF-correction plan of one block (src/bvals/fc/flux_correction_fc.cpp).  For every face /
// edge neighbour (ids: faces 0..5 = ix1,ox1,ix2,ox2,ix3,ox3; edges 0..11 = reference eid) the
// kernel needs the buffer that neighbour packed for us.
struct EmfPlan {
  const double *face_src[6];      // nullptr when no neighbour across that face
  const double *edge_src[12];     // nullptr when no neighbour across that edge
  double *face_dst[6];            // this block's send buffers
  double *edge_dst[12];
  int face_avg[6];                // 1: face is block/periodic -> interior face EMFs x 0.5
  int nedge_fine[12];             // divisor on the 12 (4 in 2-D) block edges
};

// Exact t / d for 0 <= t < 2^31 by one multiply-high and a shift (the divisor is a launch / plan
// constant; a run-time integer division costs ~20 instructions, a 64-bit one ~70).
// l = ceil(log2 d), m = ceil(2^(31+l) / d) < 2^32, t / d = umulhi(t, m) >> (l-1)
// (Granlund & Montgomery 1994, N = 31).  d = 1 is encoded as m = 0.
struct FastDiv { unsigned m, s; };
static inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u};
  if (d <= 1) return f;
  int l = 0;
  while ((1LL << l) < d) ++l;
  f.m = (unsigned)(((1ULL << (31 + l)) + (unsigned long long)d - 1ULL)/(unsigned long long)d);
  f.s = (unsigned)(l - 1);
  return f;
}

// one box copy of the ghost exchange: dst(k,j,i) = src(k+dk, j+dj, i+di) for the box
struct CopyBox {
  double *dst; const double *src;
  long dst_s3, dst_s2, src_s3, src_s2;   // strides of k and j (i stride 1); var stride below
  long dst_sv, src_sv;                    // stride between variables (0 if nvar==1)
  int nvar;
  int di0, dj0, dk0;                      // dst box origin
  int si0, sj0, sk0;                      // src box origin
  int ni, nj, nk;                         // box extent
  long offset;                            // exclusive prefix of element counts
  FastDiv d_per, d_ni, d_nj;              // divisors ni*nj*nk, ni, nj (set_box_divisors)
};
static inline void set_box_divisors(CopyBox &c) {
  c.d_per = make_fastdiv(c.ni*c.nj*c.nk); c.d_ni = make_fastdiv(c.ni); c.d_nj = make_fastdiv(c.nj);
}

// Static mesh refinement (ab_smr_kernels.cu): index geometry of a block and its coarse buffers
// (MeshBlock::cis.. / cnghost, mesh/meshblock.cpp:82-100) with the 1-D coordinate arrays the
// restriction / prolongation formulas read, and an inclusive index box.
struct SmrGeom {
  int nc1, nc2, nc3, cnc1, cnc2, cnc3, is, js, ks, cis, cjs, cks, ndim;
  const double *dx1f, *dx2f, *dx3f, *x1v, *x2v, *x3v, *cx1v, *cx2v, *cx3v;
};
struct SmrBox { int si, ei, sj, ej, sk, ek; };

struct Params {
  double gamma, dfloor, pfloor;
  int mhd, solver, xorder;
  double sfloor;                  // passive-scalar concentration floor (hydro/sfloor)
  int eos;                        // 0 adiabatic, 1 isothermal
  double iso_cs;                  // hydro/iso_sound_speed
  int char_proj;                  // time/xorder = 2c / 3c
};

}  // namespace ab
#endif
