// ab_smr_exec.h -- interpretation of the host planner's rows (ab_smr.cpp) as work on arrays:
// which restriction / copy / prolongation / flux-correction runs over which box of which block.
// Plain C++ templates over an `Ops` back end, so that the SAME interpretation drives the CUDA
// launches in ab_mesh.cu and the CPU execution in tests/hostcheck/smr_host.cpp, where it is
// compared with the oracle step by step (tests/test_smr_exec_cpu.py).
//
// Restated tasks: SendBoundaryBuffers / SetBoundaries of u (and s) between levels
// (src/bvals/cc/bvals_cc.cpp:195-470), ProlongateBoundaries (src/bvals/bvals_refine.cpp:96-570),
// hydro / scalar flux correction (src/bvals/cc/flux_correction_cc.cpp:69-290).
//
// Ops must provide:
//   restrict_box(geom, fine, coarse, nvar, box)       copy_boxes(std::vector<CopyBox>&)
//   c2p_box(geom, cu, cw, ns, cs, cr, box)            bc_box(geom, cw, nh, cr, ns, face, refl, lo, hi, box)
//   prolong_box(geom, coarse, fine, nvar, box)        prim2cons_box(block, il, iu, jl, ju, kl, ku)
//   flux_face(geom, fine_flux, coarse_flux, nvar, dir, fpos, cpos, a0, b0, na, nb)
#ifndef AB_SMR_EXEC_H_
#define AB_SMR_EXEC_H_
#include <array>
#include <cstring>
#include <vector>

#include "ab_types.h"

namespace ab {

// arrays of one MeshBlock as the SMR steps see them (device or host pointers)
struct SmrView {
  double *u = nullptr, *s = nullptr, *w = nullptr, *r = nullptr;
  double *flux[3] = {nullptr, nullptr, nullptr}, *sflux[3] = {nullptr, nullptr, nullptr};
  double *cu = nullptr, *cw = nullptr, *cs = nullptr, *cr = nullptr;
  SmrGeom g;
  int bcs[6];          // -1 block boundary, else the mesh's boundary flag (0 periodic, 1 outflow, 2 reflecting)
};

struct SmrDims {        // what every block of the mesh shares
  int nh, ns, ng, bx[3], s0[3], e0[3];
  bool fdim[3];
};

using SmrRow = std::array<long, 12>;

inline SmrBox smr_box(const long *origin, const long *extent) {
  return SmrBox{(int)origin[0], (int)(origin[0] + extent[0] - 1), (int)origin[1],
                (int)(origin[1] + extent[1] - 1), (int)origin[2], (int)(origin[2] + extent[2] - 1)};
}

// the box copy of an exchange row (kind 0 same level, 1 into the finer block's coarse buffer, 2 from
// the finer block's coarse buffer) for the hydro variables (pass 0) or the scalars (pass 1)
inline CopyBox smr_row_copy(const SmrRow &r, const SmrView &S, const SmrView &T, const SmrDims &d,
                            int pass) {
  const long ncc_s = (long)S.g.nc1*S.g.nc2*S.g.nc3, ncc_t = (long)T.g.nc1*T.g.nc2*T.g.nc3;
  const long cncc_s = (long)S.g.cnc1*S.g.cnc2*S.g.cnc3, cncc_t = (long)T.g.cnc1*T.g.cnc2*T.g.cnc3;
  CopyBox c;
  std::memset(&c, 0, sizeof(c));
  const bool src_coarse = (r[0] == 2), dst_coarse = (r[0] == 1);
  c.src = src_coarse ? (pass ? S.cs : S.cu) : (pass ? S.s : S.u);
  c.dst = dst_coarse ? (pass ? T.cs : T.cu) : (pass ? T.s : T.u);
  if (src_coarse) { c.src_s3 = (long)S.g.cnc2*S.g.cnc1; c.src_s2 = S.g.cnc1; c.src_sv = cncc_s; }
  else { c.src_s3 = (long)S.g.nc2*S.g.nc1; c.src_s2 = S.g.nc1; c.src_sv = ncc_s; }
  if (dst_coarse) { c.dst_s3 = (long)T.g.cnc2*T.g.cnc1; c.dst_s2 = T.g.cnc1; c.dst_sv = cncc_t; }
  else { c.dst_s3 = (long)T.g.nc2*T.g.nc1; c.dst_s2 = T.g.nc1; c.dst_sv = ncc_t; }
  c.nvar = pass ? d.ns : d.nh;
  c.si0 = (int)r[2]; c.sj0 = (int)r[3]; c.sk0 = (int)r[4];
  c.di0 = (int)r[6]; c.dj0 = (int)r[7]; c.dk0 = (int)r[8];
  c.ni = (int)r[9]; c.nj = (int)r[10]; c.nk = (int)r[11];
  return c;
}
// the same copy with one end in a message buffer (the box stored compactly, variable slowest):
// `to_buffer`: source block -> buffer at buf; else buffer at buf -> destination block
inline CopyBox smr_row_copy_buffered(const SmrRow &r, const SmrView &S, const SmrView &T,
                                     const SmrDims &d, int pass, double *buf, bool to_buffer) {
  CopyBox c = smr_row_copy(r, S, T, d, pass);
  const long s2 = c.ni, s3 = (long)c.ni*c.nj, sv = s3*c.nk;
  if (to_buffer) { c.dst = buf; c.dst_s2 = s2; c.dst_s3 = s3; c.dst_sv = sv; c.di0 = c.dj0 = c.dk0 = 0; }
  else { c.src = buf; c.src_s2 = s2; c.src_s3 = s3; c.src_sv = sv; c.si0 = c.sj0 = c.sk0 = 0; }
  return c;
}

// ghost exchange of u (and s): rows 2 first restrict the senders' slabs, then every row 0 / 1 / 2
// is one box copy
template <class Ops>
void smr_run_exchange(const std::vector<SmrRow> &rows, std::vector<SmrView> &v, const SmrDims &d,
                      Ops &ops) {
  for (const auto &r : rows) {
    if (r[0] != 2) continue;
    SmrView &S = v[r[1]];
    const SmrBox bx = smr_box(&r[2], &r[9]);
    ops.restrict_box(S.g, S.u, S.cu, d.nh, bx);
    if (d.ns > 0) ops.restrict_box(S.g, S.s, S.cs, d.ns, bx);
  }
  std::vector<CopyBox> boxes;
  for (const auto &r : rows) {
    if (r[0] < 0 || r[0] > 2) continue;
    SmrView &S = v[r[1]], &T = v[r[5]];
    for (int pass = 0; pass < (d.ns > 0 ? 2 : 1); ++pass) {
      CopyBox c = smr_row_copy(r, S, T, d, pass);
      boxes.push_back(c);
    }
  }
  long total = 0;
  for (auto &c : boxes) { c.offset = total; total += (long)c.ni*c.nj*c.nk*c.nvar; }
  ops.copy_boxes(boxes, total);
}

// BoundaryValues::ProlongateBoundaries of block `lid` (its gid: the rows name blocks by gid): rows 12 (restriction of the ghost cells
// that same-level neighbours filled), then per coarser neighbour rows 10 / 11: ConservedToPrimitive
// on the coarse box with its margins, the boundary functions of the mesh faces this block
// touches, prolongation of the primitives, PrimitiveToConserved on the fine ghost cells
template <class Ops>
void smr_run_prolongate(const std::vector<SmrRow> &rows, int lid, SmrView &B, const SmrDims &d,
                        Ops &ops) {
  const int cs[3] = {B.g.cis, B.g.cjs, B.g.cks};
  int ce[3];
  for (int k = 0; k < 3; ++k) ce[k] = d.fdim[k] ? cs[k] + d.bx[k]/2 - 1 : 0;
  for (size_t n = 0; n < rows.size(); ++n) {
    const auto &r = rows[n];
    if (r[1] != lid) continue;
    if (r[0] == 12) {
      const SmrBox bx = smr_box(&r[2], &r[9]);
      ops.restrict_box(B.g, B.u, B.cu, d.nh, bx);
      if (d.ns > 0) ops.restrict_box(B.g, B.s, B.cs, d.ns, bx);
    } else if (r[0] == 10) {
      const auto &r2 = rows[n + 1];               // row 11 follows its row 10
      const SmrBox pbx = smr_box(&r[2], &r[9]);
      const SmrBox cbx{(int)r[6], (int)r2[2], (int)r[7], (int)r2[3], (int)r[8], (int)r2[4]};
      const int ox[3] = {(int)r2[6], (int)r2[7], (int)r2[8]};
      ops.c2p_box(B.g, B.cu, B.cw, d.ns, B.cs, B.cr, cbx);
      for (int k = 0; k < 3; ++k) {
        if (!d.fdim[k] || ox[k] != 0) continue;
        for (int side = 0; side < 2; ++side) {
          const int face = 2*k + side, bc = B.bcs[face];
          if (bc != 1 && bc != 2) continue;       // outflow, reflecting
          SmrBox t = pbx;                         // transverse range; the normal index is lo / hi
          if (k == 0) { t.si = t.ei = 0; } else if (k == 1) { t.sj = t.ej = 0; } else { t.sk = t.ek = 0; }
          ops.bc_box(B.g, B.cw, d.nh, B.cr, d.ns, face, bc == 2, cs[k], ce[k], t);
        }
      }
      ops.prolong_box(B.g, B.cw, B.w, d.nh, pbx);
      if (d.ns > 0) ops.prolong_box(B.g, B.cr, B.r, d.ns, pbx);
      int f0[3], f1[3];
      const int p0[3] = {pbx.si, pbx.sj, pbx.sk}, p1[3] = {pbx.ei, pbx.ej, pbx.ek};
      for (int k = 0; k < 3; ++k) {
        if (d.fdim[k]) { f0[k] = (p0[k] - cs[k])*2 + d.s0[k]; f1[k] = (p1[k] - cs[k])*2 + d.s0[k] + 1; }
        else { f0[k] = d.s0[k]; f1[k] = d.e0[k]; }
      }
      ops.prim2cons_box(lid, f0[0], f1[0], f0[1], f1[1], f0[2], f1[2]);
    }
  }
}

// SendFluxCorrection / ReceiveFluxCorrection: rows 20 {fine gid, its face, -, -, coarse gid, its
// face, fi1, fi2}
template <class Ops>
void smr_run_flux_correction(const std::vector<SmrRow> &rows, std::vector<SmrView> &v,
                             const SmrDims &d, Ops &ops) {
  for (const auto &r : rows) {
    if (r[0] != 20) continue;
    SmrView &F = v[r[1]], &Cb = v[r[5]];
    const int ffid = (int)r[2], cfid = (int)r[6], fi1 = (int)r[7], fi2 = (int)r[8];
    const int dir = ffid >> 1;
    const int fpos = d.s0[dir] + (d.e0[dir] - d.s0[dir] + 1)*(ffid & 1);   // face index, fine block
    const int cpos = d.s0[dir] + (d.e0[dir] - d.s0[dir] + 1)*(cfid & 1);   // face index, coarse block
    const int da = dir == 0 ? 1 : 0, db = dir == 2 ? 1 : 2;                // transverse directions
    const int ha = d.fdim[da] ? d.bx[da]/2 : 0, hb = d.fdim[db] ? d.bx[db]/2 : 0;
    const int a0 = d.s0[da] + (fi1 ? ha : 0), b0 = d.s0[db] + (fi2 ? hb : 0);
    const int na = d.fdim[da] ? ha : 1, nb = d.fdim[db] ? hb : 1;
    ops.flux_face(F.g, F.flux[dir], Cb.flux[dir], d.nh, dir, fpos, cpos, a0, b0, na, nb);
    if (d.ns > 0)
      ops.flux_face(F.g, F.sflux[dir], Cb.sflux[dir], d.ns, dir, fpos, cpos, a0, b0, na, nb);
  }
}

}  // namespace ab
#endif
