// ab_smr.cpp -- host-side planner for static mesh refinement (mesh/refinement = static).
//
// Plain C++ (no CUDA): everything a refined mesh needs BEFORE any kernel runs.
//   * the MeshBlock tree grown from the <refinementN> regions with the 2:1 balance rule, and its
//     Z-ordered leaf list            (src/mesh/meshblock_tree.cpp:60-352, src/mesh/mesh.cpp:323-465)
//   * the load balance over ranks    (src/mesh/amr_loadbalance.cpp:72-112)
//   * the level-aware neighbour list of every block  (src/bvals/bvals_base.cpp:299-736)
//   * the transfer plan of one cell-centred ghost exchange, sender by sender: same level,
//     fine -> coarse (restricted slab) and coarse -> fine (into the receiver's coarse buffer)
//                                     (src/bvals/cc/bvals_cc.cpp:195-470)
//   * the per-block ProlongateBoundaries work list (src/bvals/bvals_refine.cpp:96-570)
//   * the flux-correction pairs      (src/bvals/cc/flux_correction_cc.cpp:69-290)
// The plan is pure index arithmetic; tests/test_smr_plan_cpu.py compares it row by row with the
// transfer log of the C oracle, which reproduces the reference's SMR runs bit for bit.  The
// device side (restriction / prolongation kernels executing this plan) is the next step; until
// then ab_mesh_create rejects refinement and only the ab_smr_plan_* entry points use this file.
#include <algorithm>
#include <array>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../../include/athena_b200.h"

namespace {

struct Node {
  int level = 0;
  long lx[3] = {0, 0, 0};
  std::unique_ptr<Node> leaf[8];
  bool split = false;
  int gid = -1;
};

struct Nbr { int ox[3]; int type, gid, level, fi1, fi2, fid; };

struct Block {
  int gid = 0, level = 0, rank = 0;
  long lx[3] = {0, 0, 0};
  int nblevel[3][3][3];
  std::vector<Nbr> nbs;
  // index ranges (mesh/meshblock.cpp:55-100): fine is..ke, coarse cis..cke
  int s[3], e[3], cs[3], ce[3];
};

using Row = std::array<long, 12>;

}  // namespace

struct AbSmrPlan {
  AbMeshParams p;
  int ndim = 1, root_level = 0, ng = 2, cng = 2;
  bool f[3] = {true, false, false};
  int nrb[3] = {1, 1, 1}, bx[3] = {1, 1, 1};
  std::unique_ptr<Node> root;
  std::vector<Block> blocks;
  std::vector<Row> rows;

  int nleaf() const { return f[2] ? 8 : (f[1] ? 4 : 2); }
  long nblocks_at(int level, int d) const { return (long)nrb[d] << (level - root_level); }

  // ---- tree ------------------------------------------------------------------------------
  static Node *child(Node *t, int i, int j, int k) {
    auto c = std::make_unique<Node>();
    c->level = t->level + 1;
    c->lx[0] = (t->lx[0] << 1) + i; c->lx[1] = (t->lx[1] << 1) + j; c->lx[2] = (t->lx[2] << 1) + k;
    Node *raw = c.get();
    t->leaf[i + (j << 1) + (k << 2)] = std::move(c);
    return raw;
  }
  void root_grid(Node *t) {          // CreateRootGrid
    if (t->level == root_level) return;
    const long fac = 1L << (root_level - t->level - 1);
    t->split = true;
    for (int n = 0; n < nleaf(); ++n) {
      const int i = n & 1, j = (n >> 1) & 1, k = (n >> 2) & 1;
      if ((t->lx[0]*2 + i)*fac < nrb[0] && (t->lx[1]*2 + j)*fac < nrb[1]
          && (t->lx[2]*2 + k)*fac < nrb[2])
        root_grid(child(t, i, j, k));
    }
  }
  bool wrap(long &x, long n, int d) const {   // periodic wrap; false = outside the mesh
    if (x < 0) { if (p.bc[2*d] != AB_BC_PERIODIC) return false; x = n - 1; }
    if (x >= n) { if (p.bc[2*d+1] != AB_BC_PERIODIC) return false; x = 0; }
    return true;
  }
  void refine(Node *t) {              // Refine: split and force the neighbours into existence
    if (t->split) return;
    t->split = true;
    for (int n = 0; n < nleaf(); ++n) child(t, n & 1, (n >> 1) & 1, (n >> 2) & 1);
    for (int oz = (f[2] ? -1 : 0); oz <= (f[2] ? 1 : 0); ++oz)
      for (int oy = (f[1] ? -1 : 0); oy <= (f[1] ? 1 : 0); ++oy)
        for (int ox = -1; ox <= 1; ++ox) {
          if (!ox && !oy && !oz) continue;
          long x = t->lx[0] + ox, y = t->lx[1] + oy, z = t->lx[2] + oz;
          if (!wrap(x, nblocks_at(t->level, 0), 0)) continue;
          if (f[1] && !wrap(y, nblocks_at(t->level, 1), 1)) continue;
          if (f[2] && !wrap(z, nblocks_at(t->level, 2), 2)) continue;
          add(root.get(), t->level, x, y, z);
        }
  }
  void add(Node *t, int level, long x, long y, long z) {   // AddMeshBlock
    if (t->level == level) return;
    if (!t->split) refine(t);
    const int sh = level - t->level - 1;
    const int n = (int)((x >> sh) & 1) + ((int)((y >> sh) & 1) << 1) + ((int)((z >> sh) & 1) << 2);
    add(t->leaf[n].get(), level, x, y, z);
  }
  void list(Node *t) {                // GetMeshBlockList: depth first = Z-order
    if (!t->split) {
      t->gid = (int)blocks.size();
      Block b;
      b.gid = t->gid; b.level = t->level;
      for (int d = 0; d < 3; ++d) b.lx[d] = t->lx[d];
      blocks.push_back(b);
      return;
    }
    for (int n = 0; n < 8; ++n) if (t->leaf[n]) list(t->leaf[n].get());
  }
  const Node *find(int level, const long lx[3], const int o[3]) const {   // FindNeighbor
    long x[3] = {lx[0] + o[0], lx[1] + o[1], lx[2] + o[2]};
    for (int d = 0; d < 3; ++d)
      if (!wrap(x[d], nblocks_at(level, d), d)) return nullptr;
    const Node *t = root.get();
    for (int l = 0; l < level; ++l) {
      if (!t->split) return t;        // a coarser leaf
      const int sh = level - l - 1;
      t = t->leaf[(int)((x[0] >> sh) & 1) + ((int)((x[1] >> sh) & 1) << 1)
                  + ((int)((x[2] >> sh) & 1) << 2)].get();
    }
    return t;                         // same-level leaf, or the parent of finer leaves
  }

  // uniform mesh generator on block edges (mesh/mesh.hpp:389-405,467-488; mesh.cpp:1679-1688)
  double edge(long index, long nrange, int d) const {
    const double mn[3] = {p.x1min, p.x2min, p.x3min}, mx[3] = {p.x1max, p.x2max, p.x3max};
    const long a = index - nrange/2, b = index - (nrange + 1)/2;
    const double x = static_cast<double>(a + b)/(2.0*nrange);
    return 0.5*(mn[d] + mx[d]) + (x*mx[d] - x*mn[d]);
  }

  // ---- neighbours ------------------------------------------------------------------------
  void push_nb(Block &B, const Node *t, const int o[3], int type, int fi1, int fi2) {
    Nbr nb;
    for (int d = 0; d < 3; ++d) nb.ox[d] = o[d];
    nb.type = type; nb.gid = t->gid; nb.level = t->level; nb.fi1 = fi1; nb.fi2 = fi2; nb.fid = -1;
    if (type == 0) for (int d = 0; d < 3; ++d) if (o[d]) nb.fid = 2*d + (o[d] > 0);
    B.nbs.push_back(nb);
  }
  void search_neighbors(Block &B) {
    const int par[3] = {(int)(B.lx[0] & 1), (int)(B.lx[1] & 1), (int)(B.lx[2] & 1)};
    const int out[3] = {par[0]*2 - 1, f[1] ? par[1]*2 - 1 : 0, f[2] ? par[2]*2 - 1 : 0};
    const int nf1 = f[1] ? 2 : 1, nf2 = f[2] ? 2 : 1;
    for (auto &pl : B.nblevel) for (auto &row : pl) for (int &v : row) v = -1;
    B.nblevel[1][1][1] = B.level;
    auto mark = [&](const int o[3], int lev) { B.nblevel[o[2]+1][o[1]+1][o[0]+1] = lev; };
    for (int d = 0; d < ndim; ++d) for (int n = -1; n <= 1; n += 2) {       // faces
      int o[3] = {0, 0, 0}; o[d] = n;
      const Node *t = find(B.level, B.lx, o);
      if (!t) continue;
      if (!t->split) { mark(o, t->level); push_nb(B, t, o, 0, 0, 0); continue; }
      mark(o, t->level + 1);
      const int ff = 1 - (n + 1)/2;
      for (int f2 = 0; f2 < nf2; ++f2) for (int f1 = 0; f1 < nf1; ++f1) {
        int l[3];
        if (d == 0) { l[0] = ff; l[1] = f1; l[2] = f2; }
        else if (d == 1) { l[0] = f1; l[1] = ff; l[2] = f2; }
        else { l[0] = f1; l[1] = f2; l[2] = ff; }
        push_nb(B, t->leaf[l[0] + (l[1] << 1) + (l[2] << 2)].get(), o, 0, f1, f2);
      }
    }
    if (!f[1]) return;
    const int pairs[3][2] = {{0, 1}, {0, 2}, {1, 2}};                       // edges
    for (int e = 0; e < (f[2] ? 3 : 1); ++e) {
      const int da = pairs[e][0], db = pairs[e][1], dc = 3 - da - db;
      const int nfe = (e == 0) ? nf2 : nf1;
      for (int m = -1; m <= 1; m += 2) for (int n = -1; n <= 1; n += 2) {
        int o[3] = {0, 0, 0}; o[da] = n; o[db] = m;
        const Node *t = find(B.level, B.lx, o);
        if (!t) continue;
        if (t->split) {
          mark(o, t->level + 1);
          for (int f1 = 0; f1 < nfe; ++f1) {
            int l[3]; l[da] = 1 - (n + 1)/2; l[db] = 1 - (m + 1)/2; l[dc] = f1;
            push_nb(B, t->leaf[l[0] + (l[1] << 1) + (l[2] << 2)].get(), o, 1, f1, 0);
          }
        } else {
          mark(o, t->level);
          // a coarser edge neighbour is only a neighbour of the outward corner of the parent
          if (t->level >= B.level || (out[da] == n && out[db] == m)) push_nb(B, t, o, 1, 0, 0);
        }
      }
    }
    if (!f[2]) return;
    for (int l = -1; l <= 1; l += 2) for (int m = -1; m <= 1; m += 2) for (int n = -1; n <= 1; n += 2) {
      int o[3] = {n, m, l};                                                 // corners
      const Node *t = find(B.level, B.lx, o);
      if (!t) continue;
      if (t->split)
        t = t->leaf[(1 - (n + 1)/2) + ((1 - (m + 1)/2) << 1) + ((1 - (l + 1)/2) << 2)].get();
      mark(o, t->level);
      if (t->level >= B.level || (out[0] == n && out[1] == m && out[2] == l))
        push_nb(B, t, o, 2, 0, 0);
    }
  }

  // which of the finer leaves along the shared face / edge a block is, as its coarser
  // neighbour counts them
  void my_fi(const Block &B, const int o[3], int &fi1, int &fi2) const {
    const int par[3] = {(int)(B.lx[0] & 1), (int)(B.lx[1] & 1), (int)(B.lx[2] & 1)};
    std::vector<int> free_dirs;
    for (int d = 0; d < 3; ++d) if (!o[d]) free_dirs.push_back(d);
    fi1 = fi2 = 0;
    if (free_dirs.size() == 2) { fi1 = par[free_dirs[0]]; fi2 = par[free_dirs[1]]; }
    else if (free_dirs.size() == 1) fi1 = par[free_dirs[0]];
  }

  // ---- transfer plan -----------------------------------------------------------------------
  void add_row(long kind, long src, const int so[3], long dst, const int d0[3], const int n[3]) {
    rows.push_back(Row{kind, src, so[0], so[1], so[2], dst, d0[0], d0[1], d0[2], n[0], n[1], n[2]});
  }
  void plan_exchange() {
    for (const Block &S : blocks) for (const Nbr &nb : S.nbs) {
      const Block &T = blocks[nb.gid];
      int so[3], d0[3], n[3];
      if (nb.level == S.level) {
        for (int d = 0; d < 3; ++d) {
          // LoadBoundaryBufferSameLevel on S; SetBoundarySameLevel on T with the mirrored offset
          const int o = nb.ox[d];
          so[d] = (o > 0) ? S.e[d] - ng + 1 : S.s[d];
          const int se = (o < 0) ? S.s[d] + ng - 1 : S.e[d];
          n[d] = se - so[d] + 1;
          d0[d] = (o == 0) ? T.s[d] : ((-o > 0) ? T.e[d] + 1 : T.s[d] - ng);
        }
        add_row(0, S.gid, so, T.gid, d0, n);
      } else if (nb.level < S.level) {
        // LoadBoundaryBufferToCoarser: S restricts NGHOST coarse cells; SetBoundaryFromFiner on T
        int fi1, fi2;
        my_fi(S, nb.ox, fi1, fi2);
        const int cn = ng - 1;
        int dir_index = 0;      // position among the directions with zero offset, in x1,x2,x3 order
        const int nzero = (nb.ox[0] == 0) + (nb.ox[1] == 0) + (nb.ox[2] == 0);
        for (int d = 0; d < 3; ++d) {
          const int o = nb.ox[d];
          so[d] = (o > 0) ? S.ce[d] - cn : S.cs[d];
          const int se = (o < 0) ? S.cs[d] + cn : S.ce[d];
          n[d] = se - so[d] + 1;
          const int to = -o;
          if (to == 0) {
            d0[d] = T.s[d];
            if (f[d]) {
              const int fi = (nzero == 2) ? (dir_index == 0 ? fi1 : fi2) : fi1;
              if (fi == 1) d0[d] += bx[d]/2;
            }
            dir_index++;
          } else {
            d0[d] = (to > 0) ? T.e[d] + 1 : T.s[d] - ng;
          }
        }
        add_row(2, S.gid, so, T.gid, d0, n);
      } else {
        // LoadBoundaryBufferToFiner on S; SetBoundaryFromCoarser into T's coarse buffer
        const int cn = cng - 1;
        int dir_index = 0;
        const int nzero = (nb.ox[0] == 0) + (nb.ox[1] == 0) + (nb.ox[2] == 0);
        for (int d = 0; d < 3; ++d) {
          const int o = nb.ox[d];
          int a = (o > 0) ? S.e[d] - cn : S.s[d], b = (o < 0) ? S.s[d] + cn : S.e[d];
          if (o == 0) {
            if (f[d]) {
              const int fi = (nzero == 2) ? (dir_index == 0 ? nb.fi1 : nb.fi2) : nb.fi1;
              if (fi == 1) a += bx[d]/2 - cng; else b -= bx[d]/2 - cng;
            }
            dir_index++;
          }
          so[d] = a; n[d] = b - a + 1;
          const int to = -o;
          if (to == 0) {
            d0[d] = T.cs[d];
            if (f[d] && (T.lx[d] & 1)) d0[d] -= cng;
          } else {
            d0[d] = (to > 0) ? T.ce[d] + 1 : T.cs[d] - cng;
          }
        }
        add_row(1, S.gid, so, T.gid, d0, n);
      }
    }
  }

  void plan_prolongation() {
    for (const Block &B : blocks) for (const Nbr &nb : B.nbs) {
      if (nb.level >= B.level) continue;
      // Step 1: ghost cells filled by same-level neighbours are restricted into the coarse buffer
      int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      for (int d = 0; d < 3; ++d) if (f[d]) {
        lo[d] = std::max(nb.ox[d] - 1, -1); hi[d] = std::min(nb.ox[d] + 1, 1);
      }
      for (int nk = lo[2]; nk <= hi[2]; ++nk) for (int nj = lo[1]; nj <= hi[1]; ++nj)
        for (int ni = lo[0]; ni <= hi[0]; ++ni) {
          if ((!ni && !nj && !nk) || B.nblevel[nk+1][nj+1][ni+1] != B.level) continue;
          const int nn[3] = {ni, nj, nk};
          int r0[3], rn[3];
          for (int d = 0; d < 3; ++d) {
            int a, b;
            if (nn[d] == 0) {
              a = B.cs[d]; b = B.ce[d];
              if (nb.ox[d] == 1) a = B.ce[d]; else if (nb.ox[d] == -1) b = B.cs[d];
            } else if (nn[d] == 1) { a = b = B.ce[d] + 1; } else { a = b = B.cs[d] - 1; }
            r0[d] = a; rn[d] = b - a + 1;
          }
          add_row(12, B.gid, r0, B.gid, nb.ox, rn);
        }
      // the coarse ghost box to prolongate, and the margins of the coarse ConservedToPrimitive
      const int cn = cng - 1;
      int s0[3], sn[3], c0[3], c1[3];
      for (int d = 0; d < 3; ++d) {
        int a, b;
        const int o = nb.ox[d];
        if (o == 0) {
          a = B.cs[d]; b = B.ce[d];
          if (f[d]) { if ((B.lx[d] & 1) == 0) b += cn; else a -= cn; }
        } else if (o > 0) { a = B.ce[d] + 1; b = B.ce[d] + cn; } else { a = B.cs[d] - cn; b = B.cs[d] - 1; }
        int fm = 0, fp = 0;
        if (f[d]) {
          if (o == 0) {
            int lo_idx[3] = {1, 1, 1}, hi_idx[3] = {1, 1, 1};
            lo_idx[d] = 0; hi_idx[d] = 2;
            if (B.nblevel[lo_idx[2]][lo_idx[1]][lo_idx[0]] != -1) fm = 1;
            if (B.nblevel[hi_idx[2]][hi_idx[1]][hi_idx[0]] != -1) fp = 1;
          } else { fm = fp = 1; }
        }
        s0[d] = a; sn[d] = b - a + 1; c0[d] = a - fm; c1[d] = b + fp;
      }
      add_row(10, B.gid, s0, B.gid, c0, sn);
      const int zero[3] = {0, 0, 0};
      add_row(11, B.gid, c1, B.gid, nb.ox, zero);
    }
  }

  void plan_flux_correction() {
    for (const Block &S : blocks) for (const Nbr &nb : S.nbs) {
      if (nb.type != 0 || nb.level >= S.level) continue;
      int fi1, fi2;
      my_fi(S, nb.ox, fi1, fi2);
      const int so[3] = {nb.fid, 0, 0}, d0[3] = {nb.fid ^ 1, fi1, fi2}, zero[3] = {0, 0, 0};
      add_row(20, S.gid, so, nb.gid, d0, zero);
    }
  }

  // Mesh::CalculateLoadBalance with unit costs (mesh/amr_loadbalance.cpp:72-112)
  void load_balance() {
    const int nb = (int)blocks.size(), nranks = std::max(p.nranks, 1);
    double total = nb, target = total/nranks, mine = 0.0;
    int j = nranks - 1;
    for (int i = nb - 1; i >= 0; --i) {
      mine += 1.0;
      blocks[i].rank = j;
      if (mine >= target && j > 0) { --j; total -= mine; mine = 0.0; target = total/(j + 1); }
    }
  }

  std::string build(const AbMeshParams &pp, const AbRefinementRegion *reg, int nreg) {
    p = pp;
    f[1] = p.nx2 > 1; f[2] = p.nx3 > 1;
    ndim = f[2] ? 3 : (f[1] ? 2 : 1);
    bx[0] = p.bx1; bx[1] = p.bx2; bx[2] = p.bx3;
    const int nx[3] = {p.nx1, p.nx2, p.nx3};
    for (int d = 0; d < 3; ++d) {
      if (bx[d] <= 0 || nx[d] % bx[d]) return "the Mesh must be evenly divisible by the MeshBlock";
      nrb[d] = nx[d]/bx[d];
      if (f[d] && bx[d] % 2) return "the size of MeshBlock must be divisible by 2 with SMR";
    }
    ng = p.nghost;
    if (ng % 2) return "an even number of ghost cells is required with mesh refinement";
    cng = (ng + 1)/2 + 1;
    const int nbmax = std::max(nrb[0], std::max(nrb[1], nrb[2]));
    for (root_level = 0; (1 << root_level) < nbmax; ++root_level) {}
    root = std::make_unique<Node>();
    root_grid(root.get());
    const double lo[3] = {p.x1min, p.x2min, p.x3min}, hi[3] = {p.x1max, p.x2max, p.x3max};
    for (int r = 0; r < nreg; ++r) {
      const double rmin[3] = {reg[r].x1min, reg[r].x2min, reg[r].x3min},
                   rmax[3] = {reg[r].x1max, reg[r].x2max, reg[r].x3max};
      if (reg[r].level < 1) return "refinement level must be larger than 0";
      long a[3] = {0, 0, 0}, b[3] = {1, 1, 1};
      for (int d = 0; d < ndim; ++d) {
        if (rmin[d] > rmax[d]) return "invalid refinement region";
        if (rmin[d] < lo[d] || rmax[d] > hi[d]) return "refinement region must be smaller than the whole mesh";
        const long lxmax = (long)nrb[d]*(1L << reg[r].level);
        long s0, e0;
        for (s0 = 0; s0 < lxmax; ++s0) if (edge(s0 + 1, lxmax, d) > rmin[d]) break;
        for (e0 = s0; e0 < lxmax; ++e0) if (edge(e0 + 1, lxmax, d) >= rmax[d]) break;
        if (s0 % 2 == 1) --s0;
        if (e0 % 2 == 0) ++e0;
        a[d] = s0; b[d] = e0;
      }
      for (long k = a[2]; k < b[2]; k += 2) for (long j = a[1]; j < b[1]; j += 2)
        for (long i = a[0]; i < b[0]; i += 2) add(root.get(), reg[r].level + root_level, i, j, k);
    }
    list(root.get());
    for (Block &B : blocks) {
      for (int d = 0; d < 3; ++d) {
        if (f[d]) { B.s[d] = ng; B.e[d] = ng + bx[d] - 1; B.cs[d] = cng; B.ce[d] = cng + bx[d]/2 - 1; }
        else { B.s[d] = B.e[d] = B.cs[d] = B.ce[d] = 0; }
      }
      search_neighbors(B);
    }
    load_balance();
    plan_exchange();
    plan_prolongation();
    plan_flux_correction();
    return "";
  }
};

// ---- C ABI ---------------------------------------------------------------------------------
namespace { thread_local std::string g_smr_err; }

extern "C" {

const char *ab_smr_last_error(void) { return g_smr_err.c_str(); }

int ab_smr_plan_create(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                       AbSmrPlan **out) {
  if (!p || !out || nregions < 0 || (nregions > 0 && !regions)) {
    g_smr_err = "bad argument"; return AB_ERR_ARG;
  }
  auto plan = std::make_unique<AbSmrPlan>();
  const std::string err = plan->build(*p, regions, nregions);
  if (!err.empty()) { g_smr_err = err; return AB_ERR_ARG; }
  *out = plan.release();
  return AB_OK;
}

int ab_smr_plan_destroy(AbSmrPlan *plan) { delete plan; return AB_OK; }

int ab_smr_plan_nblocks(const AbSmrPlan *plan) { return plan ? (int)plan->blocks.size() : 0; }

int ab_smr_plan_blocks(const AbSmrPlan *plan, long *rows, int max_rows) {
  if (!plan) return AB_ERR_ARG;
  for (int g = 0; g < (int)plan->blocks.size() && g < max_rows && rows; ++g) {
    const Block &B = plan->blocks[g];
    long *r = rows + 5*g;
    r[0] = B.level; r[1] = B.lx[0]; r[2] = B.lx[1]; r[3] = B.lx[2]; r[4] = B.rank;
  }
  return (int)plan->blocks.size();
}

int ab_smr_plan_neighbors(const AbSmrPlan *plan, int gid, int *rows, int *nblevel) {
  if (!plan || gid < 0 || gid >= (int)plan->blocks.size()) return AB_ERR_ARG;
  const Block &B = plan->blocks[gid];
  for (size_t n = 0; n < B.nbs.size() && rows; ++n) {
    const Nbr &nb = B.nbs[n];
    int *r = rows + 8*n;
    r[0] = nb.ox[0]; r[1] = nb.ox[1]; r[2] = nb.ox[2]; r[3] = nb.type; r[4] = nb.gid;
    r[5] = nb.level; r[6] = nb.fi1; r[7] = nb.fi2;
  }
  if (nblevel) for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
    nblevel[(k*3 + j)*3 + i] = B.nblevel[k][j][i];
  return (int)B.nbs.size();
}

long ab_smr_plan_transfers(const AbSmrPlan *plan, long *rows, long max_rows) {
  if (!plan) return AB_ERR_ARG;
  for (long n = 0; n < (long)plan->rows.size() && n < max_rows && rows; ++n)
    for (int c = 0; c < 12; ++c) rows[12*n + c] = plan->rows[n][c];
  return (long)plan->rows.size();
}

}  // extern "C"
