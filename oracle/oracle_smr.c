/* oracle_smr.c -- static mesh refinement (hydro, cell-centred variables) for the CPU oracle.
 * TEST INFRASTRUCTURE ONLY.  #included at the end of oracle_mesh.c (shares its structs).
 *
 * Restates, for a single process and without message buffers:
 *   the MeshBlock tree and Z-ordered block list     src/mesh/meshblock_tree.cpp:60-352
 *   refinement regions of the Mesh ctor              src/mesh/mesh.cpp:323-465
 *   the level-aware neighbour search                 src/bvals/bvals_base.cpp:299-736
 *   cc ghost exchange between levels                 src/bvals/cc/bvals_cc.cpp:195-470
 *   restriction / prolongation                       src/mesh/mesh_refinement.cpp:106-176,386-540
 *   ProlongateBoundaries                             src/bvals/bvals_refine.cpp:96-570
 *   flux correction fine -> coarse                   src/bvals/cc/flux_correction_cc.cpp:69-290
 * Scope: hydro (no face-centred field, no passive scalars), periodic / outflow / reflecting
 * mesh boundaries.  ao_create rejects anything else with refinement.
 */

#ifndef SIGN
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
#endif

struct TNode {
  int level; long lx1, lx2, lx3;
  struct TNode *leaf[8];
  int has_leaf, gid;
};

static TNode *tnode_new(const TNode *parent, int ox1, int ox2, int ox3) {
  TNode *t = (TNode *)calloc(1, sizeof(TNode));
  if (parent) {
    t->lx1 = (parent->lx1 << 1) + ox1; t->lx2 = (parent->lx2 << 1) + ox2;
    t->lx3 = (parent->lx3 << 1) + ox3; t->level = parent->level + 1;
  }
  t->gid = -1;
  return t;
}

static void tree_free(TNode *t) {
  if (!t) return;
  for (int n = 0; n < 8; ++n) tree_free(t->leaf[n]);
  free(t);
}

static int tree_nleaf(const AoMesh *m) { return m->f3 ? 8 : (m->f2 ? 4 : 2); }

/* MeshBlockTree::CreateRootGrid (meshblock_tree.cpp:87-110) */
static void tree_create_root(const AoMesh *m, TNode *t) {
  if (t->level == m->root_level) return;
  long levfac = 1L << (m->root_level - t->level - 1);
  t->has_leaf = 1;
  for (int n = 0; n < tree_nleaf(m); ++n) {
    int i = n & 1, j = (n >> 1) & 1, k = (n >> 2) & 1;
    if ((t->lx3*2 + k)*levfac < m->nrbx3 && (t->lx2*2 + j)*levfac < m->nrbx2
        && (t->lx1*2 + i)*levfac < m->nrbx1) {
      t->leaf[n] = tnode_new(t, i, j, k);
      tree_create_root(m, t->leaf[n]);
    }
  }
}

static void tree_add(AoMesh *m, TNode *t, int level, long lx1, long lx2, long lx3);

/* MeshBlockTree::Refine (meshblock_tree.cpp:166-262): split, then make every same-level
 * neighbour location exist (keeps the 2:1 balance) */
static void tree_refine(AoMesh *m, TNode *t) {
  if (t->has_leaf) return;
  t->has_leaf = 1;
  for (int n = 0; n < tree_nleaf(m); ++n)
    t->leaf[n] = tnode_new(t, n & 1, (n >> 1) & 1, (n >> 2) & 1);
  int dl = t->level - m->root_level;
  long nxmax = (long)m->nrbx1 << dl, nymax = m->f2 ? (long)m->nrbx2 << dl : 1,
       nzmax = m->f3 ? (long)m->nrbx3 << dl : 1;
  const int *bc = m->p.bc;
  for (int oz = (m->f3 ? -1 : 0); oz <= (m->f3 ? 1 : 0); ++oz) {
    long z = t->lx3 + oz;
    if (z < 0) { if (bc[4] != AO_BC_PERIODIC) continue; z = nzmax - 1; }
    if (z >= nzmax) { if (bc[5] != AO_BC_PERIODIC) continue; z = 0; }
    for (int oy = (m->f2 ? -1 : 0); oy <= (m->f2 ? 1 : 0); ++oy) {
      long y = t->lx2 + oy;
      if (y < 0) { if (bc[2] != AO_BC_PERIODIC) continue; y = nymax - 1; }
      if (y >= nymax) { if (bc[3] != AO_BC_PERIODIC) continue; y = 0; }
      for (int ox = -1; ox <= 1; ++ox) {
        if (ox == 0 && oy == 0 && oz == 0) continue;
        long x = t->lx1 + ox;
        if (x < 0) { if (bc[0] != AO_BC_PERIODIC) continue; x = nxmax - 1; }
        if (x >= nxmax) { if (bc[1] != AO_BC_PERIODIC) continue; x = 0; }
        tree_add(m, m->root, t->level, x, y, z);
      }
    }
  }
}

/* MeshBlockTree::AddMeshBlock (meshblock_tree.cpp:117-135) */
static void tree_add(AoMesh *m, TNode *t, int level, long lx1, long lx2, long lx3) {
  if (t->level == level) return;
  if (!t->has_leaf) tree_refine(m, t);
  int sh = level - t->level - 1;
  int n = (int)((lx1 >> sh) & 1) + ((int)((lx2 >> sh) & 1) << 1) + ((int)((lx3 >> sh) & 1) << 2);
  tree_add(m, t->leaf[n], level, lx1, lx2, lx3);
}

/* MeshBlockTree::GetMeshBlockList (meshblock_tree.cpp:336-352): depth-first = Z-order */
static void tree_list(TNode *t, TNode **list, int *count) {
  if (!t->has_leaf) {
    if (list) list[*count] = t;
    t->gid = (*count)++;
    return;
  }
  for (int n = 0; n < 8; ++n) if (t->leaf[n]) tree_list(t->leaf[n], list, count);
}

/* MeshBlockTree::FindNeighbor (meshblock_tree.cpp:360-460): the leaf at the same or the next
 * coarser level, or the node whose leaves are the finer neighbours; NULL outside the mesh */
static TNode *tree_find_neighbor(const AoMesh *m, int ll, long lx, long ly, long lz, int ox1,
                                 int ox2, int ox3) {
  const int *bc = m->p.bc;
  int dl = ll - m->root_level;
  lx += ox1; ly += ox2; lz += ox3;
  long nx = (long)m->nrbx1 << dl, ny = (long)m->nrbx2 << dl, nz = (long)m->nrbx3 << dl;
  if (lx < 0) { if (bc[0] == AO_BC_PERIODIC) lx = nx - 1; else return NULL; }
  if (lx >= nx) { if (bc[1] == AO_BC_PERIODIC) lx = 0; else return NULL; }
  if (ly < 0) { if (bc[2] == AO_BC_PERIODIC) ly = ny - 1; else return NULL; }
  if (ly >= ny) { if (bc[3] == AO_BC_PERIODIC) ly = 0; else return NULL; }
  if (lz < 0) { if (bc[4] == AO_BC_PERIODIC) lz = nz - 1; else return NULL; }
  if (lz >= nz) { if (bc[5] == AO_BC_PERIODIC) lz = 0; else return NULL; }
  TNode *bt = m->root;
  if (ll < 1) return bt;
  for (int level = 0; level < ll; ++level) {
    if (!bt->has_leaf) return bt;     /* coarser leaf (one level up in a balanced tree) */
    int sh = ll - level - 1;
    bt = bt->leaf[(int)((lx >> sh) & 1) + ((int)((ly >> sh) & 1) << 1) + ((int)((lz >> sh) & 1) << 2)];
  }
  return bt;                          /* same-level leaf, or a node with finer leaves */
}

/* ------------------------------------------------------------------ mesh construction */

static void smr_build_tree(AoMesh *m) {
  const AoParams *p = &m->p;
  int nbmax = m->nrbx1 > m->nrbx2 ? m->nrbx1 : m->nrbx2;
  if (m->nrbx3 > nbmax) nbmax = m->nrbx3;
  for (m->root_level = 0; (1 << m->root_level) < nbmax; m->root_level++) {}
  m->root = tnode_new(NULL, 0, 0, 0);
  tree_create_root(m, m->root);
  const double mmin[3] = {p->x1min, p->x2min, p->x3min}, mmax[3] = {p->x1max, p->x2max, p->x3max};
  const int nrb[3] = {m->nrbx1, m->nrbx2, m->nrbx3}, nxm[3] = {p->nx1, p->nx2, p->nx3};
  /* refinement regions (mesh.cpp:330-465): logical range at the refined level, widened to
   * even/odd pairs, every pair of blocks added */
  for (int r = 0; r < p->nref; ++r) {
    int ref_lev = p->ref_level[r], lrlev = ref_lev + m->root_level;
    long lmin[3] = {0, 0, 0}, lmax[3] = {0, 0, 0};
    for (int d = 0; d < m->ndim; ++d) {
      long lxmax = (long)nrb[d]*(1L << ref_lev);
      double rmin = p->ref[r][2*d], rmax = p->ref[r][2*d+1];
      long a, b;
      for (a = 0; a < lxmax; a++)
        if (block_edge(a + 1, (int)lxmax, mmin[d], mmax[d], m->xrat[d], nxm[d]) > rmin) break;
      for (b = a; b < lxmax; b++)
        if (block_edge(b + 1, (int)lxmax, mmin[d], mmax[d], m->xrat[d], nxm[d]) >= rmax) break;
      if (a % 2 == 1) a--;
      if (b % 2 == 0) b++;
      lmin[d] = a; lmax[d] = b;
    }
    if (m->ndim == 1) { lmin[1] = 0; lmax[1] = 1; }
    if (m->ndim <= 2) { lmin[2] = 0; lmax[2] = 1; }
    for (long k = lmin[2]; k < lmax[2]; k += 2) for (long j = lmin[1]; j < lmax[1]; j += 2)
      for (long i = lmin[0]; i < lmax[0]; i += 2) tree_add(m, m->root, lrlev, i, j, k);
  }
}

/* restricted 1-D coordinates of the MeshRefinement's coarse Coordinates (coarse_flag branch of
 * coordinates.cpp:92-160, cartesian.cpp:25-75) */
static void make_coarse_coords(int nrootmesh, int bx, int cng, long lx, double mmin, double mmax,
                               double bmin, double bmax, int cnc, int refl_in, int refl_out,
                               double **xf, double **xv) {
  *xf = dalloc(cnc + 1); *xv = dalloc(cnc);
  if (cnc == 1) { (*xf)[0] = bmin; (*xf)[1] = bmax; (*xv)[0] = 0.5*((*xf)[1] + (*xf)[0]); return; }
  int il = cng, iu = cng + bx/2 - 1;
  double *dxf = dalloc(cnc);
  double dx = (bmax - bmin)/(iu - il + 1);
  for (int i = il - cng; i <= iu + cng + 1; ++i) {
    long noffset = (long)(i - il)*2 + lx*bx;
    (*xf)[i] = uniform_gen(mesh_gen_x(noffset, nrootmesh), mmin, mmax);
  }
  (*xf)[il] = bmin; (*xf)[iu+1] = bmax;
  for (int i = il - cng; i <= iu + cng; ++i) dxf[i] = dx;
  if (refl_in) for (int i = 1; i <= cng; ++i) {
    dxf[il-i] = dxf[il+i-1]; (*xf)[il-i] = (*xf)[il-i+1] - dxf[il-i]; }
  if (refl_out) for (int i = 1; i <= cng; ++i) {
    dxf[iu+i] = dxf[iu-i+1]; (*xf)[iu+i+1] = (*xf)[iu+i] + dxf[iu+i]; }
  for (int i = il - cng; i <= iu + cng; ++i) (*xv)[i] = 0.5*((*xf)[i+1] + (*xf)[i]);
  free(dxf);
}

#define CCC(B,n,k,j,i) ((((long)(n)*(B)->cnc3 + (k))*(B)->cnc2 + (j))*(B)->cnc1 + (i))

static void smr_add_neighbor(AoBlock *B, int gid, int level, int o1, int o2, int o3, int type,
                             int fi1, int fi2) {
  Nb *nb = &B->nb[B->nnb++];
  memset(nb, 0, sizeof(*nb));
  nb->ox1 = o1; nb->ox2 = o2; nb->ox3 = o3; nb->type = type; nb->gid = gid;
  nb->level = level; nb->fi1 = fi1; nb->fi2 = fi2;
  nb->fid = -1; nb->eid = -1; nb->bufid = -1; nb->targetid = -1;
  if (type == 0) {
    if (o1 == -1) nb->fid = 0; else if (o1 == 1) nb->fid = 1;
    else if (o2 == -1) nb->fid = 2; else if (o2 == 1) nb->fid = 3;
    else if (o3 == -1) nb->fid = 4; else nb->fid = 5;
  }
}

/* BoundaryBase::SearchAndSetNeighbors (bvals_base.cpp:299-736) without buffer ids */
static void smr_search_neighbors(AoMesh *m, AoBlock *B) {
  int myfx[3] = {(int)(B->lx1 & 1), (int)(B->lx2 & 1), (int)(B->lx3 & 1)};
  int myox[3] = {myfx[0]*2 - 1, m->f2 ? myfx[1]*2 - 1 : 0, m->f3 ? myfx[2]*2 - 1 : 0};
  int nf1 = m->f2 ? 2 : 1, nf2 = m->f3 ? 2 : 1;
  for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
    B->nblevel[k][j][i] = -1;
  B->nblevel[1][1][1] = B->level;
  B->nnb = 0;
  /* faces */
  for (int d = 0; d < m->ndim; ++d) for (int n = -1; n <= 1; n += 2) {
    int o[3] = {0, 0, 0}; o[d] = n;
    TNode *t = tree_find_neighbor(m, B->level, B->lx1, B->lx2, B->lx3, o[0], o[1], o[2]);
    if (!t) continue;
    if (t->has_leaf) {
      int ff = 1 - (n + 1)/2;
      B->nblevel[o[2]+1][o[1]+1][o[0]+1] = t->level + 1;
      for (int f2 = 0; f2 < nf2; ++f2) for (int f1 = 0; f1 < nf1; ++f1) {
        int l[3];
        if (d == 0) { l[0] = ff; l[1] = f1; l[2] = f2; }
        else if (d == 1) { l[0] = f1; l[1] = ff; l[2] = f2; }
        else { l[0] = f1; l[1] = f2; l[2] = ff; }
        TNode *nf = t->leaf[l[0] + (l[1] << 1) + (l[2] << 2)];
        smr_add_neighbor(B, nf->gid, nf->level, o[0], o[1], o[2], 0, f1, f2);
      }
    } else {
      B->nblevel[o[2]+1][o[1]+1][o[0]+1] = t->level;
      smr_add_neighbor(B, t->gid, t->level, o[0], o[1], o[2], 0, 0, 0);
    }
  }
  if (!m->f2) return;
  /* edges: x1x2, x1x3, x2x3 */
  for (int e = 0; e < 3; ++e) {
    if (e > 0 && !m->f3) break;
    int da = (e == 2) ? 1 : 0, db = (e == 0) ? 1 : 2, dc = 3 - da - db;
    int nfe = (e == 0) ? nf2 : nf1;
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) {
      int o[3] = {0, 0, 0}; o[da] = n; o[db] = mm;
      TNode *t = tree_find_neighbor(m, B->level, B->lx1, B->lx2, B->lx3, o[0], o[1], o[2]);
      if (!t) continue;
      if (t->has_leaf) {
        B->nblevel[o[2]+1][o[1]+1][o[0]+1] = t->level + 1;
        for (int f1 = 0; f1 < nfe; ++f1) {
          int l[3]; l[da] = 1 - (n + 1)/2; l[db] = 1 - (mm + 1)/2; l[dc] = f1;
          TNode *nf = t->leaf[l[0] + (l[1] << 1) + (l[2] << 2)];
          smr_add_neighbor(B, nf->gid, nf->level, o[0], o[1], o[2], 1, f1, 0);
        }
      } else {
        B->nblevel[o[2]+1][o[1]+1][o[0]+1] = t->level;
        if (t->level >= B->level || (myox[da] == n && myox[db] == mm))
          smr_add_neighbor(B, t->gid, t->level, o[0], o[1], o[2], 1, 0, 0);
      }
    }
  }
  if (!m->f3) return;
  /* corners */
  for (int l = -1; l <= 1; l += 2) for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) {
    TNode *t = tree_find_neighbor(m, B->level, B->lx1, B->lx2, B->lx3, n, mm, l);
    if (!t) continue;
    if (t->has_leaf) {
      int ff1 = 1 - (n + 1)/2, ff2 = 1 - (mm + 1)/2, ff3 = 1 - (l + 1)/2;
      t = t->leaf[ff1 + (ff2 << 1) + (ff3 << 2)];
    }
    B->nblevel[l+1][mm+1][n+1] = t->level;
    if (t->level >= B->level || (myox[0] == n && myox[1] == mm && myox[2] == l))
      smr_add_neighbor(B, t->gid, t->level, n, mm, l, 2, 0, 0);
  }
}

/* ------------------------------------------------------------------ restriction / prolongation */

/* MeshRefinement::RestrictCellCenteredValues (mesh_refinement.cpp:106-176): fine u of B ->
 * coarse (B's coarse_u layout) over coarse indices [csi..cei] x [csj..cej] x [csk..cek] */
static void smr_restrict(const AoMesh *m, const AoBlock *B, const double *fine, double *coarse,
                         int nvar, int csi, int cei, int csj, int cej, int csk, int cek) {
  for (int n = 0; n < nvar; ++n) {
    if (m->f3) {
      for (int ck = csk; ck <= cek; ck++) {
        int k = (ck - B->cks)*2 + B->ks;
        for (int cj = csj; cj <= cej; cj++) {
          int j = (cj - B->cjs)*2 + B->js;
          for (int ci = csi; ci <= cei; ci++) {
            int i = (ci - B->cis)*2 + B->is;
#define VOL(kk,jj,ii) (B->dx1f[ii]*B->dx2f[jj]*B->dx3f[kk])
            double v000 = VOL(k,j,i), v010 = VOL(k,j+1,i), v001 = VOL(k,j,i+1), v011 = VOL(k,j+1,i+1);
            double v100 = VOL(k+1,j,i), v110 = VOL(k+1,j+1,i), v101 = VOL(k+1,j,i+1),
                   v111 = VOL(k+1,j+1,i+1);
            double tvol = ((v000 + v010) + (v001 + v011)) + ((v100 + v110) + (v101 + v111));
            coarse[CCC(B,n,ck,cj,ci)] =
                (((fine[CC(B,n,k,j,i)]*v000 + fine[CC(B,n,k,j+1,i)]*v010)
                  + (fine[CC(B,n,k,j,i+1)]*v001 + fine[CC(B,n,k,j+1,i+1)]*v011))
                 + ((fine[CC(B,n,k+1,j,i)]*v100 + fine[CC(B,n,k+1,j+1,i)]*v110)
                    + (fine[CC(B,n,k+1,j,i+1)]*v101 + fine[CC(B,n,k+1,j+1,i+1)]*v111)))/tvol;
          }
        }
      }
    } else if (m->f2) {
      for (int cj = csj; cj <= cej; cj++) {
        int j = (cj - B->cjs)*2 + B->js;
        for (int ci = csi; ci <= cei; ci++) {
          int i = (ci - B->cis)*2 + B->is;
          double v00 = VOL(0,j,i), v10 = VOL(0,j+1,i), v01 = VOL(0,j,i+1), v11 = VOL(0,j+1,i+1);
          double tvol = (v00 + v10) + (v01 + v11);
          coarse[CCC(B,n,0,cj,ci)] =
              ((fine[CC(B,n,0,j,i)]*v00 + fine[CC(B,n,0,j+1,i)]*v10)
               + (fine[CC(B,n,0,j,i+1)]*v01 + fine[CC(B,n,0,j+1,i+1)]*v11))/tvol;
        }
      }
    } else {
      for (int ci = csi; ci <= cei; ci++) {
        int i = (ci - B->cis)*2 + B->is;
        double v0 = VOL(0,0,i), v1 = VOL(0,0,i+1);
        double tvol = v0 + v1;
        coarse[CCC(B,n,0,0,ci)] = (fine[CC(B,n,0,0,i)]*v0 + fine[CC(B,n,0,0,i+1)]*v1)/tvol;
#undef VOL
      }
    }
  }
}

static double smr_minmod_grad(double gm, double gp) {
  return 0.5*(SIGN(gm) + SIGN(gp))*mn(fabs(gm), fabs(gp));
}

/* MeshRefinement::ProlongateCellCenteredValues (mesh_refinement.cpp:386-540) */
static void smr_prolongate(const AoMesh *m, const AoBlock *B, const double *coarse, double *fine,
                           int nvar, int si, int ei, int sj, int ej, int sk, int ek) {
  for (int n = 0; n < nvar; ++n)
    for (int k = sk; k <= ek; k++) {
      int fk = m->f3 ? (k - B->cks)*2 + B->ks : B->ks;
      double dx3m = 0, dx3p = 0, dx3fm = 0, dx3fp = 0;
      if (m->f3) {
        double x3m = B->cx3v[k-1], x3c = B->cx3v[k], x3p = B->cx3v[k+1];
        dx3m = x3c - x3m; dx3p = x3p - x3c;
        dx3fm = x3c - B->x3v[fk]; dx3fp = B->x3v[fk+1] - x3c;
      }
      for (int j = sj; j <= ej; j++) {
        int fj = m->f2 ? (j - B->cjs)*2 + B->js : B->js;
        double dx2m = 0, dx2p = 0, dx2fm = 0, dx2fp = 0;
        if (m->f2) {
          double x2m = B->cx2v[j-1], x2c = B->cx2v[j], x2p = B->cx2v[j+1];
          dx2m = x2c - x2m; dx2p = x2p - x2c;
          dx2fm = x2c - B->x2v[fj]; dx2fp = B->x2v[fj+1] - x2c;
        }
        for (int i = si; i <= ei; i++) {
          int fi = (i - B->cis)*2 + B->is;
          double x1m = B->cx1v[i-1], x1c = B->cx1v[i], x1p = B->cx1v[i+1];
          double dx1m = x1c - x1m, dx1p = x1p - x1c;
          double dx1fm = x1c - B->x1v[fi], dx1fp = B->x1v[fi+1] - x1c;
          double ccval = coarse[CCC(B,n,k,j,i)];
          double gx1c = smr_minmod_grad((ccval - coarse[CCC(B,n,k,j,i-1)])/dx1m,
                                        (coarse[CCC(B,n,k,j,i+1)] - ccval)/dx1p);
          if (m->f3) {
            double gx2c = smr_minmod_grad((ccval - coarse[CCC(B,n,k,j-1,i)])/dx2m,
                                          (coarse[CCC(B,n,k,j+1,i)] - ccval)/dx2p);
            double gx3c = smr_minmod_grad((ccval - coarse[CCC(B,n,k-1,j,i)])/dx3m,
                                          (coarse[CCC(B,n,k+1,j,i)] - ccval)/dx3p);
            fine[CC(B,n,fk  ,fj  ,fi  )] = ccval - (gx1c*dx1fm + gx2c*dx2fm + gx3c*dx3fm);
            fine[CC(B,n,fk  ,fj  ,fi+1)] = ccval + (gx1c*dx1fp - gx2c*dx2fm - gx3c*dx3fm);
            fine[CC(B,n,fk  ,fj+1,fi  )] = ccval - (gx1c*dx1fm - gx2c*dx2fp + gx3c*dx3fm);
            fine[CC(B,n,fk  ,fj+1,fi+1)] = ccval + (gx1c*dx1fp + gx2c*dx2fp - gx3c*dx3fm);
            fine[CC(B,n,fk+1,fj  ,fi  )] = ccval - (gx1c*dx1fm + gx2c*dx2fm - gx3c*dx3fp);
            fine[CC(B,n,fk+1,fj  ,fi+1)] = ccval + (gx1c*dx1fp - gx2c*dx2fm + gx3c*dx3fp);
            fine[CC(B,n,fk+1,fj+1,fi  )] = ccval - (gx1c*dx1fm - gx2c*dx2fp - gx3c*dx3fp);
            fine[CC(B,n,fk+1,fj+1,fi+1)] = ccval + (gx1c*dx1fp + gx2c*dx2fp + gx3c*dx3fp);
          } else if (m->f2) {
            double gx2c = smr_minmod_grad((ccval - coarse[CCC(B,n,k,j-1,i)])/dx2m,
                                          (coarse[CCC(B,n,k,j+1,i)] - ccval)/dx2p);
            fine[CC(B,n,fk,fj  ,fi  )] = ccval - (gx1c*dx1fm + gx2c*dx2fm);
            fine[CC(B,n,fk,fj  ,fi+1)] = ccval + (gx1c*dx1fp - gx2c*dx2fm);
            fine[CC(B,n,fk,fj+1,fi  )] = ccval - (gx1c*dx1fm - gx2c*dx2fp);
            fine[CC(B,n,fk,fj+1,fi+1)] = ccval + (gx1c*dx1fp + gx2c*dx2fp);
          } else {
            fine[CC(B,n,fk,fj,fi  )] = ccval - gx1c*dx1fm;
            fine[CC(B,n,fk,fj,fi+1)] = ccval + gx1c*dx1fp;
          }
        }
      }
    }
}

/* ------------------------------------------------------------------ ghost exchange */

static void box_copy(const double *src, long sv_s, long s2s, long s1s, int si, int sj, int sk,
                     double *dst, long sv_d, long s2d, long s1d, int di, int dj, int dk,
                     int ni, int nj, int nk, int nvar) {
  for (int n = 0; n < nvar; ++n) for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j)
    for (int i = 0; i < ni; ++i)
      dst[n*sv_d + ((long)(dk+k)*s2d + (dj+j))*s1d + (di+i)] =
          src[n*sv_s + ((long)(sk+k)*s2s + (sj+j))*s1s + (si+i)];
}

/* optional log of the transfers of one exchange (test hook: ao_smr_transfers) */
static long *g_xfer_rows = NULL; static long g_xfer_n = 0, g_xfer_cap = 0;
static void smr_log(int kind, int src, int si, int sj, int sk, int dst, int di, int dj, int dk,
                    int ni, int nj, int nk) {
  if (!g_xfer_rows) return;
  if (g_xfer_n < g_xfer_cap) {
    long *r = g_xfer_rows + 12*g_xfer_n;
    r[0] = kind; r[1] = src; r[2] = si; r[3] = sj; r[4] = sk; r[5] = dst; r[6] = di; r[7] = dj;
    r[8] = dk; r[9] = ni; r[10] = nj; r[11] = nk;
  }
  g_xfer_n++;
}

/* pass 0: pack the source box into a fresh buffer; pass 1: unpack it into the destination box */
static void smr_stage(int pass, double **buf, const double *src, long sv_s, long s2s, long s1s,
                      int si, int sj, int sk, double *dst, long sv_d, long s2d, long s1d, int di,
                      int dj, int dk, int ni, int nj, int nk, int nvar) {
  long cnt = (long)ni*nj*nk;
  if (pass == 0) {
    *buf = dalloc(cnt*nvar);
    box_copy(src, sv_s, s2s, s1s, si, sj, sk, *buf, cnt, nj, ni, 0, 0, 0, ni, nj, nk, nvar);
  } else {
    box_copy(*buf, cnt, nj, ni, 0, 0, 0, dst, sv_d, s2d, s1d, di, dj, dk, ni, nj, nk, nvar);
    free(*buf); *buf = NULL;
  }
}

/* index of B among the finer leaves of the face / edge it shares with a coarser block, as the
 * coarser block's neighbour entry would carry it (bvals_base.cpp: FindBufferID arguments) */
static void smr_my_fi(const AoMesh *m, const AoBlock *B, int o1, int o2, int o3, int *fi1, int *fi2) {
  int fx1 = (int)(B->lx1 & 1), fx2 = (int)(B->lx2 & 1), fx3 = (int)(B->lx3 & 1);
  int nz = (o1 != 0) + (o2 != 0) + (o3 != 0);
  *fi1 = 0; *fi2 = 0;
  (void)m;
  if (nz == 1) {
    if (o1 != 0) { *fi1 = fx2; *fi2 = fx3; }
    else if (o2 != 0) { *fi1 = fx1; *fi2 = fx3; }
    else { *fi1 = fx1; *fi2 = fx2; }
  } else if (nz == 2) {
    if (o3 == 0) *fi1 = fx3; else if (o2 == 0) *fi1 = fx2; else *fi1 = fx1;
  }
}

/* SendBoundaryBuffers + ReceiveBoundaryBuffers + SetBoundaries of u on a multilevel mesh, from
 * the receiver's side (bvals_cc.cpp:195-470; bvals_var.cpp:212-296) */
static void smr_exchange_cc(AoMesh *m, int scalars) {
  int ng = m->p.ng, nv = scalars ? m->p.nscalars : NHYDRO;
  if (nv <= 0) return;
  /* Phase 1 packs every message from the sender's arrays as they are now (the reference packs
   * at send time; with MeshBlocks narrower than 2*NGHOST a restricted slab reaches into the
   * sender's own ghost zones, i.e. the previous exchange's data), phase 2 unpacks them. */
  double ***stage = (double ***)calloc((size_t)m->nb, sizeof(double **));
  for (int pass = 0; pass < 2; ++pass)
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];
    long svf = (long)B->nc3*B->nc2*B->nc1, svc = (long)B->cnc3*B->cnc2*B->cnc1;
    if (pass == 0) stage[g] = (double **)calloc((size_t)B->nnb, sizeof(double *));
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      AoBlock *N = &m->blk[nb->gid];
      double *Bf = scalars ? B->s : B->u, *Bc = scalars ? B->coarse_s : B->coarse_u;
      double *Nf = scalars ? N->s : N->u, *Nc = scalars ? N->coarse_s : N->coarse_u;
      int o1 = nb->ox1, o2 = nb->ox2, o3 = nb->ox3;
      int si, ei, sj, ej, sk, ek;      /* destination box in B */
      int ti, tj, tk;                  /* source origin in N */
      if (nb->level == B->level) {
        if (o1 == 0) { si = B->is; ei = B->ie; } else if (o1 > 0) { si = B->ie + 1; ei = B->ie + ng; }
        else { si = B->is - ng; ei = B->is - 1; }
        if (o2 == 0) { sj = B->js; ej = B->je; } else if (o2 > 0) { sj = B->je + 1; ej = B->je + ng; }
        else { sj = B->js - ng; ej = B->js - 1; }
        if (o3 == 0) { sk = B->ks; ek = B->ke; } else if (o3 > 0) { sk = B->ke + 1; ek = B->ke + ng; }
        else { sk = B->ks - ng; ek = B->ks - 1; }
        /* LoadBoundaryBufferSameLevel on N with offsets (-o) */
        ti = (-o1 > 0) ? (N->ie - ng + 1) : N->is;
        tj = (-o2 > 0) ? (N->je - ng + 1) : N->js;
        tk = (-o3 > 0) ? (N->ke - ng + 1) : N->ks;
        if (pass == 0) smr_log(0, nb->gid, ti, tj, tk, g, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1);
        smr_stage(pass, &stage[g][n], Nf, svf, N->nc2, N->nc1, ti, tj, tk, Bf, svf, B->nc2,
                  B->nc1, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1, nv);
      } else if (nb->level < B->level) {
        /* SetBoundaryFromCoarser: destination in B's coarse buffer */
        int cng = B->cng;
        if (o1 == 0) { si = B->cis; ei = B->cie; if ((B->lx1 & 1) == 0) ei += cng; else si -= cng; }
        else if (o1 > 0) { si = B->cie + 1; ei = B->cie + cng; } else { si = B->cis - cng; ei = B->cis - 1; }
        if (o2 == 0) { sj = B->cjs; ej = B->cje;
          if (m->f2) { if ((B->lx2 & 1) == 0) ej += cng; else sj -= cng; } }
        else if (o2 > 0) { sj = B->cje + 1; ej = B->cje + cng; } else { sj = B->cjs - cng; ej = B->cjs - 1; }
        if (o3 == 0) { sk = B->cks; ek = B->cke;
          if (m->f3) { if ((B->lx3 & 1) == 0) ek += cng; else sk -= cng; } }
        else if (o3 > 0) { sk = B->cke + 1; ek = B->cke + cng; } else { sk = B->cks - cng; ek = B->cks - 1; }
        /* LoadBoundaryBufferToFiner on N: offsets (-o), fi1 / fi2 = where B sits */
        int fi1, fi2, p1 = -o1, p2 = -o2, p3 = -o3, cn = cng - 1;
        smr_my_fi(m, B, o1, o2, o3, &fi1, &fi2);
        int a = (p1 > 0) ? (N->ie - cn) : N->is, b = (p1 < 0) ? (N->is + cn) : N->ie;
        int c = (p2 > 0) ? (N->je - cn) : N->js, d = (p2 < 0) ? (N->js + cn) : N->je;
        int e = (p3 > 0) ? (N->ke - cn) : N->ks, f = (p3 < 0) ? (N->ks + cn) : N->ke;
        int h1 = m->p.bx1/2 - cng, h2 = m->p.bx2/2 - cng, h3 = m->p.bx3/2 - cng;
        if (p1 == 0) { if (fi1 == 1) a += h1; else b -= h1; }
        if (p2 == 0 && m->f2) {
          if (p1 != 0) { if (fi1 == 1) c += h2; else d -= h2; }
          else { if (fi2 == 1) c += h2; else d -= h2; }
        }
        if (p3 == 0 && m->f3) {
          if (p1 != 0 && p2 != 0) { if (fi1 == 1) e += h3; else f -= h3; }
          else { if (fi2 == 1) e += h3; else f -= h3; }
        }
        (void)b; (void)d; (void)f;
        if (pass == 0) smr_log(1, nb->gid, a, c, e, g, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1);
        smr_stage(pass, &stage[g][n], Nf, (long)N->nc3*N->nc2*N->nc1, N->nc2, N->nc1, a, c, e,
                  Bc, svc, B->cnc2, B->cnc1, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1, nv);
      } else {
        /* SetBoundaryFromFiner: destination in B's fine array, data = N's restricted slab */
        int fi1 = nb->fi1, fi2 = nb->fi2;
        if (o1 == 0) { si = B->is; ei = B->ie; if (fi1 == 1) si += m->p.bx1/2; else ei -= m->p.bx1/2; }
        else if (o1 > 0) { si = B->ie + 1; ei = B->ie + ng; } else { si = B->is - ng; ei = B->is - 1; }
        if (o2 == 0) { sj = B->js; ej = B->je;
          if (m->f2) {
            if (o1 != 0) { if (fi1 == 1) sj += m->p.bx2/2; else ej -= m->p.bx2/2; }
            else { if (fi2 == 1) sj += m->p.bx2/2; else ej -= m->p.bx2/2; }
          } }
        else if (o2 > 0) { sj = B->je + 1; ej = B->je + ng; } else { sj = B->js - ng; ej = B->js - 1; }
        if (o3 == 0) { sk = B->ks; ek = B->ke;
          if (m->f3) {
            if (o1 != 0 && o2 != 0) { if (fi1 == 1) sk += m->p.bx3/2; else ek -= m->p.bx3/2; }
            else { if (fi2 == 1) sk += m->p.bx3/2; else ek -= m->p.bx3/2; }
          } }
        else if (o3 > 0) { sk = B->ke + 1; ek = B->ke + ng; } else { sk = B->ks - ng; ek = B->ks - 1; }
        /* LoadBoundaryBufferToCoarser on N with offsets (-o) */
        int cn = ng - 1;
        ti = (-o1 > 0) ? (N->cie - cn) : N->cis;
        tj = (-o2 > 0) ? (N->cje - cn) : N->cjs;
        tk = (-o3 > 0) ? (N->cke - cn) : N->cks;
        if (pass == 0) {   /* the sender restricts this slab now (RestrictCellCenteredValues) */
          int te = (-o1 < 0) ? (N->cis + cn) : N->cie, ue = (-o2 < 0) ? (N->cjs + cn) : N->cje,
              ve = (-o3 < 0) ? (N->cks + cn) : N->cke;
          smr_restrict(m, N, Nf, Nc, nv, ti, te, tj, ue, tk, ve);
        }
        if (pass == 0) smr_log(2, nb->gid, ti, tj, tk, g, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1);
        smr_stage(pass, &stage[g][n], Nc, (long)N->cnc3*N->cnc2*N->cnc1, N->cnc2, N->cnc1,
                  ti, tj, tk, Bf, svf, B->nc2, B->nc1, si, sj, sk, ei-si+1, ej-sj+1, ek-sk+1, nv);
      }
    }
  }
  for (int g = 0; g < m->nb; ++g) free(stage[g]);
  free(stage);
}

/* ------------------------------------------------------------------ ProlongateBoundaries */

/* coarse-level ConservedToPrimitive (hydro): coarse_u -> coarse_w with floors written back */
static void smr_coarse_cons2prim(const AoMesh *m, AoBlock *B, int il, int iu, int jl, int ju,
                                 int kl, int ku) {
  double gm1 = m->p.gamma - 1.0, dfl = m->p.dfloor, pfl = m->p.pfloor;
  for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
    double *u_d = &B->coarse_u[CCC(B,IDN,k,j,i)];
    double u_m1 = B->coarse_u[CCC(B,IM1,k,j,i)], u_m2 = B->coarse_u[CCC(B,IM2,k,j,i)],
           u_m3 = B->coarse_u[CCC(B,IM3,k,j,i)];
    *u_d = (*u_d > dfl) ? *u_d : dfl;
    double w_d = *u_d;
    double di = 1.0/(*u_d);
    B->coarse_w[CCC(B,IDN,k,j,i)] = w_d;
    B->coarse_w[CCC(B,IVX,k,j,i)] = u_m1*di;
    B->coarse_w[CCC(B,IVY,k,j,i)] = u_m2*di;
    B->coarse_w[CCC(B,IVZ,k,j,i)] = u_m3*di;
    if (!ISO(m)) {
      double *u_e = &B->coarse_u[CCC(B,IEN,k,j,i)];
      double e_k = 0.5*di*(SQR(u_m1) + SQR(u_m2) + SQR(u_m3));
      double w_p = gm1*(*u_e - e_k);
      *u_e = (w_p > pfl) ? *u_e : ((pfl/gm1) + e_k);
      w_p = (w_p > pfl) ? w_p : pfl;
      B->coarse_w[CCC(B,IPR,k,j,i)] = w_p;
    }
  }
}

/* outflow / reflecting boundary functions on the coarse primitive buffer with ngh = 1
 * (bvals_refine.cpp:385-440; cc/outflow_cc.cpp, cc/hydro/reflect_hydro.cpp) */
static void smr_coarse_phys_bc(const AoMesh *m, AoBlock *B, int face, int il, int iu, int jl,
                               int ju, int kl, int ku) {
  int bc = B->bcs[face];
  if (bc != AO_BC_OUTFLOW && bc != AO_BC_REFLECT) return;
  int dir = face/2, outer = face & 1;
  for (int n = 0; n < NHYDRO + m->p.nscalars; ++n) {
    double sign = (bc == AO_BC_REFLECT && n == IVX + dir) ? -1.0 : 1.0;
    double *cw = B->coarse_w;
    int v = n;
    if (n >= NHYDRO) { cw = B->coarse_r; v = n - NHYDRO; sign = 1.0; }
    if (dir == 0) {
      for (int k = kl; k <= ku; ++k) for (int j = jl; j <= ju; ++j) {
        if (!outer) cw[CCC(B,v,k,j,il-1)] = sign*cw[CCC(B,v,k,j,il)];
        else cw[CCC(B,v,k,j,iu+1)] = sign*cw[CCC(B,v,k,j,iu)];
      }
    } else if (dir == 1) {
      for (int k = kl; k <= ku; ++k) for (int i = il; i <= iu; ++i) {
        if (!outer) cw[CCC(B,v,k,jl-1,i)] = sign*cw[CCC(B,v,k,jl,i)];
        else cw[CCC(B,v,k,ju+1,i)] = sign*cw[CCC(B,v,k,ju,i)];
      }
    } else {
      for (int j = jl; j <= ju; ++j) for (int i = il; i <= iu; ++i) {
        if (!outer) cw[CCC(B,v,kl-1,j,i)] = sign*cw[CCC(B,v,kl,j,i)];
        else cw[CCC(B,v,ku+1,j,i)] = sign*cw[CCC(B,v,ku,j,i)];
      }
    }
  }
}

/* BoundaryValues::ProlongateBoundaries for one block (bvals_refine.cpp:96-570, hydro) */
static void smr_prolongate_boundaries(AoMesh *m, int g) {
  AoBlock *B = &m->blk[g];
  int nv = NHYDRO, ns = m->p.nscalars;
  for (int n = 0; n < B->nnb; ++n) {
    const Nb *nb = &B->nb[n];
    if (nb->level >= B->level) continue;
    int o1 = nb->ox1, o2 = nb->ox2, o3 = nb->ox3;
    int nis = (o1 - 1 > -1) ? o1 - 1 : -1, nie = (o1 + 1 < 1) ? o1 + 1 : 1;
    int njs = 0, nje = 0, nks = 0, nke = 0;
    if (m->f2) { njs = (o2 - 1 > -1) ? o2 - 1 : -1; nje = (o2 + 1 < 1) ? o2 + 1 : 1; }
    if (m->f3) { nks = (o3 - 1 > -1) ? o3 - 1 : -1; nke = (o3 + 1 < 1) ? o3 + 1 : 1; }
    /* Step 1: RestrictGhostCellsOnSameLevel */
    for (int nk = nks; nk <= nke; nk++) for (int nj = njs; nj <= nje; nj++)
      for (int ni = nis; ni <= nie; ni++) {
        int ntype = abs(ni) + abs(nj) + abs(nk);
        if (ntype == 0 || B->nblevel[nk+1][nj+1][ni+1] != B->level) continue;
        int ris, rie, rjs, rje, rks, rke;
        if (ni == 0) { ris = B->cis; rie = B->cie; if (o1 == 1) ris = B->cie; else if (o1 == -1) rie = B->cis; }
        else if (ni == 1) { ris = B->cie + 1; rie = B->cie + 1; } else { ris = B->cis - 1; rie = B->cis - 1; }
        if (nj == 0) { rjs = B->cjs; rje = B->cje; if (o2 == 1) rjs = B->cje; else if (o2 == -1) rje = B->cjs; }
        else if (nj == 1) { rjs = B->cje + 1; rje = B->cje + 1; } else { rjs = B->cjs - 1; rje = B->cjs - 1; }
        if (nk == 0) { rks = B->cks; rke = B->cke; if (o3 == 1) rks = B->cke; else if (o3 == -1) rke = B->cks; }
        else if (nk == 1) { rks = B->cke + 1; rke = B->cke + 1; } else { rks = B->cks - 1; rke = B->cks - 1; }
        smr_log(12, g, ris, rjs, rks, g, o1, o2, o3, rie-ris+1, rje-rjs+1, rke-rks+1);
        smr_restrict(m, B, B->u, B->coarse_u, nv, ris, rie, rjs, rje, rks, rke);
        if (ns > 0) smr_restrict(m, B, B->s, B->coarse_s, ns, ris, rie, rjs, rje, rks, rke);
      }
    /* loop limits of the ghost zones on the coarse level */
    int cn = B->cng - 1, si, ei, sj, ej, sk, ek;
    if (o1 == 0) { si = B->cis; ei = B->cie; if ((B->lx1 & 1) == 0) ei += cn; else si -= cn; }
    else if (o1 > 0) { si = B->cie + 1; ei = B->cie + cn; } else { si = B->cis - cn; ei = B->cis - 1; }
    if (o2 == 0) { sj = B->cjs; ej = B->cje; if (m->f2) { if ((B->lx2 & 1) == 0) ej += cn; else sj -= cn; } }
    else if (o2 > 0) { sj = B->cje + 1; ej = B->cje + cn; } else { sj = B->cjs - cn; ej = B->cjs - 1; }
    if (o3 == 0) { sk = B->cks; ek = B->cke; if (m->f3) { if ((B->lx3 & 1) == 0) ek += cn; else sk -= cn; } }
    else if (o3 > 0) { sk = B->cke + 1; ek = B->cke + cn; } else { sk = B->cks - cn; ek = B->cks - 1; }
    /* Step 2: ApplyPhysicalBoundariesOnCoarseLevel */
    int f1m = 0, f1p = 0, f2m = 0, f2p = 0, f3m = 0, f3p = 0;
    if (o1 == 0) { if (B->nblevel[1][1][0] != -1) f1m = 1; if (B->nblevel[1][1][2] != -1) f1p = 1; }
    else { f1m = 1; f1p = 1; }
    if (m->f2) {
      if (o2 == 0) { if (B->nblevel[1][0][1] != -1) f2m = 1; if (B->nblevel[1][2][1] != -1) f2p = 1; }
      else { f2m = 1; f2p = 1; }
    }
    if (m->f3) {
      if (o3 == 0) { if (B->nblevel[0][1][1] != -1) f3m = 1; if (B->nblevel[2][1][1] != -1) f3p = 1; }
      else { f3m = 1; f3p = 1; }
    }
    smr_coarse_cons2prim(m, B, si-f1m, ei+f1p, sj-f2m, ej+f2p, sk-f3m, ek+f3p);
    /* PassiveScalarConservedToPrimitive on the coarse buffers (eos_scalars.cpp:31-60) */
    for (int v = 0; v < ns; ++v) for (int k = sk-f3m; k <= ek+f3p; ++k)
      for (int j = sj-f2m; j <= ej+f2p; ++j) for (int i = si-f1m; i <= ei+f1p; ++i) {
        double d = B->coarse_u[CCC(B,IDN,k,j,i)];
        double *s_n = &B->coarse_s[CCC(B,v,k,j,i)];
        *s_n = (*s_n < m->p.sfloor*d) ? m->p.sfloor*d : *s_n;
        B->coarse_r[CCC(B,v,k,j,i)] = *s_n/d;
      }
    if (o1 == 0) {
      if (B->bcs[0] >= 0) smr_coarse_phys_bc(m, B, 0, B->cis, B->cie, sj, ej, sk, ek);
      if (B->bcs[1] >= 0) smr_coarse_phys_bc(m, B, 1, B->cis, B->cie, sj, ej, sk, ek);
    }
    if (o2 == 0 && m->f2) {
      if (B->bcs[2] >= 0) smr_coarse_phys_bc(m, B, 2, si, ei, B->cjs, B->cje, sk, ek);
      if (B->bcs[3] >= 0) smr_coarse_phys_bc(m, B, 3, si, ei, B->cjs, B->cje, sk, ek);
    }
    if (o3 == 0 && m->f3) {
      if (B->bcs[4] >= 0) smr_coarse_phys_bc(m, B, 4, si, ei, sj, ej, B->cks, B->cke);
      if (B->bcs[5] >= 0) smr_coarse_phys_bc(m, B, 5, si, ei, sj, ej, B->cks, B->cke);
    }
    smr_log(10, g, si, sj, sk, g, si-f1m, sj-f2m, sk-f3m, ei-si+1, ej-sj+1, ek-sk+1);
    smr_log(11, g, ei+f1p, ej+f2p, ek+f3p, g, o1, o2, o3, 0, 0, 0);
    /* Step 3: ProlongateGhostCells on primitives, then PrimitiveToConserved on the fine cells */
    smr_prolongate(m, B, B->coarse_w, B->w, nv, si, ei, sj, ej, sk, ek);
    if (ns > 0) smr_prolongate(m, B, B->coarse_r, B->r, ns, si, ei, sj, ej, sk, ek);
    int fsi = (si - B->cis)*2 + B->is, fei = (ei - B->cis)*2 + B->is + 1;
    int fsj = B->js, fej = B->je, fsk = B->ks, fek = B->ke;
    if (m->f2) { fsj = (sj - B->cjs)*2 + B->js; fej = (ej - B->cjs)*2 + B->js + 1; }
    if (m->f3) { fsk = (sk - B->cks)*2 + B->ks; fek = (ek - B->cks)*2 + B->ks + 1; }
    ao_prim2cons(m, g, fsi, fei, fsj, fej, fsk, fek);
    if (ns > 0) ao_scalar_prim2cons(m, g, fsi, fei, fsj, fej, fsk, fek);
  }
}

/* ------------------------------------------------------------------ flux correction */

/* SendFluxCorrection / ReceiveFluxCorrection of the hydro fluxes (flux_correction_cc.cpp:69-290):
 * the area-weighted average of the fine fluxes replaces the coarse flux on a shared face */
static void smr_flux_correction(AoMesh *m, int scalars) {
  int nvf = scalars ? m->p.nscalars : NHYDRO;
  if (nvf <= 0) return;
  for (int g = 0; g < m->nb; ++g) {
    AoBlock *B = &m->blk[g];            /* coarse receiver */
    for (int n = 0; n < B->nnb; ++n) {
      const Nb *nb = &B->nb[n];
      if (nb->type != 0 || nb->level <= B->level) continue;
      AoBlock *N = &m->blk[nb->gid];    /* fine sender; its face towards B is the opposite one */
      if (!scalars) smr_log(20, nb->gid, nb->fid ^ 1, 0, 0, g, nb->fid, nb->fi1, nb->fi2, 0, 0, 0);
      int fid = nb->fid, sfid = fid ^ 1;
      int hx1 = m->p.bx1/2, hx2 = m->f2 ? m->p.bx2/2 : 0, hx3 = m->f3 ? m->p.bx3/2 : 0;
      for (int nn = 0; nn < nvf; ++nn) {
        if (fid < 2) {
          int i = N->is + (N->ie - N->is + 1)*sfid;
          int il = B->is + (B->ie - B->is)*fid + fid;
          int jl = B->js, kl = B->ks;
          if (nb->fi1 != 0) jl += hx2;
          if (nb->fi2 != 0) kl += hx3;
          const double *fx = scalars ? N->sflux[0] : N->flux[0]; double *bf = scalars ? B->sflux[0] : B->flux[0];
          if (m->f3) {
            for (int k = N->ks, ck = kl; k <= N->ke; k += 2, ++ck)
              for (int j = N->js, cj = jl; j <= N->je; j += 2, ++cj) {
                double amm = N->dx2f[j]*N->dx3f[k], amp = N->dx2f[j+1]*N->dx3f[k];
                double apm = N->dx2f[j]*N->dx3f[k+1], app = N->dx2f[j+1]*N->dx3f[k+1];
                double tarea = amm + amp + apm + app;
                bf[FL1(B,nn,ck,cj,il)] =
                    (fx[FL1(N,nn,k,j,i)]*amm + fx[FL1(N,nn,k,j+1,i)]*amp
                     + fx[FL1(N,nn,k+1,j,i)]*apm + fx[FL1(N,nn,k+1,j+1,i)]*app)/tarea;
              }
          } else if (m->f2) {
            int k = N->ks;
            for (int j = N->js, cj = jl; j <= N->je; j += 2, ++cj) {
              double am = N->dx2f[j]*N->dx3f[k], ap = N->dx2f[j+1]*N->dx3f[k];
              double tarea = am + ap;
              bf[FL1(B,nn,B->ks,cj,il)] =
                  (fx[FL1(N,nn,k,j,i)]*am + fx[FL1(N,nn,k,j+1,i)]*ap)/tarea;
            }
          } else {
            bf[FL1(B,nn,B->ks,B->js,il)] = fx[FL1(N,nn,N->ks,N->js,i)];
          }
        } else if (fid < 4) {
          int j = N->js + (N->je - N->js + 1)*(sfid & 1);
          int jl = B->js + (B->je - B->js)*(fid & 1) + (fid & 1);
          int il = B->is, kl = B->ks;
          if (nb->fi1 != 0) il += hx1;
          if (nb->fi2 != 0) kl += hx3;
          const double *fx = scalars ? N->sflux[1] : N->flux[1]; double *bf = scalars ? B->sflux[1] : B->flux[1];
          if (m->f3) {
            for (int k = N->ks, ck = kl; k <= N->ke; k += 2, ++ck)
              for (int i = N->is, ci = il; i <= N->ie; i += 2, ++ci) {
                double a00 = N->dx1f[i]*N->dx3f[k], a01 = N->dx1f[i+1]*N->dx3f[k];
                double a10 = N->dx1f[i]*N->dx3f[k+1], a11 = N->dx1f[i+1]*N->dx3f[k+1];
                double tarea = a00 + a01 + a10 + a11;
                bf[FL2(B,nn,ck,jl,ci)] =
                    (fx[FL2(N,nn,k,j,i)]*a00 + fx[FL2(N,nn,k,j,i+1)]*a01
                     + fx[FL2(N,nn,k+1,j,i)]*a10 + fx[FL2(N,nn,k+1,j,i+1)]*a11)/tarea;
              }
          } else {
            int k = N->ks;
            for (int i = N->is, ci = il; i <= N->ie; i += 2, ++ci) {
              double a0 = N->dx1f[i]*N->dx3f[k], a1 = N->dx1f[i+1]*N->dx3f[k];
              double tarea = a0 + a1;
              bf[FL2(B,nn,B->ks,jl,ci)] =
                  (fx[FL2(N,nn,k,j,i)]*a0 + fx[FL2(N,nn,k,j,i+1)]*a1)/tarea;
            }
          }
        } else {
          int k = N->ks + (N->ke - N->ks + 1)*(sfid & 1);
          int kl = B->ks + (B->ke - B->ks)*(fid & 1) + (fid & 1);
          int il = B->is, jl = B->js;
          if (nb->fi1 != 0) il += hx1;
          if (nb->fi2 != 0) jl += hx2;
          const double *fx = scalars ? N->sflux[2] : N->flux[2]; double *bf = scalars ? B->sflux[2] : B->flux[2];
          for (int j = N->js, cj = jl; j <= N->je; j += 2, ++cj)
            for (int i = N->is, ci = il; i <= N->ie; i += 2, ++ci) {
              double a00 = N->dx1f[i]*N->dx2f[j], a01 = N->dx1f[i+1]*N->dx2f[j];
              double a10 = N->dx1f[i]*N->dx2f[j+1], a11 = N->dx1f[i+1]*N->dx2f[j+1];
              double tarea = a00 + a01 + a10 + a11;
              bf[FL3(B,nn,kl,cj,ci)] =
                  (fx[FL3(N,nn,k,j,i)]*a00 + fx[FL3(N,nn,k,j,i+1)]*a01
                   + fx[FL3(N,nn,k,j+1,i)]*a10 + fx[FL3(N,nn,k,j+1,i+1)]*a11)/tarea;
            }
        }
      }
    }
  }
}

/* test hook: the transfer list of one ghost exchange + ProlongateBoundaries + flux correction
 * (rows of 12 longs: kind, src gid, src origin i j k, dst gid, dst origin i j k, extent ni nj nk;
 * kind 0 same level fine->fine, 1 coarser fine -> coarse buffer, 2 finer's restricted slab ->
 * fine, 10/11 prolongation box + coarse cons2prim margins / offsets, 12 restriction of own
 * ghost cells, 20 flux correction {fine gid, its face, -, -, coarse gid, face, fi1, fi2}).
 * Runs on a scratch copy of nothing: it performs a real exchange on the mesh's current state. */
long ao_smr_transfers(AoMesh *m, long *rows, long max_rows) {
  if (!m->multilevel) return 0;
  static long dummy[12];
  g_xfer_rows = rows ? rows : dummy; g_xfer_n = 0; g_xfer_cap = rows ? max_rows : 0;
  smr_exchange_cc(m, 0);
  for (int g = 0; g < m->nb; ++g) smr_prolongate_boundaries(m, g);
  smr_flux_correction(m, 0);
  g_xfer_rows = NULL;
  return g_xfer_n;
}

/* test hooks: restriction of u into coarse_u / prolongation of coarse_w into w over one coarse box
 * of block b (all variables) */
void ao_smr_restrict_box(AoMesh *m, int b, const int *box) {
  AoBlock *B = &m->blk[b];
  smr_restrict(m, B, B->u, B->coarse_u, NHYDRO, box[0], box[1], box[2], box[3], box[4], box[5]);
}
void ao_smr_prolong_box(AoMesh *m, int b, const int *box) {
  AoBlock *B = &m->blk[b];
  smr_prolongate(m, B, B->coarse_w, B->w, NHYDRO, box[0], box[1], box[2], box[3], box[4], box[5]);
}

/* test hook: one SMR step on the mesh's current state: 0 ghost exchange of u and s, 1
 * ProlongateBoundaries of every block, 2 flux correction of the hydro and scalar fluxes */
void ao_smr_step(AoMesh *m, int what) {
  if (!m->multilevel) return;
  if (what == 0) { smr_exchange_cc(m, 0); smr_exchange_cc(m, 1); }
  else if (what == 1) { for (int g = 0; g < m->nb; ++g) smr_prolongate_boundaries(m, g); }
  else { smr_flux_correction(m, 0); smr_flux_correction(m, 1); }
}
