// b200_bridge.cpp -- see b200_bridge.hpp.  Written against the reference's public headers
// (mesh/mesh.hpp, hydro/hydro.hpp, field/field.hpp, scalars/scalars.hpp, eos/eos.hpp,
// reconstruct/reconstruction.hpp, parameter_input.hpp); no reference source is modified.
#include "b200_bridge.hpp"

#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "athena_arrays.hpp"
#include "bvals/bvals.hpp"
#include "coordinates/coordinates.hpp"
#include "eos/eos.hpp"
#include "field/field.hpp"
#include "field/field_diffusion/field_diffusion.hpp"
#include "hydro/hydro.hpp"
#include "hydro/hydro_diffusion/hydro_diffusion.hpp"
#include "hydro/srcterms/hydro_srcterms.hpp"
#include "mesh/mesh.hpp"
#include "parameter_input.hpp"
#include "reconstruct/reconstruction.hpp"
#include "scalars/scalars.hpp"

namespace b200 {

Bridge &Bridge::Get() {
  static Bridge b;
  return b;
}

void Bridge::Check(int rc, const char *where) {
  if (rc >= 0) return;
  std::stringstream msg;
  msg << "### FATAL ERROR in " << where << std::endl
      << "libathena_b200 error " << rc << ": " << ab_last_error() << std::endl;
  ATHENA_ERROR(msg);
}

namespace {

void Unsupported(const char *what) {
  std::stringstream msg;
  msg << "### FATAL ERROR in b200::Bridge::Create" << std::endl
      << what << " is not on the B200 path (see DESIGN.md, out of scope)" << std::endl;
  ATHENA_ERROR(msg);
}

int SolverId() {
  const std::string s = RIEMANN_SOLVER;
  if (s == "hlle") return AB_SOLVER_HLLE;
  if (s == "hllc") return AB_SOLVER_HLLC;
  if (s == "hlld") return AB_SOLVER_HLLD;
  if (s == "roe") return AB_SOLVER_ROE;
  if (s == "lhllc") return AB_SOLVER_LHLLC;
  if (s == "lhlld") return AB_SOLVER_LHLLD;
  if (s == "llf") return AB_SOLVER_LLF;
  Unsupported(("Riemann solver " + s).c_str());
  return -1;
}

int BcId(BoundaryFlag f) {
  switch (f) {
    case BoundaryFlag::periodic: return AB_BC_PERIODIC;
    case BoundaryFlag::outflow: return AB_BC_OUTFLOW;
    case BoundaryFlag::reflect: return AB_BC_REFLECT;
    case BoundaryFlag::user: return AB_BC_USER;
    case BoundaryFlag::undef: return AB_BC_PERIODIC;     // direction not present (nx = 1)
    default: Unsupported("this boundary flag (polar / shear_periodic / block)");
  }
  return -1;
}

// The library hands the user hooks host staging arrays in AthenaArray layout; the reference's
// hook signatures take the block's AthenaArrays.  The staging data is moved through the
// block's own arrays around the call.
void BcTrampoline(void *user, int lid, double *prim, double *b1, double *b2, double *b3,
                  double time, double dt, int il, int iu, int jl, int ju, int kl, int ku,
                  int ngh) {
  const int face = static_cast<int>(reinterpret_cast<std::intptr_t>(user));
  Bridge &B = Bridge::Get();
  MeshBlock *pmb = B.pm->my_blocks(lid);
  Hydro *ph = pmb->phydro;
  Field *pf = pmb->pfield;
  std::memcpy(ph->w.data(), prim, sizeof(Real)*ph->w.GetSize());
  if (MAGNETIC_FIELDS_ENABLED) {
    std::memcpy(pf->b.x1f.data(), b1, sizeof(Real)*pf->b.x1f.GetSize());
    std::memcpy(pf->b.x2f.data(), b2, sizeof(Real)*pf->b.x2f.GetSize());
    std::memcpy(pf->b.x3f.data(), b3, sizeof(Real)*pf->b.x3f.GetSize());
  }
  B.user_bc[face](pmb, pmb->pcoord, ph->w, pf->b, time, dt, il, iu, jl, ju, kl, ku, ngh);
  std::memcpy(prim, ph->w.data(), sizeof(Real)*ph->w.GetSize());
  if (MAGNETIC_FIELDS_ENABLED) {
    std::memcpy(b1, pf->b.x1f.data(), sizeof(Real)*pf->b.x1f.GetSize());
    std::memcpy(b2, pf->b.x2f.data(), sizeof(Real)*pf->b.x2f.GetSize());
    std::memcpy(b3, pf->b.x3f.data(), sizeof(Real)*pf->b.x3f.GetSize());
  }
}

void SrcTrampoline(void *, int lid, double time, double dt, const double *prim,
                   const double *prim_scalar, const double *bcc, double *cons,
                   double *cons_scalar) {
  Bridge &B = Bridge::Get();
  MeshBlock *pmb = B.pm->my_blocks(lid);
  Hydro *ph = pmb->phydro;
  Field *pf = pmb->pfield;
  PassiveScalars *ps = pmb->pscalars;
  std::memcpy(ph->w.data(), prim, sizeof(Real)*ph->w.GetSize());
  std::memcpy(ph->u.data(), cons, sizeof(Real)*ph->u.GetSize());
  if (MAGNETIC_FIELDS_ENABLED && bcc)
    std::memcpy(pf->bcc.data(), bcc, sizeof(Real)*pf->bcc.GetSize());
  if (NSCALARS > 0) {
    std::memcpy(ps->r.data(), prim_scalar, sizeof(Real)*ps->r.GetSize());
    std::memcpy(ps->s.data(), cons_scalar, sizeof(Real)*ps->s.GetSize());
  }
  B.user_src(pmb, time, dt, ph->w, ps->r, pf->bcc, ph->u, ps->s);
  std::memcpy(cons, ph->u.data(), sizeof(Real)*ph->u.GetSize());
  if (NSCALARS > 0) std::memcpy(cons_scalar, ps->s.data(), sizeof(Real)*ps->s.GetSize());
}

}  // namespace

void Bridge::Create(MeshBlock *pmb0, const BValFunc bc[6], SrcTermFunc src,
                    const bool uniform_gen[3]) {
  Mesh *m = pmb0->pmy_mesh;
  pm = m;
  if (m->multilevel) Unsupported("mesh refinement through the reference-side shim");
  if (Globals::nranks > 1) Unsupported("MPI (add the ab_comm_unique_id broadcast of INTEGRATION.md)");
  if (RELATIVISTIC_DYNAMICS || GENERAL_EOS || STS_ENABLED || SELF_GRAVITY_ENABLED)
    Unsupported("relativity / general EOS / super-time-stepping / self-gravity");
  if (std::strcmp(COORDINATE_SYSTEM, "cartesian") != 0) Unsupported("non-Cartesian coordinates");
  if (m->orbital_advection != 0 || m->shear_periodic) Unsupported("orbital advection / shearing box");
  if (pmb0->phydro->hdif.hydro_diffusion_defined) Unsupported("hydro diffusion");
  if (MAGNETIC_FIELDS_ENABLED && pmb0->pfield->fdif.field_diffusion_defined)
    Unsupported("field diffusion");
  AbMeshParams p;
  std::memset(&p, 0, sizeof(p));
  const RegionSize &ms = m->mesh_size, &bs = pmb0->block_size;
  p.nx1 = ms.nx1; p.nx2 = ms.nx2; p.nx3 = ms.nx3;
  p.bx1 = bs.nx1; p.bx2 = bs.nx2; p.bx3 = bs.nx3;
  p.x1min = ms.x1min; p.x1max = ms.x1max; p.x2min = ms.x2min; p.x2max = ms.x2max;
  p.x3min = ms.x3min; p.x3max = ms.x3max;
  p.xrat[0] = ms.x1rat; p.xrat[1] = ms.x2rat; p.xrat[2] = ms.x3rat;
  for (int d = 0; d < 3; ++d)
    if (!uniform_gen[d] && p.xrat[d] == 1.0) Unsupported("a user-enrolled mesh generator");
  for (int f = 0; f < 6; ++f) {
    p.bc[f] = BcId(m->mesh_bcs[f]);
    user_bc[f] = bc[f];
  }
  user_src = src;
  p.nghost = NGHOST;
  p.mhd = MAGNETIC_FIELDS_ENABLED;
  p.solver = SolverId();
  p.xorder = pmb0->precon->xorder;
  p.char_proj = pmb0->precon->characteristic_projection ? 1 : 0;
  if (integrator == "vl2") p.integrator = AB_INT_VL2;
  else if (integrator == "rk1") p.integrator = AB_INT_RK1;
  else if (integrator == "rk2") p.integrator = AB_INT_RK2;
  else if (integrator == "rk3") p.integrator = AB_INT_RK3;
  else Unsupported(("time/integrator = " + integrator).c_str());
  EquationOfState *peos = pmb0->peos;
  p.eos = NON_BAROTROPIC_EOS ? AB_EOS_ADIABATIC : AB_EOS_ISOTHERMAL;
  p.gamma = NON_BAROTROPIC_EOS ? peos->GetGamma() : 0.0;
  p.iso_sound_speed = NON_BAROTROPIC_EOS ? 0.0 : peos->GetIsoSoundSpeed();
  p.dfloor = peos->GetDensityFloor();
  p.pfloor = peos->GetPressureFloor();
  p.sfloor = pin->GetOrAddReal("hydro", "sfloor", 0.0);      // eos ctor has added the default
  p.cfl_number = m->cfl_number;
  p.tlim = m->tlim;
  p.start_time = m->time;
  p.rank = 0; p.nranks = 1;
  p.device = pin->GetOrAddInteger("b200", "device", 0);
  p.nscalars = NSCALARS;
  p.grav_acc[0] = pin->GetOrAddReal("hydro", "grav_acc1", 0.0);
  p.grav_acc[1] = pin->GetOrAddReal("hydro", "grav_acc2", 0.0);
  p.grav_acc[2] = pin->GetOrAddReal("hydro", "grav_acc3", 0.0);
  sync_every_cycle = pin->GetOrAddBoolean("b200", "sync_every_cycle", false);
  Check(ab_mesh_create(&p, &mesh), "b200::Bridge::Create (ab_mesh_create)");
  if (ab_mesh_nblocks_local(mesh) != m->nblocal) Unsupported("a MeshBlock list that differs from Mesh's");
  // the library orders blocks as MeshBlockTree::GetMeshBlockList does; check block 0 .. n-1
  for (int i = 0; i < m->nblocal; ++i) {
    long info[16] = {0};
    Check(ab_block_info(mesh, i, info), "b200::Bridge::Create (ab_block_info)");
    const LogicalLocation &loc = m->my_blocks(i)->loc;
    if (info[1] != loc.lx1 || info[2] != loc.lx2 || info[3] != loc.lx3)
      Unsupported("a MeshBlock order that differs from Mesh's");
  }
  for (int f = 0; f < 6; ++f)
    if (p.bc[f] == AB_BC_USER && m->mesh_size.nx1 > 0) {
      const bool present = (f < 2) || (f < 4 && m->f2) || m->f3;
      if (!present) continue;
      if (user_bc[f] == nullptr) Unsupported("a user boundary flag without an enrolled function");
      Check(ab_enroll_user_boundary_function(mesh, f, BcTrampoline,
                                             reinterpret_cast<void *>(static_cast<std::intptr_t>(f))),
            "b200::Bridge::Create (ab_enroll_user_boundary_function)");
    }
  if (user_src != nullptr)
    Check(ab_enroll_user_explicit_source_function(mesh, SrcTrampoline, nullptr),
          "b200::Bridge::Create (ab_enroll_user_explicit_source_function)");
  uploaded.assign(m->nblocal, 0);
  if (Globals::my_rank == 0)
    std::cout << "[b200] device mesh created: " << m->nbtotal << " MeshBlocks of " << bs.nx1
              << "x" << bs.nx2 << "x" << bs.nx3 << " on CUDA device " << p.device << std::endl;
}

void Bridge::Upload(MeshBlock *pmb) {
  const int lid = pmb->lid;
  Hydro *ph = pmb->phydro;
  Check(ab_upload(mesh, lid, AB_U, ph->u.data()), "b200::Bridge::Upload (u)");
  if (MAGNETIC_FIELDS_ENABLED) {
    Field *pf = pmb->pfield;
    Check(ab_upload(mesh, lid, AB_B_X1F, pf->b.x1f.data()), "b200::Bridge::Upload (b.x1f)");
    Check(ab_upload(mesh, lid, AB_B_X2F, pf->b.x2f.data()), "b200::Bridge::Upload (b.x2f)");
    Check(ab_upload(mesh, lid, AB_B_X3F, pf->b.x3f.data()), "b200::Bridge::Upload (b.x3f)");
  }
  if (NSCALARS > 0)
    Check(ab_upload(mesh, lid, AB_S, pmb->pscalars->s.data()), "b200::Bridge::Upload (s)");
  // w, bcc (r): Mesh::Initialize has produced them on the host from exactly these u, b with
  // the same arithmetic; they are uploaded too so that the device starts from the host's bits
  Check(ab_upload(mesh, lid, AB_W, ph->w.data()), "b200::Bridge::Upload (w)");
  if (MAGNETIC_FIELDS_ENABLED)
    Check(ab_upload(mesh, lid, AB_BCC, pmb->pfield->bcc.data()), "b200::Bridge::Upload (bcc)");
  if (NSCALARS > 0)
    Check(ab_upload(mesh, lid, AB_R, pmb->pscalars->r.data()), "b200::Bridge::Upload (r)");
  uploaded[lid] = 1;
}

void Bridge::Download(MeshBlock *pmb) {
  const int lid = pmb->lid;
  Hydro *ph = pmb->phydro;
  Check(ab_download(mesh, lid, AB_U, ph->u.data()), "b200::Bridge::Download (u)");
  Check(ab_download(mesh, lid, AB_W, ph->w.data()), "b200::Bridge::Download (w)");
  if (MAGNETIC_FIELDS_ENABLED) {
    Field *pf = pmb->pfield;
    Check(ab_download(mesh, lid, AB_B_X1F, pf->b.x1f.data()), "b200::Bridge::Download (b.x1f)");
    Check(ab_download(mesh, lid, AB_B_X2F, pf->b.x2f.data()), "b200::Bridge::Download (b.x2f)");
    Check(ab_download(mesh, lid, AB_B_X3F, pf->b.x3f.data()), "b200::Bridge::Download (b.x3f)");
    Check(ab_download(mesh, lid, AB_BCC, pf->bcc.data()), "b200::Bridge::Download (bcc)");
  }
  if (NSCALARS > 0) {
    Check(ab_download(mesh, lid, AB_S, pmb->pscalars->s.data()), "b200::Bridge::Download (s)");
    Check(ab_download(mesh, lid, AB_R, pmb->pscalars->r.data()), "b200::Bridge::Download (r)");
  }
}

bool Bridge::HostStateNeededAfterThisCycle() const {
  if (sync_every_cycle) return true;
  const Real t_new = pm->time + pm->dt;
  const int ncycle_new = pm->ncycle + 1;
  if (t_new >= pm->tlim) return true;                              // main.cpp:430, final outputs
  if (pm->nlim >= 0 && ncycle_new >= pm->nlim) return true;
  for (InputBlock *pib = pin->pfirst_block; pib != nullptr; pib = pib->pnext) {
    if (pib->block_name.compare(0, 6, "output") != 0) continue;
    const std::string &name = pib->block_name;
    const Real odt = pin->DoesParameterExist(name, "dt") ? pin->GetReal(name, "dt") : 0.0;
    const int dcycle = pin->DoesParameterExist(name, "dcycle") ? pin->GetInteger(name, "dcycle") : 0;
    if (odt > 0.0 && pin->DoesParameterExist(name, "next_time")
        && t_new >= pin->GetReal(name, "next_time")) return true;   // outputs.cpp:791
    if (dcycle > 0 && ncycle_new % dcycle == 0) return true;        // outputs.cpp:792
  }
  return false;
}

}  // namespace b200
