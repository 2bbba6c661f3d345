// b200_new_blockdt.cpp -- replaces src/hydro/new_blockdt.cpp of the reference in a B200 build:
// Hydro::NewBlockTimeStep (hydro/new_blockdt.cpp:42-190) computed on the device.
//
// It is also where the device mesh comes to life.  Mesh::Initialize finishes on the host, as in
// the reference, by calling this function for every MeshBlock (mesh.cpp:1640-1644, inside an
// OpenMP loop): the first call creates the device mesh from the reference's own objects, every
// block's first call uploads that block's state with the ghost zones Initialize has filled.
// Hydro is a friend of Mesh and MeshBlock, so the user-function tables the library needs
// (Mesh::BoundaryFunction_, UserSourceTerm_) and MeshBlock::new_block_dt_ are reachable here
// exactly as they are in the reference's version of this function.
#include <algorithm>
#include <limits>
#include <mutex>

#include "athena.hpp"
#include "athena_arrays.hpp"
#include "field/field.hpp"
#include "hydro/hydro.hpp"
#include "mesh/mesh.hpp"
#include "scalars/scalars.hpp"

#include "b200_bridge.hpp"

void Hydro::NewBlockTimeStep() {
  MeshBlock *pmb = pmy_block;
  Mesh *pm = pmb->pmy_mesh;
  b200::Bridge &B = b200::Bridge::Get();
  {
    std::lock_guard<std::mutex> lock(B.mu);
    if (B.mesh == nullptr)
      B.Create(pmb, pm->BoundaryFunction_, pm->UserSourceTerm_, pm->use_uniform_meshgen_fn_);
    if (!B.uploaded[pmb->lid]) B.Upload(pmb);
  }
  const Real real_max = std::numeric_limits<Real>::max();
  double dt_hyp = real_max;
  b200::Bridge::Check(ab_new_block_dt(B.mesh, pmb->lid, &dt_hyp), "Hydro::NewBlockTimeStep");
  Real min_dt = std::min(real_max, static_cast<Real>(dt_hyp));
  Real min_dt_user = real_max;
  if (UserTimeStep_ != nullptr) {      // a host function of the pgen: give it current arrays
    B.Download(pmb);
    min_dt_user = UserTimeStep_(pmb);
    min_dt = std::min(min_dt, min_dt_user);
  }
  pmb->new_block_dt_ = min_dt;
  pmb->new_block_dt_hyperbolic_ = dt_hyp;
  pmb->new_block_dt_parabolic_ = real_max*pm->cfl_number;
  pmb->new_block_dt_user_ = min_dt_user;
}
