#ifndef SHIM_B200_BRIDGE_HPP_
#define SHIM_B200_BRIDGE_HPP_
// b200_bridge.hpp -- the reference-side binding of libathena_b200 (include/athena_b200.h).
//
// These files are compiled INTO the reference (Athena++ fork, /root/reference) in place of two
// of its translation units:
//     src/task_list/time_integrator.cpp  ->  shim/b200_time_integrator.cpp
//     src/hydro/new_blockdt.cpp          ->  shim/b200_new_blockdt.cpp
// Everything else of the reference is compiled unchanged: main(), ParameterInput, Mesh and
// MeshBlock construction, the C++ problem generators, Mesh::Initialize, the polling task
// scheduler (task_list/task_list.cpp), Mesh::NewTimeStep, outputs.  The task bodies of the
// TimeIntegratorTaskList call the C ABI; the state lives on the GPU between outputs.
//
// Host/device coherence:
//   * Mesh::Initialize runs on the host as in the reference and ends with
//     Hydro::NewBlockTimeStep for every block (mesh.cpp:1640-1644); the shim's version of that
//     function creates the device mesh on its first call and uploads the block's u, b (with
//     ghost zones, as Initialize left them) before it computes the block's dt on the device.
//   * The last stage's USERWORK task downloads u, w, b, bcc (s, r) into the block's own
//     AthenaArrays when an output is due after this cycle, when the run ends with this cycle,
//     or always with <b200> sync_every_cycle = true (needed by a pgen whose UserWorkInLoop
//     reads the arrays).
#include <athena_b200.h>

#include <mutex>
#include <string>
#include <vector>

#include "athena.hpp"

class Mesh;
class MeshBlock;
class ParameterInput;

namespace b200 {

struct Bridge {
  AbMesh *mesh = nullptr;
  Mesh *pm = nullptr;
  ParameterInput *pin = nullptr;
  std::string integrator;
  bool sync_every_cycle = false;
  std::vector<char> uploaded;         // per local block
  std::mutex mu;
  BValFunc user_bc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  SrcTermFunc user_src = nullptr;

  static Bridge &Get();
  // non-zero return code of the C ABI -> ATHENA_ERROR (the reference's error path)
  static void Check(int rc, const char *where);
  // called once, from the first Hydro::NewBlockTimeStep (it alone may read Mesh's private
  // user-function tables): creates the device mesh from the reference's own objects
  void Create(MeshBlock *pmb, const BValFunc bc[6], SrcTermFunc src, const bool uniform_gen[3]);
  void Upload(MeshBlock *pmb);
  void Download(MeshBlock *pmb);
  // Outputs::MakeOutputs's conditions (outputs/outputs.cpp:786-803) evaluated for the state
  // after this cycle, plus the end-of-run conditions of main.cpp:430
  bool HostStateNeededAfterThisCycle() const;
};

}  // namespace b200
#endif  // SHIM_B200_BRIDGE_HPP_
