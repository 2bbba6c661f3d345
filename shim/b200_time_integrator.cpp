// b200_time_integrator.cpp -- replaces src/task_list/time_integrator.cpp of the reference in a
// B200 build.  The TimeIntegratorTaskList keeps its declaration (task_list/task_list.hpp), its
// stage structure, its task ids and dependencies (time_integrator.cpp:899-1098) and is run by
// the reference's own scheduler (TaskList::DoTaskListOneStage, task_list/task_list.cpp:66-91,
// OpenMP loop over MeshBlocks included); the body of every task is a call into the C ABI of
// libathena_b200 for that MeshBlock.  Host AthenaArrays are refreshed from the device only when
// the host needs them (b200_bridge.hpp).
//
// In scope: one-level meshes; vl2 / rk1 / rk2 / rk3; hydro, MHD, passive scalars; periodic,
// outflow, reflecting and user-enrolled boundaries; constant-acceleration and user-enrolled
// source terms.  Everything else is rejected with the reference's error mechanism.
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "athena.hpp"
#include "field/field.hpp"
#include "hydro/hydro.hpp"
#include "hydro/srcterms/hydro_srcterms.hpp"
#include "mesh/mesh.hpp"
#include "parameter_input.hpp"
#include "reconstruct/reconstruction.hpp"
#include "scalars/scalars.hpp"
#include "task_list/task_list.hpp"

#include "b200_bridge.hpp"

using b200::Bridge;

namespace {
inline AbMesh *Dev() { return Bridge::Get().mesh; }
inline void Ck(int rc, const char *where) { Bridge::Check(rc, where); }
}  // namespace

//----------------------------------------------------------------------------------------
// constructor: integrator weights (time_integrator.cpp:199-235 vl2, :237-262 rk1, :425-458
// rk2, :462-520 rk3; start/end times of the stages :863-877), CFL limit (:884-895), task graph

TimeIntegratorTaskList::TimeIntegratorTaskList(ParameterInput *pin, Mesh *pm) {
  integrator = pin->GetOrAddString("time", "integrator", "vl2");
  ORBITAL_ADVECTION = (pm->orbital_advection != 0);
  SHEAR_PERIODIC = pm->shear_periodic;
  if (ORBITAL_ADVECTION || SHEAR_PERIODIC || STS_ENABLED || pm->multilevel) {
    std::stringstream msg;
    msg << "### FATAL ERROR in TimeIntegratorTaskList constructor (B200 build)" << std::endl
        << "orbital advection, shearing box, super-time-stepping and mesh refinement are "
        << "not on the B200 path" << std::endl;
    ATHENA_ERROR(msg);
  }
  for (int l = 0; l < MAX_NSTAGE; ++l) {
    stage_wghts[l].delta = 0.0; stage_wghts[l].gamma_1 = 0.0; stage_wghts[l].gamma_2 = 1.0;
    stage_wghts[l].gamma_3 = 0.0; stage_wghts[l].beta = 0.0;
    stage_wghts[l].sbeta = 0.0; stage_wghts[l].ebeta = 1.0;
    stage_wghts[l].main_stage = true; stage_wghts[l].orbital_stage = false;
  }
  stage_wghts[0].delta = 1.0;
  if (integrator == "vl2") {
    nstages = nstages_main = 2;
    cfl_limit = (pm->ndim == 1) ? 1.0 : 0.5;
    stage_wghts[0].beta = 0.5; stage_wghts[0].ebeta = 0.5;
    stage_wghts[1].beta = 1.0; stage_wghts[1].sbeta = 0.5;
  } else if (integrator == "rk1") {
    nstages = nstages_main = 1;
    cfl_limit = 1.0;
    stage_wghts[0].beta = 1.0;
  } else if (integrator == "rk2") {
    nstages = nstages_main = 2;
    cfl_limit = 1.0;
    stage_wghts[0].beta = 1.0;
    stage_wghts[1].gamma_1 = 0.5; stage_wghts[1].gamma_2 = 0.5; stage_wghts[1].beta = 0.5;
    stage_wghts[1].sbeta = 1.0;
  } else if (integrator == "rk3") {
    nstages = nstages_main = 3;
    cfl_limit = 1.0;
    stage_wghts[0].beta = 1.0;
    stage_wghts[1].gamma_1 = 0.25; stage_wghts[1].gamma_2 = 0.75; stage_wghts[1].beta = 0.25;
    stage_wghts[1].sbeta = 1.0; stage_wghts[1].ebeta = 0.5;
    stage_wghts[2].gamma_1 = TWO_3RD; stage_wghts[2].gamma_2 = ONE_3RD;
    stage_wghts[2].beta = TWO_3RD; stage_wghts[2].sbeta = 0.5;
  } else {
    std::stringstream msg;
    msg << "### FATAL ERROR in TimeIntegratorTaskList constructor (B200 build)" << std::endl
        << "integrator=" << integrator << " is not on the B200 path (vl2, rk1, rk2, rk3)"
        << std::endl;
    ATHENA_ERROR(msg);
  }
  Real cfl_number = pin->GetReal("time", "cfl_number");
  if (cfl_number > cfl_limit && pm->fluid_setup == FluidFormulation::evolve) {
    std::cout << "### Warning in TimeIntegratorTaskList constructor" << std::endl
              << "User CFL number " << cfl_number << " must be smaller than " << cfl_limit
              << " for integrator=" << integrator << " in " << pm->ndim
              << "D simulation" << std::endl << "Setting to limit" << std::endl;
    cfl_number = cfl_limit;
  }
  pm->cfl_number = cfl_number;

  Bridge &B = Bridge::Get();
  B.pin = pin;
  B.pm = pm;
  B.integrator = integrator;

  // the reference's graph for a one-level mesh without diffusion / STS / orbital / shear
  {using namespace HydroIntegratorTaskNames; // NOLINT (build/namespace)
    AddTask(CALC_HYDFLX, NONE);
    if (NSCALARS > 0) AddTask(CALC_SCLRFLX, CALC_HYDFLX);
    AddTask(INT_HYD, CALC_HYDFLX);
    if (NSCALARS > 0) AddTask(SRC_TERM, (INT_HYD|INT_SCLR));
    else AddTask(SRC_TERM, INT_HYD);
    AddTask(SEND_HYD, SRC_TERM);
    AddTask(RECV_HYD, NONE);
    AddTask(SETB_HYD, (RECV_HYD|SRC_TERM));
    if (NSCALARS > 0) {
      AddTask(INT_SCLR, CALC_SCLRFLX);
      AddTask(SEND_SCLR, SRC_TERM);
      AddTask(RECV_SCLR, NONE);
      AddTask(SETB_SCLR, (RECV_SCLR|SRC_TERM));
    }
    if (MAGNETIC_FIELDS_ENABLED) {
      AddTask(CALC_FLDFLX, CALC_HYDFLX);
      AddTask(SEND_FLDFLX, CALC_FLDFLX);
      AddTask(RECV_FLDFLX, SEND_FLDFLX);
      AddTask(INT_FLD, RECV_FLDFLX);
      AddTask(SEND_FLD, INT_FLD);
      AddTask(RECV_FLD, NONE);
      AddTask(SETB_FLD, (RECV_FLD|INT_FLD));
      if (NSCALARS > 0) AddTask(CONS2PRIM, (SETB_HYD|SETB_FLD|SETB_SCLR));
      else AddTask(CONS2PRIM, (SETB_HYD|SETB_FLD));
    } else {
      if (NSCALARS > 0) AddTask(CONS2PRIM, (SETB_HYD|SETB_SCLR));
      else AddTask(CONS2PRIM, SETB_HYD);
    }
    AddTask(PHY_BVAL, CONS2PRIM);
    AddTask(USERWORK, PHY_BVAL);
    AddTask(NEW_DT, USERWORK);
    AddTask(CLEAR_ALLBND, NEW_DT);
  }
}

//----------------------------------------------------------------------------------------
// AddTask: task id -> member function (time_integrator.cpp:1105-1378)

void TimeIntegratorTaskList::AddTask(const TaskID &id, const TaskID &dep) {
  using namespace HydroIntegratorTaskNames; // NOLINT (build/namespace)
  typedef TaskStatus (TaskList::*Fn)(MeshBlock *, int);
  struct Row { const TaskID *id; TaskStatus (TimeIntegratorTaskList::*fn)(MeshBlock *, int); bool lb; };
  const Row table[] = {
    {&CLEAR_ALLBND, &TimeIntegratorTaskList::ClearAllBoundary, false},
    {&CALC_HYDFLX, &TimeIntegratorTaskList::CalculateHydroFlux, true},
    {&CALC_FLDFLX, &TimeIntegratorTaskList::CalculateEMF, true},
    {&SEND_FLDFLX, &TimeIntegratorTaskList::SendEMF, true},
    {&RECV_FLDFLX, &TimeIntegratorTaskList::ReceiveAndCorrectEMF, false},
    {&INT_HYD, &TimeIntegratorTaskList::IntegrateHydro, true},
    {&INT_FLD, &TimeIntegratorTaskList::IntegrateField, true},
    {&SRC_TERM, &TimeIntegratorTaskList::AddSourceTerms, true},
    {&SEND_HYD, &TimeIntegratorTaskList::SendHydro, true},
    {&SEND_FLD, &TimeIntegratorTaskList::SendField, true},
    {&RECV_HYD, &TimeIntegratorTaskList::ReceiveHydro, false},
    {&RECV_FLD, &TimeIntegratorTaskList::ReceiveField, false},
    {&SETB_HYD, &TimeIntegratorTaskList::SetBoundariesHydro, true},
    {&SETB_FLD, &TimeIntegratorTaskList::SetBoundariesField, true},
    {&CONS2PRIM, &TimeIntegratorTaskList::Primitives, true},
    {&PHY_BVAL, &TimeIntegratorTaskList::PhysicalBoundary, true},
    {&USERWORK, &TimeIntegratorTaskList::UserWork, true},
    {&NEW_DT, &TimeIntegratorTaskList::NewBlockTimeStep, true},
    {&CALC_SCLRFLX, &TimeIntegratorTaskList::CalculateScalarFlux, true},
    {&INT_SCLR, &TimeIntegratorTaskList::IntegrateScalars, true},
    {&SEND_SCLR, &TimeIntegratorTaskList::SendScalars, true},
    {&RECV_SCLR, &TimeIntegratorTaskList::ReceiveScalars, false},
    {&SETB_SCLR, &TimeIntegratorTaskList::SetBoundariesScalars, true},
  };
  task_list_[ntasks].task_id = id;
  task_list_[ntasks].dependency = dep;
  bool found = false;
  for (const Row &r : table) {
    if (id == *r.id) {
      task_list_[ntasks].TaskFunc = static_cast<Fn>(r.fn);
      task_list_[ntasks].lb_time = r.lb;
      found = true;
      break;
    }
  }
  if (!found) {
    std::stringstream msg;
    msg << "### FATAL ERROR in TimeIntegratorTaskList::AddTask (B200 build)" << std::endl
        << "Invalid Task is specified" << std::endl;
    ATHENA_ERROR(msg);
  }
  ntasks++;
}

//----------------------------------------------------------------------------------------
// StartupTaskList (time_integrator.cpp:1384-1436): the storage registers start from zero

void TimeIntegratorTaskList::StartupTaskList(MeshBlock *pmb, int stage) {
  if (stage == 1) {
    Ck(ab_zero(Dev(), pmb->lid, AB_U1), "StartupTaskList (u1)");
    if (MAGNETIC_FIELDS_ENABLED) Ck(ab_zero(Dev(), pmb->lid, AB_B1_X1F), "StartupTaskList (b1)");
    if (NSCALARS > 0) Ck(ab_zero(Dev(), pmb->lid, AB_S1), "StartupTaskList (s1)");
  }
}

TaskStatus TimeIntegratorTaskList::ClearAllBoundary(MeshBlock *pmb, int stage) {
  Ck(ab_clear_boundary(Dev(), pmb->lid), "ClearAllBoundary");
  return TaskStatus::success;
}

//----------------------------------------------------------------------------------------
// fluxes (time_integrator.cpp:1442-1486): first-order fluxes in the predictor stage of vl2

TaskStatus TimeIntegratorTaskList::CalculateHydroFlux(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  const int order = (stage == 1 && integrator == "vl2") ? 1 : pmb->precon->xorder;
  Ck(ab_calc_fluxes(Dev(), pmb->lid, order, pmb->pmy_mesh->dt), "CalculateHydroFlux");
  return TaskStatus::next;
}

TaskStatus TimeIntegratorTaskList::CalculateEMF(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  Ck(ab_corner_e(Dev(), pmb->lid), "CalculateEMF");
  return TaskStatus::next;
}

TaskStatus TimeIntegratorTaskList::CalculateScalarFlux(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  const int order = (stage == 1 && integrator == "vl2") ? 1 : pmb->precon->xorder;
  Ck(ab_calc_scalar_fluxes(Dev(), pmb->lid, order), "CalculateScalarFlux");
  return TaskStatus::next;
}

//----------------------------------------------------------------------------------------
// EMF correction (time_integrator.cpp:1508-1558)

TaskStatus TimeIntegratorTaskList::SendEMF(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  Ck(ab_emf_send(Dev(), pmb->lid), "SendEMF");
  return TaskStatus::success;
}

TaskStatus TimeIntegratorTaskList::ReceiveAndCorrectEMF(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  const int rc = ab_emf_recv_try(Dev(), pmb->lid);
  Ck(rc, "ReceiveAndCorrectEMF");
  return rc ? TaskStatus::next : TaskStatus::fail;      // fail = polled again by the scheduler
}

//----------------------------------------------------------------------------------------
// integrate (time_integrator.cpp:1563-1650, 2141-2185): u1 = u1 + delta*u;
// u = gamma_1*u + gamma_2*u1 (a register swap when that is the identity on u1); u -= beta*dt*div F

namespace {
void Integrate(const TimeIntegratorTaskList::IntegratorWeight &sw, int lid, int reg, int reg1,
               const char *where) {
  const double w1[5] = {1.0, sw.delta, 0.0, 0.0, 0.0};
  Ck(ab_weighted_ave(Dev(), lid, reg1, reg, w1), where);
  if (sw.gamma_1 == 0.0 && sw.gamma_2 == 1.0 && sw.gamma_3 == 0.0) {
    Ck(ab_swap(Dev(), lid, reg), where);
  } else {
    const double w2[5] = {sw.gamma_1, sw.gamma_2, sw.gamma_3, 0.0, 0.0};
    Ck(ab_weighted_ave(Dev(), lid, reg, reg1, w2), where);
  }
}
}  // namespace

TaskStatus TimeIntegratorTaskList::IntegrateHydro(MeshBlock *pmb, int stage) {
  if (pmb->pmy_mesh->fluid_setup != FluidFormulation::evolve) return TaskStatus::next;
  if (stage > nstages) return TaskStatus::fail;
  Integrate(stage_wghts[stage-1], pmb->lid, AB_U, AB_U1, "IntegrateHydro");
  Ck(ab_add_flux_div(Dev(), pmb->lid, stage_wghts[stage-1].beta*pmb->pmy_mesh->dt),
     "IntegrateHydro (AddFluxDivergence)");
  return TaskStatus::next;
}

TaskStatus TimeIntegratorTaskList::IntegrateField(MeshBlock *pmb, int stage) {
  if (pmb->pmy_mesh->fluid_setup != FluidFormulation::evolve) return TaskStatus::next;
  if (stage > nstages) return TaskStatus::fail;
  Integrate(stage_wghts[stage-1], pmb->lid, AB_B_X1F, AB_B1_X1F, "IntegrateField");
  Ck(ab_ct(Dev(), pmb->lid, stage_wghts[stage-1].beta*pmb->pmy_mesh->dt), "IntegrateField (CT)");
  return TaskStatus::next;
}

TaskStatus TimeIntegratorTaskList::IntegrateScalars(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  Integrate(stage_wghts[stage-1], pmb->lid, AB_S, AB_S1, "IntegrateScalars");
  Ck(ab_add_scalar_flux_div(Dev(), pmb->lid, stage_wghts[stage-1].beta*pmb->pmy_mesh->dt),
     "IntegrateScalars (AddFluxDivergence)");
  return TaskStatus::next;
}

// AddSourceTerms (time_integrator.cpp:1655-1678): start-of-stage time, beta*dt
TaskStatus TimeIntegratorTaskList::AddSourceTerms(MeshBlock *pmb, int stage) {
  if (!(pmb->phydro->hsrc.hydro_sourceterms_defined)
      || pmb->pmy_mesh->fluid_setup != FluidFormulation::evolve) return TaskStatus::next;
  if (stage > nstages) return TaskStatus::fail;
  const Real t_start_stage = pmb->pmy_mesh->time + stage_wghts[stage-1].sbeta*pmb->pmy_mesh->dt;
  const Real dt = stage_wghts[stage-1].beta*pmb->pmy_mesh->dt;
  Ck(ab_add_source_terms(Dev(), pmb->lid, t_start_stage, dt), "AddSourceTerms");
  return TaskStatus::next;
}

//----------------------------------------------------------------------------------------
// ghost zones (time_integrator.cpp:1735-1814, 2187-2232)

#define B200_BVALS_TASKS(NAME, VAR)                                                          \
  TaskStatus TimeIntegratorTaskList::Send##NAME(MeshBlock *pmb, int stage) {                 \
    if (stage > nstages) return TaskStatus::fail;                                            \
    Ck(ab_bvals_send(Dev(), pmb->lid, VAR), "Send" #NAME);                                   \
    return TaskStatus::success;                                                              \
  }                                                                                          \
  TaskStatus TimeIntegratorTaskList::Receive##NAME(MeshBlock *pmb, int stage) {              \
    if (stage > nstages) return TaskStatus::fail;                                            \
    const int rc = ab_bvals_recv_try(Dev(), pmb->lid, VAR);                                  \
    Ck(rc, "Receive" #NAME);                                                                 \
    return rc ? TaskStatus::success : TaskStatus::fail;                                      \
  }                                                                                          \
  TaskStatus TimeIntegratorTaskList::SetBoundaries##NAME(MeshBlock *pmb, int stage) {        \
    if (stage > nstages) return TaskStatus::fail;                                            \
    Ck(ab_bvals_set(Dev(), pmb->lid, VAR), "SetBoundaries" #NAME);                           \
    return TaskStatus::success;                                                              \
  }
B200_BVALS_TASKS(Hydro, AB_VAR_HYDRO)
B200_BVALS_TASKS(Field, AB_VAR_FIELD)
B200_BVALS_TASKS(Scalars, AB_VAR_SCALARS)
#undef B200_BVALS_TASKS

//----------------------------------------------------------------------------------------
// Primitives (time_integrator.cpp:1965-2040), PhysicalBoundary (:2043-2068), UserWork,
// NewBlockTimeStep (:2071-2085)

TaskStatus TimeIntegratorTaskList::Primitives(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  Ck(ab_primitives(Dev(), pmb->lid), "Primitives");
  return TaskStatus::success;
}

TaskStatus TimeIntegratorTaskList::PhysicalBoundary(MeshBlock *pmb, int stage) {
  if (stage > nstages) return TaskStatus::fail;
  const Real t_end_stage = pmb->pmy_mesh->time + stage_wghts[stage-1].ebeta*pmb->pmy_mesh->dt;
  const Real dt = stage_wghts[stage-1].beta*pmb->pmy_mesh->dt;
  Ck(ab_physical_bcs_at(Dev(), pmb->lid, t_end_stage, dt), "PhysicalBoundary");
  return TaskStatus::success;
}

TaskStatus TimeIntegratorTaskList::UserWork(MeshBlock *pmb, int stage) {
  if (stage != nstages) return TaskStatus::success;      // only do on last stage
  Bridge &B = Bridge::Get();
  // the block's state of this cycle is final: bring it to the host when the host will read it
  if (B.HostStateNeededAfterThisCycle()) B.Download(pmb);
  pmb->UserWorkInLoop();
  return TaskStatus::success;
}

TaskStatus TimeIntegratorTaskList::NewBlockTimeStep(MeshBlock *pmb, int stage) {
  if (stage != nstages) return TaskStatus::success;      // only do on last stage
  pmb->phydro->NewBlockTimeStep();                       // shim/b200_new_blockdt.cpp
  return TaskStatus::success;
}

//----------------------------------------------------------------------------------------
// Tasks of features outside the B200 path.  The reference's SuperTimeStepTaskList
// (task_list/sts_task_list.cpp, compiled unchanged) takes their addresses, so they must exist;
// none of them can be reached: the constructor above rejects those configurations.

#define B200_NOT_ON_PATH(NAME)                                                               \
  TaskStatus TimeIntegratorTaskList::NAME(MeshBlock *, int) {                                \
    std::stringstream msg;                                                                   \
    msg << "### FATAL ERROR in TimeIntegratorTaskList::" #NAME << std::endl                  \
        << "this task is not on the B200 path" << std::endl;                                 \
    ATHENA_ERROR(msg);                                                                       \
    return TaskStatus::fail;                                                                 \
  }
B200_NOT_ON_PATH(SendHydroFlux)
B200_NOT_ON_PATH(ReceiveAndCorrectHydroFlux)
B200_NOT_ON_PATH(SendHydroFluxShear)
B200_NOT_ON_PATH(ReceiveHydroFluxShear)
B200_NOT_ON_PATH(SendHydroShear)
B200_NOT_ON_PATH(ReceiveHydroShear)
B200_NOT_ON_PATH(SendFieldShear)
B200_NOT_ON_PATH(ReceiveFieldShear)
B200_NOT_ON_PATH(SendEMFShear)
B200_NOT_ON_PATH(ReceiveEMFShear)
B200_NOT_ON_PATH(SendScalarFlux)
B200_NOT_ON_PATH(ReceiveScalarFlux)
B200_NOT_ON_PATH(SendScalarsShear)
B200_NOT_ON_PATH(ReceiveScalarsShear)
B200_NOT_ON_PATH(SendScalarsFluxShear)
B200_NOT_ON_PATH(ReceiveScalarsFluxShear)
B200_NOT_ON_PATH(DiffuseHydro)
B200_NOT_ON_PATH(DiffuseField)
B200_NOT_ON_PATH(DiffuseScalars)
#undef B200_NOT_ON_PATH
