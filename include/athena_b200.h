/* athena_b200.h -- C ABI of libathena_b200.so: the B200 (sm_100a) implementation of the
 * per-MeshBlock finite-volume hydro/MHD update of Athena++ (fork pabolmasov/Athena-gamma).
 *
 * Plain C types only (pointers, ints, doubles); no exceptions cross this boundary: every call
 * returns 0 on success or a negative error code, and ab_last_error() gives the message (the
 * host shim turns that into ATHENA_ERROR, src/defs.hpp.in:119-123).
 *
 * The reference has no plugin layer; its "operator API" for this path is the set of member
 * functions the task list calls (src/task_list/time_integrator.cpp:1442-2083) plus
 * Mesh::Initialize / Mesh::NewTimeStep.  Each entry point below names the member function it
 * replaces.  Blocks are addressed as (mesh handle, local block index `lid`) = pmb->lid.
 * All device memory is owned by the library; host pointers are borrowed for the call only.
 * There is no CPU fallback: every compute entry point fails with AB_ERR_NO_DEVICE when no
 * CUDA device is usable.
 */
#ifndef ATHENA_B200_H_
#define ATHENA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { AB_OK = 0, AB_ERR_ARG = -1, AB_ERR_NO_DEVICE = -2, AB_ERR_CUDA = -3, AB_ERR_NCCL = -4,
       AB_ERR_STATE = -5 };
enum { AB_BC_PERIODIC = 0, AB_BC_OUTFLOW = 1, AB_BC_REFLECT = 2, AB_BC_USER = 3 }; /* mesh/ix1_bc ... */
enum { AB_SOLVER_HLLE = 0, AB_SOLVER_HLLC = 1, AB_SOLVER_HLLD = 2, AB_SOLVER_ROE = 3,
       AB_SOLVER_LHLLC = 4, AB_SOLVER_LHLLD = 5, AB_SOLVER_LLF = 6 };
       /* --flux=hlle|hllc|hlld|roe|lhllc|lhlld|llf */
enum { AB_INT_VL2 = 0, AB_INT_RK2 = 1, AB_INT_RK1 = 2, AB_INT_RK3 = 3 };
/* registers (Hydro::u,u1,w ; Field::b,b1,bcc,e,wght ; Hydro::flux ; face EMFs) */
enum { AB_U = 0, AB_U1 = 1, AB_W = 2, AB_BCC = 3,
       AB_B_X1F = 4, AB_B_X2F = 5, AB_B_X3F = 6, AB_B1_X1F = 7, AB_B1_X2F = 8, AB_B1_X3F = 9,
       AB_FLUX_X1 = 10, AB_FLUX_X2 = 11, AB_FLUX_X3 = 12,
       AB_E_X1E = 13, AB_E_X2E = 14, AB_E_X3E = 15,
       AB_WGHT_X1F = 16, AB_WGHT_X2F = 17, AB_WGHT_X3F = 18,
       AB_E3_X1F = 19, AB_E2_X1F = 20, AB_E1_X2F = 21, AB_E3_X2F = 22, AB_E2_X3F = 23,
       AB_E1_X3F = 24,
       /* PassiveScalars::s, s1, r, s_flux[3] (src/scalars/scalars.hpp:40-56) */
       AB_S = 25, AB_S1 = 26, AB_R = 27, AB_SFLUX_X1 = 28, AB_SFLUX_X2 = 29, AB_SFLUX_X3 = 30,
       AB_NREG = 31 };
enum { AB_EOS_ADIABATIC = 0, AB_EOS_ISOTHERMAL = 1 };   /* configure.py --eos */

/* What configure.py flags + the athinput <mesh>/<meshblock>/<time>/<hydro> blocks fix
 * (src/defs.hpp.in:18-102, src/mesh/mesh.cpp:63-120, src/eos/adiabatic_mhd.cpp:26-31). */
typedef struct {
  int nx1, nx2, nx3;            /* <mesh> nx?            */
  int bx1, bx2, bx3;            /* <meshblock> nx?       */
  double x1min, x1max, x2min, x2max, x3min, x3max;
  int bc[6];                    /* ix1,ox1,ix2,ox2,ix3,ox3 : AB_BC_* */
  int nghost;                   /* NGHOST                */
  int mhd;                      /* MAGNETIC_FIELDS_ENABLED */
  int solver;                   /* RIEMANN_SOLVER        */
  int xorder;                   /* time/xorder 1,2,3     */
  int integrator;               /* time/integrator       */
  double gamma, dfloor, pfloor; /* hydro/gamma,dfloor,pfloor */
  double cfl_number, tlim, start_time;
  int rank, nranks;             /* this process / number of processes (one GPU each) */
  int device;                   /* CUDA device ordinal for this process */
  int nscalars;                 /* NSCALARS (configure.py --nscalars)   */
  int eos;                      /* AB_EOS_* (configure.py --eos)        */
  double sfloor;                /* hydro/sfloor (0 -> eos ctor default sqrt(1024*FLT_MIN)) */
  double iso_sound_speed;       /* hydro/iso_sound_speed (isothermal EOS) */
  double grav_acc[3];           /* hydro/grav_acc1..3: constant acceleration source term */
  int char_proj;                /* time/xorder = "2c" / "3c": characteristic reconstruction */
  double xrat[3];               /* mesh/x1rat..x3rat: geometric cell-size ratio (0 or 1 = uniform),
                                   Mesh ctor src/mesh/mesh.cpp:69-75,278-289 */
} AbMeshParams;

typedef struct AbMesh AbMesh;

const char *ab_last_error(void);
int ab_device_count(void);                                /* 0 when no usable GPU */

/* ---- construction: Mesh ctor + MeshBlock ctors + SearchAndSetNeighbors
 *      (src/mesh/mesh.cpp:63-548, src/bvals/bvals_base.cpp:299-480) */
int ab_mesh_create(const AbMeshParams *p, AbMesh **out);
int ab_mesh_destroy(AbMesh *m);
int ab_mesh_nblocks_total(const AbMesh *m);
int ab_mesh_nblocks_local(const AbMesh *m);
/* info[0]=gid, [1..3]=lx1..3, [4..6]=nc1..3 (cells incl. ghosts), [7..12]=is,ie,js,je,ks,ke */
int ab_block_info(const AbMesh *m, int lid, long *info);
/* number of doubles in a register of block lid */
long ab_reg_size(const AbMesh *m, int lid, int reg);

/* Host-only twin of ab_mesh_create (no GPU needed, owns no device memory): the MeshBlock
 * list, load balance and cross-rank message plan of rank p->rank, for inspection / CPU tests.
 * ab_plan_messages rows: {dir 0 send|1 recv, peer rank, key = dst_gid*64+dst_bufid, doubles,
 * local block}; kind 0 = ghost zones of u (and b), 1 = EMF correction; returns the row count.
 * Only ab_mesh_nblocks_*, ab_block_info, ab_plan_* and ab_mesh_destroy accept such a handle. */
int ab_plan_create(const AbMeshParams *p, AbMesh **out);
int ab_plan_messages(const AbMesh *m, int kind, long *out, int max_rows);
int ab_plan_ranklist(const AbMesh *m, int *out, int max_n);
/* Host-computed geometry of local block lid along dir 0..2 (Coordinates / Reconstruction ctors:
 * src/coordinates/coordinates.cpp:92-160, src/reconstruct/reconstruction.cpp:196-213,434-461):
 * what 0 x?f, 1 x?v, 2 dx?f, 3 / 4 PLM face weights, 5 nonuniform-reconstruction table (13
 * doubles per cell index; exists only when mesh/x?rat != 1), 6 / 7 weights of
 * Field::CalculateCellCenteredField, 8 cell centres of the MeshRefinement's coarse buffers
 * (refined meshes only).  Returns the number of doubles (0: no such array). */
int ab_plan_geometry(const AbMesh *m, int lid, int dir, int what, double *out, int max_n);

/* ---- static mesh refinement, host-side planner (no GPU needed; csrc/ab_smr.cpp).  The Mesh
 * ctor's <refinementN> handling, MeshBlockTree, CalculateLoadBalance and
 * BoundaryBase::SearchAndSetNeighbors on a refined mesh (src/mesh/mesh.cpp:323-465,
 * src/mesh/meshblock_tree.cpp:60-460, src/bvals/bvals_base.cpp:299-736) plus the transfer plan of
 * the cell-centred ghost exchange between levels, ProlongateBoundaries and the flux correction
 * (src/bvals/cc/bvals_cc.cpp:195-470, src/bvals/bvals_refine.cpp:96-570,
 * src/bvals/cc/flux_correction_cc.cpp:69-290) as index lists.  The device kernels that execute
 * the plan are not in this version: ab_mesh_create has no refinement input yet.
 *   blocks:    5 longs per MeshBlock in gid (Z) order {level, lx1, lx2, lx3, rank}
 *   neighbors: 8 ints per neighbour {ox1, ox2, ox3, type 0 face|1 edge|2 corner, gid, level,
 *              fi1, fi2}; nblevel = 27 ints [k][j][i] (-1: none)
 *   transfers: 12 longs per row {kind, src gid, src origin i j k, dst gid, dst origin i j k,
 *              extent ni nj nk}; kind 0 same level (fine -> fine ghost cells), 1 to a finer block
 *              (fine -> its coarse buffer), 2 to a coarser block (the sender's restricted slab in
 *              coarse indices -> fine ghost cells), 10 / 11 / 12 the ProlongateBoundaries work of
 *              one (block, coarser neighbour), 20 flux correction {fine gid, its face, -, -,
 *              coarse gid, its face, fi1, fi2} */
typedef struct {
  double x1min, x1max, x2min, x2max, x3min, x3max;   /* <refinementN> x?min / x?max */
  int level;                                          /* <refinementN> level (root grid = 0) */
} AbRefinementRegion;
typedef struct AbSmrPlan AbSmrPlan;
const char *ab_smr_last_error(void);
int ab_smr_plan_create(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                       AbSmrPlan **out);
int ab_smr_plan_destroy(AbSmrPlan *plan);
int ab_smr_plan_nblocks(const AbSmrPlan *plan);
int ab_smr_plan_blocks(const AbSmrPlan *plan, long *rows, int max_rows);
int ab_smr_plan_neighbors(const AbSmrPlan *plan, int gid, int *rows, int *nblevel);
long ab_smr_plan_transfers(const AbSmrPlan *plan, long *rows, long max_rows);
/* Mesh ctor with mesh/refinement = static: like ab_mesh_create, MeshBlocks from the planner, each
 * with the MeshRefinement's coarse buffers; the cycle then also runs the level-aware ghost
 * exchange, ProlongateBoundaries and the flux correction.  MeshBlocks of all levels are sharded
 * over the ranks like Mesh::CalculateLoadBalance (contiguous Z-order gid ranges, unit costs); the
 * transfers between blocks of different ranks travel over NCCL (ab_comm_init).  This version:
 * hydro (+ passive scalars), MeshBlocks of at least 2*NGHOST cells, no user-enrolled boundaries. */
int ab_mesh_create_refined(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                           AbMesh **out);
int ab_block_level(const AbMesh *m, int lid);    /* LogicalLocation::level (0 on a one-level mesh) */
/* host-only twin (no GPU needed), for ab_block_info / ab_block_level / ab_plan_geometry (what 8 =
 * cell centres of the coarse buffers) / ab_mesh_destroy */
int ab_plan_create_refined(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                           AbMesh **out);

/* ---- user-enrolled boundary functions: Mesh::EnrollUserBoundaryFunction (src/mesh/mesh.cpp)
 * with the BValFunc signature of src/athena.hpp:179-182 on plain arrays.  `face`: 0..5 =
 * inner_x1, outer_x1, ..., outer_x3 of a face whose flag is AB_BC_USER ("user").  The function
 * runs on the HOST: for every MeshBlock touching that face the library stages w (and the three
 * face-field arrays when MHD; NULL otherwise) in host memory in AthenaArray layout, calls fn,
 * and copies them back (host round trip per stage; time = end-of-stage time, dt = beta*dt as in
 * time_integrator.cpp:2045-2062).  Must be enrolled before ab_mesh_initialize, which fails
 * like bvals.cpp:328-335 otherwise. */
typedef void (*AbBValFunc)(void *user, int lid, double *prim, double *b_x1f, double *b_x2f,
                           double *b_x3f, double time, double dt, int il, int iu, int jl,
                           int ju, int kl, int ku, int ngh);
int ab_enroll_user_boundary_function(AbMesh *m, int face, AbBValFunc fn, void *user);

/* ---- user-enrolled explicit source terms: Mesh::EnrollUserExplicitSourceFunction with the
 * SrcTermFunc signature of src/athena.hpp:185-189 (the fork's production pgen sg_tde.cpp enrols
 * its black-hole gravity this way).  Called last in AddSourceTerms
 * (hydro/srcterms/hydro_srcterms.cpp:150-153) once per stage and MeshBlock after INT_HYD /
 * INT_SCLR, with time = start-of-stage time and dt = beta*dt (time_integrator.cpp:1655-1678).
 *   host variant:   prim, prim_scalar, bcc, cons, cons_scalar are HOST arrays in AthenaArray
 *                   layout (NULL where the build has none); the library stages them around the
 *                   call (download w,r,bcc,u,s; upload u,s).
 *   device variant: the same arguments are DEVICE pointers and `stream` is the library's
 *                   cudaStream_t; the function must only enqueue work on that stream (no copies,
 *                   no synchronisation) -- the native way to keep a production source term on
 *                   the GPU. */
typedef void (*AbSrcTermFunc)(void *user, int lid, double time, double dt, const double *prim,
                              const double *prim_scalar, const double *bcc, double *cons,
                              double *cons_scalar);
typedef void (*AbSrcTermFuncDevice)(void *user, int lid, double time, double dt,
                                    const double *prim, const double *prim_scalar,
                                    const double *bcc, double *cons, double *cons_scalar,
                                    void *stream);
int ab_enroll_user_explicit_source_function(AbMesh *m, AbSrcTermFunc fn, void *user);
int ab_enroll_user_explicit_source_function_device(AbMesh *m, AbSrcTermFuncDevice fn, void *user);

/* ---- host <-> device mirror of the AthenaArrays (same layout as AthenaArray::data()) */
int ab_upload(AbMesh *m, int lid, int reg, const double *host);     /* after ProblemGenerator */
int ab_download(AbMesh *m, int lid, int reg, double *host);         /* before outputs / hooks */
int ab_download_coord(AbMesh *m, int lid, int which, double *host); /* 0..8: x1f x2f x3f x1v x2v x3v dx1f dx2f dx3f */

/* ---- pipelined staging for callers that stream a state in and a result out EVERY step (the
 * bench's end-to-end leg; a coupled code exchanging fields with a host-side solver).  The chosen
 * registers of all local blocks get device staging buffers; uploads and downloads run on their
 * own copy streams ordered against the compute stream by events, so the PCIe transfers of the
 * neighbouring steps overlap this step's kernels.  host[] = one PINNED host pointer per
 * (local block, register), index lid*nregs + r.  Sequence per step:
 *   ab_stage_commit            staging -> registers (waits for the upload; frees the buffer)
 *   ab_stage_upload_all        next step's input, returns at once
 *   ab_mesh_initialize / ab_mesh_cycles ...
 *   ab_stage_download_all      registers -> staging -> host, returns at once
 * and ab_stage_sync before the host touches the downloaded arrays. */
int ab_stage_begin(AbMesh *m, const int *regs, int nregs);
int ab_stage_upload_all(AbMesh *m, const double *const *host);
int ab_stage_commit(AbMesh *m);
int ab_stage_download_all(AbMesh *m, double *const *host);
int ab_stage_sync(AbMesh *m);

/* ---- multi-process plumbing (replaces MPI_Init + persistent requests, bvals_cc.cpp:528-633).
 * rank 0 calls ab_comm_unique_id and the host broadcasts the 128 bytes (MPI_Bcast /
 * torch.distributed); then every rank calls ab_comm_init before the first exchange. */
int ab_comm_unique_id(unsigned char id[128]);
int ab_comm_init(AbMesh *m, const unsigned char id[128]);

/* ---- task bodies, one block (src/task_list/time_integrator.cpp) -------------------------- */
/* EquationOfState::ConservedToPrimitive (eos/adiabatic_{hydro,mhd}.cpp:39-90) */
int ab_cons2prim(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku);
/* EquationOfState::PrimitiveToConserved (eos/adiabatic_{hydro,mhd}.cpp:89-136) */
int ab_prim2cons(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku);
/* TimeIntegratorTaskList::Primitives range logic (:1965-1983) + ConservedToPrimitive */
int ab_primitives(AbMesh *m, int lid);
/* Hydro::CalculateFluxes(w,b,bcc,order) (hydro/calculate_fluxes.cpp:36-378); dt = pmesh->dt */
int ab_calc_fluxes(AbMesh *m, int lid, int order, double dt);
/* Field::ComputeCornerE (field/calculate_corner_e.cpp:28-236) */
int ab_corner_e(AbMesh *m, int lid);
/* MeshBlock::WeightedAve (mesh/weighted_ave.cpp): out = f(w[0]*out, w[1]*in); regs AB_U/AB_U1
 * (cell-centred), AB_S/AB_S1 (passive scalars) or AB_B_X1F/AB_B1_X1F (all three face arrays
 * of b / b1) */
int ab_weighted_ave(AbMesh *m, int lid, int out_reg, int in_reg, const double w[5]);
/* AthenaArray::SwapAthenaArray on (u,u1) when reg==AB_U, (b,b1) when reg==AB_B_X1F, (s,s1)
 * when reg==AB_S */
int ab_swap(AbMesh *m, int lid, int reg);
/* AthenaArray::ZeroClear on u1 (AB_U1), b1 (AB_B1_X1F) or s1 (AB_S1)
 * (time_integrator.cpp:1386-1410) */
int ab_zero(AbMesh *m, int lid, int reg);
/* Hydro::AddFluxDivergence(wght, u) (hydro/add_flux_divergence.cpp:39-96) */
int ab_add_flux_div(AbMesh *m, int lid, double wght);
/* HydroSourceTerms::AddSourceTerms(time, dt, ...) (hydro/srcterms/hydro_srcterms.cpp:117-156):
 * constant acceleration hydro/grav_acc1..3 (constant_acc.cpp:25-77) on u with the current w,
 * then the user-enrolled explicit source function; time = start-of-stage time, dt = beta*dt
 * (time_integrator.cpp:1655-1678).  ab_mesh_cycles fuses the constant acceleration into the
 * IntegrateHydro kernel. */
int ab_add_source_terms(AbMesh *m, int lid, double time, double dt);
/* Field::CT(wght, b) (field/ct.cpp:31-116) */
int ab_ct(AbMesh *m, int lid, double wght);
/* ---- passive scalars (NSCALARS > 0), src/scalars + src/eos/eos_scalars.cpp ---------------- */
/* PassiveScalars::CalculateFluxes(r, order) (scalars/calculate_scalar_fluxes.cpp:41-382);
 * needs the mass flux of ab_calc_fluxes */
int ab_calc_scalar_fluxes(AbMesh *m, int lid, int order);
/* PassiveScalars::AddFluxDivergence(wght, s) (scalars/add_scalar_flux_divergence.cpp:43-97) */
int ab_add_scalar_flux_div(AbMesh *m, int lid, double wght);
/* EquationOfState::PassiveScalarConservedToPrimitive / PrimitiveToConserved
 * (eos/eos_scalars.cpp:31-60,133-152) on a cell range (uses the current u(IDN)) */
int ab_scalar_cons2prim(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku);
int ab_scalar_prim2cons(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku);
/* BoundaryValues::ApplyPhysicalBoundaries (bvals/bvals.cpp:436-620): outflow and reflecting
 * faces (cc/outflow_cc.cpp, cc/hydro/reflect_hydro.cpp, fc/outflow_fc.cpp, fc/reflect_fc.cpp) */
int ab_physical_bcs(AbMesh *m, int lid);
/* the same with the (time, dt) arguments of ApplyPhysicalBoundaries, which user-enrolled
 * boundary functions receive: end-of-stage time and beta*dt from the PhysicalBoundary task
 * (task_list/time_integrator.cpp:2045-2062) */
int ab_physical_bcs_at(AbMesh *m, int lid, double time, double dt);
/* Hydro::NewBlockTimeStep (hydro/new_blockdt.cpp:42-190): *dt_out = new_block_dt_ (synchronises) */
int ab_new_block_dt(AbMesh *m, int lid, double *dt_out);

/* ---- boundary tasks of ONE MeshBlock, for the reference's polling scheduler
 * (TaskList::DoAllAvailableTasks, task_list/task_list.cpp:29-59: a task that answers
 * TaskStatus::fail is retried later).  var: AB_VAR_HYDRO (hbvar), AB_VAR_FIELD (fbvar),
 * AB_VAR_SCALARS (sbvar).  Safe to call for different blocks from different host threads
 * (task_list.cpp:71-88 runs the block loop under OpenMP): every per-block entry point of this
 * header locks the mesh while it enqueues. */
enum { AB_VAR_HYDRO = 0, AB_VAR_FIELD = 1, AB_VAR_SCALARS = 2 };
/* BoundaryVariable::SendBoundaryBuffers (bvals/bvals_var.cpp:212-236): the block's active zones
 * are final for this stage; packs the buffers of neighbours on other ranks.  Never waits. */
int ab_bvals_send(AbMesh *m, int lid, int var);
/* BoundaryVariable::ReceiveBoundaryBuffers (bvals_var.cpp:242-270): 1 = every neighbour has
 * sent (TaskStatus::success), 0 = not yet (TaskStatus::fail, poll again), < 0 = error */
int ab_bvals_recv_try(AbMesh *m, int lid, int var);
/* BoundaryVariable::SetBoundaries (bvals_var.cpp:276-296): fills the block's ghost zones */
int ab_bvals_set(AbMesh *m, int lid, int var);
/* FaceCenteredBoundaryVariable::SendFluxCorrection (bvals/fc/flux_correction_fc.cpp:623-680) */
int ab_emf_send(AbMesh *m, int lid);
/* FaceCenteredBoundaryVariable::ReceiveFluxCorrection (flux_correction_fc.cpp:1610-1749):
 * 1 = all surface EMFs arrived, summed in neighbour-list order and averaged; 0 = not yet */
int ab_emf_recv_try(AbMesh *m, int lid);
/* BoundaryValues::ClearBoundarySubset (bvals/bvals.cpp:390-418): end of the stage's
 * communication for this block (reports a variable that was sent but never received) */
int ab_clear_boundary(AbMesh *m, int lid);

/* ---- boundary communication over all local blocks (src/bvals): Send + Receive + Set.
 * Same-device neighbours are copied device-to-device; neighbours on other ranks go through
 * NCCL send/recv of packed buffers in the reference's buffer layout. */
/* FaceCenteredBoundaryVariable::SendFluxCorrection + ReceiveFluxCorrection
 * (bvals/fc/flux_correction_fc.cpp:623-680,1610-1749) */
int ab_emf_exchange(AbMesh *m);
/* hbvar / fbvar SendBoundaryBuffers + ReceiveBoundaryBuffers + SetBoundaries
 * (bvals/bvals_var.cpp:212-296) for u (and b when MHD, s when NSCALARS > 0) */
int ab_bvals_exchange(AbMesh *m);

/* ---- whole-mesh driver (host side of the path): Mesh::Initialize after ProblemGenerator
 * (mesh/mesh.cpp:1416-1649) and the main loop body (main.cpp:430-515 without outputs). */
int ab_mesh_initialize(AbMesh *m);
/* run ncycles cycles (all stages of TimeIntegratorTaskList + Mesh::NewTimeStep) without
 * host synchronisation; stops early when time >= tlim.  dt stays on the device. */
int ab_mesh_cycles(AbMesh *m, int ncycles);
/* async=1: ab_mesh_cycles never synchronises (caller guarantees tlim is not reached) */
int ab_mesh_set_async(AbMesh *m, int async);
/* read back {time, dt, ncycle} (synchronises the stream) */
int ab_mesh_state(AbMesh *m, double *time, double *dt, long *ncycle);
int ab_mesh_set_time_dt(AbMesh *m, double time, double dt);
/* HistoryOutput::WriteOutputFile sums (outputs/history.cpp:69-169) reduced on the device over
 * the active cells of all MeshBlocks of all ranks (NCCL sum when nranks > 1): mass, 1-mom,
 * 2-mom, 3-mom, 1-KE, 2-KE, 3-KE, tot-E, [1-ME, 2-ME, 3-ME when MHD], [one per scalar].
 * Returns the number of values written (<= max_n) or a negative error.  The summation order
 * is fixed (reproducible run to run) but is a tree, not the reference's running sum: values
 * agree with the reference to ~1e-15 relative to the sum of magnitudes. */
int ab_history(AbMesh *m, double *out, int max_n);
/* per-cycle dt history of the last ab_mesh_cycles call (dt used by each cycle) */
int ab_mesh_dt_history(AbMesh *m, double *out, int max_n);
/* CUDA-event timing of the reconstruct+Riemann kernels on the compute stream (roofline
 * evidence): enable, run cycles, then read out[0..8]=ms and out[9..17]=launches per
 * slot = dir*3 + (order-1); reading resets the accumulators. */
int ab_mesh_profile(AbMesh *m, int enable);
int ab_mesh_profile_read(AbMesh *m, double *out);
/* count of kernel launches issued by this mesh since creation (for bench accounting) */
long ab_mesh_launch_count(const AbMesh *m);
void *ab_mesh_stream(AbMesh *m);                          /* cudaStream_t of the compute stream */
int ab_mesh_sync(AbMesh *m);

#ifdef __cplusplus
}
#endif
#endif /* ATHENA_B200_H_ */
