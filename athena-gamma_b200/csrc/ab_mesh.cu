// ab_mesh.cu -- host runtime behind the C ABI (include/athena_b200.h): MeshBlock bookkeeping,
// device memory, neighbour tables, ghost-exchange / EMF-correction plans, NCCL transport and
// the TimeIntegratorTaskList stage loop.  All physics runs in the kernels of ab_kernels.cu;
// nothing here computes on the host (there is no CPU fallback).
//
// Reference structure mirrored: Mesh ctor / MeshBlockTree Z-ordering (src/mesh/mesh.cpp,
// meshblock_tree.cpp), Coordinates (src/coordinates), BoundaryBase::SearchAndSetNeighbors
// (src/bvals/bvals_base.cpp), TimeIntegratorTaskList (src/task_list/time_integrator.cpp).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <float.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <array>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/athena_b200.h"
#include "ab_kernels.h"
#include "ab_smr_exec.h"

namespace ab {
extern std::atomic<long> g_launches;
}

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess)                                                             \
      return fail(AB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
  } while (0)

// ------------------------------------------------------------------ NCCL through dlopen
// (the library has no link-time NCCL dependency: single-GPU use needs none, and inside a
//  torch process the already-loaded libnccl.so.2 is picked up)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    // AB_NCCL_LIB: which NCCL to load (an absolute path; default: the libnccl.so.2 already in the
    // process, else the system's)
    if (const char *e = getenv("AB_NCCL_LIB")) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define LD(name) name = (decltype(name))dlsym(h, "nccl" #name)
    LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(Send); LD(Recv); LD(AllReduce);
    LD(GroupStart); LD(GroupEnd); LD(GetErrorString);
#undef LD
    return GetUniqueId && CommInitRank && Send && Recv && AllReduce && GroupStart && GroupEnd;
  }
};
Nccl g_nccl;
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble
constexpr int NCCL_MIN = 3;       // ncclMin
constexpr int NCCL_SUM = 0;       // ncclSum

#define NK(call)                                                                       \
  do {                                                                                 \
    int r_ = (call);                                                                   \
    if (r_ != 0)                                                                       \
      return fail(AB_ERR_NCCL, std::string(#call) + ": " +                            \
                  (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); \
  } while (0)

// ------------------------------------------------------------------ mesh geometry helpers
// src/mesh/mesh.hpp:389-405
double mesh_gen_x(long index, long nrange) {
  long noffset = index - (nrange)/2;
  long noffset_ceil = index - (nrange+1)/2;
  return static_cast<double>(noffset + noffset_ceil)/(2.0*nrange);
}
// src/mesh/mesh.hpp:467-488
double uniform_gen(double x, double xmin, double xmax) {
  return 0.5*(xmin + xmax) + (x*xmax - x*xmin);
}

// DefaultMeshGeneratorX1..3 (src/mesh/mesh.hpp:411-455): geometric spacing with cell-size ratio
// `rat`, logical x in [0,1]; used in a direction whose mesh/x?rat != 1
// (Mesh::use_uniform_meshgen_fn_, mesh/mesh.cpp:278-289)
double ratio_gen(double x, double xmin, double xmax, double rat, int nx) {
  double lw, rw;
  if (rat == 1.0) {
    rw = x; lw = 1.0 - x;
  } else {
    double ratn = std::pow(rat, nx);
    double rnx = std::pow(rat, x*nx);
    lw = (rnx - ratn)/(1.0 - ratn);
    rw = 1.0 - lw;
  }
  return xmin*lw + xmax*rw;
}
// logical block / face position -> coordinate (mesh.cpp:1679-1688, coordinates.cpp:100-145)
double gen_x(long index, long nrange, double xmin, double xmax, double rat, int nx) {
  if (rat != 1.0)
    return ratio_gen(static_cast<double>(index)/static_cast<double>(nrange), xmin, xmax, rat, nx);
  return uniform_gen(mesh_gen_x(index, nrange), xmin, xmax);
}

struct Nb {
  int ox1, ox2, ox3, type;   // 0 face, 1 edge, 2 corner
  int gid, rank, bufid, targetid, fid, eid;
};

struct HostBlock {
  int gid = 0, rank = 0;
  long lx[3] = {0, 0, 0};
  int bcs[6];                 // -1: neighbour block; otherwise the physical AB_BC_*
  int nblevel[3][3][3];
  std::vector<Nb> nbs;
  int nedge_fine[12];
  double bmin[3], bmax[3];
  int level = 0, dl = 0;      // refined meshes: logical level and its height above the root grid
};

struct PeerBuf {              // staging for one peer rank and one exchange kind
  double *send = nullptr, *recv = nullptr;
  size_t nsend = 0, nrecv = 0;
  // Direct exchange over NVLink peer memory (AbMesh::p2p): my receive buffers, two of them used
  // alternately, are exported with cudaIpcGetMemHandle; p2p_send[] are the PEER's receive
  // buffers for my messages mapped into this process -- the pack kernel stores straight into
  // them and no ncclSend / ncclRecv moves the ghost zones.
  double *p2p_recv[2] = {nullptr, nullptr};
  double *p2p_send[2] = {nullptr, nullptr};
};

struct LocalBlock {
  HostBlock *hb = nullptr;
  ab::BlkDev d;
  ab::ReconGeom g;
  ab::EmfPlan emf;
  char *base = nullptr;                   // this block's slab inside AbMesh::slab
  size_t nbytes = 0;
  long regsize[AB_NREG];
  double *coord_dev[9];
  long coord_n[9];
  double *emf_send = nullptr;             // send buffers toward same-rank neighbours
  long face_off[6], edge_off[12];
  int parity = 0;                         // (u,u1) swap state
  int parity_b = 0;                       // (b,b1) swap state
  int parity_s = 0;                       // (s,s1) swap state
  unsigned long long *dtmin = nullptr;    // slot in the mesh-wide array
  std::vector<void *> debug_allocs;       // AB_DEBUG_ALLOC=1: per-array allocations
  cudaStream_t stream = nullptr;          // stream of this block's per-block tasks inside a cycle
  bool cc_e_valid = false;                // cc_e written by the fused cons2prim
  bool has_phys_bc = false;
};

}  // namespace

struct AbMesh {
  // The reference runs its task bodies from an OpenMP loop over MeshBlocks
  // (task_list.cpp:71-88): every entry point that takes a block index locks the mesh, so calls
  // for different blocks may come from different host threads (they only enqueue work).
  std::recursive_mutex mu;
  AbMeshParams p;
  ab::Params kp;
  int ndim = 1, f2 = 0, f3 = 0;
  int nh = 5;                             // NHYDRO: 5 adiabatic, 4 isothermal
  int nrb[3] = {1, 1, 1};
  double xrat[3] = {1.0, 1.0, 1.0};       // mesh/x?rat (0 read as 1; 1 in a degenerate direction)
  int nbtotal = 0;
  int nc[3] = {1, 1, 1}, is = 0, ie = 0, js = 0, je = 0, ks = 0, ke = 0;
  std::vector<HostBlock> hb;              // all blocks of the mesh (every rank knows them)
  std::vector<int> gid_of;                // [lx3][lx2][lx1] -> gid
  std::vector<LocalBlock> lb;             // blocks owned by this rank, in gid order
  std::vector<HostBlock *> lb_hb;         // their host descriptors (also for dry plans)
  bool dry = false;                       // host-only plan (no device resources)
  int gid_start = 0;
  int nstages = 2;
  double beta[4], delta[4], g1[4], g2[4], g3[4], ebeta[4];
  double cfl = 0.0;
  // user-enrolled boundary functions (Mesh::BoundaryFunction_) and their host staging
  AbBValFunc user_bc[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void *user_bc_arg[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool has_user_bc = false;
  // user-enrolled explicit source function (HydroSourceTerms::UserSourceTerm)
  AbSrcTermFunc user_src = nullptr;
  AbSrcTermFuncDevice user_src_dev = nullptr;
  void *user_src_arg = nullptr;
  double sbeta[4];
  std::vector<double> stage_u, stage_r, stage_s, stage_bcc;
  double bc_time = 0.0, bc_dt = 0.0;      // (time, dt) of the PhysicalBoundary task
  std::vector<double> stage_w, stage_b[3];
  cudaStream_t stream = nullptr;
  // all local blocks' slabs in one allocation, bstride bytes apart (0: separate allocations,
  // AB_DEBUG_ALLOC); the EMF-correction plans of all local blocks for the batched launches
  char *slab = nullptr;
  long bstride = 0;
  ab::EmfPlan *emf_plans_dev = nullptr;
  bool batch = false;                     // inside a cycle that runs every task as one launch
  // overlapped schedule (AB_OVERLAP=1): NCCL transfers on comm_stream while the compute stream
  // works on data that does not depend on them.  Opt-in: at 2 GPUs x 512^3 it measured 1.2 %
  // SLOWER than the in-order schedule (73.10 vs 72.21 ms per cycle, profiles/r1_tuning_log.md):
  // over NVLink the transfers are shorter than the six extra ghost-shell launches cost.
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_recv = nullptr;
  bool overlap = false;
  // Static mesh refinement (ab_mesh_create_refined): one process, hydro + passive scalars.  The
  // host planner (ab_smr.cpp) supplies the block list and the transfer rows; every block gets the
  // MeshRefinement's coarse buffers and coarse cell centres.
  bool smr = false;
  AbSmrPlan *smr_plan = nullptr;
  std::vector<ab::SmrRow> smr_rows;
  struct SmrBlk {
    double *cu = nullptr, *cw = nullptr, *cs = nullptr, *cr = nullptr, *cxv[3] = {nullptr, nullptr, nullptr};
    ab::SmrGeom g;
  };
  std::vector<SmrBlk> smr_blk;
  ab::CopyBox *smr_boxes = nullptr; size_t smr_boxes_cap = 0;
  std::map<int, PeerBuf> peer_smr, peer_smr_flux;   // refined meshes across ranks: message buffers
  // Pipelined host <-> device staging (ab_stage_*): a caller that streams a new state in and
  // the result out every step (bench.py's e2e leg) overlaps the PCIe copies of the neighbouring
  // steps with this step's kernels.  Device staging buffers `in` / `out` hold the chosen
  // registers of every local block; copies run on their own streams, ordered by events.
  struct Staging {
    std::vector<int> regs;
    std::vector<size_t> off;              // [lid*nregs + r] -> element offset into in / out
    size_t elems = 0;
    double *in = nullptr, *out = nullptr;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t in_ready = nullptr, in_free = nullptr, out_ready = nullptr, out_free = nullptr;
    bool active = false;
  } stg;
  // Many MeshBlocks per GPU: the per-block tasks of a stage are spread round-robin over a few
  // extra streams (forked from / joined into the main stream around every exchange), so that the
  // ramp-up and tail of one block's kernels overlap the next block's (AB_BLOCK_STREAMS, default
  // min(#blocks, 4); 0 = everything on the main stream)
  std::vector<cudaStream_t> bstream;
  std::vector<cudaEvent_t> ev_b;
  cudaEvent_t ev_main = nullptr;
  double *state = nullptr;                // device: time, dt, tlim, cfl, min, ncycle
  double *dt_hist = nullptr;              // device ring of per-cycle dt
  int hist_cap = 0, hist_n = 0;
  unsigned long long *dtmin = nullptr;    // device per-local-block min (bits)
  double *hist_partial = nullptr, *hist_out = nullptr;   // HistoryOutput reduction scratch
  double h_time = 0.0, h_dt = DBL_MAX;
  long h_ncycle = 0;
  int async = 0;
  // exchange plans (two variants: register parity 0 / 1)
  struct Plan {
    bool built = false;
    ab::CopyBox *pack = nullptr; int npack = 0; long maxpack = 0;     // to peer send buffers
    ab::CopyBox *phase1 = nullptr; int n1 = 0; long max1 = 0;         // ghost fill, local sources
    ab::CopyBox *phase1r = nullptr; int n1r = 0; long max1r = 0;      // ghost fill from peer buffers
    ab::CopyBox *phase2 = nullptr; int n2 = 0; long max2 = 0;         // 1-D/2-D duplicates
  } plan[24];                             // [0,8): NCCL path by register-swap parity; [8,24): direct exchange, parity | buffer << 3
  std::map<int, PeerBuf> peer_state, peer_emf;
  // per-block boundary tasks (ab_bvals_send / recv_try / set, ab_emf_send / recv_try): how many
  // times each local block has sent / received each variable, the number of completed NCCL
  // rounds, and the per-(block, variable, register parity) copy plans
  struct BlockComm { long sent[3] = {0, 0, 0}, recvd[3] = {0, 0, 0}, emf_sent = 0, emf_recvd = 0; };
  std::vector<BlockComm> bcomm;
  long nccl_state_round = 0, nccl_emf_round = 0;
  // p2p: 0 off (NCCL moves the ghost zones), 1 requested (AB_P2P=1), 2 active (peer buffers
  // mapped), -1 unavailable (cudaIpc* failed: NCCL path)
  int p2p = 0;
  long p2p_round = 0;                     // whole-mesh ghost exchanges done (selects the buffer)
  double *p2p_token = nullptr;            // 1 double: operand of the barrier all-reduce
  std::map<long, Plan> bplan;
  ncclComm_t comm = nullptr;
  bool emf_built = false;
  // optional CUDA-event timing of the flux kernels (bench.py roofline): slot = dir*3+(order-1)
  int profile = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
  std::vector<int> prof_slot;
  double prof_ms[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  long prof_n[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace {

using ab::CopyBox;

struct Msg { int peer; long key; int lid; int nbi; long count; };

// ---- buffer box ranges (same for every block: all blocks have identical shape) ------------
struct Box { int si, ei, sj, ej, sk, ek; long count() const {
  return (long)(ei-si+1)*(ej-sj+1)*(ek-sk+1); } };

// LoadBoundaryBufferSameLevel (bvals/cc/bvals_cc.cpp:201-216)
Box cc_send_box(const AbMesh *m, int ox1, int ox2, int ox3) {
  int ng = m->p.nghost;
  Box b;
  b.si = (ox1 > 0) ? (m->ie - ng + 1) : m->is; b.ei = (ox1 < 0) ? (m->is + ng - 1) : m->ie;
  b.sj = (ox2 > 0) ? (m->je - ng + 1) : m->js; b.ej = (ox2 < 0) ? (m->js + ng - 1) : m->je;
  b.sk = (ox3 > 0) ? (m->ke - ng + 1) : m->ks; b.ek = (ox3 < 0) ? (m->ks + ng - 1) : m->ke;
  return b;
}
// SetBoundarySameLevel (bvals/cc/bvals_cc.cpp:300-336)
Box cc_recv_box(const AbMesh *m, int ox1, int ox2, int ox3) {
  int ng = m->p.nghost;
  Box b;
  if (ox1 == 0) { b.si = m->is; b.ei = m->ie; }
  else if (ox1 > 0) { b.si = m->ie + 1; b.ei = m->ie + ng; }
  else { b.si = m->is - ng; b.ei = m->is - 1; }
  if (ox2 == 0) { b.sj = m->js; b.ej = m->je; }
  else if (ox2 > 0) { b.sj = m->je + 1; b.ej = m->je + ng; }
  else { b.sj = m->js - ng; b.ej = m->js - 1; }
  if (ox3 == 0) { b.sk = m->ks; b.ek = m->ke; }
  else if (ox3 > 0) { b.sk = m->ke + 1; b.ek = m->ke + ng; }
  else { b.sk = m->ks - ng; b.ek = m->ks - 1; }
  return b;
}
// FaceCentered LoadBoundaryBufferSameLevel (bvals/fc/bvals_fc.cpp:344-397), comp = 0,1,2
Box fc_send_box(const AbMesh *m, int comp, int ox1, int ox2, int ox3) {
  int ng = m->p.nghost;
  int is = m->is, ie = m->ie, js = m->js, je = m->je, ks = m->ks, ke = m->ke;
  Box b;
  // i
  if (comp == 0) {
    if (ox1 == 0) { b.si = is; b.ei = ie + 1; }
    else if (ox1 > 0) { b.si = ie - ng + 1; b.ei = ie; }
    else { b.si = is + 1; b.ei = is + ng; }
  } else {
    if (ox1 == 0) { b.si = is; b.ei = ie; }
    else if (ox1 > 0) { b.si = ie - ng + 1; b.ei = ie; }
    else { b.si = is; b.ei = is + ng - 1; }
  }
  // j
  if (comp == 1) {
    if (!m->f2) { b.sj = js; b.ej = je; }
    else if (ox2 == 0) { b.sj = js; b.ej = je + 1; }
    else if (ox2 > 0) { b.sj = je - ng + 1; b.ej = je; }
    else { b.sj = js + 1; b.ej = js + ng; }
  } else {
    if (ox2 == 0) { b.sj = js; b.ej = je; }
    else if (ox2 > 0) { b.sj = je - ng + 1; b.ej = je; }
    else { b.sj = js; b.ej = js + ng - 1; }
  }
  // k
  if (comp == 2) {
    if (!m->f3) { b.sk = ks; b.ek = ke; }
    else if (ox3 == 0) { b.sk = ks; b.ek = ke + 1; }
    else if (ox3 > 0) { b.sk = ke - ng + 1; b.ek = ke; }
    else { b.sk = ks + 1; b.ek = ks + ng; }
  } else {
    if (ox3 == 0) { b.sk = ks; b.ek = ke; }
    else if (ox3 > 0) { b.sk = ke - ng + 1; b.ek = ke; }
    else { b.sk = ks; b.ek = ks + ng - 1; }
  }
  return b;
}
// FaceCentered SetBoundarySameLevel (bvals/fc/bvals_fc.cpp:583-684)
Box fc_recv_box(const AbMesh *m, int comp, int ox1, int ox2, int ox3) {
  int ng = m->p.nghost;
  int is = m->is, ie = m->ie, js = m->js, je = m->je, ks = m->ks, ke = m->ke;
  Box b;
  if (comp == 0) {
    if (ox1 == 0) { b.si = is; b.ei = ie + 1; }
    else if (ox1 > 0) { b.si = ie + 2; b.ei = ie + ng + 1; }
    else { b.si = is - ng; b.ei = is - 1; }
  } else {
    if (ox1 == 0) { b.si = is; b.ei = ie; }
    else if (ox1 > 0) { b.si = ie + 1; b.ei = ie + ng; }
    else { b.si = is - ng; b.ei = is - 1; }
  }
  if (comp == 1) {
    if (!m->f2) { b.sj = js; b.ej = je; }
    else if (ox2 == 0) { b.sj = js; b.ej = je + 1; }
    else if (ox2 > 0) { b.sj = je + 2; b.ej = je + ng + 1; }
    else { b.sj = js - ng; b.ej = js - 1; }
  } else {
    if (ox2 == 0) { b.sj = js; b.ej = je; }
    else if (ox2 > 0) { b.sj = je + 1; b.ej = je + ng; }
    else { b.sj = js - ng; b.ej = js - 1; }
  }
  if (comp == 2) {
    if (!m->f3) { b.sk = ks; b.ek = ke; }
    else if (ox3 == 0) { b.sk = ks; b.ek = ke + 1; }
    else if (ox3 > 0) { b.sk = ke + 2; b.ek = ke + ng + 1; }
    else { b.sk = ks - ng; b.ek = ks - 1; }
  } else {
    if (ox3 == 0) { b.sk = ks; b.ek = ke; }
    else if (ox3 > 0) { b.sk = ke + 1; b.ek = ke + ng; }
    else { b.sk = ks - ng; b.ek = ks - 1; }
  }
  return b;
}

// strides (k stride, j stride) of the register arrays
void fc_strides(const AbMesh *m, int comp, long &s3, long &s2) {
  int n1 = m->nc[0] + (comp == 0), n2 = m->nc[1] + (comp == 1);
  s2 = n1; s3 = (long)n1*n2;
}

long state_msg_count(const AbMesh *m, int ox1, int ox2, int ox3) {
  long n = m->nh*cc_send_box(m, ox1, ox2, ox3).count();
  if (m->p.mhd) for (int c = 0; c < 3; ++c) n += fc_send_box(m, c, ox1, ox2, ox3).count();
  n += m->p.nscalars*cc_send_box(m, ox1, ox2, ox3).count();   // s rides behind u and b
  return n;
}

// EMF message sizes (bvals/fc/bvals_fc.cpp:304-338 / flux_correction_fc.cpp:51-307)
long emf_face_count(const AbMesh *m, int fid) {
  int nx1 = m->p.bx1, nx2 = m->p.bx2, nx3 = m->p.bx3;
  if (m->f3) {
    if (fid < 2) return (long)(nx3+1)*nx2 + (long)nx3*(nx2+1);
    if (fid < 4) return (long)(nx3+1)*nx1 + (long)nx3*(nx1+1);
    return (long)(nx2+1)*nx1 + (long)nx2*(nx1+1);
  } else if (m->f2) {
    if (fid < 2) return nx2 + nx2 + 1;
    return nx1 + nx1 + 1;
  }
  return 2;
}
long emf_edge_count(const AbMesh *m, int eid) {
  if (eid < 4) return m->p.bx3;
  if (eid < 8) return m->p.bx2;
  return m->p.bx1;
}
int opposite_eid(int eid) { return (eid & ~3) | ((eid & 3) ^ 3); }

int owner_lid(const AbMesh *m, int gid) { return gid - m->gid_start; }

// ------------------------------------------------------------------ construction
void set_integrator(AbMesh *m) {
  // src/task_list/time_integrator.cpp:104-604 (vl2, rk1, rk2, rk3)
  double cfl_limit = 1.0;
  for (int s = 0; s < 4; ++s) { m->g1[s] = 0; m->g2[s] = 1; m->g3[s] = 0; m->delta[s] = 0; m->beta[s] = 0; }
  m->delta[0] = 1.0;
  for (int s = 0; s < 4; ++s) { m->ebeta[s] = 1.0; m->sbeta[s] = 0.0; }   // stage_wghts[].ebeta/sbeta
  switch (m->p.integrator) {
    case AB_INT_VL2:
      m->nstages = 2; m->beta[0] = 0.5; m->beta[1] = 1.0; m->ebeta[0] = 0.5; m->sbeta[1] = 0.5;
      if (m->ndim >= 2) cfl_limit = 0.5;
      break;
    case AB_INT_RK1:
      m->nstages = 1; m->beta[0] = 1.0;
      break;
    case AB_INT_RK2:
      m->nstages = 2; m->beta[0] = 1.0; m->beta[1] = 0.5; m->sbeta[1] = 1.0;
      m->g1[1] = 0.5; m->g2[1] = 0.5;
      break;
    default:
      m->nstages = 3; m->beta[0] = 1.0; m->beta[1] = 0.25; m->beta[2] = 0.66666666666666667;
      m->ebeta[1] = 0.5; m->sbeta[1] = 1.0; m->sbeta[2] = 0.5;
      m->g1[1] = 0.25; m->g2[1] = 0.75;
      m->g1[2] = 0.66666666666666667; m->g2[2] = 0.33333333333333333;
      break;
  }
  m->cfl = m->p.cfl_number;
  if (m->cfl > cfl_limit) m->cfl = cfl_limit;   // time_integrator.cpp:886-894
}

void build_block_list(AbMesh *m) {
  const AbMeshParams &p = m->p;
  m->nrb[0] = p.nx1/p.bx1; m->nrb[1] = p.nx2/p.bx2; m->nrb[2] = p.nx3/p.bx3;
  m->nbtotal = m->nrb[0]*m->nrb[1]*m->nrb[2];
  // Z-order = Morton order with x1 the fastest bit (meshblock_tree.cpp:87-110,336-352)
  struct Key { unsigned long long key; int i, j, k; };
  std::vector<Key> keys;
  for (int k = 0; k < m->nrb[2]; ++k) for (int j = 0; j < m->nrb[1]; ++j)
    for (int i = 0; i < m->nrb[0]; ++i) {
      unsigned long long key = 0;
      for (int bit = 0; bit < 20; ++bit)
        key |= ((unsigned long long)((i >> bit) & 1) << (3*bit))
             | ((unsigned long long)((j >> bit) & 1) << (3*bit + 1))
             | ((unsigned long long)((k >> bit) & 1) << (3*bit + 2));
      keys.push_back({key, i, j, k});
    }
  std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.key < b.key; });
  m->hb.resize(m->nbtotal);
  m->gid_of.assign(m->nbtotal, -1);
  for (int g = 0; g < m->nbtotal; ++g) {
    HostBlock &B = m->hb[g];
    B.gid = g; B.lx[0] = keys[g].i; B.lx[1] = keys[g].j; B.lx[2] = keys[g].k;
    m->gid_of[(B.lx[2]*m->nrb[1] + B.lx[1])*m->nrb[0] + B.lx[0]] = g;
  }
  // Mesh::CalculateLoadBalance with unit costs (mesh/amr_loadbalance.cpp:72-112)
  {
    int nranks = p.nranks;
    double totalcost = m->nbtotal;
    int j = nranks - 1;
    double targetcost = totalcost/nranks, mycost = 0.0;
    for (int i = m->nbtotal - 1; i >= 0; i--) {
      mycost += 1.0;
      m->hb[i].rank = j;
      if (mycost >= targetcost && j > 0) {
        j--;
        totalcost -= mycost;
        mycost = 0.0;
        targetcost = totalcost/(j + 1);
      }
    }
  }
  // block extents + boundary flags (mesh/mesh.cpp:1668-1751)
  const double mmin[3] = {p.x1min, p.x2min, p.x3min}, mmax[3] = {p.x1max, p.x2max, p.x3max};
  const int nxm[3] = {p.nx1, p.nx2, p.nx3};
  for (auto &B : m->hb) {
    for (int d = 0; d < 3; ++d) {
      if (d > 0 && nxm[d] == 1) {
        B.bmin[d] = mmin[d]; B.bmax[d] = mmax[d];
        B.bcs[2*d] = p.bc[2*d]; B.bcs[2*d+1] = p.bc[2*d+1];
        continue;
      }
      if (B.lx[d] == 0) { B.bmin[d] = mmin[d]; B.bcs[2*d] = p.bc[2*d]; }
      else { B.bmin[d] = gen_x(B.lx[d], m->nrb[d], mmin[d], mmax[d], m->xrat[d], nxm[d]); B.bcs[2*d] = -1; }
      if (B.lx[d] == m->nrb[d] - 1) { B.bmax[d] = mmax[d]; B.bcs[2*d+1] = p.bc[2*d+1]; }
      else { B.bmax[d] = gen_x(B.lx[d] + 1, m->nrb[d], mmin[d], mmax[d], m->xrat[d], nxm[d]); B.bcs[2*d+1] = -1; }
    }
  }
  // canonical neighbour enumeration = buffer ids (bvals/bvals_base.cpp:153-256)
  std::vector<std::array<int,3>> ni;
  for (int n = -1; n <= 1; n += 2) ni.push_back({n, 0, 0});
  if (m->ndim >= 2) for (int n = -1; n <= 1; n += 2) ni.push_back({0, n, 0});
  if (m->ndim == 3) for (int n = -1; n <= 1; n += 2) ni.push_back({0, 0, n});
  if (m->ndim >= 2)
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) ni.push_back({n, mm, 0});
  if (m->ndim == 3) {
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) ni.push_back({n, 0, mm});
    for (int mm = -1; mm <= 1; mm += 2) for (int n = -1; n <= 1; n += 2) ni.push_back({0, n, mm});
    for (int l = -1; l <= 1; l += 2) for (int mm = -1; mm <= 1; mm += 2)
      for (int n = -1; n <= 1; n += 2) ni.push_back({n, mm, l});
  }
  auto find_ni = [&](int a, int b, int c) {
    for (size_t n = 0; n < ni.size(); ++n)
      if (ni[n][0] == a && ni[n][1] == b && ni[n][2] == c) return (int)n;
    return -1;
  };
  // SearchAndSetNeighbors, same level (bvals/bvals_base.cpp:299-480)
  for (auto &B : m->hb) {
    for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
      B.nblevel[k][j][i] = -1;
    B.nblevel[1][1][1] = 0;
    for (size_t n = 0; n < ni.size(); ++n) {
      int o[3] = {ni[n][0], ni[n][1], ni[n][2]};
      long l[3];
      bool ok = true;
      for (int d = 0; d < 3; ++d) {
        l[d] = B.lx[d] + o[d];
        if (l[d] < 0) { if (p.bc[2*d] == AB_BC_PERIODIC) l[d] = m->nrb[d] - 1; else ok = false; }
        if (l[d] >= m->nrb[d]) { if (p.bc[2*d+1] == AB_BC_PERIODIC) l[d] = 0; else ok = false; }
      }
      if (!ok) continue;
      Nb nb;
      nb.ox1 = o[0]; nb.ox2 = o[1]; nb.ox3 = o[2];
      nb.type = (o[0] != 0) + (o[1] != 0) + (o[2] != 0) - 1;
      nb.gid = m->gid_of[(l[2]*m->nrb[1] + l[1])*m->nrb[0] + l[0]];
      nb.rank = m->hb[nb.gid].rank;
      nb.bufid = (int)n;
      nb.targetid = find_ni(-o[0], -o[1], -o[2]);
      nb.fid = -1; nb.eid = -1;
      if (nb.type == 0) {          // NeighborBlock::SetNeighbor (bvals_base.cpp:54-66)
        if (o[0] == -1) nb.fid = 0; else if (o[0] == 1) nb.fid = 1;
        else if (o[1] == -1) nb.fid = 2; else if (o[1] == 1) nb.fid = 3;
        else if (o[2] == -1) nb.fid = 4; else nb.fid = 5;
      } else if (nb.type == 1) {
        if (o[2] == 0) nb.eid = (((o[0] + 1) >> 1) | ((o[1] + 1) & 2));
        else if (o[1] == 0) nb.eid = (4 + (((o[0] + 1) >> 1) | ((o[2] + 1) & 2)));
        else nb.eid = (8 + (((o[1] + 1) >> 1) | ((o[2] + 1) & 2)));
      }
      B.nblevel[o[2]+1][o[1]+1][o[0]+1] = 0;
      B.nbs.push_back(nb);
    }
    // CountFineEdges (bvals/fc/bvals_fc.cpp:1129-1193), single level
    for (int e = 0; e < 12; ++e) B.nedge_fine[e] = 1;
    int eid = 0;
    auto lo = [](int o) { return (o-1 > -1) ? o-1 : -1; };
    auto hi = [](int o) { return (o+1 < 1) ? o+1 : 1; };
    if (m->f2)
      for (int o2 = -1; o2 <= 1; o2 += 2) for (int o1 = -1; o1 <= 1; o1 += 2) {
        int nf = 0;
        for (int nj = lo(o2); nj <= hi(o2); nj++) for (int nii = lo(o1); nii <= hi(o1); nii++)
          if (B.nblevel[1][nj+1][nii+1] == 0) nf++;
        B.nedge_fine[eid++] = nf;
      }
    if (m->f3) {
      for (int o3 = -1; o3 <= 1; o3 += 2) for (int o1 = -1; o1 <= 1; o1 += 2) {
        int nf = 0;
        for (int nk = lo(o3); nk <= hi(o3); nk++) for (int nii = lo(o1); nii <= hi(o1); nii++)
          if (B.nblevel[nk+1][1][nii+1] == 0) nf++;
        B.nedge_fine[eid++] = nf;
      }
      for (int o3 = -1; o3 <= 1; o3 += 2) for (int o2 = -1; o2 <= 1; o2 += 2) {
        int nf = 0;
        for (int nk = lo(o3); nk <= hi(o3); nk++) for (int nj = lo(o2); nj <= hi(o2); nj++)
          if (B.nblevel[nk+1][nj+1][1] == 0) nf++;
        B.nedge_fine[eid++] = nf;
      }
    }
  }
}

// Coordinates ctor (coordinates.cpp:92-160: uniform branch, or the mesh-generator branch with
// dx?f = face differences when x?rat != 1) + Cartesian x?v (cartesian.cpp:25-75)
void make_coords(int nx_mesh, int bx, int ng, long lx, double mmin, double mmax, double bmin,
                 double bmax, int nc, bool refl_in, bool refl_out, double rat,
                 std::vector<double> &xf, std::vector<double> &xv, std::vector<double> &dxf) {
  xf.assign(nc + 1, 0.0); xv.assign(nc, 0.0); dxf.assign(nc, 0.0);
  if (nc == 1) {
    dxf[0] = bmax - bmin;
    xf[0] = bmin; xf[1] = bmax;
    xv[0] = 0.5*(xf[1] + xf[0]);
    return;
  }
  int il = ng, iu = ng + bx - 1;
  double dx = (bmax - bmin)/(iu - il + 1);
  for (int i = il - ng; i <= iu + ng + 1; ++i) {
    long noffset = (long)(i - il) + lx*bx;
    xf[i] = gen_x(noffset, nx_mesh, mmin, mmax, rat, nx_mesh);
  }
  xf[il] = bmin;
  xf[iu+1] = bmax;
  for (int i = il - ng; i <= iu + ng; ++i) dxf[i] = (rat != 1.0) ? xf[i+1] - xf[i] : dx;
  // reflecting boundaries mirror the ghost-zone spacing (coordinates.cpp:147-160)
  if (refl_in) for (int i = 1; i <= ng; ++i) {
    dxf[il-i] = dxf[il+i-1];
    xf[il-i] = xf[il-i+1] - dxf[il-i];
  }
  if (refl_out) for (int i = 1; i <= ng; ++i) {
    dxf[iu+i] = dxf[iu-i+1];
    xf[iu+i+1] = xf[iu+i] + dxf[iu+i];
  }
  for (int i = il - ng; i <= iu + ng; ++i) xv[i] = 0.5*(xf[i+1] + xf[i]);
}

// Geometry of the nonuniform reconstruction along one direction (x?rat != 1): NUG doubles per
// cell index (layout: ab_physics.cuh) -- the factors plm.cpp:85-93,198-204,308-313 evaluate on
// the fly from dx?f, dx?v (= differences of x?v, cartesian.cpp:31-75) and the PPM weights of the
// Reconstruction ctor (reconstruction.cpp:434-461; the x2 / x3 loops :559-582,609-631 start one
// index later than x1's, entries outside keep NewAthenaArray's zero) -- and the
// CalculateCellCenteredField weights (field.cpp:139-172).
void make_recon_table(int dir, int nc, int s, int e, int ng, const std::vector<double> &xf,
                      const std::vector<double> &xv, const std::vector<double> &dxf,
                      std::vector<double> &tab, std::vector<double> &lw, std::vector<double> &rw) {
  tab.assign((size_t)nc*ab::NUG, 0.0);
  lw.assign(nc, 0.5); rw.assign(nc, 0.5);
  std::vector<double> dxv(nc, 0.0);
  for (int i = s - ng; i <= e + ng - 1; ++i) dxv[i] = xv[i+1] - xv[i];
  for (int i = 0; i < nc; ++i) {
    double *t = &tab[(size_t)i*ab::NUG];
    const double dvm = (i > 0) ? dxv[i-1] : 0.0;
    t[0] = dxf[i]; t[1] = dxv[i]; t[2] = dvm;
    t[3] = dxv[i]/(xf[i+1] - xv[i]);         // cf (Mignone eq 33)
    t[4] = dvm/(xv[i] - xf[i]);              // cb
    t[5] = dxf[i]/dxv[i];
    t[6] = dxf[i]/dvm;
    lw[i] = (xf[i+1] - xv[i])/dxf[i];
    rw[i] = (xv[i] - xf[i])/dxf[i];
  }
  const int first = (dir == 0) ? s - ng + 1 : s - ng + 2;
  for (int i = first; i <= e + ng - 1; ++i) {
    double *t = &tab[(size_t)i*ab::NUG];
    const double dm1 = dxf[i-1], d0 = dxf[i], dp1 = dxf[i+1];
    const double qe = d0/(dm1 + d0 + dp1);               // CW eq 1.7
    t[7] = qe*(2.0*dm1+d0)/(dp1 + d0);
    t[8] = qe*(2.0*dp1+d0)/(dm1 + d0);
    if (i > s - ng + 1) {
      const double dm2 = dxf[i-2];
      const double qa = dm2 + dm1 + d0 + dp1;
      double qb = dm1/(dm1 + d0);
      const double qc = (dm2 + dm1)/(2.0*dm1 + d0);
      const double qd = (dp1 + d0)/(2.0*d0 + dm1);
      qb = qb + 2.0*d0*qb/qa*(qc-qd);
      t[9] = 1.0 - qb;
      t[10] = qb;
      t[11] = d0/qa*qd;
      t[12] = -dm1/qa*qc;
    }
  }
}

// cell centres of the MeshRefinement's coarse Coordinates along direction d (coarse_flag branch of
// coordinates.cpp:92-160 + cartesian.cpp:25-75): faces from the mesh generator at every second
// fine index, block edges pinned, reflecting ghost spacing mirrored
std::vector<double> coarse_centres(const AbMesh *m, const HostBlock &B, int d) {
  const AbMeshParams &p = m->p;
  const double mmin[3] = {p.x1min, p.x2min, p.x3min}, mmax[3] = {p.x1max, p.x2max, p.x3max};
  const int nxm[3] = {p.nx1, p.nx2, p.nx3}, bxs[3] = {p.bx1, p.bx2, p.bx3};
  const bool fdim[3] = {true, (bool)m->f2, (bool)m->f3};
  const int cng = (p.nghost + 1)/2 + 1;
  const int cnc = fdim[d] ? bxs[d]/2 + 2*cng : 1;
  std::vector<double> xf(cnc + 1, 0.0), xv(cnc, 0.0);
  if (cnc == 1) {
    xf[0] = B.bmin[d]; xf[1] = B.bmax[d]; xv[0] = 0.5*(xf[1] + xf[0]);
    return xv;
  }
  const int il = cng, iu = cng + bxs[d]/2 - 1;
  const long nroot = (long)nxm[d] << B.dl;
  std::vector<double> dxf(cnc, (B.bmax[d] - B.bmin[d])/(iu - il + 1));
  for (int i = il - cng; i <= iu + cng + 1; ++i)
    xf[i] = gen_x((long)(i - il)*2 + B.lx[d]*bxs[d], nroot, mmin[d], mmax[d], 1.0, nxm[d]);
  xf[il] = B.bmin[d]; xf[iu+1] = B.bmax[d];
  if (B.bcs[2*d] == AB_BC_REFLECT) for (int i = 1; i <= cng; ++i) {
    dxf[il-i] = dxf[il+i-1]; xf[il-i] = xf[il-i+1] - dxf[il-i]; }
  if (B.bcs[2*d+1] == AB_BC_REFLECT) for (int i = 1; i <= cng; ++i) {
    dxf[iu+i] = dxf[iu-i+1]; xf[iu+i+1] = xf[iu+i] + dxf[iu+i]; }
  for (int i = il - cng; i <= iu + cng; ++i) xv[i] = 0.5*(xf[i+1] + xf[i]);
  return xv;
}

long reg_size(const AbMesh *m, int reg) {
  long n1 = m->nc[0], n2 = m->nc[1], n3 = m->nc[2];
  long ncc = n1*n2*n3;
  long nf1 = n3*n2*(n1+1), nf2 = n3*(n2+1)*n1, nf3 = (n3+1)*n2*n1;
  switch (reg) {
    case AB_U: case AB_U1: case AB_W: return m->nh*ncc;
    case AB_FLUX_X1: return m->nh*nf1;
    case AB_FLUX_X2: return m->nh*nf2;
    case AB_FLUX_X3: return m->nh*nf3;
    case AB_S: case AB_S1: case AB_R: return m->p.nscalars*ncc;
    case AB_SFLUX_X1: return m->p.nscalars*nf1;
    case AB_SFLUX_X2: return m->p.nscalars*nf2;
    case AB_SFLUX_X3: return m->p.nscalars*nf3;
    default: break;
  }
  if (!m->p.mhd) return 0;
  switch (reg) {
    case AB_BCC: return 3*ncc;
    case AB_B_X1F: case AB_B1_X1F: case AB_WGHT_X1F: case AB_E3_X1F: case AB_E2_X1F: return nf1;
    case AB_B_X2F: case AB_B1_X2F: case AB_WGHT_X2F: case AB_E1_X2F: case AB_E3_X2F: return nf2;
    case AB_B_X3F: case AB_B1_X3F: case AB_WGHT_X3F: case AB_E2_X3F: case AB_E1_X3F: return nf3;
    case AB_E_X1E: return (n3+1)*(n2+1)*n1;
    case AB_E_X2E: return (n3+1)*n2*(n1+1);
    case AB_E_X3E: return n3*(n2+1)*(n1+1);
    default: return 0;
  }
}

double **reg_slot(LocalBlock &L, int reg) {
  ab::BlkDev &d = L.d;
  switch (reg) {
    case AB_U: return &d.u; case AB_U1: return &d.u1; case AB_W: return &d.w;
    case AB_BCC: return &d.bcc;
    case AB_B_X1F: return &d.b[0]; case AB_B_X2F: return &d.b[1]; case AB_B_X3F: return &d.b[2];
    case AB_B1_X1F: return &d.b1[0]; case AB_B1_X2F: return &d.b1[1]; case AB_B1_X3F: return &d.b1[2];
    case AB_FLUX_X1: return &d.flux[0]; case AB_FLUX_X2: return &d.flux[1];
    case AB_FLUX_X3: return &d.flux[2];
    case AB_E_X1E: return &d.e[0]; case AB_E_X2E: return &d.e[1]; case AB_E_X3E: return &d.e[2];
    case AB_WGHT_X1F: return &d.wght[0]; case AB_WGHT_X2F: return &d.wght[1];
    case AB_WGHT_X3F: return &d.wght[2];
    case AB_E3_X1F: return &d.ef[0][0]; case AB_E2_X1F: return &d.ef[0][1];
    case AB_E1_X2F: return &d.ef[1][0]; case AB_E3_X2F: return &d.ef[1][1];
    case AB_E2_X3F: return &d.ef[2][0]; case AB_E1_X3F: return &d.ef[2][1];
    case AB_S: return &d.s; case AB_S1: return &d.s1; case AB_R: return &d.r;
    case AB_SFLUX_X1: return &d.sflux[0]; case AB_SFLUX_X2: return &d.sflux[1];
    case AB_SFLUX_X3: return &d.sflux[2];
    default: return nullptr;
  }
}

size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

int alloc_blocks(AbMesh *m) {
  const AbMeshParams &p = m->p;
  int ng = p.nghost;
  for (auto *pB : m->lb_hb) {
    LocalBlock L;
    L.hb = pB;
    m->lb.push_back(L);
  }
  if (m->lb.empty()) return fail(AB_ERR_ARG, "this rank owns no MeshBlock: use fewer ranks or smaller MeshBlocks");
  CK(cudaMalloc(&m->dtmin, sizeof(unsigned long long)*m->lb.size()*ab::DT_SLOTS));
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    HostBlock &B = *L.hb;
    memset(&L.d, 0, sizeof(L.d));
    ab::BlkDev &d = L.d;
    d.nc1 = m->nc[0]; d.nc2 = m->nc[1]; d.nc3 = m->nc[2];
    d.is = m->is; d.ie = m->ie; d.js = m->js; d.je = m->je; d.ks = m->ks; d.ke = m->ke;
    d.ng = ng; d.f2 = m->f2; d.f3 = m->f3;
    d.ns = p.nscalars;
    d.nh = m->nh;
    // sizes
    size_t tot = 0;
    for (int r = 0; r < AB_NREG; ++r) { L.regsize[r] = reg_size(m, r); tot += align256(L.regsize[r]*8); }
    long ncc = (long)m->nc[0]*m->nc[1]*m->nc[2];
    size_t cce = p.mhd ? align256(3*ncc*8) : 0;
    tot += cce;
    // coordinates: 9 arrays + 6 PLM weight arrays (+ the nonuniform-spacing tables)
    std::vector<double> xf[3], xv[3], dxf[3], wp[3], wm[3], nutab[3], bcw;
    bool any_nu = false;
    const double mmin[3] = {p.x1min, p.x2min, p.x3min}, mmax[3] = {p.x1max, p.x2max, p.x3max};
    const int nxm[3] = {p.nx1, p.nx2, p.nx3}, bxs[3] = {p.bx1, p.bx2, p.bx3};
    for (int dd = 0; dd < 3; ++dd) {
      make_coords(nxm[dd] << B.dl, bxs[dd], ng, B.lx[dd], mmin[dd], mmax[dd], B.bmin[dd], B.bmax[dd],
                  m->nc[dd], B.bcs[2*dd] == AB_BC_REFLECT, B.bcs[2*dd+1] == AB_BC_REFLECT,
                  m->xrat[dd], xf[dd], xv[dd], dxf[dd]);
      wp[dd].assign(m->nc[dd], 0.0); wm[dd].assign(m->nc[dd], 0.0);
      for (int c = 0; c < m->nc[dd]; ++c) {   // plm.cpp:114-119 / 226-227 / 332-333
        wp[dd][c] = (xf[dd][c+1] - xv[dd][c])/dxf[dd][c];
        wm[dd][c] = (xv[dd][c] - xf[dd][c])/dxf[dd][c];
      }
      std::vector<double> lw(m->nc[dd], 0.5), rw(m->nc[dd], 0.5);
      if (m->xrat[dd] != 1.0) {
        const int s0[3] = {m->is, m->js, m->ks}, e0[3] = {m->ie, m->je, m->ke};
        make_recon_table(dd, m->nc[dd], s0[dd], e0[dd], ng, xf[dd], xv[dd], dxf[dd], nutab[dd],
                         lw, rw);
        any_nu = true;
      }
      bcw.insert(bcw.end(), lw.begin(), lw.end());
      bcw.insert(bcw.end(), rw.begin(), rw.end());
    }
    size_t coord_bytes = 0;
    for (int dd = 0; dd < 3; ++dd)
      coord_bytes += align256((m->nc[dd]+1)*8) + 4*align256(m->nc[dd]*8) +
                     align256(nutab[dd].size()*8);
    if (any_nu) coord_bytes += align256(bcw.size()*8);
    tot += coord_bytes;
    // EMF send buffers for every face / edge (used for same-rank neighbours)
    size_t emf_elems = 0;
    if (p.mhd) {
      for (int f = 0; f < 6; ++f) { L.face_off[f] = emf_elems; emf_elems += emf_face_count(m, f); }
      for (int e = 0; e < 12; ++e) { L.edge_off[e] = emf_elems; emf_elems += emf_edge_count(m, e); }
    }
    tot += align256(emf_elems*8);
    L.nbytes = tot;
    // AB_DEBUG_ALLOC=1: one cudaMalloc per array instead of one slab per block, so that
    // compute-sanitizer memcheck sees every array's bounds (debug aid; same layout otherwise)
    const char *dbg = getenv("AB_DEBUG_ALLOC");
    const bool split_alloc = dbg && dbg[0] == '1';
    char *cur = nullptr;
    if (!split_alloc) {
      // one allocation for all local blocks, slabs bstride bytes apart with identical layouts:
      // a task can then run over every block in ONE launch (ab_batch.cuh)
      if (l == 0) {
        m->bstride = (long)align256(tot);
        CK(cudaMalloc(&m->slab, (size_t)m->bstride*m->lb.size()));
      }
      if ((long)align256(tot) != m->bstride) return fail(AB_ERR_STATE, "MeshBlock slabs differ in size");
      L.base = m->slab + (size_t)l*m->bstride;
      CK(cudaMemsetAsync(L.base, 0, tot, m->stream));   // AthenaArray storage is zero-initialised
      cur = (char *)L.base;
      d.bstride = m->bstride;
    }
    auto carve = [&](size_t bytes) -> char * {
      if (!split_alloc) { char *q = cur; cur += align256(bytes); return q; }
      void *q = nullptr;
      cudaMalloc(&q, bytes ? bytes : 8);
      cudaMemsetAsync(q, 0, bytes ? bytes : 8, m->stream);
      L.debug_allocs.push_back(q);
      return (char *)q;
    };
    for (int r = 0; r < AB_NREG; ++r) {
      double **slot = reg_slot(L, r);
      if (L.regsize[r] > 0) *slot = (double *)carve(L.regsize[r]*8);
      else *slot = nullptr;
    }
    if (p.mhd) d.cc_e = (double *)carve(3*ncc*8);
    auto put = [&](const std::vector<double> &v, int idx) -> const double * {
      double *dst = (double *)carve(v.size()*8);
      cudaMemcpyAsync(dst, v.data(), v.size()*8, cudaMemcpyHostToDevice, m->stream);
      if (idx >= 0) { L.coord_dev[idx] = dst; L.coord_n[idx] = (long)v.size(); }
      return dst;
    };
    d.x1f = put(xf[0], 0); d.x2f = put(xf[1], 1); d.x3f = put(xf[2], 2);
    d.x1v = put(xv[0], 3); d.x2v = put(xv[1], 4); d.x3v = put(xv[2], 5);
    d.dx1f = put(dxf[0], 6); d.dx2f = put(dxf[1], 7); d.dx3f = put(dxf[2], 8);
    for (int dd = 0; dd < 3; ++dd) { L.g.wp[dd] = put(wp[dd], -1); L.g.wm[dd] = put(wm[dd], -1); }
    for (int dd = 0; dd < 3; ++dd) L.g.nu[dd] = nutab[dd].empty() ? nullptr : put(nutab[dd], -1);
    d.bcw = any_nu ? put(bcw, -1) : nullptr;
    CK(cudaStreamSynchronize(m->stream));   // host vectors go out of scope
    L.emf_send = (double *)carve(emf_elems*8);
    L.dtmin = m->dtmin + l*ab::DT_SLOTS;
    L.has_phys_bc = false;
    for (int f = 0; f < 2*m->ndim; ++f)
      if (B.bcs[f] != -1 && B.bcs[f] != AB_BC_PERIODIC) L.has_phys_bc = true;
  }
  return AB_OK;
}

// ------------------------------------------------------------------ exchange plans
// Messages this rank exchanges with every peer rank, per exchange kind (0: ghost zones of u
// [and b], 1: EMF correction).  Both sides order a peer's messages by (destination gid,
// destination buffer id), so one concatenated buffer per peer can be sent with a single
// ncclSend and split identically by the receiver (the reference tags each message with
// (lid, bufid), bvals_base.cpp:287; here the position in the buffer plays that role).
void peer_messages(const AbMesh *m, int kind, std::map<int, std::vector<Msg>> &sends,
                   std::map<int, std::vector<Msg>> &recvs) {
  for (size_t l = 0; l < m->lb_hb.size(); ++l) {
    const HostBlock &B = *m->lb_hb[l];
    for (size_t n = 0; n < B.nbs.size(); ++n) {
      const Nb &nb = B.nbs[n];
      if (nb.rank == m->p.rank) continue;
      long cnt;
      if (kind == 0) {
        cnt = state_msg_count(m, nb.ox1, nb.ox2, nb.ox3);
      } else {
        if (nb.type > 1) continue;
        cnt = nb.type == 0 ? emf_face_count(m, nb.fid) : emf_edge_count(m, nb.eid);
      }
      sends[nb.rank].push_back({nb.rank, (long)nb.gid*64 + nb.targetid, (int)l, (int)n, cnt});
      recvs[nb.rank].push_back({nb.rank, (long)B.gid*64 + nb.bufid, (int)l, (int)n, cnt});
    }
  }
  auto by_key = [](const Msg &a, const Msg &b) { return a.key < b.key; };
  for (auto &kv : sends) std::sort(kv.second.begin(), kv.second.end(), by_key);
  for (auto &kv : recvs) std::sort(kv.second.begin(), kv.second.end(), by_key);
}

// ghost-exchange plan for the current register parity: box copies, pack lists, peer buffers
// lid_filter >= 0: only the ghost zones (and send buffers) of that local block; vars: bit 0 the
// hydro registers, bit 1 the face field, bit 2 the passive scalars (the per-block, per-variable
// plans behind ab_bvals_send / ab_bvals_set).  Buffer offsets never depend on the filters.
int build_state_plan(AbMesh *m, AbMesh::Plan &P, int lid_filter = -1, int vars = 7, int p2p_buf = -1) {
  std::vector<CopyBox> pack, ph1, ph1r, ph2;
  const int mhd = m->p.mhd;
  const long ncc = (long)m->nc[0]*m->nc[1]*m->nc[2];
  // per-peer message lists (sorted by (dst gid, dst bufid))
  std::map<int, std::vector<Msg>> sends, recvs;
  peer_messages(m, 0, sends, recvs);
  std::map<std::pair<int,int>, long> send_off, recv_off;   // (lid, nbi) -> element offset
  for (auto &kv : sends) {
    long off = 0;
    for (auto &ms : kv.second) { send_off[{ms.lid, ms.nbi}] = off; off += ms.count; }
    PeerBuf &pb = m->peer_state[kv.first];
    if (!pb.send) { pb.nsend = off; CK(cudaMalloc(&pb.send, std::max<size_t>(off, 1)*8)); }
  }
  for (auto &kv : recvs) {
    long off = 0;
    for (auto &ms : kv.second) { recv_off[{ms.lid, ms.nbi}] = off; off += ms.count; }
    PeerBuf &pb = m->peer_state[kv.first];
    if (!pb.recv) { pb.nrecv = off; CK(cudaMalloc(&pb.recv, std::max<size_t>(off, 1)*8)); }
  }
  // where a peer's messages land / are read: the NCCL staging buffers, or (direct exchange) the
  // peer's own receive buffer of this round mapped into this process and mine
  auto send_base = [&](int rank) -> double * {
    PeerBuf &pb = m->peer_state[rank];
    return p2p_buf >= 0 ? pb.p2p_send[p2p_buf] : pb.send;
  };
  auto recv_base = [&](int rank) -> double * {
    PeerBuf &pb = m->peer_state[rank];
    return p2p_buf >= 0 ? pb.p2p_recv[p2p_buf] : pb.recv;
  };
  auto add_box = [](std::vector<CopyBox> &v, double *dst, long ds3, long ds2, long dsv,
                    const double *src, long ss3, long ss2, long ssv, int nvar, const Box &db,
                    int si0, int sj0, int sk0) {
    CopyBox c;
    c.dst = dst; c.src = src; c.dst_s3 = ds3; c.dst_s2 = ds2; c.src_s3 = ss3; c.src_s2 = ss2;
    c.dst_sv = dsv; c.src_sv = ssv; c.nvar = nvar;
    c.di0 = db.si; c.dj0 = db.sj; c.dk0 = db.sk; c.si0 = si0; c.sj0 = sj0; c.sk0 = sk0;
    c.ni = db.ei-db.si+1; c.nj = db.ej-db.sj+1; c.nk = db.ek-db.sk+1; c.offset = 0;
    if (c.ni > 0 && c.nj > 0 && c.nk > 0) v.push_back(c);
  };
  const long cs2 = m->nc[0], cs3 = (long)m->nc[0]*m->nc[1];
  const bool v_hyd = vars & 1, v_fld = (vars & 2) && mhd, v_scl = (vars & 4) && m->p.nscalars > 0;
  for (size_t l = 0; l < m->lb.size(); ++l) {
    if (lid_filter >= 0 && (int)l != lid_filter) continue;
    LocalBlock &L = m->lb[l];
    HostBlock &B = *L.hb;
    for (size_t n = 0; n < B.nbs.size(); ++n) {
      const Nb &nb = B.nbs[n];
      const bool local = (nb.rank == m->p.rank);
      // ---- receiving side: fill my ghost zones from the neighbour's active zones
      Box rb = cc_recv_box(m, nb.ox1, nb.ox2, nb.ox3);
      Box sb = cc_send_box(m, -nb.ox1, -nb.ox2, -nb.ox3);      // what the neighbour loads
      if (!v_hyd) {
      } else if (local) {
        LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
        add_box(ph1, L.d.u, cs3, cs2, ncc, N.d.u, cs3, cs2, ncc, m->nh, rb, sb.si, sb.sj, sb.sk);
      } else {
        double *src = recv_base(nb.rank) + recv_off[{(int)l, (int)n}];
        Box z = {0, rb.ei-rb.si, 0, rb.ej-rb.sj, 0, rb.ek-rb.sk};
        long s2 = z.ei+1, s3 = s2*(z.ej+1);
        add_box(ph1r, L.d.u, cs3, cs2, ncc, src, s3, s2, s3*(z.ek+1), m->nh, rb, 0, 0, 0);
      }
      long roff = m->nh*rb.count();
      const int ns = m->p.nscalars;
      if (mhd) {
        for (int c = 0; c < 3; ++c) {
          Box frb = fc_recv_box(m, c, nb.ox1, nb.ox2, nb.ox3);
          Box fsb = fc_send_box(m, c, -nb.ox1, -nb.ox2, -nb.ox3);
          long s3, s2;
          fc_strides(m, c, s3, s2);
          if (local) {
            LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
            if (v_fld) add_box(ph1, L.d.b[c], s3, s2, 0, N.d.b[c], s3, s2, 0, 1, frb, fsb.si, fsb.sj, fsb.sk);
          } else {
            double *src = recv_base(nb.rank) + recv_off[{(int)l, (int)n}] + roff;
            long t2 = frb.ei-frb.si+1, t3 = t2*(frb.ej-frb.sj+1);
            if (v_fld) add_box(ph1r, L.d.b[c], s3, s2, 0, src, t3, t2, 0, 1, frb, 0, 0, 0);
            roff += frb.count();
          }
          if (!v_fld) continue;
          // 1-D / 2-D duplicate faces (bvals_fc.cpp:641-645, 669-675)
          if (c == 1 && !m->f2) {
            Box dup = frb; dup.sj = frb.sj + 1; dup.ej = frb.sj + 1;
            add_box(ph2, L.d.b[1], s3, s2, 0, L.d.b[1], s3, s2, 0, 1, dup, frb.si, frb.sj, frb.sk);
          }
          if (c == 2 && !m->f3) {
            Box dup = frb; dup.sk = frb.sk + 1; dup.ek = frb.sk + 1;
            add_box(ph2, L.d.b[2], s3, s2, 0, L.d.b[2], s3, s2, 0, 1, dup, frb.si, frb.sj, frb.sk);
          }
        }
      }
      if (v_scl) {    // PassiveScalars::sbvar: same boxes as u (scalars.cpp:59-72)
        if (local) {
          LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
          add_box(ph1, L.d.s, cs3, cs2, ncc, N.d.s, cs3, cs2, ncc, ns, rb, sb.si, sb.sj, sb.sk);
        } else {
          double *src = recv_base(nb.rank) + recv_off[{(int)l, (int)n}] + roff;
          Box z = {0, rb.ei-rb.si, 0, rb.ej-rb.sj, 0, rb.ek-rb.sk};
          long s2 = z.ei+1, s3 = s2*(z.ej+1);
          add_box(ph1r, L.d.s, cs3, cs2, ncc, src, s3, s2, s3*(z.ek+1), ns, rb, 0, 0, 0);
        }
      }
      // ---- sending side (remote only): pack my active zones into the peer buffer
      if (!local) {
        double *dst = send_base(nb.rank) + send_off[{(int)l, (int)n}];
        Box lb2 = cc_send_box(m, nb.ox1, nb.ox2, nb.ox3);
        Box z = {0, lb2.ei-lb2.si, 0, lb2.ej-lb2.sj, 0, lb2.ek-lb2.sk};
        long s2 = z.ei+1, s3 = s2*(z.ej+1);
        if (v_hyd) add_box(pack, dst, s3, s2, s3*(z.ek+1), L.d.u, cs3, cs2, ncc, m->nh, z, lb2.si, lb2.sj, lb2.sk);
        long soff = m->nh*lb2.count();
        if (mhd) for (int c = 0; c < 3; ++c) {
          Box fb = fc_send_box(m, c, nb.ox1, nb.ox2, nb.ox3);
          long fs3, fs2;
          fc_strides(m, c, fs3, fs2);
          Box fz = {0, fb.ei-fb.si, 0, fb.ej-fb.sj, 0, fb.ek-fb.sk};
          long t2 = fz.ei+1, t3 = t2*(fz.ej+1);
          if (v_fld) add_box(pack, dst + soff, t3, t2, 0, L.d.b[c], fs3, fs2, 0, 1, fz, fb.si, fb.sj, fb.sk);
          soff += fb.count();
        }
        if (v_scl)
          add_box(pack, dst + soff, s3, s2, s3*(z.ek+1), L.d.s, cs3, cs2, ncc, ns, z, lb2.si, lb2.sj, lb2.sk);
      }
    }
  }
  auto upload = [&](std::vector<CopyBox> &v, CopyBox *&dev, int &n, long &mx) -> int {
    n = (int)v.size(); mx = 0;     // mx = total element count; offsets = exclusive prefix
    for (auto &c : v) { c.offset = mx; mx += (long)c.ni*c.nj*c.nk*c.nvar; ab::set_box_divisors(c); }
    if (dev) { cudaFree(dev); dev = nullptr; }
    if (n == 0) return AB_OK;
    // behind the boxes: for every chunk of 256 consecutive elements the box of its first element
    // (k_copy_boxes scans forward from it)
    std::vector<int> first((size_t)((mx + 255)/256), 0);
    {
      int q = 0;
      for (size_t c = 0; c < first.size(); ++c) {
        const long t = (long)c*256;
        while (q + 1 < n && v[q + 1].offset <= t) ++q;
        first[c] = q;
      }
    }
    CK(cudaMalloc(&dev, sizeof(CopyBox)*n + sizeof(int)*first.size()));
    CK(cudaMemcpyAsync(dev + n, first.data(), sizeof(int)*first.size(), cudaMemcpyHostToDevice, m->stream));
    // Stream-ordered copy + sync.  A plain cudaMemcpy from pageable memory may return before the
    // DMA has landed and orders only against the NULL stream; the kernels that read this table run
    // on a non-blocking stream and occasionally saw it half-written (flaky illegal address with
    // >64 KB tables, i.e. many MeshBlocks).
    CK(cudaMemcpyAsync(dev, v.data(), sizeof(CopyBox)*n, cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    return AB_OK;
  };
  int rc;
  if ((rc = upload(pack, P.pack, P.npack, P.maxpack))) return rc;
  if ((rc = upload(ph1, P.phase1, P.n1, P.max1))) return rc;
  if ((rc = upload(ph1r, P.phase1r, P.n1r, P.max1r))) return rc;
  if ((rc = upload(ph2, P.phase2, P.n2, P.max2))) return rc;
  P.built = true;
  return AB_OK;
}

int build_emf_plan(AbMesh *m) {
  if (!m->p.mhd) { m->emf_built = true; return AB_OK; }
  std::map<int, std::vector<Msg>> sends, recvs;
  peer_messages(m, 1, sends, recvs);
  std::map<std::pair<int,int>, long> send_off, recv_off;
  for (auto &kv : sends) {
    long off = 0;
    for (auto &ms : kv.second) { send_off[{ms.lid, ms.nbi}] = off; off += ms.count; }
    PeerBuf &pb = m->peer_emf[kv.first];
    pb.nsend = off; CK(cudaMalloc(&pb.send, std::max<size_t>(off, 1)*8));
  }
  for (auto &kv : recvs) {
    long off = 0;
    for (auto &ms : kv.second) { recv_off[{ms.lid, ms.nbi}] = off; off += ms.count; }
    PeerBuf &pb = m->peer_emf[kv.first];
    pb.nrecv = off; CK(cudaMalloc(&pb.recv, std::max<size_t>(off, 1)*8));
  }
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    HostBlock &B = *L.hb;
    ab::EmfPlan &E = L.emf;
    memset(&E, 0, sizeof(E));
    for (int e = 0; e < 12; ++e) E.nedge_fine[e] = B.nedge_fine[e];
    for (int f = 0; f < 2*m->ndim; ++f)
      E.face_avg[f] = (B.bcs[f] == -1 || B.bcs[f] == AB_BC_PERIODIC) ? 1 : 0;
    for (size_t n = 0; n < B.nbs.size(); ++n) {
      const Nb &nb = B.nbs[n];
      if (nb.type > 1) continue;
      const bool local = (nb.rank == m->p.rank);
      if (nb.type == 0) {
        if (local) {
          LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
          E.face_dst[nb.fid] = L.emf_send + L.face_off[nb.fid];
          E.face_src[nb.fid] = N.emf_send + N.face_off[nb.fid ^ 1];
        } else {
          E.face_dst[nb.fid] = m->peer_emf[nb.rank].send + send_off[{(int)l, (int)n}];
          E.face_src[nb.fid] = m->peer_emf[nb.rank].recv + recv_off[{(int)l, (int)n}];
        }
      } else {
        if (local) {
          LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
          E.edge_dst[nb.eid] = L.emf_send + L.edge_off[nb.eid];
          E.edge_src[nb.eid] = N.emf_send + N.edge_off[opposite_eid(nb.eid)];
        } else {
          E.edge_dst[nb.eid] = m->peer_emf[nb.rank].send + send_off[{(int)l, (int)n}];
          E.edge_src[nb.eid] = m->peer_emf[nb.rank].recv + recv_off[{(int)l, (int)n}];
        }
      }
    }
  }
  // the plans of all local blocks in device memory, for the one-launch-over-all-blocks path
  {
    std::vector<ab::EmfPlan> all;
    for (auto &L : m->lb) all.push_back(L.emf);
    if (!m->emf_plans_dev) CK(cudaMalloc(&m->emf_plans_dev, all.size()*sizeof(ab::EmfPlan)));
    CK(cudaMemcpyAsync(m->emf_plans_dev, all.data(), all.size()*sizeof(ab::EmfPlan),
                       cudaMemcpyHostToDevice, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  m->emf_built = true;
  return AB_OK;
}

int peer_exchange(AbMesh *m, std::map<int, PeerBuf> &peers, cudaStream_t st = nullptr) {
  if (!st) st = m->stream;
  if (peers.empty()) return AB_OK;
  if (!m->comm) return fail(AB_ERR_STATE, "blocks on other ranks but ab_comm_init was not called");
  NK(g_nccl.GroupStart());
  for (auto &kv : peers) {
    if (kv.second.nsend) NK(g_nccl.Send(kv.second.send, kv.second.nsend, NCCL_FLOAT64, kv.first, m->comm, st));
    if (kv.second.nrecv) NK(g_nccl.Recv(kv.second.recv, kv.second.nrecv, NCCL_FLOAT64, kv.first, m->comm, st));
  }
  NK(g_nccl.GroupEnd());
  return AB_OK;
}

// Direct ghost-zone exchange (AB_P2P=1): every rank exports two receive buffers per peer with
// cudaIpcGetMemHandle, the handles travel over NCCL, the peers map them.  Collective: every rank
// gets here in its first whole-mesh exchange.  Any failure on any rank -> all fall back to NCCL.
int p2p_setup(AbMesh *m) {
#ifdef AB_HOST_EMU
  m->p2p = -1;
  return AB_OK;
#else
  if (!m->comm) { m->p2p = -1; return AB_OK; }
  std::map<int, std::vector<Msg>> sends, recvs;
  peer_messages(m, 0, sends, recvs);
  double ok = 1.0;
  std::map<int, std::array<cudaIpcMemHandle_t, 2>> mine, theirs;
  for (auto &kv : recvs) {
    long n = 0;
    for (auto &ms : kv.second) n += ms.count;
    PeerBuf &pb = m->peer_state[kv.first];
    for (int b = 0; b < 2; ++b) {
      if (cudaMalloc(&pb.p2p_recv[b], std::max<size_t>(n, 1)*8) != cudaSuccess ||
          cudaIpcGetMemHandle(&mine[kv.first][b], pb.p2p_recv[b]) != cudaSuccess) ok = 0.0;
    }
  }
  cudaGetLastError();
  // handles: 2 x 64 bytes per peer, moved as 16 doubles
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
  const size_t npeer = recvs.size();
  double *stage = nullptr;
  CK(cudaMalloc(&stage, std::max<size_t>(npeer, 1)*2*128));
  {
    size_t q = 0;
    for (auto &kv : recvs) {
      CK(cudaMemcpyAsync((char *)stage + q*256, mine[kv.first].data(), 128, cudaMemcpyHostToDevice, m->stream));
      ++q;
    }
    NK(g_nccl.GroupStart());
    q = 0;
    for (auto &kv : recvs) {
      NK(g_nccl.Send((char *)stage + q*256, 16, NCCL_FLOAT64, kv.first, m->comm, m->stream));
      NK(g_nccl.Recv((char *)stage + q*256 + 128, 16, NCCL_FLOAT64, kv.first, m->comm, m->stream));
      ++q;
    }
    NK(g_nccl.GroupEnd());
    q = 0;
    for (auto &kv : recvs) {
      CK(cudaMemcpyAsync(theirs[kv.first].data(), (char *)stage + q*256 + 128, 128, cudaMemcpyDeviceToHost, m->stream));
      ++q;
    }
    CK(cudaStreamSynchronize(m->stream));
  }
  cudaFree(stage);
  if (ok != 0.0) for (auto &kv : recvs) {
    PeerBuf &pb = m->peer_state[kv.first];
    for (int b = 0; b < 2; ++b)
      if (cudaIpcOpenMemHandle((void **)&pb.p2p_send[b], theirs[kv.first][b],
                               cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.0; pb.p2p_send[b] = nullptr; }
  }
  cudaGetLastError();
  if (!m->p2p_token) CK(cudaMalloc(&m->p2p_token, 8));
  CK(cudaMemcpyAsync(m->p2p_token, &ok, 8, cudaMemcpyHostToDevice, m->stream));
  NK(g_nccl.AllReduce(m->p2p_token, m->p2p_token, 1, NCCL_FLOAT64, NCCL_MIN, m->comm, m->stream));
  double all_ok = 0.0;
  CK(cudaMemcpyAsync(&all_ok, m->p2p_token, 8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  m->p2p = (all_ok != 0.0) ? 2 : -1;
  if (const char *e = getenv("AB_P2P_VERBOSE")) if (e[0] == '1')
    fprintf(stderr, "[athena_b200] rank %d: direct ghost-zone exchange over peer memory %s (%zu peers)\n",
            m->p.rank, m->p2p == 2 ? "ACTIVE" : "unavailable, using NCCL", npeer);
  return AB_OK;
#endif
}

int plan_index(AbMesh *m) {
  // all local blocks swap registers in lockstep inside the driver; a mixed state (possible
  // only through per-block ab_swap calls) forces a rebuild
  int pu = m->lb[0].parity, pb = m->lb[0].parity_b, ps = m->lb[0].parity_s;
  for (auto &L : m->lb) if (L.parity != pu || L.parity_b != pb || L.parity_s != ps) return -1;
  return pu | (pb << 1) | (ps << 2);
}

int p2p_setup(AbMesh *m);

int bvals_exchange(AbMesh *m) {
  int idx = plan_index(m);
  if (m->p2p == 1) { int rc = p2p_setup(m); if (rc) return rc; }   // first exchange: map the peers
  // Direct exchange: the pack kernel stores into the peers' receive buffers over NVLink; one
  // 8-byte all-reduce is the barrier between everybody's stores and everybody's unpack.  The two
  // receive buffers alternate by round: a rank that is a whole round ahead writes the buffer its
  // peer is not reading (it cannot be two ahead: the barrier in between needs the peer).
  const bool direct = (m->p2p == 2);
  const int buf = direct ? (int)(m->p2p_round & 1) : -1;
  int use = idx < 0 ? 0 : idx;
  if (direct) use = 8 + (use | (buf << 3));
  if (idx < 0) m->plan[use].built = false;
  if (!m->plan[use].built) { int rc = build_state_plan(m, m->plan[use], -1, 7, buf); if (rc) return rc; }
  if (idx < 0) m->plan[use].built = false;   // mixed state: never cache
  AbMesh::Plan &P = m->plan[use];
  if (P.npack) ab::launch_copy_boxes(P.pack, P.npack, P.maxpack, m->stream, direct ? 3 : 1);
  int rc = AB_OK;
  if (direct) {
    NK(g_nccl.AllReduce(m->p2p_token, m->p2p_token, 1, NCCL_FLOAT64, NCCL_MIN, m->comm, m->stream));
    ++m->p2p_round;
  } else {
    rc = peer_exchange(m, m->peer_state);
  }
  if (rc) return rc;
  ab::launch_copy_boxes(P.phase1, P.n1, P.max1, m->stream, 1);
  ab::launch_copy_boxes(P.phase1r, P.n1r, P.max1r, m->stream, 1);
  ab::launch_copy_boxes(P.phase2, P.n2, P.max2, m->stream, 1);
  CK(cudaGetLastError());
  return AB_OK;
}

// The two halves of bvals_exchange for the overlapped schedule: everything that only needs this
// rank's data (pack, start of the NCCL transfer on the comm stream, same-rank ghost copies) and
// everything that needs the peers' data.
int bvals_exchange_begin(AbMesh *m) {
  int idx = plan_index(m);
  int use = idx < 0 ? 0 : idx;
  if (idx < 0) m->plan[0].built = false;
  if (!m->plan[use].built) { int rc = build_state_plan(m, m->plan[use]); if (rc) return rc; }
  AbMesh::Plan &P = m->plan[use];
  if (P.npack) ab::launch_copy_boxes(P.pack, P.npack, P.maxpack, m->stream, 1);
  CK(cudaEventRecord(m->ev_pack, m->stream));
  CK(cudaStreamWaitEvent(m->comm_stream, m->ev_pack, 0));
  int rc = peer_exchange(m, m->peer_state, m->comm_stream);
  if (rc) return rc;
  CK(cudaEventRecord(m->ev_recv, m->comm_stream));
  ab::launch_copy_boxes(P.phase1, P.n1, P.max1, m->stream, 1);
  return AB_OK;
}
int bvals_exchange_end(AbMesh *m) {
  int idx = plan_index(m);
  AbMesh::Plan &P = m->plan[idx < 0 ? 0 : idx];
  CK(cudaStreamWaitEvent(m->stream, m->ev_recv, 0));
  ab::launch_copy_boxes(P.phase1r, P.n1r, P.max1r, m->stream, 1);
  ab::launch_copy_boxes(P.phase2, P.n2, P.max2, m->stream, 1);
  if (idx < 0) m->plan[0].built = false;   // mixed state: never cache
  CK(cudaGetLastError());
  return AB_OK;
}

// EMF pack / apply of every local block: one launch over all blocks inside a batched cycle
void emf_pack_all(AbMesh *m) {
  if (m->batch) ab::launch_emf_pack(m->lb[0].d, m->emf_plans_dev, m->stream, (int)m->lb.size());
  else for (size_t l = 0; l < m->lb.size(); ++l) ab::launch_emf_pack(m->lb[l].d, m->emf_plans_dev + l, m->stream);
}
void emf_apply_all(AbMesh *m) {
  if (m->batch) ab::launch_emf_apply(m->lb[0].d, m->emf_plans_dev, m->stream, (int)m->lb.size());
  else for (size_t l = 0; l < m->lb.size(); ++l) ab::launch_emf_apply(m->lb[l].d, m->emf_plans_dev + l, m->stream);
}

int emf_exchange(AbMesh *m) {
  if (!m->p.mhd) return AB_OK;
  if (!m->emf_built) { int rc = build_emf_plan(m); if (rc) return rc; }
  emf_pack_all(m);
  int rc = peer_exchange(m, m->peer_emf);
  if (rc) return rc;
  emf_apply_all(m);
  CK(cudaGetLastError());
  return AB_OK;
}

int emf_exchange_begin(AbMesh *m) {
  if (!m->p.mhd) return AB_OK;
  if (!m->emf_built) { int rc = build_emf_plan(m); if (rc) return rc; }
  emf_pack_all(m);
  CK(cudaEventRecord(m->ev_pack, m->stream));
  CK(cudaStreamWaitEvent(m->comm_stream, m->ev_pack, 0));
  int rc = peer_exchange(m, m->peer_emf, m->comm_stream);
  if (rc) return rc;
  CK(cudaEventRecord(m->ev_recv, m->comm_stream));
  return AB_OK;
}
int emf_exchange_end(AbMesh *m) {
  if (!m->p.mhd) return AB_OK;
  CK(cudaStreamWaitEvent(m->stream, m->ev_recv, 0));
  emf_apply_all(m);
  CK(cudaGetLastError());
  return AB_OK;
}

// ---------------------------------------------------------------------------------------------
// Boundary tasks of ONE MeshBlock, for a host scheduler that polls (task_list.cpp:66-91):
// Send never waits, ReceiveTry answers 0 ("not yet", TaskStatus::fail) until every neighbour of
// the block has sent for the same round, Set fills the block's ghost zones
// (bvals_var.cpp:212-296).  Same-GPU neighbours exchange nothing: Set gathers straight from the
// neighbour's active zones, which the neighbour's Send declared final for this stage.  Blocks on
// other ranks: Send packs into the peer buffer; the grouped NCCL send/recv of a round is issued
// by the last local Send of that round, and ReceiveTry of a block with remote neighbours waits
// for it.  Everything is enqueued on the mesh's stream in call order; nothing synchronises.
// ---------------------------------------------------------------------------------------------
int var_applies(const AbMesh *m, int var) {
  return var == 0 || (var == 1 && m->p.mhd) || (var == 2 && m->p.nscalars > 0);
}
// for_send: the plan is used for its pack list only (Send reads nothing but the block's own
// arrays, so the same-rank neighbours may still be a task behind, in the other swap state); the
// gather lists of such a plan carry the neighbours' pointers of that moment and are never used --
// Set builds its own plan, when every neighbour has sent
int block_plan(AbMesh *m, int lid, int var, AbMesh::Plan **out, bool for_send = false) {
  LocalBlock &L = m->lb[lid];
  const int par = (var == 0) ? L.parity : ((var == 1) ? L.parity_b : L.parity_s);
  // gathers read the neighbours' registers: their swap state is part of the key
  long key = (((long)lid*3 + var)*2 + par)*2 + (for_send ? 1 : 0);
  if (!for_send) for (const Nb &nb : L.hb->nbs) if (nb.rank == m->p.rank) {
    const LocalBlock &N = m->lb[owner_lid(m, nb.gid)];
    const int np = (var == 0) ? N.parity : ((var == 1) ? N.parity_b : N.parity_s);
    if (np != par) return fail(AB_ERR_STATE, "neighbouring MeshBlocks are in different register "
                                             "swap states (integrate all blocks of a stage first)");
  }
  AbMesh::Plan &P = m->bplan[key];
  if (!P.built) { int rc = build_state_plan(m, P, lid, 1 << var); if (rc) return rc; }
  *out = &P;
  return AB_OK;
}
bool has_remote_nb(const AbMesh *m, const LocalBlock &L, int max_type = 2) {
  for (const Nb &nb : L.hb->nbs) if (nb.rank != m->p.rank && nb.type <= max_type) return true;
  return false;
}
int block_bvals_send(AbMesh *m, int lid, int var) {
  if (m->bcomm.size() != m->lb.size()) m->bcomm.assign(m->lb.size(), AbMesh::BlockComm());
  LocalBlock &L = m->lb[lid];
  if (has_remote_nb(m, L)) {
    AbMesh::Plan *P;
    int rc = block_plan(m, lid, var, &P, true);
    if (rc) return rc;
    if (P->npack) ab::launch_copy_boxes(P->pack, P->npack, P->maxpack, m->stream, 1);
  }
  m->bcomm[lid].sent[var]++;
  if (!m->peer_state.empty() || m->p.nranks > 1) {
    long lo = m->bcomm[0].sent[0];
    for (auto &c : m->bcomm) for (int v = 0; v < 3; ++v) if (var_applies(m, v)) lo = std::min(lo, c.sent[v]);
    if (lo > m->nccl_state_round) {          // every block has sent every variable of this round
      int rc = peer_exchange(m, m->peer_state);
      if (rc) return rc;
      m->nccl_state_round = lo;
    }
  }
  CK(cudaGetLastError());
  return AB_OK;
}
int block_bvals_recv_try(AbMesh *m, int lid, int var) {
  if (m->bcomm.size() != m->lb.size()) m->bcomm.assign(m->lb.size(), AbMesh::BlockComm());
  const long need = m->bcomm[lid].recvd[var] + 1;
  for (const Nb &nb : m->lb[lid].hb->nbs) {
    if (nb.rank == m->p.rank) { if (m->bcomm[owner_lid(m, nb.gid)].sent[var] < need) return 0; }
    else if (m->nccl_state_round < need) return 0;
  }
  m->bcomm[lid].recvd[var] = need;
  return 1;
}
int block_bvals_set(AbMesh *m, int lid, int var) {
  AbMesh::Plan *P;
  int rc = block_plan(m, lid, var, &P);
  if (rc) return rc;
  ab::launch_copy_boxes(P->phase1, P->n1, P->max1, m->stream, 1);
  ab::launch_copy_boxes(P->phase1r, P->n1r, P->max1r, m->stream, 1);
  ab::launch_copy_boxes(P->phase2, P->n2, P->max2, m->stream, 1);
  CK(cudaGetLastError());
  return AB_OK;
}
int block_emf_send(AbMesh *m, int lid) {
  if (m->bcomm.size() != m->lb.size()) m->bcomm.assign(m->lb.size(), AbMesh::BlockComm());
  if (!m->emf_built) { int rc = build_emf_plan(m); if (rc) return rc; }
  LocalBlock &L = m->lb[lid];
  ab::launch_emf_pack(L.d, m->emf_plans_dev + lid, m->stream);
  m->bcomm[lid].emf_sent++;
  if (!m->peer_emf.empty()) {
    long lo = m->bcomm[0].emf_sent;
    for (auto &c : m->bcomm) lo = std::min(lo, c.emf_sent);
    if (lo > m->nccl_emf_round) {
      int rc = peer_exchange(m, m->peer_emf);
      if (rc) return rc;
      m->nccl_emf_round = lo;
    }
  }
  CK(cudaGetLastError());
  return AB_OK;
}
int block_emf_recv_try(AbMesh *m, int lid) {
  if (m->bcomm.size() != m->lb.size()) m->bcomm.assign(m->lb.size(), AbMesh::BlockComm());
  LocalBlock &L = m->lb[lid];
  const long need = m->bcomm[lid].emf_recvd + 1;
  if (m->bcomm[lid].emf_sent < need) return 0;      // own buffers feed the averages too
  for (const Nb &nb : L.hb->nbs) {
    if (nb.type > 1) continue;                      // faces and edges only
    if (nb.rank == m->p.rank) { if (m->bcomm[owner_lid(m, nb.gid)].emf_sent < need) return 0; }
    else if (m->nccl_emf_round < need) return 0;
  }
  ab::launch_emf_apply(L.d, m->emf_plans_dev + lid, m->stream);
  m->bcomm[lid].emf_recvd = need;
  CK(cudaGetLastError());
  return 1;
}

// Primitives task split for the overlapped schedule: part 0 = the active cells (with the fused
// CFL reduction), part 1 = the ghost shell as up to six slabs in one launch (after the ghost zones
// arrived).  nb > 1: the same for nb blocks starting at L in one launch (same ranges, same flags).
void primitives_part(AbMesh *m, LocalBlock &L, int part, int with_dt, int nb = 1) {
  HostBlock &B = *L.hb;
  int ng = m->p.nghost;
  const int is = m->is, ie = m->ie, js = m->js, je = m->je, ks = m->ks, ke = m->ke;
  int il = is, iu = ie, jl = js, ju = je, kl = ks, ku = ke;
  if (B.nblevel[1][1][0] != -1) il -= ng;
  if (B.nblevel[1][1][2] != -1) iu += ng;
  if (B.nblevel[1][0][1] != -1) jl -= ng;
  if (B.nblevel[1][2][1] != -1) ju += ng;
  if (B.nblevel[0][1][1] != -1) kl -= ng;
  if (B.nblevel[2][1][1] != -1) ku += ng;
  const int cce = (m->p.mhd && !L.has_phys_bc) ? 1 : 0;
  if (part == 0) {
    ab::launch_cons2prim(L.d, m->kp, is, ie, js, je, ks, ke, m->stream, cce | (with_dt ? 2 : 0),
                         L.dtmin, nb);
    ab::launch_scalar_eos(L.d, m->kp, 0, is, ie, js, je, ks, ke, m->stream, nb);
  } else {
    const int box[6][6] = {{il, iu, jl, ju, kl, ks-1}, {il, iu, jl, ju, ke+1, ku},
                           {il, iu, jl, js-1, ks, ke}, {il, iu, je+1, ju, ks, ke},
                           {il, is-1, js, je, ks, ke}, {ie+1, iu, js, je, ks, ke}};
    ab::launch_cons2prim_boxes(L.d, m->kp, 6, box, m->stream, cce, nb);
    if (m->p.nscalars > 0) for (int q = 0; q < 6; ++q)
      if (box[q][0] <= box[q][1] && box[q][2] <= box[q][3] && box[q][4] <= box[q][5])
        ab::launch_scalar_eos(L.d, m->kp, 0, box[q][0], box[q][1], box[q][2], box[q][3], box[q][4],
                              box[q][5], m->stream, nb);
    for (int l = 0; l < nb; ++l) (&L)[l].cc_e_valid = cce != 0;
  }
}

// true when every local block converts the same cell range with the same flags (always on a
// periodic mesh): the Primitives task can then run over all of them in one launch
bool primitives_uniform(const AbMesh *m) {
  for (size_t l = 1; l < m->lb.size(); ++l) {
    const HostBlock &A = *m->lb[0].hb, &B = *m->lb[l].hb;
    if (m->lb[0].has_phys_bc != m->lb[l].has_phys_bc) return false;
    for (int d = 0; d < 3; ++d) for (int sgn = 0; sgn < 3; sgn += 2) {
      const int a = (d == 0) ? A.nblevel[1][1][sgn] : ((d == 1) ? A.nblevel[1][sgn][1] : A.nblevel[sgn][1][1]);
      const int b = (d == 0) ? B.nblevel[1][1][sgn] : ((d == 1) ? B.nblevel[1][sgn][1] : B.nblevel[sgn][1][1]);
      if ((a != -1) != (b != -1)) return false;
    }
  }
  return true;
}

void primitives(AbMesh *m, LocalBlock &L, int with_dt = 0) {
  // TimeIntegratorTaskList::Primitives (time_integrator.cpp:1965-1983)
  HostBlock &B = *L.hb;
  int ng = m->p.nghost;
  int il = m->is, iu = m->ie, jl = m->js, ju = m->je, kl = m->ks, ku = m->ke;
  if (B.nblevel[1][1][0] != -1) il -= ng;
  if (B.nblevel[1][1][2] != -1) iu += ng;
  if (B.nblevel[1][0][1] != -1) jl -= ng;
  if (B.nblevel[1][2][1] != -1) ju += ng;
  if (B.nblevel[0][1][1] != -1) kl -= ng;
  if (B.nblevel[2][1][1] != -1) ku += ng;
  // blocks without physical boundaries get cc_e from the same pass (all cells that
  // ComputeCornerE reads, [is-1,ie+1]^dim, lie inside the cons2prim range)
  int flags = (m->p.mhd && !L.has_phys_bc ? 1 : 0) | (with_dt ? 2 : 0);
  ab::launch_cons2prim(L.d, m->kp, il, iu, jl, ju, kl, ku, L.stream, flags, L.dtmin);
  ab::launch_scalar_eos(L.d, m->kp, 0, il, iu, jl, ju, kl, ku, L.stream);
  L.cc_e_valid = (flags & 1) != 0;
}

// Primitives of every local block in one launch when they all convert the same cell range with
// the same flags (always on a periodic mesh); else block by block.
void primitives_all(AbMesh *m, int with_dt) {
  const int ng = m->p.nghost;
  int r0[6] = {0, 0, 0, 0, 0, 0}, f0 = 0;
  bool same = true;
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    HostBlock &B = *L.hb;
    int r[6] = {m->is, m->ie, m->js, m->je, m->ks, m->ke};
    if (B.nblevel[1][1][0] != -1) r[0] -= ng;
    if (B.nblevel[1][1][2] != -1) r[1] += ng;
    if (B.nblevel[1][0][1] != -1) r[2] -= ng;
    if (B.nblevel[1][2][1] != -1) r[3] += ng;
    if (B.nblevel[0][1][1] != -1) r[4] -= ng;
    if (B.nblevel[2][1][1] != -1) r[5] += ng;
    const int flags = (m->p.mhd && !L.has_phys_bc ? 1 : 0) | (with_dt ? 2 : 0);
    if (l == 0) { for (int q = 0; q < 6; ++q) r0[q] = r[q]; f0 = flags; }
    else { for (int q = 0; q < 6; ++q) same = same && (r[q] == r0[q]); same = same && (flags == f0); }
  }
  if (!same) {
    for (auto &L : m->lb) primitives(m, L, with_dt);
    return;
  }
  const int nb = (int)m->lb.size();
  LocalBlock &L0 = m->lb[0];
  ab::launch_cons2prim(L0.d, m->kp, r0[0], r0[1], r0[2], r0[3], r0[4], r0[5], m->stream, f0,
                       L0.dtmin, nb);
  ab::launch_scalar_eos(L0.d, m->kp, 0, r0[0], r0[1], r0[2], r0[3], r0[4], r0[5], m->stream, nb);
  for (auto &L : m->lb) L.cc_e_valid = (f0 & 1) != 0;
}

// DispatchBoundaryFunctions for a user-enrolled face (bvals.cpp:617-620): host round trip of
// w (and b) of this block around the callback.
int user_bc_face(AbMesh *m, LocalBlock &L, int face, int il, int iu, int jl, int ju, int kl,
                 int ku) {
  if (!m->user_bc[face]) return fail(AB_ERR_STATE, "user boundary function not enrolled");
  cudaStream_t s = m->stream;
  m->stage_w.resize(L.regsize[AB_W]);
  CK(cudaMemcpyAsync(m->stage_w.data(), L.d.w, L.regsize[AB_W]*8, cudaMemcpyDeviceToHost, s));
  double *bp[3] = {nullptr, nullptr, nullptr};
  if (m->p.mhd) for (int c = 0; c < 3; ++c) {
    m->stage_b[c].resize(L.regsize[AB_B_X1F + c]);
    bp[c] = m->stage_b[c].data();
    CK(cudaMemcpyAsync(bp[c], L.d.b[c], L.regsize[AB_B_X1F + c]*8, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  int lid = (int)(&L - &m->lb[0]);
  m->user_bc[face](m->user_bc_arg[face], lid, m->stage_w.data(), bp[0], bp[1], bp[2],
                   m->bc_time, m->bc_dt, il, iu, jl, ju, kl, ku, m->p.nghost);
  CK(cudaMemcpyAsync(L.d.w, m->stage_w.data(), L.regsize[AB_W]*8, cudaMemcpyHostToDevice, s));
  if (m->p.mhd) for (int c = 0; c < 3; ++c)
    CK(cudaMemcpyAsync(L.d.b[c], bp[c], L.regsize[AB_B_X1F + c]*8, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));    // the staging vectors are reused by the next face
  L.cc_e_valid = false;
  return AB_OK;
}

// HydroSourceTerms::UserSourceTerm (hydro_srcterms.cpp:150-153) for one block
int user_source(AbMesh *m, LocalBlock &L, double time, double dt) {
  int lid = (int)(&L - &m->lb[0]);
  cudaStream_t s = m->stream;
  if (m->user_src_dev) {
    m->user_src_dev(m->user_src_arg, lid, time, dt, L.d.w, L.d.r, L.d.bcc, L.d.u, L.d.s, (void *)s);
    return AB_OK;
  }
  auto down = [&](std::vector<double> &v, const double *src, long n) -> double * {
    if (!src || n <= 0) return nullptr;
    v.resize(n);
    cudaMemcpyAsync(v.data(), src, n*8, cudaMemcpyDeviceToHost, s);
    return v.data();
  };
  double *hw = down(m->stage_w, L.d.w, L.regsize[AB_W]);
  double *hr = down(m->stage_r, L.d.r, L.regsize[AB_R]);
  double *hb = down(m->stage_bcc, L.d.bcc, L.regsize[AB_BCC]);
  double *hu = down(m->stage_u, L.d.u, L.regsize[AB_U]);
  double *hs = down(m->stage_s, L.d.s, L.regsize[AB_S]);
  CK(cudaStreamSynchronize(s));
  m->user_src(m->user_src_arg, lid, time, dt, hw, hr, hb, hu, hs);
  CK(cudaMemcpyAsync(L.d.u, hu, L.regsize[AB_U]*8, cudaMemcpyHostToDevice, s));
  if (hs) CK(cudaMemcpyAsync(L.d.s, hs, L.regsize[AB_S]*8, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  return AB_OK;
}

int physical_bcs(AbMesh *m, LocalBlock &L) {
  // BoundaryValues::ApplyPhysicalBoundaries (bvals/bvals.cpp:436-620): outflow, reflecting
  HostBlock &B = *L.hb;
  int ng = m->p.nghost, mhd = m->p.mhd;
  int is = m->is, ie = m->ie, js = m->js, je = m->je, ks = m->ks, ke = m->ke;
  int bis = is - ng, bie = ie + ng, bjs = js, bje = je, bks = ks, bke = ke;
  bool app[6];
  for (int f = 0; f < 6; ++f) app[f] = (B.bcs[f] != -1 && B.bcs[f] != AB_BC_PERIODIC);
  if (!m->f2) app[2] = app[3] = false;
  if (!m->f3) app[4] = app[5] = false;
  if (!app[2] && m->f2) bjs = js - ng;
  if (!app[3] && m->f2) bje = je + ng;
  if (!app[4] && m->f3) bks = ks - ng;
  if (!app[5] && m->f3) bke = ke + ng;
  cudaStream_t s = L.stream;
  if (app[0]) {
    if (B.bcs[0] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 0, is, ie, bjs, bje, bks, bke); if (rcu) return rcu; }

    else ab::launch_phys_bc(L.d, mhd, 0, B.bcs[0] == AB_BC_REFLECT, is, ie, bjs, bje, bks, bke, s);
    if (mhd) ab::launch_calc_bcc(L.d, is-ng, is-1, bjs, bje, bks, bke, s);
    ab::launch_prim2cons(L.d, m->kp, is-ng, is-1, bjs, bje, bks, bke, s);
    ab::launch_scalar_eos(L.d, m->kp, 1, is-ng, is-1, bjs, bje, bks, bke, s);
  }
  if (app[1]) {
    if (B.bcs[1] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 1, is, ie, bjs, bje, bks, bke); if (rcu) return rcu; }

    else ab::launch_phys_bc(L.d, mhd, 1, B.bcs[1] == AB_BC_REFLECT, is, ie, bjs, bje, bks, bke, s);
    if (mhd) ab::launch_calc_bcc(L.d, ie+1, ie+ng, bjs, bje, bks, bke, s);
    ab::launch_prim2cons(L.d, m->kp, ie+1, ie+ng, bjs, bje, bks, bke, s);
    ab::launch_scalar_eos(L.d, m->kp, 1, ie+1, ie+ng, bjs, bje, bks, bke, s);
  }
  if (m->f2) {
    if (app[2]) {
      if (B.bcs[2] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 2, bis, bie, js, je, bks, bke); if (rcu) return rcu; }

      else ab::launch_phys_bc(L.d, mhd, 2, B.bcs[2] == AB_BC_REFLECT, bis, bie, js, je, bks, bke, s);
      if (mhd) ab::launch_calc_bcc(L.d, bis, bie, js-ng, js-1, bks, bke, s);
      ab::launch_prim2cons(L.d, m->kp, bis, bie, js-ng, js-1, bks, bke, s);
      ab::launch_scalar_eos(L.d, m->kp, 1, bis, bie, js-ng, js-1, bks, bke, s);
    }
    if (app[3]) {
      if (B.bcs[3] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 3, bis, bie, js, je, bks, bke); if (rcu) return rcu; }

      else ab::launch_phys_bc(L.d, mhd, 3, B.bcs[3] == AB_BC_REFLECT, bis, bie, js, je, bks, bke, s);
      if (mhd) ab::launch_calc_bcc(L.d, bis, bie, je+1, je+ng, bks, bke, s);
      ab::launch_prim2cons(L.d, m->kp, bis, bie, je+1, je+ng, bks, bke, s);
      ab::launch_scalar_eos(L.d, m->kp, 1, bis, bie, je+1, je+ng, bks, bke, s);
    }
  }
  if (m->f3) {
    bjs = js - ng; bje = je + ng;
    if (app[4]) {
      if (B.bcs[4] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 4, bis, bie, bjs, bje, ks, ke); if (rcu) return rcu; }

      else ab::launch_phys_bc(L.d, mhd, 4, B.bcs[4] == AB_BC_REFLECT, bis, bie, bjs, bje, ks, ke, s);
      if (mhd) ab::launch_calc_bcc(L.d, bis, bie, bjs, bje, ks-ng, ks-1, s);
      ab::launch_prim2cons(L.d, m->kp, bis, bie, bjs, bje, ks-ng, ks-1, s);
      ab::launch_scalar_eos(L.d, m->kp, 1, bis, bie, bjs, bje, ks-ng, ks-1, s);
    }
    if (app[5]) {
      if (B.bcs[5] == AB_BC_USER) { const int rcu = user_bc_face(m, L, 5, bis, bie, bjs, bje, ks, ke); if (rcu) return rcu; }

      else ab::launch_phys_bc(L.d, mhd, 5, B.bcs[5] == AB_BC_REFLECT, bis, bie, bjs, bje, ks, ke, s);
      if (mhd) ab::launch_calc_bcc(L.d, bis, bie, bjs, bje, ke+1, ke+ng, s);
      ab::launch_prim2cons(L.d, m->kp, bis, bie, bjs, bje, ke+1, ke+ng, s);
      ab::launch_scalar_eos(L.d, m->kp, 1, bis, bie, bjs, bje, ke+1, ke+ng, s);
    }
  }
  return AB_OK;
}

int new_time_step(AbMesh *m, int advance, bool blocks_done = false) {
  // NewBlockTimeStep on every block (unless already reduced by the fused cons2prim), then
  // Mesh::NewTimeStep (mesh/mesh.cpp:1078-1119)
  if (!blocks_done) {
    ab::launch_fill_u64(m->dtmin, (int)m->lb.size()*ab::DT_SLOTS, 0x7FEFFFFFFFFFFFFFull /* DBL_MAX */, m->stream);
    for (auto &L : m->lb) ab::launch_new_block_dt(L.d, m->kp, L.dtmin, m->stream);
  }
  ab::launch_mesh_new_dt(m->state, m->dtmin, (int)m->lb.size(), 0 /*phase 0*/, m->stream);
  if (m->p.nranks > 1) {
    if (!m->comm) return fail(AB_ERR_STATE, "nranks > 1 but ab_comm_init was not called");
    NK(g_nccl.AllReduce(m->state + 4, m->state + 4, 1, NCCL_FLOAT64, NCCL_MIN, m->comm, m->stream));
  }
  ab::launch_mesh_new_dt(m->state, m->dtmin, (int)m->lb.size(), 2 | (advance ? 1 : 0), m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}

int read_state(AbMesh *m) {
  double h[6];
  CK(cudaMemcpyAsync(h, m->state, sizeof(h), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  m->h_time = h[0]; m->h_dt = h[1]; m->h_ncycle = (long)h[5];
  return AB_OK;
}

void swap_cc(LocalBlock &L) { std::swap(L.d.u, L.d.u1); L.parity ^= 1; }
void swap_fc(LocalBlock &L) {
  for (int d = 0; d < 3; ++d) std::swap(L.d.b[d], L.d.b1[d]);
  L.parity_b ^= 1;
}

void swap_sc(LocalBlock &L) { std::swap(L.d.s, L.d.s1); L.parity_s ^= 1; }

// AB_DEBUG_SYNC=1: synchronise after every task group of the cycle and name the group that
// faulted (debug aid; the production path never synchronises inside a cycle)
bool debug_sync() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("AB_DEBUG_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
#define DBG(m, label)                                                               \
  do {                                                                              \
    if (debug_sync()) {                                                             \
      cudaError_t e_ = cudaStreamSynchronize((m)->stream);                          \
      if (e_ != cudaSuccess)                                                        \
        return fail(AB_ERR_CUDA, std::string("after ") + (label) + ": " + cudaGetErrorString(e_)); \
    }                                                                               \
  } while (0)

// ------------------------------------------------------------------ static mesh refinement
// Execution of the host planner's rows on the device.  The interpretation of the rows is the
// back-end independent code of ab_smr_exec.h (also run on the CPU by tests/hostcheck against the
// oracle); this back end turns each piece of work into a kernel launch (ab_smr_kernels.cu).
struct SmrDeviceOps {
  AbMesh *m;
  cudaStream_t st;
  int rc = AB_OK;
  void restrict_box(const ab::SmrGeom &g, const double *fine, double *coarse, int nvar,
                    const ab::SmrBox &bx) { ab::launch_smr_restrict(g, fine, coarse, nvar, bx, st); }
  void prolong_box(const ab::SmrGeom &g, const double *coarse, double *fine, int nvar,
                   const ab::SmrBox &bx) { ab::launch_smr_prolong(g, coarse, fine, nvar, bx, st); }
  void c2p_box(const ab::SmrGeom &g, double *cu, double *cw, int ns, double *cs, double *cr,
               const ab::SmrBox &bx) { ab::launch_smr_c2p(g, m->kp, cu, cw, ns, cs, cr, bx, st); }
  void bc_box(const ab::SmrGeom &g, double *cw, int nh, double *cr, int ns, int face, bool refl,
              int lo, int hi, const ab::SmrBox &bx) {
    ab::launch_smr_bc(g, cw, nh, cr, ns, face, refl ? 1 : 0, lo, hi, bx, st);
  }
  void prim2cons_box(int gid, int il, int iu, int jl, int ju, int kl, int ku) {
    LocalBlock &L = m->lb[gid - m->gid_start];
    ab::launch_prim2cons(L.d, m->kp, il, iu, jl, ju, kl, ku, st);
    ab::launch_scalar_eos(L.d, m->kp, 1, il, iu, jl, ju, kl, ku, st);
  }
  void flux_face(const ab::SmrGeom &g, const double *ff, double *cf, int nvar, int dir, int fpos,
                 int cpos, int a0, int b0, int na, int nb) {
    ab::launch_smr_flux(g, ff, cf, nvar, dir, fpos, cpos, a0, b0, na, nb, st);
  }
  // room for n box descriptors in the device table.  Across ranks this must happen BEFORE the
  // round's NCCL operation is enqueued: cudaFree / cudaMalloc synchronise the device, and blocking
  // in them while a send/recv kernel waits for its peer can deadlock with NCCL's own progress.
  void reserve_boxes(size_t n) {
    if (n <= m->smr_boxes_cap) return;
    if (m->smr_boxes) cudaFree(m->smr_boxes);
    m->smr_boxes = nullptr;
    m->smr_boxes_cap = n;
    if (cudaMalloc(&m->smr_boxes, sizeof(ab::CopyBox)*m->smr_boxes_cap) != cudaSuccess) rc = AB_ERR_CUDA;
  }
  void copy_boxes(std::vector<ab::CopyBox> &v, long total) {
    if (v.empty()) return;
    reserve_boxes(v.size());
    if (rc) return;
    // the table carries the CURRENT register pointers (u / u1 swap every stage); stream-ordered
    // upload + sync because the source is pageable host memory (see build_state_plan)
    for (auto &c : v) ab::set_box_divisors(c);
    if (cudaMemcpyAsync(m->smr_boxes, v.data(), sizeof(ab::CopyBox)*v.size(), cudaMemcpyHostToDevice, st) != cudaSuccess
        || cudaStreamSynchronize(st) != cudaSuccess) { rc = AB_ERR_CUDA; return; }
    ab::launch_copy_boxes(m->smr_boxes, (int)v.size(), total, st);
  }
};

ab::SmrDims smr_dims(const AbMesh *m) {
  ab::SmrDims d;
  d.nh = m->nh; d.ns = m->p.nscalars; d.ng = m->p.nghost;
  d.bx[0] = m->p.bx1; d.bx[1] = m->p.bx2; d.bx[2] = m->p.bx3;
  d.s0[0] = m->is; d.s0[1] = m->js; d.s0[2] = m->ks;
  d.e0[0] = m->ie; d.e0[1] = m->je; d.e0[2] = m->ke;
  d.fdim[0] = true; d.fdim[1] = m->f2; d.fdim[2] = m->f3;
  return d;
}

// views with the registers' CURRENT pointers (u <-> u1 swap by pointer every stage), one per
// MeshBlock of the MESH in gid order (the planner's rows name blocks by gid); the blocks of other
// ranks carry the index geometry only (all blocks have the same shape), their pointers are null
std::vector<ab::SmrView> smr_views(AbMesh *m) {
  std::vector<ab::SmrView> v(m->hb.size());
  for (size_t g = 0; g < m->hb.size(); ++g) {
    ab::SmrView &x = v[g];
    const HostBlock &B = m->hb[g];
    for (int f = 0; f < 6; ++f) x.bcs[f] = B.bcs[f];
    x.g = m->smr_blk[0].g;                       // nc*, cnc*, is.., cis..: identical on every block
    x.g.dx1f = x.g.dx2f = x.g.dx3f = x.g.x1v = x.g.x2v = x.g.x3v = nullptr;
    x.g.cx1v = x.g.cx2v = x.g.cx3v = nullptr;
    if (B.rank != m->p.rank) continue;
    const int l = (int)g - m->gid_start;
    const LocalBlock &L = m->lb[l];
    const AbMesh::SmrBlk &sb = m->smr_blk[l];
    x.u = L.d.u; x.s = L.d.s; x.w = L.d.w; x.r = L.d.r;
    for (int d = 0; d < 3; ++d) { x.flux[d] = L.d.flux[d]; x.sflux[d] = L.d.sflux[d]; }
    x.cu = sb.cu; x.cw = sb.cw; x.cs = sb.cs; x.cr = sb.cr;
    x.g = sb.g;
  }
  return v;
}

// message buffers of one exchange kind sized for this round (grown on demand)
int smr_peer_buffers(std::map<int, PeerBuf> &peers, const std::map<int, long> &nsend,
                     const std::map<int, long> &nrecv) {
  for (auto &kv : nsend) {
    PeerBuf &pb = peers[kv.first];
    if ((size_t)kv.second > pb.nsend || !pb.send) {
      cudaFree(pb.send);
      CK(cudaMalloc(&pb.send, std::max<long>(kv.second, 1)*8));
    }
    pb.nsend = (size_t)kv.second;
  }
  for (auto &kv : nrecv) {
    PeerBuf &pb = peers[kv.first];
    if ((size_t)kv.second > pb.nrecv || !pb.recv) {
      cudaFree(pb.recv);
      CK(cudaMalloc(&pb.recv, std::max<long>(kv.second, 1)*8));
    }
    pb.nrecv = (size_t)kv.second;
  }
  return AB_OK;
}

int smr_exchange(AbMesh *m) {
  // disjoint sources / destinations within the one copy launch: ab_mesh_create_refined requires
  // MeshBlocks of at least 2*NGHOST cells, so restricted slabs stay inside the active coarse cells
  SmrDeviceOps ops{m, m->stream};
  std::vector<ab::SmrView> v = smr_views(m);
  const ab::SmrDims d = smr_dims(m);
  if (m->p.nranks == 1) {
    ab::smr_run_exchange(m->smr_rows, v, d, ops);
    if (ops.rc) return fail(ops.rc, "SMR exchange: CUDA error");
    CK(cudaGetLastError());
    return AB_OK;
  }
  // Across ranks (bvals_cc.cpp:195-470 with MPI): every rank walks the SAME row list, so the
  // messages of a pair of ranks are in the same order on both sides: the sender packs the box of
  // a row (after restricting it, kind 2) into its buffer for the receiver's rank, one grouped
  // send/recv per peer moves the buffers, the receiver unpacks in the same order.
  const int me = m->p.rank;
  const int npass = d.ns > 0 ? 2 : 1;
  std::map<int, long> nsend, nrecv;
  for (const auto &r : m->smr_rows) {
    if (r[0] < 0 || r[0] > 2) continue;
    const int rs = m->hb[r[1]].rank, rt = m->hb[r[5]].rank;
    if (r[0] == 2 && rs == me) {
      ab::SmrView &S = v[r[1]];
      const ab::SmrBox bx = ab::smr_box(&r[2], &r[9]);
      ops.restrict_box(S.g, S.u, S.cu, d.nh, bx);
      if (d.ns > 0) ops.restrict_box(S.g, S.s, S.cs, d.ns, bx);
    }
    if (rs == rt) continue;
    const long cnt = (long)r[9]*r[10]*r[11]*(d.nh + d.ns);
    if (rs == me) nsend[rt] += cnt;
    if (rt == me) nrecv[rs] += cnt;
  }
  { int rcb = smr_peer_buffers(m->peer_smr, nsend, nrecv); if (rcb) return rcb; }
  std::vector<ab::CopyBox> local, pack, unpack;
  std::map<int, long> soff, roff;
  for (const auto &r : m->smr_rows) {
    if (r[0] < 0 || r[0] > 2) continue;
    const int rs = m->hb[r[1]].rank, rt = m->hb[r[5]].rank;
    if (rs != me && rt != me) continue;
    ab::SmrView &S = v[r[1]], &T = v[r[5]];
    for (int pass = 0; pass < npass; ++pass) {
      const long cnt = (long)r[9]*r[10]*r[11]*(pass ? d.ns : d.nh);
      if (rs == me && rt == me) {
        local.push_back(ab::smr_row_copy(r, S, T, d, pass));
      } else if (rs == me) {
        pack.push_back(ab::smr_row_copy_buffered(r, S, T, d, pass, m->peer_smr[rt].send + soff[rt], true));
        soff[rt] += cnt;
      } else {
        unpack.push_back(ab::smr_row_copy_buffered(r, S, T, d, pass, m->peer_smr[rs].recv + roff[rs], false));
        roff[rs] += cnt;
      }
    }
  }
  auto launch = [&](std::vector<ab::CopyBox> &b) {
    long total = 0;
    for (auto &c : b) { c.offset = total; total += (long)c.ni*c.nj*c.nk*c.nvar; }
    ops.copy_boxes(b, total);
  };
  ops.reserve_boxes(std::max(pack.size(), local.size() + unpack.size()));   // no allocation once NCCL is in flight
  launch(pack);
  if (ops.rc) return fail(ops.rc, "SMR exchange: CUDA error");
  { int rcx = peer_exchange(m, m->peer_smr); if (rcx) return rcx; }
  local.insert(local.end(), unpack.begin(), unpack.end());
  launch(local);
  if (ops.rc) return fail(ops.rc, "SMR exchange: CUDA error");
  CK(cudaGetLastError());
  return AB_OK;
}

int smr_prolongate(AbMesh *m, int lid) {
  SmrDeviceOps ops{m, m->lb[lid].stream};
  std::vector<ab::SmrView> v = smr_views(m);
  const int gid = m->lb[lid].hb->gid;
  ab::smr_run_prolongate(m->smr_rows, gid, v[gid], smr_dims(m), ops);
  if (ops.rc) return fail(ops.rc, "SMR prolongation: CUDA error");
  CK(cudaGetLastError());
  return AB_OK;
}

int smr_flux_correction(AbMesh *m) {
  SmrDeviceOps ops{m, m->stream};
  std::vector<ab::SmrView> v = smr_views(m);
  const ab::SmrDims d = smr_dims(m);
  if (m->p.nranks == 1) {
    ab::smr_run_flux_correction(m->smr_rows, v, d, ops);
    if (ops.rc) return fail(ops.rc, "SMR flux correction: CUDA error");
    CK(cudaGetLastError());
    return AB_OK;
  }
  // Across ranks (flux_correction_cc.cpp:69-290 with MPI): the fine block's rank forms the
  // area-weighted coarse fluxes straight into its message buffer, the coarse block's rank copies
  // them onto its face; same row order on both sides.
  const int me = m->p.rank;
  struct Geo { int dir, fpos, cpos, a0, b0, na, nb; };
  auto geo = [&](const ab::SmrRow &r) {
    Geo q;
    const int ffid = (int)r[2], cfid = (int)r[6], fi1 = (int)r[7], fi2 = (int)r[8];
    q.dir = ffid >> 1;
    q.fpos = d.s0[q.dir] + (d.e0[q.dir] - d.s0[q.dir] + 1)*(ffid & 1);
    q.cpos = d.s0[q.dir] + (d.e0[q.dir] - d.s0[q.dir] + 1)*(cfid & 1);
    const int da = q.dir == 0 ? 1 : 0, db = q.dir == 2 ? 1 : 2;
    const int ha = d.fdim[da] ? d.bx[da]/2 : 0, hb = d.fdim[db] ? d.bx[db]/2 : 0;
    q.a0 = d.s0[da] + (fi1 ? ha : 0); q.b0 = d.s0[db] + (fi2 ? hb : 0);
    q.na = d.fdim[da] ? ha : 1; q.nb = d.fdim[db] ? hb : 1;
    return q;
  };
  std::map<int, long> nsend, nrecv;
  for (const auto &r : m->smr_rows) {
    if (r[0] != 20) continue;
    const int rf = m->hb[r[1]].rank, rc_ = m->hb[r[5]].rank;
    if (rf == rc_) continue;
    const Geo q = geo(r);
    const long cnt = (long)q.na*q.nb*(d.nh + d.ns);
    if (rf == me) nsend[rc_] += cnt;
    if (rc_ == me) nrecv[rf] += cnt;
  }
  { int rcb = smr_peer_buffers(m->peer_smr_flux, nsend, nrecv); if (rcb) return rcb; }
  std::vector<ab::CopyBox> unpack;
  std::map<int, long> soff, roff;
  const int n1 = m->nc[0], n2 = m->nc[1], n3 = m->nc[2];
  for (const auto &r : m->smr_rows) {
    if (r[0] != 20) continue;
    const int rf = m->hb[r[1]].rank, rc_ = m->hb[r[5]].rank;
    if (rf != me && rc_ != me) continue;
    ab::SmrView &F = v[r[1]], &Cb = v[r[5]];
    const Geo q = geo(r);
    for (int pass = 0; pass < (d.ns > 0 ? 2 : 1); ++pass) {
      const int nvar = pass ? d.ns : d.nh;
      const long cnt = (long)q.na*q.nb*nvar;
      if (rf == me && rc_ == me) {
        ops.flux_face(F.g, pass ? F.sflux[q.dir] : F.flux[q.dir], pass ? Cb.sflux[q.dir] : Cb.flux[q.dir],
                      nvar, q.dir, q.fpos, q.cpos, q.a0, q.b0, q.na, q.nb);
      } else if (rf == me) {
        ab::launch_smr_flux(F.g, pass ? F.sflux[q.dir] : F.flux[q.dir], m->peer_smr_flux[rc_].send + soff[rc_],
                            nvar, q.dir, q.fpos, q.cpos, q.a0, q.b0, q.na, q.nb, m->stream, 1);
        soff[rc_] += cnt;
      } else {
        // buffer [nvar][nb][na] -> the coarse block's face array of direction dir at (cpos, a0.., b0..)
        ab::CopyBox c;
        memset(&c, 0, sizeof(c));
        c.src = m->peer_smr_flux[rf].recv + roff[rf];
        c.dst = pass ? Cb.sflux[q.dir] : Cb.flux[q.dir];
        c.nvar = nvar;
        c.src_sv = (long)q.na*q.nb;
        if (q.dir == 0) {          // x1 faces: (a, b) = (j, k)
          c.dst_s2 = n1 + 1; c.dst_s3 = (long)n2*(n1 + 1); c.dst_sv = (long)n3*n2*(n1 + 1);
          c.di0 = q.cpos; c.dj0 = q.a0; c.dk0 = q.b0; c.ni = 1; c.nj = q.na; c.nk = q.nb;
          c.src_s2 = 1; c.src_s3 = q.na;
        } else if (q.dir == 1) {   // x2 faces: (a, b) = (i, k)
          c.dst_s2 = n1; c.dst_s3 = (long)(n2 + 1)*n1; c.dst_sv = (long)n3*(n2 + 1)*n1;
          c.di0 = q.a0; c.dj0 = q.cpos; c.dk0 = q.b0; c.ni = q.na; c.nj = 1; c.nk = q.nb;
          c.src_s2 = q.na; c.src_s3 = q.na;
        } else {                   // x3 faces: (a, b) = (i, j)
          c.dst_s2 = n1; c.dst_s3 = (long)n2*n1; c.dst_sv = (long)(n3 + 1)*n2*n1;
          c.di0 = q.a0; c.dj0 = q.b0; c.dk0 = q.cpos; c.ni = q.na; c.nj = q.nb; c.nk = 1;
          c.src_s2 = q.na; c.src_s3 = (long)q.na*q.nb;
        }
        unpack.push_back(c);
        roff[rf] += cnt;
      }
    }
  }
  ops.reserve_boxes(unpack.size());            // no allocation once NCCL is in flight
  if (ops.rc) return fail(ops.rc, "SMR flux correction: CUDA error");
  { int rcx = peer_exchange(m, m->peer_smr_flux); if (rcx) return rcx; }
  long total = 0;
  for (auto &c : unpack) { c.offset = total; total += (long)c.ni*c.nj*c.nk*c.nvar; }
  ops.copy_boxes(unpack, total);
  if (ops.rc) return fail(ops.rc, "SMR flux correction: CUDA error");
  CK(cudaGetLastError());
  return AB_OK;
}

int one_cycle(AbMesh *m) {
  const double *dtp = m->state + 6;     // dt of this cycle (0 past tlim, see k_mesh_new_dt)
  const bool user_src = (m->user_src || m->user_src_dev);
  if (m->has_user_bc || user_src) { int rc0 = read_state(m); if (rc0) return rc0; }   // host needs time, dt
  // per-block streams (see AbMesh::bstream); host-hook modes keep everything on the main stream
  // One launch per task over all local MeshBlocks (ab_batch.cuh) whenever the blocks are in one
  // allocation and in the same register-swap state and no host hook has to run between the tasks
  // of a block; refined meshes keep the block-by-block order of their prolongation hooks.
  static const bool no_batch = [] { const char *e = getenv("AB_NO_BATCH"); return e && e[0] == '1'; }();
  const bool bt = !no_batch && m->bstride != 0 && !m->smr && !m->has_user_bc && !user_src &&
                  !debug_sync() && plan_index(m) >= 0;
  const int nb = (int)m->lb.size();
  m->batch = bt;
  struct Leave {   // also on the error returns: task-level entry points run unbatched on the main stream
    AbMesh *m;
    ~Leave() { m->batch = false; for (auto &L : m->lb) L.stream = m->stream; }
  } leave{m};
  const bool ms = !bt && !m->bstream.empty() && !m->has_user_bc && !user_src && !debug_sync();
  for (size_t l = 0; l < m->lb.size(); ++l)
    m->lb[l].stream = ms ? m->bstream[l % m->bstream.size()] : m->stream;
  auto fork = [&]() {     // block streams continue after everything queued on the main stream
    if (!ms) return;
    cudaEventRecord(m->ev_main, m->stream);
    for (auto st : m->bstream) cudaStreamWaitEvent(st, m->ev_main, 0);
  };
  auto join = [&]() {     // the main stream continues after all block streams
    if (!ms) return;
    for (size_t i = 0; i < m->bstream.size(); ++i) {
      cudaEventRecord(m->ev_b[i], m->bstream[i]);
      cudaStreamWaitEvent(m->stream, m->ev_b[i], 0);
    }
  };
  for (int stage = 1; stage <= m->nstages; ++stage) {
    const int s = stage - 1;
    const int order = (m->p.integrator == AB_INT_VL2 && stage == 1) ? 1 : m->p.xorder;
    fork();
    if (bt) {
      LocalBlock &L0 = m->lb[0];
      for (int dir = 0; dir < m->ndim; ++dir) {
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (m->profile) {
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0, m->stream);
        }
        ab::launch_flux_dir(L0.d, L0.g, m->kp, order, dir, 0.0, dtp, m->stream, nb);
        if (m->profile) {
          cudaEventRecord(e1, m->stream);
          m->prof_ev.push_back({e0, e1});
          m->prof_slot.push_back(dir*3 + (order - 1));
        }
      }
      if (m->p.mhd) {
        bool all_cce = true;
        for (auto &L : m->lb) all_cce = all_cce && L.cc_e_valid;
        if (all_cce) ab::launch_corner_e(L0.d, m->stream, 1, nb);
        else for (auto &L : m->lb) ab::launch_corner_e(L.d, m->stream, L.cc_e_valid ? 1 : 0);
      }
      ab::launch_scalar_fluxes(L0.d, L0.g, m->kp, order, m->stream, nb);   // CALC_SCLRFLX
    } else for (auto &L : m->lb) {
      for (int dir = 0; dir < m->ndim; ++dir) {
        if (m->profile) {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0); cudaEventCreate(&e1);
          cudaEventRecord(e0, L.stream);
          ab::launch_flux_dir(L.d, L.g, m->kp, order, dir, 0.0, dtp, L.stream);
          cudaEventRecord(e1, L.stream);
          m->prof_ev.push_back({e0, e1});
          m->prof_slot.push_back(dir*3 + (order - 1));
        } else {
          ab::launch_flux_dir(L.d, L.g, m->kp, order, dir, 0.0, dtp, L.stream);
        }
        DBG(m, "flux dir " + std::to_string(dir) + " order " + std::to_string(order));
      }
      if (m->p.mhd) ab::launch_corner_e(L.d, L.stream, L.cc_e_valid ? 1 : 0);
      DBG(m, "corner_e");
      ab::launch_scalar_fluxes(L.d, L.g, m->kp, order, L.stream);   // CALC_SCLRFLX
      DBG(m, "scalar fluxes");
    }
    join();
    if (m->smr) { int rcs = smr_flux_correction(m); if (rcs) return rcs; }   // SEND/RECV_HYDFLX
    // EMF correction.  Overlapped schedule: the NCCL transfer runs on the comm stream while
    // IntegrateHydro / IntegrateScalars (which do not read EMFs) run on the compute stream.
    int rc = m->overlap ? emf_exchange_begin(m) : emf_exchange(m);
    if (rc) return rc;
    DBG(m, "emf exchange");
    const bool swap = (m->g1[s] == 0.0 && m->g2[s] == 1.0 && m->g3[s] == 0.0);
    const int zero_init = (stage == 1);   // StartupTaskList: u1, b1 ZeroClear (:1386-1397)
    fork();
    if (bt) {   // INT_HYD (+ SRC_TERM), INT_SCLR of every block: registers swap in lockstep
      if (swap) for (auto &L : m->lb) swap_cc(L);
      ab::launch_integrate_cc(m->lb[0].d, swap ? 1 : 2, zero_init, m->delta[s], swap ? 0 : m->g1[s],
                              swap ? 0 : m->g2[s], m->beta[s], 0.0, dtp, m->stream, -1, -1, 0, 0,
                              m->p.grav_acc, nb);
      if (m->p.nscalars > 0) {
        if (swap) for (auto &L : m->lb) swap_sc(L);
        ab::launch_integrate_cc(m->lb[0].d, swap ? 1 : 2, zero_init, m->delta[s], swap ? 0 : m->g1[s],
                                swap ? 0 : m->g2[s], m->beta[s], 0.0, dtp, m->stream, -1, -1, 0, 1,
                                nullptr, nb);
      }
    } else for (auto &L : m->lb) {   // INT_HYD (+ SRC_TERM), INT_SCLR (time_integrator.cpp:2141-2185)
      if (swap) {
        swap_cc(L);
        ab::launch_integrate_cc(L.d, 1, zero_init, m->delta[s], 0, 0, m->beta[s], 0.0, dtp,
                                L.stream, -1, -1, 0, 0, m->p.grav_acc);
        if (m->p.nscalars > 0) {
          swap_sc(L);
          ab::launch_integrate_cc(L.d, 1, zero_init, m->delta[s], 0, 0, m->beta[s], 0.0, dtp,
                                  L.stream, -1, -1, 0, 1);
        }
      } else {
        ab::launch_integrate_cc(L.d, 2, zero_init, m->delta[s], m->g1[s], m->g2[s], m->beta[s], 0.0,
                                dtp, L.stream, -1, -1, 0, 0, m->p.grav_acc);
        if (m->p.nscalars > 0)
          ab::launch_integrate_cc(L.d, 2, zero_init, m->delta[s], m->g1[s], m->g2[s], m->beta[s],
                                  0.0, dtp, L.stream, -1, -1, 0, 1);
      }
    }
    DBG(m, "integrate_cc");
    if (user_src) {   // SRC_TERM, user part: start-of-stage time, beta*dt
      for (auto &L : m->lb) {
        rc = user_source(m, L, m->h_time + m->sbeta[s]*m->h_dt, m->beta[s]*m->h_dt);
        if (rc) return rc;
      }
    }
    if (m->overlap) { rc = emf_exchange_end(m); if (rc) return rc; }
    if (m->p.mhd && bt) {   // INT_FLD of every block
      if (swap) for (auto &L : m->lb) swap_fc(L);
      ab::launch_integrate_fc(m->lb[0].d, swap ? 1 : 2, zero_init, m->delta[s], swap ? 0 : m->g1[s],
                              swap ? 0 : m->g2[s], m->beta[s], 0.0, dtp, m->stream, nb);
    } else if (m->p.mhd) for (auto &L : m->lb) {   // INT_FLD
      if (swap) {
        swap_fc(L);
        ab::launch_integrate_fc(L.d, 1, zero_init, m->delta[s], 0, 0, m->beta[s], 0.0, dtp, L.stream);
      } else {
        ab::launch_integrate_fc(L.d, 2, zero_init, m->delta[s], m->g1[s], m->g2[s], m->beta[s], 0.0, dtp, L.stream);
      }
    }
    join();
    DBG(m, "integrate_fc");
    if (m->has_user_bc) {   // PhysicalBoundary: t_end_stage, beta*dt (time_integrator.cpp:2045-2062)
      m->bc_time = m->h_time + m->ebeta[s]*m->h_dt;
      m->bc_dt = m->beta[s]*m->h_dt;
    }
    const int last = (stage == m->nstages);
    if (!m->overlap) {
      rc = m->smr ? smr_exchange(m) : bvals_exchange(m);
      if (rc) return rc;
      DBG(m, "bvals exchange");
      if (last) ab::launch_fill_u64(m->dtmin, (int)m->lb.size()*ab::DT_SLOTS, 0x7FEFFFFFFFFFFFFFull, m->stream);
      fork();
      if (bt) {
        primitives_all(m, last);
        for (auto &L : m->lb) { rc = physical_bcs(m, L); if (rc) return rc; }
      } else for (size_t l = 0; l < m->lb.size(); ++l) {
        LocalBlock &L = m->lb[l];
        if (m->smr) { rc = smr_prolongate(m, (int)l); if (rc) return rc; }   // PROLONG
        primitives(m, L, last);
        DBG(m, "primitives");
        rc = physical_bcs(m, L);
        if (rc) return rc;
        DBG(m, "physical bcs");
      }
      join();
    } else {
      // ghost zones travel while ConservedToPrimitive (+ CFL reduction) covers the active cells;
      // the ghost shell follows once they arrived
      rc = bvals_exchange_begin(m);
      if (rc) return rc;
      if (last) ab::launch_fill_u64(m->dtmin, (int)m->lb.size()*ab::DT_SLOTS, 0x7FEFFFFFFFFFFFFFull, m->stream);
      const bool one = bt && primitives_uniform(m);
      if (one) primitives_part(m, m->lb[0], 0, last, nb);
      else for (auto &L : m->lb) primitives_part(m, L, 0, last);
      rc = bvals_exchange_end(m);
      if (rc) return rc;
      if (one) primitives_part(m, m->lb[0], 1, 0, nb);
      else for (auto &L : m->lb) primitives_part(m, L, 1, 0);
      for (auto &L : m->lb) { rc = physical_bcs(m, L); if (rc) return rc; }
    }
    if (stage == m->nstages) {
      // record the dt this cycle used, then time += dt, ncycle++, NewTimeStep
      if (m->hist_n < m->hist_cap)
        CK(cudaMemcpyAsync(m->dt_hist + m->hist_n, m->state + 6, 8, cudaMemcpyDeviceToDevice, m->stream));
      m->hist_n++;
      rc = new_time_step(m, 1, true);
      if (rc) return rc;
      DBG(m, "new_time_step");
    }
  }
  for (auto &L : m->lb) L.stream = m->stream;   // task-level entry points run on the main stream
  m->batch = false;
  CK(cudaGetLastError());
  return AB_OK;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

const char *ab_last_error(void) { return g_err.c_str(); }

int ab_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

static int validate_params(const AbMeshParams *p);
static void host_setup(AbMesh *m, const AbMeshParams *p);

int ab_mesh_create(const AbMeshParams *p, AbMesh **out) {
  if (!p || !out) return fail(AB_ERR_ARG, "null argument");
  *out = nullptr;
  int vrc = validate_params(p);
  if (vrc) return vrc;
  if (ab_device_count() <= 0)
    return fail(AB_ERR_NO_DEVICE, "no CUDA device: libathena_b200 has no CPU fallback");
  CK(cudaSetDevice(p->device));
  AbMesh *m = new AbMesh();
  host_setup(m, p);
  CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&m->ev_pack, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&m->ev_recv, cudaEventDisableTiming));
  {
    const char *e = getenv("AB_OVERLAP");
    m->overlap = e && e[0] == '1';
  }
  int rc = alloc_blocks(m);
  if (rc) { delete m; return rc; }
  {
    int ns = (int)std::min<size_t>(m->lb.size(), 4);
    if (const char *e = getenv("AB_BLOCK_STREAMS")) ns = std::min<int>(atoi(e), (int)m->lb.size());
    if (m->lb.size() < 2 || ns < 2 || m->overlap) ns = 0;
    CK(cudaEventCreateWithFlags(&m->ev_main, cudaEventDisableTiming));
    for (int i = 0; i < ns; ++i) {
      cudaStream_t st; cudaEvent_t ev;
      CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      m->bstream.push_back(st); m->ev_b.push_back(ev);
    }
    // outside a cycle every block's work goes to the main stream: the task-level entry points
    // (ab_primitives, ab_physical_bcs ...) enqueue there in call order; one_cycle hands the
    // block streams out itself, between its fork and join
    for (size_t l = 0; l < m->lb.size(); ++l) m->lb[l].stream = m->stream;
  }
  CK(cudaMalloc(&m->state, 8*sizeof(double)));
  m->hist_cap = 1 << 16;
  CK(cudaMalloc(&m->dt_hist, sizeof(double)*m->hist_cap));
  m->h_time = p->start_time; m->h_dt = DBL_MAX; m->h_ncycle = 0;
  double h[8] = {p->start_time, DBL_MAX, p->tlim, m->cfl, DBL_MAX, 0.0, 0.0, 0.0};
  CK(cudaMemcpyAsync(m->state, h, sizeof(h), cudaMemcpyHostToDevice, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  *out = m;
  return AB_OK;
}

// Mesh ctor with mesh/refinement = static (src/mesh/mesh.cpp:323-548): the block list, levels,
// boundary flags and neighbour levels come from the host planner (ab_smr.cpp); every MeshBlock
// additionally gets the MeshRefinement's coarse buffers (src/mesh/mesh_refinement.cpp:40-100).
// Scope of this version: one process, hydro (+ passive scalars), uniformly spaced levels,
// MeshBlocks of at least 2*NGHOST cells, no user-enrolled boundary functions.
static int refined_host_setup(const AbMeshParams *p, const AbRefinementRegion *regions,
                              int nregions, bool dry, AbMesh **out) {
  if (!p || !out) return fail(AB_ERR_ARG, "null argument");
  *out = nullptr;
  int vrc = validate_params(p);
  if (vrc) return vrc;
  if (p->mhd) return fail(AB_ERR_ARG, "mesh refinement with MHD is not on the device path");
  for (int f = 0; f < 6; ++f)
    if (p->bc[f] == AB_BC_USER) return fail(AB_ERR_ARG, "mesh refinement with user-enrolled boundaries is not on the device path");
  for (int d = 0; d < 3; ++d)
    if (p->xrat[d] != 0.0 && p->xrat[d] != 1.0)
      return fail(AB_ERR_ARG, "mesh refinement needs uniformly spaced levels (x?rat = 1)");
  if (p->bx1 < 2*p->nghost || (p->nx2 > 1 && p->bx2 < 2*p->nghost) || (p->nx3 > 1 && p->bx3 < 2*p->nghost))
    return fail(AB_ERR_ARG, "mesh refinement on the device path needs MeshBlocks of at least 2*NGHOST cells");
  AbSmrPlan *plan = nullptr;
  if (ab_smr_plan_create(p, regions, nregions, &plan) != AB_OK)
    return fail(AB_ERR_ARG, std::string("refinement: ") + ab_smr_last_error());
  if (!dry && ab_device_count() <= 0) {
    ab_smr_plan_destroy(plan);
    return fail(AB_ERR_NO_DEVICE, "no CUDA device: libathena_b200 has no CPU fallback");
  }
  AbMesh *m = new AbMesh();
  m->dry = dry;
  m->smr = true; m->smr_plan = plan;
  // host_setup without build_block_list
  {
    AbMeshParams q = *p;
    q.nx1 = q.bx1; q.nx2 = q.bx2; q.nx3 = q.bx3;      // one dummy block: sets params, index ranges
    host_setup(m, &q);
    m->p.nx1 = p->nx1; m->p.nx2 = p->nx2; m->p.nx3 = p->nx3;
    m->f2 = p->nx2 > 1; m->f3 = p->nx3 > 1;
  }
  m->hb.clear(); m->lb_hb.clear(); m->gid_of.clear();
  m->nrb[0] = p->nx1/p->bx1; m->nrb[1] = p->nx2/p->bx2; m->nrb[2] = p->nx3/p->bx3;
  const int nb = ab_smr_plan_nblocks(plan);
  m->nbtotal = nb;
  std::vector<long> rows(5*(size_t)nb);
  ab_smr_plan_blocks(plan, rows.data(), nb);
  int root_level = 0;
  { int nbmax = std::max(m->nrb[0], std::max(m->nrb[1], m->nrb[2]));
    for (root_level = 0; (1 << root_level) < nbmax; ++root_level) {} }
  m->hb.resize(nb);
  const double mmin[3] = {p->x1min, p->x2min, p->x3min}, mmax[3] = {p->x1max, p->x2max, p->x3max};
  const int nxm[3] = {p->nx1, p->nx2, p->nx3};
  for (int g = 0; g < nb; ++g) {
    HostBlock &B = m->hb[g];
    B.gid = g; B.rank = 0; B.level = (int)rows[5*g]; B.dl = B.level - root_level;
    for (int d = 0; d < 3; ++d) B.lx[d] = rows[5*g + 1 + d];
    for (int d = 0; d < 3; ++d) {      // SetBlockSizeAndBoundaries at the block's level
      const long nrb = (long)m->nrb[d] << B.dl;
      if (d > 0 && nxm[d] == 1) {
        B.bmin[d] = mmin[d]; B.bmax[d] = mmax[d];
        B.bcs[2*d] = p->bc[2*d]; B.bcs[2*d+1] = p->bc[2*d+1];
        continue;
      }
      if (B.lx[d] == 0) { B.bmin[d] = mmin[d]; B.bcs[2*d] = p->bc[2*d]; }
      else { B.bmin[d] = gen_x(B.lx[d], nrb, mmin[d], mmax[d], 1.0, nxm[d]); B.bcs[2*d] = -1; }
      if (B.lx[d] == nrb - 1) { B.bmax[d] = mmax[d]; B.bcs[2*d+1] = p->bc[2*d+1]; }
      else { B.bmax[d] = gen_x(B.lx[d] + 1, nrb, mmin[d], mmax[d], 1.0, nxm[d]); B.bcs[2*d+1] = -1; }
    }
    int nbl[27];
    ab_smr_plan_neighbors(plan, g, nullptr, nbl);
    for (int k = 0; k < 3; ++k) for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
      B.nblevel[k][j][i] = nbl[(k*3 + j)*3 + i];
    for (int e = 0; e < 12; ++e) B.nedge_fine[e] = 1;
  }
  // Mesh::CalculateLoadBalance with unit costs over the Z-ordered block list of all levels
  // (mesh/amr_loadbalance.cpp:72-112): contiguous gid ranges per rank
  {
    int nranks = p->nranks;
    double totalcost = nb;
    int j = nranks - 1;
    double targetcost = totalcost/nranks, mycost = 0.0;
    for (int i = nb - 1; i >= 0; i--) {
      mycost += 1.0;
      m->hb[i].rank = j;
      if (mycost >= targetcost && j > 0) {
        j--;
        totalcost -= mycost;
        mycost = 0.0;
        targetcost = totalcost/(j + 1);
      }
    }
  }
  m->gid_start = -1;
  for (auto &B : m->hb) if (B.rank == p->rank) {
    if (m->gid_start < 0) m->gid_start = B.gid;
    m->lb_hb.push_back(&B);
  }
  if (m->lb_hb.empty()) {
    ab_smr_plan_destroy(plan);
    delete m;
    return fail(AB_ERR_ARG, "this rank owns no MeshBlock: use fewer ranks or smaller MeshBlocks");
  }
  {
    const long n = ab_smr_plan_transfers(plan, nullptr, 0);
    std::vector<long> tr(12*(size_t)n);
    ab_smr_plan_transfers(plan, tr.data(), n);
    m->smr_rows.resize(n);
    for (long i = 0; i < n; ++i) for (int c = 0; c < 12; ++c) m->smr_rows[i][c] = tr[12*i + c];
  }
  *out = m;
  return AB_OK;
}

int ab_mesh_create_refined(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                           AbMesh **out) {
  AbMesh *m = nullptr;
  int hrc = refined_host_setup(p, regions, nregions, false, &m);
  if (hrc) return hrc;
  *out = nullptr;
  CK(cudaSetDevice(p->device));
  CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&m->ev_pack, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&m->ev_recv, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&m->ev_main, cudaEventDisableTiming));
  m->overlap = false;
  int rc = alloc_blocks(m);
  if (rc) { ab_mesh_destroy(m); return rc; }
  for (auto &L : m->lb) L.stream = m->stream;
  // coarse buffers, coarse cell centres (coarse_flag branch of coordinates.cpp:92-160)
  const int ng = p->nghost, cng = (ng + 1)/2 + 1;
  const int bxs[3] = {p->bx1, p->bx2, p->bx3};
  m->smr_blk.resize(m->lb.size());
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    HostBlock &B = *L.hb;
    AbMesh::SmrBlk &sb = m->smr_blk[l];
    ab::SmrGeom &g = sb.g;
    memset(&g, 0, sizeof(g));
    g.nc1 = m->nc[0]; g.nc2 = m->nc[1]; g.nc3 = m->nc[2];
    g.is = m->is; g.js = m->js; g.ks = m->ks; g.ndim = m->ndim;
    g.cnc1 = bxs[0]/2 + 2*cng; g.cis = cng;
    g.cnc2 = m->f2 ? bxs[1]/2 + 2*cng : 1; g.cjs = m->f2 ? cng : 0;
    g.cnc3 = m->f3 ? bxs[2]/2 + 2*cng : 1; g.cks = m->f3 ? cng : 0;
    g.dx1f = L.d.dx1f; g.dx2f = L.d.dx2f; g.dx3f = L.d.dx3f;
    g.x1v = L.d.x1v; g.x2v = L.d.x2v; g.x3v = L.d.x3v;
    for (int d = 0; d < 3; ++d) {
      std::vector<double> xv = coarse_centres(m, B, d);
      CK(cudaMalloc(&sb.cxv[d], xv.size()*8));
      CK(cudaMemcpyAsync(sb.cxv[d], xv.data(), xv.size()*8, cudaMemcpyHostToDevice, m->stream));
      CK(cudaStreamSynchronize(m->stream));
    }
    g.cx1v = sb.cxv[0]; g.cx2v = sb.cxv[1]; g.cx3v = sb.cxv[2];
    const size_t cncc = (size_t)g.cnc1*g.cnc2*g.cnc3;
    CK(cudaMalloc(&sb.cu, cncc*m->nh*8)); CK(cudaMemsetAsync(sb.cu, 0, cncc*m->nh*8, m->stream));
    CK(cudaMalloc(&sb.cw, cncc*m->nh*8)); CK(cudaMemsetAsync(sb.cw, 0, cncc*m->nh*8, m->stream));
    if (p->nscalars > 0) {
      CK(cudaMalloc(&sb.cs, cncc*p->nscalars*8)); CK(cudaMemsetAsync(sb.cs, 0, cncc*p->nscalars*8, m->stream));
      CK(cudaMalloc(&sb.cr, cncc*p->nscalars*8)); CK(cudaMemsetAsync(sb.cr, 0, cncc*p->nscalars*8, m->stream));
    }
  }
  CK(cudaMalloc(&m->state, 8*sizeof(double)));
  m->hist_cap = 1 << 16;
  CK(cudaMalloc(&m->dt_hist, sizeof(double)*m->hist_cap));
  m->h_time = p->start_time; m->h_dt = DBL_MAX; m->h_ncycle = 0;
  double h[8] = {p->start_time, DBL_MAX, p->tlim, m->cfl, DBL_MAX, 0.0, 0.0, 0.0};
  CK(cudaMemcpyAsync(m->state, h, sizeof(h), cudaMemcpyHostToDevice, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  *out = m;
  return AB_OK;
}

// Host-only twin of ab_mesh_create_refined (no GPU needed): block list, levels, extents, boundary
// flags, neighbour levels; for ab_block_info / ab_block_level / ab_plan_geometry / ab_mesh_destroy.
int ab_plan_create_refined(const AbMeshParams *p, const AbRefinementRegion *regions, int nregions,
                           AbMesh **out) {
  if (!p || !out) return fail(AB_ERR_ARG, "null argument");
  return refined_host_setup(p, regions, nregions, true, out);
}

// Host-only plan of the same mesh (no device needed): used to check the cross-rank message
// pairing on CPU-only machines (tests with the gloo backend).
int ab_plan_create(const AbMeshParams *p, AbMesh **out) {
  if (!p || !out) return fail(AB_ERR_ARG, "null argument");
  *out = nullptr;
  int vrc = validate_params(p);
  if (vrc) return vrc;
  AbMesh *m = new AbMesh();
  m->dry = true;
  host_setup(m, p);
  *out = m;
  return AB_OK;
}

// rows of 5 longs: {direction (0 send, 1 recv), peer rank, key = dst_gid*64 + dst_bufid,
// element count, local block index}; kind 0 = ghost zones, 1 = EMF correction.
// Returns the number of rows (also when out == NULL or max_rows is too small).
int ab_plan_messages(const AbMesh *m, int kind, long *out, int max_rows) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  std::map<int, std::vector<Msg>> sends, recvs;
  peer_messages(m, kind, sends, recvs);
  int n = 0;
  for (int dir = 0; dir < 2; ++dir)
    for (auto &kv : (dir == 0 ? sends : recvs))
      for (auto &ms : kv.second) {
        if (out && n < max_rows) {
          long *r = out + 5L*n;
          r[0] = dir; r[1] = kv.first; r[2] = ms.key; r[3] = ms.count; r[4] = ms.lid;
        }
        n++;
      }
  return n;
}

// owner rank of every MeshBlock in gid order (Mesh::CalculateLoadBalance)
int ab_plan_ranklist(const AbMesh *m, int *out, int max_n) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  for (int g = 0; g < m->nbtotal && g < max_n; ++g) out[g] = m->hb[g].rank;
  return m->nbtotal;
}

static int validate_params(const AbMeshParams *p) {
  if (p->nx1 <= 0 || p->nx2 <= 0 || p->nx3 <= 0 || p->bx1 <= 0 || p->bx2 <= 0 || p->bx3 <= 0)
    return fail(AB_ERR_ARG, "mesh/nx? and meshblock/nx? must be positive");
  if (p->nx1 % p->bx1 || p->nx2 % p->bx2 || p->nx3 % p->bx3)
    return fail(AB_ERR_ARG, "the Mesh must be evenly divisible by the MeshBlock");
  // mesh.cpp:260-267: ghost zones are filled from the neighbour's ACTIVE cells
  if (p->bx1 < p->nghost || (p->nx2 > 1 && p->bx2 < p->nghost) ||
      (p->nx3 > 1 && p->bx3 < p->nghost))
    return fail(AB_ERR_ARG, "block_size must be larger than or equal to NGHOST cells with "
                            "uniform grid.");
  if (p->xorder < 1 || p->xorder > 3) return fail(AB_ERR_ARG, "time/xorder must be 1, 2 or 3");
  if (p->xorder == 3 && p->nghost < 3)
    return fail(AB_ERR_ARG, "xorder=3 (PPM) needs nghost >= 3 (reconstruction.cpp:90-99)");
  if (p->nghost < 2) return fail(AB_ERR_ARG, "nghost must be >= 2");
  if (p->solver < 0 || p->solver > AB_SOLVER_LLF) return fail(AB_ERR_ARG, "unknown Riemann solver");
  // configure.py:310-325
  if (p->mhd && p->solver == AB_SOLVER_HLLC) return fail(AB_ERR_ARG, "HLLC flux cannot be used with MHD");
  if (p->mhd && p->solver == AB_SOLVER_LHLLC) return fail(AB_ERR_ARG, "LHLLC flux cannot be used with MHD");
  if (!p->mhd && p->solver == AB_SOLVER_HLLD) return fail(AB_ERR_ARG, "HLLD flux can only be used with MHD");
  if (!p->mhd && p->solver == AB_SOLVER_LHLLD) return fail(AB_ERR_ARG, "LHLLD flux can only be used with MHD");
  for (int f = 0; f < 6; ++f)
    if (p->bc[f] != AB_BC_PERIODIC && p->bc[f] != AB_BC_OUTFLOW && p->bc[f] != AB_BC_REFLECT &&
        p->bc[f] != AB_BC_USER)
      return fail(AB_ERR_ARG, "unsupported boundary flag");
  {
    long n1 = p->bx1 + 2L*p->nghost + 1, n2 = (p->nx2 > 1 ? p->bx2 + 2L*p->nghost : 1) + 1,
         n3 = (p->nx3 > 1 ? p->bx3 + 2L*p->nghost : 1) + 1;
    const long nvar = p->nscalars > 5 ? p->nscalars : 5;   // scalar registers use int offsets too
    if (nvar*n1*n2*n3 >= (1L << 31))
      return fail(AB_ERR_ARG, "MeshBlock too large: registers must have < 2^31 elements "
                              "(kernels use 32-bit element offsets); use smaller MeshBlocks");
  }
  if (p->nranks < 1 || p->rank < 0 || p->rank >= p->nranks) return fail(AB_ERR_ARG, "bad rank/nranks");
  if (p->nscalars < 0 || p->nscalars > 16) return fail(AB_ERR_ARG, "nscalars must be in [0, 16]");
  if (p->eos != AB_EOS_ADIABATIC && p->eos != AB_EOS_ISOTHERMAL) return fail(AB_ERR_ARG, "unknown EOS");
  if (p->char_proj && p->eos == AB_EOS_ISOTHERMAL)
    return fail(AB_ERR_ARG, "characteristic reconstruction with isothermal EOS is not on the device path");
  if (p->eos == AB_EOS_ISOTHERMAL) {
    // configure.py:311-322
    if (p->solver == AB_SOLVER_HLLC || p->solver == AB_SOLVER_LHLLC || p->solver == AB_SOLVER_LHLLD)
      return fail(AB_ERR_ARG, "HLLC / LHLLC / LHLLD flux cannot be used with isothermal EOS");
    if (!(p->iso_sound_speed > 0.0)) return fail(AB_ERR_ARG, "hydro/iso_sound_speed must be set");
  }
  return AB_OK;
}

static void host_setup(AbMesh *m, const AbMeshParams *p) {
  m->p = *p;
  m->f2 = p->nx2 > 1; m->f3 = p->nx3 > 1;
  m->ndim = m->f3 ? 3 : (m->f2 ? 2 : 1);
  m->kp.gamma = p->gamma; m->kp.dfloor = p->dfloor; m->kp.pfloor = p->pfloor;
  m->kp.mhd = p->mhd; m->kp.solver = p->solver; m->kp.xorder = p->xorder;
  // EquationOfState ctor: scalar_floor_ = hydro/sfloor, default sqrt(1024*FLT_MIN)
  if (m->p.sfloor == 0.0) m->p.sfloor = std::sqrt(1024.0*(double)FLT_MIN);
  m->kp.sfloor = m->p.sfloor;
  m->kp.eos = p->eos; m->kp.iso_cs = p->iso_sound_speed;
  m->kp.char_proj = p->char_proj;
  const int nxd[3] = {p->nx1, p->nx2, p->nx3};
  for (int d = 0; d < 3; ++d)
    m->xrat[d] = (p->xrat[d] == 0.0 || nxd[d] == 1) ? 1.0 : p->xrat[d];
  m->nh = (p->eos == AB_EOS_ISOTHERMAL) ? 4 : 5;     // configure.py:374-377
  int ng = p->nghost;
  // MeshBlock index ranges (mesh/meshblock.cpp:55-80)
  m->is = ng; m->ie = ng + p->bx1 - 1; m->nc[0] = p->bx1 + 2*ng;
  if (m->f2) { m->js = ng; m->je = ng + p->bx2 - 1; m->nc[1] = p->bx2 + 2*ng; }
  if (m->f3) { m->ks = ng; m->ke = ng + p->bx3 - 1; m->nc[2] = p->bx3 + 2*ng; }
  set_integrator(m);
  build_block_list(m);
  m->gid_start = -1;
  for (auto &B : m->hb) if (B.rank == p->rank) {
    if (m->gid_start < 0) m->gid_start = B.gid;
    m->lb_hb.push_back(&B);
  }
}

int ab_mesh_destroy(AbMesh *m) {
  if (!m) return AB_OK;
  if (m->dry) {
    if (m->smr_plan) ab_smr_plan_destroy(m->smr_plan);
    delete m;
    return AB_OK;
  }
  cudaSetDevice(m->p.device);
  cudaStreamSynchronize(m->stream);
  for (auto &L : m->lb) for (void *q : L.debug_allocs) cudaFree(q);
  cudaFree(m->slab);
  cudaFree(m->emf_plans_dev);
  for (int i = 0; i < 24; ++i) { cudaFree(m->plan[i].pack); cudaFree(m->plan[i].phase1); cudaFree(m->plan[i].phase1r); cudaFree(m->plan[i].phase2); }
#ifndef AB_HOST_EMU
  for (auto &kv : m->peer_state) for (int b = 0; b < 2; ++b) {
    if (kv.second.p2p_send[b]) cudaIpcCloseMemHandle(kv.second.p2p_send[b]);
    cudaFree(kv.second.p2p_recv[b]);
  }
#endif
  cudaFree(m->p2p_token);
  for (auto &kv : m->bplan) { cudaFree(kv.second.pack); cudaFree(kv.second.phase1); cudaFree(kv.second.phase1r); cudaFree(kv.second.phase2); }
  for (auto &kv : m->peer_state) { cudaFree(kv.second.send); cudaFree(kv.second.recv); }
  for (auto &kv : m->peer_emf) { cudaFree(kv.second.send); cudaFree(kv.second.recv); }
  for (auto &kv : m->peer_smr) { cudaFree(kv.second.send); cudaFree(kv.second.recv); }
  for (auto &kv : m->peer_smr_flux) { cudaFree(kv.second.send); cudaFree(kv.second.recv); }
  cudaFree(m->state); cudaFree(m->dt_hist); cudaFree(m->dtmin);
  cudaFree(m->hist_partial); cudaFree(m->hist_out);
  for (auto &sb : m->smr_blk) {
    cudaFree(sb.cu); cudaFree(sb.cw); cudaFree(sb.cs); cudaFree(sb.cr);
    for (int d = 0; d < 3; ++d) cudaFree(sb.cxv[d]);
  }
  cudaFree(m->smr_boxes);
  if (m->smr_plan) ab_smr_plan_destroy(m->smr_plan);
  if (m->stg.active) {
    cudaStreamSynchronize(m->stg.h2d); cudaStreamSynchronize(m->stg.d2h);
    cudaFree(m->stg.in); cudaFree(m->stg.out);
    cudaEventDestroy(m->stg.in_ready); cudaEventDestroy(m->stg.in_free);
    cudaEventDestroy(m->stg.out_ready); cudaEventDestroy(m->stg.out_free);
    cudaStreamDestroy(m->stg.h2d); cudaStreamDestroy(m->stg.d2h);
  }
  if (m->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(m->comm);
  for (auto st : m->bstream) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  for (auto ev : m->ev_b) cudaEventDestroy(ev);
  if (m->ev_main) cudaEventDestroy(m->ev_main);
  if (m->ev_pack) cudaEventDestroy(m->ev_pack);
  if (m->ev_recv) cudaEventDestroy(m->ev_recv);
  if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
  cudaStreamDestroy(m->stream);
  delete m;
  return AB_OK;
}

int ab_mesh_nblocks_total(const AbMesh *m) { return m ? m->nbtotal : 0; }
int ab_block_level(const AbMesh *m, int lid) {
  if (!m || lid < 0 || lid >= (int)m->lb_hb.size()) return fail(AB_ERR_ARG, "bad argument");
  return m->lb_hb[lid]->level;
}
int ab_mesh_nblocks_local(const AbMesh *m) { return m ? (int)m->lb_hb.size() : 0; }

#define GET_L(m, lid)                                                              \
  if (!(m) || (lid) < 0 || (lid) >= (int)(m)->lb.size())                           \
    return fail(AB_ERR_ARG, "bad mesh handle or block index");                     \
  std::lock_guard<std::recursive_mutex> guard_((m)->mu);                           \
  LocalBlock &L = (m)->lb[lid];                                                    \
  (void)L

int ab_block_info(const AbMesh *m, int lid, long *info) {
  if (!m || lid < 0 || lid >= (int)m->lb_hb.size() || !info) return fail(AB_ERR_ARG, "bad argument");
  const HostBlock &B = *m->lb_hb[lid];
  info[0] = B.gid; info[1] = B.lx[0]; info[2] = B.lx[1]; info[3] = B.lx[2];
  info[4] = m->nc[0]; info[5] = m->nc[1]; info[6] = m->nc[2];
  info[7] = m->is; info[8] = m->ie; info[9] = m->js; info[10] = m->je; info[11] = m->ks;
  info[12] = m->ke;
  return AB_OK;
}

long ab_reg_size(const AbMesh *m, int lid, int reg) {
  if (!m || lid < 0 || lid >= (int)m->lb.size() || reg < 0 || reg >= AB_NREG) return -1;
  return m->lb[lid].regsize[reg];
}

int ab_upload(AbMesh *m, int lid, int reg, const double *host) {
  GET_L(m, lid);
  if (reg == AB_W || reg == AB_BCC) L.cc_e_valid = false;
  if (reg < 0 || reg >= AB_NREG || !host) return fail(AB_ERR_ARG, "bad register");
  double **slot = reg_slot(L, reg);
  if (!slot || !*slot) return fail(AB_ERR_ARG, "register not allocated in this configuration");
  CK(cudaMemcpyAsync(*slot, host, L.regsize[reg]*8, cudaMemcpyHostToDevice, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return AB_OK;
}

int ab_download(AbMesh *m, int lid, int reg, double *host) {
  GET_L(m, lid);
  if (reg < 0 || reg >= AB_NREG || !host) return fail(AB_ERR_ARG, "bad register");
  double **slot = reg_slot(L, reg);
  if (!slot || !*slot) return fail(AB_ERR_ARG, "register not allocated in this configuration");
  CK(cudaMemcpyAsync(host, *slot, L.regsize[reg]*8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return AB_OK;
}

// ---- pipelined staging ------------------------------------------------------------------------
int ab_stage_begin(AbMesh *m, const int *regs, int nregs) {
  if (!m || m->dry || !regs || nregs <= 0) return fail(AB_ERR_ARG, "bad argument");
  if (m->stg.active) return fail(AB_ERR_STATE, "staging already set up");
  CK(cudaSetDevice(m->p.device));
  AbMesh::Staging &S = m->stg;
  S.regs.assign(regs, regs + nregs);
  S.off.clear(); S.elems = 0;
  for (auto &L : m->lb)
    for (int r : S.regs) {
      if (r < 0 || r >= AB_NREG || L.regsize[r] <= 0) return fail(AB_ERR_ARG, "bad register");
      S.off.push_back(S.elems);
      S.elems += (size_t)((L.regsize[r] + 31)/32*32);     // keep every array 256-byte aligned
    }
  CK(cudaMalloc(&S.in, S.elems*8));
  CK(cudaMalloc(&S.out, S.elems*8));
  CK(cudaStreamCreateWithFlags(&S.h2d, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&S.d2h, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&S.in_ready, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&S.in_free, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&S.out_ready, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&S.out_free, cudaEventDisableTiming));
  S.active = true;
  return AB_OK;
}

// host[lid*nregs + r] (pinned) -> device staging, on the upload stream; returns at once.  The
// copies wait until the previous ab_stage_commit has drained the staging buffer.
int ab_stage_upload_all(AbMesh *m, const double *const *host) {
  if (!m || !m->stg.active || !host) return fail(AB_ERR_STATE, "staging not set up");
  AbMesh::Staging &S = m->stg;
  CK(cudaStreamWaitEvent(S.h2d, S.in_free, 0));      // no-op before the first commit
  const size_t nr = S.regs.size();
  for (size_t l = 0; l < m->lb.size(); ++l)
    for (size_t r = 0; r < nr; ++r)
      CK(cudaMemcpyAsync(S.in + S.off[l*nr + r], host[l*nr + r], m->lb[l].regsize[S.regs[r]]*8,
                         cudaMemcpyHostToDevice, S.h2d));
  CK(cudaEventRecord(S.in_ready, S.h2d));
  return AB_OK;
}

// staging -> registers on the compute stream, behind the upload; frees the staging buffer for
// the next ab_stage_upload_all.  Does not block the host.
int ab_stage_commit(AbMesh *m) {
  if (!m || !m->stg.active) return fail(AB_ERR_STATE, "staging not set up");
  AbMesh::Staging &S = m->stg;
  CK(cudaStreamWaitEvent(m->stream, S.in_ready, 0));
  const size_t nr = S.regs.size();
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    for (size_t r = 0; r < nr; ++r) {
      const int reg = S.regs[r];
      if (reg == AB_W || reg == AB_BCC) L.cc_e_valid = false;
      CK(cudaMemcpyAsync(*reg_slot(L, reg), S.in + S.off[l*nr + r], L.regsize[reg]*8,
                         cudaMemcpyDeviceToDevice, m->stream));
    }
  }
  CK(cudaEventRecord(S.in_free, m->stream));
  return AB_OK;
}

// registers -> staging on the compute stream (behind everything enqueued so far), then
// staging -> host[lid*nregs + r] (pinned) on the download stream.  Does not block the host;
// ab_stage_sync waits for the copies.
int ab_stage_download_all(AbMesh *m, double *const *host) {
  if (!m || !m->stg.active || !host) return fail(AB_ERR_STATE, "staging not set up");
  AbMesh::Staging &S = m->stg;
  CK(cudaStreamWaitEvent(m->stream, S.out_free, 0));  // previous download has left the buffer
  const size_t nr = S.regs.size();
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    for (size_t r = 0; r < nr; ++r)
      CK(cudaMemcpyAsync(S.out + S.off[l*nr + r], *reg_slot(L, S.regs[r]),
                         L.regsize[S.regs[r]]*8, cudaMemcpyDeviceToDevice, m->stream));
  }
  CK(cudaEventRecord(S.out_ready, m->stream));
  CK(cudaStreamWaitEvent(S.d2h, S.out_ready, 0));
  for (size_t l = 0; l < m->lb.size(); ++l)
    for (size_t r = 0; r < nr; ++r)
      CK(cudaMemcpyAsync(host[l*nr + r], S.out + S.off[l*nr + r],
                         m->lb[l].regsize[S.regs[r]]*8, cudaMemcpyDeviceToHost, S.d2h));
  CK(cudaEventRecord(S.out_free, S.d2h));
  return AB_OK;
}

int ab_stage_sync(AbMesh *m) {
  if (!m || !m->stg.active) return fail(AB_ERR_STATE, "staging not set up");
  CK(cudaStreamSynchronize(m->stg.h2d));
  CK(cudaStreamSynchronize(m->stream));
  CK(cudaStreamSynchronize(m->stg.d2h));
  return AB_OK;
}

// Host-computed geometry of local block `lid` along direction dir (0..2), no device needed:
// what 0 x?f (nc+1), 1 x?v, 2 dx?f, 3 / 4 PLM face weights wp / wm, 5 nonuniform-reconstruction
// table (nc*13, only when x?rat != 1), 6 / 7 cell-centred-field weights lw / rw.
// Returns the number of doubles (also when out == NULL), 0 when that array does not exist.
int ab_plan_geometry(const AbMesh *m, int lid, int dir, int what, double *out, int max_n) {
  if (!m || lid < 0 || lid >= (int)m->lb_hb.size() || dir < 0 || dir > 2 || what < 0 || what > 8)
    return fail(AB_ERR_ARG, "bad argument");
  const HostBlock &B = *m->lb_hb[lid];
  if (what == 8) {    // refined meshes: cell centres of the coarse buffers
    if (!m->smr) return 0;
    const std::vector<double> v = coarse_centres(m, B, dir);
    if (out) for (size_t i = 0; i < v.size() && (int)i < max_n; ++i) out[i] = v[i];
    return (int)v.size();
  }
  const AbMeshParams &p = m->p;
  const double mmin[3] = {p.x1min, p.x2min, p.x3min}, mmax[3] = {p.x1max, p.x2max, p.x3max};
  const int nxm[3] = {p.nx1, p.nx2, p.nx3}, bxs[3] = {p.bx1, p.bx2, p.bx3};
  const int s0[3] = {m->is, m->js, m->ks}, e0[3] = {m->ie, m->je, m->ke};
  const int nc = m->nc[dir];
  std::vector<double> xf, xv, dxf, tab, lw(nc, 0.5), rw(nc, 0.5), wp(nc), wm(nc);
  make_coords(nxm[dir] << B.dl, bxs[dir], p.nghost, B.lx[dir], mmin[dir], mmax[dir], B.bmin[dir],
              B.bmax[dir], nc, B.bcs[2*dir] == AB_BC_REFLECT, B.bcs[2*dir+1] == AB_BC_REFLECT,
              m->xrat[dir], xf, xv, dxf);
  for (int c = 0; c < nc; ++c) {
    wp[c] = (xf[c+1] - xv[c])/dxf[c];
    wm[c] = (xv[c] - xf[c])/dxf[c];
  }
  if (m->xrat[dir] != 1.0)
    make_recon_table(dir, nc, s0[dir], e0[dir], p.nghost, xf, xv, dxf, tab, lw, rw);
  const std::vector<double> *src[8] = {&xf, &xv, &dxf, &wp, &wm, &tab, &lw, &rw};
  const std::vector<double> &v = *src[what];
  if (out) for (size_t i = 0; i < v.size() && (int)i < max_n; ++i) out[i] = v[i];
  return (int)v.size();
}

int ab_download_coord(AbMesh *m, int lid, int which, double *host) {
  GET_L(m, lid);
  if (which < 0 || which > 8 || !host) return fail(AB_ERR_ARG, "bad coordinate id");
  CK(cudaMemcpyAsync(host, L.coord_dev[which], L.coord_n[which]*8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return AB_OK;
}

int ab_comm_unique_id(unsigned char id[128]) {
  if (!g_nccl.load()) return fail(AB_ERR_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId u;
  NK(g_nccl.GetUniqueId(&u));
  memcpy(id, u.internal, 128);
  return AB_OK;
}

int ab_comm_init(AbMesh *m, const unsigned char id[128]) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  if (m->p.nranks <= 1) return AB_OK;
  if (!g_nccl.load()) return fail(AB_ERR_NCCL, "cannot load libnccl.so.2");
  CK(cudaSetDevice(m->p.device));
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  NK(g_nccl.CommInitRank(&m->comm, m->p.nranks, u, m->p.rank));
  // direct ghost-zone exchange over peer memory unless AB_P2P=0 (falls back to NCCL by itself when
  // the peers' buffers cannot be mapped); the overlapped schedule keeps its NCCL transfers
  m->p2p = 1;
  if (const char *e = getenv("AB_P2P")) m->p2p = (e[0] == '0') ? 0 : 1;
  return AB_OK;
}

int ab_cons2prim(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku) {
  GET_L(m, lid);
  L.cc_e_valid = false;
  ab::launch_cons2prim(L.d, m->kp, il, iu, jl, ju, kl, ku, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_prim2cons(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku) {
  GET_L(m, lid);
  ab::launch_prim2cons(L.d, m->kp, il, iu, jl, ju, kl, ku, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_primitives(AbMesh *m, int lid) {
  GET_L(m, lid);
  primitives(m, L);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_calc_fluxes(AbMesh *m, int lid, int order, double dt) {
  GET_L(m, lid);
  if (order < 1 || order > 3) return fail(AB_ERR_ARG, "order must be 1, 2 or 3");
  ab::launch_fluxes(L.d, L.g, m->kp, order, dt, nullptr, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_corner_e(AbMesh *m, int lid) {
  GET_L(m, lid);
  if (!m->p.mhd) return fail(AB_ERR_STATE, "ComputeCornerE needs MAGNETIC_FIELDS_ENABLED");
  ab::launch_corner_e(L.d, m->stream, L.cc_e_valid ? 1 : 0);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_weighted_ave(AbMesh *m, int lid, int out_reg, int in_reg, const double w[5]) {
  GET_L(m, lid);
  if (!w) return fail(AB_ERR_ARG, "null weights");
  if (w[2] != 0.0 || w[3] != 0.0 || w[4] != 0.0)
    return fail(AB_ERR_ARG, "only two-register averages (u,u1 / b,b1) are supported");
  if ((out_reg == AB_U || out_reg == AB_U1) && (in_reg == AB_U || in_reg == AB_U1)) {
    ab::launch_weighted_ave_cc(L.d, out_reg == AB_U ? L.d.u : L.d.u1,
                               in_reg == AB_U ? L.d.u : L.d.u1, w[0], w[1], m->stream, m->nh);
  } else if ((out_reg == AB_S || out_reg == AB_S1) && (in_reg == AB_S || in_reg == AB_S1) &&
             m->p.nscalars > 0) {
    ab::launch_weighted_ave_cc(L.d, out_reg == AB_S ? L.d.s : L.d.s1,
                               in_reg == AB_S ? L.d.s : L.d.s1, w[0], w[1], m->stream,
                               m->p.nscalars);
  } else if ((out_reg == AB_B_X1F || out_reg == AB_B1_X1F) &&
             (in_reg == AB_B_X1F || in_reg == AB_B1_X1F) && m->p.mhd) {
    ab::launch_weighted_ave_fc(L.d, out_reg == AB_B_X1F ? L.d.b : L.d.b1,
                               in_reg == AB_B_X1F ? L.d.b : L.d.b1, w[0], w[1], m->stream);
  } else {
    return fail(AB_ERR_ARG, "bad register pair for WeightedAve");
  }
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_swap(AbMesh *m, int lid, int reg) {
  GET_L(m, lid);
  if (reg == AB_U) swap_cc(L);
  else if (reg == AB_B_X1F && m->p.mhd) swap_fc(L);
  else if (reg == AB_S && m->p.nscalars > 0) swap_sc(L);
  else return fail(AB_ERR_ARG, "ab_swap: reg must be AB_U, AB_B_X1F or AB_S");
  return AB_OK;
}
int ab_zero(AbMesh *m, int lid, int reg) {
  GET_L(m, lid);
  if (reg == AB_U1) {
    CK(cudaMemsetAsync(L.d.u1, 0, L.regsize[AB_U1]*8, m->stream));
  } else if (reg == AB_B1_X1F && m->p.mhd) {
    for (int d = 0; d < 3; ++d) CK(cudaMemsetAsync(L.d.b1[d], 0, L.regsize[AB_B1_X1F + d]*8, m->stream));
  } else if (reg == AB_S1 && m->p.nscalars > 0) {
    CK(cudaMemsetAsync(L.d.s1, 0, L.regsize[AB_S1]*8, m->stream));
  } else {
    return fail(AB_ERR_ARG, "ab_zero: reg must be AB_U1, AB_B1_X1F or AB_S1");
  }
  ab::g_launches++;
  return AB_OK;
}
int ab_add_flux_div(AbMesh *m, int lid, double wght) {
  GET_L(m, lid);
  ab::launch_integrate_cc(L.d, 0, 0, 0.0, 0.0, 0.0, 1.0, wght, nullptr, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_add_source_terms(AbMesh *m, int lid, double time, double dt) {
  GET_L(m, lid);
  const double *g = m->p.grav_acc;
  if (g[0] != 0.0 || g[1] != 0.0 || g[2] != 0.0) ab::launch_const_accel(L.d, g, dt, m->stream);
  CK(cudaGetLastError());
  if (m->user_src || m->user_src_dev) return user_source(m, L, time, dt);
  return AB_OK;
}
int ab_calc_scalar_fluxes(AbMesh *m, int lid, int order) {
  GET_L(m, lid);
  if (m->p.nscalars <= 0) return fail(AB_ERR_STATE, "NSCALARS == 0");
  if (order < 1 || order > 3) return fail(AB_ERR_ARG, "order must be 1, 2 or 3");
  ab::launch_scalar_fluxes(L.d, L.g, m->kp, order, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_add_scalar_flux_div(AbMesh *m, int lid, double wght) {
  GET_L(m, lid);
  if (m->p.nscalars <= 0) return fail(AB_ERR_STATE, "NSCALARS == 0");
  ab::launch_integrate_cc(L.d, 0, 0, 0.0, 0.0, 0.0, 1.0, wght, nullptr, m->stream, -1, -1, 0, 1);
  CK(cudaGetLastError());
  return AB_OK;
}
static int scalar_eos_range(AbMesh *m, int lid, int to_cons, int il, int iu, int jl, int ju,
                            int kl, int ku) {
  GET_L(m, lid);
  if (m->p.nscalars <= 0) return fail(AB_ERR_STATE, "NSCALARS == 0");
  if (il < 0 || iu >= m->nc[0] || jl < 0 || ju >= m->nc[1] || kl < 0 || ku >= m->nc[2] ||
      il > iu || jl > ju || kl > ku)
    return fail(AB_ERR_ARG, "cell range outside the MeshBlock");
  ab::launch_scalar_eos(L.d, m->kp, to_cons, il, iu, jl, ju, kl, ku, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_scalar_cons2prim(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku) {
  return scalar_eos_range(m, lid, 0, il, iu, jl, ju, kl, ku);
}
int ab_scalar_prim2cons(AbMesh *m, int lid, int il, int iu, int jl, int ju, int kl, int ku) {
  return scalar_eos_range(m, lid, 1, il, iu, jl, ju, kl, ku);
}
int ab_ct(AbMesh *m, int lid, double wght) {
  GET_L(m, lid);
  if (!m->p.mhd) return fail(AB_ERR_STATE, "CT needs MAGNETIC_FIELDS_ENABLED");
  ab::launch_integrate_fc(L.d, 0, 0, 0.0, 0.0, 0.0, 1.0, wght, nullptr, m->stream);
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_physical_bcs(AbMesh *m, int lid) {
  GET_L(m, lid);
  { const int rcb = physical_bcs(m, L); if (rcb) return rcb; }
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_physical_bcs_at(AbMesh *m, int lid, double time, double dt) {
  GET_L(m, lid);
  m->bc_time = time; m->bc_dt = dt;      // what user-enrolled boundary functions are handed
  { const int rcb = physical_bcs(m, L); if (rcb) return rcb; }
  CK(cudaGetLastError());
  return AB_OK;
}
int ab_new_block_dt(AbMesh *m, int lid, double *dt_out) {
  GET_L(m, lid);
  ab::launch_fill_u64(L.dtmin, ab::DT_SLOTS, 0x7FEFFFFFFFFFFFFFull, m->stream);
  ab::launch_new_block_dt(L.d, m->kp, L.dtmin, m->stream);
  unsigned long long bits[ab::DT_SLOTS];
  CK(cudaMemcpyAsync(bits, L.dtmin, sizeof(bits), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  double v = DBL_MAX;
  for (int q = 0; q < ab::DT_SLOTS; ++q) {
    double t;
    memcpy(&t, &bits[q], 8);
    if (t < v) v = t;
  }
  if (dt_out) *dt_out = v*m->cfl;     // new_blockdt.cpp:164
  return AB_OK;
}

int ab_bvals_send(AbMesh *m, int lid, int var) {
  GET_L(m, lid);
  if (m->smr) return fail(AB_ERR_STATE, "per-block boundary tasks are not available on a refined mesh");
  if (var < 0 || var > 2 || !var_applies(m, var)) return fail(AB_ERR_ARG, "boundary variable not in this build");
  return block_bvals_send(m, lid, var);
}
int ab_bvals_recv_try(AbMesh *m, int lid, int var) {
  GET_L(m, lid);
  if (var < 0 || var > 2 || !var_applies(m, var)) return fail(AB_ERR_ARG, "boundary variable not in this build");
  return block_bvals_recv_try(m, lid, var);
}
int ab_bvals_set(AbMesh *m, int lid, int var) {
  GET_L(m, lid);
  if (m->smr) return fail(AB_ERR_STATE, "per-block boundary tasks are not available on a refined mesh");
  if (var < 0 || var > 2 || !var_applies(m, var)) return fail(AB_ERR_ARG, "boundary variable not in this build");
  if (!m->bcomm.empty() && m->bcomm[lid].recvd[var] == 0)
    return fail(AB_ERR_STATE, "ab_bvals_set before ab_bvals_recv_try succeeded");
  return block_bvals_set(m, lid, var);
}
int ab_emf_send(AbMesh *m, int lid) {
  GET_L(m, lid);
  if (!m->p.mhd) return fail(AB_ERR_STATE, "EMF correction needs MAGNETIC_FIELDS_ENABLED");
  return block_emf_send(m, lid);
}
int ab_emf_recv_try(AbMesh *m, int lid) {
  GET_L(m, lid);
  if (!m->p.mhd) return fail(AB_ERR_STATE, "EMF correction needs MAGNETIC_FIELDS_ENABLED");
  return block_emf_recv_try(m, lid);
}
int ab_clear_boundary(AbMesh *m, int lid) {
  GET_L(m, lid);
  // nothing is pending between stages (no persistent requests to reset); a block that sent
  // without every neighbour having received would be a scheduler error worth reporting
  if (m->bcomm.size() == m->lb.size())
    for (int v = 0; v < 3; ++v)
      if (m->bcomm[lid].recvd[v] != m->bcomm[lid].sent[v])
        return fail(AB_ERR_STATE, "ClearBoundary: a boundary variable was sent but not received");
  return AB_OK;
}

int ab_emf_exchange(AbMesh *m) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  std::lock_guard<std::recursive_mutex> guard_(m->mu);
  return emf_exchange(m);
}
int ab_bvals_exchange(AbMesh *m) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  std::lock_guard<std::recursive_mutex> guard_(m->mu);
  return m->smr ? smr_exchange(m) : bvals_exchange(m);
}

int ab_enroll_user_explicit_source_function(AbMesh *m, AbSrcTermFunc fn, void *user) {
  if (!m || !fn) return fail(AB_ERR_ARG, "null mesh or function");
  m->user_src = fn; m->user_src_dev = nullptr; m->user_src_arg = user;
  return AB_OK;
}
int ab_enroll_user_explicit_source_function_device(AbMesh *m, AbSrcTermFuncDevice fn, void *user) {
  if (!m || !fn) return fail(AB_ERR_ARG, "null mesh or function");
  m->user_src_dev = fn; m->user_src = nullptr; m->user_src_arg = user;
  return AB_OK;
}
int ab_enroll_user_boundary_function(AbMesh *m, int face, AbBValFunc fn, void *user) {
  if (!m || face < 0 || face > 5 || !fn) return fail(AB_ERR_ARG, "bad face or null function");
  // Mesh::EnrollUserBoundaryFunction: the face must carry the "user" flag
  if (m->p.bc[face] != AB_BC_USER)
    return fail(AB_ERR_ARG, "boundary function enrolled on a face whose flag is not user");
  m->user_bc[face] = fn; m->user_bc_arg[face] = user;
  return AB_OK;
}

int ab_mesh_initialize(AbMesh *m) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  CK(cudaSetDevice(m->p.device));
  m->has_user_bc = false;
  for (int f = 0; f < 2*m->ndim; ++f) if (m->p.bc[f] == AB_BC_USER) {
    if (!m->user_bc[f])   // bvals.cpp:328-335
      return fail(AB_ERR_STATE, "a user-defined boundary is specified but the actual boundary "
                                "function is not enrolled");
    m->has_user_bc = true;
  }
  m->bc_time = m->h_time; m->bc_dt = 0.0;   // ApplyPhysicalBoundaries(time, 0.0, ...) mesh.cpp:1574
  for (auto &L : m->lb) L.stream = m->stream;
  int rc = m->smr ? smr_exchange(m) : bvals_exchange(m);
  if (rc) return rc;
  for (size_t l = 0; l < m->lb.size(); ++l) {
    LocalBlock &L = m->lb[l];
    if (m->smr) { rc = smr_prolongate(m, (int)l); if (rc) return rc; }   // mesh.cpp:1527-1528
    primitives(m, L);
    rc = physical_bcs(m, L);
    if (rc) return rc;
  }
  rc = new_time_step(m, 0);
  if (rc) return rc;
  return read_state(m);
}

int ab_mesh_cycles(AbMesh *m, int ncycles) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  CK(cudaSetDevice(m->p.device));
  m->hist_n = 0;
  for (int c = 0; c < ncycles; ++c) {
    if (!m->async && !(m->h_time < m->p.tlim)) break;       // main.cpp:430
    int rc = one_cycle(m);
    if (rc) return rc;
    if (!m->async) { rc = read_state(m); if (rc) return rc; }
  }
  return AB_OK;
}

int ab_mesh_set_async(AbMesh *m, int async) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  m->async = async;
  return AB_OK;
}

int ab_mesh_state(AbMesh *m, double *time, double *dt, long *ncycle) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  int rc = read_state(m);
  if (rc) return rc;
  if (time) *time = m->h_time;
  if (dt) *dt = m->h_dt;
  if (ncycle) *ncycle = m->h_ncycle;
  return AB_OK;
}

int ab_mesh_set_time_dt(AbMesh *m, double time, double dt) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  double h[2] = {time, dt}, eff = (time < m->p.tlim) ? dt : 0.0;
  CK(cudaMemcpyAsync(m->state, h, sizeof(h), cudaMemcpyHostToDevice, m->stream));
  CK(cudaMemcpyAsync(m->state + 6, &eff, sizeof(eff), cudaMemcpyHostToDevice, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  m->h_time = time; m->h_dt = dt;
  return AB_OK;
}

int ab_history(AbMesh *m, double *out, int max_n) {
  if (!m || !out) return fail(AB_ERR_ARG, "null argument");
  if (m->dry) return fail(AB_ERR_STATE, "host-only plan");
  CK(cudaSetDevice(m->p.device));
  const int nq = m->nh + 3 + (m->p.mhd ? 3 : 0) + m->p.nscalars;
  if (max_n < nq) return fail(AB_ERR_ARG, "ab_history: output array too small");
  if (!m->hist_partial) {
    CK(cudaMalloc(&m->hist_partial, sizeof(double)*32*ab::history_grid()));
    CK(cudaMalloc(&m->hist_out, sizeof(double)*32));
  }
  int first = 1;
  for (auto &L : m->lb) {
    ab::launch_history(L.d, m->p.mhd, nq, first, m->hist_partial, m->hist_out, m->stream);
    first = 0;
  }
  CK(cudaGetLastError());
  if (m->p.nranks > 1) {   // MPI_Reduce(MPI_SUM) of history.cpp:203-211
    if (!m->comm) return fail(AB_ERR_STATE, "nranks > 1 but ab_comm_init was not called");
    NK(g_nccl.AllReduce(m->hist_out, m->hist_out, nq, NCCL_FLOAT64, NCCL_SUM, m->comm, m->stream));
  }
  CK(cudaMemcpyAsync(out, m->hist_out, sizeof(double)*nq, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return nq;
}

int ab_mesh_dt_history(AbMesh *m, double *out, int max_n) {
  if (!m || !out) return fail(AB_ERR_ARG, "null argument");
  int n = std::min(std::min(m->hist_n, m->hist_cap), max_n);
  if (n > 0) {
    CK(cudaMemcpyAsync(out, m->dt_hist, 8*(size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  return n;
}

int ab_mesh_profile(AbMesh *m, int enable) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  m->profile = enable;
  return AB_OK;
}

// out[0..8] = accumulated ms, out[9..17] = launch counts; slot = dir*3 + (order-1); resets
int ab_mesh_profile_read(AbMesh *m, double *out) {
  if (!m || !out) return fail(AB_ERR_ARG, "null argument");
  CK(cudaStreamSynchronize(m->stream));
  for (size_t i = 0; i < m->prof_ev.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->prof_ev[i].first, m->prof_ev[i].second);
    m->prof_ms[m->prof_slot[i]] += ms;
    m->prof_n[m->prof_slot[i]] += 1;
    cudaEventDestroy(m->prof_ev[i].first);
    cudaEventDestroy(m->prof_ev[i].second);
  }
  m->prof_ev.clear(); m->prof_slot.clear();
  for (int i = 0; i < 9; ++i) { out[i] = m->prof_ms[i]; out[9+i] = (double)m->prof_n[i]; m->prof_ms[i] = 0; m->prof_n[i] = 0; }
  return AB_OK;
}

long ab_mesh_launch_count(const AbMesh *m) { (void)m; return ab::g_launches; }
void *ab_mesh_stream(AbMesh *m) { return m ? (void *)m->stream : nullptr; }
int ab_mesh_sync(AbMesh *m) {
  if (!m) return fail(AB_ERR_ARG, "null mesh");
  CK(cudaStreamSynchronize(m->stream));
  return AB_OK;
}

}  // extern "C"
