// ab_kernels.h -- launchers of the sm_100a kernels (implemented in ab_kernels.cu).
#ifndef AB_KERNELS_H_
#define AB_KERNELS_H_
#include <cuda_runtime.h>
#include <atomic>
#include "ab_types.h"

namespace ab {

// per-direction PLM face weights precomputed on the host with the reference's expression
// (x?f(i+1)-x?v(i))/dx?f(i) and (x?v(i)-x?f(i))/dx?f(i)  (reconstruct/plm.cpp:114-119)
// nu[dir]: nullptr for uniform spacing, else the per-index table of the nonuniform PLM / PPM
// branches (NUG doubles per cell index, ab_physics.cuh); bcw: nullptr when every direction is
// uniform, else the CalculateCellCenteredField weights lw, rw of x1, x2, x3 back to back
// (field/field.cpp:139-172)
struct ReconGeom { const double *wp[3]; const double *wm[3]; const double *nu[3]; };

// nb (last argument of the per-block task launchers): one launch over nb MeshBlocks of a rank,
// `b` / `g` being the views of the first of them (ab_batch.cuh); 1 = that block alone.
// `dt_ptr` (device) wins over `dt_val` when non-null: the cycle loop keeps dt on the device.
// flags bit0: also write cc_e; bit1: also reduce NewBlockTimeStep over active cells into dtmin
void launch_cons2prim(const BlkDev &b, const Params &p, int il, int iu, int jl, int ju, int kl,
                      int ku, cudaStream_t s, int flags = 0, unsigned long long *dtmin = nullptr,
                      int nb = 1);
// the same over up to six cell boxes {il,iu,jl,ju,kl,ku} in one launch (no CFL reduction)
void launch_cons2prim_boxes(const BlkDev &b, const Params &p, int nbox, const int (*box)[6],
                            cudaStream_t s, int flags = 0, int nb = 1);
void launch_prim2cons(const BlkDev &b, const Params &p, int il, int iu, int jl, int ju, int kl,
                      int ku, cudaStream_t s);
void launch_calc_bcc(const BlkDev &b, int il, int iu, int jl, int ju, int kl, int ku,
                     cudaStream_t s);
void launch_fluxes(const BlkDev &b, const ReconGeom &g, const Params &p, int order,
                   double dt_val, const double *dt_ptr, cudaStream_t s);
void launch_flux_dir_nu(const BlkDev &b, const ReconGeom &g, const Params &p, int order, int dir,
                        double dt_val, const double *dt_ptr, cudaStream_t s, int nb = 1);
void launch_flux_dir(const BlkDev &b, const ReconGeom &g, const Params &p, int order, int dir,
                     double dt_val, const double *dt_ptr, cudaStream_t s, int nb = 1);
// have_cc_e: cc_e was already written by cons2prim (flags bit0) over [is-1,ie+1]^dim
void launch_corner_e(const BlkDev &b, cudaStream_t s, int have_cc_e = 0, int nb = 1);
// plans_dev: device array of the plans of the nb blocks (first element = block b's plan)
void launch_emf_pack(const BlkDev &b, const EmfPlan *plans_dev, cudaStream_t s, int nb = 1);
void launch_emf_apply(const BlkDev &b, const EmfPlan *plans_dev, cudaStream_t s, int nb = 1);

// WeightedAve special-casing of the reference (mesh/weighted_ave.cpp) for out = w0*out + w1*in
void launch_weighted_ave_cc(const BlkDev &b, double *out, const double *in, double w0,
                            double w1, cudaStream_t s, int nvar);
void launch_weighted_ave_fc(const BlkDev &b, double *const out[3], double *const in[3],
                            double w0, double w1, cudaStream_t s);

// Fused IntegrateHydro on active cells (task_list/time_integrator.cpp:1563-1612 +
// hydro/add_flux_divergence.cpp:39-96).  mode 0: u -= wght*div only.  mode 1 (registers
// already pointer-swapped): u = (zero_init ? 0 : u) [+ delta*u1] ; u -= wght*div.
// mode 2: u1 = (zero_init ? 0 : u1) [+ delta*u]; u = wave(u; u1; g1, g2); u -= wght*div.
// wght = beta * dt.
// gacc != nullptr: constant acceleration g[3] added in the same pass (SRC_TERM task).
// [kl,ku]: plane range (kl < 0: all active planes); grid > 0 caps the number of CTAs
// (grid-stride loop) so that the kernel can share the SMs with a concurrent flux kernel.
void launch_integrate_cc(const BlkDev &b, int mode, int zero_init, double delta, double g1,
                         double g2, double beta, double dt_val, const double *dt_ptr,
                         cudaStream_t s, int kl = -1, int ku = -1, int grid = 0,
                         int scalars = 0, const double *gacc = nullptr, int nb = 1);
// passive scalars: s_flux from r and the hydro mass flux; r <-> s conversions on a cell range
void launch_scalar_fluxes(const BlkDev &b, const ReconGeom &g, const Params &p, int order,
                          cudaStream_t s, int nb = 1);
void launch_scalar_eos(const BlkDev &b, const Params &p, int to_cons, int il, int iu, int jl,
                       int ju, int kl, int ku, cudaStream_t s, int nb = 1);
void launch_const_accel(const BlkDev &b, const double *g, double dt, cudaStream_t s);
// Same for the face field + Field::CT (field/ct.cpp:31-116)
void launch_integrate_fc(const BlkDev &b, int mode, int zero_init, double delta, double g1,
                         double g2, double beta, double dt_val, const double *dt_ptr,
                         cudaStream_t s, int nb = 1);

// ghost exchange: apply `n` box copies (descriptors in device memory, CopyBox::offset = exclusive
// prefix of the element counts, total_elems = their sum)
// chunked != 0: the per-256-element first-box table follows the n boxes in the same allocation
void launch_copy_boxes(const CopyBox *boxes_dev, int n, long total_elems, cudaStream_t s,
                       int chunked = 0);
// static mesh refinement (ab_smr_kernels.cu)
void launch_smr_restrict(const SmrGeom &g, const double *fine, double *coarse, int nvar,
                         const SmrBox &bx, cudaStream_t s);
void launch_smr_prolong(const SmrGeom &g, const double *coarse, double *fine, int nvar,
                        const SmrBox &bx, cudaStream_t s);
void launch_smr_c2p(const SmrGeom &g, const Params &p, double *cu, double *cw, int ns, double *cs,
                    double *cr, const SmrBox &bx, cudaStream_t s);
void launch_smr_bc(const SmrGeom &g, double *cw, int nh, double *cr, int ns, int face, int refl,
                   int lo, int hi, const SmrBox &bx, cudaStream_t s);
// compact != 0: coarse_flux is a message buffer, values stored as [nvar][nb][na]
void launch_smr_flux(const SmrGeom &gf, const double *fine_flux, double *coarse_flux, int nvar,
                     int dir, int fpos, int cpos, int a0, int b0, int na, int nb, cudaStream_t s,
                     int compact = 0);

// outflow (refl=0) / reflecting (refl=1) physical boundary on primitives and face fields
void launch_phys_bc(const BlkDev &b, int mhd, int face, int refl, int il, int iu, int jl,
                    int ju, int kl, int ku, cudaStream_t s);

// NewBlockTimeStep: min over active cells of dx/(|v|+c) -> atomicMin into *out_bits
// (out must be pre-set to DBL_MAX bits); result NOT yet multiplied by cfl.
void launch_new_block_dt(const BlkDev &b, const Params &p, unsigned long long *out_bits,
                         cudaStream_t s);
// Mesh::NewTimeStep on the device: state = {time, dt, tlim, cfl}; blk_min[nb]
void launch_mesh_new_dt(double *state, const unsigned long long *blk_min, int nb,
                        int advance_time, cudaStream_t s);
// HistoryOutput sums of one block added onto `out` (first != 0: start from zero); `partial` is
// device scratch of history_grid()*32 doubles
int history_grid();
void launch_history(const BlkDev &b, int mhd, int nq, int first, double *partial, double *out,
                    cudaStream_t s);
void launch_fill_u64(unsigned long long *p, int n, unsigned long long v, cudaStream_t s);

}  // namespace ab
#endif
