// ab_flux_nu.cu -- the Riemann-sweep kernels for nonuniform (geometric) mesh spacing,
// mesh/x?rat != 1: the nonuniform PLM / PPM branches of reconstruct/plm.cpp:81-105,194-214,
// 306-322 and ppm.cpp:196-207,282-300 with the weights of reconstruction.cpp:434-461.
// A separate translation unit so that it compiles in parallel with ab_kernels.cu and the
// production (uniform) kernels keep their code and register allocation untouched.
// Must be compiled with -fmad=false like the rest.
#include "ab_flux.cuh"

namespace ab {

void launch_flux_dir_nu(const BlkDev &b, const ReconGeom &g, const Params &p, int order, int dir,
                        double dt_val, const double *dt_ptr, cudaStream_t s, int nb) {
  launch_flux_dir_t<true>(b, g, p, order, dir, dt_val, dt_ptr, s, nb);
}

}  // namespace ab
