// annulus_build.cu -- device builder for init_annulus (src/GridAnnulus.jl:57-70).  [placeholder until the
// closed-form kernels land: the entry point reports RT_ERR_UNSUPPORTED instead of falling back to the host]
#include "mesh2d.cuh"

int annulus_build_device(rt_mesh* h, i64 ntheta, i64 nr, double spacing) {
  (void)h; (void)ntheta; (void)nr; (void)spacing;
  rt_set_error("rt_annulus_build: device builder not implemented yet");
  return RT_ERR_UNSUPPORTED;
}
