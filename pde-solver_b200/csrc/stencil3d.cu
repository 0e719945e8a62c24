// Specialised streaming kernels for the 3D hot paths (heat / mass 15-point, elasticity 3x3-block).
// launch_stencil_fast() claims a launch when a specialised kernel applies; otherwise the generic
// table kernel in kernels.cu runs.
#include "device.cuh"

int launch_stencil_fast(pde_ctx* c, const Grid& g, const BcDev& bc, const OpDev& op, const StencilArgs& a,
                        bool* handled) {
  (void)c; (void)g; (void)bc; (void)op; (void)a;
  *handled = false;
  return 0;
}
