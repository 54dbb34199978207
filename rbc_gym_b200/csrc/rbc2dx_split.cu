// rbc2dx_split.cu — the pressure-split instantiations of the cluster kernel (pressure channels requested:
// rbc_sim2D_api.jl:114-115 returns pHY' and pNHS with the state).  A translation unit of its own so that it
// compiles next to rbc2dx_lib.cu.
#include "rbc2dx_kernel.cuh"

namespace rbc2dx_api {

int create_split(Plan* p)
{
    const int nx = p->nx, nz = p->nz, precision = p->precision;
    if (nx == 192 && nz == 128) {
        if (precision == 32) return create_impl<Grid<192, 128, 4, 2>, float, false, true>(p, "rbc2dx_env_kernel<192x128,cl4,f32,split>");
        return create_impl<Grid<192, 128, 8, 2>, double, true, true>(p, "rbc2dx_env_kernel<192x128,cl8,f64,split>");
    }
    if (nx == 128 && nz == 64) {
        if (precision == 32) return create_impl<Grid<128, 64, 2, 2, 4>, float, false, true>(p, "rbc2dx_env_kernel<128x64,cl2,f32,split>");
        return create_impl<Grid<128, 64, 2, 2, 4>, double, true, true>(p, "rbc2dx_env_kernel<128x64,cl2,f64,split>");
    }
    if (nx == 96 && nz == 64) {
        if (precision == 32) return create_impl<Grid<96, 64, 1, 4>, float, false, true>(p, "rbc2dx_env_kernel<96x64,cl1,f32,split>");
        return create_impl<Grid<96, 64, 2, 4>, double, true, true>(p, "rbc2dx_env_kernel<96x64,cl2,f64,split>");
    }
    return rbc_fail("rbc2dx: no pressure-split kernel registered for this grid");
}

}  // namespace rbc2dx_api
