// rbc2dx_more.cuh — further registered 2D grids of the cluster kernel (state_shape is a free keyword of the reference environment,
// src/rbc_gym/envs/rbc2D.py:43-60): the decompositions below cover widths 64, 96, 128, 192 against heights 32, 64, 128.  Instantiated in
// rbc2dx_more.cu (plain) and rbc2dx_more_split.cu (pressure channels); fp32 keeps both state buffers on-chip, fp64 keeps the predicted
// state in global memory (and splits the grid over more CTAs where the slab would not fit).
#pragma once
#include "rbc2dx_kernel.cuh"

namespace rbc2dx_api {

//                     NX   NZ  CL  strips  FFT split
using G64x32_f  = Grid< 64,  32, 1, 4, 4>;
using G64x32_d  = Grid< 64,  32, 1, 4, 4>;
using G64x64_f  = Grid< 64,  64, 1, 4, 4>;
using G64x64_d  = Grid< 64,  64, 2, 4, 4>;
using G96x32_f  = Grid< 96,  32, 1, 4, 6>;
using G96x32_d  = Grid< 96,  32, 1, 4, 6>;
using G128x32_f = Grid<128,  32, 1, 2, 4>;
using G128x32_d = Grid<128,  32, 2, 2, 4>;
using G128x128_f = Grid<128, 128, 4, 2, 4>;
using G128x128_d = Grid<128, 128, 8, 2, 4>;
using G192x64_f = Grid<192,  64, 2, 2, 6>;
using G192x64_d = Grid<192,  64, 4, 2, 6>;
using G96x128_f = Grid< 96, 128, 4, 4, 6>;
using G96x128_d = Grid< 96, 128, 4, 4, 6>;

template <bool SPLIT>
static int create_more_impl(Plan* p)
{
    const int nx = p->nx, nz = p->nz;
    const bool f = p->precision == 32;
#define RBX_MORE(NXv, NZv, GF, GD, TAG)                                                                            \
    if (nx == NXv && nz == NZv)                                                                                    \
        return f ? create_impl<GF, float, false, SPLIT>(p, SPLIT ? "rbc2dx_env_kernel<" TAG ",f32,split>" : "rbc2dx_env_kernel<" TAG ",f32>") \
                 : create_impl<GD, double, true, SPLIT>(p, SPLIT ? "rbc2dx_env_kernel<" TAG ",f64,split>" : "rbc2dx_env_kernel<" TAG ",f64>");
    RBX_MORE(64, 32, G64x32_f, G64x32_d, "64x32")
    RBX_MORE(64, 64, G64x64_f, G64x64_d, "64x64")
    RBX_MORE(96, 32, G96x32_f, G96x32_d, "96x32")
    RBX_MORE(128, 32, G128x32_f, G128x32_d, "128x32")
    RBX_MORE(128, 128, G128x128_f, G128x128_d, "128x128")
    RBX_MORE(192, 64, G192x64_f, G192x64_d, "192x64")
    RBX_MORE(96, 128, G96x128_f, G96x128_d, "96x128")
#undef RBX_MORE
    return rbc_fail("rbc2dx: no kernel registered for this grid");
}

}  // namespace rbc2dx_api
