// rbc2dx_more.cu — plain-mode instantiations of the further registered 2D grids (rbc2dx_more.cuh).
#include "rbc2dx_more.cuh"

namespace rbc2dx_api {

int more_supported(int nx, int nz)
{
    return (nx == 64 && (nz == 32 || nz == 64)) || (nx == 96 && (nz == 32 || nz == 128)) || (nx == 128 && (nz == 32 || nz == 128)) ||
           (nx == 192 && nz == 64);
}
int create_more(Plan* p) { return create_more_impl<false>(p); }

}  // namespace rbc2dx_api
