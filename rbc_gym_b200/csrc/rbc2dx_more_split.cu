// rbc2dx_more_split.cu — pressure-split instantiations of the further registered 2D grids (rbc2dx_more.cuh).
#include "rbc2dx_more.cuh"

namespace rbc2dx_api {

int create_more_split(Plan* p) { return create_more_impl<true>(p); }

}  // namespace rbc2dx_api
