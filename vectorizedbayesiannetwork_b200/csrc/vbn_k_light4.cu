// FP32-pipe schedule kernels for linear-Gaussian / table-only programs, 4 rows per thread.
#include "vbn_launch.h"
#include "vbn_schedule.cuh"

namespace vbn {
const void* light4_kernel_ptr(int rpt, int nt, int min_blocks, bool tab) {
#define VBN_SHAPE(R, N, M)                                                                              \
  if (rpt == R && nt == N && min_blocks == M)                                                           \
    return tab ? reinterpret_cast<const void*>(&schedule_kernel<R, N, false, M, true>)                  \
               : reinterpret_cast<const void*>(&schedule_kernel<R, N, false, M, false>)
  VBN_SHAPE(4, 256, 2);
  VBN_SHAPE(4, 128, 4);
  VBN_SHAPE(4, 128, 3);
#undef VBN_SHAPE
  return nullptr;
}
const void* light_kernel_ptr(int rpt, int nt, int min_blocks, bool tab) {
  return rpt == 4 ? light4_kernel_ptr(rpt, nt, min_blocks, tab) : light2_kernel_ptr(rpt, nt, min_blocks, tab);
}
}  // namespace vbn
