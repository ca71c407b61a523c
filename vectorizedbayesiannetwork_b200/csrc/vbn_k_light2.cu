// FP32-pipe schedule kernels for linear-Gaussian / table-only programs, 1-2 rows per thread.
#include "vbn_launch.h"
#include "vbn_schedule.cuh"

namespace vbn {
const void* light2_kernel_ptr(int rpt, int nt, int min_blocks, bool tab) {
#define VBN_SHAPE(R, N, M)                                                                              \
  if (rpt == R && nt == N && min_blocks == M)                                                           \
    return tab ? reinterpret_cast<const void*>(&schedule_kernel<R, N, false, M, true>)                  \
               : reinterpret_cast<const void*>(&schedule_kernel<R, N, false, M, false>)
  VBN_SHAPE(2, 256, 4);
  VBN_SHAPE(2, 256, 3);
  VBN_SHAPE(2, 256, 2);
  VBN_SHAPE(1, 256, 4);
  VBN_SHAPE(1, 256, 6);
  VBN_SHAPE(1, 128, 4);
  VBN_SHAPE(1, 64, 1);
#undef VBN_SHAPE
  return nullptr;
}
}  // namespace vbn
