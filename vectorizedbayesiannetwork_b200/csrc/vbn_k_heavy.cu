// FP32-pipe schedule kernels for programs with MLP / KDE / RFF / Gibbs ops (HEAVY = true).
#include "vbn_launch.h"
#include "vbn_schedule.cuh"

namespace vbn {
const void* heavy_kernel_ptr(int rpt, int nt, int min_blocks) {
#define VBN_SHAPE(R, N, M) \
  if (rpt == R && nt == N && min_blocks == M) return reinterpret_cast<const void*>(&schedule_kernel<R, N, true, M>)
  VBN_SHAPE(2, 128, 3);
  VBN_SHAPE(1, 128, 3);
  VBN_SHAPE(1, 64, 1);
#undef VBN_SHAPE
  return nullptr;
}
}  // namespace vbn
