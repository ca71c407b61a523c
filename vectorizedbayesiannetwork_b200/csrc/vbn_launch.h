// Internal (not part of the C ABI): the schedule kernels are instantiated in their own translation
// units so that nvcc compiles them in parallel; vbn_cuda.cu looks the entry points up through these.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vbn {

// &schedule_kernel<rpt, nt, heavy, min_blocks>, or nullptr when that shape is not instantiated
const void* heavy_kernel_ptr(int rpt, int nt, int min_blocks);
// tab: the schedule holds table-lookup ops (the kernels without their bodies serve linear-Gaussian-only schedules)
const void* light_kernel_ptr(int rpt, int nt, int min_blocks, bool tab);   // dispatches to the two light TUs
const void* light4_kernel_ptr(int rpt, int nt, int min_blocks, bool tab);  // 4 rows per thread
const void* light2_kernel_ptr(int rpt, int nt, int min_blocks, bool tab);  // 1-2 rows per thread

// tcgen05 KDE kernel (vbn_k_kde_tc.cu): bytes of the packed-point workspace, whether (dp, dx) fits, the launch
size_t kde_tc_workspace_bytes(int64_t n_points, int dp, int dx);
bool kde_tc_supported(int dp, int dx);
cudaError_t launch_kde_log_prob_tc(const float* tp, const float* ty, int64_t n_points, int dp, int dx, const float* qp,
                                   const float* qx, int64_t n_rows, float bandwidth, float parent_bandwidth,
                                   float min_scale, const float* center, float* workspace, float* out,
                                   cudaStream_t stream);

namespace tc {
// tensor-core kernel variants: nwg warpgroups per CTA, rpt 128-row tiles per warpgroup
const void* tc_kernel_ptr(int nwg, int rpt);
const void* tf32_peak_kernel_ptr();
}  // namespace tc

}  // namespace vbn
