// Internal (not part of the C ABI): the schedule kernels are instantiated in their own translation
// units so that nvcc compiles them in parallel; vbn_cuda.cu looks the entry points up through these.
#pragma once

namespace vbn {

// &schedule_kernel<rpt, nt, heavy, min_blocks>, or nullptr when that shape is not instantiated
const void* heavy_kernel_ptr(int rpt, int nt, int min_blocks);
const void* light_kernel_ptr(int rpt, int nt, int min_blocks);   // dispatches to the two light TUs
const void* light4_kernel_ptr(int rpt, int nt, int min_blocks);  // 4 rows per thread
const void* light2_kernel_ptr(int rpt, int nt, int min_blocks);  // 1-2 rows per thread

namespace tc {
// tensor-core kernel variants: nwg warpgroups per CTA, rpt 128-row tiles per warpgroup
const void* tc_kernel_ptr(int nwg, int rpt);
const void* tf32_peak_kernel_ptr();
}  // namespace tc

}  // namespace vbn
