// libvbn_cuda.so -- C ABI + host launch code.  See include/vbn_cuda.h for the contract.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

#include "vbn_cuda.h"
#include "vbn_kde.cuh"
#include "vbn_reduce.cuh"
#include "vbn_launch.h"
#include "vbn_schedule.cuh"
#include "vbn_tc_layout.h"

static_assert(sizeof(VbnOp) == 128, "VbnOp must be 128 bytes");
static_assert(sizeof(VbnView) == 24, "VbnView layout");
static_assert(sizeof(VbnNoise) == 24, "VbnNoise layout");

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess)                                                             \
      return fail(VBN_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                  \
  } while (0)

// ---- launch shapes of the schedule kernel -------------------------------------------------
// The kernels themselves are instantiated in vbn_k_heavy.cu / vbn_k_light4.cu / vbn_k_light2.cu / vbn_k_tc.cu
// (separate translation units, compiled in parallel); entry points are looked up through vbn_launch.h.
struct Shape {
  int rpt, nt, heavy, min_blocks;
};

// preferred first; later entries hold fewer rows per CTA (capacity fallbacks)
const Shape kShapes[] = {
    {2, 128, 1, 3},
    {1, 128, 1, 3},
    {1, 64, 1, 1},
    {4, 256, 0, 2},  // best on LG chains (cfg2: 64% of the level-SoA HBM model)
    {4, 128, 0, 4},
    {2, 256, 0, 4},
    {4, 128, 0, 3},
    {2, 256, 0, 3},
    {2, 256, 0, 2},
    {1, 256, 0, 4},
    {1, 256, 0, 6},
    {1, 128, 0, 4},
    {1, 64, 0, 1},
};
constexpr int kNumShapes = static_cast<int>(sizeof(kShapes) / sizeof(kShapes[0]));

const void* shape_fn(const Shape& s, bool tab) {
  return s.heavy ? vbn::heavy_kernel_ptr(s.rpt, s.nt, s.min_blocks)
                 : vbn::light_kernel_ptr(s.rpt, s.nt, s.min_blocks, tab);
}

}  // namespace

struct VbnPlan {
  VbnProgramDesc desc;
  int device;
  int num_sms;
  int shape;       // index into kShapes
  int blocks_per_sm;
  size_t smem_bytes;
  int small_shape;      // smallest-CTA shape of the same family (-1: none): used when a run has too few rows to give
  int small_blocks;     // every SM a CTA of the preferred shape (Gibbs chains, CPD-handle calls, small batches)
  size_t small_smem;
  int tab;         // light schedules: the kernel variant with the table-lookup op bodies
  int tc;          // 0, or warpgroups per CTA of vbn::tc::schedule_tc_kernel<NWG, RPT>
  int tc_rpt;      // 128-row tiles per warpgroup
  int tc_nbuf;     // weight-ring depth of the tensor-core kernel
  int tc_slot_bytes;  // size of one ring slot
};

extern "C" {

int32_t vbn_cuda_abi_version(void) { return VBN_CUDA_ABI_VERSION; }

const char* vbn_cuda_last_error(void) { return g_err; }

int32_t vbn_cuda_device_count(int32_t* out_count) {
  if (!out_count) return fail(VBN_E_INVALID, "out_count is NULL");
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  *out_count = n;
  return VBN_OK;
}

int32_t vbn_plan_create(const VbnProgramDesc* desc, VbnPlan** out_plan) {
  if (!desc || !out_plan) return fail(VBN_E_INVALID, "NULL argument");
  if (desc->n_ops <= 0 || !desc->ops_dev) return fail(VBN_E_INVALID, "empty program");
  if (desc->n_slots <= 0 || desc->n_scratch < 0) return fail(VBN_E_INVALID, "bad slot counts");
  std::unique_ptr<VbnPlan> holder(new (std::nothrow) VbnPlan());  // released to the caller only on success
  VbnPlan* p = holder.get();
  if (!p) return fail(VBN_E_INVALID, "out of host memory");
  p->desc = *desc;
  CUDA_TRY(cudaGetDevice(&p->device));
  CUDA_TRY(cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device));
  int max_smem = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p->device));
  p->shape = -1;
  p->small_shape = -1;
  p->tc = 0;
  p->tc_rpt = 1;
  p->tc_nbuf = 0;
  const size_t per_row = static_cast<size_t>(desc->n_slots + desc->n_scratch) * sizeof(float);
  if (desc->tc) {
    // tensor-core kernel: 512 rows per CTA, one CTA per SM, deepest weight ring that fits
    if (!desc->tc_list_dev || desc->n_tc <= 0) {
      return fail(VBN_E_INVALID, "tc program without a tc_list");
    }
    const int nwg = desc->tc & 0xFF, rpt = (desc->tc >> 8) > 0 ? (desc->tc >> 8) : 1;
    const void* fn = vbn::tc::tc_kernel_ptr(nwg, rpt);
    if (!fn) {
      return fail(VBN_E_INVALID, "no tensor-core kernel with %d warpgroups x %d tiles", nwg, rpt);
    }
    const int threads = nwg * vbn::tc::kWgThreads;
    // ring slot = the program's largest image (weights + descriptor tail), 128-byte granular; deepest ring that fits
    int slot = desc->tc_image_bytes > 0 ? (desc->tc_image_bytes + 127) & ~127 : vbn::tc::kWbufBytes;
    if (slot > vbn::tc::kWbufBytes) return fail(VBN_E_INVALID, "tc_image_bytes %d exceeds %d", slot, vbn::tc::kWbufBytes);
    p->tc_slot_bytes = slot;
    for (int nbuf = vbn::tc::kMaxBufs; nbuf >= 2 && !p->tc; --nbuf) {
      const size_t bytes = vbn::tc::kCtrlBytes + static_cast<size_t>(nbuf) * slot + per_row * threads * rpt;
      if (bytes > static_cast<size_t>(max_smem)) continue;
      // the attribute is per FUNCTION, not per plan: always raise it to the device maximum so that plans with
      // different footprints can coexist (a later, smaller plan must not lower the limit of an earlier one)
      CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      int occ = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, bytes));
      if (occ < 1) continue;
      p->tc = nwg;
      p->tc_rpt = rpt;
      p->tc_nbuf = nbuf;
      p->blocks_per_sm = 1;
      p->smem_bytes = bytes;
    }
    if (p->tc) {
      *out_plan = holder.release();
      return VBN_OK;
    }
    // does not fit: fall through to the FFMA shapes (the ops keep their FFMA parameter blocks)
  }
  const bool tab = desc->has_tables != 0;
  p->tab = tab ? 1 : 0;
  const char* force = desc->heavy ? nullptr : std::getenv("VBN_SHAPE");  // dev knob (light schedules): force a shape
  for (int i = 0; i < kNumShapes; ++i) {
    const Shape& s = kShapes[i];
    if (s.heavy != (desc->heavy ? 1 : 0)) continue;
    if (force && std::atoi(force) != i) continue;
    if (!force && !s.heavy && desc->rows_per_thread == 2 && s.rpt > 2) continue;
    const size_t bytes = per_row * s.rpt * s.nt;
    if (bytes > static_cast<size_t>(max_smem)) continue;
    CUDA_TRY(cudaFuncSetAttribute(shape_fn(s, tab), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shape_fn(s, tab), s.nt, bytes));
    if (occ < 1) continue;
    p->shape = i;
    p->blocks_per_sm = occ;
    p->smem_bytes = bytes;
    break;
  }
  if (p->shape >= 0 && !force) {
    for (int i = kNumShapes - 1; i > p->shape; --i) {  // fewest rows per CTA first
      const Shape& s = kShapes[i];
      if (s.heavy != (desc->heavy ? 1 : 0) || s.rpt * s.nt >= kShapes[p->shape].rpt * kShapes[p->shape].nt) continue;
      const size_t bytes = per_row * s.rpt * s.nt;
      CUDA_TRY(cudaFuncSetAttribute(shape_fn(s, tab), cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
      int occ = 0;
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shape_fn(s, tab), s.nt, bytes));
      if (occ < 1) continue;
      p->small_shape = i;
      p->small_blocks = occ;
      p->small_smem = bytes;
      break;
    }
  }
  if (p->shape < 0) {
    return fail(VBN_E_CAPACITY,
                "schedule needs %d value slots + %d scratch floats per row: exceeds %d bytes of "
                "shared memory even at 64 rows per CTA",
                desc->n_slots, desc->n_scratch, max_smem);
  }
  *out_plan = holder.release();
  return VBN_OK;
}

int32_t vbn_plan_destroy(VbnPlan* plan) {
  delete plan;
  return VBN_OK;
}

int32_t vbn_run_forward_launches(const VbnPlan* plan) { return plan ? 1 : 0; }

int32_t vbn_run_forward(const VbnPlan* plan, const VbnRunDesc* run, void* stream) {
  if (!plan || !run) return fail(VBN_E_INVALID, "NULL argument");
  if (run->n_queries <= 0 || run->n_samples <= 0) return fail(VBN_E_INVALID, "empty run");
  if (!plan->tc && plan->shape < 0) return fail(VBN_E_INVALID, "plan has no launch shape");
  const Shape& s = kShapes[plan->shape < 0 ? 0 : plan->shape];
  vbn::ScheduleArgs a;
  std::memset(&a, 0, sizeof(a));
  a.ops = plan->desc.ops_dev;
  a.par_slots = plan->desc.par_slots_dev;
  a.params = plan->desc.params_dev;
  a.n_ops = plan->desc.n_ops;
  a.n_slots = plan->desc.n_slots;
  a.n_scratch = plan->desc.n_scratch;
  a.logp_as_pdf = run->logp_as_pdf;
  a.logw_accumulate = run->logw_accumulate;
  a.n_queries = run->n_queries;
  a.n_samples = run->n_samples;
  a.n_rows = run->n_queries * run->n_samples;
  a.query_offset = static_cast<uint32_t>(run->query_offset);
  a.sample_offset = static_cast<uint32_t>(run->sample_offset);
  a.key0 = static_cast<uint32_t>(run->seed);
  a.key1 = static_cast<uint32_t>(run->seed >> 32);
  a.call_offset = static_cast<uint32_t>(run->call_offset);
  vbn::fill_round_keys(a);
  a.fixed = run->fixed_dev;
  a.inputs = run->inputs_dev;
  a.stores = run->stores_dev;
  a.noise = run->noise_dev;
  a.logw = run->logw_dev;
  a.logp = run->logp_dev;
  a.error_flag = run->error_flag_dev;
  a.seg = run->seg_dev;
  a.seg_per_query = run->seg_per_query;
  a.seg_slot = run->seg_slot;
  a.seg_classes = run->seg_classes;
  if (a.seg) {
    if (a.seg_per_query < (a.n_samples + 31) / 32 + 1) return fail(VBN_E_INVALID, "seg_per_query too small");
    if (a.seg_slot >= plan->desc.n_slots || a.seg_classes < 0 || a.seg_classes > 8)
      return fail(VBN_E_INVALID, "bad seg_slot / seg_classes");
  }
  if (plan->tc) {
    const int threads = plan->tc * vbn::tc::kWgThreads;
    const int64_t rows_per_cta = static_cast<int64_t>(threads) * plan->tc_rpt;
    const int64_t n_tiles = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
    const unsigned grid = static_cast<unsigned>(n_tiles < plan->num_sms ? n_tiles : plan->num_sms);
    int nbuf = plan->tc_nbuf;
    a.tc_list = reinterpret_cast<const int2*>(plan->desc.tc_list_dev);
    a.n_tc = plan->desc.n_tc;
    int slot_bytes = plan->tc_slot_bytes;
    const char* tune_env = std::getenv("VBN_TC_TUNE");  // dev knob: bit 0 = issuing warps on different SM sub-partitions
    int tune = tune_env ? std::atoi(tune_env) : 0;
    void* targs[] = {&a, &nbuf, &slot_bytes, &tune};
    CUDA_TRY(cudaLaunchKernel(vbn::tc::tc_kernel_ptr(plan->tc, plan->tc_rpt), dim3(grid), dim3(threads), targs, plan->smem_bytes,
                              static_cast<cudaStream_t>(stream)));
    return VBN_OK;
  }
  const int64_t rows_per_cta = static_cast<int64_t>(s.rpt) * s.nt;
  const int64_t n_tiles = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
  const bool tab = plan->tab != 0;
  void* args[] = {&a};
  if (n_tiles < plan->num_sms && plan->small_shape >= 0) {
    // fewer tiles than SMs: spread the rows over more, smaller CTAs (a 4096-chain Gibbs run is 16 CTAs of the
    // preferred heavy shape, 64 of the small one)
    const Shape& q = kShapes[plan->small_shape];
    const int64_t rows_small = static_cast<int64_t>(q.rpt) * q.nt;
    const int64_t tiles_small = (a.n_rows + rows_small - 1) / rows_small;
    const int64_t resident_small = static_cast<int64_t>(plan->num_sms) * plan->small_blocks;
    const unsigned grid_small = static_cast<unsigned>(tiles_small < resident_small ? tiles_small : resident_small);
    CUDA_TRY(cudaLaunchKernel(shape_fn(q, tab), dim3(grid_small), dim3(q.nt), args, plan->small_smem,
                              static_cast<cudaStream_t>(stream)));
    return VBN_OK;
  }
  const int64_t resident = static_cast<int64_t>(plan->num_sms) * plan->blocks_per_sm;
  const unsigned grid = static_cast<unsigned>(n_tiles < resident ? n_tiles : resident);
  CUDA_TRY(cudaLaunchKernel(shape_fn(s, tab), dim3(grid), dim3(s.nt), args, plan->smem_bytes,
                            static_cast<cudaStream_t>(stream)));
  return VBN_OK;
}

int32_t vbn_lse_partials(const float* logw_dev, int64_t n_queries, int64_t n_samples,
                         int32_t n_split, float* partials_dev, void* stream) {
  if (!logw_dev || !partials_dev || n_queries <= 0 || n_samples <= 0 || n_split <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_lse_partials");
  vbn::lse_partials_kernel<<<dim3(static_cast<unsigned>(n_queries), n_split), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(logw_dev, n_samples, n_split,
                                                                  partials_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_lse_merge(const float* partials_dev, int64_t n_queries, int32_t n_split,
                      float* stats_dev, void* stream) {
  if (!partials_dev || !stats_dev || n_queries <= 0 || n_split <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_lse_merge");
  const unsigned grid = static_cast<unsigned>((n_queries + 127) / 128);
  vbn::lse_merge_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      partials_dev, n_queries, n_split, stats_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_weights_normalize(const float* logw_dev, const float* stats_dev, int64_t n_queries,
                              int64_t n_samples, int32_t normalize, float eps, float* w_dev,
                              float* ess_dev, void* stream) {
  if (!logw_dev || !stats_dev || !w_dev || n_queries <= 0 || n_samples <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_weights_normalize");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t total = n_queries * n_samples;
  int64_t want = (total + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = static_cast<int64_t>(sms) * 8;
  const unsigned grid = static_cast<unsigned>(want < cap ? (want > 0 ? want : 1) : cap);
  vbn::weights_normalize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logw_dev, stats_dev, n_queries, n_samples, normalize, eps, w_dev, ess_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_segment_merge(const float* records_dev, int64_t n_queries, int64_t n_samples, int32_t n_per_query,
                          float ess_threshold, float* merged_dev, float* stats_dev, int32_t* flag_dev, void* stream) {
  if (!records_dev || n_queries <= 0 || n_samples < 0 || n_per_query <= 0 || (!merged_dev && !stats_dev))
    return fail(VBN_E_INVALID, "bad argument to vbn_segment_merge");
  vbn::segment_merge_kernel<<<static_cast<unsigned>(n_queries), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      records_dev, n_samples, n_per_query, ess_threshold, merged_dev, stats_dev, flag_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_ess_below(const float* stats_dev, int64_t n_queries, float threshold,
                      int32_t* flag_dev, void* stream) {
  if (!stats_dev || !flag_dev || n_queries <= 0) return fail(VBN_E_INVALID, "bad argument");
  const unsigned grid = static_cast<unsigned>((n_queries + 255) / 256);
  vbn::ess_below_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(stats_dev, n_queries,
                                                                           threshold, flag_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_row_cdf(const float* w_dev, int64_t n_queries, int64_t n_samples, float* cdf_dev, void* stream) {
  if (!w_dev || !cdf_dev || n_queries <= 0 || n_samples <= 0) return fail(VBN_E_INVALID, "bad argument to vbn_row_cdf");
  vbn::row_cdf_kernel<<<static_cast<unsigned>(n_queries), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_dev, n_samples, cdf_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_resample_indices(const float* cdf_dev, int64_t n_queries, int64_t n_samples, uint64_t seed,
                             uint64_t call_offset, int64_t query_offset, int64_t sample_offset,
                             int32_t* idx_dev, void* stream) {
  if (!cdf_dev || !idx_dev || n_queries <= 0 || n_samples <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_resample_indices");
  const int64_t rows = n_queries * n_samples;
  vbn::resample_index_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cdf_dev, n_queries, n_samples, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32),
      static_cast<uint32_t>(call_offset), static_cast<uint32_t>(query_offset), static_cast<uint32_t>(sample_offset),
      idx_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_gather_rows(const float* src_dev, float* dst_dev, const int32_t* idx_dev, int32_t n_cols,
                        int64_t n_queries, int64_t n_samples, void* stream) {
  if (!src_dev || !dst_dev || !idx_dev || n_cols <= 0 || n_queries <= 0 || n_samples <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_gather_rows");
  const int64_t rows = n_queries * n_samples;
  vbn::gather_rows_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src_dev, dst_dev, idx_dev, n_cols, n_queries, n_samples);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_weighted_sum(const float* w_dev, const float* x_dev, int64_t n_queries, int64_t n_samples,
                         int32_t width, float* out_dev, void* stream) {
  if (!w_dev || !x_dev || !out_dev || n_queries <= 0 || n_samples <= 0 || width <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_weighted_sum");
  vbn::weighted_sum_kernel<<<static_cast<unsigned>(n_queries), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_dev, x_dev, n_samples, width, out_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_weighted_histogram(const float* samples_dev, const float* w_dev, int64_t n_queries, int64_t n_samples,
                               int64_t sample_stride, int32_t n_classes, float* probs_dev, void* stream) {
  if (!samples_dev || !w_dev || !probs_dev || n_queries <= 0 || n_samples <= 0 || sample_stride <= 0 ||
      n_classes <= 0 || n_classes > vbn::kHistMaxClasses)
    return fail(VBN_E_INVALID, "bad argument to vbn_weighted_histogram (1 <= n_classes <= %d)", vbn::kHistMaxClasses);
  vbn::weighted_histogram_kernel<<<static_cast<unsigned>(n_queries), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      samples_dev, w_dev, n_samples, sample_stride, n_classes, probs_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_gaussian_mixture_grid(const float* w_dev, const float* loc_scale_dev, int64_t n_queries,
                                  int64_t n_particles, int64_t n_out, float stddevs, float min_scale,
                                  float* pdf_dev, float* grid_dev, void* stream) {
  if (!w_dev || !loc_scale_dev || !pdf_dev || !grid_dev || n_queries <= 0 || n_particles <= 0 || n_out <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_gaussian_mixture_grid");
  vbn::gaussian_mixture_grid_kernel<<<static_cast<unsigned>(n_queries), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w_dev, loc_scale_dev, n_particles, n_out, stddevs, min_scale, pdf_dev, grid_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_gaussian_grid(const float* loc_scale_dev, int64_t n_queries, int64_t n_samples, float stddevs,
                          float min_scale, float* pdf_dev, float* samples_dev, void* stream) {
  if (!loc_scale_dev || !pdf_dev || !samples_dev || n_queries <= 0 || n_samples <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_gaussian_grid");
  const int64_t total = n_queries * n_samples;
  const unsigned grid = static_cast<unsigned>(total / 256 + 1 < 148 * 8 ? total / 256 + 1 : 148 * 8);
  vbn::gaussian_grid_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      loc_scale_dev, n_queries, n_samples, stddevs, min_scale, pdf_dev, samples_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_posterior_stats(const float* pdf_dev, const float* samples_dev, int64_t n_queries,
                            int64_t n_samples, int32_t dim, int32_t n_split, float eps,
                            float* partials_dev, float* stats_dev, void* stream) {
  if (!pdf_dev || !samples_dev || !partials_dev || !stats_dev || n_queries <= 0 || n_samples <= 0 ||
      dim <= 0 || dim > vbn::kStatsMaxDim || n_split <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_posterior_stats (dim must be 1..%d)", vbn::kStatsMaxDim);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(static_cast<unsigned>(n_queries), n_split);
  const unsigned mg = static_cast<unsigned>((n_queries + 127) / 128);
  vbn::posterior_partials_kernel<0><<<grid, 256, 0, st>>>(pdf_dev, samples_dev, n_samples, dim, n_split,
                                                          stats_dev, eps, partials_dev);
  vbn::posterior_merge_kernel<0><<<mg, 128, 0, st>>>(partials_dev, pdf_dev, samples_dev, n_queries, n_samples,
                                                     dim, n_split, eps, stats_dev);
  vbn::posterior_partials_kernel<1><<<grid, 256, 0, st>>>(pdf_dev, samples_dev, n_samples, dim, n_split,
                                                          stats_dev, eps, partials_dev);
  vbn::posterior_merge_kernel<1><<<mg, 128, 0, st>>>(partials_dev, pdf_dev, samples_dev, n_queries, n_samples,
                                                     dim, n_split, eps, stats_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_kde_log_prob(const float* train_p_dev, const float* train_y_dev, int64_t n_points,
                         int32_t dp, int32_t dx, const float* query_p_dev,
                         const float* query_x_dev, int64_t n_rows, float bandwidth,
                         float parent_bandwidth, float min_scale, float* out_dev, void* stream) {
  if (!train_y_dev || !query_x_dev || !out_dev || n_points <= 0 || n_rows <= 0 || dx <= 0 || dp < 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_kde_log_prob");
  if (dp > 0 && (!train_p_dev || !query_p_dev))
    return fail(VBN_E_INVALID, "parent arrays required when dp > 0");
  return vbn::launch_kde_log_prob(train_p_dev, train_y_dev, n_points, dp, dx, query_p_dev,
                                  query_x_dev, n_rows, bandwidth, parent_bandwidth, min_scale,
                                  out_dev, static_cast<cudaStream_t>(stream)) == cudaSuccess
             ? VBN_OK
             : fail(VBN_E_CUDA, "kde launch failed: %s", cudaGetErrorString(cudaGetLastError()));
}

int32_t vbn_kde_tc_workspace_bytes(int64_t n_points, int32_t dp, int32_t dx, int64_t* out_bytes) {
  if (!out_bytes || n_points <= 0 || dx <= 0 || dp < 0) return fail(VBN_E_INVALID, "bad argument");
  *out_bytes = vbn::kde_tc_supported(dp, dx) ? static_cast<int64_t>(vbn::kde_tc_workspace_bytes(n_points, dp, dx)) : 0;
  return VBN_OK;
}

int32_t vbn_kde_log_prob_tc(const float* train_p_dev, const float* train_y_dev, int64_t n_points, int32_t dp,
                            int32_t dx, const float* query_p_dev, const float* query_x_dev, int64_t n_rows,
                            float bandwidth, float parent_bandwidth, float min_scale, const float* center_dev,
                            void* workspace_dev, float* out_dev, void* stream) {
  if (!train_y_dev || !query_x_dev || !out_dev || !workspace_dev || n_points <= 0 || n_rows <= 0 || dx <= 0 || dp < 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_kde_log_prob_tc");
  if (dp > 0 && (!train_p_dev || !query_p_dev)) return fail(VBN_E_INVALID, "parent arrays required when dp > 0");
  if (!vbn::kde_tc_supported(dp, dx)) return fail(VBN_E_CAPACITY, "dims (%d, %d) exceed the tensor-core KDE kernel", dp, dx);
  const cudaError_t e = vbn::launch_kde_log_prob_tc(train_p_dev, train_y_dev, n_points, dp, dx, query_p_dev, query_x_dev,
                                                    n_rows, bandwidth, parent_bandwidth, min_scale, center_dev,
                                                    static_cast<float*>(workspace_dev), out_dev,
                                                    static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? VBN_OK : fail(VBN_E_CUDA, "kde tensor-core launch failed: %s", cudaGetErrorString(e));
}

int32_t vbn_philox_fill(const uint32_t* ctr_dev, int64_t n, uint32_t key0, uint32_t key1,
                        uint32_t* out_dev, void* stream) {
  if (!ctr_dev || !out_dev || n <= 0) return fail(VBN_E_INVALID, "bad argument");
  const unsigned grid = static_cast<unsigned>((n + 255) / 256);
  vbn::philox_fill_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ctr_dev, n, key0, key1,
                                                                             out_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_stream_draws(uint64_t seed, uint64_t call_offset, int32_t kind, int32_t shared, int32_t block_lo,
                         int32_t n_blocks, int64_t n_queries, int64_t n_samples, int64_t query_offset,
                         int64_t sample_offset, float* out_dev, void* stream) {
  if (!out_dev || kind < 0 || kind > 2 || block_lo < 0 || n_blocks <= 0 || n_queries <= 0 || n_samples <= 0)
    return fail(VBN_E_INVALID, "bad argument to vbn_stream_draws");
  vbn::ScheduleArgs keys;
  keys.key0 = static_cast<uint32_t>(seed);
  keys.key1 = static_cast<uint32_t>(seed >> 32);
  vbn::fill_round_keys(keys);
  vbn::StreamDrawArgs a;
  std::memcpy(a.rk, keys.rk, sizeof(a.rk));
  a.call_offset = static_cast<uint32_t>(call_offset);
  a.query_offset = static_cast<uint32_t>(query_offset);
  a.sample_offset = static_cast<uint32_t>(sample_offset);
  a.kind = kind;
  a.shared = shared ? 1 : 0;
  a.block_lo = block_lo;
  a.n_blocks = n_blocks;
  a.n_queries = n_queries;
  a.n_samples = n_samples;
  const int64_t total = (shared ? 1 : n_queries) * n_samples * n_blocks;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  vbn::stream_draws_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, out_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_fma_peak(int32_t mode, int32_t iters, int32_t n_blocks, float* scratch_dev, void* stream) {
  if (!scratch_dev || iters <= 0 || n_blocks <= 0) return fail(VBN_E_INVALID, "bad argument");
  vbn::fma_peak_kernel<<<n_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(mode, iters, scratch_dev);
  CUDA_TRY(cudaGetLastError());
  return VBN_OK;
}

int32_t vbn_tf32_peak(int32_t iters, int32_t n_blocks, float* scratch_dev, void* stream) {
  if (!scratch_dev || iters <= 0 || n_blocks <= 0) return fail(VBN_E_INVALID, "bad argument");
  void* args[] = {&iters, &scratch_dev};
  CUDA_TRY(cudaLaunchKernel(vbn::tc::tf32_peak_kernel_ptr(), dim3(n_blocks), dim3(128), args, 0,
                            static_cast<cudaStream_t>(stream)));
  return VBN_OK;
}

}  // extern "C"
