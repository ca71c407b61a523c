// Weight reduction kernels: per-query softmax over the S log-weights, ESS, fallback flag.
// Replaces torch.softmax(log_weights, dim=1); ess = 1/(w**2).sum(1); torch.any(ess < thr)
// (vbn/inference/importance_sampling.py:82-86, likelihood_weighting.py:75-80).
// HBM-bound: logw is read twice (4+4 B/row) and w written once (4 B/row).
#pragma once
#include "vbn_device.cuh"

namespace vbn {

// (m, l, q) triple: m = max x, l = sum exp(x-m), q = sum exp(2(x-m)).  Merge is associative.
struct Mlq {
  float m, l, q;
};

__device__ __forceinline__ Mlq mlq_merge(Mlq a, Mlq b) {
  if (b.m == -CUDART_INF_F) return a;
  if (a.m == -CUDART_INF_F) return b;
  const float m = fmaxf(a.m, b.m);
  const float ea = __expf(a.m - m), eb = __expf(b.m - m);
  return Mlq{m, a.l * ea + b.l * eb, a.q * ea * ea + b.q * eb * eb};
}

__device__ __forceinline__ void mlq_push(Mlq& s, float x) {
  if (x == -CUDART_INF_F) return;  // weight 0 (torch.softmax gives such rows 0); exp(-inf - -inf) would be NaN
  if (x > s.m) {
    const float e = __expf(s.m - x);  // exp(-inf)=0 on the first element
    s.l = s.l * e + 1.0f;
    s.q = s.q * e * e + 1.0f;
    s.m = x;
  } else {
    const float e = __expf(x - s.m);
    s.l += e;
    s.q += e * e;
  }
}

__device__ __forceinline__ Mlq mlq_warp(Mlq s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Mlq t;
    t.m = __shfl_xor_sync(0xffffffffu, s.m, o);
    t.l = __shfl_xor_sync(0xffffffffu, s.l, o);
    t.q = __shfl_xor_sync(0xffffffffu, s.q, o);
    s = mlq_merge(s, t);
  }
  return s;
}

// grid (B, n_split), 256 threads: block (b, sp) reduces samples [sp*chunk, (sp+1)*chunk) of query b
__global__ void __launch_bounds__(256) lse_partials_kernel(const float* __restrict__ logw,
                                                            int64_t n_samples, int n_split,
                                                            float* __restrict__ partials) {
  const int64_t b = blockIdx.x;
  const int sp = blockIdx.y;
  const int64_t chunk = (n_samples + n_split - 1) / n_split;
  const int64_t lo = sp * chunk;
  const int64_t hi = lo + chunk < n_samples ? lo + chunk : n_samples;
  const float* x = logw + b * n_samples;
  Mlq s{-CUDART_INF_F, 0.0f, 0.0f};
  // 128-bit loads when the query row is 16-byte aligned
  const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((lo & 3) == 0);
  if (vec) {
    const int64_t n4 = (hi - lo) >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x + lo);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(x4 + i);
      mlq_push(s, v.x);
      mlq_push(s, v.y);
      mlq_push(s, v.z);
      mlq_push(s, v.w);
    }
    for (int64_t i = lo + (n4 << 2) + threadIdx.x; i < hi; i += blockDim.x) mlq_push(s, __ldg(x + i));
  } else {
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) mlq_push(s, __ldg(x + i));
  }
  s = mlq_warp(s);
  __shared__ Mlq sh[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < 8 ? sh[lane] : Mlq{-CUDART_INF_F, 0.0f, 0.0f};
    s = mlq_warp(s);
    if (lane == 0) {
      float* o = partials + (b * n_split + sp) * 3;
      o[0] = s.m;
      o[1] = s.l;
      o[2] = s.q;
    }
  }
}

// one thread per query: merge n_split partial triples.  Also used for the cross-GPU merge
// (partials = all-gathered per-rank stats).
__global__ void lse_merge_kernel(const float* __restrict__ partials, int64_t n_queries, int n_split,
                                 float* __restrict__ stats) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= n_queries) return;
  Mlq s{-CUDART_INF_F, 0.0f, 0.0f};
  for (int i = 0; i < n_split; ++i) {
    const float* p = partials + (b * n_split + i) * 3;
    s = mlq_merge(s, Mlq{p[0], p[1], p[2]});
  }
  stats[b * 3 + 0] = s.m;
  stats[b * 3 + 1] = s.l;
  stats[b * 3 + 2] = s.q;
}

// ---------------------------------------------------------------------------------------
// Merge of the per-warp records the schedule kernels emit (emit_segments, vbn_schedule.cuh): one 128-thread block per
// query folds its records into one -- (m, l, q) for the softmax / ESS, weighted mean and M2 = sum e (x - mean)^2 by
// the pairwise (Chan) update so no E[x^2] - E[x]^2 cancellation, class sums -- and optionally raises the batch-global
// IS -> LW fallback flag (importance_sampling.py:85-88: any ESS below the threshold).
//   records [B][P][16]; n_samples > 0: query b owns records 0 .. ((b+1) S - 1)/32 - b S / 32 (the kernel's geometry);
//   n_samples == 0: all P records are valid (cross-rank merge of all-gathered merged records).
//   merged [B][16] = {m, l, q, mean, 0, M2, rows, ess, class sums[8]} (a valid record itself: x0 = mean, sx = 0);
//   stats [B][3] = {m, l, q} for vbn_weights_normalize (either output may be NULL).
// ---------------------------------------------------------------------------------------
struct SegAcc {
  float m, l, q, mean, m2, n;
  float h[8];
};
__device__ __forceinline__ void seg_merge(SegAcc& a, const SegAcc& b) {
  if (b.n == 0.0f) return;
  if (a.n == 0.0f) { a = b; return; }
  const float m = fmaxf(a.m, b.m);
  const float sa = a.m > -CUDART_INF_F ? __expf(a.m - m) : 0.0f, sb = b.m > -CUDART_INF_F ? __expf(b.m - m) : 0.0f;
  const float la = a.l * sa, lb = b.l * sb, l = la + lb;
  const float d = b.mean - a.mean;
  const float fb = l > 0.0f ? __fdiv_rn(lb, l) : 0.0f;
  a.mean = fmaf(d, fb, a.mean);
  a.m2 = a.m2 * sa + b.m2 * sb + d * d * la * fb;
  a.q = a.q * sa * sa + b.q * sb * sb;
  a.l = l;
  a.m = m;
  a.n += b.n;
#pragma unroll
  for (int k = 0; k < 8; ++k) a.h[k] = a.h[k] * sa + b.h[k] * sb;
}
__device__ __forceinline__ SegAcc seg_shfl(const SegAcc& a, int o) {
  SegAcc r;
  r.m = __shfl_xor_sync(0xffffffffu, a.m, o);
  r.l = __shfl_xor_sync(0xffffffffu, a.l, o);
  r.q = __shfl_xor_sync(0xffffffffu, a.q, o);
  r.mean = __shfl_xor_sync(0xffffffffu, a.mean, o);
  r.m2 = __shfl_xor_sync(0xffffffffu, a.m2, o);
  r.n = __shfl_xor_sync(0xffffffffu, a.n, o);
#pragma unroll
  for (int k = 0; k < 8; ++k) r.h[k] = __shfl_xor_sync(0xffffffffu, a.h[k], o);
  return r;
}
__global__ void __launch_bounds__(128) segment_merge_kernel(const float* __restrict__ records, int64_t n_samples,
                                                            int n_per_query, float ess_threshold,
                                                            float* __restrict__ merged, float* __restrict__ stats,
                                                            int32_t* __restrict__ flag) {
  const int64_t b = blockIdx.x;
  int count = n_per_query;
  if (n_samples > 0) count = static_cast<int>((((b + 1) * n_samples - 1) >> 5) - ((b * n_samples) >> 5) + 1);
  SegAcc acc;
  acc.m = -CUDART_INF_F; acc.l = acc.q = acc.mean = acc.m2 = acc.n = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) acc.h[k] = 0.0f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float4* r = reinterpret_cast<const float4*>(records + (b * n_per_query + i) * 16);
    const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
    SegAcc s;
    s.m = r0.x; s.l = r0.y; s.q = r0.z; s.n = r1.z;
    // record -> (mean, M2): mean = x0 + sx / l, M2 = sxx - sx^2 / l
    const float inv = s.l > 0.0f ? __fdiv_rn(1.0f, s.l) : 0.0f;
    s.mean = fmaf(r1.x, inv, r0.w);
    s.m2 = fmaxf(r1.y - r1.x * r1.x * inv, 0.0f);
    s.h[0] = r2.x; s.h[1] = r2.y; s.h[2] = r2.z; s.h[3] = r2.w; s.h[4] = r3.x; s.h[5] = r3.y; s.h[6] = r3.z; s.h[7] = r3.w;
    seg_merge(acc, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const SegAcc other = seg_shfl(acc, o);
    seg_merge(acc, other);
  }
  __shared__ SegAcc sh[4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) seg_merge(acc, sh[w]);
    // a query whose log-weights are all -inf has l = q = 0: ESS = 0/0 = NaN and NaN weights, like torch.softmax of
    // such a row (and NaN < threshold is false: no fallback, importance_sampling.py:85)
    const float ess = __fdiv_rn(acc.l * acc.l, acc.q);
    if (merged) {
      float4* o = reinterpret_cast<float4*>(merged + b * 16);
      o[0] = make_float4(acc.m, acc.l, acc.q, acc.mean);
      o[1] = make_float4(0.0f, acc.m2, acc.n, ess);
      o[2] = make_float4(acc.h[0], acc.h[1], acc.h[2], acc.h[3]);
      o[3] = make_float4(acc.h[4], acc.h[5], acc.h[6], acc.h[7]);
    }
    if (stats) {
      stats[b * 3 + 0] = acc.m;
      stats[b * 3 + 1] = acc.l;
      stats[b * 3 + 2] = acc.q;
    }
    if (flag && ess_threshold > 0.0f && ess < ess_threshold) atomicOr(flag, 1);
  }
}

__global__ void __launch_bounds__(256) weights_normalize_kernel(
    const float* __restrict__ logw, const float* __restrict__ stats, int64_t n_queries,
    int64_t n_samples, int normalize, float eps, float* __restrict__ w, float* __restrict__ ess) {
  const int64_t total = n_queries * n_samples;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < total; r += stride) {
    const int64_t b = r / n_samples;
    const float m = __ldg(stats + b * 3), l = __ldg(stats + b * 3 + 1);
    const float e = expf(__ldg(logw + r) - m);
    w[r] = normalize ? __fdiv_rn(e, l) : fmaxf(e, eps);
    if (ess != nullptr && r == b * n_samples) {
      const float q = __ldg(stats + b * 3 + 2);
      ess[b] = __fdiv_rn(l * l, q);  // 1/sum(w^2) with w = e/l
    }
  }
}

__global__ void ess_below_kernel(const float* __restrict__ stats, int64_t n_queries, float threshold,
                                 int32_t* __restrict__ flag) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= n_queries) return;
  const float l = stats[b * 3 + 1], q = stats[b * 3 + 2];
  const float ess = __fdiv_rn(l * l, q);
  if (ess < threshold) atomicOr(flag, 1);
}

// ---------------------------------------------------------------------------------------
// Posterior summary (VBN._posterior_stats, vbn/vbn.py:483-504): per query, weights = sanitised pdf
// normalised over S (uniform when the sum is <= eps), mean / std of the weighted samples per
// output dim, ESS = 1 / sum w^2.  Two passes like the reference (mean first, then the centred
// second moment) so fp32 keeps its digits.  HBM-bound: pdf and samples are read twice.
//   pass 0 partials per (query, split): {sum w, sum w^2, sum w x_d ...}
//   pass 1 partials per (query, split): {sum w (x_d - mean_d)^2 ...}
// ---------------------------------------------------------------------------------------
constexpr int kStatsMaxDim = 8;

__device__ __forceinline__ float sanitize_weight(float w) {
  // nan_to_num(nan=0, posinf=0, neginf=0).clamp_min(0)
  return (w == w && fabsf(w) != CUDART_INF_F && w > 0.0f) ? w : 0.0f;
}

__device__ __forceinline__ float block_sum256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = lane < 8 ? sh[lane] : 0.0f;
  if (warp == 0) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}

// grid (B, n_split), 256 threads.  PASS 0: sums; PASS 1: centred second moments (needs stats).
template <int PASS>
__global__ void __launch_bounds__(256) posterior_partials_kernel(
    const float* __restrict__ pdf, const float* __restrict__ samples, int64_t n_samples, int dim,
    int n_split, const float* __restrict__ stats, float eps, float* __restrict__ partials) {
  __shared__ float sh[8];
  const int64_t b = blockIdx.x;
  const int sp = blockIdx.y;
  const int64_t chunk = (n_samples + n_split - 1) / n_split;
  const int64_t lo = sp * chunk;
  const int64_t hi = lo + chunk < n_samples ? lo + chunk : n_samples;
  const float* w = pdf + b * n_samples;
  const float* x = samples + b * n_samples * dim;
  const int width = PASS == 0 ? 2 + dim : dim;
  float acc[2 + kStatsMaxDim];
#pragma unroll
  for (int i = 0; i < 2 + kStatsMaxDim; ++i) acc[i] = 0.0f;
  float mean[kStatsMaxDim];
  bool uniform = false;
  if (PASS == 1) {
    const float* st = stats + b * (2 + 2 * dim);  // {sum w, ess, mean[d], std[d]}
    uniform = !(st[0] > eps);
    for (int d = 0; d < dim; ++d) mean[d] = st[2 + d];
  }
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float wi = sanitize_weight(__ldg(w + i));
    if (PASS == 0) {
      acc[0] += wi;
      acc[1] += wi * wi;
      for (int d = 0; d < dim; ++d) {
        const float xv = __ldg(x + i * dim + d);
        acc[2 + d] += wi * xv;
      }
    } else {
      const float wu = uniform ? 1.0f : wi;
      for (int d = 0; d < dim; ++d) {
        const float df = __ldg(x + i * dim + d) - mean[d];
        acc[d] += wu * df * df;
      }
    }
  }
  for (int k = 0; k < width; ++k) {
    const float t = block_sum256(acc[k], sh);
    if (threadIdx.x == 0) partials[(b * n_split + sp) * (2 + kStatsMaxDim) + k] = t;
  }
}

// plain sums of x for the uniform-weights fallback are folded in here: when sum w <= eps the
// reference uses w = 1/S, so the mean is the plain sample mean; pass 0b recomputes it cheaply.
// One thread per query.  PASS 0: partials -> stats {sum w, ess, mean[d], (std unset)};
// PASS 1: partials -> std[d].
template <int PASS>
__global__ void posterior_merge_kernel(const float* __restrict__ partials, const float* __restrict__ pdf,
                                       const float* __restrict__ samples, int64_t n_queries,
                                       int64_t n_samples, int dim, int n_split, float eps,
                                       float* __restrict__ stats) {
  const int64_t b = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (b >= n_queries) return;
  float* st = stats + b * (2 + 2 * dim);
  float acc[2 + kStatsMaxDim];
  for (int k = 0; k < 2 + kStatsMaxDim; ++k) acc[k] = 0.0f;
  const int width = PASS == 0 ? 2 + dim : dim;
  for (int i = 0; i < n_split; ++i)
    for (int k = 0; k < width; ++k) acc[k] += partials[(b * n_split + i) * (2 + kStatsMaxDim) + k];
  if (PASS == 0) {
    const float sw = acc[0];
    st[0] = sw;
    if (sw > eps) {
      const float den = fmaxf(sw, eps);
      st[1] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(acc[1], den * den), eps));  // 1 / sum (w/den)^2
      for (int d = 0; d < dim; ++d) st[2 + d] = __fdiv_rn(acc[2 + d], den);
    } else {  // uniform weights 1/S (degenerate query): plain mean, ESS = S
      st[1] = __fdiv_rn(1.0f, fmaxf(__fdiv_rn(1.0f, static_cast<float>(n_samples)), eps));
      for (int d = 0; d < dim; ++d) {
        float m = 0.0f;
        for (int64_t i = 0; i < n_samples; ++i) m += samples[(b * n_samples + i) * dim + d];
        st[2 + d] = __fdiv_rn(m, static_cast<float>(n_samples));
      }
    }
  } else {
    const float sw = st[0];
    const float den = sw > eps ? fmaxf(sw, eps) : static_cast<float>(n_samples);
    for (int d = 0; d < dim; ++d) st[2 + dim + d] = sqrtf(fmaxf(__fdiv_rn(acc[d], den), 0.0f));
  }
}

// ---------------------------------------------------------------------------------------
// Resampling step of resampled_importance_sampling (resampled_importance_sampling.py:33-41):
//   idx = torch.multinomial(softmax(logw), S, replacement=True); samples = samples[b, idx]
// as three kernels: per-query inclusive CDF of the weights, inverse-CDF index draws (one Philox
// uniform per output row + binary search), and a gather of the live node columns.
// ---------------------------------------------------------------------------------------
// grid B, 256 threads: cdf[b, s] = sum_{t <= s} w[b, t]; the running carry is kept in double
__global__ void __launch_bounds__(256) row_cdf_kernel(const float* __restrict__ w, int64_t n_samples,
                                                       float* __restrict__ cdf) {
  __shared__ double warp_sums[8];
  __shared__ double carry_sh;
  const int64_t b = blockIdx.x;
  const float* x = w + b * n_samples;
  float* y = cdf + b * n_samples;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_sh = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < n_samples; base += 1024) {
    const int64_t i0 = base + threadIdx.x * 4;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < n_samples ? x[i0 + k] : 0.0f;
    double t = static_cast<double>(v[0]) + v[1] + v[2] + v[3];
    double inc = t;  // inclusive scan of the per-thread sums across the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    double before = carry_sh;
    for (int q = 0; q < warp; ++q) before += warp_sums[q];
    double run = before + inc - t;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      run += v[k];
      if (i0 + k < n_samples) y[i0 + k] = static_cast<float>(run);
    }
    __syncthreads();
    if (threadIdx.x == 255) carry_sh = before + inc;
    __syncthreads();
  }
}

// one thread per output row (b, s): u ~ U(0, total_b) from the row's own Philox stream, then the
// first position whose cumulative weight exceeds u
__global__ void __launch_bounds__(256) resample_index_kernel(const float* __restrict__ cdf, int64_t n_queries,
                                                              int64_t n_samples, uint32_t key0, uint32_t key1,
                                                              uint32_t call_offset, uint32_t query_offset,
                                                              uint32_t sample_offset, int32_t* __restrict__ idx) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= n_queries * n_samples) return;
  const int64_t b = r / n_samples, s = r - b * n_samples;
  const uint4 c = make_uint4(sample_offset + static_cast<uint32_t>(s), query_offset + static_cast<uint32_t>(b),
                             0x7FFFFFFFu, call_offset);  // stream tag 1 (uniform/row), block 0x3FFFFFFF: unused by ops
  const uint4 rnd = philox4x32_10(c, make_uint2(key0, key1));
  const float* row = cdf + b * n_samples;
  const float u = u01(rnd.x) * row[n_samples - 1];
  int64_t lo = 0, hi = n_samples - 1;  // invariant: answer in [lo, hi]
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (row[mid] > u) hi = mid; else lo = mid + 1;
  }
  idx[r] = static_cast<int32_t>(lo);
}

// dst[c][b][s] = src[c][b][idx[b][s]] for n_cols columns of B*S floats
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           const int32_t* __restrict__ idx, int32_t n_cols,
                                                           int64_t n_queries, int64_t n_samples) {
  const int64_t rows = n_queries * n_samples;
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= rows) return;
  const int64_t b = r / n_samples;
  const int64_t from = b * n_samples + idx[r];
  for (int c = 0; c < n_cols; ++c) dst[c * rows + r] = __ldg(src + c * rows + from);
}

// ---------------------------------------------------------------------------------------
// Weighted class histogram of a discrete target: the benchmark adapter's Python double loop
// _estimate_discrete_posterior(_batch) (benchmarking/models/vbn.py:202-242) + _normalize_probs (:116-121):
//   hist[b][c] = sum_s w[b,s] [round_half_even(x[b,s]) == c, w finite];  probs = hist / sum(hist),
//   uniform 1/K when the total is not finite or <= 0.
// grid B, 256 threads, K <= kHistMaxClasses; classes are processed 16 at a time in register
// accumulators (deterministic: no atomics), block-reduced by warp shuffles.
// ---------------------------------------------------------------------------------------
constexpr int kHistMaxClasses = 256;

__global__ void __launch_bounds__(256) weighted_histogram_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  int64_t n_samples, int64_t x_stride, int n_classes,
                                                                  float* __restrict__ probs) {
  __shared__ float sh[8];
  __shared__ float hist[kHistMaxClasses];
  const int64_t b = blockIdx.x;
  for (int c0 = 0; c0 < n_classes; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = 0.0f;
    for (int64_t s = threadIdx.x; s < n_samples; s += blockDim.x) {
      const float wv = __ldg(w + b * n_samples + s);
      const float xv = __ldg(x + (b * n_samples + s) * x_stride);
      const bool ok = wv == wv && fabsf(wv) != CUDART_INF_F && xv == xv && fabsf(xv) < 2.0e9f;
      const int j = ok ? static_cast<int>(rintf(xv)) - c0 : -1;  // Python round(): half to even
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] += j == q ? wv : 0.0f;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float t = block_sum256(acc[q], sh);
      if (threadIdx.x == 0 && c0 + q < n_classes) hist[c0 + q] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.0f;
    for (int c = 0; c < n_classes; ++c) total += hist[c];
    const bool uniform = !(total == total) || fabsf(total) == CUDART_INF_F || total <= 0.0f;
    for (int c = 0; c < n_classes; ++c)
      probs[b * n_classes + c] = uniform ? __fdiv_rn(1.0f, static_cast<float>(n_classes)) : __fdiv_rn(hist[c], total);
  }
}

// ---------------------------------------------------------------------------------------
// rao_blackwellized_marginalization epilogues (rao_blackwellized_marginalization.py:277-317).
// ---------------------------------------------------------------------------------------
// out[b][k] = sum_s w[b,s] * x[b,s,k]   (categorical marginal; K <= 64).  grid B, 256 threads.
__global__ void __launch_bounds__(256) weighted_sum_kernel(const float* __restrict__ w, const float* __restrict__ x,
                                                            int64_t n_samples, int width, float* __restrict__ out) {
  __shared__ float sh[8];
  const int64_t b = blockIdx.x;
  for (int k = 0; k < width; ++k) {
    float acc = 0.0f;
    for (int64_t s = threadIdx.x; s < n_samples; s += blockDim.x)
      acc = fmaf(__ldg(w + b * n_samples + s), __ldg(x + (b * n_samples + s) * width + k), acc);
    const float t = block_sum256(acc, sh);
    if (threadIdx.x == 0) out[b * width + k] = t;
  }
}

// Gaussian-mixture posterior on a grid: per query b, particles s with (w, mu, sigma):
//   mix_mean = sum w mu; mix_std = sqrt(max(sum w (sigma^2 + mu^2) - mix_mean^2, min_scale^2))
//   grid[b,j] = lo + (hi - lo) * linspace(0,1,S_out)[j], lo/hi = mix_mean -/+ stddevs * mix_std
//   pdf[b,j] = sum_s w exp(-0.5 ((x - mu)/sigma)^2) / (sqrt(2 pi) sigma)
// grid B, 256 threads; loc_scale [B][S][2]; sigma sanitised like the reference (:145-149, :311).
__global__ void __launch_bounds__(256) gaussian_mixture_grid_kernel(
    const float* __restrict__ w, const float* __restrict__ loc_scale, int64_t n_particles, int64_t n_out,
    float stddevs, float min_scale, float* __restrict__ pdf, float* __restrict__ grid) {
  __shared__ float sh[8];
  __shared__ float moments[2];
  const int64_t b = blockIdx.x;
  const float* wb = w + b * n_particles;
  const float* ls = loc_scale + b * n_particles * 2;
  auto sigma_of = [&](int64_t s) {
    float sc = __ldg(ls + 2 * s + 1);
    if (!(sc == sc) || fabsf(sc) == CUDART_INF_F) sc = min_scale;
    return fmaxf(fabsf(sc), min_scale);
  };
  float m1 = 0.0f, m2 = 0.0f;
  for (int64_t s = threadIdx.x; s < n_particles; s += blockDim.x) {
    const float wi = __ldg(wb + s), mu = __ldg(ls + 2 * s), sg = sigma_of(s);
    m1 = fmaf(wi, mu, m1);
    m2 = fmaf(wi, fmaf(sg, sg, mu * mu), m2);
  }
  const float t1 = block_sum256(m1, sh);
  if (threadIdx.x == 0) moments[0] = t1;
  const float t2 = block_sum256(m2, sh);
  if (threadIdx.x == 0) moments[1] = t2;
  __syncthreads();
  const float mean = moments[0];
  const float sd = sqrtf(fmaxf(moments[1] - mean * mean, min_scale * min_scale));
  const float lo = mean - stddevs * sd, hi = mean + stddevs * sd;
  const float step = n_out > 1 ? __fdiv_rn(1.0f, static_cast<float>(n_out - 1)) : 0.0f;
  for (int64_t j = threadIdx.x; j < n_out; j += blockDim.x) {
    const float z = j < n_out / 2 ? step * static_cast<float>(j) : 1.0f - step * static_cast<float>(n_out - 1 - j);
    const float x = lo + (hi - lo) * z;
    float acc = 0.0f;
    for (int64_t s = 0; s < n_particles; ++s) {
      const float sg = sigma_of(s);
      const float zn = __fdiv_rn(x - __ldg(ls + 2 * s), sg);
      acc = fmaf(__ldg(wb + s), __fdiv_rn(expf(-0.5f * zn * zn), 2.5066282746310002f * sg), acc);
    }
    grid[b * n_out + j] = x;
    pdf[b * n_out + j] = acc;
  }
}

// gaussian_exact grid (vbn/inference/gaussian_exact.py:166-183): per query b with Normal(loc_b, scale_b)
//   z_s = linspace(-k, k, S)[s];  samples[b,s] = loc_b + scale_b * z_s;
//   pdf[b,s] = exp(-0.5 * (z_s^2 + 2 ln scale_b + ln 2 pi))
// scale is sanitised like the reference: nan/inf -> min_scale, abs, clamp_min(min_scale).
__global__ void gaussian_grid_kernel(const float* __restrict__ loc_scale, int64_t n_queries, int64_t n_samples,
                                     float stddevs, float min_scale, float* __restrict__ pdf,
                                     float* __restrict__ samples) {
  const int64_t total = n_queries * n_samples;
  const float step = n_samples > 1 ? __fdiv_rn(stddevs - (-stddevs), static_cast<float>(n_samples - 1)) : 0.0f;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < total;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = r / n_samples, s = r - b * n_samples;
    const float loc = loc_scale[2 * b];
    float sc = loc_scale[2 * b + 1];
    if (!(sc == sc) || fabsf(sc) == CUDART_INF_F) sc = min_scale;
    sc = fmaxf(fabsf(sc), min_scale);
    // torch.linspace: start + step*i in the lower half, end - step*(S-1-i) in the upper half
    const float z = s < n_samples / 2 ? fmaf(step, static_cast<float>(s), -stddevs)
                                      : stddevs - step * static_cast<float>(n_samples - 1 - s);
    samples[r] = loc + sc * z;
    pdf[r] = expf(-0.5f * (z * z + 2.0f * logf(sc) + kLog2Pi));
  }
}

__global__ void philox_fill_kernel(const uint32_t* __restrict__ ctr, int64_t n, uint32_t k0,
                                   uint32_t k1, uint32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint4 c = make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]);
  const uint4 r = philox4x32_10(c, make_uint2(k0, k1));
  out[4 * i] = r.x;
  out[4 * i + 1] = r.y;
  out[4 * i + 2] = r.z;
  out[4 * i + 3] = r.w;
}

// The draws the schedule kernels make, exported for replay checks (tests regenerate a run's noise and feed it to
// the CPU oracle).  Same code path as Ctx::normals / Ctx::uniforms with the inlined generator: counter
// (global sample, global query | 0xFFFFFFFF when shared, block | tag << 30, call offset), round keys as kernel
// parameters, normal4 / uniform4.  out[(4 * blk + i) * Bn * S + b * S + s], Bn = 1 for the shared streams.
// kind 0: normals, 1: uniforms, 2: raw generator words (bit patterns), tags as in vbn_schedule.cuh.
struct StreamDrawArgs {
  uint32_t rk[20];
  uint32_t call_offset, query_offset, sample_offset;
  int32_t kind, shared, block_lo, n_blocks;
  int64_t n_queries, n_samples;
};
__global__ void __launch_bounds__(256) stream_draws_kernel(const StreamDrawArgs a, float* __restrict__ out) {
  const int64_t bn = a.shared ? 1 : a.n_queries;
  const int64_t per_block = bn * a.n_samples;
  const int64_t total = per_block * a.n_blocks;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t blk = i / per_block, r = i - blk * per_block;
    const int64_t b = r / a.n_samples, s = r - b * a.n_samples;
    const uint32_t tag = (a.kind == 1 ? 1u : 0u) + (a.shared ? 2u : 0u);
    const uint4 c = make_uint4(a.sample_offset + static_cast<uint32_t>(s),
                               a.shared ? 0xFFFFFFFFu : a.query_offset + static_cast<uint32_t>(b),
                               static_cast<uint32_t>(a.block_lo + blk) | (tag << 30), a.call_offset);
    const uint4 w = philox4x32_10_rk(c, a.rk);
    float4 v;
    if (a.kind == 0) {
      v = normal4(w);
    } else if (a.kind == 1) {
      v = uniform4(w);
    } else {
      v = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
    }
    float* o = out + (4 * blk) * per_block + r;
    o[0] = v.x;
    o[per_block] = v.y;
    o[2 * per_block] = v.z;
    o[3 * per_block] = v.w;
  }
}

}  // namespace vbn

// FP32 FMA peak probe (bench only): the MLP layers run on the FFMA pipe in fp32 (1e-5 parity
// rules out TF32), and MEASURED_PEAKS.json has no fp32 figure, so bench.py measures the
// denominator itself.  mode 0: scalar FFMA, mode 1: packed fma.rn.f32x2 (FFMA2).
namespace vbn {
__global__ void __launch_bounds__(256) fma_peak_kernel(int mode, int iters, float* __restrict__ out) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const float m = 0.999f + out[0] * 0.0f, c = 1e-3f;
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
  } else {
    unsigned long long p[8], mm, cc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(mm), "l"(cc));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;  // keep the chain alive without a store in the common case
}
}  // namespace vbn
