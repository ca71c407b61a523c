// TEST INFRASTRUCTURE ONLY -- host emulation of libvbn_cuda.so's C ABI (see cuda_shim.h).
// The schedule kernel is the real device code compiled for the host; reductions / KDE use the
// thread-independent generic code paths or plain loops.  Pointers are HOST pointers.
#include "cuda_shim.h"

thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
namespace vbn { float smem[256 * 1024]; }

#include "vbn_cuda.h"
#include "vbn_schedule.cuh"

#include <cstdio>
#include <vector>

struct VbnPlan { VbnProgramDesc desc; };

namespace {
template <int RPT, int NT, bool HEAVY>
void run_kernel(const vbn::ScheduleArgs& a) {
  const int64_t rows = static_cast<int64_t>(RPT) * NT;
  const int64_t n_tiles = (a.n_rows + rows - 1) / rows;
  gridDim.x = static_cast<unsigned>(n_tiles);
  blockDim.x = NT;
  for (unsigned b = 0; b < gridDim.x; ++b) {
    blockIdx.x = b;
    for (unsigned t = 0; t < NT; ++t) {
      threadIdx.x = t;
      vbn::schedule_kernel<RPT, NT, HEAVY, 1>(a);
    }
  }
}
}  // namespace

extern "C" {
int32_t vbn_cuda_abi_version(void) { return VBN_CUDA_ABI_VERSION; }
const char* vbn_cuda_last_error(void) { return "emu"; }
int32_t vbn_cuda_device_count(int32_t* n) { *n = 0; return 0; }
int32_t vbn_plan_create(const VbnProgramDesc* d, VbnPlan** out) { *out = new VbnPlan{*d}; return 0; }
int32_t vbn_plan_destroy(VbnPlan* p) { delete p; return 0; }
int32_t vbn_run_forward_launches(const VbnPlan*) { return 1; }

int32_t vbn_run_forward(const VbnPlan* plan, const VbnRunDesc* run, void*) {
  vbn::ScheduleArgs a;
  std::memset(&a, 0, sizeof(a));
  a.ops = plan->desc.ops_dev; a.par_slots = plan->desc.par_slots_dev; a.params = plan->desc.params_dev;
  a.n_ops = plan->desc.n_ops; a.n_slots = plan->desc.n_slots; a.n_scratch = plan->desc.n_scratch;
  a.logp_as_pdf = run->logp_as_pdf; a.logw_accumulate = run->logw_accumulate; a.n_queries = run->n_queries; a.n_samples = run->n_samples;
  a.n_rows = run->n_queries * run->n_samples;
  a.query_offset = (uint32_t)run->query_offset; a.sample_offset = (uint32_t)run->sample_offset;
  a.key0 = (uint32_t)run->seed; a.key1 = (uint32_t)(run->seed >> 32); a.call_offset = (uint32_t)run->call_offset; vbn::fill_round_keys(a);
  a.fixed = run->fixed_dev; a.inputs = run->inputs_dev; a.stores = run->stores_dev; a.noise = run->noise_dev;
  a.logw = run->logw_dev; a.logp = run->logp_dev; a.error_flag = run->error_flag_dev;
  a.seg = run->seg_dev; a.seg_per_query = run->seg_per_query; a.seg_slot = run->seg_slot; a.seg_classes = run->seg_classes;
  if ((size_t)(a.n_slots + a.n_scratch) * 2 * 32 > sizeof(vbn::smem) / sizeof(float)) return VBN_E_CAPACITY;
  if (plan->desc.heavy) run_kernel<2, 32, true>(a); else run_kernel<2, 32, false>(a);
  return 0;
}

int32_t vbn_lse_partials(const float* x, int64_t B, int64_t S, int32_t n_split, float* part, void*) {
  const int64_t chunk = (S + n_split - 1) / n_split;
  for (int64_t b = 0; b < B; ++b)
    for (int sp = 0; sp < n_split; ++sp) {
      const int64_t lo = sp * chunk, hi = std::min<int64_t>(lo + chunk, S);
      float m = -CUDART_INF_F; double l = 0, q = 0;
      for (int64_t i = lo; i < hi; ++i) m = std::max(m, x[b * S + i]);
      for (int64_t i = lo; i < hi; ++i) { const double e = std::exp((double)x[b * S + i] - m); l += e; q += e * e; }
      float* o = part + (b * n_split + sp) * 3; o[0] = m; o[1] = (float)l; o[2] = (float)q;
    }
  return 0;
}
int32_t vbn_lse_merge(const float* part, int64_t B, int32_t n, float* st, void*) {
  for (int64_t b = 0; b < B; ++b) {
    float m = -CUDART_INF_F;
    for (int i = 0; i < n; ++i) m = std::max(m, part[(b * n + i) * 3]);
    double l = 0, q = 0;
    for (int i = 0; i < n; ++i) {
      const float* p = part + (b * n + i) * 3;
      if (p[0] == -CUDART_INF_F) continue;
      const double e = std::exp((double)p[0] - m); l += p[1] * e; q += p[2] * e * e;
    }
    st[b * 3] = m; st[b * 3 + 1] = (float)l; st[b * 3 + 2] = (float)q;
  }
  return 0;
}
int32_t vbn_weights_normalize(const float* x, const float* st, int64_t B, int64_t S, int32_t norm, float eps,
                              float* w, float* ess, void*) {
  for (int64_t b = 0; b < B; ++b) {
    const float m = st[b * 3], l = st[b * 3 + 1], q = st[b * 3 + 2];
    for (int64_t i = 0; i < S; ++i) {
      const float e = std::exp(x[b * S + i] - m);
      w[b * S + i] = norm ? e / l : std::max(e, eps);
    }
    if (ess) ess[b] = l * l / q;
  }
  return 0;
}
int32_t vbn_segment_merge(const float* rec, int64_t B, int64_t S, int32_t P, float thr, float* merged, float* stats,
                          int32_t* flag, void*) {
  for (int64_t b = 0; b < B; ++b) {
    const int64_t count = S > 0 ? (((b + 1) * S - 1) >> 5) - ((b * S) >> 5) + 1 : P;
    double m = -CUDART_INF_F, l = 0, q = 0, mean = 0, m2 = 0, n = 0, h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = 0; i < count; ++i) {
      const float* r = rec + (b * P + i) * 16;
      if (r[6] == 0.0f) continue;
      const double rl = r[1], rmean = rl > 0 ? r[3] + r[4] / rl : r[3], rm2 = rl > 0 ? std::max(0.0, r[5] - (double)r[4] * r[4] / rl) : 0.0;
      if (n == 0) { m = r[0]; l = rl; q = r[2]; mean = rmean; m2 = rm2; n = r[6]; for (int k = 0; k < 8; ++k) h[k] = r[8 + k]; continue; }
      const double mm = std::max(m, (double)r[0]);
      const double sa = m > -CUDART_INF_F ? std::exp(m - mm) : 0.0, sb = r[0] > -CUDART_INF_F ? std::exp(r[0] - mm) : 0.0;
      const double la = l * sa, lb = rl * sb, lt = la + lb, d = rmean - mean, fb = lt > 0 ? lb / lt : 0.0;
      mean += d * fb; m2 = m2 * sa + rm2 * sb + d * d * la * fb; q = q * sa * sa + r[2] * sb * sb; l = lt; m = mm; n += r[6];
      for (int k = 0; k < 8; ++k) h[k] = h[k] * sa + r[8 + k] * sb;
    }
    const double ess = l * l / q;  // 0/0 = NaN for an all -inf query, like torch.softmax
    if (merged) {
      float* o = merged + b * 16;
      o[0] = (float)m; o[1] = (float)l; o[2] = (float)q; o[3] = (float)mean; o[4] = 0; o[5] = (float)m2; o[6] = (float)n; o[7] = (float)ess;
      for (int k = 0; k < 8; ++k) o[8 + k] = (float)h[k];
    }
    if (stats) { stats[b * 3] = (float)m; stats[b * 3 + 1] = (float)l; stats[b * 3 + 2] = (float)q; }
    if (flag && thr > 0 && ess < thr) *flag |= 1;
  }
  return 0;
}
int32_t vbn_ess_below(const float* st, int64_t B, float thr, int32_t* flag, void*) {
  for (int64_t b = 0; b < B; ++b) if (st[b * 3 + 1] * st[b * 3 + 1] / st[b * 3 + 2] < thr) *flag |= 1;
  return 0;
}
int32_t vbn_row_cdf(const float* w, int64_t B, int64_t S, float* cdf, void*) {
  for (int64_t b = 0; b < B; ++b) { double run = 0; for (int64_t s = 0; s < S; ++s) { run += w[b * S + s]; cdf[b * S + s] = (float)run; } }
  return 0;
}
int32_t vbn_resample_indices(const float* cdf, int64_t B, int64_t S, uint64_t seed, uint64_t call, int64_t qo, int64_t so,
                             int32_t* idx, void*) {
  for (int64_t b = 0; b < B; ++b) for (int64_t s = 0; s < S; ++s) {
    const uint4 c = make_uint4((uint32_t)(so + s), (uint32_t)(qo + b), 0x7FFFFFFFu, (uint32_t)call);
    const uint4 r = vbn::philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float* row = cdf + b * S;
    const float u = vbn::u01(r.x) * row[S - 1];
    int64_t lo = 0, hi = S - 1;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (row[mid] > u) hi = mid; else lo = mid + 1; }
    idx[b * S + s] = (int32_t)lo;
  }
  return 0;
}
int32_t vbn_gather_rows(const float* src, float* dst, const int32_t* idx, int32_t n_cols, int64_t B, int64_t S, void*) {
  const int64_t rows = B * S;
  for (int c = 0; c < n_cols; ++c) for (int64_t r = 0; r < rows; ++r) dst[c * rows + r] = src[c * rows + (r / S) * S + idx[r]];
  return 0;
}

int32_t vbn_weighted_sum(const float* w, const float* x, int64_t B, int64_t S, int32_t K, float* out, void*) {
  for (int64_t b = 0; b < B; ++b) for (int k = 0; k < K; ++k) {
    double acc = 0; for (int64_t s = 0; s < S; ++s) acc += (double)w[b * S + s] * x[(b * S + s) * K + k];
    out[b * K + k] = (float)acc;
  }
  return 0;
}
int32_t vbn_weighted_histogram(const float* x, const float* w, int64_t B, int64_t S, int64_t stride, int32_t K, float* probs, void*) {
  for (int64_t b = 0; b < B; ++b) {
    std::vector<double> h(K, 0.0);
    for (int64_t s = 0; s < S; ++s) {
      const float wv = w[b * S + s], xv = x[(b * S + s) * stride];
      if (!(wv == wv) || std::fabs(wv) == INFINITY || !(xv == xv) || !(std::fabs(xv) < 2.0e9f)) continue;
      const long j = std::lrint(xv);  // round half to even (default rounding mode)
      if (j >= 0 && j < K) h[j] += wv;
    }
    double tot = 0; for (double v : h) tot += v;
    const bool uni = !(tot == tot) || std::fabs(tot) == INFINITY || tot <= 0;
    for (int c = 0; c < K; ++c) probs[b * K + c] = uni ? 1.0f / K : (float)(h[c] / tot);
  }
  return 0;
}
int32_t vbn_gaussian_mixture_grid(const float* w, const float* ls, int64_t B, int64_t S, int64_t N, float k, float min_scale,
                                  float* pdf, float* grid, void*) {
  for (int64_t b = 0; b < B; ++b) {
    auto sg = [&](int64_t s) { float v = ls[(b * S + s) * 2 + 1]; if (!(v == v) || std::fabs(v) == INFINITY) v = min_scale; return std::max(std::fabs(v), min_scale); };
    double m1 = 0, m2 = 0;
    for (int64_t s = 0; s < S; ++s) { const double wi = w[b * S + s], mu = ls[(b * S + s) * 2], g = sg(s); m1 += wi * mu; m2 += wi * (g * g + mu * mu); }
    const float mean = (float)m1, sd = std::sqrt(std::max((float)(m2 - m1 * m1), min_scale * min_scale));
    const float lo = mean - k * sd, hi = mean + k * sd, step = N > 1 ? 1.0f / (float)(N - 1) : 0.0f;
    for (int64_t j = 0; j < N; ++j) {
      const float z = j < N / 2 ? step * (float)j : 1.0f - step * (float)(N - 1 - j);
      const float x = lo + (hi - lo) * z;
      double acc = 0;
      for (int64_t s = 0; s < S; ++s) { const float g = sg(s); const float zn = (x - ls[(b * S + s) * 2]) / g; acc += (double)w[b * S + s] * std::exp(-0.5f * zn * zn) / (2.5066282746310002f * g); }
      grid[b * N + j] = x; pdf[b * N + j] = (float)acc;
    }
  }
  return 0;
}

int32_t vbn_gaussian_grid(const float* ls, int64_t B, int64_t S, float k, float min_scale, float* pdf, float* x, void*) {
  const float step = S > 1 ? (k - (-k)) / (float)(S - 1) : 0.0f;
  for (int64_t b = 0; b < B; ++b) {
    float sc = ls[2 * b + 1];
    if (!(sc == sc) || std::fabs(sc) == INFINITY) sc = min_scale;
    sc = std::max(std::fabs(sc), min_scale);
    for (int64_t s = 0; s < S; ++s) {
      const float z = s < S / 2 ? -k + step * (float)s : k - step * (float)(S - 1 - s);
      x[b * S + s] = ls[2 * b] + sc * z;
      pdf[b * S + s] = std::exp(-0.5f * (z * z + 2.0f * std::log(sc) + 1.8378770664093453f));
    }
  }
  return 0;
}

int32_t vbn_posterior_stats(const float* pdf, const float* x, int64_t B, int64_t S, int32_t D, int32_t, float eps,
                            float*, float* stats, void*) {
  for (int64_t b = 0; b < B; ++b) {
    float* st = stats + b * (2 + 2 * D);
    double sw = 0, sw2 = 0;
    std::vector<double> m(D, 0.0), v(D, 0.0);
    auto wt = [&](int64_t i) { float w = pdf[b * S + i]; return (w == w && std::fabs(w) != INFINITY && w > 0) ? w : 0.0f; };
    for (int64_t i = 0; i < S; ++i) sw += wt(i);
    const bool uni = !(sw > eps);
    const double den = uni ? (double)S : sw;
    for (int64_t i = 0; i < S; ++i) {
      const double w = uni ? 1.0 : wt(i);
      sw2 += (w / den) * (w / den);
      for (int d = 0; d < D; ++d) m[d] += w / den * x[(b * S + i) * D + d];
    }
    for (int64_t i = 0; i < S; ++i) {
      const double w = uni ? 1.0 : wt(i);
      for (int d = 0; d < D; ++d) { const double df = x[(b * S + i) * D + d] - m[d]; v[d] += w / den * df * df; }
    }
    st[0] = (float)sw; st[1] = (float)(1.0 / std::max(sw2, (double)eps));
    for (int d = 0; d < D; ++d) { st[2 + d] = (float)m[d]; st[2 + D + d] = (float)std::sqrt(std::max(v[d], 0.0)); }
  }
  return 0;
}

int32_t vbn_kde_log_prob(const float* tp, const float* ty, int64_t N, int32_t dp, int32_t dx, const float* qp,
                         const float* qx, int64_t M, float bw, float pbw, float ms, float* out, void*) {
  const double sy = std::max((double)bw, 1e-3) + ms, sp = std::max((double)pbw, 1e-3) + ms;
  const float hy = (float)(0.5 / (sy * sy)), hp = (float)(0.5 / (sp * sp));
  const float cy = (float)(-0.5 * dx * (1.8378770664093453 + 2.0 * std::log(sy))), ln = (float)std::log((double)N);
  for (int64_t r = 0; r < M; ++r) {
    vbn::Lse num, den; num.init(); den.init();
    for (int64_t n = 0; n < N; ++n) {
      float q1 = 0, q2 = 0;
      for (int d = 0; d < dp; ++d) { const float df = qp[r * dp + d] - tp[n * dp + d]; q1 += df * df; }
      for (int d = 0; d < dx; ++d) { const float df = qx[r * dx + d] - ty[n * dx + d]; q2 += df * df; }
      den.push(-hp * q1); num.push(-hp * q1 - hy * q2);
    }
    out[r] = dp > 0 ? num.value() - den.value() + cy : num.value() + cy - ln;
  }
  return 0;
}
int32_t vbn_stream_draws(uint64_t seed, uint64_t call, int32_t kind, int32_t shared, int32_t block_lo, int32_t n_blocks,
                         int64_t B, int64_t S, int64_t qo, int64_t so, float* out, void*) {
  vbn::ScheduleArgs keys; keys.key0 = (uint32_t)seed; keys.key1 = (uint32_t)(seed >> 32); vbn::fill_round_keys(keys);
  const int64_t bn = shared ? 1 : B, per_block = bn * S;
  const uint32_t tag = (kind == 1 ? 1u : 0u) + (shared ? 2u : 0u);
  for (int64_t blk = 0; blk < n_blocks; ++blk) for (int64_t b = 0; b < bn; ++b) for (int64_t s = 0; s < S; ++s) {
    const uint4 c = make_uint4((uint32_t)(so + s), shared ? 0xFFFFFFFFu : (uint32_t)(qo + b), (uint32_t)(block_lo + blk) | (tag << 30), (uint32_t)call);
    const uint4 w = vbn::philox4x32_10_rk(c, keys.rk);
    float v[4];
    if (kind == 0) { const float4 n = vbn::normal4(w); v[0] = n.x; v[1] = n.y; v[2] = n.z; v[3] = n.w; }
    else if (kind == 1) { const float4 u = vbn::uniform4(w); v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
    else { std::memcpy(v, &w, 16); }
    for (int i = 0; i < 4; ++i) out[(4 * blk + i) * per_block + b * S + s] = v[i];
  }
  return 0;
}
int32_t vbn_kde_tc_workspace_bytes(int64_t, int32_t, int32_t, int64_t* out) { *out = 0; return 0; }  // not covered here
int32_t vbn_kde_log_prob_tc(const float*, const float*, int64_t, int32_t, int32_t, const float*, const float*, int64_t,
                            float, float, float, const float*, void*, float*, void*) { return VBN_E_CAPACITY; }
int32_t vbn_fma_peak(int32_t, int32_t, int32_t, float*, void*) { return 0; }
int32_t vbn_tf32_peak(int32_t, int32_t, float*, void*) { return 0; }
int32_t vbn_philox_fill(const uint32_t* ctr, int64_t n, uint32_t k0, uint32_t k1, uint32_t* out, void*) {
  for (int64_t i = 0; i < n; ++i) {
    const uint4 r = vbn::philox4x32_10(make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]), make_uint2(k0, k1));
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
  }
  return 0;
}
}
