// Stand-alone KDE conditional log-density kernel (vbn/cpds/kde.py:111-149):
//   out[m] = LSE_n(log_kp[m,n] + log_ky[m,n]) - LSE_n(log_kp[m,n])      (dp > 0)
//   out[m] = LSE_n(log_ky[m,n]) - ln N                                   (dp == 0)
// with log_k = sum_d -0.5*((diff/s)^2 + ln 2pi + 2 ln s), s = max(h,1e-3)+min_scale
// (kde.py:105-109).  The per-dim constants are hoisted: the parent constant cancels between
// the two logsumexps, the target constant is added once at the end.
//
// Work decomposition: each thread owns QPT query rows and streams all N stored points from
// shared memory (tiles of TILE points staged cooperatively, double buffered with cp.async),
// keeping two online logsumexp accumulators per query in registers.  The kernel is MUFU
// (ex2) bound at small dims: 2 ex2 per (query, point) pair; everything is in the log2
// domain so each exp is one FFMA-free ex2.approx.
#pragma once
#include <cuda_pipeline.h>

#include <cstdlib>

#include "vbn_device.cuh"

namespace vbn {

constexpr int kKdeTile = 1024;  // stored points per shared-memory tile
constexpr int kKdeThreads = 128;
constexpr int kKdeQpt = 4;      // query rows per thread
constexpr int kKdeChunk = 4;    // points per online-max step (two packed pairs)

// ---- packed fp32x2 arithmetic (sm_100a: one FMA-pipe issue slot does two lanes' worth) ----------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// exp2 of a packed pair on the FMA pipe instead of the MUFU (the FlashAttention-4 trick): the kernel is
// bound by 16 MUFU lanes/clk/SM while the FMA pipe has issue slots to spare, so a fraction of the
// exponentials is evaluated as 2^round(x) * p(x - round(x)) with a degree-5 polynomial on [-0.5, 0.5]
// (max relative error 2.4e-7, the same 2 ulp as ex2.approx).  Arguments are <= 0; they are clamped at
// -126 so the exponent arithmetic cannot wrap (such terms are ~1e-38, i.e. zero for every sum the fast
// pass accepts).  12 instructions per pair of values.
__device__ __forceinline__ f32x2 ex2_poly2(f32x2 x) {
  float a, b;
  unpack2(x, a, b);
  const f32x2 xc = pack2(fmaxf(a, -126.0f), fmaxf(b, -126.0f));
  const f32x2 magic = pack2(12582912.0f, 12582912.0f);  // 1.5 * 2^23: adding it rounds to an integer
  const f32x2 t = add2(xc, magic);
  const f32x2 f = sub2(xc, sub2(t, magic));
  f32x2 p = pack2(0.0013390863314270973f, 0.0013390863314270973f);
  p = fma2(p, f, pack2(0.009676031768321991f, 0.009676031768321991f));
  p = fma2(p, f, pack2(0.055503569543361664f, 0.055503569543361664f));
  p = fma2(p, f, pack2(0.2402210682630539f, 0.2402210682630539f));
  p = fma2(p, f, pack2(0.6931471824645996f, 0.6931471824645996f));
  p = fma2(p, f, pack2(1.0000001192092896f, 1.0000001192092896f));
  float pa, pb, ta, tb;
  unpack2(p, pa, pb);
  unpack2(t, ta, tb);
  // the low mantissa bits of t hold round(x): shifted into the exponent field they scale p by 2^round(x)
  const float ra = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
  const float rb = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
  return pack2(ra, rb);
}

// One online (max, sum) accumulator, two-level: `cur` (packed pair) collects the current tile,
// `tot` the finished tiles, so the fp32 error grows with sqrt(tile) + sqrt(#tiles) instead of
// sqrt(N) (200 000 sequential adds cost ~3e-5 relative, which fails the 1e-5 parity).
struct KdeAcc {
  float m, tot;
  f32x2 cur, mm;  // mm = (m, m), kept packed so the hot path never rebuilds it
  __device__ __forceinline__ void init() {
    m = -CUDART_INF_F;
    tot = 0.0f;
    cur = pack2(0.0f, 0.0f);
    mm = pack2(m, m);
  }
  // v0, v1: four log2-domain terms (two packed pairs)
  __device__ __forceinline__ void push4(f32x2 v0, f32x2 v1) {
    float a, b, c, d;
    unpack2(v0, a, b);
    unpack2(v1, c, d);
    const float mx = fmaxf(fmaxf(a, b), fmaxf(c, d));
    if (mx > m) {
      const float sc = ex2_approx(m - mx);  // exp2(-inf) = 0 on the first chunk
      tot *= sc;
      cur = mul2(cur, pack2(sc, sc));
      m = mx;
      mm = pack2(mx, mx);
    }
    unpack2(sub2(v0, mm), a, b);
    unpack2(sub2(v1, mm), c, d);
    cur = add2(cur, pack2(ex2_approx(a), ex2_approx(b)));
    cur = add2(cur, pack2(ex2_approx(c), ex2_approx(d)));
  }
  __device__ __forceinline__ void end_tile() {
    float a, b;
    unpack2(cur, a, b);
    tot += a + b;
    cur = pack2(0.0f, 0.0f);
  }
  __device__ __forceinline__ float log2_value() const { return m + log2f(tot); }
};

// Fast accumulator: every log2-domain term is <= 0 (it is minus a squared distance), so the sum
// can be taken with a fixed shift m = 0 -- no running max, no rescale -- as long as the nearest
// stored point keeps the sum above the fp32 underflow range.  Rows whose sums fall below 2^-100
// (queries ~12 kernel widths away from every stored point) make the CTA redo the pass with KdeAcc.
// NPOLY of the two packed pairs of every push go through ex2_poly2, the rest through the MUFU.
template <int NPOLY>
struct KdeFastAcc {
  float tot;
  f32x2 cur;
  __device__ __forceinline__ void init() {
    tot = 0.0f;
    cur = pack2(0.0f, 0.0f);
  }
  __device__ __forceinline__ void push4(f32x2 v0, f32x2 v1) {
    float a, b, c, d;
    unpack2(v0, a, b);
    unpack2(v1, c, d);
    cur = add2(cur, NPOLY >= 2 ? ex2_poly2(v0) : pack2(ex2_approx(a), ex2_approx(b)));
    cur = add2(cur, NPOLY >= 1 ? ex2_poly2(v1) : pack2(ex2_approx(c), ex2_approx(d)));
  }
  __device__ __forceinline__ void end_tile() {
    float a, b;
    unpack2(cur, a, b);
    tot += a + b;
    cur = pack2(0.0f, 0.0f);
  }
  __device__ __forceinline__ float log2_value() const { return log2f(tot); }
  __device__ __forceinline__ bool underflowed() const { return !(tot > 7.888609e-31f); }  // 2^-100
};

// Shared-memory tile layout: dimension-major, [d][kKdeTile] floats, so a float2 load fetches the
// same coordinate of two consecutive points (the two lanes of the packed math).  Tile tails are
// filled with +inf parents / targets: their terms are exp2(-inf) = 0.
template <int DP, int DX, class Acc, class AccD>
__device__ __forceinline__ void kde_pass(const float* __restrict__ tp, const float* __restrict__ ty,
                                         int64_t n_points, float (*s_p)[(DP > 0 ? DP : 1) * kKdeTile],
                                         float (*s_y)[DX * kKdeTile],
                                         const f32x2 (&xp)[kKdeQpt][DP > 0 ? DP : 1],
                                         const f32x2 (&xy)[kKdeQpt][DX], float hp2, float hy2,
                                         AccD (&den)[kKdeQpt], Acc (&num)[kKdeQpt]) {
  constexpr int DPS = DP > 0 ? DP : 1;
#pragma unroll
  for (int j = 0; j < kKdeQpt; ++j) {
    den[j].init();
    num[j].init();
  }
  const f32x2 nhp = pack2(-hp2, -hp2), nhy = pack2(-hy2, -hy2);
  const int64_t n_tiles = (n_points + kKdeTile - 1) / kKdeTile;
  auto stage = [&](int64_t t, int buf) {
    const int64_t base = t * kKdeTile;
    const int cnt = static_cast<int>(n_points - base < kKdeTile ? n_points - base : kKdeTile);
    for (int i = threadIdx.x; i < kKdeTile; i += kKdeThreads) {
      if (i < cnt) {
#pragma unroll
        for (int d = 0; d < DP; ++d)
          __pipeline_memcpy_async(&s_p[buf][d * kKdeTile + i], tp + (base + i) * DP + d, sizeof(float));
#pragma unroll
        for (int d = 0; d < DX; ++d)
          __pipeline_memcpy_async(&s_y[buf][d * kKdeTile + i], ty + (base + i) * DX + d, sizeof(float));
      } else {  // sentinel: infinitely far away
#pragma unroll
        for (int d = 0; d < DP; ++d) s_p[buf][d * kKdeTile + i] = CUDART_INF_F;
#pragma unroll
        for (int d = 0; d < DX; ++d) s_y[buf][d * kKdeTile + i] = CUDART_INF_F;
      }
    }
    __pipeline_commit();
  };

  stage(0, 0);
  for (int64_t t = 0; t < n_tiles; ++t) {
    const int buf = static_cast<int>(t & 1);
    if (t + 1 < n_tiles) {
      stage(t + 1, buf ^ 1);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
#pragma unroll 2
    for (int n0 = 0; n0 < kKdeTile; n0 += kKdeChunk) {
      f32x2 a[kKdeQpt][2], cc[kKdeQpt][2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        f32x2 pv[DPS], yv[DX];
#pragma unroll
        for (int d = 0; d < DP; ++d) pv[d] = *reinterpret_cast<const f32x2*>(&s_p[buf][d * kKdeTile + n0 + 2 * i]);
#pragma unroll
        for (int d = 0; d < DX; ++d) yv[d] = *reinterpret_cast<const f32x2*>(&s_y[buf][d * kKdeTile + n0 + 2 * i]);
#pragma unroll
        for (int j = 0; j < kKdeQpt; ++j) {
          f32x2 av = pack2(0.0f, 0.0f);
          if (DP > 0) {
            f32x2 df = sub2(xp[j][0], pv[0]);
            f32x2 q1 = mul2(df, df);
#pragma unroll
            for (int d = 1; d < DP; ++d) {
              df = sub2(xp[j][d], pv[d]);
              q1 = fma2(df, df, q1);
            }
            av = mul2(q1, nhp);
          }
          f32x2 df = sub2(xy[j][0], yv[0]);
          f32x2 q2 = mul2(df, df);
#pragma unroll
          for (int d = 1; d < DX; ++d) {
            df = sub2(xy[j][d], yv[d]);
            q2 = fma2(df, df, q2);
          }
          a[j][i] = av;
          cc[j][i] = DP > 0 ? fma2(q2, nhy, av) : mul2(q2, nhy);
        }
      }
#pragma unroll
      for (int j = 0; j < kKdeQpt; ++j) {
        if (DP > 0) den[j].push4(a[j][0], a[j][1]);
        num[j].push4(cc[j][0], cc[j][1]);
      }
    }
#pragma unroll
    for (int j = 0; j < kKdeQpt; ++j) {
      den[j].end_tile();
      num[j].end_tile();
    }
    __syncthreads();
  }
}

// PN / PD: how many of the two value pairs per push the numerator / denominator accumulators evaluate
// with the polynomial (0..2 each): (PN + PD) / 4 of all exponentials leave the MUFU.
template <int DP, int DX, int PN, int PD>
__global__ void __launch_bounds__(kKdeThreads) kde_log_prob_kernel(
    const float* __restrict__ tp, const float* __restrict__ ty, int64_t n_points,
    const float* __restrict__ qp, const float* __restrict__ qx, int64_t n_rows,
    float hp2, float hy2, float const_y, float log_n, float* __restrict__ out) {
  // hp2 = 0.5*log2(e)/s_p^2, hy2 = 0.5*log2(e)/s_y^2  (log2 domain)
  constexpr int DPS = DP > 0 ? DP : 1;
  __shared__ __align__(16) float s_p[2][DPS * kKdeTile];
  __shared__ __align__(16) float s_y[2][DX * kKdeTile];

  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * kKdeThreads * kKdeQpt) + threadIdx.x;
  f32x2 xp[kKdeQpt][DPS], xy[kKdeQpt][DX];
#pragma unroll
  for (int j = 0; j < kKdeQpt; ++j) {
    int64_t r = row0 + static_cast<int64_t>(j) * kKdeThreads;
    if (r >= n_rows) r = n_rows - 1;
#pragma unroll
    for (int d = 0; d < DPS; ++d) {
      const float v = DP > 0 ? __ldg(qp + r * DP + d) : 0.0f;
      xp[j][d] = pack2(v, v);
    }
#pragma unroll
    for (int d = 0; d < DX; ++d) {
      const float v = __ldg(qx + r * DX + d);
      xy[j][d] = pack2(v, v);
    }
  }
  constexpr float kLn2 = 0.6931471805599453f;
  float res[kKdeQpt];
  bool redo = false;
  {
    KdeFastAcc<PD> den[kKdeQpt];
    KdeFastAcc<PN> num[kKdeQpt];
    kde_pass<DP, DX>(tp, ty, n_points, s_p, s_y, xp, xy, hp2, hy2, den, num);
#pragma unroll
    for (int j = 0; j < kKdeQpt; ++j) {
      const float nv = num[j].log2_value() * kLn2;
      res[j] = DP > 0 ? nv - den[j].log2_value() * kLn2 + const_y : nv + const_y - log_n;
      redo = redo || num[j].underflowed() || (DP > 0 && den[j].underflowed());
    }
  }
  if (__syncthreads_or(redo ? 1 : 0)) {  // far-tail rows in this CTA: exact online-max pass
    KdeAcc den[kKdeQpt], num[kKdeQpt];
    kde_pass<DP, DX>(tp, ty, n_points, s_p, s_y, xp, xy, hp2, hy2, den, num);
#pragma unroll
    for (int j = 0; j < kKdeQpt; ++j) {
      const float nv = num[j].log2_value() * kLn2;
      res[j] = DP > 0 ? nv - den[j].log2_value() * kLn2 + const_y : nv + const_y - log_n;
    }
  }
#pragma unroll
  for (int j = 0; j < kKdeQpt; ++j) {
    const int64_t r = row0 + static_cast<int64_t>(j) * kKdeThreads;
    if (r < n_rows) out[r] = res[j];
  }
}

// Generic dims (runtime dp/dx): one query per thread, stored points read through L1.
__global__ void __launch_bounds__(128) kde_log_prob_generic_kernel(
    const float* __restrict__ tp, const float* __restrict__ ty, int64_t n_points, int dp, int dx,
    const float* __restrict__ qp, const float* __restrict__ qx, int64_t n_rows, float hp, float hy,
    float const_y, float log_n, float* __restrict__ out) {
  const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (r >= n_rows) return;
  Lse num, den;
  num.init();
  den.init();
  for (int64_t n = 0; n < n_points; ++n) {
    float q1 = 0.0f, q2 = 0.0f;
    for (int d = 0; d < dp; ++d) {
      const float df = __ldg(qp + r * dp + d) - __ldg(tp + n * dp + d);
      q1 = fmaf(df, df, q1);
    }
    for (int d = 0; d < dx; ++d) {
      const float df = __ldg(qx + r * dx + d) - __ldg(ty + n * dx + d);
      q2 = fmaf(df, df, q2);
    }
    const float a = -hp * q1;
    den.push(a);
    num.push(fmaf(-hy, q2, a));
  }
  out[r] = dp > 0 ? num.value() - den.value() + const_y : num.value() + const_y - log_n;
}

inline cudaError_t launch_kde_log_prob(const float* tp, const float* ty, int64_t n_points, int dp,
                                       int dx, const float* qp, const float* qx, int64_t n_rows,
                                       float bandwidth, float parent_bandwidth, float min_scale,
                                       float* out, cudaStream_t stream) {
  // scale = max(bandwidth, 1e-3) + min_scale   (kde.py:106)
  const double sy = (bandwidth > 1e-3f ? static_cast<double>(bandwidth) : 1e-3) + min_scale;
  const double sp = (parent_bandwidth > 1e-3f ? static_cast<double>(parent_bandwidth) : 1e-3) + min_scale;
  const double log2e = 1.4426950408889634;
  const double ln2pi = 1.8378770664093453;
  const float hy = static_cast<float>(0.5 / (sy * sy)), hp = static_cast<float>(0.5 / (sp * sp));
  const float hy2 = static_cast<float>(0.5 * log2e / (sy * sy));
  const float hp2 = static_cast<float>(0.5 * log2e / (sp * sp));
  const float const_y = static_cast<float>(-0.5 * dx * (ln2pi + 2.0 * log(sy)));
  const float log_n = static_cast<float>(log(static_cast<double>(n_points)));
  const int64_t per_cta = static_cast<int64_t>(kKdeThreads) * kKdeQpt;
  const unsigned grid = static_cast<unsigned>((n_rows + per_cta - 1) / per_cta);
  // fraction of exponentials moved to the FMA pipe; VBN_KDE_POLY=0..4 overrides (dev knob)
  const char* env = std::getenv("VBN_KDE_POLY");
  const int poly = env ? std::atoi(env) : 1;
#define VBN_KDE_LAUNCH(DP_, DX_, PN_, PD_)                                                      \
  kde_log_prob_kernel<DP_, DX_, PN_, PD_><<<grid, kKdeThreads, 0, stream>>>(                    \
      tp, ty, n_points, qp, qx, n_rows, hp2, hy2, const_y, log_n, out)
#define VBN_KDE_CASE(DP_, DX_)                                                                  \
  if (dp == DP_ && dx == DX_) {                                                                 \
    if (poly <= 0) VBN_KDE_LAUNCH(DP_, DX_, 0, 0);                                              \
    else if (poly == 1 || DP_ == 0) VBN_KDE_LAUNCH(DP_, DX_, 1, 0);                             \
    else if (poly == 2) VBN_KDE_LAUNCH(DP_, DX_, 1, 1);                                         \
    else if (poly == 3) VBN_KDE_LAUNCH(DP_, DX_, 2, 1);                                         \
    else VBN_KDE_LAUNCH(DP_, DX_, 2, 2);                                                        \
    return cudaGetLastError();                                                                  \
  }
  VBN_KDE_CASE(0, 1)
  VBN_KDE_CASE(1, 1)
  VBN_KDE_CASE(2, 1)
  VBN_KDE_CASE(3, 1)
  VBN_KDE_CASE(1, 2)
  VBN_KDE_CASE(2, 2)
#undef VBN_KDE_CASE
#undef VBN_KDE_LAUNCH
  const unsigned g2 = static_cast<unsigned>((n_rows + 127) / 128);
  kde_log_prob_generic_kernel<<<g2, 128, 0, stream>>>(tp, ty, n_points, dp, dx, qp, qx, n_rows, hp,
                                                      hy, const_y, log_n, out);
  return cudaGetLastError();
}

}  // namespace vbn
