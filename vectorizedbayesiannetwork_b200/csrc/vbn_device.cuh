// Device-side building blocks shared by the schedule kernel and the stand-alone kernels.
// sm_100a only.  No host code here.
#pragma once
#ifndef VBN_HOST_EMU  // tests/emu/cuda_shim.h supplies these for the host-emulation test build
#include <cuda_runtime.h>
#include <math_constants.h>
#endif
#include <stdint.h>

#include "vbn_cuda.h"

namespace vbn {

constexpr float kLog2Pi = 1.8378770664093453f;      // ln(2*pi)
constexpr float kHalfLog2Pi = 0.9189385332046727f;  // ln(sqrt(2*pi))

// ---------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (Salmon et al., SC'11).  Counter layout used by the schedule
// kernel: (c0, c1, c2, c3) = (global sample s, global query b | 0xFFFFFFFF when the draw is
// shared by all queries, stream block index | stream tag << 30, call offset).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    // one 32x32->64 multiply per lane pair (IMAD.WIDE.U32) instead of separate hi / lo products
    const uint64_t p0 = static_cast<uint64_t>(M0) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(M1) * c.z;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// Same generator with the ten round keys supplied precomputed (host: key + i * Weyl).  Inlined
// where `rk` is a kernel parameter, the keys become constant-bank operands of the LOP3s and the
// twenty key-schedule adds disappear.
__device__ __forceinline__ uint4 philox4x32_10_rk(uint4 c, const uint32_t (&rk)[20]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = static_cast<uint64_t>(M0) * c.x;
    const uint64_t p1 = static_cast<uint64_t>(M1) * c.z;
    const uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
    c = make_uint4(hi1 ^ c.y ^ rk[2 * i], lo1, hi0 ^ c.w ^ rk[2 * i + 1], lo0);
  }
  return c;
}

// uint32 -> float in (0,1): 24 random bits, centred, never 0 or 1.
__device__ __forceinline__ float u01(uint32_t x) {
  // (N + 0.5) 2^-24 with N = x >> 8: one FMA -- N 2^-24 + 2^-25 is rounded once, exactly like N + 0.5 (scaling by a
  // power of two commutes with the rounding), so the bits are those of the two-step form
  return fmaf(static_cast<float>(x >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f);
}

// Two uniforms -> two standard normals (Box-Muller).  Fast intrinsics: these only shape the
// generated noise (no parity requirement on generated bits, only on the law).
#ifdef VBN_HOST_EMU
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = u01(a);
  const float u2 = u01(b) - 0.5f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}
#else
// Device version, trimmed to the instruction count that matters for LG-only schedules (one normal
// per row-node): mantissa-stuffing instead of int->float conversions, raw lg2 / sqrt approximations
// (u1 >= 2^-24 is never denormal), angle produced by one FFMA.  ~15 instructions per pair.
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float f1 = __uint_as_float((a >> 9) | 0x3F800000u);  // [1, 2), 23 random bits
  const float f2 = __uint_as_float((b >> 9) | 0x3F800000u);
  const float u1 = f1 - 0.99999994f;                         // (0, 1): 2^-24 .. 1 - 2^-24
  const float ang = fmaf(f2, 6.283185307179586f, -9.42477796076938f);  // 2 pi (f2 - 1.5) in [-pi, pi)
  float l, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -1.3862943611198906f));  // sqrt(-2 ln u1)
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(ang));
  return make_float2(r * c, r * s);
}
#endif

__device__ __forceinline__ float4 normal4(uint4 w) {
  const float2 a = box_muller(w.x, w.y);
  const float2 b = box_muller(w.z, w.w);
  return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ float4 uniform4(uint4 w) {
  return make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
}

// Out-of-line Philox + transform: the schedule kernels call these from many op bodies; keeping one
// copy keeps the hot loop inside the instruction cache.
static __device__ __noinline__ float4 philox_normal4(uint4 c, uint2 k) { return normal4(philox4x32_10(c, k)); }
static __device__ __noinline__ float4 philox_uniform4(uint4 c, uint2 k) { return uniform4(philox4x32_10(c, k)); }

__device__ __forceinline__ float lane4(const float4& v, int lane) {
  // two levels of selects on the two index bits (the chained comparison form compiled to branches)
  const float lo = (lane & 1) ? v.y : v.x;
  const float hi = (lane & 1) ? v.w : v.z;
  return (lane & 2) ? hi : lo;
}

// ---------------------------------------------------------------------------------------
// scalar math that mirrors the torch ops the reference uses
// ---------------------------------------------------------------------------------------
// F.softplus(x) (beta=1, threshold=20)  -- vbn/cpds/utils.py:6-7
__device__ __forceinline__ float softplus20(float x) { return x > 20.0f ? x : log1pf(expf(x)); }

// Same function for values that only scale generated noise (drawn-only MDN nodes): branch-free,
// ~1e-7 relative.  softplus(x) = max(x, 0) + log1p(e), e = exp(-|x|) in (0, 1];
// log1p(e) = 2 atanh(s), s = e / (2 + e) <= 1/3, odd series to s^15 (next term < 2e-9 relative).
// For x > 20 the series term is < 2.1e-9, i.e. the result is x in fp32, as with the threshold form.
// 2^x, flush-to-zero form: one multiply-free MUFU (the default form brackets it with a range test and two
// predicated multiplies for results below 2^-126, which the callers here clamp or add to >= 1 anyway)
__device__ __forceinline__ float ex2_ftz(float x) {
#ifdef VBN_HOST_EMU
  return exp2f(x);
#else
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#endif
}
__device__ __forceinline__ float softplus20_fast(float x) {
  const float e = ex2_ftz(fabsf(x) * -1.4426950408889634f);
  const float s = __fdividef(e, 2.0f + e);
  const float z = s * s;
  float p = 1.0f / 15.0f;
  p = fmaf(p, z, 1.0f / 13.0f);
  p = fmaf(p, z, 1.0f / 11.0f);
  p = fmaf(p, z, 1.0f / 9.0f);
  p = fmaf(p, z, 1.0f / 7.0f);
  p = fmaf(p, z, 1.0f / 5.0f);
  p = fmaf(p, z, 1.0f / 3.0f);
  p = fmaf(p, z, 1.0f);
  return fmaf(2.0f * s, p, fmaxf(x, 0.0f));
}

static __device__ __noinline__ float activate_slow(float x, int act) {
  switch (act) {
    case VBN_ACT_TANH: return tanhf(x);
    case VBN_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));  // nn.GELU (erf)
    default: return x > 0.0f ? x : expm1f(x);                                      // nn.ELU(alpha=1)
  }
}
__device__ __forceinline__ float activate(float x, int act) {
  return act == VBN_ACT_RELU ? fmaxf(x, 0.0f) : activate_slow(x, act);
}

// -0.5*((x-loc)^2/var + 2 ln sigma + ln 2pi) for one dim (linear_gaussian.py:214-217)
__device__ __forceinline__ float gauss_term(float x, float loc, float var, float two_log_scale) {
  const float d = x - loc;
  return __fdiv_rn(d * d, var) + two_log_scale + kLog2Pi;
}

// streaming logsumexp state
struct Lse {
  float m, l;
  __device__ __forceinline__ void init() { m = -CUDART_INF_F; l = 0.0f; }
  __device__ __forceinline__ void push(float x) {
    if (x == -CUDART_INF_F) return;  // contributes exp(-inf) = 0; the update below would form exp(-inf - -inf) = NaN
    if (x > m) {
      l = l * __expf(m - x) + 1.0f;
      m = x;
    } else {
      l += __expf(x - m);
    }
  }
  __device__ __forceinline__ void merge(const Lse& o) {
    if (o.m == -CUDART_INF_F) return;
    if (o.m > m) {
      l = l * __expf(m - o.m) + o.l;
      m = o.m;
    } else {
      l += o.l * __expf(o.m - m);
    }
  }
  __device__ __forceinline__ float value() const { return m + logf(l); }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

}  // namespace vbn
