// TEST INFRASTRUCTURE ONLY: lets g++ compile the schedule kernel's device code for the host so
// its logic can be checked against the oracle in a container without a GPU.  Threads of a CTA run
// one after another (valid because the schedule kernel has no inter-thread communication).
// Nothing in vectorizedbayesiannetwork_b200/ uses this; the product path is the CUDA build.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

#define VBN_HOST_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__

struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
struct int2 { int32_t x, y; };
struct int4 { int32_t x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x = 1, y = 1, z = 1; };
inline uint2 make_uint2(uint32_t a, uint32_t b) { return {a, b}; }
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {a, b, c, d}; }
inline float2 make_float2(float a, float b) { return {a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }

extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define CUDART_INF_F std::numeric_limits<float>::infinity()

template <typename T> inline T __ldg(const T* p) { return *p; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
inline float __logf(float x) { return std::log(x); }
inline float __expf(float x) { return std::exp(x); }
inline void __sincosf(float x, float* s, float* c) { *s = std::sin(x); *c = std::cos(x); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fdividef(float a, float b) { return a / b; }
inline int __float2int_rz(float x) { return (x != x || x >= 2147483648.0f || x < -2147483648.0f) ? 0 : static_cast<int>(x); }
inline float fmaf_(float a, float b, float c) { return std::fma(a, b, c); }
inline int atomicOr(int32_t* p, int v) { int o = *p; *p |= v; return o; }
using std::min;
using std::max;
inline float __int_as_float(int32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int32_t __float_as_int(float f) { int32_t v; std::memcpy(&v, &f, 4); return v; }
