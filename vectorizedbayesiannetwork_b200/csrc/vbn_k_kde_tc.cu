// KDE conditional log-density for high-dimensional points on the tensor cores (sm_100a: tcgen05 + TMEM + bulk copy).
//
//   out[m] = LSE_n(log_kp[m,n] + log_ky[m,n]) - LSE_n(log_kp[m,n])        (vbn/cpds/kde.py:111-149)
//
// The FP32-pipe kernel (vbn_kde.cuh) spends 3 (Dp + Dx) / 2 packed issue slots per (query, point) pair on the squared
// distances; from Dp + Dx ~ 8 on that, not the two exponentials, is the bound.  Here the pairwise term is a GEMM:
//
//   a[m,n] = -h_p |xp_m - p_n|^2 = (sqrt(2 h_p) xp_m) . (sqrt(2 h_p) p_n) - h_p |xp_m|^2 - h_p |p_n|^2
//
// i.e. one dot product of the augmented rows  A_m = [sqrt(2h) xp_m, -h |xp_m|^2, 1]  and  B_n = [sqrt(2h) p_n, 1,
// -h |p_n|^2]  (h in the log2 domain, so the accumulator IS the exp2 argument); the target part likewise.  A CTA owns
// 128 query rows (= TMEM lanes): their augmented rows sit in shared memory as K-major core-matrix images for the
// whole kernel, tiles of 64 stored points (prepacked the same way by kde_pack_kernel) stream through a two-slot
// cp.async.bulk ring, tcgen05.mma.kind::tf32 (A and B from shared memory) leaves a[128 x 64] and y[128 x 64] in TMEM,
// and every thread folds its row's 64 + 64 values into two streaming sums with ex2.approx -- the same fixed-shift
// accumulation as the FP32 kernel (every term is <= 0), two-level (per tile, then across tiles).
//
// fp32 parity: x.p suffers the cancellation |x|^2 + |p|^2 - 2 x.p, so the products must be exact to fp32 level.
// Every operand is split into THREE tf32 pieces (11 + 11 + 2 bits: exact) and the six significant piece products
// (1,1) (1,2) (2,1) (1,3) (2,2) (3,1) are accumulated in fp32: ~2^-30 relative per product.
//
// What fp32 accumulation cannot fix: the accumulator holds values of size h (|x|^2 + |p|^2) while the answer needs
// an absolute 1e-5 on exponents of size ~10, so the form is only as good as 2^-23 h |x|^2.  Both clouds are centred
// on the stored points' mean (`center`), and a query row whose scaled squared norm exceeds kMaxRowNorm (its error
// estimate passes ~5e-6) is not trusted to the GEMM: like the rows whose sums underflow 2^-100 (queries ~12 kernel
// widths from every stored point) it is written as NaN and redone by kde_fixup_kernel with direct differences and
// the exact online-max accumulation.  Several CTAs share an SM (128 TMEM columns each), so
// one CTA's MMAs overlap another's exponentials.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "vbn_launch.h"
#include "vbn_schedule_tc.cuh"

namespace vbn {
namespace kdetc {

using namespace vbn::tc;

constexpr int kRows = 128;     // query rows per CTA
constexpr int kTileN = 64;     // stored points per tile
constexpr int kThreads = 128;
constexpr int kPieces = 3;
constexpr float kMaxRowNorm = 40.0f;  // h |x - center|^2 (log2 units) above which a row takes the exact path

__host__ __device__ __forceinline__ int pad8(int x) { return (x + 7) & ~7; }
// floats of one [rows][k] K-major core-matrix image
__host__ __device__ __forceinline__ int image_floats(int rows, int k) { return rows * k; }
// element (r, c) of a [rows][k] image: core matrices of 8 rows x 4 floats, K-adjacent ones 32 floats apart,
// 8-row groups k * 8 floats apart
__host__ __device__ __forceinline__ int image_index(int r, int c, int k) {
  return ((r >> 3) * (k >> 2) + (c >> 2)) * 32 + (r & 7) * 4 + (c & 3);
}
// bytes of one packed tile of stored points: parent pieces then target pieces
__host__ __device__ __forceinline__ size_t tile_bytes(int kp, int ky) {
  return static_cast<size_t>(kPieces) * kTileN * (kp + ky) * sizeof(float);
}

__device__ __forceinline__ void split3(float x, float (&p)[3]) {
  uint32_t h1 = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  const float r1 = x - __uint_as_float(h1);
  uint32_t h2 = (__float_as_uint(r1) + 0x1000u) & 0xFFFFE000u;
  p[0] = __uint_as_float(h1);
  p[1] = __uint_as_float(h2);
  p[2] = r1 - __uint_as_float(h2);  // exact: at most the last two bits of the mantissa
}

// Augmented, scaled, split stored points -> tile images.  One thread per (tile row, tile).
// image layout per tile: [piece][parent image 64 x kp] then [piece][target image 64 x ky]
__global__ void __launch_bounds__(64) kde_pack_kernel(const float* __restrict__ tp, const float* __restrict__ ty,
                                                      int64_t n_points, int dp, int dx, int kp, int ky, float hp2,
                                                      float hy2, const float* __restrict__ center,
                                                      float* __restrict__ image) {
  const int64_t tile = blockIdx.x;
  const int r = threadIdx.x;
  const int64_t n = tile * kTileN + r;
  float* base = image + tile * (tile_bytes(kp, ky) / sizeof(float));
  const bool live = n < n_points;
  const float sp = sqrtf(2.0f * hp2), sy = sqrtf(2.0f * hy2);
  for (int part = 0; part < 2; ++part) {
    const int k = part == 0 ? kp : ky, d = part == 0 ? dp : dx;
    if (k == 0) continue;
    const float* src = part == 0 ? tp : ty;
    const float s = part == 0 ? sp : sy, h = part == 0 ? hp2 : hy2;
    float* img = base + (part == 0 ? 0 : kPieces * image_floats(kTileN, kp));
    double norm = 0.0;
    for (int c = 0; c < k; ++c) {
      float v = 0.0f;
      if (c < d) {
        const float ctr = center ? __ldg(center + (part == 0 ? 0 : dp) + c) : 0.0f;
        const float x = live ? __ldg(src + n * d + c) - ctr : 0.0f;
        v = s * x;
        norm += static_cast<double>(v) * v;  // |s x|^2 = 2 h |x|^2
      } else if (c == d) {
        v = 1.0f;  // pairs with the query's -h |x|^2
      } else if (c == d + 1) {
        // -h |p|^2 (= -|s p|^2 / 2); a row past the end of the data gets a huge negative constant: exp2 -> 0
        v = live ? static_cast<float>(-0.5 * norm) : -1.0e30f;
      }
      float pc[3];
      split3(v, pc);
#pragma unroll
      for (int q = 0; q < kPieces; ++q) img[q * image_floats(kTileN, k) + image_index(r, c, k)] = pc[q];
    }
  }
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// one part (parent or target) of one tile: D = sum over K steps and the six piece pairs
__device__ __forceinline__ void issue_part(uint32_t d, uint32_t a_img, uint32_t b_img, int k, uint32_t idesc) {
  const uint32_t a_piece = static_cast<uint32_t>(image_floats(kRows, k)) * 4u;
  const uint32_t b_piece = static_cast<uint32_t>(image_floats(kTileN, k)) * 4u;
  const uint32_t sbo = static_cast<uint32_t>(k) * 32u;
  uint32_t acc = 0u;
  for (int s = 0; s < (k >> 3); ++s) {
    // piece pairs (a, b): (1,1) (1,2) (2,1) (1,3) (2,2) (3,1), smallest contributions first
    const int pa[6] = {2, 1, 0, 1, 0, 0}, pb[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      const uint64_t ad = make_b_desc(a_img + pa[q] * a_piece, 128u, sbo) + 16u * s;
      const uint64_t bd = make_b_desc(b_img + pb[q] * b_piece, 128u, sbo) + 16u * s;
      mma_tf32_ss(d, ad, bd, idesc, acc);
      acc = 1u;
    }
  }
}

// dynamic smem: [ctrl 64 B][A images: 3 x 128 x (kp + ky) floats][2 ring slots of one tile each]
__global__ void __launch_bounds__(kThreads) kde_tc_kernel(const float* __restrict__ image, int64_t n_points,
                                                          int dp, int dx, int kp, int ky,
                                                          const float* __restrict__ qp, const float* __restrict__ qx,
                                                          int64_t n_rows, float hp2, float hy2, float const_y,
                                                          float log_n, const float* __restrict__ center,
                                                          float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t full_bar = smem_base + 16, mma_bar = smem_base + 32;
  const uint32_t a_base = smem_base + 64;
  const uint32_t a_bytes = static_cast<uint32_t>(kPieces * kRows * (kp + ky)) * 4u;
  const uint32_t t_bytes = static_cast<uint32_t>(tile_bytes(kp, ky));
  const uint32_t ring = a_base + a_bytes;
  float* a_img = reinterpret_cast<float*>(smem_raw + 64);

  if (tid == 0) {
    mbar_init(full_bar, 1);
    mbar_init(full_bar + 8, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_base, 128);

  // ---- this thread's query row -> augmented, scaled, split images in shared memory
  int64_t row = static_cast<int64_t>(blockIdx.x) * kRows + tid;
  const bool live = row < n_rows;
  if (!live) row = n_rows - 1;
  const float sp = sqrtf(2.0f * hp2), sy = sqrtf(2.0f * hy2);
  float row_norm = 0.0f;  // h |x - center|^2 over both parts: the size of the numbers the accumulators will hold
  for (int part = 0; part < 2; ++part) {
    const int k = part == 0 ? kp : ky, d = part == 0 ? dp : dx;
    if (k == 0) continue;
    const float* src = part == 0 ? qp : qx;
    const float s = part == 0 ? sp : sy;
    float* img = a_img + (part == 0 ? 0 : kPieces * image_floats(kRows, kp));
    double norm = 0.0;
    for (int c = 0; c < k; ++c) {
      float v = 0.0f;
      if (c < d) {
        const float ctr = center ? __ldg(center + (part == 0 ? 0 : dp) + c) : 0.0f;
        v = s * (__ldg(src + row * d + c) - ctr);
        norm += static_cast<double>(v) * v;
      } else if (c == d) {
        row_norm += static_cast<float>(0.5 * norm);
        v = static_cast<float>(-0.5 * norm);  // -h |x|^2, pairs with the point's 1
      } else if (c == d + 1) {
        v = 1.0f;                             // pairs with the point's -h |p|^2
      }
      float pc[3];
      split3(v, pc);
#pragma unroll
      for (int q = 0; q < kPieces; ++q) img[q * image_floats(kRows, k) + image_index(tid, c, k)] = pc[q];
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw);
  const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const uint32_t idesc = make_idesc(kTileN);
  const int64_t n_tiles = (n_points + kTileN - 1) / kTileN;
  const unsigned char* img_bytes = reinterpret_cast<const unsigned char*>(image);

  if (tid == 0) {
    for (int t = 0; t < 2 && t < n_tiles; ++t) {
      mbar_expect_tx(full_bar + 8 * t, t_bytes);
      bulk_g2s(ring + t * t_bytes, img_bytes + static_cast<size_t>(t) * t_bytes, t_bytes, full_bar + 8 * t);
    }
  }

  float den_tot = 0.0f, num_tot = 0.0f;
  uint32_t mma_phase = 0;
  for (int64_t t = 0; t < n_tiles; ++t) {
    const uint32_t buf = static_cast<uint32_t>(t & 1);
    if (tid == 0) {
      mbar_wait(full_bar + 8 * buf, static_cast<uint32_t>((t >> 1) & 1));
      tc_fence_after();
      const uint32_t b_img = ring + buf * t_bytes;
      if (kp > 0) issue_part(tmem, a_base, b_img, kp, idesc);
      issue_part(tmem + kTileN, a_base + kPieces * image_floats(kRows, kp) * 4u,
                 b_img + kPieces * image_floats(kTileN, kp) * 4u, ky, idesc);
      mma_commit(mma_bar);
    }
    mbar_wait(mma_bar, mma_phase);
    mma_phase ^= 1u;
    tc_fence_after();
    if (tid == 0 && t + 2 < n_tiles) {  // the ring slot is free again (its MMAs have completed)
      mbar_expect_tx(full_bar + 8 * buf, t_bytes);
      bulk_g2s(ring + buf * t_bytes, img_bytes + static_cast<size_t>(t + 2) * t_bytes, t_bytes, full_bar + 8 * buf);
    }
    // ---- epilogue: this thread's row, 64 points
    float den_t = 0.0f, num_t = 0.0f;
#pragma unroll
    for (int c0 = 0; c0 < kTileN; c0 += 16) {
      uint32_t ya[16], pa[16];
      tmem_ld16(t_row + kTileN + c0, ya);
      if (kp > 0) tmem_ld16(t_row + c0, pa);
      tmem_wait_ld();
      float ds = 0.0f, ns = 0.0f;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const float a = kp > 0 ? __uint_as_float(pa[q]) : 0.0f;
        const float cterm = a + __uint_as_float(ya[q]);
        float ea = 0.0f, ec;
        if (kp > 0) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(a));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ec) : "f"(cterm));
        ds += ea;
        ns += ec;
      }
      den_t += ds;
      num_t += ns;
    }
    den_tot += den_t;
    num_tot += num_t;
    tc_fence_before();
    __syncthreads();  // every warp has read this tile's accumulators: the next tile's MMAs may overwrite them
  }
  if (live) {
    constexpr float kLn2 = 0.6931471805599453f;
    const bool under = !(num_tot > 7.888609e-31f) || (kp > 0 && !(den_tot > 7.888609e-31f))  // 2^-100
                       || row_norm > kMaxRowNorm;
    float res;
    if (under) {
      res = CUDART_NAN_F;  // redone by kde_fixup_kernel with the exact online-max accumulation
    } else if (kp > 0) {
      res = (log2f(num_tot) - log2f(den_tot)) * kLn2 + const_y;
    } else {
      res = log2f(num_tot) * kLn2 + const_y - log_n;
    }
    out[row] = res;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// Rows the fast pass flagged (NaN): direct differences and the exact streaming logsumexp.  One 128-thread block per
// row (blocks of unflagged rows leave at once); the threads share the stored points and merge their (max, sum) pairs.
__device__ __forceinline__ Lse lse_shfl(const Lse& a, int o) {
  Lse r;
  r.m = __shfl_xor_sync(0xffffffffu, a.m, o);
  r.l = __shfl_xor_sync(0xffffffffu, a.l, o);
  return r;
}
__global__ void __launch_bounds__(128) kde_fixup_kernel(const float* __restrict__ tp, const float* __restrict__ ty,
                                                        int64_t n_points, int dp, int dx, const float* __restrict__ qp,
                                                        const float* __restrict__ qx, int64_t n_rows, float hp, float hy,
                                                        float const_y, float log_n, float* __restrict__ out) {
  const int64_t r = blockIdx.x;
  if (!isnan(out[r])) return;  // block-uniform
  Lse num, den;
  num.init();
  den.init();
  for (int64_t n = threadIdx.x; n < n_points; n += blockDim.x) {
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num.merge(lse_shfl(num, o));
    den.merge(lse_shfl(den, o));
  }
  __shared__ Lse sh_num[4], sh_den[4];
  if ((threadIdx.x & 31) == 0) {
    sh_num[threadIdx.x >> 5] = num;
    sh_den[threadIdx.x >> 5] = den;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 4; ++w) {
      num.merge(sh_num[w]);
      den.merge(sh_den[w]);
    }
    out[r] = dp > 0 ? num.value() - den.value() + const_y : num.value() + const_y - log_n;
  }
}

}  // namespace kdetc

size_t kde_tc_workspace_bytes(int64_t n_points, int dp, int dx) {
  const int kp = dp > 0 ? kdetc::pad8(dp + 2) : 0, ky = kdetc::pad8(dx + 2);
  const int64_t n_tiles = (n_points + kdetc::kTileN - 1) / kdetc::kTileN;
  return static_cast<size_t>(n_tiles) * kdetc::tile_bytes(kp, ky);
}

bool kde_tc_supported(int dp, int dx) {
  const int kp = dp > 0 ? kdetc::pad8(dp + 2) : 0, ky = kdetc::pad8(dx + 2);
  const size_t smem = 64 + static_cast<size_t>(kdetc::kPieces) * kdetc::kRows * (kp + ky) * 4 + 2 * kdetc::tile_bytes(kp, ky);
  return dx >= 1 && smem <= 200 * 1024;
}

cudaError_t launch_kde_log_prob_tc(const float* tp, const float* ty, int64_t n_points, int dp, int dx, const float* qp,
                                   const float* qx, int64_t n_rows, float bandwidth, float parent_bandwidth,
                                   float min_scale, const float* center, float* workspace, float* out,
                                   cudaStream_t stream) {
  using namespace kdetc;
  const double sy = (bandwidth > 1e-3f ? static_cast<double>(bandwidth) : 1e-3) + min_scale;  // kde.py:106
  const double sp = (parent_bandwidth > 1e-3f ? static_cast<double>(parent_bandwidth) : 1e-3) + min_scale;
  const double log2e = 1.4426950408889634, ln2pi = 1.8378770664093453;
  const float hy = static_cast<float>(0.5 / (sy * sy)), hp = static_cast<float>(0.5 / (sp * sp));
  const float hy2 = static_cast<float>(0.5 * log2e / (sy * sy)), hp2 = static_cast<float>(0.5 * log2e / (sp * sp));
  const float const_y = static_cast<float>(-0.5 * dx * (ln2pi + 2.0 * log(sy)));
  const float log_n = static_cast<float>(log(static_cast<double>(n_points)));
  const int kp = dp > 0 ? pad8(dp + 2) : 0, ky = pad8(dx + 2);
  const int64_t n_tiles = (n_points + kTileN - 1) / kTileN;
  kde_pack_kernel<<<static_cast<unsigned>(n_tiles), kTileN, 0, stream>>>(tp, ty, n_points, dp, dx, kp, ky, hp2, hy2,
                                                                         center, workspace);
  const size_t smem = 64 + static_cast<size_t>(kPieces) * kRows * (kp + ky) * 4 + 2 * tile_bytes(kp, ky);
  cudaError_t e = cudaFuncSetAttribute(reinterpret_cast<const void*>(&kde_tc_kernel),
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  const unsigned grid = static_cast<unsigned>((n_rows + kRows - 1) / kRows);
  kde_tc_kernel<<<grid, kThreads, smem, stream>>>(workspace, n_points, dp, dx, kp, ky, qp, qx, n_rows, hp2, hy2,
                                                  const_y, log_n, center, out);
  kde_fixup_kernel<<<static_cast<unsigned>(n_rows), 128, 0, stream>>>(tp, ty, n_points, dp, dx, qp, qx, n_rows, hp, hy,
                                                                    const_y, log_n, out);
  return cudaGetLastError();
}

}  // namespace vbn
