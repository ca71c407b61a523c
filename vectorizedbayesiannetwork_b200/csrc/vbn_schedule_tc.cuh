// Tensor-core variant of the fused schedule kernel (sm_100a: tcgen05 + TMEM + bulk async copy).
//
// Same schedule walk as schedule_kernel (vbn_schedule.cuh), but the dense contractions of every
// [Dp -> 32 -> 32 -> O] MLP CPD (gaussian_nn.py:16-34 _build_mlp, evaluated at gaussian_nn.py:235, mdn.py:199,
// softmax_nn.py:585) run on the 5th-gen tensor cores:
//
//   * one CTA = NWG warpgroups, one CTA per SM, persistent; every warpgroup owns RPT 128-row tiles
//     (<4,1> or <2,2>: 512 rows per CTA pass either way);
//   * a row is a TMEM lane: thread t of a warpgroup owns row t of each of its tiles for the whole DAG walk, so
//     hidden activations move registers <-> TMEM with tcgen05.st / tcgen05.ld (32x32b shapes) and never touch
//     shared memory or HBM;
//   * activations are the A operand, read by tcgen05.mma straight from TMEM; the weights are the B operand,
//     K-major core-matrix images prepacked by the host (cpds.py pack_mlp_tc) and streamed L2 -> shared memory
//     with cp.async.bulk through a ring of mbarrier-guarded buffers; one elected lane of one warp tops the ring
//     up (non-blocking look-ahead) each time it enters an MLP, so no warp -- and no register allocation -- is
//     spent on a dedicated producer;
//   * a first layer with <= 4 inputs (K would be padded 2 -> 8 for 64 useful MACs per row) runs as 16 Dp packed
//     FMAs per row straight from a plain copy of W1 / b1 at the head of the image: one MMA round trip less;
//   * fp32 parity (1e-5) rules out plain TF32, so every product is the error-compensated
//     3xTF32 split  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi  (hi = RN-to-tf32, lo = exact
//     remainder), accumulated in fp32 in TMEM: ~2^-21 relative per product;
//   * each warpgroup issues its own MMAs (one elected thread of warp `wg & 3`, so the issuing warps of a CTA
//     sit on different SM sub-partitions) and waits on per-tile mbarriers (tcgen05.commit); tiles drift freely
//     and hide each other's MMA / TMEM latency; the only CTA-wide coupling is the weight ring.
//
// TMEM map (512 columns allocated, 96 per tile): D fp32 accumulator [0,32) | A_hi [32,64) | A_lo [64,96); then the
// 8-column constant block (1, 0, ..., 0) of the bias step.
#pragma once
#include "vbn_schedule.cuh"
#include "vbn_tc_layout.h"

namespace vbn {
namespace tc {

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware, do not spin
      : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 16 bytes of shared memory (warp-uniform address: one broadcast wavefront) as two packed f32x2 operands
__device__ __forceinline__ void lds_f32x4_as_x2(uint32_t addr, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem descriptor], kind::tf32, issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 B (128 B
// contiguous); LBO = bytes between the two K-adjacent core matrices of one MMA (K = 8 tf32),
// SBO = bytes between 8-row groups.  Bits: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout_type=0 (SWIZZLE_NONE) [61,64).
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) |
         (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major, M = 128.
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24);
}

// x = hi + lo with hi = x rounded to nearest tf32 (10 explicit mantissa bits), lo the exact
// remainder (the tensor core reads the top 19 bits of lo: residual <= 2^-21 |x|).  Truncating hi
// instead saves one instruction per element but its one-sided 2^-20 residual fails the 2e-5
// log_prob parity of narrow fitted MDN components (readme golden), so round.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
// Two activations at once with packed fp32x2 arithmetic: c = p * 2^13 (exact), t = fl(p + c) rounds
// p's low 13 bits away (round to nearest), hi = t - c (exact) is p with 11 significant bits -- a tf32
// value -- and lo = p - hi exactly.  4 packed instructions per PAIR instead of 3 scalar ones per
// element.  Written so that an FMA contraction of any mul+add pair cannot change a value (the classic
// Veltkamp form t = p*(2^13+1); hi = t - (t - p) was contracted by ptxas and lost the split).
__device__ __forceinline__ void split_tf32_x2(float a, float b, uint32_t& hi_a, uint32_t& hi_b, uint32_t& lo_a,
                                              uint32_t& lo_b) {
  unsigned long long p, c, t, hi, lo;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
  asm("{\n\t"
      ".reg .b64 k;\n\t"
      "mov.b64 k, {%5, %5};\n\t"
      "mul.rn.f32x2 %0, %4, k;\n\t"
      "add.rn.f32x2 %1, %4, %0;\n\t"
      "sub.rn.f32x2 %2, %1, %0;\n\t"
      "sub.rn.f32x2 %3, %4, %2;\n\t"
      "}"
      : "=&l"(c), "=&l"(t), "=&l"(hi), "=&l"(lo)
      : "l"(p), "f"(8192.0f));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(hi_a), "=r"(hi_b) : "l"(hi));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo_a), "=r"(lo_b) : "l"(lo));
}
// ReLU + split of two pre-activations: q = p + |p| = 2 relu(p) is ONE packed add (the |.| is an operand
// modifier of FADD2) instead of two FMNMX; the factor 2 is exact and the host halves the weights of the layer that
// consumes these activations (cpds.pack_mlp_tc relu2), so every product -- and the accumulator -- is bit-identical
// to relu(p) * w.  NaN pre-activations stay NaN (torch.relu semantics; fmaxf would have returned 0).
__device__ __forceinline__ void split_relu2_tf32_x2(float a, float b, uint32_t& hi_a, uint32_t& hi_b, uint32_t& lo_a,
                                                    uint32_t& lo_b) {
  unsigned long long p, m, q, c, t, hi, lo;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b));
  asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(fabsf(a)), "f"(fabsf(b)));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(p), "l"(m));
  asm("{\n\t"
      ".reg .b64 k;\n\t"
      "mov.b64 k, {%5, %5};\n\t"
      "mul.rn.f32x2 %0, %4, k;\n\t"
      "add.rn.f32x2 %1, %4, %0;\n\t"
      "sub.rn.f32x2 %2, %1, %0;\n\t"
      "sub.rn.f32x2 %3, %4, %2;\n\t"
      "}"
      : "=&l"(c), "=&l"(t), "=&l"(hi), "=&l"(lo)
      : "l"(q), "f"(8192.0f));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(hi_a), "=r"(hi_b) : "l"(hi));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo_a), "=r"(lo_b) : "l"(lo));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// 3xTF32 layer: D = bias + A_lo*B_hi + A_hi*B_lo + A_hi*B_hi over K (multiple of 8), N columns.
// B images are [N][K+8] K-major core-matrix layouts (see pack_mlp_tc): LBO = 128, SBO = (K+8)*32; the last
// K = 8 step of the image is the bias column.  It is multiplied by the constant A block `ones` = (1, 0, .., 0)
// with accumulate = 0, so the layer needs no accumulator preload and its epilogue no bias add (the lo part of
// 1.0 is zero: two MMAs).  KS = number of K = 8 steps of the weights proper (compile time so the issue
// sequence is straight-line code on the uniform datapath); one step covers two core matrices = 256 B = 16
// descriptor units.
template <int KS>
__device__ __forceinline__ void issue_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t ones, uint32_t b_hi,
                                            uint32_t b_lo, int n) {
  const uint32_t idesc = make_idesc(n);
  const uint32_t sbo = static_cast<uint32_t>(KS + 1) * 256u;
  const uint64_t dhi = make_b_desc(b_hi, 128u, sbo), dlo = make_b_desc(b_lo, 128u, sbo);
  mma_tf32_ts(d, ones, dhi + 16 * KS, idesc, 0u);  // D = bias_hi
  mma_tf32_ts(d, ones, dlo + 16 * KS, idesc, 1u);  // D += bias_lo
#pragma unroll
  for (int s = 0; s < KS; ++s) mma_tf32_ts(d, a_lo + 8 * s, dhi + 16 * s, idesc, 1u);
#pragma unroll
  for (int s = 0; s < KS; ++s) mma_tf32_ts(d, a_hi + 8 * s, dlo + 16 * s, idesc, 1u);
#pragma unroll
  for (int s = 0; s < KS; ++s) mma_tf32_ts(d, a_hi + 8 * s, dhi + 16 * s, idesc, 1u);
}

// Per-thread tensor-core state, plugged into Ctx as the TC policy.
// NWG warpgroups per CTA, RPT 128-row tiles per warpgroup (thread t of a warpgroup owns row t of each of its
// tiles: TMEM lane t of RPT column regions).  <4,1>: four tiles drift against each other (thread-level
// parallelism hides the MMA round trips); <2,2>: each thread interleaves two rows -- tile 1's MMAs run while
// tile 0's accumulator is split, and the per-op control flow, descriptor fetch and weight loads are paid once
// for two rows.
template <int NWG, int RPT>
struct TcMlp {
  static constexpr bool kEnabled = true;
  static constexpr bool kLoops = false;      // Gibbs programs never carry tensor-core images
  static constexpr bool kInlineRng = true;   // inlined generator with constant-bank round keys
  static constexpr bool kTab = true;
  // Per-thread state is kept to a handful of OPAQUE values (see "values, not recipes" in the kernel): everything
  // else is an immediate offset from them, and the weight-ring producer's counters live in shared memory (one
  // lane touches them once per MLP) instead of a register in every thread.
  uint32_t t_d;        // TMEM address of tile 0's accumulator for this thread's warp (lane base folded in); tile j
                       // adds j * kColsPerTile, A_hi / A_lo add kColAhi / kColAlo
  uint32_t tm;         // TMEM base of the CTA's allocation
  uint32_t ctl;        // shared-space address of the control block: barriers at fixed offsets, weight ring behind it
  uint32_t role;       // wg | issuer << 8 | producer << 9 | (lane == 0) << 10 | (a ring slot is still held) << 11
  const int4* dptr;    // descriptor of the NEXT op of the walk (generic address): runs through the global op list, or
                       // through the descriptor tail of the ring slot held (re-pointed once per MLP, in layers_end)
  uint32_t w_buf;      // ring slot of the next MLP (w_iter % nbuf), and (w_iter / nbuf) & 1 in bit 8: running state
  uint32_t mma_phase;  // parity of the next completion of the mma barriers (all tiles of a thread move together)
  uint32_t nbuf;
  uint32_t slot_bytes; // size of one ring slot

  static constexpr uint32_t kFull = 16, kEmpty = kFull + 8 * kMaxBufs, kMma = kEmpty + 8 * kMaxBufs;
  static constexpr uint32_t kProd = kMma + 8 * kMaxTiles;  // producer words: p_iter, p_buf, p_idx, p_total, w_iter
  static_assert(kProd + 20 <= kCtrlBytes, "control block");

  __device__ __forceinline__ uint32_t td(int j) const { return t_d + static_cast<uint32_t>(j * kColsPerTile); }
  __device__ __forceinline__ uint32_t wg() const { return role & 0xFFu; }
  __device__ __forceinline__ uint32_t md(int j) const { return tm + (wg() * RPT + j) * kColsPerTile; }
  __device__ __forceinline__ uint32_t m_ones() const { return tm + NWG * RPT * kColsPerTile; }
  __device__ __forceinline__ uint32_t wbuf(uint32_t buf) const { return ctl + kCtrlBytes + buf * slot_bytes; }
  __device__ __forceinline__ uint32_t full_bar(uint32_t buf) const { return ctl + kFull + 8 * buf; }
  __device__ __forceinline__ uint32_t empty_bar(uint32_t buf) const { return ctl + kEmpty + 8 * buf; }
  __device__ __forceinline__ uint32_t mma_bar(int j) const { return ctl + kMma + 8 * (wg() * RPT + j); }
  // Descriptor of the next op, and step.  One running generic pointer serves both sources -- the global op list and
  // the copies that ride behind an MLP op's weights in the ring slot (plan.py: all ops up to and including the next
  // MLP op, or none) -- so an op costs one 64-bit add instead of a range test, a shared-window lookup (S2R) and an
  // address build.
  __device__ __forceinline__ const int4* next_desc() {
    const int4* d = dptr;
    dptr = d + 8;
    return d;
  }
  // gives the ring slot of the previous MLP back (its tail has been consumed) -- called when the next MLP starts and
  // at the end of the walk
  __device__ __forceinline__ void release_held() {
    if (role & 0x800u) {
      const uint32_t buf = w_buf & 0xFFu;
      const uint32_t ph = w_buf >> 8;
      __syncwarp();
      if (role & 0x400u) mbar_arrive(empty_bar(buf));
      w_buf = buf + 1u == nbuf ? ((ph ^ 1u) << 8) : w_buf + 1u;
      role &= ~0x800u;
    }
  }
  __device__ __forceinline__ bool issuer() const { return (role & 0x100u) != 0; }
  __device__ __forceinline__ bool producer() const { return (role & 0x200u) != 0; }

  // Issues every image whose ring slot is free, up to nbuf ahead of this warp; blocks only for the
  // image of the op this warp is about to run (its slot frees once the slowest warp has left op
  // w_iter - nbuf, which never depends on this warp).  One elected lane; its counters live in shared memory.
  __device__ __forceinline__ void produce(const ScheduleArgs& a) {
    const uint32_t st = ctl + kProd;
    uint32_t p_iter = lds_u32(st), p_buf = lds_u32(st + 4), p_idx = lds_u32(st + 8);
    const uint32_t p_total = lds_u32(st + 12), w_iter = lds_u32(st + 16);
    while (p_iter < p_total && p_iter < w_iter + nbuf) {
      const uint32_t buf = p_buf & 0xFFu;
      const uint32_t ph = p_buf >> 8;
      if (p_iter > w_iter) {
        if (!mbar_test(empty_bar(buf), ph)) break;
      } else {
        mbar_wait(empty_bar(buf), ph);
      }
      const int2 e = __ldg(a.tc_list + p_idx);
      mbar_expect_tx(full_bar(buf), static_cast<uint32_t>(e.y));
      bulk_g2s(wbuf(buf), a.params + e.x, static_cast<uint32_t>(e.y), full_bar(buf));
      ++p_iter;
      p_idx = p_idx + 1u == static_cast<uint32_t>(a.n_tc) ? 0u : p_idx + 1u;
      p_buf = buf + 1u == nbuf ? ((ph ^ 1u) << 8) : p_buf + 1u;
    }
    sts_u32(st, p_iter);
    sts_u32(st + 4, p_buf);
    sts_u32(st + 8, p_idx);
    sts_u32(st + 16, w_iter + 1u);  // this warp is entering MLP number w_iter
  }

  __device__ __forceinline__ void wg_sync() const { named_bar_sync(1 + wg(), kWgThreads); }

  // All 128 threads of the warpgroup: make this thread's TMEM stores visible to the tensor core, meet, and let
  // the issuing warp queue one layer for tiles [j0, j1) -- each tile commits to its own mbarrier.
  __device__ __forceinline__ void publish_and_issue(int j0, int j1, uint32_t b_hi, uint32_t b_lo, int k, int n) {
    tmem_wait_st();
    tc_fence_before();
    if (!issuer()) {
      // Only the issuing warp has to see all 128 rows published.  With one tile per warpgroup the other warps just
      // check in: they cannot reach the next check-in before this layer's MMAs -- which the issuing warp queues
      // after its own bar.sync -- have completed.  With two tiles a warp can publish tile 1 while the issuing warp
      // still waits for tile 0's check-ins, so there everybody waits.
      if constexpr (RPT == 1) {
        named_bar_arrive(1 + wg(), kWgThreads);
      } else {
        wg_sync();
      }
    } else {  // warp-uniform: operands stay on the uniform datapath
      wg_sync();
      tc_fence_after();
      if (elect_one()) {
        for (int j = j0; j < j1; ++j) {
          const uint32_t d = md(j);
          switch (k) {
            case 8: issue_layer<1>(d, d + kColAhi, d + kColAlo, m_ones(), b_hi, b_lo, n); break;
            case 16: issue_layer<2>(d, d + kColAhi, d + kColAlo, m_ones(), b_hi, b_lo, n); break;
            case 24: issue_layer<3>(d, d + kColAhi, d + kColAlo, m_ones(), b_hi, b_lo, n); break;
            default: issue_layer<4>(d, d + kColAhi, d + kColAlo, m_ones(), b_hi, b_lo, n); break;
          }
          mma_commit(mma_bar(j));
        }
      }
      __syncwarp();
    }
  }
  // the layer queued last for tile j has completed: its accumulator may be read, its A operand overwritten
  __device__ __forceinline__ void await(int j) const {
    mbar_wait(mma_bar(j), mma_phase);
    tc_fence_after();
  }

  // 16 activations -> 3xTF32 split -> columns [16 half, 16 half + 16) of A_hi / A_lo of tile j.
  // RELU: the values are pre-activations and the activation is ReLU (applied here, as 2 relu: the next layer's
  // weights are halved in the image).
  template <bool RELU>
  __device__ __forceinline__ void split_store(int j, int half, const float (&h)[16]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; q += 2) {
      if constexpr (RELU) {
        split_relu2_tf32_x2(h[q], h[q + 1], hi[q], hi[q + 1], lo[q], lo[q + 1]);
      } else {
        split_tf32_x2(h[q], h[q + 1], hi[q], hi[q + 1], lo[q], lo[q + 1]);
      }
    }
    tmem_st16(td(j) + kColAhi + 16 * half, hi);
    tmem_st16(td(j) + kColAlo + 16 * half, lo);
  }
  // ACT >= 0: compile-time activation (the hot path is ReLU); ACT < 0: `act` decides at run time
  template <int ACT>
  __device__ __forceinline__ void act_split_store(int j, int half, int act, float (&h)[16]) {
    if (ACT == VBN_ACT_RELU || (ACT < 0 && act == VBN_ACT_RELU)) {
      split_store<true>(j, half, h);
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) h[q] = activate_slow(h[q], act);
      split_store<false>(j, half, h);
    }
  }
  // D of tile j (32 fp32 columns, bias included) -> activation -> split -> A.  Both halves of the accumulator
  // row are requested before the first is processed, so the second TMEM read is in flight meanwhile.
  template <int ACT>
  __device__ __forceinline__ void hidden_epilogue(int j, int act) {
    uint32_t v0[16], v1[16];
    tmem_ld16(td(j), v0);
    tmem_ld16(td(j) + 16, v1);
    tmem_wait_ld();
    float h[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) h[q] = __uint_as_float(v0[q]);
    act_split_store<ACT>(j, 0, act, h);
#pragma unroll
    for (int q = 0; q < 16; ++q) h[q] = __uint_as_float(v1[q]);
    act_split_store<ACT>(j, 1, act, h);
  }

  // First layer on the FP32 pipe (<= 4 parent dims: 16 Dp packed FMAs per row).  `l1` = shared-space address of
  // the plain block at the head of the weight image: W1^T[4][32] (rows >= Dp zero), b1[32].  The MMA route
  // (K padded 2 -> 8: 5 MMAs, an input split, two TMEM stores, a commit / wait round trip and a 32-column TMEM
  // read) only pays from Dp > 4.  DP: compile-time parent count, so the body is straight-line code.
  template <int DP, int ACT, class C>
  __device__ __forceinline__ void hidden1_fma(C& c, const VbnOp& op, const float* norm, int j, uint32_t l1) {
    unsigned long long h[16];  // 32 pre-activations as f32x2 pairs
#pragma unroll
    for (int q = 0; q < 8; ++q) lds_f32x4_as_x2(l1 + 4 * (4 * kHidden + 4 * q), h[2 * q], h[2 * q + 1]);
#pragma unroll
    for (int p = 0; p < DP; ++p) {
      // (one word per parent slot, as op_lg_plain reads them, was measured slower here: 3 more live descriptor words
      // at 128 registers per thread doubled the spill traffic, cfg5 66.0 -> 68.9 ms)
      float z = c.slot((op.aux[1 + (p >> 1)] >> (16 * (p & 1))) & 0xFFFF, j);
      if (norm) z = __fdiv_rn(z - __ldg(norm + p), __ldg(norm + DP + p));
      unsigned long long zz;
      asm("mov.b64 %0, {%1, %1};" : "=l"(zz) : "f"(z));
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        unsigned long long w0, w1;
        lds_f32x4_as_x2(l1 + 4 * (p * kHidden + 4 * q), w0, w1);
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(h[2 * q]) : "l"(w0), "l"(zz));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(h[2 * q + 1]) : "l"(w1), "l"(zz));
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float f[16];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm("mov.b64 {%0, %1}, %2;" : "=f"(f[2 * q]), "=f"(f[2 * q + 1]) : "l"(h[8 * half + q]));
      act_split_store<ACT>(j, half, op.act, f);
    }
  }
  template <int ACT, class C>
  __device__ __forceinline__ void hidden1_fma_any(C& c, const VbnOp& op, const float* norm, int j, uint32_t l1) {
    switch (op.n_par) {
      case 1: hidden1_fma<1, ACT>(c, op, norm, j, l1); break;
      case 2: hidden1_fma<2, ACT>(c, op, norm, j, l1); break;
      case 3: hidden1_fma<3, ACT>(c, op, norm, j, l1); break;
      default: hidden1_fma<4, ACT>(c, op, norm, j, l1); break;
    }
  }

  // Inputs of tile j -> A (K1 columns, zero padded) for a first layer on the tensor core; gaussian_nn
  // standardises them first.
  template <class C>
  __device__ __forceinline__ void inputs_to_a(C& c, const VbnOp& op, const float* norm, const int32_t* par, int j,
                                              int k1) {
    const int dp = op.n_par;
    if (op.flags & VBN_F_PAR4) {  // <= 4 parent dims, slots packed in aux[1..2]: K1 == 8 (static field indices only:
                                  // a dynamically indexed descriptor would be demoted to local memory)
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float z = 0.0f;
        if (q < dp) {
          z = c.slot((op.aux[1 + (q >> 1)] >> (16 * (q & 1))) & 0xFFFF, j);
          if (norm) z = __fdiv_rn(z - __ldg(norm + q), __ldg(norm + dp + q));
        }
        split_tf32(z, hi[q], lo[q]);
      }
#pragma unroll
      for (int q = 4; q < 8; ++q) hi[q] = lo[q] = 0u;
      tmem_st8(td(j) + kColAhi, hi);
      tmem_st8(td(j) + kColAlo, lo);
      return;
    }
    for (int k0 = 0; k0 < k1; k0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int p = k0 + q;
        float z = 0.0f;
        if (p < dp) {
          z = c.slot(__ldg(par + p), j);
          if (norm) z = __fdiv_rn(z - __ldg(norm + p), __ldg(norm + dp + p));
        }
        split_tf32(z, hi[q], lo[q]);
      }
      tmem_st8(td(j) + kColAhi + k0, hi);
      tmem_st8(td(j) + kColAlo + k0, lo);
    }
  }

  // The MLP of one op for all RPT tiles, in two stages so that the caller can put independent work into the
  // shadow of the MMAs:
  //   stage_a : first layer (FP32 pipe, or MMA round trip), hidden layer queued              -> shadow 1
  //   stage_b : per tile: wait, split the hidden layer's accumulator, queue the last layer   -> shadow 2
  //   await(j) per tile -- D of tile j then holds the N3 outputs, bias included -- and layers_end().
  // FAST (the drawn-only MDN hot path): first layer on the FP32 pipe, ReLU, all decided at compile time.
  template <bool FAST, class C>
  __device__ __forceinline__ void stage_a(C& c, const VbnOp& op, const float* norm, const int32_t* par) {
    const int k1 = FAST ? 0 : op.tc[2];
    release_held();
    if (producer()) {  // warp-uniform
      if (elect_one()) produce(c.a);
      __syncwarp();
    }
    const uint32_t buf = w_buf & 0xFFu;
    const uint32_t ph = w_buf >> 8;
    const uint32_t img = wbuf(buf);
    // image: k1 == 0: L1 plain block (640 B) | W2hi W2lo | W3hi W3lo ; else W1hi W1lo | W2hi W2lo | W3hi W3lo
    const uint32_t w1_bytes = k1 == 0 ? kL1PlainBytes / 2 : kHidden * (k1 + kBiasK) * 4;  // per half
    const uint32_t w2hi = img + 2 * w1_bytes;
    const uint32_t w2lo = w2hi + kHidden * (kHidden + kBiasK) * 4;
    if (k1 == 0) {
      mbar_wait(full_bar(buf), ph);  // this layer reads its weights from the image itself
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        if constexpr (FAST) {
          hidden1_fma_any<VBN_ACT_RELU>(c, op, norm, j, img);
        } else {
          hidden1_fma_any<-1>(c, op, norm, j, img);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < RPT; ++j) inputs_to_a(c, op, norm, par, j, k1);
      mbar_wait(full_bar(buf), ph);
      publish_and_issue(0, RPT, img, img + w1_bytes, k1, kHidden);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        await(j);
        hidden_epilogue<-1>(j, op.act);
      }
      mma_phase ^= 1u;
    }
    publish_and_issue(0, RPT, w2hi, w2lo, kHidden, kHidden);
  }
  template <bool FAST, class C>
  __device__ __forceinline__ void stage_b(C& c, const VbnOp& op) {
    const int k1 = FAST ? 0 : op.tc[2], n3 = op.tc[3];
    const uint32_t w1_bytes = k1 == 0 ? kL1PlainBytes / 2 : kHidden * (k1 + kBiasK) * 4;
    const uint32_t w3hi = wbuf(w_buf & 0xFFu) + 2 * w1_bytes + 2 * kHidden * (kHidden + kBiasK) * 4;
    const uint32_t w3lo = w3hi + n3 * (kHidden + kBiasK) * 4;
    // tile j's accumulator is split while the other tiles' MMAs run; its last layer is queued at once
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      await(j);
      if constexpr (FAST) {
        hidden_epilogue<VBN_ACT_RELU>(j, VBN_ACT_RELU);
      } else {
        hidden_epilogue<-1>(j, op.act);
      }
      publish_and_issue(j, j + 1, w3hi, w3lo, kHidden, n3);
    }
    mma_phase ^= 1u;
  }
  // `gop` = the MLP op in the global op list; layer_dim[7] = descriptors behind its weights in the image.
  // The ring slot stays held until the next MLP starts (release_held): the ops in between read their descriptors
  // from it.
  __device__ __forceinline__ void layers_end(const VbnOp& op, const VbnOp* gop) {
    mma_phase ^= 1u;
    const int k1 = op.tc[2], n3 = op.tc[3];
    role |= 0x800u;
    if (op.layer_dim[7] != 0) {  // the ops up to and including the next MLP op ride behind the weights
      dptr = reinterpret_cast<const int4*>(
          __cvta_shared_to_generic(wbuf(w_buf & 0xFFu) + static_cast<uint32_t>(blob_bytes(k1, n3))));
    } else {
      dptr = reinterpret_cast<const int4*>(gop + 1);
    }
  }

  // Generic consumer: outputs land in scratch rows 0..n_out-1 like the FFMA paths
  // (mlp_fast32 / mlp_generic), for op_gnn / op_mdn / op_snn to pick up.
  template <class C>
  __device__ __forceinline__ void mlp(C& c, const VbnOp& op, const float* norm, const int32_t* par) {
    stage_a<false>(c, op, norm, par);
    stage_b<false>(c, op);
    const int n3 = op.tc[3], n_out = op.n_out;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      await(j);
      for (int o0 = 0; o0 < n3; o0 += 16) {
        uint32_t v[16];
        tmem_ld16(td(j) + o0, v);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 16; ++q)
          if (o0 + q < n_out) c.scr(o0 + q, j) = __uint_as_float(v[q]);
      }
    }
    layers_end(op, c.gop);
  }

  // VBN_F_MDNPLAIN: an MDN node (D = 1, K <= 5 components, <= 4 parent dims) that is only drawn
  // (mdn.py:209-235).  MLP outputs stay in registers: logits[K], then per component (loc, raw scale).
  template <int K, class C>
  __device__ __forceinline__ void mdn_tail(C& c, const VbnOp& op, const uint32_t (&v)[16], int j) {
    float q[K];
    float mx = __uint_as_float(v[0]);
#pragma unroll
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, __uint_as_float(v[k]));
    float se = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      q[k] = ex2_ftz((__uint_as_float(v[k]) - mx) * 1.4426950408889634f);  // only places the CDF edges: 2^-21
                                                                             // relative is plenty; arguments <= 0
      se += q[k];
    }
    // pi = softmax.clamp_min(1e-5) / sum (mdn.py:227-228); k ~ Categorical(pi) by inverse CDF on
    // un-normalised weights: e_k / se >= 1e-5  <=>  e_k >= 1e-5 se, so the clamp is applied to e_k
    // itself and u is scaled by the clamped total -- no division (se >= 1: the max term is 1)
    const float floor_e = 1e-5f * se;
    float tot = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      q[k] = fmaxf(q[k], floor_e);
      tot += q[k];
    }
    const float u = lane4(c.rows.ucache[j], op.u_off & 3) * tot;
    float cum = 0.0f;
    float loc = __uint_as_float(v[K + 2 * (K - 1)]), raw = __uint_as_float(v[K + 2 * (K - 1) + 1]);
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
      const float before = cum;
      cum += q[k];
      if (u >= before && u < cum) {
        loc = __uint_as_float(v[K + 2 * k]);
        raw = __uint_as_float(v[K + 2 * k + 1]);
      }
    }
    const float eps = lane4(c.rows.ncache[j], op.n_off & 3);
    const float sc = softplus20_fast(raw) + __int_as_float(op.aux[0]);
    c.slot(op.out_slot, j) = fmaf(eps, sc, loc);
  }

  template <bool FAST, class C>
  __device__ __forceinline__ void mdn_plain_body(C& c, const VbnOp& op) {
    stage_a<FAST>(c, op, nullptr, nullptr);
    {  // shadow of the hidden layer's MMAs: refill the uniform cache
      const int uq = op.u_off >> 2;
      if (uq != c.rows.cur_uq) {
        c.rows.cur_uq = uq;
#pragma unroll
        for (int j = 0; j < RPT; ++j) c.rows.ucache[j] = c.uniforms(j, uq, 1u, false);
      }
    }
    stage_b<FAST>(c, op);
    {  // shadow of the last layer's MMAs: refill the normal cache
      const int nq = op.n_off >> 2;
      if (nq != c.rows.cur_nq) {
        c.rows.cur_nq = nq;
#pragma unroll
        for (int j = 0; j < RPT; ++j) c.rows.ncache[j] = c.normals(j, nq, 0u, false);
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      await(j);
      uint32_t v[16];
      tmem_ld16(td(j), v);
      tmem_wait_ld();
      if constexpr (FAST) {
        mdn_tail<3>(c, op, v, j);
      } else {
        switch (op.k) {
          case 2: mdn_tail<2>(c, op, v, j); break;
          case 3: mdn_tail<3>(c, op, v, j); break;
          case 4: mdn_tail<4>(c, op, v, j); break;
          default: mdn_tail<5>(c, op, v, j); break;
        }
      }
    }
    layers_end(op, c.gop);
  }
  // mdn_plain_body<true>: the reference's defaults (mdn.py: n_components = 3 is what BASELINE cfg5 uses, ReLU, first
  // layer on the FP32 pipe) get a body with every one of those choices made at compile time; the plan compiler marks
  // those ops VBN_F_MDNFAST (run_ops fetches a shorter descriptor for them).
};

// The kernel.  grid <= #SMs (one CTA per SM), NWG * 128 threads, NWG * RPT * 128 rows per pass.
// dynamic smem: [ctrl][nbuf weight buffers][slots + scratch: (n_slots+n_scratch) x ROWS floats]
template <int NWG, int RPT>
__global__ void __launch_bounds__(NWG* kWgThreads, 1) schedule_tc_kernel(const ScheduleArgs a, const int nbuf, const int slot_bytes, const int tune) {
  constexpr int kThreads = NWG * kWgThreads;
  constexpr int kTiles = NWG * RPT;
  using Tc = TcMlp<NWG, RPT>;
  static_assert(kTiles <= kMaxTiles && kTiles * kColsPerTile + kBiasK <= kTmemCols, "TMEM budget");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform for the compiler
  const uint32_t smem_base = smem_u32(smem_raw);
  // Word offset of this thread's column in the slot area.  Made opaque to the optimiser: at 128
  // registers per thread ptxas otherwise REMATERIALISES it (S2R tid, nbuf * kWbufBytes, ...) at the top
  // of every op of the walk instead of keeping -- or spilling -- one register.
  int slot_word = (kCtrlBytes + nbuf * slot_bytes) / 4 + tid;
  asm volatile("" : "+r"(slot_word));
  float* slots = reinterpret_cast<float*>(smem_raw) + slot_word;

  const int64_t n_tiles = (a.n_rows + kWgThreads - 1) / kWgThreads;
  const int64_t per_round = static_cast<int64_t>(gridDim.x) * kTiles;
  const int64_t n_iter = (n_tiles + per_round - 1) / per_round;  // same for every warpgroup: the ring needs it

  if (tid == 0) {
    for (int i = 0; i < kMaxBufs; ++i) {
      mbar_init(smem_base + Tc::kFull + 8 * i, 1);               // the producer lane's arrive.expect_tx
      mbar_init(smem_base + Tc::kEmpty + 8 * i, kThreads / 32);  // one arrive per warp
    }
    for (int g = 0; g < kTiles; ++g) mbar_init(smem_base + Tc::kMma + 8 * g, 1);  // tcgen05.commit
    fence_mbar_init();
    sts_u32(smem_base + Tc::kProd, 0u);                // p_iter
    sts_u32(smem_base + Tc::kProd + 4, 1u << 8);       // p_buf
    sts_u32(smem_base + Tc::kProd + 8, 0u);            // p_idx
    sts_u32(smem_base + Tc::kProd + 12, static_cast<uint32_t>(n_iter) * static_cast<uint32_t>(a.n_tc));  // p_total
    sts_u32(smem_base + Tc::kProd + 16, 0u);           // w_iter of the producer warp
  }
  if (warp == 0) tmem_alloc(smem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *reinterpret_cast<volatile uint32_t*>(smem_raw), 0);

  {  // constant A block of the bias step: columns [kTiles*96, kTiles*96 + 8) of every lane = (1, 0, ..., 0)
    const uint32_t ones[8] = {__float_as_uint(1.0f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (warp < 4) tmem_st8(tmem_base + kTiles * kColsPerTile + (static_cast<uint32_t>(warp * 32) << 16), ones);
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  Ctx<RPT, kThreads, Tc> c(a, slots, 0);  // the thread index is already folded into `slots`
  const int wg = warp >> 2;
  c.tc.t_d = tmem_base + static_cast<uint32_t>(wg * RPT * kColsPerTile) + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  c.tc.tm = tmem_base;
  c.tc.ctl = smem_base;
  // the issuing warp of warpgroup g is its warp 0; the ring producer is a warp that issues no MMAs
  c.tc.role = static_cast<uint32_t>(wg) | ((warp & 3) == ((tune & 1) ? (wg & 3) : 0) ? 0x100u : 0u) |
              (warp == ((tune & 2) ? 3 : 6) ? 0x200u : 0u) |
              (lane == 0 ? 0x400u : 0u);
  c.tc.w_buf = 0;
  c.tc.dptr = nullptr;
  c.tc.mma_phase = 0;
  c.tc.nbuf = static_cast<uint32_t>(nbuf);
  c.tc.slot_bytes = static_cast<uint32_t>(slot_bytes);  // a kernel parameter: stays a constant-bank operand
  // values, not recipes in (tid, wg, nbuf): one register each for the whole walk
  asm volatile("" : "+r"(c.tc.t_d), "+r"(c.tc.tm), "+r"(c.tc.ctl), "+r"(c.tc.role), "+r"(c.tc.nbuf));
  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t base = (it * gridDim.x + blockIdx.x) * static_cast<int64_t>(kThreads * RPT);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      bind_row(c, j, base + j * kThreads + tid);
      // same reason as slot_word: keep the Philox counter words as values, not as a recipe (the row
      // index -> (query, sample) division chain was being re-executed per op)
      int row_ok = c.rows.valid[j] ? 1 : 0;
      asm volatile("" : "+r"(c.rows.gs[j]), "+r"(c.rows.gb[j]), "+l"(c.rows.r[j]), "+r"(row_ok));
      c.rows.valid[j] = row_ok != 0;
    }
    c.tc.dptr = reinterpret_cast<const int4*>(a.ops);
    run_ops<true>(c);
    c.tc.release_held();  // the last MLP's ring slot (the ops behind it read their descriptors from its tail)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace vbn
