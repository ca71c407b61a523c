// The fused schedule kernel: one launch walks every op (node) of a compiled schedule for a
// tile of rows; node values live in shared-memory slots, never in HBM unless the caller
// asked for the column.  sm_100a only.
//
// Replaces, per row, the Python per-node loops of the reference:
//   vbn/inference/likelihood_weighting.py:42-71, importance_sampling.py:56-80,
//   monte_carlo_marginalization.py:39-92, vbn/sampling/ancestral.py:27-41
// and the CPD bodies they call (cited at each op below).
#pragma once
#include <type_traits>

#include "vbn_device.cuh"

namespace vbn {

struct ScheduleArgs {
  const VbnOp* ops;
  const int32_t* par_slots;
  const float* params;
  int32_t n_ops;
  int32_t n_slots;
  int32_t n_scratch;
  int32_t logp_as_pdf;
  int32_t logw_accumulate;
  int64_t n_queries;  // B (local)
  int64_t n_samples;  // S (local)
  int64_t n_rows;     // B*S
  uint32_t query_offset;
  uint32_t sample_offset;
  uint32_t key0, key1;
  uint32_t call_offset;
  const float* fixed;
  const VbnView* inputs;
  const VbnView* stores;
  const VbnNoise* noise;
  float* logw;
  float* logp;
  int32_t* error_flag;
  const int2* tc_list;  // tensor-core kernel only: per tc op {image float offset, image bytes}
  int32_t n_tc;
  // fused weight reduction (emit_segments): per-warp partial records [n_queries][seg_per_query][kSegWords]
  float* seg;
  int32_t seg_per_query;
  int32_t seg_slot;     // slot of the value whose weighted moments are accumulated (D = 1), -1: log-weights only
  int32_t seg_classes;  // 1..8: also the weighted class histogram of that value (classes coded 0..k-1)
  uint32_t rk[20];      // Philox round keys: rk[2i] = key0 + i*0x9E3779B9, rk[2i+1] = key1 + i*0xBB67AE85
};

__host__ __device__ inline void fill_round_keys(ScheduleArgs& a) {
  for (int i = 0; i < 10; ++i) {
    a.rk[2 * i] = a.key0 + static_cast<uint32_t>(i) * 0x9E3779B9u;
    a.rk[2 * i + 1] = a.key1 + static_cast<uint32_t>(i) * 0xBB67AE85u;
  }
}

constexpr int kMaxGenericWidth = 128;  // widest layer the generic MLP path accepts

__host__ __device__ __forceinline__ int pad4(int x) { return (x + 3) & ~3; }

// Per-thread state for RPT rows handled together (row j of the thread is tile row j*NT+tid).
template <int RPT>
struct Rows {
  int64_t r[RPT];   // local row index, clamped into range
  int64_t lb[RPT];  // local query index
  int64_t ls[RPT];  // local sample index
  uint32_t gb[RPT], gs[RPT];  // global query / sample index (RNG counter words)
  bool valid[RPT];
  float logw[RPT];
  float logp[RPT];
  // VBN_OP_JUMP loop state: iteration number, and the Philox stream blocks one iteration consumes
  // (added to every stream block index so each iteration draws fresh numbers)
  int loop_iter, loop_nq, loop_uq;
  // Philox stream caches (per-row keyed streams): block index + 4 values per row
  int cur_nq, cur_uq;
  float4 ncache[RPT], ucache[RPT];
};

// Tensor-core hook: the tcgen05 kernel (vbn_schedule_tc.cuh) plugs its MLP evaluator in here.
struct NoTc {
  static constexpr bool kEnabled = false;
  static constexpr bool kInlineRng = false;
  static constexpr bool kLoops = true;  // carries the Gibbs glue ops (VBN_OP_TAKEW / SELECT / JUMP)
  static constexpr bool kTab = true;    // carries the table-lookup op bodies
};
// LG / table-only schedules: the generator is the hot loop, so it is inlined with constant-bank
// round keys instead of being called out of line.
// TAB = false: linear-Gaussian-only schedules get a kernel without the table op bodies (a chain's hot loop then has
// nothing but the plain LG op in it: BASELINE cfg2 lost 9 % when the table code moved into the same kernel).
template <bool TAB>
struct LightPolicyT {
  static constexpr bool kEnabled = false;
  static constexpr bool kInlineRng = true;
  static constexpr bool kLoops = false;  // the LG / table hot loops stay a plain counted walk
  static constexpr bool kTab = TAB;
};
using LightPolicy = LightPolicyT<true>;

template <int RPT, int NT, class TC = NoTc>
struct Ctx {
  const ScheduleArgs& a;
  const VbnOp* gop = nullptr;  // the op being executed, in global memory (for dynamic field access)
  float* slots;    // [n_slots][ROWS]
  float* scratch;  // [n_scratch][ROWS]
  int tid;
  TC tc;
  Rows<RPT> rows;
  static constexpr int ROWS = RPT * NT;

  __device__ __forceinline__ Ctx(const ScheduleArgs& args, float* smem, int t)
      : a(args), slots(smem), scratch(smem + static_cast<size_t>(args.n_slots) * ROWS), tid(t) {}

  __device__ __forceinline__ float& slot(int s, int j) { return slots[s * ROWS + j * NT + tid]; }
  __device__ __forceinline__ float& scr(int s, int j) { return scratch[s * ROWS + j * NT + tid]; }

  // ---- random draws ---------------------------------------------------------------------
  // stream tags in counter word c2 (top 2 bits): 0 normal/row, 1 uniform/row, 2 normal/shared,
  // 3 uniform/shared
  __device__ __forceinline__ uint4 counter(int j, uint32_t block, uint32_t tag, bool shared) const {
    return make_uint4(rows.gs[j], shared ? 0xFFFFFFFFu : rows.gb[j], block | (tag << 30), a.call_offset);
  }
  __device__ __forceinline__ float4 normals(int j, uint32_t block, uint32_t tag, bool shared) const {
    if constexpr (TC::kInlineRng) return normal4(philox4x32_10_rk(counter(j, block, tag, shared), a.rk));
    return philox_normal4(counter(j, block, tag, shared), make_uint2(a.key0, a.key1));
  }
  __device__ __forceinline__ float4 uniforms(int j, uint32_t block, uint32_t tag, bool shared) const {
    if constexpr (TC::kInlineRng) return uniform4(philox4x32_10_rk(counter(j, block, tag, shared), a.rk));
    return philox_uniform4(counter(j, block, tag, shared), make_uint2(a.key0, a.key1));
  }

  // Element of an injected array that belongs to row j: [Bn, S] entries of `width` values; inside a
  // VBN_OP_JUMP loop the arrays gain a leading iteration axis, and an op may address one member of a
  // group of draws the reference made in a single call (gop->tc[1] = group size, tc[2] = member:
  // the Gibbs sampler's candidates, sampling/gibbs.py:54).  Test / replay path only.
  __device__ __forceinline__ int64_t injected_row(int j, bool shared) const {
    if constexpr (!TC::kLoops) return shared ? rows.ls[j] : rows.r[j];
    const int64_t per_iter = shared ? a.n_samples : a.n_rows;
    int64_t r = static_cast<int64_t>(rows.loop_iter) * per_iter + (shared ? rows.ls[j] : rows.r[j]);
    const int group = __ldg(&gop->tc[0]) == 0 ? __ldg(&gop->tc[1]) : 0;  // tc[0] != 0: the fields describe a tensor-core image
    if (group > 1) r = r * group + __ldg(&gop->tc[2]);
    return r;
  }

  // normal number `index` of the op's stream, for all RPT rows
  __device__ __forceinline__ void draw_normal(const VbnOp& op, int index, int d, float (&out)[RPT]) {
    const bool shared = (op.flags & VBN_F_SHARED) != 0;
    if (op.noise_idx >= 0) {
      const float* eps = a.noise[op.noise_idx].eps;
#pragma unroll
      for (int j = 0; j < RPT; ++j) out[j] = __ldg(eps + injected_row(j, shared) * op.dim + d);
      return;
    }
    if constexpr (TC::kLoops) index += 4 * rows.loop_iter * rows.loop_nq;
    cached_normal(shared, index, out);
  }

  // Value `index & 3` of stream block `index >> 2` for all RPT rows, through the 4-value block cache.
  // One cache serves both streams (key = block | shared << 30).  Shared draws are the roots of LW / MCM /
  // ancestral passes; the plan compiler moves them to the head of the schedule (plan.py), where four of them
  // share one generator call.  A shared draw that arrives while the cache holds a per-row block (roots left
  // interleaved in a large DAG) bypasses the cache instead of evicting it.
  __device__ __forceinline__ void cached_normal(bool shared, int index, float (&out)[RPT]) {
    const int q = index >> 2, lane = index & 3;
    const int key = q | (shared ? 0x40000000 : 0);
    if (key == rows.cur_nq) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) out[j] = lane4(rows.ncache[j], lane);
      return;
    }
    const bool keep = !shared || rows.cur_nq < 0 || (rows.cur_nq & 0x40000000) != 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float4 blk = normals(j, q, shared ? 2u : 0u, shared);
      if (keep) rows.ncache[j] = blk;
      out[j] = lane4(blk, lane);
    }
    if (keep) rows.cur_nq = key;
  }
  __device__ __forceinline__ void cached_uniform(bool shared, int index, float (&out)[RPT]) {
    const int q = index >> 2, lane = index & 3;
    const int key = q | (shared ? 0x40000000 : 0);
    if (key == rows.cur_uq) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) out[j] = lane4(rows.ucache[j], lane);
      return;
    }
    const bool keep = !shared || rows.cur_uq < 0 || (rows.cur_uq & 0x40000000) != 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float4 blk = uniforms(j, q, shared ? 3u : 1u, shared);
      if (keep) rows.ucache[j] = blk;
      out[j] = lane4(blk, lane);
    }
    if (keep) rows.cur_uq = key;
  }

  // uniform number `index`; `slot_in_noise` selects which injected array element to use
  __device__ __forceinline__ void draw_uniform(const VbnOp& op, int index, int d, float (&out)[RPT]) {
    const bool shared = (op.flags & VBN_F_SHARED) != 0;
    if (op.noise_idx >= 0 && a.noise[op.noise_idx].u != nullptr) {
      const float* u = a.noise[op.noise_idx].u;
#pragma unroll
      for (int j = 0; j < RPT; ++j) out[j] = __ldg(u + injected_row(j, shared) * op.dim + d);
      return;
    }
    if constexpr (TC::kLoops) index += 4 * rows.loop_iter * rows.loop_uq;
    cached_uniform(shared, index, out);
  }

  // injected categorical pick (per row, per dim `d` of `width` dims); returns false if absent
  __device__ __forceinline__ bool injected_index(const VbnOp& op, int d, int width, int (&out)[RPT]) {
    if (op.noise_idx < 0) return false;
    const int32_t* idx = a.noise[op.noise_idx].idx;
    if (idx == nullptr) return false;
    const bool shared = (op.flags & VBN_F_SHARED) != 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) out[j] = __ldg(idx + injected_row(j, shared) * width + d);
    return true;
  }
};

// ---------------------------------------------------------------------------------------
// MLP forward (gaussian_nn.py:16-34 _build_mlp; mdn.py:199; softmax_nn.py:585).
// Parameter block layout (floats, every array padded to a multiple of 4):
//   n_layers == 0 : out[n_out] constants (root CPDs)
//   hidden layer l: WT[in][pad4(out)] (transposed), bias[pad4(out)]
//   last layer    : W[out][pad4(in)] (nn.Linear layout), bias[pad4(out)]
// Result lands in scratch rows 0..n_out-1.
// `norm` (gaussian_nn only): mean_x[Dp], std_x[Dp] -> z = (pa - mean_x)/std_x (gaussian_nn.py:105-112)
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void mlp_generic(Ctx<RPT, NT, TC>& c, const VbnOp& op, const float* P,
                                         const float* norm, const int32_t* par) {
  float bufA[kMaxGenericWidth], bufB[kMaxGenericWidth];
  for (int j = 0; j < RPT; ++j) {
    for (int p = 0; p < op.n_par; ++p) {
      float z = c.slot(__ldg(par + p), j);
      if (norm) z = __fdiv_rn(z - __ldg(norm + p), __ldg(norm + op.n_par + p));
      bufA[p] = z;
    }
    float* cur = bufA;
    float* nxt = bufB;
    int in = op.n_par;
    const float* W = P;
    for (int l = 0; l + 1 < op.n_layers; ++l) {
      const int out = __ldg(&c.gop->layer_dim[l]), outp = pad4(out);
      const float* bias = W + static_cast<size_t>(in) * outp;
      for (int o = 0; o < out; ++o) {
        float acc = __ldg(bias + o);
        for (int i = 0; i < in; ++i) acc = fmaf(__ldg(W + i * outp + o), cur[i], acc);
        nxt[o] = activate(acc, op.act);
      }
      W = bias + outp;
      float* t = cur;
      cur = nxt;
      nxt = t;
      in = out;
    }
    const int out = op.n_out, inp = pad4(in);
    const float* bias = W + static_cast<size_t>(out) * inp;
    for (int o = 0; o < out; ++o) {
      float acc = __ldg(bias + o);
      for (int i = 0; i < in; ++i) acc = fmaf(__ldg(W + o * inp + i), cur[i], acc);
      c.scr(o, j) = acc;
    }
  }
}

// Fast path for the reference's default hidden_dims [32, 32]: h1 lives in registers, the
// 32x32 layer is register-tiled 8 outputs at a time with 128-bit uniform weight loads (L1
// broadcast), and the output layer is accumulated chunk by chunk into scratch.
template <int RPT, int NT, class TC>
__device__ __forceinline__ void mlp_fast32(Ctx<RPT, NT, TC>& c, const VbnOp& op, const float* P,
                                           const float* norm, const int32_t* par) {
  constexpr int H = 32;
  const int dp = op.n_par;
  const float* W1 = P;               // [dp][32]
  const float* b1 = W1 + dp * H;     // [32]
  const float* W2 = b1 + H;          // [32][32] transposed: W2[k][j]
  const float* b2 = W2 + H * H;      // [32]
  const float* W3 = b2 + H;          // [O][32]
  const int O = op.n_out;
  const float* b3 = W3 + O * H;      // [pad4(O)]
  const int act = op.act;

  float h1[RPT][H];
#pragma unroll
  for (int q = 0; q < H / 4; ++q) {
    const float4 b = ldg4(b1 + 4 * q);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      h1[j][4 * q + 0] = b.x;
      h1[j][4 * q + 1] = b.y;
      h1[j][4 * q + 2] = b.z;
      h1[j][4 * q + 3] = b.w;
    }
  }
  for (int p = 0; p < dp; ++p) {
    float z[RPT];
    const int ps = __ldg(par + p);
#pragma unroll
    for (int j = 0; j < RPT; ++j) z[j] = c.slot(ps, j);
    if (norm) {
      const float mu = __ldg(norm + p), sd = __ldg(norm + dp + p);
#pragma unroll
      for (int j = 0; j < RPT; ++j) z[j] = __fdiv_rn(z[j] - mu, sd);
    }
#pragma unroll
    for (int q = 0; q < H / 4; ++q) {
      const float4 w = ldg4(W1 + p * H + 4 * q);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        h1[j][4 * q + 0] = fmaf(w.x, z[j], h1[j][4 * q + 0]);
        h1[j][4 * q + 1] = fmaf(w.y, z[j], h1[j][4 * q + 1]);
        h1[j][4 * q + 2] = fmaf(w.z, z[j], h1[j][4 * q + 2]);
        h1[j][4 * q + 3] = fmaf(w.w, z[j], h1[j][4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < H; ++k)
#pragma unroll
    for (int j = 0; j < RPT; ++j) h1[j][k] = activate(h1[j][k], act);

  for (int o = 0; o < O; ++o) {
    const float b = __ldg(b3 + o);
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.scr(o, j) = b;
  }

#pragma unroll 1
  for (int jc = 0; jc < H / 8; ++jc) {
    float acc[RPT][8];
    {
      const float4 ba = ldg4(b2 + 8 * jc), bb = ldg4(b2 + 8 * jc + 4);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        acc[j][0] = ba.x; acc[j][1] = ba.y; acc[j][2] = ba.z; acc[j][3] = ba.w;
        acc[j][4] = bb.x; acc[j][5] = bb.y; acc[j][6] = bb.z; acc[j][7] = bb.w;
      }
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const float4 wa = ldg4(W2 + k * H + 8 * jc), wb = ldg4(W2 + k * H + 8 * jc + 4);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float h = h1[j][k];
        acc[j][0] = fmaf(wa.x, h, acc[j][0]);
        acc[j][1] = fmaf(wa.y, h, acc[j][1]);
        acc[j][2] = fmaf(wa.z, h, acc[j][2]);
        acc[j][3] = fmaf(wa.w, h, acc[j][3]);
        acc[j][4] = fmaf(wb.x, h, acc[j][4]);
        acc[j][5] = fmaf(wb.y, h, acc[j][5]);
        acc[j][6] = fmaf(wb.z, h, acc[j][6]);
        acc[j][7] = fmaf(wb.w, h, acc[j][7]);
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[j][q] = activate(acc[j][q], act);
    for (int o = 0; o < O; ++o) {
      const float4 wa = ldg4(W3 + o * H + 8 * jc), wb = ldg4(W3 + o * H + 8 * jc + 4);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float t = c.scr(o, j);
        t = fmaf(wa.x, acc[j][0], t);
        t = fmaf(wa.y, acc[j][1], t);
        t = fmaf(wa.z, acc[j][2], t);
        t = fmaf(wa.w, acc[j][3], t);
        t = fmaf(wb.x, acc[j][4], t);
        t = fmaf(wb.y, acc[j][5], t);
        t = fmaf(wb.z, acc[j][6], t);
        t = fmaf(wb.w, acc[j][7], t);
        c.scr(o, j) = t;
      }
    }
  }
}

template <int RPT, int NT, class TC>
__device__ __forceinline__ void mlp_eval(Ctx<RPT, NT, TC>& c, const VbnOp& op, const float* P,
                                         const float* norm, const int32_t* par) {
  if constexpr (TC::kEnabled) {
    if (op.tc[0] != 0) {  // tensor-core eligible MLP (plan.py sets the tc fields)
      c.tc.mlp(c, op, norm, par);
      return;
    }
  }
  if (op.n_layers == 0) {
    for (int o = 0; o < op.n_out; ++o) {
      const float v = __ldg(P + o);
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.scr(o, j) = v;
    }
  } else if (op.flags & VBN_F_FAST32) {  // n_layers == 3, hidden dims [32, 32]
    mlp_fast32(c, op, P, norm, par);
  } else {
    mlp_generic(c, op, P, norm, par);
  }
}

// ---------------------------------------------------------------------------------------
// value sources / sinks common to all ops
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void load_fixed(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const int src = op.flags & VBN_SRC_MASK;
  if (src == VBN_SRC_FIXED_Q) {
    for (int d = 0; d < op.dim; ++d)
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        c.slot(op.out_slot + d, j) =
            __ldg(c.a.fixed + static_cast<int64_t>(op.fixed_col + d) * c.a.n_queries + c.rows.lb[j]);
  } else if (src == VBN_SRC_FIXED_ROW) {
    const VbnView v = c.a.inputs[op.fixed_col];
    for (int d = 0; d < op.dim; ++d)
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        c.slot(op.out_slot + d, j) = __ldg(v.base + c.rows.r[j] * v.row_stride + d * v.dim_stride);
  } else if (src == VBN_SRC_SLOT) {  // score the value another op left in slot fixed_col (Gibbs: a child's state)
    if (op.fixed_col != op.out_slot)
      for (int d = 0; d < op.dim; ++d)
#pragma unroll
        for (int j = 0; j < RPT; ++j) c.slot(op.out_slot + d, j) = c.slot(op.fixed_col + d, j);
  }
}

template <int RPT, int NT, class TC>
__device__ __forceinline__ void store_value(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  if (op.store_idx < 0) return;
  const VbnView v = c.a.stores[op.store_idx];
  for (int d = 0; d < op.dim; ++d)
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (c.rows.valid[j]) v.base[c.rows.r[j] * v.row_stride + d * v.dim_stride] = c.slot(op.out_slot + d, j);
}

template <int RPT, int NT, class TC>
__device__ __forceinline__ void commit_logp(Ctx<RPT, NT, TC>& c, const VbnOp& op, const float (&lp)[RPT]) {
  if (op.flags & VBN_F_ADD_LOGW) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.rows.logw[j] += lp[j];
  }
  if (op.flags & VBN_F_OUT_LOGP) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.rows.logp[j] = lp[j];
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_LG: linear_gaussian.py:163-217 (and gaussian_nn roots, gaussian_nn.py:244-254).
// params: W[Dp][D], bias[D], scale[D], two_log_scale[D], var[D]
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_lg(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, Dp = op.n_par;
  const float* W = P;
  const float* bias = W + Dp * D;
  const float* scale = bias + D;
  const float* tls = scale + D;
  const float* var = tls + D;
  const int32_t* par = c.a.par_slots + op.par_off;
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;
  float acc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) acc[j] = 0.0f;
  for (int d = 0; d < D; ++d) {
    float loc[RPT];
    const float b = __ldg(bias + d);
#pragma unroll
    for (int j = 0; j < RPT; ++j) loc[j] = 0.0f;
    for (int p = 0; p < Dp; ++p) {
      const float w = __ldg(W + p * D + d);
      const int ps = __ldg(par + p);
#pragma unroll
      for (int j = 0; j < RPT; ++j) loc[j] = fmaf(c.slot(ps, j), w, loc[j]);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) loc[j] += b;  // flat @ W + bias (linear_gaussian.py:180)
    if (sample) {
      float eps[RPT];
      c.draw_normal(op, op.n_off + d, d, eps);
      const float sc = __ldg(scale + d);
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.slot(op.out_slot + d, j) = fmaf(eps[j], sc, loc[j]);
    }
    if (want_lp) {
      const float v = __ldg(var + d), t = __ldg(tls + d);
#pragma unroll
      for (int j = 0; j < RPT; ++j) acc[j] += gauss_term(c.slot(op.out_slot + d, j), loc[j], v, t);
    }
  }
  if (want_lp) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] *= -0.5f;
    commit_logp(c, op, acc);
  }
}

// VBN_OP_LG with VBN_F_LGFAST (D = 1, Dp <= 4): parameters and parent slots ride in the op itself
// (layer_dim[] as float bits: bias, scale, 2 ln scale, var, w0..w3; aux[] = parent slots), so the
// op costs one round of independent 128-bit loads and no dependent parameter fetches.
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_lg_fast(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const int Dp = op.n_par;
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;
  float loc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) loc[j] = 0.0f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < Dp) {
      const float w = __int_as_float(op.layer_dim[4 + p]);
      const int ps = (op.aux[p >> 1] >> (16 * (p & 1))) & 0xFFFF;
#pragma unroll
      for (int j = 0; j < RPT; ++j) loc[j] = fmaf(c.slot(ps, j), w, loc[j]);
    }
  }
  const float b = __int_as_float(op.layer_dim[0]);
#pragma unroll
  for (int j = 0; j < RPT; ++j) loc[j] += b;  // flat @ W + bias (linear_gaussian.py:180)
  if (sample) {
    float eps[RPT];
    c.draw_normal(op, op.n_off, 0, eps);
    const float sc = __int_as_float(op.layer_dim[1]);
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.slot(op.out_slot, j) = fmaf(eps[j], sc, loc[j]);
  }
  if (want_lp) {
    const float t = __int_as_float(op.layer_dim[2]), v = __int_as_float(op.layer_dim[3]);
    float acc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] = -0.5f * gauss_term(c.slot(op.out_slot, j), loc[j], v, t);
    commit_logp(c, op, acc);
  }
}

// VBN_F_LGPLAIN: an LGFAST op that is simply drawn (Philox, per-row stream, no store, no density).
// Everything it needs is in quads 0, 4, 5, 6: aux[] = {parent slot 0, parent slot 1, out_slot, n_off}, parent slots 2
// and 3 in layer_dim[2..3] (where a scored LG op keeps 2 ln scale and the variance): one word per slot, so a parent's
// shared-memory address is one shift-add instead of extract + scale + add.
// This is the inner loop of linear-Gaussian chains (BASELINE cfg2) and of the LG half of cfg5.
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_lg_plain(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const int Dp = op.n_par;
  float loc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) loc[j] = 0.0f;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < Dp) {
      const float w = __int_as_float(op.layer_dim[4 + p]);
      const int ps = p < 2 ? op.aux[p] : op.layer_dim[p];
#pragma unroll
      for (int j = 0; j < RPT; ++j) loc[j] = fmaf(c.slot(ps, j), w, loc[j]);
    }
  }
  const float b = __int_as_float(op.layer_dim[0]), sc = __int_as_float(op.layer_dim[1]);
  const int n_off = op.aux[3], q = n_off >> 2, lane = n_off & 3;
  if (q != c.rows.cur_nq) {
    c.rows.cur_nq = q;
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.rows.ncache[j] = c.normals(j, q, 0u, false);
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j) c.slot(op.aux[2], j) = fmaf(lane4(c.rows.ncache[j], lane), sc, loc[j] + b);
}

// VBN_F_MDNROOT: a parent-less MDN node (D = 1, K <= 4) that is simply drawn.  The mixture is the
// same for every row, so the plan compiler has already evaluated pi = softmax(logits).clamp_min(1e-5)
// / sum and scale_k = softplus(raw_k) + min_scale (mdn.py:190-196, 227-235); the op carries the
// cumulative weights and the (loc, scale) pairs.  k ~ Categorical(pi) by inverse CDF on one uniform.
template <int K, int RPT, int NT, class TC>
__device__ __forceinline__ void mdn_root_k(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  float f[12];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __int_as_float(op.layer_dim[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) f[8 + i] = __int_as_float(op.aux[i]);
  const int out_slot = op.tc[0], n_off = op.tc[1], u_off = op.tc[2];
  const int uq = u_off >> 2, ul = u_off & 3;
  if (uq != c.rows.cur_uq) {
    c.rows.cur_uq = uq;
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.rows.ucache[j] = c.uniforms(j, uq, 1u, false);
  }
  const int nq = n_off >> 2, nl = n_off & 3;
  if (nq != c.rows.cur_nq) {
    c.rows.cur_nq = nq;
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.rows.ncache[j] = c.normals(j, nq, 0u, false);
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const float u = lane4(c.rows.ucache[j], ul);
    float loc = f[K - 1 + 2 * (K - 1)], sc = f[K - 1 + 2 * (K - 1) + 1];
#pragma unroll
    for (int k = K - 2; k >= 0; --k)
      if (u < f[k]) {
        loc = f[K - 1 + 2 * k];
        sc = f[K - 1 + 2 * k + 1];
      }
    c.slot(out_slot, j) = fmaf(lane4(c.rows.ncache[j], nl), sc, loc);
  }
}
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_mdn_root(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  switch (op.tc[3]) {
    case 2: mdn_root_k<2>(c, op); break;
    case 3: mdn_root_k<3>(c, op); break;
    default: mdn_root_k<4>(c, op); break;
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_GNN: gaussian_nn.py:215-288.
// params: mean_x[Dp], std_x[Dp], mean_y[D], std_y[D], min_scale, pad4 ; then MLP block
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_gnn(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, Dp = op.n_par;
  const float* norm = P;
  const float* mean_y = P + 2 * Dp;
  const float* std_y = mean_y + D;
  const float min_scale = __ldg(std_y + D);
  const float* mlp = P + pad4(2 * Dp + 2 * D + 1);
  const int32_t* par = c.a.par_slots + op.par_off;
  mlp_eval(c, op, mlp, norm, par);
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;
  float acc[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) acc[j] = 0.0f;
  for (int d = 0; d < D; ++d) {
    const float my = __ldg(mean_y + d), sy = __ldg(std_y + d);
    float loc[RPT], sc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      loc[j] = fmaf(c.scr(d, j), sy, my);                          // loc*std_y + mean_y
      sc[j] = (softplus20(c.scr(D + d, j)) + min_scale) * sy;      // (softplus+min)*std_y
    }
    if (sample) {
      float eps[RPT];
      c.draw_normal(op, op.n_off + d, d, eps);
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.slot(op.out_slot + d, j) = fmaf(eps[j], sc[j], loc[j]);
    }
    if (want_lp) {
#pragma unroll
      for (int j = 0; j < RPT; ++j)
        acc[j] += gauss_term(c.slot(op.out_slot + d, j), loc[j], sc[j] * sc[j], 2.0f * logf(sc[j]));
    }
  }
  if (want_lp) {
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] *= -0.5f;
    commit_logp(c, op, acc);
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_MDN: mdn.py:185-272.  params: min_scale, pad4 ; MLP block.
// MLP outputs: logits[K], then per component k: loc[D], raw_scale[D].
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_mdn(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  // (VBN_F_MDNPLAIN ops of the tensor-core kernel never get here: run_ops hands them to TcMlp::mdn_plain_body)
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, K = op.k;
  const float min_scale = __ldg(P);
  const int32_t* par = c.a.par_slots + op.par_off;
  mlp_eval(c, op, P + 4, nullptr, par);
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;

  // pi = softmax(logits).clamp_min(1e-5); pi /= pi.sum().clamp_min(1e-12)   (mdn.py:227-228).
  // The clamped probabilities q_k = max(softmax_k, 1e-5) overwrite the logits in scratch;
  // pi_k = q_k / tot is formed where it is consumed.
  float tot[RPT];
  {
    float mx[RPT], se[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) mx[j] = -CUDART_INF_F;
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) mx[j] = fmaxf(mx[j], c.scr(k, j));
#pragma unroll
    for (int j = 0; j < RPT; ++j) se[j] = 0.0f;
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float e = expf(c.scr(k, j) - mx[j]);
        c.scr(k, j) = e;
        se[j] += e;
      }
#pragma unroll
    for (int j = 0; j < RPT; ++j) tot[j] = 0.0f;
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float q = fmaxf(__fdiv_rn(c.scr(k, j), se[j]), 1e-5f);
        c.scr(k, j) = q;
        tot[j] += q;
      }
#pragma unroll
    for (int j = 0; j < RPT; ++j) tot[j] = fmaxf(tot[j], 1e-12f);
  }

  if (sample) {
    int pick[RPT];
    if (!c.injected_index(op, 0, 1, pick)) {
      // k ~ Categorical(pi): inverse CDF on the un-normalised q_k with u scaled by tot
      float u[RPT];
      c.draw_uniform(op, op.u_off, 0, u);
      float cum[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) { cum[j] = 0.0f; pick[j] = K - 1; u[j] *= tot[j]; }
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const float before = cum[j];
          cum[j] += c.scr(k, j);
          if (u[j] >= before && u[j] < cum[j]) pick[j] = k;
        }
    }
    for (int d = 0; d < D; ++d) {
      float eps[RPT];
      c.draw_normal(op, op.n_off + d, d, eps);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int kk = min(max(pick[j], 0), K - 1);
        const float loc = c.scr(K + kk * 2 * D + d, j);
        const float sc = softplus20(c.scr(K + kk * 2 * D + D + d, j)) + min_scale;
        c.slot(op.out_slot + d, j) = fmaf(eps[j], sc, loc);
      }
    }
  }
  if (want_lp) {
    // t_k = log pi_k + log_comp_k, then logsumexp over k (mdn.py:264-272); t_k overwrites q_k
    float tmax[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) tmax[j] = -CUDART_INF_F;
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float pk = __fdiv_rn(c.scr(k, j), tot[j]);
        float q = 0.0f;
        for (int d = 0; d < D; ++d) {
          const float loc = c.scr(K + k * 2 * D + d, j);
          const float ls = logf(softplus20(c.scr(K + k * 2 * D + D + d, j)) + min_scale);
          const float var = expf(2.0f * ls);  // mdn.py:265
          q += gauss_term(c.slot(op.out_slot + d, j), loc, var, 2.0f * ls);
        }
        const float t = logf(pk) + (-0.5f * q);
        c.scr(k, j) = t;
        tmax[j] = fmaxf(tmax[j], t);
      }
    }
    float lp[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) lp[j] = 0.0f;
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) lp[j] += expf(c.scr(k, j) - tmax[j]);
#pragma unroll
    for (int j = 0; j < RPT; ++j) lp[j] = (tmax[j] == -CUDART_INF_F) ? -CUDART_INF_F : tmax[j] + logf(lp[j]);
    commit_logp(c, op, lp);
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_SNN: softmax_nn.py:581-759.
// params: {min_bin_width, within_bin_scale, temperature, 0}, bin_edges[D][C+1],
//         class_values[D][C], sample_values[D][C], is_discrete[D], pad4 ; MLP block
// MLP outputs: logits[D][C]
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_snn(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, C = op.k;
  const float min_bw = __ldg(P), wb_scale = __ldg(P + 1), temperature = __ldg(P + 2);
  const float* edges = P + 4;
  const float* class_values = edges + D * (C + 1);
  const float* sample_values = class_values + D * C;
  const float* is_disc = sample_values + D * C;
  const float* mlp = P + pad4(4 + D * (C + 1) + 2 * D * C + D);
  const int32_t* par = c.a.par_slots + op.par_off;
  const int within = op.aux[0];
  const bool clip = op.aux[1] != 0;
  mlp_eval(c, op, mlp, nullptr, par);
  if (op.n_layers > 0 && temperature != 1.0f) {
    for (int o = 0; o < D * C; ++o)
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.scr(o, j) = __fdiv_rn(c.scr(o, j), temperature);
  }
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;
  float lp[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) lp[j] = 0.0f;

  for (int d = 0; d < D; ++d) {
    const bool disc = __ldg(is_disc + d) != 0.0f;
    const float* e = edges + d * (C + 1);
    float mx[RPT], se[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) mx[j] = -CUDART_INF_F;
    for (int k = 0; k < C; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) mx[j] = fmaxf(mx[j], c.scr(d * C + k, j));
#pragma unroll
    for (int j = 0; j < RPT; ++j) se[j] = 0.0f;
    for (int k = 0; k < C; ++k)
#pragma unroll
      for (int j = 0; j < RPT; ++j) se[j] += expf(c.scr(d * C + k, j) - mx[j]);

    if (sample) {
      int pick[RPT];
      if (!c.injected_index(op, d, D, pick)) {  // Categorical(logits).sample() (softmax_nn.py:651)
        float u[RPT], cum[RPT];
        c.draw_uniform(op, op.u_off + d, d, u);
#pragma unroll
        for (int j = 0; j < RPT; ++j) { cum[j] = 0.0f; pick[j] = C - 1; u[j] *= se[j]; }
        for (int k = 0; k < C; ++k)
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            const float before = cum[j];
            cum[j] += expf(c.scr(d * C + k, j) - mx[j]);
            if (u[j] >= before && u[j] < cum[j]) pick[j] = k;
          }
      }
      float val[RPT];
      if (disc) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) val[j] = __ldg(sample_values + d * C + min(max(pick[j], 0), C - 1));
      }
      // the reference draws the within-bin variate for every dim, discrete or not
      // (softmax_nn.py:664-679); the stream index is consumed either way.
      float w[RPT];
      if (within == VBN_WB_GAUSSIAN) {
        c.draw_normal(op, op.n_off + d, d, w);
      } else {
        // injected u for within-bin shares the [Bn,S,D] layout of noise.u
        c.draw_uniform(op, op.u_off + D + d, d, w);
      }
      if (!disc) {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int kk = min(max(pick[j], 0), C - 1);
          const float left = __ldg(e + kk), right = __ldg(e + min(kk + 1, C));
          const float width = fmaxf(right - left, min_bw);
          float v;
          if (within == VBN_WB_UNIFORM) {
            v = fmaf(w[j], width, left);
          } else if (within == VBN_WB_TRIANGULAR) {
            const float lv = fmaf(width, sqrtf(fmaxf(w[j] * 0.5f, 0.0f)), left);
            const float rv = right - width * sqrtf(fmaxf((1.0f - w[j]) * 0.5f, 0.0f));
            v = w[j] < 0.5f ? lv : rv;
          } else {
            const float center = 0.5f * (left + right);
            v = fmaf(w[j], fmaxf(wb_scale * width, min_bw), center);
          }
          if (clip) v = fminf(fmaxf(v, left), right);
          val[j] = v;
        }
      }
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.slot(op.out_slot + d, j) = val[j];
    }

    if (want_lp) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float x = c.slot(op.out_slot + d, j);
        int bin;
        if (disc) {  // exact class match (softmax_nn.py:618-627)
          bin = -1;
          for (int k = C - 1; k >= 0; --k)
            if (x == __ldg(class_values + d * C + k)) bin = k;
          if (bin < 0) {
            if (c.rows.valid[j] && c.a.error_flag) atomicOr(c.a.error_flag, 1);
            bin = 0;
          }
        } else {     // (x >= edges).sum() - 1, clamped (softmax_nn.py:615-616)
          int cnt = 0;
          for (int k = 0; k <= C; ++k) cnt += (x >= __ldg(e + k)) ? 1 : 0;
          bin = min(max(cnt - 1, 0), C - 1);
        }
        const float log_bin = (c.scr(d * C + bin, j) - mx[j]) - logf(se[j]);  // log_softmax
        float log_within = 0.0f;
        if (!disc) {
          const float left = __ldg(e + bin), right = __ldg(e + min(bin + 1, C));
          const float width = fmaxf(right - left, min_bw);
          const float center = 0.5f * (left + right);
          const float xu = clip ? fminf(fmaxf(x, left), right) : x;
          const bool inside = (x >= left) && (x <= right);
          if (within == VBN_WB_UNIFORM) {
            log_within = -logf(width);
            if (!clip && !inside) log_within = -CUDART_INF_F;
          } else if (within == VBN_WB_TRIANGULAR) {
            const float dl = fmaxf(width * (center - left), min_bw * min_bw);
            const float dr = fmaxf(width * (right - center), min_bw * min_bw);
            const float lpdf = __fdiv_rn(2.0f * (xu - left), dl);
            const float rpdf = __fdiv_rn(2.0f * (right - xu), dr);
            const float pdf = fmaxf(xu <= center ? lpdf : rpdf, 0.0f);
            log_within = logf(fmaxf(pdf, 1e-12f));
            if (!clip && !inside) log_within = -CUDART_INF_F;
          } else {
            const float sigma = fmaxf(wb_scale * width, min_bw);
            const float dd = xu - center;
            log_within = -__fdiv_rn(dd * dd, 2.0f * sigma * sigma) - logf(sigma) - kHalfLog2Pi;
          }
        }
        lp[j] += log_bin + log_within;
      }
    }
  }
  if (want_lp) commit_logp(c, op, lp);
}

// ---------------------------------------------------------------------------------------
// VBN_OP_TAB: a softmax_nn node in discrete mode (softmax_nn.py:581-759, D = 1) whose parents are all
// discrete too.  Its logits depend only on the parent configuration, so the plan compiler evaluates
// the node's own log-density once per configuration (through this library's MLP path, i.e. the
// same numbers the per-row evaluation produces) and the node becomes a table lookup.
// The same op serves categorical_table CPDs (vbn/cpds/categorical_table.py:359-417), whose
// parameters ARE such a table; those set `strict`: a parent value outside its support raises
// (categorical_table.py:12-21), where a softmax_nn would have fed it to its MLP.
// params: {C, n_cfg, cpad, strict}, per parent p < Dp: {card, stride, 0, 0, class_values[cpad]},
//         sample_values[cpad], class_values[cpad], cdf[n_cfg][C] (running sums of exp(logp)),
//         logp[n_cfg][C]
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_tab(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int C = op.k, Dp = op.n_par;
  const int n_cfg = static_cast<int>(__ldg(P + 1)), cpad = static_cast<int>(__ldg(P + 2));
  const bool strict = __ldg(P + 3) != 0.0f;
  const float* pinfo = P + 4;
  const float* sample_values = pinfo + (4 + cpad) * Dp;
  const float* class_values = sample_values + cpad;
  const float* cdf = class_values + cpad;
  const float* logp = cdf + n_cfg * C;
  const int32_t* par = c.a.par_slots + op.par_off;
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;

  int cfg[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) cfg[j] = 0;
  for (int p = 0; p < Dp; ++p) {
    const float* pi = pinfo + (4 + cpad) * p;
    const int card = static_cast<int>(__ldg(pi)), stride = static_cast<int>(__ldg(pi + 1));
    const int ps = __ldg(par + p);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float v = c.slot(ps, j);
      int ci = -1;
      for (int q = 0; q < card; ++q)
        if (v == __ldg(pi + 4 + q)) ci = q;
      if (ci < 0) {  // softmax_nn tables: flagged where that parent is itself scored
        if (strict && c.rows.valid[j] && c.a.error_flag) atomicOr(c.a.error_flag, 1);
        ci = 0;
      }
      cfg[j] += ci * stride;
    }
  }
  if (sample) {
    int pick[RPT];
    if (!c.injected_index(op, 0, 1, pick)) {  // Categorical(logits).sample() (softmax_nn.py:651)
      float u[RPT];
      c.draw_uniform(op, op.u_off, 0, u);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float* row = cdf + cfg[j] * C;
        const float t = u[j] * __ldg(row + C - 1);
        int k = 0;
        for (int q = 0; q < C - 1; ++q) k += (t >= __ldg(row + q)) ? 1 : 0;
        pick[j] = k;
      }
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      c.slot(op.out_slot, j) = __ldg(sample_values + min(max(pick[j], 0), C - 1));
  }
  if (want_lp) {
    float lp[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float x = c.slot(op.out_slot, j);
      int bin = -1;  // exact class match (softmax_nn.py:618-627)
      for (int q = C - 1; q >= 0; --q)
        if (x == __ldg(class_values + q)) bin = q;
      if (bin < 0) {
        if (c.rows.valid[j] && c.a.error_flag) atomicOr(c.a.error_flag, 1);
        bin = 0;
      }
      lp[j] = __ldg(logp + cfg[j] * C + bin);
    }
    commit_logp(c, op, lp);
  }
}

// ---------------------------------------------------------------------------------------
// VBN_F_OUT_PARAMS: write the op's conditional-distribution parameters for every row instead of
// drawing it -- the numbers CPDHandle.conditional() reports (vbn/core/cpd_handle.py:40-118, 348-402):
//   LG / GNN : loc[D], scale[D]                       ("normal_params")
//   MDN      : softmax(logits)[K], loc[K][D], scale[K][D]   ("mixture_params"; unclamped softmax)
//   SNN      : softmax(logits)[D][C]                  ("categorical_probs")
// element i of row r goes to stores[store_idx] (r, i).
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_params(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, Dp = op.n_par;
  const int32_t* par = c.a.par_slots + op.par_off;
  const VbnView v = c.a.stores[op.store_idx];
  auto put = [&](int i, int j, float x) {
    if (c.rows.valid[j]) v.base[c.rows.r[j] * v.row_stride + i * v.dim_stride] = x;
  };
  if (op.kind == VBN_OP_LG) {
    const float* W = P;
    const float* bias = W + Dp * D;
    const float* scale = bias + D;
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        float loc = 0.0f;
        for (int p = 0; p < Dp; ++p) loc = fmaf(c.slot(__ldg(par + p), j), __ldg(W + p * D + d), loc);
        put(d, j, loc + __ldg(bias + d));
        put(D + d, j, __ldg(scale + d));
      }
  } else if (op.kind == VBN_OP_GNN) {
    const float* mean_y = P + 2 * Dp;
    const float* std_y = mean_y + D;
    const float min_scale = __ldg(std_y + D);
    mlp_eval(c, op, P + pad4(2 * Dp + 2 * D + 1), P, par);
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float sy = __ldg(std_y + d);
        put(d, j, fmaf(c.scr(d, j), sy, __ldg(mean_y + d)));
        put(D + d, j, (softplus20(c.scr(D + d, j)) + min_scale) * sy);
      }
  } else if (op.kind == VBN_OP_RFF) {
    const float* tail = P + 4 + pad4(2 * Dp);
    rff_loc(c, op, P);
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        put(d, j, c.scr(Dp + d, j));
        put(D + d, j, __ldg(tail + 3 * D + d));
      }
  } else if (op.kind == VBN_OP_MDN) {
    const int K = op.k;
    const float min_scale = __ldg(P);
    mlp_eval(c, op, P + 4, nullptr, par);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      float mx = -CUDART_INF_F, se = 0.0f;
      for (int k = 0; k < K; ++k) mx = fmaxf(mx, c.scr(k, j));
      for (int k = 0; k < K; ++k) se += expf(c.scr(k, j) - mx);
      for (int k = 0; k < K; ++k) put(k, j, __fdiv_rn(expf(c.scr(k, j) - mx), se));
      for (int k = 0; k < K; ++k)
        for (int d = 0; d < D; ++d) {
          put(K + k * D + d, j, c.scr(K + k * 2 * D + d, j));
          put(K + K * D + k * D + d, j, softplus20(c.scr(K + k * 2 * D + D + d, j)) + min_scale);
        }
    }
  } else if (op.kind == VBN_OP_TAB) {  // exp(logp[cfg][:]) of the row's parent configuration
    const int C = op.k;
    const int n_cfg = static_cast<int>(__ldg(P + 1)), cpad = static_cast<int>(__ldg(P + 2));
    const float* pinfo = P + 4;
    const float* logp = pinfo + (4 + cpad) * Dp + 2 * cpad + n_cfg * C;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      int cfg = 0;
      for (int p = 0; p < Dp; ++p) {
        const float* pi = pinfo + (4 + cpad) * p;
        const int card = static_cast<int>(__ldg(pi)), stride = static_cast<int>(__ldg(pi + 1));
        const float v = c.slot(__ldg(par + p), j);
        int ci = 0;
        for (int q = 1; q < card; ++q)
          if (v == __ldg(pi + 4 + q)) ci = q;
        cfg += ci * stride;
      }
      for (int k = 0; k < C; ++k) put(k, j, expf(__ldg(logp + cfg * C + k)));
    }
  } else if (op.kind == VBN_OP_SNN) {
    const int C = op.k;
    const float temperature = __ldg(P + 2);
    mlp_eval(c, op, P + pad4(4 + D * (C + 1) + 2 * D * C + D), nullptr, par);
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      for (int d = 0; d < D; ++d) {
        float mx = -CUDART_INF_F, se = 0.0f;
        const float it = (op.n_layers > 0 && temperature != 1.0f) ? temperature : 1.0f;
        for (int k = 0; k < C; ++k) mx = fmaxf(mx, __fdiv_rn(c.scr(d * C + k, j), it));
        for (int k = 0; k < C; ++k) se += expf(__fdiv_rn(c.scr(d * C + k, j), it) - mx);
        for (int k = 0; k < C; ++k) put(d * C + k, j, __fdiv_rn(expf(__fdiv_rn(c.scr(d * C + k, j), it) - mx), se));
      }
  }
}

// VBN_F_TABPLAIN: a TAB op with <= 4 parents and <= 4 classes whose parents' and own classes are coded 0..k-1, so
// the class index IS the value: no per-parent search, no value gather, no parameter-block header reads.  Two uses:
//   drawn   (VBN_SRC_SAMPLE, no store / density): k = #{q < C-1 : u * total >= cdf_q}    -- one 128-bit row load of
//           the padded table cdf4[n_cfg] = {c_0, .., c_{C-2}, +inf.., total}
//   scored  (VBN_SRC_FIXED_Q | VBN_F_ADD_LOGW: an evidence node): logw += logp4[cfg][class of the fixed value]
// Everything else rides in quads 0,4,5,6.  The body is one short straight-line block for every parent / class
// count (BASELINE cfg3: 37 of 37 ALARM ops): the per-count variants of the first version made the hot loop spill out
// of the instruction cache (a quarter of the stall samples were instruction fetch).
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_tab_plain(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const int Dp = op.n_par;
  const int C = op.layer_dim[5] & 0xFFFF;
  const bool strict = (op.layer_dim[5] >> 16) != 0;
  const float4* tab = reinterpret_cast<const float4*>(c.a.params + op.layer_dim[4]);
  int cfg[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) cfg[j] = 0;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (p < Dp) {
      const int ps = op.aux[p];
      const int stride = op.layer_dim[p] & 0xFFFF, card = (op.layer_dim[p] >> 16) & 0x7FFF;
      if (op.layer_dim[p] < 0) {  // bit 31: the parent is drawn by a plain table op of this schedule -- always a valid index
#pragma unroll
        for (int j = 0; j < RPT; ++j) cfg[j] += __float2int_rz(c.slot(ps, j)) * stride;
      } else {
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const float v = c.slot(ps, j);
          const int ci = __float2int_rz(v);
          const bool ok = static_cast<float>(ci) == v && static_cast<unsigned>(ci) < static_cast<unsigned>(card);
          if (!ok && strict && c.rows.valid[j] && c.a.error_flag) atomicOr(c.a.error_flag, 1);
          cfg[j] += (ok ? ci : 0) * stride;
        }
      }
    }
  }
  const int out_slot = op.layer_dim[6];
  if ((op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE) {
    float uu[RPT];
    c.cached_uniform((op.flags & VBN_F_SHARED) != 0, op.layer_dim[7], uu);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float4 r = __ldg(tab + cfg[j]);
      const float t = uu[j] * r.w;
      // class index as a float straight from the three compares (FSET.BF gives 1.0 / 0.0): no integer sum, no
      // int -> float conversion
      c.slot(out_slot, j) = ((t >= r.x) ? 1.0f : 0.0f) + ((t >= r.y) ? 1.0f : 0.0f) + ((t >= r.z) ? 1.0f : 0.0f);
    }
  } else {  // evidence: the per-query value is its own class index (softmax_nn.py:618-627 exact class match)
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float x = __ldg(c.a.fixed + static_cast<int64_t>(op.layer_dim[7]) * c.a.n_queries + c.rows.lb[j]);
      c.slot(out_slot, j) = x;
      const int ci = __float2int_rz(x);
      const bool ok = static_cast<float>(ci) == x && static_cast<unsigned>(ci) < static_cast<unsigned>(C);
      if (!ok && c.rows.valid[j] && c.a.error_flag) atomicOr(c.a.error_flag, 1);
      const float4 r = __ldg(tab + cfg[j]);
      const int b = ok ? ci : 0;
      const float lo = (b & 1) ? r.y : r.x, hi = (b & 1) ? r.w : r.z;
      c.rows.logw[j] += (b & 2) ? hi : lo;
    }
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_RFF: rff_gaussian.py:131-146 (_normalize_parents, _features), 185-206 (_params),
// 254-291 (sample, log_prob).
// params: {F, sqrt(2/F), 0, 0}, mean_x[Dp], std_x[Dp] (pad4), {bias, mean_y, std_y, scale,
//         2 ln scale, var}[D] (pad4), then per feature f: w_f[Dp], b_f  (stride Dp+1), coef[F][D]
// loc lands in scratch rows Dp..Dp+D-1 (rows 0..Dp-1 hold the standardised parents).
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void rff_loc(Ctx<RPT, NT, TC>& c, const VbnOp& op, const float* P) {
  const int D = op.dim, Dp = op.n_par;
  const int F = static_cast<int>(__ldg(P));
  const float phi = __ldg(P + 1);
  const float* mean_x = P + 4;
  const float* std_x = mean_x + Dp;
  const float* tail = P + 4 + pad4(2 * Dp);
  const float* wb = tail + pad4(6 * D);
  const float* coef = wb + F * (Dp + 1);
  const int32_t* par = c.a.par_slots + op.par_off;
  for (int p = 0; p < Dp; ++p) {
    const float m = __ldg(mean_x + p), sd = __ldg(std_x + p);
    const int ps = __ldg(par + p);
#pragma unroll
    for (int j = 0; j < RPT; ++j) c.scr(p, j) = __fdiv_rn(c.slot(ps, j) - m, sd);
  }
  for (int d0 = 0; d0 < D; d0 += 4) {
    float acc[RPT][4];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[j][q] = 0.0f;
    for (int f = 0; f < F; ++f) {
      const float* w = wb + f * (Dp + 1);
      float proj[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) proj[j] = 0.0f;
      for (int p = 0; p < Dp; ++p) {
        const float wp = __ldg(w + p);
#pragma unroll
        for (int j = 0; j < RPT; ++j) proj[j] = fmaf(c.scr(p, j), wp, proj[j]);
      }
      const float bf = __ldg(w + Dp);
      float cf[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) cf[q] = d0 + q < D ? __ldg(coef + f * D + d0 + q) : 0.0f;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const float ph = phi * cosf(proj[j] + bf);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[j][q] = fmaf(ph, cf[q], acc[j][q]);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (d0 + q < D) {
        const int d = d0 + q;
        const float b = __ldg(tail + d), my = __ldg(tail + D + d), sy = __ldg(tail + 2 * D + d);
#pragma unroll
        for (int j = 0; j < RPT; ++j) c.scr(Dp + d, j) = fmaf(acc[j][q] + b, sy, my);
      }
  }
}

template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_rff(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, Dp = op.n_par;
  const float* tail = P + 4 + pad4(2 * Dp);
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;
  rff_loc(c, op, P);
  if (sample) {
    for (int d = 0; d < D; ++d) {
      float eps[RPT];
      c.draw_normal(op, op.n_off + d, d, eps);
      const float sc = __ldg(tail + 3 * D + d);
#pragma unroll
      for (int j = 0; j < RPT; ++j) c.slot(op.out_slot + d, j) = fmaf(eps[j], sc, c.scr(Dp + d, j));
    }
  }
  if (want_lp) {
    float acc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] = 0.0f;
    for (int d = 0; d < D; ++d) {
      const float t = __ldg(tail + 4 * D + d), v = __ldg(tail + 5 * D + d);
#pragma unroll
      for (int j = 0; j < RPT; ++j) acc[j] += gauss_term(c.slot(op.out_slot + d, j), c.scr(Dp + d, j), v, t);
    }
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] *= -0.5f;
    commit_logp(c, op, acc);
  }
}

// ---------------------------------------------------------------------------------------
// VBN_OP_KDE: kde.py:105-182.
// params: {hy, hp, const_y, noise_scale, log_n, 0,0,0}, parents[N][Dp] (pad4), targets[N][D]
//   hy = 0.5/s_y^2, hp = 0.5/s_p^2  ->  log_k = -h*diff^2 + const
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_kde(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const float* P = c.a.params + op.param_off;
  const int D = op.dim, Dp = op.n_par, N = op.k;
  const float hy = __ldg(P), hp = __ldg(P + 1), const_y = __ldg(P + 2), noise_scale = __ldg(P + 3);
  const float log_n = __ldg(P + 4);
  const float* tp = P + 8;
  const float* ty = tp + pad4(N * Dp);
  const int32_t* par = c.a.par_slots + op.par_off;
  const bool sample = (op.flags & VBN_SRC_MASK) == VBN_SRC_SAMPLE;
  const bool want_lp = (op.flags & (VBN_F_ADD_LOGW | VBN_F_OUT_LOGP)) != 0;

  if (sample) {
    int pick[RPT];
    if (!c.injected_index(op, 0, 1, pick)) {
      float u[RPT];
      c.draw_uniform(op, op.u_off, 0, u);
      if (Dp == 0) {  // torch.randint(0, N) (kde.py:170)
#pragma unroll
        for (int j = 0; j < RPT; ++j) pick[j] = min(static_cast<int>(u[j] * N), N - 1);
      } else {        // multinomial(softmax(log_kp)) (kde.py:172-178): two passes, inverse CDF
        Lse acc[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) acc[j].init();
        for (int n = 0; n < N; ++n)
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            float q = 0.0f;
            for (int p = 0; p < Dp; ++p) {
              const float df = c.slot(__ldg(par + p), j) - __ldg(tp + n * Dp + p);
              q = fmaf(df, df, q);
            }
            acc[j].push(-hp * q);
          }
        float cum[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) { cum[j] = 0.0f; pick[j] = N - 1; u[j] *= acc[j].l; }
        for (int n = 0; n < N; ++n)
#pragma unroll
          for (int j = 0; j < RPT; ++j) {
            float q = 0.0f;
            for (int p = 0; p < Dp; ++p) {
              const float df = c.slot(__ldg(par + p), j) - __ldg(tp + n * Dp + p);
              q = fmaf(df, df, q);
            }
            const float before = cum[j];
            cum[j] += __expf(-hp * q - acc[j].m);
            if (u[j] >= before && u[j] < cum[j]) pick[j] = n;
          }
      }
    }
    for (int d = 0; d < D; ++d) {
      float eps[RPT];
      c.draw_normal(op, op.n_off + d, d, eps);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int nn = min(max(pick[j], 0), N - 1);
        c.slot(op.out_slot + d, j) = fmaf(eps[j], noise_scale, __ldg(ty + nn * D + d));
      }
    }
  }
  if (want_lp) {
    // two-level logsumexp: blocks of 256 stored points are reduced on their own and then merged,
    // so the fp32 summation error does not grow with sqrt(N) (matters from N ~ 1e5 on)
    Lse num[RPT], den[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) { num[j].init(); den[j].init(); }
    for (int n0 = 0; n0 < N; n0 += 256) {
      const int n1 = min(n0 + 256, N);
      Lse bn[RPT], bd[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) { bn[j].init(); bd[j].init(); }
      for (int n = n0; n < n1; ++n)
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          float qp = 0.0f, qy = 0.0f;
          for (int p = 0; p < Dp; ++p) {
            const float df = c.slot(__ldg(par + p), j) - __ldg(tp + n * Dp + p);
            qp = fmaf(df, df, qp);
          }
          for (int d = 0; d < D; ++d) {
            const float df = c.slot(op.out_slot + d, j) - __ldg(ty + n * D + d);
            qy = fmaf(df, df, qy);
          }
          const float a = -hp * qp;
          bd[j].push(a);
          bn[j].push(fmaf(-hy, qy, a));
        }
#pragma unroll
      for (int j = 0; j < RPT; ++j) { num[j].merge(bn[j]); den[j].merge(bd[j]); }
    }
    float lp[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      lp[j] = (Dp == 0) ? (num[j].value() + const_y - log_n) : (num[j].value() - den[j].value() + const_y);
    commit_logp(c, op, lp);
  }
}

// ---------------------------------------------------------------------------------------
// the kernels' shared body
// ---------------------------------------------------------------------------------------
// Gibbs sweeps (sampling/gibbs.py:38-82): candidate / score ops are ordinary CPD ops; these three
// tie them together.
//   VBN_OP_TAKEW : slot[out_slot] = accumulated log-weight of the row; log-weight = 0
//   VBN_OP_SELECT: among k candidates (values in slots aux[1] + c*dim + d, scores in slots aux[0] + c),
//                  draw c ~ softmax(scores) (gibbs.py:76-77) and copy its value to out_slot
//   VBN_OP_JUMP  : back to op aux[0] until the loop has run k times; layer_dim[0..1] = Philox stream
//                  blocks (normal, uniform) one iteration consumes
// ---------------------------------------------------------------------------------------
template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_takew(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    c.slot(op.out_slot, j) = c.rows.logw[j];
    c.rows.logw[j] = 0.0f;
  }
}

template <int RPT, int NT, class TC>
__device__ __forceinline__ void op_select(Ctx<RPT, NT, TC>& c, const VbnOp& op) {
  const int K = __ldg(&c.gop->k), D = op.dim;  // read from the descriptor: the light kernels do not fetch quad 3
  const int s_score = __ldg(&c.gop->aux[0]), s_cand = __ldg(&c.gop->aux[1]);
  int pick[RPT];
  if (!c.injected_index(op, 0, 1, pick)) {
    float u[RPT];
    c.draw_uniform(op, op.u_off, 0, u);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      float mx = -CUDART_INF_F;
      for (int k = 0; k < K; ++k) mx = fmaxf(mx, c.slot(s_score + k, j));
      float tot = 0.0f;
      for (int k = 0; k < K; ++k) tot += expf(c.slot(s_score + k, j) - mx);
      const float t = u[j] * tot;
      float cum = 0.0f;
      pick[j] = K - 1;
      for (int k = 0; k < K; ++k) {
        const float before = cum;
        cum += expf(c.slot(s_score + k, j) - mx);
        if (t >= before && t < cum) { pick[j] = k; break; }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int k = min(max(pick[j], 0), K - 1);
    for (int d = 0; d < D; ++d) c.slot(op.out_slot + d, j) = c.slot(s_cand + k * D + d, j);
  }
}

// ---------------------------------------------------------------------------------------
// Fused weight reduction (replaces the first pass of torch.softmax(log_weights, 1) -- importance_sampling.py:82-84,
// likelihood_weighting.py:75-80 -- and, when a value slot is given, the sums behind VBN._posterior_stats,
// vbn/vbn.py:495-503, and the benchmark adapter's class histogram, benchmarking/models/vbn.py:202-242).
// The 32 lanes of a warp hold 32 consecutive rows; every run of rows of one query inside them becomes one record
//   [0] m = max logw   [1] l = sum e   [2] q = sum e^2   [3] x0   [4] sum e (x - x0)   [5] sum e (x - x0)^2
//   [6] rows   [7] -   [8..15] sum e [x == k]          with e = exp(logw - m)
// at seg[(b * seg_per_query + (r / 32 - b S / 32)) * kSegWords]; vbn_segment_merge folds a query's records.
// ---------------------------------------------------------------------------------------
constexpr int kSegWords = 16;

template <int RPT, int NT, class TC>
__device__ __forceinline__ void emit_segments(Ctx<RPT, NT, TC>& c) {
  const ScheduleArgs& a = c.a;
  const bool moments = a.seg_slot >= 0;
  const int classes = a.seg_classes;
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const float lw = c.rows.logw[j];
    const float x = moments ? c.slot(a.seg_slot, j) : 0.0f;
    const int64_t b = c.rows.lb[j];
    const int64_t seg_index = (c.rows.r[j] >> 5) - ((b * a.n_samples) >> 5);
    float* rec = a.seg + (b * a.seg_per_query + seg_index) * kSegWords;
#ifdef VBN_HOST_EMU
    // host emulation: threads run one after another, so each row is merged into its (zero-initialised) record
    if (!c.rows.valid[j]) continue;
    if (rec[6] == 0.0f) {
      rec[0] = lw; rec[1] = lw > -CUDART_INF_F ? 1.0f : 0.0f; rec[2] = rec[1]; rec[3] = x; rec[4] = 0.0f; rec[5] = 0.0f;
      for (int k = 0; k < 8; ++k) rec[8 + k] = (k < classes && x == static_cast<float>(k)) ? rec[1] : 0.0f;
    } else {
      const float m = fmaxf(rec[0], lw);
      const float so = rec[0] > -CUDART_INF_F ? expf(rec[0] - m) : 0.0f, e = lw > -CUDART_INF_F ? expf(lw - m) : 0.0f;
      const float dx = x - rec[3];
      rec[0] = m; rec[1] = rec[1] * so + e; rec[2] = rec[2] * so * so + e * e;
      rec[4] = rec[4] * so + e * dx; rec[5] = rec[5] * so + e * dx * dx;
      for (int k = 0; k < classes; ++k) rec[8 + k] = rec[8 + k] * so + (x == static_cast<float>(k) ? e : 0.0f);
    }
    rec[6] += 1.0f;
#else
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(full, c.rows.valid[j]);
    while (todo) {  // one pass per query present among the warp's rows (one, unless S is small or a query ends here)
      const int leader = __ffs(todo) - 1;
      const int64_t b0 = __shfl_sync(full, b, leader);
      const bool mine = c.rows.valid[j] && b == b0;
      todo &= ~__ballot_sync(full, mine);
      float m = mine ? lw : -CUDART_INF_F;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(full, m, o));
      const float e = (mine && lw > -CUDART_INF_F) ? __expf(lw - m) : 0.0f;
      const float x0 = __shfl_sync(full, x, leader);
      const float dx = x - x0;
      float l = e, q = e * e, sx = e * dx, sxx = e * dx * dx, n = mine ? 1.0f : 0.0f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        l += __shfl_xor_sync(full, l, o);
        q += __shfl_xor_sync(full, q, o);
        n += __shfl_xor_sync(full, n, o);
      }
      if (moments) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sx += __shfl_xor_sync(full, sx, o);
          sxx += __shfl_xor_sync(full, sxx, o);
        }
      }
      float h = 0.0f;  // lane k ends up with class k's sum
      for (int k = 0; k < classes; ++k) {
        float hk = (x == static_cast<float>(k)) ? e : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hk += __shfl_xor_sync(full, hk, o);
        if (lane == k) h = hk;
      }
      float* r0 = reinterpret_cast<float*>(__shfl_sync(full, reinterpret_cast<unsigned long long>(rec), leader));
      if (lane == leader) {
        reinterpret_cast<float4*>(r0)[0] = make_float4(m, l, q, x0);
        reinterpret_cast<float4*>(r0)[1] = make_float4(sx, sxx, n, 0.0f);
      }
      if (lane < 8) r0[8 + lane] = h;
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------
// binds row j of the thread to local row index r (clamped into range; `valid` masks the stores)
template <class C>
__device__ __forceinline__ void bind_row(C& c, int j, int64_t r) {
  const ScheduleArgs& a = c.a;
  c.rows.valid[j] = r < a.n_rows;
  const int64_t rc = c.rows.valid[j] ? r : a.n_rows - 1;
  c.rows.r[j] = rc;
  const int64_t b = rc / a.n_samples;
  c.rows.lb[j] = b;
  c.rows.ls[j] = rc - b * a.n_samples;
  c.rows.gb[j] = a.query_offset + static_cast<uint32_t>(b);
  c.rows.gs[j] = a.sample_offset + static_cast<uint32_t>(c.rows.ls[j]);
  c.rows.logw[j] = 0.0f;
  c.rows.logp[j] = 0.0f;
}

// row j of the thread = its row j-1 + STEP, for n_samples >= STEP: the (query, sample) pair moves by at most one query,
// so no division is needed (the 64-bit one costs ~100 instructions per row: 13 % of a 37-op walk at 4 rows per thread)
template <int STEP, class C>
__device__ __forceinline__ void bind_next_row(C& c, int j) {
  const ScheduleArgs& a = c.a;
  const int64_t r = c.rows.r[j - 1] + STEP;  // (row j-1 is only clamped when it is itself out of range)
  c.rows.valid[j] = c.rows.valid[j - 1] && r < a.n_rows;
  if (c.rows.valid[j]) {
    int64_t b = c.rows.lb[j - 1], s = c.rows.ls[j - 1] + STEP;
    if (s >= a.n_samples) {
      s -= a.n_samples;
      ++b;
    }
    c.rows.r[j] = r;
    c.rows.lb[j] = b;
    c.rows.ls[j] = s;
  } else {  // clamp to the last row, like bind_row
    c.rows.r[j] = a.n_rows - 1;
    c.rows.lb[j] = a.n_queries - 1;
    c.rows.ls[j] = a.n_samples - 1;
  }
  c.rows.gb[j] = a.query_offset + static_cast<uint32_t>(c.rows.lb[j]);
  c.rows.gs[j] = a.sample_offset + static_cast<uint32_t>(c.rows.ls[j]);
  c.rows.logw[j] = 0.0f;
  c.rows.logp[j] = 0.0f;
}

// walks every op of the schedule for the rows bound to this thread, then writes logw / logp
template <bool HEAVY, int RPT, int NT, class TC>
__device__ __forceinline__ void run_ops(Ctx<RPT, NT, TC>& c) {
  const ScheduleArgs& a = c.a;
  c.rows.cur_nq = -1;
  c.rows.cur_uq = -1;
  c.rows.loop_iter = 0;
  c.rows.loop_nq = 0;
  c.rows.loop_uq = 0;
  for (int i = 0; i < a.n_ops; ++i) {
    VbnOp op;  // only the 16-byte quads this op kind reads are fetched (layer_dim: via c.gop)
    {
      const int4* src = reinterpret_cast<const int4*>(a.ops + i);
      // The plain ops get their own descriptor object: `op` below is indexed dynamically by the generic
      // paths and therefore lives in local memory, which would cost the hot ops a store per quad.
      if constexpr (TC::kEnabled) {
        // Tensor-core kernel: the whole 128-byte descriptor is requested in ONE round of independent loads.
        // Fetching the kind-specific quads only after the flag test made every op pay two dependent L1
        // round trips (~125 + ~210 cycles measured: 17 % of the kernel's stall samples).
        // Ops that follow an MLP op are served from the descriptor tail of its weight image in shared memory.
        const int4* dsc = c.tc.next_desc();
        const int4 q0 = dsc[0];  // kind, flags, dim, n_par -- decides which other quads the op reads (a second
                                 // round trip, ~30 cycles from shared memory; loading all eight up front costs 32
                                 // registers at the top of every op)
        if (q0.y & (VBN_F_LGPLAIN | VBN_F_MDNROOT | VBN_F_MDNPLAIN)) {
          VbnOp lop;
          int4* ld = reinterpret_cast<int4*>(&lop);
          ld[0] = q0;
          c.gop = a.ops + i;
          if (q0.y & VBN_F_LGPLAIN) {
            ld[4] = dsc[4]; ld[5] = dsc[5]; ld[6] = dsc[6];
            op_lg_plain(c, lop);
          } else if (q0.y & VBN_F_MDNROOT) {
            ld[4] = dsc[4]; ld[5] = dsc[5]; ld[6] = dsc[6]; ld[7] = dsc[7];
            op_mdn_root(c, lop);
          } else if (q0.y & VBN_F_MDNFAST) {
            // the reference's default MDN (K = 3, ReLU, first layer on the FP32 pipe): three quads instead of six --
            // out_slot and the tail count ride in tc[1], which the kernel does not read otherwise
            ld[2] = dsc[2]; ld[6] = dsc[6]; ld[7] = dsc[7];
            lop.out_slot = lop.tc[1] & 0xFFFF;
            lop.layer_dim[7] = static_cast<int>(static_cast<uint32_t>(lop.tc[1]) >> 16);
            c.tc.template mdn_plain_body<true>(c, lop);
            store_value(c, lop);
          } else {
            ld[1] = dsc[1]; ld[2] = dsc[2]; ld[3] = dsc[3]; ld[5] = dsc[5]; ld[6] = dsc[6]; ld[7] = dsc[7];
            c.tc.template mdn_plain_body<false>(c, lop);
            store_value(c, lop);  // a plain node may still be the stored target
          }
          continue;
        }
        const int4 q1 = dsc[1], q2 = dsc[2], q3 = dsc[3], q4 = dsc[4], q5 = dsc[5], q6 = dsc[6], q7 = dsc[7];
        int4* dst = reinterpret_cast<int4*>(&op);
        dst[0] = q0; dst[1] = q1; dst[2] = q2; dst[3] = q3; dst[4] = q4; dst[5] = q5; dst[6] = q6; dst[7] = q7;
      } else {
      // quads 4-6 (what a plain LG op needs) are requested together with quad 0: one L1 round trip per op
      const int4 q0 = __ldg(src + 0);  // kind, flags, dim, n_par
      const int4 q4 = __ldg(src + 4), q5 = __ldg(src + 5), q6 = __ldg(src + 6);
      if (q0.y & VBN_F_LGPLAIN) {
        VbnOp lop;
        int4* ld = reinterpret_cast<int4*>(&lop);
        ld[0] = q0;
        ld[4] = q4;  // bias, scale, 2 ln scale, var
        ld[5] = q5;  // w0..w3
        ld[6] = q6;  // packed parent slots, out_slot, n_off
        c.gop = a.ops + i;
        op_lg_plain(c, lop);
        continue;
      }
      // (only in the <= 2 rows-per-thread shapes, which is where schedules with table ops are placed: the
      // 4-row linear-Gaussian shapes keep their loop body small; there the op takes the generic lookup)
      if (TC::kTab && (q0.y & VBN_F_TABPLAIN)) {
        VbnOp lop;
        int4* ld = reinterpret_cast<int4*>(&lop);
        ld[0] = q0;
        ld[4] = q4;  // stride | card << 16 per parent
        ld[5] = q5;  // table offset, C | strict << 16, out_slot, u_off (drawn) / fixed_col (scored)
        ld[6] = q6;  // packed parent slots
        c.gop = a.ops + i;
        op_tab_plain(c, lop);
        continue;
      }
      if (HEAVY && (q0.y & VBN_F_MDNROOT)) {
        VbnOp lop;
        int4* ld = reinterpret_cast<int4*>(&lop);
        ld[0] = q0;
        ld[4] = __ldg(src + 4);  // cumulative weights, (loc, scale) pairs ...
        ld[5] = __ldg(src + 5);
        ld[6] = __ldg(src + 6);
        ld[7] = __ldg(src + 7);  // out_slot, n_off, u_off, K
        c.gop = a.ops + i;
        op_mdn_root(c, lop);
        continue;
      }
      int4* dst = reinterpret_cast<int4*>(&op);
      dst[0] = q0;
      dst[1] = __ldg(src + 1);  // out_slot, par_off, param_off, fixed_col
      dst[2] = __ldg(src + 2);  // store_idx, noise_idx, n_off, u_off
      if (op.kind == VBN_OP_LG) {
        if (op.flags & VBN_F_LGFAST) {
          dst[4] = __ldg(src + 4);  // bias, scale, 2 ln scale, var
          dst[5] = __ldg(src + 5);  // w0..w3
          dst[6] = __ldg(src + 6);  // parent slots
        }
      } else if (op.kind == VBN_OP_TAB) {
        dst[3] = __ldg(src + 3);  // k = classes
      } else if (HEAVY && op.kind >= VBN_OP_GNN) {
        dst[3] = __ldg(src + 3);  // n_layers, act, n_out, k
        dst[6] = __ldg(src + 6);  // aux
        dst[7] = __ldg(src + 7);  // tc
      }
      }  // !TC::kEnabled
    }
    c.gop = a.ops + i;
    if (TC::kLoops && op.kind >= VBN_OP_TAKEW) {  // Gibbs glue ops: FFMA-pipe HEAVY kernels only (plan.compile_gibbs)
      if (op.kind == VBN_OP_TAKEW) {
        op_takew(c, op);
      } else if (op.kind == VBN_OP_SELECT) {
        op_select(c, op);
        store_value(c, op);
      } else if (op.kind == VBN_OP_JUMP) {
        const int k = __ldg(&c.gop->k);
        c.rows.loop_nq = __ldg(&c.gop->layer_dim[0]);
        c.rows.loop_uq = __ldg(&c.gop->layer_dim[1]);
        if (c.rows.loop_iter + 1 < k) {
          ++c.rows.loop_iter;
          c.rows.cur_nq = -1;
          c.rows.cur_uq = -1;
          i = __ldg(&c.gop->aux[0]) - 1;
        }
      }
      continue;
    }
    if (HEAVY && (op.flags & VBN_F_OUT_PARAMS)) {
      op_params(c, op);
      continue;
    }
    load_fixed(c, op);
    switch (op.kind) {
      case VBN_OP_LG:
        if (op.flags & VBN_F_LGFAST) op_lg_fast(c, op); else op_lg(c, op);
        break;
      case VBN_OP_GNN: if (HEAVY) op_gnn(c, op); break;
      case VBN_OP_MDN: if (HEAVY) op_mdn(c, op); break;
      case VBN_OP_SNN: if (HEAVY) op_snn(c, op); break;
      case VBN_OP_KDE: if (HEAVY) op_kde(c, op); break;
      case VBN_OP_RFF: if (HEAVY) op_rff(c, op); break;
      case VBN_OP_TAB: if (TC::kTab) op_tab(c, op); break;
      default: break;
    }
    store_value(c, op);
  }
  if (a.logw) {
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (c.rows.valid[j]) a.logw[c.rows.r[j]] = (a.logw_accumulate ? a.logw[c.rows.r[j]] : 0.0f) + c.rows.logw[j];
  }
  if (a.logp) {
#pragma unroll
    for (int j = 0; j < RPT; ++j)
      if (c.rows.valid[j]) a.logp[c.rows.r[j]] = a.logp_as_pdf ? expf(c.rows.logp[j]) : c.rows.logp[j];
  }
  if (a.seg) emit_segments(c);
}

template <int RPT, int NT, bool HEAVY, int MIN_BLOCKS, bool TAB = true>
__global__ void __launch_bounds__(NT, MIN_BLOCKS) schedule_kernel(const ScheduleArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int ROWS = RPT * NT;
  // The thread index and the Philox counter words are made opaque to the optimiser (empty asm): otherwise ptxas
  // rematerialises them -- S2R tid in front of every slot access, the row -> (query, sample) division chain in
  // front of every generator refill -- instead of keeping one register each (same finding as in the tcgen05 kernel).
  int tid = threadIdx.x;
  asm volatile("" : "+r"(tid));
  // The thread's column of the slot area as ONE opaque pointer that is still known to point to shared memory:
  // otherwise ptxas rebuilds the address of every slot access from the shared-window base (S2UR CgaCtaId, UMOV,
  // ULEA) plus the thread index -- five instructions at the head of each parent read instead of one multiply-add.
  float* col = smem + tid;
#ifdef __CUDA_ARCH__
  asm volatile("" : "+l"(col));
  __builtin_assume(__isShared(col));
#endif
  Ctx<RPT, NT, typename std::conditional<HEAVY, NoTc, LightPolicyT<TAB>>::type> c(a, col, 0);
  const int64_t n_tiles = (a.n_rows + ROWS - 1) / ROWS;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * ROWS;
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (j == 0 || a.n_samples < NT) {
        bind_row(c, j, base + j * NT + tid);  // 64-bit division: row -> (query, sample)
      } else {
        bind_next_row<NT>(c, j);              // row j = row j-1 + NT < one query further: a compare instead
      }
      asm volatile("" : "+r"(c.rows.gs[j]), "+r"(c.rows.gb[j]));
    }
    run_ops<HEAVY>(c);
  }
}

}  // namespace vbn
