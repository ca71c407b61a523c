/*
 * vbn_cuda.h -- C ABI of libvbn_cuda.so, the B200 (sm_100a) backend for VBN's batched
 * posterior-inference hot path.
 *
 * The reference (Giovannibriglia/VectorizedBayesianNetwork, package `vbn` 0.3.0) is pure
 * Python/PyTorch and has no FFI; each entry point below names the reference code it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, a negative VBN_E_* code otherwise; the message
 *     is available from vbn_cuda_last_error() (thread local).  Nothing throws.
 *   - all `*_dev` / device-side pointers are CUDA device pointers owned by the CALLER
 *     (PyTorch tensors); the library allocates no device memory and keeps no global state.
 *   - every launch goes to the cudaStream_t passed as `stream` (a void* here so the header
 *     needs no CUDA include); no implicit synchronisation.
 *   - all floating point is IEEE fp32; categorical indices are int32.
 *   - row r of a run is (query b, sample s) with r = b*S + s, b in [0,B), s in [0,S).
 */
#ifndef VBN_CUDA_H_
#define VBN_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 6: descriptor layout of VBN_F_LGPLAIN / VBN_F_TABPLAIN ops (one word per parent slot), VBN_F_MDNFAST, all-or-nothing descriptor
   tails of tensor-core images -- a library and a plan compiler of different versions must not be mixed */
#define VBN_CUDA_ABI_VERSION 6

/* error codes */
#define VBN_OK 0
#define VBN_E_INVALID (-1)   /* bad argument / malformed program */
#define VBN_E_CUDA (-2)      /* CUDA runtime error, see last_error */
#define VBN_E_CAPACITY (-3)  /* program needs more on-chip slots than one CTA can hold */

/* ---- op kinds: one per CPD family on the path --------------------------------------- */
#define VBN_OP_NONE 0   /* value only (do-node, or parent supplied by the caller)          */
#define VBN_OP_LG 1     /* vbn/cpds/linear_gaussian.py:163-217 (also gaussian_nn roots)    */
#define VBN_OP_GNN 2    /* vbn/cpds/gaussian_nn.py:105-119,215-288                         */
#define VBN_OP_MDN 3    /* vbn/cpds/mdn.py:185-272                                         */
#define VBN_OP_SNN 4    /* vbn/cpds/softmax_nn.py:581-759                                  */
#define VBN_OP_KDE 5    /* vbn/cpds/kde.py:105-182                                         */
#define VBN_OP_TAB 6     /* log-density table per parent configuration: softmax_nn in discrete mode with
                            all-discrete parents (table built by the plan compiler), categorical_table
                            (vbn/cpds/categorical_table.py:359-417) and categorical_embedded_softmax
                            (vbn/cpds/categorical_embedded_softmax.py:316-511), both with strict supports */

#define VBN_OP_RFF 7     /* vbn/cpds/rff_gaussian.py:131-146,185-206,254-291: random Fourier features,
                            loc = (sqrt(2/F) cos(z W^T + b)) coef + bias, de-standardised; constant scale */

/* Gibbs sweeps (vbn/sampling/gibbs.py:38-82): candidates and their scores are ordinary CPD ops; these
 * glue them together.  One row = one chain. */
#define VBN_OP_TAKEW 8   /* slot[out_slot] = accumulated log-weight of the row; log-weight = 0          */
#define VBN_OP_SELECT 9  /* draw c ~ softmax(slots aux[0] .. aux[0]+k-1) (gibbs.py:76-77), copy candidate
                            c (slots aux[1] + c*dim ..) to out_slot; u from u_off or noise idx           */
#define VBN_OP_JUMP 10   /* back to op aux[0] until the loop has run k times; layer_dim[0], [1] = Philox
                            stream blocks (normal, uniform) one iteration consumes                      */

/* ---- op flags ------------------------------------------------------------------------ */
#define VBN_SRC_MASK 0x3
#define VBN_SRC_SAMPLE 0x0     /* draw x ~ p(x | parents)                                    */
#define VBN_SRC_FIXED_Q 0x1    /* x = fixed[fixed_col + d][b]   (evidence / do, per query)   */
#define VBN_SRC_FIXED_ROW 0x2  /* x = inputs[fixed_col] element (r, d)  (per row, CPD API)   */
#define VBN_SRC_SLOT 0x3       /* x = the value another op left in slots fixed_col .. (Gibbs: a child's
                                  current state, scored under a candidate parent value)       */
#define VBN_F_ADD_LOGW 0x4     /* logw[r] += log p(x | parents)  (evidence node)             */
#define VBN_F_OUT_LOGP 0x8     /* logp[r]  = log p(x | parents)  (MCM target, CPD.log_prob)  */
#define VBN_F_SHARED 0x10      /* draw is shared by all queries: noise keyed by s, not (b,s) */
#define VBN_F_FAST32 0x20      /* MLP has exactly 3 Linear layers with hidden dims [32, 32]    */
#define VBN_F_LGFAST 0x40      /* LG op with D = 1, Dp <= 4: layer_dim[] holds the float bits of
                                  {bias, scale, 2 ln scale, var, w0..w3}, aux[] =
                                  {parent slots 0|1<<16, 2|3<<16, out_slot, n_off}               */
#define VBN_F_PAR4 0x100       /* GNN / MDN op with <= 4 parent dims: aux[1] = slots 0|1<<16,
                                  aux[2] = slots 2|3<<16 (no par_slots[] lookup); aux[0] = float
                                  bits of min_scale                                            */
#define VBN_F_MDNPLAIN 0x200   /* MDN op, D = 1, 2 <= K <= 5, PAR4, tensor-core MLP, that is only
                                  drawn (Philox, per-row stream, no density): register-resident
                                  tail in the tcgen05 kernel; ignored by the FP32-pipe kernel   */
#define VBN_F_OUT_PARAMS 0x400  /* write the conditional-distribution parameters of every row to
                                  stores[store_idx] instead of drawing (CPDHandle.conditional)  */
#define VBN_F_MDNROOT 0x800    /* MDN op without parents (mdn.py:190-196 root parameters), D = 1,
                                  2 <= K <= 4, that is only drawn (Philox, per-row stream, no store,
                                  no density).  Its mixture is the same for every row, so the plan
                                  compiler evaluates it once: layer_dim[0..7], aux[0..3] hold the
                                  float bits of {cumulative weights c_0..c_{K-2} of the clamped,
                                  renormalised pi (mdn.py:227-228); then (loc_k, scale_k) for
                                  k = 0..K-1, scale_k = softplus(raw_k) + min_scale}; tc[] =
                                  {out_slot, n_off, u_off, K}.  The kernel reads quads 0,4,5,6,7  */
#define VBN_F_TABPLAIN 0x1000   /* TAB op with <= 4 parents and <= 4 classes, every parent's class values and its own
                                  coded 0..k-1 (the class index IS the value: no per-parent search, no value gather),
                                  either only drawn (VBN_SRC_SAMPLE: Philox, no store, no density) or only scored
                                  (VBN_SRC_FIXED_Q | VBN_F_ADD_LOGW: an evidence node).
                                  layer_dim[p] = stride_p | card_p << 16 (p < 4); layer_dim[4] = float offset in the
                                  parameter blob of the 128-bit-row table the op reads -- drawn: cdf4[n_cfg][4] =
                                  {c_0 .. c_{C-2}, +inf .., total}, scored: logp4[n_cfg][4] --; layer_dim[5] =
                                  C | strict << 16; layer_dim[6] = out_slot; layer_dim[7] = u_off (drawn) or the
                                  fixed[][B] row (scored); aux[0..3] = parent slots, one word each.
                                  The kernel reads quads 0,4,5,6                                            */
#define VBN_F_MDNFAST 0x2000   /* MDNPLAIN op with K = 3, ReLU and the first layer on the FP32 pipe (tc[2] == 0): the
                                  tcgen05 kernel reads quads 0,2,6,7 only; tc[1] = out_slot | tail count << 16         */
#define VBN_F_LGPLAIN 0x80     /* LGFAST op that is only drawn: Philox, per-row stream, no store,
                                  no density -- the kernel reads nothing but quads 0,4,5,6.  Its parent
                                  slots are one word each: aux[0], aux[1], layer_dim[2], layer_dim[3]
                                  (2 ln scale / var are only read by scored ops)                */

/* activations of the MLP CPDs (gaussian_nn.py:16-34) */
#define VBN_ACT_RELU 0
#define VBN_ACT_TANH 1
#define VBN_ACT_GELU 2
#define VBN_ACT_ELU 3

/* softmax_nn within-bin laws (softmax_nn.py:664-679) */
#define VBN_WB_UNIFORM 0
#define VBN_WB_TRIANGULAR 1
#define VBN_WB_GAUSSIAN 2

#define VBN_MAX_LAYERS 8

/*
 * One node of the schedule.  The host-side plan compiler
 * (vectorizedbayesiannetwork_b200/plan.py) emits these in topological order; it replaces
 * InferenceState / get_inference_state / prepare_fixed_values (vbn/inference/_core.py:13-135)
 * and the per-node Python loop bodies of likelihood_weighting.py:42-71,
 * importance_sampling.py:56-80, monte_carlo_marginalization.py:39-92, ancestral.py:27-41.
 * 128 bytes, 16-byte aligned so the kernel fetches it with eight 128-bit loads.
 */
typedef struct VbnOp {
  int32_t kind;       /* VBN_OP_*                                                          */
  int32_t flags;      /* VBN_SRC_* | VBN_F_*                                               */
  int32_t dim;        /* D: output dims of the node                                        */
  int32_t n_par;      /* Dp: total parent dims                                             */
  int32_t out_slot;   /* first of D consecutive value slots, -1 if the value is never read */
  int32_t par_off;    /* offset into par_slots[] of the Dp parent slot ids                 */
  int32_t param_off;  /* float offset of this node's block in the parameter blob           */
  int32_t fixed_col;  /* FIXED_Q: first row of fixed[][B]; FIXED_ROW: index into inputs[]  */
  int32_t store_idx;  /* index into stores[] where x is written, -1 none                   */
  int32_t noise_idx;  /* index into noise[] (injected draws), -1 = Philox                  */
  int32_t n_off;      /* first index in the per-row normal stream used by this op          */
  int32_t u_off;      /* first index in the per-row uniform stream used by this op         */
  int32_t n_layers;   /* MLP: number of Linear layers; 0 = outputs are constants (root)    */
  int32_t act;        /* VBN_ACT_*                                                         */
  int32_t n_out;      /* MLP output width O                                                */
  int32_t k;          /* MDN: components K; SNN: classes C; KDE: stored points N           */
  int32_t layer_dim[VBN_MAX_LAYERS]; /* widths after each Linear layer (last == n_out)      */
  int32_t aux[4];     /* SNN: {within_bin, clip, any_discrete, 0}                          */
  /* injected draws only, ops without a tensor-core image (tc[0] == 0): tc[1] > 1 = this op replays member tc[2] of a group of tc[1] draws the reference
     made in one call (Gibbs candidates); arrays inside a VBN_OP_JUMP loop carry a leading iteration axis */
  int32_t tc[4];      /* tensor-core MLP image (hidden dims [32,32], Dp <= 32, O <= 32), written by
                         the plan compiler: {1 if present, float offset of the image in the
                         parameter blob (16-byte aligned), K1 = Dp padded to 8 -- or 0: first layer on the
                         FP32 pipe from the plain W1^T[4][32], b1[32] block at the head of the image (GNN /
                         MDN ops with VBN_F_PAR4 only) --, N3 = O padded to 16};
                         image layout: see csrc/vbn_schedule_tc.cuh and cpds.py pack_mlp_tc        */
} VbnOp;

/* strided view of caller memory: element (r, d) lives at base[r*row_stride + d*dim_stride] */
typedef struct VbnView {
  float* base;
  int64_t row_stride;
  int64_t dim_stride;
} VbnView;

/* injected draws for one node (record/replay parity).  Layout of each
 * array is [Bn, S, D] row-major with Bn = 1 when the op has VBN_F_SHARED, else B.
 * eps: normals; u: uniforms; idx: categorical picks (MDN [Bn,S]; SNN [Bn,S,D]; KDE [Bn,S]). */
typedef struct VbnNoise {
  const float* eps;
  const float* u;
  const int32_t* idx;
} VbnNoise;

/* A compiled schedule.  All pointers are DEVICE pointers that must outlive the plan. */
typedef struct VbnProgramDesc {
  const VbnOp* ops_dev;
  int32_t n_ops;
  const int32_t* par_slots_dev;
  int32_t n_par_slots;
  const float* params_dev;
  int64_t n_params;
  int32_t n_slots;    /* value slots a row needs at once (after liveness analysis)         */
  int32_t n_scratch;  /* per-row scratch floats (max MLP output width over the program)    */
  int32_t heavy;      /* 1 if the program contains MLP / KDE ops (picks the launch shape)  */
  int32_t tc;         /* 0: FP32-pipe kernel; else the tcgen05 kernel (ops carry tensor-core MLP images) with
                         (tc & 0xFF) warpgroups per CTA, each owning (tc >> 8, 0 = 1) 128-row tiles:
                         4 | 1 << 8 or 2 | 2 << 8                                          */
  const int32_t* tc_list_dev; /* [n_tc][2] {image float offset, image bytes} of every op with
                         tc[0] != 0, in schedule order (the weight ring's fetch list)       */
  int32_t n_tc;
  int32_t tc_image_bytes;  /* largest {image bytes} of tc_list (one slot of the kernel's weight ring), 0 = the
                         format's maximum (30720)                                           */
  int32_t has_tables;      /* 1 if the program contains VBN_OP_TAB ops (FP32-pipe kernels without MLP ops come in a
                         variant without the table op bodies for linear-Gaussian-only schedules)           */
  int32_t rows_per_thread; /* FP32-pipe kernel, schedules without MLP/KDE ops: 0 = default (4 rows
                         per thread, best for drawn linear-Gaussian chains), 2 = table-lookup
                         heavy schedules (fewer registers, more resident warps)             */
} VbnProgramDesc;

typedef struct VbnPlan VbnPlan;

/* One execution of a plan over B*S rows. */
typedef struct VbnRunDesc {
  int64_t n_queries;       /* B (local)                                                    */
  int64_t n_samples;       /* S (local)                                                    */
  int64_t query_offset;    /* global index of local query 0   (multi-GPU sharding)         */
  int64_t sample_offset;   /* global index of local sample 0                               */
  uint64_t seed;           /* Philox key                                                   */
  uint64_t call_offset;    /* Philox counter word 3: caller bumps it per call              */
  const float* fixed_dev;  /* [n_fixed_cols][B] evidence / do values, may be NULL          */
  const VbnView* inputs_dev; /* per-row inputs referenced by FIXED_ROW ops                 */
  const VbnView* stores_dev; /* outputs referenced by store_idx                            */
  const VbnNoise* noise_dev; /* injected draws referenced by noise_idx, may be NULL        */
  float* logw_dev;         /* [B*S] log-weights, required if any op has VBN_F_ADD_LOGW      */
  float* logp_dev;         /* [B*S] log-density, required if any op has VBN_F_OUT_LOGP      */
  int32_t logp_as_pdf;     /* 1: write exp(logp) (MCM pdf, monte_carlo_marginalization.py:57,91) */
  int32_t logw_accumulate; /* 1: logw[r] += this run's log-weight (schedules run in segments,
                              resampled_importance_sampling.py:69-100); 0: overwrite          */
  int32_t* error_flag_dev; /* set to 1 when a softmax_nn discrete value has no class
                              (softmax_nn.py:620-625 raises ValueError); may be NULL       */
  /* Fused weight reduction (optional): the kernel folds every run of rows of one query inside a warp's 32
   * consecutive rows into one 16-float record {m = max logw, l = sum e, q = sum e^2, x0, sum e (x - x0),
   * sum e (x - x0)^2, rows, -, sum e [x == k] for k < 8}, e = exp(logw - m), at
   * seg_dev[(b * seg_per_query + (r / 32 - b S / 32)) * 16]; vbn_segment_merge folds a query's records.  This is the
   * first pass of torch.softmax(log_weights, 1) (importance_sampling.py:82-84, likelihood_weighting.py:75-80) and,
   * with seg_slot >= 0, the sums of VBN._posterior_stats (vbn/vbn.py:495-503) / the benchmark adapter's class
   * histogram (benchmarking/models/vbn.py:202-242) without a [B, S] tensor ever reaching HBM.                  */
  float* seg_dev;          /* [B][seg_per_query][16], NULL = off                              */
  int32_t seg_per_query;   /* >= ceil(S / 32) + 1                                             */
  int32_t seg_slot;        /* value slot whose weighted moments are accumulated (D = 1), -1 = log-weights only */
  int32_t seg_classes;     /* 0, or 1..8: also the weighted histogram of that value over classes 0..k-1 */
} VbnRunDesc;

int32_t vbn_cuda_abi_version(void);
const char* vbn_cuda_last_error(void);
int32_t vbn_cuda_device_count(int32_t* out_count);

/* Validates the program and records launch geometry for the current device. */
int32_t vbn_plan_create(const VbnProgramDesc* desc, VbnPlan** out_plan);
int32_t vbn_plan_destroy(VbnPlan* plan);

/* Replaces the whole per-node loop of LW / IS / MCM / ancestral and CPD.sample/log_prob:
 * one launch runs every op of the schedule for every row. */
int32_t vbn_run_forward(const VbnPlan* plan, const VbnRunDesc* run, void* stream);
/* number of kernel launches vbn_run_forward issues (for bench accounting) */
int32_t vbn_run_forward_launches(const VbnPlan* plan);

/*
 * Weight reduction: torch.softmax(log_weights, dim=1), ess = 1/sum(w^2)
 * (importance_sampling.py:82-84, likelihood_weighting.py:75-80).
 *   partials_dev : [B][n_split][3] workspace (m, l = sum exp(x-m), q = sum exp(2(x-m)))
 *   stats_dev    : [B][3] merged (m, l, q); ess = l*l/q
 * vbn_lse_partials + vbn_lse_merge compute stats; a multi-GPU caller all-gathers stats
 * across ranks and calls vbn_lse_merge again on the gathered [B][n_ranks][3] array.
 * vbn_weights_normalize then writes w = exp(x-m)/l (normalize=1) or
 * max(exp(x-m), eps) (normalize=0) and, if ess_dev != NULL, ess[b].
 */
int32_t vbn_lse_partials(const float* logw_dev, int64_t n_queries, int64_t n_samples,
                         int32_t n_split, float* partials_dev, void* stream);
int32_t vbn_lse_merge(const float* partials_dev, int64_t n_queries, int32_t n_split,
                      float* stats_dev, void* stream);
int32_t vbn_weights_normalize(const float* logw_dev, const float* stats_dev, int64_t n_queries,
                              int64_t n_samples, int32_t normalize, float eps, float* w_dev,
                              float* ess_dev, void* stream);
/*
 * Folds the records written through VbnRunDesc.seg_dev: merged_dev [B][16] = {m, l, q, weighted mean, 0,
 * M2 = sum e (x - mean)^2, rows, ess = l^2 / q, class sums[8]} (itself a valid record), stats_dev [B][3] = {m, l, q}
 * for vbn_weights_normalize (either may be NULL); with ess_threshold > 0 and flag_dev != NULL also the IS -> LW
 * fallback test any(ess < threshold) (importance_sampling.py:85-88).  n_samples > 0: records follow the kernel's
 * geometry (query b owns ((b+1) S - 1)/32 - b S/32 + 1 of its n_per_query slots); n_samples == 0: all n_per_query
 * records of every query are valid -- the cross-GPU merge of all-gathered merged records (samples sharded).
 */
int32_t vbn_segment_merge(const float* records_dev, int64_t n_queries, int64_t n_samples, int32_t n_per_query,
                          float ess_threshold, float* merged_dev, float* stats_dev, int32_t* flag_dev, void* stream);
/* any(ess < threshold) -> *flag_dev (int32), the IS->LW fallback test
 * (importance_sampling.py:85-88) */
int32_t vbn_ess_below(const float* stats_dev, int64_t n_queries, float threshold,
                      int32_t* flag_dev, void* stream);

/*
 * Resampling step of resampled_importance_sampling (vbn/inference/resampled_importance_sampling.py:33-41,
 * torch.multinomial(weights, S, replacement=True) + row gather), three launches:
 *   vbn_row_cdf          : cdf[b,s] = running sum of w[b,:]          (w need not be normalised)
 *   vbn_resample_indices : idx[b,s] = first position with cdf > u * cdf[b,S-1], u from the row's Philox stream
 *   vbn_gather_rows      : dst[c][b][s] = src[c][b][idx[b][s]] for n_cols live node columns ([n_cols][B][S] each)
 */
int32_t vbn_row_cdf(const float* w_dev, int64_t n_queries, int64_t n_samples, float* cdf_dev, void* stream);
int32_t vbn_resample_indices(const float* cdf_dev, int64_t n_queries, int64_t n_samples, uint64_t seed,
                             uint64_t call_offset, int64_t query_offset, int64_t sample_offset,
                             int32_t* idx_dev, void* stream);
int32_t vbn_gather_rows(const float* src_dev, float* dst_dev, const int32_t* idx_dev, int32_t n_cols,
                        int64_t n_queries, int64_t n_samples, void* stream);

/*
 * rao_blackwellized_marginalization epilogues (vbn/inference/rao_blackwellized_marginalization.py:277-317):
 *   vbn_weighted_sum          : out[b][k] = sum_s w[b,s] x[b,s,k]            (categorical marginal)
 *   vbn_gaussian_mixture_grid : mixture of N(mu_s, sigma_s) weighted by w on a +-stddevs grid around the
 *                               mixture mean; loc_scale_dev [B][S][2]; pdf_dev / grid_dev [B][n_out]
 */
int32_t vbn_weighted_sum(const float* w_dev, const float* x_dev, int64_t n_queries, int64_t n_samples,
                         int32_t width, float* out_dev, void* stream);
int32_t vbn_gaussian_mixture_grid(const float* w_dev, const float* loc_scale_dev, int64_t n_queries,
                                  int64_t n_particles, int64_t n_out, float stddevs, float min_scale,
                                  float* pdf_dev, float* grid_dev, void* stream);

/*
 * gaussian_exact support grid (vbn/inference/gaussian_exact.py:166-183): loc_scale_dev [B][2];
 * samples[b,s] = loc + scale * linspace(-stddevs, stddevs, S)[s], pdf[b,s] = N(samples; loc, scale).
 */
int32_t vbn_gaussian_grid(const float* loc_scale_dev, int64_t n_queries, int64_t n_samples, float stddevs,
                          float min_scale, float* pdf_dev, float* samples_dev, void* stream);

/*
 * Posterior summary, VBN._posterior_stats (vbn/vbn.py:483-504): weights = pdf sanitised
 * (nan/inf -> 0, clamp >= 0) and normalised over S (uniform 1/S when their sum <= eps);
 * weighted mean / std per output dim and ESS = 1 / sum w^2.
 *   pdf_dev [B,S], samples_dev [B,S,D] contiguous, D <= 8
 *   partials_dev : [B][n_split][10] workspace
 *   stats_dev    : [B][2+2D] = {sum of raw weights, ess, mean[D], std[D]}
 * Four launches (two-pass moments like the reference).
 */
int32_t vbn_posterior_stats(const float* pdf_dev, const float* samples_dev, int64_t n_queries,
                            int64_t n_samples, int32_t dim, int32_t n_split, float eps,
                            float* partials_dev, float* stats_dev, void* stream);

/*
 * Weighted class histogram of a discrete target -- replaces the benchmark adapter's Python double
 * loop _estimate_discrete_posterior(_batch) (benchmarking/models/vbn.py:202-242) and _normalize_probs
 * (:116-121): probs[b][c] = sum_s w[b,s] [round_half_even(x[b,s]) == c and w finite], divided by the
 * row total; uniform 1/K when the total is not finite or <= 0.  Sample (b, s) is read at
 * samples_dev[(b*S + s) * sample_stride] (stride D picks dim 0 of a [B,S,D] tensor, like the
 * reference's samples[:, :, 0]).  1 <= n_classes <= 256.  probs_dev: [B][n_classes].
 */
int32_t vbn_weighted_histogram(const float* samples_dev, const float* w_dev, int64_t n_queries, int64_t n_samples,
                               int64_t sample_stride, int32_t n_classes, float* probs_dev, void* stream);

/*
 * KDE conditional log-density over M query rows against N stored points
 * (vbn/cpds/kde.py:111-149): out[m] = LSE_n(log_kp + log_ky) - LSE_n(log_kp)
 * (root, dp == 0: LSE_n(log_ky) - ln N).  Stored points / queries are row-major
 * [N,dp],[N,dx],[M,dp],[M,dx].
 */
int32_t vbn_kde_log_prob(const float* train_p_dev, const float* train_y_dev, int64_t n_points,
                         int32_t dp, int32_t dx, const float* query_p_dev,
                         const float* query_x_dev, int64_t n_rows, float bandwidth,
                         float parent_bandwidth, float min_scale, float* out_dev, void* stream);

/*
 * The same density with the pairwise |x - y|^2 term on the tensor cores (tcgen05.mma kind::tf32, three-piece
 * operand split = fp32-exact products, streaming logsumexp in the epilogue): pays from dp + dx ~ 8, where the
 * FP32-pipe kernel is bound by the distance arithmetic instead of the two exponentials per pair.
 * workspace_dev: vbn_kde_tc_workspace_bytes() bytes (the stored points re-packed as augmented K-major tiles on every
 * call; 0 bytes = these dims are not covered, use vbn_kde_log_prob).  center_dev: dp + dx floats (parents then
 * targets) subtracted from both clouds, normally the stored points' mean, or NULL.  The GEMM form holds numbers of
 * size h |x - center|^2 in fp32 accumulators: rows beyond 40 (log2 units) -- and rows whose sums underflow -- are
 * recomputed with direct differences by a scalar pass, so accuracy never depends on the data, only speed does.
 */
int32_t vbn_kde_tc_workspace_bytes(int64_t n_points, int32_t dp, int32_t dx, int64_t* out_bytes);
int32_t vbn_kde_log_prob_tc(const float* train_p_dev, const float* train_y_dev, int64_t n_points, int32_t dp,
                            int32_t dx, const float* query_p_dev, const float* query_x_dev, int64_t n_rows,
                            float bandwidth, float parent_bandwidth, float min_scale, const float* center_dev,
                            void* workspace_dev, float* out_dev, void* stream);

/* Philox4x32-10 known-answer hook: out[i] = philox(ctr[i], key) for n counters (uint32 x4). */
int32_t vbn_philox_fill(const uint32_t* ctr_dev, int64_t n, uint32_t key0, uint32_t key1,
                        uint32_t* out_dev, void* stream);

/* Replay hook: the random draws vbn_run_forward makes for a run with the same (seed, call_offset, offsets), through
 * the same inlined generator as the hot kernels (csrc/vbn_schedule.cuh Ctx::normals / uniforms: Philox4x32-10 on
 * counter (global sample, global query | 0xFFFFFFFF when shared, block | tag << 30, call_offset), Box-Muller / u01).
 * kind 0: normals, 1: uniforms, 2: raw generator words (uint32 bit patterns).  Value i of an op's stream
 * (VbnOp.n_off / u_off + dim) is element i & 3 of block i >> 2.  out_dev: [4 * n_blocks][Bn][S] floats for blocks
 * block_lo .. block_lo + n_blocks - 1, Bn = n_queries (per-row streams) or 1 (shared != 0: the streams of
 * VBN_F_SHARED ops).  Tests regenerate a run's noise with it and replay it through the CPU oracle; the words are
 * checked against the oracle's numpy Philox (Random123 known answers). */
int32_t vbn_stream_draws(uint64_t seed, uint64_t call_offset, int32_t kind, int32_t shared, int32_t block_lo,
                         int32_t n_blocks, int64_t n_queries, int64_t n_samples, int64_t query_offset,
                         int64_t sample_offset, float* out_dev, void* stream);

/* Bench-only probe of the FP32 FMA pipe (the MLP layers' roofline denominator): n_blocks x 256
 * threads each run 16*iters FMAs.  mode 0 = scalar FFMA, 1 = packed fma.rn.f32x2.
 * flops = 2 * 16 * iters * 256 * n_blocks.  scratch_dev: >= 1 float. */
int32_t vbn_fma_peak(int32_t mode, int32_t iters, int32_t n_blocks, float* scratch_dev, void* stream);

/* Bench-only probe of the tensor pipe in the mode the MLP layers use (tcgen05.mma kind::tf32, A in
 * TMEM): n_blocks CTAs (one per SM) each issue `iters` M=128 N=256 K=8 MMAs.
 * flops = 2 * 128 * 256 * 8 * iters * n_blocks.  scratch_dev: >= 1 float. */
int32_t vbn_tf32_peak(int32_t iters, int32_t n_blocks, float* scratch_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBN_CUDA_H_ */
