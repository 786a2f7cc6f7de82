/*
 * vtc_b200.h -- C ABI of the B200-native contrastive-retrieval hot path of unitaryai/VTC.
 *
 * The reference has no FFI/plugin registry: its extension mechanism is name lookup on Python
 * modules (train.py:85-89, utils/parse_config.py:97-112).  The Python shims in
 * vtc_b200/{model,evaluation}/ keep those symbols and call the entry points below through
 * ctypes with `tensor.data_ptr()` and `torch.cuda.current_stream().cuda_stream`.
 * Each entry point names the reference call site it replaces (path:line in unitaryai/VTC).
 *
 * Conventions (all entry points):
 *   - extern "C", return int: 0 = VTC_OK, < 0 = error (see vtc_strerror); never throw.
 *   - never allocate device memory, never synchronise: all work is enqueued on `stream`;
 *     the caller owns every buffer including the workspace (size from vtc_workspace_bytes);
 *   - all pointers are DEVICE pointers unless said otherwise; matrices are row-major with an
 *     explicit leading dimension in ELEMENTS;
 *   - re-entrant: no mutable global state besides read-only lazily-resolved driver entry points.
 */
#ifndef VTC_B200_H_
#define VTC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define VTC_ABI_VERSION 3

typedef void* vtc_stream_t; /* a cudaStream_t */

/* storage type of an embedding matrix handed to the library */
enum { VTC_F32 = 0, VTC_BF16 = 1 };

/* retrieval score: DOT ranks by -q.x, L2 by ||x||^2 - 2 q.x (what faiss' IndexFlatL2 orders by;
 * the gallery is NOT assumed unit-norm, evaluation/retrieval_evaluation.py:254-259) */
enum { VTC_METRIC_DOT = 0, VTC_METRIC_L2 = 1 };

/*
 * arithmetic of the similarity pass.
 *   VTC_PREC_EXACT : the values ranked are the caller's fp32 inputs.  Tensor-core pass on a
 *                    3-term bf16 split (hi*hi + hi*lo + lo*hi), guard band, fp64 re-check of the
 *                    ambiguous pairs => ranks / top-k indices identical to fp64-sequential
 *                    arithmetic (oracle/vtc_oracle.c).  Loss/similarity within 1e-4 rel.
 *   VTC_PREC_BF16  : inputs are first rounded to bf16 (RN-even); one tensor-core pass; the same
 *                    guard band + fp64 re-check => ranks identical to fp64-sequential arithmetic
 *                    ON THE ROUNDED inputs.  Loss/similarity within 2e-2 rel of fp32.
 *   VTC_PREC_BRUTE : fp64 SIMT brute force over the fp32 inputs (no tensor cores); the anchor the
 *                    other two are tested against and their overflow fallback.
 */
enum { VTC_PREC_EXACT = 0, VTC_PREC_BF16 = 1, VTC_PREC_BRUTE = 2 };

/* ops for vtc_workspace_bytes */
enum {
  VTC_OP_SIM_RANK = 0,
  VTC_OP_SIM_TOPK = 1,
  VTC_OP_INFONCE_FWD = 2,
  VTC_OP_SIM_MATRIX = 3,
  VTC_OP_INFONCE_BWD = 4,
  VTC_OP_GT_SCORES = 5,
  VTC_OP_LINEAR = 6 /* N = rows, M = out_features, D = in_features */
};

/* CAM read-out modes (model/model.py:156-161, :356-362) */
enum { VTC_CAM_READOUT_AVG = 0, VTC_CAM_READOUT_RESIDUAL_ONLY = 1, VTC_CAM_READOUT_UNIFORM = 2 };

/* residual activations applied to the CAM residual (model/model.py:30-77, :168-171):
 * NORMALIZE_EPS = "normalize", SQUASH = "squash*" (res_scale = 1, 10, 1.2, 1.5, 1.8), TANH,
 * AFFINE = eval-mode "sub_mean" / "bn": (x - res_shift) * res_mul with the BatchNorm running stats */
enum {
  VTC_RESACT_NONE = 0,
  VTC_RESACT_NORMALIZE_EPS = 1,
  VTC_RESACT_SQUASH = 2,
  VTC_RESACT_TANH = 3,
  VTC_RESACT_AFFINE = 4
};

enum {
  VTC_OK = 0,
  VTC_ERR_INVALID_ARG = -1,
  VTC_ERR_UNSUPPORTED_SHAPE = -2, /* beyond the built limits: rows <= 2^29 per side, D <= 8192, k <= 16,
                                     L <= 16 tokens; or a dtype / precision combination an entry
                                     point documents as unsupported */
  VTC_ERR_WORKSPACE = -3,
  VTC_ERR_NO_DEVICE = -4,
  VTC_ERR_DRIVER = -5,
  VTC_ERR_CUDA_BASE = -1000 /* -1000 - cudaError_t */
};

int vtc_abi_version(void);
const char* vtc_strerror(int code);

/* bytes of caller-owned workspace an op needs; 0 on invalid arguments. */
size_t vtc_workspace_bytes(int op, int64_t N, int64_t M, int D, int precision);

/* ---- H1: normalize(x) = x / ||x||_2, no eps (model/model.py:26-27) ---------------------------
 * vtc_row_norms writes inv_norm[r] = 1/||x_r|| and sq_norm[r] = ||x_r||^2 (either may be NULL);
 * vtc_normalize writes Y = X / ||X|| (Y may alias X).  A zero row yields NaN like the reference. */
int vtc_row_norms(const void* X, int64_t rows, int D, int64_t ldx, int dtype, float* inv_norm,
                  float* sq_norm, vtc_stream_t stream);
int vtc_normalize(const void* X, int64_t rows, int D, int64_t ldx, int dtype, void* Y, int64_t ldy,
                  vtc_stream_t stream);

/* ---- H2: sim = (scale * A) @ B^T materialised (model/model.py:369,478,504,621) ---------------
 * Only for callers that really need the tensor; the fused ops below never form it.
 * A [N,D], B [M,D]; out fp32 [N,M] with leading dimension ldo; *scale is a device scalar. */
int vtc_sim_matrix(const void* A, const void* B, int64_t N, int64_t M, int D, int dtype,
                   int precision, const float* scale, float* out, int64_t ldo, void* ws,
                   size_t ws_bytes, vtc_stream_t stream);

/* ---- R1/R3: fused similarity + rank of ground truth ------------------------------------------
 * Replaces faiss.GpuIndexFlatL2.add/search + the host loop of RecallAtK.compute
 * (model/metric.py:137-161).  Q [N,D] queries, G [M,D] gallery (a chunk of the gallery when
 * col_offset/gt_score are used).  Ground truth of query t is gallery row gt[t] (global index) or
 * t + row_offset when gt == NULL.  rank0[t] (int32) receives, or with accumulate != 0 is
 * incremented by,
 *     #{ j in this G : j_glob != gt, d(t,j) < d(t,gt) } + #{ j_glob < gt : d(t,j) == d(t,gt) }
 * with j_glob = j + col_offset.  gt_score [N] (fp64, device) is d(t,gt); when NULL it is computed
 * from this G (gt must then lie inside it) and, if gt_score_out != NULL, stored there.
 * Finish with vtc_rank_finalize once all gallery chunks are accumulated. */
int vtc_sim_rank(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                 const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                 int precision, const double* gt_score, double* gt_score_out, int accumulate,
                 int32_t* rank0, void* ws, size_t ws_bytes, vtc_stream_t stream);

/* ---- R1 + R3 in one call: similarity + rank + R@K hit counts + median rank -----------------------
 * The whole of RecallAtK.compute (model/metric.py:137-161) for one (queries, gallery) pair that is
 * resident on the device: vtc_sim_rank over the full gallery followed by vtc_rank_finalize, as
 * memset + row prologue + tensor-core pass + a chain of three short launches (re-check, commit +
 * hit counts, median select).  hits [nk] int64 and medr (fp64, nullable) are device pointers; gt_score_out [N] (nullable) receives d(t,gt).
 * The workspace is that of VTC_OP_SIM_RANK. */
int vtc_rank_eval(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                  const int64_t* gt, int metric, int precision, const int* k_vals, int nk,
                  int32_t* rank0, int64_t* hits, double* medr, double* gt_score_out, void* ws,
                  size_t ws_bytes, vtc_stream_t stream);

/* ---- chunked ranking: per-row quantities computed once, reused by every call ---------------------
 * A chunked evaluation (host staging pipelined against the ranking, gallery shards gathered over
 * NVLink) calls the ranking many times on the same rows.  Three per-row quantities can be cached
 * across those calls, each either handed IN or computed by this call and stored OUT:
 *   sq64 [M] fp64  canonical ||x_j||^2 of THIS G            (sq64     in | sq64_out     out)
 *   qq   [N] fp32  upper bound of ||q_t||^2 (guard band)   (qq_up    in | qq_out       out)
 *   d(t,gt) [N] fp64                                        (gt_score in | gt_score_out out; when
 *           computed, rows whose ground truth lies outside this G receive NaN)
 * (exactly one of each pair must be non-NULL).  With all three handed in and bf16 rows of whole
 * 128-byte swizzle atoms (D % 64 == 0), no row is read outside the tensor-core pass and the re-check.
 * The cached values are tied to the canonical rows: the bf16 roundings with VTC_PREC_BF16, the fp32
 * values with VTC_PREC_EXACT.  Same results as vtc_sim_rank.  vtc_rank_prepare computes sq64 and / or
 * qq_up for a block of rows on its own (either may be NULL); its rows must already be the canonical
 * values (bf16 rows with VTC_PREC_BF16, fp32 rows otherwise: VTC_ERR_UNSUPPORTED_SHAPE if not). */
int vtc_rank_prepare(const void* X, int64_t rows, int D, int dtype, int precision, double* sq64,
                     float* qq_up, vtc_stream_t stream);
int vtc_sim_rank_prepared(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                          const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                          int precision, const double* gt_score, double* gt_score_out,
                          const double* sq64, double* sq64_out, const float* qq_up, float* qq_out,
                          int accumulate, int32_t* rank0, void* ws, size_t ws_bytes,
                          vtc_stream_t stream);

/* fp64-sequential d(t,gt) for each query (also the pre-pass of vtc_sim_rank). */
int vtc_gt_scores(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                  const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                  int precision, double* gt_score, void* ws, size_t ws_bytes,
                  vtc_stream_t stream);

/* rank0[t] = M_total where gt_score[t] is NaN ("never retrieved"); hits[i] = #{t: rank0[t] <
 * k_vals[i]} (int64, device; replaces model/metric.py:149-160); k_vals is a HOST array, nk <= 8.
 * medr (device double, may be NULL) = median(rank0) + 1 with numpy semantics.
 * hist_ws: >= (3*65536+8)*4 bytes of workspace when medr != NULL (three-level radix select). */
int vtc_rank_finalize(int32_t* rank0, const double* gt_score, int64_t N, int64_t M_total,
                      const int* k_vals, int nk, int64_t* hits, double* medr, void* hist_ws,
                      size_t hist_ws_bytes, vtc_stream_t stream);

/* ---- K7: fused similarity + streaming top-k ----------------------------------------------------
 * The `search(b, max(k)+1)` of model/metric.py:144-146 without the N x M matrix.  k <= 16.
 * out_val fp32 [N,k] ascending score d (L2: ||x||^2 - 2q.x + ||q||^2 like faiss; DOT: -q.x),
 * out_idx int64 [N,k] = j + col_offset, ties by lower index, -1/+inf fill when M < k. */
int vtc_sim_topk(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype, int metric,
                 int precision, int k, int64_t col_offset, float* out_val, int64_t* out_idx,
                 void* ws, size_t ws_bytes, vtc_stream_t stream);

/* merge `parts` sorted candidate lists vals/idx [parts,N,k] into the k best per row. */
int vtc_topk_merge(const float* vals, const int64_t* idx, int parts, int64_t N, int k,
                   float* out_val, int64_t* out_idx, vtc_stream_t stream);

/* ---- H2+H3: fused similarity + symmetric InfoNCE (model/loss.py:18-22) -------------------------
 * A [n,D], B [n,D] (already normalised); logits = (*scale) * A B^T never reach HBM.
 * loss (device float) = 0.5 * (mean_i[row_lse_i - diag_i] + mean_j[col_lse_j - diag_j]);
 * row_lse/col_lse/diag [n] fp32 are saved for the backward. */
int vtc_infonce_fwd(const void* A, const void* B, int64_t n, int D, int dtype, int precision,
                    const float* scale, float* loss, float* row_lse, float* col_lse, float* diag,
                    void* ws, size_t ws_bytes, vtc_stream_t stream);

/* backward of vtc_infonce_fwd (loss.backward() of model/loss.py:18-22, trainer/trainer.py:79):
 * dA, dB fp32 [n,D], dscale (device float); grad_loss device float; `precision` as in the forward
 * (the logits are recomputed from the same operands the saved log-sum-exps were reduced from).
 * n <= 2048: fp32 SIMT tiles; larger: tcgen05 -- logit tiles recomputed, gradient weights handed on
 * as bf16 operand strips of <= 64 MB, no n x n array (fp32 features only). */
int vtc_infonce_bwd(const void* A, const void* B, int64_t n, int D, int dtype, int precision,
                    const float* scale, const float* row_lse, const float* col_lse,
                    const float* grad_loss, float* dA, float* dB, float* dscale, void* ws,
                    size_t ws_bytes, vtc_stream_t stream);

/* clip_loss on a MATERIALISED sim [n, n] fp32 (what the reference's own forward returns,
 * model/model.py:369 -> model/loss.py:19): row / column log-sum-exp + diagonal in one pass over the
 * matrix, then the scalar; the backward is elementwise (d loss / d sim). */
int vtc_infonce_dense_fwd(const float* sim, int64_t n, int64_t ld, float* loss, float* row_lse,
                          float* col_lse, float* diag, vtc_stream_t stream);
int vtc_infonce_dense_bwd(const float* sim, int64_t n, int64_t ld, const float* row_lse,
                          const float* col_lse, const float* grad_loss, float* dsim, int64_t ldd,
                          vtc_stream_t stream);

/* ---- H4: Context Adapter Module pieces (model/model.py:141-205) --------------------------------
 * vtc_cam_stack_normalize: X[l] = normalize(l == 0 ? main : aux[l-1]) -> X [L,b,D]   (:150-151)
 * vtc_layernorm          : fp32 LayerNorm over the last dim, eps 1e-5 (clip.model.LayerNorm)
 * vtc_cam_attn_core      : per (token batch, head) softmax(q k^T / sqrt(hd)) v over L <= 16
 *                          tokens; QKV [L,b,3D] (in_proj output), out [L,b,D]       (:155, MHA)
 * vtc_bias_act           : Y = act(X + bias) (+ residual); act 0 = none, 1 = QuickGELU
 * vtc_cam_readout        : AVG:  res = normalize(mean_l normalize(T_l))            (:156-159)
 *                          RESIDUAL_ONLY: res = `res_in` [b,D] (final_linear done by caller, :161)
 *                          UNIFORM: out = normalize(mean_l T_l)  (averaging fusion, :356-366)
 *                          then (AVG / RESIDUAL_ONLY) zero rows where skip_mask != 0 (:199-201)
 *                          and out = normalize(normalize(main) + res)               (:203) */
int vtc_cam_stack_normalize(const float* main, const float* aux, int L, int64_t b, int D,
                            float* X, vtc_stream_t stream);
int vtc_layernorm(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                  float eps, float* Y, vtc_stream_t stream);
int vtc_cam_attn_core(const float* QKV, int L, int64_t b, int D, int heads, float* out,
                      vtc_stream_t stream);
int vtc_bias_act(const float* X, const float* bias, const float* residual, int64_t rows, int D,
                 int act, float* Y, vtc_stream_t stream);
int vtc_cam_readout(const float* T, const float* main, const float* res_in,
                    const uint8_t* skip_mask, int L, int64_t b, int D, int mode, int res_act,
                    float res_scale, const float* res_shift, const float* res_mul, float* out,
                    vtc_stream_t stream);

/* ---- dense linear on the tensor cores (CAM projections / MLP) ----------------------------------
 * Y[rows,out_f] = act(X[rows,in_f] @ W[out_f,in_f]^T + bias) + residual, fused in the GEMM
 * epilogue (fp32 in/out; bias [out_f] and residual [rows,out_f] nullable; act 0 none / 1
 * QuickGELU; precision EXACT or BF16). */
int vtc_linear(const float* X, const float* W, const float* bias, const float* residual,
               int64_t rows, int in_f, int out_f, int act, int precision, float* Y, void* ws,
               size_t ws_bytes, vtc_stream_t stream);

/* ---- H4 in one call: prepared linears + the whole CAM forward -----------------------------------
 * vtc_linear_prepare writes a weight once as the gallery-side bf16 operand (+ padded bias) into a
 * caller-owned buffer of vtc_linear_prepared_bytes(); vtc_cam_forward then runs
 *   stack+normalize -> layers x [LN1+prep, QKV GEMM, attention core, out_proj GEMM(+residual),
 *                                LN2+prep, c_fc GEMM(+QuickGELU, bf16 operand out), c_proj GEMM(+res)]
 *   -> read-out (AVG, or final_linear(token 0) for RESIDUAL_ONLY) -> normalize(normalize(main)+res)
 * i.e. PretrainedCLIPBase._adapt_feature (model/model.py:141-205) in 2 + 7*layers launches.
 * `layers_params` is a HOST array of `layers` structs holding DEVICE pointers. */
typedef struct {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;  /* LayerNorm weight / bias [D] */
  const void *qkv, *out, *fc, *proj;           /* prepared linears: D->3D, D->D, D->4D, 4D->D */
} vtc_cam_layer;
size_t vtc_linear_prepared_bytes(int in_f, int out_f, int precision);
int vtc_linear_prepare(const float* W, const float* bias, int in_f, int out_f, int precision,
                       void* prepared, vtc_stream_t stream);
size_t vtc_cam_workspace_bytes(int L, int64_t b, int D, int precision);
int vtc_cam_forward(const float* main, const float* aux, int L, int64_t b, int D, int heads,
                    int layers, const vtc_cam_layer* layers_params, int readout_mode,
                    const void* final_linear_prepared, const uint8_t* skip_mask, int res_act,
                    float res_scale, const float* res_shift, const float* res_mul, int precision,
                    float* out, void* ws, size_t ws_bytes, vtc_stream_t stream);

/* ---- backward of the CAM (training: trainer/trainer.py:79 back-propagates through
 * _adapt_feature).  Dense products of the backward are vtc_linear on transposed copies
 * (dX = dY W = linear(dY, W^T), dW = dY^T X = linear(dY^T, X^T)); the rest is here.
 * vtc_layernorm_bwd: dX = LN'(dY) + dres (dres nullable), dgamma / dbeta [D] overwritten.
 * vtc_cam_readout_bwd: same residual-activation arguments as vtc_cam_readout (the Jacobians of
 * model/model.py:30-77; AFFINE = sub_mean / bn on running statistics, whose own parameters get no
 * gradient here); AVG writes dT [L,b,D], RESIDUAL_ONLY writes dres [b,D]; dmain [b,D] is the part
 * through normalize(main) at model.py:203; UNIFORM (averaging fusion) writes dT only.
 * vtc_cam_stack_normalize_bwd: dX [L,b,D] -> dmain [b,D] (token 0), daux [L-1,b,D]. */
int vtc_normalize_bwd(const float* X, const float* dY, int64_t rows, int D, float* dX,
                      vtc_stream_t stream); /* dX of Y = X / |X| (model/model.py:26-27) */
int vtc_transpose(const float* in, int64_t rows, int64_t cols, float* out, vtc_stream_t stream);
int vtc_gelu_bwd(const float* dF, const float* U, int64_t n, float* dU, vtc_stream_t stream);
int vtc_colsum(const float* X, int64_t rows, int64_t cols, float* out, vtc_stream_t stream);
int vtc_layernorm_bwd(const float* dY, const float* X, const float* gamma, int64_t rows, int D,
                      float eps, const float* dres, float* dX, float* dgamma, float* dbeta,
                      vtc_stream_t stream);
int vtc_cam_attn_core_bwd(const float* QKV, const float* dO, int L, int64_t b, int D, int heads,
                          float* dQKV, vtc_stream_t stream);
int vtc_cam_stack_normalize_bwd(const float* main, const float* aux, const float* dX, int L,
                                int64_t b, int D, float* dmain, float* daux, vtc_stream_t stream);
int vtc_cam_readout_bwd(const float* T, const float* main, const float* res_in,
                        const uint8_t* skip_mask, const float* dout, int L, int64_t b, int D,
                        int mode, int res_act, float res_scale, const float* res_shift,
                        const float* res_mul, float* dT, float* dres, float* dmain,
                        vtc_stream_t stream);

/* The whole backward of _adapt_feature in ONE call (the chain round 1 drove from Python with ~156
 * launches): read-out backward -> layers (reversed) x [c_proj, QuickGELU, c_fc, LN2, out_proj,
 * attention core, in_proj, LN1] -> stacked-input normalisation.  Per layer the caller hands in the
 * activations its forward kept (fp32 [L*b, .]: X the layer input, H1 = LN1(X), QKV, A the attention
 * output, X2, H2 = LN2(X2), U = c_fc(H2), Fa = QuickGELU(U)), the LayerNorm weights, the prepared
 * linears of the TRANSPOSED weights (vtc_linear_prepare on W^T, no bias) and the buffers that receive
 * the parameter gradients.  T [L,b,D] is the transformer output, res_in the final_linear output
 * (RESIDUAL_ONLY).  dmain [b,D], daux [L-1,b,D], dflw [D,D] (nullable).  HOST array of structs
 * holding DEVICE pointers. */
typedef struct {
  const float *X, *H1, *QKV, *A, *X2, *H2, *U, *Fa;
  const float *ln1_g, *ln2_g;
  const void *qkv_t, *out_t, *fc_t, *proj_t;
  float *dWqkv, *dbqkv, *dWo, *dbo, *dg1, *db1, *dWfc, *dbfc, *dWpr, *dbpr, *dg2, *db2;
} vtc_cam_layer_bwd;
size_t vtc_cam_backward_workspace_bytes(int L, int64_t b, int D, int precision);
int vtc_cam_backward(const float* dout, const float* main, const float* aux, const float* T,
                     const float* res_in, const uint8_t* skip_mask, int L, int64_t b, int D,
                     int heads, int layers, const vtc_cam_layer_bwd* layers_params,
                     int readout_mode, const void* final_linear_t, int res_act, float res_scale,
                     const float* res_shift, const float* res_mul, int precision, float* dmain,
                     float* daux, float* dflw, void* ws, size_t ws_bytes, vtc_stream_t stream);

/* number of kernels this library has launched since load (for bench.py's gpu_launches). */
uint64_t vtc_launch_count(void);

/* Opt-in profiling aid: while enabled, every tensor-core similarity launch is bracketed by CUDA
 * events on its own stream.  vtc_kernel_timer_read synchronises those events, returns their summed
 * duration (HOST pointers: total_ms, count) and clears the list.  Off by default. */
int vtc_kernel_timer_enable(int on);
int vtc_kernel_timer_read(double* total_ms, int* count);

/* Profiling only (VTC_DBG_PROF=1 in the environment when the library is loaded): the tensor-core
 * kernel's MMA issuer and first epilogue warp record, per CTA, 8 x uint64 {clock64 ticks,
 * globaltimer ns, ticks the issuer waited for a free accumulator, ticks it waited for operand
 * stages, tiles, ticks the epilogue waited for a full accumulator, epilogue ticks, 0}.
 * Synchronises the device, copies the words of the LAST launch to the HOST buffer and clears them.
 * VTC_ERR_INVALID_ARG when profiling is off. */
int vtc_debug_prof_read(unsigned long long* out, int max_words);

/* Profiling only: launch trace.  Between vtc_trace_begin(stream) and vtc_trace_end the library
 * records one CUDA event on `stream` after every kernel it launches; vtc_trace_end synchronises,
 * writes one line "<source file>:<line> <microseconds>" per launch into the HOST buffer (the gap
 * to the previous event: the launch as it ran back to back in the stream, warm caches -- unlike
 * ncu's serialised cold-cache times) and returns the number of launches (< 0: error).  Single
 * stream, not thread-safe against concurrent library calls; off by default. */
int vtc_trace_begin(vtc_stream_t stream);
int vtc_trace_end(char* buf, size_t cap);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VTC_B200_H_ */
