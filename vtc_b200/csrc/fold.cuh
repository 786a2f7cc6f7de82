// fold.cuh -- operands of the EPI_RANK_FOLD tensor-core pass (fold.cu; see RankFoldEpi in
// sim_tc_kernel.cuh): the per-column bias and the per-row ground-truth score as one extra K16 step
// of the MMA.
#pragma once
#include "common.cuh"

namespace vtc {

// bf16 per fold-operand row in global memory.  64 = one whole 128-byte swizzle atom, laid out like
// any other k-block.  16 = only the K16 slice that is multiplied: the TMA box stays {64, rows} and
// the 48 out-of-range columns are zero-filled (as for a K tail), so a tile's fold block costs 4 KB of
// L2 -> shared-memory traffic instead of 16 KB.  VTC_FOLD_COLS selects (default 64).
constexpr int FOLD_COLS_MAX = 64;

// Qx [N, cols] bf16 (cols = 16 or 64) = [ m'_t (three bf16 pieces) | 1 1 1 | 0 ... ] and
// fold_w [N]: half-width of the guard band around acc' = 0, derived from the (lo, hi) thresholds and d(t,gt) that
// launch_gt_score produced.  Rows whose ground-truth score is NaN get m' = -1e30, w = -1 (they count
// nothing and never push; vtc_rank_finalize gives them rank M); an infinite score sets *invalid = 1
// (brute-force fallback, as for launch_fold_g).
int launch_fold_q(const float2* thr, const double* dgt, const unsigned int* max_sq_bits, int64_t N,
                  int metric, float guard_rel, __nv_bfloat16* Qx, int cols, float* fold_w,
                  unsigned int* invalid, cudaStream_t s);

// Gx [Mpad, cols] bf16 = [ 1 1 1 | h_j (three bf16 pieces) | 0 ... ], h_j = -||x_j||^2 / 2 (L2) or 0
// (DOT); padding rows j >= M carry h = -1e30 (never closer, never inside the band).  A gallery row
// whose squared norm is not finite sets *invalid = 1: the caller's brute-force fallback then
// recomputes the whole call in canonical arithmetic.
int launch_fold_g(const double* sq64, int64_t M, int64_t Mpad, int metric, __nv_bfloat16* Gx,
                  int cols, unsigned int* invalid, cudaStream_t s);

// ---- prepared ranking (vtc_rank_prepare / vtc_sim_rank_prepared): per-row quantities computed once
// per gallery / query chunk and reused by every library call that touches the chunk.
// qq_up[r] = fp32 upper bound of ||x_r||^2 (coalesced warp-per-row sum, inflated by 1e-4).
int launch_qnorm_up(const void* X, bool bf16, int64_t ldx, int64_t rows, int D, float* qq_up,
                    cudaStream_t s);
// bias[j] = (float)sq64[j] (L2) or 0 (DOT) for j < M, +inf for M <= j < Mpad, and
// *max_sq_bits = max over the finite (float)sq64[j]  (what sqnorm64 + fill_bias produce together)
int launch_bias_max(const double* sq64, int64_t M, int64_t Mpad, int metric, float* bias,
                    unsigned int* max_sq_bits, cudaStream_t s);
// thresholds from cached quantities: d(t,gt) and the upper bound of ||q_t||^2 (no row walks)
int launch_thr_cached(const float* qq_up, const double* dgt, const unsigned int* max_sq_bits,
                      int64_t N, int metric, float guard_rel, float2* thr, cudaStream_t s);

// Opt-in (VTC_FAST_THR=1): the (lo, hi) guard-band thresholds of launch_gt_score from a coalesced
// fp32 row norm instead of a second fp64-sequential walk over every query row.  The band only needs
// an UPPER bound of ||q_t|| (the scores themselves stay canonical), so ||q||^2 is summed in fp32 by a
// warp and inflated by 1e-4; everything else is gt_score_kernel's formula.  Q / ldq / bf16 / D are
// the canonical query rows (ExactArgs), dgt the canonical d(t,gt).
int launch_thr_fast(const void* Q, bool bf16, int64_t ldq, int64_t N, int D, const double* dgt,
                    const unsigned int* max_sq_bits, int metric, float guard_rel, float2* thr,
                    cudaStream_t s);

}  // namespace vtc
