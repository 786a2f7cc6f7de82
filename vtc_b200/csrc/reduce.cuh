// reduce.cuh -- launchers of the post-GEMM reductions (reduce.cu): InfoNCE loss from the fused
// log-sum-exp partials, exact top-k selection from the streamed candidate pools, top-k merge.
#pragma once
#include "common.cuh"
#include "exact.cuh"

namespace vtc {

// row_lse[t] = ln2 * (max + log2(sum)) merged over `splits` partials [splits, n]
// (run_flag != NULL: the launch does nothing unless *run_flag != 0)
int launch_lse_merge(const float2* part, int splits, int64_t n, float* lse, cudaStream_t s,
                     const unsigned int* run_flag = nullptr);
// one-pass column log-sum-exp of the symmetric loss (LseEpi in sim_tc_kernel.cuh): per-column
// reference = the positive logit, and the merge of the per-(query tile, lane quarter) partial sums
int launch_nce_colref(const __nv_bfloat16* opA, const __nv_bfloat16* opB, int64_t n, int64_t npad,
                      int Kp, const float* scale_ptr, float* ref, cudaStream_t s);
int launch_nce_col_merge(const float* col_part, int parts, int64_t ld, const float* ref, int64_t n,
                         float* col_lse, unsigned int* flag, cudaStream_t s);
// loss = 0.5 * (mean(row_lse - diag) + mean(col_lse - diag)); diag = scale * diag_raw (written back)
int launch_infonce_loss(const float* row_lse, const float* col_lse, const float* diag_raw,
                        const float* scale_ptr, int64_t n, float* diag_out, float* loss,
                        cudaStream_t s);

struct TopkSelectArgs {
  ExactArgs ex;              // canonical operands (Q, G, sq64, metric, col_offset ...)
  const float2* pool_buf;    // [splits, N, pool] (approx score, column index as int bits)
  const float2* pool_meta;   // [splits, N] (entries, tau)
  int splits, pool, k;
  float guard_rel;
  const unsigned int* max_sq_bits;
  float* out_val;            // [N, k]
  int64_t* out_idx;          // [N, k]
  unsigned int* row_flag;    // [N] 1 = pool could not prove completeness -> brute-force row
};
int launch_topk_select(const TopkSelectArgs& a, cudaStream_t s);
// tau0[t] = (k-th smallest approximate score in the pools) + 3 delta_t, +inf if fewer than k entries
int launch_topk_tau(const TopkSelectArgs& a, const float* scores, int cols, int64_t ld, float* tau0,
                    cudaStream_t s);
int launch_topk_brute_rows(const TopkSelectArgs& a, cudaStream_t s);
int launch_topk_merge(const float* vals, const int64_t* idx, int parts, int64_t N, int k,
                      float* out_val, int64_t* out_idx, cudaStream_t s);

}  // namespace vtc
