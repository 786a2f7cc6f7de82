// exact.cuh -- launchers of the fp64-sequential kernels (exact.cu).
#pragma once
#include "common.cuh"

namespace vtc {

// Operands of the exact kernels: "canonical" values, i.e. the caller's fp32 inputs
// (VTC_PREC_EXACT / BRUTE) or their bf16 roundings (VTC_PREC_BF16).
struct ExactArgs {
  const void* Q;
  const void* G;
  int64_t ldq, ldg;
  bool bf16;  // element type of Q and G
  int64_t N, M;
  int D;
  const double* sq64;  // [M] canonical ||x_j||^2 (L2 only)
  const int64_t* gt;   // device, nullable
  int64_t row_offset, col_offset;
  int metric;
};

int launch_sqnorm64(const void* X, bool bf16, int64_t rows, int D, int64_t ld, double* sq64,
                    float* sq32, unsigned int* max_sq_bits, cudaStream_t s);
int launch_gt_score(const ExactArgs& a, const double* gt_in, double* gt_out, float2* thr,
                    const unsigned int* max_sq_bits, float guard_rel, cudaStream_t s);
int launch_rank_brute(const ExactArgs& a, const double* dgt, int* rank,
                      const unsigned int* run_flag, cudaStream_t s);
int launch_fill_bias(float* dst, const float* src, int64_t M, int64_t Mpad, float pad,
                     cudaStream_t s);
int launch_zero_if_flag(int* buf, int64_t n, const unsigned int* flag, cudaStream_t s);
int launch_rank_commit(const int* tmp, int* rank, int64_t n, int accumulate, cudaStream_t s);
int launch_rank_finalize(int* rank, const double* dgt, int64_t N, int64_t M_total,
                         const int* k_vals, int nk, int64_t* hits, double* medr, void* hist_ws,
                         cudaStream_t s);

}  // namespace vtc
