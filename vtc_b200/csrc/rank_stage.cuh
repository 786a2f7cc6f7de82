// rank_stage.cuh -- the two kernels around the tensor-core pass of vtc_sim_rank / vtc_rank_eval
// (rank_stage.cu): ONE prologue launch (bf16 operands, canonical norms, ground-truth scores, epilogue
// bias, guard-band inputs) and the epilogue chain (fp64 re-check of the guard-band groups -- or, decided
// up front by every block, the canonical recount of the whole call --, commit + R@K hit counts,
// radix-select median), short launches chained by programmatic dependent launch.  A retrieval
// evaluation is memset + prologue + tensor-core pass + 3 epilogue launches instead of the 17 launches
// of round 1.
#pragma once
#include "common.cuh"
#include "exact.cuh"

namespace vtc {

constexpr int STAGE_NONE = -1;  // mode_q / mode_g: the input rows ARE the tensor-core operands

struct RankPrologueArgs {
  const void* Q;   // [N, ldq] fp32 or bf16
  const void* G;   // [M, ldg]
  int in_bf16;
  int64_t N, M;
  int D;
  int64_t ldq, ldg;
  int mode_q, mode_g;  // STAGE_NONE or PREP_PLAIN / PREP_SPLIT_A / PREP_SPLIT_B (prep.cuh)
  int round_bf16;      // canonical values = bf16 roundings of the fp32 inputs (VTC_PREC_BF16)
  __nv_bfloat16* opQ;  // [N, Kp]
  __nv_bfloat16* opG;  // [M, Kp]
  int Kp;
  int metric;
  // gallery side: canonical ||x_j||^2 (computed here unless sq64_in is given), fp32 epilogue bias
  // padded with +inf to Mpad, largest finite norm (atomicMax on the float bits; zeroed by the caller)
  const double* sq64_in;
  double* sq64;
  float* bias;                // nullable (with max_sq_bits) when only the norms are wanted
  int64_t Mpad;               // gallery rows incl. bias padding (>= M)
  unsigned int* max_sq_bits;
  // query side: d(t, gt) (computed here unless gt_in is given) and an upper bound of ||q_t||^2
  // (computed here unless qq_in is given)
  const double* gt_in;
  double* dgt;                // nullable: no ground-truth scores wanted
  const float* qq_in;
  float* qq;                  // nullable: no norm bounds wanted
  const int64_t* gt;
  int64_t row_offset, col_offset;
  // split modes only (nullable): the norms of the split's pieces for the per-row guard band
  // (sim_tc.cuh::rank_split_bound) -- per query upper bounds of (|lo(q_t)|, |e(q_t)|), over the
  // gallery the float bits of max_j |lo(x_j)|^2 and max_j |e(x_j)|^2 (atomicMax; zeroed by the caller)
  float2* qsplit;
  unsigned int* split_max_bits;
  // set to 1 when a 3-term split operand would have to carry inf / NaN (x - bf16(x) is NaN there):
  // the epilogue then recomputes the call with the canonical brute-force kernel
  unsigned int* fallback;
};
int launch_rank_prologue(const RankPrologueArgs& a, cudaStream_t s);

struct RankEpilogueArgs {
  ExactArgs ex;  // canonical operands
  const int2* amb_list;
  const unsigned int* seg_count;
  int nseg;
  unsigned int seg_cap;
  const double* dgt;
  int* rank_tmp;           // [N] counts of this call (tensor-core pass + re-check)
  int* rank_alt;           // [N] zeroed by the caller: counts of the canonical recount, taken instead
                           // of rank_tmp when the call falls back (rank_needs_fallback)
  const unsigned int* fallback;  // set by the prologue: split operands that cannot carry inf / NaN
  int32_t* rank0;          // [N] result
  int accumulate;
  // optional finalisation (vtc_rank_eval): NaN ground truth -> rank M_total, hits, median
  int finalize;
  int64_t M_total;
  int k_vals[8];
  int nk;
  unsigned long long* hits;  // [nk]
  double* medr;              // nullable
  unsigned int* hist;        // [3 * 2 * 2048 + 8] scratch (zeroed by the first launch)
};
int launch_rank_epilogue(const RankEpilogueArgs& a, cudaStream_t s);
size_t rank_epilogue_hist_words();

}  // namespace vtc
