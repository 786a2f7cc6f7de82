// sim_tc.cuh -- interface of the tcgen05 similarity GEMM with fused epilogues (sim_tc.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vtc {
namespace tc {

constexpr int BM = 128;  // query rows per CTA tile  (UMMA M, TMEM lanes)
constexpr int BN = 256;  // gallery rows per tile    (UMMA N, TMEM columns per accumulator stage)
constexpr int BK = 64;   // bf16 per k-block = one 128-byte swizzle atom

enum Epilogue { EPI_RANK = 0, EPI_LSE = 1, EPI_STORE = 2, EPI_TOPK = 3 };

constexpr int TOPK_POOL = 64;  // buffer entries per (row, gallery split)
constexpr int TOPK_KEEP_MAX = 32;  // entries kept when a buffer is compacted: 16 for k <= 12, else 32

struct Params {
  int64_t N, M;          // valid rows of A (queries) and B (gallery)
  int num_kb;            // K' / 64
  int q_tiles, g_tiles;  // ceil(N/128), ceil(M/256)
  int g_splits;          // gallery is cut into g_splits contiguous tile ranges
  int tiles_per_split;
  // work items: q_groups * g_splits, of which those from tail_first on (the last, partial round of
  // the round-robin deal) are cut into tail_parts gallery sub-ranges each (1 = not cut)
  int num_items, tail_first, tail_parts;
  // score = scale * acc + col_bias[j].  col_bias is REQUIRED and padded to a multiple of BN
  // entries; the padding holds the epilogue's neutral value (+inf rank/top-k, -inf LSE, 0 store).
  const float* col_bias;
  const float* scale_ptr;  // optional device scalar multiplied into `scale`
  float scale;
  // EPI_RANK: the guard band (lo, hi) around d(t, gt) is derived per work item from the canonical
  // ground-truth score, an upper bound of ||q_t||^2 and the largest gallery norm (rank_band())
  const double* dgt;                 // [N] canonical d(t, gt); NaN = no ground truth in this call
  const float* qq;                   // [N] upper bound of ||q_t||^2
  const unsigned int* max_sq_bits;   // float bits of max_j ||x_j||^2 over the finite gallery rows
  float guard_rel;                   // bound of |tensor-core dot - exact dot| / (|q||x|)
  // VTC_PREC_EXACT: what the 3-term bf16 split drops, bounded per ROW from the norms of the split's
  // own pieces (rank_split_bound()) instead of the worst-case constant folded into guard_rel
  const float2* qsplit;                 // [N] upper bounds of (|lo(q_t)|, |e(q_t)|); nullable
  const unsigned int* split_max_bits;   // float bits of max_j |lo(x_j)|^2, max_j |e(x_j)|^2
  int metric_l2;                     // 1: score = ||x||^2 - 2 q.x, 0: score = -q.x
  int* rank;          // [N] += #{j : score < lo}
  int2* amb_list;     // amb_pack(t, j0, mask): row t has a score in [lo, hi] at the columns j0 + i of
                      // [j0, j0 + 8) whose mask bit i is set; one segment of amb_seg_cap entries per CTA
  unsigned int* amb_seg_count;  // [grid] entries each CTA wanted to push (may exceed the capacity)
  unsigned int amb_seg_cap;
  // ground-truth column of row t in THIS launch's column numbering:
  // (gt ? gt[t] : t + gt_row_offset) - gt_col_offset.  Its own score is inside the guard band by
  // construction; a group whose only in-band column is that one needs no re-check.
  const int64_t* gt;
  int64_t gt_row_offset, gt_col_offset;
  // EPI_LSE  (score is the logit in log2 units: scale already includes log2(e))
  float2* lse_part;     // [2 * g_splits, N] running (max, sum) in log2 domain (part = 2*split + half)
  float* diag;          // [N] raw accumulator of column t + diag_offset (nullable)
  int64_t diag_offset;
  // column log-sum-exp of the SAME pass (symmetric InfoNCE): per column j a reference col_ref[j]
  // (log2 units; its own positive logit, +inf on padding columns) and, per (query tile, TMEM lane
  // quarter), the partial sum over its 32 rows of 2^(logit - col_ref[j]) at
  // col_part[(qt * 4 + quarter) * col_ld + j]   (nullable: row statistics only)
  const float* col_ref;
  float* col_part;
  int64_t col_ld;
  // any epilogue: when non-NULL and *run_flag == 0 the whole launch exits at once (device-side
  // gating of a fallback pass, no host synchronisation)
  const unsigned int* run_flag;
  // EPI_STORE
  float* out;  // [N, ldo] (nullable when only out_op is wanted)
  int64_t ldo;
  __nv_bfloat16* out_op;  // optional: the result as query-side bf16 operand [N, out_op_kp]
  int out_op_kp;
  int out_op_split;       // 0: [x], 1: [hi | hi | lo] (VTC_PREC_EXACT)
  const float* residual;  // optional [N, ldo]
  int act;                // 0 none, 1 QuickGELU (applied after bias, before residual),
                          // 2 InfoNCE gradient weight (below)
  // act == 2 (backward of the symmetric InfoNCE, model/loss.py:18-22): with x = scale * acc the logit,
  //   z = coef * (exp(x - row_stat[t]) + exp(x - col_bias[j]) - 2 [t + diag_offset == j]),
  // coef = *coef_ptr * coef_scale (= grad_loss / 2n); col_bias holds the other side's log-sum-exp
  // (+inf on padding columns).  z leaves as the bf16 operand of the gradient product (out_op);
  // ds_part[part * N + t] = sum_j z * acc for the gradient of the logit scale.
  const float* row_stat;
  const float* coef_ptr;
  float coef_scale;
  float* ds_part;
  // EPI_TOPK
  float2* pool;       // [2 * g_splits, N, TOPK_POOL] (score, column index as int bits)
  float2* pool_meta;  // [2 * g_splits, N] (entries, tau): every column outside the pool scores >= tau
  int topk_keep;      // entries kept by a compaction (>= k, <= TOPK_KEEP_MAX)
  const float* tau_init;  // optional [N]: initial per-row threshold (from a sample pass); NULL = +inf
  int dbg_skip_epilogue;  // profiling only (VTC_DBG_SKIP_EPILOGUE=1): drain TMEM but reduce nothing
  int dbg_stages;         // profiling only (VTC_DBG_STAGES=n): use only n stages of the operand ring
  // profiling only (VTC_DBG_PROF=1): per CTA 8 x u64 {clock64 ticks, globaltimer ns, MMA-issuer
  // ticks waiting for a free accumulator (epilogue-bound), ticks waiting for operand stages
  // (load-bound), tiles, epilogue-warp ticks waiting for a full accumulator, ...}
  unsigned long long* dbg_prof;
};

// Row-major bf16 [rows, cols] with leading dimension ld (elements) -> 2-D TMA descriptor with a
// {64, box_rows} box and 128-byte swizzle.
int make_operand_tmap(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                      CUtensorMap* out);

// cluster size (1, 2 or 4) used to multicast the gallery stream; honours VTC_CLUSTER
int choose_cluster(int64_t N, int64_t M);

struct Plan {
  int cluster;  // CTAs per cluster sharing each gallery tile
  int grid;     // CTAs to launch (multiple of cluster)
  bool pair;    // cluster == 2 run as a CTA pair: one M256 cta_group::2 MMA instead of multicast
  int bn;       // gallery rows per tile: BN (256) or 128 (EPI_STORE only)
};
// fills q_tiles / g_tiles / g_splits / tiles_per_split for the given cluster size and tile width;
// min_tiles_per_split is in 256-row tiles
// balance_tail: cut the items of the last round so that every cluster gets a share (epilogues whose
// results are additive over gallery ranges, i.e. EPI_RANK)
Plan plan_tiles(Params& p, int max_splits, int cluster, int min_tiles_per_split = 8, int bn = BN,
                bool balance_tail = false);


// a_resident: keep the whole 128 x K' query tile in shared memory (needs num_kb <= 8).
// tmB must have been built with box_rows = pl.bn / pl.cluster.
int launch_sim_tc(int epilogue, bool a_resident, const Plan& pl, const CUtensorMap& tmA,
                  const CUtensorMap& tmB, const Params& p, cudaStream_t s);

// opt-in CUDA-event timing of the tensor-core launches (bench.py roofline)
void kernel_timer_enable(bool on);
int kernel_timer_read(double* total_ms, int* count);
// VTC_DBG_PROF=1: copies the per-CTA profile words of the LAST launch to the host (synchronises)
int debug_prof_read(unsigned long long* out, int max_words);

// The guard band of the rank epilogue: every tensor-core score of query t is within delta of the
// canonical one (DESIGN.md "guard band"), so a column is certainly closer than the ground truth
// below lo = d(t,gt) - delta and certainly not above hi = d(t,gt) + delta.
//   dot error <= guard_rel * |q| * max|x|;  L2: d = sq32 - 2 acc in fp32 adds the roundings of sq32
//   and of the FMA.  qq is an upper bound of |q|^2, gmax_sq the largest finite gallery norm^2.
// One entry of the guard-band list: row t (< 2^29), first column j0 of an 8-column group (a multiple
// of 8, < 2^30) and the 8-bit mask of its columns that need the canonical re-check.
__host__ __device__ inline int2 amb_pack(int t, int j0, unsigned int mask) {
  return make_int2((int)((unsigned int)t | ((mask >> 5) << 29)),
                   (int)(((unsigned int)j0 >> 3) | ((mask & 31u) << 27)));
}
__host__ __device__ inline void amb_unpack(int2 e, int* t, int* j0, unsigned int* mask) {
  const unsigned int x = (unsigned int)e.x, y = (unsigned int)e.y;
  *t = (int)(x & 0x1fffffffu);
  *j0 = (int)((y & 0x07ffffffu) << 3);
  *mask = ((x >> 29) << 5) | (y >> 27);
}

// What hi*hi + hi*lo + lo*hi drops of q.x, with q = hi + lo + e per element (hi = bf16(q),
// lo = bf16(q - hi), e = q - hi - lo, all exact in fp32):  q.x - (hq.hx + hq.lx + lq.hx) =
// q.ex + lq.lx + eq.(x - ex), so by Cauchy-Schwarz
//   |dropped| <= |q| max|ex| + |lq| max|lx| + |eq| (max|x| + max|ex|).
// The worst case over all inputs is 3 * 2^-16 |q||x| (tests/test_guard_band.py); the norms of real
// rows' pieces give a bound ~10x smaller (|lo| ~ 1.1e-3 |x|, |e| ~ 1e-6 |x|), i.e. ~10x fewer column
// groups in the fp64 re-check, and it is just as rigorous: it uses the pieces actually fed to the MMA.
__host__ __device__ inline double rank_split_bound(double qn, double lqn, double eqn, double gn,
                                                   double lxn, double exn) {
  return qn * exn + lqn * lxn + eqn * (gn + exn);
}

__host__ __device__ inline void rank_band(double d0, double qq, double gmax_sq, int metric_l2,
                                          float guard_rel, double split_abs, float* lo, float* hi) {
  const double qn = sqrt(qq), gn = sqrt(gmax_sq);
  double delta;
  if (metric_l2)
    delta = 2.0 * (guard_rel * qn * gn + split_abs) + 2.4e-7 * (gmax_sq + 2.0 * qn * gn);
  else
    delta = (double)guard_rel * qn * gn + split_abs + 1.2e-7 * qn * gn;
#ifdef __CUDA_ARCH__
  *lo = __double2float_rd(d0 - delta);
  *hi = __double2float_ru(d0 + delta);
#else
  *lo = (float)(d0 - delta);
  *hi = (float)(d0 + delta);
#endif
  if (!(qq == qq) || !(d0 == d0) || !(delta == delta)) *lo = *hi = nanf("");
}

}  // namespace tc
}  // namespace vtc
