// api.cu -- the extern "C" entry points declared in include/vtc_b200.h.
//
// Each entry point validates its arguments, carves the caller's workspace, and enqueues kernels on
// the caller's stream.  Nothing here allocates device memory or synchronises.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "cam.cuh"
#include "exact.cuh"
#include "prep.cuh"
#include "rank_stage.cuh"
#include "reduce.cuh"
#include "sim_tc.cuh"

namespace vtc {
std::atomic<uint64_t> g_launch_count{0};

bool pdl_enabled() {
  static const bool on = []() {
    const char* e = getenv("VTC_PDL");
    return !(e && *e && atoi(e) == 0);
  }();
  return on;
}

// ---- launch trace: a CUDA event on the traced stream after every launch of the library, so that
// the gaps between consecutive events are the launches' durations as they run back to back in the
// caller's stream (ncu's per-launch times are serialised and cold-cache).  Debug aid, off by default.
std::atomic<int> g_trace_on{0};
namespace {
struct LaunchTrace {
  std::mutex mu;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> events;           // events[0] = begin
  std::vector<std::pair<const char*, int>> where;
};
LaunchTrace g_trace;
}  // namespace
void trace_mark(const char* file, int line) {
  std::lock_guard<std::mutex> lk(g_trace.mu);
  if (g_trace.events.empty() || g_trace.events.size() > 4096) return;
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, g_trace.stream);
  g_trace.events.push_back(e);
  g_trace.where.emplace_back(file, line);
}

// infonce_small.cu
size_t infonce_small_ws_bytes(int64_t n);
int launch_infonce_small(const void* A, const void* B, int64_t n, int D, bool in_bf16,
                         bool round_bf16, const float* scale, float* loss, float* row_lse,
                         float* col_lse, float* diag, void* wsp, size_t ws_bytes, cudaStream_t s);

namespace {

// Bound on |tensor-core dot - exact dot| / (|q| |x|) for operands of padded depth Kp; see
// DESIGN.md "guard band".  Each of the Kp/16 accumulation steps of the fp32 accumulator may lose
// up to 1 ulp (2^-24, i.e. truncation rather than round-to-nearest) of the FULL-SCALE value |q||x|
// -- partial sums are bounded by it (Cauchy-Schwarz) -- plus 8 steps of slack for the epilogue's
// own fp32 roundings (measured total on B200: <= 1.8e-7 at K = 768, 13x inside the bound).
//   BF16 : products of bf16 are exact in fp32, so that is the whole error;
//   EXACT: x = hi + lo + e with hi = bf16(x), lo = bf16(x - hi).  bf16 carries 8 significand bits,
//          so its unit roundoff is 2^-8: |x - hi| <= 2^-8 |x|, |e| <= 2^-16 |x|, |lo| <= 2^-8 |x|.
//          hi*hi + hi*lo + lo*hi drops lo_q*lo_x + e_q*x + q*e_x (+ third-order terms), i.e. at most
//          3 * 2^-16 (1 + 2^-7) < 4.62e-5 of |q_k||x_k| per product, and by Cauchy-Schwarz of
//          |q||x| per row.
// (the rank path bounds the split term per row instead -- sim_tc.cuh::rank_split_bound -- and keeps
// only the accumulation term: guard_accum_rel)
float guard_accum_rel(int Kp) { return (float)(Kp / 16 + 8) * 5.9604645e-08f; }
float guard_rel_for(int precision, int Kp) {
  const char* e = getenv(precision == VTC_PREC_BF16 ? "VTC_GUARD_REL_BF16" : "VTC_GUARD_REL_EXACT");
  if (e && *e) {
    const float v = (float)atof(e);
    if (v > 0.f) return v;
  }
  const float accum = (float)(Kp / 16 + 8) * 5.9604645e-08f;
  return precision == VTC_PREC_BF16 ? accum : 4.62e-05f + accum;
}

bool valid_dtype(int d) { return d == VTC_F32 || d == VTC_BF16; }
bool valid_metric(int m) { return m == VTC_METRIC_DOT || m == VTC_METRIC_L2; }
bool valid_prec(int p) { return p == VTC_PREC_EXACT || p == VTC_PREC_BF16 || p == VTC_PREC_BRUTE; }

// Operand plan of a tensor-core pass: which layout the bf16 operands take and which arrays hold
// the canonical values the exact kernels read.
struct OperandPlan {
  bool split;  // 3-term bf16 split (fp32 inputs ranked exactly)
  int Kp;      // padded K' of the bf16 operands
};
OperandPlan plan_operands(int D, int dtype, int precision) {
  OperandPlan o;
  o.split = (precision == VTC_PREC_EXACT && dtype == VTC_F32);
  o.Kp = round_up(o.split ? 3 * D : D, tc::BK);
  return o;
}

constexpr int64_t kMaxRows = (int64_t)1 << 29;  // (guard-band list entries pack row, column group and a mask)

// bf16 inputs whose rows already are whole 128-byte swizzle atoms ARE the tensor-core operands:
// no prep launch, no copy (callers that keep bf16 embeddings, the multi-GPU gather and the
// pipelined host staging hand such rows in).
static bool operand_is_input(const void* X, int D, int dtype, const OperandPlan& o) {
  return dtype == VTC_BF16 && !o.split && o.Kp == D && (reinterpret_cast<uintptr_t>(X) & 15) == 0;
}

// ------------------------------------------------------------------------------------ sim_rank
struct RankWs {
  __nv_bfloat16 *opQ, *opG;
  double *sq64, *dgt;
  float *sq32, *qq;
  float2* qsplit;  // exact mode: norms of the split pieces of the query rows
  unsigned int* scalars;  // [0] max_sq_bits, [2] fallback flag, [64..] per-CTA list segment counts
  int* rank_tmp;          // directly behind `scalars`, rank_alt behind it: one memset clears all three
  int* rank_alt;          // counts of the canonical recount (fallback)
  unsigned int* hist;     // scratch of the fused finalisation (vtc_rank_eval)
  int2* amb;
  size_t amb_cap;
};
constexpr size_t kRankScalars = 64 + 256;

size_t amb_entries_wanted(int64_t N) {
  const int64_t want = 64 * N;
  return (size_t)(want < (1 << 20) ? (1 << 20) : want);
}

RankWs carve_rank(Workspace& ws, int64_t N, int64_t M, int D, int dtype, int precision,
                  bool sizing) {
  RankWs r;
  memset(&r, 0, sizeof(r));
  r.scalars = ws.take<unsigned int>(kRankScalars);  // 1280 bytes: a whole number of 256-byte granules
  r.rank_tmp = ws.take<int>(N);
  r.rank_alt = ws.take<int>(N);
  r.sq64 = ws.take<double>(M);
  r.dgt = ws.take<double>(N);
  r.hist = ws.take<unsigned int>(rank_epilogue_hist_words());
  if (precision != VTC_PREC_BRUTE) {
    const OperandPlan o = plan_operands(D, dtype, precision);
    r.opQ = ws.take<__nv_bfloat16>((size_t)N * o.Kp);
    r.opG = ws.take<__nv_bfloat16>((size_t)M * o.Kp);
    r.sq32 = ws.take<float>(round_up<int64_t>(M, tc::BN));
    r.qq = ws.take<float>(N);
    r.qsplit = ws.take<float2>(N);
    if (sizing) {
      r.amb_cap = amb_entries_wanted(N);
      ws.take<int2>(r.amb_cap);
    } else {
      size_t left = ws.size > ws.used ? ws.size - ws.used : 0;
      left -= left % 256;  // take() hands out 256-byte granules
      size_t cap = left / sizeof(int2);
      if (cap > 0x7fffffffu) cap = 0x7fffffffu;
      r.amb_cap = cap;
      r.amb = cap ? ws.take<int2>(cap) : nullptr;
    }
  }
  return r;
}

// What the fused finalisation of vtc_rank_eval needs (NULL for plain vtc_sim_rank calls).
struct RankFinalize {
  int64_t M_total;
  const int* k_vals;
  int nk;
  int64_t* hits;
  double* medr;
};

// One retrieval evaluation (or one gallery chunk of it) = memset + prologue + tensor-core pass +
// epilogue (+ its device-gated fallback launch), chained by programmatic dependent launch.
// Cached per-row quantities (vtc_sim_rank_prepared): canonical ||x_j||^2 of THIS gallery chunk
// (sq64_in, or computed into sq64_out), an upper bound of ||q_t||^2 (qq_in, or computed into
// qq_out) and d(t,gt) (gt_score, or computed into gt_score_out).  With all three given (and bf16 rows
// that are the operands) the prologue touches no row at all.
struct RankCache {
  const double* sq64_in;
  double* sq64_out;
  const float* qq_in;
  float* qq_out;
};
int sim_rank_impl(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                  const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                  int precision, const double* gt_score, double* gt_score_out, int accumulate,
                  int32_t* rank0, void* wsp, size_t ws_bytes, cudaStream_t s,
                  const RankCache* cache = nullptr, const RankFinalize* fin = nullptr) {
  const double* sq64_in = cache ? cache->sq64_in : nullptr;
  const float* qq_in = cache ? cache->qq_in : nullptr;
  if (N < 0 || M < 0 || D <= 0 || !valid_dtype(dtype) || !valid_metric(metric) ||
      !valid_prec(precision))
    return VTC_ERR_INVALID_ARG;
  if (cache && (precision == VTC_PREC_BRUTE || M == 0)) return VTC_ERR_INVALID_ARG;
  if (fin && (fin->nk < 0 || fin->nk > 8 || (fin->nk > 0 && (!fin->k_vals || !fin->hits))))
    return VTC_ERR_INVALID_ARG;
  if (N == 0) {  // no queries: nothing to rank (pointers may be NULL)
    if (fin)
      return launch_rank_finalize(rank0, nullptr, 0, fin->M_total, fin->k_vals, fin->nk, fin->hits,
                                  fin->medr, nullptr, s);
    return VTC_OK;
  }
  if (!Q || (!G && M > 0) || !rank0) return VTC_ERR_INVALID_ARG;
  if (N > kMaxRows || M > kMaxRows || D > 8192) return VTC_ERR_UNSUPPORTED_SHAPE;
  Workspace ws(wsp, ws_bytes);
  RankWs w = carve_rank(ws, N, M, D, dtype, precision, false);
  if (!ws.ok() || !w.sq64 || !w.dgt || !w.scalars || !w.rank_tmp || !w.rank_alt || !w.hist)
    return VTC_ERR_WORKSPACE;
  const bool in_bf16 = dtype == VTC_BF16;
  cudaError_t e = cudaMemsetAsync(
      w.scalars, 0, kRankScalars * sizeof(unsigned int) + 2 * round_up<size_t>(sizeof(int) * N, 256), s);
  if (e != cudaSuccess) return cuda_err(e);

  if (precision == VTC_PREC_BRUTE || M == 0) {
    ExactArgs ex{Q, G, D, D, in_bf16, N, M, D, w.sq64, gt, row_offset, col_offset, metric};
    if (M > 0 && metric == VTC_METRIC_L2)
      VTC_RETURN_IF_ERROR(launch_sqnorm64(G, in_bf16, M, D, D, w.sq64, nullptr, nullptr, s));
    VTC_RETURN_IF_ERROR(launch_gt_score(ex, gt_score, w.dgt, nullptr, nullptr, 0.f, s));
    if (gt_score_out) {
      e = cudaMemcpyAsync(gt_score_out, w.dgt, sizeof(double) * N, cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return cuda_err(e);
    }
    if (!accumulate) {
      e = cudaMemsetAsync(rank0, 0, sizeof(int32_t) * N, s);
      if (e != cudaSuccess) return cuda_err(e);
    }
    VTC_RETURN_IF_ERROR(launch_rank_brute(ex, w.dgt, rank0, nullptr, s));
    if (fin)
      return launch_rank_finalize(rank0, w.dgt, N, fin->M_total, fin->k_vals, fin->nk, fin->hits,
                                  fin->medr, w.hist, s);
    return VTC_OK;
  }

  if (!w.opQ || !w.opG || !w.sq32 || !w.qq || !w.amb || w.amb_cap < 1024) return VTC_ERR_WORKSPACE;
  const OperandPlan o = plan_operands(D, dtype, precision);
  const bool aliasQ = operand_is_input(Q, D, dtype, o), aliasG = operand_is_input(G, D, dtype, o);
  const __nv_bfloat16* opQ = aliasQ ? static_cast<const __nv_bfloat16*>(Q) : w.opQ;
  const __nv_bfloat16* opG = aliasG ? static_cast<const __nv_bfloat16*>(G) : w.opG;
  // computed quantities land where the caller wants them
  double* sq64_w = (cache && cache->sq64_out) ? cache->sq64_out : w.sq64;
  float* qq_w = (cache && cache->qq_out) ? cache->qq_out : w.qq;
  double* dgt_w = (!gt_score && gt_score_out) ? gt_score_out : w.dgt;
  const double* sq64 = sq64_in ? sq64_in : sq64_w;
  const double* dgt = gt_score ? gt_score : dgt_w;
  const float* qq = qq_in ? qq_in : qq_w;
  // canonical values: the fp32 inputs (split) or the bf16 operands
  ExactArgs ex;
  if (o.split)
    ex = ExactArgs{Q, G, D, D, false, N, M, D, sq64, gt, row_offset, col_offset, metric};
  else
    ex = ExactArgs{opQ, opG, o.Kp, o.Kp, true, N, M, D, sq64, gt, row_offset, col_offset, metric};
  // 1. prologue: operands, canonical norms, ground-truth scores, bias, guard-band inputs
  const int64_t Mpad = round_up<int64_t>(M, tc::BN);
  RankPrologueArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.Q = Q, pa.G = G, pa.in_bf16 = in_bf16 ? 1 : 0, pa.N = N, pa.M = M, pa.D = D;
  pa.ldq = D, pa.ldg = D;
  pa.mode_q = aliasQ ? STAGE_NONE : (o.split ? PREP_SPLIT_A : PREP_PLAIN);
  pa.mode_g = aliasG ? STAGE_NONE : (o.split ? PREP_SPLIT_B : PREP_PLAIN);
  pa.round_bf16 = (!o.split && !in_bf16) ? 1 : 0;
  pa.opQ = w.opQ, pa.opG = w.opG, pa.Kp = o.Kp, pa.metric = metric;
  pa.sq64_in = sq64_in, pa.sq64 = sq64_w, pa.bias = w.sq32, pa.Mpad = Mpad;
  pa.max_sq_bits = &w.scalars[0];
  pa.gt_in = gt_score, pa.dgt = dgt_w, pa.qq_in = qq_in, pa.qq = qq_w;
  pa.gt = gt, pa.row_offset = row_offset, pa.col_offset = col_offset;
  pa.fallback = &w.scalars[2];
  pa.qsplit = o.split ? w.qsplit : nullptr, pa.split_max_bits = &w.scalars[4];
  VTC_RETURN_IF_ERROR(launch_rank_prologue(pa, s));
  if (gt_score && gt_score_out && gt_score_out != gt_score) {
    e = cudaMemcpyAsync(gt_score_out, gt_score, sizeof(double) * N, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return cuda_err(e);
  }
  // 2. tensor-core pass
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = N, p.M = M, p.num_kb = o.Kp / tc::BK;
  p.gt = gt, p.gt_row_offset = row_offset, p.gt_col_offset = col_offset;
  p.col_bias = w.sq32;
  p.scale = metric == VTC_METRIC_L2 ? -2.f : -1.f;
  p.dgt = dgt, p.qq = qq, p.max_sq_bits = &w.scalars[0];
  // (an override from the environment is a constant relative bound, as in round 1)
  const bool row_split = o.split && !getenv("VTC_GUARD_REL_EXACT");
  p.guard_rel = row_split ? guard_accum_rel(o.Kp) : guard_rel_for(precision, o.Kp);
  p.qsplit = row_split ? w.qsplit : nullptr, p.split_max_bits = &w.scalars[4];
  p.metric_l2 = metric == VTC_METRIC_L2 ? 1 : 0;
  p.rank = w.rank_tmp, p.amb_list = w.amb, p.amb_seg_count = &w.scalars[64];
  // (rank counts are additive over gallery ranges: the last round of work items is balanced)
  tc::Plan pl = tc::plan_tiles(p, 64, tc::choose_cluster(N, M), 8, tc::BN, true);
  if (pl.grid > 256) return VTC_ERR_UNSUPPORTED_SHAPE;
  p.amb_seg_cap = (unsigned int)(w.amb_cap / (size_t)pl.grid);
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opQ, N, o.Kp, o.Kp, tc::BM, &tmA));
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opG, M, o.Kp, o.Kp, pl.bn / pl.cluster, &tmB));
  VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_RANK, p.num_kb <= 8, pl, tmA, tmB, p, s));
  // 3. epilogue: exact re-check of the guard-band groups (brute force if the list overflowed),
  //    commit, and -- for vtc_rank_eval -- hit counts and the median rank
  RankEpilogueArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.ex = ex, ea.amb_list = w.amb, ea.seg_count = &w.scalars[64], ea.nseg = pl.grid;
  ea.seg_cap = p.amb_seg_cap, ea.dgt = dgt, ea.rank_tmp = w.rank_tmp, ea.rank_alt = w.rank_alt;
  ea.fallback = &w.scalars[2];
  ea.rank0 = rank0, ea.accumulate = accumulate;
  if (fin) {
    ea.finalize = 1, ea.M_total = fin->M_total, ea.nk = fin->nk;
    for (int i = 0; i < fin->nk; ++i) ea.k_vals[i] = fin->k_vals[i];
    ea.hits = reinterpret_cast<unsigned long long*>(fin->hits), ea.medr = fin->medr;
    ea.hist = w.hist;
  }
  return launch_rank_epilogue(ea, s);
}

// ------------------------------------------------------------------------------------ sim_topk
struct TopkWs {
  __nv_bfloat16 *opQ, *opG;
  double* sq64;
  float* sq32;
  unsigned int* scalars;
  float2* pool;
  float2* pool_meta;
  unsigned int* row_flag;
  float* tau0;
  float* sample;  // [N, sample columns] approximate scores of the sample pass
};
constexpr int kTopkMaxSplits = 8;

// Gallery tiles (of 256 rows) scored densely to seed the per-row top-k thresholds: ~1/32 of the
// gallery, 8..32 tiles, capped so that the [N, columns] fp32 scratch stays below 512 MB; 0 = no
// sample pass (small galleries are cheap to stream without a threshold).
static int64_t topk_sample_tiles(int64_t N, int64_t M) {
  const int64_t g_tiles = ceil_div<int64_t>(M, tc::BN);
  if (g_tiles < 128 || N <= 0) return 0;
  int64_t t = g_tiles / 32;
  if (t < 8) t = 8;
  if (t > 32) t = 32;
  // tuning knob (VTC_TOPK_SAMPLE_TILES=4..64): fewer tiles = a cheaper sample pass (its dense scores
  // are written and read back once) but looser thresholds, i.e. more candidates in the main pass
  if (const char* e = getenv("VTC_TOPK_SAMPLE_TILES")) {
    const int64_t v = atoll(e);
    if (v >= 4 && v <= 64 && v < g_tiles) t = v;
  }
  const int64_t cap = ((int64_t)512 << 20) / (N * tc::BN * 4);
  if (t > cap) t = cap;
  return t >= 4 ? t : 0;
}

TopkWs carve_topk(Workspace& ws, int64_t N, int64_t M, int D, int dtype, int precision) {
  TopkWs t;
  memset(&t, 0, sizeof(t));
  const int prec = precision == VTC_PREC_BRUTE ? VTC_PREC_EXACT : precision;
  const OperandPlan o = plan_operands(D, dtype, prec);
  t.sq64 = ws.take<double>(M);
  t.sq32 = ws.take<float>(round_up<int64_t>(M, tc::BN));
  t.scalars = ws.take<unsigned int>(64);
  t.row_flag = ws.take<unsigned int>(N);
  t.opQ = ws.take<__nv_bfloat16>((size_t)N * o.Kp);
  t.opG = ws.take<__nv_bfloat16>((size_t)M * o.Kp);
  t.pool = ws.take<float2>((size_t)2 * kTopkMaxSplits * N * tc::TOPK_POOL);
  t.pool_meta = ws.take<float2>((size_t)2 * kTopkMaxSplits * N);
  t.tau0 = ws.take<float>(N);
  const int64_t st = precision == VTC_PREC_BRUTE ? 0 : topk_sample_tiles(N, M);
  t.sample = st ? ws.take<float>((size_t)N * st * tc::BN) : nullptr;
  return t;
}

int sim_topk_impl(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype, int metric,
                  int precision, int k, int64_t col_offset, float* out_val, int64_t* out_idx,
                  void* wsp, size_t ws_bytes, cudaStream_t s) {
  if (N < 0 || M < 0 || D <= 0 || !valid_dtype(dtype) || !valid_metric(metric) ||
      !valid_prec(precision) || k < 1)
    return VTC_ERR_INVALID_ARG;
  if (N == 0) return VTC_OK;  // no queries: nothing to search (pointers may be NULL)
  if (!Q || (!G && M > 0) || !out_val || !out_idx) return VTC_ERR_INVALID_ARG;
  if (k > 16 || N > kMaxRows || M > kMaxRows || D > 8192) return VTC_ERR_UNSUPPORTED_SHAPE;
  Workspace ws(wsp, ws_bytes);
  TopkWs w = carve_topk(ws, N, M, D, dtype, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const bool in_bf16 = dtype == VTC_BF16;
  cudaError_t e = cudaMemsetAsync(w.scalars, 0, 64 * sizeof(unsigned int), s);
  if (e != cudaSuccess) return cuda_err(e);
  const bool brute = precision == VTC_PREC_BRUTE || M == 0;
  const OperandPlan o = plan_operands(D, dtype, brute ? VTC_PREC_EXACT : precision);
  const bool canon_inputs = brute || o.split;

  TopkSelectArgs a;
  memset(&a, 0, sizeof(a));
  const bool aliasQ = !brute && operand_is_input(Q, D, dtype, o);
  const bool aliasG = !brute && operand_is_input(G, D, dtype, o);
  const __nv_bfloat16* opQ = aliasQ ? static_cast<const __nv_bfloat16*>(Q) : w.opQ;
  const __nv_bfloat16* opG = aliasG ? static_cast<const __nv_bfloat16*>(G) : w.opG;
  if (canon_inputs)
    a.ex = ExactArgs{Q, G, D, D, in_bf16, N, M, D, w.sq64, nullptr, 0, col_offset, metric};
  else
    a.ex = ExactArgs{opQ, opG, o.Kp, o.Kp, true, N, M, D, w.sq64, nullptr, 0, col_offset, metric};
  a.pool_buf = w.pool, a.pool_meta = w.pool_meta;
  a.pool = tc::TOPK_POOL, a.k = k;
  a.guard_rel = guard_rel_for(precision, o.Kp);
  a.max_sq_bits = &w.scalars[0];
  a.out_val = out_val, a.out_idx = out_idx, a.row_flag = w.row_flag;

  if (brute) {
    if (M > 0)
      VTC_RETURN_IF_ERROR(launch_sqnorm64(G, in_bf16, M, D, D, w.sq64, nullptr, nullptr, s));
    e = cudaMemsetAsync(w.row_flag, 0xff, sizeof(unsigned int) * N, s);  // every row brute-forced
    if (e != cudaSuccess) return cuda_err(e);
    return launch_topk_brute_rows(a, s);
  }

  // one staged pass over the rows (rank_stage.cu): operands of both sides, canonical ||x||^2, the
  // padded epilogue bias and the largest norm -- round 1 spent four launches on this
  {
    RankPrologueArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.Q = Q, pa.G = G, pa.in_bf16 = in_bf16 ? 1 : 0, pa.N = N, pa.M = M, pa.D = D;
    pa.ldq = D, pa.ldg = D;
    pa.mode_q = aliasQ ? STAGE_NONE : (o.split ? PREP_SPLIT_A : PREP_PLAIN);
    pa.mode_g = aliasG ? STAGE_NONE : (o.split ? PREP_SPLIT_B : PREP_PLAIN);
    pa.round_bf16 = (!o.split && !in_bf16) ? 1 : 0;
    pa.opQ = w.opQ, pa.opG = w.opG, pa.Kp = o.Kp, pa.metric = metric;
    pa.sq64 = w.sq64, pa.bias = w.sq32, pa.Mpad = round_up<int64_t>(M, tc::BN);
    pa.max_sq_bits = &w.scalars[0];
    if (aliasQ) pa.N = 0;  // nothing to do on the query side
    VTC_RETURN_IF_ERROR(launch_rank_prologue(pa, s));
  }
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = N, p.M = M, p.num_kb = o.Kp / tc::BK;
  p.col_bias = w.sq32;
  p.scale = metric == VTC_METRIC_L2 ? -2.f : -1.f;
  p.pool = w.pool, p.pool_meta = w.pool_meta;
  p.topk_keep = k <= 12 ? 16 : tc::TOPK_KEEP_MAX;
  const int cluster = tc::choose_cluster(N, M);
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opQ, N, o.Kp, o.Kp, tc::BM, &tmA));
  // Sample pass: the first ~1/32 of the gallery is scored densely (plain tensor-core product into
  // a scratch matrix) and reduced to a per-row threshold that the k best of the WHOLE gallery
  // provably beat, so the main pass appends a handful of candidates per row and (almost) never
  // has to re-sort a buffer.
  const int64_t sample_tiles = topk_sample_tiles(N, M);
  if (sample_tiles > 0 && w.sample && !getenv("VTC_TOPK_NO_SAMPLE")) {
    tc::Params ps = p;
    ps.M = sample_tiles * tc::BN;
    ps.pool = nullptr, ps.pool_meta = nullptr;
    ps.out = w.sample, ps.ldo = ps.M;
    const tc::Plan pls = tc::plan_tiles(ps, 64, 1, 1);
    CUtensorMap tmS;
    VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opG, ps.M, o.Kp, o.Kp, tc::BN, &tmS));
    VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_STORE, ps.num_kb <= 8, pls, tmA, tmS, ps, s));
    VTC_RETURN_IF_ERROR(launch_topk_tau(a, w.sample, (int)ps.M, ps.M, w.tau0, s));
    p.tau_init = w.tau0;
  }
  const tc::Plan pl = tc::plan_tiles(p, kTopkMaxSplits, cluster);
  a.splits = 2 * p.g_splits;  // two column halves per gallery split
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opG, M, o.Kp, o.Kp, tc::BN / pl.cluster, &tmB));
  VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_TOPK, p.num_kb <= 8, pl, tmA, tmB, p, s));
  VTC_RETURN_IF_ERROR(launch_topk_select(a, s));
  return launch_topk_brute_rows(a, s);
}

// Tile width of a dense product: the widest of 256 / 128 / 64 gallery rows per tile that still gives
// two thirds of the SMs a tile (the skinny projections of the CAM: 1536 x 512 outputs are 24 tiles of
// 256 columns but 96 of 64).  These launches are bound by the operand bytes each CTA pulls from L2.
int store_tile_width(int64_t rows, int64_t cols) {
  const int64_t q_tiles = ceil_div<int64_t>(rows, tc::BM);
  for (int bn = tc::BN; bn > 64; bn /= 2)
    if (q_tiles * ceil_div<int64_t>(cols, bn) * 3 >= kNumSMs * 2) return bn;
  return 64;
}

// --------------------------------------------------------------------------- dense TC products
// out[N,M] = act(scale * A B^T + bias) + residual, fp32 in/out.
struct GemmWs {
  __nv_bfloat16 *opA, *opB;
  float* bias;  // padded per-column bias
};
GemmWs carve_gemm(Workspace& ws, int64_t N, int64_t M, int D, int dtype, int precision) {
  const OperandPlan o = plan_operands(D, dtype, precision);
  GemmWs g;
  g.opA = ws.take<__nv_bfloat16>((size_t)N * o.Kp);
  g.opB = ws.take<__nv_bfloat16>((size_t)M * o.Kp);
  g.bias = ws.take<float>(round_up<int64_t>(M, tc::BN));
  return g;
}

int gemm_store_impl(const void* A, const void* B, int64_t N, int64_t M, int D, int dtype,
                    int precision, const float* scale_ptr, const float* bias,
                    const float* residual, int act, float* out, int64_t ldo, void* wsp,
                    size_t ws_bytes, cudaStream_t s) {
  if (!A || !B || !out || N < 0 || M < 0 || D <= 0 || !valid_dtype(dtype) || ldo < M ||
      (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16))
    return VTC_ERR_INVALID_ARG;
  if (N > kMaxRows || M > kMaxRows || D > 8192) return VTC_ERR_UNSUPPORTED_SHAPE;
  if (N == 0 || M == 0) return VTC_OK;
  Workspace ws(wsp, ws_bytes);
  GemmWs g = carve_gemm(ws, N, M, D, dtype, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const OperandPlan o = plan_operands(D, dtype, precision);
  const bool in_bf16 = dtype == VTC_BF16;
  VTC_RETURN_IF_ERROR(launch_prep_operand(A, in_bf16, N, D, D, o.split ? PREP_SPLIT_A : PREP_PLAIN,
                                          g.opA, o.Kp, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(B, in_bf16, M, D, D, o.split ? PREP_SPLIT_B : PREP_PLAIN,
                                          g.opB, o.Kp, s));
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = N, p.M = M, p.num_kb = o.Kp / tc::BK;
  VTC_RETURN_IF_ERROR(launch_fill_bias(g.bias, bias, M, round_up<int64_t>(M, tc::BN), 0.f, s));
  p.col_bias = g.bias, p.scale_ptr = scale_ptr, p.scale = 1.f;
  p.out = out, p.ldo = ldo, p.residual = residual, p.act = act;
  const tc::Plan pl = tc::plan_tiles(p, 64, 1, 1, store_tile_width(N, M));
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(g.opA, N, o.Kp, o.Kp, tc::BM, &tmA));
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(g.opB, M, o.Kp, o.Kp, pl.bn, &tmB));
  return tc::launch_sim_tc(tc::EPI_STORE, p.num_kb <= 8, pl, tmA, tmB, p, s);
}

// ------------------------------------------------------------------------------------- InfoNCE
struct NceWs {
  __nv_bfloat16 *opA, *opB;  // A as query-side operand, B as gallery-side operand
  float2 *part_row, *part_col;
  float* diag_raw;
  float* bias;     // zeros, -inf padding
  float* col_ref;  // per-column reference of the one-pass column sums (+inf padding)
  float* col_part; // [q_tiles * 4, npad] partial column sums
  unsigned int* flag;
};
constexpr int kNceMaxSplits = 8;
constexpr int64_t kNceSmallMax = 2048;
NceWs carve_nce(Workspace& ws, int64_t n, int D, int dtype, int precision) {
  const OperandPlan o = plan_operands(D, dtype, precision);
  const int64_t npad = round_up<int64_t>(n, tc::BN);
  NceWs w;
  memset(&w, 0, sizeof(w));
  w.opA = ws.take<__nv_bfloat16>((size_t)n * o.Kp);
  w.opB = ws.take<__nv_bfloat16>((size_t)n * o.Kp);
  w.part_row = ws.take<float2>((size_t)2 * kNceMaxSplits * n);
  w.part_col = ws.take<float2>((size_t)2 * kNceMaxSplits * n);
  w.diag_raw = ws.take<float>(n);
  w.bias = ws.take<float>(npad);
  w.col_ref = ws.take<float>(npad);
  w.col_part = ws.take<float>((size_t)ceil_div<int64_t>(n, tc::BM) * 4 * npad);
  w.flag = ws.take<unsigned int>(64);
  return w;
}

int infonce_fwd_impl(const void* A, const void* B, int64_t n, int D, int dtype, int precision,
                     const float* scale, float* loss, float* row_lse, float* col_lse, float* diag,
                     void* wsp, size_t ws_bytes, cudaStream_t s) {
  if (!A || !B || !scale || !loss || !row_lse || !col_lse || !diag || n <= 0 || D <= 0 ||
      !valid_dtype(dtype) || (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16))
    return VTC_ERR_INVALID_ARG;
  if (n > kMaxRows || D > 8192) return VTC_ERR_UNSUPPORTED_SHAPE;
  // training-sized batches are launch-latency bound: one fused SIMT launch (infonce_small.cu);
  // larger ones stream through the tcgen05 kernel with the online-LSE epilogue
  const bool force_tc = getenv("VTC_INFONCE_FORCE_TC") != nullptr;  // test knob (tests/ only)
  if (n <= kNceSmallMax && !force_tc)
    return launch_infonce_small(A, B, n, D, dtype == VTC_BF16, precision == VTC_PREC_BF16, scale,
                                loss, row_lse, col_lse, diag, wsp, ws_bytes, s);
  Workspace ws(wsp, ws_bytes);
  NceWs w = carve_nce(ws, n, D, dtype, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const OperandPlan o = plan_operands(D, dtype, precision);
  const bool in_bf16 = dtype == VTC_BF16;
  const int64_t npad = round_up<int64_t>(n, tc::BN);
  cudaError_t e = cudaMemsetAsync(w.flag, 0, 64 * sizeof(unsigned int), s);
  if (e != cudaSuccess) return cuda_err(e);
  VTC_RETURN_IF_ERROR(launch_prep_operand(A, in_bf16, n, D, D, o.split ? PREP_SPLIT_A : PREP_PLAIN,
                                          w.opA, o.Kp, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(B, in_bf16, n, D, D, o.split ? PREP_SPLIT_B : PREP_PLAIN,
                                          w.opB, o.Kp, s));
  VTC_RETURN_IF_ERROR(launch_fill_bias(w.bias, nullptr, n, npad, -INFINITY, s));
  VTC_RETURN_IF_ERROR(launch_nce_colref(w.opA, w.opB, n, npad, o.Kp, scale, w.col_ref, s));
  // ONE pass over the logits: online row log-sum-exp, diagonal, and the column sums against the
  // per-column reference (model/loss.py:21 evaluates cross_entropy on sim and on sim.t())
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = n, p.M = n, p.num_kb = o.Kp / tc::BK;
  p.scale_ptr = scale, p.scale = 1.4426950408889634f;  // logits in log2 units
  p.col_bias = w.bias;
  p.lse_part = w.part_row, p.diag = w.diag_raw, p.diag_offset = 0;
  p.col_ref = w.col_ref, p.col_part = w.col_part, p.col_ld = npad;
  tc::Plan pl = tc::plan_tiles(p, kNceMaxSplits, 1);
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(w.opA, n, o.Kp, o.Kp, tc::BM, &tmA));
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(w.opB, n, o.Kp, o.Kp, tc::BN, &tmB));
  VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_LSE, p.num_kb <= 8, pl, tmA, tmB, p, s));
  VTC_RETURN_IF_ERROR(launch_lse_merge(w.part_row, 2 * p.g_splits, n, row_lse, s));
  VTC_RETURN_IF_ERROR(launch_nce_col_merge(w.col_part, p.q_tiles * 4, npad, w.col_ref, n, col_lse,
                                           &w.flag[0], s));
  // fallback, gated on the device (launches that exit at once unless a column sum overflowed): the
  // transposed product with the online row statistics.  The split operands need no second
  // preparation: [hi|lo|hi] . [hi|hi|lo] multiplies the same three term pairs with the roles swapped.
  {
    tc::Params q;
    memset(&q, 0, sizeof(q));
    q.N = n, q.M = n, q.num_kb = o.Kp / tc::BK;
    q.scale_ptr = scale, q.scale = 1.4426950408889634f;
    q.col_bias = w.bias;
    q.lse_part = w.part_col;
    q.run_flag = &w.flag[0];
    tc::Plan pl2 = tc::plan_tiles(q, kNceMaxSplits, 1);
    CUtensorMap tmA2, tmB2;
    VTC_RETURN_IF_ERROR(tc::make_operand_tmap(w.opB, n, o.Kp, o.Kp, tc::BM, &tmA2));
    VTC_RETURN_IF_ERROR(tc::make_operand_tmap(w.opA, n, o.Kp, o.Kp, tc::BN, &tmB2));
    VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_LSE, q.num_kb <= 8, pl2, tmA2, tmB2, q, s));
    VTC_RETURN_IF_ERROR(launch_lse_merge(w.part_col, 2 * q.g_splits, n, col_lse, s, &w.flag[0]));
  }
  return launch_infonce_loss(row_lse, col_lse, w.diag_raw, scale, n, diag, loss, s);
}

// Backward of the symmetric InfoNCE on the tensor cores (n > kNceSmallMax; smaller batches take the
// single SIMT path of infonce_bwd.cu):  W = g/2n (softmax_row + softmax_col - 2I),  dA = s W B,
// dB = s W^T A,  ds = sum W .* (A B^T)   (the autograd of model/loss.py:18-22 through
// model/model.py:369).  W is never formed as an n x n matrix: for a block of R query rows
//   (1) the logit tiles are RECOMPUTED by the tcgen05 GEMM (same operands and precision as the
//       forward, so the exponentials meet the saved log-sum-exps they were reduced from) and leave
//       the epilogue as gradient weights in bf16 OPERAND form (StoreEpi act = 2): an R x n strip,
//   (2) a second tcgen05 GEMM multiplies the strip with B^T-as-operand: dA[block] = s W_blk B.
// dB is the same computation with the roles of A and B swapped (W^T of (A, B) is W of (B, A) with
// the log-sum-exps exchanged), so no transposed strip and no accumulation across blocks is needed:
// 4 n^2 D tensor-core FLOPs instead of the minimal 3, and scratch O(R n) (<= 64 MB) for any n.
constexpr size_t kNceBwdStripBytes = (size_t)64 << 20;
struct NceBwdWs {
  __nv_bfloat16 *opA, *opB, *opAt, *opBt, *strip;
  float *tA, *tB, *colb, *rowb, *zb, *dsp;
  int64_t R;
  int KpD, KpN;
};
NceBwdWs carve_nce_bwd(Workspace& ws, int64_t n, int D, int precision) {
  NceBwdWs w;
  memset(&w, 0, sizeof(w));
  const OperandPlan od = plan_operands(D, VTC_F32, precision);
  const int64_t kn = round_up<int64_t>((od.split ? 3 : 1) * n, tc::BK);
  w.KpD = od.Kp, w.KpN = (int)kn;
  int64_t R = (int64_t)(kNceBwdStripBytes / ((size_t)kn * 2)) / tc::BM * tc::BM;
  if (R < tc::BM) R = tc::BM;
  if (R > n) R = n;
  w.R = R;
  const int64_t npad = round_up<int64_t>(n, tc::BN);
  w.opA = ws.take<__nv_bfloat16>((size_t)n * od.Kp);
  w.opB = ws.take<__nv_bfloat16>((size_t)n * od.Kp);
  w.tA = ws.take<float>((size_t)D * n);
  w.tB = ws.take<float>((size_t)D * n);
  w.opAt = ws.take<__nv_bfloat16>((size_t)D * kn);
  w.opBt = ws.take<__nv_bfloat16>((size_t)D * kn);
  w.strip = ws.take<__nv_bfloat16>((size_t)R * kn);
  w.colb = ws.take<float>(npad);
  w.rowb = ws.take<float>(npad);
  w.zb = ws.take<float>(round_up<int64_t>(D, tc::BN));
  w.dsp = ws.take<float>((size_t)128 * R);
  return w;
}

}  // namespace

// cam_bwd.cu / infonce_bwd.cu
int launch_transpose(const float* in, int64_t R, int64_t C, float* out, cudaStream_t s);
int launch_nce_ds_reduce(const float* part, int64_t count, float* out, cudaStream_t s);

int infonce_bwd_tc_impl(const float* A, const float* B, int64_t n, int D, int precision,
                        const float* scale, const float* row_lse, const float* col_lse,
                        const float* grad_loss, float* dA, float* dB, float* dscale, void* wsp,
                        size_t ws_bytes, cudaStream_t s) {
  if (n > kMaxRows || D > 8192 || 3 * n > 0x7ffffff0) return VTC_ERR_UNSUPPORTED_SHAPE;
  Workspace ws(wsp, ws_bytes);
  NceBwdWs w = carve_nce_bwd(ws, n, D, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const OperandPlan od = plan_operands(D, VTC_F32, precision);
  const int split = od.split ? 1 : 0;
  const int64_t npad = round_up<int64_t>(n, tc::BN);
  VTC_RETURN_IF_ERROR(launch_prep_operand(A, false, n, D, D, split ? PREP_SPLIT_A : PREP_PLAIN, w.opA, w.KpD, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(B, false, n, D, D, split ? PREP_SPLIT_B : PREP_PLAIN, w.opB, w.KpD, s));
  // A^T, B^T as gallery-side operands [D, K' = n]: the right-hand sides of the gradient products
  VTC_RETURN_IF_ERROR(launch_transpose(A, n, D, w.tA, s));
  VTC_RETURN_IF_ERROR(launch_transpose(B, n, D, w.tB, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(w.tA, false, D, (int)n, n, split ? PREP_SPLIT_B : PREP_PLAIN, w.opAt, w.KpN, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(w.tB, false, D, (int)n, n, split ? PREP_SPLIT_B : PREP_PLAIN, w.opBt, w.KpN, s));
  VTC_RETURN_IF_ERROR(launch_fill_bias(w.colb, col_lse, n, npad, INFINITY, s));
  VTC_RETURN_IF_ERROR(launch_fill_bias(w.rowb, row_lse, n, npad, INFINITY, s));
  VTC_RETURN_IF_ERROR(launch_fill_bias(w.zb, nullptr, D, round_up<int64_t>(D, tc::BN), 0.f, s));
  // padding columns of the strip stay zero for the whole call (the epilogue never writes them)
  cudaError_t e = cudaMemsetAsync(w.strip, 0, (size_t)w.R * w.KpN * sizeof(__nv_bfloat16), s);
  if (e == cudaSuccess) e = cudaMemsetAsync(dscale, 0, sizeof(float), s);
  if (e != cudaSuccess) return cuda_err(e);
  for (int dir = 0; dir < 2; ++dir) {
    // dir 0: rows of A against B -> dA (and ds);  dir 1: rows of B against A -> dB
    const __nv_bfloat16* opX = dir == 0 ? w.opA : w.opB;
    const __nv_bfloat16* opY = dir == 0 ? w.opB : w.opA;
    const __nv_bfloat16* opYt = dir == 0 ? w.opBt : w.opAt;
    const float* rstat = dir == 0 ? row_lse : col_lse;
    const float* cbias = dir == 0 ? w.colb : w.rowb;
    float* dX = dir == 0 ? dA : dB;
    for (int64_t r0 = 0; r0 < n; r0 += w.R) {
      const int64_t rows = n - r0 < w.R ? n - r0 : w.R;
      tc::Params p;
      memset(&p, 0, sizeof(p));
      p.N = rows, p.M = n, p.num_kb = w.KpD / tc::BK;
      p.col_bias = cbias, p.scale_ptr = scale, p.scale = 1.f;
      p.act = 2, p.row_stat = rstat + r0, p.coef_ptr = grad_loss, p.coef_scale = 0.5f / (float)n;
      p.diag_offset = r0, p.ds_part = dir == 0 ? w.dsp : nullptr;
      p.out_op = w.strip, p.out_op_kp = w.KpN, p.out_op_split = split;
      const tc::Plan pl = tc::plan_tiles(p, 64, 1, 1, store_tile_width(rows, n));
      CUtensorMap tmA, tmB;
      VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opX + (size_t)r0 * w.KpD, rows, w.KpD, w.KpD, tc::BM, &tmA));
      VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opY, n, w.KpD, w.KpD, pl.bn, &tmB));
      VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_STORE, false, pl, tmA, tmB, p, s));
      if (dir == 0)
        VTC_RETURN_IF_ERROR(launch_nce_ds_reduce(w.dsp, (int64_t)2 * p.g_splits * rows, dscale, s));
      tc::Params q;
      memset(&q, 0, sizeof(q));
      q.N = rows, q.M = D, q.num_kb = w.KpN / tc::BK;
      q.col_bias = w.zb, q.scale_ptr = scale, q.scale = 1.f;
      q.out = dX + (size_t)r0 * D, q.ldo = D;
      const tc::Plan pl2 = tc::plan_tiles(q, 64, 1, 1, store_tile_width(rows, D));
      CUtensorMap tmW, tmY;
      VTC_RETURN_IF_ERROR(tc::make_operand_tmap(w.strip, rows, w.KpN, w.KpN, tc::BM, &tmW));
      VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opYt, D, w.KpN, w.KpN, pl2.bn, &tmY));
      VTC_RETURN_IF_ERROR(tc::launch_sim_tc(tc::EPI_STORE, false, pl2, tmW, tmY, q, s));
    }
  }
  return VTC_OK;
}

namespace {


// ----------------------------------------------------------------------- prepared linears / CAM
// A "prepared" linear holds its weight as the gallery-side bf16 operand [out_f, Kp] followed by the
// bias padded to a multiple of 256 floats; both are written once (vtc_linear_prepare) and reused
// by every forward, so a linear is ONE launch: TMA-fed tcgen05 GEMM with bias / QuickGELU /
// residual in the epilogue, output as fp32 and/or as the next GEMM's bf16 operand.
size_t prepared_linear_bytes(int in_f, int out_f, int precision) {
  const OperandPlan o = plan_operands(in_f, VTC_F32, precision);
  return round_up<size_t>((size_t)out_f * o.Kp * sizeof(__nv_bfloat16), 256) +
         round_up<size_t>(round_up<int64_t>(out_f, tc::BN) * sizeof(float), 256);
}
const float* prepared_bias(const void* prepared, int in_f, int out_f, int precision) {
  const OperandPlan o = plan_operands(in_f, VTC_F32, precision);
  return (const float*)((const char*)prepared +
                        round_up<size_t>((size_t)out_f * o.Kp * sizeof(__nv_bfloat16), 256));
}

int linear_prepared(const __nv_bfloat16* Xop, const void* prepared, const float* residual,
                    int64_t rows, int in_f, int out_f, int act, int precision, float* Y,
                    __nv_bfloat16* Yop, int Yop_kp, cudaStream_t s) {
  const OperandPlan o = plan_operands(in_f, VTC_F32, precision);
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = rows, p.M = out_f, p.num_kb = o.Kp / tc::BK;
  p.col_bias = prepared_bias(prepared, in_f, out_f, precision);
  p.scale = 1.f;
  p.out = Y, p.ldo = out_f, p.residual = residual, p.act = act;
  p.out_op = Yop, p.out_op_kp = Yop_kp, p.out_op_split = o.split ? 1 : 0;
  // skinny products (few output columns): 128-column tiles put twice as many SMs to work
  const int bn = store_tile_width(rows, out_f);
  const tc::Plan pl = tc::plan_tiles(p, 64, 1, 1, bn);
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(Xop, rows, o.Kp, o.Kp, tc::BM, &tmA));
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(prepared, out_f, o.Kp, o.Kp, pl.bn, &tmB));
  return tc::launch_sim_tc(tc::EPI_STORE, p.num_kb <= 8, pl, tmA, tmB, p, s);
}

struct CamWs {
  float *X0, *X1, *QKV, *res;
  __nv_bfloat16 *Hop, *Fop;
};
CamWs carve_cam(Workspace& ws, int L, int64_t b, int D, int precision) {
  const int64_t rows = (int64_t)L * b;
  const OperandPlan od = plan_operands(D, VTC_F32, precision);
  const OperandPlan of = plan_operands(4 * D, VTC_F32, precision);
  CamWs w;
  w.X0 = ws.take<float>((size_t)rows * D);
  w.X1 = ws.take<float>((size_t)rows * D);
  w.QKV = ws.take<float>((size_t)rows * 3 * D);
  w.res = ws.take<float>((size_t)b * D);
  w.Hop = ws.take<__nv_bfloat16>((size_t)rows * od.Kp);
  w.Fop = ws.take<__nv_bfloat16>((size_t)rows * of.Kp);
  return w;
}

// ---- CAM backward in one call (autograd through _adapt_feature, model/model.py:141-205 under
// loss.backward(), trainer/trainer.py:79).  Round 1 drove ~156 launches from Python; here the same
// chain is enqueued from C: per layer 8 tensor-core products (dX = dY W through the prepared
// TRANSPOSED weights, dW = dY^T X on transposed operand copies), the fp32 row kernels of cam_bwd.cu
// between them.  Operand preparation is the only extra traffic: every product reads bf16 operands.
struct CamBwdWs {
  float *dXa, *dXb, *dA, *dQKV, *dF, *dU, *tX, *tY, *zb, *dres, *dT;
  __nv_bfloat16 *opX, *opT1, *opT2;
};
CamBwdWs carve_cam_bwd(Workspace& ws, int L, int64_t b, int D, int precision) {
  const int64_t rows = (int64_t)L * b;
  const OperandPlan of = plan_operands(4 * D, VTC_F32, precision);
  const OperandPlan orow = plan_operands((int)rows, VTC_F32, precision);
  CamBwdWs w;
  w.dXa = ws.take<float>((size_t)rows * D);
  w.dXb = ws.take<float>((size_t)rows * D);
  w.dA = ws.take<float>((size_t)rows * D);
  w.dQKV = ws.take<float>((size_t)rows * 3 * D);
  w.dF = ws.take<float>((size_t)rows * 4 * D);
  w.dU = ws.take<float>((size_t)rows * 4 * D);
  w.tX = ws.take<float>((size_t)rows * 4 * D);
  w.tY = ws.take<float>((size_t)rows * 4 * D);
  w.zb = ws.take<float>(round_up<int64_t>(4 * D, tc::BN));
  w.dres = ws.take<float>((size_t)b * D);
  w.dT = ws.take<float>((size_t)rows * D);
  w.opX = ws.take<__nv_bfloat16>((size_t)rows * of.Kp);
  w.opT1 = ws.take<__nv_bfloat16>((size_t)4 * D * orow.Kp);
  w.opT2 = ws.take<__nv_bfloat16>((size_t)4 * D * orow.Kp);
  return w;
}

// out[N, M] = X[N, K] Y[M, K]^T from fp32 row-major matrices: operands prepared here (X as the
// query side, Y as the gallery side), no bias
int gemm_f32(const float* X, const float* Y, int64_t N, int64_t M, int K, int precision,
             __nv_bfloat16* opX, __nv_bfloat16* opY, const float* zero_bias, float* out,
             cudaStream_t s) {
  const OperandPlan o = plan_operands(K, VTC_F32, precision);
  VTC_RETURN_IF_ERROR(launch_prep_operand(X, false, N, K, K, o.split ? PREP_SPLIT_A : PREP_PLAIN, opX, o.Kp, s));
  VTC_RETURN_IF_ERROR(launch_prep_operand(Y, false, M, K, K, o.split ? PREP_SPLIT_B : PREP_PLAIN, opY, o.Kp, s));
  tc::Params p;
  memset(&p, 0, sizeof(p));
  p.N = N, p.M = M, p.num_kb = o.Kp / tc::BK;
  p.col_bias = zero_bias, p.scale = 1.f, p.out = out, p.ldo = M;
  const tc::Plan pl = tc::plan_tiles(p, 64, 1, 1, store_tile_width(N, M));
  CUtensorMap tmA, tmB;
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opX, N, o.Kp, o.Kp, tc::BM, &tmA));
  VTC_RETURN_IF_ERROR(tc::make_operand_tmap(opY, M, o.Kp, o.Kp, pl.bn, &tmB));
  return tc::launch_sim_tc(tc::EPI_STORE, false, pl, tmA, tmB, p, s);
}

}  // namespace
int launch_transpose(const float* in, int64_t R, int64_t C, float* out, cudaStream_t s);
int launch_gelu_bwd(const float* dF, const float* U, int64_t n, float* dU, cudaStream_t s);
int launch_colsum(const float* X, int64_t R, int64_t C, float* out, cudaStream_t s);
int launch_layernorm_bwd(const float* dY, const float* X, const float* gamma, int64_t rows, int D,
                         float eps, const float* dres, float* dX, float* dgamma, float* dbeta,
                         cudaStream_t s);
int launch_cam_attn_core_bwd(const float* QKV, const float* dO, int L, int64_t b, int D, int heads,
                             float* dQKV, cudaStream_t s);
int launch_cam_stack_normalize_bwd(const float* main, const float* aux, const float* dX, int L,
                                   int64_t b, int D, float* dmain, float* daux, cudaStream_t s);
int launch_cam_readout_bwd(const float* T, const float* main, const float* res_in,
                           const uint8_t* skip_mask, const float* dout, int L, int64_t b, int D,
                           int mode, int res_act, float res_scale, const float* res_shift,
                           const float* res_mul, float* dT, float* dres, float* dmain,
                           cudaStream_t s);
namespace {

// dX = dY W (through the prepared transposed weight) and dW = dY^T Act, db = colsum(dY)
int linear_bwd(const float* dY, const float* Act, const void* Wt_prepared, int64_t rows, int in_f,
               int out_f, int precision, const CamBwdWs& w, float* dX, float* dW, float* db,
               cudaStream_t s) {
  const OperandPlan oo = plan_operands(out_f, VTC_F32, precision);
  // dX [rows, in_f] = dY [rows, out_f] . (W^T)[in_f, out_f]^T
  VTC_RETURN_IF_ERROR(launch_prep_operand(dY, false, rows, out_f, out_f,
                                          oo.split ? PREP_SPLIT_A : PREP_PLAIN, w.opX, oo.Kp, s));
  VTC_RETURN_IF_ERROR(linear_prepared(w.opX, Wt_prepared, nullptr, rows, out_f, in_f, 0, precision,
                                      dX, nullptr, 0, s));
  // dW [out_f, in_f] = (dY^T)[out_f, rows] . (Act^T)[in_f, rows]^T
  VTC_RETURN_IF_ERROR(launch_transpose(dY, rows, out_f, w.tX, s));
  VTC_RETURN_IF_ERROR(launch_transpose(Act, rows, in_f, w.tY, s));
  VTC_RETURN_IF_ERROR(gemm_f32(w.tX, w.tY, out_f, in_f, (int)rows, precision, w.opT1, w.opT2, w.zb,
                               dW, s));
  return launch_colsum(dY, rows, out_f, db, s);
}

int cam_backward_impl(const float* dout, const float* main, const float* aux, const float* T,
                      const float* res_in, const uint8_t* skip_mask, int L, int64_t b, int D,
                      int heads, int layers, const vtc_cam_layer_bwd* lp, int readout_mode,
                      const void* flw_t, int res_act, float res_scale, const float* res_shift,
                      const float* res_mul, int precision, float* dmain, float* daux, float* dflw,
                      void* wsp, size_t ws_bytes, cudaStream_t s) {
  Workspace ws(wsp, ws_bytes);
  CamBwdWs w = carve_cam_bwd(ws, L, b, D, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const int64_t rows = (int64_t)L * b;
  VTC_RETURN_IF_ERROR(launch_fill_bias(w.zb, nullptr, 4 * D, round_up<int64_t>(4 * D, tc::BN), 0.f, s));
  // read-out (model/model.py:156-161, 168-171, 199-203): dmain first receives the part through
  // normalize(main) at :203; the stacked-input part is added at the end
  float* dX = w.dXa;
  if (readout_mode == VTC_CAM_READOUT_AVG) {
    VTC_RETURN_IF_ERROR(launch_cam_readout_bwd(T, main, nullptr, skip_mask, dout, L, b, D,
                                               VTC_CAM_READOUT_AVG, res_act, res_scale, res_shift,
                                               res_mul, dX, nullptr, dmain, s));
  } else {
    VTC_RETURN_IF_ERROR(launch_cam_readout_bwd(nullptr, main, res_in, skip_mask, dout, L, b, D,
                                               VTC_CAM_READOUT_RESIDUAL_ONLY, res_act, res_scale,
                                               res_shift, res_mul, nullptr, w.dres, dmain, s));
    // final_linear(token 0) (model/model.py:161): d token0 = dres W, dW = dres^T token0
    cudaError_t e = cudaMemsetAsync(dX, 0, (size_t)rows * D * sizeof(float), s);
    if (e != cudaSuccess) return cuda_err(e);
    const OperandPlan od = plan_operands(D, VTC_F32, precision);
    VTC_RETURN_IF_ERROR(launch_prep_operand(w.dres, false, b, D, D,
                                            od.split ? PREP_SPLIT_A : PREP_PLAIN, w.opX, od.Kp, s));
    VTC_RETURN_IF_ERROR(linear_prepared(w.opX, flw_t, nullptr, b, D, D, 0, precision, dX, nullptr, 0, s));
    if (dflw) {
      VTC_RETURN_IF_ERROR(launch_transpose(w.dres, b, D, w.tX, s));
      VTC_RETURN_IF_ERROR(launch_transpose(T, b, D, w.tY, s));
      VTC_RETURN_IF_ERROR(gemm_f32(w.tX, w.tY, D, D, (int)b, precision, w.opT1, w.opT2, w.zb, dflw, s));
    }
  }
  float* dX2 = w.dXb;
  for (int i = layers - 1; i >= 0; --i) {  // timesformer_clip_alt.py:112-124, backwards
    const vtc_cam_layer_bwd& l = lp[i];
    // x = x2 + c_proj(gelu(c_fc(ln_2(x2))))
    VTC_RETURN_IF_ERROR(linear_bwd(dX, l.Fa, l.proj_t, rows, 4 * D, D, precision, w, w.dF, l.dWpr, l.dbpr, s));
    VTC_RETURN_IF_ERROR(launch_gelu_bwd(w.dF, l.U, rows * 4 * D, w.dU, s));
    VTC_RETURN_IF_ERROR(linear_bwd(w.dU, l.H2, l.fc_t, rows, D, 4 * D, precision, w, w.dA, l.dWfc, l.dbfc, s));
    VTC_RETURN_IF_ERROR(launch_layernorm_bwd(w.dA, l.X2, l.ln2_g, rows, D, 1e-5f, dX, dX2, l.dg2, l.db2, s));
    // x2 = x + out_proj(attn(in_proj(ln_1(x))))
    VTC_RETURN_IF_ERROR(linear_bwd(dX2, l.A, l.out_t, rows, D, D, precision, w, w.dA, l.dWo, l.dbo, s));
    VTC_RETURN_IF_ERROR(launch_cam_attn_core_bwd(l.QKV, w.dA, L, b, D, heads, w.dQKV, s));
    VTC_RETURN_IF_ERROR(linear_bwd(w.dQKV, l.H1, l.qkv_t, rows, D, 3 * D, precision, w, w.dA, l.dWqkv, l.dbqkv, s));
    VTC_RETURN_IF_ERROR(launch_layernorm_bwd(w.dA, l.X, l.ln1_g, rows, D, 1e-5f, dX2, dX, l.dg1, l.db1, s));
  }
  // X = normalize(stack([main, *aux])) (model/model.py:150-151): dmain += the token-0 part
  VTC_RETURN_IF_ERROR(launch_cam_stack_normalize_bwd(main, aux, dX, L, b, D, w.dres, daux, s));
  return launch_bias_act(dmain, nullptr, w.dres, b, D, 0, dmain, s);
}

}  // namespace
}  // namespace vtc

using namespace vtc;

extern "C" {

int vtc_abi_version(void) { return VTC_ABI_VERSION; }

const char* vtc_strerror(int code) {
  switch (code) {
    case VTC_OK: return "ok";
    case VTC_ERR_INVALID_ARG: return "invalid argument";
    case VTC_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
    case VTC_ERR_WORKSPACE: return "workspace missing or too small (see vtc_workspace_bytes)";
    case VTC_ERR_NO_DEVICE: return "no CUDA device";
    case VTC_ERR_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
    default: break;
  }
  if (code <= VTC_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(VTC_ERR_CUDA_BASE - code));
  return "unknown error";
}

uint64_t vtc_launch_count(void) { return g_launch_count.load(); }

int vtc_kernel_timer_enable(int on) {
  tc::kernel_timer_enable(on != 0);
  return VTC_OK;
}

int vtc_kernel_timer_read(double* total_ms, int* count) {
  return tc::kernel_timer_read(total_ms, count);
}

int vtc_debug_prof_read(unsigned long long* out, int max_words) {
  return tc::debug_prof_read(out, max_words);
}

int vtc_trace_begin(vtc_stream_t stream) {
  std::lock_guard<std::mutex> lk(g_trace.mu);
  for (cudaEvent_t e : g_trace.events) cudaEventDestroy(e);
  g_trace.events.clear();
  g_trace.where.clear();
  g_trace.stream = (cudaStream_t)stream;
  cudaEvent_t e = nullptr;
  cudaError_t rc = cudaEventCreate(&e);
  if (rc == cudaSuccess) rc = cudaEventRecord(e, g_trace.stream);
  if (rc != cudaSuccess) return cuda_err(rc);
  g_trace.events.push_back(e);
  g_trace_on.store(1);
  return VTC_OK;
}

int vtc_trace_end(char* buf, size_t cap) {
  g_trace_on.store(0);
  std::lock_guard<std::mutex> lk(g_trace.mu);
  size_t off = 0;
  int n = 0;
  cudaError_t rc = cudaSuccess;
  if (!g_trace.events.empty()) rc = cudaEventSynchronize(g_trace.events.back());
  for (size_t i = 1; i < g_trace.events.size() && rc == cudaSuccess; ++i) {
    float ms = 0.f;
    rc = cudaEventElapsedTime(&ms, g_trace.events[i - 1], g_trace.events[i]);
    const char* f = g_trace.where[i - 1].first;
    const char* slash = strrchr(f, '/');
    if (buf && off < cap) {
      const int w = snprintf(buf + off, cap - off, "%s:%d %.3f\n", slash ? slash + 1 : f,
                             g_trace.where[i - 1].second, ms * 1e3f);
      if (w > 0) off += (size_t)w;
    }
    ++n;
  }
  for (cudaEvent_t e : g_trace.events) cudaEventDestroy(e);
  g_trace.events.clear();
  g_trace.where.clear();
  if (buf && cap) buf[off < cap ? off : cap - 1] = 0;
  return rc == cudaSuccess ? n : cuda_err(rc);
}

size_t vtc_workspace_bytes(int op, int64_t N, int64_t M, int D, int precision) {
  if (N < 0 || M < 0 || D <= 0 || !valid_prec(precision)) return 0;
  Workspace ws(nullptr, 0);
  // sizes do not depend on the storage dtype except through the split; assume fp32 inputs (worst)
  switch (op) {
    case VTC_OP_SIM_RANK: carve_rank(ws, N, M, D, VTC_F32, precision, true); break;
    case VTC_OP_SIM_TOPK: carve_topk(ws, N, M, D, VTC_F32, precision); break;
    case VTC_OP_INFONCE_FWD: {
      carve_nce(ws, N, D, VTC_F32, precision == VTC_PREC_BRUTE ? VTC_PREC_EXACT : precision);
      const size_t small = N <= kNceSmallMax ? infonce_small_ws_bytes(N) : 0;
      if (small > ws.used) ws.used = small;
      break;
    }
    case VTC_OP_SIM_MATRIX:
      carve_gemm(ws, N, M, D, VTC_F32, precision == VTC_PREC_BRUTE ? VTC_PREC_EXACT : precision);
      break;
    case VTC_OP_LINEAR:
      carve_gemm(ws, N, M, D, VTC_F32, precision == VTC_PREC_BRUTE ? VTC_PREC_EXACT : precision);
      break;
    case VTC_OP_INFONCE_BWD: {
      carve_nce_bwd(ws, N, D, precision == VTC_PREC_BRUTE ? VTC_PREC_EXACT : precision);
      const size_t small = N <= kNceSmallMax ? (size_t)N * N * sizeof(float) + 256 : 0;
      if (small > ws.used) ws.used = small;
      break;
    }
    case VTC_OP_GT_SCORES:
      ws.take<double>(M);
      ws.take<__nv_bfloat16>((size_t)N * round_up(D, tc::BK));
      ws.take<__nv_bfloat16>((size_t)M * round_up(D, tc::BK));
      break;
    default: return 0;
  }
  return ws.used + 512;
}

int vtc_row_norms(const void* X, int64_t rows, int D, int64_t ldx, int dtype, float* inv_norm,
                  float* sq_norm, vtc_stream_t stream) {
  if (!X || rows < 0 || D <= 0 || ldx < D || !valid_dtype(dtype)) return VTC_ERR_INVALID_ARG;
  return launch_row_norms(X, dtype == VTC_BF16, rows, D, ldx, inv_norm, sq_norm,
                          (cudaStream_t)stream);
}

int vtc_normalize(const void* X, int64_t rows, int D, int64_t ldx, int dtype, void* Y, int64_t ldy,
                  vtc_stream_t stream) {
  if (!X || !Y || rows < 0 || D <= 0 || ldx < D || ldy < D || !valid_dtype(dtype))
    return VTC_ERR_INVALID_ARG;
  return launch_normalize(X, dtype == VTC_BF16, rows, D, ldx, Y, ldy, (cudaStream_t)stream);
}

int vtc_sim_matrix(const void* A, const void* B, int64_t N, int64_t M, int D, int dtype,
                   int precision, const float* scale, float* out, int64_t ldo, void* ws,
                   size_t ws_bytes, vtc_stream_t stream) {
  return gemm_store_impl(A, B, N, M, D, dtype, precision, scale, nullptr, nullptr, 0, out, ldo, ws,
                         ws_bytes, (cudaStream_t)stream);
}

int vtc_linear(const float* X, const float* W, const float* bias, const float* residual,
               int64_t rows, int in_f, int out_f, int act, int precision, float* Y, void* ws,
               size_t ws_bytes, vtc_stream_t stream) {
  if (act != 0 && act != 1) return VTC_ERR_INVALID_ARG;
  return gemm_store_impl(X, W, rows, out_f, in_f, VTC_F32, precision, nullptr, bias, residual, act,
                         Y, out_f, ws, ws_bytes, (cudaStream_t)stream);
}

int vtc_sim_rank(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                 const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                 int precision, const double* gt_score, double* gt_score_out, int accumulate,
                 int32_t* rank0, void* ws, size_t ws_bytes, vtc_stream_t stream) {
  return sim_rank_impl(Q, G, N, M, D, dtype, gt, row_offset, col_offset, metric, precision,
                       gt_score, gt_score_out, accumulate, rank0, ws, ws_bytes,
                       (cudaStream_t)stream);
}

int vtc_rank_eval(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                  const int64_t* gt, int metric, int precision, const int* k_vals, int nk,
                  int32_t* rank0, int64_t* hits, double* medr, double* gt_score_out, void* ws,
                  size_t ws_bytes, vtc_stream_t stream) {
  if (nk < 0 || nk > 8 || (nk > 0 && (!k_vals || !hits))) return VTC_ERR_INVALID_ARG;
  RankFinalize fin{M, k_vals, nk, hits, medr};
  return sim_rank_impl(Q, G, N, M, D, dtype, gt, 0, 0, metric, precision, nullptr, gt_score_out, 0,
                       rank0, ws, ws_bytes, (cudaStream_t)stream, nullptr, &fin);
}

int vtc_rank_prepare(const void* X, int64_t rows, int D, int dtype, int precision, double* sq64,
                     float* qq_up, vtc_stream_t stream) {
  if (!X || rows < 0 || D <= 0 || !valid_dtype(dtype) || !valid_prec(precision) || (!sq64 && !qq_up))
    return VTC_ERR_INVALID_ARG;
  // the canonical values must be the rows as handed in: bf16 rows in the bf16 mode, fp32 rows
  // otherwise (fp32 rows in the bf16 mode would have to be rounded first -- hand in the rounded rows)
  if ((precision == VTC_PREC_BF16) != (dtype == VTC_BF16)) return VTC_ERR_UNSUPPORTED_SHAPE;
  // one staged pass over the rows (rank_stage.cu): as gallery rows for the canonical norms, as
  // query rows for the norm bounds
  RankPrologueArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.in_bf16 = dtype == VTC_BF16 ? 1 : 0, pa.D = D, pa.ldq = D, pa.ldg = D;
  pa.mode_q = STAGE_NONE, pa.mode_g = STAGE_NONE, pa.metric = VTC_METRIC_L2;
  if (sq64) pa.G = X, pa.M = rows, pa.Mpad = rows, pa.sq64 = sq64;
  if (qq_up) pa.Q = X, pa.N = rows, pa.qq = qq_up;
  return launch_rank_prologue(pa, (cudaStream_t)stream);
}

int vtc_sim_rank_prepared(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                          const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                          int precision, const double* gt_score, double* gt_score_out,
                          const double* sq64, double* sq64_out, const float* qq_up, float* qq_out,
                          int accumulate, int32_t* rank0, void* ws, size_t ws_bytes,
                          vtc_stream_t stream) {
  if ((!sq64 && !sq64_out) || (!qq_up && !qq_out) || (!gt_score && !gt_score_out))
    return VTC_ERR_INVALID_ARG;
  if (precision == VTC_PREC_BRUTE) return VTC_ERR_UNSUPPORTED_SHAPE;
  if (N > 0 && M == 0) {  // an empty gallery chunk adds nothing
    if (accumulate) return VTC_OK;
    return sim_rank_impl(Q, G, N, M, D, dtype, gt, row_offset, col_offset, metric, precision,
                         gt_score, gt_score_out, accumulate, rank0, ws, ws_bytes,
                         (cudaStream_t)stream);
  }
  const RankCache cache{sq64, sq64_out, qq_up, qq_out};
  return sim_rank_impl(Q, G, N, M, D, dtype, gt, row_offset, col_offset, metric, precision,
                       gt_score, gt_score_out, accumulate, rank0, ws, ws_bytes, (cudaStream_t)stream,
                       &cache);
}

int vtc_gt_scores(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype,
                  const int64_t* gt, int64_t row_offset, int64_t col_offset, int metric,
                  int precision, double* gt_score, void* wsp, size_t ws_bytes,
                  vtc_stream_t stream) {
  if (N < 0 || M < 0 || D <= 0 || !valid_dtype(dtype) || !valid_metric(metric) ||
      !valid_prec(precision))
    return VTC_ERR_INVALID_ARG;
  if (N == 0) return VTC_OK;  // no queries (pointers may be NULL, e.g. an empty shard)
  if (!Q || (!G && M > 0) || !gt_score) return VTC_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  Workspace ws(wsp, ws_bytes);
  double* sq64 = ws.take<double>(M);
  const bool in_bf16 = dtype == VTC_BF16;
  // VTC_PREC_BF16 ranks the bf16 roundings of fp32 inputs: round them first
  __nv_bfloat16 *opQ = nullptr, *opG = nullptr;
  const bool round_first = precision == VTC_PREC_BF16 && !in_bf16;
  const int Kp = round_up(D, tc::BK);
  if (round_first) {
    opQ = ws.take<__nv_bfloat16>((size_t)N * Kp);
    opG = ws.take<__nv_bfloat16>((size_t)M * Kp);
  }
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  ExactArgs ex;
  if (round_first) {
    VTC_RETURN_IF_ERROR(launch_prep_operand(Q, false, N, D, D, PREP_PLAIN, opQ, Kp, s));
    VTC_RETURN_IF_ERROR(launch_prep_operand(G, false, M, D, D, PREP_PLAIN, opG, Kp, s));
    ex = ExactArgs{opQ, opG, Kp, Kp, true, N, M, D, sq64, gt, row_offset, col_offset, metric};
  } else {
    ex = ExactArgs{Q, G, D, D, in_bf16, N, M, D, sq64, gt, row_offset, col_offset, metric};
  }
  if (metric == VTC_METRIC_L2 && M > 0)
    VTC_RETURN_IF_ERROR(launch_sqnorm64(ex.G, ex.bf16, M, D, ex.ldg, sq64, nullptr, nullptr, s));
  return launch_gt_score(ex, nullptr, gt_score, nullptr, nullptr, 0.f, s);
}

int vtc_rank_finalize(int32_t* rank0, const double* gt_score, int64_t N, int64_t M_total,
                      const int* k_vals, int nk, int64_t* hits, double* medr, void* hist_ws,
                      size_t hist_ws_bytes, vtc_stream_t stream) {
  if (!rank0 || N < 0 || nk < 0 || nk > 8 || (nk > 0 && (!k_vals || !hits)))
    return VTC_ERR_INVALID_ARG;
  if (medr && (!hist_ws || hist_ws_bytes < (3 * 65536 + 8) * sizeof(unsigned int)))
    return VTC_ERR_WORKSPACE;
  if (N == 0 || !hist_ws)  // no scratch: the stand-alone kernels (hit counts only)
    return launch_rank_finalize(rank0, gt_score, N, M_total, k_vals, nk, hits, medr, hist_ws,
                                (cudaStream_t)stream);
  // one block: NaN ground truth -> M_total, hit counts, radix-select median
  RankEpilogueArgs ea;
  memset(&ea, 0, sizeof(ea));
  ea.ex.N = N, ea.ex.bf16 = false;
  ea.dgt = gt_score, ea.rank0 = rank0, ea.accumulate = 1;
  ea.finalize = 1, ea.M_total = M_total, ea.nk = nk;
  for (int i = 0; i < nk; ++i) ea.k_vals[i] = k_vals[i];
  ea.hits = reinterpret_cast<unsigned long long*>(hits), ea.medr = medr;
  ea.hist = static_cast<unsigned int*>(hist_ws);
  return launch_rank_epilogue(ea, (cudaStream_t)stream);
}

int vtc_sim_topk(const void* Q, const void* G, int64_t N, int64_t M, int D, int dtype, int metric,
                 int precision, int k, int64_t col_offset, float* out_val, int64_t* out_idx,
                 void* ws, size_t ws_bytes, vtc_stream_t stream) {
  return sim_topk_impl(Q, G, N, M, D, dtype, metric, precision, k, col_offset, out_val, out_idx, ws,
                       ws_bytes, (cudaStream_t)stream);
}

int vtc_topk_merge(const float* vals, const int64_t* idx, int parts, int64_t N, int k,
                   float* out_val, int64_t* out_idx, vtc_stream_t stream) {
  if (!vals || !idx || !out_val || !out_idx || N < 0) return VTC_ERR_INVALID_ARG;
  return launch_topk_merge(vals, idx, parts, N, k, out_val, out_idx, (cudaStream_t)stream);
}

int vtc_infonce_fwd(const void* A, const void* B, int64_t n, int D, int dtype, int precision,
                    const float* scale, float* loss, float* row_lse, float* col_lse, float* diag,
                    void* ws, size_t ws_bytes, vtc_stream_t stream) {
  return infonce_fwd_impl(A, B, n, D, dtype, precision, scale, loss, row_lse, col_lse, diag, ws,
                          ws_bytes, (cudaStream_t)stream);
}

int vtc_cam_stack_normalize(const float* main, const float* aux, int L, int64_t b, int D, float* X,
                            vtc_stream_t stream) {
  if (!main || !X || L < 1 || (L > 1 && !aux) || b < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_cam_stack_normalize(main, aux, L, b, D, X, (cudaStream_t)stream);
}

int vtc_layernorm(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                  float eps, float* Y, vtc_stream_t stream) {
  if (!X || !gamma || !beta || !Y || rows < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_layernorm(X, gamma, beta, rows, D, eps, Y, (cudaStream_t)stream);
}

int vtc_cam_attn_core(const float* QKV, int L, int64_t b, int D, int heads, float* out,
                      vtc_stream_t stream) {
  if (!QKV || !out || b < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_cam_attn_core(QKV, L, b, D, heads, out, nullptr, 0, 0, (cudaStream_t)stream);
}

int vtc_bias_act(const float* X, const float* bias, const float* residual, int64_t rows, int D,
                 int act, float* Y, vtc_stream_t stream) {
  if (!X || !Y || rows < 0 || D <= 0 || (act != 0 && act != 1)) return VTC_ERR_INVALID_ARG;
  return launch_bias_act(X, bias, residual, rows, D, act, Y, (cudaStream_t)stream);
}

int vtc_cam_readout(const float* T, const float* main, const float* res_in,
                    const uint8_t* skip_mask, int L, int64_t b, int D, int mode, int res_act,
                    float res_scale, const float* res_shift, const float* res_mul, float* out,
                    vtc_stream_t stream) {
  if (!out || b < 0 || D <= 0 || L < 1) return VTC_ERR_INVALID_ARG;
  if (mode == VTC_CAM_READOUT_RESIDUAL_ONLY ? (!res_in || !main)
                                            : (!T || (mode == VTC_CAM_READOUT_AVG && !main)))
    return VTC_ERR_INVALID_ARG;
  if (mode < 0 || mode > 2) return VTC_ERR_INVALID_ARG;
  return launch_cam_readout(T, main, res_in, skip_mask, L, b, D, mode, res_act, res_scale, res_shift,
                            res_mul, out, (cudaStream_t)stream);
}

size_t vtc_linear_prepared_bytes(int in_f, int out_f, int precision) {
  if (in_f <= 0 || out_f <= 0 || (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16)) return 0;
  return prepared_linear_bytes(in_f, out_f, precision);
}

int vtc_linear_prepare(const float* W, const float* bias, int in_f, int out_f, int precision,
                       void* prepared, vtc_stream_t stream) {
  if (!W || !prepared || in_f <= 0 || out_f <= 0 ||
      (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16))
    return VTC_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const OperandPlan o = plan_operands(in_f, VTC_F32, precision);
  VTC_RETURN_IF_ERROR(launch_prep_operand(W, false, out_f, in_f, in_f,
                                          o.split ? PREP_SPLIT_B : PREP_PLAIN,
                                          (__nv_bfloat16*)prepared, o.Kp, s));
  return launch_fill_bias((float*)prepared_bias(prepared, in_f, out_f, precision), bias, out_f,
                          round_up<int64_t>(out_f, tc::BN), 0.f, s);
}

size_t vtc_cam_workspace_bytes(int L, int64_t b, int D, int precision) {
  if (L < 1 || b < 0 || D <= 0 || (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16)) return 0;
  Workspace ws(nullptr, 0);
  carve_cam(ws, L, b, D, precision);
  return ws.used + 512;
}

int vtc_cam_forward(const float* main, const float* aux, int L, int64_t b, int D, int heads,
                    int layers, const vtc_cam_layer* lp, int readout_mode, const void* final_linear,
                    const uint8_t* skip_mask, int res_act, float res_scale, const float* res_shift,
                    const float* res_mul, int precision, float* out, void* wsp, size_t ws_bytes,
                    vtc_stream_t stream) {
  if (!main || !out || L < 1 || (L > 1 && !aux) || b < 0 || D <= 0 || heads < 1 || layers < 0 ||
      (layers > 0 && !lp) || (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16) ||
      (readout_mode != VTC_CAM_READOUT_AVG && readout_mode != VTC_CAM_READOUT_RESIDUAL_ONLY) ||
      (readout_mode == VTC_CAM_READOUT_RESIDUAL_ONLY && !final_linear))
    return VTC_ERR_INVALID_ARG;
  if (b == 0) return VTC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Workspace ws(wsp, ws_bytes);
  CamWs w = carve_cam(ws, L, b, D, precision);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  const int64_t rows = (int64_t)L * b;
  const OperandPlan od = plan_operands(D, VTC_F32, precision);
  const OperandPlan of = plan_operands(4 * D, VTC_F32, precision);
  const int split = od.split ? 1 : 0;
  if (of.Kp != (split ? 3 : 1) * 4 * D) {  // padding columns of the MLP operand must be zero
    cudaError_t e = cudaMemsetAsync(w.Fop, 0, (size_t)rows * of.Kp * sizeof(__nv_bfloat16), s);
    if (e != cudaSuccess) return cuda_err(e);
  }
  VTC_RETURN_IF_ERROR(launch_cam_stack_normalize(main, aux, L, b, D, w.X0, s));  // model.py:150-151
  float *X = w.X0, *X2 = w.X1;
  for (int i = 0; i < layers; ++i) {  // clip.model.Transformer block (timesformer_clip_alt.py:112-124)
    const vtc_cam_layer& l = lp[i];
    VTC_RETURN_IF_ERROR(launch_layernorm_prep(X, l.ln1_g, l.ln1_b, rows, D, 1e-5f, split, w.Hop, od.Kp, s));
    VTC_RETURN_IF_ERROR(linear_prepared(w.Hop, l.qkv, nullptr, rows, D, 3 * D, 0, precision, w.QKV,
                                        nullptr, 0, s));
    VTC_RETURN_IF_ERROR(launch_cam_attn_core(w.QKV, L, b, D, heads, nullptr, w.Hop, od.Kp, split, s));
    VTC_RETURN_IF_ERROR(linear_prepared(w.Hop, l.out, X, rows, D, D, 0, precision, X2, nullptr, 0, s));
    VTC_RETURN_IF_ERROR(launch_layernorm_prep(X2, l.ln2_g, l.ln2_b, rows, D, 1e-5f, split, w.Hop, od.Kp, s));
    VTC_RETURN_IF_ERROR(linear_prepared(w.Hop, l.fc, nullptr, rows, D, 4 * D, 1, precision, nullptr,
                                        w.Fop, of.Kp, s));
    VTC_RETURN_IF_ERROR(linear_prepared(w.Fop, l.proj, X2, rows, 4 * D, D, 0, precision, X, nullptr, 0, s));
  }
  if (readout_mode == VTC_CAM_READOUT_AVG)
    return launch_cam_readout(X, main, nullptr, skip_mask, L, b, D, VTC_CAM_READOUT_AVG, res_act,
                              res_scale, res_shift, res_mul, out, s);
  // final_linear(token 0)  (model/model.py:161): token 0 = the first b rows of X
  VTC_RETURN_IF_ERROR(launch_prep_operand(X, false, b, D, D, split ? PREP_SPLIT_A : PREP_PLAIN, w.Hop,
                                          od.Kp, s));
  VTC_RETURN_IF_ERROR(linear_prepared(w.Hop, final_linear, nullptr, b, D, D, 0, precision, w.res,
                                      nullptr, 0, s));
  return launch_cam_readout(nullptr, main, w.res, skip_mask, L, b, D, VTC_CAM_READOUT_RESIDUAL_ONLY,
                            res_act, res_scale, res_shift, res_mul, out, s);
}

size_t vtc_cam_backward_workspace_bytes(int L, int64_t b, int D, int precision) {
  if (L < 1 || b < 0 || D <= 0 || (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16)) return 0;
  Workspace ws(nullptr, 0);
  carve_cam_bwd(ws, L, b, D, precision);
  return ws.used + 512;
}

int vtc_cam_backward(const float* dout, const float* main, const float* aux, const float* T,
                     const float* res_in, const uint8_t* skip_mask, int L, int64_t b, int D,
                     int heads, int layers, const vtc_cam_layer_bwd* lp, int readout_mode,
                     const void* final_linear_t, int res_act, float res_scale,
                     const float* res_shift, const float* res_mul, int precision, float* dmain,
                     float* daux, float* dflw, void* ws, size_t ws_bytes, vtc_stream_t stream) {
  if (!dout || !main || !T || !dmain || L < 1 || (L > 1 && (!aux || !daux)) || b < 0 || D <= 0 ||
      heads < 1 || layers < 0 || (layers > 0 && !lp) ||
      (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16) ||
      (readout_mode != VTC_CAM_READOUT_AVG && readout_mode != VTC_CAM_READOUT_RESIDUAL_ONLY) ||
      (readout_mode == VTC_CAM_READOUT_RESIDUAL_ONLY && (!final_linear_t || !res_in)))
    return VTC_ERR_INVALID_ARG;
  if (b == 0) return VTC_OK;
  return cam_backward_impl(dout, main, aux, T, res_in, skip_mask, L, b, D, heads, layers, lp,
                           readout_mode, final_linear_t, res_act, res_scale, res_shift, res_mul,
                           precision, dmain, daux, dflw, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
