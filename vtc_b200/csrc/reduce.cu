// reduce.cu -- reductions that follow the fused similarity GEMM.
//   * InfoNCE (model/loss.py:18-22): merge the per-split online log-sum-exp partials and form
//     0.5 * (CE(sim, arange) + CE(sim.t(), arange)) without ever seeing `sim`;
//   * exact top-k (model/metric.py:144-146 `search(b, k)`): re-score the streamed candidate pools
//     in fp64-sequential arithmetic, select the k best by (score, index), prove completeness
//     against the guard band or flag the row for the brute-force kernel;
//   * k-way merge of per-shard top-k lists (gallery-sharded multi-GPU search).
#include "reduce.cuh"

namespace vtc {

// ------------------------------------------------------------------------------------- InfoNCE
// run_flag (nullable): the launch does nothing unless *run_flag != 0 (fallback pass)
__global__ void lse_merge_kernel(const float2* __restrict__ part, int splits, int64_t n,
                                 float* __restrict__ lse, const unsigned int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0u) return;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, part[(int64_t)s * n + t].x);
  float l = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 p = part[(int64_t)s * n + t];
    if (p.x > -INFINITY) l += p.y * exp2f(p.x - m);
  }
  lse[t] = 0.6931471805599453f * (m + log2f(l));
}

// Reference of column j for the one-pass column sums: its own positive logit in log2 units,
// ref[j] = scale * log2(e) * <opA_j, opB_j> over the bf16 operand rows (plain, or the 3-term split
// [hi|hi|lo].[hi|lo|hi]); +inf on the padding columns j >= n.  One warp per row.
__global__ void __launch_bounds__(256)
nce_colref_kernel(const __nv_bfloat16* __restrict__ opA, const __nv_bfloat16* __restrict__ opB,
                  int64_t n, int64_t npad, int Kp, const float* __restrict__ scale_ptr,
                  float* __restrict__ ref) {
  const int64_t j = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= npad) return;
  const int lane = threadIdx.x & 31;
  if (j >= n) {
    if (lane == 0) ref[j] = INFINITY;
    return;
  }
  const uint4* a = reinterpret_cast<const uint4*>(opA + j * Kp);
  const uint4* b = reinterpret_cast<const uint4*>(opB + j * Kp);
  float acc = 0.f;
  for (int k = lane; k < Kp / 8; k += 32) {  // Kp % 64 == 0: whole 16-byte groups
    const uint4 ua = __ldg(a + k), ub = __ldg(b + k);
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc = fmaf(__uint_as_float(wa[i] << 16), __uint_as_float(wb[i] << 16), acc);
      acc = fmaf(__uint_as_float(wa[i] & 0xffff0000u), __uint_as_float(wb[i] & 0xffff0000u), acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) ref[j] = *scale_ptr * 1.4426950408889634f * acc;
}

// col_lse[j] = ln2 * (ref[j] + log2(sum over the partial rows of col_part)); a column whose sum is
// not a positive finite number (a negative pair beat the positive one by more than 2^127) raises
// *flag, which gates the second pass on.
__global__ void nce_col_merge_kernel(const float* __restrict__ col_part, int parts, int64_t ld,
                                     const float* __restrict__ ref, int64_t n,
                                     float* __restrict__ col_lse, unsigned int* __restrict__ flag) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // fixed order, four independent chains
  int p = 0;
  for (; p + 4 <= parts; p += 4) {
    s0 += col_part[(int64_t)p * ld + j];
    s1 += col_part[(int64_t)(p + 1) * ld + j];
    s2 += col_part[(int64_t)(p + 2) * ld + j];
    s3 += col_part[(int64_t)(p + 3) * ld + j];
  }
  for (; p < parts; ++p) s0 += col_part[(int64_t)p * ld + j];
  const float s = (s0 + s1) + (s2 + s3);
  if (!(s > 0.f && s < INFINITY)) *flag = 1u;
  col_lse[j] = 0.6931471805599453f * (ref[j] + log2f(s));
}

__global__ void __launch_bounds__(1024)
infonce_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                    const float* __restrict__ diag_raw, const float* __restrict__ scale_ptr,
                    int64_t n, float* __restrict__ diag_out, float* __restrict__ loss) {
  __shared__ double sh[32];
  const float scale = scale_ptr ? *scale_ptr : 1.f;
  double acc = 0.0;
  for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
    const float d = scale * diag_raw[t];
    if (diag_out) diag_out[t] = d;
    acc += ((double)row_lse[t] - (double)d) + ((double)col_lse[t] - (double)d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sh[i];
    *loss = (float)(0.5 * tot / (double)n);
  }
}

int launch_lse_merge(const float2* part, int splits, int64_t n, float* lse, cudaStream_t s,
                     const unsigned int* run_flag) {
  if (n == 0) return VTC_OK;
  lse_merge_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(part, splits, n, lse, run_flag);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_nce_colref(const __nv_bfloat16* opA, const __nv_bfloat16* opB, int64_t n, int64_t npad,
                      int Kp, const float* scale_ptr, float* ref, cudaStream_t s) {
  if (npad == 0) return VTC_OK;
  nce_colref_kernel<<<(unsigned)ceil_div<int64_t>(npad, 8), 256, 0, s>>>(opA, opB, n, npad, Kp,
                                                                        scale_ptr, ref);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_nce_col_merge(const float* col_part, int parts, int64_t ld, const float* ref, int64_t n,
                         float* col_lse, unsigned int* flag, cudaStream_t s) {
  if (n == 0) return VTC_OK;
  nce_col_merge_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(col_part, parts, ld, ref, n,
                                                                          col_lse, flag);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_infonce_loss(const float* row_lse, const float* col_lse, const float* diag_raw,
                        const float* scale_ptr, int64_t n, float* diag_out, float* loss,
                        cudaStream_t s) {
  infonce_loss_kernel<<<1, 1024, 0, s>>>(row_lse, col_lse, diag_raw, scale_ptr, n, diag_out, loss);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

// --------------------------------------------------------------------------------------- top-k
constexpr int SEL_WARPS = 2;
constexpr int SEL_MAX_PARTS = 16;                  // 2 column halves x 8 gallery splits
constexpr int SEL_MAX_CAND = SEL_MAX_PARTS * 64;   // pool = 64 entries per part
constexpr int SEL_MAX_SURV = 256;  // candidates re-scored in fp64 per row; more -> brute-force row

template <typename T>
__device__ __forceinline__ double exact_score(const T* __restrict__ q, const T* __restrict__ x, int D,
                                              int metric, double sq) {
  const double acc = dot_seq64(q, x, D);
  return metric == VTC_METRIC_L2 ? sq - 2.0 * acc : -acc;
}

// lexicographic (d, j) less-than; j < 0 marks "no candidate"
__device__ __forceinline__ bool cand_less(double d1, int j1, double d2, int j2) {
  if (j1 < 0) return false;
  if (j2 < 0) return true;
  return d1 < d2 || (d1 == d2 && j1 < j2);
}

template <typename T>
__global__ void __launch_bounds__(SEL_WARPS * 32)
topk_select_kernel(TopkSelectArgs a) {
  __shared__ double cd[SEL_WARPS][SEL_MAX_SURV];
  __shared__ float ca[SEL_WARPS][SEL_MAX_CAND];
  __shared__ int cj[SEL_WARPS][SEL_MAX_CAND];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * SEL_WARPS + w;
  if (t >= a.ex.N) return;
  const T* Q = (const T*)a.ex.Q;
  const T* G = (const T*)a.ex.G;
  const T* q = Q + t * a.ex.ldq;
  const unsigned lt_mask = (1u << lane) - 1u;
  // 1. gather the pooled candidates (approximate score, column).  Lane l walks part l % 16 from
  // entry l / 16 in steps of 2: the fill counts are read once, only filled entries are touched and
  // four independent loads are in flight per lane (this phase is pure latency).
  const int part = lane & (SEL_MAX_PARTS - 1), half = lane / SEL_MAX_PARTS;
  int fill = 0;
  int64_t slot = 0;
  if (part < a.splits) {
    slot = (int64_t)part * a.ex.N + t;
    fill = min((int)a.pool_meta[slot].x, a.pool);
  }
  int maxfill = fill;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxfill = max(maxfill, __shfl_xor_sync(0xffffffffu, maxfill, o));
  const float2* __restrict__ pb = a.pool_buf + slot * a.pool;
  int ncand = 0;
  for (int i0 = 0; i0 < maxfill; i0 += 8) {
    float2 e[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + half + 2 * u;
      e[u] = i < fill ? pb[i] : make_float2(NAN, __int_as_float(-1));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = __float_as_int(e[u].y);
      const float ap = e[u].x;
      const bool ok = (i0 + half + 2 * u) < fill && j >= 0 && j < a.ex.M && ap == ap;
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int o = ncand + __popc(m & lt_mask);
        ca[w][o] = ap;
        cj[w][o] = j;
      }
      ncand += __popc(m);
    }
  }
  __syncwarp();
  // ||q||^2 for the guard band and the reported L2 distances (not rank-deciding: a warp-parallel sum;
  // round 1 had every lane walk the row sequentially, ~20 us of pure latency per warp)
  const double qq = warp_sq64(q, a.ex.D);
  const double qn = sqrt(qq) * (1.0 + 1e-12);
  const double gmax_sq = (double)__uint_as_float(*a.max_sq_bits);
  const double gn = sqrt(gmax_sq);
  const double delta = a.ex.metric == VTC_METRIC_L2
                           ? 2.0 * a.guard_rel * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn)
                           : (double)a.guard_rel * qn * gn + 1.2e-7 * qn * gn;
  // 2. k-th smallest APPROXIMATE score a_k (k rounds of "smallest (value, slot) above the last")
  float last_v = -INFINITY;
  int last_c = -1;
  int have = 0;
  for (int r = 0; r < a.k; ++r) {
    float bv = INFINITY;
    int bc = 0x7fffffff;
    for (int c = lane; c < ncand; c += 32) {
      const float x = ca[w][c];
      const bool above = x > last_v || (x == last_v && c > last_c);
      if (above && (x < bv || (x == bv && c < bc))) bv = x, bc = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ov < bv || (ov == bv && oc < bc)) bv = ov, bc = oc;
    }
    if (bc == 0x7fffffff) break;
    last_v = bv, last_c = bc;
    ++have;
  }
  // every member of the exact top-k has approximate score <= a_k + 2 delta
  const double thr = have == a.k ? (double)last_v + 2.0 * delta : INFINITY;
  // 3. compact the candidates that can still be in the top-k to the front of the list (in place:
  // a pass only writes slots it or an earlier pass has already read), so that the fp64 re-scoring
  // below runs with full warps instead of once per 32-slot stripe with a few live lanes
  int nsurv = 0;
  for (int c0 = 0; c0 < ncand; c0 += 32) {
    const int c = c0 + lane;
    const int j = c < ncand ? cj[w][c] : -1;
    const bool keep = j >= 0 && (double)ca[w][c] <= thr;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) cj[w][nsurv + __popc(m & lt_mask)] = j;
    nsurv += __popc(m);
    __syncwarp();
  }
  if (nsurv > SEL_MAX_SURV) {  // pathological ties: let the brute-force kernel do this row
    if (lane == 0) a.row_flag[t] = 1u;
    return;
  }
  // exact re-scoring, one survivor per lane
  for (int c = lane; c < nsurv; c += 32) {
    int j = cj[w][c];
    const double d = exact_score(q, G + (int64_t)j * a.ex.ldg, a.ex.D, a.ex.metric,
                                 a.ex.metric == VTC_METRIC_L2 ? a.ex.sq64[j] : 0.0);
    if (d != d) j = -1;  // NaN scores are never selected
    cd[w][c] = d;
    cj[w][c] = j;
  }
  __syncwarp();
  // 4. exact selection by (score, column)
  double dk = -INFINITY;  // exact score of the last selected candidate
  int found = 0;
  for (int r = 0; r < a.k; ++r) {
    double bd = 0.0;
    int bj = -1, bc = -1;
    for (int c = lane; c < nsurv; c += 32) {
      if (cand_less(cd[w][c], cj[w][c], bd, bj)) bd = cd[w][c], bj = cj[w][c], bc = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (cand_less(od, oj, bd, bj)) bd = od, bj = oj, bc = oc;
    }
    if (bj >= 0) {
      if (lane == 0) {
        a.out_val[t * a.k + r] =
            (float)(a.ex.metric == VTC_METRIC_L2 ? bd + qq : bd);
        a.out_idx[t * a.k + r] = (int64_t)bj + a.ex.col_offset;
      }
      if ((bc & 31) == lane) cj[w][bc] = -1;  // owner retires it
      dk = bd;
      ++found;
    } else if (lane == 0) {
      a.out_val[t * a.k + r] = INFINITY;
      a.out_idx[t * a.k + r] = -1;
    }
    __syncwarp();
  }
  // 5. completeness: every column outside a pool has approx score >= its tau (+inf while the pool
  // still holds every column seen), hence exact score >= tau - delta; the selection is provably
  // right when the k-th exact score is below that.
  if (lane == 0) {
    bool ok = qq == qq;
    for (int s = 0; s < a.splits; ++s) {
      const float2 meta = a.pool_meta[(int64_t)s * a.ex.N + t];
      if (meta.y < INFINITY) {
        if (found < a.k || !(dk < (double)meta.y - delta)) ok = false;
      }
    }
    a.row_flag[t] = ok ? 0u : 1u;
  }
}

// Threshold for the main top-k pass from a SAMPLE of the gallery: `scores` [N, cols] holds the
// approximate scores of every query against the first `cols` gallery rows (a plain tensor-core
// product, EPI_STORE).  The k-th smallest a_k of a row bounds the k-th smallest exact score of the
// whole gallery (d_k(all) <= d_k(sample) <= a_k + delta), so every member of the exact top-k has
// approximate score < a_k + 3 delta.  One warp per query row: every lane keeps the 16 smallest of
// its share in sorted registers (branch-free bubble insert, taken only when a value beats the
// lane's worst), then k rounds of warp-min pop the k-th smallest overall.
template <typename T>
__global__ void __launch_bounds__(SEL_WARPS * 32)
topk_tau_kernel(TopkSelectArgs a, const float* __restrict__ scores, int cols, int64_t ld,
                float* __restrict__ tau0) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  (void)w;
  const int64_t t = (int64_t)blockIdx.x * SEL_WARPS + (threadIdx.x >> 5);
  if (t >= a.ex.N) return;
  const T* q = (const T*)a.ex.Q + t * a.ex.ldq;
  // every lane keeps the KEEP smallest of its share: the k-th smallest of the 32 * KEEP kept values
  // is >= the k-th smallest of the row (a subset), i.e. still a valid -- at worst looser -- threshold;
  // 4 per lane instead of 16 quarters the insertion work of this HBM-bound pass
  constexpr int KEEP = 4;
  float best[KEEP];
#pragma unroll
  for (int i = 0; i < KEEP; ++i) best[i] = INFINITY;
  const float4* row = reinterpret_cast<const float4*>(scores + t * ld);
  const int nv = cols / 4;
  for (int c0 = lane; c0 < nv; c0 += 128) {  // four independent 16-byte loads in flight per lane
    float4 v4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + 32 * u;
      v4[u] = c < nv ? __ldg(row + c) : make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float vs[4] = {v4[u].x, v4[u].y, v4[u].z, v4[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = vs[e];
        if (v < best[KEEP - 1]) {  // NaN never enters
#pragma unroll
          for (int i = 0; i < KEEP; ++i) {
            const float lo = fminf(best[i], v);
            v = fmaxf(best[i], v);
            best[i] = lo;
          }
        }
      }
    }
  }
  float kth = INFINITY;
  int have = 0;
  for (int r = 0; r < a.k; ++r) {
    float m = best[0];
    int src = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int os = __shfl_xor_sync(0xffffffffu, src, o);
      if (om < m || (om == m && os < src)) m = om, src = os;
    }
    if (!(m < INFINITY)) break;
    kth = m;
    ++have;
    if (lane == src) {  // pop this lane's head
#pragma unroll
      for (int i = 0; i + 1 < KEEP; ++i) best[i] = best[i + 1];
      best[KEEP - 1] = INFINITY;
    }
  }
  const double qq = warp_sq64(q, a.ex.D);
  const double qn = sqrt(qq) * (1.0 + 1e-12);
  const double gmax_sq = (double)__uint_as_float(*a.max_sq_bits);
  const double gn = sqrt(gmax_sq);
  const double delta = a.ex.metric == VTC_METRIC_L2
                           ? 2.0 * a.guard_rel * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn)
                           : (double)a.guard_rel * qn * gn + 1.2e-7 * qn * gn;
  if (lane == 0)
    tau0[t] = (have == a.k && qq == qq) ? __double2float_ru((double)kth + 3.0 * delta) : INFINITY;
}

int launch_topk_tau(const TopkSelectArgs& a, const float* scores, int cols, int64_t ld, float* tau0,
                    cudaStream_t s) {
  if (a.ex.N == 0) return VTC_OK;
  if (a.k > 16 || cols % 4 || ld % 4) return VTC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)ceil_div<int64_t>(a.ex.N, SEL_WARPS);
  if (a.ex.bf16)
    topk_tau_kernel<__nv_bfloat16><<<grid, SEL_WARPS * 32, 0, s>>>(a, scores, cols, ld, tau0);
  else
    topk_tau_kernel<float><<<grid, SEL_WARPS * 32, 0, s>>>(a, scores, cols, ld, tau0);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

constexpr int BRUTE_THREADS = 128;
constexpr int BRUTE_KMAX = 16;

template <typename T>
__global__ void __launch_bounds__(BRUTE_THREADS)
topk_brute_rows_kernel(TopkSelectArgs a) {
  const int64_t t = blockIdx.x;
  if (a.row_flag[t] == 0) return;
  __shared__ double ld[BRUTE_THREADS][BRUTE_KMAX];
  __shared__ int lj[BRUTE_THREADS][BRUTE_KMAX];
  __shared__ double rd[BRUTE_THREADS / 32];
  __shared__ int rj[BRUTE_THREADS / 32], rt[BRUTE_THREADS / 32];
  __shared__ int winner;
  const int tid = threadIdx.x;
  const T* q = (const T*)a.ex.Q + t * a.ex.ldq;
  const T* G = (const T*)a.ex.G;
  const int k = a.k;
  int fill = 0;
  for (int64_t j = tid; j < a.ex.M; j += BRUTE_THREADS) {
    const double d = exact_score(q, G + j * a.ex.ldg, a.ex.D, a.ex.metric,
                                 a.ex.metric == VTC_METRIC_L2 ? a.ex.sq64[j] : 0.0);
    if (d != d) continue;
    if (fill == k && !(d < ld[tid][k - 1])) continue;  // j grows: ties keep the earlier index
    int pos = fill < k ? fill : k - 1;
    while (pos > 0 && d < ld[tid][pos - 1]) {
      ld[tid][pos] = ld[tid][pos - 1];
      lj[tid][pos] = lj[tid][pos - 1];
      --pos;
    }
    ld[tid][pos] = d;
    lj[tid][pos] = (int)j;
    if (fill < k) ++fill;
  }
  const double qq = sq_seq64(q, a.ex.D);
  int head = 0;
  for (int r = 0; r < k; ++r) {
    double bd = head < fill ? ld[tid][head] : 0.0;
    int bj = head < fill ? lj[tid][head] : -1;
    int bt = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, bd, o);
      const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
      if (cand_less(od, oj, bd, bj)) bd = od, bj = oj, bt = ot;
    }
    if ((tid & 31) == 0) rd[tid >> 5] = bd, rj[tid >> 5] = bj, rt[tid >> 5] = bt;
    __syncthreads();
    if (tid == 0) {
      for (int i = 1; i < BRUTE_THREADS / 32; ++i)
        if (cand_less(rd[i], rj[i], rd[0], rj[0])) rd[0] = rd[i], rj[0] = rj[i], rt[0] = rt[i];
      winner = rj[0] >= 0 ? rt[0] : -1;
      a.out_val[t * k + r] =
          rj[0] >= 0 ? (float)(a.ex.metric == VTC_METRIC_L2 ? rd[0] + qq : rd[0]) : INFINITY;
      a.out_idx[t * k + r] = rj[0] >= 0 ? (int64_t)rj[0] + a.ex.col_offset : -1;
    }
    __syncthreads();
    if (winner == tid) ++head;
    __syncthreads();
  }
}

__global__ void topk_merge_kernel(const float* __restrict__ vals, const int64_t* __restrict__ idx,
                                  int parts, int64_t N, int k, float* __restrict__ out_val,
                                  int64_t* __restrict__ out_idx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  unsigned char head[64];
  for (int p = 0; p < parts; ++p) head[p] = 0;
  for (int r = 0; r < k; ++r) {
    int best = -1;
    float bv = 0.f;
    int64_t bi = -1;
    for (int p = 0; p < parts; ++p) {
      if (head[p] >= k) continue;
      const int64_t off = ((int64_t)p * N + t) * k + head[p];
      const int64_t ci = idx[off];
      if (ci < 0) continue;
      const float cv = vals[off];
      if (best < 0 || cv < bv || (cv == bv && ci < bi)) best = p, bv = cv, bi = ci;
    }
    if (best >= 0) {
      out_val[t * k + r] = bv;
      out_idx[t * k + r] = bi;
      ++head[best];
    } else {
      out_val[t * k + r] = INFINITY;
      out_idx[t * k + r] = -1;
    }
  }
}

int launch_topk_select(const TopkSelectArgs& a, cudaStream_t s) {
  if (a.ex.N == 0) return VTC_OK;
  if (a.splits > SEL_MAX_PARTS || a.pool > 64 || a.k > BRUTE_KMAX) return VTC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)ceil_div<int64_t>(a.ex.N, SEL_WARPS);
  if (a.ex.bf16)
    topk_select_kernel<__nv_bfloat16><<<grid, SEL_WARPS * 32, 0, s>>>(a);
  else
    topk_select_kernel<float><<<grid, SEL_WARPS * 32, 0, s>>>(a);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_topk_brute_rows(const TopkSelectArgs& a, cudaStream_t s) {
  if (a.ex.N == 0) return VTC_OK;
  if (a.k > BRUTE_KMAX) return VTC_ERR_UNSUPPORTED_SHAPE;
  if (a.ex.bf16)
    topk_brute_rows_kernel<__nv_bfloat16><<<(unsigned)a.ex.N, BRUTE_THREADS, 0, s>>>(a);
  else
    topk_brute_rows_kernel<float><<<(unsigned)a.ex.N, BRUTE_THREADS, 0, s>>>(a);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_topk_merge(const float* vals, const int64_t* idx, int parts, int64_t N, int k,
                      float* out_val, int64_t* out_idx, cudaStream_t s) {
  if (N == 0) return VTC_OK;
  if (parts < 1 || parts > 64 || k < 1 || k > 255) return VTC_ERR_INVALID_ARG;
  topk_merge_kernel<<<(unsigned)ceil_div<int64_t>(N, 128), 128, 0, s>>>(vals, idx, parts, N, k,
                                                                        out_val, out_idx);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc
