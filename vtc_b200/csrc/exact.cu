// exact.cu -- fp64-sequential SIMT kernels: the bit-exactness anchor of the retrieval path.
//
// "fp64-sequential" arithmetic (identical to oracle/vtc_oracle.c):
//   dot(q,x) = fold_{k=0..D-1} acc = fma((double)q[k], (double)x[k], acc)   (products are exact)
//   sq(x) = dot(x,x);  L2 score d = sq(x) - 2 dot(q,x);  DOT score d = -dot(q,x)
// Every dot product below is accumulated by ONE thread in k order, so results are bit-identical
// to the CPU oracle regardless of tiling.
//
// Kernels: canonical row norms, ground-truth scores + guard-band thresholds, brute-force rank,
// re-check of ambiguous pairs emitted by the tensor-core pass, R@K / median finalisation.
// Replaces faiss.GpuIndexFlatL2.search + the host loop at model/metric.py:140-160.
#include "exact.cuh"

namespace vtc {

// ------------------------------------------------------------------------------------------------
// canonical squared norms (thread per row, sequential in k) + max for the guard band
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void sqnorm64_kernel(const T* __restrict__ X, int64_t rows, int D, int64_t ld,
                                double* __restrict__ sq64, float* __restrict__ sq32,
                                unsigned int* __restrict__ max_sq_bits) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mine = 0.f;
  if (r < rows) {
    const double acc = sq_seq64(X + r * ld, D);
    if (sq64) sq64[r] = acc;
    const float f = (float)acc;
    if (sq32) sq32[r] = f;
    // NaN / inf rows do not take part in the guard-band scale (their scores are NaN / inf anyway)
    if (f == f && f < 3.0e38f) mine = f;
  }
  if (max_sq_bits) {
    mine = warp_max(mine);
    if ((threadIdx.x & 31) == 0 && mine > 0.f) atomicMax(max_sq_bits, __float_as_uint(mine));
  }
}

// ------------------------------------------------------------------------------------------------
// d(t, gt(t)) and the guard-band thresholds of the tensor-core pass
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gt_score_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G,
                                int64_t ldg, const double* __restrict__ sq64, int64_t N, int64_t M,
                                int D, const int64_t* __restrict__ gt, int64_t row_offset,
                                int64_t col_offset, int metric, const double* __restrict__ gt_in,
                                double* __restrict__ gt_out, float2* __restrict__ thr,
                                const unsigned int* __restrict__ max_sq_bits, float guard_rel) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const T* q = Q + t * ldq;
  double d0;
  if (gt_in) {
    d0 = gt_in[t];
  } else {
    const int64_t g = (gt ? gt[t] : t + row_offset) - col_offset;
    if (g >= 0 && g < M) {
      const double acc = dot_seq64(q, G + g * ldg, D);
      d0 = metric == VTC_METRIC_L2 ? sq64[g] - 2.0 * acc : -acc;
    } else {
      d0 = nan("");
    }
  }
  if (gt_out) gt_out[t] = d0;
  if (thr) {
    const double qq = sq_seq64(q, D);
    const double qn = sqrt(qq);
    const double gmax_sq = (double)__uint_as_float(*max_sq_bits);
    const double gn = sqrt(gmax_sq);
    // |approx - exact| <= delta for every pair of this query (see DESIGN.md "guard band"):
    //   dot error  <= guard_rel * |q| * max|x|
    //   L2: d = sq32 - 2*acc in fp32 adds the rounding of sq32 and of the FMA.
    double delta;
    if (metric == VTC_METRIC_L2)
      delta = 2.0 * guard_rel * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn);
    else
      delta = (double)guard_rel * qn * gn + 1.2e-7 * qn * gn;
    float lo = __double2float_rd(d0 - delta);
    float hi = __double2float_ru(d0 + delta);
    if (!(qq == qq) || !(d0 == d0)) lo = hi = nanf("");
    thr[t] = make_float2(lo, hi);
  }
}

// ------------------------------------------------------------------------------------------------
// brute-force rank: register-tiled fp64 "GEMM" whose accumulators run sequentially in k
// ------------------------------------------------------------------------------------------------
constexpr int BR_T = 64;   // block tile (queries x gallery rows)
constexpr int BR_K = 16;   // k chunk
constexpr int BR_PAD = 2;  // doubles of padding per smem row

template <typename T>
__global__ void __launch_bounds__(256)
rank_brute_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G, int64_t ldg,
                  const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N,
                  int64_t M, int D, const int64_t* __restrict__ gt, int64_t row_offset,
                  int64_t col_offset, int metric, int* __restrict__ rank,
                  const unsigned int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0) return;
  __shared__ double Qs[BR_K][BR_T + BR_PAD];
  __shared__ double Gs[BR_K][BR_T + BR_PAD];
  __shared__ int cnt_s[BR_T];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t q_tiles = ceil_div<int64_t>(N, BR_T), g_tiles = ceil_div<int64_t>(M, BR_T);
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader mapping: 64 rows x 16 k
  for (int64_t tile = blockIdx.x; tile < q_tiles * g_tiles; tile += gridDim.x) {
    // consecutive blocks share the gallery tile (L2 reuse), queries vary fastest
    const int64_t q0 = (tile % q_tiles) * BR_T, g0 = (tile / q_tiles) * BR_T;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    if (tid < BR_T) cnt_s[tid] = 0;
    for (int k0 = 0; k0 < D; k0 += BR_K) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lk + e;
        const int64_t qr = q0 + lrow, gr = g0 + lrow;
        Qs[lk + e][lrow] = (qr < N && k < D) ? to_f64(Q[qr * ldq + k]) : 0.0;
        Gs[lk + e][lrow] = (gr < M && k < D) ? to_f64(G[gr * ldg + k]) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BR_K; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Qs[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Gs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t t = q0 + ty * 4 + i;
      if (t >= N) continue;
      const double d0 = dgt[t];
      const int64_t g = gt ? gt[t] : t + row_offset;
      int c = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t jl = g0 + tx * 4 + j;
        if (jl >= M) continue;
        const int64_t jg = jl + col_offset;
        if (jg == g) continue;
        const double d = metric == VTC_METRIC_L2 ? sq64[jl] - 2.0 * acc[i][j] : -acc[i][j];
        c += (d < d0) || (d == d0 && jg < g);
      }
      if (c) atomicAdd(&cnt_s[ty * 4 + i], c);
    }
    __syncthreads();
    if (tid < BR_T && cnt_s[tid] && q0 + tid < N) atomicAdd(&rank[q0 + tid], cnt_s[tid]);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// re-check of the ambiguous column groups emitted by the tensor-core pass
// ------------------------------------------------------------------------------------------------
// The list has one segment per CTA of the tensor-core launch; entry (t, j0) means "row t has a
// score inside the guard band among gallery columns [j0, j0 + 8)": the tensor-core pass added
// nothing for that group, so all 8 columns are decided here in canonical arithmetic.
constexpr int RECHECK_GROUP = 8;
constexpr int RECHECK_PARTS = 4;  // blocks per segment

template <typename T>
__global__ void __launch_bounds__(256)
recheck_kernel(const int2* __restrict__ list, const unsigned int* __restrict__ seg_count,
               unsigned int seg_cap, const T* __restrict__ Q, int64_t ldq,
               const T* __restrict__ G, int64_t ldg, const double* __restrict__ sq64,
               const double* __restrict__ dgt, int64_t N, int64_t M, int D,
               const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset, int metric,
               int* __restrict__ rank, unsigned int* __restrict__ overflow) {
  const int seg = blockIdx.x / RECHECK_PARTS, part = blockIdx.x % RECHECK_PARTS;
  const unsigned int n = seg_count[seg];
  if (n > seg_cap) {
    if (part == 0 && threadIdx.x == 0) *overflow = 1u;
    return;  // the brute-force fallback recomputes everything
  }
  const int2* seg_list = list + (size_t)seg * seg_cap;
  for (unsigned int u = part * blockDim.x + threadIdx.x; u < n * RECHECK_GROUP;
       u += RECHECK_PARTS * blockDim.x) {
    const int2 e = seg_list[u / RECHECK_GROUP];
    const int64_t t = e.x, jl = (int64_t)e.y + (u % RECHECK_GROUP);
    if (t >= N || jl >= M) continue;  // zero-padded tile rows / columns
    const int64_t g = gt ? gt[t] : t + row_offset;
    const int64_t jg = jl + col_offset;
    if (jg == g) continue;
    const double acc = dot_seq64(Q + t * ldq, G + jl * ldg, D);
    const double d = metric == VTC_METRIC_L2 ? sq64[jl] - 2.0 * acc : -acc;
    const double d0 = dgt[t];
    if ((d < d0) || (d == d0 && jg < g)) atomicAdd(&rank[t], 1);
  }
}

// dst[j] = j < M ? (src ? src[j] : 0) : pad   for j in [0, Mpad)   (in place allowed)
__global__ void fill_bias_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t M,
                                 int64_t Mpad, float pad) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < Mpad) dst[j] = j < M ? (src ? src[j] : 0.f) : pad;
}

__global__ void zero_if_flag_kernel(int* __restrict__ buf, int64_t n,
                                    const unsigned int* __restrict__ flag) {
  if (*flag == 0) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    buf[i] = 0;
}

__global__ void rank_commit_kernel(const int* __restrict__ tmp, int* __restrict__ rank, int64_t n,
                                   int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rank[i] = (accumulate ? rank[i] : 0) + tmp[i];
}

// ------------------------------------------------------------------------------------------------
// finalisation: NaN ground truth -> rank = M_total; R@K hit counts; median rank
// ------------------------------------------------------------------------------------------------
struct KVals {
  int k[8];
  int nk;
};

__global__ void rank_finalize_kernel(int* __restrict__ rank, const double* __restrict__ dgt,
                                     int64_t N, int M_total, KVals kv,
                                     unsigned long long* __restrict__ hits) {
  __shared__ int sh[8];
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  int local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < N;
       t += (int64_t)gridDim.x * blockDim.x) {
    int r = rank[t];
    if (dgt) {
      const double d0 = dgt[t];
      if (d0 != d0) {
        r = M_total;
        rank[t] = r;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < kv.nk) local[i] += r < kv.k[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < kv.nk) {
      int v = local[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[i], v);
    }
  }
  __syncthreads();
  if (threadIdx.x < kv.nk && sh[threadIdx.x] && hits)
    atomicAdd(&hits[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

// median(rank)+1 with numpy semantics by a two-level radix select over multi-block histograms.
// ws layout (uint32): hist0[65536] | hist1a[65536] | hist1b[65536] | sel[8]
//   level 0 bins by rank >> shift (shift chosen so that bins cover [0, M_total]), level 1 by the
//   low `shift` bits of the ranks inside the selected bucket(s).
constexpr int MED_BINS = 65536;

__global__ void med_hist0_kernel(const int* __restrict__ rank, int64_t N, int shift,
                                 unsigned int* __restrict__ hist0) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    unsigned int b = ((unsigned int)rank[i]) >> shift;
    if (b >= MED_BINS) b = MED_BINS - 1;
    atomicAdd(&hist0[b], 1u);
  }
}

// one block: locate the bin holding order statistic `target` (exclusive prefix <= target)
__device__ void med_find(const unsigned int* __restrict__ hist, unsigned int target,
                         unsigned int* part /*[1024] smem*/, unsigned int* bin_out,
                         unsigned int* rem_out) {
  const int tid = threadIdx.x;
  unsigned int s = 0;
  for (int i = 0; i < MED_BINS / 1024; ++i) s += hist[tid * (MED_BINS / 1024) + i];
  part[tid] = s;
  __syncthreads();
  if (tid == 0) {
    unsigned int run = 0;
    int b = 0;
    while (b < 1023 && run + part[b] <= target) run += part[b++];
    int bin = b * (MED_BINS / 1024);
    const int last = bin + MED_BINS / 1024 - 1;
    while (bin < last && run + hist[bin] <= target) run += hist[bin++];
    *bin_out = (unsigned int)bin;
    *rem_out = target - run;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
med_select0_kernel(const unsigned int* __restrict__ hist0, int64_t N, unsigned int* __restrict__ sel) {
  __shared__ unsigned int part[1024];
  med_find(hist0, (unsigned int)((N - 1) / 2), part, &sel[0], &sel[1]);
  med_find(hist0, (unsigned int)(N / 2), part, &sel[2], &sel[3]);
}

__global__ void med_hist1_kernel(const int* __restrict__ rank, int64_t N, int shift,
                                 const unsigned int* __restrict__ sel,
                                 unsigned int* __restrict__ hist1a,
                                 unsigned int* __restrict__ hist1b) {
  const unsigned int ba = sel[0], bb = sel[2];
  const unsigned int mask = (1u << shift) - 1u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned int r = (unsigned int)rank[i];
    unsigned int b = r >> shift;
    if (b >= MED_BINS) b = MED_BINS - 1;
    if (b == ba) atomicAdd(&hist1a[r & mask], 1u);
    if (b == bb && bb != ba) atomicAdd(&hist1b[r & mask], 1u);
  }
}

__global__ void __launch_bounds__(1024)
med_select1_kernel(const unsigned int* __restrict__ hist1a, const unsigned int* __restrict__ hist1b,
                   int shift, const unsigned int* __restrict__ sel, double* __restrict__ medr) {
  __shared__ unsigned int part[1024];
  __shared__ unsigned int lo[2], rem[2];
  med_find(hist1a, sel[1], part, &lo[0], &rem[0]);
  med_find(sel[2] != sel[0] ? hist1b : hist1a, sel[3], part, &lo[1], &rem[1]);
  if (threadIdx.x == 0) {
    const double v0 = (double)((sel[0] << shift) | lo[0]);
    const double v1 = (double)((sel[2] << shift) | lo[1]);
    *medr = 0.5 * (v0 + v1) + 1.0;
  }
}

__global__ void med_nan_kernel(double* medr) { *medr = nan(""); }

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
template <typename T>
static int launch_sqnorm64_t(const void* X, int64_t rows, int D, int64_t ld, double* sq64,
                             float* sq32, unsigned int* max_sq_bits, cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  sqnorm64_kernel<T><<<(unsigned)ceil_div<int64_t>(rows, 128), 128, 0, s>>>(
      (const T*)X, rows, D, ld, sq64, sq32, max_sq_bits);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_sqnorm64(const void* X, bool bf16, int64_t rows, int D, int64_t ld, double* sq64,
                    float* sq32, unsigned int* max_sq_bits, cudaStream_t s) {
  return bf16 ? launch_sqnorm64_t<__nv_bfloat16>(X, rows, D, ld, sq64, sq32, max_sq_bits, s)
              : launch_sqnorm64_t<float>(X, rows, D, ld, sq64, sq32, max_sq_bits, s);
}

template <typename T>
static int launch_gt_score_t(const ExactArgs& a, const double* gt_in, double* gt_out, float2* thr,
                             const unsigned int* max_sq_bits, float guard_rel, cudaStream_t s) {
  if (a.N == 0) return VTC_OK;
  gt_score_kernel<T><<<(unsigned)ceil_div<int64_t>(a.N, 128), 128, 0, s>>>(
      (const T*)a.Q, a.ldq, (const T*)a.G, a.ldg, a.sq64, a.N, a.M, a.D, a.gt, a.row_offset,
      a.col_offset, a.metric, gt_in, gt_out, thr, max_sq_bits, guard_rel);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_gt_score(const ExactArgs& a, const double* gt_in, double* gt_out, float2* thr,
                    const unsigned int* max_sq_bits, float guard_rel, cudaStream_t s) {
  return a.bf16 ? launch_gt_score_t<__nv_bfloat16>(a, gt_in, gt_out, thr, max_sq_bits, guard_rel, s)
                : launch_gt_score_t<float>(a, gt_in, gt_out, thr, max_sq_bits, guard_rel, s);
}

template <typename T>
static int launch_rank_brute_t(const ExactArgs& a, const double* dgt, int* rank,
                               const unsigned int* run_flag, cudaStream_t s) {
  if (a.N == 0 || a.M == 0) return VTC_OK;
  const int64_t tiles = ceil_div<int64_t>(a.N, BR_T) * ceil_div<int64_t>(a.M, BR_T);
  const unsigned grid = (unsigned)(tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8);
  rank_brute_kernel<T><<<grid, 256, 0, s>>>((const T*)a.Q, a.ldq, (const T*)a.G, a.ldg, a.sq64, dgt,
                                            a.N, a.M, a.D, a.gt, a.row_offset, a.col_offset,
                                            a.metric, rank, run_flag);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_rank_brute(const ExactArgs& a, const double* dgt, int* rank,
                      const unsigned int* run_flag, cudaStream_t s) {
  return a.bf16 ? launch_rank_brute_t<__nv_bfloat16>(a, dgt, rank, run_flag, s)
                : launch_rank_brute_t<float>(a, dgt, rank, run_flag, s);
}

template <typename T>
static int launch_recheck_t(const ExactArgs& a, const int2* list, const unsigned int* seg_count,
                            int nseg, unsigned int seg_cap, const double* dgt, int* rank,
                            unsigned int* overflow, cudaStream_t s) {
  if (nseg <= 0) return VTC_OK;
  recheck_kernel<T><<<nseg * RECHECK_PARTS, 256, 0, s>>>(list, seg_count, seg_cap, (const T*)a.Q,
                                                        a.ldq, (const T*)a.G, a.ldg, a.sq64, dgt,
                                                        a.N, a.M, a.D, a.gt, a.row_offset,
                                                        a.col_offset, a.metric, rank, overflow);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_recheck(const ExactArgs& a, const int2* list, const unsigned int* seg_count, int nseg,
                   unsigned int seg_cap, const double* dgt, int* rank, unsigned int* overflow,
                   cudaStream_t s) {
  return a.bf16 ? launch_recheck_t<__nv_bfloat16>(a, list, seg_count, nseg, seg_cap, dgt, rank,
                                                  overflow, s)
                : launch_recheck_t<float>(a, list, seg_count, nseg, seg_cap, dgt, rank, overflow, s);
}

int launch_fill_bias(float* dst, const float* src, int64_t M, int64_t Mpad, float pad,
                     cudaStream_t s) {
  if (Mpad <= 0) return VTC_OK;
  fill_bias_kernel<<<(unsigned)ceil_div<int64_t>(Mpad, 256), 256, 0, s>>>(dst, src, M, Mpad, pad);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_zero_if_flag(int* buf, int64_t n, const unsigned int* flag, cudaStream_t s) {
  if (n == 0) return VTC_OK;
  zero_if_flag_kernel<<<kNumSMs, 256, 0, s>>>(buf, n, flag);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_rank_commit(const int* tmp, int* rank, int64_t n, int accumulate, cudaStream_t s) {
  if (n == 0) return VTC_OK;
  rank_commit_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(tmp, rank, n, accumulate);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_rank_finalize(int* rank, const double* dgt, int64_t N, int64_t M_total,
                         const int* k_vals, int nk, int64_t* hits, double* medr, void* hist_ws,
                         cudaStream_t s) {
  KVals kv;
  kv.nk = nk;
  for (int i = 0; i < 8; ++i) kv.k[i] = i < nk ? k_vals[i] : 0;
  if (hits) {
    cudaError_t e = cudaMemsetAsync(hits, 0, sizeof(int64_t) * nk, s);
    if (e != cudaSuccess) return cuda_err(e);
  }
  if (N > 0) {
    const unsigned grid = (unsigned)(ceil_div<int64_t>(N, 256) < kNumSMs * 4
                                         ? ceil_div<int64_t>(N, 256)
                                         : kNumSMs * 4);
    rank_finalize_kernel<<<grid, 256, 0, s>>>(rank, dgt, N, (int)M_total, kv,
                                              (unsigned long long*)hits);
    VTC_LAUNCH_CHECK();
  }
  if (medr) {
    if (N <= 0) {
      med_nan_kernel<<<1, 1, 0, s>>>(medr);
      VTC_LAUNCH_CHECK();
      return VTC_OK;
    }
    unsigned int* h0 = (unsigned int*)hist_ws;
    unsigned int* h1a = h0 + MED_BINS;
    unsigned int* h1b = h1a + MED_BINS;
    unsigned int* sel = h1b + MED_BINS;
    cudaError_t e = cudaMemsetAsync(h0, 0, (3 * MED_BINS + 8) * sizeof(unsigned int), s);
    if (e != cudaSuccess) return cuda_err(e);
    int shift = 0;  // level-0 bins must cover ranks up to M_total (and any int32 beyond, clamped)
    while (shift < 16 && (M_total >> shift) >= MED_BINS) ++shift;
    const unsigned blocks = (unsigned)(ceil_div<int64_t>(N, 256) < kNumSMs * 4
                                           ? ceil_div<int64_t>(N, 256)
                                           : kNumSMs * 4);
    med_hist0_kernel<<<blocks, 256, 0, s>>>(rank, N, shift, h0);
    VTC_LAUNCH_CHECK();
    med_select0_kernel<<<1, 1024, 0, s>>>(h0, N, sel);
    VTC_LAUNCH_CHECK();
    med_hist1_kernel<<<blocks, 256, 0, s>>>(rank, N, shift, sel, h1a, h1b);
    VTC_LAUNCH_CHECK();
    med_select1_kernel<<<1, 1024, 0, s>>>(h1a, h1b, shift, sel, medr);
    VTC_LAUNCH_CHECK();
  }
  return VTC_OK;
}

}  // namespace vtc
