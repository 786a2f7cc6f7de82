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

#include "exact_dev.cuh"

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
// brute-force rank and re-check of the guard-band column groups (bodies: exact_dev.cuh)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
rank_brute_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G, int64_t ldg,
                  const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N,
                  int64_t M, int D, const int64_t* __restrict__ gt, int64_t row_offset,
                  int64_t col_offset, int metric, int* __restrict__ rank,
                  const unsigned int* __restrict__ run_flag) {
  if (run_flag && *run_flag == 0) return;
  __shared__ BruteSmem sm;
  rank_brute_tiles<T>(sm, Q, ldq, G, ldg, sq64, dgt, N, M, D, gt, row_offset, col_offset, metric,
                      rank, blockIdx.x, gridDim.x);
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

// median(rank)+1 with numpy semantics: three-level radix select (bits [21,32), [10,21), [0,10))
// over block-privatised shared-memory histograms.  Two order statistics are tracked at once
// (k_lo = (N-1)/2 and k_hi = N/2; their mean is the median), they may part ways at any level.
// ws layout (uint32): hist[3 levels][2 stats][2048] | state[8] = {prefix_lo, rem_lo, prefix_hi, rem_hi}
constexpr int MED_BINS = 2048;
__device__ __forceinline__ int med_shift(int level) { return level == 0 ? 21 : (level == 1 ? 10 : 0); }
__device__ __forceinline__ unsigned int med_mask(int level) { return level == 2 ? 1023u : 2047u; }

__global__ void __launch_bounds__(256)
med_hist_kernel(const int* __restrict__ rank, int64_t N, int level,
                const unsigned int* __restrict__ state, unsigned int* __restrict__ hist) {
  __shared__ unsigned int sh[2][MED_BINS];
  for (int i = threadIdx.x; i < 2 * MED_BINS; i += blockDim.x) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int shift = med_shift(level);
  const unsigned int mask = med_mask(level);
  // ranks must match the prefix selected so far (bits above this level's field)
  const int up = level == 0 ? 32 : med_shift(level - 1);
  const unsigned int pa = level == 0 ? 0u : state[0], pb = level == 0 ? 0u : state[2];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N;
       i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned int r = (unsigned int)rank[i];
    const unsigned int hi = up >= 32 ? 0u : (r >> up);
    const unsigned int bin = (r >> shift) & mask;
    if (hi == pa) atomicAdd(&sh[0][bin], 1u);
    if (hi == pb && pb != pa) atomicAdd(&sh[1][bin], 1u);
  }
  __syncthreads();
  unsigned int* h = hist + (size_t)level * 2 * MED_BINS;
  for (int i = threadIdx.x; i < 2 * MED_BINS; i += blockDim.x) {
    const unsigned int v = (&sh[0][0])[i];
    if (v) atomicAdd(&h[i], v);
  }
}

// one block of 1024 threads, 2 bins each: exclusive scan, then the thread whose bin straddles the
// target records (bin, remainder)
__device__ void med_pick(const unsigned int* __restrict__ h, unsigned int target,
                         unsigned int* sh_warp /*[32]*/, unsigned int* out_bin, unsigned int* out_rem) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const unsigned int c0 = h[2 * tid], c1 = h[2 * tid + 1];
  unsigned int incl = c0 + c1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sh_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    unsigned int x = sh_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += v;
    }
    sh_warp[lane] = x;  // inclusive over warps
  }
  __syncthreads();
  const unsigned int base = (w ? sh_warp[w - 1] : 0u) + incl - (c0 + c1);  // exclusive prefix
  if (target >= base && target < base + c0) {
    *out_bin = 2 * tid;
    *out_rem = target - base;
  } else if (target >= base + c0 && target < base + c0 + c1) {
    *out_bin = 2 * tid + 1;
    *out_rem = target - base - c0;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024)
med_select_kernel(const unsigned int* __restrict__ hist, int64_t N, int level,
                  unsigned int* __restrict__ state, double* __restrict__ medr) {
  __shared__ unsigned int sh_warp[32];
  __shared__ unsigned int bin[2], rem[2];
  const unsigned int* h = hist + (size_t)level * 2 * MED_BINS;
  const bool same = level == 0 || state[0] == state[2];
  const unsigned int ta = level == 0 ? (unsigned int)((N - 1) / 2) : state[1];
  const unsigned int tb = level == 0 ? (unsigned int)(N / 2) : state[3];
  if (threadIdx.x == 0) bin[0] = bin[1] = rem[0] = rem[1] = 0;
  __syncthreads();
  med_pick(h, ta, sh_warp, &bin[0], &rem[0]);
  med_pick(same ? h : h + MED_BINS, tb, sh_warp, &bin[1], &rem[1]);
  if (threadIdx.x == 0) {
    const int bits = level == 2 ? 10 : 11;
    const unsigned int pa = ((level == 0 ? 0u : state[0]) << bits) | bin[0];
    const unsigned int pb = ((level == 0 ? 0u : state[2]) << bits) | bin[1];
    state[0] = pa, state[1] = rem[0], state[2] = pb, state[3] = rem[1];
    if (level == 2) *medr = 0.5 * ((double)pa + (double)pb) + 1.0;
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
    unsigned int* hist = (unsigned int*)hist_ws;
    unsigned int* state = hist + 3 * 2 * MED_BINS;
    cudaError_t e = cudaMemsetAsync(hist, 0, (3 * 2 * MED_BINS + 8) * sizeof(unsigned int), s);
    if (e != cudaSuccess) return cuda_err(e);
    const unsigned blocks = (unsigned)(ceil_div<int64_t>(N, 1024) < kNumSMs
                                           ? ceil_div<int64_t>(N, 1024)
                                           : kNumSMs);
    for (int level = 0; level < 3; ++level) {
      med_hist_kernel<<<blocks, 256, 0, s>>>(rank, N, level, state, hist);
      VTC_LAUNCH_CHECK();
      med_select_kernel<<<1, 1024, 0, s>>>(hist, N, level, state, medr);
      VTC_LAUNCH_CHECK();
    }
  }
  return VTC_OK;
}

}  // namespace vtc
