// fold.cu -- builds the fold operands of the EPI_RANK_FOLD pass (opt-in, VTC_RANK_FOLD=1).
//
// RankEpi spends FFMA + 2 FSET + 2 FADD per logit on "scale * acc + ||x_j||^2, compare with the two
// guard-band thresholds of row t" (model/metric.py:144-160 is the faiss search + hit loop this
// replaces).  Both the per-column term and the per-row term are rank-1, so they can ride in the
// MMA: with
//     Qx[t] = [ m'_t | 1 ],  Gx[j] = [ 1 | h_j ]      (each scalar as three bf16 pieces = 24 bits)
// one extra K16 step makes the accumulator acc' = q.x + h_j + m'_t, and
//     L2 : h_j = -||x_j||^2 / 2, m'_t = d(t,gt) / 2   =>  d(t,j) < d(t,gt)  <=>  acc' > 0
//     DOT: h_j = 0,              m'_t = d(t,gt)       =>  -q.x   < d(t,gt)  <=>  acc' > 0
// The decision that matters is still taken in canonical fp64 arithmetic: |acc'| <= w_t sends the
// column group to the same re-check as before.
#include "fold.cuh"

namespace vtc {
namespace {

// v ~ p[0] + p[1] + p[2], each a bf16; the remainder is below 2^-24 |v|
__device__ __forceinline__ void split3(double v, unsigned short (&p)[3]) {
  double rem = v;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const __nv_bfloat16 b = __float2bfloat16_rn((float)rem);
    p[i] = __bfloat16_as_ushort(b);
    rem -= (double)__bfloat162float(b);
  }
}

constexpr unsigned short kOne = 0x3f80;  // bf16 1.0
constexpr double kFar = -1.0e30;         // "never closer, never in the band"

__device__ __forceinline__ void store_fold_row(__nv_bfloat16* row, int cols,
                                               const unsigned short (&e)[6]) {
  uint4* o = reinterpret_cast<uint4*>(row);
  o[0] = make_uint4((uint32_t)e[0] | ((uint32_t)e[1] << 16), (uint32_t)e[2] | ((uint32_t)e[3] << 16),
                    (uint32_t)e[4] | ((uint32_t)e[5] << 16), 0u);
  for (int i = 1; i < cols / 8; ++i) o[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void fold_q_kernel(const float2* __restrict__ thr, const double* __restrict__ dgt,
                              const unsigned int* __restrict__ max_sq_bits, int64_t N, int metric,
                              float guard_rel, __nv_bfloat16* __restrict__ Qx, int cols,
                              float* __restrict__ fold_w, unsigned int* __restrict__ invalid) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  const float2 th = thr[t];
  const double d0 = dgt[t];
  double m = kFar;
  float w = -1.f;
  if (d0 == d0 && fabs(d0) < 1.0e30 && th.x == th.x && th.y == th.y) {
    // (hi - lo) / 2 >= the delta RankEpi's thresholds were built from (they are rounded outwards)
    const double delta = 0.5 * ((double)th.y - (double)th.x);
    const double gmax_sq = (double)__uint_as_float(*max_sq_bits);
    double wd;
    if (metric == VTC_METRIC_L2) {
      m = 0.5 * d0;
      // delta >= 2 g |q| max|x|; the fold step adds three more products per scalar to an
      // accumulator whose magnitude is bounded by S: 8 ulp (2^-24 each) of slack on top
      const double S = delta / (2.0 * (double)guard_rel) + 0.5 * gmax_sq + fabs(m);
      wd = 0.5 * delta + 4.8e-7 * S;
    } else {
      m = d0;
      const double S = delta / (double)guard_rel + fabs(m);
      wd = delta + 4.8e-7 * S;
    }
    w = __double2float_ru(wd * (1.0 + 1.0e-6));
  } else if (d0 == d0) {
    // an infinite (or absurdly large) ground-truth score still orders against finite scores in
    // canonical arithmetic, but cannot ride in a bf16 operand: let the brute-force fallback decide
    *invalid = 1u;
  }
  unsigned short p[3];
  split3(m, p);
  const unsigned short e[6] = {p[0], p[1], p[2], kOne, kOne, kOne};
  store_fold_row(Qx + t * cols, cols, e);
  fold_w[t] = w;
}

__global__ void fold_g_kernel(const double* __restrict__ sq64, int64_t M, int64_t Mpad, int metric,
                              __nv_bfloat16* __restrict__ Gx, int cols,
                              unsigned int* __restrict__ invalid) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Mpad) return;
  double h = kFar;
  if (j < M) {
    const double sq = sq64[j];
    if (fabs(sq) < 1.0e30)  // false for NaN and inf
      h = metric == VTC_METRIC_L2 ? -0.5 * sq : 0.0;
    else
      *invalid = 1u;
  }
  unsigned short p[3];
  split3(h, p);
  const unsigned short e[6] = {kOne, kOne, kOne, p[0], p[1], p[2]};
  store_fold_row(Gx + j * cols, cols, e);
}

// (lo, hi) of exact.cu::gt_score_kernel from d(t,gt) and an upper bound qq of ||q_t||^2
__device__ __forceinline__ float2 thresholds_from(double d0, double qq, double gmax_sq, int metric,
                                                  float guard_rel) {
  const double qn = sqrt(qq), gn = sqrt(gmax_sq);
  double delta;
  if (metric == VTC_METRIC_L2)
    delta = 2.0 * guard_rel * qn * gn + 2.4e-7 * (gmax_sq + 2.0 * qn * gn);
  else
    delta = (double)guard_rel * qn * gn + 1.2e-7 * qn * gn;
  float lo = __double2float_rd(d0 - delta);
  float hi = __double2float_ru(d0 + delta);
  if (!(qq == qq) || !(d0 == d0)) lo = hi = nanf("");
  return make_float2(lo, hi);
}

template <typename T>
__global__ void __launch_bounds__(256)
thr_fast_kernel(const T* __restrict__ Q, int64_t ldq, int64_t N, int D,
                const double* __restrict__ dgt, const unsigned int* __restrict__ max_sq_bits,
                int metric, float guard_rel, float2* __restrict__ thr) {
  const int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= N) return;
  const int lane = threadIdx.x & 31;
  const T* q = Q + t * ldq;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = to_f32(q[k]);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane != 0) return;
  // fp32 accumulation of D non-negative terms is within (D/32 + 5) ulp of the exact sum: 1e-4 covers
  // every D the library accepts (<= 8192)
  thr[t] = thresholds_from(dgt[t], (double)s * (1.0 + 1.0e-4),
                           (double)__uint_as_float(*max_sq_bits), metric, guard_rel);
}

template <typename T>
__global__ void __launch_bounds__(256)
qnorm_up_kernel(const T* __restrict__ X, int64_t ldx, int64_t rows, int D,
                float* __restrict__ qq_up) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + r * ldx;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = to_f32(x[k]);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  // (D/32 + 5) ulp of fp32 accumulation error at most: 1e-4 covers every D the library accepts
  if (lane == 0) qq_up[r] = s * (1.0f + 1.0e-4f);
}

__global__ void bias_max_kernel(const double* __restrict__ sq64, int64_t M, int64_t Mpad, int metric,
                                float* __restrict__ bias, unsigned int* __restrict__ max_sq_bits) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mine = 0.f;
  if (j < Mpad) {
    float b = INFINITY;
    if (j < M) {
      const float f = (float)sq64[j];
      b = metric == VTC_METRIC_L2 ? f : 0.f;
      if (f == f && f < 3.0e38f) mine = f;  // as sqnorm64_kernel: NaN / inf rows do not scale the band
    }
    bias[j] = b;
  }
  mine = warp_max(mine);
  if ((threadIdx.x & 31) == 0 && mine > 0.f) atomicMax(max_sq_bits, __float_as_uint(mine));
}

__global__ void thr_cached_kernel(const float* __restrict__ qq_up, const double* __restrict__ dgt,
                                  const unsigned int* __restrict__ max_sq_bits, int64_t N, int metric,
                                  float guard_rel, float2* __restrict__ thr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  thr[t] = thresholds_from(dgt[t], (double)qq_up[t], (double)__uint_as_float(*max_sq_bits), metric,
                           guard_rel);
}

}  // namespace

int launch_qnorm_up(const void* X, bool bf16, int64_t ldx, int64_t rows, int D, float* qq_up,
                    cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, 8);
  if (bf16)
    qnorm_up_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)X, ldx, rows, D, qq_up);
  else
    qnorm_up_kernel<float><<<grid, 256, 0, s>>>((const float*)X, ldx, rows, D, qq_up);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_bias_max(const double* sq64, int64_t M, int64_t Mpad, int metric, float* bias,
                    unsigned int* max_sq_bits, cudaStream_t s) {
  if (Mpad == 0) return VTC_OK;
  bias_max_kernel<<<(unsigned)ceil_div<int64_t>(Mpad, 256), 256, 0, s>>>(sq64, M, Mpad, metric, bias,
                                                                       max_sq_bits);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_thr_cached(const float* qq_up, const double* dgt, const unsigned int* max_sq_bits,
                      int64_t N, int metric, float guard_rel, float2* thr, cudaStream_t s) {
  if (N == 0) return VTC_OK;
  thr_cached_kernel<<<(unsigned)ceil_div<int64_t>(N, 256), 256, 0, s>>>(qq_up, dgt, max_sq_bits, N,
                                                                      metric, guard_rel, thr);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_thr_fast(const void* Q, bool bf16, int64_t ldq, int64_t N, int D, const double* dgt,
                    const unsigned int* max_sq_bits, int metric, float guard_rel, float2* thr,
                    cudaStream_t s) {
  if (N == 0) return VTC_OK;
  const unsigned grid = (unsigned)ceil_div<int64_t>(N, 8);
  if (bf16)
    thr_fast_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)Q, ldq, N, D, dgt,
                                                       max_sq_bits, metric, guard_rel, thr);
  else
    thr_fast_kernel<float><<<grid, 256, 0, s>>>((const float*)Q, ldq, N, D, dgt, max_sq_bits, metric,
                                                guard_rel, thr);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_fold_q(const float2* thr, const double* dgt, const unsigned int* max_sq_bits, int64_t N,
                  int metric, float guard_rel, __nv_bfloat16* Qx, int cols, float* fold_w,
                  unsigned int* invalid, cudaStream_t s) {
  if (N == 0) return VTC_OK;
  if (cols != 16 && cols != 64) return VTC_ERR_INVALID_ARG;
  fold_q_kernel<<<(unsigned)ceil_div<int64_t>(N, 128), 128, 0, s>>>(
      thr, dgt, max_sq_bits, N, metric, guard_rel, Qx, cols, fold_w, invalid);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_fold_g(const double* sq64, int64_t M, int64_t Mpad, int metric, __nv_bfloat16* Gx,
                  int cols, unsigned int* invalid, cudaStream_t s) {
  if (Mpad == 0) return VTC_OK;
  if (cols != 16 && cols != 64) return VTC_ERR_INVALID_ARG;
  fold_g_kernel<<<(unsigned)ceil_div<int64_t>(Mpad, 128), 128, 0, s>>>(sq64, M, Mpad, metric, Gx,
                                                                     cols, invalid);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc
