// prep.cu -- HBM-bound row kernels around the similarity GEMM:
//   * H1 `normalize` (model/model.py:26-27): row norms and x / ||x|| (no eps);
//   * operand preparation for the tcgen05 pass: fp32/bf16 rows -> K-major bf16 rows padded to a
//     multiple of 64 columns, optionally as the 3-term split used by VTC_PREC_EXACT
//       A side: [ hi | hi | lo ]      B side: [ hi | lo | hi ]
//     so that A'.B'^T = hi.hi + hi.lo + lo.hi  (hi = bf16(x), lo = bf16(x - hi)).
// One warp per row, 128-bit loads/stores where alignment allows; algorithmic traffic is
// rows*D*sizeof(in) read + rows*Kp*2 written.
#include "prep.cuh"

namespace vtc {

constexpr int ROWS_PER_BLOCK = 8;  // 8 warps

template <typename T>
__device__ __forceinline__ float load_elem(const T* row, int k) {
  return to_f32(row[k]);
}

template <typename T>
__global__ void __launch_bounds__(256)
row_norms_kernel(const T* __restrict__ X, int64_t rows, int D, int64_t ld,
                 float* __restrict__ inv_norm, float* __restrict__ sq_norm) {
  const int64_t r = (int64_t)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + r * ld;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = load_elem(x, k);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    if (sq_norm) sq_norm[r] = s;
    if (inv_norm) inv_norm[r] = 1.0f / sqrtf(s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
normalize_kernel(const T* __restrict__ X, int64_t rows, int D, int64_t ldx, T* __restrict__ Y,
                 int64_t ldy) {
  const int64_t r = (int64_t)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + r * ldx;
  T* y = Y + r * ldy;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float v = load_elem(x, k);
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  const float nrm = sqrtf(s);  // x / x.norm(): a zero row gives 0/0 = NaN like the reference
  for (int k = lane; k < D; k += 32) {
    const float v = load_elem(x, k) / nrm;
    if constexpr (sizeof(T) == 4)
      y[k] = v;
    else
      y[k] = __float2bfloat16_rn(v);
  }
}

// mode: 0 = plain [x | 0], 1 = split A side [hi|hi|lo|0], 2 = split B side [hi|lo|hi|0]
template <typename T>
__global__ void __launch_bounds__(256)
prep_operand_kernel(const T* __restrict__ X, int64_t rows, int D, int64_t ldx, int mode,
                    __nv_bfloat16* __restrict__ out, int Kp) {
  const int64_t r = (int64_t)blockIdx.x * ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + r * ldx;
  __nv_bfloat16* o = out + r * (int64_t)Kp;
  const int used = mode == 0 ? D : 3 * D;
  for (int k = lane; k < D; k += 32) {
    const float v = load_elem(x, k);
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    if (mode == 0) {
      o[k] = hi;
    } else {
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      o[k] = hi;
      o[D + k] = mode == 1 ? hi : lo;
      o[2 * D + k] = mode == 1 ? lo : hi;
    }
  }
  for (int k = used + lane; k < Kp; k += 32) o[k] = __float2bfloat16_rn(0.f);
}

template <typename T>
static int row_norms_t(const void* X, int64_t rows, int D, int64_t ld, float* inv_norm,
                       float* sq_norm, cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  row_norms_kernel<T><<<(unsigned)ceil_div<int64_t>(rows, ROWS_PER_BLOCK), 256, 0, s>>>(
      (const T*)X, rows, D, ld, inv_norm, sq_norm);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_row_norms(const void* X, bool bf16, int64_t rows, int D, int64_t ld, float* inv_norm,
                     float* sq_norm, cudaStream_t s) {
  return bf16 ? row_norms_t<__nv_bfloat16>(X, rows, D, ld, inv_norm, sq_norm, s)
              : row_norms_t<float>(X, rows, D, ld, inv_norm, sq_norm, s);
}

template <typename T>
static int normalize_t(const void* X, int64_t rows, int D, int64_t ldx, void* Y, int64_t ldy,
                       cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  normalize_kernel<T><<<(unsigned)ceil_div<int64_t>(rows, ROWS_PER_BLOCK), 256, 0, s>>>(
      (const T*)X, rows, D, ldx, (T*)Y, ldy);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_normalize(const void* X, bool bf16, int64_t rows, int D, int64_t ldx, void* Y,
                     int64_t ldy, cudaStream_t s) {
  return bf16 ? normalize_t<__nv_bfloat16>(X, rows, D, ldx, Y, ldy, s)
              : normalize_t<float>(X, rows, D, ldx, Y, ldy, s);
}

template <typename T>
static int prep_t(const void* X, int64_t rows, int D, int64_t ldx, int mode, __nv_bfloat16* out,
                  int Kp, cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  prep_operand_kernel<T><<<(unsigned)ceil_div<int64_t>(rows, ROWS_PER_BLOCK), 256, 0, s>>>(
      (const T*)X, rows, D, ldx, mode, out, Kp);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_prep_operand(const void* X, bool bf16, int64_t rows, int D, int64_t ldx, int mode,
                        __nv_bfloat16* out, int Kp, cudaStream_t s) {
  return bf16 ? prep_t<__nv_bfloat16>(X, rows, D, ldx, mode, out, Kp, s)
              : prep_t<float>(X, rows, D, ldx, mode, out, Kp, s);
}

}  // namespace vtc
