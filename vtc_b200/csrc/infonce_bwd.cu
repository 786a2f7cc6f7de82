// infonce_bwd.cu -- backward of the fused symmetric InfoNCE (model/loss.py:18-22 under
// `loss.backward()`, trainer/trainer.py:79), from the saved row / column log-sum-exp:
//
//   W_ij   = g/(2n) * ( exp(s a_i.b_j - row_lse_i) + exp(s a_i.b_j - col_lse_j) - 2 [i == j] )
//   dA     = s * W   B        dB = s * W^T A        ds = sum_ij W_ij (a_i.b_j)
//
// Two implementations behind vtc_infonce_bwd:
//   n <= 2048 (training batches are 50..256, < 0.2 GFLOP: launch-latency bound): fp32 SIMT tiles,
//     W staged in the caller's workspace (<= 16 MB), three launches -- this file;
//   larger n: the tcgen05 path of api.cu::infonce_bwd_tc_impl -- logit tiles recomputed on the tensor
//     cores, gradient weights handed on as bf16 operand strips, no n x n array, any n.
#include <cstdlib>

#include "common.cuh"

namespace vtc {

constexpr int TB = 64;  // block tile
constexpr int TKK = 16; // k chunk

template <typename T>
__global__ void __launch_bounds__(256)
infonce_w_kernel(const T* __restrict__ A, const T* __restrict__ B, int64_t n, int D,
                 const float* __restrict__ scale_ptr, const float* __restrict__ row_lse,
                 const float* __restrict__ col_lse, const float* __restrict__ grad_loss,
                 float* __restrict__ W, float* __restrict__ dscale_acc, int round_bf16) {
  __shared__ float As[TKK][TB + 1];
  __shared__ float Bs[TKK][TB + 1];
  __shared__ float red[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * TB, j0 = (int64_t)blockIdx.x * TB;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += TKK) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      // VTC_PREC_BF16: the forward reduced the logits of the bf16-rounded features; recompute those
      float av = (i0 + lrow < n && k < D) ? to_f32(A[(i0 + lrow) * D + k]) : 0.f;
      float bv = (j0 + lrow < n && k < D) ? to_f32(B[(j0 + lrow) * D + k]) : 0.f;
      if (round_bf16) {
        av = __bfloat162float(__float2bfloat16_rn(av));
        bv = __bfloat162float(__float2bfloat16_rn(bv));
      }
      As[lk + e][lrow] = av;
      Bs[lk + e][lrow] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TKK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const float s = *scale_ptr;
  const float g = *grad_loss / (2.0f * (float)n);
  float ds = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = i0 + ty * 4 + i;
    if (r >= n) continue;
    const float rl = row_lse[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = j0 + tx * 4 + j;
      if (c >= n) continue;
      const float logit = s * acc[i][j];
      const float w = g * (expf(logit - rl) + expf(logit - col_lse[c]) - (r == c ? 2.f : 0.f));
      W[r * n + c] = w;
      ds = fmaf(w, acc[i][j], ds);
    }
  }
  ds = warp_sum(ds);
  if ((tid & 31) == 0) red[tid >> 5] = ds;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(dscale_acc, t);
  }
}

// out[m, d] = alpha * op(W)[m, k] * Y[k, d];  op = identity or transpose (W stored [n, n]).
template <typename T, bool kTrans>
__global__ void __launch_bounds__(256)
infonce_grad_kernel(const float* __restrict__ W, const T* __restrict__ Y, int64_t n, int D,
                    const float* __restrict__ alpha_ptr, float* __restrict__ out) {
  __shared__ float Ws[TKK][TB + 1];
  __shared__ float Ys[TKK][TB + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.y * TB;  // output rows
  const int d0 = blockIdx.x * TB;               // output cols
  float acc[4][4] = {};
  for (int64_t k0 = 0; k0 < n; k0 += TKK) {
    __syncthreads();
    // W tile: 64 rows x 16 k
    {
      const int lrow = tid >> 2, lk = (tid & 3) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t r = i0 + lrow, k = k0 + lk + e;
        float v = 0.f;
        if (r < n && k < n) v = kTrans ? W[k * n + r] : W[r * n + k];
        Ws[lk + e][lrow] = v;
      }
    }
    // Y tile: 16 k x 64 cols
    {
      const int lk = tid >> 4, lc = (tid & 15) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t k = k0 + lk;
        const int c = d0 + lc + e;
        Ys[lk][lc + e] = (k < n && c < D) ? to_f32(Y[k * D + c]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TKK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Ws[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ys[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  const float alpha = *alpha_ptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = i0 + ty * 4 + i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = d0 + tx * 4 + j;
      if (c < D) out[r * D + c] = alpha * acc[i][j];
    }
  }
}

// *out += sum(part[0 .. count)) in a fixed order (one block): the gradient of the logit scale from
// the per-(part, row) partial sums of the tensor-core path
__global__ void __launch_bounds__(256)
nce_ds_reduce_kernel(const float* __restrict__ part, int64_t count, float* __restrict__ out) {
  __shared__ double red[256];
  griddep_launch();
  griddep_wait();
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < count; i += 256) acc += (double)part[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out += (float)red[0];
}
int launch_nce_ds_reduce(const float* part, int64_t count, float* out, cudaStream_t s) {
  launch_pdl(nce_ds_reduce_kernel, dim3(1), dim3(256), 0, s, part, count, out);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int infonce_bwd_tc_impl(const float* A, const float* B, int64_t n, int D, int precision,
                        const float* scale, const float* row_lse, const float* col_lse,
                        const float* grad_loss, float* dA, float* dB, float* dscale, void* wsp,
                        size_t ws_bytes, cudaStream_t s);  // api.cu

template <typename T>
static int infonce_bwd_t(const void* A, const void* B, int64_t n, int D, const float* scale,
                         const float* row_lse, const float* col_lse, const float* grad_loss,
                         float* dA, float* dB, float* dscale, float* W, int round_bf16,
                         cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(dscale, 0, sizeof(float), s);
  if (e != cudaSuccess) return cuda_err(e);
  const dim3 gw((unsigned)ceil_div<int64_t>(n, TB), (unsigned)ceil_div<int64_t>(n, TB));
  infonce_w_kernel<T><<<gw, 256, 0, s>>>((const T*)A, (const T*)B, n, D, scale, row_lse, col_lse,
                                         grad_loss, W, dscale, round_bf16);
  VTC_LAUNCH_CHECK();
  const dim3 gg((unsigned)ceil_div(D, TB), (unsigned)ceil_div<int64_t>(n, TB));
  infonce_grad_kernel<T, false><<<gg, 256, 0, s>>>(W, (const T*)B, n, D, scale, dA);
  VTC_LAUNCH_CHECK();
  infonce_grad_kernel<T, true><<<gg, 256, 0, s>>>(W, (const T*)A, n, D, scale, dB);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc

using namespace vtc;

extern "C" int vtc_infonce_bwd(const void* A, const void* B, int64_t n, int D, int dtype,
                               int precision, const float* scale, const float* row_lse,
                               const float* col_lse,
                               const float* grad_loss, float* dA, float* dB, float* dscale,
                               void* ws, size_t ws_bytes, vtc_stream_t stream) {
  if (!A || !B || !scale || !row_lse || !col_lse || !grad_loss || !dA || !dB || !dscale || n <= 0 ||
      D <= 0 || (dtype != VTC_F32 && dtype != VTC_BF16) ||
      (precision != VTC_PREC_EXACT && precision != VTC_PREC_BF16))
    return VTC_ERR_INVALID_ARG;
  const bool force_tc = getenv("VTC_INFONCE_FORCE_TC") != nullptr;  // test knob (tests/ only)
  if (n > 2048 || force_tc) {
    // the tensor-core path takes fp32 features (what training hands in)
    if (dtype != VTC_F32) return VTC_ERR_UNSUPPORTED_SHAPE;
    return infonce_bwd_tc_impl((const float*)A, (const float*)B, n, D, precision, scale, row_lse,
                               col_lse, grad_loss, dA, dB, dscale, ws, ws_bytes,
                               (cudaStream_t)stream);
  }
  Workspace w(ws, ws_bytes);
  float* W = w.take<float>((size_t)n * n);
  if (!w.ok() || !W) return VTC_ERR_WORKSPACE;
  return dtype == VTC_BF16
             ? infonce_bwd_t<__nv_bfloat16>(A, B, n, D, scale, row_lse, col_lse, grad_loss, dA, dB,
                                            dscale, W, 0, (cudaStream_t)stream)
             : infonce_bwd_t<float>(A, B, n, D, scale, row_lse, col_lse, grad_loss, dA, dB, dscale,
                                    W, precision == VTC_PREC_BF16 ? 1 : 0, (cudaStream_t)stream);
}
