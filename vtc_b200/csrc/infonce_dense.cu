// infonce_dense.cu -- clip_loss on a MATERIALISED similarity matrix (model/loss.py:18-22 when the
// caller hands in a tensor instead of this package's LazySim): the loss needs three reductions of
// `sim` -- row log-sum-exp, column log-sum-exp, diagonal -- and its gradient is elementwise,
//   d loss / d sim_ij = g/2n * (exp(sim_ij - row_lse_i) + exp(sim_ij - col_lse_j) - 2 [i == j]).
// Round 1 ran this through the feature path with B = I (an O(n^3) product and a second rounding of
// logits the caller had already computed).  HBM-bound: n^2 * 4 bytes read twice, written once in the
// backward.
#include "common.cuh"

namespace vtc {

constexpr int DN_WARPS = 8;

// blocks [0, row_blocks): one warp per row, online (max, sum) over the row, coalesced;
// blocks [row_blocks, ...): 32 columns per block, 8 row lanes per column, merged through shared memory
__global__ void __launch_bounds__(256)
infonce_dense_lse_kernel(const float* __restrict__ sim, int64_t n, int64_t ld, int row_blocks,
                         float* __restrict__ row_lse, float* __restrict__ col_lse,
                         float* __restrict__ diag) {
  __shared__ float sm_m[DN_WARPS][33], sm_l[DN_WARPS][33];
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < row_blocks) {
    const int64_t i = (int64_t)blockIdx.x * DN_WARPS + warp;
    if (i >= n) return;
    const float* row = sim + i * ld;
    float m = -INFINITY, l = 0.f;
    for (int64_t j = lane; j < n; j += 32) {
      const float x = row[j];
      const float mn = fmaxf(m, x);
      if (mn > -INFINITY) l = l * __expf(m - mn) + __expf(x - mn);
      m = mn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
      const float mn = fmaxf(m, m2);
      if (mn > -INFINITY) l = l * __expf(m - mn) + l2 * __expf(m2 - mn);
      m = mn;
    }
    if (lane == 0) {
      row_lse[i] = m + logf(l);
      diag[i] = row[i];
    }
    return;
  }
  const int64_t j = ((int64_t)blockIdx.x - row_blocks) * 32 + lane;
  float m = -INFINITY, l = 0.f;
  if (j < n)
    for (int64_t i = warp; i < n; i += DN_WARPS) {
      const float x = sim[i * ld + j];
      const float mn = fmaxf(m, x);
      if (mn > -INFINITY) l = l * __expf(m - mn) + __expf(x - mn);
      m = mn;
    }
  sm_m[warp][lane] = m, sm_l[warp][lane] = l;
  __syncthreads();
  if (warp == 0 && j < n) {
    for (int w = 1; w < DN_WARPS; ++w) {
      const float m2 = sm_m[w][lane], l2 = sm_l[w][lane];
      const float mn = fmaxf(m, m2);
      if (mn > -INFINITY) l = l * __expf(m - mn) + l2 * __expf(m2 - mn);
      m = mn;
    }
    col_lse[j] = m + logf(l);
  }
}

// loss = 0.5 * (mean_i[row_lse_i - diag_i] + mean_j[col_lse_j - diag_j]); one block, fixed order
__global__ void __launch_bounds__(256)
infonce_dense_loss_kernel(const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                          const float* __restrict__ diag, int64_t n, float* __restrict__ loss) {
  __shared__ double red[256];
  griddep_launch();
  griddep_wait();
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256)
    acc += ((double)row_lse[i] - (double)diag[i]) + ((double)col_lse[i] - (double)diag[i]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(0.5 * red[0] / (double)n);
}

__global__ void __launch_bounds__(256)
infonce_dense_bwd_kernel(const float* __restrict__ sim, int64_t n, int64_t ld,
                         const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                         const float* __restrict__ grad_loss, float* __restrict__ dsim,
                         int64_t ldd) {
  griddep_launch();
  griddep_wait();
  const float g = *grad_loss / (2.f * (float)n);
  const int64_t total = n * n;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
    const int64_t i = e / n, j = e % n;
    const float x = sim[i * ld + j];
    dsim[i * ldd + j] = g * (expf(x - row_lse[i]) + expf(x - col_lse[j]) - (i == j ? 2.f : 0.f));
  }
}

}  // namespace vtc

using namespace vtc;

extern "C" int vtc_infonce_dense_fwd(const float* sim, int64_t n, int64_t ld, float* loss,
                                     float* row_lse, float* col_lse, float* diag,
                                     vtc_stream_t stream) {
  if (!sim || !loss || !row_lse || !col_lse || !diag || n <= 0 || ld < n) return VTC_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  const int row_blocks = (int)ceil_div<int64_t>(n, DN_WARPS);
  const int col_blocks = (int)ceil_div<int64_t>(n, 32);
  launch_pdl(infonce_dense_lse_kernel, dim3((unsigned)(row_blocks + col_blocks)), dim3(256), 0, s, sim,
             n, ld, row_blocks, row_lse, col_lse, diag);
  VTC_LAUNCH_CHECK();
  launch_pdl(infonce_dense_loss_kernel, dim3(1), dim3(256), 0, s, (const float*)row_lse,
             (const float*)col_lse, (const float*)diag, n, loss);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

extern "C" int vtc_infonce_dense_bwd(const float* sim, int64_t n, int64_t ld, const float* row_lse,
                                     const float* col_lse, const float* grad_loss, float* dsim,
                                     int64_t ldd, vtc_stream_t stream) {
  if (!sim || !row_lse || !col_lse || !grad_loss || !dsim || n <= 0 || ld < n || ldd < n)
    return VTC_ERR_INVALID_ARG;
  const int64_t blocks = ceil_div<int64_t>(n * n, 256);
  launch_pdl(infonce_dense_bwd_kernel, dim3((unsigned)(blocks < kNumSMs * 16 ? blocks : kNumSMs * 16)),
             dim3(256), 0, (cudaStream_t)stream, sim, n, ld, row_lse, col_lse, grad_loss, dsim, ldd);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
