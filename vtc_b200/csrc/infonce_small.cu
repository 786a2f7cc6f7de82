// infonce_small.cu -- single-launch fused symmetric InfoNCE for training-sized batches.
//
// model/loss.py:18-22 on sim = (s*A) B^T with b = 50..256 rows (configs/*.jsonc) is launch-latency
// bound: 67 MFLOP at b = 256.  One kernel computes 32 x 32 logit tiles in fp32 FMA arithmetic,
// reduces each tile to per-row and per-column (max, sum-exp) partials, and the last block to finish
// (atomic ticket) merges the partials into row_lse / col_lse / diag and the scalar loss -- the
// logits never leave the SM.  Used for n <= 2048; larger batches take the tcgen05 path
// (sim_tc_kernel<LseEpi>).  VTC_PREC_BF16 rounds the inputs to bf16 on load (fp32 accumulate),
// the same semantics as the tensor-core path.
#include "common.cuh"

namespace vtc {

constexpr int NS_T = 32;   // tile edge
constexpr int NS_K = 32;   // k chunk

template <typename T>
__device__ __forceinline__ float ld_elem(const T* p, bool round_bf16) {
  const float v = to_f32(*p);
  return round_bf16 ? __bfloat162float(__float2bfloat16_rn(v)) : v;
}

// merge (m, l) log2-domain partials
__device__ __forceinline__ void lse_merge(float& m, float& l, float m2, float l2) {
  if (m2 == -INFINITY) return;
  const float mn = fmaxf(m, m2);
  l = l * exp2f(m - mn) + l2 * exp2f(m2 - mn);
  m = mn;
}

template <typename T>
__global__ void __launch_bounds__(256)
infonce_small_kernel(const T* __restrict__ A, const T* __restrict__ B, int n, int D,
                     const float* __restrict__ scale_ptr, int round_bf16,
                     float2* __restrict__ part_row, float2* __restrict__ part_col,
                     float* __restrict__ diag, unsigned int* __restrict__ ticket,
                     float* __restrict__ row_lse, float* __restrict__ col_lse,
                     float* __restrict__ loss) {
  __shared__ float As[NS_K][NS_T + 1];
  __shared__ float Bs[NS_K][NS_T + 1];
  __shared__ float Ls[NS_T][NS_T + 1];
  __shared__ double red[8];
  __shared__ bool last;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 2 x 2 outputs each
  const int tiles = (n + NS_T - 1) / NS_T;
  const int bx = blockIdx.x % tiles, by = blockIdx.x / tiles;
  const int i0 = by * NS_T, j0 = bx * NS_T;
  const bool rb = round_bf16 != 0;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const int lr = tid >> 3, lk = (tid & 7) * 4;  // loader: 32 rows x 32 k, 4 k per thread
  for (int k0 = 0; k0 < D; k0 += NS_K) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      As[lk + e][lr] = (i0 + lr < n && k < D) ? ld_elem(A + (int64_t)(i0 + lr) * D + k, rb) : 0.f;
      Bs[lk + e][lr] = (j0 + lr < n && k < D) ? ld_elem(B + (int64_t)(j0 + lr) * D + k, rb) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < NS_K; ++kk) {
      const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
      const float b0 = Bs[kk][tx * 2], b1 = Bs[kk][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
  }
  const float s = *scale_ptr;
  const float sl2 = s * 1.4426950408889634f;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = ty * 2 + i, c = tx * 2 + j;
      Ls[r][c] = (i0 + r < n && j0 + c < n) ? sl2 * acc[i][j] : -INFINITY;
      if (bx == by && r == c && i0 + r < n) diag[i0 + r] = s * acc[i][j];
    }
  __syncthreads();
  if (tid < 2 * NS_T) {
    const int idx = tid & (NS_T - 1);
    const bool is_row = tid < NS_T;
    float m = -INFINITY;
    for (int e = 0; e < NS_T; ++e) m = fmaxf(m, is_row ? Ls[idx][e] : Ls[e][idx]);
    float l = 0.f;
    if (m > -INFINITY)
      for (int e = 0; e < NS_T; ++e) l += exp2f((is_row ? Ls[idx][e] : Ls[e][idx]) - m);
    if (is_row) {
      if (i0 + idx < n) part_row[(int64_t)bx * n + i0 + idx] = make_float2(m, l);
    } else {
      if (j0 + idx < n) part_col[(int64_t)by * n + j0 + idx] = make_float2(m, l);
    }
  }
  // last block to finish merges the partials
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  double accd = 0.0;
  for (int t = tid; t < n; t += blockDim.x) {
    float mr = -INFINITY, lrw = 0.f, mc = -INFINITY, lc = 0.f;
    for (int b = 0; b < tiles; ++b) {
      const float2 pr = __ldcg(&part_row[(int64_t)b * n + t]);
      const float2 pc = __ldcg(&part_col[(int64_t)b * n + t]);
      lse_merge(mr, lrw, pr.x, pr.y);
      lse_merge(mc, lc, pc.x, pc.y);
    }
    const float rl = 0.6931471805599453f * (mr + log2f(lrw));
    const float cl = 0.6931471805599453f * (mc + log2f(lc));
    row_lse[t] = rl;
    col_lse[t] = cl;
    const float d = __ldcg(&diag[t]);
    accd += ((double)rl - (double)d) + ((double)cl - (double)d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, o);
  if ((tid & 31) == 0) red[tid >> 5] = accd;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    *loss = (float)(0.5 * tot / (double)n);
    *ticket = 0;  // leave the workspace reusable
  }
}

size_t infonce_small_ws_bytes(int64_t n) {
  const int64_t tiles = ceil_div<int64_t>(n, NS_T);
  return round_up<size_t>(2 * tiles * n * sizeof(float2), 256) + 256;
}

int launch_infonce_small(const void* A, const void* B, int64_t n, int D, bool in_bf16,
                         bool round_bf16, const float* scale, float* loss, float* row_lse,
                         float* col_lse, float* diag, void* wsp, size_t ws_bytes, cudaStream_t s) {
  Workspace ws(wsp, ws_bytes);
  const int64_t tiles = ceil_div<int64_t>(n, NS_T);
  float2* part_row = ws.take<float2>((size_t)tiles * n);
  float2* part_col = ws.take<float2>((size_t)tiles * n);
  unsigned int* ticket = ws.take<unsigned int>(1);
  if (!ws.ok()) return VTC_ERR_WORKSPACE;
  cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(unsigned int), s);
  if (e != cudaSuccess) return cuda_err(e);
  const unsigned grid = (unsigned)(tiles * tiles);
  if (in_bf16)
    infonce_small_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(
        (const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (int)n, D, scale, 0, part_row, part_col,
        diag, ticket, row_lse, col_lse, loss);
  else
    infonce_small_kernel<float><<<grid, 256, 0, s>>>((const float*)A, (const float*)B, (int)n, D,
                                                     scale, round_bf16 ? 1 : 0, part_row, part_col,
                                                     diag, ticket, row_lse, col_lse, loss);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc
