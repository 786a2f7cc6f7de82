// exact_dev.cuh -- device bodies of the fp64-sequential kernels that more than one kernel runs:
// the brute-force rank tile loop and the re-check of guard-band column groups.  Used by the
// stand-alone kernels of exact.cu and by the cooperative rank epilogue (rank_stage.cu).
// Arithmetic: see exact.cu ("fp64-sequential", identical to oracle/vtc_oracle.c).
#pragma once
#include "common.cuh"

namespace vtc {

constexpr int BR_T = 64;   // block tile (queries x gallery rows)
constexpr int BR_K = 16;   // k chunk
constexpr int BR_PAD = 2;  // doubles of padding per smem row
constexpr int BR_SMEM_BYTES = 2 * BR_K * (BR_T + BR_PAD) * 8 + BR_T * 4;

struct BruteSmem {
  double Qs[BR_K][BR_T + BR_PAD];
  double Gs[BR_K][BR_T + BR_PAD];
  int cnt[BR_T];
};

// rank[t] += #{j != gt : d(t,j) < d(t,gt) or (== and j < gt)} over the whole N x M problem, tiles
// dealt to blocks `first, first + stride, ...`; 256 threads, 4 x 4 register tile, every accumulator
// runs sequentially in k (bit-identical to one thread walking the row).
template <typename T>
__device__ __forceinline__ void rank_brute_tiles(
    BruteSmem& sm, const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G, int64_t ldg,
    const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N, int64_t M, int D,
    const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset, int metric,
    int* __restrict__ rank, int64_t first, int64_t stride) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t q_tiles = ceil_div<int64_t>(N, BR_T), g_tiles = ceil_div<int64_t>(M, BR_T);
  const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader mapping: 64 rows x 16 k
  for (int64_t tile = first; tile < q_tiles * g_tiles; tile += stride) {
    // consecutive blocks share the gallery tile (L2 reuse), queries vary fastest
    const int64_t q0 = (tile % q_tiles) * BR_T, g0 = (tile / q_tiles) * BR_T;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    if (tid < BR_T) sm.cnt[tid] = 0;
    for (int k0 = 0; k0 < D; k0 += BR_K) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lk + e;
        const int64_t qr = q0 + lrow, gr = g0 + lrow;
        sm.Qs[lk + e][lrow] = (qr < N && k < D) ? to_f64(Q[qr * ldq + k]) : 0.0;
        sm.Gs[lk + e][lrow] = (gr < M && k < D) ? to_f64(G[gr * ldg + k]) : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BR_K; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sm.Qs[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sm.Gs[kk][tx * 4 + j];
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
      if (c) atomicAdd(&sm.cnt[ty * 4 + i], c);
    }
    __syncthreads();
    if (tid < BR_T && sm.cnt[tid] && q0 + tid < N) atomicAdd(&rank[q0 + tid], sm.cnt[tid]);
    __syncthreads();
  }
}

// The tensor-core pass lists (t, j0): "row t has a score inside the guard band among gallery columns
// [j0, j0 + 8)" in one segment per CTA and adds nothing for such a group; all 8 columns are decided
// here in canonical arithmetic.  One virtual block = one quarter of a segment.
constexpr int RECHECK_GROUP = 8;
constexpr int RECHECK_PARTS = 4;

template <typename T>
__device__ __forceinline__ void recheck_part(
    int vblock, const int2* __restrict__ list, const unsigned int* __restrict__ seg_count,
    unsigned int seg_cap, const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G,
    int64_t ldg, const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N,
    int64_t M, int D, const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset,
    int metric, int* __restrict__ rank, unsigned int* __restrict__ overflow) {
  const int seg = vblock / RECHECK_PARTS, part = vblock % RECHECK_PARTS;
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

}  // namespace vtc
