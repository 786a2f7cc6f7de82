// exact_dev.cuh -- device bodies of the fp64-sequential kernels that more than one kernel runs:
// the brute-force rank tile loop and the re-check of guard-band column groups.  Used by the
// stand-alone kernels of exact.cu and by the rank epilogue (rank_stage.cu).
// Arithmetic: see exact.cu ("fp64-sequential", identical to oracle/vtc_oracle.c).
#pragma once
#include "common.cuh"
#include "sim_tc.cuh"

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

// four consecutive elements [k, k + 4) of a row as floats; zero beyond D or for an invalid row
__device__ __forceinline__ void load_quad(const float* row, int k, int D, bool vec, float (&v)[4]) {
  if (row != nullptr && vec && k + 4 <= D) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(row + k));
    v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (row != nullptr && k + e < D) ? __ldg(row + k + e) : 0.f;
  }
}
__device__ __forceinline__ void load_quad(const __nv_bfloat16* row, int k, int D, bool vec,
                                          float (&v)[4]) {
  if (row != nullptr && vec && k + 4 <= D) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(row + k));
    v[0] = __uint_as_float(u.x << 16), v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16), v[3] = __uint_as_float(u.y & 0xffff0000u);
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e)
      v[e] = (row != nullptr && k + e < D) ? __bfloat162float(row[k + e]) : 0.f;
  }
}
template <typename T>
__device__ __forceinline__ bool rows_vectorisable(const T* base, int64_t ld) {
  constexpr uintptr_t kAlign = sizeof(T) == 4 ? 15 : 7;  // 4 elements
  return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & kAlign) == 0;
}

// The tensor-core pass lists (t, j0, mask): "row t scores inside the guard band at the columns j0 + i
// of [j0, j0 + 8) whose mask bit i is set" in one segment per CTA; it has already counted the
// group's columns that are certainly below the band, the masked ones are decided here in canonical
// arithmetic (round 2a re-decided all 8 columns of a group: 4.5x the L2 traffic for typically one
// in-band column).
//
// One warp decides FOUR groups at a time, a lane per column: the 4 x (query row + 8 gallery rows) are
// staged 32 columns at a time in a warp-private shared-memory tile by the whole warp (128-bit loads,
// 8 lanes per row segment, the next chunk already in flight), and every lane walks one column's fp64
// chain in k order.  A thread walking two rows on its own (round 1) is a chain of D/16 dependent L2
// round trips (~20 us at D = 512); here the chain itself (D dependent DFMAs, ~3 us) is what is left,
// with all 32 lanes busy -- this sits on the critical path of every chunked call.
constexpr int RECHECK_GROUP = 8;
constexpr int RC_GROUPS = 4;                       // groups per warp and step
constexpr int RC_KC = 32;                          // staged columns per step
constexpr int RC_LD = RC_KC + 4;                   // floats per staged row: 16-byte aligned rows, and
                                                   // 8 consecutive rows cover all 32 banks (LDS.128)
constexpr int RC_GROUP_ROWS = RECHECK_GROUP + 1;   // the query row + the group's gallery rows
constexpr int RC_ROWS = RC_GROUPS * RC_GROUP_ROWS; // 36
constexpr int RC_WARP_FLOATS = RC_ROWS * RC_LD;       // 1296 floats

// `e`: the packed entry of this lane's group (lanes 8g .. 8g+7 hold group g), if `has`.
template <typename T>
__device__ __forceinline__ void recheck_groups_warp(
    float* __restrict__ st, int2 e, bool has, const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G,
    int64_t ldg, const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N,
    int64_t M, int D, const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset,
    int metric, int* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  unsigned int cmask = 0;  // the group's columns that need the canonical decision
  if (has) {
    int t_, j_;
    tc::amb_unpack(e, &t_, &j_, &cmask);
    e = make_int2(t_, j_);
  } else {
    e = make_int2(-1, 0);
  }
  if (e.x >= N || e.y >= M) e.x = -1;  // zero-padded tile rows / columns
  if (e.x < 0) cmask = 0;
  const bool vq = rows_vectorisable(Q, ldq), vg = rows_vectorisable(G, ldg);
  // loader view: lane owns column quad (lane % 8) of staged rows 4 i + lane / 8, i = 0..8;
  // gallery rows outside the mask are not loaded at all (their chains run on zeros, unused)
  const T* src[RC_GROUP_ROWS];
  bool is_q[RC_GROUP_ROWS];
#pragma unroll
  for (int i = 0; i < RC_GROUP_ROWS; ++i) {
    const int row = 4 * i + (lane >> 3);
    const int g = row / RC_GROUP_ROWS, rr = row % RC_GROUP_ROWS;
    const int t = __shfl_sync(0xffffffffu, e.x, 8 * g);
    const int j0 = __shfl_sync(0xffffffffu, e.y, 8 * g);
    const unsigned int gm = __shfl_sync(0xffffffffu, cmask, 8 * g);
    is_q[i] = rr == 0;
    src[i] = nullptr;
    if (t >= 0 && gm != 0) {
      if (rr == 0)
        src[i] = Q + (int64_t)t * ldq;
      else if (((gm >> (rr - 1)) & 1u) && (int64_t)j0 + rr - 1 < M)
        src[i] = G + ((int64_t)j0 + rr - 1) * ldg;
    }
  }
  const int kq = 4 * (lane & 7);
  float v[RC_GROUP_ROWS][4];
#pragma unroll
  for (int i = 0; i < RC_GROUP_ROWS; ++i) load_quad(src[i], kq, D, is_q[i] ? vq : vg, v[i]);
  // walker view: lane = column (lane % 8) of group (lane / 8)
  const float* qs = st + (lane >> 3) * RC_GROUP_ROWS * RC_LD;
  const float* xs = qs + (1 + (lane & 7)) * RC_LD;
  double acc = 0.0;
  for (int k0 = 0; k0 < D; k0 += RC_KC) {
    __syncwarp();  // the walkers have finished the previous chunk
#pragma unroll
    for (int i = 0; i < RC_GROUP_ROWS; ++i)
      *reinterpret_cast<float4*>(st + (4 * i + (lane >> 3)) * RC_LD + kq) =
          make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
    __syncwarp();
    if (k0 + RC_KC < D) {
#pragma unroll
      for (int i = 0; i < RC_GROUP_ROWS; ++i)
        load_quad(src[i], k0 + RC_KC + kq, D, is_q[i] ? vq : vg, v[i]);
    }
    const int kn = min(RC_KC, D - k0);
    int k = 0;
    for (; k + 4 <= kn; k += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(qs + k);
      const float4 x4 = *reinterpret_cast<const float4*>(xs + k);
      acc = fma((double)q4.x, (double)x4.x, acc);
      acc = fma((double)q4.y, (double)x4.y, acc);
      acc = fma((double)q4.z, (double)x4.z, acc);
      acc = fma((double)q4.w, (double)x4.w, acc);
    }
    for (; k < kn; ++k) acc = fma((double)qs[k], (double)xs[k], acc);
  }
  if (e.x >= 0 && ((cmask >> (lane & 7)) & 1u)) {
    const int64_t t = e.x, jl = (int64_t)e.y + (lane & 7);
    if (jl < M) {
      const int64_t g = gt ? gt[t] : t + row_offset;
      const int64_t jg = jl + col_offset;
      if (jg != g) {
        const double d = metric == VTC_METRIC_L2 ? sq64[jl] - 2.0 * acc : -acc;
        const double d0 = dgt[t];
        if ((d < d0) || (d == d0 && jg < g)) atomicAdd(&rank[t], 1);
      }
    }
  }
}

// One group per warp, 128 columns per step (lanes 0..7 walk): a quarter of the steps of the
// four-group variant and loads a whole chunk ahead, i.e. the shorter critical path -- used for
// segments with so few groups that every warp gets at most a couple.
constexpr int RC1_KC = 128;
constexpr int RC1_LD = RC1_KC + 4;
static_assert(RC_GROUP_ROWS * RC1_LD <= RC_WARP_FLOATS, "the staging tile serves both variants");

template <typename T>
__device__ __forceinline__ void recheck_group_warp(
    float* __restrict__ st, int2 entry, const T* __restrict__ Q, int64_t ldq,
    const T* __restrict__ G, int64_t ldg, const double* __restrict__ sq64,
    const double* __restrict__ dgt, int64_t N, int64_t M, int D, const int64_t* __restrict__ gt,
    int64_t row_offset, int64_t col_offset, int metric, int* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  int t_, j_;
  unsigned int cmask;
  tc::amb_unpack(entry, &t_, &j_, &cmask);
  const int64_t t = t_, j0 = j_;
  if (t >= N || j0 >= M || cmask == 0) return;  // zero-padded tile rows / columns (warp-uniform)
  const bool vq = rows_vectorisable(Q, ldq), vg = rows_vectorisable(G, ldg);
  const T* qrow = Q + t * ldq;
  const T* xrow[RECHECK_GROUP];  // only the columns of the mask are loaded
#pragma unroll
  for (int r = 0; r < RECHECK_GROUP; ++r)
    xrow[r] = (((cmask >> r) & 1u) && j0 + r < M) ? G + (j0 + r) * ldg : (const T*)nullptr;
  float v[RC_GROUP_ROWS][4];
  load_quad(qrow, 4 * lane, D, vq, v[0]);
#pragma unroll
  for (int r = 0; r < RECHECK_GROUP; ++r) load_quad(xrow[r], 4 * lane, D, vg, v[1 + r]);
  double acc = 0.0;
  for (int k0 = 0; k0 < D; k0 += RC1_KC) {
    __syncwarp();  // the walkers have finished the previous chunk
#pragma unroll
    for (int r = 0; r < RC_GROUP_ROWS; ++r)
      *reinterpret_cast<float4*>(st + r * RC1_LD + 4 * lane) = make_float4(v[r][0], v[r][1], v[r][2], v[r][3]);
    __syncwarp();
    if (k0 + RC1_KC < D) {
      load_quad(qrow, k0 + RC1_KC + 4 * lane, D, vq, v[0]);
#pragma unroll
      for (int r = 0; r < RECHECK_GROUP; ++r) load_quad(xrow[r], k0 + RC1_KC + 4 * lane, D, vg, v[1 + r]);
    }
    if (lane < RECHECK_GROUP) {
      const int kn = min(RC1_KC, D - k0);
      const float* qs = st;
      const float* xs = st + (1 + lane) * RC1_LD;
      int k = 0;
      for (; k + 4 <= kn; k += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + k);
        const float4 x4 = *reinterpret_cast<const float4*>(xs + k);
        acc = fma((double)q4.x, (double)x4.x, acc);
        acc = fma((double)q4.y, (double)x4.y, acc);
        acc = fma((double)q4.z, (double)x4.z, acc);
        acc = fma((double)q4.w, (double)x4.w, acc);
      }
      for (; k < kn; ++k) acc = fma((double)qs[k], (double)xs[k], acc);
    }
  }
  const int64_t jl = j0 + lane;
  if (lane < RECHECK_GROUP && ((cmask >> lane) & 1u) && jl < M) {
    const int64_t g = gt ? gt[t] : t + row_offset;
    const int64_t jg = jl + col_offset;
    if (jg != g) {
      const double d = metric == VTC_METRIC_L2 ? sq64[jl] - 2.0 * acc : -acc;
      const double d0 = dgt[t];
      if ((d < d0) || (d == d0 && jg < g)) atomicAdd(&rank[t], 1);
    }
  }
}

// All list segments, dealt to the `num_warps` warps of the launch (`gw` = this warp's index):
// warps gw, gw + nseg, ... share segment gw % nseg.  `st` is this warp's RC_WARP_FLOATS staging tile.
template <typename T>
__device__ __forceinline__ void recheck_all(
    float* __restrict__ st, int gw, int num_warps, const int2* __restrict__ list,
    const unsigned int* __restrict__ seg_count, int nseg, unsigned int seg_cap,
    const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G, int64_t ldg,
    const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N, int64_t M, int D,
    const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset, int metric,
    int* __restrict__ rank, unsigned int* __restrict__ overflow) {
  if (nseg <= 0) return;
  const int lane = threadIdx.x & 31;
  const bool shared_segs = num_warps >= nseg;
  for (int seg = shared_segs ? gw % nseg : gw; seg < nseg; seg += shared_segs ? nseg : num_warps) {
    const unsigned int first = shared_segs ? (unsigned int)(gw / nseg) : 0u;
    const unsigned int step = shared_segs ? (unsigned int)((num_warps - seg + nseg - 1) / nseg) : 1u;
    const unsigned int n = seg_count[seg];
    if (n > seg_cap) {
      if (first == 0 && lane == 0) *overflow = 1u;
      continue;  // the brute-force fallback recomputes everything
    }
    const int2* seg_list = list + (size_t)seg * seg_cap;
    if (n < 2u * RC_GROUPS * step) {
      // a few groups per warp at most: the variant with the shorter critical path
      for (unsigned int u = first; u < n; u += step) {
        recheck_group_warp<T>(st, seg_list[u], Q, ldq, G, ldg, sq64, dgt, N, M, D, gt, row_offset,
                              col_offset, metric, rank);
      }
      continue;
    }
    for (unsigned int u = first * RC_GROUPS; u < n; u += step * RC_GROUPS) {
      const unsigned int mine = u + (lane >> 3);
      const int2 e = mine < n ? seg_list[mine] : make_int2(0, 0);
      recheck_groups_warp<T>(st, e, mine < n, Q, ldq, G, ldg, sq64, dgt, N, M, D, gt, row_offset, col_offset,
                             metric, rank);
    }
  }
}

}  // namespace vtc
