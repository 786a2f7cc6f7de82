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
// A warp stages the rows of a step 32 columns at a time in a warp-private shared-memory tile (128-bit
// loads, 8 lanes per row segment, the next chunk already in flight) and its lanes walk one fp64 chain
// each in k order.  A thread walking two rows on its own (round 1) is a chain of D/16 dependent L2
// round trips (~20 us at D = 512); here the chain itself (D dependent DFMAs, ~3 us) is what is left
// -- this sits on the critical path of every chunked call.
constexpr int RECHECK_GROUP = 8;
constexpr int RC_GROUPS = 4;                       // (sizes the staging tile: 36 rows)
constexpr int RC_KC = 32;                          // staged columns per step
constexpr int RC_LD = RC_KC + 4;                   // floats per staged row: 16-byte aligned rows, and
                                                   // 8 consecutive rows cover all 32 banks (LDS.128)
constexpr int RC_GROUP_ROWS = RECHECK_GROUP + 1;   // the query row + the group's gallery rows
constexpr int RC_ROWS = RC_GROUPS * RC_GROUP_ROWS; // 36
constexpr int RC_WARP_FLOATS = RC_ROWS * RC_LD;       // 1296 floats

// SIXTEEN (row, column) pairs at a time.  A listed group typically has ONE in-band column, so with a
// lane per column of four groups (the earlier round-2 layout) ~5 of the 32 chains of a step were
// real.  Here a warp takes up to 32 list entries (a lane each), expands their masks into (t, j) pairs
// without a buffer -- pair p of the batch is the (p - off[b])-th set lane of ballot b, bits first --
// and decides 16 pairs per step: 16 query rows + 16 gallery rows staged 32 columns at a time by the
// whole warp exactly as above, lanes 0..15 walk one pair's fp64 chain each.  3x fewer steps for the
// same list (100k x 100k: 141 -> 67 us bf16, 322 -> 81 us exact).
constexpr int RP_PAIRS = 16;
static_assert(2 * RP_PAIRS * RC_LD <= RC_WARP_FLOATS, "the staging tile serves the pair variant");

template <typename T>
__device__ __forceinline__ void recheck_pairs_warp(
    float* __restrict__ st, int2 e, bool has, const T* __restrict__ Q, int64_t ldq, const T* __restrict__ G,
    int64_t ldg, const double* __restrict__ sq64, const double* __restrict__ dgt, int64_t N,
    int64_t M, int D, const int64_t* __restrict__ gt, int64_t row_offset, int64_t col_offset,
    int metric, int* __restrict__ rank) {
  const int lane = threadIdx.x & 31;
  unsigned int cmask = 0;
  int t_ = -1, j_ = 0;
  if (has) {
    tc::amb_unpack(e, &t_, &j_, &cmask);
    if (t_ >= N || j_ >= M) cmask = 0;  // zero-padded tile rows / columns
    if (cmask) {
      const int64_t room = M - (int64_t)j_;  // columns of the group inside this gallery chunk
      if (room < RECHECK_GROUP) cmask &= (1u << room) - 1u;
      const int64_t g = (gt ? gt[t_] : (int64_t)t_ + row_offset) - col_offset - (int64_t)j_;
      if (g >= 0 && g < RECHECK_GROUP) cmask &= ~(1u << g);  // the ground truth is no competitor
    }
  }
  unsigned int bal[RECHECK_GROUP];
  int off[RECHECK_GROUP + 1];
  off[0] = 0;
#pragma unroll
  for (int b = 0; b < RECHECK_GROUP; ++b) {
    bal[b] = __ballot_sync(0xffffffffu, (cmask >> b) & 1u);
    off[b + 1] = off[b] + __popc(bal[b]);
  }
  const int total = off[RECHECK_GROUP];
  const bool vq = rows_vectorisable(Q, ldq), vg = rows_vectorisable(G, ldg);
  const int kq = 4 * (lane & 7);
  for (int base = 0; base < total; base += RP_PAIRS) {
    // lanes 0..15: pair base + lane -> (source lane, bit)
    int pt = -1, pj = 0;
    {
      const int p = base + (lane & (RP_PAIRS - 1));
      int b = 0, srcl = 0;
      if (p < total) {
#pragma unroll
        for (int bb = 1; bb < RECHECK_GROUP; ++bb)
          if (p >= off[bb]) b = bb;
        unsigned int bm = bal[0];
        int ob = off[0];
#pragma unroll
        for (int bb = 1; bb < RECHECK_GROUP; ++bb)
          if (b == bb) bm = bal[bb], ob = off[bb];
        srcl = __fns(bm, 0, p - ob + 1);
      }
      const int st_ = __shfl_sync(0xffffffffu, t_, srcl);
      const int sj_ = __shfl_sync(0xffffffffu, j_, srcl);
      if (p < total) pt = st_, pj = sj_ + b;
    }
    // loader view: lane owns column quad (lane % 8) of staged rows lane / 8 + 4 i, i = 0..7;
    // rows 0..15 = the pairs' query rows, rows 16..31 = their gallery rows
    const T* src[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = (lane >> 3) + 4 * i;
      const int pr = row & (RP_PAIRS - 1);
      const int tt = __shfl_sync(0xffffffffu, pt, pr);
      const int jj = __shfl_sync(0xffffffffu, pj, pr);
      src[i] = tt < 0 ? (const T*)nullptr
                      : (row < RP_PAIRS ? Q + (int64_t)tt * ldq : G + (int64_t)jj * ldg);
    }
    float v[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) load_quad(src[i], kq, D, (i < 4) ? vq : vg, v[i]);
    // walker view: lane p < 16 walks pair p
    const float* qs = st + (lane & (RP_PAIRS - 1)) * RC_LD;
    const float* xs = qs + RP_PAIRS * RC_LD;
    double acc = 0.0;
    for (int k0 = 0; k0 < D; k0 += RC_KC) {
      __syncwarp();  // the walkers have finished the previous chunk
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(st + ((lane >> 3) + 4 * i) * RC_LD + kq) =
            make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
      __syncwarp();
      if (k0 + RC_KC < D) {
#pragma unroll
        for (int i = 0; i < 8; ++i) load_quad(src[i], k0 + RC_KC + kq, D, (i < 4) ? vq : vg, v[i]);
      }
      if (lane < RP_PAIRS) {
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
    }
    if (lane < RP_PAIRS && pt >= 0) {
      const int64_t t = pt, jl = pj;
      const int64_t g = gt ? gt[t] : t + row_offset;
      const int64_t jg = jl + col_offset;
      const double d = metric == VTC_METRIC_L2 ? sq64[jl] - 2.0 * acc : -acc;
      const double d0 = dgt[t];
      if ((d < d0) || (d == d0 && jg < g)) atomicAdd(&rank[t], 1);
    }
    __syncwarp();
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
    int* __restrict__ rank) {
  if (nseg <= 0) return;
  const int lane = threadIdx.x & 31;
  const bool shared_segs = num_warps >= nseg;
  for (int seg = shared_segs ? gw % nseg : gw; seg < nseg; seg += shared_segs ? nseg : num_warps) {
    const unsigned int first = shared_segs ? (unsigned int)(gw / nseg) : 0u;
    const unsigned int step = shared_segs ? (unsigned int)((num_warps - seg + nseg - 1) / nseg) : 1u;
    const unsigned int n = seg_count[seg];
    if (n > seg_cap) continue;  // (an overflowed segment routes the whole call to the canonical
                                // recount before any re-check starts: rank_needs_fallback)
    const int2* seg_list = list + (size_t)seg * seg_cap;
    const unsigned int per_warp = (n + step - 1u) / step;  // entries per warp of this segment
    if (per_warp <= 2u) {
      // a couple of groups per warp at most: the variant with the shorter critical path
      for (unsigned int u = first; u < n; u += step) {
        recheck_group_warp<T>(st, seg_list[u], Q, ldq, G, ldg, sq64, dgt, N, M, D, gt, row_offset,
                              col_offset, metric, rank);
      }
      continue;
    }
    // batches of up to 32 consecutive entries per warp, sized so that every warp of the segment
    // gets work; each batch is decided 16 (row, column) pairs at a time
    const unsigned int batch = per_warp < 32u ? per_warp : 32u;
    for (unsigned int u = first * batch; u < n; u += step * batch) {
      const unsigned int mine = u + (unsigned int)lane;
      const bool has = (unsigned int)lane < batch && mine < n;
      const int2 e = has ? seg_list[mine] : make_int2(0, 0);
      recheck_pairs_warp<T>(st, e, has, Q, ldq, G, ldg, sq64, dgt, N, M, D, gt, row_offset, col_offset,
                            metric, rank);
    }
  }
}

}  // namespace vtc
