// rank_stage.cu -- everything a retrieval evaluation does around its tensor-core pass, in two
// launches (replaces the faiss index build / search bookkeeping and the Python hit loop of
// model/metric.py:140-160; round 1 spent 16 small launches on it).
//
// rank_prologue_kernel -- one pass over the gallery rows and one over the query rows:
//   * the K-major bf16 tensor-core operands (plain, or the 3-term split of VTC_PREC_EXACT),
//   * canonical ||x_j||^2 (fp64-sequential, exact.cu), the fp32 epilogue bias padded with +inf, the
//     largest finite norm (guard band),
//   * canonical d(t, gt) and an upper bound of ||q_t||^2 (guard band).
//   Rows are staged through shared memory 32 rows x 128 columns at a time: the whole block loads
//   (128-bit, coalesced), converts and stores the operands, while ONE warp walks the 32 staged rows,
//   a lane per row, accumulating in k order -- the fp64-sequential definition needs one thread per
//   dot product, the memory system needs a warp per row; the staging gives both.  HBM-bound:
//   rows * D * sizeof(in) read + rows * K' * 2 written.
//
// rank_epilogue_kernel -- cooperative (grid-wide barriers), persistent:
//   re-check of the guard-band groups -> [flag: zero + brute-force] -> commit into rank0 ->
//   [NaN ground truth -> M, R@K hit counts, median rank by radix select].
#include "rank_stage.cuh"

#include <cooperative_groups.h>

#include "exact_dev.cuh"
#include "prep.cuh"

namespace cg = cooperative_groups;

namespace vtc {

// ------------------------------------------------------------------------------------ prologue
constexpr int PR_ROWS = 32;
constexpr int PR_KC = 128;
constexpr int PR_LD = PR_KC + 4;  // floats per staged row: 16-byte aligned rows, conflict-free LDS.128
constexpr int PR_THREADS = 256;
constexpr int PR_LIGHT_ROWS = 256;

__device__ __forceinline__ float bf16_rn(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
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

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
         ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// the operand columns of four consecutive input elements (canonical values v, already rounded in the
// plain mode): mode PREP_PLAIN [x], PREP_SPLIT_A [hi | hi | lo], PREP_SPLIT_B [hi | lo | hi]
__device__ __forceinline__ void emit_quad(__nv_bfloat16* o, int k, int D, int mode,
                                          const float (&v)[4], unsigned int* fallback) {
  if (k >= D) return;
  if (mode == PREP_PLAIN) {
    if (k + 4 <= D) {
      *reinterpret_cast<uint2*>(o + k) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
    } else {
      for (int e = 0; e < 4 && k + e < D; ++e) o[k + e] = __float2bfloat16_rn(v[e]);
    }
    return;
  }
  float hi[4], lo[4];
  bool bad = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hi[e] = bf16_rn(v[e]);
    lo[e] = v[e] - hi[e];
    // inf / NaN (or an fp32 value that rounds to inf) has no 3-term split: x - hi is NaN
    bad |= !(fabsf(hi[e]) <= 3.0e38f);
  }
  if (bad && fallback) *fallback = 1u;
  const int o1 = mode == PREP_SPLIT_A ? D : 2 * D;  // second copy of hi
  const int o2 = mode == PREP_SPLIT_A ? 2 * D : D;  // lo
  if (k + 4 <= D && (D & 3) == 0) {
    const uint2 h = make_uint2(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]));
    *reinterpret_cast<uint2*>(o + k) = h;
    *reinterpret_cast<uint2*>(o + o1 + k) = h;
    *reinterpret_cast<uint2*>(o + o2 + k) = make_uint2(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]));
  } else {
    for (int e = 0; e < 4 && k + e < D; ++e) {
      const __nv_bfloat16 h = __float2bfloat16_rn(hi[e]);
      o[k + e] = h;
      o[o1 + k + e] = h;
      o[o2 + k + e] = __float2bfloat16_rn(lo[e]);
    }
  }
}

template <typename T>
__device__ __forceinline__ bool rows_vectorisable(const T* base, int64_t ld) {
  constexpr uintptr_t kAlign = sizeof(T) == 4 ? 15 : 7;  // 4 elements
  return (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(base) & kAlign) == 0;
}

// zero the padding columns [used, Kp) of the 32 operand rows of this block
__device__ __forceinline__ void zero_pad_columns(__nv_bfloat16* op, int Kp, int used,
                                                 const int64_t* srow) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (used >= Kp) return;
  for (int i = warp; i < PR_ROWS; i += PR_THREADS / 32) {
    const int64_t r = srow[i];
    if (r < 0) continue;
    for (int k = used + lane; k < Kp; k += 32) op[r * (int64_t)Kp + k] = __float2bfloat16_rn(0.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(PR_THREADS)
rank_prologue_kernel(const RankPrologueArgs a, int g_blocks, int g_light) {
  __shared__ __align__(16) float tile_a[PR_ROWS * PR_LD];  // gallery rows / query rows
  __shared__ __align__(16) float tile_b[PR_ROWS * PR_LD];  // ground-truth gallery rows of the queries
  __shared__ int64_t srow_a[PR_ROWS], srow_b[PR_ROWS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* Qb = static_cast<const T*>(a.Q);
  const T* Gb = static_cast<const T*>(a.G);
  const bool round = a.round_bf16 != 0;

  if ((int)blockIdx.x < g_blocks) {
    // ------------------------------------------------------------------------ gallery rows
    if (g_light) {
      // norms given, operands alias the input: only the padded fp32 bias and the largest norm
      const int64_t j = (int64_t)blockIdx.x * PR_LIGHT_ROWS + tid;
      float mine = 0.f;
      if (j < a.Mpad) {
        float b = INFINITY;
        if (j < a.M) {
          const float f = (float)a.sq64_in[j];
          b = a.metric == VTC_METRIC_L2 ? f : 0.f;
          if (f == f && f < 3.0e38f) mine = f;  // NaN / inf rows do not scale the guard band
        }
        a.bias[j] = b;
      }
      mine = warp_max(mine);
      if (lane == 0 && mine > 0.f) atomicMax(a.max_sq_bits, __float_as_uint(mine));
      return;  // (the light path is only taken with bias and max_sq_bits given)
    }
    const int64_t r0 = (int64_t)blockIdx.x * PR_ROWS;
    if (tid < PR_ROWS) srow_a[tid] = r0 + tid < a.M ? r0 + tid : -1;
    __syncthreads();
    const bool walk = a.sq64_in == nullptr;
    const bool emit = a.mode_g != STAGE_NONE;
    double sq = 0.0;
    if ((walk || emit) && r0 < a.M) {
      const bool vec = rows_vectorisable(Gb, a.ldg);
      const int nchunks = ceil_div(a.D, PR_KC);
      float v[4][4];
      const T* rp[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t r = srow_a[warp + 8 * j];
        rp[j] = r >= 0 ? Gb + r * a.ldg : nullptr;
        load_quad(rp[j], 4 * lane, a.D, vec, v[j]);
      }
      for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * PR_KC;
        __syncthreads();  // the walker warp has finished the previous chunk
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = warp + 8 * j;
          if (round) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[j][e] = bf16_rn(v[j][e]);
          }
          *reinterpret_cast<float4*>(&tile_a[row * PR_LD + 4 * lane]) =
              make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
          if (emit && rp[j])
            emit_quad(a.opG + srow_a[row] * (int64_t)a.Kp, k0 + 4 * lane, a.D, a.mode_g, v[j],
                      a.fallback);
        }
        __syncthreads();
        if (c + 1 < nchunks) {
#pragma unroll
          for (int j = 0; j < 4; ++j) load_quad(rp[j], k0 + PR_KC + 4 * lane, a.D, vec, v[j]);
        }
        if (walk && warp == 0) {
          const int kn = min(PR_KC, a.D - k0);
          const float* row = &tile_a[lane * PR_LD];
          for (int k = 0; k < kn; k += 4) {
            const float4 x = *reinterpret_cast<const float4*>(row + k);
            const double x0 = x.x, x1 = x.y, x2 = x.z, x3 = x.w;
            sq = fma(x0, x0, sq);
            sq = fma(x1, x1, sq);
            sq = fma(x2, x2, sq);
            sq = fma(x3, x3, sq);
          }
        }
      }
      if (emit) zero_pad_columns(a.opG, a.Kp, a.mode_g == PREP_PLAIN ? a.D : 3 * a.D, srow_a);
    }
    if (warp == 0) {
      const int64_t j = r0 + lane;
      float mine = 0.f;
      if (j < a.Mpad) {
        float b = INFINITY;
        if (j < a.M) {
          const double s64 = walk ? sq : a.sq64_in[j];
          if (walk && a.sq64) a.sq64[j] = s64;
          const float f = (float)s64;
          b = a.metric == VTC_METRIC_L2 ? f : 0.f;
          if (f == f && f < 3.0e38f) mine = f;
        }
        if (a.bias) a.bias[j] = b;
      }
      if (a.max_sq_bits) {
        mine = warp_max(mine);
        if (lane == 0 && mine > 0.f) atomicMax(a.max_sq_bits, __float_as_uint(mine));
      }
    }
    return;
  }

  // -------------------------------------------------------------------------- query rows
  const int64_t t0 = (int64_t)(blockIdx.x - g_blocks) * PR_ROWS;
  const bool need_gt = a.gt_in == nullptr && a.dgt != nullptr;
  const bool need_qq = a.qq_in == nullptr && a.qq != nullptr;
  const bool emit = a.mode_q != STAGE_NONE;
  if (tid < PR_ROWS) {
    const int64_t t = t0 + tid;
    srow_a[tid] = t < a.N ? t : -1;
    int64_t g = -1;
    if (t < a.N && need_gt) {
      g = (a.gt ? a.gt[t] : t + a.row_offset) - a.col_offset;
      if (g < 0 || g >= a.M) g = -1;
    }
    srow_b[tid] = g;
  }
  __syncthreads();
  double dot = 0.0, sqx = 0.0, qq = 0.0;
  {
    const bool vq = rows_vectorisable(Qb, a.ldq), vg = rows_vectorisable(Gb, a.ldg);
    const int nchunks = ceil_div(a.D, PR_KC);
    float v[4][4], w[4][4];
    const T *qp[4], *gp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t t = srow_a[warp + 8 * j], g = srow_b[warp + 8 * j];
      qp[j] = t >= 0 ? Qb + t * a.ldq : nullptr;
      gp[j] = g >= 0 ? Gb + g * a.ldg : nullptr;
      load_quad(qp[j], 4 * lane, a.D, vq, v[j]);
      if (need_gt) load_quad(gp[j], 4 * lane, a.D, vg, w[j]);
    }
    for (int c = 0; c < nchunks; ++c) {
      const int k0 = c * PR_KC;
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int row = warp + 8 * j;
        if (round) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[j][e] = bf16_rn(v[j][e]);
            if (need_gt) w[j][e] = bf16_rn(w[j][e]);
          }
        }
        *reinterpret_cast<float4*>(&tile_a[row * PR_LD + 4 * lane]) =
            make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
        if (need_gt)
          *reinterpret_cast<float4*>(&tile_b[row * PR_LD + 4 * lane]) =
              make_float4(w[j][0], w[j][1], w[j][2], w[j][3]);
        if (emit && qp[j])
          emit_quad(a.opQ + srow_a[row] * (int64_t)a.Kp, k0 + 4 * lane, a.D, a.mode_q, v[j],
                    a.fallback);
      }
      __syncthreads();
      if (c + 1 < nchunks) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          load_quad(qp[j], k0 + PR_KC + 4 * lane, a.D, vq, v[j]);
          if (need_gt) load_quad(gp[j], k0 + PR_KC + 4 * lane, a.D, vg, w[j]);
        }
      }
      if (warp == 0 && (need_gt || need_qq)) {
        const int kn = min(PR_KC, a.D - k0);
        const float* qr = &tile_a[lane * PR_LD];
        const float* xr = &tile_b[lane * PR_LD];
        if (need_gt) {
          for (int k = 0; k < kn; k += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(qr + k);
            const float4 x4 = *reinterpret_cast<const float4*>(xr + k);
            const double q0 = q4.x, q1 = q4.y, q2 = q4.z, q3 = q4.w;
            const double x0 = x4.x, x1 = x4.y, x2 = x4.z, x3 = x4.w;
            dot = fma(q0, x0, dot), sqx = fma(x0, x0, sqx), qq = fma(q0, q0, qq);
            dot = fma(q1, x1, dot), sqx = fma(x1, x1, sqx), qq = fma(q1, q1, qq);
            dot = fma(q2, x2, dot), sqx = fma(x2, x2, sqx), qq = fma(q2, q2, qq);
            dot = fma(q3, x3, dot), sqx = fma(x3, x3, sqx), qq = fma(q3, q3, qq);
          }
        } else {
          for (int k = 0; k < kn; k += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(qr + k);
            const double q0 = q4.x, q1 = q4.y, q2 = q4.z, q3 = q4.w;
            qq = fma(q0, q0, qq), qq = fma(q1, q1, qq), qq = fma(q2, q2, qq), qq = fma(q3, q3, qq);
          }
        }
      }
    }
    if (emit) zero_pad_columns(a.opQ, a.Kp, a.mode_q == PREP_PLAIN ? a.D : 3 * a.D, srow_a);
  }
  if (warp == 0 && t0 + lane < a.N) {
    const int64_t t = t0 + lane;
    if (need_gt) {
      double d0 = nan("");
      if (srow_b[lane] >= 0) d0 = a.metric == VTC_METRIC_L2 ? sqx - 2.0 * dot : -dot;
      a.dgt[t] = d0;
    }
    // the guard band only needs an UPPER bound of ||q||^2: round up, one part in 10^6 of slack
    if (need_qq) a.qq[t] = __double2float_ru(qq * (1.0 + 1.0e-6));
  }
}

int launch_rank_prologue(const RankPrologueArgs& a, cudaStream_t s) {
  if (a.N <= 0 && a.M <= 0) return VTC_OK;
  const bool g_work = a.sq64_in == nullptr || a.mode_g != STAGE_NONE;
  const int g_light = g_work ? 0 : 1;
  const int64_t g_blocks = a.Mpad > 0 ? ceil_div<int64_t>(a.Mpad, g_light ? PR_LIGHT_ROWS : PR_ROWS) : 0;
  const bool q_work = (a.gt_in == nullptr && a.dgt != nullptr) ||
                      (a.qq_in == nullptr && a.qq != nullptr) || a.mode_q != STAGE_NONE;
  const int64_t q_blocks = q_work ? ceil_div<int64_t>(a.N, PR_ROWS) : 0;
  if (g_blocks + q_blocks <= 0) return VTC_OK;
  if (g_blocks + q_blocks > 0x7fffffff) return VTC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)(g_blocks + q_blocks);
  if (a.in_bf16)
    rank_prologue_kernel<__nv_bfloat16><<<grid, PR_THREADS, 0, s>>>(a, (int)g_blocks, g_light);
  else
    rank_prologue_kernel<float><<<grid, PR_THREADS, 0, s>>>(a, (int)g_blocks, g_light);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

// ------------------------------------------------------------------------------------ epilogue
constexpr int EP_THREADS = 256;
constexpr int MED_BINS = 2048;
constexpr int MED_LEVELS = 3;  // bits [21,32), [10,21), [0,10)
size_t rank_epilogue_hist_words() { return (size_t)MED_LEVELS * 2 * MED_BINS + 8; }

__device__ __forceinline__ int med_shift(int level) { return level == 0 ? 21 : (level == 1 ? 10 : 0); }
__device__ __forceinline__ unsigned int med_mask(int level) { return level == 2 ? 1023u : 2047u; }

union EpilogueSmem {
  BruteSmem brute;
  unsigned int hist[2][MED_BINS];
};

// the bin (and the remainder inside it) that holds order statistic `target` of a 2048-bin
// histogram; every thread of the block returns the same pair
// (h was filled by other blocks' atomics in this same launch: read it through L2, not the
// non-coherent path)
__device__ __forceinline__ uint2 med_pick(const unsigned int* h, unsigned int target,
                                          unsigned int* sh_warp, unsigned int* sh_out) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  constexpr int PER = MED_BINS / EP_THREADS;  // 8 bins per thread
  unsigned int c[PER], tot = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    c[i] = __ldcg(h + tid * PER + i);
    tot += c[i];
  }
  unsigned int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncthreads();  // sh_warp / sh_out may still be read from a previous call
  if (lane == 31) sh_warp[w] = incl;
  __syncthreads();
  unsigned int base = incl - tot;
  for (int i = 0; i < w; ++i) base += sh_warp[i];
  if (target >= base && target < base + tot) {
    unsigned int run = base;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (target >= run && target < run + c[i]) {
        sh_out[0] = (unsigned int)(tid * PER + i);
        sh_out[1] = target - run;
      }
      run += c[i];
    }
  }
  __syncthreads();
  return make_uint2(sh_out[0], sh_out[1]);
}

template <typename T>
__global__ void __launch_bounds__(EP_THREADS)
rank_epilogue_kernel(const RankEpilogueArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ EpilogueSmem sm;
  __shared__ unsigned int sh_warp[EP_THREADS / 32], sh_out[2];
  __shared__ int sh_hits[8];
  const int tid = threadIdx.x;
  const int64_t gtid = (int64_t)blockIdx.x * EP_THREADS + tid;
  const int64_t gthreads = (int64_t)gridDim.x * EP_THREADS;
  const T* Q = static_cast<const T*>(a.ex.Q);
  const T* G = static_cast<const T*>(a.ex.G);
  const int64_t N = a.ex.N;

  // ---- phase 1: scratch of the finalisation + re-check of the guard-band groups
  if (a.finalize) {
    const int64_t words = (int64_t)MED_LEVELS * 2 * MED_BINS + 8;
    for (int64_t i = gtid; i < words; i += gthreads) a.hist[i] = 0u;
    if (blockIdx.x == 0 && tid < a.nk) a.hits[tid] = 0ull;
  }
  // (rank_tmp == NULL: finalisation only -- vtc_rank_finalize on ranks that are already complete)
  if (a.rank_tmp)
    for (int vb = blockIdx.x; vb < a.nseg * RECHECK_PARTS; vb += gridDim.x)
      recheck_part<T>(vb, a.amb_list, a.seg_count, a.seg_cap, Q, a.ex.ldq, G, a.ex.ldg, a.ex.sq64,
                      a.dgt, N, a.ex.M, a.ex.D, a.ex.gt, a.ex.row_offset, a.ex.col_offset,
                      a.ex.metric, a.rank_tmp, a.fallback);
  grid.sync();

  // ---- phase 2 (rare): the list overflowed, or the split operands could not carry the inputs:
  // recount everything in canonical arithmetic
  if (a.rank_tmp && *reinterpret_cast<volatile unsigned int*>(a.fallback) != 0u) {
    for (int64_t i = gtid; i < N; i += gthreads) a.rank_tmp[i] = 0;
    grid.sync();
    rank_brute_tiles<T>(sm.brute, Q, a.ex.ldq, G, a.ex.ldg, a.ex.sq64, a.dgt, N, a.ex.M, a.ex.D,
                        a.ex.gt, a.ex.row_offset, a.ex.col_offset, a.ex.metric, a.rank_tmp,
                        blockIdx.x, gridDim.x);
    grid.sync();
  }

  // ---- phase 3: commit (+ NaN ground truth -> M_total, R@K hit counts, first histogram level)
  if (!a.finalize) {
    for (int64_t i = gtid; i < N; i += gthreads)
      a.rank0[i] = (a.accumulate ? a.rank0[i] : 0) + a.rank_tmp[i];
    return;
  }
  const bool want_med = a.medr != nullptr;
  const int level0 = (a.M_total >> 21) == 0 ? ((a.M_total >> 10) == 0 ? 2 : 1) : 0;
  for (int i = tid; i < 2 * MED_BINS; i += EP_THREADS) (&sm.hist[0][0])[i] = 0u;
  if (tid < 8) sh_hits[tid] = 0;
  __syncthreads();
  {
    int local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int shift = med_shift(level0);
    const unsigned int mask = med_mask(level0);
    for (int64_t i = gtid; i < N; i += gthreads) {
      int r = (a.accumulate ? a.rank0[i] : 0) + (a.rank_tmp ? a.rank_tmp[i] : 0);
      if (a.dgt) {
        const double d0 = a.dgt[i];
        if (d0 != d0) r = (int)a.M_total;  // no ground truth: never retrieved
      }
      a.rank0[i] = r;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < a.nk) local[k] += r < a.k_vals[k];
      if (want_med) atomicAdd(&sm.hist[0][((unsigned int)r >> shift) & mask], 1u);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k < a.nk) {
        int v = local[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0 && v) atomicAdd(&sh_hits[k], v);
      }
    }
  }
  __syncthreads();
  if (tid < a.nk && sh_hits[tid]) atomicAdd(&a.hits[tid], (unsigned long long)sh_hits[tid]);
  if (!want_med) return;
  if (N <= 0) {
    if (gtid == 0) *a.medr = nan("");
    return;
  }
  // ---- median(rank0) + 1 with numpy semantics: radix select over 11/11/10-bit digits, both middle
  // order statistics at once (they may part ways at any level).  Every block derives the same
  // (prefix, remainder) state from the global histogram, so one grid barrier per level suffices.
  unsigned int pa = 0, pb = 0, ra = (unsigned int)((N - 1) / 2), rb = (unsigned int)(N / 2);
  for (int level = level0; level < MED_LEVELS; ++level) {
    unsigned int* h = a.hist + (size_t)level * 2 * MED_BINS;
    if (level > level0) {
      // histogram of this level's digit among the ranks that match the prefix selected so far
      for (int i = tid; i < 2 * MED_BINS; i += EP_THREADS) (&sm.hist[0][0])[i] = 0u;
      __syncthreads();
      const int shift = med_shift(level), up = med_shift(level - 1);
      const unsigned int mask = med_mask(level);
      for (int64_t i = gtid; i < N; i += gthreads) {
        const unsigned int r = (unsigned int)a.rank0[i];
        const unsigned int hi = r >> up, bin = (r >> shift) & mask;
        if (hi == pa) atomicAdd(&sm.hist[0][bin], 1u);
        if (hi == pb && pb != pa) atomicAdd(&sm.hist[1][bin], 1u);
      }
      __syncthreads();
    }
    for (int i = tid; i < 2 * MED_BINS; i += EP_THREADS) {
      const unsigned int v = (&sm.hist[0][0])[i];
      if (v) atomicAdd(&h[i], v);
    }
    grid.sync();
    const bool same = pa == pb;
    const uint2 sa = med_pick(h, ra, sh_warp, sh_out);
    const uint2 sb = med_pick(same ? h : h + MED_BINS, rb, sh_warp, sh_out);
    const int bits = level == 2 ? 10 : 11;
    pa = (pa << bits) | sa.x, ra = sa.y;
    pb = (pb << bits) | sb.x, rb = sb.y;
  }
  if (gtid == 0) *a.medr = 0.5 * ((double)pa + (double)pb) + 1.0;
}

template <typename T>
static int launch_rank_epilogue_t(const RankEpilogueArgs& a, cudaStream_t s) {
  static int blocks_per_sm = []() {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, rank_epilogue_kernel<T>, EP_THREADS, 0) !=
        cudaSuccess)
      return 0;
    return n;
  }();
  if (blocks_per_sm <= 0) return VTC_ERR_NO_DEVICE;
  const int grid = kNumSMs * (blocks_per_sm < 2 ? blocks_per_sm : 2);
  void* args[] = {const_cast<RankEpilogueArgs*>(&a)};
  const cudaError_t e = cudaLaunchCooperativeKernel((const void*)rank_epilogue_kernel<T>, dim3(grid),
                                                    dim3(EP_THREADS), args, 0, s);
  if (e != cudaSuccess) return cuda_err(e);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_rank_epilogue(const RankEpilogueArgs& a, cudaStream_t s) {
  if (a.ex.N <= 0) return VTC_OK;
  return a.ex.bf16 ? launch_rank_epilogue_t<__nv_bfloat16>(a, s) : launch_rank_epilogue_t<float>(a, s);
}

}  // namespace vtc
