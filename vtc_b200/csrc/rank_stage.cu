// rank_stage.cu -- everything a retrieval evaluation does around its tensor-core pass, in two
// launches (replaces the faiss index build / search bookkeeping and the Python hit loop of
// model/metric.py:140-160; round 1 spent 16 small launches on it).
//
// rank_prologue_kernel -- one pass over the gallery rows and one over the query rows:
//   * the K-major bf16 tensor-core operands (plain, or the 3-term split of VTC_PREC_EXACT),
//   * canonical ||x_j||^2 (fp64-sequential, exact.cu), the fp32 epilogue bias padded with +inf, the
//     largest finite norm (guard band),
//   * canonical d(t, gt) and an upper bound of ||q_t||^2 (guard band).
//   Rows are staged through shared memory 128 rows x 32 columns at a time: the whole block loads
//   (128-bit; the 8 lanes of a row segment read 128 contiguous bytes), converts and stores the
//   operands, and 128 threads walk one staged row each, accumulating in k order -- the
//   fp64-sequential definition needs one thread per dot product, the memory system needs several
//   lanes per row; the staging gives both, and every row of the block has its own walker (round 2a
//   walked 32 rows per block with one warp: 2.8x off the HBM bound).  HBM-bound:
//   rows * D * sizeof(in) read + rows * K' * 2 written.
//
// the epilogue chain -- see "epilogue" below:
//   re-check of the guard-band groups (or, rarely, the canonical recount of the whole call) -> commit
//   into rank0 -> [NaN ground truth -> M, R@K hit counts, median rank by radix select].
#include "rank_stage.cuh"

#include "exact_dev.cuh"
#include "prep.cuh"

namespace vtc {

// ------------------------------------------------------------------------------------ prologue
constexpr int PR_ROWS = 128;      // rows per block: one walker thread per row
constexpr int PR_KC = 32;         // staged columns per step
constexpr int PR_LD = PR_KC + 4;  // floats per staged row: 16-byte aligned rows, and 8 consecutive rows
                                  // cover all 32 banks (conflict-free LDS.128 for the walkers)
constexpr int PR_THREADS = 256;
constexpr int PR_Q = PR_ROWS * (PR_KC / 4) / PR_THREADS;  // quads a thread loads per tile and step (4)
constexpr int PR_LIGHT_ROWS = 256;

__device__ __forceinline__ float bf16_rn(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// an upper bound of a sum of squares accumulated in fp32 in any order (D <= 8192 terms: the
// relative rounding error is below 8192 * 2^-24 = 4.9e-4)
__device__ __forceinline__ float piece_up(float s) { return s * 1.001f; }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
         ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

// the operand columns of four consecutive input elements (canonical values v, already rounded in the
// plain mode): mode PREP_PLAIN [x], PREP_SPLIT_A [hi | hi | lo], PREP_SPLIT_B [hi | lo | hi]
// (split modes: *s_lo / *s_e gain the squared pieces lo = bf16(x - hi) and e = x - hi - lo)
__device__ __forceinline__ void emit_quad(__nv_bfloat16* o, int k, int D, int mode,
                                          const float (&v)[4], unsigned int* fallback,
                                          float* s_lo = nullptr, float* s_e = nullptr) {
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
    if (s_lo && k + e < D) {
      const float lb = bf16_rn(lo[e]), ee = lo[e] - lb;  // (x - hi and lo - bf16(lo) are exact in fp32)
      *s_lo = fmaf(lb, lb, *s_lo);
      *s_e = fmaf(ee, ee, *s_e);
    }
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
rank_prologue_kernel(const RankPrologueArgs a, int g_blocks, int g_light, int paired) {
  __shared__ __align__(16) float tile_a[PR_ROWS * PR_LD];  // gallery rows / query rows
  __shared__ __align__(16) float tile_b[PR_ROWS * PR_LD];  // ground-truth gallery rows of the queries
  __shared__ int64_t srow_a[PR_ROWS], srow_b[PR_ROWS];
  __shared__ float blk_max[PR_THREADS / 32][3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  griddep_launch();
  griddep_wait();
  const T* Qb = static_cast<const T*>(a.Q);
  const T* Gb = static_cast<const T*>(a.G);
  const bool round = a.round_bf16 != 0;
  // loader view: thread owns column quad (tid % 8) of staged rows tid / 8 + 32 i, i = 0..3 -- the 8
  // lanes of a row segment read 128 contiguous bytes (fp32) per step
  const int lrow = tid >> 3, kq = 4 * (tid & 7);
  const int nchunks = ceil_div(a.D, PR_KC);

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
    const bool pieces = emit && a.mode_g != PREP_PLAIN && a.split_max_bits != nullptr;
    float s_lo[PR_Q], s_e[PR_Q];
#pragma unroll
    for (int i = 0; i < PR_Q; ++i) s_lo[i] = s_e[i] = 0.f;
    double sq = 0.0;
    if ((walk || emit) && r0 < a.M) {
      const bool vec = rows_vectorisable(Gb, a.ldg);
      float v[PR_Q][4];
      const T* rp[PR_Q];
#pragma unroll
      for (int i = 0; i < PR_Q; ++i) {
        const int64_t r = srow_a[lrow + 32 * i];
        rp[i] = r >= 0 ? Gb + r * a.ldg : nullptr;
        load_quad(rp[i], kq, a.D, vec, v[i]);
      }
      for (int c = 0; c < nchunks; ++c) {
        const int k0 = c * PR_KC;
        __syncthreads();  // the walkers have finished the previous chunk
#pragma unroll
        for (int i = 0; i < PR_Q; ++i) {
          const int row = lrow + 32 * i;
          if (round) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[i][e] = bf16_rn(v[i][e]);
          }
          *reinterpret_cast<float4*>(&tile_a[row * PR_LD + kq]) =
              make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
          if (emit && rp[i])
            emit_quad(a.opG + srow_a[row] * (int64_t)a.Kp, k0 + kq, a.D, a.mode_g, v[i], a.fallback,
                      pieces ? &s_lo[i] : nullptr, pieces ? &s_e[i] : nullptr);
        }
        __syncthreads();
        if (c + 1 < nchunks) {
#pragma unroll
          for (int i = 0; i < PR_Q; ++i) load_quad(rp[i], k0 + PR_KC + kq, a.D, vec, v[i]);
        }
        if (walk && tid < PR_ROWS) {
          const int kn = min(PR_KC, a.D - k0);
          const float* row = &tile_a[tid * PR_LD];
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
    // per-row results (walker threads) and the block's maxima: one set of global atomics per block
    float m_sq = 0.f, m_lo = 0.f, m_e = 0.f;
    if (tid < PR_ROWS) {
      const int64_t j = r0 + tid;
      if (j < a.Mpad) {
        float b = INFINITY;
        if (j < a.M) {
          const double s64 = walk ? sq : a.sq64_in[j];
          if (walk && a.sq64) a.sq64[j] = s64;
          const float f = (float)s64;
          b = a.metric == VTC_METRIC_L2 ? f : 0.f;
          if (f == f && f < 3.0e38f) m_sq = f;
        }
        if (a.bias) a.bias[j] = b;
      }
    }
    if (pieces) {
#pragma unroll
      for (int i = 0; i < PR_Q; ++i) {
        // a row's 8 quads per step live in 8 consecutive lanes
        float l = s_lo[i], e = s_e[i];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          l += __shfl_xor_sync(0xffffffffu, l, o);
          e += __shfl_xor_sync(0xffffffffu, e, o);
        }
        l = piece_up(l), e = piece_up(e);
        if (l < 3.0e38f) m_lo = fmaxf(m_lo, l);  // (NaN / inf rows raise the fallback flag instead)
        if (e < 3.0e38f) m_e = fmaxf(m_e, e);
      }
    }
    m_sq = warp_max(m_sq), m_lo = warp_max(m_lo), m_e = warp_max(m_e);
    if (lane == 0) blk_max[warp][0] = m_sq, blk_max[warp][1] = m_lo, blk_max[warp][2] = m_e;
    __syncthreads();
    if (tid < 3) {
      float m = 0.f;
      for (int w = 0; w < PR_THREADS / 32; ++w) m = fmaxf(m, blk_max[w][tid]);
      unsigned int* dst = tid == 0 ? a.max_sq_bits : (pieces ? a.split_max_bits + (tid - 1) : nullptr);
      if (dst && m > 0.f) atomicMax(dst, __float_as_uint(m));
    }
    return;
  }

  // -------------------------------------------------------------------------- query rows
  const int64_t t0 = (int64_t)(blockIdx.x - g_blocks) * PR_ROWS;
  const bool need_gt = a.gt_in == nullptr && a.dgt != nullptr;
  const bool need_qq = a.qq_in == nullptr && a.qq != nullptr;
  const bool emit = a.mode_q != STAGE_NONE;
  const bool pieces = emit && a.mode_q != PREP_PLAIN && a.qsplit != nullptr;
  // paired rows (gt(t) = t over the whole gallery, norms not given): the ground-truth row of query t
  // IS gallery row t, which this block stages anyway -- it also emits that row's operand and takes
  // its canonical norm (the sqx chain below is the gallery walker's chain, same order), so the
  // gallery is read once instead of twice and there are no gallery blocks at all
  const bool emit_g = paired != 0 && a.mode_g != STAGE_NONE;
  const bool pieces_g = emit_g && a.mode_g != PREP_PLAIN && a.split_max_bits != nullptr;
  float s_lo[PR_Q], s_e[PR_Q], g_lo[PR_Q], g_e[PR_Q];
#pragma unroll
  for (int i = 0; i < PR_Q; ++i) s_lo[i] = s_e[i] = g_lo[i] = g_e[i] = 0.f;
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
    float v[PR_Q][4], w[PR_Q][4];
    const T *qp[PR_Q], *gp[PR_Q];
#pragma unroll
    for (int i = 0; i < PR_Q; ++i) {
      const int64_t t = srow_a[lrow + 32 * i], g = srow_b[lrow + 32 * i];
      qp[i] = t >= 0 ? Qb + t * a.ldq : nullptr;
      gp[i] = g >= 0 ? Gb + g * a.ldg : nullptr;
      load_quad(qp[i], kq, a.D, vq, v[i]);
      if (need_gt) load_quad(gp[i], kq, a.D, vg, w[i]);
    }
    for (int c = 0; c < nchunks; ++c) {
      const int k0 = c * PR_KC;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < PR_Q; ++i) {
        const int row = lrow + 32 * i;
        if (round) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[i][e] = bf16_rn(v[i][e]);
            if (need_gt) w[i][e] = bf16_rn(w[i][e]);
          }
        }
        *reinterpret_cast<float4*>(&tile_a[row * PR_LD + kq]) =
            make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
        if (need_gt)
          *reinterpret_cast<float4*>(&tile_b[row * PR_LD + kq]) =
              make_float4(w[i][0], w[i][1], w[i][2], w[i][3]);
        if (emit && qp[i])
          emit_quad(a.opQ + srow_a[row] * (int64_t)a.Kp, k0 + kq, a.D, a.mode_q, v[i], a.fallback,
                    pieces ? &s_lo[i] : nullptr, pieces ? &s_e[i] : nullptr);
        if (emit_g && gp[i])
          emit_quad(a.opG + srow_b[row] * (int64_t)a.Kp, k0 + kq, a.D, a.mode_g, w[i], a.fallback,
                    pieces_g ? &g_lo[i] : nullptr, pieces_g ? &g_e[i] : nullptr);
      }
      __syncthreads();
      if (c + 1 < nchunks) {
#pragma unroll
        for (int i = 0; i < PR_Q; ++i) {
          load_quad(qp[i], k0 + PR_KC + kq, a.D, vq, v[i]);
          if (need_gt) load_quad(gp[i], k0 + PR_KC + kq, a.D, vg, w[i]);
        }
      }
      if (tid < PR_ROWS && (need_gt || need_qq)) {
        const int kn = min(PR_KC, a.D - k0);
        const float* qr = &tile_a[tid * PR_LD];
        const float* xr = &tile_b[tid * PR_LD];
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
    if (emit_g) zero_pad_columns(a.opG, a.Kp, a.mode_g == PREP_PLAIN ? a.D : 3 * a.D, srow_b);
    if (pieces) {
#pragma unroll
      for (int i = 0; i < PR_Q; ++i) {
        float l = s_lo[i], e = s_e[i];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          l += __shfl_xor_sync(0xffffffffu, l, o);
          e += __shfl_xor_sync(0xffffffffu, e, o);
        }
        const int64_t t = srow_a[lrow + 32 * i];
        if ((tid & 7) == 0 && t >= 0)
          a.qsplit[t] = make_float2(sqrtf(piece_up(l)) * 1.000001f, sqrtf(piece_up(e)) * 1.000001f);
      }
    }
  }
  if (tid < PR_ROWS && t0 + tid < a.N) {
    const int64_t t = t0 + tid;
    if (need_gt) {
      double d0 = nan("");
      if (srow_b[tid] >= 0) d0 = a.metric == VTC_METRIC_L2 ? sqx - 2.0 * dot : -dot;
      a.dgt[t] = d0;
    }
    // the guard band only needs an UPPER bound of ||q||^2: round up, one part in 10^6 of slack
    if (need_qq) a.qq[t] = __double2float_ru(qq * (1.0 + 1.0e-6));
  }
  if (paired) {
    // the gallery side of these rows: canonical norm, epilogue bias (the last block also pads it
    // with +inf up to Mpad), and the block's maxima -- as the gallery blocks do it
    float m_sq = 0.f, m_lo = 0.f, m_e = 0.f;
    if (tid < PR_ROWS && srow_b[tid] >= 0) {
      const int64_t j = srow_b[tid];
      if (a.sq64) a.sq64[j] = sqx;
      const float f = (float)sqx;
      if (a.bias) a.bias[j] = a.metric == VTC_METRIC_L2 ? f : 0.f;
      if (f == f && f < 3.0e38f) m_sq = f;
    }
    if (a.bias && t0 + PR_ROWS >= a.M)
      for (int64_t j = a.M + tid; j < a.Mpad; j += PR_THREADS) a.bias[j] = INFINITY;
    if (pieces_g) {
#pragma unroll
      for (int i = 0; i < PR_Q; ++i) {
        float l = g_lo[i], e = g_e[i];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          l += __shfl_xor_sync(0xffffffffu, l, o);
          e += __shfl_xor_sync(0xffffffffu, e, o);
        }
        l = piece_up(l), e = piece_up(e);
        if (l < 3.0e38f) m_lo = fmaxf(m_lo, l);
        if (e < 3.0e38f) m_e = fmaxf(m_e, e);
      }
    }
    m_sq = warp_max(m_sq), m_lo = warp_max(m_lo), m_e = warp_max(m_e);
    if (lane == 0) blk_max[warp][0] = m_sq, blk_max[warp][1] = m_lo, blk_max[warp][2] = m_e;
    __syncthreads();
    if (tid < 3) {
      float m = 0.f;
      for (int w = 0; w < PR_THREADS / 32; ++w) m = fmaxf(m, blk_max[w][tid]);
      unsigned int* dst = tid == 0 ? a.max_sq_bits : (pieces_g ? a.split_max_bits + (tid - 1) : nullptr);
      if (dst && m > 0.f) atomicMax(dst, __float_as_uint(m));
    }
  }
}

int launch_rank_prologue(const RankPrologueArgs& a, cudaStream_t s) {
  if (a.N <= 0 && a.M <= 0) return VTC_OK;
  const bool g_work = a.sq64_in == nullptr || a.mode_g != STAGE_NONE;
  const int g_light = g_work ? 0 : 1;
  // paired rows: every query's ground truth is the gallery row of the same index and that covers
  // the whole gallery (one evaluation, gt(t) = t): the query blocks do the gallery side as well
  const int paired = (a.sq64_in == nullptr && a.gt_in == nullptr && a.dgt != nullptr &&
                      a.gt == nullptr && a.row_offset == a.col_offset && a.N == a.M && a.N > 0 &&
                      a.sq64 != nullptr && a.bias != nullptr && a.max_sq_bits != nullptr) ? 1 : 0;
  const int64_t g_blocks = (paired || a.Mpad <= 0) ? 0 : ceil_div<int64_t>(a.Mpad, g_light ? PR_LIGHT_ROWS : PR_ROWS);
  const bool q_work = (a.gt_in == nullptr && a.dgt != nullptr) ||
                      (a.qq_in == nullptr && a.qq != nullptr) || a.mode_q != STAGE_NONE;
  const int64_t q_blocks = q_work ? ceil_div<int64_t>(a.N, PR_ROWS) : 0;
  if (g_blocks + q_blocks <= 0) return VTC_OK;
  if (g_blocks + q_blocks > 0x7fffffff) return VTC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)(g_blocks + q_blocks);
  if (a.in_bf16)
    launch_pdl(rank_prologue_kernel<__nv_bfloat16>, dim3(grid), dim3(PR_THREADS), 0, s, a,
               (int)g_blocks, g_light, paired);
  else
    launch_pdl(rank_prologue_kernel<float>, dim3(grid), dim3(PR_THREADS), 0, s, a, (int)g_blocks,
               g_light, paired);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

// ------------------------------------------------------------------------------------ epilogue
// A chain of short plain launches under programmatic dependent launch (round 2a used one cooperative
// launch with grid-wide barriers; a kernel boundary under PDL is cheaper than a grid barrier plus the
// cooperative launch, and every stage gets the grid size that suits it):
//   rank_recheck_kernel   zero the finalisation scratch; every warp re-checks its share of the
//                         guard-band groups in canonical arithmetic.  If the call has to fall back
//                         (a list segment overflowed, or split operands could not carry inf / NaN --
//                         every block of the chain decides that on its own from the list counts and
//                         the prologue's flag, rank_needs_fallback) the grid recounts the whole call
//                         in canonical arithmetic into a spare count array instead; until the last
//                         session of round 2 that was a launch of its own that exited at once
//   rank_commit_kernel    rank0 = (accumulate ? rank0 : 0) + counts (the spare array after a
//                         fallback); with `finalize`: NaN ground truth -> M_total, R@K hit counts,
//                         histogram of the first radix digit
//   rank_select_kernel    (per further digit) pick the digit(s) of the two middle order statistics,
//                         histogram of the next digit among the ranks that match; the LAST block of
//                         the last level writes median(rank0) + 1 (numpy semantics).
constexpr int EP_THREADS = 256;
constexpr int MED_BINS = 2048;
constexpr int MED_LEVELS = 3;  // bits [21,32), [10,21), [0,10)
constexpr int EP_ITEMS = 4;    // ranks per thread in the commit / select passes
size_t rank_epilogue_hist_words() { return (size_t)MED_LEVELS * 2 * MED_BINS + 8; }

__device__ __forceinline__ int med_shift(int level) { return level == 0 ? 21 : (level == 1 ? 10 : 0); }
__device__ __forceinline__ unsigned int med_mask(int level) { return level == 2 ? 1023u : 2047u; }
__host__ __device__ inline int med_level0(int64_t M_total) {
  return (M_total >> 21) == 0 ? ((M_total >> 10) == 0 ? 2 : 1) : 0;
}

union __align__(16) EpilogueSmem {
  BruteSmem brute;
  unsigned int hist[2][MED_BINS];
  float recheck[EP_THREADS / 32][RC_WARP_FLOATS];  // one staging tile per warp
};

// the bin (and the remainder inside it) that holds order statistic `target` of a 2048-bin
// histogram in global memory (filled by other blocks' atomics: read through L2); every thread of
// the block returns the same pair
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

// true in every thread of exactly one block of the grid: the last one to get here.  All global
// writes of the other blocks made before their call are visible to it afterwards.
__device__ __forceinline__ bool last_block_here(unsigned int* ticket, unsigned int* sh_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *sh_flag = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  const bool last = *sh_flag != 0u;
  if (last) __threadfence();
  return last;
}

// hist layout: [level][2][MED_BINS] counters, then 8 words: tickets of the select levels
__device__ __forceinline__ unsigned int* hist_ticket(unsigned int* hist, int level) {
  return hist + (size_t)MED_LEVELS * 2 * MED_BINS + level;
}

// Does this call have to be recounted in canonical arithmetic?  Decided identically by every block of
// every launch of the chain from data that is complete before the chain starts: the prologue's flag
// (split operands that cannot carry inf / NaN) and the tensor-core pass's per-CTA list counts (a
// segment that overflowed its capacity).  No launch has to wait for another one's verdict.
__device__ __forceinline__ bool rank_needs_fallback(const RankEpilogueArgs& a) {
  int over = __ldcg(a.fallback) != 0u ? 1 : 0;
  for (int i = threadIdx.x; i < a.nseg; i += blockDim.x) over |= __ldcg(a.seg_count + i) > a.seg_cap ? 1 : 0;
  return __syncthreads_or(over) != 0;
}

template <typename T>
__global__ void __launch_bounds__(EP_THREADS)
rank_recheck_kernel(const RankEpilogueArgs a) {
  __shared__ EpilogueSmem sm;
  griddep_launch();
  griddep_wait();
  const int tid = threadIdx.x;
  if (a.finalize) {
    const int64_t words = (int64_t)(MED_LEVELS * 2 * MED_BINS + 8);
    for (int64_t i = (int64_t)blockIdx.x * EP_THREADS + tid; i < words; i += (int64_t)gridDim.x * EP_THREADS)
      a.hist[i] = 0u;
    if (blockIdx.x == 0 && tid < a.nk) a.hits[tid] = 0ull;
  }
  if (!a.rank_tmp) return;  // finalisation only (vtc_rank_finalize): nothing to re-check
  if (rank_needs_fallback(a)) {
    // rare (adversarial duplicates overflowing the list, inf / NaN rows in the exact mode): the whole
    // grid recounts the call in canonical arithmetic into the spare count array, which the commit
    // launch then takes instead of the tensor-core counts
    rank_brute_tiles<T>(sm.brute, static_cast<const T*>(a.ex.Q), a.ex.ldq, static_cast<const T*>(a.ex.G),
                        a.ex.ldg, a.ex.sq64, a.dgt, a.ex.N, a.ex.M, a.ex.D, a.ex.gt, a.ex.row_offset,
                        a.ex.col_offset, a.ex.metric, a.rank_alt, blockIdx.x, gridDim.x);
    return;
  }
  recheck_all<T>(sm.recheck[tid >> 5], (int)blockIdx.x * (EP_THREADS / 32) + (tid >> 5),
                 (int)gridDim.x * (EP_THREADS / 32), a.amb_list, a.seg_count, a.nseg, a.seg_cap,
                 static_cast<const T*>(a.ex.Q), a.ex.ldq, static_cast<const T*>(a.ex.G), a.ex.ldg,
                 a.ex.sq64, a.dgt, a.ex.N, a.ex.M, a.ex.D, a.ex.gt, a.ex.row_offset,
                 a.ex.col_offset, a.ex.metric, a.rank_tmp);
}

// the median's radix-select state after the levels before `level`: prefixes (pa, pb) and remaining
// order statistics (ra, rb) of the lower and upper middle element
struct MedState {
  unsigned int pa, pb, ra, rb;
};
__device__ __forceinline__ MedState med_state_before(const unsigned int* hist, int level0, int level,
                                                     int64_t N, unsigned int* sh_warp,
                                                     unsigned int* sh_out) {
  MedState m{0u, 0u, (unsigned int)((N - 1) / 2), (unsigned int)(N / 2)};
  for (int l = level0; l < level; ++l) {
    const unsigned int* h = hist + (size_t)l * 2 * MED_BINS;
    const bool same = m.pa == m.pb;
    const uint2 sa = med_pick(h, m.ra, sh_warp, sh_out);
    const uint2 sb = med_pick(same ? h : h + MED_BINS, m.rb, sh_warp, sh_out);
    const int bits = l == 2 ? 10 : 11;
    m.pa = (m.pa << bits) | sa.x, m.ra = sa.y;
    m.pb = (m.pb << bits) | sb.x, m.rb = sb.y;
  }
  return m;
}

__global__ void __launch_bounds__(EP_THREADS)
rank_commit_kernel(const RankEpilogueArgs a) {
  __shared__ unsigned int hist[MED_BINS];
  __shared__ unsigned int sh_warp[EP_THREADS / 32], sh_out[2], sh_flag;
  __shared__ int sh_hits[8];
  griddep_launch();
  griddep_wait();
  const int tid = threadIdx.x;
  const int64_t N = a.ex.N;
  const int64_t i0 = ((int64_t)blockIdx.x * EP_THREADS + tid) * EP_ITEMS;
  // the counts of this call: the tensor-core pass + re-check, or the canonical recount
  const int* counts = a.rank_tmp ? (rank_needs_fallback(a) ? a.rank_alt : a.rank_tmp) : nullptr;
  if (!a.finalize) {
#pragma unroll
    for (int e = 0; e < EP_ITEMS; ++e)
      if (i0 + e < N) a.rank0[i0 + e] = (a.accumulate ? a.rank0[i0 + e] : 0) + __ldcg(counts + i0 + e);
    return;
  }
  const bool want_med = a.medr != nullptr;
  const int level0 = med_level0(a.M_total);
  if (want_med)
    for (int i = tid; i < MED_BINS; i += EP_THREADS) hist[i] = 0u;
  if (tid < 8) sh_hits[tid] = 0;
  __syncthreads();
  int r[EP_ITEMS];
#pragma unroll
  for (int e = 0; e < EP_ITEMS; ++e) {
    r[e] = -1;
    if (i0 + e < N) {
      r[e] = (a.accumulate ? a.rank0[i0 + e] : 0) + (counts ? __ldcg(counts + i0 + e) : 0);
      if (a.dgt) {
        const double d0 = a.dgt[i0 + e];
        if (d0 != d0) r[e] = (int)a.M_total;  // no ground truth: never retrieved
      }
    }
  }
  int local[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int shift = med_shift(level0);
  const unsigned int mask = med_mask(level0);
#pragma unroll
  for (int e = 0; e < EP_ITEMS; ++e) {
    if (i0 + e >= N) continue;
    a.rank0[i0 + e] = r[e];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < a.nk) local[k] += r[e] < a.k_vals[k];
    if (want_med) atomicAdd(&hist[((unsigned int)r[e] >> shift) & mask], 1u);
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
  __syncthreads();
  if (tid < a.nk && sh_hits[tid]) atomicAdd(&a.hits[tid], (unsigned long long)sh_hits[tid]);
  if (!want_med) return;
  unsigned int* h = a.hist + (size_t)level0 * 2 * MED_BINS;
  for (int i = tid; i < MED_BINS; i += EP_THREADS)
    if (hist[i]) atomicAdd(&h[i], hist[i]);
  if (level0 < MED_LEVELS - 1) return;  // further digits: rank_select_kernel
  // a single digit (M_total < 1024): the last block picks the median here
  if (!last_block_here(hist_ticket(a.hist, level0), &sh_flag)) return;
  const MedState m = med_state_before(a.hist, level0, MED_LEVELS, N, sh_warp, sh_out);
  if (tid == 0) *a.medr = 0.5 * ((double)m.pa + (double)m.pb) + 1.0;
}

// one further radix digit (`level` > level0)
__global__ void __launch_bounds__(EP_THREADS)
rank_select_kernel(const RankEpilogueArgs a, int level) {
  __shared__ unsigned int hist[2][MED_BINS];
  __shared__ unsigned int sh_warp[EP_THREADS / 32], sh_out[2], sh_flag;
  griddep_launch();
  griddep_wait();
  const int tid = threadIdx.x;
  const int64_t N = a.ex.N;
  const int level0 = med_level0(a.M_total);
  for (int i = tid; i < 2 * MED_BINS; i += EP_THREADS) (&hist[0][0])[i] = 0u;
  const MedState m = med_state_before(a.hist, level0, level, N, sh_warp, sh_out);  // (syncs inside)
  const int64_t i0 = ((int64_t)blockIdx.x * EP_THREADS + tid) * EP_ITEMS;
  const int shift = med_shift(level), up = med_shift(level - 1);
  const unsigned int mask = med_mask(level);
#pragma unroll
  for (int e = 0; e < EP_ITEMS; ++e) {
    if (i0 + e >= N) continue;
    const unsigned int r = (unsigned int)a.rank0[i0 + e];
    const unsigned int hi = r >> up, bin = (r >> shift) & mask;
    if (hi == m.pa) atomicAdd(&hist[0][bin], 1u);
    if (hi == m.pb && m.pb != m.pa) atomicAdd(&hist[1][bin], 1u);
  }
  __syncthreads();
  unsigned int* h = a.hist + (size_t)level * 2 * MED_BINS;
  for (int i = tid; i < 2 * MED_BINS; i += EP_THREADS) {
    const unsigned int v = (&hist[0][0])[i];
    if (v) atomicAdd(&h[i], v);
  }
  if (level < MED_LEVELS - 1) return;
  if (!last_block_here(hist_ticket(a.hist, level), &sh_flag)) return;
  const MedState f = med_state_before(a.hist, level0, MED_LEVELS, N, sh_warp, sh_out);
  if (tid == 0) *a.medr = 0.5 * ((double)f.pa + (double)f.pb) + 1.0;
}

template <typename T>
static int launch_rank_epilogue_t(const RankEpilogueArgs& a, cudaStream_t s) {
  const int64_t N = a.ex.N;
  if (a.finalize && !a.hist) return VTC_ERR_WORKSPACE;
  if (a.rank_tmp && (!a.rank_alt || !a.fallback || !a.seg_count)) return VTC_ERR_INVALID_ARG;
  if (a.rank_tmp || a.finalize) {
    // finalisation only: the scratch is 12 K words -- a handful of blocks zero it
    launch_pdl(rank_recheck_kernel<T>, dim3(a.rank_tmp ? 2 * kNumSMs : 8), dim3(EP_THREADS), 0, s, a);
    VTC_LAUNCH_CHECK();
  }
  const unsigned blocks = (unsigned)ceil_div<int64_t>(N, EP_THREADS * EP_ITEMS);
  launch_pdl(rank_commit_kernel, dim3(blocks), dim3(EP_THREADS), 0, s, a);
  VTC_LAUNCH_CHECK();
  if (a.finalize && a.medr) {
    for (int level = med_level0(a.M_total) + 1; level < MED_LEVELS; ++level) {
      launch_pdl(rank_select_kernel, dim3(blocks), dim3(EP_THREADS), 0, s, a, level);
      VTC_LAUNCH_CHECK();
    }
  }
  return VTC_OK;
}

int launch_rank_epilogue(const RankEpilogueArgs& a, cudaStream_t s) {
  if (a.ex.N <= 0) return VTC_OK;
  return a.ex.bf16 ? launch_rank_epilogue_t<__nv_bfloat16>(a, s) : launch_rank_epilogue_t<float>(a, s);
}

}  // namespace vtc
