// sim_tc_kernel.cuh -- the similarity GEMM of the retrieval hot path on the 5th-gen tensor cores.
//
//   S[t, j] = scale * <A_t, B_j> + bias_j        A: queries [N, K'] bf16, B: gallery [M, K'] bf16
//
// replaces `feats_a @ feats_b.t()` (model/model.py:369,478,504,621) and the SGEMM inside
// faiss.GpuIndexFlatL2.search (model/metric.py:144-146).  S is never written to HBM: each
// 128 x 256 fp32 tile lives in TMEM and is consumed by a fused epilogue:
//   EPI_RANK   count columns whose score beats the ground-truth score (guard-banded; ambiguous
//              pairs go to a list that exact.cu re-checks in fp64)        -> rank / R@K / MedR
//   EPI_LSE    online row log-sum-exp + diagonal                           -> symmetric InfoNCE
//   EPI_TOPK   streaming top-k candidate pool per row                      -> faiss-style search
//   EPI_STORE  materialise (bias / QuickGELU / residual)                   -> sim tensor, linears
//
// Structure (one CTA per SM, persistent over work items, 384 threads):
//   warp 0    TMA producer  : cp.async.bulk.tensor 2-D loads, 128B-swizzled K-major tiles
//   warp 1    MMA issuer    : one thread issues tcgen05.mma (M128 x N256 x K16 per CTA, or one
//                             M256 cta_group::2 MMA per CTA pair; bf16 -> fp32 in TMEM)
//   warp 2    TMEM allocator: 512 columns = 2 accumulator stages of 256
//   warp 4-11 epilogue      : tcgen05.ld 32x32b (thread = tile row; two warps per scheduler, each
//                             one column half of the tile), fused reduction
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// With `kRes` the 128 x K' query tile stays resident in shared memory for a whole work item
// (K' <= 512), so only gallery tiles stream from L2: 64 B/clk/SM instead of 96.
//
// Work item = (query tile, gallery split); items are ordered so that CTAs running concurrently
// sweep the same gallery range (B tiles are shared through the 126 MB L2).
#pragma once
#include <cuda_bf16.h>

#include <climits>
#include <cstring>

#include "ptx.cuh"
#include "sim_tc.cuh"

namespace vtc {
namespace tc {

using namespace ptx;

constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_WARPS_C = 8;
constexpr int SMEM_LIMIT = 232448;         // 227 KB per CTA
constexpr int MAX_RES_KB = 8;              // resident A: up to K' = 512
constexpr int NUM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;      // two per scheduler: each TMEM lane quarter is split in two column halves
constexpr int TMEM_COLS = 512;

// kPair: the two CTAs of a cluster issue one M256 cta_group::2 MMA; each CTA stages only ITS half
// of every gallery tile (16 KB), so the ring is twice as deep in the same shared memory.
// kBN: gallery rows per tile (UMMA N, TMEM columns per accumulator stage): 256, or 128 for small
// dense products that need more CTAs.
// The ring takes what the resident query tile leaves of the 227 KB (at most 8 stages).
// kEpiStage: bytes of warp-private staging each epilogue warp gets behind the bias buffers (StoreEpi
// transposes its 32 x 32 blocks there so that global stores and residual loads are whole 128-byte rows).
template <bool kRes, bool kPair = false, int kBN = BN, int kEpiStage = 0>
struct SmemLayout {
  static constexpr int kResKb = MAX_RES_KB;
  static constexpr int kBBytes = (kBN * BK * 2) / (kPair ? 2 : 1);
  static constexpr int kStageBytes = kRes ? kBBytes : (A_TILE_BYTES + kBBytes);
  static constexpr int kResBytes = kRes ? kResKb * A_TILE_BYTES : 0;
  static constexpr int kFit =
      (SMEM_LIMIT - 256 - EPI_WARPS_C * (64 * 4 + kEpiStage) - kResBytes) / kStageBytes;
  static constexpr int kStages = kFit > 8 ? 8 : kFit;
  static constexpr int kStagesOff = kResBytes;
  static constexpr int kBarOff = kStagesOff + kStages * kStageBytes;
  static constexpr int kBiasOff = kBarOff + 256;  // 8 epilogue warps x 64 floats
  static constexpr int kNumBars = 2 * kStages + kResKb + 1 + 2 + 2;
  static constexpr int kEpiStageOff = kBiasOff + EPI_WARPS * 64 * 4;
  static constexpr int kTotal = kEpiStageOff + EPI_WARPS * kEpiStage;
  static_assert(kStages >= 3, "ring too shallow");
  static_assert(kNumBars * 8 + 8 <= 256, "barrier area");
  static_assert(kTotal <= SMEM_LIMIT, "exceeds 227 KB of shared memory");
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------ epilogues
// Each epilogue thread owns one tile row (= TMEM lane) and one column half of every tile.
// `begin_item(p, row, part)` / `chunk` / `end_item(p, part)` with part = 2 * split + half.
// `bias` points at 32 consecutive entries of the per-column bias staged in this WARP'S PRIVATE
// shared-memory buffer (128 columns at a time; the global array is padded to a multiple of 256
// columns by the host, out-of-range columns hold the epilogue's neutral value): reads are
// warp-uniform 128-bit LDS broadcasts and the epilogue warps never synchronise with each other.
// Code size matters here (one warp per scheduler, 32 KB of L1.5 I-cache): rare paths are kept tiny
// and out of line.

constexpr int RANK_GROUP = 8;  // columns re-checked together when any of them is in the guard band

struct RankEpi {
  static constexpr int kStageBytes = 0;
  __device__ __forceinline__ void bind_stage(float*) {}
  float lo, hi;
  int cnt;
  int gtc;  // ground-truth column of this row (launch-local), or a value no group can contain
  int64_t t;
  __device__ __forceinline__ void begin_item(const Params& p, int64_t t_, int) {
    t = t_;
    cnt = 0;
    lo = hi = nanf("");
    gtc = INT_MIN / 2;
    if (t < p.N) {
      band(p.dgt[t], p.qq[t], p.max_sq_bits, p.metric_l2, p.guard_rel,
           p.qsplit ? p.qsplit + t : nullptr, p.split_max_bits, &lo, &hi);
      const int64_t g = (p.gt ? p.gt[t] : t + p.gt_row_offset) - p.gt_col_offset;
      if (g >= 0 && g < p.M) gtc = (int)g;
    }
  }
  // once per row and work item (dozens of tiles): out of line, so that the double-precision square
  // roots stay out of the instruction-cache footprint of the per-column loop
  static __device__ __noinline__ void band(double d0, float qq, const unsigned int* max_sq_bits,
                                           int metric_l2, float guard_rel, const float2* qsplit,
                                           const unsigned int* split_max_bits, float* lo, float* hi) {
    const double gmax_sq = (double)__uint_as_float(*max_sq_bits);
    double split_abs = 0.0;
    if (qsplit) {
      const float2 qs = *qsplit;
      split_abs = rank_split_bound(sqrt((double)qq), (double)qs.x, (double)qs.y, sqrt(gmax_sq),
                                   sqrt((double)__uint_as_float(split_max_bits[0])),
                                   sqrt((double)__uint_as_float(split_max_bits[1])));
    }
    rank_band(d0, (double)qq, gmax_sq, metric_l2, guard_rel, split_abs, lo, hi);
  }
  // rare: this row has scores inside the guard band among columns [j, j+8): hand the group and the
  // mask of those columns to the fp64 re-check.  List segments are per CTA, slots come from a
  // shared-memory counter, so there is no global atomic hot spot.
  // (static + by-value arguments: a member function would force the epilogue state through
  // `this`, i.e. into local memory, on the hot path)
  static __device__ __noinline__ void push_group(int2* __restrict__ seg_list, unsigned int seg_cap,
                                                 unsigned int* seg_count, int t, int j0, int gtc,
                                                 float lo, float hi, float d0, float d1, float d2,
                                                 float d3, float d4, float d5, float d6, float d7) {
    // which columns score inside [lo, hi]: those -- and only those -- go to the fp64 re-check; the
    // ground truth's own column is in the band by construction and is not a competitor
    unsigned int mask = (unsigned int)(d0 <= hi && !(d0 < lo)) | ((unsigned int)(d1 <= hi && !(d1 < lo)) << 1) |
                        ((unsigned int)(d2 <= hi && !(d2 < lo)) << 2) | ((unsigned int)(d3 <= hi && !(d3 < lo)) << 3) |
                        ((unsigned int)(d4 <= hi && !(d4 < lo)) << 4) | ((unsigned int)(d5 <= hi && !(d5 < lo)) << 5) |
                        ((unsigned int)(d6 <= hi && !(d6 < lo)) << 6) | ((unsigned int)(d7 <= hi && !(d7 < lo)) << 7);
    if ((unsigned)(gtc - j0) < (unsigned)RANK_GROUP) mask &= ~(1u << (gtc - j0));
    if (mask == 0) return;
    const unsigned int slot = atomicAdd(seg_count, 1u);
    if (slot < seg_cap) seg_list[slot] = amb_pack(t, j0, mask);
  }
  __device__ __forceinline__ void chunk(const Params& p, const uint32_t (&v)[32],
                                        const float* __restrict__ bias, float scale, int64_t jbase,
                                        unsigned int* seg_count) {
    // counts kept as floats: (d < lo) ? 1.f : 0.f is one FSET, the add one FADD (exact, <= 8)
    float csum = 0.f;
#pragma unroll
    for (int g = 0; g < 32 / RANK_GROUP; ++g) {
      const float4 b0 = reinterpret_cast<const float4*>(bias)[2 * g];
      const float4 b1 = reinterpret_cast<const float4*>(bias)[2 * g + 1];
      const float d0 = fmaf(scale, __uint_as_float(v[8 * g + 0]), b0.x);
      const float d1 = fmaf(scale, __uint_as_float(v[8 * g + 1]), b0.y);
      const float d2 = fmaf(scale, __uint_as_float(v[8 * g + 2]), b0.z);
      const float d3 = fmaf(scale, __uint_as_float(v[8 * g + 3]), b0.w);
      const float d4 = fmaf(scale, __uint_as_float(v[8 * g + 4]), b1.x);
      const float d5 = fmaf(scale, __uint_as_float(v[8 * g + 5]), b1.y);
      const float d6 = fmaf(scale, __uint_as_float(v[8 * g + 6]), b1.z);
      const float d7 = fmaf(scale, __uint_as_float(v[8 * g + 7]), b1.w);
      const float lt = ((d0 < lo ? 1.f : 0.f) + (d1 < lo ? 1.f : 0.f)) +
                       ((d2 < lo ? 1.f : 0.f) + (d3 < lo ? 1.f : 0.f)) +
                       ((d4 < lo ? 1.f : 0.f) + (d5 < lo ? 1.f : 0.f)) +
                       ((d6 < lo ? 1.f : 0.f) + (d7 < lo ? 1.f : 0.f));
      const float le = ((d0 <= hi ? 1.f : 0.f) + (d1 <= hi ? 1.f : 0.f)) +
                       ((d2 <= hi ? 1.f : 0.f) + (d3 <= hi ? 1.f : 0.f)) +
                       ((d4 <= hi ? 1.f : 0.f) + (d5 <= hi ? 1.f : 0.f)) +
                       ((d6 <= hi ? 1.f : 0.f) + (d7 <= hi ? 1.f : 0.f));
      csum += lt;  // columns certainly below the band count here, whatever the rest of the group does
      if (le != lt)  // rare: some column of the group scores inside [lo, hi]
        push_group(p.amb_list + (size_t)blockIdx.x * p.amb_seg_cap, p.amb_seg_cap, seg_count, (int)t,
                   (int)jbase + RANK_GROUP * g, gtc, lo, hi, d0, d1, d2, d3, d4, d5, d6, d7);
    }
    cnt += (int)csum;
  }
  __device__ __forceinline__ void end_item(const Params& p, int) {
    if (t < p.N && cnt) atomicAdd(&p.rank[t], cnt);
  }
};

// Online row log-sum-exp (running max + rescaled sum, log2 domain) and -- in the SAME pass -- the
// column sums of the symmetric loss: sim^T is never formed and the product is not run a second time
// (model/loss.py:21 calls cross_entropy on sim and on sim.t()).  Columns cannot keep a running max
// across the row tiles of different CTAs, so every column j is summed against a fixed reference, its
// own positive logit col_ref[j]: the positive pair contributes 2^~0, so the sum cannot underflow, and
// it overflows only if some negative beats the positive by > 88 nats (caught by the merge kernel,
// which then gates a second pass on).  A warp holds a 32 x 32 block of 2^(x - ref): a butterfly
// transpose-reduce (31 shuffles) leaves lane i with the block's sum of column i, written once per
// (query tile, lane quarter, column) -- deterministic, no atomics.
struct LseEpi {
  static constexpr int kStageBytes = 0;
  __device__ __forceinline__ void bind_stage(float*) {}
  float m, l, dv;
  int64_t t, jd;
  float* cpart;  // this (query tile, lane quarter)'s row of col_part, or NULL
  bool rowvalid;
  __device__ __forceinline__ void begin_item(const Params& p, int64_t t_, int) {
    t = t_;
    m = -INFINITY;
    l = 0.f;
    dv = nanf("");
    jd = p.diag ? t + p.diag_offset : -1;
    rowvalid = t < p.N;
    cpart = p.col_part ? p.col_part + ((t / BM) * 4 + ((t % BM) >> 5)) * p.col_ld : nullptr;
  }
  __device__ __forceinline__ void chunk(const Params& p, const uint32_t (&v)[32],
                                        const float* __restrict__ bias, float scale, int64_t jbase,
                                        unsigned int*) {
    float x[32];
    float cmax = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = reinterpret_cast<const float4*>(bias)[i];
      x[4 * i + 0] = fmaf(scale, __uint_as_float(v[4 * i + 0]), b.x);
      x[4 * i + 1] = fmaf(scale, __uint_as_float(v[4 * i + 1]), b.y);
      x[4 * i + 2] = fmaf(scale, __uint_as_float(v[4 * i + 2]), b.z);
      x[4 * i + 3] = fmaf(scale, __uint_as_float(v[4 * i + 3]), b.w);
      cmax = fmaxf(cmax, fmaxf(fmaxf(x[4 * i], x[4 * i + 1]), fmaxf(x[4 * i + 2], x[4 * i + 3])));
    }
    if (cmax > -INFINITY) {
      const float mn = fmaxf(m, cmax);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s += ex2_approx(x[i] - mn);
      l = l * ex2_approx(m - mn) + s;
      m = mn;
    }
    if (jd >= jbase && jd < jbase + 32) {
      const int sel = (int)(jd - jbase);
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i == sel) dv = __uint_as_float(v[i]);
    }
    if (cpart) {
      float e[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(p.col_ref + jbase) + i);
        // zero-filled tile rows beyond N must not count; padding columns have x = -inf, ref = +inf
        e[4 * i + 0] = rowvalid ? ex2_approx(x[4 * i + 0] - r.x) : 0.f;
        e[4 * i + 1] = rowvalid ? ex2_approx(x[4 * i + 1] - r.y) : 0.f;
        e[4 * i + 2] = rowvalid ? ex2_approx(x[4 * i + 2] - r.z) : 0.f;
        e[4 * i + 3] = rowvalid ? ex2_approx(x[4 * i + 3] - r.w) : 0.f;
      }
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;  // this lane keeps the upper half of its values
#pragma unroll
        for (int k = 0; k < o; ++k) {
          const float send = up ? e[k] : e[k + o];
          const float keep = up ? e[k + o] : e[k];
          e[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      cpart[jbase + lane] = e[0];  // lane i: column jbase + i, summed over this warp's 32 rows
    }
  }
  __device__ __forceinline__ void end_item(const Params& p, int split) {
    if (t < p.N) {
      p.lse_part[(int64_t)split * p.N + t] = make_float2(m, l);
      if (p.diag && dv == dv) p.diag[t] = dv;
    }
  }
};

// Materialising epilogue.  A thread owns one tile ROW, so a warp's registers hold a 32 x 32 block
// row-per-lane, while memory wants whole 128-byte rows per instruction: the block goes through a
// warp-private 4 KB shared-memory tile (128-bit accesses, XOR-swizzled by row so that both the
// row-per-lane and the transposed accesses are bank-conflict free) on its way out -- and the residual
// on its way in.  Straight from registers a float4 store touched 32 different rows, i.e. 32
// transactions per instruction instead of 4.
struct StoreEpi {
  static constexpr int kStageBytes = 32 * 32 * 4;
  int64_t t, t0w;  // this thread's row, first row of the warp's 32-row block
  float* st;       // warp-private 32 x 32 fp32 staging tile
  float rstat, coef, dsum;  // act == 2: this row's log-sum-exp, grad_loss / 2n, sum_j z * acc
  __device__ __forceinline__ void bind_stage(float* s) { st = s; }
  __device__ __forceinline__ void begin_item(const Params& p, int64_t t_, int) {
    t = t_;
    t0w = t_ - (threadIdx.x & 31);
    dsum = 0.f;
    rstat = INFINITY;
    coef = 0.f;
    if (p.act == 2) {
      if (t < p.N) rstat = p.row_stat[t];  // (rows of the tile padding: exp(x - inf) = 0)
      coef = *p.coef_ptr * p.coef_scale;
    }
  }
  // float4 slot q (0..7) of staged row r lives at r * 32 + 4 * (q ^ (r & 7))
  static __device__ __forceinline__ int slot(int r, int q) { return r * 32 + 4 * (q ^ (r & 7)); }
  static __device__ __forceinline__ uint32_t pack2(float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
  }
  __device__ __forceinline__ void chunk(const Params& p, const uint32_t (&v)[32],
                                        const float* __restrict__ bias, float scale, int64_t jbase,
                                        unsigned int*) {
    const int lane = threadIdx.x & 31;
    if (t0w >= p.N || jbase >= p.M) return;  // warp-uniform: nothing of this block is in range
    const bool full = jbase + 32 <= p.M && t0w + 32 <= p.N;
    const bool vec = full && ((p.ldo & 3) == 0) &&
                     (!p.out || (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                     (!p.residual || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    const int rr = lane >> 3, q = lane & 7;  // transposed view: 4 rows x 8 float4 per instruction
    float y[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = reinterpret_cast<const float4*>(bias)[i];
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float z = fmaf(scale, __uint_as_float(v[4 * i + e]), bb[e]);
        // QuickGELU: z * sigmoid(1.702 z).  __fdividef (MUFU.RCP, 2 ulp): the IEEE division is a
        // ~40-instruction subroutine per element and made c_fc's epilogue 4x its main loop
        if (p.act == 1) z = __fdividef(z, 1.f + __expf(-1.702f * z));
        if (p.act == 2) {
          const float acc = __uint_as_float(v[4 * i + e]), x = scale * acc;
          // (the split operands carry fp32 logits: accurate exponentials there)
          const float er = p.out_op_split ? expf(x - rstat) : __expf(x - rstat);
          const float ec = p.out_op_split ? expf(x - bb[e]) : __expf(x - bb[e]);
          z = coef * (er + ec);
          if (t + p.diag_offset == jbase + 4 * i + e) z -= 2.f * coef;
          dsum = fmaf(z, acc, dsum);
        }
        y[4 * i + e] = z;
      }
    }
    if (p.residual) {
      if (vec) {
        // whole 128-byte rows in, four rows per instruction; each lane then reads its own row
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rr;
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(p.residual + (t0w + r) * p.ldo + jbase) + q);
          *reinterpret_cast<float4*>(st + slot(r, q)) = x4;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x4 = *reinterpret_cast<const float4*>(st + slot(lane, i));
          y[4 * i] += x4.x, y[4 * i + 1] += x4.y, y[4 * i + 2] += x4.z, y[4 * i + 3] += x4.w;
        }
      } else if (t < p.N) {
        const float* r = p.residual + t * p.ldo + jbase;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (jbase + i < p.M) y[i] += r[i];
      }
    }
    const bool ovec = p.out_op && full && ((p.M & 7) == 0) && ((p.out_op_kp & 7) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out_op) & 15) == 0);
    if ((p.out && vec) || ovec) {
      __syncwarp();  // the residual reads of the staging tile are complete
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(st + slot(lane, i)) =
            make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
      __syncwarp();
    }
    if (p.out) {
      if (vec) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + rr;
          *reinterpret_cast<float4*>(p.out + (t0w + r) * p.ldo + jbase + 4 * q) =
              *reinterpret_cast<const float4*>(st + slot(r, q));
        }
      } else if (t < p.N) {
        float* o = p.out + t * p.ldo + jbase;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (jbase + i < p.M) o[i] = y[i];
      }
    }
    if (p.out_op) {
      // the result as the query-side bf16 operand of the NEXT GEMM: [x] or the split [hi|hi|lo]
      if (ovec) {
        // 64-byte operand rows: 8 rows x 4 groups of 8 bf16 per instruction
        const int r8 = lane >> 2, g8 = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + r8;
          const float4 a4 = *reinterpret_cast<const float4*>(st + slot(r, 2 * g8));
          const float4 b4 = *reinterpret_cast<const float4*>(st + slot(r, 2 * g8 + 1));
          const float f[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
          const uint4 hi = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]),
                                      pack2(f[6], f[7]));
          __nv_bfloat16* oo = p.out_op + (t0w + r) * (int64_t)p.out_op_kp + jbase + 8 * g8;
          *reinterpret_cast<uint4*>(oo) = hi;
          if (p.out_op_split) {
            float l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) l[e] = f[e] - __bfloat162float(__float2bfloat16_rn(f[e]));
            *reinterpret_cast<uint4*>(oo + p.M) = hi;
            *reinterpret_cast<uint4*>(oo + 2 * p.M) =
                make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
          }
        }
      } else if (t < p.N) {
        __nv_bfloat16* oo = p.out_op + t * (int64_t)p.out_op_kp + jbase;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (jbase + i < p.M) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(y[i]);
            oo[i] = hi;
            if (p.out_op_split) {
              oo[p.M + i] = hi;
              oo[2 * p.M + i] = __float2bfloat16_rn(y[i] - __bfloat162float(hi));
            }
          }
        }
      }
    }
  }
  __device__ __forceinline__ void end_item(const Params& p, int part) {
    if (p.act == 2 && p.ds_part && t < p.N) p.ds_part[(int64_t)part * p.N + t] = dsum;
  }
};

// Streaming top-k: every row appends the scores that beat its threshold `tau` to a 64-entry
// buffer in global memory (fire-and-forget stores); when a buffer is nearly full the WARP sorts
// it cooperatively (bitonic network over 64 entries, two per lane), keeps the TOPK_KEEP smallest
// and lowers tau to the largest kept score.  Invariant: every column that is not in the buffer has
// score >= tau, which is what topk_select_kernel's completeness proof needs.  A row compacts only
// O(log(M / 64)) times, so the per-column cost stays at one FFMA + one compare.
struct TopkEpi {
  static constexpr int kStageBytes = 0;
  __device__ __forceinline__ void bind_stage(float*) {}
  float tau;
  int cnt;
  int64_t t;
  float2* buf;  // (score, column as int bits), TOPK_POOL entries

  __device__ __forceinline__ float2* buf_of(const Params& p, int64_t row, int split) const {
    return p.pool + ((int64_t)split * p.N + row) * TOPK_POOL;
  }
  __device__ __forceinline__ void begin_item(const Params& p, int64_t t_, int split) {
    t = t_;
    tau = INFINITY;
    cnt = 0;
    buf = buf_of(p, t < p.N ? t : 0, split);
    if (t >= p.N)
      tau = -INFINITY;  // padded rows never append
    else if (p.tau_init)
      tau = p.tau_init[t];  // threshold proven by a sample pass: appends are rare from column 0
  }
  // warp-cooperative: sort lane `r`'s buffer (`my_buf` / `my_cnt` are each lane's own), keep the
  // TOPK_KEEP smallest in place and return the new threshold to every lane.  Static with by-value
  // arguments so that the epilogue state stays in registers (no `this` escaping to local memory).
  static __device__ __noinline__ float compact(float2* my_buf, int my_cnt, int r, int keep) {
    const int lane = threadIdx.x & 31;
    const unsigned long long bp = __shfl_sync(0xffffffffu, (unsigned long long)my_buf, r);
    float2* rb = reinterpret_cast<float2*>(bp);
    const int n = __shfl_sync(0xffffffffu, my_cnt, r);
    __syncwarp();  // lane r's appends are visible to the whole warp
    float2 e0 = lane < n ? rb[lane] : make_float2(INFINITY, 0.f);
    float2 e1 = lane + 32 < n ? rb[lane + 32] : make_float2(INFINITY, 0.f);
    // bitonic sort, ascending, of the 64 entries at positions p = lane (e0) and lane + 32 (e1)
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j >= 1; j >>= 1) {
        if (j == 32) {  // partner is this lane's other slot; k == 64: ascending everywhere
          if (e1.x < e0.x) {
            const float2 tmp = e0;
            e0 = e1;
            e1 = tmp;
          }
        } else {
          const bool lower = (lane & j) == 0;
          {
            const float ox = __shfl_xor_sync(0xffffffffu, e0.x, j);
            const float oy = __shfl_xor_sync(0xffffffffu, e0.y, j);
            const bool asc = (lane & k) == 0;  // position p = lane
            const bool take = (lower == asc) ? (ox < e0.x) : (e0.x < ox);
            if (take) e0 = make_float2(ox, oy);
          }
          {
            const float ox = __shfl_xor_sync(0xffffffffu, e1.x, j);
            const float oy = __shfl_xor_sync(0xffffffffu, e1.y, j);
            const bool asc = ((lane + 32) & k) == 0;  // position p = lane + 32
            const bool take = (lower == asc) ? (ox < e1.x) : (e1.x < ox);
            if (take) e1 = make_float2(ox, oy);
          }
        }
      }
    }
    if (lane < keep) rb[lane] = e0;  // positions 0..keep-1 = the `keep` smallest
    const float new_tau = __shfl_sync(0xffffffffu, e0.x, keep - 1);
    __syncwarp();
    return new_tau;
  }
  __device__ __forceinline__ void chunk(const Params& p, const uint32_t (&v)[32],
                                        const float* __restrict__ bias, float scale, int64_t jbase,
                                        unsigned int*) {
    const int keep = p.topk_keep;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float4 b0 = reinterpret_cast<const float4*>(bias)[2 * g];
      const float4 b1 = reinterpret_cast<const float4*>(bias)[2 * g + 1];
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      // branch-free appends: with 32 independent rows per warp SOME lane beats its threshold in
      // almost every group, so a "rare path" branch would be taken all the time; predicated stores
      // cost three issue slots per column and no divergence
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = fmaf(scale, __uint_as_float(v[8 * g + i]), bb[i]);
        if (d < tau) {
          buf[cnt] = make_float2(d, __int_as_float((int)(jbase + 8 * g + i)));
          ++cnt;
        }
      }
      // cnt <= TOPK_POOL - 8 before a group, so a group can never overflow the buffer
      unsigned int full = __ballot_sync(0xffffffffu, cnt > TOPK_POOL - 8);
      while (full) {
        const int r = __ffs(full) - 1;
        full &= full - 1;
        const float nt = compact(buf, cnt, r, keep);
        if ((int)(threadIdx.x & 31) == r) {
          cnt = keep;
          tau = nt;
        }
      }
    }
  }
  __device__ __forceinline__ void end_item(const Params& p, int split) {
    if (t < p.N) p.pool_meta[(int64_t)split * p.N + t] = make_float2((float)cnt, tau);
  }
};

// ------------------------------------------------------------------------------------- kernel
// Work items.  Item i < tail_first is (query-tile group i % q_groups, gallery split i / q_groups):
// tiles_per_split consecutive gallery tiles.  Items are dealt round-robin to the co-resident
// clusters, so the last round would keep only (items mod clusters) of them busy for a whole item;
// instead each item of that round is cut into tail_parts sub-ranges (Params::tail_first /
// tail_parts, EPI_RANK only: its counts are additive over gallery ranges), one per idle cluster.
struct Work {
  int qg, split, t0, t1;
};
// (out of line, scalar arguments by value: called once per work item by each role; inlined five times
// it grew the kernel by a quarter, and the epilogue-bound D = 256 shape lost 8 % to the instruction
// cache -- a reference to Params would instead push the kernel parameters through local memory)
static __device__ __noinline__ Work decode_work_impl(int item, int q_groups, int tail_first, int tail_parts,
                                              int tiles_per_split, int g_tiles) {
  int base = item, part = 0, parts = 1;
  if (item >= tail_first) {
    const int u = item - tail_first;
    base = tail_first + u / tail_parts;
    part = u % tail_parts;
    parts = tail_parts;
  }
  Work w;
  w.qg = base % q_groups;
  w.split = base / q_groups;
  w.t0 = w.split * tiles_per_split;
  w.t1 = min(g_tiles, w.t0 + tiles_per_split);
  if (parts > 1) {
    const int len = w.t1 - w.t0, lo = w.t0;
    // (len <= 2^22 tiles and parts <= 148: the products fit 32 bits)
    w.t0 = lo + (int)(((unsigned)len * (unsigned)part) / (unsigned)parts);
    w.t1 = lo + (int)(((unsigned)len * (unsigned)(part + 1)) / (unsigned)parts);
  }
  return w;
}
__device__ __forceinline__ Work decode_work(const Params& p, int item, int q_groups) {
  return decode_work_impl(item, q_groups, p.tail_first, p.tail_parts, p.tiles_per_split, p.g_tiles);
}

template <typename Epi, bool kRes, int kC, bool kPair = false, int kBN = BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
              const Params p) {
  static_assert(!kPair || kC == 2, "a CTA pair is a cluster of two");
  static_assert(kBN == 256 || kBN == 128 || kBN == 64, "tile widths built: 256, 128, 64 gallery rows");
  using L = SmemLayout<kRes, kPair, kBN, Epi::kStageBytes>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* res_a = smem;
  uint8_t* stages = smem + L::kStagesOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* full = bars;
  uint64_t* empty = full + L::kStages;
  uint64_t* a_full = empty + L::kStages;
  uint64_t* a_empty = a_full + L::kResKb;
  uint64_t* tmem_full = a_empty + 1;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  unsigned int* seg_count = tmem_slot + 1;  // EPI_RANK: entries this CTA pushed to its list segment

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // programmatic dependent launch (common.cuh): the next kernel of the stream may be scheduled once
  // every CTA of this one is running; this one touches global memory only below griddep_wait()
  griddep_launch();
  // a gated fallback pass that is not needed: every CTA sees the same flag and leaves before any
  // barrier, cluster or tensor-memory state exists
  if (p.run_flag) {
    griddep_wait();
    if (*p.run_flag == 0u) return;
  }

  if (threadIdx.x == 0) {
    *seg_count = 0;
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < L::kStages; ++i) {
      mbar_init(&full[i], 1);
      // one tcgen05.commit arrival from every MMA issuer of the cluster (a pair has one)
      mbar_init(&empty[i], kPair ? 1 : kC);
    }
    for (int i = 0; i < L::kResKb; ++i) mbar_init(&a_full[i], 1);
    mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      // one arrival per epilogue warp; the leader of a pair collects both CTAs' warps
      mbar_init(&tmem_empty[i], kPair ? 2 * EPI_WARPS : EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (kPair) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kC > 1) cluster_sync_all();  // peers' barriers are initialised before anything crosses CTAs
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // set-up done (shared memory, tensor memory, descriptor prefetch only): from here on the kernel
  // reads what its predecessors wrote
  griddep_wait();

  // Work items are dealt to CLUSTERS: the kC CTAs of a cluster take kC consecutive query tiles
  // against the same gallery tiles, so every gallery stage is fetched from L2 once per cluster
  // (each CTA loads 1/kC of it and multicasts).
  const int cta_rank = kC > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / kC;
  const int num_clusters = gridDim.x / kC;
  const int q_groups = (p.q_tiles + kC - 1) / kC;
  const int num_items = p.num_items;  // incl. the sub-items of the balanced last round
  const int nkb = p.num_kb;
  // ring depth in use (profiling: VTC_DBG_STAGES limits it to show how much load latency the ring hides)
  const uint32_t nstages = p.dbg_stages > 0 && p.dbg_stages < L::kStages ? (uint32_t)p.dbg_stages
                                                                         : (uint32_t)L::kStages;
  const bool swap_pipelined = kRes && nstages <= (uint32_t)nkb;
  constexpr uint16_t kMask = (uint16_t)((1u << kC) - 1);
  // CTA pair: rank 0 issues the MMAs and owns the `full` / `a_full` / `tmem_empty` barriers; both
  // CTAs' TMA loads and epilogue warps signal ITS barriers (shared::cluster addresses)
  const bool leader = !kPair || cta_rank == 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, it = 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const Work wk = decode_work(p, item, q_groups);
        const int qt = wk.qg * kC + cta_rank, t0 = wk.t0, t1 = wk.t1;
        if (t0 >= t1) continue;  // an empty sub-range: skipped by every role alike
        // The resident query tile is swapped k-block by k-block behind the previous item's last tile
        // when the ring is no deeper than the tile (nstages <= nkb): the wait for ring slot P below
        // guarantees that the MMAs of position P - nstages -- hence of P - nkb, the last reader of this
        // k-block of the query tile -- have retired.  Otherwise wait for the whole item to drain.
        if (kRes && !swap_pipelined) mbar_wait(a_empty, (it & 1) ^ 1);
        ++it;
        for (int tile = t0; tile < t1; ++tile) {
          for (int kb = 0; kb < nkb; ++kb) {
            if (kPair) {
              // each CTA loads its own query rows and its half of the gallery tile into its own
              // shared memory; the bytes of both are expected on the leader's barrier
              mbar_wait(&empty[stage], phase ^ 1);
              if (kRes && tile == t0) {
                if (leader) mbar_arrive_expect_tx(&a_full[kb], 2 * A_TILE_BYTES);
                tma_load_2d_pair(res_a + kb * A_TILE_BYTES, &tmA,
                                 mapa_shared(smem_u32(&a_full[kb]), 0), kb * BK, qt * BM);
              }
              if (leader) mbar_arrive_expect_tx(&full[stage], 2 * L::kStageBytes);
              const uint32_t fbar = mapa_shared(smem_u32(&full[stage]), 0);
              uint8_t* st = stages + stage * L::kStageBytes;
              if (!kRes) tma_load_2d_pair(st, &tmA, fbar, kb * BK, qt * BM);
              tma_load_2d_pair(st + (kRes ? 0 : A_TILE_BYTES), &tmB, fbar, kb * BK,
                               tile * kBN + cta_rank * (kBN / 2));
              if (++stage == nstages) stage = 0, phase ^= 1;
              continue;
            }
            mbar_wait(&empty[stage], phase ^ 1);
            if (kRes && tile == t0) {
              mbar_arrive_expect_tx(&a_full[kb], A_TILE_BYTES);
              tma_load_2d(res_a + kb * A_TILE_BYTES, &tmA, &a_full[kb], kb * BK, qt * BM);
            }
            mbar_arrive_expect_tx(&full[stage], L::kStageBytes);
            uint8_t* st = stages + stage * L::kStageBytes;
            if (!kRes) tma_load_2d(st, &tmA, &full[stage], kb * BK, qt * BM);
            uint8_t* bdst = st + (kRes ? 0 : A_TILE_BYTES);
            if (kC == 1) {
              tma_load_2d(bdst, &tmB, &full[stage], kb * BK, tile * kBN);
            } else {
              constexpr int kSliceRows = kBN / kC;
              tma_load_2d_multicast(bdst + cta_rank * kSliceRows * (BK * 2), &tmB, &full[stage],
                                    kb * BK, tile * kBN + cta_rank * kSliceRows, kMask);
            }
            if (++stage == nstages) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(kPair ? 2 * BM : BM, kBN);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0, it = 0;
      // VTC_DBG_PROF: where the issuer waits (one thread; two clock reads per wait when enabled)
      const bool prof = p.dbg_prof != nullptr;
      unsigned long long w_acc = 0, w_ld = 0, n_tiles = 0;
      const long long c_begin = prof ? clock64() : 0;
      const uint64_t g_begin = prof ? globaltimer_ns() : 0;
      for (int item = cluster_id; item < num_items; item += num_clusters) {
        const Work wk = decode_work(p, item, q_groups);
        const int t0 = wk.t0, t1 = wk.t1;
        if (t0 >= t1) continue;
        for (int tile = t0; tile < t1; ++tile) {
          {
            // (one copy of every wait: the profiling reads are predicated around it)
            const long long c0 = prof ? clock64() : 0;
            mbar_wait(&tmem_empty[as], aphase ^ 1);  // epilogue has drained this accumulator
            if (prof) w_acc += (unsigned long long)(clock64() - c0), ++n_tiles;
          }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * kBN;
          for (int kb = 0; kb < nkb; ++kb) {
            if (kRes && tile == t0) mbar_wait(&a_full[kb], it & 1);
            {
              const long long c0 = prof ? clock64() : 0;
              mbar_wait(&full[stage], phase);
              if (prof) w_ld += (unsigned long long)(clock64() - c0);
            }
            tc_fence_after();
            uint8_t* st = stages + stage * L::kStageBytes;
            const uint64_t adesc =
                make_smem_desc_sw128(smem_u32(kRes ? res_a + kb * A_TILE_BYTES : st));
            const uint64_t bdesc = make_smem_desc_sw128(smem_u32(st + (kRes ? 0 : A_TILE_BYTES)));
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              if (kPair)
                umma_bf16_pair(d_tmem, desc_advance(adesc, k4 * 32), desc_advance(bdesc, k4 * 32),
                               idesc, (uint32_t)((kb | k4) != 0));
              else
                umma_bf16(d_tmem, desc_advance(adesc, k4 * 32), desc_advance(bdesc, k4 * 32), idesc,
                          (uint32_t)((kb | k4) != 0));
            }
            // smem slot free (in every CTA of the cluster) once these MMAs retire
            if (kPair)
              umma_commit_pair(&empty[stage], kMask);
            else if (kC == 1)
              umma_commit(&empty[stage]);
            else
              umma_commit_multicast(&empty[stage], kMask);
            if (++stage == nstages) stage = 0, phase ^= 1;
          }
          // accumulator complete (in both CTAs of a pair)
          if (kPair)
            umma_commit_pair(&tmem_full[as], kMask);
          else
            umma_commit(&tmem_full[as]);
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
        if (kRes) {
          if (kPair)
            umma_commit_pair(a_empty, kMask);
          else
            umma_commit(a_empty);
        }
        ++it;
      }
      if (prof) {
        unsigned long long* o = p.dbg_prof + (size_t)blockIdx.x * 8;
        o[0] = (unsigned long long)(clock64() - c_begin);
        o[1] = globaltimer_ns() - g_begin;
        o[2] = w_acc;
        o[3] = w_ld;
        o[4] = n_tiles;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue
    // Warp w reads TMEM lane quarter w % 4 (a hardware rule) and column half (w - 4) / 4 of every
    // accumulator: 8 epilogue warps = two per scheduler, so one warp's TMEM / shared-memory
    // latencies hide behind the other's arithmetic.  Each (row, half) keeps its own epilogue state;
    // partial results are indexed by part = 2 * split + half.
    const int q4 = warp & 3;
    const int half = (warp - EPI_WARP0) >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const float scale = p.scale * (p.scale_ptr ? *p.scale_ptr : 1.0f);
    constexpr int kHalfCols = kBN / 2;  // 128 (64) columns = 4 (2) chunks of 32
    uint32_t as = 0, aphase = 0;
    Epi epi;
    float* wbias = reinterpret_cast<float*>(smem + L::kBiasOff) + (warp - EPI_WARP0) * 64;
    epi.bind_stage(reinterpret_cast<float*>(smem + L::kEpiStageOff +
                                            (warp - EPI_WARP0) * Epi::kStageBytes));
    // column bias, staged 64 columns at a time in this warp's private buffer; lanes 0..15 carry the
    // next 64 values in registers (prefetched one step ahead so the L2 latency is never exposed)
    // (pre_j: the column bias_pre was fetched for; a wrong guess at an item boundary is re-fetched)
    float4 bias_pre = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t pre_j = -1;
    const bool eprof = p.dbg_prof != nullptr;
    unsigned long long e_wait = 0;
    const long long e_begin = eprof ? clock64() : 0;
    for (int item = cluster_id; item < num_items; item += num_clusters) {
      const Work wk = decode_work(p, item, q_groups);
      const int qt = wk.qg * kC + cta_rank, split = wk.split, t0 = wk.t0, t1 = wk.t1;
      if (t0 >= t1) continue;
      epi.begin_item(p, (int64_t)qt * BM + row, 2 * split + half);
      for (int tile = t0; tile < t1; ++tile) {
        const int64_t j0 = (int64_t)tile * kBN + half * kHalfCols;
        // first column this warp will process in its next tile (for the bias prefetch)
        int64_t next_j0 = j0 + kBN;
        if (tile + 1 >= t1) {
          const int nitem = item + num_clusters;
          next_j0 = nitem < num_items
                        ? (int64_t)decode_work(p, nitem, q_groups).t0 * kBN + half * kHalfCols
                        : 0;
        }
        {
          const long long c0 = eprof ? clock64() : 0;
          mbar_wait(&tmem_full[as], aphase);
          if (eprof) e_wait += (unsigned long long)(clock64() - c0);
        }
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_off + as * kBN + half * kHalfCols;
        uint32_t va[32], vb[32];
        tmem_ld_32x32(taddr, va);
        tmem_ld_wait(va);
        if constexpr (kHalfCols == 32) {
          // 64-column tiles (small dense products): one chunk per warp and tile
          (void)vb, (void)next_j0;
          __syncwarp();
          if (lane < 8)
            reinterpret_cast<float4*>(wbias)[lane] =
                __ldg(reinterpret_cast<const float4*>(p.col_bias + j0) + lane);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair)
              mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
            else
              mbar_arrive(&tmem_empty[as]);
          }
          if (!p.dbg_skip_epilogue) epi.chunk(p, va, wbias, scale, j0, seg_count);
        } else {
#pragma unroll 1
        for (int c = 0; c < kHalfCols / 32; c += 2) {
          // stage the bias of columns [32c, 32c + 64) and start fetching the following 64
          __syncwarp();
          if (pre_j != j0 + c * 32 && lane < 16)  // first chunk of the launch, or a skipped item
            bias_pre = __ldg(reinterpret_cast<const float4*>(p.col_bias + j0 + c * 32) + lane);
          if (lane < 16) reinterpret_cast<float4*>(wbias)[lane] = bias_pre;
          __syncwarp();
          pre_j = c + 2 < kHalfCols / 32 ? j0 + (c + 2) * 32 : next_j0;
          if (lane < 16) bias_pre = __ldg(reinterpret_cast<const float4*>(p.col_bias + pre_j) + lane);
          tmem_ld_32x32(taddr + (c + 1) * 32, vb);
          if (!p.dbg_skip_epilogue) epi.chunk(p, va, wbias, scale, j0 + c * 32, seg_count);
          tmem_ld_wait(vb);
          if (c + 2 < kHalfCols / 32) {
            tmem_ld_32x32(taddr + (c + 2) * 32, va);
          } else {
            // this warp's TMEM reads of the accumulator are complete: one of the 8 arrivals that
            // hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (kPair)
                mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
              else
                mbar_arrive(&tmem_empty[as]);
            }
          }
          if (!p.dbg_skip_epilogue) epi.chunk(p, vb, wbias + 32, scale, j0 + (c + 1) * 32, seg_count);
          if (c + 2 < kHalfCols / 32) tmem_ld_wait(va);
        }
        }
        as ^= 1;
        if (as == 0) aphase ^= 1;
      }
      epi.end_item(p, 2 * split + half);
    }
    if (eprof && warp == EPI_WARP0 && lane == 0) {
      p.dbg_prof[(size_t)blockIdx.x * 8 + 5] = e_wait;
      p.dbg_prof[(size_t)blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - e_begin);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && p.amb_seg_count) p.amb_seg_count[blockIdx.x] = *seg_count;
  if (kC > 1) cluster_sync_all();  // no CTA leaves while peers may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (kPair)
      tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else
      tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------ launch helper
// Launches one instance with a thread-block-cluster dimension of kC (1 = plain launch).
template <typename Epi, bool kRes, int kC, bool kPair = false, int kBN = BN>
int launch_instance(const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, int grid,
                    cudaStream_t s) {
  using L = SmemLayout<kRes, kPair, kBN, Epi::kStageBytes>;
  auto* kern = &sim_tc_kernel<Epi, kRes, kC, kPair, kBN>;
  // the attribute is per function and per device: cheap, so set it on every launch
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
  if (e != cudaSuccess) return cuda_err(e);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kC > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = kC;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p);
  if (e != cudaSuccess) return cuda_err(e);
  return VTC_OK;
}

// how many clusters of kC CTAs of this kernel can be co-resident on the current device
template <typename Epi, bool kRes, int kC, bool kPair = false>
int max_active_clusters() {
  using L = SmemLayout<kRes, kPair, BN, Epi::kStageBytes>;
  auto* kern = &sim_tc_kernel<Epi, kRes, kC, kPair>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal) != cudaSuccess)
    return 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(kNumSMs / kC * kC));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::kTotal;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
  return n;
}

// dispatch over (resident, cluster) for one epilogue (256-column tiles)
template <typename Epi>
int launch_epilogue(bool a_resident, int cluster, const CUtensorMap& tmA, const CUtensorMap& tmB,
                    const Params& p, int grid, cudaStream_t s, bool pair = false) {
  if (pair && cluster == 2)
    return a_resident ? launch_instance<Epi, true, 2, true>(tmA, tmB, p, grid, s)
                      : launch_instance<Epi, false, 2, true>(tmA, tmB, p, grid, s);
  if (a_resident) {
    if (cluster == 4) return launch_instance<Epi, true, 4>(tmA, tmB, p, grid, s);
    if (cluster == 2) return launch_instance<Epi, true, 2>(tmA, tmB, p, grid, s);
    return launch_instance<Epi, true, 1>(tmA, tmB, p, grid, s);
  }
  if (cluster == 4) return launch_instance<Epi, false, 4>(tmA, tmB, p, grid, s);
  if (cluster == 2) return launch_instance<Epi, false, 2>(tmA, tmB, p, grid, s);
  return launch_instance<Epi, false, 1>(tmA, tmB, p, grid, s);
}
template <typename Epi>
int launch_epilogue_c1(bool a_resident, const CUtensorMap& tmA, const CUtensorMap& tmB,
                       const Params& p, int grid, cudaStream_t s) {
  return a_resident ? launch_instance<Epi, true, 1>(tmA, tmB, p, grid, s)
                    : launch_instance<Epi, false, 1>(tmA, tmB, p, grid, s);
}

}  // namespace tc
}  // namespace vtc
