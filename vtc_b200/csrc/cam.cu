// cam.cu -- Context Adapter Module kernels (model/model.py:141-205; clip.model.Transformer block
// structure per model/timesformer_clip_alt.py:22-33,43-67,112-124).  All fp32, HBM/latency-bound:
// vectorised, coalesced row kernels (one warp per row) and a one-warp-per-(sample, head) attention
// core for the short token axis (L = 1 + #comments <= 16).  The dense projections run on the
// tensor cores through vtc_linear (sim_tc.cu, EPI_STORE).
#include "cam.cuh"

namespace vtc {

constexpr int WARPS = 8;
constexpr int MAX_VEC = 8;  // float4 per lane: rows up to D = 1024

// row r of the stacked input: l == 0 -> main[b], else aux[l-1][b]
__global__ void __launch_bounds__(256)
cam_stack_normalize_kernel(const float* __restrict__ main, const float* __restrict__ aux, int L,
                           int64_t b, int D, float* __restrict__ X) {
  griddep_launch();
  griddep_wait();
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= (int64_t)L * b) return;
  const int lane = threadIdx.x & 31;
  const int64_t l = r / b, bi = r % b;
  const float* src = l == 0 ? main + bi * D : aux + ((l - 1) * b + bi) * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s = fmaf(src[k], src[k], s);
  const float nrm = sqrtf(warp_sum(s));
  float* dst = X + r * D;
  for (int k = lane; k < D; k += 32) dst[k] = src[k] / nrm;
}

__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                 const float* __restrict__ beta, int64_t rows, int D, float eps,
                 float* __restrict__ Y) {
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* x = X + r * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s += x[k];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float d = x[k] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)D + eps);
  float* y = Y + r * D;
  for (int k = lane; k < D; k += 32) y[k] = (x[k] - mean) * rstd * gamma[k] + beta[k];
}

// bf16 operand element(s) of value v at column k of a row laid out as [x] or [hi | hi | lo]
__device__ __forceinline__ void put_operand(__nv_bfloat16* row, int k, int D, int split, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  row[k] = hi;
  if (split) {
    row[D + k] = hi;
    row[2 * D + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
  }
}

// LayerNorm fused with operand preparation: the normalised row goes straight out as the
// query-side bf16 operand of the following linear (no fp32 round trip, one launch less).
__global__ void __launch_bounds__(256)
layernorm_prep_kernel(const float* __restrict__ X, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int64_t rows, int D, float eps, int split,
                      __nv_bfloat16* __restrict__ out, int Kp) {
  griddep_launch();
  griddep_wait();
  const int64_t r = (int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* x = X + r * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s += x[k];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float d = x[k] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)D + eps);
  __nv_bfloat16* o = out + r * (int64_t)Kp;
  for (int k = lane; k < D; k += 32)
    put_operand(o, k, D, split, (x[k] - mean) * rstd * gamma[k] + beta[k]);
  for (int k = (split ? 3 : 1) * D + lane; k < Kp; k += 32) o[k] = __float2bfloat16_rn(0.f);
}

// softmax(q k^T / sqrt(hd)) v for the short token axis of the CAM (L = 1 + #comments <= 16).
// QKV is the in_proj output [L, b, 3D] (q | k | v along the last dim, heads contiguous inside each),
// out is [L, b, D].  LPH lanes per (sample, head), each owning FOUR consecutive dims of the head
// (head_dim <= 4 * LPH): every q / k / v access is one 128-bit load, the LPH lanes of a head read
// 16 * LPH contiguous bytes (256 B at head_dim 64), and all 3L loads of a lane are issued before the
// first use.  Dot products reduce over the LPH lanes with log2(LPH) shuffles.
template <int MAXL, int LPH>
__global__ void __launch_bounds__(256)
cam_attn_core_kernel(const float* __restrict__ QKV, int L, int64_t b, int D, int heads,
                     float* __restrict__ out, __nv_bfloat16* __restrict__ out_op, int Kp, int split) {
  griddep_launch();
  griddep_wait();
  constexpr int HPW = 32 / LPH;  // heads per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPH, li = lane % LPH;
  const int64_t w = ((int64_t)blockIdx.x * WARPS + (threadIdx.x >> 5)) * HPW + sub;
  const bool live = w < b * heads;  // (no early return: the shuffles below are warp-wide)
  const int64_t bi = live ? w / heads : 0;
  const int h = live ? (int)(w % heads) : 0;
  const int hd = D / heads;
  const int d0 = 4 * li;
  const bool has = live && d0 < hd;  // hd % 4 == 0 is checked by the launcher
  const float scaling = rsqrtf((float)hd);
  float4 q[MAXL], k[MAXL], v[MAXL];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
    q[l] = k[l] = v[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < L && has) {
      const float4* base =
          reinterpret_cast<const float4*>(QKV + ((int64_t)l * b + bi) * 3 * D + h * hd + d0);
      q[l] = __ldg(base);
      k[l] = __ldg(base + D / 4);
      v[l] = __ldg(base + D / 2);
    }
  }
#pragma unroll
  for (int i = 0; i < MAXL; ++i) {
    if (i >= L) break;
    float s[MAXL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      // q = q * head_dim^-0.5 (timesformer_clip_alt.py:43)
      float part = (q[i].x * k[j].x + q[i].y * k[j].y + q[i].z * k[j].z + q[i].w * k[j].w) * scaling;
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      s[j] = j < L ? part : -INFINITY;
      mx = fmaxf(mx, s[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      s[j] = j < L ? __expf(s[j] - mx) : 0.f;
      den += s[j];
    }
    const float inv = 1.f / den;
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      const float p = s[j] * inv;
      o4.x = fmaf(p, v[j].x, o4.x), o4.y = fmaf(p, v[j].y, o4.y);
      o4.z = fmaf(p, v[j].z, o4.z), o4.w = fmaf(p, v[j].w, o4.w);
    }
    if (has) {
      const int64_t row = (int64_t)i * b + bi;
      const int col = h * hd + d0;
      if (out) *reinterpret_cast<float4*>(out + row * D + col) = o4;
      if (out_op) {
        __nv_bfloat16* op = out_op + row * Kp + col;
        const float y[4] = {o4.x, o4.y, o4.z, o4.w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          hi[e] = __float2bfloat16_rn(y[e]);
          lo[e] = __float2bfloat16_rn(y[e] - __bfloat162float(hi[e]));
        }
        const uint2 h2 = make_uint2(
            (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16),
            (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16));
        *reinterpret_cast<uint2*>(op) = h2;  // (col % 4 == 0, Kp % 64 == 0: 8-byte aligned)
        if (split) {
          *reinterpret_cast<uint2*>(op + D) = h2;
          *reinterpret_cast<uint2*>(op + 2 * D) = make_uint2(
              (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16),
              (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16));
        }
      }
    }
  }
}

__global__ void bias_act_kernel(const float* __restrict__ X, const float* __restrict__ bias,
                                const float* __restrict__ residual, int64_t total, int D, int act,
                                float* __restrict__ Y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    float z = X[i] + (bias ? bias[i % D] : 0.f);
    if (act == 1) z = z / (1.f + __expf(-1.702f * z));
    Y[i] = z + (residual ? residual[i] : 0.f);
  }
}

// One block per sample, T is [L, b, D].  The L token rows are read by L different warps at the same
// time (each: 128-bit loads, its own norm, the weighted row into shared memory); warp 0 then sums the
// partial rows and finishes the sample (mean, residual activation, normalize(normalize(main) + res)).
// A warp per sample walking its tokens one after the other is a chain of L dependent
// load -> reduce round trips: 14 us at b = 256, L = 6 for 3.6 MB of traffic.
constexpr int RO_MAX_D = 128 * MAX_VEC;
__global__ void __launch_bounds__(256)
cam_readout_kernel(const float* __restrict__ T, const float* __restrict__ main,
                   const float* __restrict__ res_in, const uint8_t* __restrict__ skip_mask, int L,
                   int64_t b, int D, int mode, int res_act, float res_scale,
                   const float* __restrict__ res_shift, const float* __restrict__ res_mul,
                   float* __restrict__ out) {
  griddep_launch();
  griddep_wait();
  __shared__ __align__(16) float part[WARPS][RO_MAX_D];
  const int64_t bi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[MAX_VEC];
  const int nvec = D / 4;  // D % 4 == 0 checked by the launcher
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (mode == VTC_CAM_READOUT_AVG || mode == VTC_CAM_READOUT_UNIFORM) {
    for (int l = warp; l < L; l += WARPS) {
      const float4* x = reinterpret_cast<const float4*>(T + ((int64_t)l * b + bi) * D);
      float4 xv[MAX_VEC];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        const int c = lane + 32 * i;
        xv[i] = c < nvec ? x[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += xv[i].x * xv[i].x + xv[i].y * xv[i].y + xv[i].z * xv[i].z + xv[i].w * xv[i].w;
      }
      // AVG: mean of normalised tokens (model.py:156-159); UNIFORM: plain mean (:356-362)
      const float w = mode == VTC_CAM_READOUT_AVG ? 1.f / sqrtf(warp_sum(s)) : 1.f;
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        acc[i].x = fmaf(xv[i].x, w, acc[i].x);
        acc[i].y = fmaf(xv[i].y, w, acc[i].y);
        acc[i].z = fmaf(xv[i].z, w, acc[i].z);
        acc[i].w = fmaf(xv[i].w, w, acc[i].w);
      }
    }
    const int nw = L < WARPS ? L : WARPS;  // warps that hold a partial row
    if (warp > 0 && warp < nw) {
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) reinterpret_cast<float4*>(part[warp])[c] = acc[i];
      }
    }
    __syncthreads();
    if (warp != 0) return;
    // fixed summation order (warp 1, 2, ...): the result does not depend on scheduling
    for (int w2 = 1; w2 < nw; ++w2) {
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
          const float4 pv = reinterpret_cast<const float4*>(part[w2])[c];
          acc[i].x += pv.x, acc[i].y += pv.y, acc[i].z += pv.z, acc[i].w += pv.w;
        }
      }
    }
    const float invL = 1.f / (float)L;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      acc[i].x *= invL, acc[i].y *= invL, acc[i].z *= invL, acc[i].w *= invL;
      s += acc[i].x * acc[i].x + acc[i].y * acc[i].y + acc[i].z * acc[i].z + acc[i].w * acc[i].w;
    }
    const float nrm = sqrtf(warp_sum(s));
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i)
      acc[i].x /= nrm, acc[i].y /= nrm, acc[i].z /= nrm, acc[i].w /= nrm;
    if (mode == VTC_CAM_READOUT_UNIFORM) {
      float4* o = reinterpret_cast<float4*>(out + bi * D);
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) o[c] = acc[i];
      }
      return;
    }
  } else {
    if (warp != 0) return;
    const float4* x = reinterpret_cast<const float4*>(res_in + bi * D);
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) acc[i] = x[c];
    }
  }
  // residual activation (model/model.py:30-77, applied at :168-171)
  if (res_act != VTC_RESACT_NONE) {
    if (res_act == VTC_RESACT_NORMALIZE_EPS || res_act == VTC_RESACT_SQUASH) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        if (lane + 32 * i < nvec) {
          acc[i].x += 1e-9f, acc[i].y += 1e-9f, acc[i].z += 1e-9f, acc[i].w += 1e-9f;
          s += acc[i].x * acc[i].x + acc[i].y * acc[i].y + acc[i].z * acc[i].z + acc[i].w * acc[i].w;
        }
      }
      const float mag_sq = warp_sum(s);
      const float mag = sqrtf(mag_sq);
      // normalize_eps: (x + eps) / |x + eps|;  squash: c * mag^2 / (1 + mag^2) * (x + eps) / mag
      const float f = res_act == VTC_RESACT_NORMALIZE_EPS
                          ? 1.f / mag
                          : res_scale * (mag_sq / (1.f + mag_sq)) / mag;
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) acc[i].x *= f, acc[i].y *= f, acc[i].z *= f, acc[i].w *= f;
    } else if (res_act == VTC_RESACT_TANH) {
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i)
        acc[i] = make_float4(tanhf(acc[i].x), tanhf(acc[i].y), tanhf(acc[i].z), tanhf(acc[i].w));
    } else if (res_act == VTC_RESACT_AFFINE) {  // eval-mode sub_mean / bn: (x - shift) * mul
#pragma unroll
      for (int i = 0; i < MAX_VEC; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
          const float4 sh = reinterpret_cast<const float4*>(res_shift)[c];
          const float4 mu = res_mul ? reinterpret_cast<const float4*>(res_mul)[c]
                                    : make_float4(1.f, 1.f, 1.f, 1.f);
          acc[i] = make_float4((acc[i].x - sh.x) * mu.x, (acc[i].y - sh.y) * mu.y,
                               (acc[i].z - sh.z) * mu.z, (acc[i].w - sh.w) * mu.w);
        }
      }
    }
  }
  if (skip_mask && skip_mask[bi]) {  // random adapter skip (model.py:199-201)
#pragma unroll
    for (int i = 0; i < MAX_VEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // adapted = normalize(normalize(main) + res)   (model.py:203)
  const float4* m = reinterpret_cast<const float4*>(main + bi * D);
  float4 mv[MAX_VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = lane + 32 * i;
    mv[i] = c < nvec ? m[c] : make_float4(0.f, 0.f, 0.f, 0.f);
    s += mv[i].x * mv[i].x + mv[i].y * mv[i].y + mv[i].z * mv[i].z + mv[i].w * mv[i].w;
  }
  const float mn = sqrtf(warp_sum(s));
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    acc[i].x += mv[i].x / mn, acc[i].y += mv[i].y / mn;
    acc[i].z += mv[i].z / mn, acc[i].w += mv[i].w / mn;
    s2 += acc[i].x * acc[i].x + acc[i].y * acc[i].y + acc[i].z * acc[i].z + acc[i].w * acc[i].w;
  }
  const float n2 = sqrtf(warp_sum(s2));
  float4* o = reinterpret_cast<float4*>(out + bi * D);
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) o[c] = make_float4(acc[i].x / n2, acc[i].y / n2, acc[i].z / n2, acc[i].w / n2);
  }
}

// --------------------------------------------------------------------------------- launchers
int launch_cam_stack_normalize(const float* main, const float* aux, int L, int64_t b, int D,
                               float* X, cudaStream_t s) {
  const int64_t rows = (int64_t)L * b;
  if (rows == 0) return VTC_OK;
  launch_pdl(cam_stack_normalize_kernel, dim3((unsigned)ceil_div<int64_t>(rows, WARPS)), dim3(256), 0, s,
             main, aux, L, b, D, X);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_layernorm(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                     float eps, float* Y, cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  layernorm_kernel<<<(unsigned)ceil_div<int64_t>(rows, WARPS), 256, 0, s>>>(X, gamma, beta, rows, D,
                                                                           eps, Y);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

template <int MAXL>
static void launch_attn_t(int lph, unsigned blocks_for, const float* QKV, int L, int64_t b, int D,
                          int heads, float* out, __nv_bfloat16* out_op, int Kp, int split,
                          cudaStream_t s) {
  (void)blocks_for;
  const int64_t units = b * heads;
  if (lph == 8)
    launch_pdl(cam_attn_core_kernel<MAXL, 8>, dim3((unsigned)ceil_div<int64_t>(units, WARPS * 4)),
               dim3(256), 0, s, QKV, L, b, D, heads, out, out_op, Kp, split);
  else if (lph == 16)
    launch_pdl(cam_attn_core_kernel<MAXL, 16>, dim3((unsigned)ceil_div<int64_t>(units, WARPS * 2)),
               dim3(256), 0, s, QKV, L, b, D, heads, out, out_op, Kp, split);
  else
    launch_pdl(cam_attn_core_kernel<MAXL, 32>, dim3((unsigned)ceil_div<int64_t>(units, WARPS)),
               dim3(256), 0, s, QKV, L, b, D, heads, out, out_op, Kp, split);
}

int launch_cam_attn_core(const float* QKV, int L, int64_t b, int D, int heads, float* out,
                         __nv_bfloat16* out_op, int Kp, int split, cudaStream_t s) {
  if (b == 0) return VTC_OK;
  const int hd = heads > 0 ? D / heads : 0;
  // 128-bit accesses: head_dim a multiple of 4 (so is D then), bases 16-byte aligned
  if (L < 1 || L > 16 || heads < 1 || D % heads || hd > 128 || hd % 4 ||
      (reinterpret_cast<uintptr_t>(QKV) & 15) || (out && (reinterpret_cast<uintptr_t>(out) & 15)) ||
      (out_op && ((reinterpret_cast<uintptr_t>(out_op) & 7) || (Kp & 3))))
    return VTC_ERR_UNSUPPORTED_SHAPE;
  const int lph = hd <= 32 ? 8 : (hd <= 64 ? 16 : 32);
  if (L <= 8)
    launch_attn_t<8>(lph, 0, QKV, L, b, D, heads, out, out_op, Kp, split, s);
  else
    launch_attn_t<16>(lph, 0, QKV, L, b, D, heads, out, out_op, Kp, split, s);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_layernorm_prep(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                          float eps, int split, __nv_bfloat16* out, int Kp, cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  launch_pdl(layernorm_prep_kernel, dim3((unsigned)ceil_div<int64_t>(rows, WARPS)), dim3(256), 0, s, X,
             gamma, beta, rows, D, eps, split, out, Kp);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_bias_act(const float* X, const float* bias, const float* residual, int64_t rows, int D,
                    int act, float* Y, cudaStream_t s) {
  const int64_t total = rows * D;
  if (total == 0) return VTC_OK;
  const int64_t blocks = ceil_div<int64_t>(total, 256);
  bias_act_kernel<<<(unsigned)(blocks < kNumSMs * 16 ? blocks : kNumSMs * 16), 256, 0, s>>>(
      X, bias, residual, total, D, act, Y);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_cam_readout(const float* T, const float* main, const float* res_in,
                       const uint8_t* skip_mask, int L, int64_t b, int D, int mode, int res_act,
                       float res_scale, const float* res_shift, const float* res_mul, float* out,
                       cudaStream_t s) {
  if (b == 0) return VTC_OK;
  if (D % 4 || D > 128 * MAX_VEC) return VTC_ERR_UNSUPPORTED_SHAPE;
  if (res_act < VTC_RESACT_NONE || res_act > VTC_RESACT_AFFINE ||
      (res_act == VTC_RESACT_AFFINE && !res_shift))
    return VTC_ERR_INVALID_ARG;
  launch_pdl(cam_readout_kernel, dim3((unsigned)b), dim3(256), 0, s, T, main, res_in, skip_mask, L, b,
             D, mode, res_act, res_scale, res_shift, res_mul, out);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc
