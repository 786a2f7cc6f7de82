// cam_bwd.cu -- backward kernels of the Context Adapter Module (model/model.py:141-205 under
// `loss.backward()`, trainer/trainer.py:79).  The dense products of the backward (dX = dY W,
// dW = dY^T X) run on the tensor cores through vtc_linear on transposed copies (vtc_transpose);
// everything else -- LayerNorm, QuickGELU, the short-sequence attention core, the read-out and the
// input normalisation -- is here: fp32, one warp per row / per (sample, head), coalesced.
#include "common.cuh"

namespace vtc {

constexpr int BW = 8;  // warps per block

// ------------------------------------------------------------------------------- small helpers
__global__ void transpose_kernel(const float* __restrict__ in, int64_t R, int64_t C,
                                 float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? in[r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) out[c * R + r] = tile[threadIdx.x][i];
  }
}

// dU = dF * d/du [u * sigmoid(1.702 u)]
__global__ void gelu_bwd_kernel(const float* __restrict__ dF, const float* __restrict__ U, int64_t n,
                                float* __restrict__ dU) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float u = U[i];
    const float s = 1.f / (1.f + __expf(-1.702f * u));
    dU[i] = dF[i] * (s + 1.702f * u * s * (1.f - s));
  }
}

// out[c] = sum_r X[r, c]   (bias gradients); block = 32 columns x 8 row-lanes
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int64_t R, int64_t C, float* __restrict__ out) {
  __shared__ float sh[BW][33];
  const int64_t c = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int ty = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int64_t r = ty; r < R; r += BW) s += X[r * C + c];
  sh[ty][threadIdx.x & 31] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < BW; ++i) t += sh[i][threadIdx.x & 31];
    out[c] = t;
  }
}

// ------------------------------------------------------------------------------- LayerNorm
// dX = rstd * (g - mean(g) - xhat * mean(g * xhat)) + dres,  g = gamma * dY;
// dgamma += dY * xhat, dbeta += dY (atomics; both zeroed by the launcher)
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ X,
                     const float* __restrict__ gamma, int64_t rows, int D, float eps,
                     const float* __restrict__ dres, float* __restrict__ dX,
                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int64_t r = (int64_t)blockIdx.x * BW + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* x = X + r * D;
  const float* dy = dY + r * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s += x[k];
  const float mean = warp_sum(s) / (float)D;
  float v = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float d = x[k] - mean;
    v = fmaf(d, d, v);
  }
  const float rstd = rsqrtf(warp_sum(v) / (float)D + eps);
  float m1 = 0.f, m2 = 0.f;
  for (int k = lane; k < D; k += 32) {
    const float xh = (x[k] - mean) * rstd;
    const float g = gamma[k] * dy[k];
    m1 += g;
    m2 = fmaf(g, xh, m2);
  }
  m1 = warp_sum(m1) / (float)D;
  m2 = warp_sum(m2) / (float)D;
  for (int k = lane; k < D; k += 32) {
    const float xh = (x[k] - mean) * rstd;
    const float g = gamma[k] * dy[k];
    dX[r * D + k] = rstd * (g - m1 - xh * m2) + (dres ? dres[r * D + k] : 0.f);
    atomicAdd(&dgamma[k], dy[k] * xh);
    atomicAdd(&dbeta[k], dy[k]);
  }
}

// ------------------------------------------------------------------------------- attention core
// One warp per (sample, head): recompute P = softmax(q k^T / sqrt(hd)), then
// dV = P^T dO, dP = dO V^T, dS = P * (dP - rowsum(dP * P)), dQ = dS K / sqrt(hd), dK = dS^T Q / sqrt(hd)
template <int MAXL, int PER>
__global__ void __launch_bounds__(256)
cam_attn_core_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ dO, int L,
                         int64_t b, int D, int heads, float* __restrict__ dQKV) {
  const int64_t w = (int64_t)blockIdx.x * BW + (threadIdx.x >> 5);
  if (w >= b * heads) return;
  const int lane = threadIdx.x & 31;
  const int64_t bi = w / heads;
  const int h = (int)(w % heads);
  const int hd = D / heads;
  const int per = (hd + 31) / 32;  // <= PER
  const float scaling = rsqrtf((float)hd);
  float q[MAXL][PER], k[MAXL][PER], v[MAXL][PER], go[MAXL][PER];
  float dq[MAXL][PER], dk[MAXL][PER], dv[MAXL][PER];
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      q[l][e] = k[l][e] = v[l][e] = go[l][e] = 0.f;
      dq[l][e] = dk[l][e] = dv[l][e] = 0.f;
      const int d = lane + 32 * e;
      if (l < L && e < per && d < hd) {
        const float* base = QKV + ((int64_t)l * b + bi) * 3 * D + h * hd + d;
        q[l][e] = base[0];
        k[l][e] = base[D];
        v[l][e] = base[2 * D];
        go[l][e] = dO[((int64_t)l * b + bi) * D + h * hd + d];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXL; ++i) {
    if (i >= L) break;
    float p[MAXL], dp[MAXL];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      float part = 0.f, part2 = 0.f;
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        part = fmaf(q[i][e], k[j][e], part);
        part2 = fmaf(go[i][e], v[j][e], part2);
      }
      p[j] = j < L ? warp_sum(part) * scaling : -INFINITY;
      dp[j] = j < L ? warp_sum(part2) : 0.f;
      mx = fmaxf(mx, p[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      p[j] = j < L ? __expf(p[j] - mx) : 0.f;
      den += p[j];
    }
    const float inv = 1.f / den;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      p[j] *= inv;
      dot = fmaf(dp[j], p[j], dot);
    }
#pragma unroll
    for (int j = 0; j < MAXL; ++j) {
      if (j >= L) break;
      const float ds = p[j] * (dp[j] - dot) * scaling;  // d loss / d (q_i . k_j)
#pragma unroll
      for (int e = 0; e < PER; ++e) {
        dv[j][e] = fmaf(p[j], go[i][e], dv[j][e]);
        dq[i][e] = fmaf(ds, k[j][e], dq[i][e]);
        dk[j][e] = fmaf(ds, q[i][e], dk[j][e]);
      }
    }
  }
#pragma unroll
  for (int l = 0; l < MAXL; ++l) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int d = lane + 32 * e;
      if (l < L && e < per && d < hd) {
        float* base = dQKV + ((int64_t)l * b + bi) * 3 * D + h * hd + d;
        base[0] = dq[l][e];
        base[D] = dk[l][e];
        base[2 * D] = dv[l][e];
      }
    }
  }
}

// ------------------------------------------------------------------------------- row helpers
// dx = (dy - y (y . dy)) / |x| for y = x / |x|, one warp, D <= 1024 handled by strided loops
__device__ __forceinline__ void normalize_bwd_row(const float* __restrict__ x,
                                                  const float* __restrict__ dy, int D, int lane,
                                                  float* __restrict__ dx, float scale_in) {
  float s = 0.f, c = 0.f;
  for (int k = lane; k < D; k += 32) {
    s = fmaf(x[k], x[k], s);
    c = fmaf(x[k], dy[k], c);
  }
  s = warp_sum(s);
  c = warp_sum(c) * scale_in;
  const float nrm = sqrtf(s);
  // y . dy = (x . dy) / |x|
  const float ydy = c / nrm;
  for (int k = lane; k < D; k += 32) dx[k] = (scale_in * dy[k] - (x[k] / nrm) * ydy) / nrm;
}

// backward of normalize (model/model.py:26-27): one warp per row
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ X, const float* __restrict__ dY, int64_t rows, int D,
                     float* __restrict__ dX) {
  const int64_t r = (int64_t)blockIdx.x * BW + (threadIdx.x >> 5);
  if (r >= rows) return;
  normalize_bwd_row(X + r * D, dY + r * D, D, threadIdx.x & 31, dX + r * D, 1.f);
}

// backward of the averaging fusion normalize(mean_l T_l) (model/model.py:356-366): dT [L,b,D]
__global__ void __launch_bounds__(256)
uniform_readout_bwd_kernel(const float* __restrict__ T, const float* __restrict__ dout, int L,
                           int64_t b, int D, float* __restrict__ dT) {
  const int64_t bi = (int64_t)blockIdx.x * BW + (threadIdx.x >> 5);
  if (bi >= b) return;
  const int lane = threadIdx.x & 31;
  // m = mean_l T_l is rebuilt on the fly; dm = (dout - y (y . dout)) / |m|, dT_l = dm / L
  float mm = 0.f, md = 0.f;
  for (int k = lane; k < D; k += 32) {
    float m = 0.f;
    for (int l = 0; l < L; ++l) m += T[((int64_t)l * b + bi) * D + k];
    m /= (float)L;
    mm = fmaf(m, m, mm);
    md = fmaf(m, dout[bi * D + k], md);
  }
  mm = warp_sum(mm);
  md = warp_sum(md);
  const float nrm = sqrtf(mm);
  for (int k = lane; k < D; k += 32) {
    float m = 0.f;
    for (int l = 0; l < L; ++l) m += T[((int64_t)l * b + bi) * D + k];
    m /= (float)L;
    const float dm = (dout[bi * D + k] - (m / nrm) * (md / nrm)) / nrm / (float)L;
    for (int l = 0; l < L; ++l) dT[((int64_t)l * b + bi) * D + k] = dm;
  }
}

// backward of normalize(stack([main, *aux])) (model/model.py:150-151): dX [L,b,D] -> dmain, daux
__global__ void __launch_bounds__(256)
cam_stack_normalize_bwd_kernel(const float* __restrict__ main, const float* __restrict__ aux,
                               const float* __restrict__ dX, int L, int64_t b, int D,
                               float* __restrict__ dmain, float* __restrict__ daux) {
  const int64_t r = (int64_t)blockIdx.x * BW + (threadIdx.x >> 5);
  if (r >= (int64_t)L * b) return;
  const int lane = threadIdx.x & 31;
  const int64_t l = r / b, bi = r % b;
  const float* src = l == 0 ? main + bi * D : aux + ((l - 1) * b + bi) * D;
  float* dst = l == 0 ? dmain + bi * D : daux + ((l - 1) * b + bi) * D;
  normalize_bwd_row(src, dX + r * D, D, lane, dst, 1.f);
}

// backward of the read-out (model/model.py:156-161,199-203), residual activation = identity:
//   AVG:            n_l = T_l/|T_l|, m = mean_l n_l, r = m/|m|
//   RESIDUAL_ONLY:  r = res_in
//   r = 0 where skipped;  u = main/|main|;  z = u + r;  out = z/|z|
// given dout [b,D] writes dT [L,b,D] (AVG) or dres [b,D] (RESIDUAL_ONLY) and dmain_direct [b,D].
__global__ void __launch_bounds__(256)
cam_readout_bwd_kernel(const float* __restrict__ T, const float* __restrict__ main,
                       const float* __restrict__ res_in, const uint8_t* __restrict__ skip_mask,
                       const float* __restrict__ dout, int L, int64_t b, int D, int mode,
                       int res_act, float res_scale, const float* __restrict__ res_shift,
                       const float* __restrict__ res_mul, float* __restrict__ dT,
                       float* __restrict__ dres, float* __restrict__ dmain) {
  extern __shared__ float smem[];  // per warp: r[D], z[D], dz[D], act aux[D]
  const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t bi = (int64_t)blockIdx.x * BW + wi;
  if (bi >= b) return;
  float* rv = smem + (size_t)wi * 4 * D;
  float* zv = rv + D;
  float* dz = zv + D;
  float* av = dz + D;
  const bool skipped = skip_mask && skip_mask[bi];
  float mnorm = 1.f;
  // forward recompute: r
  if (mode == VTC_CAM_READOUT_AVG) {
    for (int k = lane; k < D; k += 32) rv[k] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float* t = T + ((int64_t)l * b + bi) * D;
      float s = 0.f;
      for (int k = lane; k < D; k += 32) s = fmaf(t[k], t[k], s);
      const float inv = 1.f / sqrtf(warp_sum(s));
      for (int k = lane; k < D; k += 32) rv[k] += t[k] * inv;
    }
    float s = 0.f;
    for (int k = lane; k < D; k += 32) {
      rv[k] /= (float)L;
      s = fmaf(rv[k], rv[k], s);
    }
    mnorm = sqrtf(warp_sum(s));
    for (int k = lane; k < D; k += 32) rv[k] /= mnorm;  // r = m / |m|
  } else {
    for (int k = lane; k < D; k += 32) rv[k] = res_in[bi * D + k];
  }
  // residual activation a = act(r) (model/model.py:30-77), kept in dz[] until dz is formed;
  // av keeps what the activation's Jacobian needs: s = r + eps (normalize_eps / squash), a (tanh)
  float act_mag = 1.f;
  if (res_act == VTC_RESACT_NORMALIZE_EPS || res_act == VTC_RESACT_SQUASH) {
    float q = 0.f;
    for (int k = lane; k < D; k += 32) {
      av[k] = rv[k] + 1e-9f;
      q = fmaf(av[k], av[k], q);
    }
    const float mag_sq = warp_sum(q);
    act_mag = sqrtf(mag_sq);
    const float f = res_act == VTC_RESACT_NORMALIZE_EPS
                        ? 1.f / act_mag
                        : res_scale * (mag_sq / (1.f + mag_sq)) / act_mag;
    for (int k = lane; k < D; k += 32) dz[k] = av[k] * f;
  } else if (res_act == VTC_RESACT_TANH) {
    for (int k = lane; k < D; k += 32) dz[k] = av[k] = tanhf(rv[k]);
  } else if (res_act == VTC_RESACT_AFFINE) {
    for (int k = lane; k < D; k += 32) dz[k] = (rv[k] - res_shift[k]) * (res_mul ? res_mul[k] : 1.f);
  } else {
    for (int k = lane; k < D; k += 32) dz[k] = rv[k];
  }
  // u, z, out
  const float* mp = main + bi * D;
  float s = 0.f;
  for (int k = lane; k < D; k += 32) s = fmaf(mp[k], mp[k], s);
  const float un = sqrtf(warp_sum(s));
  float zs = 0.f;
  for (int k = lane; k < D; k += 32) {
    zv[k] = mp[k] / un + (skipped ? 0.f : dz[k]);
    zs = fmaf(zv[k], zv[k], zs);
  }
  const float zn = sqrtf(warp_sum(zs));
  // dz = (dout - out (out . dout)) / |z|
  const float* go = dout + bi * D;
  float c = 0.f;
  for (int k = lane; k < D; k += 32) c = fmaf(zv[k] / zn, go[k], c);
  c = warp_sum(c);
  for (int k = lane; k < D; k += 32) dz[k] = (go[k] - (zv[k] / zn) * c) / zn;
  __syncwarp();
  // dmain through u = main / |main|
  float cu = 0.f;
  for (int k = lane; k < D; k += 32) cu = fmaf(mp[k] / un, dz[k], cu);
  cu = warp_sum(cu);
  for (int k = lane; k < D; k += 32) dmain[bi * D + k] = (dz[k] - (mp[k] / un) * cu) / un;
  // d r = J_act(r)^T d a   (d a = dz)
  if (res_act == VTC_RESACT_NORMALIZE_EPS) {
    // a = s/|s|:  ds = (da - a (a . da)) / |s|
    float c1 = 0.f;
    for (int k = lane; k < D; k += 32) c1 = fmaf(av[k] / act_mag, dz[k], c1);
    c1 = warp_sum(c1);
    for (int k = lane; k < D; k += 32) dz[k] = (dz[k] - (av[k] / act_mag) * c1) / act_mag;
  } else if (res_act == VTC_RESACT_SQUASH) {
    // a = c f(m) s, f(m) = m / (1 + m^2):  ds = c (f da + f'(m)/m s (s . da)), f' = (1-m^2)/(1+m^2)^2
    float c1 = 0.f;
    for (int k = lane; k < D; k += 32) c1 = fmaf(av[k], dz[k], c1);
    c1 = warp_sum(c1);
    const float m2 = act_mag * act_mag, den = 1.f + m2;
    const float f = act_mag / den, fp = (1.f - m2) / (den * den);
    for (int k = lane; k < D; k += 32)
      dz[k] = res_scale * (f * dz[k] + (fp / act_mag) * av[k] * c1);
  } else if (res_act == VTC_RESACT_TANH) {
    for (int k = lane; k < D; k += 32) dz[k] *= 1.f - av[k] * av[k];
  } else if (res_act == VTC_RESACT_AFFINE) {
    if (res_mul)
      for (int k = lane; k < D; k += 32) dz[k] *= res_mul[k];
  }
  __syncwarp();
  // dr = dz (0 if skipped)
  if (mode != VTC_CAM_READOUT_AVG) {
    for (int k = lane; k < D; k += 32) dres[bi * D + k] = skipped ? 0.f : dz[k];
    return;
  }
  // dm = (dr - r (r . dr)) / |m|;  dn_l = dm / L;  dT_l = (dn_l - n_l (n_l . dn_l)) / |T_l|
  float cr = 0.f;
  for (int k = lane; k < D; k += 32) cr = fmaf(rv[k], dz[k], cr);
  cr = warp_sum(cr);
  for (int k = lane; k < D; k += 32)
    zv[k] = skipped ? 0.f : (dz[k] - rv[k] * cr) / (mnorm * (float)L);  // zv now holds dn_l
  __syncwarp();
  for (int l = 0; l < L; ++l) {
    const float* t = T + ((int64_t)l * b + bi) * D;
    float* g = dT + ((int64_t)l * b + bi) * D;
    float tt = 0.f, td = 0.f;
    for (int k = lane; k < D; k += 32) {
      tt = fmaf(t[k], t[k], tt);
      td = fmaf(t[k], zv[k], td);
    }
    tt = warp_sum(tt);
    td = warp_sum(td);
    const float tn = sqrtf(tt);
    for (int k = lane; k < D; k += 32) g[k] = (zv[k] - (t[k] / tn) * (td / tn)) / tn;
  }
}

// --------------------------------------------------------------------------------- launchers
int launch_transpose(const float* in, int64_t R, int64_t C, float* out, cudaStream_t s) {
  if (R == 0 || C == 0) return VTC_OK;
  dim3 grid((unsigned)ceil_div<int64_t>(C, 32), (unsigned)ceil_div<int64_t>(R, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, R, C, out);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_gelu_bwd(const float* dF, const float* U, int64_t n, float* dU, cudaStream_t s) {
  if (n == 0) return VTC_OK;
  const int64_t blocks = ceil_div<int64_t>(n, 256);
  gelu_bwd_kernel<<<(unsigned)(blocks < kNumSMs * 16 ? blocks : kNumSMs * 16), 256, 0, s>>>(dF, U, n, dU);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_colsum(const float* X, int64_t R, int64_t C, float* out, cudaStream_t s) {
  if (C == 0) return VTC_OK;
  colsum_kernel<<<(unsigned)ceil_div<int64_t>(C, 32), 256, 0, s>>>(X, R, C, out);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_layernorm_bwd(const float* dY, const float* X, const float* gamma, int64_t rows, int D,
                         float eps, const float* dres, float* dX, float* dgamma, float* dbeta,
                         cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(dgamma, 0, sizeof(float) * D, s);
  if (e == cudaSuccess) e = cudaMemsetAsync(dbeta, 0, sizeof(float) * D, s);
  if (e != cudaSuccess) return cuda_err(e);
  if (rows == 0) return VTC_OK;
  layernorm_bwd_kernel<<<(unsigned)ceil_div<int64_t>(rows, BW), 256, 0, s>>>(dY, X, gamma, rows, D, eps,
                                                                           dres, dX, dgamma, dbeta);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_cam_attn_core_bwd(const float* QKV, const float* dO, int L, int64_t b, int D, int heads,
                             float* dQKV, cudaStream_t s) {
  if (b == 0) return VTC_OK;
  if (L < 1 || L > 16 || heads < 1 || D % heads || D / heads > 128) return VTC_ERR_UNSUPPORTED_SHAPE;
  const unsigned grid = (unsigned)ceil_div<int64_t>(b * heads, BW);
  const bool narrow = D / heads <= 64;
  if (L <= 8 && narrow)
    cam_attn_core_bwd_kernel<8, 2><<<grid, 256, 0, s>>>(QKV, dO, L, b, D, heads, dQKV);
  else if (L <= 8)
    cam_attn_core_bwd_kernel<8, 4><<<grid, 256, 0, s>>>(QKV, dO, L, b, D, heads, dQKV);
  else if (narrow)
    cam_attn_core_bwd_kernel<16, 2><<<grid, 256, 0, s>>>(QKV, dO, L, b, D, heads, dQKV);
  else
    cam_attn_core_bwd_kernel<16, 4><<<grid, 256, 0, s>>>(QKV, dO, L, b, D, heads, dQKV);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_cam_stack_normalize_bwd(const float* main, const float* aux, const float* dX, int L,
                                   int64_t b, int D, float* dmain, float* daux, cudaStream_t s) {
  const int64_t rows = (int64_t)L * b;
  if (rows == 0) return VTC_OK;
  cam_stack_normalize_bwd_kernel<<<(unsigned)ceil_div<int64_t>(rows, BW), 256, 0, s>>>(main, aux, dX, L,
                                                                                     b, D, dmain, daux);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}
int launch_normalize_bwd(const float* X, const float* dY, int64_t rows, int D, float* dX,
                         cudaStream_t s) {
  if (rows == 0) return VTC_OK;
  normalize_bwd_kernel<<<(unsigned)ceil_div<int64_t>(rows, BW), 256, 0, s>>>(X, dY, rows, D, dX);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

int launch_cam_readout_bwd(const float* T, const float* main, const float* res_in,
                           const uint8_t* skip_mask, const float* dout, int L, int64_t b, int D,
                           int mode, int res_act, float res_scale, const float* res_shift,
                           const float* res_mul, float* dT, float* dres, float* dmain,
                           cudaStream_t s) {
  if (b == 0) return VTC_OK;
  if (D > 1024) return VTC_ERR_UNSUPPORTED_SHAPE;
  if (mode == VTC_CAM_READOUT_UNIFORM) {
    uniform_readout_bwd_kernel<<<(unsigned)ceil_div<int64_t>(b, BW), 256, 0, s>>>(T, dout, L, b, D, dT);
    VTC_LAUNCH_CHECK();
    return VTC_OK;
  }
  const size_t smem = (size_t)BW * 4 * D * sizeof(float);
  auto* kern = &cam_readout_bwd_kernel;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_err(e);
  }
  kern<<<(unsigned)ceil_div<int64_t>(b, BW), 256, smem, s>>>(T, main, res_in, skip_mask, dout, L, b, D,
                                                            mode, res_act, res_scale, res_shift,
                                                            res_mul, dT, dres, dmain);
  VTC_LAUNCH_CHECK();
  return VTC_OK;
}

}  // namespace vtc

using namespace vtc;

extern "C" {

int vtc_transpose(const float* in, int64_t rows, int64_t cols, float* out, vtc_stream_t stream) {
  if (!in || !out || rows < 0 || cols < 0) return VTC_ERR_INVALID_ARG;
  return launch_transpose(in, rows, cols, out, (cudaStream_t)stream);
}
int vtc_gelu_bwd(const float* dF, const float* U, int64_t n, float* dU, vtc_stream_t stream) {
  if (!dF || !U || !dU || n < 0) return VTC_ERR_INVALID_ARG;
  return launch_gelu_bwd(dF, U, n, dU, (cudaStream_t)stream);
}
int vtc_colsum(const float* X, int64_t rows, int64_t cols, float* out, vtc_stream_t stream) {
  if (!X || !out || rows < 0 || cols < 0) return VTC_ERR_INVALID_ARG;
  return launch_colsum(X, rows, cols, out, (cudaStream_t)stream);
}
int vtc_layernorm_bwd(const float* dY, const float* X, const float* gamma, int64_t rows, int D,
                      float eps, const float* dres, float* dX, float* dgamma, float* dbeta,
                      vtc_stream_t stream) {
  if (!dY || !X || !gamma || !dX || !dgamma || !dbeta || rows < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_layernorm_bwd(dY, X, gamma, rows, D, eps, dres, dX, dgamma, dbeta, (cudaStream_t)stream);
}
int vtc_cam_attn_core_bwd(const float* QKV, const float* dO, int L, int64_t b, int D, int heads,
                          float* dQKV, vtc_stream_t stream) {
  if (!QKV || !dO || !dQKV || b < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_cam_attn_core_bwd(QKV, dO, L, b, D, heads, dQKV, (cudaStream_t)stream);
}
int vtc_cam_stack_normalize_bwd(const float* main, const float* aux, const float* dX, int L,
                                int64_t b, int D, float* dmain, float* daux, vtc_stream_t stream) {
  if (!main || !dX || !dmain || L < 1 || (L > 1 && (!aux || !daux)) || b < 0 || D <= 0)
    return VTC_ERR_INVALID_ARG;
  return launch_cam_stack_normalize_bwd(main, aux, dX, L, b, D, dmain, daux, (cudaStream_t)stream);
}
int vtc_normalize_bwd(const float* X, const float* dY, int64_t rows, int D, float* dX,
                      vtc_stream_t stream) {
  if (!X || !dY || !dX || rows < 0 || D <= 0) return VTC_ERR_INVALID_ARG;
  return launch_normalize_bwd(X, dY, rows, D, dX, (cudaStream_t)stream);
}
int vtc_cam_readout_bwd(const float* T, const float* main, const float* res_in,
                        const uint8_t* skip_mask, const float* dout, int L, int64_t b, int D,
                        int mode, int res_act, float res_scale, const float* res_shift,
                        const float* res_mul, float* dT, float* dres, float* dmain,
                        vtc_stream_t stream) {
  if (mode == VTC_CAM_READOUT_UNIFORM) {
    if (!T || !dout || !dT || b < 0 || D <= 0 || L < 1) return VTC_ERR_INVALID_ARG;
    return launch_cam_readout_bwd(T, nullptr, nullptr, nullptr, dout, L, b, D, mode,
                                  VTC_RESACT_NONE, 1.f, nullptr, nullptr, dT, nullptr, nullptr,
                                  (cudaStream_t)stream);
  }
  if (res_act < VTC_RESACT_NONE || res_act > VTC_RESACT_AFFINE ||
      (res_act == VTC_RESACT_AFFINE && !res_shift))
    return VTC_ERR_INVALID_ARG;
  if (!main || !dout || !dmain || b < 0 || D <= 0 || L < 1) return VTC_ERR_INVALID_ARG;
  if (mode == VTC_CAM_READOUT_AVG ? (!T || !dT) : (mode != VTC_CAM_READOUT_RESIDUAL_ONLY || !res_in || !dres))
    return VTC_ERR_INVALID_ARG;
  return launch_cam_readout_bwd(T, main, res_in, skip_mask, dout, L, b, D, mode, res_act, res_scale,
                                res_shift, res_mul, dT, dres, dmain, (cudaStream_t)stream);
}

}  // extern "C"
