// cam.cuh -- launchers of the Context Adapter Module kernels (cam.cu).
#pragma once
#include "common.cuh"

namespace vtc {

int launch_cam_stack_normalize(const float* main, const float* aux, int L, int64_t b, int D,
                               float* X, cudaStream_t s);
int launch_layernorm(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                     float eps, float* Y, cudaStream_t s);
// out (fp32 [L,b,D]) and/or out_op (bf16 query-side operand [L*b, Kp], split = [hi|hi|lo])
int launch_cam_attn_core(const float* QKV, int L, int64_t b, int D, int heads, float* out,
                         __nv_bfloat16* out_op, int Kp, int split, cudaStream_t s);
int launch_layernorm_prep(const float* X, const float* gamma, const float* beta, int64_t rows, int D,
                          float eps, int split, __nv_bfloat16* out, int Kp, cudaStream_t s);
int launch_bias_act(const float* X, const float* bias, const float* residual, int64_t rows, int D,
                    int act, float* Y, cudaStream_t s);
int launch_cam_readout(const float* T, const float* main, const float* res_in,
                       const uint8_t* skip_mask, int L, int64_t b, int D, int mode, int res_act,
                       float res_scale, const float* res_shift, const float* res_mul, float* out,
                       cudaStream_t s);

}  // namespace vtc
