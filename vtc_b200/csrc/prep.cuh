// prep.cuh -- launchers of the row kernels (prep.cu).
#pragma once
#include "common.cuh"

namespace vtc {

enum { PREP_PLAIN = 0, PREP_SPLIT_A = 1, PREP_SPLIT_B = 2 };

int launch_row_norms(const void* X, bool bf16, int64_t rows, int D, int64_t ld, float* inv_norm,
                     float* sq_norm, cudaStream_t s);
int launch_normalize(const void* X, bool bf16, int64_t rows, int D, int64_t ldx, void* Y,
                     int64_t ldy, cudaStream_t s);
int launch_prep_operand(const void* X, bool bf16, int64_t rows, int D, int64_t ldx, int mode,
                        __nv_bfloat16* out, int Kp, cudaStream_t s);

}  // namespace vtc
