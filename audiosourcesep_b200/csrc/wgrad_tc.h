// Weight-gradient GEMM on tcgen05 (wgrad_tc.cu): out[i][n] += sum_p A[p][i] * B[p][n] for i < 512, n < n_valid.
#pragma once
#include "common.cuh"

namespace asep {
// A [P, 512] bf16, B [P, ldb] bf16 (ldb a multiple of 64; columns >= n_valid are padding), out [512, ldo] fp32 (+=)
void wgrad_tc(const __nv_bfloat16* A, const __nv_bfloat16* B, int ldb, int n_valid, float* out, int ldo, long long P,
              cudaStream_t s);
}  // namespace asep
