///
/// gemm.cuh — shared declarations of the GEMM kernels.
///

#ifndef LLO_CUDA_GEMM_CUH
#define LLO_CUDA_GEMM_CUH

#include "common.cuh"

namespace llo_cuda
{

/// C[m,n] = sum_k A[m,k] * B[k,n]; element (i,j) of X is X[i*x_rs + j*x_cs]
struct GemmParams
{
	uint64_t M;
	uint64_t N;
	uint64_t K;
	int64_t a_rs, a_cs;
	int64_t b_rs, b_cs;
	int64_t c_rs, c_cs;
};

/// sgemm.cu: FFMA fast path for fp32 (handled = false when the layout is
/// not one it supports and the generic kernel should run)
int sgemm_launch (llo_cuda_ctx* ctx, const GemmParams& p, float* C,
	const float* A, const float* B, int precision, bool& handled);

/// tf32gemm.cu: opt-in 3xTF32 tensor-core path (tcgen05 + TMEM) for fp32
int tf32x3_launch (llo_cuda_ctx* ctx, const GemmParams& p, float* C,
	const float* A, const float* B, bool& handled);

/// skinny.cu: products with one very short output dimension (N or M <= 16) and
/// a long k: bandwidth-bound kernels, fp32 / fp64, same contract
template <typename T>
int skinny_launch (llo_cuda_ctx* ctx, const GemmParams& p, T* C, const T* A,
	const T* B, bool& handled);

/// dgemm.cu: DFMA path for fp64, same contract
int dgemm_launch (llo_cuda_ctx* ctx, const GemmParams& p, double* C,
	const double* A, const double* B, bool& handled);

}

#endif // LLO_CUDA_GEMM_CUH
