///
/// gemm.cu — K6/K7: the GEMM the composed matmul sub-graph lowers to.
///
/// age::matmul builds reduce_sum(mul(permute(extend(a)), permute(extend(b))), 2)
/// (llo/src/helper.cpp:92-96): five [M,N,C] tensors.  The planner recognises
/// that contraction (and the ones its gradients produce) and calls this
/// entry point instead; nothing of size M*N*C is ever materialised.
///
/// Generic path (every dtype): shared-memory tiled kernel, one output per
/// thread, k ascending — the reference's accumulation order, so integer
/// results are bit-exact and floating results differ from the reference only
/// by FMA contraction.
///

#include "gemm.cuh"
#include "ops.cuh"

namespace llo_cuda
{

/// C[m,n] = sum_k A[m,k] * B[k,n]; 32x32 tile, 32x32 threads' worth of
/// outputs computed by 256 threads (4 rows each)
template <typename T>
__global__ void __launch_bounds__(256)
gemm_generic_kernel (T* __restrict__ C, const T* __restrict__ A,
	const T* __restrict__ B, const __grid_constant__ GemmParams p)
{
	constexpr int TILE = 32;
	__shared__ T sa[TILE][TILE + 1];
	__shared__ T sb[TILE][TILE + 1];
	int tx = threadIdx.x & 31;
	int ty = threadIdx.x >> 5; // 0..7
	uint64_t m0 = (uint64_t) blockIdx.y * TILE;
	uint64_t n0 = (uint64_t) blockIdx.x * TILE;
	using P = typename promoted<T>::type;
	P acc[4] = {0, 0, 0, 0};
	for (uint64_t k0 = 0; k0 < p.K; k0 += TILE)
	{
#pragma unroll
		for (int r = 0; r < 4; ++r)
		{
			int row = ty * 4 + r;
			uint64_t m = m0 + row;
			uint64_t k = k0 + tx;
			sa[row][tx] = (m < p.M && k < p.K) ?
				A[(int64_t) m * p.a_rs + (int64_t) k * p.a_cs] : T();
			uint64_t kb = k0 + row;
			uint64_t n = n0 + tx;
			sb[row][tx] = (kb < p.K && n < p.N) ?
				B[(int64_t) kb * p.b_rs + (int64_t) n * p.b_cs] : T();
		}
		__syncthreads();
#pragma unroll 8
		for (int kk = 0; kk < TILE; ++kk)
		{
			P bv = (P) sb[kk][tx];
#pragma unroll
			for (int r = 0; r < 4; ++r)
			{
				if constexpr (std::is_same<T,float>::value)
				{
					acc[r] = fmaf(sa[ty * 4 + r][kk], bv, acc[r]);
				}
				else if constexpr (std::is_same<T,double>::value)
				{
					acc[r] = fma(sa[ty * 4 + r][kk], bv, acc[r]);
				}
				else
				{
					acc[r] += (P) sa[ty * 4 + r][kk] * bv;
				}
			}
		}
		__syncthreads();
	}
#pragma unroll
	for (int r = 0; r < 4; ++r)
	{
		uint64_t m = m0 + ty * 4 + r;
		uint64_t n = n0 + tx;
		if (m < p.M && n < p.N)
		{
			C[(int64_t) m * p.c_rs + (int64_t) n * p.c_cs] = (T) acc[r];
		}
	}
}

}

using namespace llo_cuda;

extern "C" int llo_cuda_gemm (llo_cuda_ctx* ctx, int dtype,
	uint64_t M, uint64_t N, uint64_t K,
	const void* A, int64_t a_rs, int64_t a_cs,
	const void* B, int64_t b_rs, int64_t b_cs,
	void* C, int64_t c_rs, int64_t c_cs, int precision)
{
	if (0 == M || 0 == N || 0 == K)
	{
		return fail(LLO_ERR_FATAL, "cannot gemm with an empty dimension");
	}
	GemmParams p = {M, N, K, a_rs, a_cs, b_rs, b_cs, c_rs, c_cs};
	if (LLO_DT_FLOAT == dtype || (LLO_DT_DOUBLE == dtype && LLO_GEMM_EXACT == precision))
	{
		bool handled = false;
		if (LLO_DT_FLOAT == dtype)
		{
			LLO_TRY(skinny_launch<float>(ctx, p, (float*) C, (const float*) A,
				(const float*) B, handled));
		}
		else
		{
			LLO_TRY(skinny_launch<double>(ctx, p, (double*) C, (const double*) A,
				(const double*) B, handled));
		}
		if (handled)
		{
			return LLO_OK;
		}
	}
	if (LLO_DT_FLOAT == dtype)
	{
		bool handled = false;
		LLO_TRY(sgemm_launch(ctx, p, (float*) C, (const float*) A,
			(const float*) B, precision, handled));
		if (handled)
		{
			return LLO_OK;
		}
	}
	else if (LLO_GEMM_EXACT != precision)
	{
		return fail(LLO_ERR_FATAL, "the 3xTF32 path is defined for FLOAT only");
	}
	else if (LLO_DT_DOUBLE == dtype)
	{
		bool handled = false;
		LLO_TRY(dgemm_launch(ctx, p, (double*) C, (const double*) A,
			(const double*) B, handled));
		if (handled)
		{
			return LLO_OK;
		}
	}
	dim3 grid((unsigned) ((N + 31) / 32), (unsigned) ((M + 31) / 32));
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		gemm_generic_kernel<T><<<grid, 256, 0, ctx->stream>>>(
			(T*) C, (const T*) A, (const T*) B, p);
		return launched(ctx, "gemm_generic_kernel");
	})
	return LLO_OK;
}
