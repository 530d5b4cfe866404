///
/// skinny.cu — GEMMs with one very short dimension (fp32 / fp64).
///
/// The 784-1024-1024-10 MLP of BASELINE.json config 5 multiplies by a
/// 10-wide layer three times per gradient evaluation:
///   forward  C[B,10]    = h[B,1024] W[1024,10]         skinny output, A k-contiguous
///   dW       C[1024,10] = h^T[1024,B] d[B,10]          skinny output, A m-contiguous, long k
///   dX       C[B,1024]  = d[B,10] W^T[10,1024]         short k (stays on the tiled kernels:
///                                                       a register-blocked version measured slower)
/// On 128 x 128 tiles more than 90% of the FMAs of the first two multiply
/// padding.  They are bandwidth problems; these kernels read the long operand
/// once, coalesced, stage the short one in shared memory, and sum in a fixed
/// order (run-to-run deterministic; per-lane chains in ascending k, then a
/// fixed shuffle tree / an ascending fold over k ranges).  Measured
/// (tools/gemm_bench.py, same box): 8192x10x1024 75 -> 43 us, 1024x10x8192
/// 70 -> 54 us, 65536x10x1024 384 -> 273 us, 1024x10x65536 356 -> 292 us --
/// still latency-bound (about 1 TB/s), a round-2 item.
///

#include <algorithm>
#include <utility>

#include "gemm.cuh"

namespace llo_cuda
{

namespace
{

constexpr int SKINNY_MAX = 16;   // the short dimension is at most this

template <typename T>
__device__ __forceinline__ T fma_t (T a, T b, T c)
{
	if constexpr (sizeof(T) == 4) { return fmaf(a, b, c); }
	else { return fma(a, b, c); }
}

// ---- skinny output, A rows contiguous along k: one warp per R rows ----------------
// B is staged chunk by chunk (SKINNY_KC k's) in shared memory with an odd row
// pitch; lane l multiplies k = l, l + 32, ... of the chunk (so a warp reads 128
// contiguous bytes of each A row per step and conflict-free words of B), the
// per-lane sums run in ascending k and the 32 of them fold through a fixed
// xor-shuffle tree

// k's of B staged at a time: 512 (fp32) / 256 (fp64), <= 35 KB of shared memory
template <typename T>
constexpr int skinny_kc (void) { return 2048 / (int) sizeof(T); }
constexpr int SKINNY_PITCH = SKINNY_MAX + 1;

template <typename T, int R>
__global__ void __launch_bounds__(256)
skinny_rows_kernel (T* __restrict__ C, const T* __restrict__ A,
	const T* __restrict__ B, GemmParams p)
{
	constexpr int SKINNY_KC = skinny_kc<T>();
	__shared__ T bs[SKINNY_KC * SKINNY_PITCH];
	const int lane = threadIdx.x & 31;
	const uint64_t row0 = ((uint64_t) blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
	const int N = (int) p.N;
	const uint32_t K = (uint32_t) p.K;
	T acc[R][SKINNY_MAX];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
#pragma unroll
		for (int n = 0; n < SKINNY_MAX; ++n)
		{
			acc[r][n] = T(0);
		}
	}
	const T* arow[R];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
		uint64_t m = row0 + r < p.M ? row0 + r : p.M - 1;
		arow[r] = A + (int64_t) m * p.a_rs;
	}
	for (uint32_t kc = 0; kc < K; kc += SKINNY_KC)
	{
		uint32_t len = min((uint32_t) SKINNY_KC, K - kc);
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < len * (uint32_t) N; i += 256)
		{
			uint32_t k = i / (uint32_t) N;
			uint32_t n = i - k * (uint32_t) N;
			bs[k * SKINNY_PITCH + n] =
				__ldg(B + (int64_t) (kc + k) * p.b_rs + (int64_t) n * p.b_cs);
		}
		__syncthreads();
#pragma unroll 2
		for (uint32_t k = (uint32_t) lane; k < len; k += 32)
		{
			T a[R];
#pragma unroll
			for (int r = 0; r < R; ++r)
			{
				a[r] = __ldcs(arow[r] + kc + k);
			}
			const T* brow = bs + k * SKINNY_PITCH;
#pragma unroll
			for (int n = 0; n < SKINNY_MAX; ++n)
			{
				if (n < N)
				{
					T bv = brow[n];
#pragma unroll
					for (int r = 0; r < R; ++r)
					{
						acc[r][n] = fma_t(a[r], bv, acc[r][n]);
					}
				}
			}
		}
	}
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
#pragma unroll
		for (int n = 0; n < SKINNY_MAX; ++n)
		{
			if (n < N)
			{
				T v = acc[r][n];
#pragma unroll
				for (int off = 16; off > 0; off >>= 1)
				{
					v = v + __shfl_xor_sync(0xffffffffu, v, off);
				}
				if (0 == lane && row0 + r < p.M)
				{
					C[(int64_t) (row0 + r) * p.c_rs + (int64_t) n * p.c_cs] = v;
				}
			}
		}
	}
}

// ---- skinny output, A contiguous along m (stored [k][m]), long k ---------------------
// thread = one m (a warp reads 128 contiguous bytes of A per k), block y = one
// k range whose rows of B sit in shared memory (every lane reads the same
// word: a broadcast); partial sums part[y][m][n] are added in ascending y by
// skinny_fold_kernel

constexpr int SKINNY_COLS_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(SKINNY_COLS_THREADS)
skinny_cols_kernel (T* __restrict__ part, const T* __restrict__ A,
	const T* __restrict__ B, GemmParams p, uint32_t k_per_block)
{
	constexpr int SKINNY_KC = skinny_kc<T>();
	__shared__ T bs[SKINNY_KC * SKINNY_MAX];
	uint64_t m = (uint64_t) blockIdx.x * SKINNY_COLS_THREADS + threadIdx.x;
	const int N = (int) p.N;
	uint32_t k_begin = blockIdx.y * k_per_block;
	uint32_t k_end = min((uint32_t) p.K, k_begin + k_per_block);
	T acc[SKINNY_MAX];
#pragma unroll
	for (int n = 0; n < SKINNY_MAX; ++n)
	{
		acc[n] = T(0);
	}
	const T* acol = A + (m < p.M ? m : p.M - 1); // a_rs == 1
	for (uint32_t kc = k_begin; kc < k_end; kc += SKINNY_KC)
	{
		uint32_t len = min((uint32_t) SKINNY_KC, k_end - kc);
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < len * SKINNY_MAX; i += SKINNY_COLS_THREADS)
		{
			uint32_t k = i / SKINNY_MAX;
			uint32_t n = i % SKINNY_MAX;
			bs[i] = (int) n < N ?
				__ldg(B + (int64_t) (kc + k) * p.b_rs + (int64_t) n * p.b_cs) : T(0);
		}
		__syncthreads();
#pragma unroll 8
		for (uint32_t k = 0; k < len; ++k)
		{
			T a = __ldcs(acol + (int64_t) (kc + k) * p.a_cs);
			const T* brow = bs + k * SKINNY_MAX;
#pragma unroll
			for (int n = 0; n < SKINNY_MAX; ++n)
			{
				if (n < N)
				{
					acc[n] = fma_t(a, brow[n], acc[n]);
				}
			}
		}
	}
	if (m < p.M)
	{
		T* dst = part + ((size_t) blockIdx.y * p.M + m) * SKINNY_MAX;
#pragma unroll
		for (int n = 0; n < SKINNY_MAX; ++n)
		{
			dst[n] = acc[n];
		}
	}
}

template <typename T>
__global__ void __launch_bounds__(256)
skinny_fold_kernel (T* __restrict__ C, const T* __restrict__ part, GemmParams p,
	uint32_t splits)
{
	uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	uint64_t m = i / SKINNY_MAX;
	int n = (int) (i % SKINNY_MAX);
	if (m >= p.M || n >= (int) p.N)
	{
		return;
	}
	const T* src = part + m * SKINNY_MAX + n;
	T acc = src[0];
	for (uint32_t s = 1; s < splits; ++s)
	{
		acc = acc + src[(size_t) s * p.M * SKINNY_MAX];
	}
	C[(int64_t) m * p.c_rs + (int64_t) n * p.c_cs] = acc;
}

/// the transposed problem C^T = B^T A^T
void transpose_problem (GemmParams& p, const void*& A, const void*& B)
{
	std::swap(p.M, p.N);
	std::swap(A, B);
	int64_t a_rs = p.b_cs, a_cs = p.b_rs, b_rs = p.a_cs, b_cs = p.a_rs;
	p.a_rs = a_rs; p.a_cs = a_cs; p.b_rs = b_rs; p.b_cs = b_cs;
	std::swap(p.c_rs, p.c_cs);
}

}

template <typename T>
int skinny_launch (llo_cuda_ctx* ctx, const GemmParams& p0, T* C, const T* A0,
	const T* B0, bool& handled)
{
	handled = false;
	GemmParams p = p0;
	const void* A = A0;
	const void* B = B0;
	if (p.M >= (1ull << 31) || p.N >= (1ull << 31) || p.K >= (1ull << 31))
	{
		return LLO_OK;
	}
	// ---- skinny output, long k --------------------------------------------------------
	if (std::min(p.M, p.N) > SKINNY_MAX || p.K < 64 || std::max(p.M, p.N) < 256)
	{
		return LLO_OK;
	}
	if (p.N > SKINNY_MAX)
	{
		transpose_problem(p, A, B); // the short dimension becomes N
	}
	if (1 == p.a_cs)
	{
		// A rows run along k
		constexpr int R = 4;
		uint64_t warps = (p.M + R - 1) / R;
		skinny_rows_kernel<T,R><<<(unsigned) ((warps + 7) / 8), 256, 0,
			ctx->stream>>>(C, (const T*) A, (const T*) B, p);
		handled = true;
		return launched(ctx, "skinny_rows_kernel");
	}
	if (1 == p.a_rs)
	{
		// A stored [k][m]: split k over enough blocks to fill the machine
		uint32_t mblocks = (uint32_t) ((p.M + SKINNY_COLS_THREADS - 1) / SKINNY_COLS_THREADS);
		uint32_t want = (uint32_t) std::max<uint64_t>(1,
			((uint64_t) ctx->sm_count * 4 + mblocks - 1) / mblocks);
		uint32_t k_per_block = (uint32_t) std::max<uint64_t>(64, (p.K + want - 1) / want);
		uint32_t splits = (uint32_t) ((p.K + k_per_block - 1) / k_per_block);
		if (splits > 65535)
		{
			return LLO_OK;
		}
		Scratch scratch(ctx);
		void* ws = nullptr;
		LLO_TRY(scratch.get(&ws, (size_t) splits * p.M * SKINNY_MAX * sizeof(T)));
		skinny_cols_kernel<T><<<dim3(mblocks, splits), SKINNY_COLS_THREADS, 0, ctx->stream>>>(
			(T*) ws, (const T*) A, (const T*) B, p, k_per_block);
		LLO_TRY(launched(ctx, "skinny_cols_kernel"));
		uint64_t total = p.M * SKINNY_MAX;
		skinny_fold_kernel<T><<<(unsigned) ((total + 255) / 256), 256, 0,
			ctx->stream>>>(C, (const T*) ws, p, splits);
		handled = true;
		return launched(ctx, "skinny_fold_kernel");
	}
	return LLO_OK;
}

template int skinny_launch<float> (llo_cuda_ctx*, const GemmParams&, float*,
	const float*, const float*, bool&);
template int skinny_launch<double> (llo_cuda_ctx*, const GemmParams&, double*,
	const double*, const double*, bool&);

}
