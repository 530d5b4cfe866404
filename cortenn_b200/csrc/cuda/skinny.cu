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
/// (tools/gemm_bench.py): 8192x10x1024 75 -> 30 us, 1024x10x8192 70 -> 54 us,
/// 65536x10x1024 384 -> 176 us, 1024x10x65536 356 -> 292 us -- still
/// latency-bound (1.0-1.5 TB/s).
///

#include <algorithm>
#include <type_traits>
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

// ---- skinny output, A rows contiguous along k: a warp per R rows at a time -------
// B is staged chunk by chunk (skinny_kc k's) in shared memory TRANSPOSED,
// bs[n][k]: lane l multiplies the E = 16 / sizeof(T) consecutive k at
// E (l + 32 i) of the chunk, so it reads 16 bytes of each of its A rows per
// step (a warp: 512 contiguous bytes per row) and one LDS.128 per column of B
// (consecutive lanes, consecutive 16 bytes: conflict-free).  Per-lane sums run
// in ascending k; the 32 of them fold through a fixed xor-shuffle tree.  Each
// warp walks `groups` row groups per staged chunk to amortise the staging.

// k's of B staged at a time: 512 (fp32) / 256 (fp64), <= 35 KB of shared memory
template <typename T>
constexpr int skinny_kc (void) { return 2048 / (int) sizeof(T); }

template <typename T, int R, int GROUPS>
__global__ void __launch_bounds__(256)
skinny_rows_kernel (T* __restrict__ C, const T* __restrict__ A,
	const T* __restrict__ B, GemmParams p, int a_vec)
{
	constexpr int E = 16 / (int) sizeof(T);
	constexpr int KC = skinny_kc<T>();
	constexpr int PITCH = KC + E;
	using V = typename std::conditional<8 == sizeof(T), double2, float4>::type;
	__shared__ __align__(16) T bs[SKINNY_MAX * PITCH];
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int N = (int) p.N;
	const uint32_t K = (uint32_t) p.K;
	const uint64_t cta_row0 = (uint64_t) blockIdx.x * (8 * R * GROUPS);
	T acc[GROUPS][R][SKINNY_MAX];
#pragma unroll
	for (int g = 0; g < GROUPS; ++g)
	{
#pragma unroll
		for (int r = 0; r < R; ++r)
		{
#pragma unroll
			for (int n = 0; n < SKINNY_MAX; ++n)
			{
				acc[g][r][n] = T(0);
			}
		}
	}
	for (uint32_t kc = 0; kc < K; kc += KC)
	{
		uint32_t len = min((uint32_t) KC, K - kc);
		uint32_t padded = (len + E - 1) / E * E;
		__syncthreads();
		// thread t stages column t % 16 of rows t / 16, + 16, ... (zero beyond N / len)
		{
			int n = threadIdx.x & (SKINNY_MAX - 1);
			for (uint32_t k = threadIdx.x / SKINNY_MAX; k < padded; k += 256 / SKINNY_MAX)
			{
				bs[n * PITCH + k] = (n < N && k < len) ?
					__ldg(B + (int64_t) (kc + k) * p.b_rs + (int64_t) n * p.b_cs) : T(0);
			}
		}
		__syncthreads();
#pragma unroll
		for (int g = 0; g < GROUPS; ++g)
		{
			uint64_t row0 = cta_row0 + (uint64_t) (g * 8 + warp) * R;
			if (row0 >= p.M)
			{
				continue;
			}
			// the loads of step i + 1 are issued before the FMAs of step i
			auto fetch = [&] (T (&dst)[R][E], uint32_t k0)
			{
#pragma unroll
				for (int r = 0; r < R; ++r)
				{
					uint64_t m = row0 + r < p.M ? row0 + r : p.M - 1;
					const T* src = A + (int64_t) m * p.a_rs + kc + k0;
					if (a_vec && k0 + E <= len)
					{
						V v = __ldcs(reinterpret_cast<const V*>(src));
						const T* t = reinterpret_cast<const T*>(&v);
#pragma unroll
						for (int e = 0; e < E; ++e)
						{
							dst[r][e] = t[e];
						}
					}
					else
					{
#pragma unroll
						for (int e = 0; e < E; ++e)
						{
							dst[r][e] = k0 + e < len ? __ldg(src + e) : T(0);
						}
					}
				}
			};
			T nxt[R][E];
			if ((uint32_t) lane * E < padded)
			{
				fetch(nxt, (uint32_t) lane * E);
			}
			for (uint32_t k0 = (uint32_t) lane * E; k0 < padded; k0 += 32 * E)
			{
				T a[R][E];
#pragma unroll
				for (int r = 0; r < R; ++r)
				{
#pragma unroll
					for (int e = 0; e < E; ++e)
					{
						a[r][e] = nxt[r][e];
					}
				}
				if (k0 + 32 * E < padded)
				{
					fetch(nxt, k0 + 32 * E);
				}
#pragma unroll
				for (int n = 0; n < SKINNY_MAX; ++n)
				{
					if (n < N)
					{
						V bv = *reinterpret_cast<const V*>(bs + n * PITCH + k0);
						const T* bt = reinterpret_cast<const T*>(&bv);
#pragma unroll
						for (int e = 0; e < E; ++e)
						{
#pragma unroll
							for (int r = 0; r < R; ++r)
							{
								acc[g][r][n] = fma_t(a[r][e], bt[e], acc[g][r][n]);
							}
						}
					}
				}
			}
		}
	}
#pragma unroll
	for (int g = 0; g < GROUPS; ++g)
	{
		uint64_t row0 = cta_row0 + (uint64_t) (g * 8 + warp) * R;
#pragma unroll
		for (int r = 0; r < R; ++r)
		{
#pragma unroll
			for (int n = 0; n < SKINNY_MAX; ++n)
			{
				if (n < N)
				{
					T v = acc[g][r][n];
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
		// rows per CTA = 8 warps x R x GROUPS; more groups amortise the staging
		// of B but need enough rows to fill the machine
		constexpr int R = 2;
		constexpr int E = 16 / (int) sizeof(T);
		int a_vec = (0 == ((uintptr_t) A & 15u) && 0 == p.a_rs % E) ? 1 : 0;
		uint64_t per1 = 8 * R;
		constexpr int G = 4 == sizeof(T) ? 2 : 1; // (fp64 accumulators of 2 groups would spill)
		if (G > 1 && p.M >= per1 * G * (uint64_t) ctx->sm_count * 2)
		{
			skinny_rows_kernel<T,R,G><<<(unsigned) ((p.M + per1 * G - 1) / (per1 * G)), 256, 0,
				ctx->stream>>>(C, (const T*) A, (const T*) B, p, a_vec);
		}
		else
		{
			skinny_rows_kernel<T,R,1><<<(unsigned) ((p.M + per1 - 1) / per1), 256, 0,
				ctx->stream>>>(C, (const T*) A, (const T*) B, p, a_vec);
		}
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
