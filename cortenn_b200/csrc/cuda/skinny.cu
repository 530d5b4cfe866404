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
/// (tools/gemm_bench.py, tiled kernel -> round 1 -> round 2): 8192x10x1024
/// 75 -> 30 -> 23 us, 1024x10x8192 70 -> 54 -> 37 us, 65536x10x1024
/// 384 -> 176 -> 73 us (3.7 TB/s), 1024x10x65536 356 -> 292 -> 93 us with its
/// fold -- what is left is global-load latency at 12-16 warps per SM.
///

#include <algorithm>
#include <cstdlib>
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

// packed fp32 FMA (SASS FFMA2, see sgemm.cu): two IEEE fmaf per issue slot; a
// pair built from one scalar twice folds into the broadcast operand form
using f32x2 = unsigned long long;

__device__ __forceinline__ f32x2 pack2 (float lo, float hi)
{
	f32x2 r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
	return r;
}

__device__ __forceinline__ void unpack2 (f32x2 v, float& lo, float& hi)
{
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__device__ __forceinline__ void fma2 (f32x2& acc, f32x2 a, f32x2 b)
{
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

/// Sum C values per lane over the 32 lanes with the xor tree 16, 8, 4, 2, 1.
/// While the count is even a stage HALVES it: partners swap the halves they do
/// not keep, so a stage costs C / 2 shuffles instead of C and the tree as a
/// whole ~C instead of 5 C.  Every sum is the same tree of the same operands as
/// the plain all-lanes butterfly (a + b == b + a), i.e. bit-identical to it.
/// On return the lane holds elements [base, base + reduced_count<C, 16>()) in
/// v[0 ..]; `owner` is false on lanes that hold a duplicate.
template <int C, int BIT>
constexpr int reduced_count (void)
{
	if constexpr (BIT < 1) { return C; }
	else if constexpr (0 == C % 2) { return reduced_count<C / 2, BIT / 2>(); }
	else { return reduced_count<C, BIT / 2>(); }
}

template <typename T, int C0, int C, int BIT>
__device__ __forceinline__ void halving_tree (T (&v)[C0], int lane, int& base, bool& owner)
{
	if constexpr (BIT >= 1)
	{
		const bool upper = 0 != (lane & BIT);
		if constexpr (0 == C % 2)
		{
			constexpr int H = C / 2;
#pragma unroll
			for (int j = 0; j < H; ++j)
			{
				T mine = upper ? v[H + j] : v[j];
				T send = upper ? v[j] : v[H + j];
				v[j] = mine + __shfl_xor_sync(0xffffffffu, send, BIT);
			}
			base += upper ? H : 0;
			halving_tree<T,C0,H,BIT / 2>(v, lane, base, owner);
		}
		else
		{
#pragma unroll
			for (int j = 0; j < C; ++j)
			{
				v[j] = v[j] + __shfl_xor_sync(0xffffffffu, v[j], BIT);
			}
			owner = owner && false == upper;
			halving_tree<T,C0,C,BIT / 2>(v, lane, base, owner);
		}
	}
}

// ---- skinny output, A rows contiguous along k: a warp per R rows at a time -------
// B is staged (whole when it fits, else chunk by chunk) in shared memory
// k-major, bs[k / E][k % E][n] (skinny_block): lane l multiplies the
// E = 16 / sizeof(T) consecutive k at E (l + 32 i) of the chunk, so it reads
// 16 bytes of each of its A rows per step (a warp: 512 contiguous bytes per
// row) and NN / E 128-bit vectors of B per k, each holding E neighbouring
// columns (the 16-byte pad per block keeps the lanes of a quarter warp on
// different bank groups).  Per-lane sums run in ascending k; the 32 of them
// fold through a fixed xor tree (halving_tree).
//
// These are bandwidth problems and what decides them is BYTES IN FLIGHT
// (round 1: 1.0-1.5 TB/s with 32 bytes per lane outstanding).  So a warp owns
// R = 4 rows, the loads of the next TWO steps are outstanding while a step is
// multiplied (128-192 bytes per lane, ~100 KB per SM), and the short dimension
// is a template argument (NN = N rounded up to 4) so that padding columns cost
// neither registers nor FMAs and two CTAs stay resident.
// Then issue slots (ncu: 4047 warp instructions per 4 rows, 1536 of them FFMA):
// column pairs share an FFMA2 (B staged k-major so that a row vector holds
// neighbouring columns) and the 32 per-lane sums fold through a halving tree.

// k's of B staged at a time: 512 (fp32) / 256 (fp64), <= 35 KB of shared memory
template <typename T>
constexpr int skinny_kc (void) { return 2048 / (int) sizeof(T); }

/// rows a warp owns: as many as the accumulators leave registers for
template <typename T, int NN>
constexpr int skinny_rows_r (void)
{
	return (int) sizeof(T) * NN <= 48 ? 4 : ((int) sizeof(T) * NN <= 96 ? 2 : 1);
}

/// shared-memory words of B for `k` staged k's: blocks of E k's x NN columns
/// (one block = what a lane multiplies per step), each followed by 16 bytes so
/// that the 8 lanes of a quarter warp hit 8 different 16-byte bank groups
template <typename T, int NN>
__host__ __device__ constexpr uint32_t skinny_block (void)
{
	return (uint32_t) (16 / sizeof(T)) * NN + (uint32_t) (16 / sizeof(T));
}

template <typename T, int NN>
__global__ void __launch_bounds__(256, 2)
skinny_rows_kernel (T* __restrict__ C, const T* __restrict__ A,
	const T* __restrict__ B, GemmParams p, int a_vec, uint32_t kcap)
{
	constexpr int R = skinny_rows_r<T,NN>();
	constexpr int E = 16 / (int) sizeof(T);
	constexpr uint32_t BLOCK = skinny_block<T,NN>();
	constexpr bool F32 = 4 == sizeof(T);
	using V = typename std::conditional<8 == sizeof(T), double2, float4>::type;
	// B is staged `kcap` k's at a time (all of it when it fits), k-major:
	// bs[k / E][k % E][n], see skinny_block
	extern __shared__ __align__(16) unsigned char skinny_smem[];
	T* bs = reinterpret_cast<T*>(skinny_smem);
	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int N = (int) p.N;
	const uint32_t K = (uint32_t) p.K;
	// When B fits the staging area it is staged ONCE and the CTA walks row
	// groups blockIdx.x, + gridDim.x, ... (the launch is then persistent:
	// staging 48 KB per 32 rows cost a third of the kernel); otherwise one
	// row group per CTA and B in chunks.
	const uint64_t ngroups = (p.M + 8 * R - 1) / (8 * R);
	for (uint64_t group = blockIdx.x; group < ngroups; group += gridDim.x)
	{
	const uint64_t row0 = (group * 8 + warp) * R;
	const bool live = row0 < p.M;
	// fp32: pairs of neighbouring columns share an FFMA2 (B's row vectors hold
	// them in consecutive registers, A's value is the broadcast operand)
	T acc[F32 ? 1 : R][F32 ? 1 : NN];
	f32x2 acc2[F32 ? R : 1][F32 ? NN / 2 : 1];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
#pragma unroll
		for (int n = 0; n < NN; ++n)
		{
			if constexpr (F32) { acc2[r][n / 2] = 0ull; }
			else { acc[r][n] = T(0); }
		}
	}
	const T* rowp[R];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
		uint64_t m = row0 + r < p.M ? row0 + r : p.M - 1;
		rowp[r] = A + (int64_t) m * p.a_rs;
	}
	for (uint32_t kc = 0; kc < K; kc += kcap)
	{
		uint32_t len = min(kcap, K - kc);
		uint32_t padded = (len + E - 1) / E * E;
		if (kcap < K || group == blockIdx.x)
		{
		__syncthreads();
		// thread = one k (zero beyond N / len)
		for (uint32_t k = threadIdx.x; k < padded; k += 256)
		{
			const T* brow = B + (int64_t) (kc + k) * p.b_rs;
			T* dst = bs + (k / E) * BLOCK + (k % E) * NN;
#pragma unroll
			for (int n = 0; n < NN; ++n)
			{
				dst[n] = (n < N && k < len) ?
					__ldg(brow + (int64_t) n * p.b_cs) : T(0);
			}
		}
		__syncthreads();
		}
		if (false == live)
		{
			continue;
		}
		auto fetch = [&] (T (&dst)[R][E], uint32_t k0)
		{
			if (k0 >= padded)
			{
				return;
			}
#pragma unroll
			for (int r = 0; r < R; ++r)
			{
				const T* src = rowp[r] + kc + k0;
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
		// one step: the lane's E consecutive k against every column, k ascending
		auto multiply = [&] (const T (&a)[R][E], uint32_t k0)
		{
			const V* blk = reinterpret_cast<const V*>(bs + (k0 / E) * BLOCK);
#pragma unroll
			for (int e = 0; e < E; ++e)
			{
#pragma unroll
				for (int q = 0; q < NN / E; ++q)
				{
					V bv = blk[e * (NN / E) + q];
					if constexpr (F32)
					{
						f32x2 b01 = pack2(bv.x, bv.y);
						f32x2 b23 = pack2(bv.z, bv.w);
#pragma unroll
						for (int r = 0; r < R; ++r)
						{
							f32x2 ar = pack2(a[r][e], a[r][e]);
							fma2(acc2[r][2 * q], ar, b01);
							fma2(acc2[r][2 * q + 1], ar, b23);
						}
					}
					else
					{
#pragma unroll
						for (int r = 0; r < R; ++r)
						{
							acc[r][2 * q] = fma_t(a[r][e], (T) bv.x, acc[r][2 * q]);
							acc[r][2 * q + 1] = fma_t(a[r][e], (T) bv.y, acc[r][2 * q + 1]);
						}
					}
				}
			}
		};
		// three rotating buffers: the loads of the next two steps are
		// outstanding while a step is multiplied
		constexpr uint32_t S = 32 * E;
		uint32_t k0 = (uint32_t) lane * E;
		T b0[R][E], b1[R][E], b2[R][E];
		fetch(b0, k0);
		fetch(b1, k0 + S);
		for (; k0 < padded; k0 += 3 * S)
		{
			fetch(b2, k0 + 2 * S);
			multiply(b0, k0);
			if (k0 + S < padded)
			{
				fetch(b0, k0 + 3 * S);
				multiply(b1, k0 + S);
			}
			if (k0 + 2 * S < padded)
			{
				fetch(b1, k0 + 4 * S);
				multiply(b2, k0 + 2 * S);
			}
		}
	}
	// the 32 per-lane sums of each output fold through the xor tree; every
	// warp of the CTA takes part (shuffles need all lanes), dead ones write nothing
	T v[R * NN];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
#pragma unroll
		for (int n = 0; n < NN; ++n)
		{
			if constexpr (F32)
			{
				if (0 == n % 2)
				{
					unpack2(acc2[r][n / 2], v[r * NN + n], v[r * NN + n + 1]);
				}
			}
			else
			{
				v[r * NN + n] = acc[r][n];
			}
		}
	}
	int base = 0;
	bool owner = true;
	halving_tree<T,R * NN,R * NN,16>(v, lane, base, owner);
	if (live && owner)
	{
		constexpr int LEFT = reduced_count<R * NN,16>();
#pragma unroll
		for (int t = 0; t < LEFT; ++t)
		{
			int r = (base + t) / NN;
			int n = (base + t) % NN;
			if (row0 + r < p.M && n < N)
			{
				C[(int64_t) (row0 + r) * p.c_rs + (int64_t) n * p.c_cs] = v[t];
			}
		}
	}
	} // row groups
}

// ---- skinny output, A contiguous along m (stored [k][m]), long k ---------------------
// thread = E consecutive m (a warp reads 512 contiguous bytes of A per k, one
// 128-bit load per lane; UNROLL k's in flight = 128 bytes per lane), block y =
// one k range whose rows of B sit in shared memory (every lane reads the same
// words: broadcasts); partial sums part[y][m][n] are added in ascending y by
// skinny_fold_kernel.  Per output the k's of a range are added in ascending
// order.

constexpr int SKINNY_COLS_THREADS = 128;

template <typename T, int NN>
__global__ void __launch_bounds__(SKINNY_COLS_THREADS)
skinny_cols_kernel (T* __restrict__ part, const T* __restrict__ A,
	const T* __restrict__ B, GemmParams p, uint32_t k_per_block, int a_vec)
{
	constexpr int E = 16 / (int) sizeof(T);
	constexpr int SKINNY_KC = skinny_kc<T>();
	constexpr int UNROLL = 8;
	using V = typename std::conditional<8 == sizeof(T), double2, float4>::type;
	__shared__ __align__(16) T bs[SKINNY_KC * NN];
	uint64_t m0 = ((uint64_t) blockIdx.x * SKINNY_COLS_THREADS + threadIdx.x) * E;
	const int N = (int) p.N;
	uint32_t k_begin = blockIdx.y * k_per_block;
	uint32_t k_end = min((uint32_t) p.K, k_begin + k_per_block);
	// fp32: the lane's neighbouring rows (m, m + 1) share an FFMA2 -- they sit
	// in consecutive registers of the 128-bit load, B's value is the broadcast
	constexpr bool F32 = 4 == sizeof(T);
	T acc[F32 ? 1 : E][F32 ? 1 : NN];
	f32x2 acc2[F32 ? E / 2 : 1][F32 ? NN : 1];
#pragma unroll
	for (int e = 0; e < E; ++e)
	{
#pragma unroll
		for (int n = 0; n < NN; ++n)
		{
			if constexpr (F32) { acc2[e / 2][n] = 0ull; }
			else { acc[e][n] = T(0); }
		}
	}
	// one k: every column of B's row (read as 128-bit vectors, all lanes the
	// same address: broadcasts) against the lane's E rows
	auto multiply = [&] (const T (&a)[E], const T* brow)
	{
		const V* bq = reinterpret_cast<const V*>(brow);
#pragma unroll
		for (int q = 0; q < NN / E; ++q)
		{
			V bv = bq[q];
			const T* bt = reinterpret_cast<const T*>(&bv);
#pragma unroll
			for (int i = 0; i < E; ++i)
			{
				int n = q * E + i;
				if constexpr (F32)
				{
					f32x2 bb = pack2(bt[i], bt[i]);
					fma2(acc2[0][n], pack2(a[0], a[1]), bb);
					fma2(acc2[1][n], pack2(a[2], a[3]), bb);
				}
				else
				{
#pragma unroll
					for (int e = 0; e < E; ++e)
					{
						acc[e][n] = fma_t(a[e], bt[i], acc[e][n]);
					}
				}
			}
		}
	};
	const bool whole = a_vec && m0 + E <= p.M;   // one aligned 128-bit load per k
	auto fetch = [&] (T (&dst)[E], uint32_t k)
	{
		const T* src = A + (int64_t) k * p.a_cs + m0; // a_rs == 1
		if (whole)
		{
			V v = __ldcs(reinterpret_cast<const V*>(src));
			const T* t = reinterpret_cast<const T*>(&v);
#pragma unroll
			for (int e = 0; e < E; ++e)
			{
				dst[e] = t[e];
			}
		}
		else
		{
#pragma unroll
			for (int e = 0; e < E; ++e)
			{
				dst[e] = m0 + e < p.M ? __ldg(src + e) : T(0);
			}
		}
	};
	for (uint32_t kc = k_begin; kc < k_end; kc += SKINNY_KC)
	{
		uint32_t len = min((uint32_t) SKINNY_KC, k_end - kc);
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < len * NN; i += SKINNY_COLS_THREADS)
		{
			uint32_t k = i / NN;
			uint32_t n = i % NN;
			bs[i] = (int) n < N ?
				__ldg(B + (int64_t) (kc + k) * p.b_rs + (int64_t) n * p.b_cs) : T(0);
		}
		__syncthreads();
		if (m0 >= p.M)
		{
			continue;
		}
		uint32_t k = 0;
		for (; k + UNROLL <= len; k += UNROLL)
		{
			T a[UNROLL][E];
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
			{
				fetch(a[u], kc + k + u);
			}
#pragma unroll
			for (int u = 0; u < UNROLL; ++u)
			{
				multiply(a[u], bs + (k + u) * NN);
			}
		}
		for (; k < len; ++k)
		{
			T a[E];
			fetch(a, kc + k);
			multiply(a, bs + k * NN);
		}
	}
#pragma unroll
	for (int e = 0; e < E; ++e)
	{
		if (m0 + e < p.M)
		{
			T* dst = part + ((size_t) blockIdx.y * p.M + m0 + e) * SKINNY_MAX;
#pragma unroll
			for (int n = 0; n < NN; ++n)
			{
				if constexpr (F32)
				{
					float lo, hi;
					unpack2(acc2[e / 2][n], lo, hi);
					dst[n] = 0 == e % 2 ? lo : hi;
				}
				else
				{
					dst[n] = acc[e][n];
				}
			}
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
	int nn = (int) ((p.N + 3) / 4 * 4);
	if (1 == p.a_cs)
	{
		// A rows run along k: 8 warps x R rows per CTA; B staged whole when
		// it fits 48 KB (two CTAs per SM), else in chunks
		constexpr int E = 16 / (int) sizeof(T);
		int a_vec = (0 == ((uintptr_t) A & 15u) && 0 == p.a_rs % E) ? 1 : 0;
		uint32_t kpad = (uint32_t) ((p.K + E - 1) / E * E);
		// (up to 96 KB: two CTAs per SM still fit)
		uint32_t kfit = (uint32_t) (96 * 1024 / ((nn + 1) * sizeof(T))) / (32 * E) * (32 * E);
		uint32_t kcap = std::min(kpad, std::max<uint32_t>(kfit, 32 * E));
		// (kcap / E blocks of E x nn values + 16 bytes each, see skinny_block)
		size_t smem = (size_t) (kcap / E) * (E * nn + E) * sizeof(T);
		switch (nn)
		{
#define LLO_ROWS(NN) case NN:\
			{\
				constexpr int R = skinny_rows_r<T,NN>();\
				unsigned blocks = (unsigned) ((p.M + 8 * R - 1) / (8 * R));\
				if (kcap >= kpad)\
				{\
					blocks = std::min(blocks, (unsigned) ctx->sm_count * 2);\
				}\
				LLO_CUDA_TRY(cudaFuncSetAttribute(skinny_rows_kernel<T,NN>,\
					cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));\
				skinny_rows_kernel<T,NN><<<blocks, 256, smem, ctx->stream>>>(\
					C, (const T*) A, (const T*) B, p, a_vec, kcap);\
			}\
			break;
			LLO_ROWS(4) LLO_ROWS(8) LLO_ROWS(12) LLO_ROWS(16)
#undef LLO_ROWS
			default: return LLO_OK;
		}
		handled = true;
		return launched(ctx, "skinny_rows_kernel");
	}
	if (1 == p.a_rs)
	{
		// A stored [k][m]: split k over enough blocks to fill the machine
		constexpr int E = 16 / (int) sizeof(T);
		int a_vec = (0 == ((uintptr_t) A & 15u) && 0 == p.a_cs % E) ? 1 : 0;
		uint64_t per_block = (uint64_t) SKINNY_COLS_THREADS * E;
		uint32_t mblocks = (uint32_t) ((p.M + per_block - 1) / per_block);
		// (four CTAs are resident per SM -- 118 registers x 128 threads; 6 per SM
		// ran as 1.5 waves.  Measured, 1024x10x65536 with the fold pass, whose
		// reads grow with the splits: 2 per SM 103 us, 3: 93, 4: 97, 6: 117, 8: 129)
		uint32_t per_sm = 3;
		if (const char* env = std::getenv("LLO_SKINNY_COLS_PER_SM"))
		{
			per_sm = (uint32_t) std::max(1, std::atoi(env)); // experiments only
		}
		uint32_t want = (uint32_t) std::max<uint64_t>(1,
			((uint64_t) ctx->sm_count * per_sm) / mblocks);
		uint32_t k_per_block = (uint32_t) std::max<uint64_t>(64, (p.K + want - 1) / want);
		uint32_t splits = (uint32_t) ((p.K + k_per_block - 1) / k_per_block);
		if (splits > 65535)
		{
			return LLO_OK;
		}
		Scratch scratch(ctx);
		void* ws = nullptr;
		LLO_TRY(scratch.get(&ws, (size_t) splits * p.M * SKINNY_MAX * sizeof(T)));
		dim3 grid(mblocks, splits);
		switch (nn)
		{
#define LLO_COLS(NN) case NN: skinny_cols_kernel<T,NN><<<grid, SKINNY_COLS_THREADS, 0, ctx->stream>>>(\
			(T*) ws, (const T*) A, (const T*) B, p, k_per_block, a_vec); break;
			LLO_COLS(4) LLO_COLS(8) LLO_COLS(12) LLO_COLS(16)
#undef LLO_COLS
			default: return LLO_OK;
		}
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
