///
/// gemm_tma.cuh — what the TMA-staged FMA-pipe GEMMs share (sgemm.cu fp32,
/// dgemm.cu fp64): mbarrier / TMA device helpers, tensor-map construction,
/// re-pitching of operands TMA cannot address, split-K with a fixed-order
/// second pass, and the host driver that normalises a GemmParams problem.
///

#ifndef LLO_CUDA_GEMM_TMA_CUH
#define LLO_CUDA_GEMM_TMA_CUH

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <utility>

#include <cuda.h>

#include "gemm.cuh"

namespace llo_cuda
{

namespace tma
{

template <typename T>
struct GemmArgs
{
	T* C;
	int64_t c_rs;      // row stride of C in elements (columns are contiguous)
	uint32_t M;
	uint32_t N;
	uint32_t K;
	uint32_t tiles_m;
	uint32_t tiles_n;
	uint32_t vec_store; // C rows allow aligned vector stores
	// split-K: CTA (tile, y) reduces k-tiles [y * kt_per_split, +kt_per_split)
	// into the partial product C + y * split_stride
	uint32_t kt_per_split;
	int64_t split_stride;
	// persistent kernels: work units [0, full_tiles) are whole tiles summed
	// over every k-tile straight into C; each tile >= full_tiles (the last,
	// partial round of the grid) is cut into tail_splits ranges of
	// kt_per_split k-tiles whose partial products go to
	// part[(tile - full_tiles) * tail_splits + split][BM][BN] and are summed
	// in split order by tile_fold_kernel
	uint32_t ktiles;
	uint32_t full_tiles;
	uint32_t tail_splits;
	T* part;
};

/// Tile number -> tile coordinates: walk GROUP tile-rows per tile-column so
/// the CTAs in flight share their A panels and B panels in L2
__host__ __device__ __forceinline__ void tile_coords (uint32_t tile,
	uint32_t tiles_m, uint32_t tiles_n, uint32_t& tile_m, uint32_t& tile_n)
{
	constexpr uint32_t GROUP = 16;
	uint32_t per_group = GROUP * tiles_n;
	uint32_t first_m = (tile / per_group) * GROUP;
	uint32_t gsize = tiles_m - first_m < GROUP ? tiles_m - first_m : GROUP;
	tile_m = first_m + (tile % per_group) % gsize;
	tile_n = (tile % per_group) / gsize;
}

__device__ __forceinline__ uint32_t smem_u32 (const void* p)
{
	return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init (uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
		:: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx (uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		:: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive (uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
		:: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait (uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra DONE;\n"
		"bra LAB_WAIT;\n"
		"DONE:\n"
		"}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_2d (void* dst, const CUtensorMap* map,
	uint64_t* bar, int c0, int c1)
{
	asm volatile(
		"cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
		" [%0], [%1, {%3, %4}], [%2];"
		:: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
		: "memory");
}


// ---- host side -----------------------------------------------------------------

using EncodeFn = CUresult (*) (CUtensorMap*, CUtensorMapDataType, cuuint32_t,
	void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
	const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
	CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeFn tensor_map_encoder (void)
{
	static EncodeFn fn = nullptr;
	static bool tried = false;
	if (false == tried)
	{
		tried = true;
		void* sym = nullptr;
		cudaDriverEntryPointQueryResult status;
		if (cudaSuccess == cudaGetDriverEntryPoint("cuTensorMapEncodeTiled",
			&sym, cudaEnableDefault, &status) &&
			cudaDriverEntryPointSuccess == status)
		{
			fn = (EncodeFn) sym;
		}
		cudaGetLastError();
	}
	return fn;
}

/// 2-D tensor map: `inner` contiguous elements, `outer` rows of
/// `outer_stride` elements; box = box_inner x box_outer
template <typename T>
bool make_map (CUtensorMap& map, const T* base, uint64_t inner,
	uint64_t outer, int64_t outer_stride, uint32_t box_inner,
	uint32_t box_outer, bool swizzle)
{
	EncodeFn encode = tensor_map_encoder();
	if (nullptr == encode)
	{
		return false;
	}
	cuuint64_t dims[2] = {inner, outer};
	cuuint64_t strides[1] = {(cuuint64_t) outer_stride * sizeof(T)};
	cuuint32_t box[2] = {box_inner, box_outer};
	cuuint32_t estrides[2] = {1, 1};
	CUresult rc = encode(&map, 8 == sizeof(T) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 :
		CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
		(void*) base, dims, strides, box, estrides,
		CU_TENSOR_MAP_INTERLEAVE_NONE,
		swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
		CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
		CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return CUDA_SUCCESS == rc;
}

template <typename T>
bool operand_ok (const T* base, int64_t lead)
{
	return 0 == ((uintptr_t) base & 15u) && lead > 0 &&
		0 == (lead * (int64_t) sizeof(T)) % 16 && lead < (1ll << 37);
}

/// dst[r * dst_ld + c] = src[r * src_ld + c]: re-pitch an operand whose rows
/// TMA cannot address (pitch not a multiple of 16 bytes, e.g. the 10-wide
/// last layer of the MLP)
template <typename T>
__global__ void __launch_bounds__(256)
repitch_kernel (T* __restrict__ dst, const T* __restrict__ src,
	uint64_t rows, uint32_t cols, int64_t src_ld, int64_t dst_ld)
{
	uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	uint64_t r = i / cols;
	uint32_t c = (uint32_t) (i - r * cols);
	if (r < rows)
	{
		dst[r * dst_ld + c] = __ldg(src + r * src_ld + c);
	}
}

/// C[m * c_rs + n] = sum over s (ascending) of part[s][m][n]: the fixed-order
/// second pass of split-K
template <typename T>
__global__ void __launch_bounds__(256)
splitk_fold_kernel (T* __restrict__ C, int64_t c_rs,
	const T* __restrict__ part, uint32_t M, uint32_t N, int64_t ld,
	int64_t split_stride, uint32_t splits)
{
	uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	uint64_t m = i / N;
	uint32_t n = (uint32_t) (i - m * N);
	if (m >= M)
	{
		return;
	}
	const T* p = part + m * ld + n;
	T acc = __ldcs(p);
	for (uint32_t s = 1; s < splits; ++s)
	{
		acc = acc + __ldcs(p + (int64_t) s * split_stride); // no contraction possible
	}
	C[m * c_rs + n] = acc;
}

/// Second pass of the persistent kernels' tail split: C tile = sum over s
/// (ascending) of part[(t * splits + s)][BM][BN]; one CTA per tail tile
template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256)
tile_fold_kernel (T* __restrict__ C, int64_t c_rs, const T* __restrict__ part,
	uint32_t M, uint32_t N, uint32_t tiles_m, uint32_t tiles_n,
	uint32_t full_tiles, uint32_t splits, uint32_t vec_store)
{
	constexpr int E = 16 / (int) sizeof(T);
	using V = typename std::conditional<8 == sizeof(T), double2, float4>::type;
	uint32_t t = blockIdx.x;
	uint32_t tile_m, tile_n;
	tile_coords(full_tiles + t, tiles_m, tiles_n, tile_m, tile_n);
	const T* base = part + (size_t) t * splits * (BM * BN);
	// blockIdx.y strides over the tile's vectors with the threads
	for (uint32_t i = blockIdx.y * 256 + threadIdx.x; i < BM * BN / E;
		i += 256 * gridDim.y)
	{
		uint32_t r = i / (BN / E);
		uint32_t c = (i % (BN / E)) * E;
		uint32_t m = tile_m * BM + r;
		uint32_t n = tile_n * BN + c;
		if (m >= M || n >= N)
		{
			continue;
		}
		V acc = __ldcs(reinterpret_cast<const V*>(base + r * BN + c));
		T* a = reinterpret_cast<T*>(&acc);
		for (uint32_t s = 1; s < splits; ++s)
		{
			V nxt = __ldcs(reinterpret_cast<const V*>(
				base + (size_t) s * (BM * BN) + r * BN + c));
			const T* b = reinterpret_cast<const T*>(&nxt);
#pragma unroll
			for (int e = 0; e < E; ++e)
			{
				a[e] = a[e] + b[e];
			}
		}
		T* dst = C + (int64_t) m * c_rs + n;
		if (vec_store && n + E <= N)
		{
			*reinterpret_cast<V*>(dst) = acc;
		}
		else
		{
#pragma unroll
			for (int e = 0; e < E; ++e)
			{
				if (n + e < N)
				{
					dst[e] = a[e];
				}
			}
		}
	}
}

/// dst[c * dst_ld + r] = src[r * src_ld + c] through 32 x 32 (+1) smem tiles:
/// both sides coalesced.  Used to hand the FMA kernels their fastest operand
/// layout ([k][row], see relayout below).
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel (T* __restrict__ dst, const T* __restrict__ src,
	uint32_t rows, uint32_t cols, int64_t src_ld, int64_t dst_ld)
{
	__shared__ T tile[32][33];
	uint32_t tiles_c = (cols + 31) / 32;
	uint32_t tc = blockIdx.x % tiles_c;
	uint32_t tr = blockIdx.x / tiles_c;
	uint32_t tx = threadIdx.x & 31;
	uint32_t ty = threadIdx.x >> 5;
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		uint32_t r = tr * 32 + ty + 8 * q;
		uint32_t c = tc * 32 + tx;
		if (r < rows && c < cols)
		{
			tile[ty + 8 * q][tx] = __ldcs(src + (int64_t) r * src_ld + c);
		}
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		uint32_t c = tc * 32 + ty + 8 * q;
		uint32_t r = tr * 32 + tx;
		if (r < rows && c < cols)
		{
			dst[(int64_t) c * dst_ld + r] = tile[tx][ty + 8 * q];
		}
	}
}

/// Make `base` (rows x cols, row pitch `lead`) addressable by TMA, copying it
/// to a 16-byte pitched scratch buffer when it is not
template <typename T>
int make_ready (llo_cuda_ctx* ctx, Scratch& scratch, const T*& base,
	int64_t& lead, uint64_t rows, uint64_t cols)
{
	if (operand_ok(base, lead))
	{
		return LLO_OK;
	}
	if (lead < (int64_t) cols && rows > 1)
	{
		return fail(LLO_ERR_FATAL, "gemm operand rows overlap");
	}
	int64_t unit = (int64_t) (16 / sizeof(T));
	int64_t ld = (int64_t) ((cols + unit - 1) / unit * unit);
	void* copy = nullptr;
	LLO_TRY(scratch.get(&copy, (size_t) rows * ld * sizeof(T)));
	uint64_t total = rows * cols;
	repitch_kernel<T><<<(unsigned) ((total + 255) / 256), 256, 0, ctx->stream>>>(
		(T*) copy, base, rows, (uint32_t) cols, lead, ld);
	LLO_TRY(launched(ctx, "repitch_kernel"));
	base = (const T*) copy;
	lead = ld;
	return LLO_OK;
}

/// Normalise the problem (row-major C, contiguous operand directions), build
/// the tensor maps, choose split-K, and call
///   launch(ak, bk, map_a, map_b, args, splits) -> status
/// `vec_a` / `vec_b` give the C store vector width (elements) per B layout.
/// PERSIST: the kernel is persistent (one CTA per SM walking work units,
/// operand pipeline kept full across units) and takes the tail-split plan
/// (full_tiles / tail_splits / part) instead of the (tile, split) grid.
/// DUAL: CTAs per SM of the kernel variant whose operands are both [k][row]
/// tiles (the fp32 one fits two; every other variant runs one)
template <typename T, int BM, int BN, int BK, bool PERSIST = false, int DUAL = 1,
	typename Launch>
int drive (llo_cuda_ctx* ctx, const GemmParams& p0, T* C, const T* A,
	const T* B, bool& handled, int vec_nc, int vec_kc, Launch launch)
{
	handled = false;
	GemmParams p = p0;
	// C must have contiguous columns; a column-major C is the row-major C^T
	// of the transposed product B^T A^T
	if (1 != p.c_cs && p.N > 1)
	{
		if (1 != p.c_rs && p.M > 1)
		{
			return LLO_OK;
		}
		std::swap(p.M, p.N);
		std::swap(A, B);
		int64_t a_rs = p.b_cs, a_cs = p.b_rs, b_rs = p.a_cs, b_cs = p.a_rs;
		p.a_rs = a_rs; p.a_cs = a_cs; p.b_rs = b_rs; p.b_cs = b_cs;
		std::swap(p.c_rs, p.c_cs);
	}
	if (p.M * p.N * p.K < 32768 ||
		p.M >= (1ull << 31) || p.N >= (1ull << 31) || p.K >= (1ull << 31))
	{
		return LLO_OK; // tiny / huge problems take the generic kernel
	}
	if (nullptr == tensor_map_encoder())
	{
		return LLO_OK;
	}
	// a 1-long dimension has no meaningful stride: pick the contiguous form
	if (1 == p.K) { p.a_cs = 1; p.b_rs = 1; }
	if (1 == p.M) { p.a_rs = (1 == p.a_cs) ? (int64_t) p.K : 1; }
	if (1 == p.N) { p.b_cs = (1 == p.b_rs) ? (int64_t) p.K : 1; p.c_cs = 1; }
	bool ak = (1 == p.a_cs);
	bool bk = (1 == p.b_rs);
	if ((false == ak && 1 != p.a_rs) || (false == bk && 1 != p.b_cs))
	{
		return LLO_OK;
	}
	int64_t lead_a = ak ? p.a_rs : p.a_cs;
	int64_t lead_b = bk ? p.b_cs : p.b_rs;
	if (lead_a <= 0 || lead_b <= 0)
	{
		return LLO_OK; // broadcast or mirrored operands
	}
	Scratch scratch(ctx);
	// The kernels run fastest with both operands as [k][row] tiles (row
	// vectors per k: no 4-k register fragments, 119 registers -> two CTAs per
	// SM; 59 vs 50-52 TFLOP/s fp32).  When the other dimension is large
	// enough to amortise one extra pass, a k-contiguous operand is transposed
	// into scratch first (the products and their order are unchanged):
	// 65536x1024x784 50.4 -> 55.1 TFLOP/s including the pass (round 2 A/B,
	// gpurun_out/gemm_ab2.log), break-even near 512.
	auto relayout = [&] (const T*& base, int64_t& lead, bool& kc,
		uint64_t rows, uint64_t other, uint64_t threshold) -> int
	{
		if (false == kc || other < threshold || rows < 64 || p.K < 64)
		{
			return LLO_OK;
		}
		int64_t unit = (int64_t) (16 / sizeof(T));
		int64_t ld = (int64_t) ((rows + unit - 1) / unit * unit);
		void* copy = nullptr;
		LLO_TRY(scratch.get(&copy, (size_t) p.K * ld * sizeof(T)));
		uint64_t blocks = ((rows + 31) / 32) * ((p.K + 31) / 32);
		transpose_kernel<T><<<(unsigned) blocks, 256, 0, ctx->stream>>>(
			(T*) copy, base, (uint32_t) rows, (uint32_t) p.K, lead, ld);
		LLO_TRY(launched(ctx, "transpose_kernel"));
		base = (const T*) copy;
		lead = ld;
		kc = false;
		return LLO_OK;
	};
	// (the pass costs ~8 bytes per operand element and buys ~10% of the
	// 2 * other flops spent on each of them)
	uint64_t relayout_min = 512;
	if (const char* env = std::getenv("LLO_GEMM_RELAYOUT_MIN"))
	{
		relayout_min = (uint64_t) std::atoll(env); // experiments only
	}
	LLO_TRY(relayout(A, lead_a, ak, p.M, p.N, relayout_min));
	LLO_TRY(relayout(B, lead_b, bk, p.N, p.M, relayout_min));
	LLO_TRY(make_ready(ctx, scratch, A, lead_a, ak ? p.M : p.K, ak ? p.K : p.M));
	LLO_TRY(make_ready(ctx, scratch, B, lead_b, bk ? p.N : p.K, bk ? p.K : p.N));
	CUtensorMap ma, mb;
	bool ok = ak ?
		make_map(ma, A, p.K, p.M, lead_a, BK, BM, true) :
		make_map(ma, A, p.M, p.K, lead_a, BM, BK, false);
	ok = ok && (bk ?
		make_map(mb, B, p.K, p.N, lead_b, BK, BN, true) :
		make_map(mb, B, p.N, p.K, lead_b, BN, BK, false));
	if (false == ok)
	{
		return LLO_OK;
	}
	GemmArgs<T> args;
	args.M = (uint32_t) p.M;
	args.N = (uint32_t) p.N;
	args.K = (uint32_t) p.K;
	args.tiles_m = (uint32_t) ((p.M + BM - 1) / BM);
	args.tiles_n = (uint32_t) ((p.N + BN - 1) / BN);
	uint64_t tiles = (uint64_t) args.tiles_m * args.tiles_n;
	if (tiles >= 0x7fffffffull)
	{
		return LLO_OK;
	}
	uint32_t ktiles = (uint32_t) ((p.K + BK - 1) / BK);
	int64_t need = bk ? vec_kc : vec_nc;
	if (PERSIST)
	{
		// Work units are handed round-robin to one CTA per SM.  Whole rounds
		// of the grid run whole tiles; the tiles of the last, partial round
		// may be cut along k into `splits` ranges (summed in range order by a
		// second pass) when that shortens the round.  Cost model in k-tile
		// times: rounds x (k-tiles per unit + ~2 for the epilogue and the
		// pipeline bubble) + 2 for the fold pass.  With fewer tiles than SMs
		// this is plain split-K.
		// (co-resident CTAs share the SM: a round of `sms` units takes as
		// long whether one or two CTAs per SM work through it)
		uint64_t sms = (uint64_t) ctx->sm_count * ((false == ak && false == bk) ? DUAL : 1);
		uint64_t rest = tiles % sms;
		uint32_t splits = 1;
		if (0 != rest)
		{
			uint64_t best = (uint64_t) ktiles + 2;
			uint64_t most = std::min<uint64_t>(ktiles / 4, 64);
			for (uint64_t cand = 2; cand <= most; ++cand)
			{
				uint64_t per = (ktiles + cand - 1) / cand;
				uint64_t real = (ktiles + per - 1) / per;
				uint64_t rounds = (rest * real + sms - 1) / sms;
				uint64_t c = rounds * (per + 2) + 2;
				if (c * 100 < best * 95) // worth it only for a clear gain
				{
					best = c;
					splits = (uint32_t) real;
				}
			}
		}
		if (const char* forced = std::getenv("LLO_GEMM_SPLITS"))
		{
			// experiments only (tools/gemm_bench.py)
			splits = (uint32_t) std::max(1, std::min(std::atoi(forced), (int) ktiles));
		}
		args.ktiles = ktiles;
		args.kt_per_split = (ktiles + splits - 1) / splits;
		splits = (ktiles + args.kt_per_split - 1) / args.kt_per_split;
		args.tail_splits = splits;
		args.full_tiles = splits > 1 ? (uint32_t) (tiles - rest) : (uint32_t) tiles;
		args.part = nullptr;
		args.split_stride = 0;
		args.C = C;
		args.c_rs = p.c_rs;
		args.vec_store = (0 == ((uintptr_t) C & (uintptr_t) (need * sizeof(T) - 1)) &&
			0 == p.c_rs % need) ? 1 : 0;
		uint32_t tail_tiles = (uint32_t) tiles - args.full_tiles;
		if (tail_tiles > 0)
		{
			void* ws = nullptr;
			LLO_TRY(scratch.get(&ws, (size_t) tail_tiles * splits * BM * BN * sizeof(T)));
			args.part = (T*) ws;
		}
		handled = true;
		LLO_TRY(launch(ak, bk, ma, mb, args, splits));
		if (tail_tiles > 0)
		{
			int64_t unit = (int64_t) (16 / sizeof(T));
			uint32_t vec = (0 == ((uintptr_t) C & 15u) && 0 == p.c_rs % unit) ? 1 : 0;
			tile_fold_kernel<T,BM,BN><<<dim3(tail_tiles, 8), 256, 0, ctx->stream>>>(
				C, p.c_rs, args.part, (uint32_t) p.M, (uint32_t) p.N,
				args.tiles_m, args.tiles_n, args.full_tiles, splits, vec);
			LLO_TRY(launched(ctx, "tile_fold_kernel"));
		}
		return LLO_OK;
	}
	args.tail_splits = 1;
	args.full_tiles = (uint32_t) tiles;
	args.ktiles = ktiles;
	args.part = nullptr;
	// Units mode (kernels that take it: DUAL > 1 marks the fp32 tile kernel):
	// with more tiles than CTA slots the last wave may be mostly empty (512
	// tiles on 296 slots: 216 of 296).  Its tiles are then cut along k so that
	// they fill the slots as the whole tiles drain.  Cost in k-tile times of
	// the tail: slots-fuls of units x (k-tiles per unit + 1) against
	// (k-tiles + 1) for leaving it whole.
	if (DUAL > 1 && false == ak && false == bk && nullptr == std::getenv("LLO_GEMM_NO_UNITS"))
	{
		uint64_t slots = (uint64_t) ctx->sm_count * DUAL;
		uint64_t rest = tiles % slots;
		if (tiles > slots && 0 != rest && rest * 100 < slots * 85 && ktiles >= 8)
		{
			uint64_t best = (uint64_t) ktiles + 1;
			uint32_t chosen = 1;
			for (uint64_t cand = 2; cand <= std::min<uint64_t>(ktiles / 3, 16); ++cand)
			{
				uint64_t per = (ktiles + cand - 1) / cand;
				uint64_t real = (ktiles + per - 1) / per;
				uint64_t c = (rest * real + slots - 1) / slots * (per + 1) + 1;
				if (c * 100 < best * 92)
				{
					best = c;
					chosen = (uint32_t) real;
				}
			}
			if (chosen > 1)
			{
				args.kt_per_split = (ktiles + chosen - 1) / chosen;
				args.tail_splits = (ktiles + args.kt_per_split - 1) / args.kt_per_split;
				args.full_tiles = (uint32_t) (tiles - rest);
				args.split_stride = 0;
				args.C = C;
				args.c_rs = p.c_rs;
				args.vec_store = (0 == ((uintptr_t) C & (uintptr_t) (need * sizeof(T) - 1)) &&
					0 == p.c_rs % need) ? 1 : 0;
				void* ws = nullptr;
				LLO_TRY(scratch.get(&ws, (size_t) rest * args.tail_splits * BM * BN * sizeof(T)));
				args.part = (T*) ws;
				handled = true;
				LLO_TRY(launch(ak, bk, ma, mb, args, 0));
				int64_t unit = (int64_t) (16 / sizeof(T));
				uint32_t vec = (0 == ((uintptr_t) C & 15u) && 0 == p.c_rs % unit) ? 1 : 0;
				tile_fold_kernel<T,BM,BN><<<dim3((unsigned) rest, 8), 256, 0, ctx->stream>>>(
					C, p.c_rs, args.part, (uint32_t) p.M, (uint32_t) p.N,
					args.tiles_m, args.tiles_n, args.full_tiles, args.tail_splits, vec);
				LLO_TRY(launched(ctx, "tile_fold_kernel"));
				return LLO_OK;
			}
		}
	}
	// split-K: the k-tiles of every output tile may be spread over `splits`
	// CTAs (partial products summed in split order by a second pass) when
	// that fills the SMs better.  Cost model in k-tile times: waves of CTAs
	// x (k-tiles per CTA + ~3 for pipeline fill and epilogue), plus one for
	// the fold pass.
	uint32_t splits = 1;
	{
		// Measured on the dW shapes of the MLP (round 2, LLO_GEMM_SPLITS sweep,
		// 1024x1024x8192: 2 splits 0.363 ms, 9 splits 0.327, 16 splits 0.324;
		// 784x1024x65536: 5 splits 2.15 ms, 16 splits 2.01): an SM works through
		// its CTAs at the same rate whether one or two are resident, so the time
		// is (CTAs per SM, rounded up) x (k-tiles per CTA + ~1 of fill that the
		// co-resident CTA does not hide), plus the fold pass (~0.35 k-tile times
		// per split for a 1024 x 1024 output).  Edge tiles whose warps lie
		// outside C skip their math (sgemm.cu) and count for the live fraction.
		uint64_t sms = (uint64_t) ctx->sm_count;
		uint64_t live_m = ((p.M % BM) + 63) / 64;     // live 64-row warp halves of the last tile row
		uint64_t live_n = ((p.N % BN) + 31) / 32;     // live 32-column warp quarters of the last tile column
		uint64_t tiles_x8 = 0;                        // tiles in eighths of a tile's warps
		{
			uint64_t full_m = p.M / BM, full_n = p.N / BN;
			tiles_x8 = full_m * full_n * 8 + (0 != p.M % BM ? full_n * live_m * 4 : 0) +
				(0 != p.N % BN ? full_m * live_n * 2 : 0) +
				((0 != p.M % BM && 0 != p.N % BN) ? live_m * live_n : 0);
			if (4 != sizeof(T))
			{
				tiles_x8 = tiles * 8;                 // (only the fp32 kernel skips dead warps)
			}
		}
		uint64_t fold_x100 = 35 * (p.M * p.N) / (1024 * 1024) + 1;
		// Three regimes (all in hundredths of a k-tile time):
		//  * every CTA is resident at once (<= 2 per SM): the SMs holding two
		//    take twice as long as those holding one;
		//  * more CTAs, all alike: they finish in waves, CTAs per SM rounded up;
		//  * more CTAs, some of them short (edge tiles): the short ones break
		//    the waves and the hardware scheduler balances the SMs; a quarter of
		//    a CTA is left over as imbalance
		auto cost = [&] (uint64_t cand) -> uint64_t
		{
			uint64_t per = (ktiles + cand - 1) / cand;
			uint64_t ctas = tiles * cand;
			uint64_t fold = cand > 1 ? cand * fold_x100 : 0;
			if (ctas <= 2 * sms)
			{
				return (ctas > sms ? 2 : 1) * (per + 1) * 100 + fold;
			}
			if (tiles_x8 == tiles * 8)
			{
				return (ctas + sms - 1) / sms * (per + 1) * 100 + fold;
			}
			return tiles_x8 * cand * (ktiles * 100 / cand + 100) / (8 * sms) +
				25 * (per + 1) + fold;
		};
		uint64_t best = cost(1);
		uint64_t most = std::min<uint64_t>(ktiles / 4, 32);
		for (uint64_t cand = 2; cand <= most; ++cand)
		{
			uint64_t c = cost(cand);
			if (c * 100 < best * 99)
			{
				best = c;
				splits = (uint32_t) cand;
			}
		}
	}
	if (const char* forced = std::getenv("LLO_GEMM_SPLITS"))
	{
		// experiments only (tools/gemm_bench.py)
		splits = (uint32_t) std::max(1, std::min(std::atoi(forced), (int) ktiles));
	}
	args.kt_per_split = (ktiles + splits - 1) / splits;
	splits = (ktiles + args.kt_per_split - 1) / args.kt_per_split;
	T* part = nullptr;
	int64_t part_ld = 0;
	if (splits > 1)
	{
		int64_t unit = (int64_t) (16 / sizeof(T));
		part_ld = (int64_t) ((p.N + unit - 1) / unit * unit);
		void* ws = nullptr;
		LLO_TRY(scratch.get(&ws, (size_t) splits * p.M * part_ld * sizeof(T)));
		part = (T*) ws;
		args.C = part;
		args.c_rs = part_ld;
		args.split_stride = (int64_t) p.M * part_ld;
		args.vec_store = 1;
	}
	else
	{
		args.C = C;
		args.c_rs = p.c_rs;
		args.split_stride = 0;
		args.vec_store = (0 == ((uintptr_t) C & (uintptr_t) (need * sizeof(T) - 1)) &&
			0 == p.c_rs % need) ? 1 : 0;
	}
	handled = true;
	LLO_TRY(launch(ak, bk, ma, mb, args, splits));
	if (splits > 1)
	{
		uint64_t total = p.M * p.N;
		splitk_fold_kernel<T><<<(unsigned) ((total + 255) / 256), 256, 0,
			ctx->stream>>>(C, p.c_rs, part, (uint32_t) p.M, (uint32_t) p.N,
			part_ld, args.split_stride, splits);
		LLO_TRY(launched(ctx, "splitk_fold_kernel"));
	}
	return LLO_OK;
}

}

}

#endif // LLO_CUDA_GEMM_TMA_CUH
