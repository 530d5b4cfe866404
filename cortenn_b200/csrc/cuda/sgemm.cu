///
/// sgemm.cu — K6/K7: exact fp32 GEMM on the FFMA pipe, TMA-staged.
///
/// C[m,n] = sum_k A[m,k] * B[k,n], one FFMA chain per output with k ascending
/// (the reference's accumulation order, llo/operator.hpp:378-392 applied to
/// the matmul sub-graph of llo/src/helper.cpp:92-96; the only difference to
/// the CPU is the fused multiply-add rounding).
///
/// Shape of the kernel
///   CTA tile 128 x 128, k-tile 32, 8 warps (2 x 4, warp tile 64 x 32,
///   8 x 8 outputs per thread in registers; two warps per scheduler leave
///   each thread the whole 255-register budget).
///   One elected lane moves both operand tiles with cp.async.bulk.tensor
///   (TMA) into a 4-stage shared-memory ring; full/empty mbarriers hand the
///   stages over and the (cheap) producer duty rotates over the warps, so no
///   thread spends issue slots on global loads or address arithmetic and the
///   FFMA pipe sees ~94% of the instruction stream.  Operand fragments are
///   double-buffered in registers: the LDS of step k+1 fly under the 64
///   FFMAs of step k.
///   Operand layouts (all four combinations, i.e. forward, dW and dX GEMMs):
///     k-contiguous operand  -> smem [row][k], 128-byte TMA swizzle; a thread
///                              reads 4 consecutive k of one row per LDS.128
///                              and owns rows r, r+8, ... so the XOR swizzle
///                              spreads a warp's rows over all banks
///     m/n-contiguous operand -> smem [k][row], unswizzled; a thread reads 4
///                              consecutive rows of one k per LDS.128
///   Out-of-range parts of edge tiles are zero-filled by TMA and add 0.
///

#include "gemm_tma.cuh"

namespace llo_cuda
{

namespace
{

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 32;
constexpr int STAGES = 4;
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int A_STAGE_FLOATS = BM * BK;
constexpr int B_STAGE_FLOATS = BK * BN;
constexpr uint32_t STAGE_BYTES = (A_STAGE_FLOATS + B_STAGE_FLOATS) * 4;
// fragment prefetches run one k / k-group past the end of a stage (the
// values are never used), so the tiles are followed by a pad
constexpr size_t TAIL_PAD = 1024;
constexpr size_t SMEM_BYTES = 1024 /* alignment slack */ +
	(size_t) STAGES * STAGE_BYTES + TAIL_PAD + 2 * STAGES * sizeof(uint64_t);
// k-tiles the TMA loads run ahead of the FFMAs; STAGES - LEAD stages of
// slack keep the refilling warp from ever waiting on a slower warp
constexpr uint32_t LEAD = STAGES - 2;

using SgemmArgs = tma::GemmArgs<float>;
using namespace tma;

/// Packed fp32 pair in one 64-bit register pair.  sm_100a multiplies two
/// independent IEEE fp32 FMAs per issue slot (fma.rn.f32x2 -> SASS FFMA2);
/// a pair built from the same scalar twice costs no instruction: ptxas folds
/// it into the scalar-broadcast operand form (`FFMA2 Rd, Ra.F32, Rb.F32x2, Rc`).
/// Each lane of the pair is the same single-rounding fused multiply-add as
/// fmaf, so results are bit-identical to the scalar kernel's.
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

/// acc(i, j) += a[i] * b[j] for the 8 x 8 outputs of a thread: 32 FFMA2.
/// The pairs run along the operand whose fragment already holds neighbours
/// in consecutive registers (a [k][row] tile read with LDS.128); the other
/// operand is the broadcast scalar.
///   PAIR_I false: acc2[i][jp] = (c[i][2jp], c[i][2jp+1])
///   PAIR_I true : acc2[j][ip] = (c[2ip][j], c[2ip+1][j])
template <bool PAIR_I>
__device__ __forceinline__ void outer8x8 (f32x2 (&acc)[8][4], const float (&a)[8],
	const float (&b)[8])
{
	f32x2 vp[4];
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		vp[q] = PAIR_I ? pack2(a[2 * q], a[2 * q + 1]) : pack2(b[2 * q], b[2 * q + 1]);
	}
#pragma unroll
	for (int s = 0; s < 8; ++s)
	{
		f32x2 sd = PAIR_I ? pack2(b[s], b[s]) : pack2(a[s], a[s]);
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			fma2(acc[s][q], sd, vp[q]);
		}
	}
}

/// the 8 x 8 outputs as scalars
template <bool PAIR_I>
__device__ __forceinline__ void unpack8x8 (float (&c)[8][8], const f32x2 (&acc)[8][4])
{
#pragma unroll
	for (int s = 0; s < 8; ++s)
	{
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			if (PAIR_I)
			{
				unpack2(acc[s][q], c[2 * q][s], c[2 * q + 1][s]);
			}
			else
			{
				unpack2(acc[s][q], c[s][2 * q], c[s][2 * q + 1]);
			}
		}
	}
}

__device__ __forceinline__ float comp (const float4& v, int k)
{
	return 0 == k ? v.x : (1 == k ? v.y : (2 == k ? v.z : v.w));
}

/// Row (A side) / column (B side) a thread owns, relative to the warp tile.
/// k-contiguous operands interleave (see file header), the others block by 4.
template <bool AK>
__device__ __forceinline__ int row_of (int tm, int i)
{
	return AK ? tm + 8 * i : tm * 4 + (i & 3) + 32 * (i >> 2);
}

template <bool BKC>
__device__ __forceinline__ int col_of (int tn, int j)
{
	return BKC ? tn * 2 + (j & 1) + 8 * (j >> 1) :
		tn * 4 + (j & 3) + 16 * (j >> 2);
}

/// Operand fragments of one thread, double-buffered so the loads of the next
/// k (or k-group) are in flight while the current one feeds the FFMAs.
///   KC (k-contiguous, swizzled [row][k] tile): 8 rows x 4 consecutive k per
///       k-group as float4; the next group's 8 vectors are fetched two per k
///   otherwise ([k][row] tile): 8 rows of one k as two float4 per k
template <bool KC>
struct Frag
{
	float4 v[2][KC ? 8 : 2];
};

// Two CTAs per SM when both operands are [k][row] tiles: that variant needs
// 127 registers, so with a 3-stage ring (100 KB) two CTAs fit and every
// scheduler has 4 warps to pick from instead of 2.
template <bool AK, bool BKC>
struct GridShape
{
	static constexpr int CTAS = (AK || BKC) ? 1 : 2;
	static constexpr int ST = (AK || BKC) ? STAGES : 3;
	static constexpr uint32_t AHEAD = ST - 2;
	static constexpr size_t SMEM = 1024 /* alignment slack */ +
		(size_t) ST * STAGE_BYTES + TAIL_PAD + 2 * ST * sizeof(uint64_t);
};

template <bool AK, bool BKC>
__global__ void __launch_bounds__(THREADS, 1)
sgemm_persist_kernel (const __grid_constant__ CUtensorMap map_a,
	const __grid_constant__ CUtensorMap map_b,
	const __grid_constant__ SgemmArgs args)
{
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	// FFMA2 pairs run along i when only A's fragments hold row neighbours
	constexpr bool PAIR_I = BKC && false == AK;
	// 128-byte swizzled tiles must start on a 1024-byte boundary
	// (offset arithmetic on the __shared__ symbol keeps the address space,
	// so tile reads compile to LDS rather than generic LD)
	unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	float* tiles_a = (float*) smem;
	float* tiles_b = tiles_a + STAGES * A_STAGE_FLOATS;
	uint64_t* full = (uint64_t*) (smem + (size_t) STAGES * STAGE_BYTES + TAIL_PAD);
	uint64_t* empty = full + STAGES;

	// Persistent: one CTA per SM walks work units (two CTAs per SM with a
	// 3-stage ring measured slower: 8192x1024x784 49.5 -> 44.8 TFLOP/s).  Units [0, full_tiles)
	// are whole tiles; the others are k ranges of the tail tiles (see
	// GemmArgs).  The operand pipeline runs across unit boundaries: while a
	// unit's last k-tiles are multiplied and its outputs stored, the TMA
	// loads of the next unit are already in flight.
	// Unit sequence of CTA c: c, c + grid, c + 2 grid, ...
	const uint32_t grid = gridDim.x;
	const uint32_t units = args.full_tiles +
		(args.tiles_m * args.tiles_n - args.full_tiles) * args.tail_splits;
	// k-tiles of unit u
	auto unit_ktiles = [&] (uint32_t u) -> uint32_t
	{
		if (u < args.full_tiles)
		{
			return args.ktiles;
		}
		uint32_t split = (u - args.full_tiles) % args.tail_splits;
		return min(args.ktiles - split * args.kt_per_split, args.kt_per_split);
	};

	int warp = threadIdx.x >> 5;
	int lane = threadIdx.x & 31;

	// one elected lane moves both operand tiles of k-tile `kt` of unit `u`
	// into stage `s` with TMA
	auto issue = [&] (uint32_t u, uint32_t kt, uint32_t s)
	{
		uint32_t tile = u, kt_begin = 0;
		if (u >= args.full_tiles)
		{
			uint32_t t = u - args.full_tiles;
			tile = args.full_tiles + t / args.tail_splits;
			kt_begin = (t % args.tail_splits) * args.kt_per_split;
		}
		uint32_t tile_m, tile_n;
		tile_coords(tile, args.tiles_m, args.tiles_n, tile_m, tile_n);
		mbar_expect_tx(&full[s], STAGE_BYTES);
		int k0 = (int) ((kt_begin + kt) * BK);
		if (AK)
		{
			tma_load_2d(tiles_a + s * A_STAGE_FLOATS, &map_a, &full[s],
				k0, (int) (tile_m * BM));
		}
		else
		{
			tma_load_2d(tiles_a + s * A_STAGE_FLOATS, &map_a, &full[s],
				(int) (tile_m * BM), k0);
		}
		if (BKC)
		{
			tma_load_2d(tiles_b + s * B_STAGE_FLOATS, &map_b, &full[s],
				k0, (int) (tile_n * BN));
		}
		else
		{
			tma_load_2d(tiles_b + s * B_STAGE_FLOATS, &map_b, &full[s],
				(int) (tile_n * BN), k0);
		}
	};

	// producer cursor (the k-tile LEAD items ahead of the consumer), tracked
	// by every thread because the duty rotates
	uint32_t c_unit = blockIdx.x;   // the unit being multiplied
	uint32_t p_unit = blockIdx.x;
	uint32_t p_kt = 0;
	uint32_t p_cnt = p_unit < units ? unit_ktiles(p_unit) : 0;
	auto p_advance = [&] ()
	{
		if (++p_kt == p_cnt)
		{
			p_unit += grid;
			p_kt = 0;
			p_cnt = p_unit < units ? unit_ktiles(p_unit) : 0;
		}
	};

	if (0 == threadIdx.x)
	{
		for (int s = 0; s < STAGES; ++s)
		{
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b) : "memory");
	}
	for (uint32_t g = 0; g < LEAD; ++g)
	{
		if (p_unit < units)
		{
			if (0 == threadIdx.x)
			{
				issue(p_unit, p_kt, g);
			}
			p_advance();
		}
	}
	__syncthreads();

	int wm = warp >> 2;   // 0..1
	int wn = warp & 3;    // 0..3
	int tm = lane >> 2;   // 0..7
	int tn = lane & 3;    // 0..3
	f32x2 acc2[8][4];
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			acc2[i][j] = 0ull;
		}
	}

	// thread-constant parts of the fragment addresses (in float4 units)
	//   KC : row * 8 + (kg ^ (row & 7)), rows r0 + 8i resp. cols c(j)
	//   else: k * 32 + row / 4
	const int a_row4 = AK ? (wm * 64 + tm) * (BK / 4) : (wm * 64 + tm * 4) / 4;
	const int b_row4 = BKC ? 0 : (wn * 32 + tn * 4) / 4;
	const int a_key = tm;
	int b_col[8];
	if (BKC)
	{
#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			b_col[j] = wn * 32 + col_of<true>(tn, j);
		}
	}

	uint32_t c_kt = 0;
	uint32_t c_cnt = c_unit < units ? unit_ktiles(c_unit) : 0;
	for (uint32_t g = 0; c_unit < units; ++g)
	{
		uint32_t s = g % STAGES;
		// the producer duty rotates over the warps: item g + LEAD goes into
		// the stage every warp released STAGES - LEAD iterations ago
		if (p_unit < units)
		{
			if ((int) (g % WARPS) == warp && 0 == lane)
			{
				uint32_t pg = g + LEAD;
				if (pg >= STAGES)
				{
					uint32_t prev = pg - STAGES;
					mbar_wait(&empty[prev % STAGES], (prev / STAGES) & 1);
				}
				issue(p_unit, p_kt, pg % STAGES);
			}
			p_advance();
		}
		__syncwarp();
		mbar_wait(&full[s], (g / STAGES) & 1);
		const float4* As4 = (const float4*) (tiles_a + s * A_STAGE_FLOATS);
		const float4* Bs4 = (const float4*) (tiles_b + s * B_STAGE_FLOATS);

		Frag<AK> fa;
		Frag<BKC> fb;
		// ---- prologue: fragments of k-group 0 / k 0 ----
		if (AK)
		{
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				fa.v[0][i] = As4[a_row4 + i * 8 * (BK / 4) + (0 ^ a_key)];
			}
		}
		else
		{
			fa.v[0][0] = As4[a_row4];
			fa.v[0][1] = As4[a_row4 + 8];
		}
		if (BKC)
		{
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				fb.v[0][j] = Bs4[b_col[j] * (BK / 4) + (0 ^ (b_col[j] & 7))];
			}
		}
		else
		{
			fb.v[0][0] = Bs4[b_row4];
			fb.v[0][1] = Bs4[b_row4 + 4];
		}

#pragma unroll 1
		for (int kgp = 0; kgp < BK / 8; ++kgp)
		{
#pragma unroll
			for (int h = 0; h < 2; ++h)
			{
				const int kg = 2 * kgp + h;
				// (the prefetch of the last step reads past the stage: unused)
				const bool more_kg = true;
#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
				{
					const int k = kg * 4 + kk;
					const bool more_k = (kk < 3) || more_kg;
					// ---- prefetch ----
					if (AK)
					{
						if (more_kg)
						{
							int chunk = (kg + 1) ^ a_key;
							fa.v[h ^ 1][2 * kk] = As4[a_row4 +
								(2 * kk) * 8 * (BK / 4) + chunk];
							fa.v[h ^ 1][2 * kk + 1] = As4[a_row4 +
								(2 * kk + 1) * 8 * (BK / 4) + chunk];
						}
					}
					else if (more_k)
					{
						fa.v[(kk + 1) & 1][0] = As4[(k + 1) * (BM / 4) + a_row4];
						fa.v[(kk + 1) & 1][1] = As4[(k + 1) * (BM / 4) + a_row4 + 8];
					}
					if (BKC)
					{
						if (more_kg)
						{
#pragma unroll
							for (int q = 0; q < 2; ++q)
							{
								int j = 2 * kk + q;
								fb.v[h ^ 1][j] = Bs4[b_col[j] * (BK / 4) +
									((kg + 1) ^ (b_col[j] & 7))];
							}
						}
					}
					else if (more_k)
					{
						fb.v[(kk + 1) & 1][0] = Bs4[(k + 1) * (BN / 4) + b_row4];
						fb.v[(kk + 1) & 1][1] = Bs4[(k + 1) * (BN / 4) + b_row4 + 4];
					}
					// ---- this k ----
					float a[8];
					float b[8];
					if (AK)
					{
#pragma unroll
						for (int i = 0; i < 8; ++i)
						{
							a[i] = comp(fa.v[h][i], kk);
						}
					}
					else
					{
						float4 lo = fa.v[kk & 1][0];
						float4 hi = fa.v[kk & 1][1];
						a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
						a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
					}
					if (BKC)
					{
#pragma unroll
						for (int j = 0; j < 8; ++j)
						{
							b[j] = comp(fb.v[h][j], kk);
						}
					}
					else
					{
						float4 lo = fb.v[kk & 1][0];
						float4 hi = fb.v[kk & 1][1];
						b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w;
						b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
					}
					outer8x8<PAIR_I>(acc2, a, b);
				}
			}
		}
		__syncwarp();
		if (0 == lane)
		{
			mbar_arrive(&empty[s]);
		}
		if (++c_kt < c_cnt)
		{
			continue;
		}

		// ---- the unit is complete: registers -> C (or its partial tile) ----
		float acc[8][8];
		unpack8x8<PAIR_I>(acc, acc2);
		if (c_unit >= args.full_tiles)
		{
			// partial product of a tail tile: dense [BM][BN], always in range
			float* ptile = args.part + (size_t) (c_unit - args.full_tiles) * (BM * BN) +
				(wm * 64) * BN + wn * 32;
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				float* prow = ptile + row_of<AK>(tm, i) * BN;
				if (BKC)
				{
#pragma unroll
					for (int j = 0; j < 8; j += 2)
					{
						*(float2*) (prow + col_of<true>(tn, j)) =
							make_float2(acc[i][j], acc[i][j + 1]);
					}
				}
				else
				{
#pragma unroll
					for (int j = 0; j < 8; j += 4)
					{
						*(float4*) (prow + col_of<false>(tn, j)) = make_float4(
							acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
					}
				}
			}
		}
		else
		{
			uint32_t tile_m, tile_n;
			tile_coords(c_unit, args.tiles_m, args.tiles_n, tile_m, tile_n);
			uint32_t m_base = tile_m * BM + wm * 64;
			uint32_t n_base = tile_n * BN + wn * 32;
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				uint32_t m = m_base + row_of<AK>(tm, i);
				if (m >= args.M)
				{
					continue;
				}
				float* crow = args.C + (int64_t) m * args.c_rs;
				if (BKC)
				{
#pragma unroll
					for (int j = 0; j < 8; j += 2)
					{
						uint32_t n = n_base + col_of<true>(tn, j);
						if (args.vec_store && n + 1 < args.N)
						{
							*(float2*) (crow + n) = make_float2(acc[i][j], acc[i][j + 1]);
						}
						else
						{
							if (n < args.N) { crow[n] = acc[i][j]; }
							if (n + 1 < args.N) { crow[n + 1] = acc[i][j + 1]; }
						}
					}
				}
				else
				{
#pragma unroll
					for (int j = 0; j < 8; j += 4)
					{
						uint32_t n = n_base + col_of<false>(tn, j);
						if (args.vec_store && n + 3 < args.N)
						{
							*(float4*) (crow + n) = make_float4(acc[i][j], acc[i][j + 1],
								acc[i][j + 2], acc[i][j + 3]);
						}
						else
						{
#pragma unroll
							for (int e = 0; e < 4; ++e)
							{
								if (n + e < args.N) { crow[n + e] = acc[i][j + e]; }
							}
						}
					}
				}
			}
		}
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				acc2[i][j] = 0ull;
			}
		}
		c_unit += grid;
		c_kt = 0;
		c_cnt = c_unit < units ? unit_ktiles(c_unit) : 0;
	}
}

// ---- one CTA per (tile, split): the hardware hands tiles to SMs -----------------

template <bool AK, bool BKC>
__global__ void __launch_bounds__(THREADS, GridShape<AK,BKC>::CTAS)
sgemm_tma_kernel (const __grid_constant__ CUtensorMap map_a,
	const __grid_constant__ CUtensorMap map_b,
	const __grid_constant__ SgemmArgs args)
{
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	// FFMA2 pairs run along i when only A's fragments hold row neighbours
	constexpr bool PAIR_I = BKC && false == AK;
	// 128-byte swizzled tiles must start on a 1024-byte boundary
	// (offset arithmetic on the __shared__ symbol keeps the address space,
	// so tile reads compile to LDS rather than generic LD)
	unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	constexpr int STAGES = GridShape<AK,BKC>::ST;       // (shadow the file-level ones)
	constexpr uint32_t LEAD = GridShape<AK,BKC>::AHEAD;
	float* tiles_a = (float*) smem;
	float* tiles_b = tiles_a + STAGES * A_STAGE_FLOATS;
	uint64_t* full = (uint64_t*) (smem + (size_t) STAGES * STAGE_BYTES + TAIL_PAD);
	uint64_t* empty = full + STAGES;

	// tile order: walk GROUP tile-rows per tile-column so the CTAs in flight
	// share their A panels and B panels in L2
	uint32_t bid = blockIdx.x;
	uint32_t tile_m, tile_n;
	uint32_t kt_begin = blockIdx.y * args.kt_per_split;
	uint32_t ktiles = min((args.K + BK - 1) / BK - kt_begin, args.kt_per_split);
	// Units mode (tail_splits > 1, a 1-D grid): the first full_tiles CTAs are
	// whole tiles -- whole waves of the chip --, the tiles of the last, partly
	// empty wave are cut along k into tail_splits CTAs each, which the
	// hardware hands out as slots free up; their partial tiles are summed in
	// split order by tile_fold_kernel.
	float* part_tile = nullptr;
	if (args.tail_splits > 1)
	{
		uint32_t tile = bid;
		kt_begin = 0;
		ktiles = args.ktiles;
		if (bid >= args.full_tiles)
		{
			uint32_t t = bid - args.full_tiles;
			tile = args.full_tiles + t / args.tail_splits;
			kt_begin = (t % args.tail_splits) * args.kt_per_split;
			ktiles = min(args.ktiles - kt_begin, args.kt_per_split);
			part_tile = args.part + (size_t) t * (BM * BN);
		}
		tile_coords(tile, args.tiles_m, args.tiles_n, tile_m, tile_n);
	}
	else
	{
		tile_coords(bid, args.tiles_m, args.tiles_n, tile_m, tile_n);
	}

	int warp = threadIdx.x >> 5;
	int lane = threadIdx.x & 31;

	// one elected lane moves both operand tiles of k-tile `kt` with TMA
	auto issue = [&] (uint32_t kt)
	{
		uint32_t s = kt % STAGES;
		mbar_expect_tx(&full[s], STAGE_BYTES);
		int k0 = (int) ((kt_begin + kt) * BK);
		if (AK)
		{
			tma_load_2d(tiles_a + s * A_STAGE_FLOATS, &map_a, &full[s],
				k0, (int) (tile_m * BM));
		}
		else
		{
			tma_load_2d(tiles_a + s * A_STAGE_FLOATS, &map_a, &full[s],
				(int) (tile_m * BM), k0);
		}
		if (BKC)
		{
			tma_load_2d(tiles_b + s * B_STAGE_FLOATS, &map_b, &full[s],
				k0, (int) (tile_n * BN));
		}
		else
		{
			tma_load_2d(tiles_b + s * B_STAGE_FLOATS, &map_b, &full[s],
				(int) (tile_n * BN), k0);
		}
	};

	if (0 == threadIdx.x)
	{
		for (int s = 0; s < STAGES; ++s)
		{
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], WARPS);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&map_a) : "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&map_b) : "memory");
		for (uint32_t kt = 0; kt < LEAD && kt < ktiles; ++kt)
		{
			issue(kt);
		}
	}
	__syncthreads();

	int wm = warp >> 2;   // 0..1
	int wn = warp & 3;    // 0..3
	int tm = lane >> 2;   // 0..7
	int tn = lane & 3;    // 0..3
	// A warp whose 64 x 32 part of an edge tile lies wholly outside C keeps
	// the barrier protocol (and its producer turns) but skips the math: the
	// FMA pipe of its scheduler goes to the warp that shares it, so e.g. the
	// 16-row last tile row of M = 784 costs half a tile, not a whole one.
	const bool live = tile_m * BM + wm * 64 < args.M && tile_n * BN + wn * 32 < args.N;
	f32x2 acc2[8][4];
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			acc2[i][j] = 0ull;
		}
	}

	// thread-constant parts of the fragment addresses (in float4 units)
	//   KC : row * 8 + (kg ^ (row & 7)), rows r0 + 8i resp. cols c(j)
	//   else: k * 32 + row / 4
	const int a_row4 = AK ? (wm * 64 + tm) * (BK / 4) : (wm * 64 + tm * 4) / 4;
	const int b_row4 = BKC ? 0 : (wn * 32 + tn * 4) / 4;
	const int a_key = tm;
	int b_col[8];
	if (BKC)
	{
#pragma unroll
		for (int j = 0; j < 8; ++j)
		{
			b_col[j] = wn * 32 + col_of<true>(tn, j);
		}
	}

	for (uint32_t kt = 0; kt < ktiles; ++kt)
	{
		uint32_t s = kt % STAGES;
		// the producer duty rotates over the warps: k-tile kt + LEAD goes
		// into the stage every warp released STAGES - LEAD iterations ago
		if ((int) (kt % WARPS) == warp && 0 == lane && kt + LEAD < ktiles)
		{
			if (kt + LEAD >= STAGES)
			{
				uint32_t prev = kt + LEAD - STAGES;
				mbar_wait(&empty[prev % STAGES], (prev / STAGES) & 1);
			}
			issue(kt + LEAD);
		}
		__syncwarp();
		mbar_wait(&full[s], (kt / STAGES) & 1);
		const float4* As4 = (const float4*) (tiles_a + s * A_STAGE_FLOATS);
		const float4* Bs4 = (const float4*) (tiles_b + s * B_STAGE_FLOATS);

		if (live)
		{
		Frag<AK> fa;
		Frag<BKC> fb;
		// ---- prologue: fragments of k-group 0 / k 0 ----
		if (AK)
		{
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				fa.v[0][i] = As4[a_row4 + i * 8 * (BK / 4) + (0 ^ a_key)];
			}
		}
		else
		{
			fa.v[0][0] = As4[a_row4];
			fa.v[0][1] = As4[a_row4 + 8];
		}
		if (BKC)
		{
#pragma unroll
			for (int j = 0; j < 8; ++j)
			{
				fb.v[0][j] = Bs4[b_col[j] * (BK / 4) + (0 ^ (b_col[j] & 7))];
			}
		}
		else
		{
			fb.v[0][0] = Bs4[b_row4];
			fb.v[0][1] = Bs4[b_row4 + 4];
		}

#pragma unroll 1
		for (int kgp = 0; kgp < BK / 8; ++kgp)
		{
#pragma unroll
			for (int h = 0; h < 2; ++h)
			{
				const int kg = 2 * kgp + h;
				// (the prefetch of the last step reads past the stage: unused)
				const bool more_kg = true;
#pragma unroll
				for (int kk = 0; kk < 4; ++kk)
				{
					const int k = kg * 4 + kk;
					const bool more_k = (kk < 3) || more_kg;
					// ---- prefetch ----
					if (AK)
					{
						if (more_kg)
						{
							int chunk = (kg + 1) ^ a_key;
							fa.v[h ^ 1][2 * kk] = As4[a_row4 +
								(2 * kk) * 8 * (BK / 4) + chunk];
							fa.v[h ^ 1][2 * kk + 1] = As4[a_row4 +
								(2 * kk + 1) * 8 * (BK / 4) + chunk];
						}
					}
					else if (more_k)
					{
						fa.v[(kk + 1) & 1][0] = As4[(k + 1) * (BM / 4) + a_row4];
						fa.v[(kk + 1) & 1][1] = As4[(k + 1) * (BM / 4) + a_row4 + 8];
					}
					if (BKC)
					{
						if (more_kg)
						{
#pragma unroll
							for (int q = 0; q < 2; ++q)
							{
								int j = 2 * kk + q;
								fb.v[h ^ 1][j] = Bs4[b_col[j] * (BK / 4) +
									((kg + 1) ^ (b_col[j] & 7))];
							}
						}
					}
					else if (more_k)
					{
						fb.v[(kk + 1) & 1][0] = Bs4[(k + 1) * (BN / 4) + b_row4];
						fb.v[(kk + 1) & 1][1] = Bs4[(k + 1) * (BN / 4) + b_row4 + 4];
					}
					// ---- this k ----
					float a[8];
					float b[8];
					if (AK)
					{
#pragma unroll
						for (int i = 0; i < 8; ++i)
						{
							a[i] = comp(fa.v[h][i], kk);
						}
					}
					else
					{
						float4 lo = fa.v[kk & 1][0];
						float4 hi = fa.v[kk & 1][1];
						a[0] = lo.x; a[1] = lo.y; a[2] = lo.z; a[3] = lo.w;
						a[4] = hi.x; a[5] = hi.y; a[6] = hi.z; a[7] = hi.w;
					}
					if (BKC)
					{
#pragma unroll
						for (int j = 0; j < 8; ++j)
						{
							b[j] = comp(fb.v[h][j], kk);
						}
					}
					else
					{
						float4 lo = fb.v[kk & 1][0];
						float4 hi = fb.v[kk & 1][1];
						b[0] = lo.x; b[1] = lo.y; b[2] = lo.z; b[3] = lo.w;
						b[4] = hi.x; b[5] = hi.y; b[6] = hi.z; b[7] = hi.w;
					}
					outer8x8<PAIR_I>(acc2, a, b);
				}
			}
		}
		} // live
		__syncwarp();
		if (0 == lane)
		{
			mbar_arrive(&empty[s]);
		}
	}

	// ---- epilogue: registers -> C ----
	if (false == live)
	{
		return;
	}
	float acc[8][8];
	unpack8x8<PAIR_I>(acc, acc2);
	if (nullptr != part_tile)
	{
		// partial product of a tail tile: dense [BM][BN], always in range
		float* ptile = part_tile + (wm * 64) * BN + wn * 32;
#pragma unroll
		for (int i = 0; i < 8; ++i)
		{
			float* prow = ptile + row_of<AK>(tm, i) * BN;
			if (BKC)
			{
#pragma unroll
				for (int j = 0; j < 8; j += 2)
				{
					*(float2*) (prow + col_of<true>(tn, j)) =
						make_float2(acc[i][j], acc[i][j + 1]);
				}
			}
			else
			{
#pragma unroll
				for (int j = 0; j < 8; j += 4)
				{
					*(float4*) (prow + col_of<false>(tn, j)) = make_float4(
						acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
				}
			}
		}
		return;
	}
	uint32_t m_base = tile_m * BM + wm * 64;
	uint32_t n_base = tile_n * BN + wn * 32;
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
		uint32_t m = m_base + row_of<AK>(tm, i);
		if (m >= args.M)
		{
			continue;
		}
		float* crow = args.C + (int64_t) blockIdx.y * args.split_stride +
			(int64_t) m * args.c_rs;
		if (BKC)
		{
#pragma unroll
			for (int j = 0; j < 8; j += 2)
			{
				uint32_t n = n_base + col_of<true>(tn, j);
				if (args.vec_store && n + 1 < args.N)
				{
					*(float2*) (crow + n) = make_float2(acc[i][j], acc[i][j + 1]);
				}
				else
				{
					if (n < args.N) { crow[n] = acc[i][j]; }
					if (n + 1 < args.N) { crow[n + 1] = acc[i][j + 1]; }
				}
			}
		}
		else
		{
#pragma unroll
			for (int j = 0; j < 8; j += 4)
			{
				uint32_t n = n_base + col_of<false>(tn, j);
				if (args.vec_store && n + 3 < args.N)
				{
					*(float4*) (crow + n) = make_float4(acc[i][j], acc[i][j + 1],
						acc[i][j + 2], acc[i][j + 3]);
				}
				else
				{
#pragma unroll
					for (int e = 0; e < 4; ++e)
					{
						if (n + e < args.N) { crow[n + e] = acc[i][j + e]; }
					}
				}
			}
		}
	}
}

// ---- host side -----------------------------------------------------------------

template <bool AK, bool BKC>
int launch_persist (llo_cuda_ctx* ctx, const CUtensorMap& ma,
	const CUtensorMap& mb, const SgemmArgs& args, uint32_t splits)
{
	constexpr size_t smem = SMEM_BYTES;
	LLO_CUDA_TRY(cudaFuncSetAttribute(sgemm_persist_kernel<AK,BKC>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	(void) splits;
	uint64_t units = (uint64_t) args.full_tiles +
		((uint64_t) args.tiles_m * args.tiles_n - args.full_tiles) * args.tail_splits;
	unsigned grid = (unsigned) std::min<uint64_t>(units, (uint64_t) ctx->sm_count);
	sgemm_persist_kernel<AK,BKC><<<grid, THREADS, smem,
		ctx->stream>>>(ma, mb, args);
	return launched(ctx, "sgemm_persist_kernel");
}

template <bool AK, bool BKC>
int launch_variant (llo_cuda_ctx* ctx, const CUtensorMap& ma,
	const CUtensorMap& mb, const SgemmArgs& args, uint32_t splits)
{
	constexpr size_t smem = GridShape<AK,BKC>::SMEM;
	LLO_CUDA_TRY(cudaFuncSetAttribute(sgemm_tma_kernel<AK,BKC>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	dim3 grid(args.tiles_m * args.tiles_n, splits);
	if (0 == splits)
	{
		// units mode: whole tiles, then the k-split tiles of the last wave
		grid = dim3(args.full_tiles + (args.tiles_m * args.tiles_n - args.full_tiles) *
			args.tail_splits, 1);
	}
	sgemm_tma_kernel<AK,BKC><<<grid, THREADS, smem, ctx->stream>>>(ma, mb, args);
	return launched(ctx, "sgemm_tma_kernel");
}

}

int sgemm_launch (llo_cuda_ctx* ctx, const GemmParams& p, float* C,
	const float* A, const float* B, int precision, bool& handled)
{
	handled = false;
	if (LLO_GEMM_3XTF32 == precision)
	{
		LLO_TRY(tf32x3_launch(ctx, p, C, A, B, handled));
		if (handled)
		{
			return LLO_OK;
		}
		// shapes the tensor-core kernel declines run exactly
	}
	// The persistent kernel keeps the operand pipeline full across tiles and
	// splits only the last partial round of the grid: it wins when tiles are
	// short (few k-tiles, so the per-tile fill / drain of the tile-grid kernel
	// weighs 10-15%) and there is at least one full round; long-k problems run
	// ~4% faster on the tile grid (measured A/B on one box: 8192x1024x784
	// 40.7 -> 44.0, 8192x1024x1024 41.2 -> 45.6, 8192^3 54.8 -> 52.6 TFLOP/s).
	bool persist = false;
	{
		uint64_t tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
		uint64_t tiles_t = ((p.N + BM - 1) / BM) * ((p.M + BN - 1) / BN);
		uint64_t ktiles = (p.K + BK - 1) / BK;
		// (with FFMA2 and the [k][row] relayout the tile grid is ahead again
		// once there are many rounds of tiles: 65536x1024x784 55.1 vs 53.5,
		// 8192x1024x784 45.8 vs 49.2 TFLOP/s tile grid vs persistent)
		uint64_t least = std::min(tiles, tiles_t);
		persist = least >= (uint64_t) ctx->sm_count &&
			least <= 12 * (uint64_t) ctx->sm_count && ktiles <= 64;
	}
	if (const char* env = std::getenv("LLO_GEMM_PERSIST"))
	{
		persist = 0 != std::atoi(env); // experiments (tools/gemm_bench.py)
	}
	if (persist)
	{
		return tma::drive<float,BM,BN,BK,true,1>(ctx, p, C, A, B, handled, 4, 2,
			[ctx] (bool ak, bool bk, const CUtensorMap& ma, const CUtensorMap& mb,
				const SgemmArgs& args, uint32_t splits) -> int
			{
				if (ak)
				{
					return bk ? launch_persist<true,true>(ctx, ma, mb, args, splits) :
						launch_persist<true,false>(ctx, ma, mb, args, splits);
				}
				return bk ? launch_persist<false,true>(ctx, ma, mb, args, splits) :
					launch_persist<false,false>(ctx, ma, mb, args, splits);
			});
	}
	return tma::drive<float,BM,BN,BK,false,2>(ctx, p, C, A, B, handled, 4, 2,
		[ctx] (bool ak, bool bk, const CUtensorMap& ma, const CUtensorMap& mb,
			const SgemmArgs& args, uint32_t splits) -> int
		{
			if (ak)
			{
				return bk ? launch_variant<true,true>(ctx, ma, mb, args, splits) :
					launch_variant<true,false>(ctx, ma, mb, args, splits);
			}
			return bk ? launch_variant<false,true>(ctx, ma, mb, args, splits) :
				launch_variant<false,false>(ctx, ma, mb, args, splits);
		});
}

}
