///
/// dgemm.cu — K6/K7: exact fp64 GEMM on the DFMA pipe, TMA-staged.
///
/// Same structure as sgemm.cu (TMA + mbarrier ring, rotating producer duty,
/// register double-buffered fragments, one DFMA chain per output with k
/// ascending), re-sized for 8-byte elements:
///   CTA tile 128 x 64, k-tile 16 (= one 128-byte swizzle row), 8 warps
///   (2 x 4, warp tile 64 x 16), 8 x 4 outputs per thread.
///   k-contiguous operand  -> smem [row][16 k], 128-byte TMA swizzle; a thread
///       reads 2 consecutive k of one row per LDS.128 and owns rows r, r + 8, ...
///   m/n-contiguous operand -> smem [k][row], unswizzled; a thread reads 2
///       consecutive rows of one k per LDS.128
///

#include "gemm_tma.cuh"

namespace llo_cuda
{

namespace
{

using namespace tma;

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int BK = 16;
constexpr int STAGES = 4;
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int A_STAGE_ELEMS = BM * BK;
constexpr int B_STAGE_ELEMS = BK * BN;
constexpr uint32_t STAGE_BYTES = (A_STAGE_ELEMS + B_STAGE_ELEMS) * 8;
constexpr size_t TAIL_PAD = 1024;
constexpr size_t SMEM_BYTES = 1024 + (size_t) STAGES * STAGE_BYTES + TAIL_PAD +
	2 * STAGES * sizeof(uint64_t);
constexpr uint32_t LEAD = STAGES - 2;

using DgemmArgs = GemmArgs<double>;

/// rows (A side, 8 per thread) / columns (B side, 4 per thread) relative to
/// the warp tile; k-contiguous operands interleave, the others block by 2
template <bool AK>
__device__ __forceinline__ int row_of (int tm, int i)
{
	return AK ? tm + 8 * i : tm * 2 + (i & 1) + 16 * (i >> 1);
}

template <bool BKC>
__device__ __forceinline__ int col_of (int tn, int j)
{
	return BKC ? tn + 4 * j : tn * 2 + (j & 1) + 8 * (j >> 1);
}

// (two CTAs per SM when both operands are [k][row] tiles: that variant fits 128
// registers, the 4-stage ring is 100 KB, and four warps per scheduler instead
// of two keep the FP64 pipe fed)
template <bool AK, bool BKC>
__global__ void __launch_bounds__(THREADS, (AK || BKC) ? 1 : 2)
dgemm_tma_kernel (const __grid_constant__ CUtensorMap map_a,
	const __grid_constant__ CUtensorMap map_b,
	const __grid_constant__ DgemmArgs args)
{
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	double* tiles_a = (double*) smem;
	double* tiles_b = tiles_a + STAGES * A_STAGE_ELEMS;
	uint64_t* full = (uint64_t*) (smem + (size_t) STAGES * STAGE_BYTES + TAIL_PAD);
	uint64_t* empty = full + STAGES;

	constexpr uint32_t GROUP = 16;
	uint32_t bid = blockIdx.x;
	uint32_t per_group = GROUP * args.tiles_n;
	uint32_t first_m = (bid / per_group) * GROUP;
	uint32_t gsize = min(args.tiles_m - first_m, GROUP);
	uint32_t tile_m = first_m + (bid % per_group) % gsize;
	uint32_t tile_n = (bid % per_group) / gsize;
	uint32_t kt_begin = blockIdx.y * args.kt_per_split;
	uint32_t ktiles = min((args.K + BK - 1) / BK - kt_begin, args.kt_per_split);

	int warp = threadIdx.x >> 5;
	int lane = threadIdx.x & 31;

	auto issue = [&] (uint32_t kt)
	{
		uint32_t s = kt % STAGES;
		mbar_expect_tx(&full[s], STAGE_BYTES);
		int k0 = (int) ((kt_begin + kt) * BK);
		if (AK)
		{
			tma_load_2d(tiles_a + s * A_STAGE_ELEMS, &map_a, &full[s],
				k0, (int) (tile_m * BM));
		}
		else
		{
			tma_load_2d(tiles_a + s * A_STAGE_ELEMS, &map_a, &full[s],
				(int) (tile_m * BM), k0);
		}
		if (BKC)
		{
			tma_load_2d(tiles_b + s * B_STAGE_ELEMS, &map_b, &full[s],
				k0, (int) (tile_n * BN));
		}
		else
		{
			tma_load_2d(tiles_b + s * B_STAGE_ELEMS, &map_b, &full[s],
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
	double acc[8][4];
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			acc[i][j] = 0.0;
		}
	}

	// fragment addresses in double2 (16-byte) units
	//   KC : row * 8 + (kg ^ (row & 7)),  else: k * (rows / 2) + row / 2
	const int a_row2 = AK ? (wm * 64 + tm) * (BK / 2) : (wm * 64 + tm * 2) / 2;
	const int b_row2 = BKC ? 0 : (wn * 16 + tn * 2) / 2;
	int b_col[4];
	if (BKC)
	{
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			b_col[j] = wn * 16 + col_of<true>(tn, j);
		}
	}

	for (uint32_t kt = 0; kt < ktiles; ++kt)
	{
		uint32_t s = kt % STAGES;
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
		const double2* As2 = (const double2*) (tiles_a + s * A_STAGE_ELEMS);
		const double2* Bs2 = (const double2*) (tiles_b + s * B_STAGE_ELEMS);

		// fragments: KC operands hold 2 k of every row per k-group (fa 8 / fb 4
		// vectors), the others hold one k (fa 4 / fb 2 vectors); two buffers
		double2 fa[2][AK ? 8 : 4];
		double2 fb[2][BKC ? 4 : 2];
		if (AK)
		{
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				fa[0][i] = As2[a_row2 + i * 8 * (BK / 2) + (0 ^ tm)];
			}
		}
		else
		{
#pragma unroll
			for (int q = 0; q < 4; ++q)
			{
				fa[0][q] = As2[a_row2 + 8 * q];
			}
		}
		if (BKC)
		{
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				fb[0][j] = Bs2[b_col[j] * (BK / 2) + (0 ^ (b_col[j] & 7))];
			}
		}
		else
		{
			fb[0][0] = Bs2[b_row2];
			fb[0][1] = Bs2[b_row2 + 4];
		}

#pragma unroll 1
		for (int kgq = 0; kgq < BK / 8; ++kgq)
		{
#pragma unroll
			for (int h4 = 0; h4 < 4; ++h4)
			{
				const int kg = 4 * kgq + h4;   // k-group of 2 k
				const int h = h4 & 1;
#pragma unroll
				for (int kk = 0; kk < 2; ++kk)
				{
					const int k = kg * 2 + kk;
					// ---- prefetch (runs one step past the stage: unused) ----
					if (AK)
					{
						int chunk = (kg + 1) ^ tm;
#pragma unroll
						for (int q = 0; q < 4; ++q)
						{
							int i = 4 * kk + q;
							fa[h ^ 1][i] = As2[a_row2 + i * 8 * (BK / 2) + chunk];
						}
					}
					else
					{
#pragma unroll
						for (int q = 0; q < 4; ++q)
						{
							fa[(kk + 1) & 1][q] = As2[(k + 1) * (BM / 2) + a_row2 + 8 * q];
						}
					}
					if (BKC)
					{
#pragma unroll
						for (int q = 0; q < 2; ++q)
						{
							int j = 2 * kk + q;
							fb[h ^ 1][j] = Bs2[b_col[j] * (BK / 2) +
								((kg + 1) ^ (b_col[j] & 7))];
						}
					}
					else
					{
						fb[(kk + 1) & 1][0] = Bs2[(k + 1) * (BN / 2) + b_row2];
						fb[(kk + 1) & 1][1] = Bs2[(k + 1) * (BN / 2) + b_row2 + 4];
					}
					// ---- this k ----
					double a[8];
					double b[4];
					if (AK)
					{
#pragma unroll
						for (int i = 0; i < 8; ++i)
						{
							a[i] = 0 == kk ? fa[h][i].x : fa[h][i].y;
						}
					}
					else
					{
#pragma unroll
						for (int q = 0; q < 4; ++q)
						{
							a[2 * q] = fa[kk & 1][q].x;
							a[2 * q + 1] = fa[kk & 1][q].y;
						}
					}
					if (BKC)
					{
#pragma unroll
						for (int j = 0; j < 4; ++j)
						{
							b[j] = 0 == kk ? fb[h][j].x : fb[h][j].y;
						}
					}
					else
					{
						b[0] = fb[kk & 1][0].x; b[1] = fb[kk & 1][0].y;
						b[2] = fb[kk & 1][1].x; b[3] = fb[kk & 1][1].y;
					}
#pragma unroll
					for (int i = 0; i < 8; ++i)
					{
#pragma unroll
						for (int j = 0; j < 4; ++j)
						{
							acc[i][j] = fma(a[i], b[j], acc[i][j]);
						}
					}
				}
			}
		}
		__syncwarp();
		if (0 == lane)
		{
			mbar_arrive(&empty[s]);
		}
	}

	// ---- epilogue ----
	uint32_t m_base = tile_m * BM + wm * 64;
	uint32_t n_base = tile_n * BN + wn * 16;
#pragma unroll
	for (int i = 0; i < 8; ++i)
	{
		uint32_t m = m_base + row_of<AK>(tm, i);
		if (m >= args.M)
		{
			continue;
		}
		double* crow = args.C + (int64_t) blockIdx.y * args.split_stride +
			(int64_t) m * args.c_rs;
		if (BKC)
		{
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				uint32_t n = n_base + col_of<true>(tn, j);
				if (n < args.N) { crow[n] = acc[i][j]; }
			}
		}
		else
		{
#pragma unroll
			for (int j = 0; j < 4; j += 2)
			{
				uint32_t n = n_base + col_of<false>(tn, j);
				if (args.vec_store && n + 1 < args.N)
				{
					*(double2*) (crow + n) = make_double2(acc[i][j], acc[i][j + 1]);
				}
				else
				{
					if (n < args.N) { crow[n] = acc[i][j]; }
					if (n + 1 < args.N) { crow[n + 1] = acc[i][j + 1]; }
				}
			}
		}
	}
}

template <bool AK, bool BKC>
int launch_variant (llo_cuda_ctx* ctx, const CUtensorMap& ma,
	const CUtensorMap& mb, const DgemmArgs& args, uint32_t splits)
{
	LLO_CUDA_TRY(cudaFuncSetAttribute(dgemm_tma_kernel<AK,BKC>,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int) SMEM_BYTES));
	dim3 grid(args.tiles_m * args.tiles_n, splits);
	dgemm_tma_kernel<AK,BKC><<<grid, THREADS, SMEM_BYTES,
		ctx->stream>>>(ma, mb, args);
	return launched(ctx, "dgemm_tma_kernel");
}

}

int dgemm_launch (llo_cuda_ctx* ctx, const GemmParams& p, double* C,
	const double* A, const double* B, bool& handled)
{
	return tma::drive<double,BM,BN,BK>(ctx, p, C, A, B, handled, 2, 1,
		[ctx] (bool ak, bool bk, const CUtensorMap& ma, const CUtensorMap& mb,
			const DgemmArgs& args, uint32_t splits) -> int
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
