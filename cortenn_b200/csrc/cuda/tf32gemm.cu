///
/// tf32gemm.cu — K6, opt-in: fp32 GEMM on the 5th-generation tensor cores by
/// the 3xTF32 split (precision = LLO_GEMM_3XTF32; the exact FFMA kernel of
/// sgemm.cu stays the default).
///
///   a = a_hi + a_lo,  a_hi = tf32(a) (round to nearest), a_lo = tf32(a - a_hi)
///   C = sum_k a_hi b_hi + a_hi b_lo + a_lo b_hi        (fp32 accumulation)
/// The dropped a_lo b_lo term and the tf32 rounding of the low parts are
/// O(2^-22) relative to |a||b|: results agree with the exact kernel to ~1e-6
/// of sum |a||b| (tests/test_gpu_ops.py), not bit-for-bit — hence opt-in.
///
/// Two kernels:
///   tf32_split_kernel   one pass over an operand: writes its hi and lo parts
///                       as K-MAJOR [rows][K] arrays whatever the operand's own
///                       layout (transposing through shared memory when its
///                       rows are the contiguous direction), so the tensor-core
///                       kernel sees exactly one layout
///   tf32x3_kernel       128 x 128 output tile per CTA, k-tile 32 (one 128-byte
///                       swizzle row).  Warp 0: TMA producer (4 tiles per stage:
///                       A_hi, A_lo, B_hi, B_lo, 3-stage mbarrier ring).  Warp 1:
///                       allocates 256 TMEM columns (two accumulators) and issues
///                       tcgen05.mma kind::tf32 (M 128, N 128, K 8; 12 per stage)
///                       from one elected lane; tcgen05.commit frees the stage /
///                       hands an accumulator over every 128 k.  Warps 2-5:
///                       tcgen05.ld that accumulator (each warp its TMEM lane
///                       quarter), add it to a running fp32 sum with rounded
///                       FADDs (the tensor core's own accumulation truncates),
///                       and finally store C.
///

#include "gemm_tma.cuh"

namespace llo_cuda
{

namespace
{

using namespace tma;

constexpr int TBM = 128;
constexpr int TBN = 128;
constexpr int TBK = 32;
constexpr int TSTAGES = 3;
constexpr int TILE_BYTES = TBM * TBK * 4;                  // 16 KB, each of the 4 tiles
constexpr uint32_t TSTAGE_BYTES = 4 * TILE_BYTES;
constexpr int TTHREADS = 192;
constexpr uint32_t TMEM_COLS = 256;     // two 128-column accumulators
constexpr uint32_t CHUNK_KT = 4;        // k-tiles accumulated in TMEM before promotion
constexpr size_t TSMEM_BYTES = 1024 + (size_t) TSTAGES * TSTAGE_BYTES + 128;

struct Tf32Args
{
	float* C;
	int64_t c_rs;
	uint32_t M;
	uint32_t N;
	uint32_t K;
	uint32_t tiles_m;
	uint32_t tiles_n;
	uint32_t vec_store;
};

/// smem matrix descriptor of a K-major tile with 128-byte swizzle: rows of
/// 128 bytes, 8-row groups 1024 bytes apart (SBO), LBO 1, descriptor
/// version 1 (sm_100), layout type SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc (uint32_t smem_addr)
{
	uint64_t d = 0;
	d |= (uint64_t) ((smem_addr >> 4) & 0x3fffu);
	d |= (uint64_t) 1 << 16;
	d |= (uint64_t) (1024 >> 4) << 32;
	d |= (uint64_t) 1 << 46;
	d |= (uint64_t) 2 << 61;
	return d;
}

/// instruction descriptor: D fp32, A / B tf32, both K-major, M 128, N 128
constexpr uint32_t TF32_IDESC = (1u << 4) | (2u << 7) | (2u << 10) |
	((uint32_t) (TBN >> 3) << 17) | ((uint32_t) (TBM >> 4) << 24);

__device__ __forceinline__ void umma_tf32 (uint32_t tmem_d, uint64_t desc_a,
	uint64_t desc_b, uint32_t accumulate)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
		"}\n"
		:: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(TF32_IDESC),
		"r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit (uint64_t* bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
		:: "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(TTHREADS, 1)
tf32x3_kernel (const __grid_constant__ CUtensorMap map_ahi,
	const __grid_constant__ CUtensorMap map_alo,
	const __grid_constant__ CUtensorMap map_bhi,
	const __grid_constant__ CUtensorMap map_blo,
	const __grid_constant__ Tf32Args args)
{
	extern __shared__ __align__(1024) unsigned char smem_raw[];
	unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
	uint64_t* full = (uint64_t*) (smem + (size_t) TSTAGES * TSTAGE_BYTES);
	uint64_t* empty = full + TSTAGES;
	uint64_t* acc_full = empty + TSTAGES;    // [2] MMA -> epilogue
	uint64_t* acc_empty = acc_full + 2;      // [2] epilogue -> MMA
	uint32_t* tmem_slot = (uint32_t*) (acc_empty + 2);

	constexpr uint32_t GROUP = 16;
	uint32_t bid = blockIdx.x;
	uint32_t per_group = GROUP * args.tiles_n;
	uint32_t first_m = (bid / per_group) * GROUP;
	uint32_t gsize = min(args.tiles_m - first_m, GROUP);
	uint32_t tile_m = first_m + (bid % per_group) % gsize;
	uint32_t tile_n = (bid % per_group) / gsize;
	uint32_t ktiles = (args.K + TBK - 1) / TBK;

	int warp = threadIdx.x >> 5;
	int lane = threadIdx.x & 31;
	if (0 == threadIdx.x)
	{
		for (int s = 0; s < TSTAGES; ++s)
		{
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], 1);
		}
		for (int b = 0; b < 2; ++b)
		{
			mbar_init(&acc_full[b], 1);
			mbar_init(&acc_empty[b], 4);   // one arrival per epilogue warp
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (1 == warp)
	{
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
			:: "r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t tmem_acc = *tmem_slot;

	if (0 == warp)
	{
		// ---- TMA producer ----
		if (0 == lane)
		{
			asm volatile("prefetch.tensormap [%0];" :: "l"(&map_ahi) : "memory");
			asm volatile("prefetch.tensormap [%0];" :: "l"(&map_alo) : "memory");
			asm volatile("prefetch.tensormap [%0];" :: "l"(&map_bhi) : "memory");
			asm volatile("prefetch.tensormap [%0];" :: "l"(&map_blo) : "memory");
			for (uint32_t kt = 0; kt < ktiles; ++kt)
			{
				uint32_t s = kt % TSTAGES;
				mbar_wait(&empty[s], ((kt / TSTAGES) & 1) ^ 1);
				mbar_expect_tx(&full[s], TSTAGE_BYTES);
				unsigned char* stage = smem + (size_t) s * TSTAGE_BYTES;
				int k0 = (int) (kt * TBK);
				tma_load_2d(stage, &map_ahi, &full[s], k0, (int) (tile_m * TBM));
				tma_load_2d(stage + TILE_BYTES, &map_alo, &full[s], k0, (int) (tile_m * TBM));
				tma_load_2d(stage + 2 * TILE_BYTES, &map_bhi, &full[s], k0, (int) (tile_n * TBN));
				tma_load_2d(stage + 3 * TILE_BYTES, &map_blo, &full[s], k0, (int) (tile_n * TBN));
			}
		}
	}
	else if (1 == warp)
	{
		// ---- MMA issuer: one elected lane ----
		// The tensor core adds into its fp32 accumulator with truncation, a
		// one-sided error that grows with the length of the chain; so a TMEM
		// accumulator only ever holds CHUNK_KT k-tiles (128 k).  The epilogue
		// warps add each chunk to a running sum with round-to-nearest FADDs
		// while the next chunk accumulates in the other TMEM buffer.
		if (0 == lane)
		{
			uint32_t nchunks = (ktiles + CHUNK_KT - 1) / CHUNK_KT;
			for (uint32_t c = 0; c < nchunks; ++c)
			{
				uint32_t buf = c & 1;
				mbar_wait(&acc_empty[buf], ((c >> 1) & 1) ^ 1);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				uint32_t tmem_d = tmem_acc + buf * 128;
				uint32_t kt_end = min(ktiles, (c + 1) * CHUNK_KT);
				for (uint32_t kt = c * CHUNK_KT; kt < kt_end; ++kt)
				{
					uint32_t s = kt % TSTAGES;
					mbar_wait(&full[s], (kt / TSTAGES) & 1);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					uint32_t base = smem_u32(smem + (size_t) s * TSTAGE_BYTES);
					uint64_t ahi = umma_desc(base);
					uint64_t alo = umma_desc(base + TILE_BYTES);
					uint64_t bhi = umma_desc(base + 2 * TILE_BYTES);
					uint64_t blo = umma_desc(base + 3 * TILE_BYTES);
#pragma unroll
					for (uint32_t k8 = 0; k8 < TBK / 8; ++k8)
					{
						// 8 tf32 = 32 bytes further along the swizzled row
						uint64_t step = (uint64_t) (k8 * 32 >> 4);
						uint32_t first = (kt == c * CHUNK_KT && 0 == k8) ? 0u : 1u;
						umma_tf32(tmem_d, alo + step, bhi + step, first);
						umma_tf32(tmem_d, ahi + step, blo + step, 1);
						umma_tf32(tmem_d, ahi + step, bhi + step, 1);
					}
					umma_commit(&empty[s]);   // stage free once these MMAs retire
				}
				umma_commit(&acc_full[buf]);
			}
		}
	}
	else
	{
		// ---- epilogue: TMEM -> registers (running sum) -> C.  A warp
		// reaches the TMEM lanes of its quarter (warp id mod 4) only; a
		// thread owns one output row of the tile ----
		uint32_t quarter = (uint32_t) warp & 3u;
		uint32_t row = quarter * 32 + (uint32_t) lane;
		uint32_t m = tile_m * TBM + row;
		float sum[TBN];
#pragma unroll
		for (int j = 0; j < TBN; ++j)
		{
			sum[j] = 0.0f;
		}
		uint32_t nchunks = (ktiles + CHUNK_KT - 1) / CHUNK_KT;
		for (uint32_t c = 0; c < nchunks; ++c)
		{
			uint32_t buf = c & 1;
			mbar_wait(&acc_full[buf], (c >> 1) & 1);
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
			for (uint32_t cc = 0; cc < TBN / 32; ++cc)
			{
				uint32_t v[32];
				uint32_t taddr = tmem_acc + ((quarter * 32) << 16) + buf * 128 + cc * 32;
				asm volatile(
					"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
					"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
					"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
					: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
					"=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
					"=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
					"=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
					"=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
					"=r"(v[30]), "=r"(v[31])
					: "r"(taddr) : "memory");
				asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
				for (int j = 0; j < 32; ++j)
				{
					sum[cc * 32 + j] = __fadd_rn(sum[cc * 32 + j], __uint_as_float(v[j]));
				}
			}
			// this buffer may be overwritten by the chunk after next
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (0 == lane)
			{
				mbar_arrive(&acc_empty[buf]);
			}
		}
		if (m < args.M)
		{
			float* crow = args.C + (int64_t) m * args.c_rs;
			uint32_t n0 = tile_n * TBN;
			if (args.vec_store && n0 + TBN <= args.N)
			{
#pragma unroll
				for (int j = 0; j < TBN; j += 4)
				{
					*(float4*) (crow + n0 + j) = make_float4(sum[j], sum[j + 1],
						sum[j + 2], sum[j + 3]);
				}
			}
			else
			{
#pragma unroll
				for (int j = 0; j < TBN; ++j)
				{
					if (n0 + j < args.N)
					{
						crow[n0 + j] = sum[j];
					}
				}
			}
		}
	}

	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (1 == warp)
	{
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;"
			:: "r"(tmem_acc), "r"(TMEM_COLS) : "memory");
	}
}

__device__ __forceinline__ void split_tf32 (float v, float& hi, float& lo)
{
	uint32_t h, l;
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
	hi = __uint_as_float(h);
	float rest = __fsub_rn(v, hi);
	asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(rest));
	lo = __uint_as_float(l);
}

/// element (r, k) of the operand is src[r * rs + k * cs]; hi / lo are
/// [rows][ld] with k contiguous.  CONTIG_K: cs == 1 (straight pass);
/// otherwise rs == 1 and 32 x 32 tiles are transposed through smem.
template <bool CONTIG_K>
__global__ void __launch_bounds__(256)
tf32_split_kernel (float* __restrict__ hi, float* __restrict__ lo,
	const float* __restrict__ src, uint32_t rows, uint32_t K, int64_t rs,
	int64_t cs, int64_t ld)
{
	if (CONTIG_K)
	{
		uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x;
		uint64_t r = i / K;
		uint32_t k = (uint32_t) (i - r * K);
		if (r < rows)
		{
			float h, l;
			split_tf32(__ldg(src + r * rs + k), h, l);
			hi[r * ld + k] = h;
			lo[r * ld + k] = l;
		}
		return;
	}
	__shared__ float tile[32][33];
	uint32_t tiles_r = (rows + 31) / 32;
	uint32_t tr = blockIdx.x % tiles_r;
	uint32_t tk = blockIdx.x / tiles_r;
	uint32_t tx = threadIdx.x & 31;
	uint32_t ty = threadIdx.x >> 5;   // 0..7
	// read along r (contiguous), 4 k per thread
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		uint32_t k = tk * 32 + ty + 8 * q;
		uint32_t r = tr * 32 + tx;
		tile[ty + 8 * q][tx] = (k < K && r < rows) ? __ldg(src + r + (int64_t) k * cs) : 0.0f;
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < 4; ++q)
	{
		uint32_t r = tr * 32 + ty + 8 * q;
		uint32_t k = tk * 32 + tx;
		if (r < rows && k < K)
		{
			float h, l;
			split_tf32(tile[tx][ty + 8 * q], h, l);
			hi[(int64_t) r * ld + k] = h;
			lo[(int64_t) r * ld + k] = l;
		}
	}
}

/// hi / lo parts of an operand as K-major scratch arrays
int split_operand (llo_cuda_ctx* ctx, Scratch& scratch, const float* src,
	uint64_t rows, uint64_t K, int64_t rs, int64_t cs, float*& hi, float*& lo,
	int64_t& ld)
{
	ld = (int64_t) ((K + 3) / 4 * 4);
	void* ph = nullptr;
	void* pl = nullptr;
	LLO_TRY(scratch.get(&ph, (size_t) rows * ld * sizeof(float)));
	LLO_TRY(scratch.get(&pl, (size_t) rows * ld * sizeof(float)));
	hi = (float*) ph;
	lo = (float*) pl;
	if (1 == cs || 1 == K)
	{
		uint64_t total = rows * K;
		tf32_split_kernel<true><<<(unsigned) ((total + 255) / 256), 256, 0,
			ctx->stream>>>(hi, lo, src, (uint32_t) rows, (uint32_t) K, rs, cs, ld);
	}
	else
	{
		uint64_t blocks = ((rows + 31) / 32) * ((K + 31) / 32);
		tf32_split_kernel<false><<<(unsigned) blocks, 256, 0, ctx->stream>>>(
			hi, lo, src, (uint32_t) rows, (uint32_t) K, rs, cs, ld);
	}
	return launched(ctx, "tf32_split_kernel");
}

}

int tf32x3_launch (llo_cuda_ctx* ctx, const GemmParams& p0, float* C,
	const float* A, const float* B, bool& handled)
{
	handled = false;
	GemmParams p = p0;
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
	if (p.M < 128 || p.N < 128 || p.K < 32 || nullptr == tensor_map_encoder() ||
		p.M >= (1ull << 31) || p.N >= (1ull << 31) || p.K >= (1ull << 31))
	{
		return LLO_OK; // small problems: the exact kernel is as fast
	}
	bool a_ok = (1 == p.a_cs || 1 == p.a_rs) && p.a_rs > 0 && p.a_cs > 0;
	bool b_ok = (1 == p.b_cs || 1 == p.b_rs) && p.b_rs > 0 && p.b_cs > 0;
	if (false == a_ok || false == b_ok)
	{
		return LLO_OK;
	}
	Scratch scratch(ctx);
	float* ahi; float* alo; float* bhi; float* blo;
	int64_t lda, ldb;
	// A(m, k) = A[m * a_rs + k * a_cs]; B as an N x K operand: B(k, n) -> row n
	LLO_TRY(split_operand(ctx, scratch, A, p.M, p.K, p.a_rs, p.a_cs, ahi, alo, lda));
	LLO_TRY(split_operand(ctx, scratch, B, p.N, p.K, p.b_cs, p.b_rs, bhi, blo, ldb));
	CUtensorMap mahi, malo, mbhi, mblo;
	bool ok = make_map(mahi, ahi, p.K, p.M, lda, TBK, TBM, true) &&
		make_map(malo, alo, p.K, p.M, lda, TBK, TBM, true) &&
		make_map(mbhi, bhi, p.K, p.N, ldb, TBK, TBN, true) &&
		make_map(mblo, blo, p.K, p.N, ldb, TBK, TBN, true);
	if (false == ok)
	{
		return LLO_OK;
	}
	Tf32Args args;
	args.C = C;
	args.c_rs = p.c_rs;
	args.M = (uint32_t) p.M;
	args.N = (uint32_t) p.N;
	args.K = (uint32_t) p.K;
	args.tiles_m = (uint32_t) ((p.M + TBM - 1) / TBM);
	args.tiles_n = (uint32_t) ((p.N + TBN - 1) / TBN);
	args.vec_store = (0 == ((uintptr_t) C & 15u) && 0 == p.c_rs % 4) ? 1 : 0;
	uint64_t tiles = (uint64_t) args.tiles_m * args.tiles_n;
	if (tiles >= 0x7fffffffull)
	{
		return LLO_OK;
	}
	LLO_CUDA_TRY(cudaFuncSetAttribute(tf32x3_kernel,
		cudaFuncAttributeMaxDynamicSharedMemorySize, (int) TSMEM_BYTES));
	handled = true;
	tf32x3_kernel<<<(unsigned) tiles, TTHREADS, TSMEM_BYTES, ctx->stream>>>(
		mahi, malo, mbhi, mblo, args);
	return launched(ctx, "tf32x3_kernel");
}

}
