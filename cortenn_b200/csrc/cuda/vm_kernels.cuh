///
/// vm_kernels.cuh — the kernels of the fused elementwise engine (vm.cuh):
/// flat launch, padded-tile launch, cp.async / swizzled tile launch for 4-byte
/// types and the pair kernel (a tensor and its transpose in one program).
///
/// Device code only.  The file is compiled twice:
///   * by nvcc into libllo_cuda.so with the program INTERPRETED (VmParams::code
///     read at run time) -- the general path, any program, no set-up cost;
///   * by NVRTC at run time with LLO_VM_STATIC defined and the program, the
///     input kinds and the rank supplied as constexpr tables (specialize.cu):
///     the same kernels with the interpreter, the operand-kind branches and
///     the accumulator copies compiled away.  Large launches use that build.
///

#ifndef LLO_CUDA_VM_KERNELS_CUH
#define LLO_CUDA_VM_KERNELS_CUH

#include "vm.cuh"

namespace llo_cuda
{

// ---- flat launch: 256 threads x V consecutive outputs per CTA -------------------

template <typename T, bool FULL>
__device__ __forceinline__ void vm_flat_body (const VmParams& p, uint64_t base,
	bool full, T* spill)
{
	using G = VmGeom<T>;
	FlatLoader<T,FULL> loader = {p, base, full};
	T acc[G::V];
	vm_exec<T>(acc, p, loader, spill);
	T* out = static_cast<T*>(p.out) + loader.first(0);
#pragma unroll
	for (int u = 0; u < G::U; ++u)
	{
		store_chunk<T,G::E>(out + u * VM_THREADS * G::E, acc + u * G::E,
			LLO_VM_OUT_VEC(p), loader.valid(u));
	}
}

template <typename T>
__global__ void __launch_bounds__(VM_THREADS, 4)
vm_flat_kernel (const __grid_constant__ VmParams p)
{
	using G = VmGeom<T>;
	extern __shared__ __align__(16) unsigned char vm_smem[];
	uint64_t base = (uint64_t) blockIdx.x * (VM_THREADS * G::V);
	bool full = base + VM_THREADS * G::V <= p.n;
#ifdef LLO_VM_STATIC
	// the specialised program is short: a copy without any tail handling for
	// the CTAs that lie wholly inside the output, the general one for the last
	if (full)
	{
		vm_flat_body<T,true>(p, base, true, reinterpret_cast<T*>(vm_smem));
		return;
	}
#endif
	vm_flat_body<T,false>(p, base, full, reinterpret_cast<T*>(vm_smem));
}

// ---- tile launch: transposed inputs staged through shared memory ---------------

/// Offset contributed by the dims beyond the tile plane (>= 2); `brest` is
/// the block's linear index over those dims
__device__ __forceinline__ int64_t rest_offset (const VmParams& p,
	const int64_t (&stride)[LLO_RANK], uint32_t brest)
{
	int64_t off = 0;
	for (int d = 2; d < LLO_VM_NDIMS(p); ++d)
	{
		uint32_t dim = (uint32_t) p.dims[d];
		uint32_t q = brest / dim;
		off += (int64_t) (brest - q * dim) * stride[d];
		brest = q;
	}
	return off;
}

/// Input access of the tile launch.  Chunk u of thread t is row
/// t / 16 + 16 u of the tile, columns (t % 16) * E .. +E
template <typename T>
struct TileLoader
{
	using G = VmGeom<T>;

	const VmParams& p;
	const T* tiles;          // [slot][TILEP][PITCH]
	uint32_t c0;             // dim-0 coordinate of this thread's chunks
	uint32_t c1;             // staged-dim coordinate of chunk 0
	uint32_t dim1;
	uint32_t brest;
	int valid0;              // elements of a chunk inside dim 0

	__device__ __forceinline__ int valid (int u) const
	{
		return (c1 + 16 * u < dim1) ? valid0 : 0;
	}

	__device__ __forceinline__ void load (T (&x)[G::V], int k,
		const VmKind& kind) const
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE == kind.mode)
		{
			const T* src = tiles + ((size_t) kind.vec_ok * G::TILEP +
				(threadIdx.x >> 4)) * G::PITCH + (threadIdx.x & 15) * G::E;
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
#pragma unroll
				for (int e = 0; e < G::E; ++e)
				{
					x[u * G::E + e] = src[u * 16 * G::PITCH + e];
				}
			}
		}
		else
		{
			int64_t off = in.offset + (int64_t) c0 * in.stride[0] +
				(int64_t) c1 * in.stride[1] + rest_offset(p, in.stride, brest);
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				int ok = valid(u);
				load_strided_chunk<T,G::E>(x + u * G::E, in, kind,
					ok > 0 ? off + (int64_t) (16 * u) * in.stride[1] : in.offset,
					ok);
			}
		}
	}
};

template <typename T>
__global__ void __launch_bounds__(VM_THREADS, 3)
vm_tile_kernel (const __grid_constant__ VmParams p)
{
	using G = VmGeom<T>;
	extern __shared__ __align__(16) unsigned char vm_smem[];
	T* tiles = reinterpret_cast<T*>(vm_smem);
	T* spill = tiles + (size_t) VM_MAX_TILE_INPUTS * G::TILEP * G::PITCH;

	// block -> (tile along dim 0, tile along dim 1, every other dim)
	uint32_t b = blockIdx.x;
	uint32_t t0 = b % p.tiles0;
	b /= p.tiles0;
	uint32_t tp = b % p.tilesp;
	uint32_t brest = b / p.tilesp;
	uint32_t o0 = t0 * G::TILE0;
	uint32_t o1 = tp * G::TILEP;
	uint32_t dim0 = (uint32_t) p.dims[0];
	uint32_t dim1 = (uint32_t) p.dims[1];

	// stage: lanes run along dim 1, the staged inputs' contiguous direction
	for (int k = 0; k < p.ninputs; ++k)
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE != in.mode)
		{
			continue;
		}
		const T* data = static_cast<const T*>(in.data);
		int64_t origin = in.offset + (int64_t) o0 * in.stride[0] +
			(int64_t) o1 * in.stride[1] + rest_offset(p, in.stride, brest);
		T* tile = tiles + (size_t) in.vec_ok * G::TILEP * G::PITCH;
		// a warp covers consecutive elements of the input's contiguous
		// direction; 256 threads sweep the TILE0 x TILEP tile in V steps,
		// all loads issued before any is used
		T v[G::V];
#pragma unroll
		for (int j = 0; j < G::V; ++j)
		{
			int e = threadIdx.x + VM_THREADS * j;
			int lp = e % G::TILEP;
			int r0 = e / G::TILEP;
			v[j] = T();
			if (o1 + lp < dim1 && o0 + r0 < dim0)
			{
				v[j] = __ldcs(data + origin + (int64_t) r0 * in.stride[0] + lp);
			}
		}
#pragma unroll
		for (int j = 0; j < G::V; ++j)
		{
			int e = threadIdx.x + VM_THREADS * j;
			tile[(e % G::TILEP) * G::PITCH + e / G::TILEP] = v[j];
		}
	}
	__syncthreads();

	TileLoader<T> loader = {p, tiles, 0, 0, dim1, brest, 0};
	loader.c0 = o0 + (threadIdx.x & 15) * G::E;
	loader.c1 = o1 + (threadIdx.x >> 4);
	if (loader.c0 < dim0)
	{
		uint32_t left = dim0 - loader.c0;
		loader.valid0 = left < (uint32_t) G::E ? (int) left : G::E;
	}
	T acc[G::V];
	vm_exec<T>(acc, p, loader, spill);
	T* out = static_cast<T*>(p.out) + (int64_t) loader.c0 * p.out_stride[0] +
		(int64_t) loader.c1 * p.out_stride[1] +
		rest_offset(p, p.out_stride, brest);
#pragma unroll
	for (int u = 0; u < G::U; ++u)
	{
		int ok = loader.valid(u);
		if (ok > 0)
		{
			store_chunk<T,G::E>(out + (int64_t) (16 * u) * p.out_stride[1],
				acc + u * G::E, LLO_VM_OUT_VEC(p), ok);
		}
	}
}

// ---- tile launch for 4-byte types: cp.async staging, swizzled tiles ------------
//
// 64 x 64 tile.  The staged (transposed) input tile is copied as it lies in
// memory, 16 bytes per cp.async, into smem [r0][p] (r0: output dim 0, p: the
// input's contiguous direction = output dim 1) with the 16-byte chunk index
// XORed by (r0 >> 2) & 7.
// Compute phase: the 256 threads form a 16 x 16 grid, thread (tx, ty) =
// (t & 15, t >> 4) owns the 4 x 4 outputs c0 = 4 tx .. +3 (dim 0) by
// c1 = 4 ty .. +3 (dim 1); element u * 4 + e of its value set is
// (c0 + e, c1 + u), i.e. chunk u is a 128-bit piece of output row c1 + u (a
// warp covers two 256-byte row segments per chunk: full sectors on the straight
// loads and on the stores).  The transposed operand is four LDS.128 -- row
// 4 tx + e, chunk ty of the staged tile -- whose words are the four c1 of
// element e (a transpose in register names only); the XOR spreads the 8
// lanes of an LDS phase (8 consecutive tx, one ty) over all 8 chunk columns.

constexpr int VM_TS = 64;

__device__ __forceinline__ void cp_async16 (void* smem_dst, const void* src,
	int src_bytes)
{
	uint32_t dst = (uint32_t) __cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
		:: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

/// full 16-byte copy to a shared-window address
__device__ __forceinline__ void cp_async16_full (uint32_t dst, const void* src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
		:: "r"(dst), "l"(src) : "memory");
}

/// 32-bit form of rest_offset (the tile launch requires 32-bit addressing)
__device__ __forceinline__ int32_t rest_offset32 (const VmParams& p,
	const int64_t (&stride)[LLO_RANK], uint32_t brest)
{
	if (LLO_VM_NDIMS(p) <= 3)
	{
		// the usual cases: no extra dim, or one (its coordinate is brest)
		return LLO_VM_NDIMS(p) < 3 ? 0 : (int32_t) brest * (int32_t) stride[2];
	}
	int32_t off = 0;
	for (int d = 2; d < LLO_VM_NDIMS(p); ++d)
	{
		uint32_t dim = (uint32_t) p.dims[d];
		uint32_t q = brest / dim;
		off += (int32_t) (brest - q * dim) * (int32_t) stride[d];
		brest = q;
	}
	return off;
}

/// Thread-constant read positions inside a staged tile (elements)
struct Tile4Thread
{
	uint32_t tx4;        // 4 tx: first dim-0 coordinate inside the tile
	uint32_t ty4;        // 4 ty: first dim-1 coordinate inside the tile
	uint32_t tile_off;   // transposed read, e = 0 (row 4 tx, chunk ty); +step per e
	uint32_t alias_off;  // straight read of the mirror tile, u = 0 (row 4 ty, chunk tx)
	uint32_t step;       // elements between the rows 4 t + e and 4 t + e + 1 of a tile
};

/// Positions for tiles written by TMA (vm_pair4_tma_kernel): the tile is
/// fetched through a 4-D view of the tensor -- (p % 32, r / 4, r % 4, p / 32)
/// -- with the 128-byte hardware swizzle, so that row r of the tile lands in
/// shared-memory row s = (r % 4) * 16 + r / 4 of the half p / 32 and its
/// 16-byte chunks are XORed with s % 8 = (r / 4) % 8: the key the conflict-free
/// 4 x 4 register transposition needs (rows 4 t + e of eight neighbouring
/// lanes get eight different keys), from a swizzle the hardware only offers
/// as a function of the shared-memory row.
__device__ __forceinline__ Tile4Thread tile4_thread_tma ()
{
	uint32_t tx = threadIdx.x & 15;
	uint32_t ty = threadIdx.x >> 4;
	Tile4Thread th;
	th.tx4 = 4 * tx;
	th.ty4 = 4 * ty;
	th.tile_off = (ty >> 3) * 2048 + tx * 32 + (((ty & 7) ^ (tx & 7)) << 2);
	th.alias_off = (tx >> 3) * 2048 + ty * 32 + (((tx & 7) ^ (ty & 7)) << 2);
	th.step = 16 * 32;
	return th;
}

__device__ __forceinline__ Tile4Thread tile4_thread ()
{
	uint32_t tx = threadIdx.x & 15;
	uint32_t ty = threadIdx.x >> 4;
	Tile4Thread th;
	th.tx4 = 4 * tx;
	th.ty4 = 4 * ty;
	th.tile_off = 4 * tx * VM_TS + ((ty ^ (tx & 7)) << 2);
	th.alias_off = 4 * ty * VM_TS + ((tx ^ (ty & 7)) << 2);
	th.step = VM_TS;
	return th;
}

template <typename T, bool FULL>
struct Tile4Loader
{
	using G = VmGeom<T>;

	const VmParams& p;
	const T* tile_rd;        // tiles + tile_off; slot k at + k * 64 * 64
	const T* alias_rd;       // pair mode: mirror tile + alias_off
	uint32_t c0;             // output coordinates of element 0; element u * 4 + e
	uint32_t c1;             // sits at (c0 + e, c1 + u)
	uint32_t dim0;
	uint32_t dim1;
	uint32_t brest;
	bool full;               // FULL = false: the whole tile lies inside the output
	const uint4* pre;        // specialised build: row-broadcast operands fetched
	                         // ahead of the barrier (tile4_prefetch), else null
	uint32_t step;           // Tile4Thread::step

	__device__ __forceinline__ int valid (int u) const
	{
		if (FULL || full)
		{
			return 4;
		}
		if (c1 + u >= dim1 || c0 >= dim0)
		{
			return 0;
		}
		uint32_t left = dim0 - c0;
		return left < 4u ? (int) left : 4;
	}

	__device__ __forceinline__ void load (T (&x)[G::V], int k,
		const VmKind& kind) const
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE == kind.mode)
		{
			const T* src = tile_rd + kind.vec_ok * VM_TS * VM_TS;
#pragma unroll
			for (int e = 0; e < 4; ++e)
			{
				uint4 tmp = *reinterpret_cast<const uint4*>(src + e * step);
				const T* t = reinterpret_cast<const T*>(&tmp);
#pragma unroll
				for (int u = 0; u < 4; ++u)
				{
					x[u * 4 + e] = t[u];
				}
			}
			return;
		}
		if (VM_IN_ALIAS == kind.mode)
		{
			// element (c0, c1) of the straight view = [row c1][col c0] of the
			// mirror tile
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				uint4 tmp = *reinterpret_cast<const uint4*>(alias_rd + u * step);
				const T* t = reinterpret_cast<const T*>(&tmp);
#pragma unroll
				for (int e = 0; e < 4; ++e)
				{
					x[u * 4 + e] = t[e];
				}
			}
			return;
		}
		const T* data = static_cast<const T*>(in.data);
		int32_t s0 = (int32_t) in.stride[0];
		int32_t s1 = (int32_t) in.stride[1];
		int32_t off = (int32_t) in.offset + rest_offset32(p, in.stride, brest) +
			(int32_t) c0 * s0 + (int32_t) c1 * s1;
		if (1 == kind.s0)
		{
			if ((FULL || full) && kind.vec_ok)
			{
				if (0 == kind.s1)
				{
					// broadcast along dim 1: the four chunks are the same 16 bytes
					uint4 tmp = (nullptr != pre) ? pre[k] :
						__ldg(reinterpret_cast<const uint4*>(data + off));
					const T* t = reinterpret_cast<const T*>(&tmp);
#pragma unroll
					for (int u = 0; u < 4; ++u)
					{
#pragma unroll
						for (int e = 0; e < 4; ++e)
						{
							x[u * 4 + e] = t[e];
						}
					}
					return;
				}
#pragma unroll
				for (int u = 0; u < 4; ++u)
				{
					uint4 tmp = __ldg(reinterpret_cast<const uint4*>(
						data + off + u * s1));
					const T* t = reinterpret_cast<const T*>(&tmp);
#pragma unroll
					for (int e = 0; e < 4; ++e)
					{
						x[u * 4 + e] = t[e];
					}
				}
				return;
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				load_chunk<T,4,false>(x + u * 4, data + off + u * s1,
					kind.vec_ok, valid(u));
			}
		}
		else if (0 == kind.s0)
		{
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				T v = valid(u) > 0 ? __ldg(data + off + u * s1) : T();
#pragma unroll
				for (int e = 0; e < 4; ++e)
				{
					x[u * 4 + e] = v;
				}
			}
		}
		else
		{
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				int ok = valid(u);
				const T* src = data + off + u * s1;
#pragma unroll
				for (int e = 0; e < 4; ++e)
				{
					x[u * 4 + e] = (e < ok) ? __ldg(src + e * s0) : T();
				}
			}
		}
	}
};

/// Tile number -> origin of the tile (o0, o1) and index over the other dims
__device__ __forceinline__ void tile_origin (const VmParams& p, uint32_t tile,
	uint32_t& o0, uint32_t& o1, uint32_t& brest)
{
	uint32_t per = p.tiles0 * p.tilesp;
	brest = tile / per;
	uint32_t b = tile - brest * per;
	uint32_t t0, tp;
	if (p.mirror)
	{
		// tiles (i,j) and (j,i) get neighbouring numbers, so an input that
		// is read both straight and transposed (x and permute(x)) comes from
		// DRAM once and from L2 the second time
		uint32_t side = p.tiles0;
		uint32_t off_diag = side * (side - 1);
		if (b < off_diag)
		{
			uint32_t q = b >> 1; // pair (i < j), q = j (j - 1) / 2 + i
			uint32_t j = (uint32_t) ((1.0f + sqrtf(1.0f + 8.0f * (float) q)) * 0.5f);
			while (j * (j - 1) / 2 > q) { --j; }
			while ((j + 1) * j / 2 <= q) { ++j; }
			uint32_t i = q - j * (j - 1) / 2;
			t0 = (b & 1) ? j : i;
			tp = (b & 1) ? i : j;
		}
		else
		{
			t0 = tp = b - off_diag;
		}
	}
	else
	{
		t0 = b % p.tiles0;
		tp = b / p.tiles0;
	}
	o0 = t0 * VM_TS;
	o1 = tp * VM_TS;
}

/// Stage every transposed input's tile at (o0, o1) into `tiles`
template <typename T, bool ALIGNED>
__device__ __forceinline__ void stage_tiles (const VmParams& p, T* tiles,
	uint32_t o0, uint32_t o1, uint32_t brest)
{
	uint32_t dim0 = (uint32_t) p.dims[0];
	uint32_t dim1 = (uint32_t) p.dims[1];
	bool full = o0 + VM_TS <= dim0 && o1 + VM_TS <= dim1;
	for (int k = 0; k < p.ninputs; ++k)
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE != in.mode)
		{
			continue;
		}
		const T* data = static_cast<const T*>(in.data);
		int32_t s0 = (int32_t) in.stride[0];
		int32_t origin = (int32_t) in.offset + rest_offset32(p, in.stride, brest) +
			(int32_t) o0 * s0 + (int32_t) o1 * (int32_t) in.stride[1];
		T* tile = tiles + in.vec_ok * VM_TS * VM_TS;
		if (ALIGNED)
		{
			// chunk ci = t + 256 j: row r0 = t / 16 + 16 j, 16-byte column t % 16;
			// 16 rows further the swizzle key (r0 >> 2) & 7 moves by 4
			int r0 = threadIdx.x >> 4;
			int pc = threadIdx.x & 15;
			const T* src = data + origin + r0 * s0 + pc * 4;
			if (full)
			{
#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					cp_async16(tile + (r0 + 16 * j) * VM_TS +
						((pc ^ (((r0 >> 2) + 4 * j) & 7)) << 2),
						src + 16 * j * s0, 16);
				}
			}
			else
			{
				int left = (int) dim1 - (int) (o1 + pc * 4);
				int row_bytes = left > 0 ? (left < 4 ? left * 4 : 16) : 0;
#pragma unroll
				for (int j = 0; j < 4; ++j)
				{
					int bytes = (o0 + r0 + 16 * j < dim0) ? row_bytes : 0;
					cp_async16(tile + (r0 + 16 * j) * VM_TS +
						((pc ^ (((r0 >> 2) + 4 * j) & 7)) << 2),
						bytes > 0 ? src + 16 * j * s0 : data, bytes);
				}
			}
		}
		else
		{
			T v[16];
#pragma unroll
			for (int j = 0; j < 16; ++j)
			{
				int e = threadIdx.x + VM_THREADS * j;
				int lp = e & 63;
				int r0 = e >> 6;
				v[j] = T();
				if (o1 + lp < dim1 && o0 + r0 < dim0)
				{
					v[j] = __ldcs(data + origin + r0 * s0 + lp);
				}
			}
#pragma unroll
			for (int j = 0; j < 16; ++j)
			{
				int e = threadIdx.x + VM_THREADS * j;
				int lp = e & 63;
				int r0 = e >> 6;
				tile[r0 * VM_TS + (((lp >> 2) ^ ((r0 >> 2) & 7)) << 2) + (lp & 3)] = v[j];
			}
		}
	}
}

#ifdef LLO_VM_STATIC
/// Operands that are contiguous along dim 0 and broadcast along dim 1
/// (extend(b): 16 bytes per thread and tile) do not depend on the staged
/// tiles, so the specialised build fetches them BEFORE the barrier the tile
/// waits at instead of stalling on them right behind it (ncu on the chain:
/// half of all stall samples sat on the multiply that consumes b).
struct Tile4Pre
{
	uint4 v[LLO_VM_S_NINPUTS];
};

template <int K, typename T>
__device__ __forceinline__ void tile4_prefetch (Tile4Pre& pre, const VmParams& p,
	uint32_t c0, uint32_t brest)
{
	if constexpr (K < LLO_VM_S_NINPUTS)
	{
		if constexpr (VM_IN_STRIDED == LLO_VM_S_MODE[K] && 1 == LLO_VM_S_S0[K] &&
			0 == LLO_VM_S_S1[K] && 0 != LLO_VM_S_VEC[K])
		{
			const VmInput& in = p.in[K];
			int32_t off = (int32_t) in.offset + rest_offset32(p, in.stride, brest) +
				(int32_t) c0;
			pre.v[K] = __ldg(reinterpret_cast<const uint4*>(
				static_cast<const T*>(in.data) + off));
		}
		tile4_prefetch<K + 1,T>(pre, p, c0, brest);
	}
}
#endif

/// Run the program for the output tile at (o0, o1) and store it
template <typename T, bool FULL>
__device__ __forceinline__ void run_tile (const VmParams& p,
	const Tile4Thread& th, const T* tiles, T* spill,
	uint32_t o0, uint32_t o1, uint32_t brest, const T* alias = nullptr,
	const uint4* pre = nullptr)
{
	using G = VmGeom<T>;
	Tile4Loader<T,FULL> loader = {p, tiles + th.tile_off,
		alias + th.alias_off, o0 + th.tx4, o1 + th.ty4,
		(uint32_t) p.dims[0], (uint32_t) p.dims[1], brest,
		FULL || (o0 + VM_TS <= (uint32_t) p.dims[0] &&
			o1 + VM_TS <= (uint32_t) p.dims[1]), pre, th.step};
	T acc[G::V];
	// (the compact interpreter for the pair kernel, FULL = true, only:
	// vm_tile4_kernel lost 2% to it on a plain permute)
	vm_exec<T,Tile4Loader<T,FULL>,FULL>(acc, p, loader, spill);
	int32_t os1 = (int32_t) p.out_stride[1];
	T* out = static_cast<T*>(p.out) + rest_offset32(p, p.out_stride, brest) +
		(int32_t) loader.c0 + (int32_t) loader.c1 * os1;
	if (loader.full && LLO_VM_OUT_VEC(p))
	{
#pragma unroll
		for (int u = 0; u < 4; ++u)
		{
			uint4 tmp;
			T* t = reinterpret_cast<T*>(&tmp);
#pragma unroll
			for (int e = 0; e < 4; ++e)
			{
				t[e] = acc[u * 4 + e];
			}
			__stcs(reinterpret_cast<uint4*>(out + u * os1), tmp);
		}
		return;
	}
#pragma unroll
	for (int u = 0; u < 4; ++u)
	{
		int ok = loader.valid(u);
		if (ok > 0)
		{
			store_chunk<T,4>(out + u * os1, acc + u * 4, LLO_VM_OUT_VEC(p), ok);
		}
	}
}

/// Work-unit hand-out of the persistent tile kernels.  A CTA starts with unit
/// blockIdx.x; further units come from a global counter (static c + grid
/// strides left the SMs that sit further from L2 up to 11% behind the others:
/// profiles/r1_permute_details.txt).  Thread 0 claims one unit ahead -- the
/// atomic's latency hides under the current unit's work -- and publishes it
/// through shared memory at the barrier that already ends every iteration.
struct UnitClaim
{
	uint32_t* sched;
	uint32_t* slot;     // shared word
	uint32_t claimed;   // thread 0: the unit claimed for the iteration after next

	__device__ __forceinline__ uint32_t claim (void) const
	{
		return gridDim.x + atomicAdd(sched, 1u);
	}

	/// the last CTA out resets the counter for the next launch
	__device__ __forceinline__ void finish (void) const
	{
		if (0 == threadIdx.x)
		{
			__threadfence();
			if (gridDim.x - 1 == atomicAdd(sched + 1, 1u))
			{
				sched[0] = 0;
				sched[1] = 0;
				__threadfence();
			}
		}
	}
};

/// Persistent: CTA c handles tiles c, then whatever the counter hands out.  With ALIGNED staging the
/// cp.async copies of the next tile fly while the current one is computed
/// (two tile buffers); the scalar fallback stages and computes in turn.
template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(VM_THREADS, 3)
vm_tile4_kernel (const __grid_constant__ VmParams p)
{
	static_assert(4 == sizeof(T), "4-byte element types only");
	extern __shared__ __align__(16) unsigned char vm_smem[];
	T* tiles = reinterpret_cast<T*>(vm_smem);
	const int buf_elems = p.nstaged * VM_TS * VM_TS;
	T* spill = tiles + (ALIGNED ? 2 : 1) * buf_elems;
	const Tile4Thread th = tile4_thread();
	uint32_t o0, o1, brest;
	if (ALIGNED)
	{
		__shared__ uint32_t next_slot;
		UnitClaim units = {p.sched, &next_slot, 0};
		uint32_t tile = blockIdx.x;
		if (0 == threadIdx.x)
		{
			next_slot = units.claim();
		}
		tile_origin(p, tile, o0, o1, brest);
		stage_tiles<T,true>(p, tiles, o0, o1, brest);
		asm volatile("cp.async.commit_group;" ::: "memory");
		__syncthreads();
		uint32_t next = next_slot;
		int buf = 0;
		while (tile < p.total_tiles)
		{
			if (0 == threadIdx.x)
			{
				units.claimed = units.claim();
			}
			uint32_t n0 = 0, n1 = 0, nrest = 0;
			if (next < p.total_tiles)
			{
				tile_origin(p, next, n0, n1, nrest);
				stage_tiles<T,true>(p, tiles + (buf ^ 1) * buf_elems, n0, n1, nrest);
				asm volatile("cp.async.commit_group;" ::: "memory");
				asm volatile("cp.async.wait_group 1;" ::: "memory");
			}
			else
			{
				asm volatile("cp.async.wait_group 0;" ::: "memory");
			}
			__syncthreads();
			run_tile<T,false>(p, th, tiles + buf * buf_elems, spill, o0, o1, brest);
			if (0 == threadIdx.x)
			{
				next_slot = units.claimed;
			}
			__syncthreads(); // everyone is done with this buffer
			buf ^= 1;
			tile = next;
			next = next_slot;
			o0 = n0;
			o1 = n1;
			brest = nrest;
		}
		units.finish();
	}
	else
	{
		for (uint32_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x)
		{
			tile_origin(p, tile, o0, o1, brest);
			stage_tiles<T,false>(p, tiles, o0, o1, brest);
			__syncthreads();
			run_tile<T,false>(p, th, tiles, spill, o0, o1, brest);
			__syncthreads();
		}
	}
}

/// q-th strictly-lower index pair of the triangular numbering: q = b (b - 1) / 2 + a, a < b
__device__ __forceinline__ void tri_index (uint32_t q, uint32_t& a, uint32_t& b)
{
	uint32_t bb = (uint32_t) ((1.0f + sqrtf(1.0f + 8.0f * (float) q)) * 0.5f);
	while (bb * (bb - 1) / 2 > q) { --bb; }
	while ((bb + 1) * bb / 2 <= q) { ++bb; }
	b = bb;
	a = q - bb * (bb - 1) / 2;
}

/// Work unit q of one slice of the pair kernels -> (i, j): i < j names the tile
/// pair {(i,j), (j,i)}; i > j the diagonal tiles (j,j) and (i,i) = (j+1,j+1);
/// i == j one diagonal tile.  Consecutive units share j, so the mirror tiles
/// (j, i), (j, i + 1) of units that run at the same time are neighbours along
/// the contiguous direction.  (Walking the triangle in 8 x 8 blocks of tiles,
/// which gives both tiles of a pair such neighbours, measured no better:
/// 0.356 vs 0.351 ms on the chain, round 2.)
__device__ __forceinline__ void pair_unit (uint32_t q, uint32_t side,
	uint32_t pairs, uint32_t& i, uint32_t& j)
{
	if (q < pairs)
	{
		tri_index(q, i, j);
	}
	else
	{
		j = 2 * (q - pairs);
		i = j + 1 < side ? j + 1 : j;
	}
}

/// Pair mode (x and permute(x) in one program, square grid of full tiles): a
/// CTA takes the tile pair {(i,j), (j,i)}, stages the two input tiles ONCE
/// (cp.async, two slots, double buffered over pairs) and computes both output
/// tiles: for output (i,j) the transposed operand reads slot 0 transposed and
/// the straight operand reads slot 1 as it lies; for output (j,i) the slots
/// swap roles.  Every element of x then crosses DRAM -> L2 -> SM once.
/// Staging addresses are affine in the tile indices: everything that depends
/// on the thread only (source offset inside a tile, swizzled destination) is
/// computed once per kernel.
#ifdef LLO_VM_STATIC
#define LLO_VM_PAIR_RING LLO_VM_PAIR_NBUF
#define LLO_VM_PAIR_CTAS LLO_VM_PAIR_MINCTAS
#else
#define LLO_VM_PAIR_RING 2
#define LLO_VM_PAIR_CTAS 3
#endif
template <typename T>
__global__ void __launch_bounds__(VM_THREADS, LLO_VM_PAIR_CTAS)
vm_pair4_kernel (const __grid_constant__ VmParams p)
{
	static_assert(4 == sizeof(T), "4-byte element types only");
	extern __shared__ __align__(16) unsigned char vm_smem[];
	T* tiles = reinterpret_cast<T*>(vm_smem);
	constexpr int SLOT = VM_TS * VM_TS;
	T* spill = tiles + 2 * LLO_VM_PAIR_RING * SLOT;
	const uint32_t side = p.tiles0;
	const uint32_t pairs = side * (side - 1) / 2;
	// work units per slice of the other dims: the off-diagonal pairs, then the
	// diagonal tiles two at a time (every unit stages and computes two tiles)
	const uint32_t per = pairs + (side + 1) / 2;
	const uint32_t total = p.total_tiles;    // per * (product of the other dims)
	const Tile4Thread th = tile4_thread();

	// staging: thread t copies 16-byte column t % 16 of rows t / 16 + 16 j
	const VmInput& tin = p.in[p.tile_input];
	const int32_t s0 = (int32_t) tin.stride[0];
	const uint32_t r0 = threadIdx.x >> 4;
	const uint32_t pc = threadIdx.x & 15;
	const T* const tsrc = static_cast<const T*>(tin.data) + tin.offset +
		(int32_t) r0 * s0 + (int32_t) pc * 4;
	const int64_t rstep = (int64_t) 16 * s0;
	const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(tiles);
	// rows r0 + 16 j: the swizzle key (r0 >> 2) + 4 j flips bit 2 for odd j
	const uint32_t dst_even = r0 * (VM_TS * 4) + ((pc ^ (r0 >> 2)) << 4);
	const uint32_t dst_odd = dst_even ^ 64u;

	// unit -> (i, j, brest): i < j names the pair {(i,j), (j,i)}; i > j the
	// diagonal tiles (j,j) and (i,i) = (j+1,j+1); i == j one diagonal tile
	auto unit_of = [&] (uint32_t unit, uint32_t& i, uint32_t& j, uint32_t& brest)
	{
		brest = unit / per;
		pair_unit(unit - brest * per, side, pairs, i, j);
	};
	// input tile of output tile (ti, tj) -> the slot at shared address `slot`
	auto stage_one = [&] (uint32_t slot, uint32_t ti, uint32_t tj, int32_t rest)
	{
		const T* src = tsrc + (rest + (int32_t) (ti * VM_TS) * s0 +
			(int32_t) (tj * VM_TS));
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			cp_async16_full(slot + ((j & 1) ? dst_odd : dst_even) +
				j * (16 * VM_TS * 4), src + j * rstep);
		}
	};
	// pair: slot 0 <- input tile of output (i,j), slot 1 <- that of (j,i);
	// diagonal: slot 0 <- tile (j,j), slot 1 <- tile (i,i)
	auto stage_pair = [&] (int buf, uint32_t i, uint32_t j, uint32_t brest)
	{
		int32_t rest = rest_offset32(p, tin.stride, brest);
		uint32_t slot = smem0 + (uint32_t) buf * (2 * SLOT * 4);
		if (i < j)
		{
			stage_one(slot, i, j, rest);
			stage_one(slot + SLOT * 4, j, i, rest);
		}
		else
		{
			stage_one(slot, j, j, rest);
			if (i != j)
			{
				stage_one(slot + SLOT * 4, i, i, rest);
			}
		}
	};

#ifdef LLO_VM_STATIC
	// ---- specialised build ------------------------------------------------------
	// No longer instruction-bound, this build is tuned for bytes in flight
	// (tools/probes/tile_probe2.cu: a persistent tile loop needs >= 64 KB of
	// loads outstanding per SM and counter-based hand-out to reach the rate of
	// a flat copy: 6.1 -> 6.9 TB/s):
	//   * a ring of NBUF unit buffers (two 16 KB tiles each): the copies of
	//     units k + 1 .. k + NBUF - 1 are in flight while unit k is computed;
	//   * units come from the global counter, claimed NBUF iterations ahead by
	//     thread 0 and published through the ring `uq` (slot m & 3 holds the
	//     unit of iteration m; it is written between barriers m - NBUF and
	//     m - NBUF + 1 and last read behind barrier m - 1... m: never both
	//     between the same two barriers);
	//   * row-broadcast operands of a unit are fetched ahead of its barrier.
	constexpr int NBUF = LLO_VM_PAIR_NBUF;
	static_assert(NBUF >= 2 && NBUF <= 4, "ring depth");
	__shared__ uint32_t uq[4];
	UnitClaim units = {p.sched, uq, 0};
	if (0 == threadIdx.x)
	{
		uq[0] = blockIdx.x;
		for (int m = 1; m < NBUF; ++m)
		{
			uq[m] = units.claim();
		}
	}
	__syncthreads();
	auto stage_unit = [&] (uint32_t unit, int buf)
	{
		if (unit < total)
		{
			uint32_t si, sj, srest;
			unit_of(unit, si, sj, srest);
			stage_pair(buf, si, sj, srest);
		}
		asm volatile("cp.async.commit_group;" ::: "memory"); // (possibly empty: keeps the group count in step)
	};
	for (int m = 0; m < NBUF - 1; ++m)
	{
		stage_unit(uq[m], m);
	}
	uint32_t unit = blockIdx.x;
	for (uint32_t k = 0; unit < total; ++k)
	{
		uint32_t i, j, brest;
		unit_of(unit, i, j, brest);
		Tile4Pre pre0, pre1;
		tile4_prefetch<0,T>(pre0, p, ((i < j) ? i : j) * VM_TS + th.tx4, brest);
		// this unit's copies have landed once at most NBUF - 2 younger groups
		// are still pending
		if (NBUF == 2) { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
		if (NBUF == 3) { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
		if (NBUF == 4) { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
		__syncthreads();
		// behind barrier k: buffer (k - 1) % NBUF is free again -> unit k + NBUF - 1
		uint32_t next = uq[(k + 1) & 3];
		stage_unit(uq[(k + NBUF - 1) & 3], (int) ((k + NBUF - 1) % NBUF));
		if (0 == threadIdx.x)
		{
			units.claimed = units.claim();
		}
		int buf = (int) (k % NBUF);
		T* slot0 = tiles + buf * 2 * SLOT;
		T* slot1 = slot0 + SLOT;
		int halves = (i == j) ? 1 : 2;
		if (2 == halves)
		{
			tile4_prefetch<0,T>(pre1, p, ((i < j) ? j : i) * VM_TS + th.tx4, brest);
		}
		{
			const T* other = (i < j) ? slot1 : slot0;
			uint32_t t0 = (i < j) ? i : j;
			uint32_t t1 = (i < j) ? j : t0;
			run_tile<T,true>(p, th, slot0, spill, t0 * VM_TS, t1 * VM_TS, brest, other, pre0.v);
		}
		if (2 == halves)
		{
			const T* other = (i < j) ? slot0 : slot1;
			uint32_t t0 = (i < j) ? j : i;
			uint32_t t1 = (i < j) ? i : t0;
			run_tile<T,true>(p, th, slot1, spill, t0 * VM_TS, t1 * VM_TS, brest, other, pre1.v);
		}
		if (0 == threadIdx.x)
		{
			uq[(k + NBUF) & 3] = units.claimed;
		}
		unit = next;
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
	units.finish();
#else
	// ---- interpreted build --------------------------------------------------------
	// (units go round-robin here: handing them out through the counter, which
	// helps vm_tile4_kernel, measured 4% slower on this instruction-bound kernel)
	uint32_t unit = blockIdx.x;    // (the grid never exceeds the unit count)
	uint32_t i, j, brest;
	unit_of(unit, i, j, brest);
	stage_pair(0, i, j, brest);
	asm volatile("cp.async.commit_group;" ::: "memory");
	int buf = 0;
	while (unit < total)
	{
		// One barrier per unit: once every thread's copies of this unit have
		// landed and every thread is past it, the unit's tiles are visible to
		// all AND everybody has finished computing from the other buffer, so
		// the next unit may be staged into it right away (its copies fly under
		// this unit's math).
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
		uint32_t next = unit + gridDim.x;
		uint32_t ni = 0, nj = 0, nrest = 0;
		if (next < total)
		{
			unit_of(next, ni, nj, nrest);
			stage_pair(buf ^ 1, ni, nj, nrest);
			asm volatile("cp.async.commit_group;" ::: "memory");
		}
		T* slot0 = tiles + buf * 2 * SLOT;
		T* slot1 = slot0 + SLOT;
		int halves = (i == j) ? 1 : 2;
#pragma unroll 1
		for (int h = 0; h < halves; ++h)
		{
			// pair (i < j): output (i,j) reads slot 0 transposed and slot 1
			// straight, output (j,i) the other way round; diagonal tiles read
			// their own slot both ways
			const T* self = h ? slot1 : slot0;
			const T* other = (i < j) ? (h ? slot0 : slot1) : self;
			uint32_t t0 = (i < j) ? (h ? j : i) : (h ? i : j);
			uint32_t t1 = (i < j) ? (h ? i : j) : t0;
			run_tile<T,true>(p, th, self, spill, t0 * VM_TS, t1 * VM_TS, brest, other);
		}
		buf ^= 1;
		unit = next;
		i = ni;
		j = nj;
		brest = nrest;
	}
#endif
}

#ifdef LLO_VM_STATIC

// ---- pair kernel, TMA-staged and warp-specialised (specialised build only) ------
//
// Same work units and the same math as vm_pair4_kernel; what changes is how the
// tiles arrive and who waits for what:
//   * one elected lane of a ninth warp is the producer: it claims units from the
//     global counter, and per unit issues TWO bulk tensor copies
//     (cp.async.bulk.tensor.4d/5d -> SASS UTMALDG) instead of 256 threads x 8
//     LDGSTS with their address arithmetic;
//   * tiles land in the layout of tile4_thread_tma (4-D view + 128-byte swizzle);
//   * full / empty mbarriers per unit buffer replace the CTA-wide barrier: a
//     consumer warp that is done with a unit moves on to the next one while its
//     neighbours are still storing, and the producer refills a buffer as soon as
//     the eighth warp has released it.

struct alignas(64) VmTensorMap
{
	unsigned long long opaque[16];   // CUtensorMap
};

__device__ __forceinline__ uint32_t vm_smem_u32 (const void* ptr)
{
	return (uint32_t) __cvta_generic_to_shared(ptr);
}

__device__ __forceinline__ void vm_mbar_init (uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
		:: "r"(vm_smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void vm_mbar_expect_tx (uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
		:: "r"(vm_smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void vm_mbar_arrive (uint64_t* bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
		:: "r"(vm_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void vm_mbar_wait (uint64_t* bar, uint32_t parity)
{
	asm volatile(
		"{\n"
		".reg .pred P1;\n"
		"VM_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
		"@P1 bra VM_DONE;\n"
		"bra VM_WAIT;\n"
		"VM_DONE:\n"
		"}" :: "r"(vm_smem_u32(bar)), "r"(parity) : "memory");
}

/// one 64 x 64 tile: box (32, 16, 4, 2[, 1]) of the (p % 32, r / 4, r % 4, p / 32[, rest]) view
template <int RANK>
__device__ __forceinline__ void vm_tma_tile (void* dst, const VmTensorMap* map,
	uint64_t* bar, uint32_t row_tile, uint32_t col_tile, uint32_t rest)
{
	if constexpr (4 == RANK)
	{
		asm volatile(
			"cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
			" [%0], [%1, {%3, %4, %5, %6}], [%2];"
			:: "r"(vm_smem_u32(dst)), "l"(map), "r"(vm_smem_u32(bar)),
			"r"(0), "r"((int) (row_tile * 16)), "r"(0), "r"((int) (col_tile * 2))
			: "memory");
	}
	else
	{
		asm volatile(
			"cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
			" [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
			:: "r"(vm_smem_u32(dst)), "l"(map), "r"(vm_smem_u32(bar)),
			"r"(0), "r"((int) (row_tile * 16)), "r"(0), "r"((int) (col_tile * 2)),
			"r"((int) rest)
			: "memory");
	}
}

template <typename T>
__global__ void __launch_bounds__(VM_THREADS + 32, LLO_VM_PAIR_MINCTAS)
vm_pair4_tma_kernel (const __grid_constant__ VmParams p,
	const __grid_constant__ VmTensorMap tmap)
{
	static_assert(4 == sizeof(T), "4-byte element types only");
	constexpr int NBUF = LLO_VM_PAIR_NBUF;
	constexpr int SLOT = VM_TS * VM_TS;
	constexpr int MAP_RANK = (LLO_VM_STATIC_NDIMS == 3) ? 5 : 4;
	extern __shared__ __align__(16) unsigned char vm_smem[];
	// swizzled TMA destinations start on 1024-byte boundaries
	unsigned char* base = vm_smem + ((1024u - (vm_smem_u32(vm_smem) & 1023u)) & 1023u);
	T* tiles = reinterpret_cast<T*>(base);
	T* spill = tiles + 2 * NBUF * SLOT;
	__shared__ uint64_t full[NBUF];
	__shared__ uint64_t empty[NBUF];
	__shared__ uint32_t uq[NBUF];

	const uint32_t side = p.tiles0;
	const uint32_t pairs = side * (side - 1) / 2;
	const uint32_t per = pairs + (side + 1) / 2;
	const uint32_t total = p.total_tiles;
	auto unit_of = [&] (uint32_t unit, uint32_t& i, uint32_t& j, uint32_t& brest)
	{
		brest = unit / per;
		pair_unit(unit - brest * per, side, pairs, i, j);
	};

	if (0 == threadIdx.x)
	{
		for (int b = 0; b < NBUF; ++b)
		{
			vm_mbar_init(&full[b], 1);
			vm_mbar_init(&empty[b], VM_THREADS / 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("prefetch.tensormap [%0];" :: "l"(&tmap) : "memory");
	}
	__syncthreads();

	if (threadIdx.x >= VM_THREADS)
	{
		// ---- producer warp: one elected lane ----
		if (VM_THREADS == threadIdx.x)
		{
			UnitClaim units = {p.sched, uq, 0};
			uint32_t unit = blockIdx.x;
			uint32_t ahead = units.claim();   // claimed one unit early: the atomic's
			                                  // round trip hides under a buffer wait
			for (uint32_t k = 0; ; ++k)
			{
				uint32_t b = k % NBUF;
				if (k >= NBUF)
				{
					vm_mbar_wait(&empty[b], ((k / NBUF) & 1) ^ 1);
				}
				uq[b] = unit;
				if (unit >= total)
				{
					vm_mbar_arrive(&full[b]);   // the consumers' stop sign
					break;
				}
				uint32_t i, j, brest;
				unit_of(unit, i, j, brest);
				T* slot0 = tiles + b * 2 * SLOT;
				// pair: slot 0 <- input tile of output (i,j), slot 1 <- that of (j,i);
				// diagonal: slot 0 <- tile (j,j), slot 1 <- tile (i,i)
				uint32_t a0 = (i < j) ? i : j, a1 = j;
				uint32_t b0 = (i < j) ? j : i, b1 = i;
				bool two = i != j;
				vm_mbar_expect_tx(&full[b], (two ? 2u : 1u) * SLOT * 4u);
				vm_tma_tile<MAP_RANK>(slot0, &tmap, &full[b], a0, a1, brest);
				if (two)
				{
					vm_tma_tile<MAP_RANK>(slot0 + SLOT, &tmap, &full[b], b0, b1, brest);
				}
				unit = ahead;
				ahead = units.claim();
			}
			// the last CTA out re-arms the counter for the next launch
			__threadfence();
			if (gridDim.x - 1 == atomicAdd(p.sched + 1, 1u))
			{
				p.sched[0] = 0;
				p.sched[1] = 0;
				__threadfence();
			}
		}
		return;
	}

	// ---- consumer warps ----
	const Tile4Thread th = tile4_thread_tma();
	const uint32_t lane = threadIdx.x & 31;
	for (uint32_t k = 0; ; ++k)
	{
		uint32_t b = k % NBUF;
		vm_mbar_wait(&full[b], (k / NBUF) & 1);
		uint32_t unit = uq[b];
		if (unit >= total)
		{
			break;
		}
		uint32_t i, j, brest;
		unit_of(unit, i, j, brest);
		T* slot0 = tiles + b * 2 * SLOT;
		T* slot1 = slot0 + SLOT;
		{
			const T* other = (i < j) ? slot1 : slot0;
			uint32_t t0 = (i < j) ? i : j;
			uint32_t t1 = (i < j) ? j : t0;
			run_tile<T,true>(p, th, slot0, spill, t0 * VM_TS, t1 * VM_TS, brest, other);
		}
		if (i != j)
		{
			const T* other = (i < j) ? slot0 : slot1;
			uint32_t t0 = (i < j) ? j : i;
			uint32_t t1 = (i < j) ? i : t0;
			run_tile<T,true>(p, th, slot1, spill, t0 * VM_TS, t1 * VM_TS, brest, other);
		}
		__syncwarp();
		if (0 == lane)
		{
			vm_mbar_arrive(&empty[b]);
		}
	}
}

#endif // LLO_VM_STATIC

}

#endif // LLO_CUDA_VM_KERNELS_CUH
