///
/// reduce.cu — K4: SUM / PROD / MIN / MAX reductions.
///
/// In the reference a reduction is the push branch of llo::nnary
/// (llo/operator.hpp:376-393) driven by a reduce map (llo/src/helper.cpp:57-65):
/// every source element is added into out[index(M * coord)] one by one.
/// Here the push map is inverted on the host (mapview.hpp) and each output
/// folds its own source elements:
///   row kernel     folded elements contiguous: 128-bit loads, in-thread
///                  partial accumulators, warp-shuffle tree, shared-memory
///                  tree across warps; long rows are split over CTAs
///   column kernel  kept elements contiguous (the layout reduce_*(x, k)
///                  produces): coalesced across columns, 8 independent
///                  accumulators per thread, rows split into slabs
///   small-K kernel few kept elements (bias gradients): rows x K slab read
///                  as one contiguous stream
///   seq kernel     one thread per output, reference order: the bit-parity
///                  mode and the fallback for irregular layouts
/// Split kernels write partials that a second launch of the same kernel
/// folds, so the summation tree is a pure function of the shapes: results
/// are reproducible run to run and GPU to GPU.
///

#include <limits>

#include "reduce.cuh"
#include "ops.cuh"

namespace llo_cuda
{

template <int OP, typename T>
__device__ __forceinline__ T fold_identity (void)
{
	if constexpr (LLO_OP_SUM == OP) { return (T) 0; }
	else if constexpr (LLO_OP_PROD == OP) { return (T) 1; }
	else if constexpr (LLO_OP_MIN == OP)
	{
		if constexpr (is_fp<T>::value) { return (T) INFINITY; }
		else { return std::numeric_limits<T>::max(); }
	}
	else
	{
		if constexpr (is_fp<T>::value) { return (T) -INFINITY; }
		else { return std::numeric_limits<T>::lowest(); }
	}
}

template <typename T>
__device__ __forceinline__ T shfl_xor_any (T v, int lane_mask)
{
	if constexpr (sizeof(T) == 8)
	{
		long long bits;
		memcpy(&bits, &v, 8);
		bits = __shfl_xor_sync(0xffffffffu, bits, lane_mask);
		memcpy(&v, &bits, 8);
		return v;
	}
	else
	{
		int bits = 0;
		memcpy(&bits, &v, sizeof(T));
		bits = __shfl_xor_sync(0xffffffffu, bits, lane_mask);
		memcpy(&v, &bits, sizeof(T));
		return v;
	}
}

/// Fixed-shape tree over the 256 threads of a CTA; result valid in thread 0
template <int OP, typename T>
__device__ __forceinline__ T block_fold (T v, T* smem)
{
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1)
	{
		T other = shfl_xor_any(v, m);
		// lower lane on the left keeps the tree shape identical per lane
		v = (threadIdx.x & m) ? fold_step<OP,T>(other, v) :
			fold_step<OP,T>(v, other);
	}
	int warp = threadIdx.x >> 5;
	int lane = threadIdx.x & 31;
	if (0 == lane)
	{
		smem[warp] = v;
	}
	__syncthreads();
	if (0 == warp)
	{
		int nwarps = blockDim.x >> 5;
		T w = lane < nwarps ? smem[lane] : fold_identity<OP,T>();
#pragma unroll
		for (int m = 4; m >= 1; m >>= 1)
		{
			T other = shfl_xor_any(w, m);
			w = (lane & m) ? fold_step<OP,T>(other, w) :
				fold_step<OP,T>(w, other);
		}
		v = w;
	}
	return v;
}

struct OutDecode
{
	uint64_t outshape[LLO_RANK];
	int64_t stride[LLO_RANK];
	int64_t offset;
};

__device__ __forceinline__ int64_t source_base (const OutDecode& d, uint64_t o)
{
	int64_t base = d.offset;
	uint64_t rem = o;
#pragma unroll
	for (int k = 0; k < LLO_RANK; ++k)
	{
		uint64_t dim = d.outshape[k];
		if (dim > 1)
		{
			uint64_t q = rem / dim;
			base += (int64_t) (rem - q * dim) * d.stride[k];
			rem = q;
		}
	}
	return base;
}

// ---- row kernel -----------------------------------------------------------

struct RowParams
{
	OutDecode dec;
	uint64_t nrows;
	uint64_t len;       // folded elements per row
	uint64_t seg;       // elements per split (multiple of the vector width)
	uint32_t splits;
	int32_t vec_ok;
	int32_t accumulate; // fold into dst (only when splits == 1)
	int32_t pad;
};

template <int OP, typename T>
__global__ void __launch_bounds__(256)
row_fold_kernel (T* __restrict__ dst, const T* __restrict__ src,
	const __grid_constant__ RowParams p)
{
	constexpr int VEC = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;
	__shared__ T smem[8];
	uint64_t row = blockIdx.x / p.splits;
	uint32_t split = blockIdx.x - (uint32_t) (row * p.splits);
	const T* base = src + source_base(p.dec, row);
	uint64_t begin = (uint64_t) split * p.seg;
	uint64_t end = begin + p.seg < p.len ? begin + p.seg : p.len;

	T acc[4];
#pragma unroll
	for (int u = 0; u < 4; ++u)
	{
		acc[u] = fold_identity<OP,T>();
	}
	if (p.vec_ok)
	{
		// 4 x 128-bit loads in flight per thread
		uint64_t nvec = (end - begin) / VEC;
		const uint4* v = reinterpret_cast<const uint4*>(base + begin);
		uint64_t i = threadIdx.x;
		for (; i + 3 * 256 < nvec; i += 4 * 256)
		{
			uint4 raw[4];
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				raw[u] = __ldcs(v + i + u * 256);
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				const T* t = reinterpret_cast<const T*>(&raw[u]);
#pragma unroll
				for (int k = 0; k < VEC; ++k)
				{
					acc[u] = fold_step<OP,T>(acc[u], t[k]);
				}
			}
		}
		for (; i < nvec; i += 256)
		{
			uint4 raw = __ldcs(v + i);
			const T* t = reinterpret_cast<const T*>(&raw);
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				acc[0] = fold_step<OP,T>(acc[0], t[k]);
			}
		}
		for (uint64_t j = begin + nvec * VEC + threadIdx.x; j < end; j += 256)
		{
			acc[1] = fold_step<OP,T>(acc[1], base[j]);
		}
	}
	else
	{
		uint64_t j = begin + threadIdx.x;
		for (; j + 3 * 256 < end; j += 4 * 256)
		{
			T raw[4];
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				raw[u] = base[j + u * 256];
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				acc[u] = fold_step<OP,T>(acc[u], raw[u]);
			}
		}
		for (; j < end; j += 256)
		{
			acc[0] = fold_step<OP,T>(acc[0], base[j]);
		}
	}
	T v = fold_step<OP,T>(fold_step<OP,T>(acc[0], acc[1]),
		fold_step<OP,T>(acc[2], acc[3]));
	v = block_fold<OP,T>(v, smem);
	if (0 == threadIdx.x)
	{
		T* o = dst + blockIdx.x;
		*o = p.accumulate ? fold_step<OP,T>(*o, v) : v;
	}
}

// ---- column kernel ---------------------------------------------------------

struct ColParams
{
	uint64_t ncols;      // kept elements, contiguous
	uint64_t nrows;      // folded elements
	int64_t row_stride;  // distance between folded elements
	int64_t offset;
	uint64_t slab;       // rows per slab
	uint32_t nslabs;
	uint32_t colblocks;
	int32_t accumulate;
	int32_t pad;
};

template <int OP, typename T>
__global__ void __launch_bounds__(256)
col_fold_kernel (T* __restrict__ dst, const T* __restrict__ src,
	const __grid_constant__ ColParams p)
{
	uint32_t slab = blockIdx.x / p.colblocks;
	uint32_t cb = blockIdx.x - slab * p.colblocks;
	uint64_t col = (uint64_t) cb * 256 + threadIdx.x;
	if (col >= p.ncols)
	{
		return;
	}
	uint64_t r0 = (uint64_t) slab * p.slab;
	uint64_t r1 = r0 + p.slab < p.nrows ? r0 + p.slab : p.nrows;
	const T* base = src + p.offset + col;
	T acc[8];
#pragma unroll
	for (int u = 0; u < 8; ++u)
	{
		acc[u] = fold_identity<OP,T>();
	}
	uint64_t r = r0;
	for (; r + 8 <= r1; r += 8)
	{
		T raw[8];
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
			raw[u] = __ldcs(base + (int64_t) (r + u) * p.row_stride);
		}
#pragma unroll
		for (int u = 0; u < 8; ++u)
		{
			acc[u] = fold_step<OP,T>(acc[u], raw[u]);
		}
	}
	for (; r < r1; ++r)
	{
		acc[0] = fold_step<OP,T>(acc[0], base[(int64_t) r * p.row_stride]);
	}
	T v01 = fold_step<OP,T>(acc[0], acc[1]);
	T v23 = fold_step<OP,T>(acc[2], acc[3]);
	T v45 = fold_step<OP,T>(acc[4], acc[5]);
	T v67 = fold_step<OP,T>(acc[6], acc[7]);
	T v = fold_step<OP,T>(fold_step<OP,T>(v01, v23),
		fold_step<OP,T>(v45, v67));
	T* o = dst + (uint64_t) slab * p.ncols + col;
	*o = p.accumulate ? fold_step<OP,T>(*o, v) : v;
}

// ---- small-K kernel: [nrows][ncols] dense, ncols <= 128 --------------------

template <int OP, typename T>
__global__ void __launch_bounds__(256)
smallk_fold_kernel (T* __restrict__ dst, const T* __restrict__ src,
	const __grid_constant__ ColParams p)
{
	__shared__ T smem[256];
	uint32_t ncols = (uint32_t) p.ncols;
	uint32_t lanes = 256 / ncols;          // rows handled per sweep
	uint32_t active = lanes * ncols;
	uint32_t slab = blockIdx.x;
	uint64_t r0 = (uint64_t) slab * p.slab;
	uint64_t r1 = r0 + p.slab < p.nrows ? r0 + p.slab : p.nrows;
	T acc[4];
#pragma unroll
	for (int u = 0; u < 4; ++u)
	{
		acc[u] = fold_identity<OP,T>();
	}
	if (threadIdx.x < active)
	{
		uint32_t lane = threadIdx.x / ncols;
		const T* base = src + p.offset + threadIdx.x; // row `lane`, own column
		uint64_t r = r0 + lane;
		uint64_t sweep = (uint64_t) lanes * ncols;
		for (; r + 3 * lanes < r1; r += 4 * lanes)
		{
			T raw[4];
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				raw[u] = __ldcs(base + (r - lane + (uint64_t) u * lanes) * ncols);
			}
#pragma unroll
			for (int u = 0; u < 4; ++u)
			{
				acc[u] = fold_step<OP,T>(acc[u], raw[u]);
			}
		}
		for (; r < r1; r += lanes)
		{
			acc[0] = fold_step<OP,T>(acc[0], base[(r - lane) * ncols]);
		}
		(void) sweep;
	}
	smem[threadIdx.x] = fold_step<OP,T>(fold_step<OP,T>(acc[0], acc[1]),
		fold_step<OP,T>(acc[2], acc[3]));
	__syncthreads();
	if (threadIdx.x < ncols)
	{
		T v = smem[threadIdx.x];
		for (uint32_t lane = 1; lane < lanes; ++lane)
		{
			v = fold_step<OP,T>(v, smem[lane * ncols + threadIdx.x]);
		}
		T* o = dst + (uint64_t) slab * ncols + threadIdx.x;
		*o = p.accumulate ? fold_step<OP,T>(*o, v) : v;
	}
}

// ---- sequential kernel -----------------------------------------------------

struct SeqParams
{
	OutDecode dec;
	uint64_t nout;
	int32_t nfold;
	int32_t accumulate;
	uint64_t fold_size[LLO_RANK];
	int64_t fold_stride[LLO_RANK];
};

template <typename T>
__global__ void __launch_bounds__(128)
seq_fold_kernel (int opcode, T* __restrict__ dst, const T* __restrict__ src,
	const __grid_constant__ SeqParams p)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		o < p.nout; o += step)
	{
		const T* base = src + source_base(p.dec, o);
		uint64_t c[LLO_RANK];
		for (int k = 0; k < LLO_RANK; ++k)
		{
			c[k] = 0;
		}
		int64_t off = 0;
		T acc;
		bool have = false;
		if (p.accumulate)
		{
			acc = dst[o];
			have = true;
		}
		bool more = true;
		while (more)
		{
			T v = base[off];
			acc = have ? apply_binary<T>(opcode, acc, v) : v;
			have = true;
			// odometer over the folded dims, dim 0 fastest = ascending
			// source index = reference order
			more = false;
			for (int k = 0; k < p.nfold; ++k)
			{
				off += p.fold_stride[k];
				if (++c[k] < p.fold_size[k])
				{
					more = true;
					break;
				}
				off -= (int64_t) c[k] * p.fold_stride[k];
				c[k] = 0;
			}
		}
		dst[o] = acc;
	}
}

// ---- host side ---------------------------------------------------------------

template <int OP, typename T>
static int run_rows (llo_cuda_ctx* ctx, T* out, const T* src,
	const OutDecode& dec, uint64_t nrows, uint64_t len, bool aligned_rows,
	int accumulate)
{
	constexpr uint64_t lanes = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;
	// enough CTAs to fill the machine, but no split below 16K elements
	uint64_t want = (uint64_t) ctx->sm_count * 8;
	uint64_t splits = 1;
	if (nrows < want)
	{
		splits = (want + nrows - 1) / nrows;
		uint64_t maxsplits = (len + 16383) / 16384;
		if (splits > maxsplits)
		{
			splits = maxsplits;
		}
		if (splits < 1)
		{
			splits = 1;
		}
	}
	uint64_t seg = (len + splits - 1) / splits;
	seg = (seg + 1023) / 1024 * 1024; // keeps every segment 128-bit aligned
	splits = (len + seg - 1) / seg;

	RowParams p;
	p.dec = dec;
	p.nrows = nrows;
	p.len = len;
	p.seg = seg;
	p.splits = (uint32_t) splits;
	p.vec_ok = (aligned_rows && lanes > 1) ? 1 : 0;
	p.pad = 0;
	uint64_t nblocks = nrows * splits;
	if (nblocks > 0x7fffffffull)
	{
		return fail(LLO_ERR_FATAL, "reduction with %llu rows is too large",
			(unsigned long long) nrows);
	}
	if (1 == splits)
	{
		p.accumulate = accumulate;
		row_fold_kernel<OP,T><<<(unsigned) nblocks, 256, 0, ctx->stream>>>(
			out, src, p);
		return launched(ctx, "row_fold_kernel");
	}
	Scratch scratch(ctx);
	T* partial = nullptr;
	LLO_TRY(scratch.get((void**) &partial, nblocks * sizeof(T)));
	p.accumulate = 0;
	row_fold_kernel<OP,T><<<(unsigned) nblocks, 256, 0, ctx->stream>>>(
		partial, src, p);
	LLO_TRY(launched(ctx, "row_fold_kernel"));
	// second level: rows of `splits` partials
	OutDecode pdec;
	for (int k = 0; k < LLO_RANK; ++k)
	{
		pdec.outshape[k] = 1;
		pdec.stride[k] = 0;
	}
	pdec.outshape[0] = nrows;
	pdec.stride[0] = (int64_t) splits;
	pdec.offset = 0;
	return run_rows<OP,T>(ctx, out, partial, pdec, nrows, splits,
		false, accumulate);
}

template <int OP, typename T>
static int run_cols (llo_cuda_ctx* ctx, T* out, const T* src,
	uint64_t ncols, uint64_t nrows, int64_t row_stride, int64_t offset,
	int accumulate)
{
	ColParams p;
	p.ncols = ncols;
	p.nrows = nrows;
	p.row_stride = row_stride;
	p.offset = offset;
	p.pad = 0;
	bool smallk = ncols <= 128 && row_stride == (int64_t) ncols;
	uint64_t colblocks = smallk ? 1 : (ncols + 255) / 256;
	// 48 CTAs per SM, i.e. 8 waves of the 6 resident ones: with 8 per SM (1.4
	// waves) the half-empty second wave cost 10% (5.84 -> 6.39 TB/s at
	// [16384,16384]; x24: 6.17, x36: 6.35)
	uint64_t want = (uint64_t) ctx->sm_count * 48;
	uint64_t nslabs = 1;
	uint64_t minrows = smallk ? 4096 / ncols + 1 : 64;
	if (colblocks < want)
	{
		nslabs = (want + colblocks - 1) / colblocks;
		uint64_t maxslabs = (nrows + minrows - 1) / minrows;
		if (nslabs > maxslabs)
		{
			nslabs = maxslabs;
		}
		if (nslabs < 1)
		{
			nslabs = 1;
		}
	}
	uint64_t slab = (nrows + nslabs - 1) / nslabs;
	nslabs = (nrows + slab - 1) / slab;
	p.slab = slab;
	p.nslabs = (uint32_t) nslabs;
	p.colblocks = (uint32_t) colblocks;
	unsigned grid = (unsigned) (colblocks * nslabs);

	Scratch scratch(ctx);
	T* dst = out;
	if (nslabs > 1)
	{
		LLO_TRY(scratch.get((void**) &dst, nslabs * ncols * sizeof(T)));
		p.accumulate = 0;
	}
	else
	{
		p.accumulate = accumulate;
	}
	if (smallk)
	{
		smallk_fold_kernel<OP,T><<<grid, 256, 0, ctx->stream>>>(dst, src, p);
		LLO_TRY(launched(ctx, "smallk_fold_kernel"));
	}
	else
	{
		col_fold_kernel<OP,T><<<grid, 256, 0, ctx->stream>>>(dst, src, p);
		LLO_TRY(launched(ctx, "col_fold_kernel"));
	}
	if (nslabs > 1)
	{
		return run_cols<OP,T>(ctx, out, dst, ncols, nslabs,
			(int64_t) ncols, 0, accumulate);
	}
	return LLO_OK;
}

template <typename T>
static int run_seq (llo_cuda_ctx* ctx, int opcode, T* out, const T* src,
	const ReduceDesc& desc, int accumulate)
{
	SeqParams p;
	for (int k = 0; k < LLO_RANK; ++k)
	{
		p.dec.outshape[k] = desc.outshape[k];
		p.dec.stride[k] = desc.inv.stride[k];
		p.fold_size[k] = k < desc.nfold ? desc.fold_size[k] : 1;
		p.fold_stride[k] = k < desc.nfold ? desc.fold_stride[k] : 0;
	}
	p.dec.offset = desc.inv.offset;
	p.nout = n_elems(desc.outshape);
	p.nfold = desc.nfold;
	p.accumulate = accumulate;
	uint64_t blocks = (p.nout + 127) / 128;
	uint64_t cap = (uint64_t) ctx->sm_count * 16;
	seq_fold_kernel<T><<<(unsigned) (blocks < cap ? blocks : cap), 128, 0,
		ctx->stream>>>(opcode, out, src, p);
	return launched(ctx, "seq_fold_kernel");
}

template <int OP, typename T>
static int run_tree (llo_cuda_ctx* ctx, T* out, const T* src,
	const ReduceDesc& desc, int accumulate, bool& handled)
{
	handled = false;
	// merge the folded dims into one run when they are dense among themselves
	uint64_t flen = 1;
	int64_t fstride = 0;
	bool fdense = true;
	for (int k = 0; k < desc.nfold; ++k)
	{
		if (0 == k)
		{
			fstride = desc.fold_stride[0];
		}
		else if (desc.fold_stride[k] != fstride * (int64_t) flen)
		{
			fdense = false;
		}
		flen *= desc.fold_size[k];
	}
	if (0 == desc.nfold || false == fdense)
	{
		return LLO_OK;
	}
	uint64_t nout = n_elems(desc.outshape);
	OutDecode dec;
	for (int k = 0; k < LLO_RANK; ++k)
	{
		dec.outshape[k] = desc.outshape[k];
		dec.stride[k] = desc.inv.stride[k];
	}
	dec.offset = desc.inv.offset;

	if (1 == fstride)
	{
		// rows of contiguous folded elements
		constexpr int64_t lanes = 16 / sizeof(T) > 0 ? 16 / sizeof(T) : 1;
		bool aligned = (0 == ((uintptr_t) src & 15u)) &&
			(0 == desc.inv.offset % lanes);
		for (int k = 0; k < LLO_RANK && aligned; ++k)
		{
			if (desc.outshape[k] > 1 && 0 != desc.inv.stride[k] % lanes)
			{
				aligned = false;
			}
		}
		handled = true;
		return run_rows<OP,T>(ctx, out, src, dec, nout, flen, aligned,
			accumulate);
	}
	// kept elements dense and contiguous?
	bool kdense = true;
	int64_t expect = 1;
	for (int k = 0; k < LLO_RANK; ++k)
	{
		if (desc.outshape[k] > 1)
		{
			if (desc.inv.stride[k] != expect)
			{
				kdense = false;
			}
			expect *= (int64_t) desc.outshape[k];
		}
	}
	if (kdense && fstride > 0)
	{
		handled = true;
		return run_cols<OP,T>(ctx, out, src, nout, flen, fstride,
			desc.inv.offset, accumulate);
	}
	return LLO_OK;
}

int reduce_launch (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const void* src, const ReduceDesc& desc, int accumulate, int mode)
{
	if (false == is_nnary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "cannot reduce with opcode %s",
			op_name(opcode));
	}
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		bool handled = false;
		// integer and min/max folds are order-independent, so the tree is
		// already bit-exact; only floating SUM / PROD have a distinct
		// reference-order mode
		bool order_matters = is_fp<T>::value &&
			(LLO_OP_SUM == opcode || LLO_OP_PROD == opcode);
		if (LLO_REDUCE_SEQUENTIAL != mode || false == order_matters)
		{
			switch (opcode)
			{
				case LLO_OP_SUM: LLO_TRY((run_tree<LLO_OP_SUM,T>(ctx, (T*) out, (const T*) src, desc, accumulate, handled))); break;
				case LLO_OP_PROD: LLO_TRY((run_tree<LLO_OP_PROD,T>(ctx, (T*) out, (const T*) src, desc, accumulate, handled))); break;
				case LLO_OP_MIN: LLO_TRY((run_tree<LLO_OP_MIN,T>(ctx, (T*) out, (const T*) src, desc, accumulate, handled))); break;
				default: LLO_TRY((run_tree<LLO_OP_MAX,T>(ctx, (T*) out, (const T*) src, desc, accumulate, handled))); break;
			}
		}
		if (false == handled)
		{
			return run_seq<T>(ctx, opcode, (T*) out, (const T*) src, desc,
				accumulate);
		}
	})
	return LLO_OK;
}

}

using namespace llo_cuda;

extern "C" int llo_cuda_reduce (llo_cuda_ctx* ctx, int opcode, int dtype,
	void* out, const void* in, const uint64_t inshape[LLO_RANK],
	int first_dim, int mode)
{
	if (first_dim < 0 || first_dim >= LLO_RANK)
	{
		return fail(LLO_ERR_FATAL, "cannot reduce dimensions beyond "
			"rank_cap %d", LLO_RANK);
	}
	ReduceDesc desc;
	int64_t stride = 1;
	desc.nfold = 0;
	desc.inv.offset = 0;
	for (int d = 0; d < LLO_RANK; ++d)
	{
		if (d < first_dim)
		{
			desc.outshape[d] = inshape[d];
			desc.inv.stride[d] = stride;
		}
		else
		{
			desc.outshape[d] = 1;
			desc.inv.stride[d] = 0;
			if (inshape[d] > 1)
			{
				desc.fold_size[desc.nfold] = inshape[d];
				desc.fold_stride[desc.nfold] = stride;
				++desc.nfold;
			}
		}
		stride *= (int64_t) inshape[d];
	}
	if (0 == desc.nfold)
	{
		return llo_cuda_d2d(ctx, out, in, n_elems(inshape) * dtype_size(dtype));
	}
	return reduce_launch(ctx, opcode, dtype, out, in, desc, 0, mode);
}

extern "C" int llo_cuda_reduce_view (llo_cuda_ctx* ctx, int opcode, int dtype,
	void* out, const uint64_t outshape[LLO_RANK], const llo_cuda_view* src,
	int nfold, const uint64_t* fold_size, const int64_t* fold_stride,
	int accumulate, int mode)
{
	if (nfold < 0 || nfold > LLO_RANK)
	{
		return fail(LLO_ERR_FATAL, "cannot fold %d dimensions", nfold);
	}
	ReduceDesc desc;
	desc.nfold = 0;
	desc.inv.offset = src->offset;
	for (int d = 0; d < LLO_RANK; ++d)
	{
		desc.outshape[d] = outshape[d];
		desc.inv.stride[d] = outshape[d] > 1 ? src->stride[d] : 0;
	}
	for (int k = 0; k < nfold; ++k)
	{
		if (fold_size[k] > 1)
		{
			desc.fold_size[desc.nfold] = fold_size[k];
			desc.fold_stride[desc.nfold] = fold_stride[k];
			++desc.nfold;
		}
	}
	if (0 == desc.nfold)
	{
		// nothing to fold: a strided copy (or accumulate) of the view
		desc.fold_size[0] = 1;
		desc.fold_stride[0] = 0;
		desc.nfold = 1;
	}
	return reduce_launch(ctx, opcode, dtype, out, src->data, desc,
		accumulate, mode);
}
