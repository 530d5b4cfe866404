///
/// elementwise.cu — launch side of the fused elementwise engine (vm.cuh)
/// and the `llo_cuda_elementwise` entry point.
///

#include <cstring>
#include <utility>

#include "vm.cuh"

namespace llo_cuda
{

// ---- flat launch: one chunk per thread ---------------------------------------

template <typename T>
__global__ void __launch_bounds__(256, 4)
vm_flat_kernel (const __grid_constant__ VmParams p)
{
	constexpr int VEC = VmVec<T>::value;
	uint64_t base = ((uint64_t) blockIdx.x * 256 + threadIdx.x) * VEC;
	if (base >= p.n)
	{
		return;
	}
	uint64_t left = p.n - base;
	int valid = left < (uint64_t) VEC ? (int) left : VEC;
	FlatLoader<T,VEC> loader = {p, base, valid};
	T acc[VEC];
	vm_run<T,VEC>(acc, p, loader);
	store_run<T,VEC>(static_cast<T*>(p.out) + base, acc, p.out_vec_ok, valid);
}

// ---- tile launch: transposed inputs staged through shared memory ---------------

/// Offset contributed by the dims beyond the tile plane (>= 2); `brest` is
/// the block's linear index over those dims
__device__ __forceinline__ int64_t rest_offset (const VmParams& p,
	const int64_t (&stride)[LLO_RANK], uint32_t brest)
{
	int64_t off = 0;
	for (int d = 2; d < p.ndims; ++d)
	{
		uint32_t dim = (uint32_t) p.dims[d];
		uint32_t q = brest / dim;
		off += (int64_t) (brest - q * dim) * stride[d];
		brest = q;
	}
	return off;
}

template <typename T, int VEC>
struct TileLoader
{
	const VmParams& p;
	const T* tiles;          // [slot][VM_TILEP][pitch]
	uint32_t c0;             // coordinates of this thread's chunk in the plane
	uint32_t c1;
	uint32_t brest;
	int col;                 // chunk position inside the tile
	int row;
	int valid;

	static constexpr int PITCH = 8 * VEC + 1;

	__device__ __forceinline__ void load (T (&dst)[VEC], int k) const
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE == in.mode)
		{
			const T* src = tiles +
				((size_t) in.vec_ok * VM_TILEP + row) * PITCH + col;
#pragma unroll
			for (int e = 0; e < VEC; ++e)
			{
				dst[e] = src[e];
			}
		}
		else
		{
			int64_t off = in.offset + (int64_t) c0 * in.stride[0] +
				(int64_t) c1 * in.stride[1] + rest_offset(p, in.stride, brest);
			load_strided_run<T,VEC>(dst, in, off, valid);
		}
	}
};

template <typename T>
__global__ void __launch_bounds__(256, 3)
vm_tile_kernel (const __grid_constant__ VmParams p)
{
	constexpr int VEC = VmVec<T>::value;
	constexpr int TILE0 = 8 * VEC;
	constexpr int PITCH = TILE0 + 1;
	extern __shared__ __align__(16) unsigned char tile_raw[];
	T* tiles = reinterpret_cast<T*>(tile_raw);

	// block -> (tile along dim 0, tile along dim 1, every other dim)
	uint32_t b = blockIdx.x;
	uint32_t t0 = b % p.tiles0;
	b /= p.tiles0;
	uint32_t tp = b % p.tilesp;
	uint32_t brest = b / p.tilesp;
	uint32_t o0 = t0 * TILE0;
	uint32_t o1 = tp * VM_TILEP;
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
		T* tile = tiles + (size_t) in.vec_ok * VM_TILEP * PITCH;
		// a warp covers 32 consecutive elements of the input's contiguous
		// direction; 256 threads sweep the TILE0 x TILEP tile in
		// TILE0 * TILEP / 256 steps, all loads issued before any is used
		constexpr int STEPS = TILE0 * VM_TILEP / 256;
		T v[STEPS];
#pragma unroll
		for (int j = 0; j < STEPS; ++j)
		{
			int e = threadIdx.x + 256 * j;
			int lp = e % VM_TILEP;
			int r0 = e / VM_TILEP;
			v[j] = T();
			if (o1 + lp < dim1 && o0 + r0 < dim0)
			{
				v[j] = __ldcs(data + origin + (int64_t) r0 * in.stride[0] + lp);
			}
		}
#pragma unroll
		for (int j = 0; j < STEPS; ++j)
		{
			int e = threadIdx.x + 256 * j;
			tile[(e % VM_TILEP) * PITCH + e / VM_TILEP] = v[j];
		}
	}
	__syncthreads();

	// 8 chunk columns x 32 rows per pass, TILEP / 32 passes
#pragma unroll 1
	for (int pass = 0; pass < VM_TILEP / 32; ++pass)
	{
		TileLoader<T,VEC> loader = {p, tiles, 0, 0, brest, 0, 0, 0};
		loader.col = (threadIdx.x & 7) * VEC;
		loader.row = (threadIdx.x >> 3) + 32 * pass;
		loader.c0 = o0 + loader.col;
		loader.c1 = o1 + loader.row;
		if (loader.c0 >= dim0 || loader.c1 >= dim1)
		{
			continue;
		}
		uint32_t left = dim0 - loader.c0;
		loader.valid = left < (uint32_t) VEC ? (int) left : VEC;
		T acc[VEC];
		vm_run<T,VEC>(acc, p, loader);
		T* out = static_cast<T*>(p.out) + (int64_t) loader.c0 * p.out_stride[0] +
			(int64_t) loader.c1 * p.out_stride[1] +
			rest_offset(p, p.out_stride, brest);
		store_run<T,VEC>(out, acc, p.out_vec_ok, loader.valid);
	}
}

static bool aligned16 (const void* ptr, int64_t elem_offset, size_t esize)
{
	return 0 == (((uintptr_t) ptr + (uintptr_t) (elem_offset * (int64_t) esize)) & 15u);
}

/// Validate the public stack program and resolve every instruction's stack
/// slot (VmOp::sel)
static int vm_encode (VmOp* code, int dtype, int ninputs, int nconsts,
	const llo_cuda_insn* program, int ninsns)
{
	if (ninputs < 0 || ninputs > LLO_VM_MAX_INPUTS ||
		nconsts < 0 || nconsts > LLO_VM_MAX_CONSTS ||
		ninsns <= 0 || ninsns > LLO_VM_MAX_INSNS)
	{
		return fail(LLO_ERR_FATAL, "elementwise program out of limits "
			"(%d inputs, %d consts, %d instructions)",
			ninputs, nconsts, ninsns);
	}
	int depth = 0;
	for (int pc = 0; pc < ninsns; ++pc)
	{
		int op = program[pc].op;
		int arg = program[pc].arg;
		int sel = -1;
		if (LLO_VM_LOAD == op || LLO_VM_CONST == op)
		{
			int limit = LLO_VM_LOAD == op ? ninputs : nconsts;
			if (arg >= limit)
			{
				return fail(LLO_ERR_FATAL, "program loads %s %d of %d",
					LLO_VM_LOAD == op ? "input" : "constant", arg, limit);
			}
			if (depth >= LLO_VM_MAX_DEPTH)
			{
				return fail(LLO_ERR_FATAL, "program needs more than %d stack "
					"slots", LLO_VM_MAX_DEPTH);
			}
			sel = (LLO_VM_LOAD == op ? VM_SEL_LOAD : VM_SEL_CONST) + depth;
			++depth;
		}
		else if (is_unary_op(op))
		{
			if (depth < 1)
			{
				return fail(LLO_ERR_FATAL, "program underflows its stack");
			}
			if (LLO_OP_NEG == op && is_unsigned_dtype(dtype))
			{
				// std::bad_function_call (llo/src/operator.cpp:39-60)
				return fail(LLO_ERR_UNSUPPORTED_TYPED_OP,
					"NEG is not defined for %s", dtype_name(dtype));
			}
			int slot = depth - 1;
			switch (op)
			{
				case LLO_OP_ABS: sel = VM_SEL_ABS + slot; break;
				case LLO_OP_NEG: sel = VM_SEL_NEG + slot; break;
				case LLO_OP_ROUND: sel = VM_SEL_ROUND + slot; break;
				case LLO_OP_SIN: sel = VM_SEL_HEAVY + 4 * VM_H_SIN + slot; break;
				case LLO_OP_COS: sel = VM_SEL_HEAVY + 4 * VM_H_COS + slot; break;
				case LLO_OP_TAN: sel = VM_SEL_HEAVY + 4 * VM_H_TAN + slot; break;
				case LLO_OP_EXP: sel = VM_SEL_HEAVY + 4 * VM_H_EXP + slot; break;
				case LLO_OP_LOG: sel = VM_SEL_HEAVY + 4 * VM_H_LOG + slot; break;
				default: sel = VM_SEL_HEAVY + 4 * VM_H_SQRT + slot; break;
			}
		}
		else if (is_binary_op(op) || is_nnary_op(op))
		{
			if (depth < 2)
			{
				return fail(LLO_ERR_FATAL, "program underflows its stack");
			}
			int slot = depth - 2;
			switch (op)
			{
				case LLO_OP_SUM: sel = VM_SEL_SUM + slot; break;
				case LLO_OP_SUB: sel = VM_SEL_SUB + slot; break;
				case LLO_OP_PROD: sel = VM_SEL_PROD + slot; break;
				case LLO_OP_MIN: sel = VM_SEL_MIN + slot; break;
				case LLO_OP_MAX: sel = VM_SEL_MAX + slot; break;
				case LLO_OP_EQ: sel = VM_SEL_EQ + slot; break;
				case LLO_OP_NEQ: sel = VM_SEL_NEQ + slot; break;
				case LLO_OP_LT: sel = VM_SEL_LT + slot; break;
				case LLO_OP_GT: sel = VM_SEL_GT + slot; break;
				case LLO_OP_POW: sel = VM_SEL_HEAVY + 4 * VM_H_POW + slot; break;
				default: sel = VM_SEL_HEAVY + 4 * VM_H_DIV + slot; break;
			}
			--depth;
		}
		else
		{
			return fail(LLO_ERR_FATAL, "opcode %d is not elementwise", op);
		}
		code[pc].sel = (uint8_t) sel;
		code[pc].arg = (uint8_t) arg;
	}
	if (1 != depth)
	{
		return fail(LLO_ERR_FATAL, "program leaves %d values", depth);
	}
	return LLO_OK;
}

int vm_prepare (VmParams& p, int dtype, const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs, const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns, int vec)
{
	size_t esize = dtype_size(dtype);
	if (0 == esize)
	{
		return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
	}
	std::memset(&p, 0, sizeof(p));
	LLO_TRY(vm_encode(p.code, dtype, ninputs, nconsts, program, ninsns));
	p.n = n_elems(outshape);
	p.ninsns = ninsns;
	p.ninputs = ninputs;
	for (int c = 0; c < nconsts; ++c)
	{
		std::memcpy(&p.consts[c], (const char*) consts + c * esize, esize);
	}

	// drop size-1 dims, then merge neighbours every strided input agrees on
	uint64_t dims[LLO_RANK];
	int64_t strides[LLO_VM_MAX_INPUTS][LLO_RANK];
	int nd = 0;
	for (int d = 0; d < LLO_RANK; ++d)
	{
		if (outshape[d] > 1)
		{
			dims[nd] = outshape[d];
			for (int k = 0; k < ninputs; ++k)
			{
				strides[k][nd] = inputs[k].stride[d];
			}
			++nd;
		}
	}
	if (0 == nd)
	{
		dims[0] = 1;
		for (int k = 0; k < ninputs; ++k)
		{
			strides[k][0] = 0;
		}
		nd = 1;
	}
	int m = 0;
	for (int d = 1; d < nd; ++d)
	{
		bool mergeable = true;
		for (int k = 0; k < ninputs; ++k)
		{
			if (VM_IN_STRIDED == inputs[k].mode && strides[k][d] !=
				strides[k][m] * (int64_t) dims[m])
			{
				mergeable = false;
				break;
			}
		}
		if (mergeable)
		{
			dims[m] *= dims[d];
		}
		else
		{
			++m;
			dims[m] = dims[d];
			for (int k = 0; k < ninputs; ++k)
			{
				strides[k][m] = strides[k][d];
			}
		}
	}
	nd = m + 1;
	p.ndims = nd;
	bool fits32 = p.n < (1ull << 31);
	int64_t dense = 1;
	for (int d = 0; d < nd; ++d)
	{
		p.dims[d] = dims[d];
		p.out_stride[d] = dense;
		dense *= (int64_t) dims[d];
	}
	p.rows_aligned = (0 == dims[0] % (uint64_t) vec) ? 1 : 0;
	int64_t lanes16 = (int64_t) (16 / esize);
	for (int k = 0; k < ninputs; ++k)
	{
		VmInput in = inputs[k];
		if (VM_IN_STRIDED == in.mode)
		{
			// farthest element the view can reach, for 32-bit addressing
			int64_t reach = in.offset < 0 ? -in.offset : in.offset;
			for (int d = 0; d < LLO_RANK; ++d)
			{
				in.stride[d] = d < nd ? strides[k][d] : 0;
				if (d < nd)
				{
					int64_t s = in.stride[d] < 0 ? -in.stride[d] : in.stride[d];
					reach += s * (int64_t) (dims[d] - 1);
				}
			}
			fits32 = fits32 && reach < (1ll << 31);
			if (1 == nd && (1 == in.stride[0] || 1 == dims[0]))
			{
				in.mode = VM_IN_LINEAR;
			}
		}
		if (VM_IN_LINEAR == in.mode)
		{
			in.vec_ok = aligned16(in.data, in.offset, esize) ? 1 : 0;
		}
		else if (VM_IN_STRIDED == in.mode)
		{
			bool ok = p.rows_aligned && 1 == in.stride[0] &&
				aligned16(in.data, in.offset, esize);
			for (int d = 1; d < nd && ok; ++d)
			{
				ok = (0 == in.stride[d] % lanes16);
			}
			in.vec_ok = ok ? 1 : 0;
		}
		p.in[k] = in;
	}
	p.idx32 = fits32 ? 1 : 0;
	if (fits32)
	{
		for (int d = 0; d < nd; ++d)
		{
			p.div[d] = make_fastdiv((uint32_t) dims[d]);
		}
	}
	return LLO_OK;
}

/// Re-shape prepared params for the tile launch if some strided input is
/// transposed (contiguous along a collapsed dim other than 0).  Returns the
/// number of staged inputs, 0 when the flat launch should be used.
static int plan_tiles (VmParams& p, size_t esize, int vec)
{
	if (p.ndims < 2 || p.n >= (1ull << 31))
	{
		return 0;
	}
	int pdim = -1;
	for (int k = 0; k < p.ninputs && pdim < 0; ++k)
	{
		const VmInput& in = p.in[k];
		if (VM_IN_STRIDED != in.mode || 0 == in.stride[0] || 1 == in.stride[0])
		{
			continue;
		}
		for (int d = 1; d < p.ndims; ++d)
		{
			if (1 == in.stride[d] && p.dims[d] >= 16)
			{
				pdim = d;
				break;
			}
		}
	}
	if (pdim < 0 || p.dims[0] < 16)
	{
		return 0;
	}
	for (int k = 0; k < p.ninputs; ++k)
	{
		if (VM_IN_INDEXED == p.in[k].mode)
		{
			return 0; // index tables address the flat output space
		}
	}
	// linear inputs become strided views with the output's dense strides
	for (int k = 0; k < p.ninputs; ++k)
	{
		VmInput& in = p.in[k];
		if (VM_IN_LINEAR == in.mode)
		{
			for (int d = 0; d < LLO_RANK; ++d)
			{
				in.stride[d] = d < p.ndims ? p.out_stride[d] : 0;
			}
			in.mode = VM_IN_STRIDED;
		}
	}
	// move pdim to position 1 (everything is stride-addressed from here on)
	if (1 != pdim)
	{
		std::swap(p.dims[1], p.dims[pdim]);
		std::swap(p.out_stride[1], p.out_stride[pdim]);
		for (int k = 0; k < p.ninputs; ++k)
		{
			std::swap(p.in[k].stride[1], p.in[k].stride[pdim]);
		}
	}
	int tile0 = 8 * vec;
	int nstaged = 0;
	int64_t lanes16 = (int64_t) (16 / esize);
	bool rows_aligned = (0 == p.dims[0] % (uint64_t) vec);
	for (int k = 0; k < p.ninputs; ++k)
	{
		VmInput& in = p.in[k];
		if (VM_IN_STRIDED != in.mode)
		{
			continue;
		}
		bool transposed = 0 != in.stride[0] && 1 != in.stride[0] &&
			1 == in.stride[1];
		if (transposed && nstaged < VM_MAX_TILE_INPUTS)
		{
			in.mode = VM_IN_TILE;
			in.vec_ok = nstaged++;
			continue;
		}
		bool ok = rows_aligned && 1 == in.stride[0] &&
			aligned16(in.data, in.offset, esize);
		for (int d = 1; d < p.ndims && ok; ++d)
		{
			ok = (0 == in.stride[d] % lanes16);
		}
		in.vec_ok = ok ? 1 : 0;
	}
	bool out_ok = rows_aligned && 0 != p.out_vec_ok;
	for (int d = 1; d < p.ndims && out_ok; ++d)
	{
		out_ok = (0 == p.out_stride[d] % lanes16);
	}
	p.out_vec_ok = out_ok ? 1 : 0;
	p.rows_aligned = 1; // tile chunks never leave their row
	p.tiles0 = (uint32_t) ((p.dims[0] + tile0 - 1) / tile0);
	p.tilesp = (uint32_t) ((p.dims[1] + VM_TILEP - 1) / VM_TILEP);
	return nstaged;
}

int vm_launch (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs,
	const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns)
{
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		constexpr int VEC = VmVec<T>::value;
		VmParams p;
		LLO_TRY(vm_prepare(p, dtype, outshape, inputs, ninputs,
			consts, nconsts, program, ninsns, VEC));
		p.out = out;
		p.out_vec_ok = aligned16(out, 0, sizeof(T)) ? 1 : 0;
		int nstaged = plan_tiles(p, sizeof(T), VEC);
		if (nstaged > 0)
		{
			uint64_t blocks = (uint64_t) p.tiles0 * p.tilesp;
			for (int d = 2; d < p.ndims; ++d)
			{
				blocks *= p.dims[d];
			}
			if (blocks >= 0x7fffffffull)
			{
				return fail(LLO_ERR_FATAL, "elementwise launch too large");
			}
			size_t smem = (size_t) nstaged * VM_TILEP * (8 * VEC + 1) *
				sizeof(T);
			vm_tile_kernel<T><<<(unsigned) blocks, 256, smem,
				ctx->stream>>>(p);
			return launched(ctx, "vm_tile_kernel");
		}
		uint64_t nchunks = (p.n + VEC - 1) / VEC;
		uint64_t blocks = (nchunks + 255) / 256;
		if (blocks >= 0x7fffffffull)
		{
			return fail(LLO_ERR_FATAL, "elementwise launch too large");
		}
		vm_flat_kernel<T><<<(unsigned) blocks, 256, 0, ctx->stream>>>(p);
		return launched(ctx, "vm_flat_kernel");
	})
	return LLO_OK;
}

}

using namespace llo_cuda;

extern "C" int llo_cuda_elementwise (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_view* inputs, int ninputs,
	const double* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns)
{
	if (ninputs < 0 || ninputs > LLO_VM_MAX_INPUTS ||
		nconsts < 0 || nconsts > LLO_VM_MAX_CONSTS)
	{
		return fail(LLO_ERR_FATAL, "elementwise program out of limits "
			"(%d inputs, %d consts)", ninputs, nconsts);
	}
	int64_t dense[LLO_RANK];
	int64_t acc = 1;
	for (int d = 0; d < LLO_RANK; ++d)
	{
		dense[d] = acc;
		acc *= (int64_t) outshape[d];
	}
	VmInput vin[LLO_VM_MAX_INPUTS];
	for (int k = 0; k < ninputs; ++k)
	{
		std::memset(&vin[k], 0, sizeof(VmInput));
		vin[k].data = inputs[k].data;
		vin[k].offset = inputs[k].offset;
		bool linear = true;
		for (int d = 0; d < LLO_RANK; ++d)
		{
			vin[k].stride[d] = inputs[k].stride[d];
			if (outshape[d] > 1 && inputs[k].stride[d] != dense[d])
			{
				linear = false;
			}
		}
		vin[k].mode = linear ? VM_IN_LINEAR : VM_IN_STRIDED;
	}
	// constants arrive as doubles and are converted to the element type
	// with the C++ conversion the reference applies to leaves
	size_t esize = dtype_size(dtype);
	unsigned char packed[LLO_VM_MAX_CONSTS * 8];
	for (int c = 0; c < nconsts; ++c)
	{
		unsigned char* dst = packed + c * esize;
		switch (dtype)
		{
			case LLO_DT_DOUBLE: { double v = consts[c]; std::memcpy(dst, &v, 8); } break;
			case LLO_DT_FLOAT: { float v = (float) consts[c]; std::memcpy(dst, &v, 4); } break;
			case LLO_DT_INT8: { int8_t v = (int8_t) consts[c]; std::memcpy(dst, &v, 1); } break;
			case LLO_DT_UINT8: { uint8_t v = (uint8_t) consts[c]; std::memcpy(dst, &v, 1); } break;
			case LLO_DT_INT16: { int16_t v = (int16_t) consts[c]; std::memcpy(dst, &v, 2); } break;
			case LLO_DT_UINT16: { uint16_t v = (uint16_t) consts[c]; std::memcpy(dst, &v, 2); } break;
			case LLO_DT_INT32: { int32_t v = (int32_t) consts[c]; std::memcpy(dst, &v, 4); } break;
			case LLO_DT_UINT32: { uint32_t v = (uint32_t) consts[c]; std::memcpy(dst, &v, 4); } break;
			case LLO_DT_INT64: { int64_t v = (int64_t) consts[c]; std::memcpy(dst, &v, 8); } break;
			case LLO_DT_UINT64: { uint64_t v = (uint64_t) consts[c]; std::memcpy(dst, &v, 8); } break;
			default: return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
		}
	}
	return vm_launch(ctx, dtype, out, outshape, vin, ninputs,
		packed, nconsts, program, ninsns);
}
