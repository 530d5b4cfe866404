///
/// elementwise.cu — launch side of the fused elementwise engine (vm.cuh)
/// and the `llo_cuda_elementwise` entry point.
///

#include <algorithm>
#include <cstring>
#include <utility>

#include "vm_kernels.cuh"
#include "specialize.hpp"

namespace llo_cuda
{

static bool aligned16 (const void* ptr, int64_t elem_offset, size_t esize)
{
	return 0 == (((uintptr_t) ptr + (uintptr_t) (elem_offset * (int64_t) esize)) & 15u);
}

// ---- stack program -> accumulator program --------------------------------------

namespace
{

struct Expr
{
	int kind;   // 0 input, 1 constant, 2 unary, 3 binary
	int op;     // public opcode
	int arg;    // input / constant number
	int lhs;
	int rhs;
	int need;   // spill levels needed to evaluate
};

struct Encoder
{
	Expr nodes[LLO_VM_MAX_INSNS];
	VmOp* code;
	int ncode = 0;
	int depth = 0;
	int peak = 0;
	bool overflow = false;

	bool leaf (int n) const { return nodes[n].kind < 2; }

	uint8_t source (int n) const
	{
		return (uint8_t) (((0 == nodes[n].kind ? VM_SRC_IN : VM_SRC_CONST) << 4) |
			nodes[n].arg);
	}

	void emit (int sel, uint8_t arg)
	{
		if (ncode >= VM_MAX_CODE)
		{
			overflow = true;
			return;
		}
		code[ncode].sel = (uint8_t) sel;
		code[ncode].arg = arg;
		++ncode;
	}

	static int unary_sel (int op)
	{
		switch (op)
		{
			case LLO_OP_ABS: return VM_S_ABS;
			case LLO_OP_NEG: return VM_S_NEG;
			case LLO_OP_ROUND: return VM_S_ROUND;
			case LLO_OP_SIN: return VM_S_SIN;
			case LLO_OP_COS: return VM_S_COS;
			case LLO_OP_TAN: return VM_S_TAN;
			case LLO_OP_EXP: return VM_S_EXP;
			case LLO_OP_LOG: return VM_S_LOG;
			default: return VM_S_SQRT;
		}
	}

	/// acc <- op(acc, x), or op(x, acc) when reversed
	static int binary_sel (int op, bool reversed)
	{
		switch (op)
		{
			case LLO_OP_SUM: return VM_S_SUM;
			case LLO_OP_PROD: return VM_S_PROD;
			case LLO_OP_EQ: return VM_S_EQ;
			case LLO_OP_NEQ: return VM_S_NEQ;
			case LLO_OP_SUB: return reversed ? VM_S_RSUB : VM_S_SUB;
			case LLO_OP_MIN: return reversed ? VM_S_RMIN : VM_S_MIN;
			case LLO_OP_MAX: return reversed ? VM_S_RMAX : VM_S_MAX;
			case LLO_OP_LT: return reversed ? VM_S_GT : VM_S_LT;
			case LLO_OP_GT: return reversed ? VM_S_LT : VM_S_GT;
			case LLO_OP_DIV: return reversed ? VM_S_RDIV : VM_S_DIV;
			default: return reversed ? VM_S_RPOW : VM_S_POW;
		}
	}

	void gen (int n)
	{
		const Expr& x = nodes[n];
		if (x.kind < 2)
		{
			emit(VM_S_SET, source(n));
		}
		else if (2 == x.kind)
		{
			gen(x.lhs);
			emit(unary_sel(x.op), 0);
		}
		else if (leaf(x.rhs))
		{
			gen(x.lhs);
			emit(binary_sel(x.op, false), source(x.rhs));
		}
		else if (leaf(x.lhs))
		{
			gen(x.rhs);
			emit(binary_sel(x.op, true), source(x.lhs));
		}
		else
		{
			// both operands are sub-expressions: the hungrier one first
			bool lhs_first = nodes[x.lhs].need >= nodes[x.rhs].need;
			gen(lhs_first ? x.lhs : x.rhs);
			emit(VM_S_PUSH, 0);
			++depth;
			peak = depth > peak ? depth : peak;
			gen(lhs_first ? x.rhs : x.lhs);
			--depth;
			emit(binary_sel(x.op, lhs_first), (uint8_t) (VM_SRC_POP << 4));
		}
	}
};

}

/// Validate the public stack program and compile it to accumulator form
static int vm_encode (VmParams& p, int dtype, int ninputs, int nconsts,
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
	Encoder enc;
	enc.code = p.code;
	int stack[LLO_VM_MAX_DEPTH];
	int depth = 0;
	for (int pc = 0; pc < ninsns; ++pc)
	{
		int op = program[pc].op;
		int arg = program[pc].arg;
		Expr& node = enc.nodes[pc];
		node = Expr{0, op, arg, -1, -1, 0};
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
			node.kind = LLO_VM_LOAD == op ? 0 : 1;
			stack[depth++] = pc;
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
			node.kind = 2;
			node.lhs = stack[depth - 1];
			node.need = enc.nodes[node.lhs].need;
			stack[depth - 1] = pc;
		}
		else if (is_binary_op(op) || is_nnary_op(op))
		{
			if (depth < 2)
			{
				return fail(LLO_ERR_FATAL, "program underflows its stack");
			}
			node.kind = 3;
			node.lhs = stack[depth - 2];
			node.rhs = stack[depth - 1];
			int nl = enc.nodes[node.lhs].need;
			int nr = enc.nodes[node.rhs].need;
			if (enc.leaf(node.lhs) || enc.leaf(node.rhs))
			{
				node.need = nl > nr ? nl : nr;
			}
			else
			{
				int hi = nl >= nr ? nl : nr;
				int lo = nl >= nr ? nr : nl;
				node.need = hi > lo + 1 ? hi : lo + 1;
			}
			--depth;
			stack[depth - 1] = pc;
		}
		else
		{
			return fail(LLO_ERR_FATAL, "opcode %d is not elementwise", op);
		}
	}
	if (1 != depth)
	{
		return fail(LLO_ERR_FATAL, "program leaves %d values", depth);
	}
	enc.gen(stack[0]);
	if (enc.overflow || enc.peak > VM_MAX_SPILL)
	{
		return fail(LLO_ERR_FATAL, "elementwise program too large for the "
			"engine (%d instructions, %d spill levels)", enc.ncode, enc.peak);
	}
	p.ninsns = enc.ncode;
	p.spill_depth = enc.peak;
	return LLO_OK;
}

int vm_prepare (VmParams& p, int dtype, const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs, const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns, int chunk, int span)
{
	size_t esize = dtype_size(dtype);
	if (0 == esize)
	{
		return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
	}
	std::memset(&p, 0, sizeof(p));
	LLO_TRY(vm_encode(p, dtype, ninputs, nconsts, program, ninsns));
	p.n = n_elems(outshape);
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
	p.rows_aligned = (0 == dims[0] % (uint64_t) chunk) ? 1 : 0;
	p.row_uniform = (0 == dims[0] % (uint64_t) span) ? 1 : 0;
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
static int plan_tiles (VmParams& p, size_t esize, int chunk, int tile0, int tilep)
{
	if (p.ndims < 2 || 0 == p.idx32)
	{
		return 0; // the tile kernels address with 32-bit offsets
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
	int nstaged = 0;
	int64_t lanes16 = (int64_t) (16 / esize);
	bool rows_aligned = (0 == p.dims[0] % (uint64_t) chunk);
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
	p.tilesp = (uint32_t) ((p.dims[1] + tilep - 1) / tilep);
	p.nstaged = nstaged;
	p.mirror = (tile0 == tilep && p.tiles0 == p.tilesp && p.tiles0 > 1) ? 1 : 0;
	bool aligned = true;
	for (int k = 0; k < p.ninputs; ++k)
	{
		const VmInput& in = p.in[k];
		if (VM_IN_TILE != in.mode)
		{
			continue;
		}
		aligned = aligned && aligned16(in.data, in.offset, esize);
		for (int d = 0; d < p.ndims; ++d)
		{
			if (1 != d)
			{
				aligned = aligned && (0 == in.stride[d] % lanes16);
			}
		}
	}
	p.tile_aligned = aligned ? 1 : 0;
	p.pair_mode = 0;
	p.tile_input = 0;
	if (aligned && 4 == esize && p.mirror && 1 == nstaged &&
		p.dims[0] == p.dims[1] && 0 == p.dims[0] % (uint64_t) tile0)
	{
		// a straight view of the same buffer whose strides mirror the staged
		// (transposed) view's: f(x, permute(x)).  vm_pair4_kernel reads it
		// from the mirror tile in shared memory.
		int t = -1;
		for (int k = 0; k < p.ninputs; ++k)
		{
			if (VM_IN_TILE == p.in[k].mode)
			{
				t = k;
				p.tile_input = k;
			}
		}
		for (int k = 0; k < p.ninputs && t >= 0; ++k)
		{
			VmInput& in = p.in[k];
			const VmInput& tr = p.in[t];
			if (VM_IN_STRIDED != in.mode || in.data != tr.data ||
				in.offset != tr.offset || 1 != in.stride[0] ||
				in.stride[1] != tr.stride[0])
			{
				continue;
			}
			bool same = true;
			for (int d = 2; d < p.ndims; ++d)
			{
				same = same && in.stride[d] == tr.stride[d];
			}
			if (same)
			{
				in.mode = VM_IN_ALIAS;
				p.pair_mode = 1;
			}
		}
	}
	return nstaged;
}

/// Launch the tile kernel of T: the cp.async / swizzled kernel for 4-byte
/// types, the padded-tile kernel for the others
template <typename T>
static int launch_tile (llo_cuda_ctx* ctx, int dtype, VmParams& p, int nstaged,
	size_t spill, unsigned blocks)
{
	p.total_tiles = blocks;
	p.sched = reinterpret_cast<uint32_t*>(ctx->dev_error + 1);
	using G = VmGeom<T>;
	if constexpr (4 == sizeof(T))
	{
		size_t smem = (size_t) (p.tile_aligned ? 2 : 1) * nstaged * VM_TS * VM_TS *
			sizeof(T) + spill;
		unsigned cap = (unsigned) ctx->sm_count * 3;
		cap += cap & 1;
		if (blocks > cap)
		{
			// persistent: an even CTA count keeps mirror pairs side by side
			blocks = cap;
		}
		// (static shared memory counts against the 48 KB default too, so the
		// opt-in is taken from 32 KB of dynamic memory on)
		if (smem > 32 * 1024)
		{
			LLO_CUDA_TRY(cudaFuncSetAttribute(vm_tile4_kernel<T,true>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
			LLO_CUDA_TRY(cudaFuncSetAttribute(vm_tile4_kernel<T,false>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
		}
		if (p.pair_mode)
		{
			// work units: tile pairs + diagonal tiles, per slice of the other dims
			uint64_t side = p.tiles0;
			uint64_t units = (uint64_t) p.total_tiles / (side * side) *
				(side * (side - 1) / 2 + (side + 1) / 2);
			p.total_tiles = (uint32_t) units;
			size_t psmem = (size_t) 4 * VM_TS * VM_TS * sizeof(T) + spill;
			LLO_CUDA_TRY(cudaFuncSetAttribute(vm_pair4_kernel<T>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
			unsigned pblocks = (unsigned) std::min<uint64_t>(units,
				(uint64_t) ctx->sm_count * 3);
			int rc = LLO_OK;
			if (spec_wanted(ctx, p.n))
			{
				if (spec_launch_pair_tma(ctx, dtype, p, units, spill, rc))
				{
					return rc;
				}
				// the specialised build has its own ring depth / occupancy
				int nbuf, ctas;
				spec_pair_shape(nbuf, ctas);
				size_t ssmem = (size_t) 2 * nbuf * VM_TS * VM_TS * sizeof(T) + spill;
				unsigned sblocks = (unsigned) std::min<uint64_t>(units,
					(uint64_t) ctx->sm_count * ctas);
				if (spec_launch(ctx, SPEC_PAIR4, dtype, p, sblocks, ssmem, rc))
				{
					return rc;
				}
			}
			vm_pair4_kernel<T><<<pblocks, VM_THREADS, psmem, ctx->stream>>>(p);
			return launched(ctx, "vm_pair4_kernel");
		}
		int rc = LLO_OK;
		if (spec_launch(ctx, p.tile_aligned ? SPEC_TILE4_ALIGNED : SPEC_TILE4_UNALIGNED,
			dtype, p, blocks, smem, rc))
		{
			return rc;
		}
		if (p.tile_aligned)
		{
			vm_tile4_kernel<T,true><<<blocks, VM_THREADS, smem, ctx->stream>>>(p);
		}
		else
		{
			vm_tile4_kernel<T,false><<<blocks, VM_THREADS, smem, ctx->stream>>>(p);
		}
		if (LLO_OK != launched(ctx, "vm_tile4_kernel"))
		{
			std::string why = llo_cuda_last_error();
			return fail(LLO_ERR_CUDA, "%s (blocks %u, smem %zu, staged %d, spill %zu)",
				why.c_str(), blocks, smem, nstaged, spill);
		}
		return LLO_OK;
	}
	else
	{
		size_t smem = (size_t) VM_MAX_TILE_INPUTS * G::TILEP * G::PITCH *
			sizeof(T) + spill;
		if (smem > 32 * 1024)
		{
			LLO_CUDA_TRY(cudaFuncSetAttribute(vm_tile_kernel<T>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		}
		int rc = LLO_OK;
		if (spec_launch(ctx, SPEC_TILE, dtype, p, blocks, smem, rc))
		{
			return rc;
		}
		vm_tile_kernel<T><<<blocks, VM_THREADS, smem, ctx->stream>>>(p);
		return launched(ctx, "vm_tile_kernel");
	}
}

int vm_launch (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs,
	const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns)
{
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		using G = VmGeom<T>;
		VmParams p;
		LLO_TRY(vm_prepare(p, dtype, outshape, inputs, ninputs,
			consts, nconsts, program, ninsns, G::E, VM_THREADS * G::V));
		p.out = out;
		p.out_vec_ok = aligned16(out, 0, sizeof(T)) ? 1 : 0;
		size_t spill = (size_t) p.spill_depth * G::V * VM_THREADS * sizeof(T);
		int nstaged = plan_tiles(p, sizeof(T), G::E, G::TILE0, G::TILEP);
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
			return launch_tile<T>(ctx, dtype, p, nstaged, spill, (unsigned) blocks);
		}
		uint64_t span = (uint64_t) VM_THREADS * G::V;
		uint64_t blocks = (p.n + span - 1) / span;
		if (blocks >= 0x7fffffffull)
		{
			return fail(LLO_ERR_FATAL, "elementwise launch too large");
		}
		if (spill > 32 * 1024)
		{
			LLO_CUDA_TRY(cudaFuncSetAttribute(vm_flat_kernel<T>,
				cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
		}
		int rc = LLO_OK;
		if (spec_launch(ctx, SPEC_FLAT, dtype, p, (unsigned) blocks, spill, rc))
		{
			return rc;
		}
		vm_flat_kernel<T><<<(unsigned) blocks, VM_THREADS, spill,
			ctx->stream>>>(p);
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
