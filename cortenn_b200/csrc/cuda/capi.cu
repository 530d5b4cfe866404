///
/// capi.cu — the operator entry points of the C-ABI: llo_cuda_unary /
/// binary / nnary / op_exec.  Each call analyses its arguments' coordinate
/// maps once (mapview.hpp), then launches the elementwise engine, a
/// reduction, or the general-map kernels.
///

#include <cstring>

#include "common.cuh"
#include "general.cuh"
#include "mapview.hpp"
#include "reduce.cuh"
#include "vm.cuh"

namespace llo_cuda
{

static int bad_coordinate (const char* what)
{
	return fail(LLO_ERR_FATAL, "cannot get index of bad coordinate: the %s "
		"coordinate map leaves the tensor's shape", what);
}

static void strided_input (VmInput& in, const void* data, const LinearMap& lm)
{
	std::memset(&in, 0, sizeof(VmInput));
	in.data = data;
	in.offset = lm.offset;
	for (int d = 0; d < LLO_RANK; ++d)
	{
		in.stride[d] = lm.stride[d];
	}
	in.mode = VM_IN_STRIDED;
}

/// Express argument `arg` as an input of an elementwise program over
/// `outshape`.  Pull maps become strided views or index tables; push maps
/// only when they are a bijection (then the inverse is a view).  pull_ok is
/// cleared, and `in` left untouched, for push maps that need the scatter
/// path.
int make_vm_input (llo_cuda_ctx* ctx, Scratch& scratch, VmInput& in,
	const llo_cuda_arg& arg, const uint64_t outshape[LLO_RANK],
	bool& pull_ok)
{
	LinearMap lm;
	if (arg.push)
	{
		MapKind kind = classify_map(arg.coorder, arg.shape, outshape, lm);
		if (MAP_INVALID == kind)
		{
			return bad_coordinate("push");
		}
		if (MAP_LINEAR == kind)
		{
			PushShape ps = analyse_push(lm, arg.shape, outshape);
			if (ps.bijective)
			{
				strided_input(in, arg.data, ps.inverse);
				return LLO_OK;
			}
		}
		pull_ok = false;
		return LLO_OK;
	}
	MapKind kind = classify_map(arg.coorder, outshape, arg.shape, lm);
	if (MAP_INVALID == kind)
	{
		return bad_coordinate("pull");
	}
	if (MAP_LINEAR == kind)
	{
		strided_input(in, arg.data, lm);
		return LLO_OK;
	}
	int64_t* table = nullptr;
	LLO_TRY(scratch.get((void**) &table,
		n_elems(outshape) * sizeof(int64_t)));
	LLO_TRY(build_index_table(ctx, table, arg.coorder, outshape, arg.shape));
	std::memset(&in, 0, sizeof(VmInput));
	in.data = arg.data;
	in.index = table;
	in.mode = VM_IN_INDEXED;
	return LLO_OK;
}

static void linear_input (VmInput& in, const void* data)
{
	std::memset(&in, 0, sizeof(VmInput));
	in.data = data;
	in.mode = VM_IN_LINEAR;
}

/// Scatter-with-overwrite of a push argument into output space:
/// staged[o] = arg[last i mapping to o], 0 where nothing maps;
/// winner (n_out entries) tells which outputs were reached
static int stage_push (llo_cuda_ctx* ctx, Scratch& scratch, int dtype,
	void** staged, int64_t** winner, const llo_cuda_arg& arg,
	const uint64_t outshape[LLO_RANK])
{
	uint64_t nin = n_elems(arg.shape);
	uint64_t nout = n_elems(outshape);
	int64_t* key = nullptr;
	LLO_TRY(scratch.get((void**) &key, nin * sizeof(int64_t)));
	LLO_TRY(build_index_table(ctx, key, arg.coorder, arg.shape, outshape));
	LLO_TRY(scratch.get((void**) winner, nout * sizeof(int64_t)));
	LLO_TRY(scatter_winner(ctx, *winner, nout, key, nin));
	LLO_TRY(scratch.get(staged, nout * dtype_size(dtype)));
	return gather_winner(ctx, dtype, *staged, arg.data, *winner, nout);
}

}

using namespace llo_cuda;

extern "C"
{

int llo_cuda_unary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* in)
{
	if (false == is_unary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "%s is not a unary operator",
			op_name(opcode));
	}
	size_t esize = dtype_size(dtype);
	if (0 == esize)
	{
		return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
	}
	if (is_unsigned_dtype(dtype))
	{
		if (LLO_OP_NEG == opcode)
		{
			// llo/src/operator.cpp:39-60
			return fail(LLO_ERR_UNSUPPORTED_TYPED_OP,
				"NEG is not defined for %s", dtype_name(dtype));
		}
		if (LLO_OP_ABS == opcode)
		{
			// llo/src/operator.cpp:15-37: plain copy, mapper ignored
			return llo_cuda_d2d(ctx, out, in->data,
				n_elems(in->shape) * esize);
		}
	}
	llo_cuda_insn prog[2] = {{LLO_VM_LOAD, 0}, {(uint8_t) opcode, 0}};
	VmInput vin;
	if (is_identity_matrix(in->coorder))
	{
		// llo/operator.hpp:34-40: walks the argument's own elements
		linear_input(vin, in->data);
		return vm_launch(ctx, dtype, out, in->shape, &vin, 1,
			nullptr, 0, prog, 2);
	}
	Scratch scratch(ctx);
	bool pull_ok = true;
	LLO_TRY(make_vm_input(ctx, scratch, vin, *in, outshape, pull_ok));
	if (pull_ok)
	{
		return vm_launch(ctx, dtype, out, outshape, &vin, 1,
			nullptr, 0, prog, 2);
	}
	// llo/operator.hpp:41-50: scatter, later elements overwrite
	void* staged = nullptr;
	int64_t* winner = nullptr;
	LLO_TRY(stage_push(ctx, scratch, dtype, &staged, &winner, *in, outshape));
	linear_input(vin, staged);
	LLO_TRY(vm_launch(ctx, dtype, out, outshape, &vin, 1, nullptr, 0,
		prog, 2));
	return zero_unreached(ctx, dtype, out, winner, n_elems(outshape));
}

int llo_cuda_binary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* a, const llo_cuda_arg* b)
{
	if (is_rand_op(opcode))
	{
		return llo_cuda_rand(ctx, opcode, dtype, out, outshape, a, b);
	}
	if (false == is_binary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "%s is not a binary operator",
			op_name(opcode));
	}
	Scratch scratch(ctx);
	llo_cuda_insn prog[3] = {{LLO_VM_LOAD, 0}, {LLO_VM_LOAD, 1},
		{(uint8_t) opcode, 0}};
	VmInput vin[2];
	bool a_ok = true;
	bool b_ok = true;
	LLO_TRY(make_vm_input(ctx, scratch, vin[0], *a, outshape, a_ok));
	LLO_TRY(make_vm_input(ctx, scratch, vin[1], *b, outshape, b_ok));
	int64_t* b_winner = nullptr;
	if (false == a_ok)
	{
		// llo/operator.hpp:193-201: a is staged in output space
		void* staged = nullptr;
		int64_t* winner = nullptr;
		LLO_TRY(stage_push(ctx, scratch, dtype, &staged, &winner, *a,
			outshape));
		linear_input(vin[0], staged);
	}
	if (false == b_ok)
	{
		// llo/operator.hpp:211-220: only outputs b reaches are written,
		// the last b element to reach an output wins
		void* staged = nullptr;
		LLO_TRY(stage_push(ctx, scratch, dtype, &staged, &b_winner, *b,
			outshape));
		linear_input(vin[1], staged);
	}
	LLO_TRY(vm_launch(ctx, dtype, out, outshape, vin, 2, nullptr, 0,
		prog, 3));
	if (nullptr != b_winner)
	{
		return zero_unreached(ctx, dtype, out, b_winner, n_elems(outshape));
	}
	return LLO_OK;
}

int llo_cuda_nnary (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* args, int nargs)
{
	if (false == is_nnary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "%s is not an n-ary operator",
			op_name(opcode));
	}
	size_t esize = dtype_size(dtype);
	if (0 == esize)
	{
		return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
	}
	if (nargs <= 0)
	{
		return fail(LLO_ERR_FATAL, "cannot perform %s with no arguments",
			op_name(opcode));
	}
	uint64_t nout = n_elems(outshape);
	Scratch scratch(ctx);

	// classify every argument first
	enum { AS_INPUT, AS_REDUCE, AS_GENERAL };
	std::vector<int> how(nargs);
	std::vector<VmInput> vins(nargs);
	std::vector<ReduceDesc> reds(nargs);
	bool any_general = false;
	for (int i = 0; i < nargs; ++i)
	{
		const llo_cuda_arg& arg = args[i];
		LinearMap lm;
		if (arg.push)
		{
			MapKind kind = classify_map(arg.coorder, arg.shape, outshape, lm);
			if (MAP_INVALID == kind)
			{
				return bad_coordinate("push");
			}
			how[i] = AS_GENERAL;
			if (MAP_LINEAR == kind)
			{
				PushShape ps = analyse_push(lm, arg.shape, outshape);
				if (ps.bijective)
				{
					strided_input(vins[i], arg.data, ps.inverse);
					how[i] = AS_INPUT;
				}
				else if (ps.clean)
				{
					ReduceDesc& rd = reds[i];
					std::memcpy(rd.outshape, outshape, sizeof(rd.outshape));
					rd.inv = ps.inverse;
					rd.nfold = ps.nfold;
					for (int k = 0; k < LLO_RANK; ++k)
					{
						rd.fold_size[k] = k < ps.nfold ? ps.fold_size[k] : 1;
						rd.fold_stride[k] = k < ps.nfold ? ps.fold_stride[k] : 0;
					}
					how[i] = AS_REDUCE;
				}
			}
		}
		else
		{
			MapKind kind = classify_map(arg.coorder, outshape, arg.shape, lm);
			if (MAP_INVALID == kind)
			{
				return bad_coordinate("pull");
			}
			if (MAP_LINEAR == kind)
			{
				strided_input(vins[i], arg.data, lm);
				how[i] = AS_INPUT;
			}
			else
			{
				how[i] = AS_GENERAL;
			}
		}
		any_general = any_general || (AS_GENERAL == how[i]);
	}

	if (any_general)
	{
		// first-touch bookkeeping made explicit (llo/operator.hpp:371-411)
		uint8_t* touched = nullptr;
		LLO_TRY(scratch.get((void**) &touched, nout));
		LLO_CUDA_TRY(cudaMemsetAsync(touched, 0, nout, ctx->stream));
		LLO_CUDA_TRY(cudaMemsetAsync(out, 0, nout * esize, ctx->stream));
		for (int i = 0; i < nargs; ++i)
		{
			const llo_cuda_arg& arg = args[i];
			int64_t* table = nullptr;
			if (arg.push)
			{
				uint64_t nin = n_elems(arg.shape);
				LLO_TRY(scratch.get((void**) &table, nin * sizeof(int64_t)));
				LLO_TRY(build_index_table(ctx, table, arg.coorder,
					arg.shape, outshape));
				LLO_TRY(ordered_fold(ctx, opcode, dtype, out, touched,
					arg.data, table, nin, nout));
			}
			else
			{
				LLO_TRY(scratch.get((void**) &table, nout * sizeof(int64_t)));
				LLO_TRY(build_index_table(ctx, table, arg.coorder,
					outshape, arg.shape));
				LLO_TRY(masked_pull(ctx, opcode, dtype, out, touched,
					arg.data, table, nout));
			}
		}
		return LLO_OK;
	}

	// every argument reaches every output: fold runs of view arguments with
	// one elementwise program each, reductions with the reduction kernels
	bool have_out = false;
	int i = 0;
	while (i < nargs)
	{
		if (AS_REDUCE == how[i])
		{
			LLO_TRY(reduce_launch(ctx, opcode, dtype, out, args[i].data,
				reds[i], have_out ? 1 : 0, ctx->reduce_mode));
			have_out = true;
			++i;
			continue;
		}
		VmInput vin[LLO_VM_MAX_INPUTS];
		llo_cuda_insn prog[2 * LLO_VM_MAX_INPUTS];
		int nin = 0;
		int np = 0;
		void* prev = nullptr;
		if (have_out)
		{
			// the running result re-enters as input 0 through a copy, so
			// the kernel never reads the buffer it writes
			LLO_TRY(scratch.get(&prev, nout * esize));
			LLO_TRY(llo_cuda_d2d(ctx, prev, out, nout * esize));
			linear_input(vin[0], prev);
			prog[np++] = {LLO_VM_LOAD, 0};
			nin = 1;
		}
		while (i < nargs && AS_INPUT == how[i] && nin < LLO_VM_MAX_INPUTS)
		{
			vin[nin] = vins[i];
			prog[np++] = {LLO_VM_LOAD, (uint8_t) nin};
			if (nin > 0)
			{
				prog[np++] = {(uint8_t) opcode, 0};
			}
			++nin;
			++i;
		}
		LLO_TRY(vm_launch(ctx, dtype, out, outshape, vin, nin, nullptr, 0,
			prog, np));
		have_out = true;
	}
	return LLO_OK;
}

int llo_cuda_op_exec (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* args, int nargs)
{
	if (is_unary_op(opcode))
	{
		if (nargs < 1)
		{
			return fail(LLO_ERR_FATAL, "cannot %s without an argument",
				op_name(opcode));
		}
		// the reference's typed wrappers pass the ARGUMENT's shape as the
		// output shape (llo/operator.hpp:66-69 and siblings)
		return llo_cuda_unary(ctx, opcode, dtype, out, args[0].shape,
			&args[0]);
	}
	if (is_binary_op(opcode) || is_rand_op(opcode))
	{
		if (nargs < 2)
		{
			return fail(LLO_ERR_FATAL, "cannot %s without exactly 2 "
				"arguments: using %d arguments", op_name(opcode), nargs);
		}
		return llo_cuda_binary(ctx, opcode, dtype, out, outshape,
			&args[0], &args[1]);
	}
	if (is_nnary_op(opcode))
	{
		return llo_cuda_nnary(ctx, opcode, dtype, out, outshape, args, nargs);
	}
	return fail(LLO_ERR_FATAL, "unknown opcode %d", opcode);
}

}
