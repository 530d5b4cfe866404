///
/// elementwise.cu — launch side of the fused elementwise engine (vm.cuh)
/// and the `llo_cuda_elementwise` entry point.
///

#include <cstring>

#include "vm.cuh"

namespace llo_cuda
{

template <typename T>
__global__ void __launch_bounds__(256, 4)
vm_kernel (const __grid_constant__ VmParams p)
{
	constexpr int VEC = VmVec<T>::value;
	uint64_t nchunks = (p.n + VEC - 1) / VEC;
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	T* out = static_cast<T*>(p.out);
	for (uint64_t chunk = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		chunk < nchunks; chunk += step)
	{
		uint64_t base = chunk * VEC;
		uint64_t left = p.n - base;
		int valid = left < (uint64_t) VEC ? (int) left : VEC;
		T acc[VEC];
		vm_run<T,VEC>(acc, p, base, valid);
		store_run<T,VEC>(out + base, acc, p.out_vec_ok, valid);
	}
}

static bool aligned16 (const void* ptr, int64_t elem_offset, size_t esize)
{
	return 0 == (((uintptr_t) ptr + (uintptr_t) (elem_offset * (int64_t) esize)) & 15u);
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
	if (ninputs < 0 || ninputs > LLO_VM_MAX_INPUTS ||
		nconsts < 0 || nconsts > LLO_VM_MAX_CONSTS ||
		ninsns <= 0 || ninsns > LLO_VM_MAX_INSNS)
	{
		return fail(LLO_ERR_FATAL, "elementwise program out of limits "
			"(%d inputs, %d consts, %d instructions)",
			ninputs, nconsts, ninsns);
	}
	// stack discipline
	int depth = 0;
	for (int pc = 0; pc < ninsns; ++pc)
	{
		int op = program[pc].op;
		int arg = program[pc].arg;
		if (LLO_VM_LOAD == op)
		{
			if (arg >= ninputs)
			{
				return fail(LLO_ERR_FATAL, "program loads input %d of %d",
					arg, ninputs);
			}
			++depth;
		}
		else if (LLO_VM_CONST == op)
		{
			if (arg >= nconsts)
			{
				return fail(LLO_ERR_FATAL, "program loads constant %d of %d",
					arg, nconsts);
			}
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
		}
		else if (is_binary_op(op) || is_nnary_op(op))
		{
			if (depth < 2)
			{
				return fail(LLO_ERR_FATAL, "program underflows its stack");
			}
			--depth;
		}
		else
		{
			return fail(LLO_ERR_FATAL, "opcode %d is not elementwise", op);
		}
		if (depth > LLO_VM_MAX_DEPTH)
		{
			return fail(LLO_ERR_FATAL, "program needs more than %d stack "
				"slots", LLO_VM_MAX_DEPTH);
		}
	}
	if (1 != depth)
	{
		return fail(LLO_ERR_FATAL, "program leaves %d values", depth);
	}

	std::memset(&p, 0, sizeof(p));
	p.n = n_elems(outshape);
	p.ninsns = ninsns;
	p.ninputs = ninputs;
	std::memcpy(p.code, program, sizeof(llo_cuda_insn) * ninsns);
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
	p.idx32 = (p.n < (1ull << 31)) ? 1 : 0;
	for (int d = 0; d < nd; ++d)
	{
		p.dims[d] = dims[d];
		if (p.idx32)
		{
			p.div[d] = make_fastdiv((uint32_t) dims[d]);
		}
	}
	p.rows_aligned = (0 == dims[0] % (uint64_t) vec) ? 1 : 0;
	int64_t lanes16 = (int64_t) (16 / esize);
	for (int k = 0; k < ninputs; ++k)
	{
		VmInput in = inputs[k];
		if (VM_IN_STRIDED == in.mode)
		{
			for (int d = 0; d < LLO_RANK; ++d)
			{
				in.stride[d] = d < nd ? strides[k][d] : 0;
			}
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
			p.need_coords = 1;
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
	return LLO_OK;
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
		uint64_t nchunks = (p.n + VEC - 1) / VEC;
		uint64_t blocks = (nchunks + 255) / 256;
		uint64_t cap = (uint64_t) ctx->sm_count * 16;
		if (blocks > cap)
		{
			blocks = cap;
		}
		vm_kernel<T><<<(unsigned) blocks, 256, 0, ctx->stream>>>(p);
		return launched(ctx, "vm_kernel");
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
	uint64_t raw[LLO_VM_MAX_CONSTS];
	size_t esize = dtype_size(dtype);
	for (int c = 0; c < nconsts; ++c)
	{
		raw[c] = 0;
		switch (dtype)
		{
			case LLO_DT_DOUBLE: { double v = consts[c]; std::memcpy(&raw[c], &v, 8); } break;
			case LLO_DT_FLOAT: { float v = (float) consts[c]; std::memcpy(&raw[c], &v, 4); } break;
			case LLO_DT_INT8: { int8_t v = (int8_t) consts[c]; std::memcpy(&raw[c], &v, 1); } break;
			case LLO_DT_UINT8: { uint8_t v = (uint8_t) consts[c]; std::memcpy(&raw[c], &v, 1); } break;
			case LLO_DT_INT16: { int16_t v = (int16_t) consts[c]; std::memcpy(&raw[c], &v, 2); } break;
			case LLO_DT_UINT16: { uint16_t v = (uint16_t) consts[c]; std::memcpy(&raw[c], &v, 2); } break;
			case LLO_DT_INT32: { int32_t v = (int32_t) consts[c]; std::memcpy(&raw[c], &v, 4); } break;
			case LLO_DT_UINT32: { uint32_t v = (uint32_t) consts[c]; std::memcpy(&raw[c], &v, 4); } break;
			case LLO_DT_INT64: { int64_t v = (int64_t) consts[c]; std::memcpy(&raw[c], &v, 8); } break;
			case LLO_DT_UINT64: { uint64_t v = (uint64_t) consts[c]; std::memcpy(&raw[c], &v, 8); } break;
			default: return fail(LLO_ERR_FATAL, "executing bad type %d", dtype);
		}
	}
	// vm_prepare reads constants as packed elements of the dtype
	unsigned char packed[LLO_VM_MAX_CONSTS * 8];
	for (int c = 0; c < nconsts; ++c)
	{
		std::memcpy(packed + c * esize, &raw[c], esize);
	}
	return vm_launch(ctx, dtype, out, outshape, vin, ninputs,
		packed, nconsts, program, ninsns);
}
