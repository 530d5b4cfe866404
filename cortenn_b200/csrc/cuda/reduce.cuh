///
/// reduce.cuh — internal interface of the reduction kernels (K4).
///

#ifndef LLO_CUDA_REDUCE_CUH
#define LLO_CUDA_REDUCE_CUH

#include "common.cuh"
#include "mapview.hpp"

namespace llo_cuda
{

/// out[o] = fold over r of src[base(o) + sum_k r_k * fold_stride[k]]
/// where base(o) = inv.offset + sum_d o_d * inv.stride[d] over `outshape`
/// and r runs over the folded dims with fold dim 0 fastest — which is
/// ascending source index, i.e. the reference's accumulation order
/// (llo/operator.hpp:376-393).
struct ReduceDesc
{
	uint64_t outshape[LLO_RANK];
	LinearMap inv;
	int nfold;
	uint64_t fold_size[LLO_RANK];
	int64_t fold_stride[LLO_RANK];
};

/// accumulate: 1 = fold into the values already in `out`
/// (later arguments of an n-ary operator), 0 = overwrite
int reduce_launch (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	const void* src, const ReduceDesc& desc, int accumulate, int mode);

}

#endif // LLO_CUDA_REDUCE_CUH
