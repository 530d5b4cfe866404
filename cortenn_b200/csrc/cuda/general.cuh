///
/// general.cuh — internal interface of the general-map path (general.cu):
/// coordinate maps that do not reduce to integer strides (fractional scales
/// onto real dimensions, partial wraps, off-diagonal mixes) are evaluated
/// per element in double on the device, exactly as the reference does on
/// the host (coordinate -> forward -> index), into an index table.
///

#ifndef LLO_CUDA_GENERAL_CUH
#define LLO_CUDA_GENERAL_CUH

#include "common.cuh"

namespace llo_cuda
{

/// table[i] = index(to, M * coordinate(from, i)) for every i in the box
/// `from`; an out-of-shape coordinate raises the context's error flag
int build_index_table (llo_cuda_ctx* ctx, int64_t* table,
	const double m[LLO_MAT], const uint64_t from[LLO_RANK],
	const uint64_t to[LLO_RANK]);

/// winner[o] = the largest source position i with key[i] == o, or -1:
/// "later elements overwrite" of the reference's push loops
/// (llo/operator.hpp:41-50, 193-201, 211-220)
int scatter_winner (llo_cuda_ctx* ctx, int64_t* winner, uint64_t nout,
	const int64_t* key, uint64_t nin);

/// out[o] = src[winner[o]], or 0 where winner[o] < 0
int gather_winner (llo_cuda_ctx* ctx, int dtype, void* out,
	const void* src, const int64_t* winner, uint64_t nout);

/// out[o] = 0 where winner[o] < 0
int zero_unreached (llo_cuda_ctx* ctx, int dtype, void* out,
	const int64_t* winner, uint64_t nout);

/// One push argument of an n-ary operator with an arbitrary map
/// (llo/operator.hpp:376-393): the contributions src[i] are folded into
/// out[key[i]] in ascending i, starting from out[o] where touched[o] is set
/// and from the first contribution otherwise; touched is updated.
int ordered_fold (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	uint8_t* touched, const void* src, const int64_t* key,
	uint64_t nin, uint64_t nout);

/// One pull argument under a touched mask: out[o] = touched[o] ?
/// acc(out[o], src[table[o]]) : src[table[o]]; touched becomes all set
int masked_pull (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	uint8_t* touched, const void* src, const int64_t* table, uint64_t nout);

}

#endif // LLO_CUDA_GENERAL_CUH
