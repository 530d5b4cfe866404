///
/// opmap.hpp
///
/// age::op_exec — the operator plug-in point of the evaluator (call site in
/// the reference: llo/eval.hpp:90; body generated from the "operation"
/// strings of llo/cfg/llo.json:31-124).  Here the opcode x dtype switch lives
/// behind the C-ABI (llo_cuda_op_exec) and every operand is device memory.
///

#ifndef CORTENN_AGE_OPMAP_HPP
#define CORTENN_AGE_OPMAP_HPP

#include "llo/data.hpp"

#include "llo/generated/codes.hpp"

namespace age
{

/// Fill a C-ABI argument descriptor from a DataArg
void to_cuda_arg (llo_cuda_arg& out, const llo::DataArg& arg);

void op_exec (_GENERATED_OPCODE opcode, _GENERATED_DTYPE dtype,
	char* out, ade::Shape shape, llo::DataArgsT& in);

}

#endif // CORTENN_AGE_OPMAP_HPP
