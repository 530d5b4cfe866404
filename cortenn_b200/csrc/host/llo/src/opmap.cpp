#include "llo/generated/opmap.hpp"

#ifdef CORTENN_AGE_OPMAP_HPP

namespace age
{

void to_cuda_arg (llo_cuda_arg& out, const llo::DataArg& arg)
{
	out.data = arg.data_.get();
	std::copy(arg.shape_.begin(), arg.shape_.end(), out.shape);
	arg.mapper_->access([&](const ade::MatrixT& m)
	{
		for (ade::RankT i = 0; i < ade::mat_dim; ++i)
		{
			for (ade::RankT j = 0; j < ade::mat_dim; ++j)
			{
				out.coorder[i * ade::mat_dim + j] = m[i][j];
			}
		}
	});
	out.push = arg.fwd_ ? 1 : 0;
	out.reserved = 0;
}

void op_exec (_GENERATED_OPCODE opcode, _GENERATED_DTYPE dtype,
	char* out, ade::Shape shape, llo::DataArgsT& in)
{
	std::vector<llo_cuda_arg> args(in.size());
	for (size_t i = 0, n = in.size(); i < n; ++i)
	{
		to_cuda_arg(args[i], in[i]);
	}
	uint64_t outshape[LLO_RANK];
	std::copy(shape.begin(), shape.end(), outshape);
	llo::device::check(llo_cuda_op_exec(llo::device::ctx(),
		(int) opcode, (int) dtype, out, outshape,
		args.data(), (int) args.size()));
}

}

#endif
