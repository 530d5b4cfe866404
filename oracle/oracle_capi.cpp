// ORACLE (test infrastructure) — C entry points over the CPU restatement,
// shaped like include/llo_cuda.h but taking HOST pointers, so tests can feed
// the same ctypes structures to the CUDA library and to the checker.

#include <string>

#include "cpu_llo.hpp"
#include "llo_cuda.h"

static thread_local std::string oracle_error;

template <typename F>
static int guarded (F f)
{
	try
	{
		f();
		return LLO_OK;
	}
	catch (std::bad_function_call&)
	{
		oracle_error = "bad_function_call";
		return LLO_ERR_UNSUPPORTED_TYPED_OP;
	}
	catch (std::exception& e)
	{
		oracle_error = e.what();
		return LLO_ERR_FATAL;
	}
}

static ade::Shape shape_of (const uint64_t dims[LLO_RANK])
{
	return ade::Shape(std::vector<ade::DimT>(dims, dims + LLO_RANK));
}

static oracle::Arg arg_of (const llo_cuda_arg& in)
{
	const double* m = in.coorder;
	ade::CoordptrT mapper = std::make_shared<ade::CoordMap>(
		[m](ade::MatrixT fwd)
		{
			for (int i = 0; i < ade::mat_dim; ++i)
			{
				for (int j = 0; j < ade::mat_dim; ++j)
				{
					fwd[i][j] = m[i * ade::mat_dim + j];
				}
			}
		});
	return oracle::Arg{
		std::shared_ptr<char>((char*) in.data, [](char*) {}),
		shape_of(in.shape), mapper, 0 != in.push,
	};
}

extern "C"
{

const char* oracle_last_error (void)
{
	return oracle_error.c_str();
}

int oracle_seed (uint64_t seed)
{
	oracle::engine().seed(seed);
	return LLO_OK;
}

int oracle_convert (void* out, int out_dtype,
	const void* in, int in_dtype, uint64_t n)
{
	return guarded([&]()
	{
		oracle::convert((char*) out, (oracle::DType) out_dtype,
			(const char*) in, (oracle::DType) in_dtype, n);
	});
}

int oracle_unary (int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* in)
{
	return guarded([&]()
	{
		oracle::Arg arg = arg_of(*in);
		oracle::unary_exec((oracle::OpCode) opcode, (oracle::DType) dtype,
			(char*) out, shape_of(outshape), arg);
	});
}

int oracle_op_exec (int opcode, int dtype, void* out,
	const uint64_t outshape[LLO_RANK], const llo_cuda_arg* args, int nargs)
{
	return guarded([&]()
	{
		std::vector<oracle::Arg> in;
		for (int i = 0; i < nargs; ++i)
		{
			in.push_back(arg_of(args[i]));
		}
		oracle::op_exec((oracle::OpCode) opcode, (oracle::DType) dtype,
			(char*) out, shape_of(outshape), in);
	});
}

}
