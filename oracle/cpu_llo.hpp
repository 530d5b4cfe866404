///
/// cpu_llo.hpp — ORACLE (test infrastructure, not product code)
///
/// A CPU restatement of the reference's LLO evaluator hot path, used ONLY as
/// the parity checker by tests/, __graft_entry__.smoke() and bench.py's
/// cpu_baseline / --impl reference legs.  Nothing under cortenn_b200/ may
/// include, link or call it.
///
/// Parity status: PINNED.  The restatement is checked against the
/// reference's own golden vectors (tests/test_oracle_golden.py restates
/// llo/test/test_operator.cpp:12-263, llo/test/test_data.cpp:13-88,
/// llo/test/test_api.cpp:153-1159).  The reference itself cannot be built
/// here (needs bazel + Tenncor@959d51e + gtest, none present).
///
/// What it follows (reference file:line):
///   convert / copyover        llo/src/data.cpp:20-72
///   unary                     llo/operator.hpp:31-61
///   binary                    llo/operator.hpp:169-231
///   nnary                     llo/operator.hpp:367-414
///   typed ops                 llo/operator.hpp:66-165, 237-363, 419-451
///                             llo/src/operator.cpp:15-133
///   op_exec dispatch          llo/cfg/llo.json:31-124 ("operation")
///   Evaluator / eval          llo/eval.hpp:26-99, llo/src/eval.cpp:8-13
/// Like the reference it is single-threaded, scalar, does one virtual
/// CoordMap::forward (9x9 double mat-vec) and one std::function call per
/// element, re-evaluates shared sub-graphs and re-converts leaves per visit.
///

#ifndef ORACLE_CPU_LLO_HPP
#define ORACLE_CPU_LLO_HPP

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <random>
#include <vector>

#include "ade/ade.hpp"
#include "llo/generated/codes.hpp"

namespace oracle
{

using DType = age::_GENERATED_DTYPE;
using OpCode = age::_GENERATED_OPCODE;

/// llo::GenericData restated with host memory (llo/data.hpp:24-42)
struct HostData final
{
	HostData (void) = default;

	HostData (ade::Shape shape, DType dtype) :
		data_((char*) std::malloc(
			std::max<size_t>(1, shape.n_elems() * age::type_size(dtype))),
			[](char* p) { std::free(p); }),
		shape_(shape), dtype_(dtype) {}

	void copyover (const char* indata, DType intype);

	std::shared_ptr<char> data_;
	ade::Shape shape_;
	DType dtype_ = age::BAD_TYPE;
};

/// llo::DataArg restated (llo/data.hpp:224-238)
struct Arg final
{
	std::shared_ptr<char> data_;
	ade::Shape shape_;
	ade::CoordptrT mapper_;
	bool push_;
};

/// llo::VecRef restated (llo/data.hpp:245-261)
template <typename T>
struct Ref final
{
	const T* data;
	ade::Shape shape;
	ade::CoordptrT mapper;
	bool push;
};

using Engine = std::default_random_engine;

/// The single global generator (llo/src/operator.cpp:8-12)
Engine& engine (void);

// ---- element-type conversion: C++ implicit conversion per element -------

template <typename OUT, typename IN>
void convert_typed (OUT* out, const IN* in, size_t n)
{
	for (size_t i = 0; i < n; ++i)
	{
		out[i] = (OUT) in[i];
	}
}

/// out (outtype) <- in (intype), n elements
void convert (char* out, DType outtype, const char* in, DType intype, size_t n);

// ---- coordinate walking --------------------------------------------------

/// source linear index -> destination linear index through mapper
inline ade::NElemT map_index (const ade::iCoordMap& mapper,
	const ade::Shape& from, ade::NElemT i, const ade::Shape& to)
{
	ade::CoordT dst;
	ade::CoordT src = ade::coordinate(from, i);
	mapper.forward(dst.begin(), src.begin());
	return ade::index(to, dst);
}

// ---- the three generic loops ----------------------------------------------

/// llo::unary (llo/operator.hpp:31-61)
template <typename T>
void unary (T* out, const ade::Shape& outshape, Ref<T> in,
	std::function<T(const T&)> f)
{
	if (ade::is_identity(in.mapper.get()))
	{
		ade::NElemT n = in.shape.n_elems();
		for (ade::NElemT i = 0; i < n; ++i)
		{
			out[i] = f(in.data[i]);
		}
		return;
	}
	if (in.push)
	{
		// scatter: later input elements overwrite earlier ones
		ade::NElemT n = in.shape.n_elems();
		for (ade::NElemT i = 0; i < n; ++i)
		{
			out[map_index(*in.mapper, in.shape, i, outshape)] = f(in.data[i]);
		}
		return;
	}
	ade::NElemT n = outshape.n_elems();
	for (ade::NElemT i = 0; i < n; ++i)
	{
		out[i] = f(in.data[map_index(*in.mapper, outshape, i, in.shape)]);
	}
}

/// llo::binary (llo/operator.hpp:169-231)
template <typename OUT, typename A, typename B>
void binary (OUT* out, const ade::Shape& outshape, Ref<A> a, Ref<B> b,
	std::function<OUT(const A&,const B&)> f)
{
	ade::NElemT nout = outshape.n_elems();
	if (!a.push && !b.push)
	{
		for (ade::NElemT i = 0; i < nout; ++i)
		{
			out[i] = f(
				a.data[map_index(*a.mapper, outshape, i, a.shape)],
				b.data[map_index(*b.mapper, outshape, i, b.shape)]);
		}
		return;
	}
	// `a` is staged in output space (value-initialised), then combined with b
	std::vector<A> staged(nout);
	if (a.push)
	{
		for (ade::NElemT i = 0, n = a.shape.n_elems(); i < n; ++i)
		{
			staged[map_index(*a.mapper, a.shape, i, outshape)] = a.data[i];
		}
	}
	else
	{
		for (ade::NElemT i = 0; i < nout; ++i)
		{
			staged[i] = a.data[map_index(*a.mapper, outshape, i, a.shape)];
		}
	}
	if (b.push)
	{
		// only outputs that b reaches are written
		for (ade::NElemT i = 0, n = b.shape.n_elems(); i < n; ++i)
		{
			ade::NElemT o = map_index(*b.mapper, b.shape, i, outshape);
			out[o] = f(staged[o], b.data[i]);
		}
	}
	else
	{
		for (ade::NElemT i = 0; i < nout; ++i)
		{
			out[i] = f(staged[i],
				b.data[map_index(*b.mapper, outshape, i, b.shape)]);
		}
	}
}

/// llo::nnary (llo/operator.hpp:367-414): first touch assigns, later touches
/// accumulate; argument order, then element order.  (The reference keeps its
/// `visited` flags in a stack VLA; a heap vector lifts the ~8M output limit.)
template <typename T>
void nnary (T* out, const ade::Shape& outshape, std::vector<Ref<T>> args,
	std::function<void(T&,const T&)> acc)
{
	ade::NElemT nout = outshape.n_elems();
	std::vector<uint8_t> touched(nout, 0);
	auto fold = [&](ade::NElemT o, const T& v)
	{
		if (touched[o])
		{
			acc(out[o], v);
		}
		else
		{
			out[o] = v;
			touched[o] = 1;
		}
	};
	for (Ref<T>& arg : args)
	{
		if (arg.push)
		{
			for (ade::NElemT i = 0, n = arg.shape.n_elems(); i < n; ++i)
			{
				fold(map_index(*arg.mapper, arg.shape, i, outshape),
					arg.data[i]);
			}
		}
		else
		{
			for (ade::NElemT i = 0; i < nout; ++i)
			{
				fold(i, arg.data[
					map_index(*arg.mapper, outshape, i, arg.shape)]);
			}
		}
	}
}

/// age::op_exec restated: dtype x opcode dispatch onto the typed operators
/// (call site llo/eval.hpp:90).  Throws std::bad_function_call for
/// neg<unsigned> / rand_normal<integer> like the reference.
void op_exec (OpCode opcode, DType dtype, char* out,
	const ade::Shape& outshape, std::vector<Arg>& args);

/// llo::unary with a caller-chosen output shape (test_operator.cpp:12-63)
void unary_exec (OpCode opcode, DType dtype, char* out,
	const ade::Shape& outshape, Arg& arg);

/// llo::Evaluator restated (llo/eval.hpp:26-99): recursive, no memoisation
struct Evaluator final : public ade::iTraveler
{
	Evaluator (DType dtype) : dtype_(dtype) {}

	void visit (ade::iLeaf* leaf) override;

	void visit (ade::iFunctor* func) override;

	HostData out_;

private:
	DType dtype_;
};

/// llo::eval restated (llo/src/eval.cpp:8-13)
HostData eval (ade::TensptrT tens, DType dtype);

}

#endif // ORACLE_CPU_LLO_HPP
