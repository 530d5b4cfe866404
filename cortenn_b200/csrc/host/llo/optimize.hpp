///
/// optimize.hpp
/// llo
///
/// Purpose:
/// Graph -> graph optimisation passes for llo graphs, built on opt::Rewriter
/// (the reference leaves this open: todo:12, and its
/// derivative builders produce exactly the patterns removed here:
/// llo/src/helper.cpp:18-33 `(prod of args) / arg`, llo/cfg/llo.json
/// derivative rules `pow(x, 2 - 1)`, `bwd * 1`).
///
/// Every pass states whether it is exact.  llo::eval evaluates a graph as it
/// is written; a caller that wants the shorter graph asks for it:
///
///     ade::TensT grads = {llo::derive(loss, w), llo::derive(loss, b)};
///     grads = llo::optimize(grads);            // optional, explicit
///     llo::eval_device_many(grads, age::FLOAT);
///
/// The result is an ordinary ade graph: it can be printed, derived again,
/// saved with pbm and evaluated by any evaluator of the reference API.
///

#ifndef LLO_OPTIMIZE_HPP
#define LLO_OPTIMIZE_HPP

#include "ade/ade.hpp"

namespace llo
{

enum OptPass : unsigned
{
	/// prod(x, 1) -> x for pull-mapped arguments.  Exact.
	OPT_ONE_MUL = 1u << 0,
	/// sum(x, 0) -> x for pull-mapped arguments.  Exact up to the sign of a
	/// zero result (-0 + 0 is +0).
	OPT_ZERO_ADD = 1u << 1,
	/// pow(x, 1) -> x.  Exact.
	OPT_POW_ONE = 1u << 2,
	/// An accumulation of one identity-mapped argument is that argument
	/// (llo/operator.hpp:367-414 with a single pull argument copies).  Exact.
	OPT_SINGLE = 1u << 3,
	/// Structurally equal functors become one node (llo::derive repeats whole
	/// backward chains per target; helper.cpp:21-23 copies the forward
	/// product).  Exact; RAND_* functors and everything above them are never
	/// shared because every visit draws again (llo/eval.hpp:77-87).
	OPT_CSE = 1u << 4,
	/// (x * y) / x -> y, the quotient the product rule builds
	/// (llo/src/helper.cpp:18-33), also through the reversed map of an
	/// extended factor.  INEXACT: the reference rounds the product and the
	/// quotient (<= 1 ulp per term of difference) and yields NaN / inf where
	/// x == 0, the rewritten graph yields y.  Floating-point evaluation only:
	/// integer products wrap and integer division truncates, so the pass is
	/// applied only under leaves that are all FLOAT or DOUBLE.
	OPT_DIV_CANCEL = 1u << 5,

	OPT_EXACT = OPT_ONE_MUL | OPT_ZERO_ADD | OPT_POW_ONE | OPT_SINGLE | OPT_CSE,
	OPT_ALL = OPT_EXACT | OPT_DIV_CANCEL,
};

struct OptReport
{
	size_t functors_before_ = 0;
	size_t functors_after_ = 0;
	size_t one_mul_ = 0;
	size_t zero_add_ = 0;
	size_t pow_one_ = 0;
	size_t single_ = 0;
	size_t shared_ = 0;
	size_t div_cancel_ = 0;
};

/// Rewrite the graphs under `roots` together (nodes the roots share stay
/// shared) with the selected passes
ade::TensT optimize (const ade::TensT& roots, unsigned passes = OPT_ALL,
	OptReport* report = nullptr);

ade::TensptrT optimize (ade::TensptrT root, unsigned passes = OPT_ALL,
	OptReport* report = nullptr);

}

#endif // LLO_OPTIMIZE_HPP
