///
/// eval.hpp
/// llo
///
/// Purpose:
/// Evaluate the root of an equation graph on the B200.  Same interface as
/// the reference (llo/eval.hpp:26-102): llo::Evaluator is an ade::iTraveler
/// with `out_`, llo::eval(tens, dtype) returns a GenericData whose `data_`
/// host bytes read as in the reference's tests.
///
/// Differences in mechanism, not in results:
///  * every sub-graph is evaluated once per eval call (the reference
///    re-evaluates shared sub-graphs per parent, llo/eval.hpp:77-87);
///    RAND_* nodes are the exception and draw fresh values per visit
///  * leaves live on the device (Variable::device_data)
///  * llo::eval_device returns the result without the device->host copy
///

#include "ade/traveler.hpp"

#include "llo/generated/opmap.hpp"

#ifndef LLO_EVAL_HPP
#define LLO_EVAL_HPP

namespace llo
{

/// Values already computed during one evaluation, keyed by node and dtype
struct EvalMemo final
{
	std::unordered_map<ade::iTensor*,GenericData>& of (
		age::_GENERATED_DTYPE dtype)
	{
		return done_[(size_t) dtype];
	}

	/// True when `tens` is or contains a RAND_* functor: such sub-graphs are
	/// re-evaluated on every visit (fresh draws, llo/eval.hpp:77-87 of the
	/// reference re-runs every child), exactly like the planner never shares
	/// them -- the two modes then sample alike
	bool has_random (ade::iTensor* tens);

private:
	std::unordered_map<ade::iTensor*,GenericData> done_[
		age::_N_GENERATED_DTYPES];

	std::unordered_map<ade::iTensor*,bool> random_;
};

/// Visitor implementation to evaluate ade nodes according to dtype
struct Evaluator final : public ade::iTraveler
{
	Evaluator (age::_GENERATED_DTYPE dtype) :
		dtype_(dtype), memo_(std::make_shared<EvalMemo>()) {}

	Evaluator (age::_GENERATED_DTYPE dtype, std::shared_ptr<EvalMemo> memo) :
		dtype_(dtype), memo_(memo) {}

	/// Implementation of iTraveler
	void visit (ade::iLeaf* leaf) override;

	/// Implementation of iTraveler
	void visit (ade::iFunctor* func) override;

	/// Output data evaluated upon visiting node (device-resident; call
	/// out_.fetch() for host bytes)
	GenericData out_;

private:
	/// Output type when evaluating data
	age::_GENERATED_DTYPE dtype_;

	std::shared_ptr<EvalMemo> memo_;
};

/// Evaluate generic data of tens converted to specified dtype; the result's
/// host mirror `data_` is filled
GenericData eval (ade::TensptrT tens, age::_GENERATED_DTYPE dtype);

/// Same, but the result stays on the device (`device_` set, `data_` null)
GenericData eval_device (ade::TensptrT tens, age::_GENERATED_DTYPE dtype);

/// Evaluate several roots of one graph together: sub-graphs they share (the
/// forward pass under the gradients of llo::derive) are computed once.
/// Results stay on the device.
std::vector<GenericData> eval_device_many (ade::TensT roots,
	age::_GENERATED_DTYPE dtype);

}

#endif // LLO_EVAL_HPP
