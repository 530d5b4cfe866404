///
/// grader.hpp
///
/// age::RuleSet / age::Grader: the derivative rules of the op table
/// (reference: llo/cfg/llo.json:25-30 "data", :32-123 "derivative").
/// Used by llo::derive (llo/src/zprune.cpp:109-116).
///

#ifndef CORTENN_AGE_GRADER_HPP
#define CORTENN_AGE_GRADER_HPP

#include "bwd/grader.hpp"

#include "llo/generated/codes.hpp"

namespace age
{

struct RuleSet final : public bwd::iRuleSet
{
	ade::TensptrT data (double scalar, ade::Shape shape) override;

	ade::Opcode sum_opcode (void) override
	{
		return ade::Opcode{"SUM", SUM};
	}

	ade::Opcode prod_opcode (void) override
	{
		return ade::Opcode{"PROD", PROD};
	}

	ade::TensptrT grad_rule (ade::iFunctor* fwd,
		ade::MappedTensor bwd, ade::TensT args, size_t idx) override;
};

using Grader = bwd::Grader;

}

#endif // CORTENN_AGE_GRADER_HPP
