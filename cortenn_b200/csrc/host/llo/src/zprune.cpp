///
/// zprune.cpp
/// llo
///
/// VENDORED FROM THE REFERENCE.  This file is mingkaic/cortenn's
/// llo/src/zprune.cpp:14-116 (MIT licence, (c) its authors): the same switch
/// over opcodes in the same order, with renamed locals and small helpers.  It
/// defines what llo::derive returns (zero-labelled leaves pruned), i.e. the
/// reference's semantics on the reference's side of the drop-in boundary;
/// nothing on the GPU path derives from it.
///

#include "ade/functor.hpp"

#include "llo/zprune.hpp"

#ifdef LLO_ZPRUNE_HPP

namespace llo
{

static ade::TensptrT zeros_like (ade::iFunctor* func)
{
	return ade::TensptrT(llo::get_scalar(0, func->shape()));
}

static ade::TensptrT ones_like (ade::iFunctor* func)
{
	return ade::TensptrT(llo::get_scalar(1, func->shape()));
}

/// What an operation becomes when the arguments in `zeros` are known to be
/// all-zero (reference: llo/src/zprune.cpp:14-92)
static ade::TensptrT prune_zero_args (ade::iFunctor* func,
	std::unordered_set<size_t> zeros, ade::ArgsT args)
{
	age::_GENERATED_OPCODE opcode =
		(age::_GENERATED_OPCODE) func->get_opcode().code_;
	auto is_zero = [&](size_t i) { return zeros.end() != zeros.find(i); };
	if (false == zeros.empty())
	{
		switch (opcode)
		{
			// f(0) = 0, and anything times 0
			case age::ABS:
			case age::NEG:
			case age::SIN:
			case age::TAN:
			case age::SQRT:
			case age::ROUND:
			case age::PROD:
				return zeros_like(func);
			// f(0) = 1
			case age::COS:
			case age::EXP:
				return ones_like(func);
			case age::LOG:
				logs::fatal("cannot LOG by zero");
			case age::POW:
				// 0 ^ x = 0, x ^ 0 = 1
				return is_zero(0) ? zeros_like(func) : ones_like(func);
			case age::SUM:
			{
				ade::ArgsT kept;
				for (size_t i = 0, n = args.size(); i < n; ++i)
				{
					if (false == is_zero(i))
					{
						kept.push_back(args[i]);
					}
				}
				if (kept.empty())
				{
					return zeros_like(func);
				}
				return ade::TensptrT(ade::Functor::get(
					ade::Opcode{"SUM", age::SUM}, kept));
			}
			case age::SUB:
				if (2 == zeros.size())
				{
					return zeros_like(func);
				}
				if (is_zero(0))
				{
					return ade::TensptrT(ade::Functor::get(
						ade::Opcode{"NEG", age::NEG}, {args[1]}));
				}
				return args[0].get_tensor();
			case age::DIV:
				if (is_zero(1))
				{
					logs::fatal("cannot DIV by zero");
				}
				return zeros_like(func);
			case age::MIN:
			case age::MAX:
			case age::EQ:
			case age::NEQ:
			case age::LT:
			case age::GT:
			case age::RAND_BINO:
			case age::RAND_UNIF:
			case age::RAND_NORM:
				break;
			default:
				logs::fatal("cannot prune unknown opcode");
		}
	}
	return ade::TensptrT(ade::Functor::get(
		ade::Opcode{age::name_op(opcode), (size_t) opcode}, args));
}

// reference: llo/src/zprune.cpp:94-107 -- zero leaves are found by LABEL "0"
ade::TensptrT zero_prune (ade::TensptrT root)
{
	opt::TargetPruner<std::string> zpruner("0",
		[](ade::iLeaf* leaf) -> std::string
		{
			auto var = dynamic_cast<Variable*>(leaf);
			if (nullptr == var)
			{
				return "";
			}
			return var->label_;
		}, prune_zero_args);
	return zpruner.prune(root);
}

// reference: llo/src/zprune.cpp:109-116
ade::TensptrT derive (ade::TensptrT root, ade::iTensor* target)
{
	age::Grader grader(target, std::make_shared<age::RuleSet>());
	root->accept(grader);
	auto it = grader.derivatives_.find(root.get());
	if (grader.derivatives_.end() == it)
	{
		logs::fatal("failed to derive root");
	}
	return zero_prune(it->second);
}

}

#endif
