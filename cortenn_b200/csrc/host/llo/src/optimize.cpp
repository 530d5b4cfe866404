#include "opt/rewrite.hpp"

#include "llo/generated/codes.hpp"
#include "llo/data.hpp"
#include "llo/optimize.hpp"

#ifdef LLO_OPTIMIZE_HPP

namespace llo
{

namespace
{

using ArgT = ade::MappedTensor;

static ade::iFunctor* as_functor (const ade::TensptrT& tens, size_t opcode)
{
	auto func = dynamic_cast<ade::iFunctor*>(tens.get());
	if (nullptr != func && opcode == func->get_opcode().code_)
	{
		return func;
	}
	return nullptr;
}

/// True when the argument's tensor is a leaf known to hold `value` everywhere
static bool is_const (const ArgT& arg, double value)
{
	auto var = dynamic_cast<Variable*>(arg.get_tensor().get());
	return nullptr != var && var->is_filled() && value == var->fill_value();
}

static bool all_pull (const ade::ArgsT& args)
{
	return std::all_of(args.begin(), args.end(),
		[](const ArgT& arg) { return false == arg.map_io(); });
}

/// The argument reads its tensor element for element
static bool is_plain (const ArgT& arg, const ade::Shape& outshape)
{
	return false == arg.map_io() &&
		ade::is_identity(arg.get_coorder().get()) &&
		ade::is_identity(arg.get_shaper().get()) &&
		arg.get_tensor()->shape().compatible_after(outshape, 0);
}

static ade::TensptrT constant (double value, const ade::Shape& shape)
{
	// labels "0" / "1" keep llo::zero_prune working on the result
	std::string label = 0 == value ? "0" : (1 == value ? "1" : "");
	return ade::TensptrT(get_scalar<double>(value, shape, label));
}

static ade::TensptrT nnary_of (ade::Opcode opcode, const ade::ArgsT& args,
	const ade::Shape& outshape, double neutral)
{
	if (args.empty())
	{
		return constant(neutral, outshape);
	}
	if (1 == args.size() && is_plain(args[0], outshape))
	{
		return args[0].get_tensor();
	}
	return ade::TensptrT(ade::Functor::get(opcode, args));
}

struct Passes
{
	Passes (unsigned passes, OptReport& report) :
		passes_(passes), report_(&report) {}

	bool on (OptPass pass) const { return 0 != (passes_ & pass); }

	/// prod(x, 1) -> x, sum(x, 0) -> x
	ade::TensptrT neutral (ade::iFunctor* func, const ade::ArgsT& args)
	{
		size_t code = func->get_opcode().code_;
		bool prod = age::PROD == code && on(OPT_ONE_MUL);
		bool sum = age::SUM == code && on(OPT_ZERO_ADD);
		if ((false == prod && false == sum) || args.size() < 2 ||
			false == all_pull(args))
		{
			return nullptr;
		}
		double neutral = prod ? 1 : 0;
		ade::ArgsT kept;
		for (const ArgT& arg : args)
		{
			if (false == is_const(arg, neutral))
			{
				kept.push_back(arg);
			}
		}
		if (kept.size() == args.size())
		{
			return nullptr;
		}
		(prod ? report_->one_mul_ : report_->zero_add_) += args.size() - kept.size();
		return nnary_of(func->get_opcode(), kept, func->shape(), neutral);
	}

	/// pow(x, 1) -> x
	ade::TensptrT pow_one (ade::iFunctor* func, const ade::ArgsT& args)
	{
		if (false == on(OPT_POW_ONE) || age::POW != func->get_opcode().code_ ||
			2 != args.size() || false == is_plain(args[0], func->shape()) ||
			args[1].map_io() || false == is_const(args[1], 1))
		{
			return nullptr;
		}
		++report_->pow_one_;
		return args[0].get_tensor();
	}

	/// sum{x} / prod{x} / min{x} / max{x} over one plain argument -> x
	ade::TensptrT single (ade::iFunctor* func, const ade::ArgsT& args)
	{
		size_t code = func->get_opcode().code_;
		if (false == on(OPT_SINGLE) || 1 != args.size() ||
			(age::SUM != code && age::PROD != code &&
			age::MIN != code && age::MAX != code) ||
			false == is_plain(args[0], func->shape()))
		{
			return nullptr;
		}
		++report_->single_;
		return args[0].get_tensor();
	}

	/// every leaf below is FLOAT or DOUBLE (memoised)
	bool floating (ade::iTensor* tens)
	{
		auto it = floating_.find(tens);
		if (floating_.end() != it)
		{
			return it->second;
		}
		bool out = true;
		if (auto leaf = dynamic_cast<ade::iLeaf*>(tens))
		{
			out = age::FLOAT == leaf->type_code() || age::DOUBLE == leaf->type_code();
		}
		else
		{
			auto func = static_cast<ade::iFunctor*>(tens);
			for (const ArgT& child : func->get_children())
			{
				out = out && floating(child.get_tensor().get());
			}
		}
		floating_.emplace(tens, out);
		return out;
	}

	/// the product without factor k
	static ade::TensptrT product_without (ade::iFunctor* prod, size_t k)
	{
		ade::ArgsT rest;
		const ade::ArgsT& factors = prod->get_children();
		for (size_t i = 0; i < factors.size(); ++i)
		{
			if (i != k)
			{
				rest.push_back(factors[i]);
			}
		}
		return nnary_of(prod->get_opcode(), rest, prod->shape(), 1);
	}

	/// (x * y) / x -> y
	ade::TensptrT div_cancel (ade::iFunctor* func, const ade::ArgsT& args)
	{
		if (false == on(OPT_DIV_CANCEL) || age::DIV != func->get_opcode().code_ ||
			2 != args.size() || false == is_plain(args[0], func->shape()) ||
			false == is_plain(args[1], func->shape()))
		{
			return nullptr;
		}
		ade::TensptrT num = args[0].get_tensor();
		ade::iTensor* den = args[1].get_tensor().get();
		if (false == floating(func))
		{
			return nullptr;
		}
		// prod{..., x, ...} / x
		if (auto prod = as_functor(num, age::PROD))
		{
			const ade::ArgsT& factors = prod->get_children();
			if (all_pull(factors))
			{
				for (size_t k = 0; k < factors.size(); ++k)
				{
					if (factors[k].get_tensor().get() == den &&
						is_plain(factors[k], prod->shape()))
					{
						++report_->div_cancel_;
						return product_without(prod, k);
					}
				}
			}
			return nullptr;
		}
		// sum{ M^-1: prod{..., M: x, ...} } / x, what grad_prod builds for a
		// factor that enters the product through the map M
		// (llo/src/helper.cpp:25-32): every element of x divides exactly the
		// terms it was a factor of
		auto outer = as_functor(num, age::SUM);
		if (nullptr == outer || 1 != outer->get_children().size())
		{
			return nullptr;
		}
		const ArgT& lifted = outer->get_children()[0];
		auto prod = as_functor(lifted.get_tensor(), age::PROD);
		if (nullptr == prod || false == all_pull(prod->get_children()))
		{
			return nullptr;
		}
		const ade::ArgsT& factors = prod->get_children();
		for (size_t k = 0; k < factors.size(); ++k)
		{
			if (factors[k].get_tensor().get() == den &&
				lifted.map_io() == (false == factors[k].map_io()) &&
				opt::same_map(lifted.get_coorder(), factors[k].get_coorder()))
			{
				++report_->div_cancel_;
				ade::TensptrT rest = product_without(prod, k);
				ArgT relifted(rest, lifted.get_shaper(), lifted.map_io(),
					lifted.get_coorder());
				return nnary_of(outer->get_opcode(), {relifted}, func->shape(), 0);
			}
		}
		return nullptr;
	}

	unsigned passes_;

	OptReport* report_;

	std::unordered_map<ade::iTensor*,bool> floating_;
};

}

ade::TensT optimize (const ade::TensT& roots, unsigned passes, OptReport* report)
{
	OptReport local;
	OptReport& rep = nullptr != report ? *report : local;
	rep = OptReport();
	rep.functors_before_ = opt::count_functors(roots);
	Passes rules(passes, rep);
	using namespace std::placeholders;
	opt::Rewriter rewriter({
		std::bind(&Passes::neutral, &rules, _1, _2),
		std::bind(&Passes::pow_one, &rules, _1, _2),
		std::bind(&Passes::div_cancel, &rules, _1, _2),
		std::bind(&Passes::single, &rules, _1, _2),
	}, rules.on(OPT_CSE), {age::RAND_BINO, age::RAND_UNIF, age::RAND_NORM});
	ade::TensT out = rewriter.rewrite(roots);
	rep.shared_ = rewriter.shared_;
	rep.functors_after_ = opt::count_functors(out);
	return out;
}

ade::TensptrT optimize (ade::TensptrT root, unsigned passes, OptReport* report)
{
	return optimize(ade::TensT{root}, passes, report)[0];
}

}

#endif
