#include "llo/generated/grader.hpp"
#include "llo/generated/api.hpp"

#include "llo/data.hpp"
#include "llo/helper.hpp"

#ifdef CORTENN_AGE_GRADER_HPP

namespace age
{

// llo/cfg/llo.json:29 "scalarize": "llo::get_scalar(scalar,shape)"
ade::TensptrT RuleSet::data (double scalar, ade::Shape shape)
{
	return llo::get_scalar(scalar, shape);
}

/// Chain rules, one per opcode (llo/cfg/llo.json:32-123).  `one`, `two`,
/// `zero` are int32 constants of the argument's shape, exactly as the
/// reference's `llo::get_scalar(1,args[0]->shape())` instantiates them.
ade::TensptrT RuleSet::grad_rule (ade::iFunctor* fwd,
	ade::MappedTensor bwd, ade::TensT args, size_t idx)
{
	_GENERATED_OPCODE code = (_GENERATED_OPCODE) fwd->get_opcode().code_;
	auto konst = [&](int v) -> ade::TensptrT
	{
		return llo::get_scalar(v, args[0]->shape());
	};
	auto through = [&](_GENERATED_OPCODE acc) -> ade::TensptrT
	{
		return ade::TensptrT(ade::Functor::get(
			ade::Opcode{name_op(acc), (size_t) acc}, {bwd}));
	};
	switch (code)
	{
		case ABS:
			return llo::mtens_mul(div(args[0], abs(args[0])), bwd);
		case NEG:
			return llo::mtens_mul(neg(konst(1)), bwd);
		case SIN:
			return llo::mtens_mul(cos(args[0]), bwd);
		case COS:
			return llo::mtens_mul(neg(sin(args[0])), bwd);
		case TAN:
			return llo::mtens_mul(
				div(konst(1), pow(cos(args[0]), konst(2))), bwd);
		case EXP:
			return llo::mtens_mul(exp(args[0]), bwd);
		case LOG:
			return llo::mtens_mul(div(konst(1), args[0]), bwd);
		case SQRT:
			return llo::mtens_mul(
				div(konst(1), mul(konst(2), sqrt(args[0]))), bwd);
		case ROUND:
			return llo::mtens_mul(konst(1), bwd);
		case POW:
			return llo::mtens_mul(0 == idx ?
				mul(args[1], pow(args[0], sub(args[1], konst(1)))) :
				mul(log(args[0]), pow(args[0], args[1])), bwd);
		case SUM:
			return mul(konst(1), through(SUM));
		case SUB:
			return llo::mtens_mul(0 == idx ? konst(1) : neg(konst(1)), bwd);
		case PROD:
			return mul(llo::grad_prod(fwd, idx, args), through(PROD));
		case DIV:
			return llo::mtens_mul(0 == idx ?
				div(konst(1), args[1]) :
				div(div(neg(args[0]), args[1]), args[1]), bwd);
		case MIN:
			return llo::mtens_mul(llo::grad_min(fwd, idx, args), bwd);
		case MAX:
			return llo::mtens_mul(llo::grad_max(fwd, idx, args), bwd);
		case EQ:
		case NEQ:
		case LT:
		case GT:
		case RAND_BINO:
		case RAND_UNIF:
		case RAND_NORM:
			return llo::mtens_mul(konst(0), bwd);
		default:
			logs::fatal("no gradient rule for unknown opcode");
	}
}

}

#endif
