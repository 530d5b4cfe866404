#include "llo/generated/api.hpp"

#include "llo/data.hpp"
#include "llo/helper.hpp"

#ifdef CORTENN_AGE_API_HPP

namespace age
{

static inline ade::Opcode op (_GENERATED_OPCODE code)
{
	return ade::Opcode{name_op(code), (size_t) code};
}

static inline ade::TensptrT unary_node (_GENERATED_OPCODE code,
	ade::TensptrT& a)
{
	return ade::TensptrT(ade::Functor::get(op(code), {ade::identity_map(a)}));
}

static inline ade::TensptrT binary_node (_GENERATED_OPCODE code,
	ade::TensptrT& a, ade::TensptrT& b)
{
	return ade::TensptrT(ade::Functor::get(op(code),
		{ade::identity_map(a), ade::identity_map(b)}));
}

// llo/cfg/llo.json:127-197
ade::TensptrT abs (ade::TensptrT arg1) { return unary_node(ABS, arg1); }
ade::TensptrT neg (ade::TensptrT arg1) { return unary_node(NEG, arg1); }
ade::TensptrT sin (ade::TensptrT arg1) { return unary_node(SIN, arg1); }
ade::TensptrT cos (ade::TensptrT arg1) { return unary_node(COS, arg1); }
ade::TensptrT tan (ade::TensptrT arg1) { return unary_node(TAN, arg1); }
ade::TensptrT exp (ade::TensptrT arg1) { return unary_node(EXP, arg1); }
ade::TensptrT log (ade::TensptrT arg1) { return unary_node(LOG, arg1); }
ade::TensptrT sqrt (ade::TensptrT arg1) { return unary_node(SQRT, arg1); }
ade::TensptrT round (ade::TensptrT arg1) { return unary_node(ROUND, arg1); }

// llo/cfg/llo.json:199-208: flip is a SUM over a flip-mapped argument
ade::TensptrT flip (ade::TensptrT arg1, uint8_t arg2)
{
	return ade::TensptrT(ade::Functor::get(op(SUM),
		{ade::flip_map(arg1, arg2)}));
}

// llo/cfg/llo.json:210-340
ade::TensptrT pow (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(POW, arg1, arg2); }
ade::TensptrT add (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(SUM, arg1, arg2); }
ade::TensptrT sub (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(SUB, arg1, arg2); }
ade::TensptrT mul (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(PROD, arg1, arg2); }
ade::TensptrT div (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(DIV, arg1, arg2); }
ade::TensptrT eq (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(EQ, arg1, arg2); }
ade::TensptrT neq (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(NEQ, arg1, arg2); }
ade::TensptrT lt (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(LT, arg1, arg2); }
ade::TensptrT gt (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(GT, arg1, arg2); }
ade::TensptrT rand_bino (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(RAND_BINO, arg1, arg2); }
ade::TensptrT rand_unif (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(RAND_UNIF, arg1, arg2); }
ade::TensptrT rand_norm (ade::TensptrT arg1, ade::TensptrT arg2) { return binary_node(RAND_NORM, arg1, arg2); }

// llo/cfg/llo.json:342-360
ade::TensptrT n_elems (ade::TensptrT arg)
{
	return llo::get_scalar(arg->shape().n_elems(), ade::Shape(), "n_elems");
}

ade::TensptrT n_dims (ade::TensptrT arg, uint8_t rank)
{
	return llo::get_scalar(arg->shape().at(rank), ade::Shape(), "n_dims");
}

// llo/cfg/llo.json:362-392
ade::TensptrT sum (ade::TensT arg1)
{
	return ade::TensptrT(ade::Functor::get(op(SUM), ade::to_args(arg1)));
}

ade::TensptrT prod (ade::TensT arg1)
{
	return ade::TensptrT(ade::Functor::get(op(PROD), ade::to_args(arg1)));
}

ade::TensptrT min (ade::TensT arg1)
{
	return ade::TensptrT(ade::Functor::get(op(MIN), ade::to_args(arg1)));
}

ade::TensptrT max (ade::TensT arg1)
{
	return ade::TensptrT(ade::Functor::get(op(MAX), ade::to_args(arg1)));
}

// llo/cfg/llo.json:393-435
ade::TensptrT reduce_sum (ade::TensptrT arg1, uint8_t arg2) { return llo::reduce(op(SUM), arg1, arg2); }
ade::TensptrT reduce_prod (ade::TensptrT arg1, uint8_t arg2) { return llo::reduce(op(PROD), arg1, arg2); }
ade::TensptrT reduce_min (ade::TensptrT arg1, uint8_t arg2) { return llo::reduce(op(MIN), arg1, arg2); }
ade::TensptrT reduce_max (ade::TensptrT arg1, uint8_t arg2) { return llo::reduce(op(MAX), arg1, arg2); }

// llo/cfg/llo.json:437-460
ade::TensptrT permute (ade::TensptrT arg1, std::vector<uint8_t> arg2)
{
	return ade::TensptrT(ade::Functor::get(op(SUM),
		{ade::permute_map(arg1, arg2)}));
}

ade::TensptrT extend (ade::TensptrT arg1, uint8_t arg2,
	std::vector<ade::DimT> arg3)
{
	return ade::TensptrT(ade::Functor::get(op(SUM),
		{ade::extend_map(arg1, arg2, arg3)}));
}

// llo/cfg/llo.json:462-531
ade::TensptrT reduce_sum (ade::TensptrT arg1) { return reduce_sum(arg1, 0); }
ade::TensptrT reduce_prod (ade::TensptrT arg1) { return reduce_prod(arg1, 0); }
ade::TensptrT reduce_min (ade::TensptrT arg1) { return reduce_min(arg1, 0); }
ade::TensptrT reduce_max (ade::TensptrT arg1) { return reduce_max(arg1, 0); }

ade::TensptrT transpose (ade::TensptrT arg1)
{
	return permute(arg1, {1, 0});
}

ade::TensptrT reduce_mean (ade::TensptrT arg1)
{
	return div(reduce_sum(arg1), n_elems(arg1));
}

ade::TensptrT matmul (ade::TensptrT arg1, ade::TensptrT arg2)
{
	return llo::matmul(arg1, arg2);
}

ade::TensptrT convolution (ade::TensptrT arg1, ade::TensptrT arg2)
{
	return llo::convolution(arg1, arg2);
}

}

#endif
