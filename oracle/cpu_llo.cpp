// ORACLE (test infrastructure) — see cpu_llo.hpp for scope and citations.

#include "cpu_llo.hpp"

namespace oracle
{

Engine& engine (void)
{
	static Engine eng;
	return eng;
}

// ---- conversion (llo/src/data.cpp:20-72) ---------------------------------

template <typename IN>
static void convert_from (char* out, DType outtype, const IN* in, size_t n)
{
	switch (outtype)
	{
#define AGE_X(NAME, CTYPE) case age::NAME:\
		convert_typed<CTYPE,IN>((CTYPE*) out, in, n); break;
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default:
			logs::fatalf("invalid output type %s",
				age::name_type(outtype).c_str());
	}
}

void convert (char* out, DType outtype, const char* in, DType intype, size_t n)
{
	switch (intype)
	{
#define AGE_X(NAME, CTYPE) case age::NAME:\
		convert_from<CTYPE>(out, outtype, (const CTYPE*) in, n); break;
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default:
			logs::fatalf("invalid input type %s",
				age::name_type(intype).c_str());
	}
}

void HostData::copyover (const char* indata, DType intype)
{
	// the reference memcpy's same-typed data and then converts anyway
	// (data.cpp:52-55); the conversion alone gives the same bytes
	convert(data_.get(), dtype_, indata, intype, shape_.n_elems());
}

// ---- typed operators -------------------------------------------------------

template <typename T>
static Ref<T> ref_of (Arg& arg)
{
	return Ref<T>{(const T*) arg.data_.get(), arg.shape_, arg.mapper_,
		arg.push_};
}

template <typename T>
static std::vector<Ref<T>> refs_of (std::vector<Arg>& args)
{
	std::vector<Ref<T>> out;
	for (Arg& arg : args)
	{
		out.push_back(ref_of<T>(arg));
	}
	return out;
}

template <typename T>
struct is_unsigned_int : std::integral_constant<bool,
	std::is_integral<T>::value && std::is_unsigned<T>::value> {};

// unary operators (llo/operator.hpp:66-165): the output shape is the
// argument's own shape, whatever the mapper
template <typename T>
static void op_abs (T* out, Ref<T> in)
{
	if constexpr (is_unsigned_int<T>::value)
	{
		// llo/src/operator.cpp:15-37: plain copy, mapper ignored
		std::memcpy(out, in.data, sizeof(T) * in.shape.n_elems());
	}
	else
	{
		unary<T>(out, in.shape, in,
			[](const T& v) -> T { return std::abs(v); });
	}
}

template <typename T>
static void op_neg (T* out, Ref<T> in)
{
	if constexpr (is_unsigned_int<T>::value)
	{
		// llo/src/operator.cpp:39-60
		throw std::bad_function_call();
	}
	else
	{
		unary<T>(out, in.shape, in, [](const T& v) -> T { return -v; });
	}
}

// std:: math on T: float stays float, integers go through double and are
// converted back on return (llo/operator.hpp:114-165)
#define ORACLE_UNARY_MATH(NAME, FN)\
template <typename T>\
static void NAME (T* out, Ref<T> in)\
{\
	unary<T>(out, in.shape, in, [](const T& v) -> T { return FN(v); });\
}

ORACLE_UNARY_MATH(op_sin, std::sin)
ORACLE_UNARY_MATH(op_cos, std::cos)
ORACLE_UNARY_MATH(op_tan, std::tan)
ORACLE_UNARY_MATH(op_exp, std::exp)
ORACLE_UNARY_MATH(op_log, std::log)
ORACLE_UNARY_MATH(op_sqrt, std::sqrt)
ORACLE_UNARY_MATH(op_round, std::round)

#undef ORACLE_UNARY_MATH

// binary operators (llo/operator.hpp:237-301)
#define ORACLE_BINARY(NAME, EXPR)\
template <typename T>\
static void NAME (T* out, const ade::Shape& shape, Ref<T> a, Ref<T> b)\
{\
	binary<T,T,T>(out, shape, a, b,\
		[](const T& a, const T& b) -> T { return EXPR; });\
}

ORACLE_BINARY(op_pow, std::pow(a, b))
ORACLE_BINARY(op_sub, a - b)
ORACLE_BINARY(op_div, a / b)
ORACLE_BINARY(op_eq, a == b)
ORACLE_BINARY(op_neq, a != b)
ORACLE_BINARY(op_lt, a < b)
ORACLE_BINARY(op_gt, a > b)

#undef ORACLE_BINARY

// random operators (llo/operator.hpp:307-363, llo/src/operator.cpp:63-133)
template <typename T>
static void op_rand_binom (T* out, const ade::Shape& shape,
	Ref<T> a, Ref<double> b)
{
	// double draws int64 trials, float int32, integers their own type
	using TrialT = typename std::conditional<std::is_same<T,double>::value,
		int64_t, typename std::conditional<std::is_same<T,float>::value,
		int32_t, T>::type>::type;
	binary<T,T,double>(out, shape, a, b,
		[](const T& a, const double& b) -> T
		{
			std::binomial_distribution<TrialT> dist(a, b);
			return dist(engine());
		});
}

template <typename T>
static void op_rand_uniform (T* out, const ade::Shape& shape,
	Ref<T> a, Ref<T> b)
{
	if constexpr (std::is_floating_point<T>::value)
	{
		binary<T,T,T>(out, shape, a, b,
			[](const T& a, const T& b) -> T
			{
				std::uniform_real_distribution<T> dist(a, b);
				return dist(engine());
			});
	}
	else
	{
		// uniform_int_distribution is undefined for 8-bit types in the
		// standard; libstdc++ accepts them, a wider type draws the same range
		using WideT = typename std::conditional<sizeof(T) == 1,
			typename std::conditional<std::is_signed<T>::value,
				int16_t,uint16_t>::type, T>::type;
		binary<T,T,T>(out, shape, a, b,
			[](const T& a, const T& b) -> T
			{
				std::uniform_int_distribution<WideT> dist(a, b);
				return dist(engine());
			});
	}
}

template <typename T>
static void op_rand_normal (T* out, const ade::Shape& shape,
	Ref<T> a, Ref<T> b)
{
	if constexpr (std::is_floating_point<T>::value)
	{
		binary<T,T,T>(out, shape, a, b,
			[](const T& a, const T& b) -> T
			{
				std::normal_distribution<T> dist(a, b);
				return dist(engine());
			});
	}
	else
	{
		// llo/operator.hpp:352-355
		throw std::bad_function_call();
	}
}

// n-ary accumulators (llo/operator.hpp:419-451)
template <typename T>
static void op_add (T* out, const ade::Shape& shape, std::vector<Ref<T>> args)
{
	nnary<T>(out, shape, args, [](T& acc, const T& v) { acc += v; });
}

template <typename T>
static void op_mul (T* out, const ade::Shape& shape, std::vector<Ref<T>> args)
{
	nnary<T>(out, shape, args, [](T& acc, const T& v) { acc *= v; });
}

template <typename T>
static void op_min (T* out, const ade::Shape& shape, std::vector<Ref<T>> args)
{
	nnary<T>(out, shape, args,
		[](T& acc, const T& v) { acc = std::min(acc, v); });
}

template <typename T>
static void op_max (T* out, const ade::Shape& shape, std::vector<Ref<T>> args)
{
	nnary<T>(out, shape, args,
		[](T& acc, const T& v) { acc = std::max(acc, v); });
}

// ---- dispatch (llo/cfg/llo.json:31-124 "operation") -------------------------

static void need_args (OpCode opcode, std::vector<Arg>& args, size_t n)
{
	if (args.size() < n)
	{
		logs::fatalf("cannot %s without at least %d arguments: "
			"using %d arguments", age::name_op(opcode).c_str(),
			(int) n, (int) args.size());
	}
}

template <typename T>
static void exec_typed (OpCode opcode, char* rawout,
	const ade::Shape& shape, std::vector<Arg>& in)
{
	T* out = (T*) rawout;
	switch (opcode)
	{
		case age::ABS: need_args(opcode, in, 1); op_abs<T>(out, ref_of<T>(in[0])); break;
		case age::NEG: need_args(opcode, in, 1); op_neg<T>(out, ref_of<T>(in[0])); break;
		case age::SIN: need_args(opcode, in, 1); op_sin<T>(out, ref_of<T>(in[0])); break;
		case age::COS: need_args(opcode, in, 1); op_cos<T>(out, ref_of<T>(in[0])); break;
		case age::TAN: need_args(opcode, in, 1); op_tan<T>(out, ref_of<T>(in[0])); break;
		case age::EXP: need_args(opcode, in, 1); op_exp<T>(out, ref_of<T>(in[0])); break;
		case age::LOG: need_args(opcode, in, 1); op_log<T>(out, ref_of<T>(in[0])); break;
		case age::SQRT: need_args(opcode, in, 1); op_sqrt<T>(out, ref_of<T>(in[0])); break;
		case age::ROUND: need_args(opcode, in, 1); op_round<T>(out, ref_of<T>(in[0])); break;
		case age::POW: need_args(opcode, in, 2); op_pow<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::SUB: need_args(opcode, in, 2); op_sub<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::DIV: need_args(opcode, in, 2); op_div<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::EQ: need_args(opcode, in, 2); op_eq<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::NEQ: need_args(opcode, in, 2); op_neq<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::LT: need_args(opcode, in, 2); op_lt<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::GT: need_args(opcode, in, 2); op_gt<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::RAND_BINO: need_args(opcode, in, 2); op_rand_binom<T>(out, shape, ref_of<T>(in[0]), ref_of<double>(in[1])); break;
		case age::RAND_UNIF: need_args(opcode, in, 2); op_rand_uniform<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::RAND_NORM: need_args(opcode, in, 2); op_rand_normal<T>(out, shape, ref_of<T>(in[0]), ref_of<T>(in[1])); break;
		case age::SUM: op_add<T>(out, shape, refs_of<T>(in)); break;
		case age::PROD: op_mul<T>(out, shape, refs_of<T>(in)); break;
		case age::MIN: op_min<T>(out, shape, refs_of<T>(in)); break;
		case age::MAX: op_max<T>(out, shape, refs_of<T>(in)); break;
		default:
			logs::fatal("unknown opcode");
	}
}

void op_exec (OpCode opcode, DType dtype, char* out,
	const ade::Shape& outshape, std::vector<Arg>& args)
{
	switch (dtype)
	{
#define AGE_X(NAME, CTYPE) case age::NAME:\
		exec_typed<CTYPE>(opcode, out, outshape, args); break;
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default:
			logs::fatal("executing bad type");
	}
}

// llo::unary with an explicit output shape, as llo/test/test_operator.cpp:12-63
// drives it (the typed wrappers above always pass the argument's shape)
template <typename T>
static void unary_typed (OpCode opcode, T* out, const ade::Shape& outshape,
	Ref<T> in)
{
	std::function<T(const T&)> f;
	switch (opcode)
	{
		case age::ABS:
			if constexpr (is_unsigned_int<T>::value)
			{
				f = [](const T& v) -> T { return v; };
			}
			else
			{
				f = [](const T& v) -> T { return std::abs(v); };
			}
			break;
		case age::NEG:
			if constexpr (is_unsigned_int<T>::value)
			{
				throw std::bad_function_call();
			}
			else
			{
				f = [](const T& v) -> T { return -v; };
			}
			break;
		case age::SIN: f = [](const T& v) -> T { return std::sin(v); }; break;
		case age::COS: f = [](const T& v) -> T { return std::cos(v); }; break;
		case age::TAN: f = [](const T& v) -> T { return std::tan(v); }; break;
		case age::EXP: f = [](const T& v) -> T { return std::exp(v); }; break;
		case age::LOG: f = [](const T& v) -> T { return std::log(v); }; break;
		case age::SQRT: f = [](const T& v) -> T { return std::sqrt(v); }; break;
		case age::ROUND: f = [](const T& v) -> T { return std::round(v); }; break;
		default: logs::fatal("not a unary opcode");
	}
	unary<T>(out, outshape, in, f);
}

void unary_exec (OpCode opcode, DType dtype, char* out,
	const ade::Shape& outshape, Arg& arg)
{
	switch (dtype)
	{
#define AGE_X(NAME, CTYPE) case age::NAME:\
		unary_typed<CTYPE>(opcode, (CTYPE*) out, outshape,\
			ref_of<CTYPE>(arg)); break;
		AGE_DTYPES(AGE_X)
#undef AGE_X
		default:
			logs::fatal("executing bad type");
	}
}

// ---- evaluator (llo/eval.hpp:26-99) ------------------------------------------

void Evaluator::visit (ade::iLeaf* leaf)
{
	const char* data = (const char*) ((const ade::iLeaf*) leaf)->data();
	DType leaftype = (DType) leaf->type_code();
	out_ = HostData(leaf->shape(), dtype_);
	out_.copyover(data, leaftype);
}

void Evaluator::visit (ade::iFunctor* func)
{
	OpCode opcode = (OpCode) func->get_opcode().code_;
	out_ = HostData(func->shape(), dtype_);

	const ade::ArgsT& children = func->get_children();
	size_t nargs = children.size();
	std::vector<Arg> args(nargs);
	if (age::RAND_BINO == opcode && 2 != nargs)
	{
		logs::fatalf("cannot RAND_BINO without exactly 2 arguments: "
			"using %d arguments", (int) nargs);
	}
	for (size_t i = 0; i < nargs; ++i)
	{
		// the probability argument of RAND_BINO is always evaluated as
		// DOUBLE (eval.hpp:50-74); every child gets a fresh evaluator, so
		// shared sub-graphs are evaluated once per use
		DType childtype =
			(age::RAND_BINO == opcode && 1 == i) ? age::DOUBLE : dtype_;
		Evaluator sub(childtype);
		children[i].get_tensor()->accept(sub);
		args[i] = Arg{
			sub.out_.data_,
			sub.out_.shape_,
			children[i].get_coorder(),
			children[i].map_io(),
		};
	}
	op_exec(opcode, out_.dtype_, out_.data_.get(), out_.shape_, args);
}

HostData eval (ade::TensptrT tens, DType dtype)
{
	Evaluator ev(dtype);
	tens->accept(ev);
	return ev.out_;
}

}
