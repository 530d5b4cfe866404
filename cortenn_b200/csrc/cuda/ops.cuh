///
/// ops.cuh — per-element semantics of the typed llo operators on the device.
///
/// Every function mirrors the C++ expression the reference evaluates on the
/// host (llo/operator.hpp:66-165, 237-301, 419-451):
///   * float / double: the <cmath> overload of the same type
///   * integers: std:: math is computed in double and converted back to T;
///     arithmetic is done in the promoted type and truncated to T
///   * comparisons yield 0 / 1 in T
///   * min / max are std::min / std::max (comparison form, so NaN handling
///     matches), accumulation is `+=` / `*=`
/// The file is compiled with -fmad=false: no multiply-add contraction.
///

#ifndef LLO_CUDA_OPS_CUH
#define LLO_CUDA_OPS_CUH

#include <cstdint>
#include <type_traits>

#include "llo_cuda.h"

namespace llo_cuda
{

template <typename T>
struct is_fp : std::integral_constant<bool,
	std::is_same<T,float>::value || std::is_same<T,double>::value> {};

/// Type arithmetic happens in after integer promotion
template <typename T>
struct promoted
{
	using type = decltype(T() + T());
};

/// (T) of a double as x86-64 hosts compute it: truncation toward zero in
/// range; out of range (undefined in C++) the low bits of a wider signed
/// conversion rather than the device's saturation
template <typename T>
__device__ __forceinline__ T from_double (double v)
{
	if constexpr (is_fp<T>::value) { return (T) v; }
	else if constexpr (sizeof(T) <= 4) { return (T) (int64_t) v; }
	else if constexpr (std::is_same<T,uint64_t>::value)
	{
		return v < 0 ? (uint64_t) (int64_t) v : (uint64_t) v;
	}
	else { return (T) v; }
}

// ---- std:: math, per type ---------------------------------------------------

#define LLO_MATH1(NAME, F32, F64)\
template <typename T>\
__device__ __forceinline__ T NAME (T v)\
{\
	if constexpr (std::is_same<T,float>::value) { return F32(v); }\
	else if constexpr (std::is_same<T,double>::value) { return F64(v); }\
	else { return from_double<T>(F64((double) v)); }\
}

LLO_MATH1(m_sin, sinf, sin)
LLO_MATH1(m_cos, cosf, cos)
LLO_MATH1(m_tan, tanf, tan)
LLO_MATH1(m_exp, expf, exp)
LLO_MATH1(m_log, logf, log)
LLO_MATH1(m_sqrt, sqrtf, sqrt)
LLO_MATH1(m_round, roundf, round)

#undef LLO_MATH1

template <typename T>
__device__ __forceinline__ T m_abs (T v)
{
	if constexpr (std::is_same<T,float>::value) { return fabsf(v); }
	else if constexpr (std::is_same<T,double>::value) { return fabs(v); }
	else if constexpr (std::is_unsigned<T>::value) { return v; }
	else
	{
		using P = typename promoted<T>::type;
		P p = (P) v;
		return (T) (p < 0 ? -p : p);
	}
}

template <typename T>
__device__ __forceinline__ T m_neg (T v)
{
	using P = typename promoted<T>::type;
	return (T) (-(P) v);
}

/// std::pow(double, double) for integer-valued operands: exact whenever the
/// true result is representable (glibc's pow is exact there, CUDA's pow is
/// only within 2 ulp, which could flip the truncation back to T)
__device__ __forceinline__ double pow_integral (double a, double b)
{
	if (b >= 0 && b <= 64)
	{
		double result = 1;
		double base = a;
		unsigned e = (unsigned) b;
		while (e > 0)
		{
			if (e & 1u)
			{
				result *= base;
			}
			base *= base;
			e >>= 1;
		}
		if (fabs(result) <= 9007199254740992.0)
		{
			return result;
		}
	}
	return pow(a, b);
}

template <typename T>
__device__ __forceinline__ T m_pow (T a, T b)
{
	if constexpr (std::is_same<T,float>::value) { return powf(a, b); }
	else if constexpr (std::is_same<T,double>::value) { return pow(a, b); }
	else { return from_double<T>(pow_integral((double) a, (double) b)); }
}

template <typename T>
__device__ __forceinline__ T m_add (T a, T b)
{
	using P = typename promoted<T>::type;
	return (T) ((P) a + (P) b);
}

template <typename T>
__device__ __forceinline__ T m_sub (T a, T b)
{
	using P = typename promoted<T>::type;
	return (T) ((P) a - (P) b);
}

template <typename T>
__device__ __forceinline__ T m_mul (T a, T b)
{
	using P = typename promoted<T>::type;
	return (T) ((P) a * (P) b);
}

template <typename T>
__device__ __forceinline__ T m_div (T a, T b)
{
	using P = typename promoted<T>::type;
	if constexpr (is_fp<T>::value)
	{
		return a / b;
	}
	else
	{
		// the reference traps (SIGFPE) on integer division by zero and on
		// INT_MIN / -1; the device returns 0 / wraps there instead
		P pb = (P) b;
		if (0 == pb)
		{
			return (T) 0;
		}
		if constexpr (std::is_signed<P>::value)
		{
			if (pb == (P) -1)
			{
				return (T) (0 - (P) a);
			}
		}
		return (T) ((P) a / pb);
	}
}

template <typename T>
__device__ __forceinline__ T m_min (T acc, T v)
{
	return (v < acc) ? v : acc;
}

template <typename T>
__device__ __forceinline__ T m_max (T acc, T v)
{
	return (acc < v) ? v : acc;
}

/// One unary opcode on one element
template <typename T>
__device__ __forceinline__ T apply_unary (int op, T v)
{
	switch (op)
	{
		case LLO_OP_ABS: return m_abs(v);
		case LLO_OP_NEG: return m_neg(v);
		case LLO_OP_SIN: return m_sin(v);
		case LLO_OP_COS: return m_cos(v);
		case LLO_OP_TAN: return m_tan(v);
		case LLO_OP_EXP: return m_exp(v);
		case LLO_OP_LOG: return m_log(v);
		case LLO_OP_SQRT: return m_sqrt(v);
		case LLO_OP_ROUND: return m_round(v);
		default: return v;
	}
}

/// One binary opcode (or one step of an n-ary accumulation) on one element
template <typename T>
__device__ __forceinline__ T apply_binary (int op, T a, T b)
{
	switch (op)
	{
		case LLO_OP_POW: return m_pow(a, b);
		case LLO_OP_SUM: return m_add(a, b);
		case LLO_OP_SUB: return m_sub(a, b);
		case LLO_OP_PROD: return m_mul(a, b);
		case LLO_OP_DIV: return m_div(a, b);
		case LLO_OP_MIN: return m_min(a, b);
		case LLO_OP_MAX: return m_max(a, b);
		case LLO_OP_EQ: return (T) (a == b);
		case LLO_OP_NEQ: return (T) (a != b);
		case LLO_OP_LT: return (T) (a < b);
		case LLO_OP_GT: return (T) (a > b);
		default: return a;
	}
}

/// Accumulator step of the n-ary operators with a compile-time opcode
template <int OP, typename T>
__device__ __forceinline__ T fold_step (T acc, T v)
{
	if constexpr (LLO_OP_SUM == OP) { return m_add(acc, v); }
	else if constexpr (LLO_OP_PROD == OP) { return m_mul(acc, v); }
	else if constexpr (LLO_OP_MIN == OP) { return m_min(acc, v); }
	else { return m_max(acc, v); }
}

}

#endif // LLO_CUDA_OPS_CUH
