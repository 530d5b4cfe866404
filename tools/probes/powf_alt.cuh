///
/// powf_alt.cuh — an alternative fp32 pow, measured and NOT adopted: the probe
/// (powf_probe.cu) shows it no faster than CUDA's powf (416 vs 406 G evaluations/s,
/// 1.6e-7 vs 7.5e-8 worst relative error), so the engine keeps powf and a
/// stand-alone pow stays ALU-bound (~88 issue slots per element).
///
/// The reference's pow is std::pow on float (llo/operator.hpp:237-245); the
/// results only have to agree within 1e-6 relative.  CUDA's powf is ~80
/// instructions per element, which makes a stand-alone pow instruction-bound
/// (3.7 TB/s of the 6.5 TB/s the loads and stores could run at).  For a
/// positive normal base and a finite exponent this computes
///     a^b = 2^(b * log2 a)
/// with log2 a carried as an unevaluated sum of two floats (~31 bits, so the
/// product with |b * log2 a| <= 150 keeps an absolute error far below 1e-7)
/// and one MUFU.EX2 on the reduced argument (2 ulp).  Everything else
/// (zero / negative / subnormal / infinite / NaN bases, infinite / NaN
/// exponents) takes powf, whose special-case table is std::pow's.
/// Error against double pow: see tools/probes/powf_probe.cu.
///

#ifndef LLO_CUDA_POWF_CUH
#define LLO_CUDA_POWF_CUH

#include <cuda_runtime.h>
#include <cstdint>

namespace llo_cuda
{

__device__ __forceinline__ float pow_fast (float a, float b)
{
	// a in [FLT_MIN, inf), |b| < inf; the negations route NaNs to powf
	if (!(a >= 1.17549435e-38f && a < __int_as_float(0x7f800000) &&
		fabsf(b) < __int_as_float(0x7f800000)))
	{
		return powf(a, b);
	}
	// a = 2^e * m, m in [sqrt(1/2), sqrt(2))
	int32_t bits = __float_as_int(a);
	int32_t e = (bits >> 23) - 127;
	float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
	if (m > 1.41421354f)
	{
		m *= 0.5f;
		e += 1;
	}
	// log2 m = (2 / ln 2) atanh(s), s = (m - 1) / (m + 1) = f / (2 + f)
	float f = m - 1.0f;                 // exact
	float d = 2.0f + f;
	float d_lo = f - (d - 2.0f);        // exact: d + d_lo = 2 + f
	float inv;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(d)); // error absorbed by s_lo
	float s_hi = f * inv;
	float r = __fmaf_rn(-s_hi, d, f);   // f - s_hi * d, exact
	r = __fmaf_rn(-s_hi, d_lo, r);
	float s_lo = r * inv;               // s = s_hi + s_lo to ~2^-45
	float s2 = s_hi * s_hi;
	// atanh(s) / s - 1 = s^2 / 3 + s^4 / 5 + ... (|s| <= 0.1716: s^12 / 13 < 1e-10)
	float t = __fmaf_rn(s2, 0.0909090909f, 0.111111111f);
	t = __fmaf_rn(s2, t, 0.142857143f);
	t = __fmaf_rn(s2, t, 0.2f);
	t = __fmaf_rn(s2, t, 0.333333333f);
	t = t * s2;                         // relative size <= 0.01: fp32 is enough
	const float c_hi = 2.88539004f;     // 2 / ln 2 = c_hi + c_lo
	const float c_lo = 3.85192607e-08f;
	float p_hi = c_hi * s_hi;
	float p_lo = __fmaf_rn(c_hi, s_hi, -p_hi);          // exact product error
	p_lo = __fmaf_rn(c_lo, s_hi, p_lo);
	p_lo = __fmaf_rn(c_hi, s_lo, p_lo);
	p_lo = __fmaf_rn(c_hi * s_hi, t, p_lo);
	// log2 a = e + p_hi + p_lo
	float fe = (float) e;
	float l_hi = fe + p_hi;
	float l_lo = (p_hi - (l_hi - fe)) + p_lo;           // fast two-sum (|fe| >= |p_hi| or fe == 0)
	// y = b * log2 a as y_hi + y_lo
	float y_hi = b * l_hi;
	float y_lo = __fmaf_rn(b, l_hi, -y_hi);
	y_lo = __fmaf_rn(b, l_lo, y_lo);
	if (!(fabsf(y_hi) < 252.0f))
	{
		// certain overflow / underflow (or b * l_hi overflowed): let the
		// scaling below saturate (both factors stay normal numbers)
		y_hi = y_hi > 0.0f ? 252.0f : -252.0f;
		y_lo = 0.0f;
	}
	float n = rintf(y_hi);
	float frac = (y_hi - n) + y_lo;     // |frac| <= 0.5 + tiny
	float v;
	asm("ex2.approx.f32 %0, %1;" : "=f"(v) : "f"(frac));
	// v * 2^n in two exact steps (subnormal results round once more)
	int32_t ni = (int32_t) n;
	int32_t n1 = ni >> 1;
	int32_t n2 = ni - n1;
	v = v * __int_as_float((n1 + 127) << 23);
	v = v * __int_as_float((n2 + 127) << 23);
	return v;
}

}

#endif // LLO_CUDA_POWF_CUH
