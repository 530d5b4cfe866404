///
/// rand.cu — K9: RAND_BINO / RAND_UNIF / RAND_NORM fills.
///
/// The reference draws every element in sequence from one global
/// std::default_random_engine (llo/src/operator.cpp:8-12, :63-133); its
/// tests pin distributions, not streams (llo/test/test_api.cpp:1016-1159:
/// mean / variance within 10 %, bounds).  Here each element owns a
/// counter-based Philox4x32-10 stream keyed by the context seed, so the fill
/// is parallel, reproducible for a given seed and independent of launch
/// geometry.  Parameters a, b are read through their argument views, so
/// extend-mapped scalar parameters (test_api.cpp:1024-1027) cost nothing.
///   uniform, floating T : a + (b - a) * u,  u in [0, 1)   (uniform_real)
///   uniform, integer T  : a + floor(u * (b - a + 1))       (uniform_int, inclusive)
///   normal, floating T  : a + b * z, z by Box-Muller       (integer T: bad_function_call)
///   binomial(n = a, p = b as DOUBLE): sum of n Bernoulli draws for n <= 64,
///     clamped rounded normal approximation beyond
///

#include "common.cuh"
#include "vm.cuh"

namespace llo_cuda
{

struct Philox
{
	uint32_t c[4];
	uint32_t k[2];
};

__device__ __forceinline__ void philox_round (Philox& s)
{
	const uint32_t M0 = 0xD2511F53u;
	const uint32_t M1 = 0xCD9E8D57u;
	uint32_t hi0 = __umulhi(M0, s.c[0]);
	uint32_t lo0 = M0 * s.c[0];
	uint32_t hi1 = __umulhi(M1, s.c[2]);
	uint32_t lo1 = M1 * s.c[2];
	uint32_t n0 = hi1 ^ s.c[1] ^ s.k[0];
	uint32_t n1 = lo1;
	uint32_t n2 = hi0 ^ s.c[3] ^ s.k[1];
	uint32_t n3 = lo0;
	s.c[0] = n0;
	s.c[1] = n1;
	s.c[2] = n2;
	s.c[3] = n3;
	s.k[0] += 0x9E3779B9u;
	s.k[1] += 0xBB67AE85u;
}

/// 4 x 32 random bits for (seed, element, draw)
__device__ __forceinline__ void philox4 (uint32_t (&out)[4],
	uint64_t seed, uint64_t element, uint32_t draw)
{
	Philox s;
	s.c[0] = (uint32_t) element;
	s.c[1] = (uint32_t) (element >> 32);
	s.c[2] = draw;
	s.c[3] = 0x5851F42Du;
	s.k[0] = (uint32_t) seed;
	s.k[1] = (uint32_t) (seed >> 32);
#pragma unroll
	for (int r = 0; r < 10; ++r)
	{
		philox_round(s);
	}
#pragma unroll
	for (int i = 0; i < 4; ++i)
	{
		out[i] = s.c[i];
	}
}

/// uniform double in [0, 1) from 53 random bits
__device__ __forceinline__ double unit_double (uint32_t hi, uint32_t lo)
{
	uint64_t bits = (((uint64_t) hi << 32) | lo) >> 11;
	return (double) bits * (1.0 / 9007199254740992.0);
}

struct RandParams
{
	void* out;
	uint64_t n;
	uint64_t seed;
	uint64_t offset;
	VmParams view; // only the coordinate machinery and in[0], in[1] are used
};

template <typename T, typename B>
__global__ void __launch_bounds__(256)
rand_kernel (int opcode, const __grid_constant__ RandParams p)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	T* out = static_cast<T*>(p.out);
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		i < p.n; i += step)
	{
		T a[1];
		B b[1];
		a[0] = load_element<T>(p.view, 0, i);
		b[0] = load_element<B>(p.view, 1, i);
		uint32_t r[4];
		philox4(r, p.seed, p.offset + i, 0);
		T result;
		if (LLO_OP_RAND_UNIF == opcode)
		{
			double u = unit_double(r[0], r[1]);
			if constexpr (is_fp<T>::value)
			{
				T v = (T) ((double) a[0] + ((double) b[0] - (double) a[0]) * u);
				// rounding to T may land on the open end
				if (v >= (T) b[0] && (T) b[0] > a[0])
				{
					v = a[0];
				}
				result = v;
			}
			else
			{
				double span = (double) b[0] - (double) a[0] + 1.0;
				double v = floor(u * span);
				if (v >= span)
				{
					v = span - 1.0;
				}
				result = (T) ((double) a[0] + v);
			}
		}
		else if (LLO_OP_RAND_NORM == opcode)
		{
			if constexpr (is_fp<T>::value)
			{
				double u1 = 1.0 - unit_double(r[0], r[1]); // (0, 1]
				double u2 = unit_double(r[2], r[3]);
				double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
				result = (T) ((double) a[0] + (double) b[0] * z);
			}
			else
			{
				result = T();
			}
		}
		else // RAND_BINO: a = trials, b = probability (always double)
		{
			double prob = (double) b[0];
			double trials = floor((double) a[0]);
			double successes = 0;
			if (trials <= 64)
			{
				int nt = (int) trials;
				uint32_t draw = 0;
				int have = 0;
				for (int t = 0; t < nt; ++t)
				{
					if (0 == have)
					{
						philox4(r, p.seed, p.offset + i, ++draw);
						have = 2;
					}
					double u = (2 == have) ? unit_double(r[0], r[1]) :
						unit_double(r[2], r[3]);
					--have;
					successes += (u < prob) ? 1.0 : 0.0;
				}
			}
			else
			{
				double u1 = 1.0 - unit_double(r[0], r[1]);
				double u2 = unit_double(r[2], r[3]);
				double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
				double v = rint(trials * prob +
					z * sqrt(trials * prob * (1.0 - prob)));
				successes = v < 0 ? 0 : (v > trials ? trials : v);
			}
			result = (T) successes;
		}
		out[i] = result;
	}
}

}

using namespace llo_cuda;

namespace llo_cuda
{

// capi.cu: turn an argument into a VM input (view / index table)
int make_vm_input (llo_cuda_ctx* ctx, Scratch& scratch, VmInput& in,
	const llo_cuda_arg& arg, const uint64_t outshape[LLO_RANK],
	bool& pull_ok);

}

extern "C" int llo_cuda_rand (llo_cuda_ctx* ctx, int opcode, int dtype,
	void* out, const uint64_t outshape[LLO_RANK],
	const llo_cuda_arg* a, const llo_cuda_arg* b)
{
	if (false == is_rand_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "%s is not a random operator",
			op_name(opcode));
	}
	if (LLO_OP_RAND_NORM == opcode && false == is_float_dtype(dtype))
	{
		// llo/operator.hpp:352-355
		return fail(LLO_ERR_UNSUPPORTED_TYPED_OP,
			"RAND_NORM is not defined for %s", dtype_name(dtype));
	}
	if (a->push || b->push)
	{
		return fail(LLO_ERR_FATAL, "random operators take pull-mapped "
			"arguments only");
	}
	Scratch scratch(ctx);
	VmInput vin[2];
	bool ok = true;
	LLO_TRY(make_vm_input(ctx, scratch, vin[0], *a, outshape, ok));
	LLO_TRY(make_vm_input(ctx, scratch, vin[1], *b, outshape, ok));
	llo_cuda_insn prog[3] = {{LLO_VM_LOAD, 0}, {LLO_VM_LOAD, 1},
		{LLO_OP_SUM, 0}};
	RandParams p;
	// the coordinate set-up is shared with the elementwise engine; both
	// inputs are read one element at a time, so vectors never apply
	LLO_TRY(vm_prepare(p.view, dtype, outshape, vin, 2, nullptr, 0,
		prog, 3, 1, 1));
	p.view.in[0].vec_ok = 0;
	p.view.in[1].vec_ok = 0;
	p.out = out;
	p.n = n_elems(outshape);
	p.seed = ctx->seed;
	p.offset = ctx->rand_offset;
	ctx->rand_offset += p.n;
	uint64_t blocks = (p.n + 255) / 256;
	uint64_t cap = (uint64_t) ctx->sm_count * 16;
	unsigned grid = (unsigned) (blocks < cap ? blocks : cap);
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		if (LLO_OP_RAND_BINO == opcode)
		{
			rand_kernel<T,double><<<grid, 256, 0, ctx->stream>>>(opcode, p);
		}
		else
		{
			rand_kernel<T,T><<<grid, 256, 0, ctx->stream>>>(opcode, p);
		}
		return launched(ctx, "rand_kernel");
	})
	return LLO_OK;
}
