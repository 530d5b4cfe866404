///
/// vm.cuh — the fused elementwise engine (K1/K2/K3).
///
/// One kernel family evaluates a small stack program per output element:
/// LOADs read inputs through strided views (the CoordMap of each argument
/// folded into the load address: contiguous, broadcast, permuted, flipped,
/// or gathered through an index table), opcodes apply the typed operators
/// of ops.cuh, the value left on the stack is stored.  A single-operator
/// call is a 2-4 instruction program; a chain of operators is one pass over
/// HBM instead of one pass per operator.
///
/// Dispatch is warp-uniform (all threads run the same program), each thread
/// owns VEC consecutive outputs so the interpretation cost is amortised over
/// VEC elements, and the 4-slot stack lives in registers (the stack rotates,
/// so every slot is statically indexed).  Contiguous inputs / outputs move
/// as 128-bit vectors.
///

#ifndef LLO_CUDA_VM_CUH
#define LLO_CUDA_VM_CUH

#include "common.cuh"
#include "ops.cuh"

namespace llo_cuda
{

enum VmInputMode
{
	VM_IN_LINEAR = 0,  // data[offset + i], i = output linear index
	VM_IN_STRIDED,     // data[offset + sum_d c_d * stride[d]]
	VM_IN_INDEXED,     // data[index[i]]  (general-map gather table)
	VM_IN_TILE,        // staged through a transposing shared-memory tile
};

struct VmInput
{
	const void* data;
	const int64_t* index;
	int64_t stride[LLO_RANK];
	int64_t offset;
	int32_t mode;
	int32_t vec_ok; // LINEAR/STRIDED: 128-bit loads are aligned
};

struct FastDiv
{
	uint32_t d;
	uint32_t m;
	uint32_t s;
};

/// n / d for n < 2^31 with one multiply-high
inline FastDiv make_fastdiv (uint32_t d)
{
	FastDiv f;
	f.d = d;
	uint32_t s = 0;
	while (s < 32 && (1ull << s) < d)
	{
		++s;
	}
	f.s = s;
	f.m = (uint32_t) (((1ull << 32) * ((1ull << s) - d)) / d + 1);
	return f;
}

__device__ __forceinline__ uint32_t fastdiv (uint32_t n, const FastDiv& f)
{
	return (__umulhi(n, f.m) + n) >> f.s;
}

struct VmParams
{
	void* out;
	uint64_t n;             // output elements
	int32_t ndims;          // collapsed output dims
	int32_t need_coords;    // some input is STRIDED
	int32_t rows_aligned;   // dims[0] % VEC == 0: a chunk never straddles a row
	int32_t idx32;          // n < 2^31: 32-bit coordinate arithmetic
	int32_t out_vec_ok;
	int32_t ninsns;
	int32_t ninputs;
	int32_t pad;
	uint64_t dims[LLO_RANK];
	FastDiv div[LLO_RANK];
	VmInput in[LLO_VM_MAX_INPUTS];
	uint64_t consts[LLO_VM_MAX_CONSTS]; // raw bits of T
	llo_cuda_insn code[LLO_VM_MAX_INSNS];
};

template <typename T>
struct VmVec
{
	// outputs per thread and iteration
	static constexpr int value = (sizeof(T) >= 8) ? 4 : 8;
};

/// 128-bit streaming load of VEC elements when aligned, scalar otherwise
template <typename T, int VEC>
__device__ __forceinline__ void load_run (T (&dst)[VEC], const T* src,
	bool vec_ok, int valid)
{
	constexpr int bytes = VEC * sizeof(T);
	if constexpr (bytes >= 16)
	{
		if (vec_ok && VEC == valid)
		{
			constexpr int nvec = bytes / 16;
			const uint4* v = reinterpret_cast<const uint4*>(src);
			uint4 tmp[nvec];
#pragma unroll
			for (int k = 0; k < nvec; ++k)
			{
				tmp[k] = __ldcs(v + k);
			}
			const T* t = reinterpret_cast<const T*>(tmp);
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				dst[k] = t[k];
			}
			return;
		}
	}
#pragma unroll
	for (int k = 0; k < VEC; ++k)
	{
		if (k < valid)
		{
			dst[k] = __ldg(src + k);
		}
		else
		{
			dst[k] = T();
		}
	}
}

template <typename T, int VEC>
__device__ __forceinline__ void store_run (T* dst, const T (&src)[VEC],
	bool vec_ok, int valid)
{
	constexpr int bytes = VEC * sizeof(T);
	if constexpr (bytes >= 16)
	{
		if (vec_ok && VEC == valid)
		{
			constexpr int nvec = bytes / 16;
			uint4 tmp[nvec];
			T* t = reinterpret_cast<T*>(tmp);
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				t[k] = src[k];
			}
			uint4* v = reinterpret_cast<uint4*>(dst);
#pragma unroll
			for (int k = 0; k < nvec; ++k)
			{
				__stcs(v + k, tmp[k]);
			}
			return;
		}
	}
#pragma unroll
	for (int k = 0; k < VEC; ++k)
	{
		if (k < valid)
		{
			dst[k] = src[k];
		}
	}
}

/// Offset of output element `i` inside a strided input
__device__ __forceinline__ int64_t strided_offset (const VmParams& p,
	const VmInput& in, uint64_t i)
{
	int64_t off = in.offset;
	if (p.idx32)
	{
		uint32_t rem = (uint32_t) i;
		for (int d = 0; d < p.ndims - 1; ++d)
		{
			uint32_t q = fastdiv(rem, p.div[d]);
			uint32_t c = rem - q * p.div[d].d;
			off += (int64_t) c * in.stride[d];
			rem = q;
		}
		off += (int64_t) rem * in.stride[p.ndims - 1];
	}
	else
	{
		uint64_t rem = i;
		for (int d = 0; d < p.ndims - 1; ++d)
		{
			uint64_t q = rem / p.dims[d];
			uint64_t c = rem - q * p.dims[d];
			off += (int64_t) c * in.stride[d];
			rem = q;
		}
		off += (int64_t) rem * in.stride[p.ndims - 1];
	}
	return off;
}

/// Fetch VEC values of input `in` for outputs [base, base + valid)
template <typename T, int VEC>
__device__ __forceinline__ void vm_load (T (&dst)[VEC], const VmParams& p,
	const VmInput& in, uint64_t base, int valid)
{
	const T* data = static_cast<const T*>(in.data);
	if (VM_IN_LINEAR == in.mode)
	{
		load_run<T,VEC>(dst, data + in.offset + base, in.vec_ok, valid);
	}
	else if (VM_IN_STRIDED == in.mode)
	{
		if (p.rows_aligned)
		{
			// the chunk shares every coordinate but the fastest
			int64_t off = strided_offset(p, in, base);
			int64_t s0 = in.stride[0];
			if (1 == s0)
			{
				load_run<T,VEC>(dst, data + off, in.vec_ok, valid);
			}
			else if (0 == s0)
			{
				T v = __ldg(data + off);
#pragma unroll
				for (int k = 0; k < VEC; ++k)
				{
					dst[k] = v;
				}
			}
			else
			{
#pragma unroll
				for (int k = 0; k < VEC; ++k)
				{
					dst[k] = (k < valid) ? __ldg(data + off + k * s0) : T();
				}
			}
		}
		else
		{
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				dst[k] = (k < valid) ?
					__ldg(data + strided_offset(p, in, base + k)) : T();
			}
		}
	}
	else // VM_IN_INDEXED
	{
#pragma unroll
		for (int k = 0; k < VEC; ++k)
		{
			dst[k] = (k < valid) ? __ldg(data + __ldg(in.index + base + k)) : T();
		}
	}
}

/// Run the program for one chunk; the result is left in s0
template <typename T, int VEC>
__device__ __forceinline__ void vm_run (T (&s0)[VEC], const VmParams& p,
	uint64_t base, int valid)
{
	T s1[VEC];
	T s2[VEC];
	T s3[VEC];
#pragma unroll
	for (int k = 0; k < VEC; ++k)
	{
		s0[k] = s1[k] = s2[k] = s3[k] = T();
	}
	for (int pc = 0; pc < p.ninsns; ++pc)
	{
		int op = p.code[pc].op;
		int arg = p.code[pc].arg;
		if (op >= LLO_VM_LOAD)
		{
			// push
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				s3[k] = s2[k];
				s2[k] = s1[k];
				s1[k] = s0[k];
			}
			if (LLO_VM_LOAD == op)
			{
				vm_load<T,VEC>(s0, p, p.in[arg], base, valid);
			}
			else
			{
				T c;
				memcpy(&c, &p.consts[arg], sizeof(T));
#pragma unroll
				for (int k = 0; k < VEC; ++k)
				{
					s0[k] = c;
				}
			}
		}
		else if (op <= LLO_OP_ROUND)
		{
			switch (op)
			{
#define LLO_VM_CASE(OP, FN) case OP:\
				_Pragma("unroll")\
				for (int k = 0; k < VEC; ++k) { s0[k] = FN(s0[k]); }\
				break;
				LLO_VM_CASE(LLO_OP_ABS, m_abs)
				LLO_VM_CASE(LLO_OP_NEG, m_neg)
				LLO_VM_CASE(LLO_OP_SIN, m_sin)
				LLO_VM_CASE(LLO_OP_COS, m_cos)
				LLO_VM_CASE(LLO_OP_TAN, m_tan)
				LLO_VM_CASE(LLO_OP_EXP, m_exp)
				LLO_VM_CASE(LLO_OP_LOG, m_log)
				LLO_VM_CASE(LLO_OP_SQRT, m_sqrt)
				LLO_VM_CASE(LLO_OP_ROUND, m_round)
#undef LLO_VM_CASE
				default: break;
			}
		}
		else
		{
			// pop two, push one: a = s1 (pushed first), b = s0
			switch (op)
			{
#define LLO_VM_CASE(OP, EXPR) case OP:\
				_Pragma("unroll")\
				for (int k = 0; k < VEC; ++k)\
				{ T a = s1[k]; T b = s0[k]; s0[k] = EXPR; }\
				break;
				LLO_VM_CASE(LLO_OP_POW, m_pow(a, b))
				LLO_VM_CASE(LLO_OP_SUM, m_add(a, b))
				LLO_VM_CASE(LLO_OP_SUB, m_sub(a, b))
				LLO_VM_CASE(LLO_OP_PROD, m_mul(a, b))
				LLO_VM_CASE(LLO_OP_DIV, m_div(a, b))
				LLO_VM_CASE(LLO_OP_MIN, m_min(a, b))
				LLO_VM_CASE(LLO_OP_MAX, m_max(a, b))
				LLO_VM_CASE(LLO_OP_EQ, (T) (a == b))
				LLO_VM_CASE(LLO_OP_NEQ, (T) (a != b))
				LLO_VM_CASE(LLO_OP_LT, (T) (a < b))
				LLO_VM_CASE(LLO_OP_GT, (T) (a > b))
#undef LLO_VM_CASE
				default: break;
			}
#pragma unroll
			for (int k = 0; k < VEC; ++k)
			{
				s1[k] = s2[k];
				s2[k] = s3[k];
			}
		}
	}
}

/// Host side: validate and launch (elementwise.cu)
int vm_launch (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs,
	const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns);

/// Fill the shape-dependent part of VmParams (collapsing compatible dims);
/// shared with the reducing variant
int vm_prepare (VmParams& p, int dtype, const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs, const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns, int vec);

}

#endif // LLO_CUDA_VM_CUH
