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
/// Dispatch is warp-uniform (all threads run the same program) and each
/// thread owns VEC consecutive outputs, so the interpretation cost is
/// amortised over VEC elements.  The host resolves the stack depth of every
/// instruction, so the 4 stack slots are plain registers: cheap operators
/// have one code copy per slot and move nothing, expensive ones (libm-class
/// math, division) go through one shared copy.  Contiguous inputs / outputs
/// move as 128-bit vectors.  Two launch shapes:
///   flat  one chunk of VEC outputs per thread, one CTA per 256 chunks (a
///         non-persistent grid measured faster than a grid-stride loop on
///         B200: tools/probes/copy_probe.cu)
///   tile  output tiles for programs with a transposed input: the input
///         tile is read along ITS contiguous dimension, staged in padded
///         shared memory and consumed transposed, so both the load and the
///         store side stay coalesced
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
	VM_IN_TILE,        // strided, staged through a transposing smem tile
};

struct VmInput
{
	const void* data;
	const int64_t* index;
	int64_t stride[LLO_RANK];
	int64_t offset;
	int32_t mode;
	int32_t vec_ok; // LINEAR/STRIDED: 128-bit loads are aligned; TILE: slot
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

constexpr int VM_TILEP = 64;   // tile extent along the transposed dim
constexpr int VM_MAX_TILE_INPUTS = 2;

/// Internal instruction: `sel` = operation and stack slot resolved by the
/// host (see vm_encode in elementwise.cu), `arg` = input / constant number
struct VmOp
{
	uint8_t sel;
	uint8_t arg;
};

enum VmSel
{
	VM_SEL_LOAD = 0,        // +slot written (0..3)
	VM_SEL_CONST = 4,       // +slot written
	VM_SEL_ABS = 8,         // unary, +slot (0..3)
	VM_SEL_NEG = 12,
	VM_SEL_ROUND = 16,
	VM_SEL_SUM = 20,        // binary, +destination slot (0..2), rhs = slot + 1
	VM_SEL_SUB = 23,
	VM_SEL_PROD = 26,
	VM_SEL_MIN = 29,
	VM_SEL_MAX = 32,
	VM_SEL_EQ = 35,
	VM_SEL_NEQ = 38,
	VM_SEL_LT = 41,
	VM_SEL_GT = 44,
	VM_SEL_HEAVY = 64,      // + 4 * VmHeavy + slot
};

enum VmHeavy
{
	VM_H_SIN = 0, VM_H_COS, VM_H_TAN, VM_H_EXP, VM_H_LOG, VM_H_SQRT,
	VM_H_POW, VM_H_DIV, VM_H_COUNT,
};

struct VmParams
{
	void* out;
	uint64_t n;             // output elements
	int32_t ndims;          // collapsed output dims
	int32_t rows_aligned;   // dims[0] % VEC == 0: a chunk never straddles a row
	int32_t idx32;          // n < 2^31 and every input offset fits in 32 bits
	int32_t out_vec_ok;
	int32_t ninsns;
	int32_t ninputs;
	// tile launch only
	uint32_t tiles0;
	uint32_t tilesp;
	uint64_t dims[LLO_RANK];
	int64_t out_stride[LLO_RANK]; // dense strides of the collapsed dims
	FastDiv div[LLO_RANK];
	VmInput in[LLO_VM_MAX_INPUTS];
	uint64_t consts[LLO_VM_MAX_CONSTS]; // raw bits of T
	VmOp code[LLO_VM_MAX_INSNS];
};

template <typename T>
struct VmVec
{
	// outputs per thread
	static constexpr int value = (sizeof(T) >= 8) ? 4 : 8;
};

/// 128-bit load of VEC elements when aligned, scalar otherwise.  STREAM
/// marks data read exactly once (evict-first); views that revisit elements
/// (broadcasts) keep the default policy so they stay in L1 / L2.
template <typename T, int VEC, bool STREAM = true>
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
				tmp[k] = STREAM ? __ldcs(v + k) : __ldg(v + k);
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

/// Offset of output element i inside strided input `in`: the coordinates
/// of i are peeled off dim by dim and dotted with the input's strides
__device__ __forceinline__ int64_t strided_offset (const VmParams& p,
	const VmInput& in, uint64_t i)
{
	if (p.idx32)
	{
		uint32_t rem = (uint32_t) i;
		int32_t off = (int32_t) in.offset;
		int last = p.ndims - 1;
		for (int d = 0; d < last; ++d)
		{
			uint32_t q = fastdiv(rem, p.div[d]);
			off += (int32_t) (rem - q * p.div[d].d) * (int32_t) in.stride[d];
			rem = q;
		}
		off += (int32_t) rem * (int32_t) in.stride[last];
		return off;
	}
	uint64_t rem = i;
	int64_t off = in.offset;
	int last = p.ndims - 1;
	for (int d = 0; d < last; ++d)
	{
		uint64_t q = rem / p.dims[d];
		off += (int64_t) (rem - q * p.dims[d]) * in.stride[d];
		rem = q;
	}
	off += (int64_t) rem * in.stride[last];
	return off;
}

/// VEC values of a strided input starting at element offset `off` and
/// running along output dim 0
template <typename T, int VEC>
__device__ __forceinline__ void load_strided_run (T (&dst)[VEC],
	const VmInput& in, int64_t off, int valid)
{
	const T* data = static_cast<const T*>(in.data);
	int64_t s0 = in.stride[0];
	if (1 == s0)
	{
		load_run<T,VEC,false>(dst, data + off, in.vec_ok, valid);
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

/// Input access of the flat launch: chunk [base, base + valid) of the
/// output's linear index space
template <typename T, int VEC>
struct FlatLoader
{
	const VmParams& p;
	uint64_t base;
	int valid;

	__device__ __forceinline__ void load (T (&dst)[VEC], int k) const
	{
		const VmInput& in = p.in[k];
		const T* data = static_cast<const T*>(in.data);
		if (VM_IN_LINEAR == in.mode)
		{
			load_run<T,VEC>(dst, data + in.offset + base, in.vec_ok, valid);
		}
		else if (VM_IN_INDEXED == in.mode)
		{
#pragma unroll
			for (int e = 0; e < VEC; ++e)
			{
				dst[e] = (e < valid) ?
					__ldg(data + __ldg(in.index + base + e)) : T();
			}
		}
		else if (p.rows_aligned)
		{
			// the chunk shares every coordinate but the fastest
			load_strided_run<T,VEC>(dst, in, strided_offset(p, in, base),
				valid);
		}
		else
		{
			// chunks may straddle rows: address every element on its own
#pragma unroll
			for (int e = 0; e < VEC; ++e)
			{
				dst[e] = (e < valid) ?
					__ldg(data + strided_offset(p, in, base + e)) : T();
			}
		}
	}
};

#define LLO_VM_EACH(BODY) _Pragma("unroll")\
	for (int k = 0; k < VEC; ++k) { BODY; }

/// Run the program for one chunk; the result is left in s0
template <typename T, int VEC, typename Loader>
__device__ __forceinline__ void vm_run (T (&s0)[VEC], const VmParams& p,
	const Loader& loader)
{
	T s1[VEC];
	T s2[VEC];
	T s3[VEC];
	for (int pc = 0; pc < p.ninsns; ++pc)
	{
		int sel = p.code[pc].sel;
		int arg = p.code[pc].arg;
		if (sel < VM_SEL_ABS)
		{
			// LOAD / CONST: fetch once, then move into the slot
			T t[VEC];
			if (sel < VM_SEL_CONST)
			{
				loader.load(t, arg);
			}
			else
			{
				T c;
				memcpy(&c, &p.consts[arg], sizeof(T));
				LLO_VM_EACH(t[k] = c)
			}
			switch (sel & 3)
			{
				case 0: LLO_VM_EACH(s0[k] = t[k]) break;
				case 1: LLO_VM_EACH(s1[k] = t[k]) break;
				case 2: LLO_VM_EACH(s2[k] = t[k]) break;
				default: LLO_VM_EACH(s3[k] = t[k]) break;
			}
		}
		else if (sel < VM_SEL_HEAVY)
		{
			switch (sel)
			{
#define LLO_VM_UNARY(SEL, FN)\
				case SEL + 0: LLO_VM_EACH(s0[k] = FN(s0[k])) break;\
				case SEL + 1: LLO_VM_EACH(s1[k] = FN(s1[k])) break;\
				case SEL + 2: LLO_VM_EACH(s2[k] = FN(s2[k])) break;\
				case SEL + 3: LLO_VM_EACH(s3[k] = FN(s3[k])) break;
				LLO_VM_UNARY(VM_SEL_ABS, m_abs)
				LLO_VM_UNARY(VM_SEL_NEG, m_neg)
				LLO_VM_UNARY(VM_SEL_ROUND, m_round)
#undef LLO_VM_UNARY
#define LLO_VM_BINARY(SEL, EXPR)\
				case SEL + 0: LLO_VM_EACH(T a = s0[k]; T b = s1[k]; s0[k] = EXPR) break;\
				case SEL + 1: LLO_VM_EACH(T a = s1[k]; T b = s2[k]; s1[k] = EXPR) break;\
				case SEL + 2: LLO_VM_EACH(T a = s2[k]; T b = s3[k]; s2[k] = EXPR) break;
				LLO_VM_BINARY(VM_SEL_SUM, m_add(a, b))
				LLO_VM_BINARY(VM_SEL_SUB, m_sub(a, b))
				LLO_VM_BINARY(VM_SEL_PROD, m_mul(a, b))
				LLO_VM_BINARY(VM_SEL_MIN, m_min(a, b))
				LLO_VM_BINARY(VM_SEL_MAX, m_max(a, b))
				LLO_VM_BINARY(VM_SEL_EQ, (T) (a == b))
				LLO_VM_BINARY(VM_SEL_NEQ, (T) (a != b))
				LLO_VM_BINARY(VM_SEL_LT, (T) (a < b))
				LLO_VM_BINARY(VM_SEL_GT, (T) (a > b))
#undef LLO_VM_BINARY
				default: break;
			}
		}
		else
		{
			// one shared copy of the expensive math: operands are moved out
			// of, and the result back into, the slot
			int slot = sel & 3;
			int hop = (sel - VM_SEL_HEAVY) >> 2;
			T a[VEC];
			T b[VEC];
			switch (slot)
			{
				case 0: LLO_VM_EACH(a[k] = s0[k]; b[k] = s1[k]) break;
				case 1: LLO_VM_EACH(a[k] = s1[k]; b[k] = s2[k]) break;
				case 2: LLO_VM_EACH(a[k] = s2[k]; b[k] = s3[k]) break;
				default: LLO_VM_EACH(a[k] = s3[k]; b[k] = s3[k]) break;
			}
			switch (hop)
			{
				case VM_H_SIN: LLO_VM_EACH(a[k] = m_sin(a[k])) break;
				case VM_H_COS: LLO_VM_EACH(a[k] = m_cos(a[k])) break;
				case VM_H_TAN: LLO_VM_EACH(a[k] = m_tan(a[k])) break;
				case VM_H_EXP: LLO_VM_EACH(a[k] = m_exp(a[k])) break;
				case VM_H_LOG: LLO_VM_EACH(a[k] = m_log(a[k])) break;
				case VM_H_SQRT: LLO_VM_EACH(a[k] = m_sqrt(a[k])) break;
				case VM_H_POW: LLO_VM_EACH(a[k] = m_pow(a[k], b[k])) break;
				default: LLO_VM_EACH(a[k] = m_div(a[k], b[k])) break;
			}
			switch (slot)
			{
				case 0: LLO_VM_EACH(s0[k] = a[k]) break;
				case 1: LLO_VM_EACH(s1[k] = a[k]) break;
				case 2: LLO_VM_EACH(s2[k] = a[k]) break;
				default: LLO_VM_EACH(s3[k] = a[k]) break;
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
/// shared with the random fills
int vm_prepare (VmParams& p, int dtype, const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs, const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns, int vec);

}

#endif // LLO_CUDA_VM_CUH
