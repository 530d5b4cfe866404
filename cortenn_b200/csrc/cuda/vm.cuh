///
/// vm.cuh — the fused elementwise engine (K1/K2/K3).
///
/// One kernel family evaluates a small program per output element: operands
/// are read through strided views (the CoordMap of each argument folded into
/// the load address: contiguous, broadcast, permuted, flipped, or gathered
/// through an index table), opcodes apply the typed operators of ops.cuh, the
/// result is stored.  A single-operator call is a 2-3 instruction program; a
/// chain of operators is one pass over HBM instead of one pass per operator.
///
/// The public program (llo_cuda_insn) is a stack program; the host compiles
/// it to ACCUMULATOR form (vm_encode in elementwise.cu): every instruction
/// updates one accumulator with an operand that is an input, a constant or a
/// spilled sub-expression, so the device keeps exactly two value sets in
/// registers (accumulator + operand), never moves values between slots, and
/// spills (rare: both operands of an operator are sub-expressions) go to
/// shared memory.
///
/// Dispatch is warp-uniform and amortised over V = 16 elements per thread
/// (8 for 8-byte types): four 128-bit chunks, each chunk strided by the CTA
/// width so a warp's request covers 512 contiguous bytes and every thread
/// keeps 64 bytes per operand in flight.  Two launch shapes:
///   flat  CTA = 256 threads x V consecutive outputs (a non-persistent grid
///         measured faster than a grid-stride loop on B200:
///         tools/probes/copy_probe.cu)
///   tile  output tiles for programs with a transposed input: the input
///         tile is read along ITS contiguous dimension, staged in padded
///         shared memory and consumed transposed, so both the load and the
///         store side stay coalesced
///

#ifndef LLO_CUDA_VM_CUH
#define LLO_CUDA_VM_CUH

#ifndef __CUDACC_RTC__
#include "common.cuh"
#endif
#include "ops.cuh"

namespace llo_cuda
{

enum VmInputMode
{
	VM_IN_LINEAR = 0,  // data[offset + i], i = output linear index
	VM_IN_STRIDED,     // data[offset + sum_d c_d * stride[d]]
	VM_IN_INDEXED,     // data[index[i]]  (general-map gather table)
	VM_IN_TILE,        // strided, staged through a transposing smem tile
	VM_IN_ALIAS,       // the straight view of a buffer whose transposed view is
	                   // VM_IN_TILE: read from the mirror tile already in smem
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

#ifndef __CUDACC_RTC__
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
#endif

__device__ __forceinline__ uint32_t fastdiv (uint32_t n, const FastDiv& f)
{
	return (__umulhi(n, f.m) + n) >> f.s;
}

constexpr int VM_THREADS = 256;
constexpr int VM_MAX_TILE_INPUTS = 2;
constexpr int VM_MAX_CODE = 64;   // accumulator-form instructions
constexpr int VM_MAX_SPILL = 3;   // spill levels (stack depth 4)

/// Per-type geometry: E elements per 128-bit chunk, U chunks per thread
template <typename T>
struct VmGeom
{
	static constexpr int E = 16 / (int) sizeof(T);
	static constexpr int U = sizeof(T) >= 4 ? 4 : (sizeof(T) == 2 ? 2 : 1);
	static constexpr int V = E * U;
	// tile launch: TILE0 outputs along dim 0 x TILEP along the staged dim
	static constexpr int TILE0 = 16 * E;
	static constexpr int TILEP = 16 * U;
	static constexpr int PITCH = TILE0 + 1;
};

/// Operand source of an accumulator-form instruction (VmOp::arg >> 4)
enum VmSrc
{
	VM_SRC_NONE = 0,
	VM_SRC_IN = 1,     // input  (arg & 15)
	VM_SRC_CONST = 2,  // consts (arg & 15)
	VM_SRC_POP = 3,    // most recent spill
};

/// Accumulator-form operations: acc <- f(acc, x), x the operand
enum VmSel
{
	VM_S_SET = 0,   // acc = x
	VM_S_PUSH,      // spill acc
	VM_S_ABS, VM_S_NEG, VM_S_ROUND, VM_S_SIN, VM_S_COS, VM_S_TAN, VM_S_EXP,
	VM_S_LOG, VM_S_SQRT,
	VM_S_SUM, VM_S_PROD,
	VM_S_SUB, VM_S_RSUB,     // acc - x, x - acc
	VM_S_MIN, VM_S_RMIN,     // min(acc, x), min(x, acc)  (std::min argument order)
	VM_S_MAX, VM_S_RMAX,
	VM_S_EQ, VM_S_NEQ,
	VM_S_LT, VM_S_GT,        // acc < x, acc > x
	VM_S_DIV, VM_S_RDIV,
	VM_S_POW, VM_S_RPOW,
};

struct VmOp
{
	uint8_t sel;
	uint8_t arg;
};

struct VmParams
{
	void* out;
	uint64_t n;             // output elements
	int32_t ndims;          // collapsed output dims
	int32_t rows_aligned;   // dims[0] % E == 0: a chunk never straddles a row
	int32_t row_uniform;    // dims[0] % (256 * V) == 0: a CTA never does
	int32_t idx32;          // n < 2^31 and every input offset fits in 32 bits
	int32_t out_vec_ok;
	int32_t ninsns;
	int32_t ninputs;
	int32_t spill_depth;    // shared-memory spill levels the program needs
	// tile launch only
	uint32_t tiles0;
	uint32_t tilesp;
	int32_t nstaged;        // inputs staged through shared-memory tiles
	int32_t tile_aligned;   // every staged row start is 16-byte aligned
	int32_t mirror;         // square tile grid: tile (i,j) runs next to (j,i)
	uint32_t total_tiles;   // tiles of the whole launch (persistent kernels)
	int32_t pair_mode;      // one CTA computes output tiles (i,j) and (j,i) from
	                        // the two input tiles it staged once (vm_pair4_kernel)
	int32_t tile_input;     // pair mode: index of the staged input
	// persistent tile kernels: {next work unit, finished CTAs}; CTAs claim
	// their units from it so faster SMs take more (the last CTA out re-arms it)
	uint32_t* sched;
	uint64_t dims[LLO_RANK];
	int64_t out_stride[LLO_RANK]; // dense strides of the collapsed dims
	FastDiv div[LLO_RANK];
	VmInput in[LLO_VM_MAX_INPUTS];
	uint64_t consts[LLO_VM_MAX_CONSTS]; // raw bits of T
	VmOp code[VM_MAX_CODE];
};

/// What a loader branches on for one input.  In the interpreted build these
/// are read from VmParams at run time (warp-uniform branches); in the
/// specialised build (LLO_VM_STATIC, NVRTC) they are compile-time constants
/// and the branches fold away.  s0 / s1: the stride along output dim 0 / 1
/// as a class -- 0 (broadcast), 1 (contiguous) or 2 (anything else).
struct VmKind
{
	int mode;
	int s0;
	int s1;
	int vec_ok;
};

__device__ __forceinline__ int vm_stride_class (int64_t stride)
{
	return 0 == stride ? 0 : (1 == stride ? 1 : 2);
}

__device__ __forceinline__ VmKind vm_kind (const VmParams& p, int k)
{
	const VmInput& in = p.in[k];
	return VmKind{in.mode, vm_stride_class(in.stride[0]),
		vm_stride_class(in.stride[1]), in.vec_ok};
}

#ifdef LLO_VM_STATIC
// the specialised build knows the collapsed rank when it is 1..3 (0 = any) and
// the launch-wide addressing flags
#define LLO_VM_NDIMS(p) (LLO_VM_STATIC_NDIMS > 0 ? LLO_VM_STATIC_NDIMS : (p).ndims)
#define LLO_VM_IDX32(p) (0 != LLO_VM_S_IDX32)
#define LLO_VM_ROW_UNIFORM(p) (0 != LLO_VM_S_ROW_UNIFORM)
#define LLO_VM_ROWS_ALIGNED(p) (0 != LLO_VM_S_ROWS_ALIGNED)
#define LLO_VM_OUT_VEC(p) (0 != LLO_VM_S_OUT_VEC)
#else
#define LLO_VM_NDIMS(p) ((p).ndims)
#define LLO_VM_IDX32(p) (0 != (p).idx32)
#define LLO_VM_ROW_UNIFORM(p) (0 != (p).row_uniform)
#define LLO_VM_ROWS_ALIGNED(p) (0 != (p).rows_aligned)
#define LLO_VM_OUT_VEC(p) (0 != (p).out_vec_ok)
#endif

/// E elements from `src`: one 128-bit load when aligned and complete, scalar
/// loads otherwise.  STREAM marks data read exactly once (evict-first);
/// views that revisit elements (broadcasts) keep the default policy so they
/// stay in L1 / L2.
template <typename T, int E, bool STREAM>
__device__ __forceinline__ void load_chunk (T* dst, const T* src,
	bool vec_ok, int valid)
{
	if constexpr (E * sizeof(T) == 16)
	{
		if (vec_ok && E == valid)
		{
			uint4 tmp = STREAM ? __ldcs(reinterpret_cast<const uint4*>(src)) :
				__ldg(reinterpret_cast<const uint4*>(src));
			const T* t = reinterpret_cast<const T*>(&tmp);
#pragma unroll
			for (int k = 0; k < E; ++k)
			{
				dst[k] = t[k];
			}
			return;
		}
	}
#pragma unroll
	for (int k = 0; k < E; ++k)
	{
		dst[k] = (k < valid) ? __ldg(src + k) : T();
	}
}

template <typename T, int E>
__device__ __forceinline__ void store_chunk (T* dst, const T* src,
	bool vec_ok, int valid)
{
	if constexpr (E * sizeof(T) == 16)
	{
		if (vec_ok && E == valid)
		{
			uint4 tmp;
			T* t = reinterpret_cast<T*>(&tmp);
#pragma unroll
			for (int k = 0; k < E; ++k)
			{
				t[k] = src[k];
			}
			__stcs(reinterpret_cast<uint4*>(dst), tmp);
			return;
		}
	}
#pragma unroll
	for (int k = 0; k < E; ++k)
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
	if (LLO_VM_IDX32(p))
	{
		uint32_t rem = (uint32_t) i;
		int32_t off = (int32_t) in.offset;
		int last = LLO_VM_NDIMS(p) - 1;
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

/// E values of a strided input starting at element offset `off` and running
/// along output dim 0
template <typename T, int E>
__device__ __forceinline__ void load_strided_chunk (T* dst,
	const VmInput& in, const VmKind& kind, int64_t off, int valid)
{
	const T* data = static_cast<const T*>(in.data);
	int64_t s0 = in.stride[0];
	if (1 == kind.s0)
	{
		load_chunk<T,E,false>(dst, data + off, kind.vec_ok, valid);
	}
	else if (0 == kind.s0)
	{
		T v = __ldg(data + off);
#pragma unroll
		for (int k = 0; k < E; ++k)
		{
			dst[k] = v;
		}
	}
	else
	{
#pragma unroll
		for (int k = 0; k < E; ++k)
		{
			dst[k] = (k < valid) ? __ldg(data + off + k * s0) : T();
		}
	}
}

/// Single-element access through the view machinery (random fills)
template <typename T>
__device__ __forceinline__ T load_element (const VmParams& p, int k, uint64_t i)
{
	const VmInput& in = p.in[k];
	const T* data = static_cast<const T*>(in.data);
	if (VM_IN_LINEAR == in.mode)
	{
		return __ldg(data + in.offset + i);
	}
	if (VM_IN_INDEXED == in.mode)
	{
		return __ldg(data + __ldg(in.index + i));
	}
	return __ldg(data + strided_offset(p, in, i));
}

/// Input access of the flat launch.  Chunk u of thread t covers the outputs
/// [base + (u * 256 + t) * E, +E)
template <typename T, bool FULL = false>
struct FlatLoader
{
	using G = VmGeom<T>;

	const VmParams& p;
	uint64_t base;   // first output of this CTA
	bool full;       // the CTA's 256 * V outputs all exist (FULL: known statically)

	__device__ __forceinline__ uint64_t first (int u) const
	{
		return base + (uint64_t) ((u * VM_THREADS + (int) threadIdx.x) * G::E);
	}

	__device__ __forceinline__ int valid (int u) const
	{
		if (FULL || full)
		{
			return G::E;
		}
		uint64_t e0 = first(u);
		if (e0 >= p.n)
		{
			return 0;
		}
		uint64_t left = p.n - e0;
		return left < (uint64_t) G::E ? (int) left : G::E;
	}

	__device__ __forceinline__ void load (T (&x)[G::V], int k,
		const VmKind& kind) const
	{
		const VmInput& in = p.in[k];
		const T* data = static_cast<const T*>(in.data);
		if (VM_IN_LINEAR == kind.mode)
		{
			const T* src = data + in.offset + first(0);
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				load_chunk<T,G::E,true>(x + u * G::E,
					src + u * VM_THREADS * G::E, kind.vec_ok, valid(u));
			}
		}
		else if (VM_IN_INDEXED == kind.mode)
		{
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				uint64_t e0 = first(u);
				int ok = valid(u);
#pragma unroll
				for (int e = 0; e < G::E; ++e)
				{
					x[u * G::E + e] = (e < ok) ?
						__ldg(data + __ldg(in.index + e0 + e)) : T();
				}
			}
		}
		else if (LLO_VM_ROW_UNIFORM(p))
		{
			// the CTA sits inside one row: only the dim-0 coordinate moves
			int64_t off = strided_offset(p, in, first(0));
			int64_t step = (int64_t) (VM_THREADS * G::E) * in.stride[0];
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				load_strided_chunk<T,G::E>(x + u * G::E, in, kind, off + u * step,
					G::E);
			}
		}
		else if (LLO_VM_ROWS_ALIGNED(p))
		{
			// a chunk shares every coordinate but the fastest
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				int ok = valid(u);
				int64_t off = ok > 0 ? strided_offset(p, in, first(u)) : in.offset;
				load_strided_chunk<T,G::E>(x + u * G::E, in, kind, off, ok);
			}
		}
		else
		{
			// chunks may straddle rows: address every element on its own
#pragma unroll
			for (int u = 0; u < G::U; ++u)
			{
				uint64_t e0 = first(u);
				int ok = valid(u);
#pragma unroll
				for (int e = 0; e < G::E; ++e)
				{
					x[u * G::E + e] = (e < ok) ?
						__ldg(data + strided_offset(p, in, e0 + e)) : T();
				}
			}
		}
	}
};

#define LLO_VM_EACH(BODY) _Pragma("unroll")\
	for (int e = 0; e < V; ++e) { BODY; }

/// The bulkiest operators (pow is > 100 instructions per element; tan / sin /
/// cos carry their argument reduction, log / div / sqrt / round their special
/// cases) inlined 16 times each make up over a third of every kernel that
/// carries the interpreter.  The
/// pair kernel is sensitive to code size (instruction-cache misses show up
/// as `no_inst` stalls; the chain ran 2% faster on the same box without that
/// third), so it (COMPACT = true) calls these operators out of line, four
/// elements per call; the flat kernel keeps them inline (a stand-alone pow
/// lost 4.5% to the calls), and so does vm_tile4_kernel (a plain permute lost
/// 2%).
template <typename T>
struct VmQuad
{
	T a, b, c, d;
};

template <int SEL, typename T>
__device__ __forceinline__ T vm_bulky (T acc, T x)
{
	if (VM_S_POW == SEL) { return m_pow(acc, x); }
	if (VM_S_RPOW == SEL) { return m_pow(x, acc); }
	if (VM_S_TAN == SEL) { return m_tan(acc); }
	if (VM_S_SIN == SEL) { return m_sin(acc); }
	if (VM_S_COS == SEL) { return m_cos(acc); }
	if (VM_S_DIV == SEL) { return m_div(acc, x); }
	if (VM_S_RDIV == SEL) { return m_div(x, acc); }
	if (VM_S_SQRT == SEL) { return m_sqrt(acc); }
	if (VM_S_ROUND == SEL) { return m_round(acc); }
	return m_log(acc);
}

template <int SEL, typename T>
__device__ __noinline__ VmQuad<T> vm_outlined (VmQuad<T> acc, VmQuad<T> x)
{
	VmQuad<T> r;
	r.a = vm_bulky<SEL,T>(acc.a, x.a);
	r.b = vm_bulky<SEL,T>(acc.b, x.b);
	r.c = vm_bulky<SEL,T>(acc.c, x.c);
	r.d = vm_bulky<SEL,T>(acc.d, x.d);
	return r;
}

/// (unary operators have no operand: they are handed the accumulator twice so
/// that the unloaded operand registers stay dead)
#define LLO_VM_BULKY(SEL) if constexpr (COMPACT)\
	{\
		_Pragma("unroll")\
		for (int q = 0; q < V; q += 4)\
		{\
			VmQuad<T> a4 = {acc[q], acc[q + 1], acc[q + 2], acc[q + 3]};\
			VmQuad<T> x4 = a4;\
			if constexpr (VM_S_POW == SEL || VM_S_RPOW == SEL || VM_S_DIV == SEL || VM_S_RDIV == SEL)\
			{\
				x4 = VmQuad<T>{x[q], x[q + 1], x[q + 2], x[q + 3]};\
			}\
			VmQuad<T> r4 = vm_outlined<SEL,T>(a4, x4);\
			acc[q] = r4.a; acc[q + 1] = r4.b; acc[q + 2] = r4.c; acc[q + 3] = r4.d;\
		}\
	}\
	else if constexpr (VM_S_POW == SEL || VM_S_RPOW == SEL || VM_S_DIV == SEL || VM_S_RDIV == SEL)\
	{\
		LLO_VM_EACH(acc[e] = (vm_bulky<SEL,T>(acc[e], x[e])))\
	}\
	else\
	{\
		LLO_VM_EACH(acc[e] = (vm_bulky<SEL,T>(acc[e], acc[e])))\
	}

/// acc <- f(acc, x) for one accumulator-form operation (PUSH / POP are the
/// caller's business)
template <int SEL, typename T, bool COMPACT>
__device__ __forceinline__ void vm_apply (T (&acc)[VmGeom<T>::V],
	T (&x)[VmGeom<T>::V])
{
	constexpr int V = VmGeom<T>::V;
	if constexpr (VM_S_SET == SEL) { LLO_VM_EACH(acc[e] = x[e]) }
	else if constexpr (VM_S_ABS == SEL) { LLO_VM_EACH(acc[e] = m_abs(acc[e])) }
	else if constexpr (VM_S_NEG == SEL) { LLO_VM_EACH(acc[e] = m_neg(acc[e])) }
	else if constexpr (VM_S_ROUND == SEL) { LLO_VM_BULKY(VM_S_ROUND) }
	else if constexpr (VM_S_SIN == SEL) { LLO_VM_BULKY(VM_S_SIN) }
	else if constexpr (VM_S_COS == SEL) { LLO_VM_BULKY(VM_S_COS) }
	else if constexpr (VM_S_TAN == SEL) { LLO_VM_BULKY(VM_S_TAN) }
	else if constexpr (VM_S_EXP == SEL) { LLO_VM_EACH(acc[e] = m_exp(acc[e])) }
	else if constexpr (VM_S_LOG == SEL) { LLO_VM_BULKY(VM_S_LOG) }
	else if constexpr (VM_S_SQRT == SEL) { LLO_VM_BULKY(VM_S_SQRT) }
	else if constexpr (VM_S_SUM == SEL) { LLO_VM_EACH(acc[e] = m_add(acc[e], x[e])) }
	else if constexpr (VM_S_PROD == SEL) { LLO_VM_EACH(acc[e] = m_mul(acc[e], x[e])) }
	else if constexpr (VM_S_SUB == SEL) { LLO_VM_EACH(acc[e] = m_sub(acc[e], x[e])) }
	else if constexpr (VM_S_RSUB == SEL) { LLO_VM_EACH(acc[e] = m_sub(x[e], acc[e])) }
	else if constexpr (VM_S_MIN == SEL) { LLO_VM_EACH(acc[e] = m_min(acc[e], x[e])) }
	else if constexpr (VM_S_RMIN == SEL) { LLO_VM_EACH(acc[e] = m_min(x[e], acc[e])) }
	else if constexpr (VM_S_MAX == SEL) { LLO_VM_EACH(acc[e] = m_max(acc[e], x[e])) }
	else if constexpr (VM_S_RMAX == SEL) { LLO_VM_EACH(acc[e] = m_max(x[e], acc[e])) }
	else if constexpr (VM_S_EQ == SEL) { LLO_VM_EACH(acc[e] = (T) (acc[e] == x[e])) }
	else if constexpr (VM_S_NEQ == SEL) { LLO_VM_EACH(acc[e] = (T) (acc[e] != x[e])) }
	else if constexpr (VM_S_LT == SEL) { LLO_VM_EACH(acc[e] = (T) (acc[e] < x[e])) }
	else if constexpr (VM_S_GT == SEL) { LLO_VM_EACH(acc[e] = (T) (acc[e] > x[e])) }
	else if constexpr (VM_S_DIV == SEL) { LLO_VM_BULKY(VM_S_DIV) }
	else if constexpr (VM_S_RDIV == SEL) { LLO_VM_BULKY(VM_S_RDIV) }
	else if constexpr (VM_S_POW == SEL) { LLO_VM_BULKY(VM_S_POW) }
	else if constexpr (VM_S_RPOW == SEL) { LLO_VM_BULKY(VM_S_RPOW) }
}

#ifdef LLO_VM_STATIC

/// Specialised build: instruction PC of the compile-time program
/// (LLO_VM_S_SEL / LLO_VM_S_ARG and the per-input kind tables are emitted by
/// specialize.cu in front of this file).  Straight-line code: no dispatch, no
/// accumulator copies, operand-kind branches resolved, the spill depth a
/// template argument.
template <int PC, int DEPTH, typename T, typename Loader>
__device__ __forceinline__ void vm_step (T (&acc)[VmGeom<T>::V],
	const VmParams& p, const Loader& loader, T* spill)
{
	if constexpr (PC < LLO_VM_S_NINSNS)
	{
		constexpr int V = VmGeom<T>::V;
		constexpr int sel = LLO_VM_S_SEL[PC];
		constexpr int arg = LLO_VM_S_ARG[PC];
		constexpr int src = arg >> 4;
		constexpr int k = arg & 15;
		T x[V];
		if constexpr (VM_SRC_IN == src)
		{
			loader.load(x, k, VmKind{LLO_VM_S_MODE[k], LLO_VM_S_S0[k],
				LLO_VM_S_S1[k], LLO_VM_S_VEC[k]});
		}
		else if constexpr (VM_SRC_CONST == src)
		{
			T c;
			memcpy(&c, &p.consts[k], sizeof(T));
			LLO_VM_EACH(x[e] = c)
		}
		else if constexpr (VM_SRC_POP == src)
		{
			const T* sp = spill + (size_t) (DEPTH - 1) * V * VM_THREADS + threadIdx.x;
			LLO_VM_EACH(x[e] = sp[e * VM_THREADS])
		}
		if constexpr (VM_S_PUSH == sel)
		{
			T* sp = spill + (size_t) DEPTH * V * VM_THREADS + threadIdx.x;
			LLO_VM_EACH(sp[e * VM_THREADS] = acc[e])
		}
		else
		{
			vm_apply<sel,T,false>(acc, x);
		}
		vm_step<PC + 1, DEPTH + (VM_S_PUSH == sel ? 1 : 0) -
			(VM_SRC_POP == src ? 1 : 0), T, Loader>(acc, p, loader, spill);
	}
}

template <typename T, typename Loader, bool COMPACT = false>
__device__ __forceinline__ void vm_exec (T (&acc)[VmGeom<T>::V],
	const VmParams& p, const Loader& loader, T* spill)
{
	vm_step<0, 0, T, Loader>(acc, p, loader, spill);
}

#else // interpreted

/// Run the accumulator program for this thread's V outputs.  `spill` is the
/// CTA's shared-memory spill area ([level][element][thread]).
template <typename T, typename Loader, bool COMPACT = false>
__device__ __forceinline__ void vm_exec (T (&acc)[VmGeom<T>::V],
	const VmParams& p, const Loader& loader, T* spill)
{
	constexpr int V = VmGeom<T>::V;
	int depth = 0;
	// every program starts by setting the accumulator (vm_encode): peeled,
	// so inside the loop the accumulator is only ever updated in place
	{
		int arg = p.code[0].arg;
		if (VM_SRC_IN == (arg >> 4))
		{
			loader.load(acc, arg & 15, vm_kind(p, arg & 15));
		}
		else
		{
			T c;
			memcpy(&c, &p.consts[arg & 15], sizeof(T));
			LLO_VM_EACH(acc[e] = c)
		}
	}
#pragma unroll 1
	for (int pc = 1; pc < p.ninsns; ++pc)
	{
		int sel = p.code[pc].sel;
		int arg = p.code[pc].arg;
		int src = arg >> 4;
		T x[V];
		if (VM_SRC_IN == src)
		{
			loader.load(x, arg & 15, vm_kind(p, arg & 15));
		}
		else if (VM_SRC_CONST == src)
		{
			T c;
			memcpy(&c, &p.consts[arg & 15], sizeof(T));
			LLO_VM_EACH(x[e] = c)
		}
		else if (VM_SRC_POP == src)
		{
			--depth;
			const T* sp = spill + (size_t) depth * V * VM_THREADS + threadIdx.x;
			LLO_VM_EACH(x[e] = sp[e * VM_THREADS])
		}
		switch (sel)
		{
			case VM_S_PUSH:
			{
				T* sp = spill + (size_t) depth * V * VM_THREADS + threadIdx.x;
				LLO_VM_EACH(sp[e * VM_THREADS] = acc[e])
				++depth;
			}
				break;
#define LLO_VM_CASE(SEL) case SEL: vm_apply<SEL,T,COMPACT>(acc, x); break;
			LLO_VM_CASE(VM_S_SET) LLO_VM_CASE(VM_S_ABS) LLO_VM_CASE(VM_S_NEG)
			LLO_VM_CASE(VM_S_ROUND) LLO_VM_CASE(VM_S_SIN) LLO_VM_CASE(VM_S_COS)
			LLO_VM_CASE(VM_S_TAN) LLO_VM_CASE(VM_S_EXP) LLO_VM_CASE(VM_S_LOG)
			LLO_VM_CASE(VM_S_SQRT) LLO_VM_CASE(VM_S_SUM) LLO_VM_CASE(VM_S_PROD)
			LLO_VM_CASE(VM_S_SUB) LLO_VM_CASE(VM_S_RSUB) LLO_VM_CASE(VM_S_MIN)
			LLO_VM_CASE(VM_S_RMIN) LLO_VM_CASE(VM_S_MAX) LLO_VM_CASE(VM_S_RMAX)
			LLO_VM_CASE(VM_S_EQ) LLO_VM_CASE(VM_S_NEQ) LLO_VM_CASE(VM_S_LT)
			LLO_VM_CASE(VM_S_GT) LLO_VM_CASE(VM_S_DIV) LLO_VM_CASE(VM_S_RDIV)
			LLO_VM_CASE(VM_S_POW) LLO_VM_CASE(VM_S_RPOW)
#undef LLO_VM_CASE
			default: break;
		}
	}
}

#endif // LLO_VM_STATIC

#ifndef __CUDACC_RTC__
/// Host side: validate and launch (elementwise.cu)
int vm_launch (llo_cuda_ctx* ctx, int dtype, void* out,
	const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs,
	const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns);

/// Fill the shape-dependent part of VmParams (collapsing compatible dims);
/// shared with the random fills.  `chunk` = elements per 128-bit chunk,
/// `span` = outputs per CTA (rows_aligned / row_uniform are relative to them)
int vm_prepare (VmParams& p, int dtype, const uint64_t outshape[LLO_RANK],
	const VmInput* inputs, int ninputs, const void* consts, int nconsts,
	const llo_cuda_insn* program, int ninsns, int chunk, int span);
#endif // __CUDACC_RTC__

}

#endif // LLO_CUDA_VM_CUH
