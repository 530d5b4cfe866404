///
/// common.cuh — shared declarations of libllo_cuda.so (sm_100a only).
///

#ifndef LLO_CUDA_COMMON_CUH
#define LLO_CUDA_COMMON_CUH

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "llo_cuda.h"

/// Context: one device, one stream, one calling thread
struct llo_cuda_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	bool own_stream = true;
	int sm_count = 148;
	uint64_t launches = 0;
	// Philox key / running counter offset (K9)
	uint64_t seed = 0;
	uint64_t rand_offset = 0;
	int reduce_mode = LLO_REDUCE_TREE;
	// device-side error flag raised by kernels that validate coordinates
	// ([0]); [1], [2]: work-unit counter / finished CTAs of the persistent tile
	// kernels (self-resetting; one such kernel at a time per context)
	int* dev_error = nullptr;
	// copy-engine lanes of llo_cuda_h2d_async / llo_cuda_d2h_async (created on
	// first use) and the events that order them against `stream`
	cudaStream_t up_stream = nullptr;
	cudaStream_t down_stream = nullptr;
	cudaEvent_t lane_fence = nullptr;
	// snapshots of dev_error[0] taken on the download lane right after each
	// asynchronous result copy (pinned), checked by llo_cuda_event_sync
	int* async_flags = nullptr;
	static constexpr int async_slots = 256;
	uint64_t async_next = 0;
	std::vector<std::pair<cudaEvent_t,int>> async_pending;
	// scratch arena of the GEMM drivers (struct Scratch): one grow-only block
	// reused by every call -- the calls are ordered on `stream`, so handing
	// the same bytes to the next call is safe, and nothing is allocated or
	// freed per call once the high-water mark has been seen
	char* arena = nullptr;
	size_t arena_cap = 0;
	size_t arena_used = 0;     // bytes handed out from the block
	size_t arena_asked = 0;    // bytes asked for by the open scopes (may exceed cap)
	size_t arena_want = 0;     // high-water mark of arena_asked
	int arena_depth = 0;
	// K8: NCCL communicator of the batch-sharded executor (comm.cu)
	void* comm = nullptr;
	int comm_rank = 0;
	int comm_size = 1;
	// second lane for collectives that overlap later compute
	// (llo_cuda_allreduce_async / llo_cuda_comm_join)
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t comm_ready = nullptr;   // recorded on `stream`: the bucket is complete
	cudaEvent_t comm_done = nullptr;    // recorded on comm_stream: collectives so far are done
	bool comm_pending = false;
	// run-time specialisation of the elementwise engine (specialize.cu):
	// launches of at least spec_min output elements use the NVRTC build
	uint64_t spec_min = 1ull << 20;
	uint64_t spec_launches = 0;   // launches that ran a specialised kernel
	uint64_t spec_compiles = 0;   // NVRTC compilations (cache misses)
	uint64_t spec_disk_hits = 0;  // cubins taken from the on-disk cache
	uint64_t spec_failures = 0;   // builds that failed (interpreted kernel ran)
};

namespace llo_cuda
{

/// Record an error message for llo_cuda_last_error and return `code`
int fail (int code, const char* fmt, ...);

/// Record a CUDA runtime failure
int fail_cuda (cudaError_t err, const char* what);

#define LLO_CUDA_TRY(expr) do {\
	cudaError_t err__ = (expr);\
	if (cudaSuccess != err__) { return llo_cuda::fail_cuda(err__, #expr); }\
} while (0)

#define LLO_TRY(expr) do {\
	int rc__ = (expr);\
	if (LLO_OK != rc__) { return rc__; }\
} while (0)

/// Call after every kernel launch: counts it and surfaces launch errors
int launched (llo_cuda_ctx* ctx, const char* kernel);

inline size_t dtype_size (int dtype)
{
	switch (dtype)
	{
		case LLO_DT_DOUBLE: case LLO_DT_INT64: case LLO_DT_UINT64: return 8;
		case LLO_DT_FLOAT: case LLO_DT_INT32: case LLO_DT_UINT32: return 4;
		case LLO_DT_INT16: case LLO_DT_UINT16: return 2;
		case LLO_DT_INT8: case LLO_DT_UINT8: return 1;
		default: return 0;
	}
}

inline const char* dtype_name (int dtype)
{
	static const char* names[] = {"BAD_TYPE", "DOUBLE", "FLOAT", "INT8",
		"UINT8", "INT16", "UINT16", "INT32", "UINT32", "INT64", "UINT64"};
	return (dtype > 0 && dtype < LLO_DT_COUNT) ? names[dtype] : names[0];
}

inline const char* op_name (int op)
{
	static const char* names[] = {"BAD_OP", "ABS", "NEG", "SIN", "COS", "TAN",
		"EXP", "LOG", "SQRT", "ROUND", "POW", "SUM", "SUB", "PROD", "DIV",
		"MIN", "MAX", "EQ", "NEQ", "LT", "GT", "RAND_BINO", "RAND_UNIF",
		"RAND_NORM"};
	return (op > 0 && op < LLO_OP_COUNT) ? names[op] : names[0];
}

inline bool is_unsigned_dtype (int dtype)
{
	return LLO_DT_UINT8 == dtype || LLO_DT_UINT16 == dtype ||
		LLO_DT_UINT32 == dtype || LLO_DT_UINT64 == dtype;
}

inline bool is_float_dtype (int dtype)
{
	return LLO_DT_DOUBLE == dtype || LLO_DT_FLOAT == dtype;
}

inline bool is_unary_op (int op)
{
	return op >= LLO_OP_ABS && op <= LLO_OP_ROUND;
}

inline bool is_binary_op (int op)
{
	return LLO_OP_POW == op || LLO_OP_SUB == op || LLO_OP_DIV == op ||
		(op >= LLO_OP_EQ && op <= LLO_OP_GT);
}

inline bool is_nnary_op (int op)
{
	return LLO_OP_SUM == op || LLO_OP_PROD == op ||
		LLO_OP_MIN == op || LLO_OP_MAX == op;
}

inline bool is_rand_op (int op)
{
	return op >= LLO_OP_RAND_BINO && op <= LLO_OP_RAND_NORM;
}

inline uint64_t n_elems (const uint64_t shape[LLO_RANK])
{
	uint64_t n = 1;
	for (int i = 0; i < LLO_RANK; ++i)
	{
		n *= shape[i];
	}
	return n;
}

/// Dispatch a generic lambda on the C type of `dtype`:
/// LLO_DISPATCH_DTYPE(dtype, T, { kernel<T><<<...>>>(...); })
#ifdef LLO_DEV_FLOAT_ONLY
// development builds only (fast ptxas / SASS turnaround): float kernels
#define LLO_DISPATCH_DTYPE(DTYPE, T, ...)\
	switch (DTYPE)\
	{\
		case LLO_DT_FLOAT: { using T = float; __VA_ARGS__ } break;\
		default: return llo_cuda::fail(LLO_ERR_FATAL, "executing bad type %d", (int) (DTYPE));\
	}
#else
#define LLO_DISPATCH_DTYPE(DTYPE, T, ...)\
	switch (DTYPE)\
	{\
		case LLO_DT_DOUBLE: { using T = double; __VA_ARGS__ } break;\
		case LLO_DT_FLOAT: { using T = float; __VA_ARGS__ } break;\
		case LLO_DT_INT8: { using T = int8_t; __VA_ARGS__ } break;\
		case LLO_DT_UINT8: { using T = uint8_t; __VA_ARGS__ } break;\
		case LLO_DT_INT16: { using T = int16_t; __VA_ARGS__ } break;\
		case LLO_DT_UINT16: { using T = uint16_t; __VA_ARGS__ } break;\
		case LLO_DT_INT32: { using T = int32_t; __VA_ARGS__ } break;\
		case LLO_DT_UINT32: { using T = uint32_t; __VA_ARGS__ } break;\
		case LLO_DT_INT64: { using T = int64_t; __VA_ARGS__ } break;\
		case LLO_DT_UINT64: { using T = uint64_t; __VA_ARGS__ } break;\
		default: return llo_cuda::fail(LLO_ERR_FATAL, "executing bad type %d", (int) (DTYPE));\
	}
#endif

/// Scratch device allocation released (stream-ordered) at scope exit
struct Scratch final
{
	Scratch (llo_cuda_ctx* ctx) : ctx_(ctx),
		used_mark_(ctx->arena_used), asked_mark_(ctx->arena_asked)
	{
		++ctx_->arena_depth;
	}

	~Scratch (void)
	{
		for (void* p : ptrs_)
		{
			cudaFreeAsync(p, ctx_->stream);
		}
		ctx_->arena_used = used_mark_;
		ctx_->arena_asked = asked_mark_;
		if (0 == --ctx_->arena_depth && ctx_->arena_want > ctx_->arena_cap)
		{
			// nobody holds arena pointers now; work already queued keeps
			// the old block until it has run (stream-ordered free)
			if (nullptr != ctx_->arena)
			{
				cudaFreeAsync(ctx_->arena, ctx_->stream);
				ctx_->arena = nullptr;
				ctx_->arena_cap = 0;
			}
			void* block = nullptr;
			size_t cap = ctx_->arena_want + ctx_->arena_want / 8;
			if (cudaSuccess == cudaMallocAsync(&block, cap, ctx_->stream))
			{
				ctx_->arena = (char*) block;
				ctx_->arena_cap = cap;
			}
			else
			{
				cudaGetLastError(); // stay on per-call allocations
				ctx_->arena_want = 0;
			}
		}
	}

	int get (void** out, size_t nbytes)
	{
		nbytes = (nbytes + 255) & ~(size_t) 255;
		if (0 == nbytes)
		{
			nbytes = 256;
		}
		ctx_->arena_asked += nbytes;
		if (ctx_->arena_asked > ctx_->arena_want)
		{
			ctx_->arena_want = ctx_->arena_asked;
		}
		if (nullptr != ctx_->arena && ctx_->arena_used + nbytes <= ctx_->arena_cap)
		{
			*out = ctx_->arena + ctx_->arena_used;
			ctx_->arena_used += nbytes;
			return LLO_OK;
		}
		void* p = nullptr;
		cudaError_t err = cudaMallocAsync(&p, nbytes, ctx_->stream);
		if (cudaSuccess != err)
		{
			return fail_cuda(err, "scratch cudaMallocAsync");
		}
		ptrs_.push_back(p);
		*out = p;
		return LLO_OK;
	}

private:
	llo_cuda_ctx* ctx_;
	size_t used_mark_;
	size_t asked_mark_;
	std::vector<void*> ptrs_;
};

}

#endif // LLO_CUDA_COMMON_CUH
