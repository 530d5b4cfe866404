///
/// comm.cu — K8: the one exchange step of batch-sharded evaluation, the
/// allreduce of the packed gradient buffer over NVLink 5 / NVSwitch.
///
/// The reference has no communication layer at all; this is
/// the collective the north star adds.  NCCL is bound at run time
/// (dlopen of libnccl.so.2: the copy torch already loaded when the host is a
/// torch.distributed process, the system one otherwise), so the library has
/// no link-time dependency and single-GPU users never touch it.  The
/// communicator belongs to the context and its collectives are enqueued on
/// the context's stream, in order with the kernels that produce the buffer.
///

#include <dlfcn.h>

#include <cstdlib>

#include "common.cuh"

namespace llo_cuda
{

namespace
{

// the slice of nccl.h this file needs (NCCL 2.x ABI)
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
typedef int NcclResult;
enum { NCCL_SUM = 0, NCCL_PROD = 1, NCCL_MAX = 2, NCCL_MIN = 3 };
enum
{
	NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_UINT32 = 3,
	NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8,
};

struct NcclApi
{
	NcclResult (*get_unique_id) (NcclUniqueId*) = nullptr;
	NcclResult (*comm_init_rank) (NcclComm*, int, NcclUniqueId, int) = nullptr;
	NcclResult (*comm_destroy) (NcclComm) = nullptr;
	NcclResult (*all_reduce) (const void*, void*, size_t, int, int, NcclComm,
		cudaStream_t) = nullptr;
	NcclResult (*all_gather) (const void*, void*, size_t, int, NcclComm,
		cudaStream_t) = nullptr;
	const char* (*get_error_string) (NcclResult) = nullptr;
	bool ok = false;
	std::string error;
};

NcclApi& nccl (void)
{
	static NcclApi api;
	static bool tried = false;
	if (tried)
	{
		return api;
	}
	tried = true;
	void* handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (nullptr == handle)
	{
		handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	}
	if (nullptr == handle)
	{
		api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
		return api;
	}
	api.get_unique_id = (decltype(api.get_unique_id)) dlsym(handle, "ncclGetUniqueId");
	api.comm_init_rank = (decltype(api.comm_init_rank)) dlsym(handle, "ncclCommInitRank");
	api.comm_destroy = (decltype(api.comm_destroy)) dlsym(handle, "ncclCommDestroy");
	api.all_reduce = (decltype(api.all_reduce)) dlsym(handle, "ncclAllReduce");
	api.all_gather = (decltype(api.all_gather)) dlsym(handle, "ncclAllGather");
	api.get_error_string = (decltype(api.get_error_string)) dlsym(handle, "ncclGetErrorString");
	api.ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy &&
		api.all_reduce && api.all_gather && api.get_error_string;
	if (false == api.ok)
	{
		api.error = "libnccl.so.2 lacks an expected symbol";
	}
	return api;
}

int fail_nccl (NcclResult rc, const char* what)
{
	return fail(LLO_ERR_CUDA, "NCCL error %d (%s) at %s", (int) rc,
		nccl().get_error_string ? nccl().get_error_string(rc) : "?", what);
}

int nccl_dtype (int dtype)
{
	switch (dtype)
	{
		case LLO_DT_DOUBLE: return NCCL_FLOAT64;
		case LLO_DT_FLOAT: return NCCL_FLOAT32;
		case LLO_DT_INT8: return NCCL_INT8;
		case LLO_DT_UINT8: return NCCL_UINT8;
		case LLO_DT_INT32: return NCCL_INT32;
		case LLO_DT_UINT32: return NCCL_UINT32;
		case LLO_DT_INT64: return NCCL_INT64;
		case LLO_DT_UINT64: return NCCL_UINT64;
		default: return -1; // 16-bit integers have no NCCL type
	}
}

/// out[i] = fold over r of gathered[r * n + i], r ascending: the same value
/// on every rank whatever algorithm NCCL would have picked
template <typename T>
__global__ void __launch_bounds__(256)
ordered_fold_kernel (T* __restrict__ out, const T* __restrict__ gathered,
	uint64_t n, int nranks, int opcode)
{
	uint64_t i = (uint64_t) blockIdx.x * 256 + threadIdx.x;
	if (i >= n)
	{
		return;
	}
	T acc = gathered[i];
	for (int r = 1; r < nranks; ++r)
	{
		T v = gathered[(uint64_t) r * n + i];
		switch (opcode)
		{
			case LLO_OP_SUM: acc = acc + v; break;
			case LLO_OP_PROD: acc = acc * v; break;
			case LLO_OP_MIN: acc = (v < acc) ? v : acc; break;
			default: acc = (acc < v) ? v : acc; break;
		}
	}
	out[i] = acc;
}

}

}

using namespace llo_cuda;

extern "C"
{

int llo_cuda_comm_unique_id (void* id128)
{
	NcclApi& api = nccl();
	if (false == api.ok)
	{
		return fail(LLO_ERR_CUDA, "%s", api.error.c_str());
	}
	NcclUniqueId id;
	NcclResult rc = api.get_unique_id(&id);
	if (0 != rc)
	{
		return fail_nccl(rc, "ncclGetUniqueId");
	}
	memcpy(id128, &id, sizeof(id));
	return LLO_OK;
}

int llo_cuda_comm_init (llo_cuda_ctx* ctx, int rank, int nranks,
	const void* id128)
{
	if (nranks < 1 || rank < 0 || rank >= nranks)
	{
		return fail(LLO_ERR_FATAL, "rank %d of %d is not a valid position",
			rank, nranks);
	}
	if (nullptr != ctx->comm)
	{
		return fail(LLO_ERR_FATAL, "the context already has a communicator");
	}
	ctx->comm_rank = rank;
	ctx->comm_size = nranks;
	if (1 == nranks)
	{
		return LLO_OK; // nothing to exchange with
	}
	NcclApi& api = nccl();
	if (false == api.ok)
	{
		return fail(LLO_ERR_CUDA, "%s", api.error.c_str());
	}
	LLO_CUDA_TRY(cudaSetDevice(ctx->device));
	NcclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	NcclComm comm = nullptr;
	NcclResult rc = api.comm_init_rank(&comm, nranks, id, rank);
	if (0 != rc)
	{
		return fail_nccl(rc, "ncclCommInitRank");
	}
	ctx->comm = comm;
	return LLO_OK;
}

int llo_cuda_comm_destroy (llo_cuda_ctx* ctx)
{
	if (nullptr != ctx->comm_stream)
	{
		cudaStreamSynchronize(ctx->comm_stream);
		cudaStreamDestroy(ctx->comm_stream);
		cudaEventDestroy(ctx->comm_ready);
		cudaEventDestroy(ctx->comm_done);
		ctx->comm_stream = nullptr;
		ctx->comm_pending = false;
	}
	if (nullptr != ctx->comm)
	{
		cudaStreamSynchronize(ctx->stream);
		nccl().comm_destroy((NcclComm) ctx->comm);
		ctx->comm = nullptr;
	}
	ctx->comm_size = 1;
	ctx->comm_rank = 0;
	return LLO_OK;
}

int llo_cuda_comm_rank (llo_cuda_ctx* ctx)
{
	return ctx->comm_rank;
}

int llo_cuda_comm_size (llo_cuda_ctx* ctx)
{
	return ctx->comm_size;
}

int llo_cuda_allreduce_async (llo_cuda_ctx* ctx, void* buf, uint64_t n,
	int dtype, int opcode)
{
	if (false == is_nnary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "cannot allreduce with %s", op_name(opcode));
	}
	if (ctx->comm_size <= 1 || 0 == n)
	{
		return LLO_OK;
	}
	int ntype = nccl_dtype(dtype);
	if (ntype < 0)
	{
		return fail(LLO_ERR_FATAL, "cannot allreduce %s", dtype_name(dtype));
	}
	if (nullptr == ctx->comm)
	{
		return fail(LLO_ERR_FATAL, "the context has no communicator "
			"(llo_cuda_comm_init)");
	}
	if (nullptr == ctx->comm_stream)
	{
		LLO_CUDA_TRY(cudaSetDevice(ctx->device));
		// highest priority: the collective's few CTAs take the first SM slots
		// that free up instead of queueing behind every CTA of the GEMM
		// launched beside it (LLO_COMM_PRIORITY=0: experiments only)
		int lowest = 0, highest = 0;
		LLO_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lowest, &highest));
		const char* env = std::getenv("LLO_COMM_PRIORITY");
		int priority = (nullptr != env && 0 == std::atoi(env)) ? lowest : highest;
		LLO_CUDA_TRY(cudaStreamCreateWithPriority(&ctx->comm_stream,
			cudaStreamNonBlocking, priority));
		LLO_CUDA_TRY(cudaEventCreateWithFlags(&ctx->comm_ready, cudaEventDisableTiming));
		LLO_CUDA_TRY(cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
	}
	// the collective starts when everything queued on the stream so far (the
	// kernels that produced `buf`) has finished, and runs beside what is
	// queued afterwards
	LLO_CUDA_TRY(cudaEventRecord(ctx->comm_ready, ctx->stream));
	LLO_CUDA_TRY(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ready, 0));
	int nop = LLO_OP_SUM == opcode ? NCCL_SUM : (LLO_OP_PROD == opcode ?
		NCCL_PROD : (LLO_OP_MIN == opcode ? NCCL_MIN : NCCL_MAX));
	NcclResult rc = nccl().all_reduce(buf, buf, n, ntype, nop,
		(NcclComm) ctx->comm, ctx->comm_stream);
	if (0 != rc)
	{
		return fail_nccl(rc, "ncclAllReduce");
	}
	ctx->comm_pending = true;
	return LLO_OK;
}

int llo_cuda_comm_join (llo_cuda_ctx* ctx)
{
	if (ctx->comm_pending)
	{
		LLO_CUDA_TRY(cudaEventRecord(ctx->comm_done, ctx->comm_stream));
		LLO_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0));
		ctx->comm_pending = false;
	}
	return LLO_OK;
}

int llo_cuda_allreduce (llo_cuda_ctx* ctx, void* buf, uint64_t n, int dtype,
	int opcode, int mode)
{
	if (false == is_nnary_op(opcode))
	{
		return fail(LLO_ERR_FATAL, "cannot allreduce with %s", op_name(opcode));
	}
	if (ctx->comm_size <= 1 || 0 == n)
	{
		return LLO_OK;
	}
	int ntype = nccl_dtype(dtype);
	if (ntype < 0)
	{
		return fail(LLO_ERR_FATAL, "cannot allreduce %s", dtype_name(dtype));
	}
	if (nullptr == ctx->comm)
	{
		return fail(LLO_ERR_FATAL, "the context has no communicator "
			"(llo_cuda_comm_init)");
	}
	NcclApi& api = nccl();
	if (LLO_ALLREDUCE_ORDERED == mode)
	{
		// all-gather, then fold in rank order: bit-identical on every rank
		Scratch scratch(ctx);
		void* gathered = nullptr;
		size_t esize = dtype_size(dtype);
		LLO_TRY(scratch.get(&gathered, (size_t) ctx->comm_size * n * esize));
		NcclResult rc = api.all_gather(buf, gathered, n, ntype,
			(NcclComm) ctx->comm, ctx->stream);
		if (0 != rc)
		{
			return fail_nccl(rc, "ncclAllGather");
		}
		unsigned blocks = (unsigned) ((n + 255) / 256);
		switch (dtype)
		{
#define LLO_FOLD(DT, T) case DT: ordered_fold_kernel<T><<<blocks, 256, 0,\
			ctx->stream>>>((T*) buf, (const T*) gathered, n, ctx->comm_size, opcode); break;
			LLO_FOLD(LLO_DT_DOUBLE, double)
			LLO_FOLD(LLO_DT_FLOAT, float)
			LLO_FOLD(LLO_DT_INT8, int8_t)
			LLO_FOLD(LLO_DT_UINT8, uint8_t)
			LLO_FOLD(LLO_DT_INT32, int32_t)
			LLO_FOLD(LLO_DT_UINT32, uint32_t)
			LLO_FOLD(LLO_DT_INT64, int64_t)
			LLO_FOLD(LLO_DT_UINT64, uint64_t)
#undef LLO_FOLD
			default: break;
		}
		return launched(ctx, "ordered_fold_kernel");
	}
	int nop = LLO_OP_SUM == opcode ? NCCL_SUM : (LLO_OP_PROD == opcode ?
		NCCL_PROD : (LLO_OP_MIN == opcode ? NCCL_MIN : NCCL_MAX));
	NcclResult rc = api.all_reduce(buf, buf, n, ntype, nop,
		(NcclComm) ctx->comm, ctx->stream);
	if (0 != rc)
	{
		return fail_nccl(rc, "ncclAllReduce");
	}
	return LLO_OK;
}

}
