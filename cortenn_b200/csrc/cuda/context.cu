///
/// context.cu — context, stream, stream-ordered memory pool and transfers.
/// The device-resident replacement for the reference's malloc/free data
/// store (llo/src/data.cpp:8-18).
///

#include <cstdlib>

#include "common.cuh"

namespace llo_cuda
{

static thread_local std::string last_error;

int fail (int code, const char* fmt, ...)
{
	char buf[1024];
	va_list args;
	va_start(args, fmt);
	vsnprintf(buf, sizeof(buf), fmt, args);
	va_end(args);
	last_error = buf;
	return code;
}

int fail_cuda (cudaError_t err, const char* what)
{
	return fail(LLO_ERR_CUDA, "CUDA error %s (%s) at %s",
		cudaGetErrorName(err), cudaGetErrorString(err), what);
}

int launched (llo_cuda_ctx* ctx, const char* kernel)
{
	++ctx->launches;
	cudaError_t err = cudaGetLastError();
	if (cudaSuccess != err)
	{
		return fail_cuda(err, kernel);
	}
	return LLO_OK;
}

}

using namespace llo_cuda;

extern "C"
{

const char* llo_cuda_last_error (void)
{
	return last_error.c_str();
}

int llo_cuda_abi_version (void)
{
	return LLO_CUDA_ABI_VERSION;
}

int llo_cuda_init (int device, llo_cuda_ctx** out)
{
	if (nullptr == out)
	{
		return fail(LLO_ERR_FATAL, "llo_cuda_init: null context pointer");
	}
	*out = nullptr;
	int ndev = 0;
	cudaError_t err = cudaGetDeviceCount(&ndev);
	if (cudaSuccess != err || 0 == ndev)
	{
		cudaGetLastError();
		return fail(LLO_ERR_CUDA, "no CUDA device available (%s)",
			cudaSuccess != err ? cudaGetErrorString(err) : "device count 0");
	}
	if (device < 0 || device >= ndev)
	{
		return fail(LLO_ERR_FATAL, "device %d out of range [0,%d)",
			device, ndev);
	}
	LLO_CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	LLO_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10)
	{
		return fail(LLO_ERR_CUDA, "device %d is sm_%d%d; this library "
			"carries sm_100a code only", device, prop.major, prop.minor);
	}
	llo_cuda_ctx* ctx = new llo_cuda_ctx();
	ctx->device = device;
	ctx->sm_count = prop.multiProcessorCount;
	err = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
	if (cudaSuccess != err)
	{
		delete ctx;
		return fail_cuda(err, "cudaStreamCreate");
	}
	// keep freed blocks cached in the pool instead of returning them to
	// the driver at every synchronisation
	cudaMemPool_t pool;
	if (cudaSuccess == cudaDeviceGetDefaultMemPool(&pool, device))
	{
		uint64_t threshold = UINT64_MAX;
		cudaMemPoolSetAttribute(pool,
			cudaMemPoolAttrReleaseThreshold, &threshold);
	}
	err = cudaMalloc(&ctx->dev_error, 4 * sizeof(int));
	if (cudaSuccess != err)
	{
		cudaStreamDestroy(ctx->stream);
		delete ctx;
		return fail_cuda(err, "cudaMalloc(error flag)");
	}
	cudaMemset(ctx->dev_error, 0, 4 * sizeof(int));
	if (const char* env = std::getenv("LLO_SPECIALIZE_MIN"))
	{
		ctx->spec_min = std::strtoull(env, nullptr, 10);
	}
	*out = ctx;
	return LLO_OK;
}

int llo_cuda_destroy (llo_cuda_ctx* ctx)
{
	if (nullptr == ctx)
	{
		return LLO_OK;
	}
	cudaSetDevice(ctx->device);
	llo_cuda_comm_destroy(ctx);
	cudaStreamSynchronize(ctx->stream);
	cudaFree(ctx->dev_error);
	if (nullptr != ctx->arena)
	{
		cudaFreeAsync(ctx->arena, ctx->stream);
		cudaStreamSynchronize(ctx->stream);
	}
	if (nullptr != ctx->lane_fence)
	{
		cudaStreamSynchronize(ctx->up_stream);
		cudaStreamSynchronize(ctx->down_stream);
		cudaStreamDestroy(ctx->up_stream);
		cudaStreamDestroy(ctx->down_stream);
		cudaEventDestroy(ctx->lane_fence);
		cudaFreeHost(ctx->async_flags);
	}
	if (ctx->own_stream)
	{
		cudaStreamDestroy(ctx->stream);
	}
	delete ctx;
	return LLO_OK;
}

int llo_cuda_set_stream (llo_cuda_ctx* ctx, void* cuda_stream)
{
	LLO_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	if (nullptr == cuda_stream)
	{
		if (false == ctx->own_stream)
		{
			LLO_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream,
				cudaStreamNonBlocking));
			ctx->own_stream = true;
		}
		return LLO_OK;
	}
	if (ctx->own_stream)
	{
		cudaStreamDestroy(ctx->stream);
	}
	ctx->stream = (cudaStream_t) cuda_stream;
	ctx->own_stream = false;
	return LLO_OK;
}

void* llo_cuda_get_stream (llo_cuda_ctx* ctx)
{
	return (void*) ctx->stream;
}

int llo_cuda_device (llo_cuda_ctx* ctx)
{
	return ctx->device;
}

int llo_cuda_sm_count (llo_cuda_ctx* ctx)
{
	return ctx->sm_count;
}

int llo_cuda_alloc (llo_cuda_ctx* ctx, size_t bytes, void** dptr)
{
	LLO_CUDA_TRY(cudaSetDevice(ctx->device));
	LLO_CUDA_TRY(cudaMallocAsync(dptr, bytes > 0 ? bytes : 16, ctx->stream));
	return LLO_OK;
}

int llo_cuda_free (llo_cuda_ctx* ctx, void* dptr)
{
	if (nullptr == dptr)
	{
		return LLO_OK;
	}
	LLO_CUDA_TRY(cudaFreeAsync(dptr, ctx->stream));
	return LLO_OK;
}

int llo_cuda_host_alloc (llo_cuda_ctx* ctx, size_t bytes, void** hptr)
{
	LLO_CUDA_TRY(cudaSetDevice(ctx->device));
	LLO_CUDA_TRY(cudaHostAlloc(hptr, bytes > 0 ? bytes : 16,
		cudaHostAllocDefault));
	return LLO_OK;
}

int llo_cuda_host_free (llo_cuda_ctx*, void* hptr)
{
	if (nullptr != hptr)
	{
		LLO_CUDA_TRY(cudaFreeHost(hptr));
	}
	return LLO_OK;
}

int llo_cuda_h2d (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	LLO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes,
		cudaMemcpyHostToDevice, ctx->stream));
	// the host buffer may be rewritten as soon as this returns
	LLO_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	return LLO_OK;
}

static int check_device_flag (llo_cuda_ctx* ctx)
{
	int flag = 0;
	LLO_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->dev_error, sizeof(int),
		cudaMemcpyDeviceToHost, ctx->stream));
	LLO_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
	if (0 != flag)
	{
		cudaMemsetAsync(ctx->dev_error, 0, sizeof(int), ctx->stream);
		return fail(LLO_ERR_FATAL, "cannot get index of bad coordinate "
			"(a coordinate map left its tensor's shape)");
	}
	return LLO_OK;
}

int llo_cuda_d2h (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	LLO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes,
		cudaMemcpyDeviceToHost, ctx->stream));
	return check_device_flag(ctx);
}

static int ensure_lanes (llo_cuda_ctx* ctx)
{
	if (nullptr != ctx->lane_fence)
	{
		return LLO_OK;
	}
	LLO_CUDA_TRY(cudaSetDevice(ctx->device));
	LLO_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
	LLO_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
	LLO_CUDA_TRY(cudaEventCreateWithFlags(&ctx->lane_fence, cudaEventDisableTiming));
	LLO_CUDA_TRY(cudaHostAlloc((void**) &ctx->async_flags,
		llo_cuda_ctx::async_slots * sizeof(int), cudaHostAllocDefault));
	std::memset(ctx->async_flags, 0, llo_cuda_ctx::async_slots * sizeof(int));
	return LLO_OK;
}

int llo_cuda_h2d_async (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	LLO_TRY(ensure_lanes(ctx));
	// the upload lane joins the stream (dst was allocated in stream order) ...
	LLO_CUDA_TRY(cudaEventRecord(ctx->lane_fence, ctx->stream));
	LLO_CUDA_TRY(cudaStreamWaitEvent(ctx->up_stream, ctx->lane_fence, 0));
	LLO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes,
		cudaMemcpyHostToDevice, ctx->up_stream));
	// ... and the stream's later work waits for the upload
	LLO_CUDA_TRY(cudaEventRecord(ctx->lane_fence, ctx->up_stream));
	LLO_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->lane_fence, 0));
	return LLO_OK;
}

int llo_cuda_d2h_async (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes,
	void** done)
{
	if (nullptr == done)
	{
		return fail(LLO_ERR_FATAL, "llo_cuda_d2h_async: null event pointer");
	}
	LLO_TRY(ensure_lanes(ctx));
	LLO_CUDA_TRY(cudaEventRecord(ctx->lane_fence, ctx->stream));
	LLO_CUDA_TRY(cudaStreamWaitEvent(ctx->down_stream, ctx->lane_fence, 0));
	LLO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes,
		cudaMemcpyDeviceToHost, ctx->down_stream));
	// the kernels that produced `src` may have raised the coordinate-error
	// flag: snapshot it behind the copy, llo_cuda_event_sync reports it
	int slot = (int) (ctx->async_next++ % llo_cuda_ctx::async_slots);
	LLO_CUDA_TRY(cudaMemcpyAsync(ctx->async_flags + slot, ctx->dev_error,
		sizeof(int), cudaMemcpyDeviceToHost, ctx->down_stream));
	cudaEvent_t ev;
	LLO_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	LLO_CUDA_TRY(cudaEventRecord(ev, ctx->down_stream));
	ctx->async_pending.emplace_back(ev, slot);
	*done = (void*) ev;
	return LLO_OK;
}

int llo_cuda_event_sync (llo_cuda_ctx* ctx, void* event)
{
	LLO_CUDA_TRY(cudaEventSynchronize((cudaEvent_t) event));
	if (nullptr == ctx)
	{
		return LLO_OK;
	}
	for (size_t i = 0; i < ctx->async_pending.size(); ++i)
	{
		if (ctx->async_pending[i].first != (cudaEvent_t) event)
		{
			continue;
		}
		int flag = ctx->async_flags[ctx->async_pending[i].second];
		ctx->async_pending.erase(ctx->async_pending.begin() + i);
		if (0 != flag)
		{
			cudaMemsetAsync(ctx->dev_error, 0, sizeof(int), ctx->stream);
			return fail(LLO_ERR_FATAL, "cannot get index of bad coordinate "
				"(a coordinate map left its tensor's shape)");
		}
		break;
	}
	return LLO_OK;
}

int llo_cuda_d2d (llo_cuda_ctx* ctx, void* dst, const void* src, size_t bytes)
{
	LLO_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes,
		cudaMemcpyDeviceToDevice, ctx->stream));
	return LLO_OK;
}

int llo_cuda_wait_stream (llo_cuda_ctx* ctx, void* cuda_stream)
{
	// work queued on the context's stream from here on starts after
	// everything queued on `cuda_stream` so far has finished
	cudaStream_t other = (cudaStream_t) cuda_stream;
	if (other == ctx->stream)
	{
		return LLO_OK;
	}
	cudaEvent_t ev;
	LLO_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	cudaError_t err = cudaEventRecord(ev, other);
	if (cudaSuccess == err)
	{
		err = cudaStreamWaitEvent(ctx->stream, ev, 0);
	}
	cudaEventDestroy(ev);
	LLO_CUDA_TRY(err);
	return LLO_OK;
}

int llo_cuda_memset (llo_cuda_ctx* ctx, void* dst, int byte, size_t bytes)
{
	LLO_CUDA_TRY(cudaMemsetAsync(dst, byte, bytes, ctx->stream));
	return LLO_OK;
}

int llo_cuda_sync (llo_cuda_ctx* ctx)
{
	if (nullptr != ctx->lane_fence)
	{
		LLO_CUDA_TRY(cudaStreamSynchronize(ctx->up_stream));
		LLO_CUDA_TRY(cudaStreamSynchronize(ctx->down_stream));
	}
	return check_device_flag(ctx);
}

uint64_t llo_cuda_launch_count (llo_cuda_ctx* ctx)
{
	return ctx->launches;
}

int llo_cuda_event_create (llo_cuda_ctx* ctx, void** event)
{
	cudaEvent_t ev;
	LLO_CUDA_TRY(cudaSetDevice(ctx->device));
	LLO_CUDA_TRY(cudaEventCreate(&ev));
	*event = (void*) ev;
	return LLO_OK;
}

int llo_cuda_event_record (llo_cuda_ctx* ctx, void* event)
{
	LLO_CUDA_TRY(cudaEventRecord((cudaEvent_t) event, ctx->stream));
	return LLO_OK;
}

int llo_cuda_event_elapsed_ms (llo_cuda_ctx*, void* start, void* stop, float* ms)
{
	LLO_CUDA_TRY(cudaEventSynchronize((cudaEvent_t) stop));
	LLO_CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t) start, (cudaEvent_t) stop));
	return LLO_OK;
}

int llo_cuda_event_destroy (llo_cuda_ctx* ctx, void* event)
{
	if (nullptr != ctx)
	{
		for (size_t i = 0; i < ctx->async_pending.size(); ++i)
		{
			if (ctx->async_pending[i].first == (cudaEvent_t) event)
			{
				ctx->async_pending.erase(ctx->async_pending.begin() + i);
				break;
			}
		}
	}
	LLO_CUDA_TRY(cudaEventDestroy((cudaEvent_t) event));
	return LLO_OK;
}

int llo_cuda_graph_begin (llo_cuda_ctx* ctx)
{
	// (relaxed: the calls made while capturing include stream-ordered
	// allocations and attribute queries that a stricter mode would refuse)
	LLO_CUDA_TRY(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed));
	return LLO_OK;
}

int llo_cuda_graph_end (llo_cuda_ctx* ctx, void** graph_exec)
{
	cudaGraph_t graph = nullptr;
	cudaError_t err = cudaStreamEndCapture(ctx->stream, &graph);
	if (cudaSuccess != err || nullptr == graph)
	{
		cudaGetLastError();
		*graph_exec = nullptr;
		return fail_cuda(cudaSuccess != err ? err : cudaErrorUnknown,
			"cudaStreamEndCapture (a call made while capturing cannot be captured)");
	}
	cudaGraphExec_t exec = nullptr;
	err = cudaGraphInstantiate(&exec, graph, 0);
	cudaGraphDestroy(graph);
	if (cudaSuccess != err)
	{
		*graph_exec = nullptr;
		return fail_cuda(err, "cudaGraphInstantiate");
	}
	*graph_exec = (void*) exec;
	return LLO_OK;
}

int llo_cuda_graph_abort (llo_cuda_ctx* ctx)
{
	cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
	cudaStreamIsCapturing(ctx->stream, &status);
	if (cudaStreamCaptureStatusNone != status)
	{
		cudaGraph_t graph = nullptr;
		cudaStreamEndCapture(ctx->stream, &graph);
		if (nullptr != graph)
		{
			cudaGraphDestroy(graph);
		}
	}
	cudaGetLastError();
	return LLO_OK;
}

int llo_cuda_graph_launch (llo_cuda_ctx* ctx, void* graph_exec, uint64_t launches)
{
	LLO_CUDA_TRY(cudaGraphLaunch((cudaGraphExec_t) graph_exec, ctx->stream));
	ctx->launches += launches;
	return LLO_OK;
}

int llo_cuda_graph_destroy (llo_cuda_ctx*, void* graph_exec)
{
	if (nullptr != graph_exec)
	{
		LLO_CUDA_TRY(cudaGraphExecDestroy((cudaGraphExec_t) graph_exec));
	}
	return LLO_OK;
}

int llo_cuda_seed (llo_cuda_ctx* ctx, uint64_t seed)
{
	ctx->seed = seed;
	ctx->rand_offset = 0;
	return LLO_OK;
}

int llo_cuda_set_reduce_mode (llo_cuda_ctx* ctx, int mode)
{
	if (LLO_REDUCE_TREE != mode && LLO_REDUCE_SEQUENTIAL != mode)
	{
		return fail(LLO_ERR_FATAL, "unknown reduce mode %d", mode);
	}
	ctx->reduce_mode = mode;
	return LLO_OK;
}

}
