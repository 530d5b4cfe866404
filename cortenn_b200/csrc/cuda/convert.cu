///
/// convert.cu — K0: leaf element-type conversion on the device.
/// Replaces GenericData::copyover / convert<T> (llo/src/data.cpp:20-72):
/// out[i] = (OUT) in[i] with C++ implicit-conversion semantics (double ->
/// integer truncates toward zero, integer narrowing wraps).
///

#include "common.cuh"

namespace llo_cuda
{

template <typename OUT, typename IN>
__device__ __forceinline__ OUT convert_one (IN v)
{
	if constexpr (std::is_floating_point<IN>::value &&
		std::is_integral<OUT>::value && sizeof(OUT) <= 4)
	{
		// in range this is plain truncation toward zero; out of range
		// (undefined in C++) x86-64 converts through a wider signed
		// integer and keeps the low bits, where a direct narrow device
		// conversion would saturate
		return (OUT) (int64_t) v;
	}
	else if constexpr (std::is_floating_point<IN>::value &&
		std::is_same<OUT,uint64_t>::value)
	{
		return v < 0 ? (uint64_t) (int64_t) v : (uint64_t) v;
	}
	else
	{
		return (OUT) v;
	}
}

template <typename OUT, typename IN>
__global__ void __launch_bounds__(256)
convert_kernel (OUT* __restrict__ out, const IN* __restrict__ in, uint64_t n)
{
	constexpr int PER = 4;
	uint64_t step = (uint64_t) gridDim.x * blockDim.x * PER;
	for (uint64_t base = ((uint64_t) blockIdx.x * blockDim.x) * PER +
		threadIdx.x; base < n; base += step)
	{
		IN v[PER];
#pragma unroll
		for (int k = 0; k < PER; ++k)
		{
			uint64_t i = base + (uint64_t) k * blockDim.x;
			if (i < n)
			{
				v[k] = in[i];
			}
		}
#pragma unroll
		for (int k = 0; k < PER; ++k)
		{
			uint64_t i = base + (uint64_t) k * blockDim.x;
			if (i < n)
			{
				out[i] = convert_one<OUT,IN>(v[k]);
			}
		}
	}
}

template <typename IN>
static int convert_from (llo_cuda_ctx* ctx, void* out, int out_dtype,
	const IN* in, uint64_t n)
{
	uint64_t blocks = (n + 1023) / 1024;
	uint64_t cap = (uint64_t) ctx->sm_count * 16;
	unsigned grid = (unsigned) (blocks < cap ? blocks : cap);
	LLO_DISPATCH_DTYPE(out_dtype, OUT,
	{
		convert_kernel<OUT,IN><<<grid, 256, 0, ctx->stream>>>(
			(OUT*) out, in, n);
		return launched(ctx, "convert_kernel");
	})
	return LLO_OK;
}

}

using namespace llo_cuda;

extern "C" int llo_cuda_convert (llo_cuda_ctx* ctx, void* out, int out_dtype,
	const void* in, int in_dtype, uint64_t n)
{
	if (0 == n)
	{
		return LLO_OK;
	}
	if (0 == dtype_size(out_dtype))
	{
		return fail(LLO_ERR_FATAL, "invalid output type %s",
			dtype_name(out_dtype));
	}
	if (out_dtype == in_dtype)
	{
		return llo_cuda_d2d(ctx, out, in, n * dtype_size(out_dtype));
	}
	switch (in_dtype)
	{
		case LLO_DT_DOUBLE: return convert_from<double>(ctx, out, out_dtype, (const double*) in, n);
		case LLO_DT_FLOAT: return convert_from<float>(ctx, out, out_dtype, (const float*) in, n);
		case LLO_DT_INT8: return convert_from<int8_t>(ctx, out, out_dtype, (const int8_t*) in, n);
		case LLO_DT_UINT8: return convert_from<uint8_t>(ctx, out, out_dtype, (const uint8_t*) in, n);
		case LLO_DT_INT16: return convert_from<int16_t>(ctx, out, out_dtype, (const int16_t*) in, n);
		case LLO_DT_UINT16: return convert_from<uint16_t>(ctx, out, out_dtype, (const uint16_t*) in, n);
		case LLO_DT_INT32: return convert_from<int32_t>(ctx, out, out_dtype, (const int32_t*) in, n);
		case LLO_DT_UINT32: return convert_from<uint32_t>(ctx, out, out_dtype, (const uint32_t*) in, n);
		case LLO_DT_INT64: return convert_from<int64_t>(ctx, out, out_dtype, (const int64_t*) in, n);
		case LLO_DT_UINT64: return convert_from<uint64_t>(ctx, out, out_dtype, (const uint64_t*) in, n);
		default:
			return fail(LLO_ERR_FATAL, "invalid input type %s",
				dtype_name(in_dtype));
	}
}
