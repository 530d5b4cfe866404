///
/// general.cu — general coordinate maps (K5).  See general.cuh.
///
/// Not a throughput path: it exists so that ANY ade::CoordMap evaluates to
/// the reference's result (golden vectors with a fractional map:
/// llo/test/test_operator.cpp:29-63, 66-165, 168-263; convolution maps:
/// llo/src/helper.cpp:136-194).  Determinism: overwrite order is resolved
/// with an index max, accumulation order with a stable radix sort by output
/// index followed by an in-order fold per output.
///

#include <cub/device/device_radix_sort.cuh>

#include "general.cuh"
#include "ops.cuh"

namespace llo_cuda
{

struct MapParams
{
	double m[LLO_MAT];
	uint64_t from[LLO_RANK];
	uint64_t to[LLO_RANK];
	uint64_t n;
};

__global__ void __launch_bounds__(256)
index_table_kernel (int64_t* __restrict__ table,
	const __grid_constant__ MapParams p, int* __restrict__ error)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		i < p.n; i += step)
	{
		// ade::coordinate: dimension 0 fastest
		double c[LLO_RANK];
		uint64_t rem = i;
#pragma unroll
		for (int d = 0; d < LLO_RANK; ++d)
		{
			c[d] = (double) (rem % p.from[d]);
			rem /= p.from[d];
		}
		// iCoordMap::forward then ade::index, same operation order as the
		// host (-fmad=false keeps the multiply and the add separate)
		int64_t idx = 0;
		bool bad = false;
#pragma unroll
		for (int j = LLO_RANK - 1; j >= 0; --j)
		{
			double acc = 0;
#pragma unroll
			for (int d = 0; d < LLO_RANK; ++d)
			{
				acc = acc + c[d] * p.m[d * 9 + j];
			}
			acc = acc + p.m[8 * 9 + j];
			double limit = (double) p.to[j];
			if (acc >= limit)
			{
				bad = true;
			}
			if (acc < 0)
			{
				// same floored-modulo wrap as ade::index on the host
				acc = acc - limit * floor(acc / limit);
				if (acc >= limit)
				{
					acc = limit - 1;
				}
			}
			idx = idx * (int64_t) p.to[j] + (bad ? 0 : (int64_t) acc);
		}
		if (bad)
		{
			*error = 1;
			idx = 0;
		}
		table[i] = idx;
	}
}

static unsigned grid_for (llo_cuda_ctx* ctx, uint64_t n, unsigned per_block)
{
	uint64_t blocks = (n + per_block - 1) / per_block;
	uint64_t cap = (uint64_t) ctx->sm_count * 16;
	if (blocks > cap)
	{
		blocks = cap;
	}
	return (unsigned) (blocks > 0 ? blocks : 1);
}

int build_index_table (llo_cuda_ctx* ctx, int64_t* table,
	const double m[LLO_MAT], const uint64_t from[LLO_RANK],
	const uint64_t to[LLO_RANK])
{
	MapParams p;
	for (int i = 0; i < LLO_MAT; ++i)
	{
		p.m[i] = m[i];
	}
	for (int d = 0; d < LLO_RANK; ++d)
	{
		p.from[d] = from[d];
		p.to[d] = to[d];
	}
	p.n = n_elems(from);
	index_table_kernel<<<grid_for(ctx, p.n, 256), 256, 0, ctx->stream>>>(
		table, p, ctx->dev_error);
	return launched(ctx, "index_table_kernel");
}

__global__ void __launch_bounds__(256)
winner_kernel (long long* __restrict__ winner,
	const int64_t* __restrict__ key, uint64_t nin)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		i < nin; i += step)
	{
		atomicMax(winner + key[i], (long long) i);
	}
}

int scatter_winner (llo_cuda_ctx* ctx, int64_t* winner, uint64_t nout,
	const int64_t* key, uint64_t nin)
{
	LLO_CUDA_TRY(cudaMemsetAsync(winner, 0xFF, nout * sizeof(int64_t),
		ctx->stream));
	winner_kernel<<<grid_for(ctx, nin, 256), 256, 0, ctx->stream>>>(
		(long long*) winner, key, nin);
	return launched(ctx, "winner_kernel");
}

template <typename T>
__global__ void __launch_bounds__(256)
gather_winner_kernel (T* __restrict__ out, const T* __restrict__ src,
	const int64_t* __restrict__ winner, uint64_t nout)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		o < nout; o += step)
	{
		int64_t w = winner[o];
		out[o] = w >= 0 ? src[w] : T();
	}
}

int gather_winner (llo_cuda_ctx* ctx, int dtype, void* out,
	const void* src, const int64_t* winner, uint64_t nout)
{
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		gather_winner_kernel<T><<<grid_for(ctx, nout, 256), 256, 0,
			ctx->stream>>>((T*) out, (const T*) src, winner, nout);
		return launched(ctx, "gather_winner_kernel");
	})
	return LLO_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
zero_unreached_kernel (T* __restrict__ out,
	const int64_t* __restrict__ winner, uint64_t nout)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		o < nout; o += step)
	{
		if (winner[o] < 0)
		{
			out[o] = T();
		}
	}
}

int zero_unreached (llo_cuda_ctx* ctx, int dtype, void* out,
	const int64_t* winner, uint64_t nout)
{
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		zero_unreached_kernel<T><<<grid_for(ctx, nout, 256), 256, 0,
			ctx->stream>>>((T*) out, winner, nout);
		return launched(ctx, "zero_unreached_kernel");
	})
	return LLO_OK;
}

__global__ void __launch_bounds__(256)
iota_kernel (uint64_t* __restrict__ pos, uint64_t n)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		i < n; i += step)
	{
		pos[i] = i;
	}
}

/// first position in sorted keys[0, n) that is >= value
__device__ __forceinline__ uint64_t lower_bound_dev (
	const uint64_t* keys, uint64_t n, uint64_t value)
{
	uint64_t lo = 0;
	uint64_t hi = n;
	while (lo < hi)
	{
		uint64_t mid = lo + (hi - lo) / 2;
		if (keys[mid] < value)
		{
			lo = mid + 1;
		}
		else
		{
			hi = mid;
		}
	}
	return lo;
}

template <typename T>
__global__ void __launch_bounds__(256)
ordered_fold_kernel (int opcode, T* __restrict__ out,
	uint8_t* __restrict__ touched, const T* __restrict__ src,
	const uint64_t* __restrict__ sorted_key,
	const uint64_t* __restrict__ sorted_pos, uint64_t nin, uint64_t nout)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		o < nout; o += step)
	{
		uint64_t begin = lower_bound_dev(sorted_key, nin, o);
		uint64_t end = lower_bound_dev(sorted_key, nin, o + 1);
		if (begin == end)
		{
			continue;
		}
		T acc;
		uint64_t k = begin;
		if (touched[o])
		{
			acc = out[o];
		}
		else
		{
			acc = src[sorted_pos[k]];
			++k;
		}
		for (; k < end; ++k)
		{
			acc = apply_binary<T>(opcode, acc, src[sorted_pos[k]]);
		}
		out[o] = acc;
		touched[o] = 1;
	}
}

int ordered_fold (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	uint8_t* touched, const void* src, const int64_t* key,
	uint64_t nin, uint64_t nout)
{
	Scratch scratch(ctx);
	uint64_t* pos = nullptr;
	uint64_t* sorted_key = nullptr;
	uint64_t* sorted_pos = nullptr;
	LLO_TRY(scratch.get((void**) &pos, nin * sizeof(uint64_t)));
	LLO_TRY(scratch.get((void**) &sorted_key, nin * sizeof(uint64_t)));
	LLO_TRY(scratch.get((void**) &sorted_pos, nin * sizeof(uint64_t)));
	iota_kernel<<<grid_for(ctx, nin, 256), 256, 0, ctx->stream>>>(pos, nin);
	LLO_TRY(launched(ctx, "iota_kernel"));

	int end_bit = 1;
	while (end_bit < 64 && (nout >> end_bit) > 0)
	{
		++end_bit;
	}
	size_t temp_bytes = 0;
	const uint64_t* ukey = reinterpret_cast<const uint64_t*>(key);
	LLO_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes,
		ukey, sorted_key, pos, sorted_pos, nin, 0, end_bit, ctx->stream));
	void* temp = nullptr;
	LLO_TRY(scratch.get(&temp, temp_bytes));
	// radix sort is stable: equal keys keep ascending source position
	LLO_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_bytes,
		ukey, sorted_key, pos, sorted_pos, nin, 0, end_bit, ctx->stream));
	ctx->launches += 1;

	LLO_DISPATCH_DTYPE(dtype, T,
	{
		ordered_fold_kernel<T><<<grid_for(ctx, nout, 256), 256, 0,
			ctx->stream>>>(opcode, (T*) out, touched, (const T*) src,
			sorted_key, sorted_pos, nin, nout);
		return launched(ctx, "ordered_fold_kernel");
	})
	return LLO_OK;
}

template <typename T>
__global__ void __launch_bounds__(256)
masked_pull_kernel (int opcode, T* __restrict__ out,
	uint8_t* __restrict__ touched, const T* __restrict__ src,
	const int64_t* __restrict__ table, uint64_t nout)
{
	uint64_t step = (uint64_t) gridDim.x * blockDim.x;
	for (uint64_t o = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
		o < nout; o += step)
	{
		T v = src[table[o]];
		out[o] = touched[o] ? apply_binary<T>(opcode, out[o], v) : v;
		touched[o] = 1;
	}
}

int masked_pull (llo_cuda_ctx* ctx, int opcode, int dtype, void* out,
	uint8_t* touched, const void* src, const int64_t* table, uint64_t nout)
{
	LLO_DISPATCH_DTYPE(dtype, T,
	{
		masked_pull_kernel<T><<<grid_for(ctx, nout, 256), 256, 0,
			ctx->stream>>>(opcode, (T*) out, touched, (const T*) src,
			table, nout);
		return launched(ctx, "masked_pull_kernel");
	})
	return LLO_OK;
}

}
