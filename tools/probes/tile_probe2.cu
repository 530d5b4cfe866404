// Probe 2: what a persistent tile-copy loop needs to reach the flat-copy rate:
// CTAs per SM, tiles in flight per CTA, static vs counter-based tile hand-out,
// persistent vs one-tile-per-CTA grids.  16 KB tiles of 64 rows x 256 B.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tile_probe2.bin tools/probes/tile_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ uint32_t counter;

template <int INFLIGHT, bool DYNAMIC, int CTAS, int ORDER = 0>
__global__ void __launch_bounds__(256, CTAS) tile_copy (const uint4* __restrict__ src,
	uint4* __restrict__ dst, uint32_t side_chunks, uint32_t rows, uint32_t persistent)
{
	constexpr int W = 16, H = 64;
	uint32_t tiles_x = side_chunks / W;
	uint32_t total = (rows / H) * tiles_x / INFLIGHT;     // work units of INFLIGHT tiles
	__shared__ uint32_t slot;
	uint32_t t = blockIdx.x;
	while (t < total)
	{
		uint4 v[INFLIGHT][4];
#pragma unroll
		for (int f = 0; f < INFLIGHT; ++f)
		{
			uint32_t tile = t * INFLIGHT + f;
			uint32_t ty = tile / tiles_x, tx = tile % tiles_x;
			if (1 == ORDER) { tx = tile / tiles_x; ty = tile % tiles_x; }   // column-major: no concurrent row neighbours
			if (2 == ORDER) { uint32_t blk = tile / 16, in = tile % 16; uint32_t bx = blk % (tiles_x / 4), by = blk / (tiles_x / 4); tx = bx * 4 + in % 4; ty = by * 4 + in / 4; }
#pragma unroll
			for (int q = 0; q < 4; ++q)
			{
				uint32_t c = threadIdx.x + 256 * q;
				v[f][q] = __ldcs(src + (size_t) (ty * H + c / W) * side_chunks + tx * W + c % W);
			}
		}
#pragma unroll
		for (int f = 0; f < INFLIGHT; ++f)
		{
			uint32_t tile = t * INFLIGHT + f;
			uint32_t ty = tile / tiles_x, tx = tile % tiles_x;
			if (1 == ORDER) { tx = tile / tiles_x; ty = tile % tiles_x; }
			if (2 == ORDER) { uint32_t blk = tile / 16, in = tile % 16; uint32_t bx = blk % (tiles_x / 4), by = blk / (tiles_x / 4); tx = bx * 4 + in % 4; ty = by * 4 + in / 4; }
#pragma unroll
			for (int q = 0; q < 4; ++q)
			{
				uint32_t c = threadIdx.x + 256 * q;
				__stcs(dst + (size_t) (ty * H + c / W) * side_chunks + tx * W + c % W, v[f][q]);
			}
		}
		if (0 == persistent)
		{
			break;
		}
		if (DYNAMIC)
		{
			__syncthreads();
			if (0 == threadIdx.x) { slot = gridDim.x + atomicAdd(&counter, 1u); }
			__syncthreads();
			t = slot;
		}
		else
		{
			t += gridDim.x;
		}
	}
}

template <int INFLIGHT, bool DYNAMIC, int CTAS, int ORDER = 0>
void run (const char* what, const uint4* src, uint4* dst, uint32_t side, bool persistent)
{
	uint32_t side_chunks = side / 4;
	uint32_t units = (side / 64) * (side_chunks / 16) / INFLIGHT;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e9f;
	for (int rep = 0; rep < 6; ++rep)
	{
		uint32_t zero = 0;
		cudaMemcpyToSymbol(counter, &zero, 4);
		cudaEventRecord(e0);
		tile_copy<INFLIGHT,DYNAMIC,CTAS,ORDER><<<persistent ? 148 * CTAS : units, 256>>>(src, dst, side_chunks, side, persistent ? 1 : 0);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms < best) best = ms;
	}
	double bytes = 2.0 * side * (double) side * 4;
	printf("%-46s %.3f ms  %.0f GB/s\n", what, best, bytes / best / 1e6);
}

int main ()
{
	uint32_t side = 16384;
	uint4 *src, *dst;
	cudaMalloc(&src, (size_t) side * side * 4);
	cudaMalloc(&dst, (size_t) side * side * 4);
	cudaMemset(src, 1, (size_t) side * side * 4);
	run<1,false,3>("persistent 3 CTA/SM, 1 tile, static", src, dst, side, true);
	run<1,true,3>("persistent 3 CTA/SM, 1 tile, counter", src, dst, side, true);
	run<1,false,4>("persistent 4 CTA/SM, 1 tile, static", src, dst, side, true);
	run<1,true,4>("persistent 4 CTA/SM, 1 tile, counter", src, dst, side, true);
	run<1,true,6>("persistent 6 CTA/SM, 1 tile, counter", src, dst, side, true);
	run<1,true,8>("persistent 8 CTA/SM, 1 tile, counter", src, dst, side, true);
	run<2,true,3>("persistent 3 CTA/SM, 2 tiles, counter", src, dst, side, true);
	run<2,true,4>("persistent 4 CTA/SM, 2 tiles, counter", src, dst, side, true);
	run<4,true,2>("persistent 2 CTA/SM, 4 tiles, counter", src, dst, side, true);
	run<1,true,3,1>("persistent 3 CTA/SM, 1 tile, counter, col-major", src, dst, side, true);
	run<2,true,3,1>("persistent 3 CTA/SM, 2 tiles, counter, col-major", src, dst, side, true);
	run<1,true,3,2>("persistent 3 CTA/SM, 1 tile, counter, 4x4 blocks", src, dst, side, true);
	run<2,true,3,2>("persistent 3 CTA/SM, 2 tiles, counter, 4x4 blocks", src, dst, side, true);
	run<1,false,8>("grid of tiles (one per CTA)", src, dst, side, false);
	run<2,false,4>("grid of tile pairs (one pair per CTA)", src, dst, side, false);
	return 0;
}
