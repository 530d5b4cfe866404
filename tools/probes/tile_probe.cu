// Probe: HBM throughput of tiled copies as a function of the contiguous run
// per row (what a transposing kernel sees on one or both sides).  Every CTA
// copies tiles of H rows x W bytes (row pitch = matrix width) from src to dst.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tile_probe.bin tools/probes/tile_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// tile: WR 16-byte chunks per row on the read side x H rows; written back as
// tiles of WW chunks per row (same bytes, different shape) -- no transpose,
// only the access shapes differ
template <int WR, int WW>
__global__ void __launch_bounds__(256, 3) tile_copy (const uint4* __restrict__ src,
	uint4* __restrict__ dst, uint32_t side_chunks, uint32_t rows)
{
	constexpr int CHUNKS = 1024;              // 16 KB per tile
	constexpr int HR = CHUNKS / WR, HW = CHUNKS / WW;
	uint32_t tiles_x_r = side_chunks / WR, tiles_x_w = side_chunks / WW;
	uint32_t total = (rows / HR) * tiles_x_r;
	for (uint32_t t = blockIdx.x; t < total; t += gridDim.x)
	{
		uint32_t ty = t / tiles_x_r, tx = t % tiles_x_r;
		uint4 v[4];
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			uint32_t c = threadIdx.x + 256 * q;
			uint32_t r = c / WR, x = c % WR;
			v[q] = __ldcs(src + (size_t) (ty * HR + r) * side_chunks + tx * WR + x);
		}
		// same tile number on the write side, its own shape
		uint32_t wy = t / tiles_x_w, wx = t % tiles_x_w;
#pragma unroll
		for (int q = 0; q < 4; ++q)
		{
			uint32_t c = threadIdx.x + 256 * q;
			uint32_t r = c / WW, x = c % WW;
			__stcs(dst + (size_t) (wy * HW + r) * side_chunks + wx * WW + x, v[q]);
		}
	}
}

template <int WR, int WW>
void run (const uint4* src, uint4* dst, uint32_t side)
{
	uint32_t side_chunks = side / 4;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	float best = 1e9f;
	for (int rep = 0; rep < 6; ++rep)
	{
		cudaEventRecord(e0);
		tile_copy<WR,WW><<<148 * 3, 256>>>(src, dst, side_chunks, side);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms;
		cudaEventElapsedTime(&ms, e0, e1);
		if (rep > 0 && ms < best) best = ms;
	}
	double bytes = 2.0 * side * (double) side * 4;
	printf("read run %5d B, write run %5d B: %.3f ms  %.0f GB/s\n", WR * 16, WW * 16, best, bytes / best / 1e6);
}

int main ()
{
	uint32_t side = 16384;
	uint4 *src, *dst;
	cudaMalloc(&src, (size_t) side * side * 4);
	cudaMalloc(&dst, (size_t) side * side * 4);
	cudaMemset(src, 1, (size_t) side * side * 4);
	run<16,16>(src, dst, side);
	run<32,32>(src, dst, side);
	run<64,64>(src, dst, side);
	run<128,128>(src, dst, side);
	run<1024,1024>(src, dst, side);
	run<16,64>(src, dst, side);
	run<64,16>(src, dst, side);
	run<16,1024>(src, dst, side);
	run<1024,16>(src, dst, side);
	run<8,8>(src, dst, side);
	run<32,16>(src, dst, side);
	run<16,32>(src, dst, side);
	return 0;
}
