// Probe: issue rate of the packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2) against
// the scalar FFMA on a B200, alone and with shared-memory loads interleaved at
// the SGEMM inner-loop ratio (is the packed form a way to free issue slots?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/probes/ffma2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2 (unsigned long long& acc, unsigned long long a, unsigned long long b)
{
	asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// MODE 0: scalar FFMA, 64 accumulators; MODE 1: FFMA2, 32 packed accumulators
// LDS: number of 128-bit shared loads per 64 FMAs (0 = none)
template <int MODE, int LDS>
__global__ void __launch_bounds__(256) loop (float* out, float a0, float b0, int iters)
{
	__shared__ float4 sm[512];
	for (int i = threadIdx.x; i < 512; i += 256)
	{
		sm[i] = make_float4(a0, b0, a0, b0);
	}
	__syncthreads();
	float s = 0;
	if (MODE == 0)
	{
		float acc[64];
#pragma unroll
		for (int i = 0; i < 64; ++i) acc[i] = (float) (threadIdx.x + i);
		float a[8] = {a0, a0, a0, a0, a0, a0, a0, a0};
		float b[8] = {b0, b0, b0, b0, b0, b0, b0, b0};
		for (int it = 0; it < iters; ++it)
		{
#pragma unroll
			for (int l = 0; l < LDS; ++l)
			{
				float4 v = sm[(threadIdx.x + it + l * 32) & 511];
				if (l & 1) { b[(2 * l) & 7] = v.x; b[(2 * l + 1) & 7] = v.y; }
				else { a[(2 * l) & 7] = v.x; a[(2 * l + 1) & 7] = v.y; }
			}
#pragma unroll
			for (int i = 0; i < 8; ++i)
#pragma unroll
				for (int j = 0; j < 8; ++j)
					acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
		}
#pragma unroll
		for (int i = 0; i < 64; ++i) s += acc[i];
	}
	else
	{
		unsigned long long acc[32];
#pragma unroll
		for (int i = 0; i < 32; ++i) acc[i] = (unsigned long long) (threadIdx.x + i) * 0x100000001ull;
		float a[8] = {a0, a0, a0, a0, a0, a0, a0, a0};
		float2 b[4] = {{b0, b0}, {b0, b0}, {b0, b0}, {b0, b0}};
		for (int it = 0; it < iters; ++it)
		{
#pragma unroll
			for (int l = 0; l < LDS; ++l)
			{
				float4 v = sm[(threadIdx.x + it + l * 32) & 511];
				if (l & 1) { b[l & 3] = make_float2(v.x, v.y); }
				else { a[(2 * l) & 7] = v.x; a[(2 * l + 1) & 7] = v.y; }
			}
#pragma unroll
			for (int i = 0; i < 8; ++i)
			{
				float2 ad = make_float2(a[i], a[i]);
				unsigned long long ap = *reinterpret_cast<unsigned long long*>(&ad);
#pragma unroll
				for (int j = 0; j < 4; ++j)
					fma2(acc[i * 4 + j], ap, *reinterpret_cast<unsigned long long*>(&b[j]));
			}
		}
#pragma unroll
		for (int i = 0; i < 32; ++i)
		{
			float2 v = *reinterpret_cast<float2*>(&acc[i]);
			s += v.x + v.y;
		}
	}
	out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int LDS>
double run (const char* name, int sms, int ctas_per_sm)
{
	int blocks = sms * ctas_per_sm;
	float* out;
	cudaMalloc(&out, (size_t) blocks * 256 * sizeof(float));
	int iters = 40000;
	loop<MODE, LDS><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, 100);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0;
	for (int rep = 0; rep < 4; ++rep)
	{
		cudaEventRecord(e0);
		loop<MODE, LDS><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, iters);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double flops = 2.0 * 64 * (double) iters * blocks * 256;
		double tf = flops / ms / 1e9;
		if (tf > best) best = tf;
	}
	printf("%-28s ctas/sm %d: %.2f TFLOP/s\n", name, ctas_per_sm, best);
	cudaFree(out);
	return best;
}

int main ()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
	int sms = prop.multiProcessorCount;
	for (int c = 1; c <= 2; ++c)
	{
		run<0, 0>("ffma  8x8, no lds", sms, c);
		run<1, 0>("ffma2 8x4x2, no lds", sms, c);
		run<0, 4>("ffma  8x8, 4 lds.128", sms, c);
		run<1, 4>("ffma2 8x4x2, 4 lds.128", sms, c);
		run<0, 8>("ffma  8x8, 8 lds.128", sms, c);
		run<1, 8>("ffma2 8x4x2, 8 lds.128", sms, c);
	}
	return 0;
}
