// Probe: sustained FFMA / DFMA issue rate of the whole chip (the denominator of
// the GEMM roofline).  Each thread runs 16 independent FMA chains so the pipe
// latency is covered; 8 warps x 4 CTAs per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fma_peak tools/probes/fma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <typename T>
__global__ void __launch_bounds__(256) fma_loop (T* out, T a, T b, int iters)
{
	T acc[16];
#pragma unroll
	for (int i = 0; i < 16; ++i)
	{
		acc[i] = (T) (threadIdx.x + i);
	}
	for (int it = 0; it < iters; ++it)
	{
#pragma unroll
		for (int r = 0; r < 8; ++r)
		{
#pragma unroll
			for (int i = 0; i < 16; ++i)
			{
				acc[i] = acc[i] * a + b; // contracted to one FMA
			}
		}
	}
	T s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i)
	{
		s += acc[i];
	}
	out[(size_t) blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename T>
double run (const char* name, int sms, int seconds_hint)
{
	int blocks = sms * 8;
	T* out;
	cudaMalloc(&out, (size_t) blocks * 256 * sizeof(T));
	int iters = 20000;
	fma_loop<T><<<blocks, 256>>>(out, (T) 1.0000001, (T) 1e-9, 100);
	cudaDeviceSynchronize();
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	double best = 0;
	for (int rep = 0; rep < 5 * seconds_hint; ++rep)
	{
		cudaEventRecord(e0);
		fma_loop<T><<<blocks, 256>>>(out, (T) 1.0000001, (T) 1e-9, iters);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double flops = 2.0 * 16 * 8 * (double) iters * blocks * 256;
		double tf = flops / ms / 1e9;
		printf("%s rep %d: %.3f ms  %.2f TFLOP/s\n", name, rep, ms, tf);
		if (tf > best) best = tf;
	}
	cudaFree(out);
	return best;
}

int main ()
{
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, 0);
	printf("%s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
	double f = run<float>("ffma", prop.multiProcessorCount, 2);
	double d = run<double>("dfma", prop.multiProcessorCount, 1);
	printf("{\"ffma_tflops\": %.2f, \"dfma_tflops\": %.2f, \"sms\": %d}\n", f, d, prop.multiProcessorCount);
	return 0;
}
