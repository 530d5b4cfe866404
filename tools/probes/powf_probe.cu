// powf_probe.cu — accuracy and speed of the double-float pow used by the
// elementwise engine (ops.cuh: pow_fast) against powf and double pow.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I tools/probes \
//      tools/probes/powf_probe.cu -o /tmp/powf_probe && /tmp/powf_probe
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#include "powf_alt.cuh"

// instruction cost: 16 evaluations per element loaded
__global__ void cost_fast (float* out, const float* a, const float* b, size_t n)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	float x = a[i], y = b[i], acc = 0.0f;
#pragma unroll 4
	for (int k = 0; k < 16; ++k) { acc += llo_cuda::pow_fast(x + 0.03125f * k, y); }
	out[i] = acc;
}

__global__ void cost_lib (float* out, const float* a, const float* b, size_t n)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	float x = a[i], y = b[i], acc = 0.0f;
#pragma unroll 4
	for (int k = 0; k < 16; ++k) { acc += powf(x + 0.03125f * k, y); }
	out[i] = acc;
}

__global__ void run_fast (float* out, const float* a, const float* b, size_t n)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { out[i] = llo_cuda::pow_fast(a[i], b[i]); }
}

__global__ void run_lib (float* out, const float* a, const float* b, size_t n)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) { out[i] = powf(a[i], b[i]); }
}

// worst relative error of got against double pow; mismatching specials counted
__global__ void errors (const float* got, const float* a, const float* b, size_t n,
	double* worst, unsigned long long* bad)
{
	size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) { return; }
	double want = pow((double) a[i], (double) b[i]);
	float wf = (float) want;
	float g = got[i];
	if (isnan(wf) || isinf(wf) || wf == 0.0f || fabs(want) < 1.2e-38)
	{
		bool same = (isnan(wf) && isnan(g)) || (wf == g) ||
			(fabs(want) < 1.2e-38 && fabsf(g - wf) <= 2e-45f * 4);
		if (!same) { atomicAdd(bad, 1ull); }
		return;
	}
	double rel = fabs((double) g - want) / fabs(want);
	unsigned long long* w = (unsigned long long*) worst;
	unsigned long long old = *w, assumed;
	do
	{
		assumed = old;
		if (__longlong_as_double(assumed) >= rel) { break; }
		old = atomicCAS(w, assumed, __double_as_longlong(rel));
	}
	while (assumed != old);
}

static uint64_t rng_state = 88172645463325252ull;
static double urand ()
{
	rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
	return (double) (rng_state >> 11) / 9007199254740992.0;
}

int main ()
{
	const size_t n = 1u << 26;
	std::vector<float> a(n), b(n);
	for (size_t i = 0; i < n; ++i)
	{
		int kind = (int) (i / (n / 8));   // homogeneous warps; kind 0 alone is timed
		double x, y;
		switch (kind)
		{
			case 0: x = 0.5 + 1.5 * urand(); y = 0.5 + 1.5 * urand(); break;        // the bench's range
			case 1: x = exp((urand() - 0.5) * 160); y = (urand() - 0.5) * 4; break;   // wide bases
			case 2: x = 0.9 + 0.2 * urand(); y = (urand() - 0.5) * 2000; break;       // near 1, big exponents
			case 3: x = exp((urand() - 0.5) * 20); y = (urand() - 0.5) * 40; break;   // over / underflow
			case 4: x = floor(urand() * 20); y = floor(urand() * 12) - 4; break;      // small integers, zero
			case 5: x = -floor(urand() * 20) * 0.5; y = floor(urand() * 9) - 4 + (i % 3 ? 0 : 0.5); break; // negative bases
			case 6: x = urand() * 1e-38; y = urand() * 2; break;                      // subnormal bases
			default: x = urand() * 100; y = (i % 5 == 0) ? INFINITY : ((i % 7 == 0) ? NAN : 2.0); break;
		}
		a[i] = (float) x; b[i] = (float) y;
	}
	float *da, *db, *dout; double* dworst; unsigned long long* dbad;
	cudaMalloc(&da, n * 4); cudaMalloc(&db, n * 4); cudaMalloc(&dout, n * 4);
	cudaMalloc(&dworst, 8); cudaMalloc(&dbad, 8);
	cudaMemcpy(da, a.data(), n * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(db, b.data(), n * 4, cudaMemcpyHostToDevice);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	for (int which = 0; which < 2; ++which)
	{
		float best = 1e9f;
		for (int it = 0; it < 5; ++it)
		{
			// timed: the first eighth (bases and exponents in [0.5, 2]), 8 passes
			cudaEventRecord(e0);
			for (int rep = 0; rep < 8; ++rep)
			{
				if (which) { cost_fast<<<(unsigned) (n / 8 / 256), 256>>>(dout, da, db, n / 8); }
				else { cost_lib<<<(unsigned) (n / 8 / 256), 256>>>(dout, da, db, n / 8); }
			}
			cudaEventRecord(e1); cudaEventSynchronize(e1);
			float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
		}
		if (which) { run_fast<<<(unsigned) (n / 256), 256>>>(dout, da, db, n); }
		else { run_lib<<<(unsigned) (n / 256), 256>>>(dout, da, db, n); }
		cudaMemset(dworst, 0, 8); cudaMemset(dbad, 0, 8);
		errors<<<(unsigned) (n / 256), 256>>>(dout, da, db, n, dworst, dbad);
		double worst; unsigned long long bad;
		cudaMemcpy(&worst, dworst, 8, cudaMemcpyDeviceToHost);
		cudaMemcpy(&bad, dbad, 8, cudaMemcpyDeviceToHost);
		printf("%s: %.3f ms for 2^30 evaluations (%.1f G/s), worst rel err %.3e, special-value mismatches %llu, %s\n",
			which ? "pow_fast" : "powf    ", best, 16.0 * n / best / 1e6, worst, bad,
			cudaGetErrorString(cudaGetLastError()));
	}
	return 0;
}
