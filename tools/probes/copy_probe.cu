// Scratch probe: what does a streaming copy / write need on B200 to reach the
// cudaMemcpy rate?  Variants differ in cache hints, vectors in flight per
// thread and grid shape.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int HINT> __device__ __forceinline__ uint4 ld(const uint4* p) {
	if (HINT == 0) return *p;
	if (HINT == 1) return __ldg(p);
	if (HINT == 2) return __ldcs(p);
	uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x),"=r"(r.y),"=r"(r.z),"=r"(r.w) : "l"(p)); return r;
}
template <int HINT> __device__ __forceinline__ void st(uint4* p, uint4 v) {
	if (HINT == 0) { *p = v; return; }
	if (HINT == 1) { __stcs(p, v); return; }
	if (HINT == 2) { __stwt(p, v); return; }
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w));
}

// persistent grid-stride, U vectors per thread per iteration
template <int LH, int SH, int U>
__global__ void __launch_bounds__(256) copy_gs(uint4* __restrict__ out, const uint4* __restrict__ in, uint64_t nvec) {
	uint64_t step = (uint64_t) gridDim.x * blockDim.x * U;
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x * U + threadIdx.x; i < nvec; i += step) {
		uint4 v[U];
#pragma unroll
		for (int u = 0; u < U; ++u) if (i + u * 256 < nvec) v[u] = ld<LH>(in + i + u * 256);
#pragma unroll
		for (int u = 0; u < U; ++u) if (i + u * 256 < nvec) st<SH>(out + i + u * 256, v[u]);
	}
}
// thread-contiguous: each thread owns 2 adjacent vectors (like the VM's VEC=8 floats)
template <int LH, int SH>
__global__ void __launch_bounds__(256) copy_pair(uint4* __restrict__ out, const uint4* __restrict__ in, uint64_t nvec) {
	uint64_t step = (uint64_t) gridDim.x * blockDim.x * 2;
	for (uint64_t i = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 2; i < nvec; i += step) {
		uint4 a = ld<LH>(in + i), b = ld<LH>(in + i + 1);
		st<SH>(out + i, a); st<SH>(out + i + 1, b);
	}
}
template <int SH, int U>
__global__ void __launch_bounds__(256) fill_gs(uint4* __restrict__ out, uint64_t nvec) {
	uint64_t step = (uint64_t) gridDim.x * blockDim.x * U;
	uint4 v = make_uint4(1, 2, 3, 4);
	for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x * U + threadIdx.x; i < nvec; i += step) {
#pragma unroll
		for (int u = 0; u < U; ++u) if (i + u * 256 < nvec) st<SH>(out + i + u * 256, v);
	}
}

template <typename F> float timeit(F f, int iters = 10) {
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	for (int i = 0; i < 3; ++i) f();
	float best = 1e9;
	for (int i = 0; i < iters; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
	return best;
}

int main() {
	uint64_t n = 1ull << 30; uint64_t nvec = n / 16;
	uint4 *in, *out; cudaMalloc(&in, n); cudaMalloc(&out, n); cudaMemset(in, 1, n); cudaMemset(out, 0, n);
	int sms = 148;
	auto rep = [&](const char* name, float ms, double bytes) { printf("%-44s %7.3f ms %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
	rep("cudaMemcpy d2d", timeit([&]{ cudaMemcpyAsync(out, in, n, cudaMemcpyDeviceToDevice); }), 2.0 * n);
	rep("cudaMemset", timeit([&]{ cudaMemsetAsync(out, 0, n); }), 1.0 * n);
#define RUN(NAME, KERNEL, GRID, BYTES) rep(NAME, timeit([&]{ KERNEL<<<GRID, 256>>>(out, in, nvec); }), BYTES)
	int grids[] = {sms * 4, sms * 8, sms * 16, sms * 32, (int)((nvec + 255) / 256)};
	for (int g : grids) {
		char name[128];
		snprintf(name, 128, "copy plain/plain U1 grid %d", g); RUN(name, (copy_gs<0,0,1>), g, 2.0 * n);
		snprintf(name, 128, "copy plain/plain U2 grid %d", g / 2 > 0 ? g / 2 : 1); RUN(name, (copy_gs<0,0,2>), g / 2, 2.0 * n);
		snprintf(name, 128, "copy plain/plain U4 grid %d", g / 4); RUN(name, (copy_gs<0,0,4>), g / 4, 2.0 * n);
		snprintf(name, 128, "copy ldcs/stcs U2 grid %d", g / 2); RUN(name, (copy_gs<2,1,2>), g / 2, 2.0 * n);
		snprintf(name, 128, "copy ldcs/plain U2 grid %d", g / 2); RUN(name, (copy_gs<2,0,2>), g / 2, 2.0 * n);
		snprintf(name, 128, "copy nc.noalloc/noalloc U2 grid %d", g / 2); RUN(name, (copy_gs<3,3,2>), g / 2, 2.0 * n);
		snprintf(name, 128, "copy ldcs/stcs U4 grid %d", g / 4); RUN(name, (copy_gs<2,1,4>), g / 4, 2.0 * n);
		snprintf(name, 128, "copy pair ldcs/stcs grid %d", g / 2); RUN(name, (copy_pair<2,1>), g / 2, 2.0 * n);
		snprintf(name, 128, "copy pair plain grid %d", g / 2); RUN(name, (copy_pair<0,0>), g / 2, 2.0 * n);
	}
	for (int g : grids) {
		char name[128];
		snprintf(name, 128, "fill plain U1 grid %d", g); rep(name, timeit([&]{ fill_gs<0,1><<<g, 256>>>(out, nvec); }), 1.0 * n);
		snprintf(name, 128, "fill stcs U1 grid %d", g); rep(name, timeit([&]{ fill_gs<1,1><<<g, 256>>>(out, nvec); }), 1.0 * n);
		snprintf(name, 128, "fill plain U4 grid %d", g / 4); rep(name, timeit([&]{ fill_gs<0,4><<<g / 4, 256>>>(out, nvec); }), 1.0 * n);
	}
	return 0;
}
