// Measures the random-gather ceiling of HBM on this GPU: every thread reads `unroll` independent,
// random, aligned blocks of 32 / 64 / 128 bytes from a buffer far larger than L2 (non-allocating
// loads, like the index probes). Prints useful GB/s (bytes requested) per block size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

template <int BYTES, int UNROLL>
__global__ void k_gather(const u64 *__restrict__ buf, u64 nblocks, u64 per_thread, u64 *out)
{
	u64 tid = (u64)blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
	for (u64 it = 0; it < per_thread; it += UNROLL) {
		u64 v[UNROLL][BYTES / 8];
		#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const u64 *p = buf + (mix(tid * 0x9E3779B97F4A7C15ULL + it + u) & (nblocks - 1)) * (BYTES / 8);   // nblocks is a power of two
			#pragma unroll
			for (int q = 0; q < BYTES / 32; q++)
				asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[u][4 * q]), "=l"(v[u][4 * q + 1]), "=l"(v[u][4 * q + 2]), "=l"(v[u][4 * q + 3]) : "l"(p + 4 * q));
		}
		#pragma unroll
		for (int u = 0; u < UNROLL; u++)
			#pragma unroll
			for (int q = 0; q < BYTES / 8; q++) acc ^= v[u][q];
	}
	if (acc == 0x1234567) out[0] = acc;
}

template <int BYTES, int UNROLL> void run(const u64 *buf, u64 bytes, u64 *out, int sms)
{
	const u64 nblocks = bytes / BYTES, per_thread = 512;
	const int grid = sms * 8, block = 256;
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	float best = 1e30f;
	for (int rep = 0; rep < 5; rep++) {
		cudaEventRecord(a);
		k_gather<BYTES, UNROLL><<<grid, block>>>(buf, nblocks, per_thread, out);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b); if (rep && ms < best) best = ms;
	}
	double req = (double)grid * block * per_thread * BYTES;
	printf("random %3d-byte gathers, %d in flight per thread: %7.1f GB/s useful (%.0f M gathers/s)\n", BYTES, UNROLL, req / best / 1e6, req / BYTES / best / 1e3);
}

int main()
{
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	const u64 bytes = 4ull << 30;
	u64 *buf, *out; cudaMalloc(&buf, bytes); cudaMalloc(&out, 8); cudaMemset(buf, 1, bytes);
	printf("%s, %d SMs, buffer %llu MiB\n", p.name, p.multiProcessorCount, bytes >> 20);
	// buffer-size sweep: L2-resident (64 MiB), beyond L2 but inside the TLB reach, far beyond both
	for (u64 mb : {64ull, 256ull, 512ull, 1024ull, 2048ull, 4096ull}) {
		printf("buffer %llu MiB:\n", mb);
		run<32, 4>(buf, mb << 20, out, p.multiProcessorCount);
		run<64, 4>(buf, mb << 20, out, p.multiProcessorCount);
		run<128, 4>(buf, mb << 20, out, p.multiProcessorCount);
	}
	return 0;
}
