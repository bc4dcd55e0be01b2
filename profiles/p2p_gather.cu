// Peer-read micro-benchmark (one process, peer access enabled between all devices).
//  (1) device 0 gathers random 32-byte sectors (one 8-byte word each) from a 512 MiB buffer on device k.
//  (2) every device runs the same gather at the same time, each thread cycling over ALL OTHER devices
//      (the access pattern of K5 / K6 reading pivots and twins where their owners wrote them),
//      as single words and as warp-wide 256-byte rows.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o p2p_gather p2p_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct Bufs { const uint64_t *p[16]; int n, self; };
__global__ void gather(const uint64_t *__restrict__ src, uint64_t words, uint64_t *out, int per_thread)
{
	uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 12345, acc = 0;
	for (int i = 0; i < per_thread; i++) {
		x ^= x << 13; x ^= x >> 7; x ^= x << 17;
		acc += src[(x % (words / 4)) * 4];
	}
	if (acc == 42) out[0] = acc;
}
// rows != 0: a warp reads 32 consecutive words (256 bytes) at a random row of a random other device
__global__ void gather_all(Bufs B, uint64_t words, uint64_t *out, int per_thread, int rows, int include_self)
{
	const int lane = threadIdx.x & 31;
	uint64_t x = ((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> (rows ? 5 : 0)) * 0x9E3779B97F4A7C15ull + 12345 + B.self, acc = 0;
	for (int i = 0; i < per_thread; i++) {
		x ^= x << 13; x ^= x >> 7; x ^= x << 17;
		int d = (int)((x >> 40) % (include_self ? B.n : B.n - 1));
		if (!include_self && d >= B.self) d++;
		const uint64_t *src = B.p[d];
		acc += rows ? src[(x % (words / 32)) * 32 + lane] : src[(x % (words / 4)) * 4];
	}
	if (acc == 42) out[0] = acc;
}
int main()
{
	int n = 0;
	CK(cudaGetDeviceCount(&n));
	const uint64_t bytes = 512ull << 20, words = bytes / 8;
	uint64_t *buf[16] = {}, *out[16];
	cudaEvent_t a[16], b[16];
	for (int d = 0; d < n; d++) { CK(cudaSetDevice(d)); CK(cudaMalloc(&buf[d], bytes)); CK(cudaMalloc(&out[d], 8)); CK(cudaMemset(buf[d], 1, bytes)); CK(cudaEventCreate(&a[d])); CK(cudaEventCreate(&b[d])); CK(cudaDeviceSynchronize()); }
	for (int s = 0; s < n; s++) {
		CK(cudaSetDevice(s));
		for (int d = 0; d < n; d++) if (d != s) { int can = 0; CK(cudaDeviceCanAccessPeer(&can, s, d)); if (!can) { printf("%d cannot access %d\n", s, d); return 1; } CK(cudaDeviceEnablePeerAccess(d, 0)); }
	}
	const int blocks = 148 * 8, threads = 256, per = 64;
	CK(cudaSetDevice(0));
	for (int d = 0; d < n; d++) {
		gather<<<blocks, threads>>>(buf[d], words, out[0], per); CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(a[0])); gather<<<blocks, threads>>>(buf[d], words, out[0], per); CK(cudaEventRecord(b[0])); CK(cudaEventSynchronize(b[0]));
		float ms; CK(cudaEventElapsedTime(&ms, a[0], b[0]));
		printf("alone: device 0 reads device %d: %.2f G loads/s, %.3f ms\n", d, blocks * (double)threads * per / ms / 1e6, ms);
	}
	for (int active = 1; active <= n; active++)
		for (int rows = 0; rows < 2; rows++) {
			for (int rep = 0; rep < 2; rep++)
				for (int s = 0; s < active; s++) {
					CK(cudaSetDevice(s));
					Bufs B; B.n = n; B.self = s; for (int d = 0; d < n; d++) B.p[d] = buf[d];
					if (rep) CK(cudaEventRecord(a[s]));
					gather_all<<<blocks, threads>>>(B, words, out[s], per, rows, 0);
					if (rep) CK(cudaEventRecord(b[s]));
				}
			for (int s = 0; s < active; s++) {
				CK(cudaSetDevice(s)); CK(cudaDeviceSynchronize());
				float ms; CK(cudaEventElapsedTime(&ms, a[s], b[s]));
				const double reqs = blocks * (double)threads * per / (rows ? 32 : 1);
				printf("%d device(s) active, %s: device %d reads all others: %.2f G %s/s, %.3f ms\n", active, rows ? "256-byte rows" : "single words", s, reqs / ms / 1e6, rows ? "rows" : "loads", ms);
			}
		}
	return 0;
}
