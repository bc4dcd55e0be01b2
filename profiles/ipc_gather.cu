// Multi-PROCESS peer-read micro-benchmark: one process per GPU (forked before CUDA is initialised), every
// process exports a 512 MiB buffer and maps everybody else's, then all run the K5/K6 access pattern at
// once (random 8-byte words / 256-byte rows from all other GPUs). Two ways of sharing the buffers:
//   ipc : cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle (legacy CUDA IPC)
//   vmm : cuMemCreate + POSIX file descriptor (SCM_RIGHTS over abstract unix sockets) + cuMemMap (2 MiB pages)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ipc_gather ipc_gather.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <sys/wait.h>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("[%d] CUDA error %s at line %d\n", rank, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define CU(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char *s; cuGetErrorString(e, &s); printf("[%d] driver error %s at line %d\n", rank, s, __LINE__); exit(1); } } while (0)
struct Bufs { const uint64_t *p[16]; int n, self; };
__global__ void gather_all(Bufs B, uint64_t words, uint64_t *out, int per_thread, int rows)
{
	const int lane = threadIdx.x & 31;
	uint64_t x = ((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> (rows ? 5 : 0)) * 0x9E3779B97F4A7C15ull + 12345 + B.self, acc = 0;
	for (int i = 0; i < per_thread; i++) {
		x ^= x << 13; x ^= x >> 7; x ^= x << 17;
		int d = (int)((x >> 40) % (B.n - 1));
		if (d >= B.self) d++;
		const uint64_t *src = B.p[d];
		acc += rows ? src[(x % (words / 32)) * 32 + lane] : src[(x % (words / 4)) * 4];
	}
	if (acc == 42) out[0] = acc;
}
struct Shared { volatile int arrive[8]; cudaIpcMemHandle_t h[16]; };
static Shared *sh;
static int rank, n, mbytes = 512;
static void barrier(int k) { __sync_fetch_and_add(&sh->arrive[k], 1); while (sh->arrive[k] < n) usleep(100); }
static void sock_name(sockaddr_un *a, socklen_t *len, int r) { memset(a, 0, sizeof *a); a->sun_family = AF_UNIX; int l = snprintf(a->sun_path + 1, sizeof a->sun_path - 2, "ogb_ipc_gather_%d_%d", (int)getppid(), r); *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + l); }
static void send_fd(int to, int fd)
{
	int s = socket(AF_UNIX, SOCK_STREAM, 0); sockaddr_un a; socklen_t al; sock_name(&a, &al, to);
	if (connect(s, (sockaddr *)&a, al) != 0) { perror("connect"); exit(1); }
	char cb[CMSG_SPACE(sizeof(int))]; memset(cb, 0, sizeof cb); int payload = rank; iovec io = {&payload, sizeof payload};
	msghdr m = {}; m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = cb; m.msg_controllen = sizeof cb;
	cmsghdr *c = CMSG_FIRSTHDR(&m); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int)); memcpy(CMSG_DATA(c), &fd, sizeof(int));
	if (sendmsg(s, &m, 0) < 0) { perror("sendmsg"); exit(1); }
	close(s);
}
static int recv_fd(int ls, int *from)
{
	int s = accept(ls, nullptr, nullptr); if (s < 0) { perror("accept"); exit(1); }
	char cb[CMSG_SPACE(sizeof(int))]; int payload = -1; iovec io = {&payload, sizeof payload};
	msghdr m = {}; m.msg_iov = &io; m.msg_iovlen = 1; m.msg_control = cb; m.msg_controllen = sizeof cb;
	if (recvmsg(s, &m, 0) <= 0) { perror("recvmsg"); exit(1); }
	int fd = -1; cmsghdr *c = CMSG_FIRSTHDR(&m); memcpy(&fd, CMSG_DATA(c), sizeof(int)); close(s); *from = payload; return fd;
}
static int child(int mode)
{
	const uint64_t bytes = (uint64_t)mbytes << 20, words = bytes / 8;
	CK(cudaSetDevice(rank)); CK(cudaFree(0));
	Bufs B; B.n = n; B.self = rank;
	uint64_t *mine = nullptr, *out;
	CK(cudaMalloc(&out, 8));
	if (mode == 0) {
		CK(cudaMalloc(&mine, bytes)); CK(cudaMemset(mine, 1, bytes)); CK(cudaDeviceSynchronize());
		CK(cudaIpcGetMemHandle(&sh->h[rank], mine));
		barrier(0);
		for (int d = 0; d < n; d++) { if (d == rank) { B.p[d] = mine; continue; } void *p; CK(cudaIpcOpenMemHandle(&p, sh->h[d], cudaIpcMemLazyEnablePeerAccess)); B.p[d] = (const uint64_t *)p; }
	} else {
		int ls = socket(AF_UNIX, SOCK_STREAM, 0); sockaddr_un a; socklen_t al; sock_name(&a, &al, rank);
		if (bind(ls, (sockaddr *)&a, al) != 0 || listen(ls, 32) != 0) { perror("bind/listen"); exit(1); }
		CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = rank;
		prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
		size_t gran = 0; CU(cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
		if (rank == 0) printf("vmm granularity %zu\n", gran);
		CUmemGenericAllocationHandle hmine; CU(cuMemCreate(&hmine, bytes, &prop, 0));
		int fd = -1; CU(cuMemExportToShareableHandle(&fd, hmine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
		CUmemAccessDesc acc = {}; acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = rank; acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
		CUdeviceptr va; CU(cuMemAddressReserve(&va, bytes, gran, 0, 0)); CU(cuMemMap(va, bytes, 0, hmine, 0)); CU(cuMemSetAccess(va, bytes, &acc, 1));
		mine = (uint64_t *)va; CK(cudaMemset(mine, 1, bytes)); CK(cudaDeviceSynchronize());
		B.p[rank] = mine;
		barrier(0);
		for (int d = 0; d < n; d++) if (d != rank) send_fd(d, fd);
		for (int k = 0; k < n - 1; k++) {
			int from, pfd = recv_fd(ls, &from);
			CUmemGenericAllocationHandle hp; CU(cuMemImportFromShareableHandle(&hp, (void *)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
			CUdeviceptr pva; CU(cuMemAddressReserve(&pva, bytes, gran, 0, 0)); CU(cuMemMap(pva, bytes, 0, hp, 0)); CU(cuMemSetAccess(pva, bytes, &acc, 1));
			B.p[from] = (const uint64_t *)pva; close(pfd);
		}
	}
	barrier(1);
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	const int blocks = 148 * 8, threads = 256, per = 64;
	for (int rows = 0; rows < 2; rows++) {
		gather_all<<<blocks, threads>>>(B, words, out, per, rows); CK(cudaDeviceSynchronize());
		barrier(2 + 2 * rows);
		CK(cudaEventRecord(a)); gather_all<<<blocks, threads>>>(B, words, out, per, rows); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		const double reqs = blocks * (double)threads * per / (rows ? 32 : 1);
		printf("%s, %d MiB, %d processes, %s: rank %d reads all others: %.2f G %s/s, %.3f ms\n", mode ? "vmm" : "ipc", mbytes, n, rows ? "256-byte rows" : "single words", rank, reqs / ms / 1e6, rows ? "rows" : "loads", ms);
		barrier(3 + 2 * rows);
	}
	return 0;
}
int main(int argc, char **argv)
{
	const int mode = argc > 1 && !strcmp(argv[1], "vmm");
	n = argc > 2 ? atoi(argv[2]) : 2;
	if (argc > 3) mbytes = atoi(argv[3]);
	sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	memset(sh, 0, sizeof *sh);
	for (rank = 0; rank < n; rank++) { pid_t p = fork(); if (p == 0) { if (mode) { CUresult e = cuInit(0); (void)e; } return child(mode); } }
	int rc = 0; for (int i = 0; i < n; i++) { int st; wait(&st); rc |= st; }
	return rc != 0;
}
