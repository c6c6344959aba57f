// Dev microbenchmark: how fast can SMs of GPU 0 move 8-byte records from/to GPU 1 over NVLink?
//   rd8 xN   : per-thread 8-byte loads from the peer, N independent loads in flight
//   rd16 xN  : 16-byte loads
//   bulk     : cp.async.bulk (TMA 1-D) 8 KB chunks peer -> shared memory, 3-stage mbarrier ring
//   wr8/wr16 : coalesced stores to the peer
//   local_*  : the same kernels on local memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/p2p tools/ubench/p2p.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

template<typename T, int ILP>
__global__ void __launch_bounds__(256) k_read(const T* __restrict__ src, uint64_t n, unsigned long long* sink)
{
	unsigned long long acc = 0;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride * ILP)
	{
		T v[ILP];
#pragma unroll
		for(int u = 0; u < ILP; u++) if(i + u * stride < n) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
		for(int u = 0; u < ILP; u++) if(i + u * stride < n) acc += *reinterpret_cast<unsigned long long*>(&v[u]);
	}
	if(acc == 0x1234567ull) *sink = acc;
}

template<typename T>
__global__ void __launch_bounds__(256) k_write(T* __restrict__ dst, uint64_t n, T val)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = val;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CHUNK = 8192, STAGES = 3;
__global__ void __launch_bounds__(256) k_bulk(const unsigned char* __restrict__ src, uint64_t nchunks, unsigned long long* sink)
{
	__shared__ __align__(128) unsigned char buf[STAGES][CHUNK];
	__shared__ __align__(8) unsigned long long bar[STAGES];
	if(threadIdx.x == 0)
	{
		for(int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[s])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	unsigned long long acc = 0;
	auto issue = [&](uint64_t c, int s) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[s])), "r"(CHUNK) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(buf[s])), "l"(src + c * CHUNK), "r"(CHUNK), "r"(smem_u32(&bar[s])) : "memory");
	};
	uint64_t c_issue = blockIdx.x;
	if(threadIdx.x == 0)
	{
		for(int s = 0; s < STAGES && c_issue < nchunks; s++, c_issue += gridDim.x) issue(c_issue, s);
	}
	int stage = 0;
	uint32_t phase = 0;
	for(uint64_t c = blockIdx.x; c < nchunks; c += gridDim.x)
	{
		uint32_t ok = 0;
		while(!ok)
		{
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
				: "=r"(ok) : "r"(smem_u32(&bar[stage])), "r"(phase) : "memory");
		}
		const unsigned long long* w = reinterpret_cast<const unsigned long long*>(buf[stage]);
#pragma unroll
		for(int u = 0; u < CHUNK / 8 / 256; u++) acc += w[u * 256 + threadIdx.x];
		__syncthreads();
		if(threadIdx.x == 0)
		{
			const uint64_t cn = c + (uint64_t)STAGES * gridDim.x;
			if(cn < nchunks) issue(cn, stage);
		}
		stage++;
		if(stage == STAGES) { stage = 0; phase ^= 1; }
	}
	if(acc == 0x1234567ull) *sink = acc;
}

int main()
{
	int nd = 0;
	CK(cudaGetDeviceCount(&nd));
	const uint64_t bytes = 512ull << 20;
	unsigned char *loc, *rem = nullptr;
	CK(cudaSetDevice(0));
	CK(cudaMalloc(&loc, bytes));
	CK(cudaMemset(loc, 1, bytes));
	if(nd > 1)
	{
		int can = 0;
		CK(cudaDeviceCanAccessPeer(&can, 0, 1));
		printf("devices %d, peer access 0->1: %d\n", nd, can);
		CK(cudaSetDevice(1));
		CK(cudaMalloc(&rem, bytes));
		CK(cudaMemset(rem, 1, bytes));
		CK(cudaDeviceSynchronize());
		CK(cudaSetDevice(0));
		CK(cudaDeviceEnablePeerAccess(1, 0));
	}
	unsigned long long* sink;
	CK(cudaMalloc(&sink, 8));
	cudaEvent_t a, b;
	cudaEventCreate(&a);
	cudaEventCreate(&b);
	const int grids[] = {148 * 2, 148 * 8};
	for(int where = 0; where < (rem ? 2 : 1); where++)
	{
		unsigned char* p = where ? rem : loc;
		const char* wn = where ? "peer " : "local";
		for(int g : grids)
		{
			for(int op = 0; op < 9; op++)
			{
				float best = 1e9;
				for(int rep = 0; rep < 3; rep++)
				{
					cudaEventRecord(a);
					switch(op)
					{
					case 0: k_read<unsigned long long, 1><<<g, 256>>>((const unsigned long long*)p, bytes / 8, sink); break;
					case 1: k_read<unsigned long long, 4><<<g, 256>>>((const unsigned long long*)p, bytes / 8, sink); break;
					case 2: k_read<unsigned long long, 16><<<g, 256>>>((const unsigned long long*)p, bytes / 8, sink); break;
					case 3: k_read<ulonglong2, 1><<<g, 256>>>((const ulonglong2*)p, bytes / 16, sink); break;
					case 4: k_read<ulonglong2, 4><<<g, 256>>>((const ulonglong2*)p, bytes / 16, sink); break;
					case 5: k_read<ulonglong2, 8><<<g, 256>>>((const ulonglong2*)p, bytes / 16, sink); break;
					case 6: k_bulk<<<g, 256>>>(p, bytes / CHUNK, sink); break;
					case 7: k_write<unsigned long long><<<g, 256>>>((unsigned long long*)p, bytes / 8, 1ull); break;
					case 8: k_write<ulonglong2><<<g, 256>>>((ulonglong2*)p, bytes / 16, make_ulonglong2(1, 1)); break;
					}
					cudaEventRecord(b);
					CK(cudaEventSynchronize(b));
					float ms;
					cudaEventElapsedTime(&ms, a, b);
					if(ms < best) best = ms;
				}
				const char* names[] = {"rd8 x1", "rd8 x4", "rd8 x16", "rd16 x1", "rd16 x4", "rd16 x8", "bulk 8K x3", "wr8", "wr16"};
				printf("%s grid %5d %-10s %8.3f ms  %8.1f GB/s\n", wn, g, names[op], best, bytes / 1e6 / best);
			}
		}
	}
	if(rem)
	{
		float best = 1e9;
		for(int rep = 0; rep < 3; rep++)
		{
			cudaEventRecord(a);
			cudaMemcpyAsync(loc, rem, bytes, cudaMemcpyDefault);
			cudaEventRecord(b);
			CK(cudaEventSynchronize(b));
			float ms;
			cudaEventElapsedTime(&ms, a, b);
			if(ms < best) best = ms;
		}
		printf("copy engine peer->local %8.3f ms  %8.1f GB/s\n", best, bytes / 1e6 / best);
	}
	return 0;
}
