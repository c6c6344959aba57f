// Dev microbenchmark: throughput of shared-memory atomics / match / plain accesses on B200 -- what would bound a
// shared-memory hash grouping (one CTA per bucket) and an atomic-free radix ranking (match.any)?
// Reports lane-operations per clock per SM at the nominal 1965 MHz.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }
__device__ __forceinline__ uint32_t rnd(uint32_t &s) { uint32_t x = lcg(s); x ^= x >> 15; x *= 0x2c1b3c6du; x ^= x >> 12; return x; }

constexpr int ITERS = 4096;
// OP: 0 add32 returning   1 cas64   2 cas32   3 or64 (no return)   4 or32 (no return)   5 lds64+sts64   6 match.any(8 bit)
//     7 add32 no return   8 cas64 then or64 on hit (hash-insert shape)   9 alu only   10 lds64 only  11 exch64
template<int OP>
__global__ void __launch_bounds__(512) k(uint32_t slots_mask, unsigned long long *sink)
{
	extern __shared__ __align__(16) unsigned long long tab[];
	uint32_t *tab32 = reinterpret_cast<uint32_t*>(tab);
	for(uint32_t i = threadIdx.x; i <= slots_mask; i += blockDim.x) tab[i] = ~0ull;
	__syncthreads();
	uint32_t s = blockIdx.x * 977u + threadIdx.x * 31u + 7u;
	unsigned long long acc = 0;
#pragma unroll 4
	for(int it = 0; it < ITERS; it++)
	{
		const uint32_t r = rnd(s);
		const uint32_t slot = r & slots_mask;
		if(OP == 0) acc += atomicAdd(&tab32[slot], 1u);
		if(OP == 1) acc += atomicCAS(&tab[slot], ~0ull, (unsigned long long)r);
		if(OP == 2) acc += atomicCAS(&tab32[slot], ~0u, r);
		if(OP == 3) atomicOr(&tab[slot], (unsigned long long)r);
		if(OP == 4) atomicOr(&tab32[slot], r);
		if(OP == 5) { unsigned long long v = tab[slot]; tab[(slot + 1) & slots_mask] = v + r; }
		if(OP == 6) acc += __match_any_sync(0xffffffffu, r & 255u);
		if(OP == 7) atomicAdd(&tab32[slot], 1u);
		if(OP == 8)
		{
			const unsigned long long w = ((unsigned long long)(r >> 4) << 11) | 1u;
			const unsigned long long old = atomicCAS(&tab[slot], ~0ull, w);
			if(old != ~0ull) atomicOr(&tab[slot], 1024ull | (r & 1023u));
		}
		if(OP == 9) acc += r;
		if(OP == 10) acc += tab[slot];
		if(OP == 11) acc += atomicExch(&tab[slot], (unsigned long long)r);
	}
	if(acc == 0x1234567) *sink = acc;
}

template<int OP> static float run(int grid, int threads, uint32_t slots, unsigned long long *sink)
{
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
	cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(slots * 8));
	float best = 1e9;
	for(int rep = 0; rep < 3; rep++)
	{
		cudaEventRecord(a);
		k<OP><<<grid, threads, slots * 8>>>(slots - 1, sink);
		cudaEventRecord(b); cudaEventSynchronize(b);
		float ms; cudaEventElapsedTime(&ms, a, b); if(ms < best) best = ms;
	}
	return best;
}

int main()
{
	const char *names[] = {"add32_ret", "cas64", "cas32", "or64_nr", "or32_nr", "lds64+sts64", "match_any8", "add32_nr", "cas64+or64",
		"alu_only", "lds64", "exch64"};
	unsigned long long *sink; CK(cudaMalloc(&sink, 8));
	cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
	const int sms = p.multiProcessorCount;
	for(uint32_t slots : {256u, 4096u, 8192u})
	{
		for(int threads : {256, 512})
		{
			for(int cps : {1, 2})
			{
				printf("slots %u (%u KB), %d threads x %d CTAs/SM\n", slots, slots * 8 / 1024, threads, cps);
				for(int op = 0; op < 12; op++)
				{
					const int grid = sms * cps;
					float ms = 0;
					switch(op) {
					case 0: ms = run<0>(grid, threads, slots, sink); break; case 1: ms = run<1>(grid, threads, slots, sink); break;
					case 2: ms = run<2>(grid, threads, slots, sink); break; case 3: ms = run<3>(grid, threads, slots, sink); break;
					case 4: ms = run<4>(grid, threads, slots, sink); break; case 5: ms = run<5>(grid, threads, slots, sink); break;
					case 6: ms = run<6>(grid, threads, slots, sink); break; case 7: ms = run<7>(grid, threads, slots, sink); break;
					case 8: ms = run<8>(grid, threads, slots, sink); break; case 9: ms = run<9>(grid, threads, slots, sink); break;
					case 10: ms = run<10>(grid, threads, slots, sink); break; case 11: ms = run<11>(grid, threads, slots, sink); break; }
					const double ops = (double)grid * threads * ITERS;
					printf("  %-12s %8.1f us  %7.1f Gop/s  %6.2f lane-ops/clk/SM\n", names[op], ms * 1e3, ops / ms / 1e6,
						ops / sms / (ms * 1e-3 * 1.965e9));
				}
			}
		}
	}
	CK(cudaGetLastError());
	return 0;
}
