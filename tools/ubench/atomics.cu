// Dev microbenchmark: throughput of scattered L2/DRAM operations on B200 (what bounds k_insert?).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__device__ __forceinline__ uint64_t mix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdull;x^=x>>33;x*=0xc4ceb9fe1a85ec53ull;x^=x>>33;return x;}

template<int OP>
__global__ void __launch_bounds__(256) k(const uint64_t* __restrict__ recs, uint64_t n, unsigned long long* tab, uint32_t T, unsigned long long* sink)
{
	unsigned long long acc = 0;
	for(uint64_t i = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x; i < n; i += (uint64_t)gridDim.x*blockDim.x)
	{
		uint64_t r = recs[i];
		uint32_t slot = __umulhi((uint32_t)mix64(r), T);
		if(OP==0) acc += __ldcg(&tab[slot]);
		if(OP==1) acc += atomicCAS(&tab[slot], ~0ull, r);
		if(OP==2) acc += atomicCAS(reinterpret_cast<unsigned int*>(tab)+slot, ~0u, (unsigned)r);
		if(OP==3) atomicOr(&tab[slot], r);
		if(OP==4) atomicOr(reinterpret_cast<unsigned int*>(tab)+slot, (unsigned)r);
		if(OP==5) tab[slot] = r;
		if(OP==6) { unsigned long long c = __ldcg(&tab[slot]); if(c==~0ull) c = atomicCAS(&tab[slot], ~0ull, r); acc += c; }
		if(OP==7) acc += atomicExch(&tab[slot], r);
		if(OP==8) acc += atomicMax(&tab[slot], r);
		if(OP==9) { acc += __ldcg(&tab[slot]); tab[slot] = r; }
		if(OP==10) acc += mix64(r ^ slot);
	}
	if(acc == 0x1234567) *sink = acc;
}

__global__ void fill(uint64_t* recs, uint64_t n){ for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=(uint64_t)gridDim.x*blockDim.x) recs[i]=mix64(i*0x9E3779B97F4A7C15ull+12345); }

int main()
{
	const char* names[] = {"ld64","cas64","cas32","red_or64","red_or32","st64","ld+cas64","exch64","max64","ld+st64","alu_only"};
	uint64_t n = 4u<<20;
	uint64_t* recs; CK(cudaMalloc(&recs, n*8)); fill<<<1184,256>>>(recs,n);
	unsigned long long* sink; CK(cudaMalloc(&sink, 8));
	cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
	for(uint32_t T : {1u<<21, 1u<<23, 1u<<25, 1u<<28})
	{
		unsigned long long* tab; CK(cudaMalloc(&tab, (size_t)T*8));
		printf("table %u slots (%.0f MB), %llu ops per launch\n", T, T*8.0/1e6, (unsigned long long)n);
		for(int op=0; op<11; op++)
		{
			float best=1e9;
			for(int rep=0; rep<4; rep++)
			{
				CK(cudaMemset(tab, 0xFF, (size_t)T*8));
				CK(cudaDeviceSynchronize());
				cudaEventRecord(a);
				switch(op){
				case 0: k<0><<<1184,256>>>(recs,n,tab,T,sink); break; case 1: k<1><<<1184,256>>>(recs,n,tab,T,sink); break;
				case 2: k<2><<<1184,256>>>(recs,n,tab,T,sink); break; case 3: k<3><<<1184,256>>>(recs,n,tab,T,sink); break;
				case 4: k<4><<<1184,256>>>(recs,n,tab,T,sink); break; case 5: k<5><<<1184,256>>>(recs,n,tab,T,sink); break;
				case 6: k<6><<<1184,256>>>(recs,n,tab,T,sink); break; case 7: k<7><<<1184,256>>>(recs,n,tab,T,sink); break;
				case 8: k<8><<<1184,256>>>(recs,n,tab,T,sink); break; case 9: k<9><<<1184,256>>>(recs,n,tab,T,sink); break;
				case 10: k<10><<<1184,256>>>(recs,n,tab,T,sink); break; }
				cudaEventRecord(b); CK(cudaEventSynchronize(b));
				float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best) best=ms;
			}
			printf("  %-10s %8.1f us  %6.1f Gop/s\n", names[op], best*1e3, n/best/1e6);
		}
		cudaFree(tab);
	}
	return 0;
}
