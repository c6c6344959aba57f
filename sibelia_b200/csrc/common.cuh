// Shared host/device helpers of libsibgpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/sibgpu.h"

namespace sibgpu {

void set_error(const std::string &msg);

#define SIB_CUDA(expr)                                                                                          \
	do {                                                                                                        \
		cudaError_t e__ = (expr);                                                                               \
		if(e__ != cudaSuccess) {                                                                                \
			::sibgpu::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +   \
				std::to_string(__LINE__) + ")");                                                                \
			return SIBGPU_ERR_CUDA;                                                                             \
		}                                                                                                       \
	} while(0)

#define SIB_TRY(expr)                                                                                           \
	do {                                                                                                        \
		int s__ = (expr);                                                                                       \
		if(s__ != SIBGPU_OK) return s__;                                                                        \
	} while(0)

// NVTX range for the profilers' timelines (header-only NVTX 3: a no-op unless a tool is attached).
struct NvtxRange {
	explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange &operator=(const NvtxRange&) = delete;
};

// Grow-only device buffer.
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if(bytes <= cap) return SIBGPU_OK;
		if(p) cudaFree(p);
		p = nullptr;
		cap = 0;
		size_t want = bytes + bytes / 8 + 256;
		SIB_CUDA(cudaMalloc(&p, want));
		cap = want;
		return SIBGPU_OK;
	}
	void release()
	{
		if(p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
	template<class T> T *as() const { return static_cast<T*>(p); }
};

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------

// 2-bit code of an upper-case base: A=0 C=1 G=2 T=3 (lexicographic, so packed keys compare like strings).
// Works bytewise on 4 ASCII characters at once: code = ((c >> 1) ^ (c >> 2)) & 3.
__device__ __forceinline__ uint32_t encode4(uint32_t w)
{
	uint32_t c = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
	// first character (lowest byte) goes to the top bit pair
	return ((c << 6) | (c >> 4) | (c >> 14) | (c >> 24)) & 0xFFu;
}

// exact per-byte "== 0" mask (0x80 in every zero byte)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t t)
{
	return ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);
}

// 0x80 in every byte of w that is one of 'A','C','G','T','$'
__device__ __forceinline__ uint32_t legal_bytes(uint32_t w)
{
	return zero_bytes(w ^ 0x41414141u) | zero_bytes(w ^ 0x43434343u) | zero_bytes(w ^ 0x47474747u) |
		zero_bytes(w ^ 0x54545454u) | zero_bytes(w ^ 0x24242424u);
}

// reverse complement of a k-mer packed in the low 2k bits (first base most significant)
__device__ __forceinline__ uint64_t revcomp_key(uint64_t f, uint32_t k)
{
	uint64_t x = ~f;                                   // complement: 3 - c on every pair
	x = __brevll(x);                                   // reverses bits, also inside the pairs
	x = ((x & 0xAAAAAAAAAAAAAAAAull) >> 1) | ((x & 0x5555555555555555ull) << 1);
	return x >> (64 - 2 * k);
}

// murmur3 finaliser: all 64 output bits usable
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
	x ^= x >> 33;
	x *= 0xff51afd7ed558ccdull;
	x ^= x >> 33;
	x *= 0xc4ceb9fe1a85ec53ull;
	x ^= x >> 33;
	return x;
}

__device__ __forceinline__ uint32_t comp_sym(uint32_t s) { return s == 4u ? 4u : 3u - s; }

} // namespace sibgpu
