// Definitions shared by enumerate.cu and fingerprint.cu.
#pragma once
#include "context.h"

namespace sibgpu {

// ---------------------------------------------------------------------------------------------------------------
// layout constants
// ---------------------------------------------------------------------------------------------------------------
constexpr int TILE_THREADS = 256;
constexpr int POS_PER_THREAD = 16;                     // one packed 32-bit word
constexpr int TILE_POS = TILE_THREADS * POS_PER_THREAD; // 4096 text positions per tile
constexpr int MAX_PARTS = 1024;
constexpr uint64_t EMPTY64 = ~0ull;
// The partition cursors are bumped by one returning atomicAdd per (tile, partition): packed 8 bytes apart, 16 of them
// share an L2 line and the line's atomic unit serialises all CTAs (ncu: barrier + long-scoreboard stalls triple as soon
// as a second warp of every CTA has cursors to bump).  One cursor per 128-byte line spreads them over the L2 slices.
constexpr uint32_t CURSOR_STRIDE = 16;

// occurrence context, 8 bits:  [7] forward k-mer is the canonical one  [6] palindrome  [5:3] prev  [2:0] next
// (prev/next are symbols 0..3 = ACGT, 4 = chromosome end '#', already in canonical orientation)
// table payload (stored inverted so that one 0xFF memset initialises keys and payloads):
//   bits 0-4 prev-symbol set, bits 5-9 next-symbol set, bit 10 "seen more than once"
constexpr uint32_t PAY_MULTI = 1u << 10;


struct Rec16 { uint64_t a, b; };

// MODE 0: k <= 28, record = key << 7 | ctx[6:0] in one 64-bit word
// MODE 1: k <= 32, record = {key, ctx}
// MODE 2: k  > 32, record = {fingerprint a, fingerprint b << 8 | ctx}, read from the per-position array d_fp
template<int MODE> struct RecT { typedef Rec16 type; };
template<> struct RecT<0> { typedef uint64_t type; };

__device__ __forceinline__ uint64_t rec_hash(uint64_t a, uint64_t b_fp)
{
	return mix64(a ^ (b_fp * 0x9E3779B97F4A7C15ull));
}

__device__ __forceinline__ uint32_t payload_bits(uint32_t ctx)
{
	uint32_t p = (ctx >> 3) & 7u, n = ctx & 7u;
	uint32_t bits = (1u << p) | (32u << n);
	if(ctx & 64u)                                       // palindrome: the same text position is also an occurrence
	{                                                   // on the other strand, with swapped complemented neighbours
		bits |= (1u << comp_sym(n)) | (32u << comp_sym(p)) | PAY_MULTI;
	}
	return bits;
}

// The reference's predicate in closed form (vertexenumeration.cpp:67-70,330,343,348)
__device__ __forceinline__ bool is_bifurcation(uint32_t pay)
{
	uint32_t P = pay & 31u, Nn = (pay >> 5) & 31u;
	bool sep = ((P | Nn) & 16u) != 0;
	if(pay & PAY_MULTI) return __popc(P) > 1 || __popc(Nn) > 1 || sep;
	return sep;
}

// chromosome cursor of a thread: [cs, ce) is the chromosome containing (or preceding) the current position
struct ChrCursor {
	uint32_t cs, ce, nc;
	__device__ __forceinline__ void init(const TextDesc &t, uint32_t p)
	{
		uint32_t lo = 0, hi = t.nchr;                  // number of chromosomes starting at or before p
		while(lo < hi)
		{
			uint32_t mid = (lo + hi) >> 1;
			if(__ldg(t.chr_start + mid) <= p) lo = mid + 1; else hi = mid;
		}
		nc = lo;
		if(lo == 0) { cs = 0; ce = 0; }
		else { cs = __ldg(t.chr_start + lo - 1); ce = cs + __ldg(t.chr_len + lo - 1); }
	}
	__device__ __forceinline__ void advance(const TextDesc &t, uint32_t p)
	{
		while(nc < t.nchr && p >= __ldg(t.chr_start + nc))
		{
			cs = __ldg(t.chr_start + nc);
			ce = cs + __ldg(t.chr_len + nc);
			nc++;
		}
	}
};

// forward key of the k-mer starting at text position p (k <= 32), straight from the packed words
__device__ __forceinline__ uint64_t key_at(const TextDesc &t, uint32_t p, uint32_t k)
{
	const uint32_t w = p >> 4, sh = 2 * (p & 15u);
	uint64_t x0 = ((uint64_t)__ldg(t.packed + w) << 32) | __ldg(t.packed + w + 1);
	uint64_t x1 = ((uint64_t)__ldg(t.packed + w + 2) << 32);
	uint64_t x = sh ? ((x0 << sh) | (x1 >> (64 - sh))) : x0;
	return x >> (64 - 2 * k);
}


// vertex map: canonical key -> (id of the canonical k-mer, id of its reverse complement, class index); 32-byte slots
struct MapSlot { unsigned long long a, b; uint32_t idc, idr; uint32_t cls; uint32_t pad; };

// 32-base chunk m of the "virtual string" (p, dir): dir = 1 the k-mer starting at text position p, dir = 0 its reverse
// complement.  Left-aligned so that chunks of the same index compare like the strings (the last chunk may be short).
__device__ __forceinline__ uint64_t vstr_chunk(const TextDesc &t, uint32_t p, uint32_t dir, uint32_t m, uint32_t k)
{
	const uint32_t done = 32u * m;
	const uint32_t len = k - done < 32u ? k - done : 32u;
	uint64_t v = dir ? key_at(t, p + done, len) : revcomp_key(key_at(t, p + k - done - len, len), len);
	return v << (64 - 2 * len);
}

} // namespace sibgpu
