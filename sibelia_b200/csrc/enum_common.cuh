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
// MODE 2: k  > 32, record = 56-bit fingerprint << 7 | ctx[6:0]; the hash partition comes from a second fingerprint and
//         belongs to the key (a vertex class is {fingerprint, partition}: 16-byte entries in the key list)
template<int MODE> struct RecT { typedef uint64_t type; };
template<> struct RecT<1> { typedef Rec16 type; };
template<int MODE> struct KeyT { typedef typename RecT<MODE>::type type; };     // entry of the vertex-key list
template<> struct KeyT<2> { typedef Rec16 type; };

// ---------------------------------------------------------------------------------------------------------------
// rolling fingerprints (k > 32)
// ---------------------------------------------------------------------------------------------------------------
constexpr uint64_t P61 = (1ull << 61) - 1;

// a * b + c mod 2^61-1 for a, b < 2^61, c < 2^61; result fully reduced
__host__ __device__ __forceinline__ uint64_t muladdmod61(uint64_t a, uint64_t b, uint64_t c)
{
#ifdef __CUDA_ARCH__
	const uint64_t hi = __umul64hi(a, b), lo = a * b;
#else
	const unsigned __int128 z = (unsigned __int128)a * b;
	const uint64_t hi = (uint64_t)(z >> 64), lo = (uint64_t)z;
#endif
	uint64_t r = (lo & P61) + (lo >> 61) + (hi << 3) + c;      // 2^61 = 1 (mod p); hi < 2^58: the sum stays below 2^63
	r = (r & P61) + (r >> 61);
	return r >= P61 ? r - P61 : r;
}
__host__ __device__ __forceinline__ uint64_t mulmod61(uint64_t a, uint64_t b) { return muladdmod61(a, b, 0); }
__host__ __device__ __forceinline__ uint64_t addmod61(uint64_t a, uint64_t b)
{
	uint64_t r = a + b;
	return r >= P61 ? r - P61 : r;
}

// D[h][out * 4 + in]: what one rolling step adds to hash h (0: h1 forward, 1: h1 reverse, 2: h2 forward, 3: h2 reverse)
// after the multiplication by the base (forward) / its inverse (reverse)
struct FpBases { uint64_t B1, invB1, B2, invB2; };
struct FpParams { uint64_t D[4][16]; FpBases b; };
struct FpCk { uint64_t hf1, hr1, hf2, hr2; };          // the four hashes of the k-mer starting at a position = 0 (mod 16)
typedef FpCk FpState;
struct FpView { const FpCk *ck; const FpParams *prm; };

// sD = FpParams in shared memory (the lanes of a warp index the tables with different symbols: shared memory serves
// distinct 8-byte entries from distinct banks, a constant bank would replay per distinct address)
constexpr int FP_SMEM_WORDS = sizeof(FpParams) / 8;
__device__ __forceinline__ void fp_stage_params(const FpParams *__restrict__ prm, uint64_t *sD)
{
	for(uint32_t i = threadIdx.x; i < FP_SMEM_WORDS; i += blockDim.x) sD[i] = __ldg(reinterpret_cast<const uint64_t*>(prm) + i);
}
__device__ __forceinline__ FpBases fp_bases(const uint64_t *sD) { return FpBases{sD[64], sD[65], sD[66], sD[67]}; }
__device__ __forceinline__ void fp_roll(FpState &h, const uint64_t *sD, const FpBases &b, uint32_t idx)
{
	h.hf1 = muladdmod61(h.hf1, b.B1, sD[idx]);
	h.hr1 = muladdmod61(h.hr1, b.invB1, sD[16 + idx]);
	h.hf2 = h.hf2 * b.B2 + sD[32 + idx];
	h.hr2 = h.hr2 * b.invB2 + sD[48 + idx];
}
// canonical orientation = the smaller (h1, h2) pair; equal pairs = the k-mer is its own reverse complement
__device__ __forceinline__ bool fp_forward(const FpState &h) { return h.hf1 < h.hr1 || (h.hf1 == h.hr1 && h.hf2 <= h.hr2); }
__device__ __forceinline__ bool fp_palindrome(const FpState &h) { return h.hf1 == h.hr1 && h.hf2 == h.hr2; }

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
