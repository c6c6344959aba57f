// Bifurcation (vertex) enumeration on the GPU -- replaces IndexedSequence::EnumerateBifurcationsSArrayInRAM
// (/root/reference/src/vertexenumeration.cpp:263-364).  The reference sorts every suffix of both strands
// (libdivsufsort + Kasai LCP) to find the classes of equal k-mers; here the classes are found by hashing:
//
//   K0 k_pack         ASCII text -> 2-bit packed words (+ legality check)                 1 B/base read, .25 written
//   K1 k_scatter      rolling canonical k-mer keys of every position; records (mixed key, context) counting-sorted per
//                     CTA in shared memory and written to fixed-capacity hash partitions in coalesced runs
//   K2 k_split        (group_smem.cuh) every partition split once more into buckets of ~1 Ki records; tiles arrive by TMA
//   K3 k_group        (group_smem.cuh) one bucket at a time per CTA in shared memory: open-addressing table, shared
//                     atomics, the reference's predicate (vertexenumeration.cpp:67-70,330,348), warp-ballot key append
//      fallbacks      k_scan_hist + k_part_offsets (exactly sized partitions when a fixed-capacity region overflows);
//                     k_insert + k_table_scan (one L2-resident table per partition when a bucket overflows)
//   K5 k_expand / cub sort / k_build_map    vertex id = lexicographic rank among {w, revcomp(w)} (:350)
//   K6 k_mark, K7 k_emit   second scan of the packed text: positions whose canonical key is a vertex, compacted in
//                     text order -> the two (chr,pos)-sorted instance tables (:361-362)
//   k > 32            fingerprint.cu (k_fp_ckpt) + scan16_fp below: rolling fingerprints fused into K1 / K6
//
// Both strands are handled with ONE record per text position: the record carries the canonical key
// min(w, revcomp(w)) and the neighbour symbols re-expressed in the canonical orientation, so the class of w and the
// class of revcomp(w) (which the reference enumerates separately and symmetrically) are decided once.
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <type_traits>
#include <cub/cub.cuh>

#include "enum_common.cuh"
#include "group_smem.cuh"

#ifndef SIBGPU_SMEM_PART_KI
#define SIBGPU_SMEM_PART_KI 256                       // Ki records per level-1 partition of the shared-memory grouping
#endif

namespace sibgpu {

// ---------------------------------------------------------------------------------------------------------------
// K0: pack
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const uint4 *__restrict__ text, uint32_t *__restrict__ packed,
	uint32_t nwords, uint32_t *__restrict__ err)
{
	// PACK_ILP independent 128-bit loads in flight per thread at 16 CTAs per SM; streaming loads, the ASCII text is read
	// exactly once.  A 100 MB text is ~20 us of HBM time inside a ~37 us kernel: launch ramp and tail dominate, and
	// 2 / 4 / 8 loads in flight at 4 / 8 / 16 CTAs per SM all land within 37-42 us (profiles/r2_k_pack_variants.txt)
	constexpr int PACK_ILP = 2;
	uint32_t bad = 0;
	const uint32_t stride = gridDim.x * blockDim.x;
	for(uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nwords; i0 += stride * PACK_ILP)
	{
		uint4 v[PACK_ILP];
#pragma unroll
		for(int u = 0; u < PACK_ILP; u++)
		{
			const uint32_t i = i0 + u * stride;
			v[u] = i < nwords ? __ldcs(text + i) : make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);
		}
#pragma unroll
		for(int u = 0; u < PACK_ILP; u++)
		{
			const uint32_t i = i0 + u * stride;
			bad |= ~(legal_bytes(v[u].x) & legal_bytes(v[u].y) & legal_bytes(v[u].z) & legal_bytes(v[u].w)) & 0x80808080u;
			if(i < nwords) packed[i] = (encode4(v[u].x) << 24) | (encode4(v[u].y) << 16) | (encode4(v[u].z) << 8) | encode4(v[u].w);
		}
	}
	if(__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31) == 0) atomicOr(err, 1u);
}

// ---------------------------------------------------------------------------------------------------------------
// per-thread scan of 16 consecutive text positions
// ---------------------------------------------------------------------------------------------------------------
// Stages the packed words of a tile (one back-halo word, 256 own, 4 forward-halo) into shared memory.
__device__ __forceinline__ void stage_tile(const TextDesc &t, uint32_t tile, uint32_t *sw)
{
	int64_t base = (int64_t)tile * TILE_THREADS - 1;
	for(int j = threadIdx.x; j < TILE_THREADS + 5; j += TILE_THREADS)
	{
		int64_t w = base + j;
		sw[j] = (w >= 0 && w < (int64_t)t.nwords) ? __ldg(t.packed + w) : 0u;
	}
}

// Calls f(i, a, b, ctx) for every valid k-mer start among the thread's 16 positions (exact modes, k <= 32).
//   a   = canonical key min(w, revcomp(w));  ctx as documented above (bit 7 = forward is canonical)
template<class F>
__device__ __forceinline__ void scan16_exact(const TextDesc &t, const uint32_t *sw, uint32_t tile, uint32_t k, F f_emit)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t p0 = tile * TILE_POS + tid * POS_PER_THREAD;
	if(p0 >= t.M) return;
	const uint32_t wm1 = sw[tid], w0 = sw[tid + 1], w1 = sw[tid + 2], w2 = sw[tid + 3], w3 = sw[tid + 4];
	const uint64_t hi = ((uint64_t)w0 << 32) | w1, lo = ((uint64_t)w2 << 32) | w3;
	const uint32_t kk = 2 * k;
	const uint64_t mask = kk == 64 ? ~0ull : ((1ull << kk) - 1);
	uint64_t f = hi >> (64 - kk);
	uint64_t stream = kk == 64 ? lo : ((hi << kk) | (lo >> (64 - kk)));   // bases p0+k, p0+k+1, ...
	uint64_t r = revcomp_key(f, k);
	uint32_t prevc = wm1 & 3u;
	ChrCursor cur;
	cur.init(t, p0);
	// Interior fast path: all 16 k-mers, their predecessors and their successors lie inside one chromosome (true for
	// all but a handful of threads of the whole text): no cursor, no validity tests, no '#' contexts.
	if(p0 > cur.cs && (uint64_t)p0 + (POS_PER_THREAD - 1) + k < cur.ce)
	{
		const uint32_t top = kk - 2;
#pragma unroll
		for(int i = 0; i < POS_PER_THREAD; i++)
		{
			const uint32_t nextc = (uint32_t)(stream >> 62);
			const bool fw = f <= r;
			const uint64_t canon = fw ? f : r;
			// forward: prev << 3 | next;  reverse: comp(next) << 3 | comp(prev) = 27 - 8 next - prev
			uint32_t ctx = fw ? ((prevc << 3) | nextc) : (27u - (nextc << 3) - prevc);
			ctx |= (f == r ? 64u : 0u) | (fw ? 128u : 0u);
			f_emit(i, canon, (uint64_t)0, ctx);
			prevc = (uint32_t)(f >> top) & 3u;
			f = ((f << 2) | nextc) & mask;
			r = (r >> 2) | ((uint64_t)(3u - nextc) << top);
			stream <<= 2;
		}
		return;
	}
#pragma unroll
	for(int i = 0; i < POS_PER_THREAD; i++)
	{
		const uint32_t p = p0 + i;
		const uint32_t nextc = (uint32_t)(stream >> 62);
		cur.advance(t, p);
		if(p >= cur.cs && p + k <= cur.ce)
		{
			const uint32_t ps = p == cur.cs ? 4u : prevc;
			const uint32_t ns = p + k == cur.ce ? 4u : nextc;
			const bool fw = f <= r;
			const uint64_t canon = fw ? f : r;
			uint32_t ctx = fw ? ((ps << 3) | ns) : ((comp_sym(ns) << 3) | comp_sym(ps));
			ctx |= (f == r ? 64u : 0u) | (fw ? 128u : 0u);
			f_emit(i, canon, (uint64_t)0, ctx);
		}
		prevc = (uint32_t)(f >> (kk - 2)) & 3u;
		f = ((f << 2) | nextc) & mask;
		r = (r >> 2) | ((uint64_t)(3u - nextc) << (kk - 2));
		stream <<= 2;
	}
}

// MODE 2 (k > 32): the four rolling hashes start from the checkpoint of the thread's word (k_fp_ckpt, fingerprint.cu)
// and advance one position at a time; f(i, a, hb, ctx): a = 56-bit fingerprint of the canonical orientation, hb = the
// 32 bits of the second hash that choose the hash partition (the partition index is part of the class key)
// UNROLL < 16 for callers whose per-position work is large and divergent (k_mark: map probes): the fully unrolled body
// would be 16 copies of the rolling step + the caller's code
template<int UNROLL, class F>
__device__ __forceinline__ void scan16_fp(const TextDesc &t, const FpView &fv, const uint32_t *sw, const uint64_t *sD,
	uint32_t tile, uint32_t k, F f_emit)
{
	const uint32_t tid = threadIdx.x;
	const uint32_t p0 = tile * TILE_POS + tid * POS_PER_THREAD;
	if(p0 >= t.M) return;
	const uint32_t wm1 = sw[tid], w0 = sw[tid + 1];
	// the symbols entering the window: text positions p0 + k ... p0 + k + 15
	const uint32_t wi = (p0 + k) >> 4;
	const uint32_t in_lo = wi < t.nwords ? __ldg(t.packed + wi) : 0u, in_hi = wi + 1 < t.nwords ? __ldg(t.packed + wi + 1) : 0u;
	const uint32_t win = __funnelshift_l(in_hi, in_lo, 2u * (k & 15u));
	const FpBases bs = fp_bases(sD);
	FpState h;
	{
		const ulonglong2 *c = reinterpret_cast<const ulonglong2*>(fv.ck + (p0 >> 4));
		const ulonglong2 c0 = __ldg(c), c1 = __ldg(c + 1);
		h.hf1 = c0.x; h.hr1 = c0.y; h.hf2 = c1.x; h.hr2 = c1.y;
	}
	uint32_t prevc = wm1 & 3u;
	ChrCursor cur;
	cur.init(t, p0);
	const bool interior = p0 > cur.cs && (uint64_t)p0 + (POS_PER_THREAD - 1) + k < cur.ce;
#pragma unroll UNROLL
	for(int i = 0; i < POS_PER_THREAD; i++)
	{
		const uint32_t p = p0 + i;
		const uint32_t outc = (w0 >> (30 - 2 * i)) & 3u, nextc = (win >> (30 - 2 * i)) & 3u;
		bool valid = true;
		uint32_t ps = prevc, ns = nextc;
		if(!interior)
		{
			cur.advance(t, p);
			valid = p >= cur.cs && p + k <= cur.ce;
			if(p == cur.cs) ps = 4u;
			if(p + k == cur.ce) ns = 4u;
		}
		if(valid)
		{
			const bool fw = fp_forward(h);
			uint32_t ctx = fw ? ((ps << 3) | ns) : ((comp_sym(ns) << 3) | comp_sym(ps));
			ctx |= (fp_palindrome(h) ? 64u : 0u) | (fw ? 128u : 0u);
			f_emit(i, (fw ? h.hf1 : h.hr1) & MIX_MASK, (fw ? h.hf2 : h.hr2) >> 32, ctx);
		}
		prevc = outc;
		fp_roll(h, sD, bs, (outc << 2) | nextc);
	}
}

// the same for ONE text position (k_emit: only the hit positions are revisited)
__device__ __forceinline__ void fp_at(const TextDesc &t, const FpView &fv, const uint64_t *sD, uint32_t p, uint32_t k,
	uint64_t &a, uint64_t &hb, bool &fw)
{
	const uint32_t w = p >> 4, wi = w + (k >> 4);
	const uint32_t w0 = __ldg(t.packed + w);
	const uint32_t in_lo = wi < t.nwords ? __ldg(t.packed + wi) : 0u, in_hi = wi + 1 < t.nwords ? __ldg(t.packed + wi + 1) : 0u;
	const uint32_t win = __funnelshift_l(in_hi, in_lo, 2u * (k & 15u));
	const FpBases bs = fp_bases(sD);
	const ulonglong2 *c = reinterpret_cast<const ulonglong2*>(fv.ck + w);
	const ulonglong2 c0 = __ldg(c), c1 = __ldg(c + 1);
	FpState h = {c0.x, c0.y, c1.x, c1.y};
	for(uint32_t i = 0; i < (p & 15u); i++) fp_roll(h, sD, bs, (((w0 >> (30 - 2 * i)) & 3u) << 2) | ((win >> (30 - 2 * i)) & 3u));
	fw = fp_forward(h);
	a = (fw ? h.hf1 : h.hr1) & MIX_MASK;
	hb = (fw ? h.hf2 : h.hr2) >> 32;
}

template<int MODE, int UNROLL = POS_PER_THREAD, class F>
__device__ __forceinline__ void scan16(const TextDesc &t, const FpView &fv, const uint32_t *sw, const uint64_t *sD, uint32_t tile,
	uint32_t k, F f_emit)
{
	if(MODE == 2) scan16_fp<UNROLL>(t, fv, sw, sD, tile, k, f_emit);
	else scan16_exact(t, sw, tile, k, f_emit);
}

// hash partition of a record: plain records hash the key, mixed records (group_smem.cuh) read a bit field of it;
// fingerprint records (MODE 2) take it from the second hash (b = its top 32 bits)
template<int MODE, bool MIXED>
__device__ __forceinline__ uint32_t part_of(uint64_t a, uint64_t b, uint32_t P)
{
	if(MODE == 2) return __umulhi((uint32_t)b, P);
	if(MIXED) return MODE == 0 ? mixed_part(a, P) : __umulhi((uint32_t)(a >> 32), P);
	return __umulhi((uint32_t)(rec_hash(a, b) >> 32), P);
}

// ---------------------------------------------------------------------------------------------------------------
// K1: scan + partition histogram
// ---------------------------------------------------------------------------------------------------------------
template<int MODE>
__global__ void __launch_bounds__(TILE_THREADS) k_scan_hist(TextDesc t, const FpView fv, uint32_t k,
	uint32_t ntiles, uint32_t P, uint32_t *__restrict__ ghist)
{
	__shared__ uint32_t sw[TILE_THREADS + 5];
	__shared__ uint32_t shist[MAX_PARTS];
	__shared__ uint64_t sD[MODE == 2 ? FP_SMEM_WORDS : 1];
	if(MODE == 2) fp_stage_params(fv.prm, sD);
	for(uint32_t b = threadIdx.x; b < P; b += TILE_THREADS) shist[b] = 0;
	for(uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
	{
		__syncthreads();
		stage_tile(t, t.tile0 + tile, sw);
		__syncthreads();
		scan16<MODE>(t, fv, sw, sD, t.tile0 + tile, k, [&](int, uint64_t a, uint64_t b, uint32_t) {
			atomicAdd(&shist[part_of<MODE, false>(a, b, P)], 1u);
		});
	}
	__syncthreads();
	for(uint32_t b = threadIdx.x; b < P; b += TILE_THREADS)
	{
		uint32_t c = shist[b];
		if(c) atomicAdd(&ghist[b], c);
	}
}

// exclusive scan of the P partition counts (one CTA); also publishes the largest partition
__global__ void __launch_bounds__(MAX_PARTS) k_part_offsets(const uint32_t *__restrict__ hist, uint32_t P,
	uint64_t *__restrict__ partoff, unsigned long long *__restrict__ cursor, uint64_t *__restrict__ scalars)
{
	typedef cub::BlockScan<uint64_t, MAX_PARTS> Scan;
	typedef cub::BlockReduce<uint32_t, MAX_PARTS> Red;
	__shared__ union { typename Scan::TempStorage s; typename Red::TempStorage r; } tmp;
	uint32_t c = threadIdx.x < P ? hist[threadIdx.x] : 0u;
	uint64_t off, total;
	Scan(tmp.s).ExclusiveSum((uint64_t)c, off, total);
	__syncthreads();
	uint32_t mx = Red(tmp.r).Reduce(c, cub::Max());
	if(threadIdx.x < P) { partoff[threadIdx.x] = off; cursor[threadIdx.x * CURSOR_STRIDE] = off; }
	if(threadIdx.x == 0) { partoff[P] = total; scalars[0] = total; scalars[1] = mx; }
}

// ---------------------------------------------------------------------------------------------------------------
// K2: scan + scatter into hash partitions
// ---------------------------------------------------------------------------------------------------------------
template<int MODE> struct ScatterSmem {
	uint32_t sw[TILE_THREADS + 5];
	uint32_t cnt[MAX_PARTS];           // per-bin count of this tile, then exclusive local offset
	uint32_t gbase[MAX_PARTS];         // index of the bin's run in `out` minus its local offset (mod 2^32: + local index >= offset)
	uint32_t dropmask[MAX_PARTS / 32]; // bins whose run did not fit its fixed-capacity region
	uint32_t total, anydrop;
	uint64_t a[TILE_POS];
	uint16_t bin_of[TILE_POS];         // bin of the record at each sorted local index (plain records only)
};
template<> struct ScatterSmem<1> : ScatterSmem<0> { uint64_t b[TILE_POS]; };
template<> struct ScatterSmem<2> : ScatterSmem<0> { uint64_t sD[FP_SMEM_WORDS]; };
// dynamic shared memory of k_scatter<MODE, MIXED>: mixed 8-byte records read their partition back from the record, their
// kernel never touches bin_of (the last member of the MODE 0 layout)
template<int MODE, bool MIXED> constexpr size_t scatter_smem_bytes()
{
	return MODE == 0 && MIXED ? offsetof(ScatterSmem<0>, bin_of) : sizeof(ScatterSmem<MODE>);
}
#ifndef SIBGPU_SCATTER_OCC
#define SIBGPU_SCATTER_OCC 4
#endif

// cap != 0: partition b owns the fixed region [b * cap, (b + 1) * cap) of `out` and cursor[b] starts at b * cap (no
// histogram pass needed); a run that does not fit raises *overflow and is dropped -- the host then redoes the
// partitioning with exact sizes (k_scan_hist + k_part_offsets, cap == 0).
// MIXED: the record carries mix56(key) (8-byte records) / mix64(key) (16-byte records) instead of the key and the
// partition is a bit field of it (group_smem.cuh); fingerprint records (MODE 2) are mixed the same way but take their
// partition from the second hash
template<int MODE, bool MIXED>
__global__ void __launch_bounds__(TILE_THREADS, MODE == 0 ? (MIXED ? SIBGPU_SCATTER_OCC : 4) : (MODE == 1 ? 2 : 3)) k_scatter(TextDesc t, const FpView fv, uint32_t k,
	uint32_t ntiles, uint32_t P, unsigned long long *__restrict__ cursor, typename RecT<MODE>::type *__restrict__ out,
	unsigned long long cap, uint32_t *__restrict__ overflow)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	ScatterSmem<MODE> &s = *reinterpret_cast<ScatterSmem<MODE>*>(smem_raw);
	typedef cub::BlockScan<uint32_t, TILE_THREADS> Scan;
	__shared__ typename Scan::TempStorage scan_tmp;
	constexpr int BINS_PER_THREAD = MAX_PARTS / TILE_THREADS;
	constexpr bool NARROW = MODE != 1;                     // 8-byte records
	constexpr bool KEEP_BIN = !MIXED || MODE == 2;         // the partition cannot be read back from the record
	uint64_t *sD = nullptr;
	if constexpr(MODE == 2)
	{
		sD = s.sD;
		fp_stage_params(fv.prm, sD);
	}

	for(uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
	{
		for(uint32_t b = threadIdx.x; b < P; b += TILE_THREADS) s.cnt[b] = 0;
		if(threadIdx.x < MAX_PARTS / 32) s.dropmask[threadIdx.x] = 0;
		if(threadIdx.x == 0) s.anydrop = 0;
		stage_tile(t, t.tile0 + tile, s.sw);
		__syncthreads();

		uint64_t ra[POS_PER_THREAD], rb[NARROW ? 1 : POS_PER_THREAD];
		uint32_t binrank[POS_PER_THREAD];
		uint32_t valid = 0;
		scan16<MODE>(t, fv, s.sw, sD, t.tile0 + tile, k, [&](int i, uint64_t a, uint64_t b, uint32_t ctx) {
			if(MIXED) a = NARROW ? mix56(a) : mix64(a);
			const uint32_t bin = part_of<MODE, MIXED>(a, b, P);
			uint32_t rank = atomicAdd(&s.cnt[bin], 1u);
			binrank[i] = (bin << 16) | rank;             // rank < 4096, bin < 1024
			valid |= 1u << i;
			if(NARROW) { ra[i] = (a << 7) | (ctx & 127u); }
			else { ra[i] = a; rb[i] = (b << 8) | ctx; }
		});
		__syncthreads();

		// exclusive scan over the bins (4 per thread); the global runs are reserved (one returning atomic per bin) while
		// the records are put in bin order in shared memory -- the placement only needs the local offsets
		uint32_t c[BINS_PER_THREAD], o[BINS_PER_THREAD], sum = 0;
		unsigned long long g[BINS_PER_THREAD];
#pragma unroll
		for(int j = 0; j < BINS_PER_THREAD; j++)
		{
			uint32_t b = threadIdx.x * BINS_PER_THREAD + j;
			c[j] = b < P ? s.cnt[b] : 0u;
			sum += c[j];
		}
		uint32_t excl, total;
		Scan(scan_tmp).ExclusiveSum(sum, excl, total);
#pragma unroll
		for(int j = 0; j < BINS_PER_THREAD; j++)
		{
			uint32_t b = threadIdx.x * BINS_PER_THREAD + j;
			o[j] = excl;
			excl += c[j];
			g[j] = c[j] ? atomicAdd(&cursor[b * CURSOR_STRIDE], (unsigned long long)c[j]) : 0ull;
			if(b < P) s.cnt[b] = o[j];
		}
		if(threadIdx.x == 0) s.total = total;
		__syncthreads();

#pragma unroll
		for(int i = 0; i < POS_PER_THREAD; i++)
		{
			if(valid & (1u << i))
			{
				const uint32_t bin = binrank[i] >> 16, l = s.cnt[bin] + (binrank[i] & 0xFFFFu);
				s.a[l] = ra[i];
				if constexpr(!NARROW) s.b[l] = rb[i];
				if(KEEP_BIN) s.bin_of[l] = (uint16_t)bin;
			}
		}
#pragma unroll
		for(int j = 0; j < BINS_PER_THREAD; j++)
		{
			uint32_t b = threadIdx.x * BINS_PER_THREAD + j;
			if(b < P)
			{
				s.gbase[b] = (uint32_t)g[j] - o[j];
				if(cap && c[j] && g[j] + c[j] > (b + 1ull) * cap)
				{
					*overflow = 1u;
					atomicOr(&s.dropmask[b >> 5], 1u << (b & 31u));
					s.anydrop = 1u;
				}
			}
		}
		__syncthreads();

		// coalesced copy-out: consecutive l of the same bin go to consecutive global slots
		const uint32_t n = s.total;
		const bool drops = s.anydrop != 0;
		for(uint32_t l = threadIdx.x; l < n; l += TILE_THREADS)
		{
			const uint64_t a = s.a[l];
			const uint32_t bin = KEEP_BIN ? s.bin_of[l] : (MODE == 0 ? mixed_part(a >> 7, P) : __umulhi((uint32_t)(a >> 32), P));
			if(drops && ((s.dropmask[bin >> 5] >> (bin & 31u)) & 1u)) continue;
			const uint32_t gi = s.gbase[bin] + l;
			if constexpr(NARROW) { reinterpret_cast<uint64_t*>(out)[gi] = a; }
			else reinterpret_cast<ulonglong2*>(out)[gi] = make_ulonglong2(a, s.b[l]);
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K3: insert one partition into the L2-resident table.   K4: predicate + reset.
// ---------------------------------------------------------------------------------------------------------------
struct Slot8 { unsigned long long key; uint32_t pay; uint32_t pad; };             // 16 B

// one occurrence into the 16-byte (MODE 0) / 32-byte (MODE 1, 2) slot tables
template<int MODE> struct RecVal { typedef unsigned long long type; };
template<> struct RecVal<1> { typedef ulonglong2 type; };
template<int MODE>
__device__ __forceinline__ typename RecVal<MODE>::type load_rec(const void *__restrict__ recs, uint64_t i)
{
	return __ldcs(static_cast<const typename RecVal<MODE>::type*>(recs) + i);
}

// 16-byte slots {key, payload}: exact keys up to 64 bits (k <= 32).  A canonical key is never all ones (the reverse
// complement of T...T is A...A = 0), so EMPTY64 can mark a free slot even at k = 32.
__device__ __forceinline__ void insert_slot8(unsigned long long key, uint32_t ctx, void *__restrict__ table, uint32_t T)
{
	{
		Slot8 *tab = static_cast<Slot8*>(table);
		uint32_t slot = __umulhi((uint32_t)rec_hash(key, 0), T);
		uint32_t bits = payload_bits(ctx);
		for(;;)
		{
			// CAS first (no peek): measured 1.7x faster than load-then-CAS on B200 (tools/ubench/atomics.cu)
			unsigned long long old = atomicCAS(&tab[slot].key, EMPTY64, key);
			if(old == EMPTY64 || old == key)
			{
				if(old == key) bits |= PAY_MULTI;
				atomicAnd(&tab[slot].pay, ~bits);
				break;
			}
			slot = slot + 1 == T ? 0 : slot + 1;
		}
	}
}

template<int MODE>
__device__ __forceinline__ void insert_wide_rec(typename RecVal<MODE>::type rec, void *__restrict__ table, uint32_t T)
{
	if constexpr(MODE != 1) insert_slot8(rec >> 7, (uint32_t)rec & 127u, table, T);   // 56-bit key or fingerprint
	else insert_slot8(rec.x, (uint32_t)rec.y & 127u, table, T);      // exact 64-bit key
}

template<int MODE>
__global__ void __launch_bounds__(256) k_insert(const typename RecT<MODE>::type *__restrict__ recs, uint64_t n,
	void *__restrict__ table, uint32_t T)
{
	for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
	{
		insert_wide_rec<MODE>(load_rec<MODE>(recs, i), table, T);
	}
}

// Appends the canonical keys of the partition's bifurcation classes to `out` (the partition's own, now dead, record
// region) and resets the table for the next partition.
template<int MODE>
__global__ void __launch_bounds__(256) k_table_scan(void *__restrict__ table, uint32_t T,
	typename RecT<MODE>::type *__restrict__ out, uint32_t *__restrict__ counter)
{
	const uint32_t lane_id = threadIdx.x & 31u;
	for(uint32_t wbase = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; wbase < T; wbase += gridDim.x * blockDim.x)
	{
		const uint32_t sidx = wbase + lane_id;            // wbase is warp-uniform: the ballots below are convergent
		bool bif = false;
		unsigned long long a = 0, b = 0;
		if(sidx < T)
		{
			{
				Slot8 *tab = static_cast<Slot8*>(table);
				a = tab[sidx].key;
				if(a != EMPTY64)
				{
					bif = is_bifurcation(~tab[sidx].pay);
					tab[sidx].key = EMPTY64;
					tab[sidx].pay = ~0u;
				}
			}
		}
		// warp-aggregated append
		const uint32_t m = __ballot_sync(0xffffffffu, bif);
		if(m)
		{
			const uint32_t lane = threadIdx.x & 31u;
			uint32_t base = 0;
			if(lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(counter, (uint32_t)__popc(m));
			base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
			if(bif)
			{
				uint32_t idx = base + __popc(m & ((1u << lane) - 1));
				if(MODE != 1) reinterpret_cast<uint64_t*>(out)[idx] = a;
				else reinterpret_cast<ulonglong2*>(out)[idx] = make_ulonglong2(a, b);
			}
		}
	}
}

// Compact variant for k <= 26: key (<= 52 bits) and the 11 payload bits share ONE 64-bit slot, so an occurrence costs a
// single atomic: CAS when it creates the class, a fire-and-forget OR when the class exists.
constexpr uint32_t COMPACT_MAX_K = 26;
constexpr int INSERT_ILP = 4;                          // records in flight per thread
// VARIANT 0: peek, then CAS   1: CAS first   (x ILP = 1 or INSERT_ILP)
template<int VARIANT, int ILP>
__global__ void __launch_bounds__(256) k_insert_compact(const uint64_t *__restrict__ recs, uint64_t n,
	unsigned long long *__restrict__ tab, uint32_t T)
{
	const uint64_t chunk = (uint64_t)blockDim.x * ILP;
	for(uint64_t base = blockIdx.x * chunk; base < n; base += (uint64_t)gridDim.x * chunk)
	{
		unsigned long long word[ILP], old[ILP];
		uint32_t slot[ILP];
		bool live[ILP];
#pragma unroll
		for(int j = 0; j < ILP; j++)
		{
			const uint64_t i = base + (uint64_t)j * blockDim.x + threadIdx.x;      // coalesced per j
			live[j] = i < n;
			const uint64_t rec = live[j] ? __ldcs(recs + i) : 0ull;
			const unsigned long long key = rec >> 7;
			word[j] = (key << 11) | payload_bits((uint32_t)rec & 127u);
			slot[j] = __umulhi((uint32_t)rec_hash(key, 0), T);
		}
		if(VARIANT == 0)
		{
#pragma unroll
			for(int j = 0; j < ILP; j++) old[j] = live[j] ? __ldcg(&tab[slot[j]]) : 0ull;
#pragma unroll
			for(int j = 0; j < ILP; j++)
			{
				if(live[j] && old[j] == EMPTY64) old[j] = atomicCAS(&tab[slot[j]], EMPTY64, word[j]);
			}
		}
		else
		{
#pragma unroll
			for(int j = 0; j < ILP; j++)
			{
				if(live[j]) old[j] = atomicCAS(&tab[slot[j]], EMPTY64, word[j]);
			}
		}
#pragma unroll
		for(int j = 0; j < ILP; j++)
		{
			if(!live[j]) continue;
			for(;;)
			{
				if(old[j] == EMPTY64) break;                                         // created
				if((old[j] >> 11) == (word[j] >> 11))
				{
					atomicOr(&tab[slot[j]], (word[j] & 2047ull) | PAY_MULTI);        // existing class: OR the context in
					break;
				}
				slot[j] = slot[j] + 1 == T ? 0 : slot[j] + 1;
				if(VARIANT == 0)
				{
					old[j] = __ldcg(&tab[slot[j]]);
					if(old[j] == EMPTY64) old[j] = atomicCAS(&tab[slot[j]], EMPTY64, word[j]);
				}
				else old[j] = atomicCAS(&tab[slot[j]], EMPTY64, word[j]);
			}
		}
	}
}

static void launch_insert_compact(int variant, int sms, cudaStream_t st, const uint64_t *recs, uint64_t n, unsigned long long *tab, uint32_t T,
	int blocks_per_sm = 8)
{
	auto grid = [&](int ilp) {
		uint64_t blocks = ((n + ilp - 1) / ilp + 255) / 256, cap = (uint64_t)sms * blocks_per_sm;
		return (uint32_t)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
	};
	switch(variant)
	{
	case 0: k_insert_compact<0, 1><<<grid(1), 256, 0, st>>>(recs, n, tab, T); break;
	case 1: k_insert_compact<1, 1><<<grid(1), 256, 0, st>>>(recs, n, tab, T); break;
	case 2: k_insert_compact<1, INSERT_ILP><<<grid(INSERT_ILP), 256, 0, st>>>(recs, n, tab, T); break;
	default: k_insert_compact<0, INSERT_ILP><<<grid(INSERT_ILP), 256, 0, st>>>(recs, n, tab, T); break;
	}
}

// Sharded path: the records of one partition arrive as one segment per source rank, and a segment may live in the
// source rank's memory (CUDA IPC mapping, read over NVLink): the exchange is fused into the grouping kernel -- every
// lane streams its record straight from the peer's send buffer (coalesced 256-byte reads per warp) and inserts it
// into the local L2-resident table, so the transfer overlaps the atomics and no receive buffer is ever written.
constexpr int MAX_PEERS = 16;
struct SegList { const void *ptr[MAX_PEERS]; uint32_t cnt[MAX_PEERS]; };

__device__ __forceinline__ void insert_compact_rec(unsigned long long rec, unsigned long long *__restrict__ tab, uint32_t T)
{
	const unsigned long long key = rec >> 7;
	const unsigned long long word = (key << 11) | payload_bits((uint32_t)rec & 127u);
	uint32_t slot = __umulhi((uint32_t)rec_hash(key, 0), T);
	for(;;)
	{
		const unsigned long long old = atomicCAS(&tab[slot], EMPTY64, word);
		if(old == EMPTY64) break;
		if((old >> 11) == key)
		{
			atomicOr(&tab[slot], (word & 2047ull) | PAY_MULTI);
			break;
		}
		slot = slot + 1 == T ? 0 : slot + 1;
	}
}

// Variant with plain coalesced loads (SEG_ILP records in flight per thread).
template<int MODE, bool COMPACT, int SEG_ILP>
__global__ void __launch_bounds__(256) k_insert_seg_ld(const SegList segs, uint32_t nseg, uint64_t total,
	void *__restrict__ table, uint32_t T)
{
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for(uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * SEG_ILP)
	{
		typename RecVal<MODE>::type rec[SEG_ILP];
#pragma unroll
		for(int u = 0; u < SEG_ILP; u++)
		{
			uint64_t j = i0 + u * stride;
			if(j < total)
			{
				uint32_t sg = 0;
				while(sg + 1 < nseg && j >= segs.cnt[sg]) { j -= segs.cnt[sg]; sg++; }
				rec[u] = load_rec<MODE>(segs.ptr[sg], j);
			}
		}
#pragma unroll
		for(int u = 0; u < SEG_ILP; u++)
		{
			if(i0 + u * stride < total)
			{
				if(COMPACT) insert_compact_rec(*reinterpret_cast<unsigned long long*>(&rec[u]), static_cast<unsigned long long*>(table), T);
				else insert_wide_rec<MODE>(rec[u], table, T);
			}
		}
	}
}

// The segments are streamed through a ring of SEG_STAGES shared-memory tiles of SEG_TILE_BYTES filled by TMA bulk
// copies (cp.async.bulk + mbarrier complete_tx): one elected thread keeps up to SEG_STAGES tiles per CTA in flight,
// so the NVLink round trip of a peer read is paid by the copy engine of the SM, not by thread slots that should be
// issuing atomics.  Tile t of the launch = records [t', t' + n) of one segment; a segment starts 16-byte aligned.
constexpr int SEG_TILE_BYTES = 8192, SEG_STAGES = 2;

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template<int MODE, bool COMPACT>
__global__ void __launch_bounds__(256) k_insert_seg(const SegList segs, uint32_t nseg, uint32_t ntiles,
	void *__restrict__ table, uint32_t T)
{
	typedef typename RecVal<MODE>::type RV;
	constexpr uint32_t TILE_REC = SEG_TILE_BYTES / sizeof(RV);
	__shared__ __align__(128) unsigned char buf[SEG_STAGES][SEG_TILE_BYTES];
	__shared__ __align__(8) unsigned long long bar[SEG_STAGES];
	__shared__ uint32_t tile_n[SEG_STAGES];
	if(threadIdx.x == 0)
	{
		for(int st = 0; st < SEG_STAGES; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&bar[st])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// elected thread: start the bulk copy of tile t into stage st
	auto issue = [&](uint32_t t, int st) {
		uint32_t sg = 0, left = t;
		for(;;)
		{
			const uint32_t tiles_here = (segs.cnt[sg] + TILE_REC - 1) / TILE_REC;
			if(left < tiles_here || sg + 1 == nseg) break;
			left -= tiles_here;
			sg++;
		}
		const uint32_t first = left * TILE_REC;
		const uint32_t n = segs.cnt[sg] - first < TILE_REC ? segs.cnt[sg] - first : TILE_REC;
		const uint32_t bytes = (uint32_t)((n * sizeof(RV) + 15u) & ~15u);
		const unsigned char *src = static_cast<const unsigned char*>(segs.ptr[sg]) + (size_t)first * sizeof(RV);
		tile_n[st] = n;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(&bar[st])), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_addr(buf[st])), "l"(src), "r"(bytes), "r"(smem_addr(&bar[st])) : "memory");
	};
	if(threadIdx.x == 0)
	{
		uint32_t t = blockIdx.x;
		for(int st = 0; st < SEG_STAGES && t < ntiles; st++, t += gridDim.x) issue(t, st);
	}
	int stage = 0;
	uint32_t phase = 0;
	for(uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		uint32_t ok = 0;
		while(!ok)
		{
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
				: "=r"(ok) : "r"(smem_addr(&bar[stage])), "r"(phase) : "memory");
		}
		const uint32_t n = tile_n[stage];
		const RV *w = reinterpret_cast<const RV*>(buf[stage]);
#pragma unroll
		for(uint32_t u = 0; u < TILE_REC / 256; u++)
		{
			const uint32_t j = u * 256 + threadIdx.x;
			if(j < n)
			{
				RV rec = w[j];
				if(COMPACT) insert_compact_rec(*reinterpret_cast<unsigned long long*>(&rec), static_cast<unsigned long long*>(table), T);
				else insert_wide_rec<MODE>(rec, table, T);
			}
		}
		__syncthreads();                                   // every lane has read its records: the stage can be refilled
		if(threadIdx.x == 0)
		{
			const uint64_t tn = (uint64_t)t + (uint64_t)SEG_STAGES * gridDim.x;
			if(tn < ntiles) issue((uint32_t)tn, stage);
		}
		if(++stage == SEG_STAGES) { stage = 0; phase ^= 1u; }
	}
}

__global__ void __launch_bounds__(256) k_table_scan_compact(unsigned long long *__restrict__ tab, uint32_t T,
	uint64_t *__restrict__ out, uint32_t *__restrict__ counter)
{
	const uint32_t lane_id = threadIdx.x & 31u;
	for(uint32_t wbase = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; wbase < T; wbase += gridDim.x * blockDim.x)
	{
		const uint32_t sidx = wbase + lane_id;
		bool bif = false;
		unsigned long long w = EMPTY64;
		if(sidx < T)
		{
			w = tab[sidx];
			if(w != EMPTY64)
			{
				bif = is_bifurcation((uint32_t)w & 2047u);
				tab[sidx] = EMPTY64;
			}
		}
		const uint32_t m = __ballot_sync(0xffffffffu, bif);
		if(m)
		{
			uint32_t base = 0;
			if(lane_id == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(counter, (uint32_t)__popc(m));
			base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
			if(bif) out[base + __popc(m & ((1u << lane_id) - 1))] = w >> 11;
		}
	}
}

// exclusive scan of the per-partition vertex-class counts (one CTA)
__global__ void __launch_bounds__(MAX_PARTS) k_key_offsets(const uint32_t *__restrict__ cnt, uint32_t P,
	uint64_t *__restrict__ keyoff, uint64_t *__restrict__ scalars)
{
	typedef cub::BlockScan<uint64_t, MAX_PARTS> Scan;
	__shared__ typename Scan::TempStorage tmp;
	uint32_t c = threadIdx.x < P ? cnt[threadIdx.x] : 0u;
	uint64_t off, total;
	Scan(tmp).ExclusiveSum((uint64_t)c, off, total);
	if(threadIdx.x < P) keyoff[threadIdx.x] = off;
	if(threadIdx.x == 0) { keyoff[P] = total; scalars[2] = total; }
}

// gathers the per-partition key lists into one array; blockIdx.y = partition
template<int MODE>
__global__ void __launch_bounds__(256) k_gather_keys(const typename RecT<MODE>::type *__restrict__ recs,
	const uint64_t *__restrict__ partoff, const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ keyoff,
	typename RecT<MODE>::type *__restrict__ ckeys)
{
	const uint32_t p = blockIdx.y;
	const uint32_t n = cnt[p];
	const typename RecT<MODE>::type *src = recs + partoff[p];
	typename RecT<MODE>::type *dst = ckeys + keyoff[p];
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// fingerprint classes (MODE 2): key list entry = {fingerprint, partition}
__global__ void __launch_bounds__(256) k_gather_keys_fp(const uint64_t *__restrict__ recs, const uint64_t *__restrict__ partoff,
	const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ keyoff, Rec16 *__restrict__ ckeys, int mixed)
{
	const uint32_t p = blockIdx.y;
	const uint32_t n = cnt[p];
	const uint64_t *src = recs + partoff[p];
	Rec16 *dst = ckeys + keyoff[p];
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
		dst[i] = Rec16{mixed ? unmix56(src[i]) : src[i], (uint64_t)p};
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K5: vertex ids (exact modes): the vertex k-mers are {c, revcomp(c)}; id = rank in the sorted array
// ---------------------------------------------------------------------------------------------------------------
template<int MODE>
__global__ void __launch_bounds__(256) k_expand(const typename KeyT<MODE>::type *__restrict__ ckeys, uint64_t n, uint32_t k,
	uint64_t *__restrict__ vkeys, uint32_t *__restrict__ npal)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	uint64_t c = MODE == 0 ? reinterpret_cast<const uint64_t*>(ckeys)[i] : reinterpret_cast<const Rec16*>(ckeys)[i].a;
	uint64_t r = revcomp_key(c, k);
	vkeys[2 * i] = c;
	vkeys[2 * i + 1] = r == c ? EMPTY64 : r;           // sentinel sorts last (k = 32: revcomp of all-T is 0 != c)
	if(r == c) atomicAdd(npal, 1u);
}

__device__ __forceinline__ uint32_t lower_bound_u64(const uint64_t *__restrict__ v, uint32_t n, uint64_t x)
{
	uint32_t lo = 0, hi = n;
	while(lo < hi)
	{
		uint32_t mid = (lo + hi) >> 1;
		if(__ldg(v + mid) < x) lo = mid + 1; else hi = mid;
	}
	return lo;
}

// vertex map: canonical key -> (id of the canonical k-mer, id of its reverse complement); 32-byte slots.
// The one-hash bit filter in front of it is tested for EVERY text position by k_mark, so its index is a single
// multiplicative hash (top bits of the low product word), not the full mix the map slot uses.
__device__ __forceinline__ uint32_t filter_bit(unsigned long long a, unsigned long long b, uint32_t fshift)
{
	return (uint32_t)(((a ^ (b * 0xD6E8FEB86659FD93ull)) * 0x9E3779B97F4A7C15ull) >> fshift);
}

__device__ __forceinline__ void map_insert(MapSlot *map, uint32_t Tm, uint32_t *filter, uint32_t fshift,
	unsigned long long a, unsigned long long b, uint32_t idc, uint32_t idr, uint32_t cls)
{
	uint64_t h = rec_hash(a, b);
	uint32_t slot = __umulhi((uint32_t)h, Tm);
	for(;;)
	{
		unsigned long long old = atomicCAS(&map[slot].a, EMPTY64, a);
		if(old == EMPTY64) break;                      // canonical keys are distinct: no equal-key case
		slot = slot + 1 == Tm ? 0 : slot + 1;
	}
	map[slot].b = b;
	map[slot].idc = idc;
	map[slot].idr = idr;
	map[slot].cls = cls;
	const uint32_t bit = filter_bit(a, b, fshift);
	atomicOr(&filter[bit >> 5], 1u << (bit & 31u));
}

template<int MODE>
__global__ void __launch_bounds__(256) k_build_map(const typename KeyT<MODE>::type *__restrict__ ckeys, uint64_t n, uint32_t k,
	const uint64_t *__restrict__ vsorted, const uint32_t *__restrict__ npal, MapSlot *__restrict__ map, uint32_t Tm,
	uint32_t *__restrict__ filter, uint32_t fshift)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	const uint32_t V = MODE == 2 ? 0u : (uint32_t)(2 * n) - *npal;   // vertices = {c, revcomp(c)} minus the palindromes counted twice
	if(MODE == 2)
	{
		// fingerprint classes: ids are assigned after the lexicographic ranking (fingerprint.cu)
		Rec16 v = reinterpret_cast<const Rec16*>(ckeys)[i];
		map_insert(map, Tm, filter, fshift, v.a, v.b, 0u, 0u, (uint32_t)i);
		return;
	}
	uint64_t c = MODE == 0 ? reinterpret_cast<const uint64_t*>(ckeys)[i] : reinterpret_cast<const Rec16*>(ckeys)[i].a;
	uint32_t idc = lower_bound_u64(vsorted, V, c);
	uint32_t idr = lower_bound_u64(vsorted, V, revcomp_key(c, k));
	map_insert(map, Tm, filter, fshift, c, 0ull, idc, idr, (uint32_t)i);
}

__device__ __forceinline__ bool map_lookup(const MapSlot *__restrict__ map, uint32_t Tm, const uint32_t *__restrict__ filter,
	uint32_t fshift, unsigned long long a, unsigned long long b, uint32_t &idc, uint32_t &idr, uint32_t &cls)
{
	const uint32_t bit = filter_bit(a, b, fshift);
	if(!((__ldg(filter + (bit >> 5)) >> (bit & 31u)) & 1u)) return false;
	uint64_t h = rec_hash(a, b);
	uint32_t slot = __umulhi((uint32_t)h, Tm);
	for(;;)
	{
		const ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2*>(&map[slot]));
		if(kv.x == EMPTY64) return false;
		if(kv.x == a && kv.y == b)
		{
			const uint2 ids = __ldg(reinterpret_cast<const uint2*>(&map[slot].idc));
			idc = ids.x;
			idr = ids.y;
			cls = __ldg(&map[slot].cls);
			return true;
		}
		slot = slot + 1 == Tm ? 0 : slot + 1;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K6: mark vertex positions.   K7: emit the instance tables in text order.
// ---------------------------------------------------------------------------------------------------------------
#ifndef SIBGPU_MARK_UNROLL
#define SIBGPU_MARK_UNROLL 2
#endif
constexpr int MARK_UNROLL = SIBGPU_MARK_UNROLL;
template<int MODE>
__global__ void __launch_bounds__(TILE_THREADS) k_mark(TextDesc t, const FpView fv, uint32_t P, uint32_t k, uint32_t ntiles,
	const MapSlot *__restrict__ map, uint32_t Tm, const uint32_t *__restrict__ filter, uint32_t fshift,
	uint16_t *__restrict__ hitmask, uint64_t *__restrict__ tilecnt, unsigned long long *__restrict__ rep)
{
	__shared__ uint32_t sw[TILE_THREADS + 5];
	__shared__ uint64_t sD[MODE == 2 ? FP_SMEM_WORDS : 1];
	typedef cub::BlockReduce<uint32_t, TILE_THREADS> Red;
	__shared__ typename Red::TempStorage red_tmp;
	if(MODE == 2) fp_stage_params(fv.prm, sD);
	for(uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
	{
		__syncthreads();
		stage_tile(t, t.tile0 + tile, sw);
		__syncthreads();
		uint32_t mask = 0;
		scan16<MODE, MARK_UNROLL>(t, fv, sw, sD, t.tile0 + tile, k, [&](int i, uint64_t a, uint64_t b, uint32_t ctx) {
			uint32_t idc, idr, cls;
			if(MODE == 2) b = __umulhi((uint32_t)b, P);    // class key = {fingerprint, partition}
			if(map_lookup(map, Tm, filter, fshift, a, b, idc, idr, cls))
			{
				mask |= 1u << i;
				if(MODE == 2)
				{
					// representative occurrence of the class = its smallest text position (+ palindrome / forward flags)
					const uint32_t p = (t.tile0 + tile) * TILE_POS + threadIdx.x * POS_PER_THREAD + i;
					atomicMin(&rep[cls], ((unsigned long long)p << 2) | ((ctx >> 5) & 2u) | ((ctx >> 7) & 1u));
				}
			}
		});
		hitmask[(uint64_t)tile * TILE_THREADS + threadIdx.x] = (uint16_t)mask;
		uint32_t total = Red(red_tmp).Sum((uint32_t)__popc(mask));
		if(threadIdx.x == 0) tilecnt[tile] = total;
	}
}

template<int MODE>
__global__ void __launch_bounds__(TILE_THREADS) k_emit(TextDesc t, const FpView fv, uint32_t P, uint32_t k, uint32_t ntiles,
	const MapSlot *__restrict__ map, uint32_t Tm, const uint32_t *__restrict__ filter, uint32_t fshift,
	const uint16_t *__restrict__ hitmask, const uint64_t *__restrict__ tileoff,
	sibgpu_inst *__restrict__ pos_out, sibgpu_inst *__restrict__ neg_tmp, uint64_t inst_cap,
	const unsigned long long *__restrict__ rep, uint32_t *__restrict__ collision)
{
	typedef cub::BlockScan<uint32_t, TILE_THREADS> Scan;
	__shared__ typename Scan::TempStorage scan_tmp;
	__shared__ uint64_t sD[MODE == 2 ? FP_SMEM_WORDS : 1];
	__shared__ uint16_t hits[TILE_POS];                    // tile-local positions of the hits, in text order
	if(MODE == 2) fp_stage_params(fv.prm, sD);
	for(uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
	{
		// The hits of the tile are listed in shared memory first and then dealt out one per thread: the work per hit
		// (key or fingerprint, map probe, string verification) is long and the hits cluster, so a thread that walked
		// the hits of its own 16 positions would hold up its whole warp; consecutive hits also write consecutive rows.
		uint32_t mask = hitmask[(uint64_t)tile * TILE_THREADS + threadIdx.x];
		uint32_t rank, total;
		__syncthreads();
		Scan(scan_tmp).ExclusiveSum((uint32_t)__popc(mask), rank, total);
		while(mask)
		{
			hits[rank++] = (uint16_t)(threadIdx.x * POS_PER_THREAD + __ffs(mask) - 1);
			mask &= mask - 1;
		}
		__syncthreads();
		const uint32_t tile_p0 = (t.tile0 + tile) * TILE_POS;
		const uint64_t o0 = tileoff[tile];
		for(uint32_t h = threadIdx.x; h < total; h += TILE_THREADS)
		{
			const uint32_t p = tile_p0 + hits[h];
			const uint64_t o = o0 + h;
			unsigned long long a, b = 0;
			bool fw;
			if(MODE == 2)
			{
				uint64_t fa, hb;
				fp_at(t, fv, sD, p, k, fa, hb, fw);
				a = fa;
				b = __umulhi((uint32_t)hb, P);
			}
			else
			{
				uint64_t f = key_at(t, p, k), r = revcomp_key(f, k);
				fw = f <= r;
				a = fw ? f : r;
			}
			uint32_t idc = 0, idr = 0, cls = 0;
			map_lookup(map, Tm, filter, fshift, a, b, idc, idr, cls);
			if(MODE == 2)
			{
				// exactness: the k-mer here must spell the same string as the class representative; a fingerprint
				// collision that could change the result always surfaces here
				const unsigned long long R = rep[cls];
				const uint32_t rp = (uint32_t)(R >> 2), rfw = (uint32_t)R & 1u;
				bool same = true;
				for(uint32_t m = 0; m * 32 < k && same; m++)
				{
					same = vstr_chunk(t, p, fw ? 1u : 0u, m, k) == vstr_chunk(t, rp, rfw, m, k);
				}
				if(!same) atomicOr(collision, 1u);
			}
			ChrCursor cur;
			cur.init(t, p);
			const uint32_t c = cur.nc - 1, ppos = p - cur.cs, len = cur.ce - cur.cs;
			sibgpu_inst ip = {fw ? idc : idr, c, ppos};
			sibgpu_inst in = {fw ? idr : idc, c, len - ppos - k};
			if(o < inst_cap)                               // a full table is regrown by the host and the emission repeated
			{
				pos_out[o] = ip;
				neg_tmp[o] = in;
			}
		}
	}
}

// number of instances = offset + count of the last tile, clipped to the capacity of the tables
__device__ __forceinline__ uint64_t inst_total(const uint64_t *__restrict__ tileoff, const uint64_t *__restrict__ tilecnt,
	uint32_t ntiles, uint64_t inst_cap)
{
	const uint64_t n = tileoff[ntiles - 1] + tilecnt[ntiles - 1];
	return n < inst_cap ? n : inst_cap;
}

// chrinst[c] = first index in the (text-ordered) negative table whose chr >= c, for c in 0..nchr
__global__ void __launch_bounds__(256) k_chr_bounds(const sibgpu_inst *__restrict__ neg_tmp, const uint64_t *__restrict__ tileoff,
	const uint64_t *__restrict__ tilecnt, uint32_t ntiles, uint64_t inst_cap, uint32_t nchr, uint64_t *__restrict__ chrinst)
{
	uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c > nchr) return;
	uint64_t lo = 0, hi = inst_total(tileoff, tilecnt, ntiles, inst_cap);
	while(lo < hi)
	{
		uint64_t mid = (lo + hi) >> 1;
		if(neg_tmp[mid].chr < c) lo = mid + 1; else hi = mid;
	}
	chrinst[c] = lo;
}

// The negative-strand table is sorted by (chr, position in the reverse complement): within a chromosome that is
// descending text order, so every chromosome's run is reversed.
__global__ void __launch_bounds__(256) k_reverse_neg(const sibgpu_inst *__restrict__ neg_tmp, const uint64_t *__restrict__ tileoff,
	const uint64_t *__restrict__ tilecnt, uint32_t ntiles, uint64_t inst_cap, const uint64_t *__restrict__ chrinst,
	sibgpu_inst *__restrict__ neg_out)
{
	const uint64_t n = inst_total(tileoff, tilecnt, ntiles, inst_cap);
	for(uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x)
	{
		sibgpu_inst v = neg_tmp[j];
		neg_out[chrinst[v.chr] + chrinst[v.chr + 1] - 1 - j] = v;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------------------
int fingerprint_positions(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, uint32_t attempt, uint32_t w_begin, uint32_t w_stop);   // fingerprint.cu
int rank_fingerprint_vertices(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, uint64_t Vc, uint32_t Tm, uint32_t *V_out);   // fingerprint.cu

static inline uint32_t grid_for(uint64_t work_items, int threads, int sm_count, int waves = 8)
{
	uint64_t blocks = (work_items + threads - 1) / threads;
	uint64_t cap = (uint64_t)sm_count * waves;
	if(blocks < 1) blocks = 1;
	return (uint32_t)(blocks < cap ? blocks : cap);
}

// Vertex ids, vertex map and the two instance tables for the text tiles [t.tile0, t.tile0 + ntiles), given the
// canonical keys of ALL vertex classes (`ckeys`, Vc of them).  Shared by the single-GPU path and the sharded path.
// phase 0 = everything.  Sharded fingerprint runs (MODE 2) stop after k_mark (phase 1): the class representatives found
// in the own text range are then min-reduced over the ranks by the caller, and phase 2 resumes with the ranking.
template<int MODE>
static int ids_and_tables(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, const typename KeyT<MODE>::type *ckeys_in, uint64_t Vc,
	uint32_t ntiles, const FpView fp, uint32_t P, bool reverse_neg, bool *collision, int phase = 0)
{
	typedef typename KeyT<MODE>::type Rec;
	NvtxRange nvtx("sibgpu: vertex ids + instance tables");
	cudaStream_t st = ctx->stream;
	const int sms = ctx->sm_count;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	const uint32_t scan_grid = ntiles < (uint32_t)sms * 8 ? ntiles : (uint32_t)sms * 8;
	Rec *ckeys = const_cast<Rec*>(ckeys_in);
	*collision = false;
	// ---- vertex ids + vertex map
	uint64_t Tm64 = 2 * Vc + 64;
	const uint32_t Tm = (uint32_t)Tm64;
	uint32_t fbits_log = 16;
	while((1ull << fbits_log) < 32 * Vc && fbits_log < 32) fbits_log++;
	const uint32_t fshift = 64 - fbits_log;
	uint32_t V = 0;
	const bool front = phase != 2, back = phase != 1;
	if(front)
	{
		SIB_TRY(ctx->d_map.ensure(sizeof(MapSlot) * (size_t)Tm));
		SIB_TRY(ctx->d_filter.ensure((1ull << fbits_log) / 8));
		SIB_CUDA(cudaMemsetAsync(ctx->d_map.p, 0xFF, sizeof(MapSlot) * (size_t)Tm, st));
		SIB_CUDA(cudaMemsetAsync(ctx->d_filter.p, 0, (1ull << fbits_log) / 8, st));
	}
	if(!front) {}
	else if(MODE != 2)
	{
		SIB_TRY(ctx->d_vkeys.ensure(sizeof(uint64_t) * 2 * Vc));
		SIB_TRY(ctx->d_vkeys_alt.ensure(sizeof(uint64_t) * 2 * Vc));
		SIB_CUDA(cudaMemsetAsync(ds + 3, 0, sizeof(uint64_t), st));
		{
			ProfScope ps(ctx, "k_expand", Vc * (sizeof(Rec) + 16));
			k_expand<MODE><<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(ckeys, Vc, k,
				ctx->d_vkeys.as<uint64_t>(), reinterpret_cast<uint32_t*>(ds + 3));
		}
		size_t tmp_bytes = 0;
		cub::DoubleBuffer<uint64_t> dbuf(ctx->d_vkeys.as<uint64_t>(), ctx->d_vkeys_alt.as<uint64_t>());
		// keys have 2k significant bits; the palindrome sentinel (all ones) must still sort last: one more bit
		const int sort_bits = 2 * k + 1 < 64 ? (int)(2 * k + 1) : 64;
		SIB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, dbuf, (int)(2 * Vc), 0, sort_bits, st));
		SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
		{
			ProfScope ps(ctx, "cub_sort_vertex_keys", 2 * Vc * 8 * 2, 8);
			SIB_CUDA(cub::DeviceRadixSort::SortKeys(ctx->d_cubtmp.p, tmp_bytes, dbuf, (int)(2 * Vc), 0, sort_bits, st));
		}
		{
			// V = 2 Vc - #palindromes is read by the kernel from the device counter; the host learns it with the final sync
			ProfScope ps(ctx, "k_build_map", Vc * (sizeof(Rec) + sizeof(MapSlot)));
			k_build_map<MODE><<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(ckeys, Vc, k, dbuf.Current(), reinterpret_cast<uint32_t*>(ds + 3),
				ctx->d_map.as<MapSlot>(), Tm, ctx->d_filter.as<uint32_t>(), fshift);
		}
	}
	else
	{
		// classes only; the ids follow once every class has a representative occurrence (after k_mark).  "No occurrence
		// yet" = 0x7F7F...: larger than any {position, flags} word, also as the signed number an all-reduce takes it for
		SIB_TRY(ctx->d_rep.ensure(sizeof(uint64_t) * Vc));
		SIB_CUDA(cudaMemsetAsync(ctx->d_rep.p, 0x7F, sizeof(uint64_t) * Vc, st));
		ProfScope ps(ctx, "k_build_map", Vc * (sizeof(Rec) + sizeof(MapSlot)));
		k_build_map<MODE><<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(ckeys, Vc, k, nullptr, nullptr,
			ctx->d_map.as<MapSlot>(), Tm, ctx->d_filter.as<uint32_t>(), fshift);
	}

	if(ntiles == 0)
	{
		// a rank without text (tiny input, many ranks) still reports the global vertex count
		if(!back) return SIBGPU_OK;
		if(MODE == 2) SIB_TRY(rank_fingerprint_vertices(ctx, t, k, Vc, Tm, &V));
		SIB_CUDA(cudaMemcpyAsync(hs + 3, ds + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		ctx->n_inst = 0;
		ctx->n_vertices = MODE == 2 ? V : (uint32_t)(2 * Vc - (hs[3] & 0xFFFFFFFFu));
		return SIBGPU_OK;
	}
	// ---- instance tables
	if(front)
	{
		SIB_TRY(ctx->d_hitmask.ensure(sizeof(uint16_t) * (size_t)ntiles * TILE_THREADS));
		SIB_TRY(ctx->d_tilecnt.ensure(sizeof(uint64_t) * ntiles));
		SIB_TRY(ctx->d_tileoff.ensure(sizeof(uint64_t) * (ntiles + 1)));
		ProfScope ps(ctx, "k_mark", (MODE == 2 ? ctx->M / 2 + ctx->M * 2 : ctx->M / 4) * ntiles / ((ctx->M + TILE_POS - 1) / TILE_POS) + ctx->M / 8);
		k_mark<MODE><<<scan_grid, TILE_THREADS, 0, st>>>(t, fp, P, k, ntiles, ctx->d_map.as<MapSlot>(), Tm,
			ctx->d_filter.as<uint32_t>(), fshift, ctx->d_hitmask.as<uint16_t>(), ctx->d_tilecnt.as<uint64_t>(),
			ctx->d_rep.as<unsigned long long>());
	}
	if(!back) return SIBGPU_OK;
	if(MODE == 2) SIB_TRY(rank_fingerprint_vertices(ctx, t, k, Vc, Tm, &V));
	{
		size_t tmp_bytes = 0;
		const uint64_t *in = ctx->d_tilecnt.as<uint64_t>();
		SIB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, ctx->d_tileoff.as<uint64_t>(), (int)ntiles, st));
		SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
		ProfScope ps(ctx, "cub_scan_tile_counts", ntiles * 12ull, 2);
		SIB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, tmp_bytes, in, ctx->d_tileoff.as<uint64_t>(), (int)ntiles, st));
	}
	// The tables are emitted into the capacity at hand (grow-only buffers: the previous call's size, at least 64 Ki
	// instances) without asking the device for the instance count first; if they turn out too small the count is known
	// by then, the buffers are regrown and only the emission is repeated.
	SIB_TRY(ctx->d_chrinst.ensure(sizeof(uint64_t) * (ctx->nchr + 2)));
	SIB_TRY(ctx->d_pos.ensure(sizeof(sibgpu_inst) * 65536));
	SIB_TRY(ctx->d_negtmp.ensure(sizeof(sibgpu_inst) * 65536));
	SIB_TRY(ctx->d_neg.ensure(sizeof(sibgpu_inst) * 65536));
	uint64_t I = 0;
	for(int pass = 0; pass < 2; pass++)
	{
		const uint64_t inst_cap = std::min(std::min(ctx->d_pos.cap, ctx->d_negtmp.cap), ctx->d_neg.cap) / sizeof(sibgpu_inst);
		{
			ProfScope ps(ctx, "k_emit", ctx->M / 8 + I * 24);
			k_emit<MODE><<<scan_grid, TILE_THREADS, 0, st>>>(t, fp, P, k, ntiles, ctx->d_map.as<MapSlot>(), Tm,
				ctx->d_filter.as<uint32_t>(), fshift, ctx->d_hitmask.as<uint16_t>(), ctx->d_tileoff.as<uint64_t>(),
				ctx->d_pos.as<sibgpu_inst>(), ctx->d_negtmp.as<sibgpu_inst>(), inst_cap, ctx->d_rep.as<unsigned long long>(),
				reinterpret_cast<uint32_t*>(ds + 9));
		}
		if(reverse_neg)
		{
			{
				ProfScope ps(ctx, "k_chr_bounds", 0);
				k_chr_bounds<<<(ctx->nchr + 1 + 255) / 256, 256, 0, st>>>(ctx->d_negtmp.as<sibgpu_inst>(), ctx->d_tileoff.as<uint64_t>(),
					ctx->d_tilecnt.as<uint64_t>(), ntiles, inst_cap, ctx->nchr, ctx->d_chrinst.as<uint64_t>());
			}
			ProfScope ps(ctx, "k_reverse_neg", I * 24);
			k_reverse_neg<<<(uint32_t)sms * 8, 256, 0, st>>>(ctx->d_negtmp.as<sibgpu_inst>(), ctx->d_tileoff.as<uint64_t>(),
				ctx->d_tilecnt.as<uint64_t>(), ntiles, inst_cap, ctx->d_chrinst.as<uint64_t>(), ctx->d_neg.as<sibgpu_inst>());
		}
		SIB_CUDA(cudaMemcpyAsync(hs + 4, ctx->d_tileoff.as<uint64_t>() + (ntiles - 1), 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaMemcpyAsync(hs + 5, ctx->d_tilecnt.as<uint64_t>() + (ntiles - 1), 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaMemcpyAsync(hs + 3, ds + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaMemcpyAsync(hs + 9, ds + 9, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		I = hs[4] + hs[5];
		if(I <= inst_cap) break;
		SIB_TRY(ctx->d_pos.ensure(sizeof(sibgpu_inst) * (I + 1)));
		SIB_TRY(ctx->d_negtmp.ensure(sizeof(sibgpu_inst) * (I + 1)));
		SIB_TRY(ctx->d_neg.ensure(sizeof(sibgpu_inst) * (I + 1)));
	}
	if(MODE != 2) V = (uint32_t)(2 * Vc - (hs[3] & 0xFFFFFFFFu));
	if(MODE == 2 && (hs[9] & 1u))
	{
		// two different k-mers shared a fingerprint inside a vertex class: the caller starts over with other bases
		SIB_CUDA(cudaMemsetAsync(ds + 9, 0, sizeof(uint64_t), st));
		*collision = true;
		return SIBGPU_OK;
	}
	ctx->n_inst = I;
	ctx->n_vertices = V;
	return SIBGPU_OK;
}

// Copies the bytes [lo, hi) of the concatenated text from the caller's chromosomes (the separators are already there).
int copy_text_range(sibgpu_ctx *ctx, const HostSrc &src, uint64_t lo, uint64_t hi, cudaStream_t st)
{
	const std::vector<uint32_t> &cs = ctx->h_chr_start;
	size_t c = std::upper_bound(cs.begin(), cs.end(), (uint32_t)lo) - cs.begin();
	if(c > 0) c--;
	for(; c < ctx->nchr && cs[c] < hi; c++)
	{
		const uint64_t s = cs[c], e = s + ctx->h_chr_len[c];
		const uint64_t a = s > lo ? s : lo, b = e < hi ? e : hi;
		if(b > a) SIB_CUDA(cudaMemcpyAsync(ctx->d_text.as<char>() + a, src.chr[c] + (a - s), b - a, cudaMemcpyHostToDevice, st));
	}
	return SIBGPU_OK;
}

// Pageable host sources.  cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread at
// ~9 GB/s; here `nthreads` host threads copy the pieces of text piece c into slot c % STAGE_SLOTS of a pinned ring while
// the earlier pieces are on their way to the device, and the copies to the device start from the ring.
struct HostStager {
	struct Piece { uint64_t text_off; const char *src; uint64_t bytes; };
	sibgpu_ctx *ctx;
	std::vector<std::vector<Piece> > pieces;           // per text piece: the chromosome stretches it consists of
	std::vector<uint64_t> lo;                          // first text byte of every piece
	std::unique_ptr<std::atomic<uint32_t>[]> done;     // threads that have staged piece c
	std::atomic<uint32_t> free_upto{0};                // pieces [0, free_upto) may be written into their ring slots
	std::atomic<bool> quit{false};
	std::vector<std::thread> th;
	uint32_t nthreads = 0;

	static bool pageable(const void *p)
	{
		cudaPointerAttributes a;
		if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
		{
			cudaGetLastError();
			return true;
		}
		return a.type == cudaMemoryTypeUnregistered;
	}
	int start(sibgpu_ctx *c, const HostSrc &src, uint32_t nchunks, uint64_t chunk_bytes)
	{
		ctx = c;
		nthreads = (uint32_t)std::max(1, std::min<int>(c->stage_threads, (int)std::thread::hardware_concurrency()));
		SIB_TRY(c->ensure_stage(chunk_bytes));
		pieces.resize(nchunks);
		lo.resize(nchunks);
		const std::vector<uint32_t> &cs = c->h_chr_start;
		for(uint32_t k = 0; k < nchunks; k++)
		{
			const uint64_t a0 = (uint64_t)k * chunk_bytes, a1 = k + 1 == nchunks ? c->M : a0 + chunk_bytes;
			lo[k] = a0;
			size_t ch = std::upper_bound(cs.begin(), cs.end(), (uint32_t)a0) - cs.begin();
			if(ch > 0) ch--;
			for(; ch < c->nchr && cs[ch] < a1; ch++)
			{
				const uint64_t s = cs[ch], e = s + c->h_chr_len[ch];
				const uint64_t a = s > a0 ? s : a0, b = e < a1 ? e : a1;
				if(b > a) pieces[k].push_back(Piece{a, src.chr[ch] + (a - s), b - a});
			}
		}
		done.reset(new std::atomic<uint32_t>[nchunks]);
		for(uint32_t k = 0; k < nchunks; k++) done[k].store(0);
		free_upto.store(std::min<uint32_t>(nchunks, sibgpu_ctx::STAGE_SLOTS));
		for(uint32_t t = 0; t < nthreads; t++)
		{
			th.emplace_back([this, t, nchunks]() {
				for(uint32_t k = 0; k < nchunks; k++)
				{
					while(free_upto.load(std::memory_order_acquire) <= k) std::this_thread::yield();
					if(quit.load(std::memory_order_relaxed)) return;
					char *slot = static_cast<char*>(ctx->h_stage) + (size_t)(k % sibgpu_ctx::STAGE_SLOTS) * ctx->stage_piece;
					for(const Piece &p : pieces[k])
					{
						// stripes of whole cache lines
						const uint64_t lines = (p.bytes + 63) / 64;
						const uint64_t b = lines * t / nthreads * 64, e = std::min<uint64_t>(p.bytes, lines * (t + 1) / nthreads * 64);
						if(e > b) memcpy(slot + (p.text_off - lo[k]) + b, p.src + b, e - b);
					}
					done[k].fetch_add(1, std::memory_order_release);
				}
			});
		}
		return SIBGPU_OK;
	}
	// waits for piece k to be staged, starts its copies to the device on `st`; ev[k] must be recorded on `st` by the caller
	// afterwards.  Before that, the slot of piece k + 1 is released once its previous occupant has left for the device.
	int issue(uint32_t k, cudaStream_t st, const std::vector<cudaEvent_t> &ev)
	{
		if(k + 1 >= sibgpu_ctx::STAGE_SLOTS)
		{
			SIB_CUDA(cudaEventSynchronize(ev[k + 1 - sibgpu_ctx::STAGE_SLOTS]));
			free_upto.store(k + 2, std::memory_order_release);
		}
		while(done[k].load(std::memory_order_acquire) < nthreads) std::this_thread::yield();
		const char *slot = static_cast<const char*>(ctx->h_stage) + (size_t)(k % sibgpu_ctx::STAGE_SLOTS) * ctx->stage_piece;
		for(const Piece &p : pieces[k])
		{
			SIB_CUDA(cudaMemcpyAsync(ctx->d_text.as<char>() + p.text_off, slot + (p.text_off - lo[k]), p.bytes, cudaMemcpyHostToDevice, st));
		}
		return SIBGPU_OK;
	}
	~HostStager()
	{
		quit.store(true);
		free_upto.store(0xFFFFFFFFu, std::memory_order_release);   // nobody waits any more (error paths)
		for(std::thread &x : th) x.join();
	}
};

static int launch_pack(sibgpu_ctx *ctx, uint64_t w0, uint64_t w1)
{
	uint32_t *d_err = reinterpret_cast<uint32_t*>(ctx->d_scalars.as<uint64_t>() + 8);
	ProfScope ps(ctx, "k_pack", (w1 - w0) * 20);
	k_pack<<<grid_for(w1 - w0, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(ctx->d_text.as<uint4>() + w0,
		ctx->d_packed.as<uint32_t>() + w0, (uint32_t)(w1 - w0), d_err);
	return SIBGPU_OK;
}

static int input_error()
{
	set_error("input: a character outside ACGT reached the device; sanitise first (indexedsequence.cpp:31-37)");
	return SIBGPU_ERR_INPUT;
}

// k_split over P1 owned partitions read from ssrc (single GPU: the own record buffer; sharded: all ranks' buffers);
// R = uint64_t (8-byte records) or ulonglong2 (16-byte records)
template<class R>
static int launch_split(sibgpu_ctx *ctx, const SplitSrc &ssrc, uint32_t P1, uint32_t tiles_per_seg, uint32_t sub_bits,
	uint64_t nrec, uint32_t *d_flags)
{
	bool &attr_done = ctx->split_attr_done[sizeof(R) == 8 ? 0 : 1];
	if(!attr_done)
	{
		SIB_CUDA(cudaFuncSetAttribute(k_split<1, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<1, R>)));
		SIB_CUDA(cudaFuncSetAttribute(k_split<2, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem<2, R>)));
		attr_done = true;
	}
	const uint64_t split_tiles = (uint64_t)P1 * ssrc.W * tiles_per_seg;
	const uint64_t sms = (uint64_t)ctx->sm_count;
	ProfScope ps(ctx, "k_split", nrec * 2 * sizeof(R));
	if(ctx->split_stages == 1)
	{
		k_split<1, R><<<(uint32_t)std::min<uint64_t>(split_tiles, sms * 3), SPLIT_THREADS, sizeof(SplitSmem<1, R>), ctx->stream>>>(
			ssrc, P1, tiles_per_seg, sub_bits, ctx->d_records2.as<R>(), ctx->d_cnt2.as<uint32_t>(), GROUP_CAP, d_flags);
	}
	else
	{
		k_split<2, R><<<(uint32_t)std::min<uint64_t>(split_tiles, sms * (sizeof(R) == 8 ? SPLIT_OCC : 2)), SPLIT_THREADS, sizeof(SplitSmem<2, R>), ctx->stream>>>(
			ssrc, P1, tiles_per_seg, sub_bits, ctx->d_records2.as<R>(), ctx->d_cnt2.as<uint32_t>(), GROUP_CAP, d_flags);
	}
	return SIBGPU_OK;
}

// PKEY: the key list takes {key, partition} (fingerprint classes) instead of the bare key
template<class R, bool PKEY = false>
static int launch_group(sibgpu_ctx *ctx, uint32_t nbuckets, uint32_t sub_bits, uint64_t nrec, uint32_t *d_flags,
	typename GroupKey<R, PKEY>::type *ckeys, uint32_t ckeys_cap, uint32_t *d_nkeys, uint32_t part0 = 0)
{
	bool &attr_done = ctx->group_attr_done[(sizeof(R) == 8 ? 0 : 1) + (PKEY ? 2 : 0)];
	if(!attr_done)
	{
		SIB_CUDA(cudaFuncSetAttribute(k_group<R, PKEY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem<R>)));
		attr_done = true;
	}
	ProfScope ps(ctx, "k_group", nrec * sizeof(R));
	k_group<R, PKEY><<<std::min<uint32_t>(nbuckets, (uint32_t)ctx->sm_count * (sizeof(R) == 8 ? 4 : 2)), GROUP_THREADS, sizeof(GroupSmem<R>),
		ctx->stream>>>(ctx->d_records2.as<R>(), ctx->d_cnt2.as<uint32_t>(), nbuckets, sub_bits, part0, GROUP_CAP, d_flags, ckeys, ckeys_cap, d_nkeys);
	return SIBGPU_OK;
}

// src != nullptr: the text is still on the host (sibgpu_enumerate); it is streamed in CHUNK_TILES-tile pieces on the
// copy stream while the pack and scatter kernels of the previous pieces run (exact modes; the layout, the '$'
// separators and the chromosome tables are already on the device).  src == nullptr: text resident and packed.
constexpr uint32_t CHUNK_TILES = 2048;                 // 8 Mi text positions per piece

template<int MODE>
static int enumerate_mode(sibgpu_ctx *ctx, uint32_t k, uint64_t nrec, const HostSrc *src)
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	const int sms = ctx->sm_count;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();

	TextDesc t;
	t.packed = ctx->d_packed.as<uint32_t>();
	t.chr_start = ctx->d_chr_start.as<uint32_t>();
	t.chr_len = ctx->d_chr_len.as<uint32_t>();
	t.nchr = ctx->nchr;
	t.M = (uint32_t)ctx->M;
	t.nwords = (uint32_t)((ctx->M + 15) / 16) + 8;
	t.tile0 = 0;
	const uint32_t ntiles = (uint32_t)((ctx->M + TILE_POS - 1) / TILE_POS);
	typedef typename std::conditional<MODE != 1, uint64_t, ulonglong2>::type SR;   // record type of the shared-memory path
	typedef typename KeyT<MODE>::type Key;                 // entry of the vertex-key list
	constexpr bool FP = MODE == 2;
	const size_t scatter_smem = scatter_smem_bytes<MODE, false>(), scatter_smem_mixed = scatter_smem_bytes<MODE, true>();
	SIB_CUDA(cudaFuncSetAttribute(k_scatter<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem));
	SIB_CUDA(cudaFuncSetAttribute(k_scatter<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_smem_mixed));

	for(uint32_t attempt = 0; ; attempt++)
	{
		if(FP) SIB_TRY(fingerprint_positions(ctx, t, k, attempt, 0u, (uint32_t)((ctx->M + 15) >> 4)));
		const FpView fp = {ctx->d_fp.as<FpCk>(), ctx->d_fpprm.as<FpParams>()};

		// ---- partition plan
		// level-1 partitions of 512 Ki records, split into ~1 Ki-record buckets and grouped in shared memory
		// (group_smem.cuh); otherwise (SIBGPU_GROUP_SMEM=0, fallbacks) one L2-resident table per partition
		const bool smem_group = ctx->group_smem && !ctx->exact_hist;
		const uint64_t part_rec = smem_group && !ctx->part_explicit ? (uint64_t)SIBGPU_SMEM_PART_KI << 10 : ctx->part_records(k);
		uint64_t P64 = (nrec + part_rec - 1) / part_rec;
		const uint32_t P = (uint32_t)(P64 < 1 ? 1 : (P64 > MAX_PARTS ? MAX_PARTS : P64));
		bool mixed = false;                                // the records carry mix56(key) (shared-memory path)
		bool grouped = false;                              // vertex keys already in d_ckeys (shared-memory path)
		uint64_t Vc = 0;
		SIB_TRY(ctx->d_hist.ensure(sizeof(uint32_t) * MAX_PARTS));
		SIB_TRY(ctx->d_partoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
		SIB_TRY(ctx->d_cursor.ensure(sizeof(uint64_t) * MAX_PARTS * CURSOR_STRIDE));
		SIB_TRY(ctx->d_partcnt.ensure(sizeof(uint32_t) * MAX_PARTS));
		SIB_TRY(ctx->d_keyoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
		SIB_CUDA(cudaMemsetAsync(ctx->d_partcnt.p, 0, sizeof(uint32_t) * MAX_PARTS, st));
		std::vector<uint64_t> part_base(P + 1), part_cnt(P);
		uint64_t maxpart = 0;
		bool have_records = false;

		// ---- fast path: no histogram pass.  Hash partitions of distinct k-mers are balanced to a few sigma, so every
		// partition gets a fixed region of mean + 1/8 (+ 4096) records; a partition that outgrows it (a k-mer repeated
		// millions of times) raises the overflow flag and the exact two-pass partitioning below takes over.
		if(!ctx->exact_hist)
		{
			const uint64_t mean = (nrec + P - 1) / P;
			uint32_t sub_bits = 0;                         // level-2 fan-out of the shared-memory grouping
			while(((mean + ((uint64_t)1 << sub_bits) - 1) >> sub_bits) > GROUP_MEAN) sub_bits++;
			mixed = smem_group && sub_bits <= SUB_BITS_MAX;
			// Pieces: with a host source the text arrives in CHUNK_TILES-tile pieces.  On the shared-memory path every piece
			// scatters into its OWN set of partition regions and is split into the (global) buckets right away, so that
			// when the last byte has landed only the last piece's scatter + split, the grouping and the second scan are
			// left; otherwise all pieces share one set of regions.
			const uint32_t nchunks = src ? (ntiles + CHUNK_TILES - 1) / CHUNK_TILES : 1;
			const bool piecewise = mixed && nchunks > 1 && ctx->piecewise_split;
			const uint32_t nreg = piecewise ? nchunks : 1;     // sets of partition regions
			const uint64_t piece_mean = ((uint64_t)(CHUNK_TILES + 2) * TILE_POS + P - 1) / P;
			const uint64_t cap = piecewise ? ((piece_mean + piece_mean / 8 + ctx->part_slack) + 1) & ~1ull
				: ((P == 1 ? nrec : mean + mean / 8 + ctx->part_slack) + 1) & ~1ull;   // even: 16-byte aligned partitions
			SIB_TRY(ctx->d_records.ensure(sizeof(Rec) * cap * P * nreg + 64));
			SIB_TRY(ctx->d_cursor.ensure(sizeof(uint64_t) * (size_t)P * CURSOR_STRIDE * nreg));
			for(uint32_t p = 0; p <= P; p++) part_base[p] = (uint64_t)p * cap;
			// cursors are relative to the base of their set of regions
			std::vector<uint64_t> cur((size_t)P * CURSOR_STRIDE * nreg, 0);
			for(uint32_t r = 0; r < nreg; r++)
			{
				for(uint32_t p = 0; p < P; p++) cur[((size_t)r * P + p) * CURSOR_STRIDE] = part_base[p];
			}
			SIB_CUDA(cudaMemcpyAsync(ctx->d_cursor.p, cur.data(), sizeof(uint64_t) * cur.size(), cudaMemcpyHostToDevice, st));
			SIB_CUDA(cudaMemcpyAsync(ctx->d_partoff.p, part_base.data(), sizeof(uint64_t) * (P + 1), cudaMemcpyHostToDevice, st));
			uint32_t *d_overflow = reinterpret_cast<uint32_t*>(ds + 10);
			uint32_t *d_grp_flags = reinterpret_cast<uint32_t*>(ds + 11);
			uint32_t nbuckets = 0;
			uint32_t ckeys_cap = 0;
			const uint32_t tiles_per_seg = (uint32_t)((cap + RecOps<SR>::TILE - 1) / RecOps<SR>::TILE);
			if(mixed)
			{
				nbuckets = P << sub_bits;
				SIB_TRY(ctx->d_records2.ensure(sizeof(SR) * (size_t)nbuckets * GROUP_CAP + 64));
				SIB_TRY(ctx->d_cnt2.ensure(sizeof(uint32_t) * (size_t)nbuckets));
				SIB_TRY(ctx->d_ckeys.ensure(sizeof(Key) * (ctx->ckeys_init ? ctx->ckeys_init : 1)));
				ckeys_cap = (uint32_t)std::min<size_t>(ctx->d_ckeys.cap / sizeof(Key), 0xFFFFFFF0u);
				SIB_CUDA(cudaMemsetAsync(ctx->d_cnt2.p, 0, sizeof(uint32_t) * (size_t)nbuckets, st));
			}
			auto split_regions = [&](uint32_t r, uint64_t nrec_here) -> int {
				SplitSrc ssrc = {};
				ssrc.seg[0] = ctx->d_records.as<Rec>() + (size_t)r * P * cap;
				ssrc.cursor[0] = ctx->d_cursor.as<unsigned long long>() + (size_t)r * P * CURSOR_STRIDE;
				ssrc.seg_cap = cap;
				ssrc.W = 1;
				return launch_split<SR>(ctx, ssrc, P, tiles_per_seg, sub_bits, nrec_here, d_grp_flags);
			};
			HostStager stager;
			bool staged = false;
			if(src)
			{
				SIB_TRY(ctx->ensure_copy_stream(nchunks));
				SIB_CUDA(cudaEventRecord(ctx->ev_fork_copy, st));              // the '$' fill and the tables precede the copies
				SIB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork_copy, 0));
				const char *first = nullptr;
				for(uint32_t c = 0; c < ctx->nchr && !first; c++) first = ctx->h_chr_len[c] ? src->chr[c] : nullptr;
				staged = ctx->stage_threads > 0 && nchunks > 1 && first && HostStager::pageable(first);
				if(staged) SIB_TRY(stager.start(ctx, *src, nchunks, (uint64_t)CHUNK_TILES * TILE_POS));
				for(uint32_t c = 0; c < nchunks && !staged; c++)
				{
					const uint64_t lo = (uint64_t)c * CHUNK_TILES * TILE_POS;
					const uint64_t hi = c + 1 == nchunks ? ctx->M : lo + (uint64_t)CHUNK_TILES * TILE_POS;
					SIB_TRY(copy_text_range(ctx, *src, lo, hi, ctx->copy_stream));
					SIB_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
				}
			}
			uint32_t tiles_done = 0;
			for(uint32_t c = 0; c < nchunks; c++)
			{
				uint32_t tile_hi = ntiles;
				if(staged)
				{
					SIB_TRY(stager.issue(c, ctx->copy_stream, ctx->ev_chunk));
					SIB_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
				}
				if(src)
				{
					// pack the words of this piece; scatter the tiles whose staged words (one behind, 260 ahead) are packed
					const uint64_t w0 = (uint64_t)c * CHUNK_TILES * TILE_THREADS;
					const uint64_t w1 = c + 1 == nchunks ? t.nwords : w0 + (uint64_t)CHUNK_TILES * TILE_THREADS;
					SIB_CUDA(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
					SIB_TRY(launch_pack(ctx, w0, w1));
					if(c + 1 < nchunks) tile_hi = (uint32_t)((w1 - (TILE_THREADS + 5)) / TILE_THREADS);
				}
				if(tile_hi > tiles_done)
				{
					const uint32_t nt = tile_hi - tiles_done;
					const uint32_t occ = mixed && MODE == 0 ? SIBGPU_SCATTER_OCC : 4;
					const uint32_t g = nt < (uint32_t)sms * occ ? nt : (uint32_t)sms * occ;
					const uint32_t r = piecewise ? c : 0;
					TextDesc tc = t;
					tc.tile0 = tiles_done;
					{
						ProfScope ps(ctx, "k_scatter", (FP ? (uint64_t)nt * TILE_POS / 2 + (uint64_t)nt * TILE_POS * 2 : (uint64_t)nt * TILE_POS / 4)
							+ nrec * sizeof(Rec) * nt / ntiles);
						unsigned long long *cursor_r = ctx->d_cursor.as<unsigned long long>() + (size_t)r * P * CURSOR_STRIDE;
						Rec *out_r = ctx->d_records.as<Rec>() + (size_t)r * P * cap;
						if(mixed) k_scatter<MODE, true><<<g, TILE_THREADS, scatter_smem_mixed, st>>>(tc, fp, k, nt, P, cursor_r, out_r, cap, d_overflow);
						else k_scatter<MODE, false><<<g, TILE_THREADS, scatter_smem, st>>>(tc, fp, k, nt, P, cursor_r, out_r, cap, d_overflow);
					}
					if(piecewise) SIB_TRY(split_regions(r, nrec * nt / ntiles));
					tiles_done = tile_hi;
				}
			}
			// ---- shared-memory grouping, launched behind the scatter without a host round trip: the level-1 fill counts
			// are read on the device; the flags (input error, level-1 / level-2 overflow) are checked once, below
			bool smem_launched = false;
			if(mixed)
			{
				if(!piecewise) SIB_TRY(split_regions(0, nrec));
				SIB_TRY((launch_group<SR, FP>(ctx, nbuckets, sub_bits, nrec, d_grp_flags, ctx->d_ckeys.as<typename GroupKey<SR, FP>::type>(), ckeys_cap,
					reinterpret_cast<uint32_t*>(ds + 2))));
				smem_launched = true;
			}
			SIB_CUDA(cudaMemcpyAsync(cur.data(), ctx->d_cursor.p, sizeof(uint64_t) * cur.size(), cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaMemcpyAsync(hs + 8, ds + 8, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaMemcpyAsync(hs + 2, ds + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaStreamSynchronize(st));
			if(hs[8] & 1u) return input_error();
			src = nullptr;                                 // the whole text is resident and packed from here on
			if(hs[10] & 0xFFFFFFFFull)
			{
				SIB_CUDA(cudaMemsetAsync(ds + 10, 0, sizeof(uint64_t) * 2, st));
				SIB_CUDA(cudaMemsetAsync(ds + 2, 0, sizeof(uint64_t), st));
				ctx->hist_fallbacks++;
				mixed = false;                             // the exact path below scatters plain records again
			}
			else
			{
				uint64_t total = 0;
				for(uint32_t p = 0; p < P; p++)
				{
					part_cnt[p] = 0;
					for(uint32_t r = 0; r < nreg; r++) part_cnt[p] += cur[((size_t)r * P + p) * CURSOR_STRIDE] - part_base[p];
					total += part_cnt[p];
					if(part_cnt[p] > maxpart) maxpart = part_cnt[p];
				}
				if(total != nrec)
				{
					set_error("internal: scatter kernel wrote " + std::to_string(total) + " k-mers, expected " + std::to_string(nrec));
					return SIBGPU_ERR_INTERNAL;
				}
				have_records = !piecewise;                 // one contiguous region per partition (what the L2-table path reads)
				if(smem_launched)
				{
					if(hs[11] & 0xFFFFFFFFull)
					{
						// a bucket outgrew its fixed region (one k-mer repeated hundreds of times): the L2-table path below
						// regroups the intact level-1 partitions (scattered again, exactly sized, when they lie in pieces)
						SIB_CUDA(cudaMemsetAsync(ds + 11, 0, sizeof(uint64_t), st));
						SIB_CUDA(cudaMemsetAsync(ds + 2, 0, sizeof(uint64_t), st));
						ctx->smem_fallbacks++;
						if(piecewise) mixed = false;
					}
					else
					{
						Vc = hs[2] & 0xFFFFFFFFull;
						if(Vc > ckeys_cap)
						{
							// the key list was too small (k_group kept counting): regrow, group again -- the buckets are intact
							SIB_TRY(ctx->d_ckeys.ensure(sizeof(Key) * Vc));
							SIB_CUDA(cudaMemsetAsync(ds + 2, 0, sizeof(uint64_t), st));
							SIB_TRY((launch_group<SR, FP>(ctx, nbuckets, sub_bits, nrec, d_grp_flags, ctx->d_ckeys.as<typename GroupKey<SR, FP>::type>(),
								(uint32_t)Vc, reinterpret_cast<uint32_t*>(ds + 2))));
							SIB_CUDA(cudaStreamSynchronize(st));
						}
						grouped = true;
					}
				}
			}
		}
		else if(src)
		{
			SIB_TRY(copy_text_range(ctx, *src, 0, ctx->M, st));
			SIB_TRY(launch_pack(ctx, 0, t.nwords));
			SIB_CUDA(cudaMemcpyAsync(hs + 8, ds + 8, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaStreamSynchronize(st));
			if(hs[8] & 1u) return input_error();
			src = nullptr;
		}

		// ---- exact path: histogram pass, prefix sums, scatter into exactly sized partitions
		if(!have_records && !grouped)
		{
			SIB_TRY(ctx->d_records.ensure(sizeof(Rec) * nrec));
			SIB_CUDA(cudaMemsetAsync(ctx->d_hist.p, 0, sizeof(uint32_t) * MAX_PARTS, st));
			const uint32_t scan_grid = ntiles < (uint32_t)sms * 8 ? ntiles : (uint32_t)sms * 8;
			{
				ProfScope ps(ctx, "k_scan_hist", FP ? ctx->M / 2 + ctx->M * 2 : ctx->M / 4);
				k_scan_hist<MODE><<<scan_grid, TILE_THREADS, 0, st>>>(t, fp, k, ntiles, P, ctx->d_hist.as<uint32_t>());
			}
			{
				ProfScope ps(ctx, "k_part_offsets", 0);
				k_part_offsets<<<1, MAX_PARTS, 0, st>>>(ctx->d_hist.as<uint32_t>(), P, ctx->d_partoff.as<uint64_t>(),
					ctx->d_cursor.as<unsigned long long>(), ds);
			}
			SIB_CUDA(cudaMemcpyAsync(hs, ds, sizeof(uint64_t) * 2, cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaMemcpyAsync(part_base.data(), ctx->d_partoff.p, sizeof(uint64_t) * (P + 1), cudaMemcpyDeviceToHost, st));
			SIB_CUDA(cudaStreamSynchronize(st));
			if(hs[0] != nrec)
			{
				set_error("internal: scan kernel counted " + std::to_string(hs[0]) + " k-mers, expected " + std::to_string(nrec));
				return SIBGPU_ERR_INTERNAL;
			}
			maxpart = hs[1];
			for(uint32_t p = 0; p < P; p++) part_cnt[p] = part_base[p + 1] - part_base[p];
			{
				const uint32_t g = ntiles < (uint32_t)sms * 4 ? ntiles : (uint32_t)sms * 4;
				ProfScope ps(ctx, "k_scatter", (FP ? ctx->M / 2 + ctx->M * 2 : ctx->M / 4) + nrec * sizeof(Rec));
				k_scatter<MODE, false><<<g, TILE_THREADS, scatter_smem, st>>>(t, fp, k, ntiles, P, ctx->d_cursor.as<unsigned long long>(),
					ctx->d_records.as<Rec>(), 0ull, nullptr);
			}
		}

		// ---- per-partition L2-resident grouping
		if(!grouped)
		{
		uint64_t T64 = (uint64_t)ctx->table_factor * maxpart + 1024;
		if(T64 > 0xFFFFFF00ull)
		{
			set_error("internal: hash partition of " + std::to_string(maxpart) + " records does not fit a 32-bit table");
			return SIBGPU_ERR_INTERNAL;
		}
		const uint32_t T = (uint32_t)T64;
		const bool compact = MODE == 0 && k <= COMPACT_MAX_K && !mixed;   // a mixed key needs all 56 bits
		const size_t slot_bytes = compact ? 8 : sizeof(Slot8);
		// S independent tables on S streams: consecutive partitions overlap, so the ramp-up / tail of one partition's
		// kernels is filled by its neighbours (with S = 1 everything runs on the main stream and is timed per launch)
		const uint32_t S = ctx->n_streams < 1 ? 1 : (ctx->n_streams > 8 ? 8 : ctx->n_streams);
		const size_t table_bytes = (slot_bytes * T + 255) / 256 * 256;
		SIB_TRY(ctx->d_table.ensure(table_bytes * S));
		SIB_CUDA(cudaMemsetAsync(ctx->d_table.p, 0xFF, table_bytes * S, st));
		if(S > 1)
		{
			SIB_TRY(ctx->ensure_aux_streams(S));
			SIB_CUDA(cudaEventRecord(ctx->ev_fork, st));
			for(uint32_t i = 0; i < S; i++) SIB_CUDA(cudaStreamWaitEvent(ctx->aux_stream[i], ctx->ev_fork, 0));
		}
		const int blocks_per_sm = S > 1 ? 4 : 8;
		{
			// with overlapped streams the two kernels are timed together, as one phase on the main stream's timeline
			const bool phase_span = S > 1 && ctx->profiling;
			if(phase_span) ctx->prof_begin("k_insert+k_table_scan", nrec * sizeof(Rec));
			for(uint32_t p = 0; p < P; p++)
			{
				const uint64_t n = part_cnt[p];
				if(n == 0) continue;
				Rec *part = ctx->d_records.as<Rec>() + part_base[p];
				cudaStream_t ps_st = S > 1 ? ctx->aux_stream[p % S] : st;
				void *table = static_cast<char*>(ctx->d_table.p) + table_bytes * (p % S);
				ctx->total_launches += 2;
				if(S == 1 && ctx->profiling) ctx->prof_begin("k_insert", n * sizeof(Rec));
				if(compact) launch_insert_compact(ctx->insert_variant, sms, ps_st, reinterpret_cast<const uint64_t*>(part), n,
					static_cast<unsigned long long*>(table), T, blocks_per_sm);
				else k_insert<MODE><<<grid_for(n, 256, sms, blocks_per_sm), 256, 0, ps_st>>>(part, n, table, T);
				if(S == 1 && ctx->profiling) { ctx->prof_end(); ctx->prof_begin("k_table_scan", (uint64_t)T * slot_bytes); }
				if(compact) k_table_scan_compact<<<grid_for(T, 256, sms, blocks_per_sm), 256, 0, ps_st>>>(
					static_cast<unsigned long long*>(table), T, reinterpret_cast<uint64_t*>(part), ctx->d_partcnt.as<uint32_t>() + p);
				else k_table_scan<MODE><<<grid_for(T, 256, sms, blocks_per_sm), 256, 0, ps_st>>>(table, T, part,
					ctx->d_partcnt.as<uint32_t>() + p);
				if(S == 1 && ctx->profiling) ctx->prof_end();
			}
			if(S > 1)
			{
				for(uint32_t i = 0; i < S; i++)
				{
					SIB_CUDA(cudaEventRecord(ctx->ev_join[i], ctx->aux_stream[i]));
					SIB_CUDA(cudaStreamWaitEvent(st, ctx->ev_join[i], 0));
				}
			}
			if(phase_span) ctx->prof_end();
		}
		{
			ProfScope ps(ctx, "k_key_offsets", 0);
			k_key_offsets<<<1, MAX_PARTS, 0, st>>>(ctx->d_partcnt.as<uint32_t>(), P, ctx->d_keyoff.as<uint64_t>(), ds);
		}
		SIB_CUDA(cudaMemcpyAsync(hs + 2, ds + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		Vc = hs[2];                                        // canonical vertex classes
		}
		if(Vc == 0)
		{
			ctx->n_inst = 0;
			ctx->n_vertices = 0;
			return SIBGPU_OK;
		}
		if(2 * Vc > 0xFFFFFFF0ull)
		{
			set_error("invalid: more than 2^32 vertices");
			return SIBGPU_ERR_INVALID;
		}
		if(!grouped)
		{
			SIB_TRY(ctx->d_ckeys.ensure(sizeof(Key) * Vc));
			ProfScope ps(ctx, "k_gather_keys", Vc * (sizeof(Rec) + sizeof(Key)));
			dim3 g(8, P);
			if constexpr(FP)
			{
				k_gather_keys_fp<<<g, 256, 0, st>>>(ctx->d_records.as<uint64_t>(), ctx->d_partoff.as<uint64_t>(),
					ctx->d_partcnt.as<uint32_t>(), ctx->d_keyoff.as<uint64_t>(), ctx->d_ckeys.as<Rec16>(), mixed ? 1 : 0);
			}
			else
			{
				k_gather_keys<MODE><<<g, 256, 0, st>>>(ctx->d_records.as<Rec>(), ctx->d_partoff.as<uint64_t>(),
					ctx->d_partcnt.as<uint32_t>(), ctx->d_keyoff.as<uint64_t>(), ctx->d_ckeys.as<Rec>());
				if(mixed) k_unmix<SR><<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(ctx->d_ckeys.as<SR>(), Vc);
			}
		}

		bool collision = false;
		SIB_TRY(ids_and_tables<MODE>(ctx, t, k, ctx->d_ckeys.as<Key>(), Vc, ntiles, fp, P, true, &collision));
		if(collision)
		{
			if(attempt >= 2)
			{
				set_error("internal: fingerprint verification failed three times");
				return SIBGPU_ERR_INTERNAL;
			}
			continue;
		}
		return SIBGPU_OK;
	}
}

int enumerate_resident(sibgpu_ctx *ctx, uint32_t k, const HostSrc *src)
{
	NvtxRange nvtx("sibgpu: enumerate");
	cudaStream_t st = ctx->stream;
	ctx->have_result = false;
	ctx->prof_reset();
	ctx->total_launches = 0;
	SIB_CUDA(cudaSetDevice(ctx->device));

	const uint32_t nwords = (uint32_t)((ctx->M + 15) / 16) + 8;
	SIB_TRY(ctx->d_packed.ensure(sizeof(uint32_t) * (size_t)nwords));
	SIB_TRY(ctx->d_scalars.ensure(sizeof(uint64_t) * 64));
	SIB_CUDA(cudaMemsetAsync(ctx->d_scalars.p, 0, sizeof(uint64_t) * 64, st));
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);

	uint64_t nrec = 0;
	for(uint32_t c = 0; c < ctx->nchr; c++)
	{
		if(ctx->h_chr_len[c] >= k) nrec += ctx->h_chr_len[c] - k + 1;
	}
	ctx->last_k = k;
	// K0: pack (+ legality).  With a host source and an exact mode the packing is pipelined with the upload inside
	// enumerate_mode; otherwise the whole text is brought in (if needed) and packed here.
	const bool pipelined = src && nrec > 0 && k <= 32;
	if(!pipelined)
	{
		if(src)
		{
			const uint64_t chunk_bytes = (uint64_t)CHUNK_TILES * TILE_POS;
			const uint32_t nchunks = (uint32_t)((ctx->M + chunk_bytes - 1) / chunk_bytes);
			const char *first = nullptr;
			for(uint32_t c = 0; c < ctx->nchr && !first; c++) first = ctx->h_chr_len[c] ? src->chr[c] : nullptr;
			if(ctx->stage_threads > 0 && nchunks > 1 && first && HostStager::pageable(first))
			{
				HostStager stager;
				SIB_TRY(ctx->ensure_copy_stream(nchunks));
				SIB_TRY(stager.start(ctx, *src, nchunks, chunk_bytes));
				for(uint32_t c = 0; c < nchunks; c++)
				{
					SIB_TRY(stager.issue(c, st, ctx->ev_chunk));
					SIB_CUDA(cudaEventRecord(ctx->ev_chunk[c], st));
				}
			}
			else SIB_TRY(copy_text_range(ctx, *src, 0, ctx->M, st));
		}
		src = nullptr;
		SIB_TRY(launch_pack(ctx, 0, nwords));
		SIB_CUDA(cudaMemcpyAsync(hs + 8, ctx->d_scalars.as<uint64_t>() + 8, 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		if(hs[8] & 1u) return input_error();
	}
	int rc = SIBGPU_OK;
	if(nrec == 0)
	{
		ctx->n_inst = 0;
		ctx->n_vertices = 0;
	}
	else if(k <= 28) rc = enumerate_mode<0>(ctx, k, nrec, src);
	else if(k <= 32) rc = enumerate_mode<1>(ctx, k, nrec, src);
	else rc = enumerate_mode<2>(ctx, k, nrec, nullptr);
	if(rc != SIBGPU_OK) return rc;
	SIB_CUDA(cudaGetLastError());
	ctx->have_result = true;
	if(ctx->profiling) SIB_TRY(ctx->prof_collect());
	return SIBGPU_OK;
}

#include "dist_impl.cuh"

} // namespace sibgpu
