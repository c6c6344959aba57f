// Shared-memory grouping of the k-mer records (included by enumerate.cu; 8-byte records for k <= 28 and for the
// fingerprints of k > 32, 16-byte records for k = 29..32).
//
// The hash partitions written by k_scatter (level 1, ~256 Ki records each) are split once more into buckets of ~1 Ki
// records (k_split: one coalesced read + one coalesced write of every record), and each bucket is then grouped
// entirely inside one SM (k_group): a TMA bulk copy (cp.async.bulk + mbarrier) brings the bucket's records into shared
// memory, every record claims the slot of its key in a shared-memory open-addressing table of 32-bit entries
// {20-bit tag, index of the record that created the class} with a shared CAS and ORs its predecessor/successor
// symbols into the 32-bit payload of the class's first record; that record then evaluates the reference's predicate
// (vertexenumeration.cpp:67-70,330,348) on the final payload, and the bifurcation k-mers are appended to the global
// key list by a warp-ballot aggregated atomic.
//
// Records of this path carry the key MIXED by a bijection of the 56-bit key space (mix56): partition, bucket, table
// slot and tag are then plain bit fields of the record -- the hash is computed once, in k_scatter, and never again --
// and equal keys stay equal; the (few) vertex keys are un-mixed when they are appended (unmix56).
//
// Shared atomics run at 4.4 (CAS.32) / 2.6 (CAS.64) / 8.7 (OR.32, ADD.32) lane-operations per clock per SM on B200
// (tools/ubench/smem.cu, profiles/r2_smem_atomics.txt) against ~0.35 per clock per SM for CAS on an L2-resident table.
#pragma once

namespace sibgpu {

constexpr uint64_t MIX_MASK = (1ull << 56) - 1;
constexpr uint64_t MIX_C1 = 0x51afd7ed558ccdull, MIX_C2 = 0xceb9fe1a85ec53ull;       // odd
constexpr uint64_t MIX_I1 = 0x74430c22a54005ull, MIX_I2 = 0xb4b2f8129337dbull;       // inverses mod 2^56

// bijection of [0, 2^56): two rounds of xor-shift (by half the width: an involution) and odd multiplication
__host__ __device__ __forceinline__ uint64_t mix56(uint64_t x)
{
	x ^= x >> 28;
	x = (x * MIX_C1) & MIX_MASK;
	x ^= x >> 28;
	x = (x * MIX_C2) & MIX_MASK;
	x ^= x >> 28;
	return x;
}
__host__ __device__ __forceinline__ uint64_t unmix56(uint64_t x)
{
	x ^= x >> 28;
	x = (x * MIX_I2) & MIX_MASK;
	x ^= x >> 28;
	x = (x * MIX_I1) & MIX_MASK;
	x ^= x >> 28;
	return x;
}
// bit fields of a mixed key m:  level-1 partition = top bits, bucket = bits 0-9, table slot = bits 10-21, tag = bits 22-41
__device__ __forceinline__ uint32_t mixed_part(uint64_t m, uint32_t P) { return __umulhi((uint32_t)(m >> 24), P); }

// Wide records (16 bytes: k = 29..32 exact keys, k > 32 fingerprints) carry mix64(key), the murmur3 finaliser, which is a
// bijection of the 64-bit words: xor-shift by 33 is an involution, the multipliers are odd.
constexpr uint64_t MIX64_I1 = 0x4f74430c22a54005ull, MIX64_I2 = 0x9cb4b2f8129337dbull;   // inverses of the two multipliers mod 2^64
__host__ __device__ __forceinline__ uint64_t unmix64(uint64_t x)
{
	x ^= x >> 33;
	x *= MIX64_I2;
	x ^= x >> 33;
	x *= MIX64_I1;
	x ^= x >> 33;
	return x;
}

// Record formats of the shared-memory path.  narrow: (mixed 56-bit key << 7) | context;  wide: {mixed 64-bit key, context}.
// Bit fields of the mixed key: bucket = bits 0-9, table slot = bits 10-21, tag = bits 22-41, partition = top bits.
template<class R> struct RecOps;
constexpr int SPLIT_OCC = 2;                           // CTAs per SM of the two-stage k_split
template<> struct RecOps<uint64_t> {
	// records per k_split tile (32 KB).  Measured: 2048-record tiles at 3 CTAs per SM are slower (0.76 vs 0.59 ms per
	// 10^8 records): the runs a tile contributes to a bucket halve and the writes fall below a sector pair
	static constexpr int TILE = 4096;
	static __device__ __forceinline__ uint64_t key(uint64_t r) { return r >> 7; }
	static __device__ __forceinline__ uint32_t ctx(uint64_t r) { return (uint32_t)r & 127u; }
	static __device__ __forceinline__ bool same_key(uint64_t a, uint64_t b) { return (a >> 7) == (b >> 7); }
};
template<> struct RecOps<ulonglong2> {
	static constexpr int TILE = 2048;
	static __device__ __forceinline__ uint64_t key(const ulonglong2 &r) { return r.x; }
	static __device__ __forceinline__ uint32_t ctx(const ulonglong2 &r) { return (uint32_t)r.y & 127u; }
	static __device__ __forceinline__ bool same_key(const ulonglong2 &a, const ulonglong2 &b) { return a.x == b.x; }
};

constexpr int SPLIT_THREADS = 512;
constexpr int SPLIT_MAX_BINS = 1024;
constexpr uint32_t GROUP_THREADS = 256;
constexpr uint32_t GROUP_SLOTS = 4096;                 // shared-memory table slots (load <= 0.41)
constexpr uint32_t GROUP_MEAN = 1024;                  // target records per bucket
constexpr uint32_t GROUP_CAP = 1664;                   // fixed capacity of a bucket's region: mean + 1/2 + 128 (even: 16-byte aligned)
constexpr uint32_t GROUP_STAGES = 2;
constexpr uint32_t SUB_BITS_MAX = 10;
constexpr uint32_t EMPTY32 = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(void *bar, uint32_t parity)
{
	uint32_t ok = 0;
	while(!ok)
	{
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
			: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
	}
}
// elected thread: bulk copy of `bytes` (multiple of 16, source 16-byte aligned) into shared memory, completing `bar`
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, void *bar)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
	if(bytes)
	{
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
	}
}

template<int STAGES, class R> struct SplitSmem {
	R in[STAGES][RecOps<R>::TILE];                     // tiles as they lie in the partition (TMA destinations)
	R sorted[RecOps<R>::TILE];                         // the current tile ordered by bucket
	uint32_t cnt[SPLIT_MAX_BINS];                      // per-bin count of this tile, then exclusive local offset
	uint32_t gbase[SPLIT_MAX_BINS];                    // index of the bin's run in the partition's level-2 region minus the local offset
	uint32_t dropmask[SPLIT_MAX_BINS / 32];            // bins whose run did not fit (overflow: the caller discards the run)
	uint32_t anydrop;
	uint32_t tile_n[STAGES], tile_p[STAGES];
	unsigned long long bar[STAGES];
};

// Where the level-1 partitions lie.  Single GPU: one source, the context's own record buffer.  Sharded: every rank
// scatters the records of its text range into its own exported buffer (one fixed-capacity segment per GLOBAL partition)
// and the owner of a partition reads that partition's segment out of every rank's buffer -- the exchange is fused into
// this kernel: the TMA bulk copies read the peers' memory over NVLink (CUDA IPC mappings).
constexpr int SPLIT_MAX_SRC = 16;
constexpr uint32_t GRP_BUCKET_OVERFLOW = 1u, GRP_PEER_FAILED = 4u, GRP_TIMEOUT = 8u;
struct DistHeader {                                    // first 256 bytes of a rank's exported buffer
	unsigned long long epoch_scatter;                  // == e once the segments and fill cursors of step e are final
	unsigned long long scatter_flags;                  // != 0: a segment overflowed / illegal input character (step failed)
	unsigned long long epoch_keys;                     // == e once the vertex keys of step e are final
	unsigned long long key_flags;                      // != 0: grouping failed on this rank (bucket overflow, peer failure)
	unsigned long long nkeys;
	unsigned long long epoch_packed;                   // == e once the packed words of the own text range are final (k > 32)
	unsigned long long pad[26];
};
struct SplitSrc {
	const void *seg[SPLIT_MAX_SRC];                    // records of source s; partition q lies at [q * seg_cap, ...)
	const unsigned long long *cursor[SPLIT_MAX_SRC];   // fill cursors of source s (absolute record index, one per CURSOR_STRIDE)
	const DistHeader *header[SPLIT_MAX_SRC];           // nullptr: no waiting (own buffer on a single GPU)
	unsigned long long seg_cap;
	unsigned long long epoch;
	uint32_t W, p0;                                    // number of sources, first owned partition
	uint32_t rot;                                      // source visited first (sharded: the own rank, so that the ranks
	                                                   // pull from different peers at any one time)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}
// spins until *flag == epoch (a peer publishes its step); false after ~4 s (the peer died: never hang the GPU)
__device__ __forceinline__ bool wait_epoch(const unsigned long long *flag, unsigned long long epoch)
{
	if(ld_acquire_sys(flag) == epoch) return true;
	const unsigned long long t0 = global_ns();
	for(;;)
	{
		if(ld_acquire_sys(flag) == epoch) return true;
		if(global_ns() - t0 > 4000000000ull) return false;
		__nanosleep(200);
	}
}

// Level 2: the records of owned partition p (from every source) -> B2 buckets of fixed capacity cap2 at
// out[(p * B2 + b) * cap2].
// A tile is 32 KB of consecutive records of one segment: TMA bulk copy into shared memory (STAGES tiles in flight: the
// copies of the following tiles run while this one is sorted and written out), counting sort by bucket (rank =
// returning shared atomic), coalesced copy-out of the runs; cnt2[p * B2 + b] is the bucket's fill count.  Consecutive
// tile indices belong to different partitions, so concurrently running CTAs bump different bucket counters.
// Segments must be 16-byte aligned (seg_cap even).
template<int STAGES, class R>
__global__ void __launch_bounds__(SPLIT_THREADS, STAGES == 1 ? 3 : (sizeof(R) == 8 ? SPLIT_OCC : 2)) k_split(const SplitSrc src, uint32_t P1, uint32_t tiles_per_seg,
	uint32_t sub_bits, R *__restrict__ out, uint32_t *__restrict__ cnt2, uint32_t cap2, uint32_t *__restrict__ overflow)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SplitSmem<STAGES, R> &s = *reinterpret_cast<SplitSmem<STAGES, R>*>(smem_raw);
	constexpr uint32_t SPLIT_TILE = RecOps<R>::TILE;
	constexpr int SPLIT_PER_THREAD = SPLIT_TILE / SPLIT_THREADS;
	typedef cub::BlockScan<uint32_t, SPLIT_THREADS> Scan;
	__shared__ typename Scan::TempStorage scan_tmp;
	const uint32_t B2 = 1u << sub_bits, sub_mask = B2 - 1u;
	const uint32_t ntiles = P1 * src.W * tiles_per_seg;
	if(threadIdx.x == 0)
	{
		for(int st = 0; st < STAGES; st++) mbar_init(&s.bar[st], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// elected thread: first non-empty tile at or after t (stride gridDim.x); starts its bulk copy into stage st and
	// publishes (n, p) there; n = 0: no tile left.  Tile t = (partition t % P1, source (t / P1 + rot) % W, chunk t / P1 / W).
	uint32_t seen = 0;                                 // sources whose step flag this CTA has already observed
	auto fetch = [&](uint32_t t, int st) -> uint32_t {
		for(; t < ntiles; t += gridDim.x)
		{
			const uint32_t p = t % P1, x = t / P1, sidx = (x + src.rot) % src.W, chunk = x / src.W;
			if(src.header[sidx] && !((seen >> sidx) & 1u))
			{
				if(!wait_epoch(&src.header[sidx]->epoch_scatter, src.epoch))
				{
					atomicOr(overflow, GRP_TIMEOUT);
					break;
				}
				if(ld_relaxed_sys(&src.header[sidx]->scatter_flags))
				{
					atomicOr(overflow, GRP_PEER_FAILED);
					break;
				}
				seen |= 1u << sidx;
			}
			const uint64_t base = (uint64_t)(src.p0 + p) * src.seg_cap;
			uint64_t np = ld_relaxed_sys(src.cursor[sidx] + (size_t)(src.p0 + p) * CURSOR_STRIDE) - base;
			if(np > src.seg_cap) np = src.seg_cap;     // an overflowed segment: the caller discards this run anyway
			const uint64_t first = (uint64_t)chunk * SPLIT_TILE;
			if(first >= np) continue;
			const uint32_t n = np - first < SPLIT_TILE ? (uint32_t)(np - first) : SPLIT_TILE;
			s.tile_n[st] = n;
			s.tile_p[st] = p;
			bulk_load(s.in[st], static_cast<const R*>(src.seg[sidx]) + base + first, (n * (uint32_t)sizeof(R) + 15u) & ~15u, &s.bar[st]);
			return t;
		}
		s.tile_n[st] = 0;
		return ntiles;
	};
	uint32_t t_cur = ntiles;
	if(threadIdx.x == 0)
	{
		t_cur = fetch(blockIdx.x, 0);
		for(int st = 1; st < STAGES; st++) t_cur = fetch(t_cur + gridDim.x, st);
	}
	uint32_t stage = 0, phase = 0;
	for(;;)
	{
		for(uint32_t b = threadIdx.x; b < B2; b += SPLIT_THREADS) s.cnt[b] = 0;
		if(threadIdx.x < SPLIT_MAX_BINS / 32) s.dropmask[threadIdx.x] = 0;
		if(threadIdx.x == 0) s.anydrop = 0;
		__syncthreads();                               // also publishes tile_n / tile_p of the fetch
		const uint32_t n = s.tile_n[stage], p = s.tile_p[stage];
		if(n == 0) break;
		mbar_wait(&s.bar[stage], phase);
		const R *in = s.in[stage];

		uint32_t rank[SPLIT_PER_THREAD];
#pragma unroll
		for(int j = 0; j < SPLIT_PER_THREAD; j++)
		{
			const uint32_t i = j * SPLIT_THREADS + threadIdx.x;
			if(i < n) rank[j] = atomicAdd(&s.cnt[(uint32_t)RecOps<R>::key(in[i]) & sub_mask], 1u);
		}
		__syncthreads();

		// exclusive scan over the bins (2 per thread); the runs are reserved in the buckets while the tile is sorted
		uint32_t c[2], g[2], sum = 0;
#pragma unroll
		for(int j = 0; j < 2; j++)
		{
			const uint32_t b = threadIdx.x * 2 + j;
			c[j] = b < B2 ? s.cnt[b] : 0u;
			sum += c[j];
		}
		uint32_t excl;
		Scan(scan_tmp).ExclusiveSum(sum, excl);
#pragma unroll
		for(int j = 0; j < 2; j++)
		{
			const uint32_t b = threadIdx.x * 2 + j;
			g[j] = c[j] ? atomicAdd(&cnt2[(size_t)p * B2 + b], c[j]) : 0u;
			if(b < B2) s.cnt[b] = excl + (j ? c[0] : 0u);
		}
		__syncthreads();
#pragma unroll
		for(int j = 0; j < SPLIT_PER_THREAD; j++)
		{
			const uint32_t i = j * SPLIT_THREADS + threadIdx.x;
			if(i < n)
			{
				const R rec = in[i];
				s.sorted[s.cnt[(uint32_t)RecOps<R>::key(rec) & sub_mask] + rank[j]] = rec;
			}
		}
#pragma unroll
		for(int j = 0; j < 2; j++)
		{
			const uint32_t b = threadIdx.x * 2 + j;
			if(b < B2)
			{
				s.gbase[b] = b * cap2 + g[j] - (excl + (j ? c[0] : 0u));   // 32-bit wrap-around arithmetic: + local index >= offset
				if(g[j] + c[j] > cap2)
				{
					atomicOr(overflow, GRP_BUCKET_OVERFLOW);
					atomicOr(&s.dropmask[b >> 5], 1u << (b & 31u));
					s.anydrop = 1u;
				}
			}
		}
		__syncthreads();                               // this stage's input is consumed: the tile after the ones in flight may land
		if(threadIdx.x == 0) t_cur = fetch(t_cur + gridDim.x, stage);

		R *dst = out + (uint64_t)p * B2 * cap2;
		const bool drops = s.anydrop != 0;
		for(uint32_t l = threadIdx.x; l < n; l += SPLIT_THREADS)
		{
			const R rec = s.sorted[l];
			const uint32_t bin = (uint32_t)RecOps<R>::key(rec) & sub_mask;
			if(!drops || !((s.dropmask[bin >> 5] >> (bin & 31u)) & 1u)) dst[s.gbase[bin] + l] = rec;
		}
		if(++stage == STAGES) { stage = 0; phase ^= 1u; }
		__syncthreads();
	}
}

constexpr uint32_t GROUP_WARPS = GROUP_THREADS / 32;
constexpr uint32_t GROUP_DEFER_CAP = (GROUP_CAP + GROUP_THREADS - 1) / GROUP_THREADS * 32;   // records one warp handles per bucket

template<class R> struct GroupSmem {
	R stage[GROUP_STAGES][GROUP_CAP];
	uint32_t tab[GROUP_SLOTS];                         // {tag : 20, index of the class's first record : 12} or EMPTY32
	uint32_t pay[GROUP_CAP];                           // payload of the class whose first record has this index (else 0)
	uint16_t defer[GROUP_WARPS][GROUP_DEFER_CAP];      // per warp: records whose home slot holds another key
	uint16_t lut[128];                                 // payload_bits of every 7-bit context
	unsigned long long bar[GROUP_STAGES];
	uint32_t n_stage[GROUP_STAGES];
};

// Level 3: one bucket at a time per CTA (persistent grid, GROUP_STAGES buckets in flight per CTA through the TMA ring).
// nkeys counts every bifurcation class even when the key list is full (the caller then regrows it and runs again).
// PKEY (fingerprint records): a class is {56-bit fingerprint, level-1 partition} -- the partition was chosen by a second
// hash that the record does not carry -- and the key list takes 16-byte entries {fingerprint, partition}.
template<class R, bool PKEY> struct GroupKey { typedef R type; };
template<> struct GroupKey<uint64_t, true> { typedef Rec16 type; };

// part0 = global index of the first partition grouped here (sharded runs: the first owned partition)
template<class R, bool PKEY = false>
__global__ void __launch_bounds__(GROUP_THREADS, sizeof(R) == 8 ? 4 : 2) k_group(const R *__restrict__ recs2, const uint32_t *__restrict__ cnt2,
	uint32_t nbuckets, uint32_t sub_bits, uint32_t part0, uint32_t cap2, const uint32_t *__restrict__ overflow,
	typename GroupKey<R, PKEY>::type *__restrict__ ckeys, uint32_t ckeys_cap, uint32_t *__restrict__ nkeys)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	GroupSmem<R> &s = *reinterpret_cast<GroupSmem<R>*>(smem_raw);
	if(*overflow) return;                                  // a bucket outgrew its region: the caller takes the L2-table path
	for(uint32_t i = threadIdx.x; i < GROUP_SLOTS; i += GROUP_THREADS) s.tab[i] = EMPTY32;
	for(uint32_t i = threadIdx.x; i < GROUP_CAP; i += GROUP_THREADS) s.pay[i] = 0u;
	if(threadIdx.x < 128) s.lut[threadIdx.x] = (uint16_t)payload_bits(threadIdx.x);
	if(threadIdx.x == 0)
	{
		for(uint32_t st = 0; st < GROUP_STAGES; st++) mbar_init(&s.bar[st], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	auto issue = [&](uint32_t q, uint32_t n, uint32_t st) {
		s.n_stage[st] = n;
		bulk_load(s.stage[st], recs2 + (uint64_t)q * cap2, (n * (uint32_t)sizeof(R) + 15u) & ~15u, &s.bar[st]);
	};
	uint32_t n_ahead = 0;                                  // thread 0: fill count of the bucket it will issue next
	if(threadIdx.x == 0)
	{
		uint32_t q = blockIdx.x;
		for(uint32_t st = 0; st < GROUP_STAGES && q < nbuckets; st++, q += gridDim.x) issue(q, __ldg(cnt2 + q), st);
		if(q < nbuckets) n_ahead = __ldg(cnt2 + q);
	}
	uint32_t stage = 0, phase = 0;
	const uint32_t lane = threadIdx.x & 31u;
	for(uint32_t q = blockIdx.x; q < nbuckets; q += gridDim.x)
	{
		mbar_wait(&s.bar[stage], phase);
		const uint32_t n = s.n_stage[stage];
		const R *w = s.stage[stage];
		// the bucket after the next: its count is fetched now and used when this stage is refilled below
		const uint64_t qn = (uint64_t)q + (uint64_t)GROUP_STAGES * gridDim.x;
		const uint32_t n_next = n_ahead;
		if(threadIdx.x == 0 && qn + gridDim.x < nbuckets) n_ahead = __ldg(cnt2 + qn + gridDim.x);

		// Fast pass: one CAS on the home slot decides most records (first of its class, or the class is there already).
		// A record whose home slot holds another key is set aside in the warp's list; the list is then probed with all
		// lanes busy -- a probing loop inside the fast pass would run at the pace of the warp's unluckiest lane.
		const uint32_t n_round = (n + 31u) & ~31u;
		const uint32_t warp = threadIdx.x >> 5, lt_mask = (1u << lane) - 1u;
		uint32_t ndef = 0;
		for(uint32_t i = threadIdx.x; i < n_round; i += GROUP_THREADS)
		{
			bool later = false;
			if(i < n)
			{
				const R rec = w[i];
				const unsigned long long m = RecOps<R>::key(rec);
				const uint32_t slot = ((uint32_t)m >> SUB_BITS_MAX) & (GROUP_SLOTS - 1u);
				const uint32_t entry = ((uint32_t)(m >> 22) << 12) | i;
				uint32_t bits = s.lut[RecOps<R>::ctx(rec)], target = i;
				const uint32_t old = atomicCAS(&s.tab[slot], EMPTY32, entry);
				if(old != EMPTY32)
				{
					const uint32_t j = old & 4095u;
					if((old ^ entry) < 4096u && RecOps<R>::key(w[j]) == m)   // same tag, same key
					{
						target = j;
						bits |= PAY_MULTI;
					}
					else later = true;
				}
				if(!later) atomicOr(&s.pay[target], bits);
			}
			const uint32_t dm = __ballot_sync(0xffffffffu, later);
			if(later) s.defer[warp][ndef + __popc(dm & lt_mask)] = (uint16_t)i;
			ndef += __popc(dm);
		}
		__syncwarp();
		for(uint32_t d = lane; d < ndef; d += 32)
		{
			const uint32_t i = s.defer[warp][d];
			const R rec = w[i];
			const unsigned long long m = RecOps<R>::key(rec);
			uint32_t slot = ((uint32_t)m >> SUB_BITS_MAX) & (GROUP_SLOTS - 1u);
			const uint32_t entry = ((uint32_t)(m >> 22) << 12) | i;
			uint32_t bits = s.lut[RecOps<R>::ctx(rec)], target = i;
			for(;;)
			{
				slot = (slot + 1u) & (GROUP_SLOTS - 1u);
				const uint32_t old = atomicCAS(&s.tab[slot], EMPTY32, entry);
				if(old == EMPTY32) break;                      // first record of its class
				if((old ^ entry) < 4096u)                      // same tag: compare the keys
				{
					const uint32_t j = old & 4095u;
					if(RecOps<R>::key(w[j]) == m)
					{
						target = j;
						bits |= PAY_MULTI;
						break;
					}
				}
			}
			atomicOr(&s.pay[target], bits);
		}
		__syncthreads();
		// table reset for the next bucket (16 KB: cheaper than each class finding its slot again)
		{
			uint4 *t4 = reinterpret_cast<uint4*>(s.tab);
			for(uint32_t i = threadIdx.x; i < GROUP_SLOTS / 4; i += GROUP_THREADS) t4[i] = make_uint4(EMPTY32, EMPTY32, EMPTY32, EMPTY32);
		}
		for(uint32_t i = threadIdx.x; i < n_round; i += GROUP_THREADS)
		{
			bool bif = false;
			if(i < n)
			{
				const uint32_t pay = s.pay[i];                     // non-zero exactly for the first record of a class
				if(pay)
				{
					bif = is_bifurcation(pay);
					s.pay[i] = 0u;
				}
			}
			const uint32_t mk = __ballot_sync(0xffffffffu, bif);
			if(mk)
			{
				uint32_t base = 0;
				if(lane == (uint32_t)(__ffs(mk) - 1)) base = atomicAdd(nkeys, (uint32_t)__popc(mk));
				base = __shfl_sync(0xffffffffu, base, __ffs(mk) - 1);
				const uint32_t idx = base + __popc(mk & ((1u << lane) - 1u));
				if(bif && idx < ckeys_cap)
				{
					if constexpr(PKEY) ckeys[idx] = Rec16{unmix56(RecOps<R>::key(w[i])), (uint64_t)(part0 + (q >> sub_bits))};
					else if constexpr(sizeof(R) == 8) ckeys[idx] = unmix56(RecOps<R>::key(w[i]));
					else ckeys[idx] = make_ulonglong2(unmix64(RecOps<R>::key(w[i])), 0ull);
				}
			}
		}
		__syncthreads();                                       // stage consumed, table and payloads clean
		if(threadIdx.x == 0 && qn < nbuckets) issue((uint32_t)qn, n_next, stage);
		if(++stage == GROUP_STAGES) { stage = 0; phase ^= 1u; }
	}
}

// vertex keys of the L2-table fallback on mixed records -> plain canonical keys
template<class R>
__global__ void __launch_bounds__(256) k_unmix(R *__restrict__ keys, uint64_t n)
{
	const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n) return;
	if constexpr(sizeof(R) == 8) keys[i] = unmix56(keys[i]);
	else keys[i].x = unmix64(keys[i].x);
}

} // namespace sibgpu
