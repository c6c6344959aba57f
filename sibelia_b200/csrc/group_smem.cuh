// Shared-memory grouping of the k-mer records (included by enumerate.cu; 8-byte records, k <= 28).
//
// The hash partitions written by k_scatter (level 1, ~512 Ki records each) are split once more by other bits of the same
// hash into buckets of ~1 Ki records (k_split: one coalesced read + one coalesced write of every record), and each
// bucket is then grouped entirely inside one SM (k_group): a TMA bulk copy (cp.async.bulk + mbarrier) brings the
// bucket's records into shared memory, every record claims the slot of its key in a shared-memory open-addressing
// table with a 64-bit shared CAS and ORs its predecessor/successor symbols into the slot's 32-bit payload, and the
// record that created a class evaluates the reference's predicate (vertexenumeration.cpp:67-70,330,348) on the final
// payload; bifurcation k-mers are appended to the global key list by a warp-ballot aggregated atomic.
// Shared atomics run at 2.6 (CAS.64) / 8.7 (OR.32) lane-operations per clock per SM on B200
// (tools/ubench/smem.cu, profiles/r2_smem_atomics.txt) against ~0.35 per clock per SM for CAS on an L2-resident table.
#pragma once

namespace sibgpu {

constexpr int SPLIT_THREADS = 512;
constexpr int SPLIT_TILE = 8192;                       // records per tile (64 KB of shared memory)
constexpr int SPLIT_PER_THREAD = SPLIT_TILE / SPLIT_THREADS;
constexpr int SPLIT_MAX_BINS = 1024;
constexpr uint32_t SPLIT_DROPPED = 0x80000000u;
constexpr uint32_t GROUP_THREADS = 256;
constexpr uint32_t GROUP_SLOTS = 4096;                 // shared-memory table slots (load <= 0.41)
constexpr uint32_t GROUP_MEAN = 1024;                  // target records per bucket
constexpr uint32_t GROUP_CAP = 1664;                   // fixed capacity of a bucket's region: mean + 1/2 + 128 (even: 16-byte aligned)
constexpr uint32_t GROUP_STAGES = 2;
constexpr uint32_t SUB_BITS_MAX = 10;                  // level-2 bucket = low bits of the hash, table slot = the next 12

struct SplitSmem {
	uint64_t rec[SPLIT_TILE];
	uint32_t cnt[SPLIT_MAX_BINS];                      // per-bin count of this tile, then exclusive local offset (| SPLIT_DROPPED)
	uint32_t gbase[SPLIT_MAX_BINS];                    // index of the bin's run in the partition's level-2 region minus the local offset
};

// Level 2: partition p's records [partbase[p], cursor[p]) -> B2 buckets of fixed capacity cap2 at out[(p * B2 + b) * cap2].
// A tile is 8192 consecutive records of one partition, counting-sorted by bucket in shared memory (rank = returning
// shared atomic) and copied out in coalesced runs; cnt2[p * B2 + b] is the bucket's fill count.  Consecutive tile
// indices belong to different partitions, so concurrently running CTAs bump different bucket counters.
__global__ void __launch_bounds__(SPLIT_THREADS, 2) k_split(const uint64_t *__restrict__ recs, const uint64_t *__restrict__ partbase,
	const unsigned long long *__restrict__ cursor, uint32_t P1, uint32_t tiles_per_part, uint32_t sub_bits,
	uint64_t *__restrict__ out, uint32_t *__restrict__ cnt2, uint32_t cap2, uint32_t *__restrict__ overflow)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SplitSmem &s = *reinterpret_cast<SplitSmem*>(smem_raw);
	typedef cub::BlockScan<uint32_t, SPLIT_THREADS> Scan;
	__shared__ typename Scan::TempStorage scan_tmp;
	const uint32_t B2 = 1u << sub_bits, sub_mask = B2 - 1u;
	const uint32_t ntiles = P1 * tiles_per_part;
	for(uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const uint32_t p = t % P1, chunk = t / P1;
		const uint64_t base = partbase[p];
		uint64_t np = cursor[(size_t)p * CURSOR_STRIDE] - base;
		const uint64_t room = partbase[p + 1] - base;
		if(np > room) np = room;                       // an overflowed level-1 partition: the caller discards this run anyway
		const uint64_t first = (uint64_t)chunk * SPLIT_TILE;
		if(first >= np) continue;                      // CTA-uniform
		const uint32_t n = np - first < SPLIT_TILE ? (uint32_t)(np - first) : SPLIT_TILE;
		for(uint32_t b = threadIdx.x; b < B2; b += SPLIT_THREADS) s.cnt[b] = 0;
		__syncthreads();

		const uint64_t *src = recs + base + first;
		uint64_t r[SPLIT_PER_THREAD];
		uint32_t binrank[SPLIT_PER_THREAD];
#pragma unroll
		for(int j = 0; j < SPLIT_PER_THREAD; j++)
		{
			const uint32_t i = j * SPLIT_THREADS + threadIdx.x;
			r[j] = i < n ? __ldcs(src + i) : 0ull;
		}
#pragma unroll
		for(int j = 0; j < SPLIT_PER_THREAD; j++)
		{
			const uint32_t i = j * SPLIT_THREADS + threadIdx.x;
			if(i < n)
			{
				const uint32_t bin = (uint32_t)rec_hash(r[j] >> 7, 0) & sub_mask;
				binrank[j] = (bin << 16) | atomicAdd(&s.cnt[bin], 1u);    // rank < 8192
			}
		}
		__syncthreads();

		// exclusive scan over the bins (2 per thread) + reservation of the runs in the buckets
		uint32_t c[2], sum = 0;
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
			if(b < B2)
			{
				uint32_t o = excl, g = 0;
				if(c[j])
				{
					g = atomicAdd(&cnt2[(size_t)p * B2 + b], c[j]);
					if(g + c[j] > cap2)
					{
						*overflow = 1u;
						o |= SPLIT_DROPPED;
					}
				}
				s.cnt[b] = o;
				s.gbase[b] = b * cap2 + g - excl;          // 32-bit wrap-around arithmetic: + local index >= excl
			}
			excl += c[j];
		}
		__syncthreads();

#pragma unroll
		for(int j = 0; j < SPLIT_PER_THREAD; j++)
		{
			const uint32_t i = j * SPLIT_THREADS + threadIdx.x;
			if(i < n) s.rec[(s.cnt[binrank[j] >> 16] & ~SPLIT_DROPPED) + (binrank[j] & 0xFFFFu)] = r[j];
		}
		__syncthreads();

		uint64_t *dst = out + (uint64_t)p * B2 * cap2;
		for(uint32_t l = threadIdx.x; l < n; l += SPLIT_THREADS)
		{
			const uint64_t rec = s.rec[l];
			const uint32_t bin = (uint32_t)rec_hash(rec >> 7, 0) & sub_mask;
			if(!(s.cnt[bin] & SPLIT_DROPPED)) dst[s.gbase[bin] + l] = rec;
		}
		__syncthreads();
	}
}

struct GroupSmem {
	unsigned long long keys[GROUP_SLOTS];
	unsigned long long stage[GROUP_STAGES][GROUP_CAP];
	uint32_t pay[GROUP_SLOTS];
	unsigned long long bar[GROUP_STAGES];
	uint32_t n_stage[GROUP_STAGES];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Level 3: one bucket at a time per CTA (persistent grid, GROUP_STAGES buckets in flight per CTA through the TMA ring).
// nkeys counts every bifurcation class even when the key list is full (the caller then regrows it and runs again).
__global__ void __launch_bounds__(GROUP_THREADS, 3) k_group(const uint64_t *__restrict__ recs2, const uint32_t *__restrict__ cnt2,
	uint32_t nbuckets, uint32_t cap2, const uint32_t *__restrict__ overflow, uint64_t *__restrict__ ckeys, uint32_t ckeys_cap,
	uint32_t *__restrict__ nkeys)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	GroupSmem &s = *reinterpret_cast<GroupSmem*>(smem_raw);
	if(*overflow) return;                                  // a bucket outgrew its region: the caller takes the L2-table path
	for(uint32_t i = threadIdx.x; i < GROUP_SLOTS; i += GROUP_THREADS)
	{
		s.keys[i] = EMPTY64;
		s.pay[i] = 0u;
	}
	if(threadIdx.x == 0)
	{
		for(uint32_t st = 0; st < GROUP_STAGES; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&s.bar[st])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// elected thread: bulk copy of bucket q (n records) into stage st; completes the stage's mbarrier
	auto issue = [&](uint32_t q, uint32_t n, uint32_t st) {
		const uint32_t bytes = (n * 8u + 15u) & ~15u;
		s.n_stage[st] = n;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s.bar[st])), "r"(bytes) : "memory");
		if(bytes)
		{
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(smem_u32(s.stage[st])), "l"(recs2 + (uint64_t)q * cap2), "r"(bytes), "r"(smem_u32(&s.bar[st])) : "memory");
		}
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
		uint32_t ok = 0;
		while(!ok)
		{
			asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
				: "=r"(ok) : "r"(smem_u32(&s.bar[stage])), "r"(phase) : "memory");
		}
		const uint32_t n = s.n_stage[stage];
		unsigned long long *w = s.stage[stage];
		uint32_t *w32 = reinterpret_cast<uint32_t*>(w);
		// the bucket after the next: its count is fetched now and used when this stage is refilled below
		const uint64_t qn = (uint64_t)q + (uint64_t)GROUP_STAGES * gridDim.x;
		uint32_t n_next = n_ahead;
		if(threadIdx.x == 0 && qn + gridDim.x < nbuckets) n_ahead = __ldg(cnt2 + qn + gridDim.x);

		for(uint32_t i = threadIdx.x; i < n; i += GROUP_THREADS)
		{
			const unsigned long long rec = w[i];
			const unsigned long long key = rec >> 7;
			uint32_t slot = ((uint32_t)rec_hash(key, 0) >> SUB_BITS_MAX) & (GROUP_SLOTS - 1u);
			uint32_t bits = payload_bits((uint32_t)rec & 127u), leader = 0;
			for(;;)
			{
				const unsigned long long old = atomicCAS(&s.keys[slot], EMPTY64, key);
				if(old == EMPTY64) { leader = 0x80000000u; break; }
				if(old == key) { bits |= PAY_MULTI; break; }
				slot = (slot + 1u) & (GROUP_SLOTS - 1u);
			}
			atomicOr(&s.pay[slot], bits);
			w32[2 * i] = slot | leader;                        // the record is consumed: its low word remembers the slot
		}
		__syncthreads();
		const uint32_t n_round = (n + 31u) & ~31u;
		for(uint32_t i = threadIdx.x; i < n_round; i += GROUP_THREADS)
		{
			bool bif = false;
			unsigned long long key = 0;
			if(i < n)
			{
				const uint32_t v = w32[2 * i];
				if(v & 0x80000000u)
				{
					const uint32_t slot = v & 0x7FFFFFFFu;
					key = s.keys[slot];
					bif = is_bifurcation(s.pay[slot]);
					s.keys[slot] = EMPTY64;                        // leave the table clean for the next bucket
					s.pay[slot] = 0u;
				}
			}
			const uint32_t m = __ballot_sync(0xffffffffu, bif);
			if(m)
			{
				uint32_t base = 0;
				if(lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(nkeys, (uint32_t)__popc(m));
				base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
				const uint32_t idx = base + __popc(m & ((1u << lane) - 1u));
				if(bif && idx < ckeys_cap) ckeys[idx] = key;
			}
		}
		// the slot notes above are generic-proxy writes into a buffer the bulk copy (async proxy) overwrites next
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();                                       // stage consumed, table clean
		if(threadIdx.x == 0 && qn < nbuckets) issue((uint32_t)qn, n_next, stage);
		if(++stage == GROUP_STAGES) { stage = 0; phase ^= 1u; }
	}
}

} // namespace sibgpu
