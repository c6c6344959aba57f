// sibgpu_simplify -- one stage of BlockFinder::PerformGraphSimplifications (/root/reference/src/blockfinder.cpp:78-98).
//
// The reference calls RemoveBulges(id) (bulgeremoval.cpp:330-430) for every vertex id in order, up to maxIterations
// sweeps; 94 % of those calls only run the detection walks (AnyBulges, :158-218) and return, the rest edit the
// sequence, and every edit is visible to all later calls.  Here
//   * the detection walks of ALL vertices of a sweep run on the GPU against a snapshot of the sweep's initial state
//     (k_bulge_detect: one warp per vertex, lanes stride the walk of each instance over the flat per-position vertex
//     marks, reached vertices meet in a per-warp shared-memory table keyed by vertex id);
//   * the host then visits the vertex ids in the reference's order and runs the exact RemoveBulges logic only for the
//     vertices that the GPU flagged, or whose inputs were touched by an earlier collapse of the same sweep ("dirty":
//     an instance added/erased, or any element within reach of one of its walks modified).  For every other vertex
//     AnyBulges is a pure function of state the snapshot still describes, so its (negative) result is already known.
// The host part restates the reference's data structures on flat arrays: elements are array indices linked in
// sequence order (unrolled_list), per-strand vertex marks live on the elements (BifurcationStorage::posBifurcation_ +
// the two info bits), per-vertex instance lists keep the reference's slist semantics (push-front, lazy erase,
// Cleanup), and the bulge groups of a vertex are visited in Boost 1.54 unordered_map order (boost_order.h).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <string>
#include <thread>

#include <cub/cub.cuh>

#include "context.h"
#include "simplifier.h"

namespace sibgpu {

double since_load_ms();                                // api.cu
using namespace simp;

// ---------------------------------------------------------------------------------------------------------------
// K8: bulge detection for every vertex against the sweep snapshot.
// flag[v] = 1 iff two instances of v with different, non-empty end characters reach a common vertex within their
// detection walks (steps 1 .. D-1 along the strand, cut at the first return to v or at a chromosome end) -- which is
// AnyBulges(v) (bulgeremoval.cpp:158-218) up to false positives (the shared-memory table merges vertex ids that
// collide, and an instance whose k+1 window crosses a separator is not excluded); the host re-checks every flagged
// vertex with the exact logic, so false positives only cost time.
// ---------------------------------------------------------------------------------------------------------------
constexpr int DETECT_WARPS = 8;
constexpr int DETECT_SLOTS = 2048;

__global__ void __launch_bounds__(DETECT_WARPS * 32) k_bulge_detect(const uint8_t *__restrict__ ch,
	const uint32_t *__restrict__ mark0, const uint32_t *__restrict__ mark1, uint32_t total,
	const uint64_t *__restrict__ inst_off, const uint32_t *__restrict__ inst_elem, uint32_t nvert, uint32_t k, uint32_t D,
	uint8_t *__restrict__ flag)
{
	__shared__ uint8_t table[DETECT_WARPS][DETECT_SLOTS];
	const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
	uint8_t *tab = table[warp];
	for(uint32_t v = blockIdx.x * DETECT_WARPS + warp; v < nvert; v += gridDim.x * DETECT_WARPS)
	{
		const uint64_t b = inst_off[v], e = inst_off[v + 1];
		if(e - b < 2)
		{
			if(lane == 0) flag[v] = 0;
			continue;
		}
		for(uint32_t i = lane; i < DETECT_SLOTS / 4; i += 32) reinterpret_cast<uint32_t*>(tab)[i] = 0u;
		__syncwarp();
		bool conflict = false;
		for(uint64_t ii = b; ii < e && !conflict; ii++)
		{
			const uint32_t packed = inst_elem[ii];
			const uint32_t strand = packed >> 31, x = packed & 0x7FFFFFFFu;
			// end character: the element k steps ahead in strand direction
			const int64_t ek = strand ? (int64_t)x - k : (int64_t)x + k;
			uint32_t ec = 255;
			if(ek >= 0 && ek < (int64_t)total)
			{
				const uint8_t c = ch[ek];
				if(c != '$')
				{
					const uint32_t code = ((c >> 1) ^ (c >> 2)) & 3u;     // A C G T -> 0 1 2 3
					ec = strand ? 3u - code : code;
				}
			}
			if(ec == 255) continue;
			const uint8_t bit = (uint8_t)(1u << ec);
			const uint32_t *mk = strand ? mark1 : mark0;
			for(uint32_t base = 1; base < D; base += 32)
			{
				const uint32_t step = base + lane;
				const int64_t p = strand ? (int64_t)x - step : (int64_t)x + step;
				uint32_t m = 0xFFFFFFFFu;
				bool stop = true;
				if(step < D && p >= 0 && p < (int64_t)total)
				{
					m = mk[p];
					stop = ch[p] == '$' || m == v;
				}
				const uint32_t stops = __ballot_sync(0xffffffffu, stop);
				const bool active = stops == 0 || lane < (uint32_t)(__ffs(stops) - 1);
				if(active && m != 0xFFFFFFFFu)
				{
					const uint32_t slot = (m * 2654435761u) >> 21;            // 2048 slots
					const uint32_t word = slot >> 2, sh = (slot & 3u) * 8;
					const uint32_t old = atomicOr(reinterpret_cast<uint32_t*>(tab) + word, (uint32_t)bit << sh);
					if(((old >> sh) & 0xFFu) & ~bit) conflict = true;
				}
				conflict = __any_sync(0xffffffffu, conflict);
				if(stops || conflict) break;
			}
		}
		if(lane == 0) flag[v] = conflict ? 1 : 0;
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Device-resident stage front-end: everything k_bulge_detect needs is derived on the device from what the enumeration
// left there -- the text IS the element array of a fresh DNASequence (element index = text position,
// dnasequence.cpp:75-103), the two instance tables give the per-strand marks (IndexedSequence::Init,
// indexedsequence.cpp:51-67) and, sorted by vertex id, the per-vertex instance lists.  No host state, no upload.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_marks(const sibgpu_inst *__restrict__ pos, const sibgpu_inst *__restrict__ neg, uint64_t n,
	const uint32_t *__restrict__ chr_start, const uint32_t *__restrict__ chr_len, uint32_t *__restrict__ mark0,
	uint32_t *__restrict__ mark1, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
	{
		const sibgpu_inst p = pos[i], q = neg[i];
		const uint32_t e0 = chr_start[p.chr] + p.pos;                               // element where the k-mer starts, + strand
		const uint32_t e1 = chr_start[q.chr] + chr_len[q.chr] - 1u - q.pos;         // - strand: counted from the chromosome's end
		mark0[e0] = p.bifId;
		mark1[e1] = q.bifId;
		keys[i] = p.bifId;
		vals[i] = e0;
		keys[n + i] = q.bifId;
		vals[n + i] = e1 | 0x80000000u;
	}
}

// inst_off[v] = first index of vertex v in the id-sorted instance list, for v in 0 .. nvert (nvert + 1 entries)
__global__ void __launch_bounds__(256) k_inst_offsets(const uint32_t *__restrict__ sorted_keys, uint64_t n, uint32_t nvert,
	uint64_t *__restrict__ inst_off)
{
	const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
	if(v > nvert) return;
	uint64_t lo = 0, hi = n;
	while(lo < hi)
	{
		const uint64_t mid = (lo + hi) >> 1;
		if(sorted_keys[mid] < v) lo = mid + 1; else hi = mid;
	}
	inst_off[v] = lo;
}

__global__ void __launch_bounds__(256) k_count_flags(const uint8_t *__restrict__ flag, uint32_t n, unsigned long long *__restrict__ out)
{
	uint32_t c = 0;
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c += flag[i];
	c = __reduce_add_sync(0xffffffffu, c);
	if((threadIdx.x & 31u) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// Detection of every vertex of the stage's first sweep, from the enumeration result resident on `ctx`.
// *nflag = number of flagged vertices; the flags stay in ctx->d_s_flag (max_id + 1 bytes).
static int detect_resident(sibgpu_ctx *ctx, uint32_t k, uint32_t D, uint32_t max_id, uint64_t *nflag, float *ms)
{
	NvtxRange nvtx("sibgpu: stage front-end (marks, instance lists, k_bulge_detect)");
	cudaStream_t st = ctx->stream;
	const uint64_t total = ctx->M, n = ctx->n_inst;
	const uint32_t nvert = max_id + 1;
	DevBuf &d_m0 = ctx->d_s_m0, &d_m1 = ctx->d_s_m1, &d_off = ctx->d_s_off, &d_inst = ctx->d_s_inst, &d_flag = ctx->d_s_flag;
	SIB_TRY(d_m0.ensure(total * 4));
	SIB_TRY(d_m1.ensure(total * 4));
	SIB_TRY(d_off.ensure(((size_t)nvert + 2) * 8));
	SIB_TRY(d_inst.ensure(n * 8 + 8));
	SIB_TRY(d_flag.ensure((size_t)nvert + 8));
	SIB_TRY(ctx->d_vkeys.ensure(n * 16 + 16));             // id-sort workspace: keys | keys_alt
	SIB_TRY(ctx->d_vkeys_alt.ensure(n * 8 + 8));           //                    vals_alt
	uint32_t *keys = ctx->d_vkeys.as<uint32_t>(), *keys_alt = keys + n * 2;
	uint32_t *vals = d_inst.as<uint32_t>(), *vals_alt = ctx->d_vkeys_alt.as<uint32_t>();
	SIB_CUDA(cudaEventRecord(ctx->ev_begin, st));
	SIB_CUDA(cudaMemsetAsync(d_m0.p, 0xFF, total * 4, st));
	SIB_CUDA(cudaMemsetAsync(d_m1.p, 0xFF, total * 4, st));
	SIB_CUDA(cudaMemsetAsync(ctx->d_scalars.as<uint64_t>() + 16, 0, 8, st));
	const uint32_t grid = (uint32_t)ctx->sm_count * 8;
	if(n)
	{
		k_fill_marks<<<grid, 256, 0, st>>>(ctx->d_pos.as<sibgpu_inst>(), ctx->d_neg.as<sibgpu_inst>(), n, ctx->d_chr_start.as<uint32_t>(),
			ctx->d_chr_len.as<uint32_t>(), d_m0.as<uint32_t>(), d_m1.as<uint32_t>(), keys, vals);
		int bits = 1;
		while(bits < 32 && (1ull << bits) <= max_id) bits++;
		cub::DoubleBuffer<uint32_t> kb(keys, keys_alt), vb(vals, vals_alt);
		size_t tmp_bytes = 0;
		SIB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int)(2 * n), 0, bits, st));
		SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
		SIB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_cubtmp.p, tmp_bytes, kb, vb, (int)(2 * n), 0, bits, st));
		k_inst_offsets<<<(nvert + 1 + 255) / 256, 256, 0, st>>>(kb.Current(), 2 * n, nvert, d_off.as<uint64_t>());
		uint32_t dgrid = (nvert + DETECT_WARPS - 1) / DETECT_WARPS;
		if(dgrid > grid) dgrid = grid;
		k_bulge_detect<<<dgrid, DETECT_WARPS * 32, 0, st>>>(ctx->d_text.as<uint8_t>(), d_m0.as<uint32_t>(), d_m1.as<uint32_t>(),
			(uint32_t)total, d_off.as<uint64_t>(), vb.Current(), nvert, k, D, d_flag.as<uint8_t>());
		k_count_flags<<<grid, 256, 0, st>>>(d_flag.as<uint8_t>(), nvert, ctx->d_scalars.as<unsigned long long>() + 16);
		ctx->total_launches += 12;
	}
	SIB_CUDA(cudaEventRecord(ctx->ev_end, st));
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	SIB_CUDA(cudaMemcpyAsync(hs + 16, ctx->d_scalars.as<uint64_t>() + 16, 8, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	SIB_CUDA(cudaEventElapsedTime(ms, ctx->ev_begin, ctx->ev_end));
	*nflag = n ? hs[16] : 0;
	return SIBGPU_OK;
}

} // namespace sibgpu

using namespace sibgpu;

extern "C" int sibgpu_simplify(sibgpu_ctx *ctx, char **seq, uint32_t **origpos, uint64_t *len, uint32_t nchr,
	uint32_t k, uint32_t min_branch_size, uint32_t max_iterations, sibgpu_progress_fn progress, void *user, uint64_t *bulges)
{
	if(!ctx || !bulges || (nchr && (!seq || !origpos || !len)) || k == 0)
	{
		set_error("invalid: NULL argument or k == 0");
		return SIBGPU_ERR_INVALID;
	}
	NvtxRange nvtx_stage("sibgpu: simplify stage");
	const bool trace = getenv("SIBGPU_TRACE") != nullptr;
	auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t_mark = now();
	auto lap = [&](const char *what) {
		if(trace)
		{
			const double t = now();
			fprintf(stderr, "[sibgpu_simplify] t=%.0f ms  %-28s %8.2f ms\n", since_load_ms(), what, (t - t_mark) * 1e3);
			t_mark = t;
		}
	};
	// ---- IndexedSequence(rawSeq_, originalPos_, k, tempDir_, true): the vertex tables come from the GPU enumerator and
	// stay in HBM; the detection of the first sweep runs on them right away (no host state yet)
	SIB_TRY(enumerate_keep(ctx, seq, len, nchr, k));
	const uint32_t count = ctx->n_vertices;
	uint64_t launches = ctx->total_launches;
	float device_ms = ctx->last_ms;
	lap("enumerate (host buffers)");
	const size_t max_id = count;
	uint64_t n_flag_dev = 0;
	{
		float ms = 0.f;
		SIB_TRY(detect_resident(ctx, k, min_branch_size, count, &n_flag_dev, &ms));
		device_ms += ms;
		launches += ctx->total_launches - launches;
	}
	lap("marks + instance lists + k_bulge_detect");
	const size_t PROGRESS_STRIDE = 50;
	if(n_flag_dev == 0)
	{
		// No vertex has a bulge: the first sweep of SimplifyGraph (blockfinder.cpp:29-43) calls RemoveBulges for every id,
		// all return 0, the loop ends, and the copy-back re-spells the unchanged sequence.  Nothing to build, nothing to
		// copy: the caller's arrays stay as they are (only the progress protocol is replayed).
		if(progress)
		{
			progress(0, 0, user);
			const size_t threshold = (max_id * max_iterations) / PROGRESS_STRIDE;
			size_t cnt = 0, total_progress = 0;
			for(size_t id = 0; id <= max_id; id++)
			{
				if(++cnt >= threshold)
				{
					cnt = 0;
					total_progress = std::min(total_progress + 1, PROGRESS_STRIDE);
					progress(total_progress, 1, user);
				}
			}
			progress(PROGRESS_STRIDE, 2, user);
		}
		if(trace) fprintf(stderr, "[sibgpu_simplify] sweep 1: %zu vertices, 0 flagged by the GPU: stage leaves the sequences untouched\n", max_id + 1);
		ctx->total_launches = launches;
		ctx->last_ms = device_ms;
		*bulges = 0;
		return SIBGPU_OK;
	}
	sibgpu_inst *pos = nullptr, *neg = nullptr;
	uint64_t npos = 0, nneg = 0;
	SIB_TRY(sibgpu_download(ctx, &pos, &npos, &neg, &nneg));
	std::vector<uint8_t> flag(max_id + 1);
	SIB_CUDA(cudaMemcpyAsync(flag.data(), ctx->d_s_flag.p, max_id + 1, cudaMemcpyDeviceToHost, ctx->stream));
	SIB_CUDA(cudaStreamSynchronize(ctx->stream));
	lap("download tables + flags");
	Simplifier S;
	auto recycle = [&]() {                              // swaps the big per-element arrays with the context's pool
		S.ch.swap(ctx->pool_ch);
		S.opos.swap(ctx->pool_u32[0]);
		S.mark[0].swap(ctx->pool_u32[1]);
		S.mark[1].swap(ctx->pool_u32[2]);
		S.nxt.swap(ctx->pool_i32[0]);
		S.prv.swap(ctx->pool_i32[1]);
		S.node_of[0].swap(ctx->pool_i32[2]);
		S.node_of[1].swap(ctx->pool_i32[3]);
	};
	recycle();
	{
		NvtxRange nvtx("sibgpu: host state build");
		S.build(nchr, seq, origpos, len, k, min_branch_size, count, pos, npos, neg, nneg);
	}
	sibgpu_free(pos);
	sibgpu_free(neg);
	S.slot_of.assign((size_t)count + 1, -1);
	if(const char *e = getenv("SIBGPU_AHEAD_HELPERS")) S.ahead_helpers = (size_t)atoi(e);   // dev: 0 = no run-ahead prefetchers
	if(const char *e = getenv("SIBGPU_AHEAD_LEAD")) S.ahead_lead = (size_t)std::max(1, atoi(e));
	lap("build host state");

	// ---- SimplifyGraph, blockfinder.cpp:16-51
	size_t cnt = 0, total_bulges = 0, iterations = 0, total_progress = 0;
	if(progress) progress(total_progress, 0, user);
	const size_t threshold = (max_id * max_iterations) / PROGRESS_STRIDE;
	do
	{
		iterations++;
		NvtxRange nvtx_sweep("sibgpu: sweep (parallel screen + ordered commit)");
		// ---- first sweep: the GPU decided every vertex against the initial state (flags above).  Later sweeps need no
		// snapshot of everything: `dirty` is cleared when a vertex is visited and set by every later change its walks can
		// see, so a vertex that is clean at its next visit would repeat its last (empty) outcome -- RemoveBulges only
		// changes state through collapses, and a call that collapsed something always dirties its own vertex.
		// The ids are visited in the reference's order, in chunks: before a chunk all host threads screen its flagged and
		// dirty vertices against the current state (read-only existence test of AnyBulges), so the strictly ordered part
		// only pays for vertices that really have a bulge or are dirtied after the screen.
		if(iterations > 1) std::fill(flag.begin(), flag.end(), 0);
		size_t n_flag = 0, n_calls = 0, n_screened = 0;
		const size_t collapses_before = S.collapses;
		const size_t CHUNK = (size_t)1 << 16;
		for(size_t chunk_lo = 0; chunk_lo <= max_id; chunk_lo += CHUNK)
		{
			const size_t chunk_hi = std::min(max_id + 1, chunk_lo + CHUNK);
			for(size_t id = chunk_lo; id < chunk_hi; id++) n_flag += flag[id];
			const size_t screened_here = S.screen_range(chunk_lo, chunk_hi, flag.data());
			n_screened += screened_here;
			if(screened_here) S.ahead_start();              // helper threads prefetch the loci of the vertices ahead
			size_t survivor = 0;                            // cursor into the screen's survivors (ascending ids)
			// ---- the reference's sweep, skipping the vertices whose negative outcome is already known
			for(size_t id = chunk_lo; id < chunk_hi; id++)
			{
				if(flag[id] || S.dirty[id])
				{
					S.ahead_pos.store((uint32_t)id, std::memory_order_relaxed);
					n_calls++;
					S.dirty[id] = 0;                        // clean as of this visit; the call itself may dirty it again
					// a vertex the screen has just seen a bulge for: its exact call skips the early-out existence pass
					bool expect = false;
					if(screened_here)
					{
						while(survivor < S.ahead_id.size() && S.ahead_id[survivor] < id) survivor++;
						expect = survivor < S.ahead_id.size() && S.ahead_id[survivor] == id;
					}
					total_bulges += S.remove_bulges(id, expect);
				}
				if(++cnt >= threshold && progress)
				{
					cnt = 0;
					total_progress = std::min(total_progress + 1, PROGRESS_STRIDE);
					progress(total_progress, 1, user);
				}
			}
			S.ahead_stop();
		}
		if(trace)
		{
			fprintf(stderr, "[sibgpu_simplify] sweep %zu: %zu vertices, %zu flagged by the GPU, %zu screened by the host threads, "
				"%zu exact calls, %zu collapses\n", iterations, max_id + 1, n_flag, n_screened, n_calls, S.collapses - collapses_before);
		}
		lap("ordered host commit");
#ifdef SIBGPU_COMMIT_PROF
		if(trace)
		{
			static const char *names[10] = {"remove_bulges (whole)", "any_bulges", "fill_visit", "overlap", "max_multiplicity",
				"erase_bifurcations", "replace", "update_bifurcations", "mark_dirty_around", "cleanup"};
			for(int i = 0; i < 10; i++) fprintf(stderr, "[commit prof] %-24s %10.1f Mcycles\n", names[i], commit_prof().t[i] / 1e6);
		}
#endif
		if(S.collapses == collapses_before)
		{
			// Nothing changed in this sweep, so every further sweep of the reference (it keeps sweeping while the
			// CUMULATIVE count is positive, blockfinder.cpp:29-43) finds the same state and changes nothing: only its
			// progress ticks remain.
			while(total_bulges > 0 && iterations < max_iterations)
			{
				iterations++;
				for(size_t id = 0; id <= max_id; id++)
				{
					if(++cnt >= threshold && progress)
					{
						cnt = 0;
						total_progress = std::min(total_progress + 1, PROGRESS_STRIDE);
						progress(total_progress, 1, user);
					}
				}
			}
			break;
		}
	}
	while(total_bulges > 0 && iterations < max_iterations);
	if(progress) progress(PROGRESS_STRIDE, 2, user);
	ctx->total_launches = launches;
	ctx->last_ms = device_ms;

	// ---- copy-back, blockfinder.cpp:85-95 (one host thread per chromosome: the walks are independent).  The caller's
	// arrays are only overwritten once every chromosome has its new buffers; on failure nothing is handed out.
	{
		NvtxRange nvtx("sibgpu: copy-back");
		std::vector<char*> new_seq(nchr, nullptr);
		std::vector<uint32_t*> new_pos(nchr, nullptr);
		std::vector<uint64_t> new_len(nchr, 0);
		auto copy_chr = [&](uint32_t c) {
			size_t n = 0;
			for(int32_t e = S.nxt[S.chr_first_sep[c]]; S.ch[e] != SEP; e = S.nxt[e]) n++;
			char *out_seq = static_cast<char*>(malloc(n + 1));
			uint32_t *out_pos = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (n + 1)));
			if(!out_seq || !out_pos)
			{
				free(out_seq);
				free(out_pos);
				return;
			}
			size_t j = 0;
			for(int32_t e = S.nxt[S.chr_first_sep[c]]; S.ch[e] != SEP; e = S.nxt[e], j++)
			{
				out_seq[j] = S.ch[e];
				out_pos[j] = S.opos[e];
			}
			new_seq[c] = out_seq;
			new_pos[c] = out_pos;
			new_len[c] = n;
		};
		std::vector<std::thread> th;
		for(uint32_t c = 0; c < nchr; c += 16)
		{
			th.clear();
			for(uint32_t d = c; d < nchr && d < c + 16; d++) th.emplace_back(copy_chr, d);
			for(std::thread &x : th) x.join();
		}
		bool ok = true;
		for(uint32_t c = 0; c < nchr; c++) ok = ok && new_seq[c] && new_pos[c];
		if(!ok)
		{
			for(uint32_t c = 0; c < nchr; c++)
			{
				free(new_seq[c]);
				free(new_pos[c]);
			}
			recycle();
			set_error("invalid: host allocation failed");
			return SIBGPU_ERR_INVALID;
		}
		for(uint32_t c = 0; c < nchr; c++)
		{
			seq[c] = new_seq[c];
			origpos[c] = new_pos[c];
			len[c] = new_len[c];
		}
	}
	lap("copy-back");
	recycle();
	*bulges = total_bulges;
	return SIBGPU_OK;
}

extern "C" void sibgpu_debug_unordered_order(const uint64_t *keys, uint64_t n, uint64_t *out)
{
	BoostUnorderedOrder o;
	for(uint64_t i = 0; i < n; i++) o.insert_new(keys[i], (int)i);
	std::vector<int> ord;
	o.order(std::back_inserter(ord));
	for(size_t i = 0; i < ord.size(); i++) out[i] = keys[ord[i]];
}
