// extern "C" surface of libsibgpu (include/sibgpu.h): context, upload/download, enumerate.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "context.h"

namespace sibgpu {
static const std::chrono::steady_clock::time_point g_loaded = std::chrono::steady_clock::now();
double since_load_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_loaded).count(); }
static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
} // namespace sibgpu

using namespace sibgpu;

int sibgpu_ctx::ensure_aux_streams(uint32_t n)
{
	if(!ev_fork) SIB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
	for(uint32_t i = 0; i < n && i < 8; i++)
	{
		if(!aux_stream[i])
		{
			SIB_CUDA(cudaStreamCreateWithFlags(&aux_stream[i], cudaStreamNonBlocking));
			SIB_CUDA(cudaEventCreateWithFlags(&ev_join[i], cudaEventDisableTiming));
		}
	}
	return SIBGPU_OK;
}

int sibgpu_ctx::ensure_copy_stream(uint32_t nchunks)
{
	if(!copy_stream)
	{
		SIB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
		SIB_CUDA(cudaEventCreateWithFlags(&ev_fork_copy, cudaEventDisableTiming));
	}
	while(ev_chunk.size() < nchunks)
	{
		cudaEvent_t e;
		SIB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		ev_chunk.push_back(e);
	}
	return SIBGPU_OK;
}

int sibgpu_ctx::ensure_stage(size_t piece_bytes)
{
	if(h_stage && stage_piece >= piece_bytes) return SIBGPU_OK;
	if(h_stage) cudaFreeHost(h_stage);
	h_stage = nullptr;
	stage_piece = 0;
	SIB_CUDA(cudaMallocHost(&h_stage, piece_bytes * STAGE_SLOTS));
	stage_piece = piece_bytes;
	return SIBGPU_OK;
}

cudaEvent_t sibgpu_ctx::get_event()
{
	if(events_used == event_pool.size())
	{
		cudaEvent_t e;
		cudaEventCreate(&e);
		event_pool.push_back(e);
	}
	return event_pool[events_used++];
}

void sibgpu_ctx::prof_begin(const char *name, uint64_t bytes)
{
	Span s = {name, get_event(), get_event(), bytes};
	cudaEventRecord(s.a, stream);
	spans.push_back(s);
}

void sibgpu_ctx::prof_end() { cudaEventRecord(spans.back().b, stream); }

void sibgpu_ctx::prof_reset()
{
	spans.clear();
	events_used = 0;
	stats.clear();
}

int sibgpu_ctx::prof_collect()
{
	SIB_CUDA(cudaStreamSynchronize(stream));
	stats.clear();
	for(const Span &s : spans)
	{
		float ms = 0.f;
		SIB_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
		size_t i = 0;
		for(; i < stats.size() && strcmp(stats[i].name, s.name) != 0; i++);
		if(i == stats.size()) stats.push_back(Stat{s.name, 0u, 0.f, 0ull});
		stats[i].launches++;
		stats[i].ms += ms;
		stats[i].bytes += s.bytes;
	}
	return SIBGPU_OK;
}

extern "C" {

const char *sibgpu_last_error(void) { return g_error.c_str(); }
const char *sibgpu_version(void) { return "sibgpu 0.1 (sm_100a)"; }

int sibgpu_device_count(void)
{
	int n = 0;
	if(cudaGetDeviceCount(&n) != cudaSuccess)
	{
		cudaGetLastError();
		return 0;
	}
	return n;
}

int sibgpu_create(int device, sibgpu_ctx **out)
{
	if(!out)
	{
		set_error("invalid: out == NULL");
		return SIBGPU_ERR_INVALID;
	}
	*out = nullptr;
	int n = 0;
	SIB_CUDA(cudaGetDeviceCount(&n));
	if(device < 0 || device >= n)
	{
		set_error("cuda: device " + std::to_string(device) + " not available (" + std::to_string(n) + " visible)");
		return SIBGPU_ERR_CUDA;
	}
	SIB_CUDA(cudaSetDevice(device));
	sibgpu_ctx *c = new sibgpu_ctx();
	c->device = device;
	cudaDeviceProp prop;
	SIB_CUDA(cudaGetDeviceProperties(&prop, device));
	c->sm_count = prop.multiProcessorCount;
	SIB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	SIB_CUDA(cudaMallocHost(&c->h_scalars, 64 * sizeof(uint64_t)));
	SIB_CUDA(cudaEventCreate(&c->ev_begin));
	SIB_CUDA(cudaEventCreate(&c->ev_end));
	if(const char *e = getenv("SIBGPU_PART_RECORDS"))
	{
		uint64_t v = strtoull(e, nullptr, 10);
		if(v >= 1024) { c->part_target = v; c->part_explicit = true; }
	}
	if(const char *e = getenv("SIBGPU_PART_SLACK")) c->part_slack = strtoull(e, nullptr, 10);
	if(const char *e = getenv("SIBGPU_EXACT_HIST")) c->exact_hist = atoi(e) != 0;
	if(const char *e = getenv("SIBGPU_INSERT_VARIANT")) c->insert_variant = atoi(e);
	if(const char *e = getenv("SIBGPU_STREAMS")) c->n_streams = atoi(e);
	if(const char *e = getenv("SIBGPU_GROUP_SMEM")) c->group_smem = atoi(e);
	if(const char *e = getenv("SIBGPU_SPLIT_STAGES")) c->split_stages = atoi(e) == 1 ? 1 : 2;
	if(const char *e = getenv("SIBGPU_PIECEWISE")) c->piecewise_split = atoi(e) != 0;
	if(const char *e = getenv("SIBGPU_STAGE_THREADS")) c->stage_threads = atoi(e);
	if(const char *e = getenv("SIBGPU_CKEYS_INIT")) c->ckeys_init = strtoull(e, nullptr, 10);
	if(const char *e = getenv("SIBGPU_TABLE_FACTOR"))
	{
		int v = atoi(e);
		if(v >= 2 && v <= 16) c->table_factor = v;
	}
	*out = c;
	return SIBGPU_OK;
}

void sibgpu_destroy(sibgpu_ctx *c)
{
	if(!c) return;
	cudaSetDevice(c->device);
	DevBuf *bufs[] = {&c->d_text, &c->d_packed, &c->d_chr_start, &c->d_chr_len, &c->d_hist, &c->d_partoff, &c->d_cursor,
		&c->d_records, &c->d_table, &c->d_partcnt, &c->d_keyoff, &c->d_ckeys, &c->d_vkeys, &c->d_vkeys_alt, &c->d_cubtmp,
		&c->d_map, &c->d_filter, &c->d_hitmask, &c->d_tilecnt, &c->d_tileoff, &c->d_pos, &c->d_negtmp, &c->d_neg,
		&c->d_chrinst, &c->d_scalars, &c->d_fp, &c->d_fpprm, &c->d_rep, &c->d_order, &c->d_s_ch, &c->d_s_m0, &c->d_s_m1, &c->d_s_off,
		&c->d_s_inst, &c->d_s_flag, &c->d_edges, &c->d_edge_skip, &c->d_records2, &c->d_cnt2};
	for(void *pp : c->peer_ptr)
	{
		if(pp) cudaIpcCloseMemHandle(pp);
	}
	for(void *pp : c->peer_x)
	{
		if(pp) cudaIpcCloseMemHandle(pp);
	}
	c->d_keystage.release();
	c->d_xbuf.release();
	for(DevBuf *b : bufs) b->release();
	for(cudaEvent_t e : c->event_pool) cudaEventDestroy(e);
	if(c->h_scalars) cudaFreeHost(c->h_scalars);
	if(c->h_stage) cudaFreeHost(c->h_stage);
	c->d_sendbuf.release();
	for(void *q : c->send_retired_old) cudaFree(q);
	for(void *q : c->send_retired_new) cudaFree(q);
	if(c->ev_begin) cudaEventDestroy(c->ev_begin);
	if(c->ev_end) cudaEventDestroy(c->ev_end);
	for(int i = 0; i < 8; i++)
	{
		if(c->aux_stream[i]) cudaStreamDestroy(c->aux_stream[i]);
		if(c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
	}
	if(c->ev_fork) cudaEventDestroy(c->ev_fork);
	for(cudaEvent_t e : c->ev_chunk) cudaEventDestroy(e);
	if(c->ev_fork_copy) cudaEventDestroy(c->ev_fork_copy);
	if(c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if(c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

} // extern "C"

// Result buffers (instance tables) come from a small pool of pinned host blocks: the device-to-host copy of a table
// then runs at PCIe speed straight into the buffer the caller receives, and a caller that frees a table before asking
// for the next one (the reference binds one index at a time) gets the same pages back -- a fresh 50 MB malloc costs more
// in page faults than the whole enumeration.  sibgpu_free recognises pool blocks; everything else is free()d.
namespace {
struct PinnedBlock { void *p; size_t cap; bool used; };
std::mutex g_pool_mutex;
std::vector<PinnedBlock> g_pool;
const size_t POOL_MAX_BLOCKS = 8;
} // namespace

void *sibgpu::pool_alloc(size_t bytes)
{
	if(bytes < (1u << 16)) return malloc(bytes ? bytes : 1);           // small tables: plain memory
	std::lock_guard<std::mutex> lock(g_pool_mutex);
	size_t best = g_pool.size();
	for(size_t i = 0; i < g_pool.size(); i++)
	{
		if(!g_pool[i].used && g_pool[i].cap >= bytes && g_pool[i].cap <= 4 * bytes + (1u << 20) &&
			(best == g_pool.size() || g_pool[i].cap < g_pool[best].cap)) best = i;
	}
	if(best < g_pool.size())
	{
		g_pool[best].used = true;
		return g_pool[best].p;
	}
	if(g_pool.size() >= POOL_MAX_BLOCKS)
	{
		// drop the smallest idle block to make room; if all are in use fall back to plain memory
		size_t idle = g_pool.size();
		for(size_t i = 0; i < g_pool.size(); i++)
		{
			if(!g_pool[i].used && (idle == g_pool.size() || g_pool[i].cap < g_pool[idle].cap)) idle = i;
		}
		if(idle == g_pool.size()) return malloc(bytes);
		cudaFreeHost(g_pool[idle].p);
		g_pool.erase(g_pool.begin() + idle);
	}
	void *p = nullptr;
	const size_t cap = bytes + bytes / 4;
	if(cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess)
	{
		cudaGetLastError();
		return malloc(bytes);
	}
	g_pool.push_back(PinnedBlock{p, cap, true});
	return p;
}

namespace {
bool pool_release(void *p)
{
	std::lock_guard<std::mutex> lock(g_pool_mutex);
	for(PinnedBlock &b : g_pool)
	{
		if(b.p == p)
		{
			b.used = false;
			return true;
		}
	}
	return false;
}
} // namespace

extern "C" {

void sibgpu_free(void *p)
{
	if(p && !pool_release(p)) free(p);
}

// Plans the text layout '$' chr0 '$' chr1 '$' ... '$' (the DNASequence layout, src/dnasequence.cpp:75-103), allocates,
// fills the whole buffer with '$' and uploads the chromosome tables; the bases follow with copy_text_range.
static int upload_layout(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr)
{
	if(!c || (nchr && (!chr || !len)))
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	c->have_text = false;
	c->have_result = false;
	c->dist_result = false;
	c->dist_world = 1;
	uint64_t N = 0;
	for(uint32_t i = 0; i < nchr; i++) N += len[i];
	const uint64_t M = N + nchr + 1;
	// positions are 32-bit like the reference's Pos/Size (src/common.h:50-52: MAX_INPUT_SIZE = 1 << 30)
	if(M >= (1ull << 31))
	{
		set_error("invalid: input of " + std::to_string(N) + " bases exceeds the 32-bit position range");
		return SIBGPU_ERR_INVALID;
	}
	c->h_chr_start.resize(nchr);
	c->h_chr_len.resize(nchr);
	uint64_t at = 1;
	for(uint32_t i = 0; i < nchr; i++)
	{
		c->h_chr_start[i] = (uint32_t)at;
		c->h_chr_len[i] = (uint32_t)len[i];
		at += len[i] + 1;
	}
	const size_t nwords = (size_t)((M + 15) / 16) + 8;
	SIB_TRY(c->d_text.ensure(nwords * 16));
	SIB_TRY(c->d_chr_start.ensure(sizeof(uint32_t) * (nchr + 1)));
	SIB_TRY(c->d_chr_len.ensure(sizeof(uint32_t) * (nchr + 1)));
	SIB_CUDA(cudaMemsetAsync(c->d_text.p, '$', nwords * 16, c->stream));
	if(nchr)
	{
		SIB_CUDA(cudaMemcpyAsync(c->d_chr_start.p, c->h_chr_start.data(), sizeof(uint32_t) * nchr, cudaMemcpyHostToDevice, c->stream));
		SIB_CUDA(cudaMemcpyAsync(c->d_chr_len.p, c->h_chr_len.data(), sizeof(uint32_t) * nchr, cudaMemcpyHostToDevice, c->stream));
	}
	c->nchr = nchr;
	c->N = N;
	c->M = M;
	return SIBGPU_OK;
}

} // extern "C"

// sibgpu_enumerate without the download: host buffers in, result tables left in HBM (d_pos / d_neg, n_inst, n_vertices).
// Upload, pack and partition are pipelined piece by piece (enumerate.cu); the context ends up in the same state as
// after sibgpu_upload + sibgpu_enumerate_resident.
int sibgpu::enumerate_keep(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t k)
{
	if(!c || k == 0)
	{
		set_error("invalid: NULL context or k == 0");
		return SIBGPU_ERR_INVALID;
	}
	SIB_TRY(upload_layout(c, chr, len, nchr));
	HostSrc src = {chr, len};
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(enumerate_resident(c, k, &src));
	c->have_text = true;
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	return SIBGPU_OK;
}

extern "C" {

int sibgpu_upload(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr)
{
	SIB_TRY(upload_layout(c, chr, len, nchr));
	HostSrc src = {chr, len};
	SIB_TRY(copy_text_range(c, src, 0, c->M, c->stream));
	SIB_CUDA(cudaStreamSynchronize(c->stream));
	c->have_text = true;
	return SIBGPU_OK;
}

int sibgpu_enumerate_resident(sibgpu_ctx *c, uint32_t k, uint64_t *ninst, uint32_t *count)
{
	if(!c || k == 0)
	{
		set_error("invalid: NULL context or k == 0");
		return SIBGPU_ERR_INVALID;
	}
	if(!c->have_text || c->dist_world > 1)
	{
		// after sibgpu_dist_upload only this rank's byte range of the text is resident
		set_error("state: sibgpu_upload must precede sibgpu_enumerate_resident");
		return SIBGPU_ERR_STATE;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(enumerate_resident(c, k, nullptr));
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	if(ninst) *ninst = c->n_inst;
	if(count) *count = c->n_vertices;
	return SIBGPU_OK;
}

int sibgpu_download(sibgpu_ctx *c, sibgpu_inst **pos, uint64_t *npos, sibgpu_inst **neg, uint64_t *nneg)
{
	if(!c || !pos || !neg || !npos || !nneg)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	if(!c->have_result)
	{
		set_error("state: no enumeration result to download");
		return SIBGPU_ERR_STATE;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	const uint64_t n = c->n_inst;
	*pos = static_cast<sibgpu_inst*>(pool_alloc(sizeof(sibgpu_inst) * (n + 1)));
	*neg = static_cast<sibgpu_inst*>(pool_alloc(sizeof(sibgpu_inst) * (n + 1)));
	if(!*pos || !*neg)
	{
		sibgpu_free(*pos);
		sibgpu_free(*neg);
		*pos = *neg = nullptr;
		set_error("invalid: host allocation failed");
		return SIBGPU_ERR_INVALID;
	}
	if(n)
	{
		SIB_CUDA(cudaMemcpyAsync(*pos, c->d_pos.p, sizeof(sibgpu_inst) * n, cudaMemcpyDeviceToHost, c->stream));
		SIB_CUDA(cudaMemcpyAsync(*neg, c->dist_result ? c->d_negtmp.p : c->d_neg.p, sizeof(sibgpu_inst) * n, cudaMemcpyDeviceToHost, c->stream));
		SIB_CUDA(cudaStreamSynchronize(c->stream));
	}
	*npos = n;
	*nneg = n;
	return SIBGPU_OK;
}

int sibgpu_enumerate(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t k,
	sibgpu_inst **pos, uint64_t *npos, sibgpu_inst **neg, uint64_t *nneg, uint32_t *count)
{
	static const bool trace = getenv("SIBGPU_TRACE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	SIB_TRY(sibgpu::enumerate_keep(c, chr, len, nchr, k));
	if(count) *count = c->n_vertices;
	const int rc = sibgpu_download(c, pos, npos, neg, nneg);
	if(trace)
	{
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "[sibgpu_enumerate] t=%.0f ms  N=%llu nchr=%u k=%u -> V=%u I=%llu  %.3f ms (device %.3f ms, %llu launches)\n",
			since_load_ms(), (unsigned long long)c->N, nchr, k, c->n_vertices, (unsigned long long)c->n_inst, ms, c->last_ms,
			(unsigned long long)c->total_launches);
	}
	return rc;
}

int sibgpu_list_edges(sibgpu_ctx *c, const char *const *seq, const uint32_t *const *origpos, const uint64_t *len, uint32_t nchr,
	uint32_t k, sibgpu_edge **edges, uint64_t *nedges)
{
	if(!c || !edges || !nedges || k == 0)
	{
		set_error("invalid: NULL argument or k == 0");
		return SIBGPU_ERR_INVALID;
	}
	static const bool trace = getenv("SIBGPU_TRACE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	SIB_TRY(upload_layout(c, seq, len, nchr));
	HostSrc src = {seq, len};
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(enumerate_resident(c, k, &src));
	c->have_text = true;
	SIB_TRY(list_edges_device(c, k, edges, nedges));
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	// original coordinates (SpellOriginal, src/dnasequence.cpp:254-260): the kernel left the element indices of the
	// first and the last element of every edge; originalPos_ lives on the host
	sibgpu_edge *e = *edges;
	for(uint64_t i = 0; i < *nedges; i++)
	{
		uint32_t o1 = e[i].original_position, o2 = e[i].original_length;
		if(origpos)
		{
			o1 = origpos[e[i].chr][o1];
			o2 = origpos[e[i].chr][o2];
		}
		const uint32_t lo = o1 < o2 ? o1 : o2, hi = o1 < o2 ? o2 : o1;
		e[i].original_position = lo;
		e[i].original_length = hi + 1 - lo;
	}
	if(trace)
	{
		const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		fprintf(stderr, "[sibgpu_list_edges] t=%.0f ms  N=%llu nchr=%u k=%u -> V=%u I=%llu E=%llu  %.3f ms (device %.3f ms)\n",
			since_load_ms(), (unsigned long long)c->N, nchr, k, c->n_vertices, (unsigned long long)c->n_inst,
			(unsigned long long)*nedges, ms, c->last_ms);
	}
	return SIBGPU_OK;
}

// BlockFinder::TrimBlocks (src/synteny.cpp:31-122) without the host-side index.  The reference walks every sequence of
// the block along the block's direction and, at every vertex mark `it`, looks at all instances `kmer` of that vertex
// that lie on ANOTHER sequence (either strand); with distances measured in elements from the two ends of a sequence
// as seen along its block direction (IndexedSequence::StrandIteratorDistance, src/indexedsequence.cpp:162-167) it keeps
//     trimStart = first `it` minimising (dStart(it) + dStart(kmer), vertex id),   trimEnd likewise with dEnd.
// Only the minimum over the kmers matters, so per vertex we keep the smallest dStart/dEnd together with the sequence it
// comes from and the smallest one from any other sequence; the scan over the marks is then O(instances).
// the O(instances) part of sibgpu_trim_blocks: the two instance tables of the enumeration -> trim points
static void trim_from_tables(const sibgpu_inst *const tab[2], const uint64_t ntab[2], uint32_t count, const uint64_t *len,
	const uint8_t *direction, uint32_t nchr, sibgpu_trim *out)
{
	const uint32_t INF = 0xFFFFFFFFu;
	struct Best { uint32_t b1, c1, b2; };
	std::vector<Best> bs(count, Best{INF, INF, INF}), be(count, Best{INF, INF, INF});
	auto add = [&](std::vector<Best> &v, uint32_t id, uint32_t d, uint32_t chr) {
		Best &b = v[id];
		if(chr == b.c1) { if(d < b.b1) b.b1 = d; }
		else if(d < b.b1) { b.b2 = b.b1; b.b1 = d; b.c1 = chr; }
		else if(d < b.b2) b.b2 = d;
	};
	for(int strand = 0; strand < 2; strand++)
	{
		for(uint64_t i = 0; i < ntab[strand]; i++)
		{
			const sibgpu_inst &x = tab[strand][i];
			const uint32_t L = (uint32_t)len[x.chr];
			const uint32_t elem = strand == 0 ? x.pos : L - 1 - x.pos;
			const uint32_t dstart = direction[x.chr] == 0 ? elem : L - 1 - elem;       // along the block direction of ITS sequence
			add(bs, x.bifId, dstart, x.chr);
			add(be, x.bifId, L - 1 - dstart, x.chr);
		}
	}
	for(uint32_t chr = 0; chr < nchr; chr++) out[chr] = sibgpu_trim{0u, 0u, 0u};
	std::vector<uint64_t> best_s(nchr, ~0ull), best_e(nchr, ~0ull);              // (sum << 32 | vertex id), first minimum wins
	for(int strand = 0; strand < 2; strand++)
	{
		for(uint64_t i = 0; i < ntab[strand]; i++)
		{
			const sibgpu_inst &x = tab[strand][i];
			if((direction[x.chr] != 0) != (strand != 0)) continue;              // `it` walks the block direction only
			const uint32_t L = (uint32_t)len[x.chr];
			const Best &s0 = bs[x.bifId], &e0 = be[x.bifId];
			const uint32_t os = s0.c1 != x.chr ? s0.b1 : s0.b2, oe = e0.c1 != x.chr ? e0.b1 : e0.b2;
			if(os == INF) continue;                                             // the vertex occurs on this sequence only
			const uint32_t elem = strand == 0 ? x.pos : L - 1 - x.pos;
			const uint64_t ks = ((uint64_t)(x.pos + os) << 32) | x.bifId, ke = ((uint64_t)(L - 1 - x.pos + oe) << 32) | x.bifId;
			if(ks < best_s[x.chr]) { best_s[x.chr] = ks; out[x.chr].start = elem; }
			if(ke < best_e[x.chr]) { best_e[x.chr] = ke; out[x.chr].end = elem; }
			out[x.chr].found = 1;
		}
	}
}

int sibgpu_trim_blocks(sibgpu_ctx *c, const char *const *seq, const uint64_t *len, const uint8_t *direction, uint32_t nchr,
	uint32_t trim_k, sibgpu_trim *out)
{
	if(!c || (nchr && (!seq || !len || !direction || !out)) || trim_k == 0)
	{
		set_error("invalid: NULL argument or trim_k == 0");
		return SIBGPU_ERR_INVALID;
	}
	sibgpu_inst *tab[2] = {nullptr, nullptr};
	uint64_t ntab[2] = {0, 0};
	uint32_t count = 0;
	SIB_TRY(sibgpu_enumerate(c, seq, len, nchr, trim_k, &tab[0], &ntab[0], &tab[1], &ntab[1], &count));
	trim_from_tables(tab, ntab, count, len, direction, nchr, out);
	sibgpu_free(tab[0]);
	sibgpu_free(tab[1]);
	return SIBGPU_OK;
}

void sibgpu_debug_trim_from_tables(const sibgpu_inst *pos, uint64_t npos, const sibgpu_inst *neg, uint64_t nneg, uint32_t count,
	const uint64_t *len, const uint8_t *direction, uint32_t nchr, sibgpu_trim *out)
{
	const sibgpu_inst *tab[2] = {pos, neg};
	const uint64_t ntab[2] = {npos, nneg};
	trim_from_tables(tab, ntab, count, len, direction, nchr, out);
}

// ---------------------------------------------------------------------------------------------------------------
// sharded enumeration (one process per GPU)
// ---------------------------------------------------------------------------------------------------------------
// Layout of a sharded run: this rank's tile range, the bytes it reads (tiles + halos), buffers, '$' fill, tables.
static int dist_layout(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world)
{
	if(!c || (nchr && (!chr || !len)) || world == 0 || rank >= world || world > 64)
	{
		set_error("invalid: NULL argument or bad rank/world");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	c->have_text = false;
	c->have_result = false;
	uint64_t N = 0;
	for(uint32_t i = 0; i < nchr; i++) N += len[i];
	const uint64_t M = N + nchr + 1;
	if(M >= (1ull << 31))
	{
		set_error("invalid: input of " + std::to_string(N) + " bases exceeds the 32-bit position range");
		return SIBGPU_ERR_INVALID;
	}
	c->h_chr_start.resize(nchr);
	c->h_chr_len.resize(nchr);
	uint64_t at = 1;
	for(uint32_t i = 0; i < nchr; i++)
	{
		c->h_chr_start[i] = (uint32_t)at;
		c->h_chr_len[i] = (uint32_t)len[i];
		at += len[i] + 1;
	}
	const size_t nwords = (size_t)((M + 15) / 16) + 8;
	const uint64_t tile_pos = 4096;
	const uint64_t ntiles = (M + tile_pos - 1) / tile_pos;
	c->dist_rank = rank;
	c->dist_world = world;
	c->dist_tile_lo = (uint32_t)(ntiles * rank / world);
	c->dist_tile_hi = (uint32_t)(ntiles * (rank + 1) / world);
	// bytes this rank reads: its tiles, one word of back halo, k + 1 <= 33 bases and the staged words of forward halo
	uint64_t b_lo = (uint64_t)c->dist_tile_lo * tile_pos, b_hi = (uint64_t)c->dist_tile_hi * tile_pos + 128;
	b_lo = b_lo >= 16 ? b_lo - 16 : 0;
	if(b_hi > nwords * 16) b_hi = nwords * 16;
	if(c->dist_tile_hi == c->dist_tile_lo) b_hi = b_lo;
	c->dist_byte_lo = b_lo;
	c->dist_byte_hi = b_hi;
	SIB_TRY(c->d_text.ensure(nwords * 16));
	SIB_TRY(c->d_chr_start.ensure(sizeof(uint32_t) * (nchr + 1)));
	SIB_TRY(c->d_chr_len.ensure(sizeof(uint32_t) * (nchr + 1)));
	if(b_hi > b_lo) SIB_CUDA(cudaMemsetAsync(c->d_text.as<char>() + b_lo, '$', b_hi - b_lo, c->stream));
	if(nchr)
	{
		SIB_CUDA(cudaMemcpyAsync(c->d_chr_start.p, c->h_chr_start.data(), sizeof(uint32_t) * nchr, cudaMemcpyHostToDevice, c->stream));
		SIB_CUDA(cudaMemcpyAsync(c->d_chr_len.p, c->h_chr_len.data(), sizeof(uint32_t) * nchr, cudaMemcpyHostToDevice, c->stream));
	}
	c->nchr = nchr;
	c->N = N;
	c->M = M;
	c->dist_result = false;
	return SIBGPU_OK;
}

int sibgpu_dist_upload(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world)
{
	SIB_TRY(dist_layout(c, chr, len, nchr, rank, world));
	HostSrc src = {chr, len};
	if(c->dist_byte_hi > c->dist_byte_lo) SIB_TRY(copy_text_range(c, src, c->dist_byte_lo, c->dist_byte_hi, c->stream));
	SIB_CUDA(cudaStreamSynchronize(c->stream));
	c->have_text = true;
	return SIBGPU_OK;
}

int sibgpu_dist_upload_scatter(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world,
	uint32_t k, uint32_t *nparts_total, uint64_t *counts, uint64_t *seg_cap, int *overflow)
{
	if(!counts || !nparts_total || !seg_cap || !overflow || k == 0 || k > 32)
	{
		set_error("invalid: the sharded path supports 1 <= k <= 32");
		return SIBGPU_ERR_INVALID;
	}
	SIB_TRY(dist_layout(c, chr, len, nchr, rank, world));
	HostSrc src = {chr, len};
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(dist_scatter_local(c, k, counts, seg_cap, overflow, &src));
	c->have_text = true;
	*nparts_total = c->dist_P_total;
	return SIBGPU_OK;
}

int sibgpu_dist_scan(sibgpu_ctx *c, uint32_t k, uint32_t *nparts_total, uint32_t *hist, uint64_t *nrec_local)
{
	if(!c || !hist || !nparts_total || k == 0 || k > 32)
	{
		set_error("invalid: the sharded path supports 1 <= k <= 32");
		return SIBGPU_ERR_INVALID;
	}
	if(!c->have_text)
	{
		set_error("state: sibgpu_dist_upload must come first");
		return SIBGPU_ERR_STATE;
	}
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(dist_scan(c, k, hist));
	*nparts_total = c->dist_P_total;
	if(nrec_local) *nrec_local = c->dist_nrec_local;
	return SIBGPU_OK;
}

uint32_t sibgpu_dist_record_bytes(sibgpu_ctx *c) { return c && c->last_k > 28 ? 16u : 8u; }

int sibgpu_dist_scatter(sibgpu_ctx *c, void *send_dev)
{
	if(!c || (!send_dev && c->dist_nrec_local))
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	return dist_scatter(c, send_dev);
}

int sibgpu_dist_group(sibgpu_ctx *c, const void *recv_dev, const uint32_t *counts, uint64_t *nkeys_local)
{
	if(!c || !counts || !nkeys_local)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	return dist_group(c, recv_dev, counts, nkeys_local);
}

int sibgpu_dist_keys(sibgpu_ctx *c, void *keys_dev)
{
	if(!c) return SIBGPU_ERR_INVALID;
	if(c->dist_nkeys_local)
	{
		SIB_CUDA(cudaMemcpyAsync(keys_dev, c->d_ckeys.p, c->dist_nkeys_local * sibgpu_dist_record_bytes(c), cudaMemcpyDeviceToDevice, c->stream));
		SIB_CUDA(cudaStreamSynchronize(c->stream));
	}
	return SIBGPU_OK;
}

int sibgpu_dist_finish(sibgpu_ctx *c, const void *allkeys_dev, uint64_t nkeys_total, uint64_t *ninst_local, uint32_t *count)
{
	if(!c)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_TRY(dist_finish(c, allkeys_dev, nkeys_total));
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	if(ninst_local) *ninst_local = c->n_inst;
	if(count) *count = c->n_vertices;
	return SIBGPU_OK;
}

int sibgpu_dist_scatter_local(sibgpu_ctx *c, uint32_t k, uint32_t *nparts_total, uint64_t *counts, uint64_t *seg_cap, int *overflow)
{
	if(!c || !counts || !nparts_total || !seg_cap || !overflow || k == 0 || k > 32)
	{
		set_error("invalid: the sharded path supports 1 <= k <= 32");
		return SIBGPU_ERR_INVALID;
	}
	if(!c->have_text)
	{
		set_error("state: sibgpu_dist_upload must come first");
		return SIBGPU_ERR_STATE;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(dist_scatter_local(c, k, counts, seg_cap, overflow, nullptr));
	*nparts_total = c->dist_P_total;
	return SIBGPU_OK;
}

int sibgpu_dist_export_send(sibgpu_ctx *c, void *handle64)
{
	if(!c || !handle64 || !c->d_sendbuf.p)
	{
		set_error("invalid: NULL argument or no send buffer yet (sibgpu_dist_scatter_local comes first)");
		return SIBGPU_ERR_INVALID;
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	SIB_CUDA(cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	SIB_CUDA(cudaIpcGetMemHandle(&h, c->d_sendbuf.p));
	memcpy(handle64, &h, 64);
	return SIBGPU_OK;
}

int sibgpu_dist_import_peers(sibgpu_ctx *c, const void *handles)
{
	if(!c || !handles)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	const uint32_t W = c->dist_world;
	c->peer_ptr.resize(W, nullptr);
	c->peer_handle.resize(W);
	for(uint32_t s = 0; s < W; s++)
	{
		if(s == c->dist_rank) continue;
		const unsigned char *h = static_cast<const unsigned char*>(handles) + 64 * (size_t)s;
		if(c->peer_ptr[s] && c->peer_handle[s].size() == 64 && memcmp(c->peer_handle[s].data(), h, 64) == 0) continue;
		if(c->peer_ptr[s])
		{
			cudaIpcCloseMemHandle(c->peer_ptr[s]);
			c->peer_ptr[s] = nullptr;
		}
		cudaIpcMemHandle_t ih;
		memcpy(&ih, h, 64);
		SIB_CUDA(cudaIpcOpenMemHandle(&c->peer_ptr[s], ih, cudaIpcMemLazyEnablePeerAccess));
		c->peer_handle[s].assign(h, h + 64);
	}
	return SIBGPU_OK;
}

int sibgpu_dist_group_peer(sibgpu_ctx *c, const uint64_t *counts, const uint64_t *seg_caps, uint64_t *nkeys_local)
{
	if(!c || !counts || !seg_caps || !nkeys_local)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	return dist_group_peer(c, counts, seg_caps, nkeys_local);
}

// ---- fused sharded path
int sibgpu_fused_plan(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t rank, uint32_t world,
	uint32_t k, int resident, int *need_alloc)
{
	if(!c || !need_alloc || k == 0)
	{
		set_error("invalid: NULL argument or k == 0");
		return SIBGPU_ERR_INVALID;
	}
	if(resident)
	{
		if(!c->have_text || c->dist_world != world || c->dist_rank != rank)
		{
			set_error("state: resident sharded run needs sibgpu_dist_upload with the same rank / world first");
			return SIBGPU_ERR_STATE;
		}
	}
	else SIB_TRY(dist_layout(c, chr, len, nchr, rank, world));
	return dist2_plan(c, k, need_alloc);
}

int sibgpu_fused_release_peers(sibgpu_ctx *c) { return c ? dist2_release_peers(c) : SIBGPU_ERR_INVALID; }

int sibgpu_fused_alloc(sibgpu_ctx *c, void *handle64)
{
	if(!c || !handle64)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	return dist2_alloc(c, handle64);
}

int sibgpu_fused_import(sibgpu_ctx *c, const void *handles)
{
	if(!c || !handles)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	return dist2_import(c, handles);
}

int sibgpu_fused_run(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, int resident, uint32_t *count,
	uint64_t *ninst_local, int *status)
{
	if(!c || !status || (!resident && nchr && (!chr || !len)))
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	HostSrc src = {chr, len};
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(dist2_run(c, resident ? nullptr : &src, status));
	c->have_text = true;
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	if(ninst_local) *ninst_local = c->n_inst;
	if(count) *count = c->n_vertices;
	return SIBGPU_OK;
}

int sibgpu_fused_run_fp(sibgpu_ctx *c, const char *const *chr, const uint64_t *len, uint32_t nchr, int resident, uint32_t attempt,
	uint64_t *nclasses, void **rep_dev, int *status)
{
	if(!c || !status || !nclasses || !rep_dev || (!resident && nchr && (!chr || !len)))
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	HostSrc src = {chr, len};
	SIB_CUDA(cudaEventRecord(c->ev_begin, c->stream));
	SIB_TRY(dist2_run_fp(c, resident ? nullptr : &src, attempt, status));
	c->have_text = true;
	*nclasses = c->x_Vc;
	*rep_dev = c->x_Vc ? c->d_rep.p : nullptr;
	return SIBGPU_OK;
}

int sibgpu_fused_finish_fp(sibgpu_ctx *c, uint32_t *count, uint64_t *ninst_local, int *collision)
{
	if(!c || !collision)
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaSetDevice(c->device));
	SIB_TRY(dist2_finish_fp(c, collision));
	SIB_CUDA(cudaEventRecord(c->ev_end, c->stream));
	SIB_CUDA(cudaEventSynchronize(c->ev_end));
	SIB_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev_begin, c->ev_end));
	if(ninst_local) *ninst_local = c->n_inst;
	if(count) *count = c->n_vertices;
	return SIBGPU_OK;
}

int sibgpu_set_profiling(sibgpu_ctx *c, int enabled)
{
	if(!c) return SIBGPU_ERR_INVALID;
	c->profiling = enabled != 0;
	return SIBGPU_OK;
}

int sibgpu_kernel_stats(sibgpu_ctx *c, sibgpu_kernel_stat *out, int cap)
{
	if(!c) return 0;
	if(c->profiling && !c->spans.empty() && c->stats.empty()) c->prof_collect();   // phases run outside a whole enumerate
	int n = (int)c->stats.size();
	for(int i = 0; i < n && i < cap; i++)
	{
		out[i].name = c->stats[i].name;
		out[i].launches = c->stats[i].launches;
		out[i].ms = c->stats[i].ms;
		out[i].algo_bytes = c->stats[i].bytes;
	}
	return n;
}

uint64_t sibgpu_last_launches(sibgpu_ctx *c) { return c ? c->total_launches : 0; }
uint64_t sibgpu_partition_fallbacks(sibgpu_ctx *c) { return c ? c->hist_fallbacks : 0; }
uint64_t sibgpu_bucket_fallbacks(sibgpu_ctx *c) { return c ? c->smem_fallbacks : 0; }
float sibgpu_last_device_ms(sibgpu_ctx *c) { return c ? c->last_ms : 0.f; }

}
