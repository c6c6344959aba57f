// libsibgpu context: device buffers, stream, per-kernel timing.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "hostvec.h"

namespace sibgpu {
// device-side description of the concatenated text '$' chr0 '$' chr1 ... '$'
struct TextDesc {
	const uint32_t *packed;        // 16 bases per word, first base in the top bit pair
	const uint32_t *chr_start;     // text index of the first base of every chromosome
	const uint32_t *chr_len;
	uint32_t nchr;
	uint32_t M;                    // text length
	uint32_t nwords;               // valid words in packed[]
	uint32_t tile0;                // first 4096-position tile of this rank's text range (0 on a single GPU)
};
}

struct sibgpu_ctx {
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;

	// resident input (sibgpu_upload)
	bool have_text = false;
	std::vector<uint32_t> h_chr_start, h_chr_len;     // text coordinates of every chromosome
	uint32_t nchr = 0;
	uint64_t M = 0;                                    // text length: N + nchr + 1  ('$' chr0 '$' chr1 ... '$')
	uint64_t N = 0;                                    // total bases
	sibgpu::DevBuf d_text, d_packed, d_chr_start, d_chr_len;

	// enumeration workspace
	sibgpu::DevBuf d_hist, d_partoff, d_cursor, d_records, d_table, d_partcnt, d_keyoff, d_ckeys, d_vkeys, d_vkeys_alt,
		d_cubtmp, d_map, d_filter, d_hitmask, d_tilecnt, d_tileoff, d_pos, d_negtmp, d_neg, d_chrinst, d_scalars,
		d_fp, d_fpprm, d_rep, d_order, d_records2, d_cnt2, d_s_ch, d_s_m0, d_s_m1, d_s_off, d_s_inst, d_s_flag, d_edges, d_edge_skip;
	void *h_scalars = nullptr;                         // pinned, 64 x u64

	// last result
	bool have_result = false;
	uint64_t n_inst = 0;                               // instances per strand
	uint32_t n_vertices = 0;
	uint32_t last_k = 0;

	// sharded mode (sibgpu_dist_*)
	uint32_t dist_rank = 0, dist_world = 1, dist_tile_lo = 0, dist_tile_hi = 0, dist_P_local = 0, dist_P_total = 0;
	uint64_t dist_byte_lo = 0, dist_byte_hi = 0, dist_nrec_local = 0, dist_nkeys_local = 0;
	bool dist_result = false;
	// peer path: fixed-capacity segments in the own send buffer (d_records), peers' send buffers mapped through CUDA IPC
	uint64_t dist_seg_cap = 0;
	std::vector<void*> peer_ptr;                       // [world], nullptr for the own rank / not mapped
	std::vector<std::vector<unsigned char>> peer_handle;
	sibgpu::DevBuf d_keystage;                         // vertex keys of the owned partitions
	// the send buffer the peers map: never shared with another path, and a buffer that had to grow is freed two steps
	// later (by then every peer has closed its mapping of it: it re-imports the handles at every step)
	sibgpu::DevBuf d_sendbuf;
	std::vector<void*> send_retired_old, send_retired_new;
	// fused path (sibgpu_fused_*): ONE exported buffer per rank [header | cursors | vertex keys | segments]
	sibgpu::DevBuf d_xbuf;
	std::vector<void*> peer_x;                         // [world] mappings of the peers' exported buffers
	uint64_t x_kc = 0, x_kc_live = 0, x_off_seg = 0, x_off_seg_plan = 0, x_seg_cap = 0, x_nrec = 0, x_bytes = 0;
	uint64_t x_off_pk = 0, x_off_pk_plan = 0;             // k > 32: packed words of the own text range, pulled by the peers
	uint64_t x_Vc = 0;                                    // k > 32: classes of the step between its two phases
	uint32_t x_attempt = 0;
	uint32_t x_PL = 0, x_sub_bits = 0, x_k = 0;
	unsigned long long dist_epoch = 0;                 // step counter, the same on all ranks
	sibgpu::TextDesc dist_text = {};

	// tunables (env SIBGPU_PART_RECORDS)
	uint64_t part_target = 1u << 20;                   // records per partition for 8-byte table slots ...
	bool part_explicit = false;                        // ... scaled down for wider slots unless the env var pins it
	// records per hash partition such that one partition's table (table_factor slots per record) stays near 16 MB and
	// the tables of the overlapped streams stay L2-resident: 8-byte slots (k <= 26) 1 Mi, 16-byte slots (all other k) 512 Ki
	uint64_t part_records(uint32_t k) const
	{
		if(part_explicit) return part_target;
		const uint64_t slot = k <= 26 ? 8 : 16;
		const uint64_t r = part_target * 8 / slot;
		return r < 65536 ? 65536 : r;
	}
	uint64_t part_slack = 4096;                        // fixed-capacity partitions: mean + mean/8 + slack (env SIBGPU_PART_SLACK)
	bool exact_hist = false;                           // always size the partitions with a histogram pass (env SIBGPU_EXACT_HIST)
	uint64_t hist_fallbacks = 0;                       // times a fixed-capacity partition overflowed and the exact path ran
	// pipelined upload (sibgpu_enumerate): copies on their own stream, one event per text piece
	cudaStream_t copy_stream = nullptr;
	cudaEvent_t ev_fork_copy = nullptr;
	std::vector<cudaEvent_t> ev_chunk;
	int ensure_copy_stream(uint32_t nchunks);
	// pageable host sources (std::string::data() of the reference's facade): the pieces are copied by host threads into a
	// ring of pinned staging buffers and go to the device from there (env SIBGPU_STAGE_THREADS, 0 = let the driver stage)
	void *h_stage = nullptr;
	size_t stage_piece = 0;                            // bytes per ring slot
	static constexpr uint32_t STAGE_SLOTS = 4;
	int stage_threads = 4;                             // 4: 4.5 ms per 100 MB end to end, 8: 4.8, 16: 5.7 (driver staging: 10.9)
	int ensure_stage(size_t piece_bytes);
	int insert_variant = 1;                            // 1 = CAS first, one record per thread (env SIBGPU_INSERT_VARIANT, dev)
	int n_streams = 4;                                 // overlapped partition streams (env SIBGPU_STREAMS)
	cudaStream_t aux_stream[8] = {};
	cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
	int ensure_aux_streams(uint32_t n);
	// grouping of 8-byte records (k <= 28): 1 = buckets of ~1 Ki records grouped in shared memory (k_split + k_group),
	// 0 = one L2-resident table per hash partition (k_insert + k_table_scan; also the fallback when a bucket overflows)
	int group_smem = 1;                                // env SIBGPU_GROUP_SMEM
	bool split_attr_done[2] = {false, false}, group_attr_done[4] = {false, false, false, false};
	int piecewise_split = 1;                           // host source: split every text piece behind its scatter (env SIBGPU_PIECEWISE, dev)
	int split_stages = 2;                              // input tiles in flight per CTA of k_split (env SIBGPU_SPLIT_STAGES, dev)
	uint64_t ckeys_init = 1u << 20;                    // initial capacity of the vertex-key list (env SIBGPU_CKEYS_INIT, tests)
	uint64_t smem_fallbacks = 0;                       // times a bucket overflowed and the L2-table path took over
	int table_factor = 2;                              // slots per record of the largest partition (env SIBGPU_TABLE_FACTOR)

	// sibgpu_simplify: the per-element host arrays of a stage are recycled between stages (a fresh 30 B/element
	// allocation per stage costs more in page faults than the stage's device work)
	sibgpu::HostChars pool_ch;
	sibgpu::HostU32 pool_u32[3];
	sibgpu::HostI32 pool_i32[4];

	// profiling
	bool profiling = false;
	struct Span { const char *name; cudaEvent_t a, b; uint64_t bytes; };
	std::vector<Span> spans;
	std::vector<cudaEvent_t> event_pool;
	size_t events_used = 0;
	struct Stat { const char *name; uint32_t launches; float ms; uint64_t bytes; };
	std::vector<Stat> stats;
	uint64_t total_launches = 0;
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	float last_ms = 0.f;

	cudaEvent_t get_event();
	void prof_begin(const char *name, uint64_t bytes);
	void prof_end();
	void prof_reset();
	int prof_collect();
};

namespace sibgpu {
struct ProfScope {
	sibgpu_ctx *c;
	ProfScope(sibgpu_ctx *ctx, const char *name, uint64_t bytes, uint32_t launches = 1) : c(ctx)
	{
		c->total_launches += launches;
		if(c->profiling) c->prof_begin(name, bytes);
	}
	~ProfScope()
	{
		if(c->profiling) c->prof_end();
	}
};

// the caller's chromosomes while sibgpu_enumerate streams them in (layout already planned by upload_layout)
struct HostSrc { const char *const *chr; const uint64_t *len; };
int copy_text_range(sibgpu_ctx *ctx, const HostSrc &src, uint64_t lo, uint64_t hi, cudaStream_t st);
int enumerate_resident(sibgpu_ctx *ctx, uint32_t k, const HostSrc *src);
void *pool_alloc(size_t bytes);                        // api.cu: pinned, pooled result buffer (release with sibgpu_free)
int enumerate_keep(sibgpu_ctx *ctx, const char *const *chr, const uint64_t *len, uint32_t nchr, uint32_t k);   // api.cu
int list_edges_device(sibgpu_ctx *ctx, uint32_t k, sibgpu_edge **edges_out, uint64_t *nedges_out);   // edges.cu
int dist_scan(sibgpu_ctx *ctx, uint32_t k, uint32_t *hist_out);
int dist_scatter(sibgpu_ctx *ctx, void *send_dev);
int dist_group(sibgpu_ctx *ctx, const void *recv_dev, const uint32_t *counts, uint64_t *nkeys_local);
int dist_finish(sibgpu_ctx *ctx, const void *allkeys_dev, uint64_t nkeys_total);
int dist_scatter_local(sibgpu_ctx *ctx, uint32_t k, uint64_t *counts_out, uint64_t *seg_cap_out, int *overflow_out, const HostSrc *src);
int dist_group_peer(sibgpu_ctx *ctx, const uint64_t *counts, const uint64_t *seg_caps, uint64_t *nkeys_local);
int dist2_plan(sibgpu_ctx *ctx, uint32_t k, int *need_alloc);
int dist2_release_peers(sibgpu_ctx *ctx);
int dist2_alloc(sibgpu_ctx *ctx, void *handle64);
int dist2_import(sibgpu_ctx *ctx, const void *handles);
int dist2_run(sibgpu_ctx *ctx, const HostSrc *src, int *status);
int dist2_run_fp(sibgpu_ctx *ctx, const HostSrc *src, uint32_t attempt, int *status);
int dist2_finish_fp(sibgpu_ctx *ctx, int *collision);
} // namespace sibgpu
