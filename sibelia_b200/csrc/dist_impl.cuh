// Sharded enumeration over `world` GPUs, one process per GPU (included at the end of enumerate.cu).
//
// The concatenated text is split into `world` contiguous tile ranges; every rank scans only its range.  Records are
// partitioned by the SAME hash function on every rank into P_total = world * P_local partitions; partition p belongs
// to rank p / P_local, so after one all-to-all (done by the caller with NCCL on the send/recv device buffers handed
// in here) every rank holds ALL occurrences of the k-mer classes it owns and decides them locally.  The (few) vertex
// keys are then all-gathered so that every rank can compute the global lexicographic ids and emit the instances of its
// own text range.  k <= 32 only (the k > 32 ranking needs the whole text on every rank).
//
//   dist_scan     pack own range, histogram over P_total partitions          -> counts (host)
//   dist_scatter  records of the own range, ordered by partition             -> send buffer
//   [caller: all_gather(counts), all_to_all(records)]
//   dist_group    per owned partition: insert the world segments, predicate  -> local canonical vertex keys
//   [caller: all_gather(keys)]
//   dist_finish   global ids + map, mark/emit over the own range             -> local instance tables (text order)

template<int MODE>
static int dist_scan_mode(sibgpu_ctx *ctx, uint32_t k, uint32_t *hist_out)
{
	cudaStream_t st = ctx->stream;
	TextDesc t = ctx->dist_text;
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
	SIB_TRY(ctx->d_hist.ensure(sizeof(uint32_t) * MAX_PARTS));
	SIB_CUDA(cudaMemsetAsync(ctx->d_hist.p, 0, sizeof(uint32_t) * MAX_PARTS, st));
	if(ntiles)
	{
		const uint32_t g = ntiles < (uint32_t)ctx->sm_count * 8 ? ntiles : (uint32_t)ctx->sm_count * 8;
		ProfScope ps(ctx, "k_scan_hist", (uint64_t)ntiles * TILE_POS / 4);
		k_scan_hist<MODE><<<g, TILE_THREADS, 0, st>>>(t, FpView{}, k, ntiles, ctx->dist_P_total, ctx->d_hist.as<uint32_t>());
	}
	SIB_CUDA(cudaMemcpyAsync(hist_out, ctx->d_hist.p, sizeof(uint32_t) * ctx->dist_P_total, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	return SIBGPU_OK;
}

template<int MODE>
static int dist_scatter_mode(sibgpu_ctx *ctx, uint32_t k, void *send_dev)
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	TextDesc t = ctx->dist_text;
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
	SIB_TRY(ctx->d_partoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
	SIB_TRY(ctx->d_cursor.ensure(sizeof(uint64_t) * MAX_PARTS * CURSOR_STRIDE));
	k_part_offsets<<<1, MAX_PARTS, 0, st>>>(ctx->d_hist.as<uint32_t>(), ctx->dist_P_total, ctx->d_partoff.as<uint64_t>(),
		ctx->d_cursor.as<unsigned long long>(), ctx->d_scalars.as<uint64_t>());
	ctx->total_launches++;
	if(ntiles)
	{
		size_t smem = scatter_smem_bytes<MODE, false>();
		SIB_CUDA(cudaFuncSetAttribute(k_scatter<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		const uint32_t g = ntiles < (uint32_t)ctx->sm_count * 4 ? ntiles : (uint32_t)ctx->sm_count * 4;
		ProfScope ps(ctx, "k_scatter", (uint64_t)ntiles * TILE_POS / 4 + ctx->dist_nrec_local * sizeof(Rec));
		k_scatter<MODE, false><<<g, TILE_THREADS, smem, st>>>(t, FpView{}, k, ntiles, ctx->dist_P_total,
			ctx->d_cursor.as<unsigned long long>(), static_cast<Rec*>(send_dev), 0ull, nullptr);
	}
	SIB_CUDA(cudaStreamSynchronize(st));
	return SIBGPU_OK;
}

template<int MODE>
static int dist_group_mode(sibgpu_ctx *ctx, uint32_t k, const void *recv_dev, const uint32_t *counts, uint64_t *nkeys_local)
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	const int sms = ctx->sm_count;
	const uint32_t W = ctx->dist_world, PL = ctx->dist_P_local, PT = ctx->dist_P_total, b0 = ctx->dist_rank * PL;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	// layout of the receive buffer: for every source rank s, its records of my partitions in partition order
	std::vector<uint64_t> src_off(W + 1, 0), stage_off(PL + 1, 0);
	uint64_t maxpart = 0;
	for(uint32_t s = 0; s < W; s++)
	{
		uint64_t sum = 0;
		for(uint32_t p = 0; p < PL; p++) sum += counts[(size_t)s * PT + b0 + p];
		src_off[s + 1] = src_off[s] + sum;
	}
	for(uint32_t p = 0; p < PL; p++)
	{
		uint64_t sum = 0;
		for(uint32_t s = 0; s < W; s++) sum += counts[(size_t)s * PT + b0 + p];
		stage_off[p + 1] = stage_off[p] + sum;
		if(sum > maxpart) maxpart = sum;
	}
	const uint64_t recv_total = src_off[W];
	*nkeys_local = 0;
	ctx->dist_nkeys_local = 0;
	if(recv_total == 0) return SIBGPU_OK;
	const uint64_t T64 = (uint64_t)ctx->table_factor * maxpart + 1024;
	if(T64 > 0xFFFFFF00ull)
	{
		set_error("internal: hash partition does not fit a 32-bit table");
		return SIBGPU_ERR_INTERNAL;
	}
	const uint32_t T = (uint32_t)T64;
	const bool compact = MODE == 0 && k <= COMPACT_MAX_K;
	const size_t slot_bytes = compact ? 8 : sizeof(Slot8);
	SIB_TRY(ctx->d_table.ensure(slot_bytes * T));
	SIB_CUDA(cudaMemsetAsync(ctx->d_table.p, 0xFF, slot_bytes * T, st));
	SIB_TRY(ctx->d_records.ensure(sizeof(Rec) * recv_total));          // staging area of the vertex keys, per partition
	SIB_TRY(ctx->d_partcnt.ensure(sizeof(uint32_t) * MAX_PARTS));
	SIB_TRY(ctx->d_keyoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
	SIB_TRY(ctx->d_partoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
	SIB_CUDA(cudaMemsetAsync(ctx->d_partcnt.p, 0, sizeof(uint32_t) * MAX_PARTS, st));
	SIB_CUDA(cudaMemcpyAsync(ctx->d_partoff.p, stage_off.data(), sizeof(uint64_t) * (PL + 1), cudaMemcpyHostToDevice, st));
	const Rec *recv = static_cast<const Rec*>(recv_dev);
	std::vector<uint64_t> seg_cursor(src_off.begin(), src_off.end() - 1);
	for(uint32_t p = 0; p < PL; p++)
	{
		if(stage_off[p + 1] == stage_off[p]) continue;
		for(uint32_t s = 0; s < W; s++)
		{
			const uint64_t n = counts[(size_t)s * PT + b0 + p];
			if(n == 0) continue;
			const Rec *seg = recv + seg_cursor[s];
			seg_cursor[s] += n;
			ProfScope ps(ctx, "k_insert", n * sizeof(Rec));
			if(compact) launch_insert_compact(ctx->insert_variant, sms, st, reinterpret_cast<const uint64_t*>(seg), n,
				ctx->d_table.as<unsigned long long>(), T);
			else k_insert<MODE><<<grid_for(n, 256, sms, 8), 256, 0, st>>>(seg, n, ctx->d_table.p, T);
		}
		Rec *out = ctx->d_records.as<Rec>() + stage_off[p];
		ProfScope ps(ctx, "k_table_scan", (uint64_t)T * slot_bytes);
		if(compact) k_table_scan_compact<<<grid_for(T, 256, sms, 8), 256, 0, st>>>(ctx->d_table.as<unsigned long long>(), T,
			reinterpret_cast<uint64_t*>(out), ctx->d_partcnt.as<uint32_t>() + p);
		else k_table_scan<MODE><<<grid_for(T, 256, sms, 8), 256, 0, st>>>(ctx->d_table.p, T, out, ctx->d_partcnt.as<uint32_t>() + p);
	}
	k_key_offsets<<<1, MAX_PARTS, 0, st>>>(ctx->d_partcnt.as<uint32_t>(), PL, ctx->d_keyoff.as<uint64_t>(), ds);
	ctx->total_launches++;
	SIB_CUDA(cudaMemcpyAsync(hs + 2, ds + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	const uint64_t Vc = hs[2];
	if(Vc)
	{
		SIB_TRY(ctx->d_ckeys.ensure(sizeof(Rec) * Vc));
		ProfScope ps(ctx, "k_gather_keys", 2 * Vc * sizeof(Rec));
		dim3 g(8, PL);
		k_gather_keys<MODE><<<g, 256, 0, st>>>(ctx->d_records.as<Rec>(), ctx->d_partoff.as<uint64_t>(),
			ctx->d_partcnt.as<uint32_t>(), ctx->d_keyoff.as<uint64_t>(), ctx->d_ckeys.as<Rec>());
		SIB_CUDA(cudaStreamSynchronize(st));
	}
	ctx->dist_nkeys_local = Vc;
	*nkeys_local = Vc;
	return SIBGPU_OK;
}

template<int MODE>
static int dist_finish_mode(sibgpu_ctx *ctx, uint32_t k, const void *allkeys_dev, uint64_t nkeys_total)
{
	typedef typename RecT<MODE>::type Rec;
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
	ctx->n_inst = 0;
	ctx->n_vertices = 0;
	if(nkeys_total == 0) return SIBGPU_OK;
	if(2 * nkeys_total > 0xFFFFFFF0ull)
	{
		set_error("invalid: more than 2^32 vertices");
		return SIBGPU_ERR_INVALID;
	}
	if(ntiles == 0)
	{
		// nothing to emit here, but the vertex count is still a global quantity: rank the keys
		ctx->dist_text.tile0 = 0;
	}
	bool collision = false;
	return ids_and_tables<MODE>(ctx, ctx->dist_text, k, static_cast<const Rec*>(allkeys_dev), nkeys_total,
		ntiles ? ntiles : 0, FpView{}, 0u, false, &collision);
}

// pack the words of the own range (+ halo) unless the caller pipelines that with the upload, partition plan, text descriptor
static int dist_prepare(sibgpu_ctx *ctx, uint32_t k, bool pack = true)
{
	cudaStream_t st = ctx->stream;
	SIB_CUDA(cudaSetDevice(ctx->device));
	ctx->have_result = false;
	ctx->prof_reset();
	ctx->total_launches = 0;
	// K0 over the words of the own range (+ halo)
	const uint64_t w_lo = ctx->dist_byte_lo / 16, w_hi = (ctx->dist_byte_hi + 15) / 16;
	const uint32_t nwords_all = (uint32_t)((ctx->M + 15) / 16) + 8;
	SIB_TRY(ctx->d_packed.ensure(sizeof(uint32_t) * (size_t)nwords_all));
	SIB_TRY(ctx->d_scalars.ensure(sizeof(uint64_t) * 64));
	SIB_CUDA(cudaMemsetAsync(ctx->d_scalars.p, 0, sizeof(uint64_t) * 64, st));
	uint32_t *d_err = reinterpret_cast<uint32_t*>(ctx->d_scalars.as<uint64_t>() + 8);
	if(pack && w_hi > w_lo)
	{
		ProfScope ps(ctx, "k_pack", (w_hi - w_lo) * 20);
		k_pack<<<grid_for(w_hi - w_lo, 256, ctx->sm_count, 8), 256, 0, st>>>(ctx->d_text.as<uint4>() + w_lo,
			ctx->d_packed.as<uint32_t>() + w_lo, (uint32_t)(w_hi - w_lo), d_err);
	}
	uint64_t nrec = 0;
	for(uint32_t c = 0; c < ctx->nchr; c++)
	{
		if(ctx->h_chr_len[c] >= k) nrec += ctx->h_chr_len[c] - k + 1;
	}
	const uint32_t W = ctx->dist_world;
	uint64_t per_rank = (nrec + W - 1) / W;
	const uint64_t part_rec = ctx->part_records(k);
	uint64_t PL = (per_rank + part_rec - 1) / part_rec;
	if(PL < 1) PL = 1;
	if(PL > MAX_PARTS / W) PL = MAX_PARTS / W;
	ctx->dist_P_local = (uint32_t)PL;
	ctx->dist_P_total = (uint32_t)PL * W;
	ctx->last_k = k;
	TextDesc &t = ctx->dist_text;
	t.packed = ctx->d_packed.as<uint32_t>();
	t.chr_start = ctx->d_chr_start.as<uint32_t>();
	t.chr_len = ctx->d_chr_len.as<uint32_t>();
	t.nchr = ctx->nchr;
	t.M = (uint32_t)ctx->M;
	t.nwords = nwords_all;
	t.tile0 = ctx->dist_tile_lo;
	return SIBGPU_OK;
}

int dist_scan(sibgpu_ctx *ctx, uint32_t k, uint32_t *hist_out)
{
	cudaStream_t st = ctx->stream;
	SIB_TRY(dist_prepare(ctx, k));
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	SIB_CUDA(cudaMemcpyAsync(hs + 8, ctx->d_scalars.as<uint64_t>() + 8, 8, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	if(hs[8] & 1u) return input_error();
	int rc = k <= 28 ? dist_scan_mode<0>(ctx, k, hist_out) : dist_scan_mode<1>(ctx, k, hist_out);
	if(rc != SIBGPU_OK) return rc;
	uint64_t local = 0;
	for(uint32_t p = 0; p < ctx->dist_P_total; p++) local += hist_out[p];
	ctx->dist_nrec_local = local;
	return SIBGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Peer path: no histogram pass and no separate exchange.  Every rank scatters the records of its text range into its
// OWN send buffer, one fixed-capacity segment per global partition (segment p at p * seg_cap records); the ranks then
// swap the per-segment counts (one small all-gather, which is also the barrier), and the owner of partition p reads
// the W segments of p straight out of the W send buffers (CUDA IPC mappings, NVLink) inside its insert kernel.
// ---------------------------------------------------------------------------------------------------------------
// pack (when the own byte range is still on the host: pipelined with its upload) + scatter of the own tile range into PT
// fixed-capacity segments of `cap` records at out_dev, fill cursors at cursor_dev (already initialised to p * cap)
template<int MODE, bool MIXED>
static int dist_scatter_core(sibgpu_ctx *ctx, uint32_t k, uint32_t PT, uint64_t cap, unsigned long long *cursor_dev,
	typename RecT<MODE>::type *out_dev, const HostSrc *src, const FpView fv = FpView{})
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	TextDesc t = ctx->dist_text;
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
	if(!ntiles) return SIBGPU_OK;
	size_t smem = scatter_smem_bytes<MODE, MIXED>();
	SIB_CUDA(cudaFuncSetAttribute(k_scatter<MODE, MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	// src: the own byte range is still on the host -- stream it in pieces of CHUNK_TILES tiles on the copy stream and
	// pack + scatter every piece as it lands (same pipeline as sibgpu_enumerate, enumerate.cu)
	const uint32_t nchunks = src ? (ntiles + CHUNK_TILES - 1) / CHUNK_TILES : 1;
	auto piece_byte = [&](uint32_t c) -> uint64_t {               // first byte of piece c (multiples of 16)
		if(c == 0) return ctx->dist_byte_lo;
		if(c >= nchunks) return ctx->dist_byte_hi;
		return (uint64_t)(ctx->dist_tile_lo + c * CHUNK_TILES) * TILE_POS;
	};
	if(src)
	{
		SIB_TRY(ctx->ensure_copy_stream(nchunks));
		SIB_CUDA(cudaEventRecord(ctx->ev_fork_copy, st));
		SIB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_fork_copy, 0));
		for(uint32_t c = 0; c < nchunks; c++)
		{
			SIB_TRY(copy_text_range(ctx, *src, piece_byte(c), piece_byte(c + 1), ctx->copy_stream));
			SIB_CUDA(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
		}
	}
	uint32_t tiles_done = ctx->dist_tile_lo;                      // absolute tile index
	for(uint32_t c = 0; c < nchunks; c++)
	{
		uint32_t tile_end = ctx->dist_tile_hi;
		if(src)
		{
			const uint64_t w0 = piece_byte(c) / 16, w1 = (piece_byte(c + 1) + 15) / 16;
			SIB_CUDA(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
			if(w1 > w0) SIB_TRY(launch_pack(ctx, w0, w1));
			if(c + 1 < nchunks) tile_end = (uint32_t)((w1 - (TILE_THREADS + 5)) / TILE_THREADS);   // staged words of these tiles are packed
		}
		if(tile_end > tiles_done)
		{
			const uint32_t nt = tile_end - tiles_done;
			const uint32_t occ = MIXED && MODE == 0 ? SIBGPU_SCATTER_OCC : 4;
			const uint32_t g = nt < (uint32_t)ctx->sm_count * occ ? nt : (uint32_t)ctx->sm_count * occ;
			TextDesc tc = t;
			tc.tile0 = tiles_done;
			ProfScope ps(ctx, "k_scatter", (uint64_t)nt * TILE_POS / 4 * (MODE == 2 ? 10 : 1) + (uint64_t)nt * TILE_POS * sizeof(Rec));
			k_scatter<MODE, MIXED><<<g, TILE_THREADS, smem, st>>>(tc, fv, k, nt, PT, cursor_dev, out_dev, cap,
				reinterpret_cast<uint32_t*>(ds + 10));
			tiles_done = tile_end;
		}
	}
	return SIBGPU_OK;
}

template<int MODE>
static int dist_scatter_local_mode(sibgpu_ctx *ctx, uint32_t k, uint64_t *counts_out, uint64_t *seg_cap_out, int *overflow_out,
	const HostSrc *src)
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo, PT = ctx->dist_P_total;
	// at most one record per text position of the own range
	const uint64_t mean = ((uint64_t)ntiles * TILE_POS + PT - 1) / PT;
	const uint64_t cap = (mean + mean / 8 + ctx->part_slack + 31) / 32 * 32;   // segments start 16-byte aligned (TMA)
	ctx->dist_seg_cap = cap;
	for(void *q : ctx->send_retired_old) cudaFree(q);
	ctx->send_retired_old.swap(ctx->send_retired_new);
	ctx->send_retired_new.clear();
	const size_t send_bytes = sizeof(Rec) * cap * PT + 256;
	if(ctx->d_sendbuf.cap < send_bytes && ctx->d_sendbuf.p)
	{
		ctx->send_retired_new.push_back(ctx->d_sendbuf.p);         // peers may still have it mapped
		ctx->d_sendbuf.p = nullptr;
		ctx->d_sendbuf.cap = 0;
	}
	SIB_TRY(ctx->d_sendbuf.ensure(send_bytes));
	SIB_TRY(ctx->d_cursor.ensure(sizeof(uint64_t) * MAX_PARTS * CURSOR_STRIDE));
	std::vector<uint64_t> base(PT), cur((size_t)PT * CURSOR_STRIDE, 0);
	for(uint32_t p = 0; p < PT; p++) cur[(size_t)p * CURSOR_STRIDE] = base[p] = (uint64_t)p * cap;
	SIB_CUDA(cudaMemcpyAsync(ctx->d_cursor.p, cur.data(), sizeof(uint64_t) * cur.size(), cudaMemcpyHostToDevice, st));
	SIB_TRY((dist_scatter_core<MODE, false>(ctx, k, PT, cap, ctx->d_cursor.as<unsigned long long>(), ctx->d_sendbuf.as<Rec>(), src)));
	SIB_CUDA(cudaMemcpyAsync(cur.data(), ctx->d_cursor.p, sizeof(uint64_t) * cur.size(), cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaMemcpyAsync(hs + 8, ds + 8, sizeof(uint64_t) * 3, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	if(hs[8] & 1u) return input_error();
	*overflow_out = (hs[10] & 0xFFFFFFFFull) ? 1 : 0;
	if(*overflow_out) ctx->hist_fallbacks++;
	uint64_t local = 0;
	for(uint32_t p = 0; p < PT; p++)
	{
		counts_out[p] = cur[(size_t)p * CURSOR_STRIDE] - base[p];
		local += counts_out[p];
	}
	ctx->dist_nrec_local = local;
	*seg_cap_out = cap;
	return SIBGPU_OK;
}

int dist_scatter_local(sibgpu_ctx *ctx, uint32_t k, uint64_t *counts_out, uint64_t *seg_cap_out, int *overflow_out, const HostSrc *src)
{
	SIB_TRY(dist_prepare(ctx, k, src == nullptr));
	return k <= 28 ? dist_scatter_local_mode<0>(ctx, k, counts_out, seg_cap_out, overflow_out, src)
		: dist_scatter_local_mode<1>(ctx, k, counts_out, seg_cap_out, overflow_out, src);
}

template<int MODE>
static int dist_group_peer_mode(sibgpu_ctx *ctx, uint32_t k, const uint64_t *counts, const uint64_t *seg_caps, uint64_t *nkeys_local)
{
	typedef typename RecT<MODE>::type Rec;
	cudaStream_t st = ctx->stream;
	const int sms = ctx->sm_count;
	const uint32_t W = ctx->dist_world, PL = ctx->dist_P_local, PT = ctx->dist_P_total, b0 = ctx->dist_rank * PL;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	std::vector<uint64_t> stage_off(PL + 1, 0);
	uint64_t maxpart = 0;
	for(uint32_t p = 0; p < PL; p++)
	{
		uint64_t sum = 0;
		for(uint32_t s = 0; s < W; s++) sum += counts[(size_t)s * PT + b0 + p];
		stage_off[p + 1] = stage_off[p] + sum;
		if(sum > maxpart) maxpart = sum;
	}
	const uint64_t recv_total = stage_off[PL];
	*nkeys_local = 0;
	ctx->dist_nkeys_local = 0;
	if(recv_total == 0) return SIBGPU_OK;
	const uint64_t T64 = (uint64_t)ctx->table_factor * maxpart + 1024;
	if(T64 > 0xFFFFFF00ull)
	{
		set_error("internal: hash partition does not fit a 32-bit table");
		return SIBGPU_ERR_INTERNAL;
	}
	const uint32_t T = (uint32_t)T64;
	const bool compact = MODE == 0 && k <= COMPACT_MAX_K;
	const size_t slot_bytes = compact ? 8 : sizeof(Slot8);
	const uint32_t S = ctx->n_streams < 1 ? 1 : (ctx->n_streams > 8 ? 8 : ctx->n_streams);
	const size_t table_bytes = (slot_bytes * T + 255) / 256 * 256;
	SIB_TRY(ctx->d_table.ensure(table_bytes * S));
	SIB_CUDA(cudaMemsetAsync(ctx->d_table.p, 0xFF, table_bytes * S, st));
	// the vertex keys of partition p are staged at stage_off[p] (a partition has at most as many classes as records);
	// the send buffer cannot serve as staging here: the peers are still reading it
	SIB_TRY(ctx->d_keystage.ensure(sizeof(Rec) * recv_total));
	SIB_TRY(ctx->d_partcnt.ensure(sizeof(uint32_t) * MAX_PARTS));
	SIB_TRY(ctx->d_keyoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
	SIB_TRY(ctx->d_partoff.ensure(sizeof(uint64_t) * (MAX_PARTS + 1)));
	SIB_CUDA(cudaMemsetAsync(ctx->d_partcnt.p, 0, sizeof(uint32_t) * MAX_PARTS, st));
	SIB_CUDA(cudaMemcpyAsync(ctx->d_partoff.p, stage_off.data(), sizeof(uint64_t) * (PL + 1), cudaMemcpyHostToDevice, st));
	if(S > 1)
	{
		SIB_TRY(ctx->ensure_aux_streams(S));
		SIB_CUDA(cudaEventRecord(ctx->ev_fork, st));
		for(uint32_t i = 0; i < S; i++) SIB_CUDA(cudaStreamWaitEvent(ctx->aux_stream[i], ctx->ev_fork, 0));
	}
	const int blocks_per_sm = S > 1 ? 4 : 8;
	const bool phase_span = ctx->profiling && S > 1;
	// Measured on 2 B200s (100 M records per rank, half of them remote; profiles/r1_sharded_peer_read.txt): plain loads
	// from the peer 3.41 ms, remote segments first pulled into a local buffer by the copy engines 2.20 ms, TMA ring
	// reading the peer directly 2.21 ms at 8 CTAs per SM -- the ring hides the NVLink latency without a staging copy.
	static const int seg_ctas_per_sm = getenv("SIBGPU_SEG_CTAS") ? atoi(getenv("SIBGPU_SEG_CTAS")) : 8;
	static const int seg_kernel = getenv("SIBGPU_SEG_KERNEL") ? atoi(getenv("SIBGPU_SEG_KERNEL")) : 1;   // 0 = loads, 1 = TMA ring
	const uint64_t tile_rec = SEG_TILE_BYTES / sizeof(Rec);
	if(phase_span) ctx->prof_begin("k_insert_seg+k_table_scan", recv_total * sizeof(Rec));
	for(uint32_t p = 0; p < PL; p++)
	{
		const uint64_t n = stage_off[p + 1] - stage_off[p];
		if(n == 0) continue;
		SegList segs;
		uint32_t nseg = 0, seg_tiles = 0;
		for(uint32_t s = 0; s < W; s++)
		{
			const uint64_t c = counts[(size_t)s * PT + b0 + p];
			if(c == 0) continue;
			if(s == ctx->dist_rank) segs.ptr[nseg] = ctx->d_sendbuf.as<Rec>() + (uint64_t)(b0 + p) * seg_caps[s];
			else segs.ptr[nseg] = static_cast<const Rec*>(ctx->peer_ptr[s]) + (uint64_t)(b0 + p) * seg_caps[s];
			segs.cnt[nseg] = (uint32_t)c;
			seg_tiles += (uint32_t)((c + tile_rec - 1) / tile_rec);
			nseg++;
		}
		for(uint32_t i = nseg; i < MAX_PEERS; i++) { segs.ptr[i] = nullptr; segs.cnt[i] = 0; }
		cudaStream_t ps_st = S > 1 ? ctx->aux_stream[p % S] : st;
		void *table = static_cast<char*>(ctx->d_table.p) + table_bytes * (p % S);
		Rec *out = ctx->d_keystage.as<Rec>() + stage_off[p];
		ctx->total_launches += 2;
		const uint32_t seg_ctas = (uint32_t)sms * seg_ctas_per_sm;
		const uint32_t g = seg_tiles < seg_ctas ? seg_tiles : seg_ctas;
		if(S == 1 && ctx->profiling) ctx->prof_begin("k_insert_seg", n * sizeof(Rec));
		if(seg_kernel == 0)
		{
			const uint32_t gl = grid_for(n, 256, sms, blocks_per_sm);
			if(compact) k_insert_seg_ld<MODE, true, 1><<<gl, 256, 0, ps_st>>>(segs, nseg, n, table, T);
			else k_insert_seg_ld<MODE, false, 1><<<gl, 256, 0, ps_st>>>(segs, nseg, n, table, T);
		}
		else if(compact) k_insert_seg<MODE, true><<<g, 256, 0, ps_st>>>(segs, nseg, seg_tiles, table, T);
		else k_insert_seg<MODE, false><<<g, 256, 0, ps_st>>>(segs, nseg, seg_tiles, table, T);
		if(S == 1 && ctx->profiling) { ctx->prof_end(); ctx->prof_begin("k_table_scan", (uint64_t)T * slot_bytes); }
		if(compact) k_table_scan_compact<<<grid_for(T, 256, sms, blocks_per_sm), 256, 0, ps_st>>>(
			static_cast<unsigned long long*>(table), T, reinterpret_cast<uint64_t*>(out), ctx->d_partcnt.as<uint32_t>() + p);
		else k_table_scan<MODE><<<grid_for(T, 256, sms, blocks_per_sm), 256, 0, ps_st>>>(table, T, out, ctx->d_partcnt.as<uint32_t>() + p);
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
	k_key_offsets<<<1, MAX_PARTS, 0, st>>>(ctx->d_partcnt.as<uint32_t>(), PL, ctx->d_keyoff.as<uint64_t>(), ds);
	ctx->total_launches++;
	SIB_CUDA(cudaMemcpyAsync(hs + 2, ds + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	const uint64_t Vc = hs[2];
	if(Vc)
	{
		SIB_TRY(ctx->d_ckeys.ensure(sizeof(Rec) * Vc));
		ProfScope ps(ctx, "k_gather_keys", 2 * Vc * sizeof(Rec));
		dim3 g(8, PL);
		k_gather_keys<MODE><<<g, 256, 0, st>>>(ctx->d_keystage.as<Rec>(), ctx->d_partoff.as<uint64_t>(),
			ctx->d_partcnt.as<uint32_t>(), ctx->d_keyoff.as<uint64_t>(), ctx->d_ckeys.as<Rec>());
		SIB_CUDA(cudaStreamSynchronize(st));
	}
	ctx->dist_nkeys_local = Vc;
	*nkeys_local = Vc;
	return SIBGPU_OK;
}

int dist_group_peer(sibgpu_ctx *ctx, const uint64_t *counts, const uint64_t *seg_caps, uint64_t *nkeys_local)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	if(ctx->dist_world > (uint32_t)MAX_PEERS)
	{
		set_error("invalid: the peer path supports at most 16 ranks");
		return SIBGPU_ERR_INVALID;
	}
	for(uint32_t s = 0; s < ctx->dist_world; s++)
	{
		if(s != ctx->dist_rank && (ctx->peer_ptr.size() <= s || !ctx->peer_ptr[s]))
		{
			set_error("state: sibgpu_dist_import_peers must precede sibgpu_dist_group_peer");
			return SIBGPU_ERR_STATE;
		}
	}
	return ctx->last_k <= 28 ? dist_group_peer_mode<0>(ctx, ctx->last_k, counts, seg_caps, nkeys_local)
		: dist_group_peer_mode<1>(ctx, ctx->last_k, counts, seg_caps, nkeys_local);
}

int dist_scatter(sibgpu_ctx *ctx, void *send_dev)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	return ctx->last_k <= 28 ? dist_scatter_mode<0>(ctx, ctx->last_k, send_dev) : dist_scatter_mode<1>(ctx, ctx->last_k, send_dev);
}

int dist_group(sibgpu_ctx *ctx, const void *recv_dev, const uint32_t *counts, uint64_t *nkeys_local)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	return ctx->last_k <= 28 ? dist_group_mode<0>(ctx, ctx->last_k, recv_dev, counts, nkeys_local)
		: dist_group_mode<1>(ctx, ctx->last_k, recv_dev, counts, nkeys_local);
}

int dist_finish(sibgpu_ctx *ctx, const void *allkeys_dev, uint64_t nkeys_total)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	int rc = ctx->last_k <= 28 ? dist_finish_mode<0>(ctx, ctx->last_k, allkeys_dev, nkeys_total)
		: dist_finish_mode<1>(ctx, ctx->last_k, allkeys_dev, nkeys_total);
	if(rc != SIBGPU_OK) return rc;
	SIB_CUDA(cudaGetLastError());
	ctx->have_result = true;
	ctx->dist_result = true;
	if(ctx->profiling) SIB_TRY(ctx->prof_collect());
	return SIBGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused path (k <= 28): no collective and no host round trip inside a step.  Every rank owns ONE exported buffer
//     [DistHeader | fill cursors | vertex keys of the owned partitions | PT fixed-capacity segments]
// mapped by all peers (CUDA IPC, once).  A step e:
//   scatter   own text range -> own segments (mixed 8-byte records), then header.epoch_scatter = e
//   k_split   for every owned partition and every source rank: wait for the source's epoch_scatter == e (device-side
//             spin on the mapped header), read its fill cursor and bulk-copy (TMA) its segment tiles straight out of
//             the peer's memory over NVLink -> local ~1 Ki-record buckets          [the all-to-all, fused]
//   k_group   buckets -> vertex keys of the owned partitions in the own key region, then header.epoch_keys = e
//   k_pull    wait for every rank's epoch_keys == e, concatenate all key regions    [the all-gather, fused]
//   ids + instance tables of the own text range (ids_and_tables)
// Reuse of a buffer for step e+1 is safe without any barrier: a rank starts e+1 after it has pulled every peer's keys
// of step e, which a peer publishes only after its k_split of step e has consumed all segments; and a peer's pull of
// step e precedes its scatter flag of step e+1, which this rank's k_split(e+1) awaits before k_group(e+1) rewrites keys.
// Failures (segment / bucket / key-list overflow, illegal character) travel in the headers, so all ranks take the
// same decision: status 1 = use the phased path, 2 = regrow the key regions (collectively) and run again.
// ---------------------------------------------------------------------------------------------------------------
constexpr size_t XOFF_CURSOR = sizeof(DistHeader);
constexpr size_t XOFF_KEYS = XOFF_CURSOR + sizeof(uint64_t) * MAX_PARTS * CURSOR_STRIDE;
static_assert(sizeof(DistHeader) == 256, "header layout");

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void k_init_cursors(unsigned long long *cursor, uint32_t PT, unsigned long long cap)
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if(p < PT) cursor[(size_t)p * CURSOR_STRIDE] = (unsigned long long)p * cap;
}

// scalars[8] bit 0 = illegal character, scalars[10] = a segment overflowed
__global__ void k_publish_scatter(DistHeader *h, const uint64_t *scalars, unsigned long long epoch)
{
	h->scatter_flags = ((scalars[8] & 1u) ? 2u : 0u) | ((scalars[10] & 0xFFFFFFFFull) ? 1u : 0u);
	__threadfence_system();
	st_release_sys(&h->epoch_scatter, epoch);
}

__global__ void k_publish_keys(DistHeader *h, const uint32_t *nkeys, const uint32_t *grp_flags, uint32_t key_cap, unsigned long long epoch)
{
	h->nkeys = *nkeys;
	h->key_flags = *grp_flags | (*nkeys > key_cap ? 16u : 0u);
	__threadfence_system();
	st_release_sys(&h->epoch_keys, epoch);
}

struct PullSrc { const DistHeader *header[SPLIT_MAX_SRC]; const unsigned long long *keys[SPLIT_MAX_SRC]; };
// out[0] = total keys, out[1] = OR of all ranks' failure flags (bit 1 of the scatter flags -> 64: illegal character),
// out[2] = largest per-rank key count
// WPK = 64-bit words per key (2: fingerprint classes {fingerprint, partition})
template<int WPK>
__global__ void __launch_bounds__(256) k_pull_keys(const PullSrc src, uint32_t W, unsigned long long epoch, uint32_t key_cap,
	uint64_t *__restrict__ allkeys, uint64_t allcap, uint64_t *__restrict__ out)
{
	__shared__ unsigned long long s_off, s_n, s_total, s_flags, s_max;
	const uint32_t me = blockIdx.y;
	if(threadIdx.x == 0)
	{
		unsigned long long off = 0, total = 0, flags = 0, mx = 0, mine = 0;
		for(uint32_t r = 0; r < W; r++)
		{
			if(!wait_epoch(&src.header[r]->epoch_keys, epoch)) { flags |= GRP_TIMEOUT; break; }
			const unsigned long long n = ld_relaxed_sys(&src.header[r]->nkeys);
			flags |= ld_relaxed_sys(&src.header[r]->key_flags);
			const unsigned long long sf = ld_relaxed_sys(&src.header[r]->scatter_flags);
			flags |= (sf & 1u ? GRP_PEER_FAILED : 0u) | (sf & 2u ? 64u : 0u);
			if(r < me) off += n;
			if(r == me) mine = n;
			total += n;
			if(n > mx) mx = n;
		}
		if(total > allcap) flags |= 32u;
		s_off = off; s_n = mine; s_total = total; s_flags = flags; s_max = mx;
	}
	__syncthreads();
	if(me == 0 && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = s_total; out[1] = s_flags; out[2] = s_max; }
	if(s_flags) return;
	// 16-byte loads over NVLink (the key regions are 16-byte aligned), the odd last key on its own
	const unsigned long long *k = src.keys[me];
	const uint64_t nw = s_n * WPK, ow = s_off * WPK, pairs = nw / 2;
	for(uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < pairs; i += (uint64_t)gridDim.x * blockDim.x)
	{
		unsigned long long a, b;
		asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(k + 2 * i) : "memory");
		allkeys[ow + 2 * i] = a;
		allkeys[ow + 2 * i + 1] = b;
	}
	if((nw & 1u) && blockIdx.x == 0 && threadIdx.x == 0) allkeys[ow + nw - 1] = ld_relaxed_sys(k + nw - 1);
}

int dist2_plan(sibgpu_ctx *ctx, uint32_t k, int *need_alloc)
{
	*need_alloc = -1;
	const uint32_t W = ctx->dist_world;
	if((k > 28 && k <= 32) || W > (uint32_t)SPLIT_MAX_SRC || !ctx->group_smem) return SIBGPU_OK;
	const bool fp = k > 32;                                // fingerprint classes: 16-byte keys, packed text replicated by peer pulls
	uint64_t nrec = 0;
	for(uint32_t c = 0; c < ctx->nchr; c++)
	{
		if(ctx->h_chr_len[c] >= k) nrec += ctx->h_chr_len[c] - k + 1;
	}
	const uint64_t part_rec = ctx->part_explicit ? ctx->part_target : (uint64_t)512 << 10;
	uint64_t PL = ((nrec + W - 1) / W + part_rec - 1) / part_rec;
	if(PL < 1) PL = 1;
	if(PL > MAX_PARTS / W) PL = MAX_PARTS / W;
	const uint32_t PT = (uint32_t)PL * W;
	const uint64_t mean1 = (nrec + PT - 1) / PT;
	uint32_t sub_bits = 0;
	while(((mean1 + ((uint64_t)1 << sub_bits) - 1) >> sub_bits) > GROUP_MEAN) sub_bits++;
	if(sub_bits > SUB_BITS_MAX) return SIBGPU_OK;
	const uint64_t ntiles = (ctx->M + TILE_POS - 1) / TILE_POS;
	const uint64_t tiles_max = ntiles / W + 1;             // no rank scans more tiles than this
	const uint64_t mean_seg = (tiles_max * TILE_POS + PT - 1) / PT;
	const uint64_t seg_cap = (mean_seg + mean_seg / 8 + ctx->part_slack + 31) / 32 * 32;
	const uint64_t kc = std::max<uint64_t>(ctx->x_kc, std::max<uint64_t>(ctx->ckeys_init, 16));
	const uint64_t off_seg = (XOFF_KEYS + kc * (fp ? 16 : 8) + 255) / 256 * 256;
	const uint64_t off_pk = fp ? (off_seg + (uint64_t)PT * seg_cap * 8 + 255) / 256 * 256 : 0;
	const uint64_t nwords_all = (ctx->M + 15) / 16 + 8;
	const uint64_t bytes = (fp ? off_pk + nwords_all * 4 : off_seg + (uint64_t)PT * seg_cap * 8) + 256;
	ctx->x_PL = (uint32_t)PL;
	ctx->x_sub_bits = sub_bits;
	ctx->x_seg_cap = seg_cap;
	ctx->x_nrec = nrec;
	ctx->x_k = k;
	ctx->x_bytes = bytes;
	// the layout of a live buffer must not move (the peers computed their addresses from it): new key capacity or
	// segment offset means a new buffer
	*need_alloc = (ctx->d_xbuf.cap < bytes || ctx->x_off_seg != off_seg || ctx->x_off_pk != off_pk || ctx->x_kc_live != kc ||
		ctx->peer_x.size() != W) ? 1 : 0;
	ctx->x_kc = kc;
	ctx->x_off_seg_plan = off_seg;
	ctx->x_off_pk_plan = off_pk;
	return SIBGPU_OK;
}

int dist2_release_peers(sibgpu_ctx *ctx)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	SIB_CUDA(cudaStreamSynchronize(ctx->stream));
	for(void *&pp : ctx->peer_x)
	{
		if(pp) cudaIpcCloseMemHandle(pp);
		pp = nullptr;
	}
	ctx->peer_x.clear();
	return SIBGPU_OK;
}

int dist2_alloc(sibgpu_ctx *ctx, void *handle64)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	if(!ctx->peer_x.empty())
	{
		set_error("state: sibgpu_fused_release_peers must precede sibgpu_fused_alloc");
		return SIBGPU_ERR_STATE;
	}
	ctx->d_xbuf.release();                                 // every peer has closed its mapping (collective protocol)
	SIB_TRY(ctx->d_xbuf.ensure(ctx->x_bytes));
	SIB_CUDA(cudaMemsetAsync(ctx->d_xbuf.p, 0, XOFF_KEYS, ctx->stream));
	SIB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->x_off_seg = ctx->x_off_seg_plan;
	ctx->x_off_pk = ctx->x_off_pk_plan;
	ctx->x_kc_live = ctx->x_kc;
	cudaIpcMemHandle_t h;
	SIB_CUDA(cudaIpcGetMemHandle(&h, ctx->d_xbuf.p));
	memcpy(handle64, &h, 64);
	return SIBGPU_OK;
}

int dist2_import(sibgpu_ctx *ctx, const void *handles)
{
	SIB_CUDA(cudaSetDevice(ctx->device));
	const uint32_t W = ctx->dist_world;
	ctx->peer_x.assign(W, nullptr);
	for(uint32_t s = 0; s < W; s++)
	{
		if(s == ctx->dist_rank) continue;
		cudaIpcMemHandle_t ih;
		memcpy(&ih, static_cast<const unsigned char*>(handles) + 64 * (size_t)s, 64);
		SIB_CUDA(cudaIpcOpenMemHandle(&ctx->peer_x[s], ih, cudaIpcMemLazyEnablePeerAccess));
	}
	return SIBGPU_OK;
}

int dist2_run(sibgpu_ctx *ctx, const HostSrc *src, int *status)
{
	NvtxRange nvtx("sibgpu: fused sharded step");
	cudaStream_t st = ctx->stream;
	const uint32_t W = ctx->dist_world, rank = ctx->dist_rank, k = ctx->x_k, PL = ctx->x_PL, PT = PL * W;
	*status = 0;
	if(ctx->peer_x.size() != W || !ctx->d_xbuf.p)
	{
		set_error("state: sibgpu_fused_alloc / sibgpu_fused_import must precede sibgpu_fused_run");
		return SIBGPU_ERR_STATE;
	}
	const unsigned long long epoch = ++ctx->dist_epoch;
	SIB_TRY(dist_prepare(ctx, k, src == nullptr));
	ctx->dist_P_local = PL;
	ctx->dist_P_total = PT;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	unsigned char *xb = static_cast<unsigned char*>(ctx->d_xbuf.p);
	DistHeader *hdr = reinterpret_cast<DistHeader*>(xb);
	unsigned long long *cursor = reinterpret_cast<unsigned long long*>(xb + XOFF_CURSOR);
	uint64_t *keys = reinterpret_cast<uint64_t*>(xb + XOFF_KEYS);
	uint64_t *segs = reinterpret_cast<uint64_t*>(xb + ctx->x_off_seg);
	const uint32_t key_cap = (uint32_t)std::min<uint64_t>(ctx->x_kc_live, 0xFFFFFFF0u);
	// ---- scatter + publish
	k_init_cursors<<<(PT + 255) / 256, 256, 0, st>>>(cursor, PT, ctx->x_seg_cap);
	ctx->total_launches++;
	SIB_TRY((dist_scatter_core<0, true>(ctx, k, PT, ctx->x_seg_cap, cursor, segs, src)));
	k_publish_scatter<<<1, 1, 0, st>>>(hdr, ds, epoch);
	// ---- fused exchange + split, group, publish
	const uint32_t sub_bits = ctx->x_sub_bits, nbuckets = PL << sub_bits;
	SIB_TRY(ctx->d_records2.ensure(sizeof(uint64_t) * (size_t)nbuckets * GROUP_CAP + 64));
	SIB_TRY(ctx->d_cnt2.ensure(sizeof(uint32_t) * (size_t)nbuckets));
	SIB_CUDA(cudaMemsetAsync(ctx->d_cnt2.p, 0, sizeof(uint32_t) * (size_t)nbuckets, st));
	SplitSrc ssrc = {};
	PullSrc psrc = {};
	for(uint32_t s = 0; s < W; s++)
	{
		const unsigned char *b = s == rank ? xb : static_cast<const unsigned char*>(ctx->peer_x[s]);
		ssrc.seg[s] = reinterpret_cast<const uint64_t*>(b + ctx->x_off_seg);
		ssrc.cursor[s] = reinterpret_cast<const unsigned long long*>(b + XOFF_CURSOR);
		ssrc.header[s] = s == rank ? nullptr : reinterpret_cast<const DistHeader*>(b);
		psrc.header[s] = reinterpret_cast<const DistHeader*>(b);
		psrc.keys[s] = reinterpret_cast<const unsigned long long*>(b + XOFF_KEYS);
	}
	ssrc.seg_cap = ctx->x_seg_cap;
	ssrc.epoch = epoch;
	ssrc.W = W;
	ssrc.p0 = rank * PL;
	ssrc.rot = rank;
	const uint32_t tiles_per_seg = (uint32_t)((ctx->x_seg_cap + RecOps<uint64_t>::TILE - 1) / RecOps<uint64_t>::TILE);
	uint32_t *d_flags = reinterpret_cast<uint32_t*>(ds + 11);
	SIB_TRY(launch_split<uint64_t>(ctx, ssrc, PL, tiles_per_seg, sub_bits, ctx->x_nrec / W, d_flags));
	SIB_TRY(launch_group<uint64_t>(ctx, nbuckets, sub_bits, ctx->x_nrec / W, d_flags, keys, key_cap, reinterpret_cast<uint32_t*>(ds + 2)));
	k_publish_keys<<<1, 1, 0, st>>>(hdr, reinterpret_cast<uint32_t*>(ds + 2), d_flags, key_cap, epoch);
	// ---- all ranks' keys
	SIB_TRY(ctx->d_ckeys.ensure(sizeof(uint64_t) * std::max<uint64_t>(ctx->ckeys_init, 16)));
	for(int attempt = 0; ; attempt++)
	{
		const uint64_t allcap = ctx->d_ckeys.cap / sizeof(uint64_t);
		{
			ProfScope ps(ctx, "k_pull_keys", 0);
			k_pull_keys<1><<<dim3(32, W), 256, 0, st>>>(psrc, W, epoch, key_cap, ctx->d_ckeys.as<uint64_t>(), allcap, ds + 12);
		}
		ctx->total_launches += 2;
		SIB_CUDA(cudaMemcpyAsync(hs + 8, ds + 8, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		const uint64_t flags = hs[13];
		if(flags & GRP_TIMEOUT)
		{
			set_error("internal: a peer rank did not publish its step within 4 s");
			return SIBGPU_ERR_INTERNAL;
		}
		if((hs[8] & 1u) || (flags & 64u)) return input_error();
		if(flags & 16u)                                        // a rank's key region is too small: regrow collectively
		{
			ctx->x_kc = hs[14] + hs[14] / 4 + 1024;
			*status = 2;
			return SIBGPU_OK;
		}
		if(flags & (GRP_BUCKET_OVERFLOW | GRP_PEER_FAILED))
		{
			if(hs[10] & 0xFFFFFFFFull) ctx->hist_fallbacks++;
			if(hs[11] & GRP_BUCKET_OVERFLOW) ctx->smem_fallbacks++;
			*status = 1;
			return SIBGPU_OK;
		}
		if(flags & 32u)                                        // the local list of all keys is too small: regrow, pull again
		{
			if(attempt) { set_error("internal: key list regrow failed"); return SIBGPU_ERR_INTERNAL; }
			SIB_TRY(ctx->d_ckeys.ensure(sizeof(uint64_t) * hs[12]));
			continue;
		}
		break;
	}
	const uint64_t Vc = hs[12];
	const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
	ctx->n_inst = 0;
	ctx->n_vertices = 0;
	if(Vc)
	{
		if(2 * Vc > 0xFFFFFFF0ull)
		{
			set_error("invalid: more than 2^32 vertices");
			return SIBGPU_ERR_INVALID;
		}
		bool collision = false;
		SIB_TRY(ids_and_tables<0>(ctx, ctx->dist_text, k, ctx->d_ckeys.as<uint64_t>(), Vc, ntiles, FpView{}, 0u, false, &collision));
	}
	SIB_CUDA(cudaGetLastError());
	ctx->have_result = true;
	ctx->dist_result = true;
	if(ctx->profiling) SIB_TRY(ctx->prof_collect());
	return SIBGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Fused path, k > 32 (fingerprint classes).  Same exchange as above on 8-byte fingerprint records; what is new:
//   * the rolling hashes of a position read the text up to k bases ahead, the string ranking of the vertex classes and
//     the verification of every instance read it at arbitrary positions: the PACKED text (2 bits per base) is made
//     complete on every rank first -- each rank packs its own range, copies it into its exported buffer, publishes
//     epoch_packed, and k_pull_packed copies the peers' ranges over NVLink (the all-gather, fused)
//   * a class is {fingerprint, GLOBAL partition}: 16-byte entries in the key regions
//   * the representative of a class is its smallest text position over ALL ranks: the step stops after k_mark
//     (dist2_run_fp), the caller min-reduces the representatives (one all-reduce), dist2_finish_fp ranks the strings
//     (every rank ranks all classes: same ids everywhere) and emits + verifies the instances of the own range.
// A verification failure anywhere (the caller max-reduces the flag) repeats the step with other hash bases.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_publish_packed(DistHeader *h, unsigned long long epoch)
{
	__threadfence_system();
	st_release_sys(&h->epoch_packed, epoch);
}

struct PackedSrc {
	const DistHeader *header[SPLIT_MAX_SRC];
	const uint32_t *pk[SPLIT_MAX_SRC];                 // exported packed region of rank s (word w at pk[s] + w)
	uint32_t w_lo[SPLIT_MAX_SRC], w_hi[SPLIT_MAX_SRC]; // the words rank s owns (w_lo is a multiple of 256)
};
// blockIdx.y = source rank
__global__ void __launch_bounds__(256) k_pull_packed(const PackedSrc src, uint32_t me, unsigned long long epoch,
	uint32_t *__restrict__ packed, uint32_t *__restrict__ flags)
{
	const uint32_t s = blockIdx.y;
	if(s == me || src.w_hi[s] <= src.w_lo[s]) return;
	__shared__ uint32_t ok;
	if(threadIdx.x == 0)
	{
		ok = wait_epoch(&src.header[s]->epoch_packed, epoch) ? 1u : 0u;
		if(!ok) atomicOr(flags, GRP_TIMEOUT);
	}
	__syncthreads();
	if(!ok) return;
	const uint32_t w_lo = src.w_lo[s], n = src.w_hi[s] - w_lo;
	const unsigned long long *from = reinterpret_cast<const unsigned long long*>(src.pk[s] + w_lo);
	unsigned long long *to = reinterpret_cast<unsigned long long*>(packed + w_lo);
	const uint32_t quads = n / 4;
	for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += gridDim.x * blockDim.x)
	{
		unsigned long long a, b;
		asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(from + 2 * i) : "memory");
		to[2 * i] = a;
		to[2 * i + 1] = b;
	}
	if(blockIdx.x == 0 && threadIdx.x < (n & 3u))
	{
		const uint32_t w = w_lo + quads * 4 + threadIdx.x;
		uint32_t v;
		asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src.pk[s] + w) : "memory");
		packed[w] = v;
	}
}

int dist2_run_fp(sibgpu_ctx *ctx, const HostSrc *src, uint32_t attempt, int *status)
{
	NvtxRange nvtx("sibgpu: fused sharded step, k > 32 (front)");
	cudaStream_t st = ctx->stream;
	const uint32_t W = ctx->dist_world, rank = ctx->dist_rank, k = ctx->x_k, PL = ctx->x_PL, PT = PL * W;
	*status = 0;
	if(ctx->peer_x.size() != W || !ctx->d_xbuf.p || !ctx->x_off_pk || k <= 32)
	{
		set_error("state: sibgpu_fused_plan (k > 32) / sibgpu_fused_alloc / sibgpu_fused_import must precede sibgpu_fused_run_fp");
		return SIBGPU_ERR_STATE;
	}
	const unsigned long long epoch = ++ctx->dist_epoch;
	SIB_TRY(dist_prepare(ctx, k, src == nullptr));
	ctx->dist_P_local = PL;
	ctx->dist_P_total = PT;
	ctx->x_attempt = attempt;
	ctx->x_Vc = 0;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	unsigned char *xb = static_cast<unsigned char*>(ctx->d_xbuf.p);
	DistHeader *hdr = reinterpret_cast<DistHeader*>(xb);
	unsigned long long *cursor = reinterpret_cast<unsigned long long*>(xb + XOFF_CURSOR);
	Rec16 *keys = reinterpret_cast<Rec16*>(xb + XOFF_KEYS);
	uint64_t *segs = reinterpret_cast<uint64_t*>(xb + ctx->x_off_seg);
	uint32_t *xpk = reinterpret_cast<uint32_t*>(xb + ctx->x_off_pk);
	const uint32_t key_cap = (uint32_t)std::min<uint64_t>(ctx->x_kc_live, 0xFFFFFFF0u);
	uint32_t *d_flags = reinterpret_cast<uint32_t*>(ds + 11);
	// ---- the packed text: own range (uploaded here unless resident), exported, the peers' ranges pulled
	if(src && ctx->dist_byte_hi > ctx->dist_byte_lo)
	{
		SIB_TRY(copy_text_range(ctx, *src, ctx->dist_byte_lo, ctx->dist_byte_hi, st));
		SIB_TRY(launch_pack(ctx, ctx->dist_byte_lo / 16, (ctx->dist_byte_hi + 15) / 16));
	}
	const uint32_t nwords_all = ctx->dist_text.nwords;
	const uint64_t ntiles_all = (ctx->M + TILE_POS - 1) / TILE_POS;
	PackedSrc ksrc = {};
	for(uint32_t s = 0; s < W; s++)
	{
		const uint64_t tlo = ntiles_all * s / W, thi = ntiles_all * (s + 1) / W;
		uint64_t lo = tlo * TILE_THREADS, hi = s + 1 == W ? nwords_all : thi * TILE_THREADS;
		if(hi > nwords_all) hi = nwords_all;
		if(lo > hi) lo = hi;
		ksrc.w_lo[s] = (uint32_t)lo;
		ksrc.w_hi[s] = (uint32_t)hi;
		const unsigned char *b = s == rank ? xb : static_cast<const unsigned char*>(ctx->peer_x[s]);
		ksrc.header[s] = reinterpret_cast<const DistHeader*>(b);
		ksrc.pk[s] = reinterpret_cast<const uint32_t*>(b + ctx->x_off_pk);
	}
	if(ksrc.w_hi[rank] > ksrc.w_lo[rank])
	{
		SIB_CUDA(cudaMemcpyAsync(xpk + ksrc.w_lo[rank], ctx->d_packed.as<uint32_t>() + ksrc.w_lo[rank],
			sizeof(uint32_t) * (size_t)(ksrc.w_hi[rank] - ksrc.w_lo[rank]), cudaMemcpyDeviceToDevice, st));
	}
	k_publish_packed<<<1, 1, 0, st>>>(hdr, epoch);
	if(W > 1)
	{
		ProfScope ps(ctx, "k_pull_packed", (uint64_t)nwords_all * 4);
		k_pull_packed<<<dim3(64, W), 256, 0, st>>>(ksrc, rank, epoch, ctx->d_packed.as<uint32_t>(), d_flags);
	}
	ctx->total_launches += 2;
	// ---- checkpoints of the own words, scatter + publish
	const uint32_t w_ck_stop = (uint32_t)std::min<uint64_t>(ksrc.w_hi[rank], (ctx->M + 15) >> 4);
	SIB_TRY(fingerprint_positions(ctx, ctx->dist_text, k, attempt, ksrc.w_lo[rank], w_ck_stop));
	const FpView fv = {ctx->d_fp.as<FpCk>(), ctx->d_fpprm.as<FpParams>()};
	k_init_cursors<<<(PT + 255) / 256, 256, 0, st>>>(cursor, PT, ctx->x_seg_cap);
	ctx->total_launches++;
	SIB_TRY((dist_scatter_core<2, true>(ctx, k, PT, ctx->x_seg_cap, cursor, segs, nullptr, fv)));
	k_publish_scatter<<<1, 1, 0, st>>>(hdr, ds, epoch);
	// ---- fused exchange + split, group, publish
	const uint32_t sub_bits = ctx->x_sub_bits, nbuckets = PL << sub_bits;
	SIB_TRY(ctx->d_records2.ensure(sizeof(uint64_t) * (size_t)nbuckets * GROUP_CAP + 64));
	SIB_TRY(ctx->d_cnt2.ensure(sizeof(uint32_t) * (size_t)nbuckets));
	SIB_CUDA(cudaMemsetAsync(ctx->d_cnt2.p, 0, sizeof(uint32_t) * (size_t)nbuckets, st));
	SplitSrc ssrc = {};
	PullSrc psrc = {};
	for(uint32_t s = 0; s < W; s++)
	{
		const unsigned char *b = s == rank ? xb : static_cast<const unsigned char*>(ctx->peer_x[s]);
		ssrc.seg[s] = reinterpret_cast<const uint64_t*>(b + ctx->x_off_seg);
		ssrc.cursor[s] = reinterpret_cast<const unsigned long long*>(b + XOFF_CURSOR);
		ssrc.header[s] = s == rank ? nullptr : reinterpret_cast<const DistHeader*>(b);
		psrc.header[s] = reinterpret_cast<const DistHeader*>(b);
		psrc.keys[s] = reinterpret_cast<const unsigned long long*>(b + XOFF_KEYS);
	}
	ssrc.seg_cap = ctx->x_seg_cap;
	ssrc.epoch = epoch;
	ssrc.W = W;
	ssrc.p0 = rank * PL;
	ssrc.rot = rank;
	const uint32_t tiles_per_seg = (uint32_t)((ctx->x_seg_cap + RecOps<uint64_t>::TILE - 1) / RecOps<uint64_t>::TILE);
	SIB_TRY(launch_split<uint64_t>(ctx, ssrc, PL, tiles_per_seg, sub_bits, ctx->x_nrec / W, d_flags));
	SIB_TRY((launch_group<uint64_t, true>(ctx, nbuckets, sub_bits, ctx->x_nrec / W, d_flags, keys, key_cap,
		reinterpret_cast<uint32_t*>(ds + 2), rank * PL)));
	k_publish_keys<<<1, 1, 0, st>>>(hdr, reinterpret_cast<uint32_t*>(ds + 2), d_flags, key_cap, epoch);
	// ---- all ranks' keys
	SIB_TRY(ctx->d_ckeys.ensure(sizeof(Rec16) * std::max<uint64_t>(ctx->ckeys_init, 16)));
	for(int pass = 0; ; pass++)
	{
		const uint64_t allcap = ctx->d_ckeys.cap / sizeof(Rec16);
		{
			ProfScope ps(ctx, "k_pull_keys", 0);
			k_pull_keys<2><<<dim3(32, W), 256, 0, st>>>(psrc, W, epoch, key_cap, ctx->d_ckeys.as<uint64_t>(), allcap, ds + 12);
		}
		ctx->total_launches += 2;
		SIB_CUDA(cudaMemcpyAsync(hs + 8, ds + 8, sizeof(uint64_t) * 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		const uint64_t flags = hs[13] | (hs[11] & GRP_TIMEOUT);
		if(flags & GRP_TIMEOUT)
		{
			set_error("internal: a peer rank did not publish its step within 4 s");
			return SIBGPU_ERR_INTERNAL;
		}
		if((hs[8] & 1u) || (flags & 64u)) return input_error();
		if(flags & 16u)                                        // a rank's key region is too small: regrow collectively
		{
			ctx->x_kc = hs[14] + hs[14] / 4 + 1024;
			*status = 2;
			return SIBGPU_OK;
		}
		if(flags & (GRP_BUCKET_OVERFLOW | GRP_PEER_FAILED))
		{
			if(hs[10] & 0xFFFFFFFFull) ctx->hist_fallbacks++;
			if(hs[11] & GRP_BUCKET_OVERFLOW) ctx->smem_fallbacks++;
			*status = 1;
			return SIBGPU_OK;
		}
		if(flags & 32u)                                        // the local list of all keys is too small: regrow, pull again
		{
			if(pass) { set_error("internal: key list regrow failed"); return SIBGPU_ERR_INTERNAL; }
			SIB_TRY(ctx->d_ckeys.ensure(sizeof(Rec16) * hs[12]));
			continue;
		}
		break;
	}
	const uint64_t Vc = hs[12];
	ctx->n_inst = 0;
	ctx->n_vertices = 0;
	ctx->x_Vc = Vc;
	if(Vc)
	{
		if(2 * Vc > 0xFFFFFFF0ull)
		{
			set_error("invalid: more than 2^32 vertices");
			return SIBGPU_ERR_INVALID;
		}
		bool collision = false;
		const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
		SIB_TRY(ids_and_tables<2>(ctx, ctx->dist_text, k, ctx->d_ckeys.as<Rec16>(), Vc, ntiles, fv, PT, false, &collision, 1));
		SIB_CUDA(cudaStreamSynchronize(st));
	}
	SIB_CUDA(cudaGetLastError());
	return SIBGPU_OK;
}

// second half of the k > 32 step, after the caller's min-reduction of the class representatives (d_rep, x_Vc entries)
int dist2_finish_fp(sibgpu_ctx *ctx, int *collision_out)
{
	NvtxRange nvtx("sibgpu: fused sharded step, k > 32 (back)");
	*collision_out = 0;
	const uint64_t Vc = ctx->x_Vc;
	if(Vc)
	{
		bool collision = false;
		const uint32_t ntiles = ctx->dist_tile_hi - ctx->dist_tile_lo;
		const FpView fv = {ctx->d_fp.as<FpCk>(), ctx->d_fpprm.as<FpParams>()};
		SIB_TRY(ids_and_tables<2>(ctx, ctx->dist_text, ctx->x_k, ctx->d_ckeys.as<Rec16>(), Vc, ntiles, fv, ctx->dist_P_total, false,
			&collision, 2));
		*collision_out = collision ? 1 : 0;
	}
	SIB_CUDA(cudaGetLastError());
	ctx->have_result = true;
	ctx->dist_result = true;
	if(ctx->profiling) SIB_TRY(ctx->prof_collect());
	return SIBGPU_OK;
}
