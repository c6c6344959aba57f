// sibgpu_list_edges -- replaces  IndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_);  ListEdges(iseq.Sequence(),
// iseq.BifStorage(), k, edge);  of BlockFinder::GenerateSyntenyBlocks (/root/reference/src/synteny.cpp:238-241) and
// SerializeCondensedGraph (src/serialization.cpp:90-93).
//
// The reference builds the whole host-side index (DNASequence + BifurcationStorage) only to walk both strands once
// and emit one Edge per pair of consecutive vertex marks (BlockFinder::ListEdges, src/serialization.cpp:56-86).  The
// two instance tables of the enumeration ARE that walk, already in its order ((chr, pos) per strand, positions in
// the strand's own coordinates), so the edges are the pairs of neighbouring table rows with equal chr: one thread per
// pair, an ordered compaction whose only irregularity is the one missing pair per non-empty chromosome.
// The original-position fields need originalPos_, which lives on the host: the kernel leaves the two element indices
// in their place and the host replaces them (2 reads per edge).
#include <algorithm>

#include "enum_common.cuh"

namespace sibgpu {

__global__ void __launch_bounds__(256) k_chr_first(const sibgpu_inst *__restrict__ tab, uint64_t n, uint32_t nchr,
	uint64_t *__restrict__ chrinst)
{
	// chrinst[c] = first row with chr >= c, c in 0..nchr
	uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
	if(c > nchr) return;
	uint64_t lo = 0, hi = n;
	while(lo < hi)
	{
		uint64_t mid = (lo + hi) >> 1;
		if(tab[mid].chr < c) lo = mid + 1; else hi = mid;
	}
	chrinst[c] = lo;
}

// skip[c] = number of non-empty chromosomes before c = pairs (i, i + 1) that straddle a chromosome change before the
// rows of c.  Row pair i of chromosome c becomes edge i - skip[c] of the strand.
__global__ void __launch_bounds__(256) k_list_edges(const sibgpu_inst *__restrict__ tab, uint64_t n, uint32_t strand,
	const uint32_t *__restrict__ skip, TextDesc t, uint32_t k, sibgpu_edge *__restrict__ out)
{
	const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i + 1 >= n) return;
	const sibgpu_inst a = tab[i], b = tab[i + 1];
	if(a.chr != b.chr) return;
	const uint32_t cs = __ldg(t.chr_start + a.chr), len = __ldg(t.chr_len + a.chr);
	const uint32_t step = b.pos - a.pos;
	// first character of the edge: the base k steps after the start vertex, read along the strand (serialization.cpp:75)
	const uint32_t q = a.pos + k;
	const uint32_t tp = strand == 0 ? cs + q : cs + (len - 1 - q);
	uint32_t code = (__ldg(t.packed + (tp >> 4)) >> (30 - 2 * (tp & 15u))) & 3u;
	if(strand) code = 3u - code;
	// elements spelled by the edge: strand positions a.pos .. a.pos + step + k - 1 (SpellOriginal, dnasequence.cpp:254-260)
	const uint32_t last = a.pos + step + k - 1;
	sibgpu_edge e;
	e.chr = a.chr;
	e.direction = strand;
	e.start_vertex = a.bifId;
	e.end_vertex = b.bifId;
	e.actual_position = strand == 0 ? a.pos : len - (a.pos + step + k);
	e.actual_length = step + k;
	e.original_position = strand == 0 ? a.pos : len - 1 - a.pos;           // element index of the first element (host: -> original)
	e.original_length = strand == 0 ? last : len - 1 - last;               // element index of the last element
	e.first_char = (uint32_t)"ACGT"[code];
	out[i - skip[a.chr]] = e;
}

int list_edges_device(sibgpu_ctx *ctx, uint32_t k, sibgpu_edge **edges_out, uint64_t *nedges_out)
{
	cudaStream_t st = ctx->stream;
	const uint64_t n = ctx->n_inst;
	const uint32_t nchr = ctx->nchr;
	*edges_out = nullptr;
	*nedges_out = 0;
	if(n < 2) return SIBGPU_OK;
	TextDesc t;
	t.packed = ctx->d_packed.as<uint32_t>();
	t.chr_start = ctx->d_chr_start.as<uint32_t>();
	t.chr_len = ctx->d_chr_len.as<uint32_t>();
	t.nchr = nchr;
	t.M = (uint32_t)ctx->M;
	t.nwords = (uint32_t)((ctx->M + 15) / 16) + 8;
	t.tile0 = 0;
	SIB_TRY(ctx->d_chrinst.ensure(sizeof(uint64_t) * (nchr + 2)));
	SIB_TRY(ctx->d_edge_skip.ensure(sizeof(uint32_t) * (nchr + 1)));
	SIB_TRY(ctx->d_edges.ensure(sizeof(sibgpu_edge) * 2 * n));
	std::vector<uint64_t> first(nchr + 1);
	std::vector<uint32_t> skip(nchr);
	const sibgpu_inst *tabs[2] = {ctx->d_pos.as<sibgpu_inst>(), ctx->d_neg.as<sibgpu_inst>()};
	uint64_t total = 0;
	for(uint32_t strand = 0; strand < 2; strand++)
	{
		k_chr_first<<<(nchr + 1 + 255) / 256, 256, 0, st>>>(tabs[strand], n, nchr, ctx->d_chrinst.as<uint64_t>());
		SIB_CUDA(cudaMemcpyAsync(first.data(), ctx->d_chrinst.p, sizeof(uint64_t) * (nchr + 1), cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		uint32_t nonempty = 0;
		for(uint32_t c = 0; c < nchr; c++)
		{
			skip[c] = nonempty;
			if(first[c + 1] > first[c]) nonempty++;
		}
		SIB_CUDA(cudaMemcpyAsync(ctx->d_edge_skip.p, skip.data(), sizeof(uint32_t) * nchr, cudaMemcpyHostToDevice, st));
		k_list_edges<<<(uint32_t)((n + 255) / 256), 256, 0, st>>>(tabs[strand], n, strand, ctx->d_edge_skip.as<uint32_t>(), t, k,
			ctx->d_edges.as<sibgpu_edge>() + total);
		ctx->total_launches += 2;
		total += n - nonempty;
		SIB_CUDA(cudaStreamSynchronize(st));               // skip[] is reused by the next strand
	}
	sibgpu_edge *host = static_cast<sibgpu_edge*>(malloc(sizeof(sibgpu_edge) * (total + 1)));
	if(!host)
	{
		set_error("invalid: host allocation failed");
		return SIBGPU_ERR_INVALID;
	}
	if(total) SIB_CUDA(cudaMemcpyAsync(host, ctx->d_edges.p, sizeof(sibgpu_edge) * total, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	SIB_CUDA(cudaGetLastError());
	*edges_out = host;
	*nedges_out = total;
	return SIBGPU_OK;
}

} // namespace sibgpu
