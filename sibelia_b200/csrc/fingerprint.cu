// k > 32: the k-mer no longer fits a 64-bit word, so classes are found on fingerprints.  Two polynomial hashes are
// rolled along the text for the k-mer w and for revcomp(w):
//   h1 mod 2^61-1 (random base; 56 of its bits are the key an 8-byte record carries)
//   h2 mod 2^64   (its top bits choose the hash partition and are NOT stored: the partition index is part of the key)
// the canonical orientation being the one with the smaller h1.  The record then takes the path of the exact 8-byte
// records of k <= 28 (k_scatter -> k_split -> k_group in shared memory).  Two different k-mers are merged only if they
// agree in the 56 stored bits AND fall into the same partition: ~ n * (records per partition) / 2^57 per run (0.002 for
// 5 * 10^8 distinct k-mers); a false merge can only ADD symbols to a class, i.e. turn a non-vertex into a vertex, and
// that is caught: vertex ids have to be the lexicographic ranks of the actual k-mers (vertexenumeration.cpp:350), so the
// (few) vertex classes are ranked by comparing the strings they spell in the packed text, and every emitted instance
// is verified against its class representative (k_emit) -- a collision that could alter the result forces a re-run
// with other bases.
//
// The hashes are not materialised per position.  k_fp_ckpt stores them for every 16th position (32 bytes per packed
// word = 2 B/base); the scan kernels (k_scatter, k_mark: scan16_fp in enumerate.cu) start from the checkpoint of their
// word and roll 16 positions in registers.
#include <cub/cub.cuh>

#include "enum_common.cuh"

namespace sibgpu {

static uint64_t powmod61(uint64_t b, uint64_t e)
{
	uint64_t r = 1;
	while(e)
	{
		if(e & 1) r = mulmod61(r, b);
		b = mulmod61(b, b);
		e >>= 1;
	}
	return r;
}

static uint64_t pow64(uint64_t b, uint64_t e)
{
	uint64_t r = 1;
	while(e)
	{
		if(e & 1) r *= b;
		b *= b;
		e >>= 1;
	}
	return r;
}

static uint64_t inv64(uint64_t b)                       // inverse of an odd number mod 2^64 (Newton)
{
	uint64_t x = b;
	for(int i = 0; i < 6; i++) x *= 2 - b * x;
	return x;
}

// One thread computes the checkpoints of a run of L consecutive text positions (L a multiple of 16): Horner warm-up
// over the k bases of its first k-mer (forward for w, backward for revcomp(w)), then one rolling step per position.
//   Hf(i) = sum_j (c[i+j]+1) B^(k-1-j)        fingerprint of the k-mer at i
//   Hr(i) = sum_m (4-c[i+m]) B^m              the same polynomial evaluated on its reverse complement
// Checkpoints of the words [w_begin, w_stop) (sharded runs: the own text range).
__global__ void __launch_bounds__(128) k_fp_ckpt(TextDesc t, uint32_t k, uint32_t L, FpParams prm, FpCk *__restrict__ ck,
	uint32_t w_begin, uint32_t w_stop)
{
	__shared__ uint64_t sD[64];
	for(uint32_t i = threadIdx.x; i < 64; i += blockDim.x) sD[i] = prm.D[i >> 4][i & 15];
	__syncthreads();
	const FpBases b = prm.b;
	const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	const uint64_t w64 = w_begin + g * (L / 16);
	if(w64 >= w_stop) return;
	const uint32_t w_first = (uint32_t)w64;
	const uint32_t w_last = w_first + L / 16 < w_stop ? w_first + L / 16 : w_stop;
	auto word = [&](uint32_t w) -> uint32_t { return w < t.nwords ? __ldg(t.packed + w) : 0u; };
	FpState h = {0, 0, 0, 0};
	const uint32_t kw = k >> 4, kt = k & 15u;
	for(uint32_t wi = 0; wi < kw; wi++)
	{
		const uint32_t x = word(w_first + wi);
#pragma unroll
		for(int i = 0; i < 16; i++)
		{
			const uint32_t c = (x >> (30 - 2 * i)) & 3u;
			h.hf1 = muladdmod61(h.hf1, b.B1, c + 1);
			h.hf2 = h.hf2 * b.B2 + (c + 1);
		}
	}
	{
		const uint32_t x = word(w_first + kw);
		for(uint32_t i = 0; i < kt; i++)
		{
			const uint32_t c = (x >> (30 - 2 * i)) & 3u;
			h.hf1 = muladdmod61(h.hf1, b.B1, c + 1);
			h.hf2 = h.hf2 * b.B2 + (c + 1);
		}
		for(uint32_t i = kt; i-- > 0; )
		{
			const uint32_t c = (x >> (30 - 2 * i)) & 3u;
			h.hr1 = muladdmod61(h.hr1, b.B1, 4 - c);
			h.hr2 = h.hr2 * b.B2 + (4 - c);
		}
	}
	for(uint32_t wi = kw; wi-- > 0; )
	{
		const uint32_t x = word(w_first + wi);
#pragma unroll
		for(int i = 15; i >= 0; i--)
		{
			const uint32_t c = (x >> (30 - 2 * i)) & 3u;
			h.hr1 = muladdmod61(h.hr1, b.B1, 4 - c);
			h.hr2 = h.hr2 * b.B2 + (4 - c);
		}
	}
	const uint32_t sh = 2 * kt;
	uint32_t in_lo = word(w_first + kw);
	for(uint32_t w = w_first; w < w_last; w++)
	{
		ck[w] = FpCk{h.hf1, h.hr1, h.hf2, h.hr2};
		const uint32_t wout = word(w), in_hi = word(w + kw + 1);
		const uint32_t win = __funnelshift_l(in_hi, in_lo, sh);       // the codes of positions 16 w + k ... + 15
		in_lo = in_hi;
#pragma unroll
		for(int i = 0; i < 16; i++)
		{
			const uint32_t idx = (((wout >> (30 - 2 * i)) & 3u) << 2) | ((win >> (30 - 2 * i)) & 3u);
			fp_roll(h, sD, b, idx);
		}
	}
}

int fingerprint_positions(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, uint32_t attempt, uint32_t w_begin, uint32_t w_stop)
{
	static const uint64_t B1S[3] = {0x0F3A5C7E9B1D2E4Full, 0x1B2D4F6A8C0E1357ull, 0x0A9C8E7F6D5B4A39ull};
	static const uint64_t B2S[3] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0xD6E8FEB86659FD93ull};
	FpParams prm;
	FpBases &b = prm.b;
	b.B1 = B1S[attempt % 3] % P61;
	b.B2 = B2S[attempt % 3] | 1ull;
	b.invB1 = powmod61(b.B1, P61 - 2);
	b.invB2 = inv64(b.B2);
	// rolling step: Hf' = Hf B - (out+1) B^k + (in+1);   Hr' = Hr / B - (4-out) / B + (4-in) B^(k-1)
	const uint64_t top1 = powmod61(b.B1, k - 1), top2 = pow64(b.B2, k - 1);
	const uint64_t full1 = mulmod61(top1, b.B1), full2 = top2 * b.B2;
	for(uint32_t co = 0; co < 4; co++)
	{
		for(uint32_t ci = 0; ci < 4; ci++)
		{
			const uint32_t i = co * 4 + ci;
			prm.D[0][i] = addmod61(ci + 1, P61 - mulmod61(co + 1, full1));
			prm.D[1][i] = addmod61(mulmod61(4 - ci, top1), P61 - mulmod61(4 - co, b.invB1));
			prm.D[2][i] = (ci + 1) - (co + 1) * full2;
			prm.D[3][i] = (4 - ci) * top2 - (4 - co) * b.invB2;
		}
	}
	const uint32_t nck = (uint32_t)((t.M + 15) >> 4);
	SIB_TRY(ctx->d_fp.ensure(sizeof(FpCk) * (size_t)nck));
	SIB_TRY(ctx->d_fpprm.ensure(sizeof(FpParams)));
	SIB_CUDA(cudaMemcpyAsync(ctx->d_fpprm.p, &prm, sizeof(FpParams), cudaMemcpyHostToDevice, ctx->stream));
	if(w_stop <= w_begin) return SIBGPU_OK;
	const uint32_t L = k < 128 ? 128 : ((k + 15) / 16) * 16;
	const uint64_t threads = ((uint64_t)(w_stop - w_begin) * 16 + L - 1) / L;
	ProfScope ps(ctx, "k_fp_ckpt", (uint64_t)(w_stop - w_begin) * (8 + sizeof(FpCk)));
	k_fp_ckpt<<<(uint32_t)((threads + 127) / 128), 128, 0, ctx->stream>>>(t, k, L, prm, ctx->d_fp.as<FpCk>(), w_begin, w_stop);
	return SIBGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lexicographic ranking of the vertex k-mers
// ---------------------------------------------------------------------------------------------------------------
// item = text position << 1 | dir (dir = 1: the k-mer at that position, dir = 0: its reverse complement), carried with
// the first 32 bases of the string it spells: almost every comparison of the sort is decided on that prefix, without
// touching the text;  item value = class << 1 | (1 if the item is the reverse complement of the class's canonical string)
struct VItem { unsigned long long prefix, item; };

__global__ void __launch_bounds__(256) k_make_items(TextDesc t, uint32_t k, const unsigned long long *__restrict__ rep, uint64_t Vc,
	VItem *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ npal)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(j >= Vc) return;
	const unsigned long long R = rep[j];
	const unsigned long long p = R >> 2;
	const uint32_t pal = (uint32_t)(R >> 1) & 1u, fw = (uint32_t)R & 1u;
	keys[2 * j] = VItem{vstr_chunk(t, (uint32_t)p, fw, 0, k), (p << 1) | fw};
	vals[2 * j] = (uint32_t)(j << 1);
	keys[2 * j + 1] = pal ? VItem{EMPTY64, EMPTY64} : VItem{vstr_chunk(t, (uint32_t)p, fw ^ 1u, 0, k), (p << 1) | (fw ^ 1u)};
	vals[2 * j + 1] = (uint32_t)(j << 1) | 1u;
	if(pal) atomicAdd(npal, 1u);
}

struct VStrLess {
	TextDesc t;
	uint32_t k;
	__device__ bool operator()(const VItem &x, const VItem &y) const
	{
		if(x.item == EMPTY64 || y.item == EMPTY64) return x.item != EMPTY64 && y.item == EMPTY64;   // palindrome placeholders sort last
		if(x.prefix != y.prefix) return x.prefix < y.prefix;
		const uint32_t px = (uint32_t)(x.item >> 1), dx = (uint32_t)x.item & 1u, py = (uint32_t)(y.item >> 1), dy = (uint32_t)y.item & 1u;
		for(uint32_t m = 1; m * 32 < k; m++)
		{
			const uint64_t cx = vstr_chunk(t, px, dx, m, k), cy = vstr_chunk(t, py, dy, m, k);
			if(cx != cy) return cx < cy;
		}
		return false;
	}
};

__global__ void __launch_bounds__(256) k_assign_class_ids(const VItem *__restrict__ keys,
	const uint32_t *__restrict__ vals, uint64_t n, const unsigned long long *__restrict__ rep, uint32_t *__restrict__ classids)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n || keys[i].item == EMPTY64) return;
	const uint32_t v = vals[i], cls = v >> 1;
	classids[2 * cls + (v & 1u)] = (uint32_t)i;
	if(!(v & 1u) && ((rep[cls] >> 1) & 1ull)) classids[2 * cls + 1] = (uint32_t)i;   // palindrome: one vertex
}

__global__ void __launch_bounds__(256) k_map_assign(MapSlot *__restrict__ map, uint32_t Tm, const uint32_t *__restrict__ classids)
{
	uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= Tm || map[s].a == EMPTY64) return;
	const uint32_t cls = map[s].cls;
	map[s].idc = classids[2 * cls];
	map[s].idr = classids[2 * cls + 1];
}

int rank_fingerprint_vertices(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, uint64_t Vc, uint32_t Tm, uint32_t *V_out)
{
	cudaStream_t st = ctx->stream;
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	SIB_TRY(ctx->d_vkeys.ensure(sizeof(VItem) * 2 * Vc));
	SIB_TRY(ctx->d_vkeys_alt.ensure(sizeof(uint32_t) * 2 * Vc));
	SIB_TRY(ctx->d_order.ensure(sizeof(uint32_t) * 2 * Vc));
	VItem *keys = ctx->d_vkeys.as<VItem>();
	uint32_t *vals = ctx->d_vkeys_alt.as<uint32_t>();
	SIB_CUDA(cudaMemsetAsync(ds + 3, 0, sizeof(uint64_t), st));
	{
		ProfScope ps(ctx, "k_make_items", Vc * 48);
		k_make_items<<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(t, k, ctx->d_rep.as<unsigned long long>(), Vc, keys, vals,
			reinterpret_cast<uint32_t*>(ds + 3));
	}
	VStrLess less;
	less.t = t;
	less.k = k;
	size_t tmp_bytes = 0;
	SIB_CUDA(cub::DeviceMergeSort::SortPairs(nullptr, tmp_bytes, keys, vals, (int)(2 * Vc), less, st));
	SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
	{
		ProfScope ps(ctx, "cub_merge_sort_vertex_strings", 2 * Vc * 20 * 2, 8);
		SIB_CUDA(cub::DeviceMergeSort::SortPairs(ctx->d_cubtmp.p, tmp_bytes, keys, vals, (int)(2 * Vc), less, st));
	}
	{
		ProfScope ps(ctx, "k_assign_class_ids", 2 * Vc * 20);
		k_assign_class_ids<<<(uint32_t)((2 * Vc + 255) / 256), 256, 0, st>>>(keys, vals, 2 * Vc,
			ctx->d_rep.as<unsigned long long>(), ctx->d_order.as<uint32_t>());
	}
	{
		ProfScope ps(ctx, "k_map_assign", (uint64_t)Tm * sizeof(MapSlot));
		k_map_assign<<<(Tm + 255) / 256, 256, 0, st>>>(ctx->d_map.as<MapSlot>(), Tm, ctx->d_order.as<uint32_t>());
	}
	SIB_CUDA(cudaMemcpyAsync(hs + 3, ds + 3, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	*V_out = (uint32_t)(2 * Vc - (hs[3] & 0xFFFFFFFFull));
	return SIBGPU_OK;
}

} // namespace sibgpu
