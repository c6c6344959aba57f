// k > 32: the k-mer no longer fits a 64-bit word, so classes are found on a 61-bit fingerprint
//   a = polynomial hash mod 2^61-1, rolled along the text,
// the canonical form being the smaller of the fingerprints of w and revcomp(w); the record {a, context} then takes the
// same path as the exact 16-byte records of k = 29..32.  Two different k-mers share a fingerprint with probability
// ~ n^2 / 2^62 per run (0.05 for 5 * 10^8 distinct k-mers); a false merge can only ADD symbols to a class, i.e. turn a
// non-vertex into a vertex, and that is caught below.  Vertex ids still have to be the
// lexicographic ranks of the actual k-mers (vertexenumeration.cpp:350), so the (few) vertex classes are ranked by
// comparing the strings they spell in the packed text, and every emitted instance is verified against its class
// representative (k_emit) -- a collision that could alter the result forces a re-run with other bases.
#include <cub/cub.cuh>

#include "enum_common.cuh"

namespace sibgpu {

constexpr uint64_t P61 = (1ull << 61) - 1;

__host__ __device__ __forceinline__ uint64_t mulmod61(uint64_t a, uint64_t b)
{
#ifdef __CUDA_ARCH__
	const uint64_t hi = __umul64hi(a, b), lo = a * b;
#else
	const unsigned __int128 z = (unsigned __int128)a * b;
	const uint64_t hi = (uint64_t)(z >> 64), lo = (uint64_t)z;
#endif
	uint64_t r = (lo & P61) + (lo >> 61) + (hi << 3);      // 2^61 = 1 (mod p); a, b < 2^61 so hi < 2^58
	r = (r & P61) + (r >> 61);
	return r >= P61 ? r - P61 : r;
}

__host__ __device__ __forceinline__ uint64_t addmod61(uint64_t a, uint64_t b)
{
	uint64_t r = a + b;
	return r >= P61 ? r - P61 : r;
}

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

struct FpParams {
	uint64_t B1, invB1, B2, invB2;
	uint64_t T1f[4], T1r[4], T2f[4], T2r[4];              // (c+1) B^(k-1) and (4-c) B^(k-1) for both hashes
};

__device__ __forceinline__ uint32_t code_at(const TextDesc &t, uint32_t j)
{
	const uint32_t w = j >> 4;
	if(w >= t.nwords) return 0u;
	return (__ldg(t.packed + w) >> (30u - 2u * (j & 15u))) & 3u;
}

// One thread rolls both fingerprints over a run of L consecutive text positions (O(k) warm-up, O(1) per position).
//   Hf(i) = sum_j (c[i+j]+1) B^(k-1-j)        fingerprint of the k-mer at i
//   Hr(i) = sum_m (4-c[i+m]) B^m              the same polynomial evaluated on its reverse complement
__global__ void __launch_bounds__(128) k_fingerprint(TextDesc t, uint32_t k, uint32_t L, FpParams prm, Rec16 *__restrict__ fp)
{
	const uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	const uint64_t start64 = g * L;
	if(start64 >= t.M) return;
	const uint32_t start = (uint32_t)start64;
	const uint32_t end = start + L < t.M ? start + L : t.M;
	uint64_t hf1 = 0, hr1 = 0;
	for(uint32_t j = 0; j < k; j++)
	{
		const uint32_t c = code_at(t, start + j);
		hf1 = addmod61(mulmod61(hf1, prm.B1), c + 1);
	}
	for(uint32_t j = k; j-- > 0; )
	{
		const uint32_t c = code_at(t, start + j);
		hr1 = addmod61(mulmod61(hr1, prm.B1), 4 - c);
	}
	ChrCursor cur;
	cur.init(t, start);
	uint32_t prevc = start ? code_at(t, start - 1) : 0u;
	for(uint32_t p = start; p < end; p++)
	{
		const uint32_t cout = code_at(t, p), cin = code_at(t, p + k);
		cur.advance(t, p);
		Rec16 out;
		out.a = EMPTY64;
		out.b = 0;
		if(p >= cur.cs && p + k <= cur.ce)
		{
			const uint32_t ps = p == cur.cs ? 4u : prevc;
			const uint32_t ns = p + k == cur.ce ? 4u : cin;
			const bool pal = hf1 == hr1;
			const bool fw = hf1 <= hr1;
			uint32_t ctx = fw ? ((ps << 3) | ns) : ((comp_sym(ns) << 3) | comp_sym(ps));
			ctx |= (pal ? 64u : 0u) | (fw ? 128u : 0u);
			out.a = fw ? hf1 : hr1;
			out.b = ctx;
		}
		fp[p] = out;
		hf1 = addmod61(mulmod61(addmod61(hf1, P61 - prm.T1f[cout]), prm.B1), cin + 1);
		hr1 = addmod61(mulmod61(addmod61(hr1, P61 - (4 - cout)), prm.invB1), prm.T1r[cin]);
		prevc = cout;
	}
}

int fingerprint_positions(sibgpu_ctx *ctx, const TextDesc &t, uint32_t k, uint32_t attempt)
{
	static const uint64_t B1S[3] = {0x0F3A5C7E9B1D2E4Full, 0x1B2D4F6A8C0E1357ull, 0x0A9C8E7F6D5B4A39ull};
	static const uint64_t B2S[3] = {0x9E3779B97F4A7C15ull, 0xC2B2AE3D27D4EB4Full, 0xD6E8FEB86659FD93ull};
	FpParams prm;
	prm.B1 = B1S[attempt % 3] % P61;
	prm.B2 = B2S[attempt % 3] | 1ull;
	prm.invB1 = powmod61(prm.B1, P61 - 2);
	prm.invB2 = inv64(prm.B2);
	const uint64_t top1 = powmod61(prm.B1, k - 1), top2 = pow64(prm.B2, k - 1);
	for(uint32_t c = 0; c < 4; c++)
	{
		prm.T1f[c] = mulmod61(c + 1, top1);
		prm.T1r[c] = mulmod61(4 - c, top1);
		prm.T2f[c] = (c + 1) * top2;
		prm.T2r[c] = (4 - c) * top2;
	}
	SIB_TRY(ctx->d_fp.ensure(sizeof(Rec16) * (size_t)t.M));
	uint32_t L = k < 128 ? 128 : ((k + 15) / 16) * 16;
	const uint64_t threads = (t.M + L - 1) / L;
	ProfScope ps(ctx, "k_fingerprint", (uint64_t)t.M / 4 * 2 + (uint64_t)t.M * 16);
	k_fingerprint<<<(uint32_t)((threads + 127) / 128), 128, 0, ctx->stream>>>(t, k, L, prm, ctx->d_fp.as<Rec16>());
	return SIBGPU_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// lexicographic ranking of the vertex k-mers
// ---------------------------------------------------------------------------------------------------------------
// item key = text position << 1 | dir (dir = 1: the k-mer at that position, dir = 0: its reverse complement);
// item value = class << 1 | (1 if the item is the reverse complement of the class's canonical string)
__global__ void __launch_bounds__(256) k_make_items(const unsigned long long *__restrict__ rep, uint64_t Vc,
	unsigned long long *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ npal)
{
	uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(j >= Vc) return;
	const unsigned long long R = rep[j];
	const unsigned long long p = R >> 2;
	const uint32_t pal = (uint32_t)(R >> 1) & 1u, fw = (uint32_t)R & 1u;
	keys[2 * j] = (p << 1) | fw;
	vals[2 * j] = (uint32_t)(j << 1);
	keys[2 * j + 1] = pal ? EMPTY64 : ((p << 1) | (fw ^ 1u));
	vals[2 * j + 1] = (uint32_t)(j << 1) | 1u;
	if(pal) atomicAdd(npal, 1u);
}

struct VStrLess {
	TextDesc t;
	uint32_t k;
	__device__ bool operator()(const unsigned long long &x, const unsigned long long &y) const
	{
		if(x == EMPTY64 || y == EMPTY64) return x != EMPTY64 && y == EMPTY64;   // palindrome placeholders sort last
		const uint32_t px = (uint32_t)(x >> 1), dx = (uint32_t)x & 1u, py = (uint32_t)(y >> 1), dy = (uint32_t)y & 1u;
		for(uint32_t m = 0; m * 32 < k; m++)
		{
			const uint64_t cx = vstr_chunk(t, px, dx, m, k), cy = vstr_chunk(t, py, dy, m, k);
			if(cx != cy) return cx < cy;
		}
		return false;
	}
};

__global__ void __launch_bounds__(256) k_assign_class_ids(const unsigned long long *__restrict__ keys,
	const uint32_t *__restrict__ vals, uint64_t n, const unsigned long long *__restrict__ rep, uint32_t *__restrict__ classids)
{
	uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
	if(i >= n || keys[i] == EMPTY64) return;
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
	SIB_TRY(ctx->d_vkeys.ensure(sizeof(uint64_t) * 2 * Vc));
	SIB_TRY(ctx->d_vkeys_alt.ensure(sizeof(uint32_t) * 2 * Vc));
	SIB_TRY(ctx->d_order.ensure(sizeof(uint32_t) * 2 * Vc));
	unsigned long long *keys = ctx->d_vkeys.as<unsigned long long>();
	uint32_t *vals = ctx->d_vkeys_alt.as<uint32_t>();
	SIB_CUDA(cudaMemsetAsync(ds + 3, 0, sizeof(uint64_t), st));
	{
		ProfScope ps(ctx, "k_make_items", Vc * 32);
		k_make_items<<<(uint32_t)((Vc + 255) / 256), 256, 0, st>>>(ctx->d_rep.as<unsigned long long>(), Vc, keys, vals,
			reinterpret_cast<uint32_t*>(ds + 3));
	}
	VStrLess less;
	less.t = t;
	less.k = k;
	size_t tmp_bytes = 0;
	SIB_CUDA(cub::DeviceMergeSort::SortPairs(nullptr, tmp_bytes, keys, vals, (int)(2 * Vc), less, st));
	SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
	{
		ProfScope ps(ctx, "cub_merge_sort_vertex_strings", 2 * Vc * 12 * 2, 8);
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
