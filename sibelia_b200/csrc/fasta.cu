// sibgpu_fasta_parse -- FASTA ingest on the GPU: replaces FASTAReader::GetSequences (/root/reference/src/fasta.cpp:22-73,
// with ValidateHeader :75-90 and ValidateSequence :92-106).
//
// The reference reads the file line by line (std::getline), trims each line (boost::algorithm::trim: the six C-locale
// white-space characters), skips empty lines, starts a record at every '>' line (description = the text between '>'
// and the first blank), upper-cases and validates every sequence line against "ACGTURYKMSWBDHWNX-" and appends it to
// the current record.  Here the whole file goes to the device once and
//   cub::DeviceSelect   positions of the '\n' bytes                                         -> lines
//   k_line_info         one thread per line: trimmed extent (only the white-space runs at both ends are walked, a
//                       line may be 100 MB long), kind (empty / sequence / header), description extent
//   cub::DeviceScan     over the lines: number of the line among the non-empty ones (the reference's error line
//                       counter), record index (= headers so far), offset of the line's payload in its record
//   k_line_place        destination of every sequence line in the '$' rec0 '$' rec1 ... '$' text (the DNASequence layout
//                       the enumerator uses), header table, empty-header errors
//   k_records           record lengths, "empty sequence" errors (a header directly after a header, or nothing at the end)
//   k_copy              byte-parallel over the FILE: 16 bytes per thread, the line of a byte by binary search in the
//                       newline positions, upper-casing + validation + store at the line's destination
// so the cost does not depend on how the sequence is wrapped.  Errors are reduced with one atomicMin on
// (non-empty line number, kind/column): the first error in file order wins, exactly the one the reference throws.
// Quirks kept: sequence lines before the first header are glued to the first record; a file without any header gives
// one record with an empty description; only ' ' (not TAB) ends a description.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <cstring>

#include "context.h"

namespace sibgpu {

__device__ __forceinline__ bool fa_is_space(uint8_t c) { return c == ' ' || (c >= 9 && c <= 13); }

// upper-cased character if legal in a sequence line, 0 otherwise (VALID_CHARS of fasta.cpp:94)
__device__ __forceinline__ uint8_t fa_sequence_char(uint8_t c)
{
	const uint8_t u = (c >= 'a' && c <= 'z') ? c - 32 : c;
	switch(u)
	{
	case 'A': case 'C': case 'G': case 'T': case 'U': case 'R': case 'Y': case 'K': case 'M': case 'S': case 'W': case 'B':
	case 'D': case 'H': case 'N': case 'X': case '-':
		return u;
	}
	return 0;
}

struct NewlineFlag {
	const uint8_t *raw;
	__device__ __forceinline__ bool operator()(uint32_t i) const { return raw[i] == '\n'; }
};

constexpr uint8_t FA_EMPTY = 0, FA_SEQ = 1, FA_HEADER = 2, FA_HEADER_NONAME = 3;
constexpr unsigned long long FA_NO_ERROR = ~0ull;
// error key = (number of the line among the non-empty lines, 1-based) << 32 | sub;  sub 0 = "empty sequence" (checked
// first at a header line, fasta.cpp:43), 1 = "empty header", 2 + column = "illegal character" at that column
__device__ __forceinline__ unsigned long long fa_key(uint32_t lineno, uint32_t sub) { return ((unsigned long long)lineno << 32) | sub; }

__global__ void __launch_bounds__(256) k_line_info(const uint8_t *__restrict__ raw, uint32_t nbytes, const uint32_t *__restrict__ nlpos,
	uint32_t nnl, uint32_t *__restrict__ lb, uint32_t *__restrict__ le, uint8_t *__restrict__ kind, uint32_t *__restrict__ nonempty,
	uint32_t *__restrict__ ishdr, uint64_t *__restrict__ payload, uint32_t *__restrict__ name_len)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if(j > nnl) return;
	uint32_t b = j == 0 ? 0u : nlpos[j - 1] + 1u, e = j < nnl ? nlpos[j] : nbytes;
	while(b < e && fa_is_space(raw[b])) b++;
	while(e > b && fa_is_space(raw[e - 1])) e--;
	uint8_t kd = FA_EMPTY;
	uint32_t nl = 0;
	if(b < e)
	{
		if(raw[b] == '>')
		{
			uint32_t x = b + 1;
			while(x < e && raw[x] != ' ') x++;             // ValidateHeader: up to the first blank, or the whole rest
			nl = x - (b + 1);
			kd = nl ? FA_HEADER : FA_HEADER_NONAME;
		}
		else kd = FA_SEQ;
	}
	lb[j] = b;
	le[j] = e;
	kind[j] = kd;
	nonempty[j] = kd != FA_EMPTY;
	ishdr[j] = kd >= FA_HEADER;
	payload[j] = kd == FA_SEQ ? e - b : 0u;
	name_len[j] = nl;
}

// after the scans: lineno[] / hdrcnt[] inclusive, payoff[] exclusive
__global__ void __launch_bounds__(256) k_line_place(uint32_t nlines, const uint8_t *__restrict__ kind, const uint32_t *__restrict__ lb,
	const uint32_t *__restrict__ lineno, const uint32_t *__restrict__ hdrcnt, const uint64_t *__restrict__ payoff,
	const uint32_t *__restrict__ name_len, uint64_t *__restrict__ dest, uint32_t *__restrict__ hdr_line,
	uint32_t *__restrict__ hdr_name_off, uint32_t *__restrict__ hdr_name_len, unsigned long long *__restrict__ err)
{
	const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if(j >= nlines) return;
	const uint8_t kd = kind[j];
	const uint32_t h = hdrcnt[j];
	if(kd == FA_SEQ)
	{
		const uint32_t rec = h ? h - 1 : 0;                // lines before the first header are glued to the first record
		dest[j] = 1ull + payoff[j] + rec;
	}
	else if(kd >= FA_HEADER)
	{
		hdr_line[h - 1] = j;
		hdr_name_off[h - 1] = lb[j] + 1;
		hdr_name_len[h - 1] = name_len[j];
		if(kd == FA_HEADER_NONAME) atomicMin(err, fa_key(lineno[j], 1u));
	}
}

// rec_base[r] = payload before record r (r = 0 .. nrec), "empty sequence" errors
__global__ void __launch_bounds__(256) k_records(uint32_t nhdr, uint32_t nrec, uint64_t total, uint32_t nonempty_total,
	const uint32_t *__restrict__ hdr_line, const uint64_t *__restrict__ payoff, const uint32_t *__restrict__ lineno,
	uint64_t *__restrict__ rec_base, unsigned long long *__restrict__ err)
{
	const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
	if(r > nrec) return;
	// record r starts at header r, except that record 0 also owns whatever precedes the first header
	const uint64_t base = (r == 0 || r >= nhdr) ? (r == 0 ? 0ull : total) : payoff[hdr_line[r]];
	rec_base[r] = base;
	if(r >= 1 && r < nhdr)
	{
		// at header r the reference flushes the sequence gathered since header r - 1 (since the file start for r = 1)
		const uint64_t since = r == 1 ? 0ull : payoff[hdr_line[r - 1]];
		if(base == since) atomicMin(err, fa_key(lineno[hdr_line[r]], 0u));
	}
	if(r == nrec)
	{
		const uint64_t since = nhdr >= 2 ? payoff[hdr_line[nhdr - 1]] : 0ull;
		if(total == since) atomicMin(err, fa_key(nonempty_total + 1u, 0u));     // after the loop, fasta.cpp:63
	}
}

constexpr uint32_t COPY_BYTES = 16;
__global__ void __launch_bounds__(256) k_copy(const uint8_t *__restrict__ raw, uint32_t nbytes, const uint32_t *__restrict__ nlpos,
	uint32_t nnl, const uint8_t *__restrict__ kind, const uint32_t *__restrict__ lb, const uint32_t *__restrict__ le,
	const uint64_t *__restrict__ dest, const uint32_t *__restrict__ lineno, uint8_t *__restrict__ text,
	unsigned long long *__restrict__ err)
{
	const uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * COPY_BYTES;
	if(i0 >= nbytes) return;
	// line of byte i0 = number of newlines before it
	uint32_t lo = 0, hi = nnl;
	while(lo < hi)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if(nlpos[mid] < i0) lo = mid + 1; else hi = mid;
	}
	uint32_t j = lo;
	uint32_t line_end = j < nnl ? nlpos[j] : nbytes;       // the newline that ends line j
	uint8_t kd = kind[j];
	uint32_t b = lb[j], e = le[j];
	uint64_t d = kd == FA_SEQ ? dest[j] : 0;
	const uint32_t i1 = (uint32_t)(i0 + COPY_BYTES < nbytes ? i0 + COPY_BYTES : nbytes);
	for(uint32_t i = (uint32_t)i0; i < i1; i++)
	{
		if(i > line_end)
		{
			j++;
			line_end = j < nnl ? nlpos[j] : nbytes;
			kd = kind[j];
			b = lb[j];
			e = le[j];
			d = kd == FA_SEQ ? dest[j] : 0;
		}
		if(kd == FA_SEQ && i >= b && i < e)
		{
			const uint8_t u = fa_sequence_char(raw[i]);
			if(u) text[d + (i - b)] = u;
			else atomicMin(err, fa_key(lineno[j], 2u + (i - b)));
		}
	}
}

} // namespace sibgpu

using namespace sibgpu;

extern "C" void sibgpu_fasta_free(sibgpu_fasta *f)
{
	if(!f) return;
	sibgpu_free(f->text_block);
	free(f->name_block);
	free(f->rec);
	memset(f, 0, sizeof(*f));
}

extern "C" int sibgpu_fasta_parse(sibgpu_ctx *ctx, const char *data, uint64_t nbytes, sibgpu_fasta *out, uint64_t *err_line)
{
	if(!ctx || !out || (nbytes && !data))
	{
		set_error("invalid: NULL argument");
		return SIBGPU_ERR_INVALID;
	}
	memset(out, 0, sizeof(*out));
	if(err_line) *err_line = 0;
	if(nbytes >= 0xFFFFFF00ull)
	{
		set_error("invalid: FASTA files of 4 GB and more are not supported (the reference stops at 1 GB of sequence)");
		return SIBGPU_ERR_INVALID;
	}
	NvtxRange nvtx("sibgpu: FASTA ingest");
	SIB_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const uint32_t n = (uint32_t)nbytes;
	ctx->have_text = false;                                // the text buffer is reused below
	ctx->have_result = false;
	// ---- the file, newline positions
	DevBuf &d_raw = ctx->d_records2;                       // enumeration workspace doubles as ingest workspace
	SIB_TRY(d_raw.ensure((size_t)n + 64));
	SIB_TRY(ctx->d_scalars.ensure(sizeof(uint64_t) * 64));
	if(n) SIB_CUDA(cudaMemcpyAsync(d_raw.p, data, n, cudaMemcpyHostToDevice, st));
	uint64_t *ds = ctx->d_scalars.as<uint64_t>();
	uint64_t *hs = static_cast<uint64_t*>(ctx->h_scalars);
	SIB_TRY(ctx->d_vkeys.ensure(sizeof(uint32_t) * ((size_t)n + 16)));
	uint32_t *nlpos = ctx->d_vkeys.as<uint32_t>();
	uint32_t nnl = 0;
	if(n)
	{
		thrust::counting_iterator<uint32_t> idx(0);
		NewlineFlag pred = {d_raw.as<uint8_t>()};
		size_t tmp_bytes = 0;
		SIB_CUDA(cub::DeviceSelect::If(nullptr, tmp_bytes, idx, nlpos, reinterpret_cast<uint32_t*>(ds + 20), (int)n, pred, st));
		SIB_TRY(ctx->d_cubtmp.ensure(tmp_bytes));
		SIB_CUDA(cub::DeviceSelect::If(ctx->d_cubtmp.p, tmp_bytes, idx, nlpos, reinterpret_cast<uint32_t*>(ds + 20), (int)n, pred, st));
		SIB_CUDA(cudaMemcpyAsync(hs + 20, ds + 20, 8, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaStreamSynchronize(st));
		nnl = (uint32_t)(hs[20] & 0xFFFFFFFFu);
	}
	const uint32_t nlines = nnl + 1;
	// ---- per-line arrays (one allocation)
	const size_t L = ((size_t)nlines + 15) / 16 * 16;
	DevBuf &d_lines = ctx->d_vkeys_alt;
	SIB_TRY(d_lines.ensure(L * (4 * 8 + 8 * 3 + 1) + 256));
	unsigned char *base = static_cast<unsigned char*>(d_lines.p);
	uint64_t *payload = reinterpret_cast<uint64_t*>(base);
	uint64_t *payoff = payload + L, *dest = payoff + L;
	uint32_t *lb = reinterpret_cast<uint32_t*>(dest + L), *le = lb + L, *nonempty = le + L, *ishdr = nonempty + L;
	uint32_t *lineno = ishdr + L, *hdrcnt = lineno + L, *name_len = hdrcnt + L, *scratch32 = name_len + L;
	uint8_t *kind = reinterpret_cast<uint8_t*>(scratch32 + L);
	const uint32_t lgrid = (nlines + 255) / 256;
	k_line_info<<<lgrid, 256, 0, st>>>(d_raw.as<uint8_t>(), n, nlpos, nnl, lb, le, kind, nonempty, ishdr, payload, name_len);
	{
		size_t t1 = 0, t2 = 0;
		SIB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, t1, nonempty, lineno, (int)nlines, st));
		SIB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, payload, payoff, (int)nlines, st));
		SIB_TRY(ctx->d_cubtmp.ensure(t1 > t2 ? t1 : t2));
		SIB_CUDA(cub::DeviceScan::InclusiveSum(ctx->d_cubtmp.p, t1, nonempty, lineno, (int)nlines, st));
		SIB_CUDA(cub::DeviceScan::InclusiveSum(ctx->d_cubtmp.p, t1, ishdr, hdrcnt, (int)nlines, st));
		SIB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_cubtmp.p, t2, payload, payoff, (int)nlines, st));
	}
	// totals of the last line -> host (record count, text size)
	SIB_CUDA(cudaMemcpyAsync(hs + 21, lineno + (nlines - 1), 4, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaMemcpyAsync(hs + 22, hdrcnt + (nlines - 1), 4, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaMemcpyAsync(hs + 23, payoff + (nlines - 1), 8, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaMemcpyAsync(hs + 24, payload + (nlines - 1), 8, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	const uint32_t nonempty_total = (uint32_t)(hs[21] & 0xFFFFFFFFu), nhdr = (uint32_t)(hs[22] & 0xFFFFFFFFu);
	const uint64_t total = hs[23] + hs[24];
	const uint32_t nrec = nhdr ? nhdr : 1;
	const uint64_t M = total + nrec + 1;
	// ---- header table, record bases, text
	DevBuf &d_hdr = ctx->d_cnt2;
	SIB_TRY(d_hdr.ensure(((size_t)nrec + 2) * (3 * 4 + 8) + 64));
	uint64_t *rec_base = d_hdr.as<uint64_t>();
	uint32_t *hdr_line = reinterpret_cast<uint32_t*>(rec_base + nrec + 2), *hdr_name_off = hdr_line + nrec + 1, *hdr_name_len = hdr_name_off + nrec + 1;
	SIB_TRY(ctx->d_text.ensure(M + 64));
	SIB_CUDA(cudaMemsetAsync(ctx->d_text.p, '$', M, st));
	SIB_CUDA(cudaMemsetAsync(ds + 25, 0xFF, 8, st));
	unsigned long long *d_err = reinterpret_cast<unsigned long long*>(ds + 25);
	k_line_place<<<lgrid, 256, 0, st>>>(nlines, kind, lb, lineno, hdrcnt, payoff, name_len, dest, hdr_line, hdr_name_off, hdr_name_len, d_err);
	k_records<<<(nrec + 1 + 255) / 256, 256, 0, st>>>(nhdr, nrec, total, nonempty_total, hdr_line, payoff, lineno, rec_base, d_err);
	if(n)
	{
		const uint64_t threads = ((uint64_t)n + COPY_BYTES - 1) / COPY_BYTES;
		k_copy<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(d_raw.as<uint8_t>(), n, nlpos, nnl, kind, lb, le, dest, lineno,
			ctx->d_text.as<uint8_t>(), d_err);
	}
	ctx->total_launches = 9;
	SIB_CUDA(cudaMemcpyAsync(hs + 25, ds + 25, 8, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	if(hs[25] != FA_NO_ERROR)
	{
		// the first error in file order; the message of an illegal character needs the character: find the line on the host
		const uint32_t lineno_err = (uint32_t)(hs[25] >> 32), sub = (uint32_t)(hs[25] & 0xFFFFFFFFu);
		if(err_line) *err_line = lineno_err;
		if(sub == 0) set_error("empty sequence");
		else if(sub == 1) set_error("empty header");
		else
		{
			auto is_space = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
			uint64_t at = 0;
			uint32_t seen = 0;
			char orig = '?';
			while(at <= nbytes)
			{
				uint64_t end = at;
				while(end < nbytes && data[end] != '\n') end++;
				uint64_t b = at, e = end;
				while(b < e && is_space((unsigned char)data[b])) b++;
				while(e > b && is_space((unsigned char)data[e - 1])) e--;
				if(b < e && ++seen == lineno_err)
				{
					orig = data[b + (sub - 2)];
					break;
				}
				at = end + 1;
			}
			set_error(std::string("illegal character: ") + orig);
		}
		return SIBGPU_ERR_INPUT;
	}
	// ---- results to the host: one pinned block with the text (the sequences are slices of it), names from the caller's
	// own buffer, lengths from the record bases
	std::vector<uint64_t> h_base(nrec + 1);
	std::vector<uint32_t> h_off(nrec), h_len(nrec);
	SIB_CUDA(cudaMemcpyAsync(h_base.data(), rec_base, sizeof(uint64_t) * (nrec + 1), cudaMemcpyDeviceToHost, st));
	if(nhdr)
	{
		SIB_CUDA(cudaMemcpyAsync(h_off.data(), hdr_name_off, sizeof(uint32_t) * nhdr, cudaMemcpyDeviceToHost, st));
		SIB_CUDA(cudaMemcpyAsync(h_len.data(), hdr_name_len, sizeof(uint32_t) * nhdr, cudaMemcpyDeviceToHost, st));
	}
	char *text = static_cast<char*>(pool_alloc(M + 1));
	if(!text)
	{
		set_error("invalid: host allocation failed");
		return SIBGPU_ERR_INVALID;
	}
	SIB_CUDA(cudaMemcpyAsync(text, ctx->d_text.p, M, cudaMemcpyDeviceToHost, st));
	SIB_CUDA(cudaStreamSynchronize(st));
	uint64_t name_bytes = 0;
	for(uint32_t r = 0; r < nhdr; r++) name_bytes += (uint64_t)h_len[r] + 1;
	char *names = static_cast<char*>(malloc(name_bytes + 1));
	sibgpu_fasta_record *rec = static_cast<sibgpu_fasta_record*>(malloc(sizeof(sibgpu_fasta_record) * nrec));
	if(!names || !rec)
	{
		free(names);
		free(rec);
		sibgpu_free(text);
		set_error("invalid: host allocation failed");
		return SIBGPU_ERR_INVALID;
	}
	uint64_t at = 0;
	names[name_bytes] = 0;
	for(uint32_t r = 0; r < nrec; r++)
	{
		if(r < nhdr)
		{
			memcpy(names + at, data + h_off[r], h_len[r]);
			names[at + h_len[r]] = 0;
			rec[r].name = names + at;
			rec[r].name_len = h_len[r];
			at += (uint64_t)h_len[r] + 1;
		}
		else
		{
			rec[r].name = names + name_bytes;              // a file without any header: one record, empty description
			rec[r].name_len = 0;
		}
		rec[r].seq = text + 1 + h_base[r] + r;
		rec[r].len = h_base[r + 1] - h_base[r];
	}
	out->nrec = nrec;
	out->rec = rec;
	out->text_block = text;
	out->name_block = names;
	out->total = total;
	return SIBGPU_OK;
}
