/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product (sibelia_b200/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built from this file.
 *
 * Plain-C restatement of the reference's bifurcation enumeration
 *     IndexedSequence::EnumerateBifurcationsSArrayInRAM   /root/reference/src/vertexenumeration.cpp:263-364
 * PARITY PINNED: tests/test_oracle.py checks this restatement against oracle/_ref/libsibelia_ref.so (the
 * unmodified reference compiled here) on the reference's own example genome, on the SURVEY.md section 4
 * known-answer vector and on seeded random inputs; the resulting tables are committed under tests/golden/.
 *
 * The reference sorts all suffixes of the "super-genome"  # chr0 # chr1 # ... # rc(chr0) # rc(chr1) # ...
 * (vertexenumeration.cpp:268-286) with libdivsufsort, computes the LCP array (Kasai, :44-65) and scans maximal
 * runs of suffixes with LCP >= k (:305-328).  A run whose first k characters contain '#' never yields a vertex
 * (every member fails `pos + k <= chrLen`, :341), so the runs that matter are exactly the classes of equal
 * k-mers over both strands.  This restatement therefore sorts the proper k-mer occurrences by their k characters
 * (memcmp on the same super-genome) instead of building a suffix array; run order = lexicographic k-mer order =
 * the order in which the reference hands out ids (:350).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint32_t bifId, chr, pos; } orc_inst;

static const char SEP = '#';                       /* IndexedSequence::SEPARATION_CHAR, vertexenumeration.cpp:11 */
static const unsigned char *g_super;
static uint32_t g_k;

static char complement(char c)                     /* DNASequence::Translate, dnasequence.cpp:41-44 (ACGT only) */
{
	switch(c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; }
	return c;
}

static int cmp_kmer(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
	int c = memcmp(g_super + x, g_super + y, g_k);
	if(c) return c;
	return x < y ? -1 : (x > y ? 1 : 0);           /* deterministic; order inside a run is irrelevant (:361-362 re-sort) */
}

static int cmp_inst(const void *a, const void *b)  /* BifurcationInstance::operator<, indexedsequence.h:64-67 */
{
	const orc_inst *x = a, *y = b;
	if(x->chr != y->chr) return x->chr < y->chr ? -1 : 1;
	if(x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
	return 0;
}

static int sym(char c) { return c == SEP ? 4 : (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); }
static int popcount5(unsigned m) { int r = 0; for(; m; m &= m - 1) r++; return r; }
/* Bifurcation(CharSet), vertexenumeration.cpp:67-70: more than one symbol, or the separator is present */
static int bifurcation(unsigned mask) { return popcount5(mask) > 1 || (mask & 16u); }

void orc_free(void *p) { free(p); }

/* Input must already be sanitised to ACGT (the rand() replacement of indexedsequence.cpp:31-37 is the caller's).
 * Returns the vertex count (the reference's return value `bifurcationCount`), or (uint64_t)-1 on allocation failure. */
uint64_t orc_enumerate(uint32_t nchr, const char *const *chr, const uint64_t *len, uint32_t k,
	orc_inst **pos_out, uint64_t *npos, orc_inst **neg_out, uint64_t *nneg)
{
	uint64_t total = 0, S, i, nocc = 0;
	uint32_t c;
	for(c = 0; c < nchr; c++) total += len[c];
	S = 2 * total + 2 * (uint64_t)nchr + 1;
	unsigned char *super = malloc(S + 1);
	uint64_t *cum = malloc(sizeof(uint64_t) * (2 * (size_t)nchr + 1));
	if(!super || !cum) return (uint64_t)-1;
	/* super-genome, :268-286 */
	uint64_t w = 0;
	super[w++] = SEP;
	for(c = 0; c < nchr; c++)
	{
		cum[c] = w;
		memcpy(super + w, chr[c], len[c]);
		w += len[c];
		super[w++] = SEP;
	}
	for(c = 0; c < nchr; c++)
	{
		cum[nchr + c] = w;
		for(i = 0; i < len[c]; i++) super[w++] = complement(chr[c][len[c] - 1 - i]);
		super[w++] = SEP;
	}
	cum[2 * nchr] = w;

	/* every proper k-mer occurrence (pos + k <= chrLen, :341) on both strands */
	for(c = 0; c < nchr; c++) if(len[c] >= k) nocc += 2 * (len[c] - k + 1);
	uint64_t *occ = malloc(sizeof(uint64_t) * (nocc + 1));
	if(!occ) return (uint64_t)-1;
	uint64_t n = 0;
	for(c = 0; c < 2 * nchr; c++)
	{
		uint64_t L = len[c % nchr];
		if(L >= k) for(i = 0; i + k <= L; i++) occ[n++] = cum[c] + i;
	}
	g_super = super; g_k = k;
	qsort(occ, n, sizeof(uint64_t), cmp_kmer);

	orc_inst *out[2];
	uint64_t cnt[2] = {0, 0}, cap[2] = {1024, 1024};
	out[0] = malloc(sizeof(orc_inst) * cap[0]);
	out[1] = malloc(sizeof(orc_inst) * cap[1]);
	uint32_t count = 0;
	for(uint64_t start = 0; start < n; )
	{
		uint64_t end = start;
		unsigned prev = 0, next = 0;
		int terminal = 0;
		do                                                         /* :314-328 */
		{
			uint64_t s = occ[end];
			prev |= 1u << sym(super[s - 1]);                       /* s > 0 always: super[0] == '#' */
			next |= 1u << sym(super[s + k]);                       /* s + k < S always: super ends with '#' */
			terminal |= super[s - 1] == SEP || super[s + k] == SEP;   /* :343 */
		}
		while(++end < n && memcmp(super + occ[end], super + occ[start], k) == 0);
		if((bifurcation(prev) || bifurcation(next)) && (end - start > 1 || terminal))   /* :330, :348 */
		{
			for(i = start; i < end; i++)
			{
				uint64_t s = occ[i];
				uint32_t lo = 0, hi = 2 * nchr;                    /* upper_bound on cumSize, :337 */
				while(hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if(cum[mid] <= s) lo = mid; else hi = mid; }
				int strand = lo < nchr ? 0 : 1;
				orc_inst r = {count, lo % nchr, (uint32_t)(s - cum[lo])};
				if(cnt[strand] == cap[strand])
				{
					cap[strand] *= 2;
					out[strand] = realloc(out[strand], sizeof(orc_inst) * cap[strand]);
				}
				out[strand][cnt[strand]++] = r;
			}
			count++;                                               /* :350 */
		}
		start = end;
	}
	qsort(out[0], cnt[0], sizeof(orc_inst), cmp_inst);             /* :361-362 */
	qsort(out[1], cnt[1], sizeof(orc_inst), cmp_inst);
	*pos_out = out[0]; *npos = cnt[0];
	*neg_out = out[1]; *nneg = cnt[1];
	free(occ); free(super); free(cum);
	return count;
}
