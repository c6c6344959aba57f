// Test harness (CPU): drives the host-side committer of sibgpu_simplify (sibelia_b200/csrc/simplifier.h) without a GPU.
// The vertex tables come from the caller (the tests pass the oracle's), and EVERY vertex is treated as flagged, i.e. the
// exact RemoveBulges restatement runs for all ids in order -- which must reproduce the reference stage bit for bit.
// With dirty_mode the later sweeps visit only the vertices dirtied since their last visit (the product's policy).
// Built by tests/test_host_commit.py into tests/_build/libhostcommit.so.
#include <cstdlib>

#include "../sibelia_b200/csrc/simplifier.h"

using namespace sibgpu::simp;

extern "C" int host_simplify(uint32_t nchr, char **seq, uint32_t **origpos, uint64_t *len, uint32_t k, uint32_t D,
	uint32_t max_iterations, const sibgpu_inst *pos, uint64_t npos, const sibgpu_inst *neg, uint64_t nneg, uint32_t count,
	uint64_t *bulges, uint64_t *exact_calls, int dirty_mode)
{
	Simplifier S;
	S.build(nchr, seq, origpos, len, k, D, count, pos, npos, neg, nneg);
	S.slot_of.assign((size_t)count + 1, -1);
	size_t total_bulges = 0, iterations = 0, calls = 0;
	do
	{
		iterations++;
		if(iterations > 1 && !dirty_mode)
		{
			S.compact();
			S.compact_nodes();
		}
		for(size_t id = 0; id <= count; id++)
		{
			// dirty_mode = the sweep policy of sibgpu_simplify: every vertex in the first sweep (there the GPU flags a
			// superset of the vertices with bulges), afterwards only the vertices dirtied since their last visit
			if(dirty_mode && iterations > 1 && !S.dirty[id]) continue;
			S.dirty[id] = 0;
			total_bulges += S.remove_bulges(id);
			calls++;
		}
	}
	while(total_bulges > 0 && iterations < max_iterations);
	for(uint32_t c = 0; c < nchr; c++)
	{
		size_t n = 0;
		for(int32_t e = S.nxt[S.chr_first_sep[c]]; S.ch[e] != SEP; e = S.nxt[e]) n++;
		char *out_seq = static_cast<char*>(malloc(n + 1));
		uint32_t *out_pos = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (n + 1)));
		size_t j = 0;
		for(int32_t e = S.nxt[S.chr_first_sep[c]]; S.ch[e] != SEP; e = S.nxt[e], j++)
		{
			out_seq[j] = S.ch[e];
			out_pos[j] = S.opos[e];
		}
		seq[c] = out_seq;
		origpos[c] = out_pos;
		len[c] = n;
	}
	*bulges = total_bulges;
	if(exact_calls) *exact_calls = calls;
	return 0;
}

extern "C" void host_free(void *p) { free(p); }
