// Test harness (CPU): drives the host-side committer of sibgpu_simplify (sibelia_b200/csrc/simplifier.h) without a GPU.
// The vertex tables come from the caller (the tests pass the oracle's), and EVERY vertex is treated as flagged, i.e. the
// exact RemoveBulges restatement runs for all ids in order -- which must reproduce the reference stage bit for bit.
// With dirty_mode the later sweeps visit only the vertices dirtied since their last visit (the product's policy);
// dirty_mode 2 is the full product policy: chunks of ids are screened by all host threads against the current state
// (Simplifier::screen_range) and only the survivors, plus whatever a collapse dirties afterwards, get the exact call.
// Built by tests/test_host_commit.py into tests/_build/libhostcommit.so.
#include <cstdlib>

#include "../sibelia_b200/csrc/simplifier.h"

using namespace sibgpu;
using namespace sibgpu::simp;

// Test-only: renumbering of the element arrays / list nodes between sweeps (the product never renumbers: it tracks dirty
// vertices instead); used by the non-dirty reference mode below to show that the committer does not depend on indices.
// Renumbers the elements in sequence order so that element index == flat position again (start of a sweep).
static void compact(Simplifier &S)
{
	const size_t total = S.live;
	std::vector<int32_t> newidx(S.ch.size(), -1);
	HostChars ch2(total);
	HostU32 op2(total), m0(total), m1(total);
	HostI32 no0(total), no1(total);
	size_t j = 0;
	for(int32_t e = 0; e >= 0; e = S.nxt[e], j++)
	{
		newidx[e] = (int32_t)j;
		ch2[j] = S.ch[e];
		op2[j] = S.opos[e];
		m0[j] = S.mark[0][e];
		m1[j] = S.mark[1][e];
		no0[j] = S.node_of[0][e];
		no1[j] = S.node_of[1][e];
	}
	for(size_t n = 0; n < S.n_elem.size(); n++)
	{
		if(S.n_valid[n]) S.n_elem[n] = newidx[S.n_elem[n]];
	}
	for(size_t c = 0; c < S.chr_first_sep.size(); c++) S.chr_first_sep[c] = newidx[S.chr_first_sep[c]];
	S.last_sep = newidx[S.last_sep];
	S.ch.swap(ch2);
	S.opos.swap(op2);
	S.mark[0].swap(m0);
	S.mark[1].swap(m1);
	S.node_of[0].swap(no0);
	S.node_of[1].swap(no1);
	S.nxt.resize(total);
	S.prv.resize(total);
	for(size_t i = 0; i < total; i++)
	{
		S.nxt[i] = (int32_t)i + 1;
		S.prv[i] = (int32_t)i - 1;
	}
	S.nxt[total - 1] = -1;
}

// Drops the list nodes that Cleanup already unlinked (keeps node indices small between sweeps).
static void compact_nodes(Simplifier &S)
{
	std::vector<int32_t> remap(S.n_elem.size(), -1);
	size_t j = 0;
	for(size_t n = 0; n < S.n_elem.size(); n++)
	{
		if(S.n_valid[n]) remap[n] = (int32_t)j++;
	}
	std::vector<int32_t> e2(j), nx2(j), pv2(j);
	std::vector<uint8_t> s2(j), v2(j, 1);
	std::vector<uint32_t> id2(j);
	for(size_t n = 0; n < S.n_elem.size(); n++)
	{
		if(!S.n_valid[n]) continue;
		const int32_t m = remap[n];
		e2[m] = S.n_elem[n];
		s2[m] = S.n_strand[n];
		id2[m] = S.n_id[n];
		nx2[m] = S.n_next[n] >= 0 ? remap[S.n_next[n]] : -1;
		pv2[m] = S.n_prev[n] >= 0 ? remap[S.n_prev[n]] : -1;
		S.node_of[s2[m]][e2[m]] = m;
	}
	for(int s = 0; s < 2; s++)
	{
		for(size_t id = 0; id < S.head[s].size(); id++)
		{
			if(S.head[s][id] >= 0) S.head[s][id] = remap[S.head[s][id]];
		}
	}
	S.n_elem.swap(e2);
	S.n_next.swap(nx2);
	S.n_prev.swap(pv2);
	S.n_strand.swap(s2);
	S.n_valid.swap(v2);
	S.n_id.swap(id2);
}



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
			compact(S);
			compact_nodes(S);
		}
		for(size_t id = 0; id <= count; id++)
		{
			// dirty_mode 2: the product's chunked screen (tiny chunks here so that it interleaves with the collapses)
			if(dirty_mode == 2 && id % 64 == 0)
			{
				S.screen_min = 1;
				if(iterations == 1) std::fill(S.dirty.begin() + id, S.dirty.begin() + std::min<size_t>(id + 64, (size_t)count + 1), 1);
				S.screen_range(id, std::min<size_t>(id + 64, (size_t)count + 1), nullptr);
			}
			// dirty_mode = the sweep policy of sibgpu_simplify: every vertex in the first sweep (there the GPU flags a
			// superset of the vertices with bulges), afterwards only the vertices dirtied since their last visit
			if(dirty_mode == 2 ? !S.dirty[id] : (dirty_mode && iterations > 1 && !S.dirty[id])) continue;
			S.dirty[id] = 0;
			// as in the product: survivors of the screen skip the existence pass of any_bulges
			const bool expect = dirty_mode == 2 && std::binary_search(S.ahead_id.begin(), S.ahead_id.end(), (uint32_t)id);
			total_bulges += S.remove_bulges(id, expect);
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
