// Host-side state of one simplification stage: the reference's DNASequence + BifurcationStorage + bulgeremoval.cpp
// logic restated on flat arrays (pure C++, no CUDA) -- see simplify.cu for how the GPU detection kernel and this exact
// in-order committer share the work.
#pragma once
#ifdef SIBGPU_COMMIT_PROF
#include <x86intrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/sibgpu.h"
#include "boost_order.h"
#include "hostvec.h"

namespace sibgpu {
namespace simp {

// Dev only (-DSIBGPU_COMMIT_PROF): time-stamp-counter split of the ordered commit, printed by sibgpu_simplify's trace.
#ifdef SIBGPU_COMMIT_PROF
struct CommitProf { unsigned long long t[12] = {}; };
inline CommitProf &commit_prof() { static CommitProf p; return p; }
struct ProfSection {
	int i; unsigned long long t0;
	explicit ProfSection(int idx) : i(idx), t0(__rdtsc()) {}
	~ProfSection() { commit_prof().t[i] += __rdtsc() - t0; }
};
#define SIB_PROF(i) ProfSection prof_section_##i(i)
#else
#define SIB_PROF(i)
#endif

const uint32_t NO_BIF = 0xFFFFFFFFu;                   // BifurcationStorage::NO_BIFURCATION, bifurcationstorage.cpp:12
const char SEP = '$';                                  // DNASequence::SEPARATION_CHAR, dnasequence.cpp:33
const char EMPTY = ' ';                                // bulgeremoval.cpp:13
const uint32_t POS_MASK = (1u << 29) - 1;              // StrandIterator::PositionMask, stranditerator.cpp:19-27

inline char complement(char c)                          // DNASequence::complementary_, dnasequence.cpp:11-29
{
	switch(c)
	{
	case 'A': return 'T';
	case 'T': return 'A';
	case 'C': return 'G';
	case 'G': return 'C';
	case 'a': return 't';
	case 't': return 'a';
	case 'c': return 'g';
	case 'g': return 'c';
	}
	return c;
}

// Runs fn(begin, end) over [0, n) split across host threads (the per-element arrays of a stage are 30 B/element: at
// 500 Mbases filling them single-threaded costs more than all the device work of the stage).  fn must be data-parallel.
template<class F>
inline void parallel_ranges(size_t n, F fn)
{
	const size_t min_chunk = 1u << 20;
	size_t nt = std::thread::hardware_concurrency();
	nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
	if(nt > (n + min_chunk - 1) / min_chunk) nt = (n + min_chunk - 1) / min_chunk;
	if(nt <= 1)
	{
		fn((size_t)0, n);
		return;
	}
	std::vector<std::thread> th;
	for(size_t t = 0; t < nt; t++) th.emplace_back(fn, n * t / nt, n * (t + 1) / nt);
	for(std::thread &x : th) x.join();
}

// DNASequence::StrandIterator over the element arrays: e = element, d = 0 positive / 1 negative strand
struct It {
	int32_t e;
	uint8_t d;
};

struct VisitData {                                      // blockfinder.h:18-24
	size_t kmerId, distance;
};

struct BifurcationMark {                                // bulgeremoval.cpp:20-37
	size_t bifId, distance;
	bool operator<(const BifurcationMark &o) const
	{
		if(bifId != o.bifId) return bifId < o.bifId;
		return distance < o.distance;
	}
};

class Simplifier {
public:
	// ---- sequence (DNASequence, dnasequence.cpp:75-103): element 0 is the leading '$'
	HostChars ch;                                       // (hostvec.h: resize() does not value-initialise)
	HostU32 opos;
	HostI32 nxt, prv;
	HostU32 mark[2];                                    // vertex id whose k-mer starts here on that strand, or NO_BIF
	HostI32 node_of[2];                                 // instance-list node of that mark
	std::vector<int32_t> chr_first_sep;                 // the '$' in front of chromosome c (chr c = elements after it
	int32_t last_sep = 0;                               //   up to the next '$')
	size_t live = 0;

	// ---- BifurcationStorage::bifurcationPos_ (bifurcationstorage.h:17-121): per strand and id an slist of nodes
	std::vector<int32_t> n_elem, n_next, n_prev;
	std::vector<uint8_t> n_valid, n_strand;
	std::vector<uint32_t> n_id;
	std::vector<int32_t> head[2];
	std::vector<uint32_t> lsize[2];
	std::vector<int32_t> to_clear;
	uint32_t max_id = 0;

	// ---- snapshot bookkeeping
	std::vector<uint8_t> dirty;
	size_t k = 0, D = 0;

	// =========================================================== iterators
	inline void inc(It &it) const { it.e = it.d == 0 ? nxt[it.e] : prv[it.e]; }
	inline char deref(It it) const { return it.d == 0 ? ch[it.e] : complement(ch[it.e]); }
	inline bool at_valid(It it) const { return ch[it.e] != SEP; }
	inline It advance(It it, size_t n) const
	{
		for(size_t i = 0; i < n; i++) inc(it);
		return it;
	}
	inline It invert(It it) const                       // StrandIterator::Invert, stranditerator.cpp:193-201
	{
		It r;
		if(it.d == 0) { r.e = prv[it.e]; r.d = 1; }
		else { r.e = nxt[it.e]; r.d = 0; }
		return r;
	}
	inline bool proper_kmer(It it, size_t n) const      // ProperKMer, dnasequence.h:154-165
	{
		for(size_t i = 0; i < n; i++, inc(it))
		{
			if(!at_valid(it)) return false;
		}
		return true;
	}

	// =========================================================== BifurcationStorage
	inline uint32_t get_bif(It it) const { return mark[it.d][it.e]; }           // GetBifurcation, .cpp:157-162

	void add_point(It it, size_t bif)                   // AddPoint, bifurcationstorage.cpp:113-126 (push-front)
	{
		if(mark[it.d][it.e] != NO_BIF || bif == NO_BIF) return;
		const uint32_t id = (uint32_t)bif;
		const int32_t n = (int32_t)n_elem.size();
		n_elem.push_back(it.e);
		n_strand.push_back(it.d);
		n_id.push_back(id);
		n_valid.push_back(1);
		n_prev.push_back(-1);
		n_next.push_back(head[it.d][id]);
		if(head[it.d][id] >= 0) n_prev[head[it.d][id]] = n;
		head[it.d][id] = n;
		lsize[it.d][id]++;
		mark[it.d][it.e] = id;
		node_of[it.d][it.e] = n;
		dirty[id] = 1;
	}

	void erase_point(It it)                              // ErasePoint, .cpp:144-155: lazy, the node stays until Cleanup
	{
		const uint32_t id = mark[it.d][it.e];
		if(id == NO_BIF) return;
		const int32_t n = node_of[it.d][it.e];
		mark[it.d][it.e] = NO_BIF;
		node_of[it.d][it.e] = -1;
		n_valid[n] = 0;
		to_clear.push_back(n);
		dirty[id] = 1;
	}

	void cleanup()                                       // Cleanup, .cpp:33-41
	{
		for(size_t i = 0; i < to_clear.size(); i++)
		{
			const int32_t n = to_clear[i];
			const uint8_t s = n_strand[n];
			const uint32_t id = n_id[n];
			if(n_prev[n] >= 0) n_next[n_prev[n]] = n_next[n]; else head[s][id] = n_next[n];
			if(n_next[n] >= 0) n_prev[n_next[n]] = n_prev[n];
			lsize[s][id]--;
		}
		to_clear.clear();
	}

	inline size_t count_bifurcations(size_t id) const { return (size_t)lsize[0][id] + lsize[1][id]; }   // .cpp:71-75

	void list_positions(size_t id, std::vector<int32_t> &out) const   // ListPositions, bifurcationstorage.h:59-72
	{
		out.clear();
		for(int s = 0; s < 2; s++)
		{
			for(int32_t n = head[s][id]; n >= 0; n = n_next[n]) out.push_back(n);
		}
	}
	inline It node_it(int32_t n) const                   // IteratorProxy::operator*
	{
		It it;
		it.e = n_elem[n];
		it.d = n_strand[n];
		return it;
	}

	// =========================================================== construction
	void build(uint32_t nchr, char *const *seq, uint32_t *const *origpos, const uint64_t *len, size_t k_, size_t D_,
		uint32_t count, const sibgpu_inst *pos, uint64_t npos, const sibgpu_inst *neg, uint64_t nneg)
	{
		k = k_;
		D = D_;
		max_id = count;
		size_t total = 1;
		for(uint32_t c = 0; c < nchr; c++) total += len[c] + 1;
		// sizes first (no value-initialising pass over recycled storage), then one parallel fill of all per-element arrays;
		// collapses that lengthen a branch append elements: room for them up front instead of whole-array reallocations
		const size_t room = total + total / 16 + 4096;
		ch.reserve(room);
		opos.reserve(room);
		nxt.reserve(room);
		prv.reserve(room);
		for(int s = 0; s < 2; s++)
		{
			mark[s].reserve(room);
			node_of[s].reserve(room);
		}
		ch.resize(total);
		opos.resize(total);
		nxt.resize(total);
		prv.resize(total);
		for(int s = 0; s < 2; s++)
		{
			mark[s].resize(total);
			node_of[s].resize(total);
			head[s].assign((size_t)count + 1, -1);
			lsize[s].assign((size_t)count + 1, 0);
		}
		dirty.assign((size_t)count + 1, 0);
		chr_first_sep.resize(nchr);
		std::vector<size_t> start(nchr);
		size_t at = 1;
		for(uint32_t c = 0; c < nchr; c++)
		{
			chr_first_sep[c] = (int32_t)(at - 1);
			start[c] = at;
			at += len[c] + 1;
		}
		ch[0] = SEP;
		opos[0] = 0;
		for(uint32_t c = 0; c < nchr; c++)
		{
			const size_t s0 = start[c];
			const char *sq = seq[c];
			const uint32_t *op = origpos[c];
			parallel_ranges(len[c], [&, s0, sq, op](size_t b, size_t e) {
				memcpy(&ch[s0 + b], sq + b, e - b);
				for(size_t i = b; i < e; i++) opos[s0 + i] = op[i] & POS_MASK;
			});
			ch[s0 + len[c]] = SEP;
			opos[s0 + len[c]] = (uint32_t)len[c] & POS_MASK;   // dnasequence.cpp:96-97
		}
		parallel_ranges(total, [&](size_t b, size_t e) {
			for(size_t i = b; i < e; i++)
			{
				nxt[i] = (int32_t)i + 1;
				prv[i] = (int32_t)i - 1;
				mark[0][i] = NO_BIF;
				mark[1][i] = NO_BIF;
				node_of[0][i] = -1;
				node_of[1][i] = -1;
			}
		});
		nxt[total - 1] = -1;
		last_sep = (int32_t)total - 1;
		live = total;
		// IndexedSequence::Init, indexedsequence.cpp:51-67: strand 0 then strand 1, chromosomes and positions ascending
		n_elem.reserve(npos + nneg);
		for(uint64_t i = 0; i < npos; i++)
		{
			It it;
			it.e = (int32_t)(start[pos[i].chr] + pos[i].pos);
			it.d = 0;
			add_point(it, pos[i].bifId);
		}
		for(uint64_t i = 0; i < nneg; i++)
		{
			It it;
			it.e = (int32_t)(start[neg[i].chr] + len[neg[i].chr] - 1 - neg[i].pos);
			it.d = 1;
			add_point(it, neg[i].bifId);
		}
		std::fill(dirty.begin(), dirty.end(), 0);
	}

	// =========================================================== DNASequence::Replace
	int32_t new_element(char c)
	{
		const int32_t e = (int32_t)ch.size();
		ch.push_back(c);
		opos.push_back(0);
		nxt.push_back(-1);
		prv.push_back(-1);
		for(int s = 0; s < 2; s++)
		{
			mark[s].push_back(NO_BIF);
			node_of[s].push_back(-1);
		}
		return e;
	}

	// ReplaceDirect, dnasequence.cpp:189-230.  `target` is the first (positive order) element of the target branch.
	void replace_direct(It source, size_t sdist, int32_t target, size_t tdist)
	{
		const int32_t save = target;
		const size_t first_pos = opos[save];
		int32_t after = save;
		for(size_t i = 0; i < tdist; i++) after = nxt[after];
		const size_t last_pos = opos[after];
		const size_t common = std::min(sdist, tdist);
		for(size_t i = 0; i < common; i++)
		{
			ch[target] = deref(source);
			target = nxt[target];
			inc(source);
		}
		if(sdist < tdist)
		{
			// erase the surplus elements [target, target + tdist - sdist)
			int32_t e = target;
			const int32_t before = prv[target];
			for(size_t i = 0; i < tdist - sdist; i++) e = nxt[e];
			nxt[before] = e;
			prv[e] = before;
			live -= tdist - sdist;
		}
		else if(sdist != tdist)
		{
			// insert the rest of the source in front of `target`
			std::string buf;
			for(size_t i = 0; i < sdist - tdist; i++, inc(source)) buf.push_back(deref(source));
			int32_t before = prv[target];
			for(size_t i = 0; i < buf.size(); i++)
			{
				const int32_t e = new_element(buf[i]);
				nxt[before] = e;
				prv[e] = before;
				before = e;
			}
			nxt[before] = target;
			prv[target] = before;
			live += sdist - tdist;
		}
		// original positions of the new branch: sequential double accumulation (:221-227)
		double acc = static_cast<double>(first_pos);
		const double ssize = double(tdist) / sdist;
		int32_t e = save;
		for(size_t step = 0; step < sdist; step++, e = nxt[e], acc += ssize)
		{
			const size_t p = std::min(last_pos, size_t(acc));
			opos[e] = static_cast<uint32_t>(p) & POS_MASK;   // SetOriginalPosition also clears both info bits; the
		}                                                    // branch carries no marks at this point
	}

	void replace(It source, size_t sdist, It target, size_t tdist)    // Replace, dnasequence.cpp:232-252
	{
		if(target.d == 0)
		{
			replace_direct(source, sdist, target.e, tdist);
		}
		else
		{
			source = invert(advance(source, sdist));
			const int32_t begin = invert(advance(target, tdist)).e;
			replace_direct(source, sdist, begin, tdist);
		}
	}

	// =========================================================== bulgeremoval.cpp
	size_t max_bifurcation_multiplicity(It it, size_t distance) const   // :39-53
	{
		size_t ret = 0;
		for(size_t i = 0; i + 1 < distance; i++)
		{
			inc(it);
			const uint32_t b = get_bif(it);
			if(b != NO_BIF) ret = std::max(ret, count_bifurcations(b));
		}
		return ret;
	}

	void erase_bifurcations(const std::vector<int32_t> &start_kmer, VisitData target,
		std::vector<std::pair<size_t, size_t> > &look_forward, std::vector<std::pair<size_t, size_t> > &look_back)   // :55-95
	{
		look_back.clear();
		look_forward.clear();
		const It t0 = node_it(start_kmer[target.kmerId]);
		It amer = invert(advance(t0, k));
		It bmer = advance(t0, target.distance);
		for(size_t i = 0; i < k; i++, inc(amer), inc(bmer))
		{
			uint32_t b = get_bif(amer);
			if(b != NO_BIF)
			{
				erase_point(amer);
				look_back.push_back(std::make_pair(i, (size_t)b));
			}
			b = get_bif(bmer);
			if(b != NO_BIF)
			{
				erase_point(bmer);
				look_forward.push_back(std::make_pair(i, (size_t)b));
			}
		}
		amer = t0;
		bmer = invert(advance(amer, k + target.distance));
		for(size_t i = 0; i < k + target.distance; i++, inc(amer), inc(bmer))
		{
			if(i > 0) erase_point(amer);
			erase_point(bmer);
		}
	}

	mutable std::vector<int32_t> occur;                  // scratch of overlap()
	bool overlap(const std::vector<int32_t> &start_kmer, VisitData source, VisitData target) const   // :97-120
	{
		occur.clear();
		It it = node_it(start_kmer[source.kmerId]);
		for(size_t i = 0; i < source.distance + k; i++, inc(it)) occur.push_back(it.e);
		it = node_it(start_kmer[target.kmerId]);
		std::sort(occur.begin(), occur.end());
		for(size_t i = 0; i < target.distance + k; i++, inc(it))
		{
			if(std::binary_search(occur.begin(), occur.end(), it.e)) return true;
		}
		return false;
	}

	void fill_visit(It kmer, std::vector<BifurcationMark> &visit) const   // :122-146
	{
		visit.clear();
		const uint32_t start = get_bif(kmer);
		inc(kmer);
		for(size_t step = 1; step < D && at_valid(kmer); inc(kmer), step++)
		{
			const uint32_t b = get_bif(kmer);
			if(b == start) break;
			if(b != NO_BIF)
			{
				BifurcationMark m;
				m.bifId = b;
				m.distance = step;
				visit.push_back(m);
			}
		}
		std::sort(visit.begin(), visit.end());
	}

	struct BranchData {
		char end_char;
		std::vector<size_t> branch_ids;
	};

	// AnyBulges, :158-218.  The visit map is boost::unordered_map<size_t, BranchData>: lookups through `slot_of`,
	// iteration order through BoostUnorderedOrder.
	std::vector<int32_t> slot_of;                        // vertex id -> index into `branches` (-1 = absent), kept sparse
	std::vector<BranchData> branches;                    // pooled: entries [0, n_branches) are live, the rest keep their capacity
	size_t n_branches = 0;
	std::vector<uint32_t> touched;
	BoostUnorderedOrder order;

	std::vector<char> first_char;                        // scratch of the existence pass

	// expect_bulge: the caller has just seen a conflict for this vertex (parallel screen): the existence pass, which
	// only exists to leave early, is skipped -- the full pass alone gives the same answer.
	bool any_bulges(const std::vector<int32_t> &start_kmer, const std::vector<char> &end_char,
		std::vector<std::vector<size_t> > &bulges, bool expect_bulge = false)
	{
		bulges.clear();
		// Existence pass.  Most calls (94 % on the reference's example genome) find nothing; the same walks without the
		// BranchData / Boost-order bookkeeping decide that: both loops perform the same insertions up to the first
		// conflict (a vertex reached again with another end character), so a conflict exists here iff the full pass
		// below produces a branch with two ids.
		if(!expect_bulge)
		{
			touched.clear();
			first_char.clear();
			bool conflict = false;
			for(size_t i = 0; i < start_kmer.size() && !conflict; i++)
			{
				if(end_char[i] == EMPTY) continue;
				It kmer = node_it(start_kmer[i]);
				const uint32_t start = get_bif(kmer);
				inc(kmer);
				for(size_t step = 1; step < D && at_valid(kmer); inc(kmer), step++)
				{
					const uint32_t b = get_bif(kmer);
					if(b == start) break;
					if(b != NO_BIF)
					{
						const int32_t sl = slot_of[b];
						if(sl < 0)
						{
							slot_of[b] = (int32_t)first_char.size();
							touched.push_back(b);
							first_char.push_back(end_char[i]);
						}
						else if(first_char[sl] != end_char[i])
						{
							conflict = true;
							break;
						}
					}
				}
			}
			for(size_t i = 0; i < touched.size(); i++) slot_of[touched[i]] = -1;
			if(!conflict) return false;
		}
		n_branches = 0;
		touched.clear();
		order.clear();
		for(size_t i = 0; i < start_kmer.size(); i++)
		{
			if(end_char[i] == EMPTY) continue;
			It kmer = node_it(start_kmer[i]);
			const uint32_t start = get_bif(kmer);
			inc(kmer);
			for(size_t step = 1; step < D && at_valid(kmer); inc(kmer), step++)
			{
				const uint32_t b = get_bif(kmer);
				if(b == start) break;
				if(b != NO_BIF)
				{
					const int32_t s = slot_of[b];
					if(s < 0)
					{
						slot_of[b] = (int32_t)n_branches;
						touched.push_back(b);
						order.insert_new(b, (int)n_branches);
						if(n_branches == branches.size()) branches.emplace_back();
						BranchData &bd = branches[n_branches++];
						bd.end_char = end_char[i];
						bd.branch_ids.clear();
						bd.branch_ids.push_back(i);
					}
					else if(branches[s].end_char != end_char[i])
					{
						branches[s].branch_ids.push_back(i);
						break;
					}
				}
			}
		}
		ord.clear();
		order.order(std::back_inserter(ord));
		for(size_t i = 0; i < ord.size(); i++)
		{
			if(branches[ord[i]].branch_ids.size() > 1) bulges.push_back(branches[ord[i]].branch_ids);
		}
		for(size_t i = 0; i < touched.size(); i++) slot_of[touched[i]] = -1;
		return !bulges.empty();
	}

	// ---- read-only screen of the dirty vertices at the start of a sweep (all host threads)
	struct ScreenScratch {
		std::vector<int32_t> slot_of, start_kmer;
		std::vector<uint32_t> touched;
		std::vector<char> first_char, end_char;
		std::vector<uint32_t> surv_id, surv_off;           // vertices that kept their flag + where their instances start below
		std::vector<int32_t> surv_elem;                    // element of every instance of those vertices
	};

	// Would RemoveBulges(id) find a bulge in the CURRENT state?  Same walks as the existence pass of any_bulges, on
	// scratch of the calling thread; touches no member.
	bool exists_bulge(size_t id, ScreenScratch &sc) const
	{
		list_positions(id, sc.start_kmer);
		if(sc.start_kmer.size() < 2) return false;
		sc.end_char.assign(sc.start_kmer.size(), EMPTY);
		for(size_t i = 0; i < sc.start_kmer.size(); i++)
		{
			const It it = node_it(sc.start_kmer[i]);
			if(proper_kmer(it, k + 1)) sc.end_char[i] = deref(advance(it, k));
		}
		sc.touched.clear();
		sc.first_char.clear();
		bool conflict = false;
		for(size_t i = 0; i < sc.start_kmer.size() && !conflict; i++)
		{
			if(sc.end_char[i] == EMPTY) continue;
			It kmer = node_it(sc.start_kmer[i]);
			const uint32_t start = get_bif(kmer);
			inc(kmer);
			for(size_t step = 1; step < D && at_valid(kmer); inc(kmer), step++)
			{
				const uint32_t b = get_bif(kmer);
				if(b == start) break;
				if(b != NO_BIF)
				{
					const int32_t sl = sc.slot_of[b];
					if(sl < 0)
					{
						sc.slot_of[b] = (int32_t)sc.first_char.size();
						sc.touched.push_back(b);
						sc.first_char.push_back(sc.end_char[i]);
					}
					else if(sc.first_char[sl] != sc.end_char[i])
					{
						conflict = true;
						break;
					}
				}
			}
		}
		for(size_t i = 0; i < sc.touched.size(); i++) sc.slot_of[sc.touched[i]] = -1;
		return conflict;
	}

	// Screens the vertices of [lo, hi) that are due for an exact call (flagged by the GPU or dirty) against the state as
	// it is NOW, with all host threads: a vertex whose RemoveBulges call would find nothing loses its flag and its dirty
	// mark -- it needs the exact call only if a later collapse dirties it again (mark_dirty_around / add_point /
	// erase_point do that).  Must not run concurrently with a mutation.  Returns the number of vertices screened
	// (0: fewer than screen_min candidates, not worth the threads).
	size_t screen_min = 2048;
	std::vector<ScreenScratch> screen_scratch;          // one per thread, kept across calls (slot_of is max_id + 1 entries)
	size_t screen_range(size_t lo, size_t hi, uint8_t *flag)
	{
		std::vector<uint32_t> ids;
		for(size_t id = lo; id < hi; id++)
		{
			if(dirty[id] || (flag && flag[id])) ids.push_back((uint32_t)id);
		}
		if(ids.size() < screen_min || ids.empty()) return 0;
		size_t nt = std::thread::hardware_concurrency();
		nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
		if(nt > (ids.size() + 255) / 256) nt = (ids.size() + 255) / 256;
		if(screen_scratch.size() < nt) screen_scratch.resize(nt);
		auto work = [this, &ids, flag, nt](size_t t) {
			ScreenScratch &sc = screen_scratch[t];
			if(sc.slot_of.size() != dirty.size()) sc.slot_of.assign(dirty.size(), -1);
			sc.surv_id.clear();
			sc.surv_off.clear();
			sc.surv_elem.clear();
			// interleaved blocks: ids are lexicographic ranks, not positions; the split only has to balance the work
			const size_t block = 256;
			for(size_t b = t * block; b < ids.size(); b += nt * block)
			{
				const size_t e = std::min(b + block, ids.size());
				for(size_t i = b; i < e; i++)
				{
					if(!exists_bulge(ids[i], sc))
					{
						dirty[ids[i]] = 0;
						if(flag) flag[ids[i]] = 0;
					}
					else
					{
						sc.surv_id.push_back(ids[i]);
						sc.surv_off.push_back((uint32_t)sc.surv_elem.size());
						for(size_t j = 0; j < sc.start_kmer.size(); j++) sc.surv_elem.push_back(n_elem[sc.start_kmer[j]]);
					}
				}
			}
			sc.surv_off.push_back((uint32_t)sc.surv_elem.size());
		};
		std::vector<std::thread> th;
		for(size_t t = 1; t < nt; t++) th.emplace_back(work, t);
		work(0);
		for(std::thread &x : th) x.join();
		// the survivors in id order, for the run-ahead prefetchers of the ordered part
		ahead_id.clear();
		ahead_off.clear();
		ahead_elem.clear();
		std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t> > > order;      // id -> (thread, index)
		for(size_t t = 0; t < nt; t++)
		{
			for(size_t i = 0; i < screen_scratch[t].surv_id.size(); i++)
			{
				order.push_back(std::make_pair(screen_scratch[t].surv_id[i], std::make_pair((uint32_t)t, (uint32_t)i)));
			}
		}
		std::sort(order.begin(), order.end());
		for(size_t i = 0; i < order.size(); i++)
		{
			const ScreenScratch &sc = screen_scratch[order[i].second.first];
			const uint32_t j = order[i].second.second;
			ahead_id.push_back(order[i].first);
			ahead_off.push_back((uint32_t)ahead_elem.size());
			ahead_elem.insert(ahead_elem.end(), sc.surv_elem.begin() + sc.surv_off[j], sc.surv_elem.begin() + sc.surv_off[j + 1]);
		}
		ahead_off.push_back((uint32_t)ahead_elem.size());
		return ids.size();
	}

	// ---- run-ahead prefetchers of the ordered part.  The exact calls of a chunk jump between loci of a working set far
	// beyond the last-level cache (29 B per element), and what the screen touched is long evicted when the ordered loop
	// gets there.  While the loop works on vertex i, helper threads issue prefetches for the neighbourhoods of the
	// instances of the survivors a few vertices ahead.  They execute nothing but prefetch instructions on addresses
	// computed from the survivor list (immutable during the loop) and the array bases as of ahead_start(): no load from a
	// structure the ordered loop mutates, and a prefetch of a stale address cannot fault.  Results do not depend on them.
	std::vector<uint32_t> ahead_id, ahead_off;
	std::vector<int32_t> ahead_elem;
	std::atomic<uint32_t> ahead_pos{0};                  // vertex id the ordered loop has reached
	std::atomic<bool> ahead_quit{false};
	std::vector<std::thread> ahead_threads;
	size_t ahead_helpers = 3, ahead_lead = 24;

	void ahead_start()
	{
		ahead_stop();
		if(ahead_id.size() < 64 || ahead_helpers == 0) return;
		ahead_quit.store(false);
		ahead_pos.store(ahead_id[0]);
		const char *b_ch = ch.data();
		const char *b4[7] = {reinterpret_cast<const char*>(opos.data()), reinterpret_cast<const char*>(nxt.data()),
			reinterpret_cast<const char*>(prv.data()), reinterpret_cast<const char*>(mark[0].data()),
			reinterpret_cast<const char*>(mark[1].data()), reinterpret_cast<const char*>(node_of[0].data()),
			reinterpret_cast<const char*>(node_of[1].data())};
		const int64_t n_el = (int64_t)ch.size();
		const int64_t reach = (int64_t)(D + 2 * k + 16);
		const size_t H = ahead_helpers, lead = ahead_lead;
		for(size_t h = 0; h < H; h++)
		{
			ahead_threads.emplace_back([this, h, H, lead, b_ch, b4, n_el, reach]() {
				const char *base4[7];
				for(int a = 0; a < 7; a++) base4[a] = b4[a];
				for(size_t j = h; j < ahead_id.size(); j += H)
				{
					if(j >= lead)
					{
						const uint32_t gate = ahead_id[j - lead];
						while(ahead_pos.load(std::memory_order_relaxed) < gate)
						{
							if(ahead_quit.load(std::memory_order_relaxed)) return;
							std::this_thread::yield();
						}
					}
					for(uint32_t x = ahead_off[j]; x < ahead_off[j + 1]; x++)
					{
						const int64_t e = ahead_elem[x];
						const int64_t lo = e - reach < 0 ? 0 : e - reach, hi = e + reach + 1 > n_el ? n_el : e + reach + 1;
						for(int64_t a = lo & ~int64_t(63); a < hi; a += 64) __builtin_prefetch(b_ch + a, 0, 2);
						for(int arr = 0; arr < 7; arr++)
						{
							for(int64_t a = (lo * 4) & ~int64_t(63); a < hi * 4; a += 64) __builtin_prefetch(base4[arr] + a, 0, 2);
						}
					}
				}
			});
		}
	}
	~Simplifier() { ahead_stop(); }
	void ahead_stop()
	{
		ahead_quit.store(true);
		for(std::thread &x : ahead_threads) x.join();
		ahead_threads.clear();
	}

	void update_bifurcations(const std::vector<int32_t> &start_kmer, VisitData source, VisitData target,
		const std::vector<std::pair<size_t, size_t> > &look_forward, const std::vector<std::pair<size_t, size_t> > &look_back)   // :238-282
	{
		size_t anear = 0, bnear = 0;
		const It t0 = node_it(start_kmer[target.kmerId]);
		const It s0 = node_it(start_kmer[source.kmerId]);
		It amer = invert(advance(t0, k));
		It bmer = advance(t0, source.distance);
		for(size_t i = 0; i < k; i++, inc(amer), inc(bmer))
		{
			if(anear < look_back.size() && i == look_back[anear].first) add_point(amer, look_back[anear++].second);
			if(bnear < look_forward.size() && i == look_forward[bnear].first) add_point(bmer, look_forward[bnear++].second);
		}
		amer = t0;
		bmer = invert(advance(t0, source.distance + k));
		It src_a = s0;
		It src_b = invert(advance(s0, source.distance + k));
		for(size_t i = 0; i < source.distance + 1; i++, inc(amer), inc(bmer), inc(src_a), inc(src_b))
		{
			uint32_t b = get_bif(src_a);
			if(b != NO_BIF) add_point(amer, b);
			b = get_bif(src_b);
			if(b != NO_BIF) add_point(bmer, b);
		}
	}

	// Everything a collapse touched, in terms of the vertices whose detection walks could see it: the elements of the
	// rewritten region plus a margin of `reach` elements on both sides; the vertex of every mark found there is dirty.
	void mark_dirty_around(It t0, size_t new_distance)
	{
		const size_t reach = std::max(D, k + 1) + 1;
		const uint32_t *m0 = mark[0].data(), *m1 = mark[1].data();
		const int32_t *nx = nxt.data(), *pv = prv.data();
		uint8_t *dr = dirty.data();
		// marks are sparse: one test decides the common element that carries none (NO_BIF is all ones)
		auto both = [&](int32_t e) {
			const uint32_t v0 = m0[e], v1 = m1[e];
			if((v0 & v1) != NO_BIF)
			{
				if(v0 != NO_BIF) dr[v0] = 1;
				if(v1 != NO_BIF) dr[v1] = 1;
			}
		};
		// The region spans offsets [0, new_distance + 2k) from the target's start in its strand direction (towards smaller
		// positive positions for a negative-strand target); [lo, hi] are its end points in positive order.  One walk finds
		// them and flags the marks of both strands on the way.
		int32_t lo = t0.e, hi = t0.e;
		both(t0.e);
		if(t0.d == 1)
		{
			for(size_t i = 0; i < new_distance + 2 * k && pv[lo] >= 0; i++)
			{
				lo = pv[lo];
				both(lo);
			}
		}
		else
		{
			for(size_t i = 0; i < new_distance + 2 * k && nx[hi] >= 0; i++)
			{
				hi = nx[hi];
				both(hi);
			}
		}
		// A positive-strand instance at x reads the elements [x, x + reach], a negative-strand one [x - reach, x]: outside
		// the region only positive marks before it and negative marks after it can see it.
		int32_t e = lo;
		for(size_t i = 0; i < reach && pv[e] >= 0; i++)
		{
			e = pv[e];
			const uint32_t v = m0[e];
			if(v != NO_BIF) dr[v] = 1;
		}
		e = hi;
		for(size_t i = 0; i < reach && nx[e] >= 0; i++)
		{
			e = nx[e];
			const uint32_t v = m1[e];
			if(v != NO_BIF) dr[v] = 1;
		}
	}

	std::vector<std::pair<size_t, size_t> > look_forward, look_back;   // scratch of collapse_bulge_greedily()
	void collapse_bulge_greedily(std::vector<int32_t> &start_kmer, VisitData source, VisitData target)   // :284-327
	{
		const It t0 = node_it(start_kmer[target.kmerId]);
		{
			SIB_PROF(5);
			erase_bifurcations(start_kmer, target, look_forward, look_back);
		}
		const It source_it = node_it(start_kmer[source.kmerId]);
		{
			SIB_PROF(6);
			replace(advance(source_it, k), source.distance, advance(t0, k), target.distance);
		}
		{
			SIB_PROF(7);
			update_bifurcations(start_kmer, source, target, look_forward, look_back);
		}
		SIB_PROF(8);                                     // mark_dirty_around
		// Vertices whose walks can see the rewritten region: marks that sat INSIDE the region were erased or re-added
		// above (erase_point / add_point flag their vertices); everything else within reach still carries its mark and
		// is found by one pass over the region's surroundings (the region's end points are elements that did not move).
		mark_dirty_around(t0, source.distance);
		collapses++;
	}

	size_t collapses = 0;

	std::vector<int32_t> start_kmer;                     // scratch of remove_bulges()
	std::vector<char> end_char;
	std::vector<std::vector<size_t> > bulges;
	std::vector<BifurcationMark> visit;
	std::vector<int> ord;

	static const int PREFETCH_LINES = 10;                // 10 x 16 elements ~ the 150-element reach of the first stage
	size_t remove_bulges(size_t bif_id, bool expect_bulge = false)   // RemoveBulges, :330-430
	{
		size_t ret = 0;
		SIB_PROF(0);                                     // whole call
		list_positions(bif_id, start_kmer);
		if(start_kmer.size() < 2) return ret;
		// The instances of a vertex lie far apart (different strains): start all their cache misses at once.  The walks
		// below read ch / nxt|prv / mark[d] of up to D elements after each instance; elements are mostly contiguous in
		// index, so the lines ahead of the start element are the ones that will be needed (prefetches have no effect on
		// the result).
		for(size_t i = 0; i < start_kmer.size(); i++)
		{
			const It it = node_it(start_kmer[i]);
			const int32_t step = it.d == 0 ? 16 : -16;
			const int32_t *link = it.d == 0 ? nxt.data() : prv.data();
			const int32_t limit = (int32_t)ch.size();
			for(int32_t j = 0, e = it.e; j < PREFETCH_LINES && e >= 0 && e < limit; j++, e += step)
			{
				__builtin_prefetch(link + e);
				__builtin_prefetch(mark[it.d].data() + e);
				if((j & 3) == 0) __builtin_prefetch(ch.data() + e);
			}
		}
		end_char.assign(start_kmer.size(), EMPTY);
		for(size_t i = 0; i < start_kmer.size(); i++)
		{
			const It it = node_it(start_kmer[i]);
			if(proper_kmer(it, k + 1)) end_char[i] = deref(advance(it, k));
		}
		{
			SIB_PROF(1);
			if(!any_bulges(start_kmer, end_char, bulges, expect_bulge)) return ret;
		}
		for(size_t num_bulge = 0; num_bulge < bulges.size(); ++num_bulge)
		{
			for(size_t id_i = 0; id_i < bulges[num_bulge].size(); ++id_i)
			{
				const size_t kmer_i = bulges[num_bulge][id_i];
				if(!n_valid[start_kmer[kmer_i]]) continue;
				{
					SIB_PROF(2);
					fill_visit(node_it(start_kmer[kmer_i]), visit);
				}
				for(size_t id_j = id_i + 1; id_j < bulges[num_bulge].size(); ++id_j)
				{
					const size_t kmer_j = bulges[num_bulge][id_j];
					if(!n_valid[start_kmer[kmer_j]] || end_char[kmer_i] == end_char[kmer_j]) continue;
					It kmer = node_it(start_kmer[kmer_j]);
					inc(kmer);
					for(size_t step = 1; at_valid(kmer) && step < D; inc(kmer), step++)
					{
						const uint32_t now_bif = get_bif(kmer);
						if(now_bif == NO_BIF) continue;
						if(now_bif == bif_id) break;
						BifurcationMark probe;
						probe.bifId = now_bif;
						probe.distance = 0;
						std::vector<BifurcationMark>::iterator vt = std::lower_bound(visit.begin(), visit.end(), probe);
						if(vt != visit.end() && vt->bifId == now_bif)
						{
							VisitData jdata, idata;
							jdata.kmerId = kmer_j;
							jdata.distance = step;
							idata.kmerId = kmer_i;
							idata.distance = vt->distance;
							{
								SIB_PROF(3);
								if(overlap(start_kmer, idata, jdata) || now_bif == bif_id) break;
							}
							++ret;
							size_t imlp, jmlp;
							{
								SIB_PROF(4);
								imlp = max_bifurcation_multiplicity(node_it(start_kmer[kmer_i]), idata.distance);
								jmlp = max_bifurcation_multiplicity(node_it(start_kmer[kmer_j]), jdata.distance);
							}
							const bool iless = imlp > jmlp || (imlp == jmlp && idata.kmerId < jdata.kmerId);
							if(iless)
							{
								end_char[jdata.kmerId] = end_char[idata.kmerId];
								collapse_bulge_greedily(start_kmer, idata, jdata);
							}
							else
							{
								end_char[idata.kmerId] = end_char[jdata.kmerId];
								collapse_bulge_greedily(start_kmer, jdata, idata);
								fill_visit(node_it(start_kmer[kmer_i]), visit);
							}
							break;
						}
					}
				}
			}
		}
		{
			SIB_PROF(9);
			cleanup();
		}
		return ret;
	}
};

} // namespace simp
} // namespace sibgpu
