// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product
// (sibelia_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the library this file is built into.
//
// Thin extern "C" shim over the UNMODIFIED reference sources under /root/reference/src.
// It is compiled by oracle/Makefile together with the reference's own translation units
// (indexedsequence.cpp, vertexenumeration.cpp, bifurcationstorage.cpp, dnasequence.cpp,
// stranditerator.cpp, blockfinder.cpp, bulgeremoval.cpp, ... + libdivsufsort) into
// oracle/_ref/libsibelia_ref.so.  No reference source is copied into this repository; this
// file only *calls* the reference's public classes:
//   IndexedSequence ctor            /root/reference/src/indexedsequence.h:29-30
//   BifurcationStorage::ListPositions / GetBifurcation   src/bifurcationstorage.h:39,59-72
//   BlockFinder::PerformGraphSimplifications             src/blockfinder.h:45
//   BlockFinder::ListEdges (private)                     src/blockfinder.h:112, src/serialization.cpp:56-86
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <unordered_map>

#include "common.h"
// rawSeq_/originalPos_ are BlockFinder's inter-stage state (src/blockfinder.h:52-54); the probe
// must read and seed them, so private members of the headers below are opened for this TU only.
#define private public
#include "blockfinder.h"
#undef private

using namespace SyntenyFinder;

const std::string VERSION("3.0.7-oracle");   // src/sibelia.cpp:11 is not linked into the shim
const std::string DELIMITER(80, '-');     // src/util.cpp:9 likewise

extern "C" {

struct ref_inst { uint32_t bifId, chr, pos; };

static std::string g_err;
const char* ref_last_error() { return g_err.c_str(); }
void ref_free(void* p) { free(p); }

static double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Builds IndexedSequence(record, k, "") -- the in-RAM path (src/indexedsequence.cpp:40-43) -- and dumps
//   * maxId                              (BifurcationStorage::GetMaxId)
//   * per-strand (bifId, chr, pos) lists in (chr, pos) order, recovered by walking both strands and asking
//     GetBifurcation at every element exactly like IndexedSequence::Init's AddPoint loop (:51-67)
//   * ListPositions(id) order for every id as CSR: lp_off[maxId+2]; lp_gidx = DNASequence::GlobalIndex of the
//     instance's base element, lp_strand = 0 (+) / 1 (-)
// If want_dump == 0 only the constructor is timed (seconds) and nothing is dumped.
int ref_index(uint32_t nchr, const char* const* chr, const uint64_t* len, uint32_t k, int want_dump,
	uint32_t* maxId, ref_inst** pos, uint64_t* npos, ref_inst** neg, uint64_t* nneg,
	uint64_t** lp_off, uint32_t** lp_gidx, uint8_t** lp_strand, double* seconds)
{
	try
	{
		std::vector<std::string> record(nchr);
		for(uint32_t i = 0; i < nchr; i++)
		{
			record[i].assign(chr[i], chr[i] + len[i]);
		}

		double t0 = now_s();
		IndexedSequence iseq(record, k, "");
		double t1 = now_s();
		if(seconds) *seconds = t1 - t0;
		const DNASequence & seq = iseq.Sequence();
		BifurcationStorage & bif = iseq.BifStorage();
		if(maxId) *maxId = static_cast<uint32_t>(bif.GetMaxId());
		if(!want_dump)
		{
			return 0;
		}

		std::unordered_map<size_t, uint32_t> gidx;
		{
			uint32_t g = 0;
			for(size_t c = 0; c < seq.ChrNumber(); c++)
			{
				StrandIterator end = seq.PositiveEnd(c);
				for(StrandIterator it = seq.PositiveBegin(c); it != end; ++it, ++g)
				{
					gidx[it.GetElementId()] = g;
				}
				gidx[end.GetElementId()] = g++;   // the '$' after chr c
			}
		}

		std::vector<ref_inst> out[2];
		for(size_t strand = 0; strand < 2; strand++)
		{
			DNASequence::Direction dir = static_cast<DNASequence::Direction>(strand);
			for(size_t c = 0; c < seq.ChrNumber(); c++)
			{
				uint32_t p = 0;
				StrandIterator end = seq.End(dir, c);
				for(StrandIterator it = seq.Begin(dir, c); it != end; ++it, ++p)
				{
					size_t id = bif.GetBifurcation(it);
					if(id != BifurcationStorage::NO_BIFURCATION)
					{
						ref_inst r = {static_cast<uint32_t>(id), static_cast<uint32_t>(c), p};
						out[strand].push_back(r);
					}
				}
			}
		}

		ref_inst** dst[2] = {pos, neg};
		uint64_t* ndst[2] = {npos, nneg};
		for(int s = 0; s < 2; s++)
		{
			*ndst[s] = out[s].size();
			*dst[s] = static_cast<ref_inst*>(malloc(sizeof(ref_inst) * (out[s].size() + 1)));
			memcpy(*dst[s], out[s].data(), sizeof(ref_inst) * out[s].size());
		}

		size_t ids = bif.GetMaxId() + 1;
		std::vector<uint64_t> off(ids + 1, 0);
		std::vector<uint32_t> lg;
		std::vector<uint8_t> ls;
		for(size_t id = 0; id < ids; id++)
		{
			IteratorProxyVector v;
			bif.ListPositions(id, std::back_inserter(v));
			for(size_t i = 0; i < v.size(); i++)
			{
				StrandIterator it = *v[i];
				lg.push_back(gidx[it.GetElementId()]);
				ls.push_back(it.GetDirection() == DNASequence::positive ? 0 : 1);
			}
			off[id + 1] = lg.size();
		}

		*lp_off = static_cast<uint64_t*>(malloc(sizeof(uint64_t) * off.size()));
		memcpy(*lp_off, off.data(), sizeof(uint64_t) * off.size());
		*lp_gidx = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (lg.size() + 1)));
		memcpy(*lp_gidx, lg.data(), sizeof(uint32_t) * lg.size());
		*lp_strand = static_cast<uint8_t*>(malloc(ls.size() + 1));
		memcpy(*lp_strand, ls.data(), ls.size());
		return 0;
	}
	catch(std::exception & e)
	{
		g_err = e.what();
		return 1;
	}
}

// Iteration order of the vendored boost::unordered_map<size_t, ...> (Boost 1.54) after inserting the given DISTINCT
// keys with operator[] in the given order -- the container AnyBulges relies on (src/bulgeremoval.cpp:168,203-215).
void ref_boost_order(const uint64_t* keys, uint64_t n, uint64_t* out)
{
	boost::unordered_map<size_t, int> m;
	for(uint64_t i = 0; i < n; i++)
	{
		m[static_cast<size_t>(keys[i])] = static_cast<int>(i);
	}
	uint64_t j = 0;
	for(boost::unordered_map<size_t, int>::iterator it = m.begin(); it != m.end(); ++it)
	{
		out[j++] = it->first;
	}
}

// One stage of BlockFinder::PerformGraphSimplifications(k, D, iters) (src/blockfinder.cpp:78-98) seeded with an
// arbitrary inter-stage state (rawSeq_, originalPos_).  seq/origpos are replaced by malloc'ed outputs.
int ref_simplify(uint32_t nchr, char** seq, uint32_t** origpos, uint64_t* len,
	uint32_t k, uint32_t D, uint32_t iters, uint64_t* bulges, double* seconds)
{
	try
	{
		std::vector<FASTARecord> chrList;
		for(uint32_t i = 0; i < nchr; i++)
		{
			chrList.push_back(FASTARecord(std::string(seq[i], seq[i] + len[i]), "chr", i));
		}

		BlockFinder finder(chrList);
		for(uint32_t i = 0; i < nchr; i++)
		{
			finder.originalPos_[i].assign(origpos[i], origpos[i] + len[i]);
		}

		double t0 = now_s();
		size_t ret = finder.PerformGraphSimplifications(k, D, iters);
		double t1 = now_s();
		if(seconds) *seconds = t1 - t0;
		*bulges = ret;
		for(uint32_t i = 0; i < nchr; i++)
		{
			const std::string & s = finder.rawSeq_[i];
			len[i] = s.size();
			seq[i] = static_cast<char*>(malloc(s.size() + 1));
			memcpy(seq[i], s.data(), s.size());
			origpos[i] = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * (s.size() + 1)));
			memcpy(origpos[i], finder.originalPos_[i].data(), sizeof(uint32_t) * s.size());
		}

		return 0;
	}
	catch(std::exception & e)
	{
		g_err = e.what();
		return 1;
	}
}


// The edge list GenerateSyntenyBlocks starts from (src/synteny.cpp:238-241): IndexedSequence(rawSeq_, originalPos_, k, "")
// followed by BlockFinder::ListEdges, for an arbitrary inter-stage state.  9 x uint32 per edge.
struct ref_edge { uint32_t chr, direction, startVertex, endVertex, actualPosition, actualLength, originalPosition, originalLength, firstChar; };
int ref_list_edges(uint32_t nchr, const char* const* seq, const uint32_t* const* origpos, const uint64_t* len, uint32_t k,
	ref_edge** edges, uint64_t* nedges, double* seconds)
{
	try
	{
		std::vector<FASTARecord> chrList;
		for(uint32_t i = 0; i < nchr; i++)
		{
			chrList.push_back(FASTARecord(std::string(seq[i], seq[i] + len[i]), "chr", i));
		}

		BlockFinder finder(chrList);
		for(uint32_t i = 0; i < nchr; i++)
		{
			finder.originalPos_[i].assign(origpos[i], origpos[i] + len[i]);
		}

		std::vector<BlockFinder::Edge> edge;
		double t0 = now_s();
		{
			IndexedSequence iseq(finder.rawSeq_, finder.originalPos_, k, "");
			finder.ListEdges(iseq.Sequence(), iseq.BifStorage(), k, edge);
		}
		double t1 = now_s();
		if(seconds) *seconds = t1 - t0;
		*nedges = edge.size();
		*edges = static_cast<ref_edge*>(malloc(sizeof(ref_edge) * (edge.size() + 1)));
		for(size_t i = 0; i < edge.size(); i++)
		{
			ref_edge e = {static_cast<uint32_t>(edge[i].GetChr()), edge[i].GetDirection() == DNASequence::positive ? 0u : 1u,
				static_cast<uint32_t>(edge[i].GetStartVertex()), static_cast<uint32_t>(edge[i].GetEndVertex()),
				static_cast<uint32_t>(edge[i].GetActualPosition()), static_cast<uint32_t>(edge[i].GetActualLength()),
				static_cast<uint32_t>(edge[i].GetOriginalPosition()), static_cast<uint32_t>(edge[i].GetOriginalLength()),
				static_cast<uint32_t>(static_cast<unsigned char>(edge[i].GetFirstChar()))};
			(*edges)[i] = e;
		}

		return 0;
	}
	catch(std::exception & e)
	{
		g_err = e.what();
		return 1;
	}
}


// BlockFinder::TrimBlocks (src/synteny.cpp:31-122) on a block whose sequences are whole "chromosomes": block[i] =
// Edge(chr i, direction dir[i], original position 0, original length len[i]).  Returns the surviving edges as
// (chr, originalPosition, originalLength) triples and the function's return value (drop).
int ref_trim_blocks(uint32_t nchr, const char* const* seq, const uint64_t* len, const uint8_t* dir, uint32_t trimK, uint32_t minSize,
	uint32_t* out_triples, uint32_t* nout, int* drop)
{
	try
	{
		std::vector<FASTARecord> chrList;
		for(uint32_t i = 0; i < nchr; i++)
		{
			chrList.push_back(FASTARecord(std::string(seq[i], seq[i] + len[i]), "chr", i));
		}

		BlockFinder finder(chrList);
		std::vector<BlockFinder::Edge> block;
		for(uint32_t i = 0; i < nchr; i++)
		{
			block.push_back(BlockFinder::Edge(i, dir[i] ? DNASequence::negative : DNASequence::positive, 0, 0, 0, 0, 0, len[i], 'A'));
		}

		*drop = finder.TrimBlocks(block, trimK, minSize) ? 1 : 0;
		*nout = static_cast<uint32_t>(block.size());
		for(size_t i = 0; i < block.size(); i++)
		{
			out_triples[3 * i] = static_cast<uint32_t>(block[i].GetChr());
			out_triples[3 * i + 1] = static_cast<uint32_t>(block[i].GetOriginalPosition());
			out_triples[3 * i + 2] = static_cast<uint32_t>(block[i].GetOriginalLength());
		}

		return 0;
	}
	catch(std::exception & e)
	{
		g_err = e.what();
		return 1;
	}
}


// FASTAReader(path).GetSequences (src/fasta.cpp:22-73).  Returns 0 and the records packed as
//   names: nrec NUL-terminated descriptions back to back, seqs: the sequences back to back, lens[nrec];
// or 1 with the exception text in ref_last_error() ("parse error in <path> on line N: what").
int ref_fasta_parse(const char* path, uint32_t* nrec, char** names, uint64_t* names_bytes, char** seqs, uint64_t** lens, double* seconds)
{
	try
	{
		const double t0 = now_s();
		FASTAReader reader(path);
		if(!reader.IsOk()) throw std::runtime_error(std::string("Cannot open file ") + path);
		std::vector<FASTARecord> rec;
		reader.GetSequences(rec);
		*seconds = now_s() - t0;
		uint64_t nb = 0, sb = 0;
		for(size_t i = 0; i < rec.size(); i++)
		{
			nb += rec[i].GetDescription().size() + 1;
			sb += rec[i].GetSequence().size();
		}
		*nrec = static_cast<uint32_t>(rec.size());
		*names = static_cast<char*>(malloc(nb + 1));
		*seqs = static_cast<char*>(malloc(sb + 1));
		*lens = static_cast<uint64_t*>(malloc(sizeof(uint64_t) * (rec.size() + 1)));
		*names_bytes = nb;
		uint64_t na = 0, sa = 0;
		for(size_t i = 0; i < rec.size(); i++)
		{
			memcpy(*names + na, rec[i].GetDescription().c_str(), rec[i].GetDescription().size() + 1);
			na += rec[i].GetDescription().size() + 1;
			memcpy(*seqs + sa, rec[i].GetSequence().data(), rec[i].GetSequence().size());
			sa += rec[i].GetSequence().size();
			(*lens)[i] = rec[i].GetSequence().size();
		}
		return 0;
	}
	catch(const std::exception & e)
	{
		g_err = e.what();
		return 1;
	}
}
}
