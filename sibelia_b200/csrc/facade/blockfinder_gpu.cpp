// Drop-in replacement for the reference translation units src/blockfinder.cpp and src/bulgeremoval.cpp.
//
// Compiled against the reference's own headers and linked INSTEAD of those two units.  It defines the BlockFinder
// members the rest of the reference calls (constructors, src/blockfinder.cpp:53-76, and PerformGraphSimplifications,
// :78-98); SimplifyGraph / RemoveBulges / CollapseBulgeGreedily / UpdateBifurcations (private, only reachable from
// PerformGraphSimplifications) now run inside sibgpu_simplify.
#include "blockfinder.h"

#include "gpu_session.h"

namespace SyntenyFinder
{
	namespace
	{
		struct ProgressThunk
		{
			BlockFinder::ProgressCallBack f;
		};

		void CallProgress(size_t done, int state, void * user)
		{
			ProgressThunk * thunk = static_cast<ProgressThunk*>(user);
			if(!thunk->f.empty())
			{
				thunk->f(done, static_cast<BlockFinder::State>(state));
			}
		}
	}

	BlockFinder::BlockFinder(const std::vector<FASTARecord> & chrList):
		originalChrList_(&chrList)
	{
		Init(chrList);
	}

	BlockFinder::BlockFinder(const std::vector<FASTARecord> & chrList, const std::string & tempDir):
		originalChrList_(&chrList), tempDir_(tempDir)
	{
		Init(chrList);
	}

	void BlockFinder::Init(const std::vector<FASTARecord> & chrList)
	{
		rawSeq_.resize(chrList.size());
		originalPos_.resize(chrList.size());
		originalSize_.clear();
		for(size_t chr = 0; chr < chrList.size(); chr++)
		{
			const std::string & sequence = chrList[chr].GetSequence();
			rawSeq_[chr] = sequence;
			originalSize_.push_back(sequence.size());
			originalPos_[chr].resize(sequence.size());
			for(size_t pos = 0; pos < sequence.size(); pos++)
			{
				originalPos_[chr][pos] = static_cast<Pos>(pos);
			}
		}
	}

	size_t BlockFinder::PerformGraphSimplifications(size_t k, size_t minBranchSize, size_t maxIterations, ProgressCallBack f)
	{
		// IndexedSequence::Init sanitises a copy of the record and builds the DNASequence from that copy
		// (src/indexedsequence.cpp:31-37,50), so the replacement characters end up in rawSeq_ after the stage.  The
		// rand() stream is process-wide state and is consumed here in the reference's order.
		const size_t chrNumber = rawSeq_.size();
		std::vector<char*> seq(chrNumber);
		std::vector<uint32_t*> pos(chrNumber);
		std::vector<uint64_t> len(chrNumber);
		for(size_t chr = 0; chr < chrNumber; chr++)
		{
			std::string & record = rawSeq_[chr];
			for(size_t j = 0; j < record.size(); j++)
			{
				record[j] = IsDefiniteBase(record[j]) ? record[j] : DEFINITE_BASE[rand() % DEFINITE_BASE.size()];
			}

			seq[chr] = record.empty() ? 0 : &record[0];
			pos[chr] = originalPos_[chr].empty() ? 0 : &originalPos_[chr][0];
			len[chr] = record.size();
		}

		// IndexedSequence(rawSeq_, originalPos_, k, tempDir_, true): sanitise, then (file-backed variant only) the temp files
		ConsumeTempFileSideEffects(tempDir_);
		uint64_t bulges = 0;
		ProgressThunk thunk;
		thunk.f = f;
		GpuCheck(sibgpu_simplify(GpuSession(), chrNumber ? &seq[0] : 0, chrNumber ? &pos[0] : 0, chrNumber ? &len[0] : 0,
			static_cast<uint32_t>(chrNumber), static_cast<uint32_t>(k), static_cast<uint32_t>(minBranchSize),
			static_cast<uint32_t>(maxIterations), CallProgress, &thunk, &bulges));
		for(size_t chr = 0; chr < chrNumber; chr++)
		{
			// a stage without any bulge leaves the arrays as they were handed in (include/sibgpu.h)
			if(seq[chr] == (rawSeq_[chr].empty() ? 0 : &rawSeq_[chr][0]))
			{
				continue;
			}

			rawSeq_[chr].assign(seq[chr], seq[chr] + len[chr]);
			originalPos_[chr].assign(pos[chr], pos[chr] + len[chr]);
			sibgpu_free(seq[chr]);
			sibgpu_free(pos[chr]);
		}

		return static_cast<size_t>(bulges);
	}
}
