// Drop-in replacement for the reference translation unit src/vertexenumeration.cpp.
//
// It is compiled AGAINST THE REFERENCE'S OWN HEADERS (-I /root/reference/src) and linked INSTEAD of
// vertexenumeration.cpp (libdivsufsort is then no longer linked at all).  It defines exactly what that unit defines:
//   IndexedSequence::SEPARATION_CHAR                                  src/vertexenumeration.cpp:11
//   IndexedSequence::EnumerateBifurcationsSArrayInRAM(...)            src/vertexenumeration.cpp:263-364
//   IndexedSequence::EnumerateBifurcationsSArray(...)                 src/vertexenumeration.cpp:160-261
// Both enumerators return the same tables in the reference (one keeps the suffix array in RAM, the other streams it
// through temporary files); here both call sibgpu_enumerate.  The file-backed twin's temporary file names consume
// 24 values of the glibc rand() stream per index (src/platform.cpp:55-58); this replacement does not, i.e. it behaves
// like the reference run with --inram (only observable for inputs that contain non-ACGT characters).
#include "indexedsequence.h"

#include "gpu_session.h"

namespace SyntenyFinder
{
	const char IndexedSequence::SEPARATION_CHAR = '#';

	sibgpu_ctx * GpuSession()
	{
		static sibgpu_ctx * ctx = 0;
		if(ctx == 0)
		{
			const char * dev = getenv("SIBELIA_GPU");
			GpuCheck(sibgpu_create(dev ? atoi(dev) : 0, &ctx));
		}

		return ctx;
	}

	size_t IndexedSequence::EnumerateBifurcationsSArrayInRAM(const std::vector<std::string> & data, std::vector<BifurcationInstance> & positiveBif, std::vector<BifurcationInstance> & negativeBif)
	{
		std::vector<const char*> chr(data.size());
		std::vector<uint64_t> len(data.size());
		for(size_t i = 0; i < data.size(); i++)
		{
			chr[i] = data[i].data();
			len[i] = data[i].size();
		}

		uint32_t count = 0;
		uint64_t size[2] = {0, 0};
		sibgpu_inst * table[2] = {0, 0};
		GpuCheck(sibgpu_enumerate(GpuSession(), chr.empty() ? 0 : &chr[0], len.empty() ? 0 : &len[0], static_cast<uint32_t>(data.size()),
			static_cast<uint32_t>(k_), &table[0], &size[0], &table[1], &size[1], &count));
		std::vector<BifurcationInstance> * ret[] = {&positiveBif, &negativeBif};
		for(size_t strand = 0; strand < 2; strand++)
		{
			// BifurcationInstance (src/indexedsequence.h:57-68) and sibgpu_inst are the same three 32-bit fields: one bulk copy
			static_assert(sizeof(BifurcationInstance) == sizeof(sibgpu_inst), "BifurcationInstance layout");
			const BifurcationInstance * first = reinterpret_cast<const BifurcationInstance*>(table[strand]);
			ret[strand]->assign(first, first + size[strand]);
			sibgpu_free(table[strand]);
		}

		return count;
	}

	size_t IndexedSequence::EnumerateBifurcationsSArray(const std::vector<std::string> & data, const std::string & tempDir, std::vector<BifurcationInstance> & positiveBif, std::vector<BifurcationInstance> & negativeBif)
	{
		// The file-backed variant of the reference creates the temp directory and two TempFiles (src/vertexenumeration.cpp:
		// 187-188, 125-126, 101); each file name draws 12 values from the process-wide rand() stream (src/platform.cpp:50-58)
		// that also replaces the non-ACGT characters of every later index: the side effects are kept, the files are not.
		ConsumeTempFileSideEffects(tempDir);
		return EnumerateBifurcationsSArrayInRAM(data, positiveBif, negativeBif);
	}
}
