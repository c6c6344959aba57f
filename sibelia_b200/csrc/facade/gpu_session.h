// One process-wide sibgpu context for the reference-side hooks (the reference is single-threaded and keeps its own
// process-wide state -- rand(), TempFile::register_ -- so a process-wide GPU session matches its model).
// Device selection: environment variable SIBELIA_GPU (default 0).  No CPU fallback: failure throws the same
// std::runtime_error the reference's main() already catches (src/sibelia.cpp:351-365).
#pragma once
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "platform.h"
#include "sibgpu.h"

namespace SyntenyFinder
{
	sibgpu_ctx * GpuSession();

	// What building an index with a temp directory does to process-wide state in the reference besides the index itself:
	// IndexedSequence::EnumerateBifurcationsSArray (src/vertexenumeration.cpp:187-188) creates the directory and two
	// TempFiles (:125-126, :101) whose names draw 12 rand() values each (src/platform.cpp:50-58) -- the same stream that
	// replaces non-ACGT characters in every later index (src/indexedsequence.cpp:31-37).  An empty tempDir is the
	// --inram path: nothing happens.
	inline void ConsumeTempFileSideEffects(const std::string & tempDir)
	{
		if(!tempDir.empty())
		{
			CreateOutDirectory(tempDir);
			for(int i = 0; i < 24; i++)
			{
				rand();
			}
		}
	}

	inline void GpuCheck(int status)
	{
		if(status != SIBGPU_OK)
		{
			throw std::runtime_error(std::string("sibgpu: ") + sibgpu_last_error());
		}
	}
}
