// One process-wide sibgpu context for the reference-side hooks (the reference is single-threaded and keeps its own
// process-wide state -- rand(), TempFile::register_ -- so a process-wide GPU session matches its model).
// Device selection: environment variable SIBELIA_GPU (default 0).  No CPU fallback: failure throws the same
// std::runtime_error the reference's main() already catches (src/sibelia.cpp:351-365).
#pragma once
#include <stdexcept>
#include <string>

#include "sibgpu.h"

namespace SyntenyFinder
{
	sibgpu_ctx * GpuSession();

	inline void GpuCheck(int status)
	{
		if(status != SIBGPU_OK)
		{
			throw std::runtime_error(std::string("sibgpu: ") + sibgpu_last_error());
		}
	}
}
