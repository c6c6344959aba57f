// Drop-in replacement for the reference translation unit src/fasta.cpp (FASTAReader::GetSequences, :22-73).
//
// Compiled against the reference's own fasta.h and linked INSTEAD of fasta.cpp.  The file is read in one piece and
// parsed on the GPU (sibgpu_fasta_parse: line splitting, trimming, records, upper-casing, validation); the records,
// their order and ids, and the text of every parse error are the reference's.
#include "fasta.h"

#include "gpu_session.h"

namespace SyntenyFinder
{
	size_t FASTAReader::GetSequences(std::vector<FASTARecord> & record)
	{
		std::string all;
		inputStream_.seekg(0, std::ios::end);
		const std::streamoff size = inputStream_.tellg();
		inputStream_.seekg(0, std::ios::beg);
		if(size > 0)
		{
			all.resize(static_cast<size_t>(size));
			inputStream_.read(&all[0], size);
			all.resize(static_cast<size_t>(inputStream_.gcount()));
		}

		sibgpu_fasta parsed;
		uint64_t line = 0;
		const int status = sibgpu_fasta_parse(GpuSession(), all.data(), all.size(), &parsed, &line);
		if(status == SIBGPU_ERR_INPUT)
		{
			std::stringstream ss;
			ss << "parse error in " << fileName_ << " on line " << line << ": " << sibgpu_last_error();
			throw std::runtime_error(ss.str());
		}

		GpuCheck(status);
		size_t seqId = record.size();
		record.reserve(record.size() + parsed.nrec);
		for(uint32_t i = 0; i < parsed.nrec; i++)
		{
			const sibgpu_fasta_record & r = parsed.rec[i];
			record.push_back(FASTARecord(std::string(r.seq, r.seq + r.len), std::string(r.name, r.name + r.name_len), seqId++));
		}

		sibgpu_fasta_free(&parsed);
		return record.size();
	}

	bool FASTAReader::IsOk() const
	{
		return inputStream_.good();
	}
}
