#!/usr/bin/env python
"""Writes a copy of the reference's src/synteny.cpp (or src/serialization.cpp) with its index-building sites bound to
libsibgpu.so -- the change a maintainer of the reference would make by hand:

  * BlockFinder::GenerateSyntenyBlocks (src/synteny.cpp:238-241): `IndexedSequence iseq(...); ListEdges(...)`
    becomes one sibgpu_list_edges call;
  * BlockFinder::TrimBlocks (src/synteny.cpp:31-122): `IndexedSequence iseq(blockSeq, trimK, "")` and the walk over all
    vertex marks become one sibgpu_trim_blocks call, the final size test / Edge construction (:105-116) is kept.

  * BlockFinder::SerializeCondensedGraph (src/serialization.cpp:88-94, the -g output): the same
    `IndexedSequence iseq(...); ListEdges(...)` pair becomes the same sibgpu_list_edges call.

The reference source is read where it lies and edited by anchors; nothing of it is stored in this repository and the
output (a file in the build directory of whoever links the bound CLI) is a build artefact.

    python patch_synteny.py /root/reference/src/synteny.cpp out.cpp
    python patch_synteny.py /root/reference/src/serialization.cpp out.cpp
"""
import sys

LIST_EDGES = r'''			// sibgpu: index + ListEdges in one GPU call, no host-side IndexedSequence (include/sibgpu.h, sibgpu_list_edges).
			// The sanitising of IndexedSequence::Init (indexedsequence.cpp:31-37) stays here: it consumes rand().
			std::vector<std::string> record(rawSeq_);
			std::vector<const char*> chrPtr(record.size());
			std::vector<const uint32_t*> posPtr(record.size());
			std::vector<uint64_t> chrLen(record.size());
			for(size_t i = 0; i < record.size(); i++)
			{
				for(size_t j = 0; j < record[i].size(); j++)
				{
					record[i][j] = IsDefiniteBase(record[i][j]) ? record[i][j] : DEFINITE_BASE[rand() % DEFINITE_BASE.size()];
				}

				chrPtr[i] = record[i].data();
				posPtr[i] = originalPos_[i].empty() ? 0 : &originalPos_[i][0];
				chrLen[i] = record[i].size();
			}

			ConsumeTempFileSideEffects(tempDir_);           // IndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_)
			sibgpu_edge * gpuEdge = 0;
			uint64_t gpuEdgeCount = 0;
			GpuCheck(sibgpu_list_edges(GpuSession(), record.empty() ? 0 : &chrPtr[0], record.empty() ? 0 : &posPtr[0], record.empty() ? 0 : &chrLen[0],
				static_cast<uint32_t>(record.size()), static_cast<uint32_t>(k), &gpuEdge, &gpuEdgeCount));
			edge.clear();
			edge.reserve(gpuEdgeCount);
			for(uint64_t i = 0; i < gpuEdgeCount; i++)
			{
				const sibgpu_edge & e = gpuEdge[i];
				edge.push_back(Edge(e.chr, e.direction == 0 ? DNASequence::positive : DNASequence::negative, e.start_vertex, e.end_vertex,
					e.actual_position, e.actual_length, e.original_position, e.original_length, static_cast<char>(e.first_char)));
			}

			sibgpu_free(gpuEdge);
'''

TRIM = r'''		// sibgpu: the index of blockSeq and the search for the trim points run in sibgpu_trim_blocks (include/sibgpu.h);
		// the sanitising of IndexedSequence::Init (indexedsequence.cpp:31-37) stays here: it consumes rand().
		std::vector<const char*> chrPtr(blockSeq.size());
		std::vector<uint64_t> chrLen(blockSeq.size());
		std::vector<uint8_t> chrDir(blockSeq.size());
		for(size_t i = 0; i < blockSeq.size(); i++)
		{
			for(size_t j = 0; j < blockSeq[i].size(); j++)
			{
				blockSeq[i][j] = IsDefiniteBase(blockSeq[i][j]) ? blockSeq[i][j] : DEFINITE_BASE[rand() % DEFINITE_BASE.size()];
			}

			chrPtr[i] = blockSeq[i].data();
			chrLen[i] = blockSeq[i].size();
			chrDir[i] = block[i].GetDirection() == DNASequence::positive ? 0 : 1;
		}

		std::vector<sibgpu_trim> trim(block.size() + 1);
		GpuCheck(sibgpu_trim_blocks(GpuSession(), block.empty() ? 0 : &chrPtr[0], block.empty() ? 0 : &chrLen[0], block.empty() ? 0 : &chrDir[0],
			static_cast<uint32_t>(block.size()), static_cast<uint32_t>(trimK), &trim[0]));
		std::vector<Edge> ret;
		for(size_t chr = 0; chr < block.size(); chr++)
		{
			if(trim[chr].found)
			{
				size_t trimStart = trim[chr].start;
				size_t trimEnd = trim[chr].end;
				size_t size = std::max(trimStart, trimEnd) - std::min(trimStart, trimEnd) + trimK;
				if(size >= minSize)
				{
					trimEnd = block[chr].GetDirection() == DNASequence::positive ? trimEnd + (trimK - 1) : trimEnd - (trimK - 1);
					size_t start = block[chr].GetOriginalPosition() + std::min(trimStart, trimEnd);
					size_t end = block[chr].GetOriginalPosition() + std::max(trimStart, trimEnd) + 1;
					ret.push_back(Edge(block[chr].GetChr(), block[chr].GetDirection(), block[chr].GetStartVertex(), block[chr].GetEndVertex(),
						block[chr].GetActualPosition(), block[chr].GetActualLength(), start, end - start, block[chr].GetFirstChar()));
				}
			}
			else
			{
				drop = true;
			}
		}

'''


def between(text, start_anchor, end_anchor, replacement, what):
    a = text.find(start_anchor)
    b = text.find(end_anchor, a)
    if a < 0 or b < 0:
        sys.exit("patch_synteny: anchor for %s not found -- the reference changed" % what)
    return text[:a] + replacement + text[b:]


def main(src, dst):
    s = open(src).read()
    s = s.replace('#include "blockfinder.h"\n', '#include "blockfinder.h"\n#include "gpu_session.h"\n', 1)
    if src.endswith("serialization.cpp"):
        # SerializeCondensedGraph: drop the host-side index, list the edges on the GPU
        a = s.find("\tvoid BlockFinder::SerializeCondensedGraph(")
        if a < 0 or s.find("\t\tIndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_);\n", a) < 0:
            sys.exit("patch_synteny: anchor for SerializeCondensedGraph not found -- the reference changed")
        head, tail = s[:a], s[a:]
        tail = tail.replace("\t\tIndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_);\n", "", 1)
        tail = tail.replace("\t\tListEdges(iseq.Sequence(), iseq.BifStorage(), k, edge);\n", LIST_EDGES, 1)
        open(dst, "w").write(head + tail)
        return
    # TrimBlocks: everything from the sentinel constant to the final swap
    s = between(s, "\t\tconst size_t oo = UINT_MAX;", "\t\tblock.swap(ret);", TRIM, "TrimBlocks")
    s = s.replace("\t\tsize_t pos = 0;\n\t\tbool drop = false;", "\t\tbool drop = false;", 1)
    # GenerateSyntenyBlocks: the scoped index + ListEdges
    s = between(s, "\t\t\tIndexedSequence iseq(rawSeq_, originalPos_, k, tempDir_);", "\t\t}\n", LIST_EDGES, "GenerateSyntenyBlocks")
    open(dst, "w").write(s)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
