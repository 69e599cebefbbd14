// fasta_io.hpp — FASTA + .fai random access for the junc path.
//
// Replaces GenomeMapper::loadFastaIndex / fetchBases (lib/src/genome_mapper.cc:76-118) and
// faidx_fetch_seq (deps/htslib-1.3/faidx.c:439-476): instead of one file seek per junction window,
// whole target sequences are unwrapped once and handed to the GPU library, which keeps them packed
// in HBM.  Coordinates and the "isgraph bytes only" rule follow faidx.c:455-473.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <unordered_map>
#include "bam_io.hpp"

namespace pjio {

struct FaiEntry { std::string name; int64_t len = 0; uint64_t offset = 0; int32_t line_blen = 0, line_len = 0; };

class FastaFile {
public:
    void open(const std::string& fasta_path, const std::string& fai_path);
    const FaiEntry* find(const std::string& name) const;
    // Unwrapped bytes of one sequence (isgraph bytes only, original case), at most entry.len of them.
    void fetch_all(const FaiEntry& e, std::string& out) const;
    const std::vector<FaiEntry>& entries() const { return entries_; }
private:
    MappedFile file_;
    std::vector<FaiEntry> entries_;
    std::unordered_map<std::string, size_t> by_name_;
};

} // namespace pjio
