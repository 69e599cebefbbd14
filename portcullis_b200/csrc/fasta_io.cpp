#include "fasta_io.hpp"
#include <fstream>
#include <sstream>
#include <cctype>
#include <cstring>

namespace pjio {

void FastaFile::open(const std::string& fasta_path, const std::string& fai_path) {
    file_.open(fasta_path);
    std::ifstream in(fai_path);
    if (!in) throw IoError("cannot open FASTA index " + fai_path);
    std::string line;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        // name \t len \t offset \t line_blen \t line_len   (faidx.c:40-44)
        std::vector<std::string> f; size_t b = 0;
        for (;;) { size_t e = line.find('\t', b); f.push_back(line.substr(b, e == std::string::npos ? e : e - b)); if (e == std::string::npos) break; b = e + 1; }
        if (f.size() < 5) throw IoError("malformed .fai line: " + line);
        FaiEntry e; e.name = f[0]; e.len = std::stoll(f[1]); e.offset = std::stoull(f[2]);
        e.line_blen = std::stoi(f[3]); e.line_len = std::stoi(f[4]);
        by_name_[e.name] = entries_.size(); entries_.push_back(e);
    }
}

const FaiEntry* FastaFile::find(const std::string& name) const {
    auto it = by_name_.find(name);
    return it == by_name_.end() ? nullptr : &entries_[it->second];
}

void FastaFile::fetch_all(const FaiEntry& e, std::string& out) const {
    out.clear(); out.resize((size_t)e.len);
    const uint8_t* p = file_.data(); uint64_t n = file_.size(), o = e.offset; size_t l = 0;
    // Fast path: copy line_blen bytes per line, then verify they are all printable; fall back to the
    // byte loop of faidx.c:470-472 when a line is irregular.
    while (l < (size_t)e.len && o < n) {
        size_t want = std::min<size_t>((size_t)e.len - l, (size_t)(e.line_blen > 0 ? e.line_blen : 1));
        size_t avail = (size_t)std::min<uint64_t>(want, n - o);
        bool clean = true;
        for (size_t k = 0; k < avail; k++) if (!isgraph(p[o + k])) { clean = false; break; }
        if (clean && avail == want) {
            memcpy(&out[l], p + o, want); l += want; o += want;
            // skip the line terminator(s)
            uint64_t term = (uint64_t)(e.line_len - e.line_blen);
            for (uint64_t k = 0; k < term && o < n && !isgraph(p[o]); k++) o++;
        } else {
            while (l < (size_t)e.len && o < n) { uint8_t c = p[o++]; if (isgraph(c)) out[l++] = (char)c; if (c == '\n') break; }
        }
    }
    out.resize(l);
}

} // namespace pjio
