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

// true when every byte of [p, p+n) is printable non-space (isgraph in the C locale); written so that it vectorises
static inline bool all_graph(const uint8_t* p, size_t n) {
    unsigned bad = 0;
    for (size_t k = 0; k < n; k++) bad |= (unsigned)((uint8_t)(p[k] - 33u) > 93u);
    return bad == 0;
}

// number of bytes of [p, p+n) that are not printable non-space; a plain counting loop over a long range (vectorises)
static inline size_t count_non_graph(const uint8_t* p, size_t n) {
    size_t bad = 0;
    for (size_t k = 0; k < n; k++) bad += (size_t)((uint8_t)(p[k] - 33u) > 93u);
    return bad;
}
template <size_t BLEN> static inline void copy_lines(char* dst, const uint8_t* src, size_t n_lines, size_t blen, size_t llen) {
    for (size_t k = 0; k < n_lines; k++) memcpy(dst + k * (BLEN ? BLEN : blen), src + k * llen, BLEN ? BLEN : blen);   // a constant size is a few moves, not a call
}

void FastaFile::fetch_all(const FaiEntry& e, std::string& out) const {
    out.clear(); out.resize((size_t)e.len);
    const uint8_t* p = file_.data(); uint64_t n = file_.size(), o = e.offset; size_t l = 0;
    const size_t blen = (size_t)(e.line_blen > 0 ? e.line_blen : 1);
    const uint64_t term = (uint64_t)(e.line_len > e.line_blen ? e.line_len - e.line_blen : 0);
    // Block path (a human chromosome is millions of 60-base lines): a block of whole lines is regular when it holds exactly `term`
    // non-printable bytes per line and they sit in the terminator positions; its lines are then copied with constant-size moves.
    // The first irregular block, and the last partial line, go through the line loop below from where this one stopped.
    if (term > 0 && e.line_blen > 0) {
        const size_t llen = blen + (size_t)term, BLOCK = 2048;
        while (l + blen <= (size_t)e.len) {
            const size_t nl = std::min<size_t>(BLOCK, ((size_t)e.len - l) / blen);
            if (nl == 0 || o + nl * llen > n) break;
            const uint8_t* src = p + o;
            if (count_non_graph(src, nl * llen) != nl * (size_t)term) break;
            bool ok = true;
            for (size_t k = 0; k < nl && ok; k++) for (uint64_t q = 0; q < term; q++) ok &= !isgraph(src[k * llen + blen + q]);
            if (!ok) break;
            switch (blen) {
            case 60: copy_lines<60>(&out[l], src, nl, blen, llen); break;
            case 70: copy_lines<70>(&out[l], src, nl, blen, llen); break;
            case 80: copy_lines<80>(&out[l], src, nl, blen, llen); break;
            default: copy_lines<0>(&out[l], src, nl, blen, llen); break;
            }
            l += nl * blen; o += nl * llen;
        }
    }
    // Line path: whole lines of line_blen printable bytes followed by the terminator; anything irregular falls back to
    // the byte loop of faidx.c:470-472 (keep isgraph bytes, stop after e.len of them).
    while (l < (size_t)e.len && o < n) {
        const size_t want = std::min<size_t>((size_t)e.len - l, blen);
        if (o + want + term <= n && all_graph(p + o, want) && (want < blen || (term > 0 && !isgraph(p[o + want])))) {
            memcpy(&out[l], p + o, want); l += want; o += want;
            for (uint64_t k = 0; k < term && o < n && !isgraph(p[o]); k++) o++;
        } else {
            while (l < (size_t)e.len && o < n) { uint8_t c = p[o++]; if (isgraph(c)) out[l++] = (char)c; if (c == '\n') break; }
        }
    }
    out.resize(l);
}

} // namespace pjio
