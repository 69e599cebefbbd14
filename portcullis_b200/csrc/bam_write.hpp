// bam_write.hpp — minimal BGZF/BAM/BAI writer used by the synthetic workload generator (pjsynth).
// Written from the SAM/BAM specification (§4.1 BGZF, §4.2 BAM, §5.2 BAI).  A BAM "fragment" is a run of complete
// BGZF blocks holding records of one position slice; fragments are concatenated behind the header blocks and their
// partial indices are merged with rebased virtual offsets, which lets slices be generated and compressed in parallel.
#pragma once
#include <zlib.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace bamw {

inline void put16(std::vector<uint8_t>& v, uint16_t x) { v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); }
inline void put32(std::vector<uint8_t>& v, uint32_t x) { for (int k = 0; k < 4; k++) v.push_back((uint8_t)(x >> (8 * k))); }
inline void put64(std::vector<uint8_t>& v, uint64_t x) { for (int k = 0; k < 8; k++) v.push_back((uint8_t)(x >> (8 * k))); }

// Compress one BGZF block (<= 65280 input bytes) and append it to `out`.
inline void bgzf_block(const uint8_t* in, size_t n, std::vector<uint8_t>& out, int level) {
    if (n > 65280) throw std::runtime_error("bgzf block too large");
    uint8_t buf[70000];
    z_stream z; memset(&z, 0, sizeof z);
    if (deflateInit2(&z, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2");
    z.next_in = (Bytef*)in; z.avail_in = (uInt)n; z.next_out = buf; z.avail_out = sizeof buf;
    if (deflate(&z, Z_FINISH) != Z_STREAM_END) { deflateEnd(&z); throw std::runtime_error("deflate"); }
    const size_t clen = z.total_out;
    deflateEnd(&z);
    const size_t bsize = clen + 26;           // header 18 + data + crc32 4 + isize 4
    if (bsize > 65536) throw std::runtime_error("bgzf block expands beyond 64 KiB");
    static const uint8_t hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
    out.insert(out.end(), hdr, hdr + 12);
    out.push_back('B'); out.push_back('C'); put16(out, 2); put16(out, (uint16_t)(bsize - 1));
    out.insert(out.end(), buf, buf + clen);
    put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, (uInt)n));
    put32(out, (uint32_t)n);
}

inline int reg2bin(int64_t beg, int64_t end) {   // SAM spec §5.3
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// Index pieces of one fragment; virtual offsets are relative to the fragment's first byte.
struct PartialIndex {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;            // window -> min voffset + 1 (0 = empty)
    uint64_t n_mapped = 0, n_unmapped = 0, off_beg = 0, off_end = 0; bool any = false;
};

class FragmentWriter {
public:
    explicit FragmentWriter(int level = 1) : level_(level) { block_.reserve(65280); }
    // Append one BAM record (without the leading block_size field) covering [beg, end) on its target.
    void add_record(const std::vector<uint8_t>& rec, int64_t beg, int64_t end, bool mapped) {
        const uint64_t v0 = voff();
        uint8_t bs[4]; uint32_t n = (uint32_t)rec.size(); memcpy(bs, &n, 4);
        write(bs, 4); write(rec.data(), rec.size());
        const uint64_t v1 = voff();
        if (end <= beg) end = beg + 1;
        auto& ch = idx_.bins[(uint32_t)reg2bin(beg, end)];
        if (!ch.empty() && ch.back().second == v0) ch.back().second = v1; else ch.emplace_back(v0, v1);
        const size_t w0 = (size_t)(beg >> 14), w1 = (size_t)((end - 1) >> 14);
        if (idx_.linear.size() <= w1) idx_.linear.resize(w1 + 1, 0);
        for (size_t w = w0; w <= w1; w++) if (idx_.linear[w] == 0) idx_.linear[w] = v0 + 1;
        if (!idx_.any) { idx_.any = true; idx_.off_beg = v0; }
        idx_.off_end = v1;
        if (mapped) idx_.n_mapped++; else idx_.n_unmapped++;
    }
    void write(const uint8_t* p, size_t n) {
        while (n) {
            const size_t k = std::min(n, (size_t)65280 - block_.size());
            block_.insert(block_.end(), p, p + k); p += k; n -= k;
            if (block_.size() == 65280) flush();
        }
    }
    void flush() { if (!block_.empty()) { bgzf_block(block_.data(), block_.size(), out_, level_); block_.clear(); } }
    std::vector<uint8_t>& bytes() { return out_; }
    PartialIndex& index() { return idx_; }
private:
    // virtual offset of the next byte; a full block is flushed eagerly, so block_.size() < 65280 here
    uint64_t voff() const { return ((uint64_t)out_.size() << 16) | (uint64_t)block_.size(); }
    int level_; std::vector<uint8_t> block_, out_; PartialIndex idx_;
};

inline uint64_t rebase(uint64_t v, uint64_t base) { return (((v >> 16) + base) << 16) | (v & 0xffff); }

// Merge the partial indices of one target's fragments (in position order); `bases` are the fragments' file offsets.
inline void append_target_index(std::vector<uint8_t>& bai, const std::vector<const PartialIndex*>& parts, const std::vector<uint64_t>& bases) {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> lin; uint64_t nm = 0, nu = 0, ob = 0, oe = 0; bool any = false;
    for (size_t k = 0; k < parts.size(); k++) {
        const PartialIndex& p = *parts[k];
        for (auto& b : p.bins) for (auto& c : b.second) bins[b.first].emplace_back(rebase(c.first, bases[k]), rebase(c.second, bases[k]));
        if (lin.size() < p.linear.size()) lin.resize(p.linear.size(), 0);
        for (size_t w = 0; w < p.linear.size(); w++) if (p.linear[w] && lin[w] == 0) lin[w] = rebase(p.linear[w] - 1, bases[k]) + 1;
        nm += p.n_mapped; nu += p.n_unmapped;
        if (p.any) { if (!any) { any = true; ob = rebase(p.off_beg, bases[k]); } oe = rebase(p.off_end, bases[k]); }
    }
    put32(bai, (uint32_t)(bins.size() + (any ? 1 : 0)));
    for (auto& b : bins) {
        put32(bai, b.first); put32(bai, (uint32_t)b.second.size());
        for (auto& c : b.second) { put64(bai, c.first); put64(bai, c.second); }
    }
    if (any) { put32(bai, 37450); put32(bai, 2); put64(bai, ob); put64(bai, oe); put64(bai, nm); put64(bai, nu); }
    // linear index: empty windows take the next window's offset (htslib's convention)
    for (size_t w = lin.size(); w-- > 0;) if (lin[w] == 0 && w + 1 < lin.size()) lin[w] = lin[w + 1];
    put32(bai, (uint32_t)lin.size());
    for (uint64_t v : lin) put64(bai, v ? v - 1 : 0);
}

// Header as one or more BGZF blocks.
inline void bam_header(std::vector<uint8_t>& out, const std::string& text, const std::vector<std::string>& names, const std::vector<int32_t>& lens) {
    std::vector<uint8_t> h;
    h.insert(h.end(), {'B', 'A', 'M', 1});
    put32(h, (uint32_t)text.size()); h.insert(h.end(), text.begin(), text.end());
    put32(h, (uint32_t)names.size());
    for (size_t i = 0; i < names.size(); i++) {
        put32(h, (uint32_t)names[i].size() + 1); h.insert(h.end(), names[i].begin(), names[i].end()); h.push_back(0);
        put32(h, (uint32_t)lens[i]);
    }
    for (size_t o = 0; o < h.size(); o += 65280) bgzf_block(h.data() + o, std::min<size_t>(65280, h.size() - o), out, 6);
}

inline void bgzf_eof(std::vector<uint8_t>& out) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 0x08, 0x04, 0, 0, 0, 0, 0, 0xff, 0x06, 0, 0x42, 0x43, 0x02, 0, 0x1b, 0, 0x03, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    out.insert(out.end(), eof, eof + 28);
}

} // namespace bamw
