// bam_out.hpp — BAM output side shared by `junc --separate` (bam_separate.cpp) and `prep` (prep_driver.cpp): a writer that
// lays blocks out the way htslib's does, with BAI / CSI index construction, and a small parallel-for.
#pragma once
#include "bam_io.hpp"
#include "bam_write.hpp"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <zlib.h>

namespace pjio {

constexpr size_t BGZF_BLOCK = 0xff00;        // htslib BGZF_BLOCK_SIZE

// Runs f(k) for k in [0, n) on up to `threads` threads; the first exception is rethrown on the caller.
inline void parallel_for(size_t n, int threads, const std::function<void(size_t)>& f) {
    if (n == 0) return;
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(1, threads), n));
    if (nt == 1) { for (size_t k = 0; k < n; k++) f(k); return; }
    std::atomic<size_t> next{0}; std::exception_ptr err; std::mutex mu;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() {
        try { for (;;) { const size_t k = next.fetch_add(1); if (k >= n) return; f(k); } }
        catch (...) { std::lock_guard<std::mutex> lk(mu); if (!err) err = std::current_exception(); next.store(n); }
    });
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

// Index under construction (BAI binning generalised to CSI's min_shift / depth).
struct IndexBuilder {
    int min_shift = 14, depth = 5;
    struct Target {
        std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
        std::vector<uint64_t> linear;        // window -> smallest record-start voffset + 1 (0 = none)
        uint64_t n_mapped = 0, n_unmapped = 0, off_beg = 0, off_end = 0; bool any = false;
    };
    std::vector<Target> targets;
    uint64_t n_no_coor = 0;
    uint32_t reg2bin(int64_t beg, int64_t end) const {       // CSIv1 specification, reg2bin
        int l, s = min_shift, t = ((1 << depth * 3) - 1) / 7;
        for (--end, l = depth; l > 0; --l, s += 3, t -= 1 << l * 3)
            if (beg >> s == end >> s) return (uint32_t)(t + (beg >> s));
        return 0;
    }
    void add(int32_t tid, int64_t beg, int64_t end, bool mapped, uint64_t v0, uint64_t v1) {
        if (tid < 0) { n_no_coor++; return; }
        Target& T = targets[(size_t)tid];
        if (end <= beg) end = beg + 1;
        auto& ch = T.bins[reg2bin(beg, end)];
        if (!ch.empty() && ch.back().second == v0) ch.back().second = v1; else ch.emplace_back(v0, v1);
        const size_t w0 = (size_t)(beg >> min_shift), w1 = (size_t)((end - 1) >> min_shift);
        if (T.linear.size() <= w1) T.linear.resize(w1 + 1, 0);
        for (size_t w = w0; w <= w1; w++) if (T.linear[w] == 0) T.linear[w] = v0 + 1;
        if (!T.any) { T.any = true; T.off_beg = v0; }
        T.off_end = v1;
        if (mapped) T.n_mapped++; else T.n_unmapped++;
    }
    void fill_linear(Target& T) const { for (size_t w = T.linear.size(); w-- > 0;) if (T.linear[w] == 0 && w + 1 < T.linear.size()) T.linear[w] = T.linear[w + 1]; }
    uint32_t meta_bin() const { return (uint32_t)(((1u << (depth + 1) * 3) - 1) / 7 + 1); }
    std::vector<uint8_t> bai() {
        std::vector<uint8_t> o = {'B', 'A', 'I', 1};
        bamw::put32(o, (uint32_t)targets.size());
        for (Target& T : targets) {
            bamw::put32(o, (uint32_t)(T.bins.size() + (T.any ? 1 : 0)));
            for (auto& b : T.bins) { bamw::put32(o, b.first); bamw::put32(o, (uint32_t)b.second.size()); for (auto& c : b.second) { bamw::put64(o, c.first); bamw::put64(o, c.second); } }
            if (T.any) { bamw::put32(o, meta_bin()); bamw::put32(o, 2); bamw::put64(o, T.off_beg); bamw::put64(o, T.off_end); bamw::put64(o, T.n_mapped); bamw::put64(o, T.n_unmapped); }
            fill_linear(T);
            bamw::put32(o, (uint32_t)T.linear.size());
            for (uint64_t v : T.linear) bamw::put64(o, v ? v - 1 : 0);
        }
        bamw::put64(o, n_no_coor);
        return o;
    }
    // first window of a bin (CSIv1: bin -> level -> offset inside the level)
    size_t bin_first_window(uint32_t bin) const {
        int l = 0; uint32_t first = 0;
        for (;; l++) { const uint32_t n = 1u << (l * 3); if (bin < first + n) break; first += n; }
        return (size_t)(bin - first) << ((depth - l) * 3);
    }
    std::vector<uint8_t> csi() {
        std::vector<uint8_t> raw = {'C', 'S', 'I', 1};
        bamw::put32(raw, (uint32_t)min_shift); bamw::put32(raw, (uint32_t)depth); bamw::put32(raw, 0);
        bamw::put32(raw, (uint32_t)targets.size());
        for (Target& T : targets) {
            fill_linear(T);
            bamw::put32(raw, (uint32_t)(T.bins.size() + (T.any ? 1 : 0)));
            for (auto& b : T.bins) {
                const size_t w = bin_first_window(b.first);
                const uint64_t loff = w < T.linear.size() && T.linear[w] ? T.linear[w] - 1 : 0;
                bamw::put32(raw, b.first); bamw::put64(raw, loff); bamw::put32(raw, (uint32_t)b.second.size());
                for (auto& c : b.second) { bamw::put64(raw, c.first); bamw::put64(raw, c.second); }
            }
            if (T.any) { bamw::put32(raw, meta_bin()); bamw::put64(raw, 0); bamw::put32(raw, 2); bamw::put64(raw, T.off_beg); bamw::put64(raw, T.off_end); bamw::put64(raw, T.n_mapped); bamw::put64(raw, T.n_unmapped); }
        }
        bamw::put64(raw, n_no_coor);
        std::vector<uint8_t> o;                               // a CSI file is itself BGZF-compressed
        for (size_t p = 0; p < raw.size(); p += BGZF_BLOCK) bamw::bgzf_block(raw.data() + p, std::min(BGZF_BLOCK, raw.size() - p), o, Z_DEFAULT_COMPRESSION);
        bamw::bgzf_eof(o);
        return o;
    }
};

// One output BAM.  Records are appended to uncompressed blocks cut the way htslib cuts them; complete blocks are
// compressed in batches on all threads and written in order.
class BamOut {
public:
    BamOut(const std::string& path, bool want_index, const BamHeader& hdr, bool csi) : path_(path), want_index_(want_index), csi_(csi) {
        f_ = fopen(path.c_str(), "wb");
        if (!f_) throw IoError("Could not open output BAM file: " + path);
        if (want_index) {
            idx_.targets.resize(hdr.lens.size());
            if (csi) {       // bam_index(fp, min_shift = 14): depth grows until the longest target + 256 fits (sam.c:475-481)
                int64_t max_len = 0; for (int32_t l : hdr.lens) max_len = std::max<int64_t>(max_len, l);
                max_len += 256;
                int n = 0; for (int64_t s = 1 << 14; max_len > s; ++n, s <<= 3) {}
                idx_.depth = n;
            }
        }
        // bam_hdr_write (sam.c:224-259): magic, l_text, text, n_ref, {l_name, name, l_ref}; then bgzf_flush
        std::vector<uint8_t> h = {'B', 'A', 'M', 1};
        bamw::put32(h, (uint32_t)hdr.text.size()); h.insert(h.end(), hdr.text.begin(), hdr.text.end());
        bamw::put32(h, (uint32_t)hdr.names.size());
        for (size_t i = 0; i < hdr.names.size(); i++) {
            bamw::put32(h, (uint32_t)hdr.names[i].size() + 1); h.insert(h.end(), hdr.names[i].begin(), hdr.names[i].end()); h.push_back(0);
            bamw::put32(h, (uint32_t)hdr.lens[i]);
        }
        write(h.data(), h.size());
        flush_block();
    }
    ~BamOut() { if (f_) fclose(f_); }
    // rec points at the block_size field; len = 4 + block_size
    void add_record(const uint8_t* rec, size_t len, int32_t tid, int64_t beg, int64_t end, bool mapped) {
        if (cur_.size() + len > BGZF_BLOCK) flush_block();                   // bgzf_flush_try(fp, 4 + block_len)
        const Mark m0{(uint32_t)(n_done_ + pending_.size()), (uint32_t)cur_.size()};
        write(rec, len);
        if (want_index_) marks_.push_back(Pending{m0, Mark{(uint32_t)(n_done_ + pending_.size()), (uint32_t)cur_.size()}, tid, beg, end, mapped});
        n_records_++;
    }
    size_t pending_blocks() const { return pending_.size(); }
    // compress and write every complete block (all of them with final = true, plus the EOF marker and the index)
    void drain(int threads, bool final) {
        if (final) flush_block();
        std::vector<std::vector<uint8_t>> comp(pending_.size());
        parallel_for(pending_.size(), threads, [&](size_t k) { bamw::bgzf_block(pending_[k].data(), pending_[k].size(), comp[k], Z_DEFAULT_COMPRESSION); });
        for (auto& c : comp) {
            block_coff_.push_back(file_off_);
            if (fwrite(c.data(), 1, c.size(), f_) != c.size()) throw IoError("write failed: " + path_);
            file_off_ += c.size();
        }
        n_done_ += pending_.size(); pending_.clear();
        // index entries whose start and end blocks now have a file offset
        while (!marks_.empty()) {
            const Pending& p = marks_.front();
            const bool end_known = p.v1.block < block_coff_.size() || (p.v1.off == 0 && final);
            if (p.v0.block >= block_coff_.size() || !end_known) break;
            idx_.add(p.tid, p.beg, p.end, p.mapped, voff(p.v0), voff(p.v1));
            marks_.pop_front();
        }
        if (final) {
            std::vector<uint8_t> eof; bamw::bgzf_eof(eof);
            if (fwrite(eof.data(), 1, eof.size(), f_) != eof.size()) throw IoError("write failed: " + path_);
            if (fclose(f_) != 0) { f_ = nullptr; throw IoError("close failed: " + path_); }
            f_ = nullptr;
            if (want_index_) {
                const std::vector<uint8_t> ix = csi_ ? idx_.csi() : idx_.bai();
                const std::string ip = path_ + (csi_ ? ".csi" : ".bai");
                FILE* g = fopen(ip.c_str(), "wb");
                if (!g || fwrite(ix.data(), 1, ix.size(), g) != ix.size() || fclose(g) != 0) throw IoError("Could not write index: " + ip);
            }
        }
    }
    uint64_t n_records() const { return n_records_; }
private:
    struct Mark { uint32_t block, off; };
    struct Pending { Mark v0, v1; int32_t tid; int64_t beg, end; bool mapped; };
    // a mark at the very end of a block names the start of the next one (bgzf_tell after the block was flushed)
    uint64_t voff(const Mark& m) const { return m.block < block_coff_.size() ? (block_coff_[m.block] << 16) | m.off : (file_off_ << 16); }
    void write(const uint8_t* p, size_t n) {                                 // bgzf_write: fill, flush when full
        while (n) {
            const size_t k = std::min(n, BGZF_BLOCK - cur_.size());
            cur_.insert(cur_.end(), p, p + k); p += k; n -= k;
            if (cur_.size() == BGZF_BLOCK) flush_block();
        }
    }
    void flush_block() { if (!cur_.empty()) { pending_.emplace_back(std::move(cur_)); cur_.clear(); cur_.reserve(BGZF_BLOCK); } }
    std::string path_; FILE* f_ = nullptr; bool want_index_, csi_;
    std::vector<uint8_t> cur_; std::vector<std::vector<uint8_t>> pending_;
    size_t n_done_ = 0; uint64_t file_off_ = 0; std::vector<uint64_t> block_coff_;
    std::deque<Pending> marks_; IndexBuilder idx_; uint64_t n_records_ = 0;
};


// BAI / CSI of an existing coordinate-sorted BAM (the job of `samtools index`): one sequential pass.  bam_separate.cpp.
void index_existing_bam(const BamFile& bam, IndexBuilder& ib);

} // namespace pjio
