// bam_separate.cpp — `junc --separate`: split the sorted BAM into <prefix>.spliced.bam, .unspliced.bam and .unmapped.bam
// and index the first two.  Host work only (BGZF inflate / deflate), the counterpart of
//   JunctionBuilder::separateBams            /root/reference/src/junction_builder.cc:152-226
//   BamWriter (bgzf_open "w", bam_hdr_write, bam_write1)   /root/reference/lib/src/bam_writer.cc:45-66
//   `samtools index [-c]`                    /root/reference/lib/src/bam_master.cc:123-125
//
// The reference does this on one thread.  Here the input blocks are inflated and the output blocks deflated on all host
// threads; only the record walk in between (a few bytes per record) is sequential.  The output is laid out exactly as
// htslib lays it out — header blocks cut at 0xff00 bytes and flushed (bam_hdr_write, sam.c:258), a record never straddles
// a block unless it cannot fit one (bgzf_flush_try, sam.c:448 / bgzf.c:763-767), default deflate level, EOF marker — so
// with the same zlib the three files are byte-identical to the reference's.  The indices are written from the SAM
// specification (§5.2 BAI, CSIv1); they are equivalent to, not byte-equal with, `samtools index` output.
#include "bam_io.hpp"
#include "bam_write.hpp"
#include "inflate_fast.hpp"
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

namespace {

inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

constexpr size_t BGZF_BLOCK = 0xff00;        // htslib BGZF_BLOCK_SIZE

struct InBlock { uint64_t coff; uint32_t bsize, xlen, isize; };

// Runs f(k) for k in [0, n) on up to `threads` threads; the first exception is rethrown on the caller.
void parallel_for(size_t n, int threads, const std::function<void(size_t)>& f) {
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

} // namespace

void separate_bams(const BamFile& bam, const std::string& spliced_path, const std::string& unspliced_path, const std::string& unmapped_path,
                   bool use_csi, int threads, SeparateCounts& counts) {
    const MappedFile& mf = bam.file();
    const BamHeader& hdr = bam.header();
    threads = std::max(1, threads);
    // ---- the input's BGZF blocks ----
    std::vector<InBlock> blocks;
    for (uint64_t coff = 0; coff + 18 <= mf.size();) {
        const uint8_t* p = mf.data() + coff;
        if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) throw IoError("not a BGZF block in " + mf.path());
        const uint32_t xlen = rd16(p + 10);
        if (coff + 12 + xlen > mf.size()) throw IoError("truncated BGZF header");
        uint32_t bsize = 0;
        for (uint32_t o = 0; o + 4 <= xlen;) { const uint8_t* e = p + 12 + o; const uint32_t sl = rd16(e + 2); if (e[0] == 'B' && e[1] == 'C' && sl == 2) bsize = rd16(e + 4) + 1; o += 4 + sl; }
        if (!bsize || coff + bsize > mf.size() || bsize < 12 + xlen + 8) throw IoError("corrupt BGZF block in " + mf.path());
        const uint32_t isize = rd32(p + bsize - 4);
        if (isize > 65536) throw IoError("BGZF block too large");
        blocks.push_back(InBlock{coff, bsize, xlen, isize});
        coff += bsize;
    }
    BamOut spliced(spliced_path, true, hdr, use_csi), unspliced(unspliced_path, true, hdr, use_csi), unmapped(unmapped_path, false, hdr, use_csi);
    // ---- stream: inflate a group of blocks in parallel, walk its records, compress the finished output blocks ----
    const size_t GROUP = 512;                                  // about 32 MB of records per round
    std::vector<std::unique_ptr<pjinflate::Inflater>> fast((size_t)threads);
    std::vector<uint8_t> buf, carry;
    // the header is skipped by its uncompressed length: first_record_voff = (coffset << 16 | uoffset)
    const uint64_t first_coff = hdr.first_record_voff >> 16; uint64_t skip = hdr.first_record_voff & 0xffff;
    size_t b0 = 0; while (b0 < blocks.size() && blocks[b0].coff < first_coff) b0++;
    for (; b0 < blocks.size(); b0 += GROUP) {
        const size_t b1 = std::min(blocks.size(), b0 + GROUP);
        std::vector<uint64_t> off(b1 - b0 + 1, 0);
        for (size_t k = b0; k < b1; k++) off[k - b0 + 1] = off[k - b0] + blocks[k].isize;
        buf.resize(carry.size() + off.back() + 64);
        if (!carry.empty()) memcpy(buf.data(), carry.data(), carry.size());
        uint8_t* base = buf.data() + carry.size();
        const size_t per = (b1 - b0 + (size_t)threads - 1) / (size_t)threads;
        parallel_for((size_t)threads, threads, [&](size_t t) {
            if (!fast[t]) fast[t] = std::make_unique<pjinflate::Inflater>();
            z_stream z; memset(&z, 0, sizeof z); bool z_open = false;
            for (size_t k = b0 + t * per; k < std::min(b1, b0 + (t + 1) * per); k++) {
                const InBlock& B = blocks[k];
                if (!B.isize) continue;
                const uint8_t* src = mf.data() + B.coff + 12 + B.xlen; const size_t n = B.bsize - 12 - B.xlen - 8;
                uint8_t* dst = base + off[k - b0];
                // the fast decoder may write up to 64 bytes past the block: decode into scratch when another thread owns what follows
                uint8_t tmp[65536 + 64];
                const bool ok = B.coff + B.bsize + 16 <= mf.size() && fast[t]->run(src, n, tmp, B.isize);
                if (ok) { memcpy(dst, tmp, B.isize); continue; }
                if (!z_open) { if (inflateInit2(&z, -15) != Z_OK) throw IoError("inflateInit2 failed"); z_open = true; } else inflateReset(&z);
                z.next_in = (Bytef*)src; z.avail_in = (uInt)n; z.next_out = dst; z.avail_out = B.isize;
                if (inflate(&z, Z_FINISH) != Z_STREAM_END || z.total_out != B.isize) { inflateEnd(&z); throw IoError("BGZF inflate failed in " + mf.path()); }
            }
            if (z_open) inflateEnd(&z);
        });
        const size_t total = carry.size() + off.back();
        size_t p = 0;
        if (skip) { const size_t k = (size_t)std::min<uint64_t>(skip, total); p = k; skip -= k; }
        // JunctionBuilder::separateBams (junction_builder.cc:176-198)
        while (p + 4 <= total) {
            const uint32_t bs = rd32(buf.data() + p);
            if (bs < 32) throw IoError("corrupt BAM record (block_size < 32)");
            if (p + 4 + bs > total) break;
            const uint8_t* r = buf.data() + p + 4;
            const int32_t tid = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
            const uint32_t l_name = r[8], n_cig = rd16(r + 12); const uint16_t flag = (uint16_t)rd16(r + 14);
            if (32ull + l_name + 4ull * n_cig > bs) throw IoError("corrupt BAM record (fields exceed block_size)");
            const uint8_t* cg = r + 32 + l_name;
            int64_t rlen = 0; bool has_n = false;
            for (uint32_t k = 0; k < n_cig; k++) { const uint32_t c = rd32(cg + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; if (op == 3) has_n = true; }
            const bool mapped = !(flag & 0x4);
            const int64_t end = (mapped && n_cig > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1;      // bam_endpos
            if (has_n) { spliced.add_record(buf.data() + p, 4 + (size_t)bs, tid, pos, end, mapped); counts.spliced++; }
            else if (mapped) { unspliced.add_record(buf.data() + p, 4 + (size_t)bs, tid, pos, end, mapped); counts.unspliced++; }
            else { unmapped.add_record(buf.data() + p, 4 + (size_t)bs, tid, pos, end, mapped); counts.unmapped++; }
            p += 4 + (size_t)bs;
        }
        carry.assign(buf.begin() + (ptrdiff_t)p, buf.begin() + (ptrdiff_t)total);
        for (BamOut* o : {&spliced, &unspliced, &unmapped}) if (o->pending_blocks() >= 64) o->drain(threads, false);
    }
    if (!carry.empty()) throw IoError("truncated BAM record at the end of " + mf.path());
    for (BamOut* o : {&spliced, &unspliced, &unmapped}) o->drain(threads, true);
}

} // namespace pjio
