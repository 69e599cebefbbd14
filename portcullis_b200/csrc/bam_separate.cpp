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
#include "bam_out.hpp"
#include "inflate_fast.hpp"
#include "crc32_fast.hpp"

namespace pjio {

namespace {

inline uint32_t rd16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct InBlock { uint64_t coff; uint32_t bsize, xlen, isize; };

} // namespace

void scan_records(const BamFile& bam, int threads, const std::function<void(const uint8_t*, size_t)>& fn, const std::function<void()>& after_group) {
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
    // ---- stream: inflate a group of blocks in parallel, then walk its records ----
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
                // the fast decoder may write up to 64 bytes past the block: decode into scratch, another thread owns what follows
                uint8_t tmp[65536 + 64];
                const bool ok = B.coff + B.bsize + 16 <= mf.size() && fast[t]->run(src, n, tmp, B.isize);
                const uint32_t want_crc = rd32(mf.data() + B.coff + B.bsize - 8);
                if (ok) {
                    if (crc32_block(tmp, B.isize) != want_crc) throw IoError("BGZF block CRC32 mismatch in " + mf.path());
                    memcpy(dst, tmp, B.isize); continue;
                }
                if (!z_open) { if (inflateInit2(&z, -15) != Z_OK) throw IoError("inflateInit2 failed"); z_open = true; } else inflateReset(&z);
                z.next_in = (Bytef*)src; z.avail_in = (uInt)n; z.next_out = dst; z.avail_out = B.isize;
                if (inflate(&z, Z_FINISH) != Z_STREAM_END || z.total_out != B.isize) { inflateEnd(&z); throw IoError("BGZF inflate failed in " + mf.path()); }
                if (crc32_block(dst, B.isize) != want_crc) { inflateEnd(&z); throw IoError("BGZF block CRC32 mismatch in " + mf.path()); }
            }
            if (z_open) inflateEnd(&z);
        });
        const size_t total = carry.size() + off.back();
        size_t p = 0;
        if (skip) { const size_t k = (size_t)std::min<uint64_t>(skip, total); p = k; skip -= k; }
        while (p + 4 <= total) {
            const uint32_t bs = rd32(buf.data() + p);
            if (bs < 32) throw IoError("corrupt BAM record (block_size < 32)");
            if (p + 4 + bs > total) break;
            const uint8_t* r = buf.data() + p + 4;
            if (32ull + r[8] + 4ull * rd16(r + 12) > bs) throw IoError("corrupt BAM record (fields exceed block_size)");
            fn(buf.data() + p, 4 + (size_t)bs);
            p += 4 + (size_t)bs;
        }
        carry.assign(buf.begin() + (ptrdiff_t)p, buf.begin() + (ptrdiff_t)total);
        if (after_group) after_group();
    }
    if (!carry.empty()) throw IoError("truncated BAM record at the end of " + mf.path());
}

void index_existing_bam(const BamFile& bam, IndexBuilder& ib) {
    BgzfStream s(bam.file());
    s.seek(bam.header().first_record_voff);
    std::vector<uint8_t> rec;
    for (;;) {
        if (s.eof()) break;
        const uint64_t v0 = s.tell();
        uint8_t b4[4];
        if (s.read(b4, 4) != 4) throw IoError("truncated BAM record in " + bam.file().path());
        const uint32_t bs = rd32(b4);
        if (bs < 32) throw IoError("corrupt BAM record (block_size < 32)");
        rec.resize(bs);
        if (s.read(rec.data(), bs) != bs) throw IoError("truncated BAM record in " + bam.file().path());
        const uint64_t v1 = s.tell();
        const uint8_t* r = rec.data();
        const int32_t tid = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
        const uint32_t l_name = r[8], n_cig = rd16(r + 12); const uint16_t flag = (uint16_t)rd16(r + 14);
        if (32ull + l_name + 4ull * n_cig > bs) throw IoError("corrupt BAM record (fields exceed block_size)");
        const uint8_t* cg = r + 32 + l_name;
        int64_t rlen = 0; for (uint32_t k = 0; k < n_cig; k++) { const uint32_t c = rd32(cg + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; }
        const bool mapped = !(flag & 0x4);
        ib.add(tid, pos, (mapped && n_cig > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1, mapped, v0, v1);
    }
}

void separate_bams(const BamFile& bam, const std::string& spliced_path, const std::string& unspliced_path, const std::string& unmapped_path,
                   bool use_csi, int threads, SeparateCounts& counts) {
    const BamHeader& hdr = bam.header();
    threads = std::max(1, threads);
    BamOut spliced(spliced_path, true, hdr, use_csi), unspliced(unspliced_path, true, hdr, use_csi), unmapped(unmapped_path, false, hdr, use_csi);
    // JunctionBuilder::separateBams (junction_builder.cc:176-198)
    scan_records(bam, threads, [&](const uint8_t* rec, size_t len) {
        const uint8_t* r = rec + 4;
        const int32_t tid = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
        const uint32_t l_name = r[8], n_cig = rd16(r + 12); const uint16_t flag = (uint16_t)rd16(r + 14);
        const uint8_t* cg = r + 32 + l_name;
        int64_t rlen = 0; bool has_n = false;
        for (uint32_t k = 0; k < n_cig; k++) { const uint32_t c = rd32(cg + 4 * k), op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; if (op == 3) has_n = true; }
        const bool mapped = !(flag & 0x4);
        const int64_t end = (mapped && n_cig > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1;      // bam_endpos
        if (has_n) { spliced.add_record(rec, len, tid, pos, end, mapped); counts.spliced++; }
        else if (mapped) { unspliced.add_record(rec, len, tid, pos, end, mapped); counts.unspliced++; }
        else { unmapped.add_record(rec, len, tid, pos, end, mapped); counts.unmapped++; }
    }, [&]() { for (BamOut* o : {&spliced, &unspliced, &unmapped}) if (o->pending_blocks() >= 64) o->drain(threads, false); });
    for (BamOut* o : {&spliced, &unspliced, &unmapped}) o->drain(threads, true);
}

} // namespace pjio
