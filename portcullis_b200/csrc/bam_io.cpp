// bam_io.cpp — see bam_io.hpp.  BGZF framing: SAM spec §4.1; BAM records: §4.2; BAI: §5.2.
#include "bam_io.hpp"
#include <zlib.h>
#include "inflate_fast.hpp"
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "crc32_fast.hpp"
#include <cstring>
#include <climits>
#include <algorithm>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/stat.h>

namespace pjio {

static inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

MappedFile::~MappedFile() { if (data_ && size_) munmap((void*)data_, size_); }

void MappedFile::open(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw IoError("cannot open " + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); throw IoError("cannot stat " + path); }
    size_ = (uint64_t)st.st_size; path_ = path;
    if (size_ > 0) {
        void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { ::close(fd); throw IoError("cannot mmap " + path); }
        madvise(p, size_, MADV_SEQUENTIAL);
        data_ = (const uint8_t*)p;
    }
    ::close(fd);
}

BgzfStream::BgzfStream(const MappedFile& f) : f_(f), ubuf_(65536 + 64), fast_(new pjinflate::Inflater()) {
    z_stream* z = new z_stream; memset(z, 0, sizeof *z);
    if (inflateInit2(z, -15) != Z_OK) { delete z; throw IoError("inflateInit2 failed"); }
    z_ = z;
}
BgzfStream::~BgzfStream() { z_stream* z = (z_stream*)z_; inflateEnd(z); delete z; delete (pjinflate::Inflater*)fast_; }

bool BgzfStream::load_block(uint64_t coff) {
    have_block_ = false; ulen_ = upos_ = 0; block_coff_ = coff; next_coff_ = coff;
    if (coff + 18 > f_.size()) return false;
    const uint8_t* p = f_.data() + coff;
    if (p[0] != 31 || p[1] != 139 || p[2] != 8 || !(p[3] & 4)) throw IoError("not a BGZF block in " + f_.path());
    uint32_t xlen = rd16(p + 10);
    if (coff + 12 + xlen > f_.size()) throw IoError("truncated BGZF header");
    uint32_t bsize = 0; bool found = false;
    for (uint32_t o = 0; o + 4 <= xlen;) {
        const uint8_t* e = p + 12 + o; uint32_t slen = rd16(e + 2);
        if (e[0] == 'B' && e[1] == 'C' && slen == 2) { bsize = (uint32_t)rd16(e + 4) + 1; found = true; }
        o += 4 + slen;
    }
    if (!found || coff + bsize > f_.size() || bsize < 12 + xlen + 8) throw IoError("corrupt BGZF block in " + f_.path());
    uint32_t isize = rd32(p + bsize - 4);
    if (isize > 65536) throw IoError("BGZF block too large");
    // fast path: our own DEFLATE decoder (inflate_fast.hpp); anything it does not accept is decoded again by zlib
    const bool fast_ok = isize && coff + bsize + 16 <= f_.size() &&
                         ((pjinflate::Inflater*)fast_)->run(p + 12 + xlen, bsize - 12 - xlen - 8, ubuf_.data(), isize);
    if (isize && !fast_ok) {
        z_stream* z = (z_stream*)z_;
        inflateReset(z);
        z->next_in = (Bytef*)(p + 12 + xlen); z->avail_in = bsize - 12 - xlen - 8;
        z->next_out = ubuf_.data(); z->avail_out = 65536;
        int rc = inflate(z, Z_FINISH);
        if (rc != Z_STREAM_END || z->total_out != isize) throw IoError("BGZF inflate failed in " + f_.path());
    }
    // CRC32 trailer, like htslib's bgzf_read_block: a corrupted block (or a decoder bug) must not reach the junction counts
    if (isize && crc32_block(ubuf_.data(), isize) != rd32(p + bsize - 8)) throw IoError("BGZF block CRC32 mismatch in " + f_.path());
    ulen_ = isize; upos_ = 0; next_coff_ = coff + bsize; have_block_ = true;
    return true;
}

void BgzfStream::seek(uint64_t voff) {
    uint64_t coff = voff >> 16; uint32_t uoff = (uint32_t)(voff & 0xffff);
    if (!have_block_ || coff != block_coff_) { if (!load_block(coff)) { ulen_ = upos_ = 0; return; } }
    upos_ = std::min(uoff, ulen_);
}

uint64_t BgzfStream::tell() const {
    if (have_block_ && upos_ >= ulen_) return next_coff_ << 16;   // normalised like htslib after a block is drained
    return (block_coff_ << 16) | upos_;
}

size_t BgzfStream::read(void* dst, size_t n) {
    uint8_t* d = (uint8_t*)dst; size_t got = 0;
    while (got < n) {
        if (!have_block_ || upos_ >= ulen_) {
            uint64_t nc = have_block_ ? next_coff_ : block_coff_;
            if (!load_block(nc)) break;
            if (ulen_ == 0) continue;      // empty block (EOF marker or padding): keep going
        }
        size_t k = std::min<size_t>(n - got, ulen_ - upos_);
        memcpy(d + got, ubuf_.data() + upos_, k);
        got += k; upos_ += (uint32_t)k;
    }
    return got;
}

bool BgzfStream::eof() {
    while (!have_block_ || upos_ >= ulen_) {
        uint64_t nc = have_block_ ? next_coff_ : block_coff_;
        if (!load_block(nc)) return true;
    }
    return false;
}

bool BamHeader::coordinate_sorted() const {
    size_t e = text.find('\n');
    std::string first = text.substr(0, e);
    return first.compare(0, 3, "@HD") == 0 && first.find("SO:coordinate") != std::string::npos;
}

void ColumnarChunk::clear() {
    tid.clear(); pos.clear(); l_qseq.clear(); mtid.clear(); mpos.clear(); flag.clear(); mapq.clear(); xs.clear();
    cigar_off.assign(1, 0); cigar.clear(); seq_off.assign(1, 0); seq4.clear(); name_code.clear();
    n_cigar.clear(); seq2.clear(); seqx_pos.clear(); seqx_code.clear(); runs.clear();
}

uint64_t name_code(const char* qname, size_t len, uint16_t flag) {
    uint64_t h = 0xCBF29CE484222325ull;
    auto eat = [&](const char* s, size_t n) { for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 0x100000001B3ull; } };
    eat(qname, len);
    if (flag & 0x1) eat((flag & 0x40) ? "_R1" : (flag & 0x80) ? "_R2" : "_R?", 3);
    h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33;
    return h;
}

void ColumnarChunk::append(const ColumnarChunk& o) {
    if (lean) {
        const int64_t r0 = n(), c0 = (int64_t)cigar.size(), s0 = (int64_t)seq2.size(), x0 = (int64_t)seqx_pos.size();
        for (const Run& r : o.runs) {
            if (!runs.empty() && runs.back().tid == r.tid) continue;                  // the target continues: same stretch
            runs.push_back(Run{r.tid, 0, r0 + r.rec0, c0 + r.cig0, s0 + r.seq0, x0 + r.seqx0});
        }
        pos.insert(pos.end(), o.pos.begin(), o.pos.end()); l_qseq.insert(l_qseq.end(), o.l_qseq.begin(), o.l_qseq.end());
        mtid.insert(mtid.end(), o.mtid.begin(), o.mtid.end()); mpos.insert(mpos.end(), o.mpos.begin(), o.mpos.end());
        flag.insert(flag.end(), o.flag.begin(), o.flag.end()); mapq.insert(mapq.end(), o.mapq.begin(), o.mapq.end()); xs.insert(xs.end(), o.xs.begin(), o.xs.end());
        n_cigar.insert(n_cigar.end(), o.n_cigar.begin(), o.n_cigar.end()); cigar.insert(cigar.end(), o.cigar.begin(), o.cigar.end());
        seq2.insert(seq2.end(), o.seq2.begin(), o.seq2.end());
        for (uint64_t v : o.seqx_pos) seqx_pos.push_back(v + (uint64_t)s0 * 4ull);
        seqx_code.insert(seqx_code.end(), o.seqx_code.begin(), o.seqx_code.end());
        name_code.insert(name_code.end(), o.name_code.begin(), o.name_code.end());
        return;
    }
    uint32_t cb = (uint32_t)cigar.size(); uint64_t sb = seq4.size();
    tid.insert(tid.end(), o.tid.begin(), o.tid.end()); pos.insert(pos.end(), o.pos.begin(), o.pos.end());
    l_qseq.insert(l_qseq.end(), o.l_qseq.begin(), o.l_qseq.end());
    mtid.insert(mtid.end(), o.mtid.begin(), o.mtid.end()); mpos.insert(mpos.end(), o.mpos.begin(), o.mpos.end());
    flag.insert(flag.end(), o.flag.begin(), o.flag.end()); mapq.insert(mapq.end(), o.mapq.begin(), o.mapq.end());
    xs.insert(xs.end(), o.xs.begin(), o.xs.end());
    for (size_t i = 1; i < o.cigar_off.size(); i++) cigar_off.push_back(cb + o.cigar_off[i]);
    for (size_t i = 1; i < o.seq_off.size(); i++) seq_off.push_back(sb + o.seq_off[i]);
    cigar.insert(cigar.end(), o.cigar.begin(), o.cigar.end());
    seq4.insert(seq4.end(), o.seq4.begin(), o.seq4.end());
    name_code.insert(name_code.end(), o.name_code.begin(), o.name_code.end());
}

void BamFile::open(const std::string& bam_path) {
    file_.open(bam_path);
    BgzfStream s(file_);
    s.seek(0);
    uint8_t b4[4];
    if (s.read(b4, 4) != 4 || memcmp(b4, "BAM\1", 4) != 0) throw IoError("not a BAM file: " + bam_path);
    if (s.read(b4, 4) != 4) throw IoError("truncated BAM header");
    uint32_t l_text = rd32(b4);
    // sizes in a corrupted header must not turn into giant allocations: DEFLATE cannot expand more than about 1032:1
    const uint64_t max_unc = (uint64_t)file_.size() * 1100ull + 65536ull;
    if (l_text > max_unc) throw IoError("corrupt BAM header (l_text) in " + bam_path);
    hdr_.text.resize(l_text);
    if (l_text && s.read(&hdr_.text[0], l_text) != l_text) throw IoError("truncated BAM header text");
    while (!hdr_.text.empty() && hdr_.text.back() == '\0') hdr_.text.pop_back();
    if (s.read(b4, 4) != 4) throw IoError("truncated BAM header");
    uint32_t n_ref = rd32(b4);
    if ((uint64_t)n_ref * 9ull > max_unc) throw IoError("corrupt BAM header (n_ref) in " + bam_path);
    hdr_.names.resize(n_ref); hdr_.lens.resize(n_ref);
    for (uint32_t i = 0; i < n_ref; i++) {
        if (s.read(b4, 4) != 4) throw IoError("truncated BAM header");
        uint32_t l_name = rd32(b4);
        if (l_name > max_unc) throw IoError("corrupt BAM header (l_name) in " + bam_path);
        std::string nm(l_name, '\0');
        if (l_name && s.read(&nm[0], l_name) != l_name) throw IoError("truncated BAM header");
        while (!nm.empty() && nm.back() == '\0') nm.pop_back();
        if (s.read(b4, 4) != 4) throw IoError("truncated BAM header");
        hdr_.names[i] = nm; hdr_.lens[i] = (int32_t)rd32(b4);
    }
    hdr_.first_record_voff = s.tell();
}

bool BamFile::load_bai(const std::string& bai_path) {
    MappedFile f;
    try { f.open(bai_path); } catch (const IoError&) { return false; }
    const uint8_t* p = f.data(); uint64_t n = f.size(), o = 0;
    auto need = [&](uint64_t k) { if (o + k > n) throw IoError("truncated BAI: " + bai_path); };
    need(8);
    if (memcmp(p, "BAI\1", 4) != 0) throw IoError("not a BAI index: " + bai_path);
    uint32_t n_ref = rd32(p + 4); o = 8;
    if ((uint64_t)n_ref * 8ull > n - o) throw IoError("corrupt BAI (n_ref): " + bai_path);          // every reference takes at least n_bin + n_intv
    if (!hdr_.lens.empty() && n_ref != hdr_.lens.size()) throw IoError("the BAI index does not belong to this BAM (" + std::to_string(n_ref) + " references, header has " + std::to_string(hdr_.lens.size()) + "): " + bai_path);
    idx_.assign(n_ref, BamTargetIndex());
    for (uint32_t r = 0; r < n_ref; r++) {
        BamTargetIndex& t = idx_[r];
        need(4); uint32_t n_bin = rd32(p + o); o += 4;
        uint64_t first = 0;
        for (uint32_t b = 0; b < n_bin; b++) {
            need(8); uint32_t bin = rd32(p + o); uint32_t n_chunk = rd32(p + o + 4); o += 8;
            need(16ull * n_chunk);
            if (bin == 37450) {             // metadata pseudo-bin: (off_beg, off_end), (n_mapped, n_unmapped)
                if (n_chunk >= 2) { t.n_mapped = rd64(p + o + 16); t.n_unmapped = rd64(p + o + 24); t.has_counts = true; }
            } else {
                for (uint32_t c = 0; c < n_chunk; c++) {
                    uint64_t beg = rd64(p + o + 16ull * c);
                    if (first == 0 || beg < first) first = beg;
                }
            }
            o += 16ull * n_chunk;
        }
        need(4); uint32_t n_intv = rd32(p + o); o += 4;
        need(8ull * n_intv);
        t.ioffset.resize(n_intv);
        for (uint32_t i = 0; i < n_intv; i++) t.ioffset[i] = rd64(p + o + 8ull * i);
        o += 8ull * n_intv;
        t.first_voff = first;
    }
    return true;
}

bool BamFile::load_csi(const std::string& csi_path) {
    MappedFile f;
    try { f.open(csi_path); } catch (const IoError&) { return false; }
    std::vector<uint8_t> buf;
    {   // the whole index is one BGZF stream
        BgzfStream s(f); s.seek(0);
        uint8_t tmp[65536]; size_t g;
        while ((g = s.read(tmp, sizeof tmp)) > 0) buf.insert(buf.end(), tmp, tmp + g);
    }
    const uint8_t* p = buf.data(); uint64_t n = buf.size(), o = 0;
    auto need = [&](uint64_t k) { if (o + k > n) throw IoError("truncated CSI: " + csi_path); };
    need(16);
    if (memcmp(p, "CSI\1", 4) != 0) throw IoError("not a CSI index: " + csi_path);
    const int32_t min_shift = (int32_t)rd32(p + 4), depth = (int32_t)rd32(p + 8); const uint32_t l_aux = rd32(p + 12);
    if (min_shift < 1 || min_shift > 30 || depth < 0 || depth > 10) throw IoError("unsupported CSI geometry: " + csi_path);   // depth 0: htslib writes it for targets shorter than one window
    o = 16; need(l_aux + 4ull); o += l_aux;
    const uint32_t n_ref = rd32(p + o); o += 4;
    if ((uint64_t)n_ref * 4ull > n - o) throw IoError("corrupt CSI (n_ref): " + csi_path);
    if (!hdr_.lens.empty() && n_ref != hdr_.lens.size()) throw IoError("the CSI index does not belong to this BAM (" + std::to_string(n_ref) + " references, header has " + std::to_string(hdr_.lens.size()) + "): " + csi_path);
    auto first_bin = [](int lvl) { return (uint64_t)(((1ull << (3 * lvl)) - 1) / 7); };
    const uint64_t meta_bin = first_bin(depth + 1) + 1;
    idx_.assign(n_ref, BamTargetIndex());
    for (uint32_t r = 0; r < n_ref; r++) {
        BamTargetIndex& t = idx_[r];
        t.window_shift = min_shift;
        need(4); const uint32_t n_bin = rd32(p + o); o += 4;
        uint64_t first = 0;
        for (uint32_t b = 0; b < n_bin; b++) {
            need(16); const uint64_t bin = rd32(p + o); const uint64_t loff = rd64(p + o + 4); const uint32_t n_chunk = rd32(p + o + 12); o += 16;
            need(16ull * n_chunk);
            if (bin == meta_bin) {
                if (n_chunk >= 2) { t.n_mapped = rd64(p + o + 16); t.n_unmapped = rd64(p + o + 24); t.has_counts = true; }
            } else {
                for (uint32_t c = 0; c < n_chunk; c++) { const uint64_t beg = rd64(p + o + 16ull * c); if (first == 0 || beg < first) first = beg; }
                // loffset = linear-index value of the bin's first finest-level window (htslib update_loff)
                int lvl = depth; while (lvl > 0 && bin < first_bin(lvl)) lvl--;
                const uint64_t w = (bin - first_bin(lvl)) << (3 * (depth - lvl));
                if (loff && w < (1ull << 28)) {
                    if (t.ioffset.size() <= w) t.ioffset.resize((size_t)w + 1, 0);
                    if (t.ioffset[w] == 0 || loff < t.ioffset[w]) t.ioffset[w] = loff;
                }
            }
            o += 16ull * n_chunk;
        }
        t.first_voff = first;
    }
    return true;
}

void BamFile::plan_target(int32_t tid, uint64_t chunk_bytes, std::vector<DecodeTask>& out) const {
    if (tid < 0 || (size_t)tid >= idx_.size()) return;
    const BamTargetIndex& t = idx_[tid];
    if (t.first_voff == 0) return;          // no records on this target
    const int64_t W = 1ll << t.window_shift;
    DecodeTask cur{tid, 0, INT32_MAX, t.first_voff, 0};
    uint64_t cur_c = t.first_voff >> 16;
    for (size_t w = 1; w < t.ioffset.size(); w++) {
        uint64_t v = t.ioffset[w];
        if (v == 0) continue;
        uint64_t c = v >> 16;
        if (c >= cur_c + chunk_bytes && v > cur.voff && (int64_t)w * W < (int64_t)INT32_MAX) {
            cur.pos_hi = (int32_t)(w * W); cur.approx_bytes = c - cur_c;
            out.push_back(cur);
            cur = DecodeTask{tid, (int32_t)(w * W), INT32_MAX, v, 0}; cur_c = c;
        }
    }
    // size of the tail: up to the next target's first record (or EOF)
    uint64_t end_c = file_.size();
    for (size_t r = (size_t)tid + 1; r < idx_.size(); r++) if (idx_[r].first_voff) { end_c = idx_[r].first_voff >> 16; break; }
    cur.approx_bytes = end_c > cur_c ? end_c - cur_c : 1;
    out.push_back(cur);
}

DecodeTask BamFile::whole_file_task() const {
    return DecodeTask{-1, 0, INT32_MAX, hdr_.first_record_voff, file_.size()};
}

// XS aux lookup: the only tag the junc path reads (lib/src/bam_alignment.cc:226-231).
// Returns 0 when absent, the character for XS:A, and throws DataError for any other XS type
// (bam_aux2A yields 0 -> strandFromChar throws, bam_master.hpp:60-72).
static uint8_t find_xs(const uint8_t* a, const uint8_t* end) {
    while (a + 3 <= end) {
        uint8_t t0 = a[0], t1 = a[1], ty = a[2];
        const uint8_t* v = a + 3;
        if (t0 == 'X' && t1 == 'S') {
            if (ty != 'A' || v >= end) throw DataError(std::string("Unknown strand: XS tag is not of type A"));
            uint8_t c = *v;
            if (c != '+' && c != '-' && c != '?' && c != '.') throw DataError(std::string("Unknown strand: ") + (char)c);
            return c;
        }
        size_t sz;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'd': sz = 8; break;
        case 'Z': case 'H': { const uint8_t* q = v; while (q < end && *q) q++; sz = (size_t)(q - v) + 1; break; }
        case 'B': {
            if (v + 5 > end) return 0;
            uint8_t st = v[0]; uint32_t cnt = rd32(v + 1);
            size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
            sz = 5 + es * (size_t)cnt; break;
        }
        default: return 0;                 // unknown type: stop scanning like a failed bam_aux_get
        }
        a = v + sz;
    }
    return 0;
}

// ---- 4-bit BAM SEQ -> 2 bits per base (lean batches) ----
// out[j] holds bases 4j .. 4j+3 (base q at bits 2(q & 3)); anything that is not A/C/G/T becomes 0 and makes the function return true
// ("look closer": the padding nibble of an odd length also triggers it).
namespace {
struct Seq2Lut { uint8_t v[256]; Seq2Lut() { for (int b = 0; b < 256; b++) { auto c2 = [](int nib, bool& bad) { switch (nib) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; default: bad = true; return 0; } };
                                                                          bool bad = false; const int hi = c2(b >> 4, bad), lo = c2(b & 15, bad); v[b] = (uint8_t)(hi | (lo << 2) | (bad ? 0x80 : 0)); } } };
const Seq2Lut SEQ2_LUT;
inline bool seq4_to_2_scalar(const uint8_t* sq, size_t j0, size_t nb4, size_t nb2, uint8_t* o2) {
    bool any_bad = false;
    for (size_t j = j0; j < nb2; j++) {
        const uint8_t a = SEQ2_LUT.v[sq[2 * j]], b2 = (2 * j + 1 < nb4) ? SEQ2_LUT.v[sq[2 * j + 1]] : (uint8_t)0;
        o2[j] = (uint8_t)((a & 15) | ((b2 & 15) << 4));
        any_bad |= ((a | b2) & 0x80) != 0;
    }
    return any_bad;
}
#if defined(__GNUC__) && defined(__x86_64__)
// 32 bases per step: two nibble look-ups (pshufb), the pairs of 4-bit results folded into bytes with one multiply-add
__attribute__((target("ssse3"))) bool seq4_to_2_ssse3(const uint8_t* sq, size_t nb4, size_t nb2, uint8_t* o2) {
    const __m128i lut_code = _mm_setr_epi8(0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 0);
    const __m128i lut_bad = _mm_setr_epi8((char)0x80, 0, 0, (char)0x80, 0, (char)0x80, (char)0x80, (char)0x80, 0, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80, (char)0x80);
    const __m128i m4 = _mm_set1_epi8(0x0f), fold = _mm_set1_epi16(0x1001);
    __m128i bad = _mm_setzero_si128();
    size_t j = 0;                                              // input byte index
    for (; j + 16 <= nb4; j += 16) {
        const __m128i in = _mm_loadu_si128((const __m128i*)(sq + j));
        const __m128i hi = _mm_and_si128(_mm_srli_epi16(in, 4), m4), lo = _mm_and_si128(in, m4);
        const __m128i v = _mm_or_si128(_mm_shuffle_epi8(lut_code, hi), _mm_slli_epi16(_mm_shuffle_epi8(lut_code, lo), 2));   // first base | second base << 2
        bad = _mm_or_si128(bad, _mm_or_si128(_mm_shuffle_epi8(lut_bad, hi), _mm_shuffle_epi8(lut_bad, lo)));
        const __m128i w = _mm_maddubs_epi16(v, fold);          // v[2k] + 16 * v[2k + 1] in 16-bit lanes
        _mm_storel_epi64((__m128i*)(o2 + j / 2), _mm_packus_epi16(w, w));
    }
    bool any_bad = _mm_movemask_epi8(bad) != 0;
    any_bad |= seq4_to_2_scalar(sq, j / 2, nb4, nb2, o2);
    return any_bad;
}
const bool HAVE_SSSE3 = __builtin_cpu_supports("ssse3");
#else
const bool HAVE_SSSE3 = false;
inline bool seq4_to_2_ssse3(const uint8_t* sq, size_t nb4, size_t nb2, uint8_t* o2) { return seq4_to_2_scalar(sq, 0, nb4, nb2, o2); }
#endif
inline bool seq4_to_2(const uint8_t* sq, size_t nb4, size_t nb2, uint8_t* o2) { return HAVE_SSSE3 ? seq4_to_2_ssse3(sq, nb4, nb2, o2) : seq4_to_2_scalar(sq, 0, nb4, nb2, o2); }
} // namespace

void BamFile::decode(const DecodeTask& task, ColumnarChunk& out) const {
    BgzfStream s(file_);
    s.seek(task.voff);
    std::vector<uint8_t> rec;
    const int32_t n_ref = (int32_t)hdr_.lens.size();
    // Lean columns are written through raw pointers into vectors grown in big steps (one capacity check per record instead of one
    // per push_back — eight per record); the vectors are cut back to the filled length when the task is done, also on an exception.
    struct LeanCols {
        ColumnarChunk& o; int64_t n, cap; size_t nc, ccap, ns, scap;
        int32_t *pos = nullptr, *lq = nullptr, *mtid = nullptr, *mpos = nullptr; uint16_t *flag = nullptr, *ncig = nullptr; uint8_t *mapq = nullptr, *xs = nullptr;
        uint64_t* name = nullptr; uint32_t* cig = nullptr; uint8_t* seq2 = nullptr;
        explicit LeanCols(ColumnarChunk& c) : o(c), n(c.n()), cap(c.n()), nc(c.cigar.size()), ccap(nc), ns(c.seq2.size()), scap(ns) {}
        void size_records(size_t k) {
            o.pos.resize(k); o.flag.resize(k); o.mapq.resize(k); o.l_qseq.resize(k); o.xs.resize(k); o.n_cigar.resize(k);
            if (o.keep_mate) { o.mtid.resize(k); o.mpos.resize(k); }
            if (o.with_names) o.name_code.resize(k);
        }
        void grow_records() {
            cap = std::max<int64_t>(cap * 2, 8192); size_records((size_t)cap);
            pos = o.pos.data(); flag = o.flag.data(); mapq = o.mapq.data(); lq = o.l_qseq.data(); xs = o.xs.data(); ncig = o.n_cigar.data();
            mtid = o.mtid.data(); mpos = o.mpos.data(); name = o.name_code.data();
        }
        void grow_cigar(size_t need) { ccap = std::max<size_t>(std::max<size_t>(ccap * 2, nc + need), 32768); o.cigar.resize(ccap); cig = o.cigar.data(); }
        void grow_seq(size_t need) { scap = std::max<size_t>(std::max<size_t>(scap * 2, ns + need), (size_t)1 << 19); o.seq2.resize(scap); seq2 = o.seq2.data(); }
        ~LeanCols() { if (o.lean) { size_records((size_t)n); o.cigar.resize(nc); o.seq2.resize(ns); } }
    } L(out);
    for (;;) {
        if (task.end_voff && s.tell() >= task.end_voff) break;   // the records from the gap cut on belong to the next slice
        uint32_t bs;
        if (const uint8_t* h = s.take(4)) bs = rd32(h);
        else {
            uint8_t b4[4];
            size_t g = s.read(b4, 4);
            if (g == 0) break;
            if (g != 4) throw IoError("truncated BAM record in " + file_.path());
            bs = rd32(b4);
        }
        if (bs < 32 || bs > (1u << 29)) throw IoError("corrupt BAM record (block_size " + std::to_string(bs) + ")");
        const uint8_t* p = s.take(bs);                      // most records lie inside one BGZF block: parse them in place
        if (!p) {
            rec.resize(bs);
            if (s.read(rec.data(), bs) != bs) throw IoError("truncated BAM record in " + file_.path());
            p = rec.data();
        }
        int32_t tid = (int32_t)rd32(p), pos = (int32_t)rd32(p + 4);
        uint32_t l_name = p[8]; uint8_t mapq = p[9];
        uint32_t n_cig = rd16(p + 12); uint16_t flag = rd16(p + 14);
        int32_t l_seq = (int32_t)rd32(p + 16);
        int32_t mtid = (int32_t)rd32(p + 20), mpos = (int32_t)rd32(p + 24);
        if (task.tid >= 0) {
            if (tid != task.tid) { if (tid > task.tid || tid < 0) break; else continue; }
            if (pos >= task.pos_hi) break;
            if (pos < task.pos_lo) continue;
        }
        if (tid < 0) { if (task.tid < 0) break; else continue; }   // unplaced reads sort last
        if (tid >= n_ref) throw IoError("BAM record refers to an unknown target");
        uint64_t need = 32ull + l_name + 4ull * n_cig + (uint64_t)((l_seq + 1) / 2) + (uint64_t)(l_seq < 0 ? 0 : l_seq);
        if (l_seq < 0 || need > bs) throw IoError("corrupt BAM record (fields exceed block_size)");
        const uint8_t* cg = p + 32 + l_name;
        // record visibility, Q13: pos < target_len and endpos > 0 (hts.c:1951-1953, sam.c:336-342)
        int64_t rlen = 0; bool spliced = false;
        for (uint32_t k = 0; k < n_cig; k++) {
            uint32_t c = rd32(cg + 4 * k); uint32_t op = c & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
            if (op == 3) spliced = true;
        }
        if (pos >= hdr_.lens[tid]) { if (task.tid >= 0) break; else continue; }
        int64_t endpos = (!(flag & 0x4) && n_cig > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1;
        if (endpos <= 0) continue;
        const uint8_t* sq = cg + 4ull * n_cig;
        const uint8_t* aux = sq + (l_seq + 1) / 2 + l_seq;
        if (out.lean) {
            if (out.runs.empty() || out.runs.back().tid != tid)
                out.runs.push_back(ColumnarChunk::Run{tid, 0, L.n, (int64_t)L.nc, (int64_t)L.ns, (int64_t)out.seqx_pos.size()});
            if (n_cig > 0xffffu) throw IoError("BAM record with more than 65535 CIGAR operations");
            if (L.n == L.cap) L.grow_records();
            uint16_t fl = flag;
            if (spliced && l_seq > 0) {
                // 4-bit BAM SEQ -> 2 bits per base; anything that is not A/C/G/T is stored as 0 and listed as an exception
                const size_t nb4 = (size_t)(l_seq + 1) / 2, nb2 = (size_t)(l_seq + 3) / 4, s0 = L.ns;
                if (L.ns + nb2 + 16 > L.scap) L.grow_seq(nb2 + 16);            // the 32-base steps store 8 bytes at a time
                if (seq4_to_2(sq, nb4, nb2, L.seq2 + s0)) {                    // rare: list the exact nibbles (the padding nibble of an odd length is not a base)
                    bool real = false;
                    for (int32_t q = 0; q < l_seq; q++) {
                        const uint32_t nib = (q & 1) ? (sq[q >> 1] & 15u) : (uint32_t)(sq[q >> 1] >> 4);
                        if (nib != 1 && nib != 2 && nib != 4 && nib != 8) { out.seqx_pos.push_back((uint64_t)s0 * 4ull + (uint64_t)q); out.seqx_code.push_back((uint8_t)nib); real = true; }
                    }
                    if (real) fl = (uint16_t)(fl | 0x8000u);
                }
                L.ns += nb2;
            }
            const int64_t k = L.n;
            L.pos[k] = pos; L.flag[k] = fl; L.mapq[k] = mapq; L.lq[k] = l_seq; L.ncig[k] = (uint16_t)n_cig;
            if (out.keep_mate) { L.mtid[k] = mtid; L.mpos[k] = mpos; }
            L.xs[k] = find_xs(aux, p + bs);
            if (out.with_names) {
                size_t ln = l_name; const char* qn = (const char*)p + 32;
                while (ln && qn[ln - 1] == 0) ln--;
                L.name[k] = name_code(qn, strnlen(qn, ln), flag);
            }
            if (L.nc + n_cig > L.ccap) L.grow_cigar(n_cig);
            if (n_cig) memcpy(L.cig + L.nc, cg, 4ull * n_cig);
            L.nc += n_cig;
            L.n = k + 1;
            continue;
        }
        out.tid.push_back(tid); out.pos.push_back(pos); out.flag.push_back(flag); out.mapq.push_back(mapq);
        out.l_qseq.push_back(l_seq); out.mtid.push_back(mtid); out.mpos.push_back(mpos);
        out.xs.push_back(find_xs(aux, p + bs));
        if (out.with_names) {
            size_t ln = l_name; const char* qn = (const char*)p + 32;
            while (ln && qn[ln - 1] == 0) ln--;                 // l_read_name counts the NUL (and htslib >= 1.5 pads with more)
            out.name_code.push_back(name_code(qn, strnlen(qn, ln), flag));
        }
        size_t c0 = out.cigar.size(); out.cigar.resize(c0 + n_cig);
        if (n_cig) memcpy(&out.cigar[c0], cg, 4ull * n_cig);
        out.cigar_off.push_back((uint32_t)out.cigar.size());
        if (spliced && l_seq > 0) {
            size_t s0 = out.seq4.size(), nb = (size_t)(l_seq + 1) / 2;
            out.seq4.resize(s0 + nb); memcpy(&out.seq4[s0], sq, nb);
        }
        out.seq_off.push_back((uint64_t)out.seq4.size());
    }
}

bool BamFile::find_gap_cut(int32_t want_tid, int32_t pos_lo, uint64_t start_voff, uint64_t* cut_voff, int32_t* cut_pos) const {
    BgzfStream s(file_);
    s.seek(start_voff);
    std::vector<uint8_t> rec;
    int64_t max_end = -1;                                    // last reference base covered by a spliced record seen so far
    for (;;) {
        const uint64_t here = s.tell();
        uint32_t bs;
        if (const uint8_t* h = s.take(4)) bs = rd32(h);
        else {
            uint8_t b4[4];
            size_t g = s.read(b4, 4);
            if (g == 0) return false;
            if (g != 4) throw IoError("truncated BAM record in " + file_.path());
            bs = rd32(b4);
        }
        if (bs < 32 || bs > (1u << 29)) throw IoError("corrupt BAM record (block_size " + std::to_string(bs) + ")");
        const uint8_t* p = s.take(bs);
        if (!p) {
            rec.resize(bs);
            if (s.read(rec.data(), bs) != bs) throw IoError("truncated BAM record in " + file_.path());
            p = rec.data();
        }
        const int32_t tid = (int32_t)rd32(p), pos = (int32_t)rd32(p + 4);
        if (tid != want_tid) { if (tid > want_tid || tid < 0) return false; else continue; }
        if (pos >= hdr_.lens[(size_t)tid]) return false;     // decode() stops the target here too (Q13)
        if (pos >= pos_lo && (int64_t)pos > max_end) { *cut_voff = here; *cut_pos = pos; return true; }
        const uint32_t l_name = p[8], n_cig = rd16(p + 12);
        if (32ull + l_name + 4ull * n_cig > bs) throw IoError("corrupt BAM record (fields exceed block_size)");
        const uint8_t* cg = p + 32 + l_name;
        int64_t rlen = 0; bool spliced = false;
        for (uint32_t k = 0; k < n_cig; k++) {
            const uint32_t c = rd32(cg + 4 * k), op = c & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;
            if (op == 3) spliced = true;
        }
        if (spliced) max_end = std::max<int64_t>(max_end, (int64_t)pos + rlen - 1);
    }
}

} // namespace pjio

// Self-test of the fast inflater against zlib on synthetic streams (several data shapes, levels and strategies).
// Returns the number of mismatches; exported through pjh_inflate_selftest for the CPU test-suite.
int pjio::inflate_selftest(int n_cases) {
    pjinflate::Inflater inf;
    uint64_t rs = 0x9E3779B97F4A7C15ull;
    auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 16); };
    int bad = 0;
    static const int levels[] = {0, 1, 6, 9};
    static const int strats[] = {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE, Z_FILTERED};
    for (int t = 0; t < n_cases; t++) {
        const size_t n = 1 + rnd() % 65000; std::vector<uint8_t> d(n); const int mode = t % 5;
        for (size_t i = 0; i < n; i++) {
            if (mode == 0) d[i] = (uint8_t)rnd();
            else if (mode == 1) d[i] = (uint8_t)"ACGT"[rnd() & 3];
            else if (mode == 2) d[i] = (i % 97 < 50) ? 0xff : (uint8_t)(rnd() & 15);
            else if (mode == 3) d[i] = (i > 200 && rnd() % 3) ? d[i - 1 - (rnd() % 200)] : (uint8_t)(rnd() & 63);
            else d[i] = 0;
        }
        z_stream z; memset(&z, 0, sizeof z);
        if (deflateInit2(&z, levels[t % 4], Z_DEFLATED, -15, 8, strats[(t / 4) % 5]) != Z_OK) return -1;
        std::vector<uint8_t> c(n * 2 + 1024);
        z.next_in = d.data(); z.avail_in = (uInt)n; z.next_out = c.data(); z.avail_out = (uInt)c.size();
        deflate(&z, Z_FINISH); const size_t cn = z.total_out; deflateEnd(&z);
        std::vector<uint8_t> o(n + 64);
        if (!inf.run(c.data(), cn, o.data(), n) || memcmp(o.data(), d.data(), n) != 0) bad++;
        if (crc32_block(d.data(), n) != (uint32_t)crc32(0L, d.data(), (uInt)n)) bad++;     // crc32_fast.hpp against zlib
        // corrupted and truncated input must never read or write out of bounds (the result itself is irrelevant: the caller
        // falls back to zlib whenever run() returns false, and BGZF carries its own size trailer)
        if (cn > 8) { c[cn / 2] ^= 0x55; std::vector<uint8_t> o2(n + 64); (void)inf.run(c.data(), cn, o2.data(), n); (void)inf.run(c.data(), cn / 2, o2.data(), n); }
    }
    return bad;
}
