// bam_io.hpp — host-side BGZF / BAM / BAI reading for the junc path.
//
// Replaces the reference's BamReader (lib/src/bam_reader.cc:78-146) + htslib iterator
// (deps/htslib-1.3/hts.c:1923-1963) for ONE purpose: decode alignment records of a prepared,
// coordinate-sorted BAM into the columnar layout of include/portcullis_junc.h, in parallel.
// Written from the SAM/BAM specification; zlib is the only dependency.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <stdexcept>
#include <functional>

namespace pjio {

struct IoError : std::runtime_error { using std::runtime_error::runtime_error; };
// Input the reference itself aborts on (e.g. a non-'A' XS tag).
struct DataError : std::runtime_error { using std::runtime_error::runtime_error; };

// Read-only memory map of a file (shared by all decode threads).
class MappedFile {
public:
    MappedFile() = default;
    ~MappedFile();
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    void open(const std::string& path);
    const uint8_t* data() const { return data_; }
    uint64_t size() const { return size_; }
    const std::string& path() const { return path_; }
private:
    const uint8_t* data_ = nullptr; uint64_t size_ = 0; std::string path_;
};

// Sequential BGZF inflater over a MappedFile, addressed by virtual offsets (coffset<<16 | uoffset).
class BgzfStream {
public:
    explicit BgzfStream(const MappedFile& f);
    ~BgzfStream();
    void seek(uint64_t voff);
    uint64_t tell() const;                 // virtual offset of the next byte
    size_t read(void* dst, size_t n);      // returns bytes delivered (< n only at EOF)
    // n bytes straight out of the current block when it holds them all (no copy; valid until the next call), else nullptr
    const uint8_t* take(size_t n) { if (have_block_ && (size_t)(ulen_ - upos_) >= n) { const uint8_t* p = ubuf_.data() + upos_; upos_ += (uint32_t)n; return p; } return nullptr; }
    bool eof();
private:
    bool load_block(uint64_t coff);
    const MappedFile& f_;
    void* z_ = nullptr;                    // z_stream (fallback decoder)
    uint64_t block_coff_ = 0, next_coff_ = 0;
    std::vector<uint8_t> ubuf_; uint32_t ulen_ = 0, upos_ = 0;
    void* fast_ = nullptr;                 // pjinflate::Inflater (inflate_fast.hpp)
    bool have_block_ = false;
};

struct BamTargetIndex {
    uint64_t first_voff = 0;               // virtual offset of the target's first record (0 = no records)
    uint64_t n_mapped = 0, n_unmapped = 0; // from the metadata pseudo-bin, 0 if absent
    bool has_counts = false;
    std::vector<uint64_t> ioffset;         // linear index: smallest virtual offset of records overlapping each window (0 = none)
    int window_shift = 14;                 // window = 1 << window_shift bases (16 kb for BAI; CSI min_shift)
};

struct BamHeader {
    std::string text;
    std::vector<std::string> names;
    std::vector<int32_t> lens;
    uint64_t first_record_voff = 0;
    bool coordinate_sorted() const;
};

// One unit of parallel decode: records of `tid` with pos in [pos_lo, pos_hi), reading starts at voff.
struct DecodeTask {
    int32_t tid; int32_t pos_lo; int32_t pos_hi; uint64_t voff;
    uint64_t approx_bytes;                 // compressed-size estimate, for scheduling
    uint64_t end_voff = 0;                 // != 0: stop at the first record whose virtual offset is >= end_voff (a gap cut, find_gap_cut)
};

// Columnar records (same columns as pj_batch), owned vectors.
struct ColumnarChunk {
    std::vector<int32_t> tid, pos, l_qseq, mtid, mpos;
    std::vector<uint16_t> flag;
    std::vector<uint8_t> mapq, xs;
    std::vector<uint32_t> cigar_off{0};    // n+1
    std::vector<uint32_t> cigar;
    std::vector<uint64_t> seq_off{0};      // n+1
    std::vector<uint8_t> seq4;
    bool with_names = false;               // set before decode(): also fill name_code (the `--extra` metrics need it)
    std::vector<uint64_t> name_code;
    // ---- lean form (pj_batch.lean, include/portcullis_junc.h): set `lean` before decode().  tid / cigar_off / seq_off / seq4 stay
    // empty; n_cigar, seq2 (2 bits per base, spliced records only) and the non-ACGT exceptions are filled instead; mtid / mpos only
    // with keep_mate.  Every record of a decode task lies on one target: `runs` lists the (target, first record / CIGAR word / seq2
    // byte / exception) of each stretch, one entry after decode(), several after append(). ----
    bool lean = false, keep_mate = true;
    std::vector<uint16_t> n_cigar;
    std::vector<uint8_t> seq2;
    std::vector<uint64_t> seqx_pos;        // base index within seq2 (byte * 4 + q), ascending
    std::vector<uint8_t> seqx_code;        // BAM nibble
    struct Run { int32_t tid; int32_t pad; int64_t rec0, cig0, seq0, seqx0; };
    std::vector<Run> runs;
    int64_t n() const { return (int64_t)pos.size(); }
    void clear();
    void append(const ColumnarChunk& o);
};

class BamFile {
public:
    void open(const std::string& bam_path);                 // maps the file and parses the header
    bool load_bai(const std::string& bai_path);             // false if the file is missing
    bool load_csi(const std::string& csi_path);             // CSI (SAM spec, CSIv1): same planning data rebuilt from the per-bin loffset fields
    const BamHeader& header() const { return hdr_; }
    const std::vector<BamTargetIndex>& index() const { return idx_; }
    bool has_index() const { return !idx_.empty(); }
    // Split target `tid` into decode tasks of roughly `chunk_bytes` compressed bytes (needs the BAI).
    void plan_target(int32_t tid, uint64_t chunk_bytes, std::vector<DecodeTask>& out) const;
    // Without an index: one task per target found by a sequential scan is impossible, so a single
    // whole-file task (tid = -1 means "every target") is returned.
    DecodeTask whole_file_task() const;
    // Decode one task into `out` (appended).  Applies the reference's record visibility rule (Q13):
    // tid == target, pos < target_len, endpos > 0; iteration of a target stops at the first pos >= target_len.
    void decode(const DecodeTask& t, ColumnarChunk& out) const;
    // Sub-target cut: junctions are keyed by (refid, start, end) (lib/include/portcullis/intron.hpp:69-73), so
    // a target can be cut at any record R with  pos(R) >= pos_lo  and  pos(R) > end of every SPLICED record before it  — no read
    // with an N op spans the cut, hence no junction has reads on both sides.  Scans the records of `tid` in file order from
    // `start_voff` (the linear-index offset of the window holding pos_lo, so every record overlapping it is seen) and returns the
    // first such record: its virtual offset and position.  false: the target ends before a cut is found.
    bool find_gap_cut(int32_t tid, int32_t pos_lo, uint64_t start_voff, uint64_t* cut_voff, int32_t* cut_pos) const;
    const MappedFile& file() const { return file_; }
private:
    MappedFile file_;
    BamHeader hdr_;
    std::vector<BamTargetIndex> idx_;
};

// 64-bit code of BamAlignment::deriveName() (bam_alignment.cc:233-242): QNAME, plus "_R1"/"_R2"/"_R?" when the read is
// paired.  Stands in for std::hash<string> (junction.hpp:158): only equality matters.  FNV-1a then a 64-bit finaliser.
uint64_t name_code(const char* qname, size_t len, uint16_t flag);

// `junc --separate` (JunctionBuilder::separateBams, src/junction_builder.cc:152-226): every record of the BAM goes to one of
// three files — spliced (any N op), unspliced (mapped, no N), unmapped — laid out block for block like htslib's BamWriter
// output; the first two get a BAI (or CSI) index.  Inflate and deflate run on `threads` host threads.  bam_separate.cpp.
// Every record of a BAM in file order: fn(pointer to the block_size field, 4 + block_size).  Blocks are inflated in groups on
// `threads` threads; after_group() runs after each group (writers drain their finished blocks there).  bam_separate.cpp.
void scan_records(const BamFile& bam, int threads, const std::function<void(const uint8_t*, size_t)>& fn, const std::function<void()>& after_group);

struct SeparateCounts { uint64_t spliced = 0, unspliced = 0, unmapped = 0; };
void separate_bams(const BamFile& bam, const std::string& spliced_path, const std::string& unspliced_path, const std::string& unmapped_path,
                   bool use_csi, int threads, SeparateCounts& counts);

int inflate_selftest(int n_cases);

} // namespace pjio
