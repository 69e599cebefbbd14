// junc_host.cpp — see junc_host.hpp.
#include "junc_host.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>
#include <vector>

// ------------------------------------------------------------------------------------------------
// `--extra` metrics, host part.
// Q14: DepthParser::loadNextBatch fills `depths` for target c_k but leaves last.ref on the NEXT covered target c_k+1
// (depth_parser.cc:159-163), and JunctionSystem::calcCoverage asks getCurrentRefIndex() (depth_parser.hpp:94-96) which
// junctions to update (junction_system.cc:231-243).  So the junctions of c_k+1 are scored against the depth of c_k; only
// after the final batch (res == 0) does last.ref still name the batch's own target, which overwrites the last one.
// ------------------------------------------------------------------------------------------------
extern "C" void pj_extra_coverage_source(int32_t n_targets, const uint8_t* covered, int32_t* depth_src) {
    int32_t prev = -1, last = -1;
    for (int32_t t = 0; t < n_targets; t++) {
        depth_src[t] = -1;
        if (!covered[t]) continue;
        depth_src[t] = prev;            // -1 for the first covered target: its junctions are never visited ...
        prev = t; last = t;
    }
    if (last >= 0) depth_src[last] = last;   // ... unless it is also the last one
}

// Junction::calcMultipleMappingScore (junction.cc:914-921) and calcCoverage (:935-951): same operands, same order.
extern "C" void pj_extra_finalize(pj_junction_extra* x, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        pj_junction_extra& e = x[i];
        e.mm_score = (double)(size_t)e.mm_n / (double)e.mm_m;
        const double m10 = 1.0 / (double)(10 - 1), m11 = 1.0 / (double)10;          // multiplier = 1.0 / (b - a) for the 10- and 11-base windows
        const double donor = m10 * (double)e.cov_sum[0] - m11 * (double)e.cov_sum[1];
        const double acceptor = m11 * (double)e.cov_sum[2] - m10 * (double)e.cov_sum[3];
        e.coverage = donor + acceptor;
    }
}

// ------------------------------------------------------------------------------------------------
// A12/A13.  Reference: JunctionBuilder::findJunctions tail (src/junction_builder.cc:270-290),
// JunctionSystem::sort/index (lib/src/junction_system.cc:322-330), calcJunctionStats (:250-320),
// createJunctionGroup (:55-70).  Expressed here as passes over the sorted row array.
// ------------------------------------------------------------------------------------------------
extern "C" int pj_junctions_finalize(pj_junction* rows, int64_t n_rows, double mean_query_length) {
    if (n_rows < 0 || (n_rows > 0 && !rows)) return PJ_EINVAL;
    std::sort(rows, rows + n_rows, [](const pj_junction& a, const pj_junction& b) {
        if (a.tid != b.tid) return a.tid < b.tid;
        if (a.start != b.start) return a.start < b.start;
        return a.end < b.end;
    });
    const uint32_t NONE = 0xFFFFFFFFu;          // "-1" stored into a uint32 column
    for (int64_t i = 0; i < n_rows; i++) {
        pj_junction& j = rows[i];
        j.index = (uint32_t)i;
        j.rel2raw = (double)j.nb_rel_aln / (double)j.nb_raw_aln;
        j.mean_mismatches = (double)j.nb_mismatches / (double)j.nb_raw_aln;
        j.uniq_junc = j.primary_junc = j.pfp = 0;
        j.dist_2_up_junc = j.dist_2_down_junc = j.dist_nearest_junc = 0;
        j.mean_readlen = 0.0;
    }
    if (n_rows <= 1) return PJ_OK;               // calcJunctionStats is skipped (junction_builder.cc:285)

    // groups: maximal runs where each member shares start or end with its predecessor on the same target
    for (int64_t g0 = 0; g0 < n_rows;) {
        int64_t g1 = g0 + 1;
        while (g1 < n_rows && rows[g1].tid == rows[g1 - 1].tid &&
               (rows[g1].start == rows[g1 - 1].start || rows[g1].end == rows[g1 - 1].end)) g1++;
        int64_t best = g0; uint32_t best_reads = 0;
        for (int64_t k = g0; k < g1; k++) {
            rows[k].uniq_junc = (g1 - g0 == 1);
            if (rows[k].nb_raw_aln > best_reads) { best_reads = rows[k].nb_raw_aln; best = k; }   // strictly greater: first wins
        }
        rows[best].primary_junc = 1;
        g0 = g1;
    }

    // neighbour distances with the reference's boundary behaviour (quirk Q8)
    bool prev_pair_crossed = false;
    for (int64_t i = 0; i + 1 < n_rows; i++) {
        pj_junction& a = rows[i]; pj_junction& b = rows[i + 1];
        const bool first_pair = (i == 0), last_pair = (i == n_rows - 2);
        if (a.tid != b.tid) {
            a.dist_2_up_junc = NONE; b.dist_2_down_junc = NONE;
            if (first_pair || prev_pair_crossed) a.dist_2_down_junc = NONE;
            if (last_pair) b.dist_2_up_junc = NONE;
            prev_pair_crossed = true;
        } else {
            int32_t gap = b.start - a.end; if (gap < 0) gap = 0;
            a.dist_2_up_junc = (uint32_t)gap; b.dist_2_down_junc = (uint32_t)gap;
            if (first_pair) a.dist_2_down_junc = NONE;
            else if (last_pair) b.dist_2_up_junc = NONE;
            prev_pair_crossed = false;
        }
    }
    const double half_len = mean_query_length / 2.0;
    for (int64_t i = 0; i < n_rows; i++) {
        pj_junction& j = rows[i];
        const int32_t dn = (int32_t)j.dist_2_down_junc, up = (int32_t)j.dist_2_up_junc;
        j.dist_nearest_junc = (uint32_t)((dn == -1 || up == -1) ? std::max(dn, up) : std::min(dn, up));
        j.mean_readlen = (double)(uint32_t)mean_query_length;
        if (j.suspicious) {
            const double prob = 1.0 - std::pow((double)j.maxmmes / half_len, (double)j.nb_raw_aln);
            if (prob > 0.99) j.pfp = 1;
        }
    }
    return PJ_OK;
}

namespace pjhost {

// ------------------------------------------------------------------------------------------------
// Writers (A14).  Output must be byte-identical to the reference's ostream output, so numbers are
// formatted with the printf conversions iostreams use: default floatfield == %g with precision 6.
// ------------------------------------------------------------------------------------------------
namespace {

class Out {
public:
    explicit Out(const std::string& path) : f_(fopen(path.c_str(), "wb")), path_(path) {
        if (!f_) throw std::runtime_error("cannot open " + path + " for writing");
        buf_.reserve(1 << 22);
    }
    Out() : f_(nullptr) {}                                   // in-memory: rows formatted by a worker thread
    ~Out() { if (f_) { if (!buf_.empty()) (void)fwrite(buf_.data(), 1, buf_.size(), f_); fclose(f_); } }   // best effort only: writers call close()
    // Flushes and closes the file; a short write or a failing close (ENOSPC, EIO, quota) is an error, never a silently truncated table.
    void close() {
        if (!f_) return;
        flush();
        FILE* f = f_; f_ = nullptr;
        if (fclose(f) != 0) throw std::runtime_error("error closing " + path_ + " (disk full?)");
    }
    const std::string& str() const { return buf_; }
    void s(const char* p, size_t n) { buf_.append(p, n); if (f_ && buf_.size() > (1u << 22) - 4096) flush(); }
    void s(const char* p) { s(p, strlen(p)); }
    void s(const std::string& v) { s(v.data(), v.size()); }
    void c(char ch) { buf_.push_back(ch); }
    void u(uint64_t v) { char t[24]; int k = 24; do { t[--k] = (char)('0' + v % 10); v /= 10; } while (v); s(t + k, 24 - k); }
    void i(int64_t v) { if (v < 0) { c('-'); u((uint64_t)(-(v + 1)) + 1); } else u((uint64_t)v); }
    void g(double v, int prec = 6) { char t[48]; int k = snprintf(t, sizeof t, "%.*g", prec, v); s(t, (size_t)k); }
    void f3(double v) { char t[48]; int k = snprintf(t, sizeof t, "%.3f", v); s(t, (size_t)k); }
    void raw(const std::string& v) {                          // a finished piece: written directly (in-memory writers append)
        if (!f_) { buf_.append(v); return; }
        flush();
        if (!v.empty() && fwrite(v.data(), 1, v.size(), f_) != v.size()) throw std::runtime_error("short write to " + path_ + " (disk full?)");
    }
    void flush() {
        if (f_ && !buf_.empty()) {
            const size_t n = fwrite(buf_.data(), 1, buf_.size(), f_);
            if (n != buf_.size()) { buf_.clear(); throw std::runtime_error("short write to " + path_ + " (disk full?)"); }
            buf_.clear();
        }
    }
private:
    FILE* f_; std::string path_; std::string buf_;
};

// Formats rows [0, n) with `fmt(out, r)` on a few threads and appends the pieces to `o` in row order.
template <typename F>
void format_rows_parallel(Out& o, int64_t n, F fmt) {
    // up to half the host threads per file (the four files are written concurrently), at least 1024 rows per thread
    const int64_t hw = std::max<int64_t>(2, (int64_t)std::thread::hardware_concurrency());
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, std::max<int64_t>(4, hw / 2)), n / 1024));
    if (nt <= 1) { for (int64_t r = 0; r < n; r++) fmt(o, r); return; }
    std::vector<Out> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t]() { const int64_t a = n * t / nt, b = n * (t + 1) / nt; for (int64_t r = a; r < b; r++) fmt(parts[(size_t)t], r); });
    for (auto& x : th) x.join();
    for (auto& p : parts) o.raw(p.str());            // straight to the file: no second copy through the writer's own buffer
}

inline char strand_char(uint8_t s) { return s == PJ_STRAND_POS ? '+' : s == PJ_STRAND_NEG ? '-' : '?'; }
inline const char* strand_name(uint8_t s) { return s == PJ_STRAND_POS ? "POSITIVE" : s == PJ_STRAND_NEG ? "NEGATIVE" : "UNKNOWN"; }
inline const char* css_name(uint8_t c) { return c == 'C' ? "Canonical" : c == 'S' ? "Semi-canonical" : "No"; }

const char* const kMetricNames[] = {                      // column names, lib/src/junction.cc:50-92
    "canonical_ss", "score", "suspicious", "pfp", "nb_raw_aln", "nb_dist_aln", "nb_us_aln", "nb_ms_aln", "nb_um_aln",
    "nb_mm_aln", "nb_bpp_aln", "nb_ppp_aln", "nb_rel_aln", "rel2raw", "nb_r1_pos", "nb_r1_neg", "nb_r2_pos", "nb_r2_neg",
    "entropy", "mean_mismatches", "mean_readlen", "max_min_anc", "maxmmes", "intron_score", "hamming5p", "hamming3p",
    "coding", "pws", "splice_sig", "uniq_junc", "primary_junc", "nb_up_juncs", "nb_down_juncs", "dist_2_up_junc",
    "dist_2_down_junc", "dist_nearest_junc", "mm_score", "coverage", "up_aln", "down_aln", "nb_samples" };

const TargetInfo& target_of(const std::vector<TargetInfo>& t, int32_t tid) {
    if (tid < 0 || (size_t)tid >= t.size()) throw std::runtime_error("junction refers to an unknown target");
    return t[(size_t)tid];
}
} // namespace

std::string tab_header() {
    std::string h = "index\trefid\trefname\treflen\tstart\tend\tsize\tleft\tright\tread-strand\tss-strand\tconsensus-strand\tss1\tss2";
    for (const char* m : kMetricNames) { h += '\t'; h += m; }
    char t[16];
    for (int k = 1; k <= PJ_NB_JAD; k++) { snprintf(t, sizeof t, "\tJAD%02d", k); h += t; }
    return h;
}

void write_tab(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets, const pj_junction_extra* extra) {
    Out o(path);
    o.s(tab_header()); o.c('\n');
    for (int64_t r = 0; r < n; r++) target_of(targets, rows[r].tid);          // validate before going parallel
    format_rows_parallel(o, n, [&](Out& o, int64_t r) {
        const pj_junction& j = rows[r];
        const TargetInfo& t = targets[(size_t)j.tid];
        o.u(j.index); o.c('\t'); o.i(j.tid); o.c('\t'); o.s(t.name); o.c('\t'); o.i(t.length); o.c('\t');
        o.i(j.start); o.c('\t'); o.i(j.end); o.c('\t'); o.u((uint32_t)(j.end - j.start + 1)); o.c('\t');
        o.i(j.left); o.c('\t'); o.i(j.right); o.c('\t');
        o.c(strand_char(j.read_strand)); o.c('\t'); o.c(strand_char(j.ss_strand)); o.c('\t'); o.c(strand_char(j.consensus_strand)); o.c('\t');
        o.s(j.ss1, 2); o.c('\t'); o.s(j.ss2, 2); o.c('\t'); o.c((char)j.canonical_ss); o.c('\t');
        o.c('0'); o.c('\t');                                           // score
        o.u(j.suspicious); o.c('\t'); o.u(j.pfp); o.c('\t');
        o.u(j.nb_raw_aln); o.c('\t'); o.u(j.nb_dist_aln); o.c('\t');
        o.u((uint32_t)(j.nb_raw_aln - j.nb_ms_aln)); o.c('\t'); o.u(j.nb_ms_aln); o.c('\t');
        o.u(j.nb_um_aln); o.c('\t'); o.u((uint32_t)(j.nb_raw_aln - j.nb_um_aln)); o.c('\t');
        o.u(j.nb_bpp_aln); o.c('\t'); o.u(j.nb_ppp_aln); o.c('\t'); o.u(j.nb_rel_aln); o.c('\t');
        o.g(j.rel2raw); o.c('\t');
        o.u(j.nb_r1_pos); o.c('\t'); o.u(j.nb_r1_neg); o.c('\t'); o.u(j.nb_r2_pos); o.c('\t'); o.u(j.nb_r2_neg); o.c('\t');
        o.g(j.entropy); o.c('\t'); o.g(j.mean_mismatches); o.c('\t'); o.g(j.mean_readlen); o.c('\t');
        o.u(j.max_min_anc); o.c('\t'); o.u(j.maxmmes); o.c('\t');
        o.c('0'); o.c('\t');                                           // intron_score
        o.u(j.hamming5p); o.c('\t'); o.u(j.hamming3p); o.c('\t');
        o.s("0\t0\t0\t");                                              // coding, pws, splice_sig
        o.u(j.uniq_junc); o.c('\t'); o.u(j.primary_junc); o.c('\t');
        o.u(j.nb_up_juncs); o.c('\t'); o.u(j.nb_down_juncs); o.c('\t');
        o.u(j.dist_2_up_junc); o.c('\t'); o.u(j.dist_2_down_junc); o.c('\t'); o.u(j.dist_nearest_junc); o.c('\t');
        if (!extra) o.s("0\t0\t0\t0\t1");                              // mm_score, coverage, up_aln, down_aln, nb_samples
        else { const pj_junction_extra& x = extra[r]; o.g(x.mm_score); o.c('\t'); o.g(x.coverage); o.c('\t'); o.u(x.up_aln); o.c('\t'); o.u(x.down_aln); o.s("\t1"); }
        for (int k = 0; k < PJ_NB_JAD; k++) { o.c('\t'); o.u(j.jad[k]); }
        o.c('\n');
    });
    o.c('\n');        // `strm << js << endl` leaves one blank line at EOF (junction_system.cc:356)
    o.close();
}

void write_bed(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
               const std::string& source, const std::string& version) {
    Out o(path);
    o.s("track name=\"junctions\" description=\"Portcullis V"); o.s(version.empty() ? std::string("X.X.X") : version); o.s(" junctions\"\n");
    for (int64_t r = 0; r < n; r++) target_of(targets, rows[r].tid);
    format_rows_parallel(o, n, [&](Out& o, int64_t r) {
        const pj_junction& j = rows[r];
        const TargetInfo& t = targets[(size_t)j.tid];
        o.s(t.name); o.c('\t'); o.i(j.left); o.c('\t'); o.i((int64_t)j.right + 1); o.c('\t');
        o.s(source); o.c('_'); o.u(j.index); o.c('\t');
        o.f3((double)j.nb_raw_aln); o.c('\t');                       // the ?: makes the score a double under std::fixed, precision 3
        o.c(j.consensus_strand == PJ_STRAND_UNKNOWN ? '.' : strand_char(j.consensus_strand)); o.c('\t');
        o.i(j.start); o.c('\t'); o.i((int64_t)j.end + 1); o.c('\t');
        o.s("255,0,0\t2\t");
        o.i(j.start - j.left); o.c(','); o.i(j.right - j.end); o.c('\t');
        o.s("0,"); o.i(j.end - j.left + 1); o.c('\n');
    });
    o.close();
}

void write_exon_gff(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
                    const std::string& source) {
    Out o(path);
    for (int64_t r = 0; r < n; r++) target_of(targets, rows[r].tid);
    format_rows_parallel(o, n, [&](Out& o, int64_t r) {
        const pj_junction& j = rows[r];
        const TargetInfo& t = targets[(size_t)j.tid];
        const char strand = strand_char(j.consensus_strand);          // '?' when unknown
        auto lead = [&](const char* type, int64_t b, int64_t e) {
            o.s(t.name); o.c('\t'); o.s(source); o.c('\t'); o.s(type); o.c('\t'); o.i(b); o.c('\t'); o.i(e);
            o.s("\t0.0\t"); o.c(strand); o.s("\t.\t");
        };
        lead("match", (int64_t)j.left + 1, (int64_t)j.right + 1);
        o.s("ID=junc_"); o.u(j.index); o.s(";Name=junc_"); o.u(j.index);
        o.s(";Note=cov:"); o.u(j.nb_raw_aln); o.s("|rel:"); o.u(j.nb_rel_aln);
        o.s("|ent:"); o.g(j.entropy, 4);
        o.s("|maxmmes:"); o.u(j.maxmmes); o.s("|ham:"); o.u(std::min(j.hamming3p, j.hamming5p));
        o.s(";mult="); o.u(j.nb_raw_aln); o.s(";grp=junc_"); o.u(j.index); o.s(";src=E;");
        o.s("Strand: "); o.s(strand_name(j.consensus_strand));
        o.s(";Canonical?="); o.s(css_name(j.canonical_ss));
        o.s(";Score=0;NbAlignments="); o.u(j.nb_raw_aln); o.s(";NbDistinct="); o.u(j.nb_dist_aln);
        o.s(";NbReliable="); o.u(j.nb_rel_aln); o.s(";Entropy="); o.g(j.entropy, 9);   // stream left at precision 9
        o.s(";MaxMMES="); o.u(j.maxmmes); o.s(";HammingDistance5="); o.u(j.hamming5p); o.s(";HammingDistance3="); o.u(j.hamming3p);
        o.s(";UniqueJunction="); o.s(j.uniq_junc ? "true" : "false");
        o.s(";PrimaryJunction="); o.s(j.primary_junc ? "true" : "false"); o.s(";\n");
        lead("match_part", (int64_t)j.left + 1, (int64_t)j.start);
        o.s("ID=junc_"); o.u(j.index); o.s("_left;Parent=junc_"); o.u(j.index); o.c('\n');
        lead("match_part", (int64_t)j.end + 2, (int64_t)j.right + 1);
        o.s("ID=junc_"); o.u(j.index); o.s("_right;Parent=junc_"); o.u(j.index); o.c('\n');
    });
    o.close();
}

void write_intron_gff(const std::string& path, const pj_junction* rows, int64_t n, const std::vector<TargetInfo>& targets,
                      const std::string& source) {
    Out o(path);
    for (int64_t r = 0; r < n; r++) target_of(targets, rows[r].tid);
    format_rows_parallel(o, n, [&](Out& o, int64_t r) {
        const pj_junction& j = rows[r];
        const TargetInfo& t = targets[(size_t)j.tid];
        o.s(t.name); o.c('\t'); o.s(source); o.s("\tintron\t"); o.i((int64_t)j.start + 1); o.c('\t'); o.i((int64_t)j.end + 1); o.c('\t');
        o.u(j.nb_raw_aln); o.c('\t'); o.c(strand_char(j.consensus_strand)); o.s("\t.\tmult="); o.u(j.nb_raw_aln);
        o.s(";grp=junc_"); o.u(j.index); o.s(";src=E\n");
    });
    o.close();
}

} // namespace pjhost
