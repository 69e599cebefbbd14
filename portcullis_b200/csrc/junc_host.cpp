// junc_host.cpp — see junc_host.hpp.
#include "junc_host.hpp"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <chrono>
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
namespace {
// f(a, b) over [0, n) split into contiguous ranges, one host thread each (the row array is 256 B per junction: the passes below are
// memory-bound sweeps over hundreds of MB on a human-scale run)
template <typename F>
void parallel_ranges(int64_t n, int64_t min_per_thread, F f) {
    const int64_t hw = std::max<int64_t>(1, (int64_t)std::thread::hardware_concurrency());
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(hw, 32), n / std::max<int64_t>(min_per_thread, 1)));
    if (nt <= 1) { f((int64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) th.emplace_back([&, t]() { f(n * t / nt, n * (t + 1) / nt); });
    f((int64_t)0, n / nt);
    for (auto& x : th) x.join();
}
inline bool row_less(const pj_junction& a, const pj_junction& b) {
    if (a.tid != b.tid) return a.tid < b.tid;
    if (a.start != b.start) return a.start < b.start;
    return a.end < b.end;
}
} // namespace

extern "C" int pj_junctions_finalize(pj_junction* rows, int64_t n_rows, double mean_query_length) {
    if (n_rows < 0 || (n_rows > 0 && !rows)) return PJ_EINVAL;
    {   // the driver hands the rows over in order already (shards are coordinate ranges): check before paying for a sort of 256-byte rows
        std::atomic<bool> sorted{true};
        parallel_ranges(n_rows, 1 << 16, [&](int64_t a, int64_t b) {
            for (int64_t i = std::max<int64_t>(a, 1); i < b; i++) if (row_less(rows[i], rows[i - 1])) { sorted.store(false, std::memory_order_relaxed); return; }
        });
        if (!sorted.load()) std::sort(rows, rows + n_rows, row_less);
    }
    const uint32_t NONE = 0xFFFFFFFFu;          // "-1" stored into a uint32 column
    const bool stats = n_rows > 1;               // calcJunctionStats is skipped for a single junction (junction_builder.cc:285)
    const double half_len = mean_query_length / 2.0;
    // A member of a group shares start or end with its predecessor on the same target.
    auto continues_group = [&](int64_t k) { return k > 0 && rows[k].tid == rows[k - 1].tid && (rows[k].start == rows[k - 1].start || rows[k].end == rows[k - 1].end); };
    parallel_ranges(n_rows, 1 << 15, [&](int64_t a, int64_t b) {
        // a thread owns the groups that START inside its range (the last one may run past b)
        while (a < b && continues_group(a)) a++;
        for (int64_t g0 = a; g0 < b;) {
            int64_t g1 = g0 + 1;
            while (g1 < n_rows && continues_group(g1)) g1++;
            int64_t best = g0; uint32_t best_reads = 0;
            for (int64_t k = g0; k < g1; k++) {
                pj_junction& j = rows[k];
                j.index = (uint32_t)k;
                j.rel2raw = (double)j.nb_rel_aln / (double)j.nb_raw_aln;
                j.mean_mismatches = (double)j.nb_mismatches / (double)j.nb_raw_aln;
                j.uniq_junc = j.primary_junc = j.pfp = 0;
                j.dist_2_up_junc = j.dist_2_down_junc = j.dist_nearest_junc = 0;
                j.mean_readlen = 0.0;
                if (!stats) continue;
                j.uniq_junc = (g1 - g0 == 1);
                if (j.nb_raw_aln > best_reads) { best_reads = j.nb_raw_aln; best = k; }   // strictly greater: first wins
                // Neighbour distances with the reference's boundary behaviour (quirk Q8), per row: the reference walks the pairs (i, i + 1)
                // in order and lets later pairs overwrite earlier ones; what is left on row k is
                //   down: -1 on the first row and behind a target change, else the gap to the previous junction;
                //   up:   -1 in front of a target change, the gap to the next junction otherwise; on the LAST row -1, except that with
                //         exactly two junctions on one target it keeps its initial 0.
                auto gap = [&](int64_t x, int64_t y) { const int32_t g = rows[y].start - rows[x].end; return (uint32_t)(g < 0 ? 0 : g); };
                j.dist_2_down_junc = (k == 0 || rows[k - 1].tid != j.tid) ? NONE : gap(k - 1, k);
                if (k + 1 < n_rows) j.dist_2_up_junc = (rows[k + 1].tid != j.tid) ? NONE : gap(k, k + 1);
                else j.dist_2_up_junc = (rows[k - 1].tid != j.tid) ? NONE : (n_rows == 2 ? 0u : NONE);
                const int32_t dn = (int32_t)j.dist_2_down_junc, up = (int32_t)j.dist_2_up_junc;
                j.dist_nearest_junc = (uint32_t)((dn == -1 || up == -1) ? std::max(dn, up) : std::min(dn, up));
                j.mean_readlen = (double)(uint32_t)mean_query_length;
                if (j.suspicious) {
                    const double prob = 1.0 - std::pow((double)j.maxmmes / half_len, (double)j.nb_raw_aln);
                    if (prob > 0.99) j.pfp = 1;
                }
            }
            if (stats) rows[best].primary_junc = 1;
            g0 = g1;
        }
    });
    return PJ_OK;
}

namespace pjhost {

// ------------------------------------------------------------------------------------------------
// Writers (A14).  Output must be byte-identical to the reference's ostream output, so numbers are
// formatted with the printf conversions iostreams use: default floatfield == %g with precision 6.
// ------------------------------------------------------------------------------------------------
namespace {

// ---- number formatting ----
// The writers print ~80 numbers per junction; snprintf and std::string::append were 85 % of the writers' time.  Integers are written
// digit pairs at a time; "%.{P}g" / "%.3f" have a fast path that is only taken when it provably prints what printf prints, anything
// else (ties and near-ties of the decimal rounding, huge / tiny magnitudes, NaN, infinities, -0) goes through snprintf.
const char DIGIT_PAIRS[] = "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
                           "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
inline int n_digits(uint64_t v) {
    if (v < 100000) return v < 10 ? 1 : v < 100 ? 2 : v < 1000 ? 3 : v < 10000 ? 4 : 5;
    if (v < 1000000000) return v < 1000000 ? 6 : v < 10000000 ? 7 : v < 100000000 ? 8 : 9;
    int n = 9; v /= 1000000000; while (v) { n++; v /= 10; } return n;
}
// writes v at dst (no terminator), returns the number of characters
inline int fmt_u(char* dst, uint64_t v) {
    const int n = n_digits(v); char* p = dst + n;
    while (v >= 100) { const unsigned r = (unsigned)(v % 100); v /= 100; p -= 2; memcpy(p, DIGIT_PAIRS + 2 * r, 2); }
    if (v >= 10) { p -= 2; memcpy(p, DIGIT_PAIRS + 2 * v, 2); } else *--p = (char)('0' + v);
    return n;
}
const double P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};   // exact in binary64
// "%.{prec}g" of v into dst (at least 40 bytes); returns the length, or 0 when the value needs printf's exact arithmetic
inline int fmt_g_fast(char* dst, double v, int prec) {
    if (prec < 1 || prec > 9 || !(v == v)) return 0;
    char* p = dst;
    double a = v;
    if (v < 0) { a = -v; *p++ = '-'; } else if (v == 0) { if (std::signbit(v)) return 0; *p = '0'; return 1; }
    if (a > 1e15 || a < 1e-7) return 0;
    const double lo = P10[prec - 1], hi = P10[prec];
    int e10 = (int)std::floor(std::log10(a));                  // an estimate: corrected below against the scaled value itself
    uint64_t N = 0;
    for (int tries = 0;; tries++) {
        if (tries > 2) return 0;
        const int k = prec - 1 - e10;
        if (k > 22 || k < -22) return 0;
        const double scaled = k >= 0 ? a * P10[k] : a / P10[-k];   // one correctly rounded operation: relative error <= 2^-53
        if (scaled < lo) { e10--; continue; }
        if (scaled >= hi) { e10++; continue; }
        const double fl = std::floor(scaled), frac = scaled - fl;
        if (std::fabs(frac - 0.5) < 1e-5) return 0;            // (near-)tie: printf decides on the exact binary value
        N = (uint64_t)fl + (frac > 0.5 ? 1u : 0u);
        if (N >= (uint64_t)hi) { N /= 10; e10++; }              // 999999.7 -> 1.00000e6
        break;
    }
    char dg[24]; int nd = fmt_u(dg, N);                         // exactly `prec` digits
    while (nd > 1 && dg[nd - 1] == '0') nd--;                   // %g drops trailing zeros
    const int X = e10;
    if (X < -4 || X >= prec) {                                  // d[.ddd]e+XX
        *p++ = dg[0];
        if (nd > 1) { *p++ = '.'; memcpy(p, dg + 1, (size_t)nd - 1); p += nd - 1; }
        *p++ = 'e'; int ax = X;
        if (ax < 0) { *p++ = '-'; ax = -ax; } else *p++ = '+';
        if (ax < 10) *p++ = '0';
        p += fmt_u(p, (uint64_t)ax);
    } else if (X >= 0) {
        const int ip = X + 1;                                   // digits before the point
        if (nd <= ip) { memcpy(p, dg, (size_t)nd); p += nd; for (int z = nd; z < ip; z++) *p++ = '0'; }
        else { memcpy(p, dg, (size_t)ip); p += ip; *p++ = '.'; memcpy(p, dg + ip, (size_t)(nd - ip)); p += nd - ip; }
    } else {
        *p++ = '0'; *p++ = '.';
        for (int z = 0; z < -X - 1; z++) *p++ = '0';
        memcpy(p, dg, (size_t)nd); p += nd;
    }
    return (int)(p - dst);
}
inline int fmt_g(char* dst, size_t cap, double v, int prec) {
    // whole numbers that fit the precision print as integers (mean_readlen, rel2raw of 0 or 1, ...)
    if (v >= 0 && v < P10[prec < 1 || prec > 15 ? 1 : prec] && v == (double)(uint64_t)v && !(v == 0 && std::signbit(v))) return fmt_u(dst, (uint64_t)v);
    const int k = fmt_g_fast(dst, v, prec);
    return k ? k : snprintf(dst, cap, "%.*g", prec, v);
}

class Out {
public:
    explicit Out(const std::string& path) : f_(fopen(path.c_str(), "wb")), path_(path) {
        if (!f_) throw std::runtime_error("cannot open " + path + " for writing");
        grow((size_t)1 << 22);
    }
    Out() : f_(nullptr) {}                                   // in-memory: rows formatted by a worker thread
    Out(const Out&) = delete; Out& operator=(const Out&) = delete;
    Out(Out&& o) noexcept : f_(o.f_), path_(std::move(o.path_)), b_(o.b_), n_(o.n_), cap_(o.cap_) { o.f_ = nullptr; o.b_ = nullptr; o.n_ = o.cap_ = 0; }
    ~Out() { if (f_) { if (n_) (void)fwrite(b_, 1, n_, f_); fclose(f_); } free(b_); }   // best effort only: writers call close()
    // Flushes and closes the file; a short write or a failing close (ENOSPC, EIO, quota) is an error, never a silently truncated table.
    void close() {
        if (!f_) return;
        flush();
        FILE* f = f_; f_ = nullptr;
        if (fclose(f) != 0) throw std::runtime_error("error closing " + path_ + " (disk full?)");
    }
    const char* data() const { return b_; }
    size_t size() const { return n_; }
    void s(const char* p, size_t n) { room(n); memcpy(b_ + n_, p, n); n_ += n; }
    void s(const char* p) { s(p, strlen(p)); }
    void s(const std::string& v) { s(v.data(), v.size()); }
    void c(char ch) { room(1); b_[n_++] = ch; }
    void u(uint64_t v) { room(24); n_ += (size_t)fmt_u(b_ + n_, v); }
    void i(int64_t v) { if (v < 0) { c('-'); u((uint64_t)(-(v + 1)) + 1); } else u((uint64_t)v); }
    void g(double v, int prec = 6) { room(48); n_ += (size_t)fmt_g(b_ + n_, 48, v, prec); }
    void f3(double v) {
        room(48);
        if (v >= 0 && v < 1e15 && v == (double)(uint64_t)v) { n_ += (size_t)fmt_u(b_ + n_, (uint64_t)v); memcpy(b_ + n_, ".000", 4); n_ += 4; }
        else n_ += (size_t)snprintf(b_ + n_, 48, "%.3f", v);
    }
    void raw(const char* p, size_t n) {                       // a finished piece: written directly (in-memory writers append)
        if (!f_) { s(p, n); return; }
        flush();
        if (n && fwrite(p, 1, n, f_) != n) throw std::runtime_error("short write to " + path_ + " (disk full?)");
    }
    void flush() {
        if (f_ && n_) {
            const size_t w = fwrite(b_, 1, n_, f_);
            const bool ok = w == n_; n_ = 0;
            if (!ok) throw std::runtime_error("short write to " + path_ + " (disk full?)");
        }
    }
private:
    void grow(size_t need) {
        const size_t nc = std::max<size_t>(std::max<size_t>(cap_ * 2, n_ + need + 4096), (size_t)1 << 16);
        char* nb = (char*)realloc(b_, nc);
        if (!nb) throw std::bad_alloc();
        b_ = nb; cap_ = nc;
    }
    void room(size_t k) {
        if (n_ + k <= cap_) return;
        if (f_) flush();                                      // a file-backed writer empties its 4 MB buffer instead of growing it
        if (n_ + k > cap_) grow(k);
    }
    FILE* f_; std::string path_; char* b_ = nullptr; size_t n_ = 0, cap_ = 0;
};

// Formats rows [0, n) with `fmt(out, r)` on a few threads and appends the pieces to `o` in row order.
template <typename F>
void format_rows_parallel(Out& o, int64_t n, F fmt) {
    // up to half the host threads per file (the four files are written concurrently), at least 1024 rows per thread
    const int64_t hw = std::max<int64_t>(2, (int64_t)std::thread::hardware_concurrency());
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, std::max<int64_t>(4, hw / 2)), n / 1024));
    if (nt <= 1) { for (int64_t r = 0; r < n; r++) fmt(o, r); return; }
    static const bool trace = getenv("PJ_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<Out> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&, t]() { const int64_t a = n * t / nt, b = n * (t + 1) / nt; for (int64_t r = a; r < b; r++) fmt(parts[(size_t)t], r); });
    for (auto& x : th) x.join();
    const auto t1 = std::chrono::steady_clock::now();
    for (auto& p : parts) o.raw(p.data(), p.size());  // straight to the file: no second copy through the writer's own buffer
    if (trace) fprintf(stderr, "[pj writer] %lld rows on %d threads: format %.3f s, write %.3f s\n", (long long)n, nt,
                       std::chrono::duration<double>(t1 - t0).count(), std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
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


// Self-test of the writers' number formatting against printf (exported as pjh_format_selftest for the CPU test-suite): returns the
// number of values whose fast rendering differs from snprintf's.
int format_selftest(int n_cases) {
    uint64_t x = 0x9e3779b97f4a7c15ull; int bad = 0;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    auto check = [&](double v) {
        for (int prec : {4, 6, 9}) {
            char a[64], b[64];
            const int ka = fmt_g(a, sizeof a, v, prec); const int kb = snprintf(b, sizeof b, "%.*g", prec, v);
            if (ka != kb || memcmp(a, b, (size_t)ka) != 0) bad++;
        }
    };
    const double specials[] = {0.0, -0.0, 1.0, -1.0, 0.5, 0.25, 1e-5, 1e-4, 9.99999e-5, 0.0001234565, 123456.5, 1234565.0, 999999.5, 999999.4, 999999.6, 0.1, 0.2, 0.3,
                               1e6, 1e5, 99999.95, 1e15, 1e16, 1e-7, 1e-8, 2.5, 3.5, 100000.5, 1234.5675, 0.000123456789, 150.0, 1.0 / 3.0, 2.0 / 3.0, 1e22, 1e-300, 1e300};
    for (double v : specials) { check(v); check(-v); }
    for (int i = 0; i < n_cases; i++) {
        const uint64_t r = next();
        const double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
        check(u); check(u * 10); check(u * 1000); check(u * 1e6); check(u * 1e9); check(u * 1e-3); check(u * 1e-6);
        check((double)(r % 2000000) / 1000.0); check((double)(r % 100000) / 7.0); check((double)(r % 1000) / (double)(1 + (r >> 20) % 1000));
        check((double)(r % 20000001) / 2.0);                                       // exact ties at 7 digits
        check(-u * 100);
    }
    return bad;
}

} // namespace pjhost
