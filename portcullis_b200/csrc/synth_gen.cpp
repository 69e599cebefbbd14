// synth_gen.cpp (pjsynth) — deterministic synthetic workloads of the shapes named in BASELINE.json / SURVEY §8(d).
//
//   pjsynth --preset c2|c3|c4|c5 [--scale F] [--seed S] [--threads T] --out DIR
//
// Writes a complete *prep directory* (src/prepare.hpp:114-140 naming): portcullis.genome.fa(.fai) and a
// coordinate-sorted portcullis.sorted.alignments.bam(.bai), plus synth.json describing what was made.
// Every target is cut into position slices that are generated, sorted and BGZF-compressed independently (one
// RNG stream per slice, so the output does not depend on the thread count).
//
// Generator rules (SURVEY §8(d) "value distributions"): positions sorted; no H ops; no leading/trailing N; every
// anchor >= 1 bp; SEQ always present; junctions >= 100 bp from target ends; XS only as XS:A.
#include "bam_write.hpp"
#include <atomic>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace fs = std::filesystem;

struct Rng {                                  // splitmix64-seeded xoshiro256**
    uint64_t s[4];
    explicit Rng(uint64_t seed) { for (auto& x : s) { seed += 0x9E3779B97F4A7C15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; x = z ^ (z >> 31); } }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int64_t range(int64_t lo, int64_t hi) { return lo + (int64_t)(next() % (uint64_t)(hi - lo)); }   // [lo, hi)
    bool chance(double p) { return uni() < p; }
    double normal() { double u = uni(), v = uni(); if (u < 1e-300) u = 1e-300; return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v); }
};

struct Preset {
    std::string name;
    std::vector<int64_t> target_len; std::vector<std::string> target_name;
    int64_t n_alignments;
    bool paired; int read_lo, read_hi;
    int exons_lo, exons_hi, exon_lo, exon_hi, intron_lo, intron_hi;
    double sub_rate, indel_rate, clip_rate, retain_rate, alt_rate, unspliced_frac, lowq_frac, xs_frac, secondary_frac, n_run_frac, lower_frac;
    double depth_mu, depth_sigma;
    int n_hot; int64_t hot_lo, hot_hi;          // hot junctions (c4)
    uint64_t seed_genome, seed_reads;
};

static Preset preset(const std::string& n, double scale) {
    Preset p;
    p.name = n; p.paired = true; p.read_lo = 150; p.read_hi = 151; p.exons_lo = 2; p.exons_hi = 12; p.exon_lo = 50; p.exon_hi = 260;
    p.intron_lo = 70; p.intron_hi = 4000; p.sub_rate = 0.005; p.indel_rate = 0.0; p.clip_rate = 0.05; p.retain_rate = 0.02; p.alt_rate = 0.03;
    p.unspliced_frac = 0.12; p.lowq_frac = 0.10; p.xs_frac = 0.80; p.secondary_frac = 0.0; p.n_run_frac = 0.0; p.lower_frac = 0.0;
    p.depth_mu = 2.5; p.depth_sigma = 1.5; p.n_hot = 0; p.hot_lo = p.hot_hi = 0;
    auto scaled = [&](double v) { return (int64_t)std::max(1.0, std::floor(v * scale + 0.5)); };
    if (n == "c2" || n == "c4" || n == "c5") {
        for (int t = 0; t < 10; t++) { p.target_len.push_back(std::max<int64_t>(200000, scaled(10e6))); p.target_name.push_back("chr" + std::to_string(t + 1)); }
        p.seed_genome = 1001;
    }
    if (n == "c2") { p.n_alignments = scaled(10e6); p.seed_reads = 2002; }
    else if (n == "c3") {
        static const int64_t GRCH38[24] = {248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
                                           135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
                                           46709983, 50818468, 156040895, 57227415};
        for (int t = 0; t < 24; t++) { p.target_len.push_back(std::max<int64_t>(200000, scaled((double)GRCH38[t]))); p.target_name.push_back(t < 22 ? "chr" + std::to_string(t + 1) : (t == 22 ? "chrX" : "chrY")); }
        p.n_alignments = scaled(200e6); p.seed_genome = 3003; p.seed_reads = 3004; p.n_run_frac = 0.01; p.lower_frac = 0.40;
        p.intron_hi = 12000;
    }
    else if (n == "c4") {
        p.n_alignments = scaled(2e6); p.seed_reads = 4004; p.lowq_frac = 0.40; p.secondary_frac = 0.20;
        p.n_hot = 8; p.hot_lo = scaled(1e6); p.hot_hi = scaled(4e6);
    }
    else if (n == "c5") {
        p.n_alignments = scaled(1e6); p.seed_reads = 5005; p.paired = false; p.read_lo = 1000; p.read_hi = 10001;
        p.exons_lo = 4; p.exons_hi = 40; p.exon_lo = 80; p.exon_hi = 600; p.intron_lo = 70; p.intron_hi = 3000;
        p.sub_rate = 0.01; p.indel_rate = 0.30; p.clip_rate = 0.20; p.retain_rate = 0.10; p.alt_rate = 0.10; p.xs_frac = 0.70; p.unspliced_frac = 0.05;
    }
    else if (n != "c2") { fprintf(stderr, "unknown preset %s (c2|c3|c4|c5)\n", n.c_str()); exit(2); }
    return p;
}

struct Gene { int64_t lo; char strand; std::vector<std::pair<int64_t, int64_t>> exons; double weight; };
struct Slice { int tid; int64_t lo, hi; uint64_t idx; int64_t n_align; int64_t hot_reads; };

struct Rec { int64_t pos; std::vector<uint8_t> bytes; int64_t end; bool mapped; };

static const char* ACGT = "ACGT";
static inline uint8_t nt16(char c) { switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8; default: return 15; } }

// Build one BAM record.
static void make_record(Rec& r, uint32_t serial, int tid, int64_t pos, const std::vector<uint32_t>& cigar, const std::string& seq, int flag, int mapq,
                        int mtid, int64_t mpos, char xs) {
    int64_t rlen = 0;
    for (uint32_t c : cigar) { uint32_t op = c & 15; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; }
    r.pos = pos; r.end = pos + (rlen > 0 ? rlen : 1); r.mapped = !(flag & 4);
    char name[24]; int ln = snprintf(name, sizeof name, "r%x", serial) + 1;
    std::vector<uint8_t>& b = r.bytes; b.clear();
    b.reserve(36 + ln + 4 * cigar.size() + seq.size() * 3 / 2 + 8);
    bamw::put32(b, (uint32_t)tid); bamw::put32(b, (uint32_t)pos);
    b.push_back((uint8_t)ln); b.push_back((uint8_t)mapq); bamw::put16(b, (uint16_t)bamw::reg2bin(pos, r.end));
    bamw::put16(b, (uint16_t)cigar.size()); bamw::put16(b, (uint16_t)flag); bamw::put32(b, (uint32_t)seq.size());
    bamw::put32(b, (uint32_t)mtid); bamw::put32(b, (uint32_t)mpos); bamw::put32(b, 0);
    b.insert(b.end(), name, name + ln);
    for (uint32_t c : cigar) bamw::put32(b, c);
    for (size_t i = 0; i < seq.size(); i += 2) b.push_back((uint8_t)((nt16(seq[i]) << 4) | (i + 1 < seq.size() ? nt16(seq[i + 1]) : 0)));
    b.insert(b.end(), seq.size(), (uint8_t)0xff);
    if (xs) { b.push_back('X'); b.push_back('S'); b.push_back('A'); b.push_back((uint8_t)xs); }
}

// Alignment of transcript interval [t0, t0+len) of `exons` -> blocks, cigar (with optional indels), sequence.
static bool align_read(Rng& g, const Preset& P, const std::string& G, int64_t gbase, const std::vector<std::pair<int64_t, int64_t>>& ex, int64_t t0, int64_t len,
                       int64_t& pos, std::vector<uint32_t>& cigar, std::string& seq, int64_t& end_excl) {
    cigar.clear(); seq.clear();
    int64_t acc = 0, prev_end = -1; bool first = true;
    for (auto& e : ex) {
        const int64_t n = e.second - e.first, lo = std::max(t0, acc), hi = std::min(t0 + len, acc + n);
        if (lo < hi) {
            const int64_t s = e.first + lo - acc, en = e.first + hi - acc, bl = en - s;
            if (first) { pos = s; first = false; } else cigar.push_back((uint32_t)((s - prev_end) << 4) | 3u);
            prev_end = en;
            const char* src = G.data() + (s - gbase);
            if (bl >= 12 && P.indel_rate > 0 && g.chance(P.indel_rate)) {
                const int64_t off = g.range(1, 5), at = g.chance(0.5) ? off : bl - off - 3, l = g.range(1, 4);
                if (g.chance(0.5)) {
                    cigar.push_back((uint32_t)(at << 4)); cigar.push_back((uint32_t)(l << 4) | 1u); cigar.push_back((uint32_t)((bl - at) << 4));
                    seq.append(src, (size_t)at); for (int64_t k = 0; k < l; k++) seq.push_back(ACGT[g.next() & 3]); seq.append(src + at, (size_t)(bl - at));
                } else {
                    cigar.push_back((uint32_t)(at << 4)); cigar.push_back((uint32_t)(l << 4) | 2u); cigar.push_back((uint32_t)((bl - at - l) << 4));
                    seq.append(src, (size_t)at); seq.append(src + at + l, (size_t)(bl - at - l));
                }
            } else { cigar.push_back((uint32_t)(bl << 4)); seq.append(src, (size_t)bl); }
        }
        acc += n;
    }
    if (first) return false;
    end_excl = prev_end;
    for (auto& c : seq) { if (c >= 'a') c = (char)(c - 32); }
    if (P.sub_rate > 0) {
        // geometric skipping between substitutions
        const double lp = std::log(1.0 - P.sub_rate);
        for (double i = std::floor(std::log(1.0 - g.uni()) / lp); i < (double)seq.size(); i += 1.0 + std::floor(std::log(1.0 - g.uni()) / lp)) {
            char& c = seq[(size_t)i]; int k = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; c = ACGT[(k + 1 + (int)(g.next() % 3)) & 3];
        }
    }
    if (P.clip_rate > 0) {
        if (g.chance(P.clip_rate)) { int64_t k = g.range(1, 11); std::string c; for (int64_t i = 0; i < k; i++) c.push_back(ACGT[g.next() & 3]); seq = c + seq; cigar.insert(cigar.begin(), (uint32_t)(k << 4) | 4u); }
        if (g.chance(P.clip_rate)) { int64_t k = g.range(1, 11); for (int64_t i = 0; i < k; i++) seq.push_back(ACGT[g.next() & 3]); cigar.push_back((uint32_t)(k << 4) | 4u); }
    }
    return true;
}

struct SliceOut { bamw::FragmentWriter frag; int64_t n_rec = 0, n_spliced = 0, n_pairs = 0; std::string genome; };

static void generate_slice(const Preset& P, const Slice& S, SliceOut& out) {
    Rng gg(P.seed_genome * 0x100000001B3ull + S.idx), gr(P.seed_reads * 0x100000001B3ull + S.idx);
    const int64_t L = S.hi - S.lo;
    std::string& G = out.genome; G.resize((size_t)L);
    for (int64_t i = 0; i < L; i += 32) { uint64_t w = gg.next(); for (int64_t k = 0; k < 32 && i + k < L; k++) { G[(size_t)(i + k)] = ACGT[w & 3]; w >>= 2; } }
    // ---- gene models tiled over the slice ----
    std::vector<Gene> genes;
    const int64_t margin = 200;
    int64_t p = margin + gg.range(0, 500);
    while (true) {
        Gene ge; ge.strand = gg.chance(0.5) ? '+' : '-'; ge.lo = p;
        const int ne = (int)gg.range(P.exons_lo, P.exons_hi + 1);
        int64_t q = p; bool ok = true;
        for (int k = 0; k < ne; k++) {
            const int64_t el = gg.range(P.exon_lo, P.exon_hi);
            if (q + el + margin >= L) { ok = k >= 2; break; }
            ge.exons.emplace_back(S.lo + q, S.lo + q + el); q += el;
            if (k + 1 < ne) {
                // log-uniform intron length
                const int64_t il = (int64_t)std::floor(std::exp(std::log((double)P.intron_lo) + gg.uni() * (std::log((double)P.intron_hi) - std::log((double)P.intron_lo))));
                if (q + il + P.exon_lo + margin >= L) { ok = k >= 1; break; }
                q += il;
            }
        }
        if (ge.exons.size() >= 2) {
            // plant splice motifs: GT..AG 98%, GC..AG / AT..AC 1.5%, random 0.5%; reverse-complemented on '-'
            for (size_t k = 0; k + 1 < ge.exons.size(); k++) {
                const int64_t i0 = ge.exons[k].second - S.lo, i1 = ge.exons[k + 1].first - S.lo;   // intron [i0, i1)
                const double u = gg.uni(); const char *d = nullptr, *a = nullptr;
                if (u < 0.98) { d = "GT"; a = "AG"; } else if (u < 0.9875) { d = "GC"; a = "AG"; } else if (u < 0.995) { d = "AT"; a = "AC"; }
                if (d) {
                    if (ge.strand == '+') { G[i0] = d[0]; G[i0 + 1] = d[1]; G[i1 - 2] = a[0]; G[i1 - 1] = a[1]; }
                    else { auto comp = [](char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; };
                           G[i0] = comp(a[1]); G[i0 + 1] = comp(a[0]); G[i1 - 2] = comp(d[1]); G[i1 - 1] = comp(d[0]); }
                }
            }
            ge.weight = std::exp(P.depth_mu + P.depth_sigma * gg.normal());
            genes.push_back(std::move(ge));
        }
        if (!ok || q + 2000 >= L) break;
        p = q + gg.range(200, 3000);
        if (p + 3000 >= L) break;
    }
    // N runs / soft-masking (c3): decorate after planting so that some motifs and exons are hit, like real data
    if (P.n_run_frac > 0) { int64_t todo = (int64_t)(P.n_run_frac * L); while (todo > 0) { int64_t n = gg.range(50, 5000), s = gg.range(0, std::max<int64_t>(1, L - n)); for (int64_t k = 0; k < n && s + k < L; k++) G[(size_t)(s + k)] = 'N'; todo -= n; } }
    if (P.lower_frac > 0) { int64_t todo = (int64_t)(P.lower_frac * L); while (todo > 0) { int64_t n = gg.range(100, 3000), s = gg.range(0, std::max<int64_t>(1, L - n)); for (int64_t k = 0; k < n && s + k < L; k++) { char& c = G[(size_t)(s + k)]; if (c >= 'A' && c <= 'Z') c = (char)(c + 32); } todo -= n; } }
    if (genes.empty()) return;
    // ---- reads ----
    std::vector<Rec> recs; recs.reserve((size_t)(S.n_align + S.hot_reads + 16));
    double wsum = 0; for (auto& ge : genes) wsum += ge.weight;
    const int64_t n_bg = (int64_t)(P.unspliced_frac * S.n_align);
    const int64_t n_gene_al = S.n_align - n_bg;
    uint32_t serial = (uint32_t)(S.idx << 22);
    std::vector<uint32_t> cg, cg2; std::string sq, sq2;
    auto emit_from_gene = [&](const Gene& ge, bool force_first_junction, int64_t fixed_offset) {
        std::vector<std::pair<int64_t, int64_t>> ex = ge.exons;
        if (!force_first_junction) {
            if (ex.size() > 2 && gr.chance(P.retain_rate)) { size_t k = (size_t)gr.range(0, (int64_t)ex.size() - 1); ex[k].second = ex[k + 1].second; ex.erase(ex.begin() + (long)k + 1); }
            if (gr.chance(P.alt_rate)) {   // alternative donor/acceptor: move one internal exon boundary into the exon
                size_t k = (size_t)gr.range(0, (int64_t)ex.size() - 1); int64_t d = gr.range(3, 31);
                if (gr.chance(0.5)) { if (ex[k].second - ex[k].first > d + 10) ex[k].second -= d; } else { if (ex[k + 1].second - ex[k + 1].first > d + 10) ex[k + 1].first += d; }
            }
        }
        int64_t tl = 0; for (auto& e : ex) tl += e.second - e.first;
        const int64_t rl = std::min<int64_t>(gr.range(P.read_lo, P.read_hi), tl);
        const char xs_strand = ge.strand;
        auto flags_mapq = [&](int& flag, int& mapq) {
            mapq = gr.chance(P.lowq_frac) ? (int)(gr.next() % 3 == 0 ? 0 : gr.next() % 2 ? 1 : 3) : 60;
            if (P.secondary_frac > 0 && gr.chance(P.secondary_frac)) flag |= 0x100;
        };
        if (P.paired) {
            const int64_t frag = std::min<int64_t>(tl, std::max<int64_t>(rl, (int64_t)(300 + 50 * gr.normal())));
            int64_t t0 = force_first_junction ? fixed_offset : gr.range(0, tl - frag + 1);
            if (t0 < 0) t0 = 0;
            if (t0 + frag > tl) t0 = tl - frag;
            int64_t p1, p2, e1, e2;
            if (!align_read(gr, P, G, S.lo, ex, t0, rl, p1, cg, sq, e1)) return;
            if (!align_read(gr, P, G, S.lo, ex, t0 + frag - rl, rl, p2, cg2, sq2, e2)) return;
            const bool r1_fwd = gr.chance(0.5);
            int f1 = 0x1 | 0x2 | 0x20 | (r1_fwd ? 0x40 : 0x80), f2 = 0x1 | 0x2 | 0x10 | (r1_fwd ? 0x80 : 0x40), q1, q2;
            flags_mapq(f1, q1); flags_mapq(f2, q2);
            bool sp1 = false, sp2 = false; for (uint32_t c : cg) sp1 |= (c & 15) == 3; for (uint32_t c : cg2) sp2 |= (c & 15) == 3;
            Rec a, b;
            make_record(a, serial, S.tid, p1, cg, sq, f1, q1, S.tid, p2, (sp1 && gr.chance(P.xs_frac)) ? (gr.chance(0.97) ? xs_strand : (xs_strand == '+' ? '-' : '+')) : 0);
            make_record(b, serial, S.tid, p2, cg2, sq2, f2, q2, S.tid, p1, (sp2 && gr.chance(P.xs_frac)) ? (gr.chance(0.97) ? xs_strand : (xs_strand == '+' ? '-' : '+')) : 0);
            serial++;
            out.n_spliced += sp1 + sp2; for (uint32_t c : cg) out.n_pairs += (c & 15) == 3; for (uint32_t c : cg2) out.n_pairs += (c & 15) == 3;
            recs.push_back(std::move(a)); recs.push_back(std::move(b));
        } else {
            int64_t t0 = force_first_junction ? fixed_offset : gr.range(0, tl - rl + 1);
            if (t0 < 0) t0 = 0;
            if (t0 + rl > tl) t0 = tl - rl;
            int64_t p1, e1;
            if (!align_read(gr, P, G, S.lo, ex, t0, rl, p1, cg, sq, e1)) return;
            int f = gr.chance(0.5) ? 0x10 : 0, q; flags_mapq(f, q);
            bool sp = false; for (uint32_t c : cg) { sp |= (c & 15) == 3; out.n_pairs += (c & 15) == 3; }
            Rec a; make_record(a, serial++, S.tid, p1, cg, sq, f, q, -1, -1, (sp && gr.chance(P.xs_frac)) ? (gr.chance(0.97) ? xs_strand : (xs_strand == '+' ? '-' : '+')) : 0);
            out.n_spliced += sp; recs.push_back(std::move(a));
        }
    };
    {   // expression-weighted sampling of genes through the cumulative weights
        std::vector<double> cum(genes.size()); double c = 0; for (size_t k = 0; k < genes.size(); k++) { c += genes[k].weight / wsum; cum[k] = c; }
        const int64_t per = P.paired ? 2 : 1;
        for (int64_t i = 0; i < n_gene_al; i += per) { size_t k = (size_t)(std::lower_bound(cum.begin(), cum.end(), gr.uni()) - cum.begin()); if (k >= genes.size()) k = genes.size() - 1; emit_from_gene(genes[k], false, 0); }
    }
    if (S.hot_reads > 0) {   // one hot junction: the first intron of the middle gene; geometric start offsets (low entropy), many exact duplicates
        const Gene& ge = genes[genes.size() / 2];
        int64_t first_exon = ge.exons[0].second - ge.exons[0].first;
        const int64_t per = P.paired ? 2 : 1;
        for (int64_t i = 0; i < S.hot_reads; i += per) {
            int64_t back = 1 + (int64_t)std::floor(std::log(1.0 - gr.uni()) / std::log(0.9));   // geometric distance of the read start before the donor site
            if (gr.chance(0.30)) back = 20;
            back = std::min<int64_t>(std::min<int64_t>(back, first_exon), P.read_lo - 1);
            emit_from_gene(ge, true, first_exon - back);
        }
    }
    for (int64_t i = 0; i < n_bg; i++) {   // unspliced background + an occasional placed-unmapped record
        const int64_t rl = std::min<int64_t>(gr.range(P.read_lo, P.read_hi), 2000), s = gr.range(margin, L - rl - margin);
        cg.assign(1, (uint32_t)(rl << 4)); sq.assign(G.data() + s, (size_t)rl); for (auto& c : sq) if (c >= 'a') c = (char)(c - 32);
        Rec a; const bool unm = gr.chance(0.002);
        if (unm) { cg.clear(); make_record(a, serial++, S.tid, S.lo + s, cg, sq, 4, 0, -1, -1, 0); }
        else make_record(a, serial++, S.tid, S.lo + s, cg, sq, gr.chance(0.5) ? 16 : 0, 60, -1, -1, 0);
        recs.push_back(std::move(a));
    }
    std::stable_sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.pos < b.pos; });
    for (auto& r : recs) out.frag.add_record(r.bytes, r.pos, r.end, r.mapped);
    out.frag.flush();
    out.n_rec = (int64_t)recs.size();
}

int main(int argc, char** argv) {
    std::string preset_name = "c2", outdir; double scale = 1.0; uint64_t seed_add = 0; int threads = (int)std::thread::hardware_concurrency();
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--preset") preset_name = val(); else if (a == "--scale") scale = atof(val().c_str()); else if (a == "--seed") seed_add = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--threads") threads = atoi(val().c_str()); else if (a == "--out") outdir = val();
        else { fprintf(stderr, "usage: pjsynth --preset c2|c3|c4|c5 [--scale F] [--seed S] [--threads T] --out DIR\n"); return 2; }
    }
    if (outdir.empty()) { fprintf(stderr, "--out is required\n"); return 2; }
    if (threads < 1) threads = 1;
    Preset P = preset(preset_name, scale);
    P.seed_genome += seed_add; P.seed_reads += seed_add;
    fs::create_directories(outdir);
    // ---- slices ----
    const int64_t SLICE = 4000000;
    std::vector<Slice> slices; int64_t total_len = 0; for (auto l : P.target_len) total_len += l;
    for (size_t t = 0; t < P.target_len.size(); t++) {
        const int64_t L = P.target_len[t]; const int64_t ns = std::max<int64_t>(1, (L + SLICE / 2) / SLICE);
        for (int64_t k = 0; k < ns; k++) { Slice s; s.tid = (int)t; s.lo = L * k / ns; s.hi = L * (k + 1) / ns; s.idx = slices.size(); s.hot_reads = 0;
            s.n_align = (int64_t)((double)P.n_alignments * (double)(s.hi - s.lo) / (double)total_len); if (P.paired) s.n_align &= ~1ll; slices.push_back(s); }
    }
    if (P.n_hot > 0) { Rng h(P.seed_reads ^ 0xABCDEFull); for (int k = 0; k < P.n_hot; k++) { Slice& s = slices[(size_t)h.range(0, (int64_t)slices.size())]; s.hot_reads += h.range(P.hot_lo, P.hot_hi + 1); } }
    std::vector<SliceOut> outs(slices.size());
    std::atomic<size_t> next{0};
    auto work = [&]() { for (;;) { size_t k = next.fetch_add(1); if (k >= slices.size()) return; generate_slice(P, slices[k], outs[k]); } };
    {
        // slices are held in memory until written; bound the number in flight by writing in order as they complete
        std::vector<std::thread> th; for (int t = 0; t < threads; t++) th.emplace_back(work);
        for (auto& t : th) t.join();
    }
    // ---- FASTA + fai ----
    const std::string fa = outdir + "/portcullis.genome.fa", bam = outdir + "/portcullis.sorted.alignments.bam";
    {
        FILE* f = fopen(fa.c_str(), "wb"); FILE* fi = fopen((fa + ".fai").c_str(), "wb");
        if (!f || !fi) { fprintf(stderr, "cannot write %s\n", fa.c_str()); return 1; }
        uint64_t off = 0; std::string line;
        for (size_t t = 0; t < P.target_len.size(); t++) {
            std::string hdr = ">" + P.target_name[t] + "\n"; fwrite(hdr.data(), 1, hdr.size(), f); off += hdr.size();
            fprintf(fi, "%s\t%" PRId64 "\t%" PRIu64 "\t60\t61\n", P.target_name[t].c_str(), P.target_len[t], off);
            int col = 0; std::string buf; buf.reserve(1 << 22);
            for (auto& s : slices) if (s.tid == (int)t) {
                const std::string& G = outs[s.idx].genome;
                for (size_t i = 0; i < G.size();) { size_t k = std::min<size_t>(60 - col, G.size() - i); buf.append(G, i, k); i += k; col += (int)k; if (col == 60) { buf.push_back('\n'); col = 0; } if (buf.size() > (1u << 22) - 128) { fwrite(buf.data(), 1, buf.size(), f); off += buf.size(); buf.clear(); } }
            }
            if (col) buf.push_back('\n');
            fwrite(buf.data(), 1, buf.size(), f); off += buf.size();
        }
        fclose(f); fclose(fi);
    }
    // ---- BAM + BAI ----
    int64_t n_rec = 0, n_spliced = 0, n_pairs = 0;
    {
        std::vector<int32_t> lens; for (auto l : P.target_len) lens.push_back((int32_t)l);
        std::string text = "@HD\tVN:1.0\tSO:coordinate\n"; for (size_t t = 0; t < lens.size(); t++) text += "@SQ\tSN:" + P.target_name[t] + "\tLN:" + std::to_string(lens[t]) + "\n";
        text += "@PG\tID:pjsynth\tPN:pjsynth\tCL:preset=" + P.name + "\n";
        std::vector<uint8_t> head; bamw::bam_header(head, text, P.target_name, lens);
        FILE* f = fopen(bam.c_str(), "wb"); if (!f) { fprintf(stderr, "cannot write %s\n", bam.c_str()); return 1; }
        fwrite(head.data(), 1, head.size(), f);
        uint64_t off = head.size(); std::vector<uint64_t> base(slices.size());
        for (auto& s : slices) { base[s.idx] = off; auto& b = outs[s.idx].frag.bytes(); if (!b.empty()) fwrite(b.data(), 1, b.size(), f); off += b.size(); n_rec += outs[s.idx].n_rec; n_spliced += outs[s.idx].n_spliced; n_pairs += outs[s.idx].n_pairs; }
        std::vector<uint8_t> eof; bamw::bgzf_eof(eof); fwrite(eof.data(), 1, eof.size(), f); fclose(f);
        std::vector<uint8_t> bai = {'B', 'A', 'I', 1}; bamw::put32(bai, (uint32_t)lens.size());
        for (size_t t = 0; t < lens.size(); t++) {
            std::vector<const bamw::PartialIndex*> parts; std::vector<uint64_t> bs;
            for (auto& s : slices) if (s.tid == (int)t) { parts.push_back(&outs[s.idx].frag.index()); bs.push_back(base[s.idx]); }
            bamw::append_target_index(bai, parts, bs);
        }
        FILE* fb = fopen((bam + ".bai").c_str(), "wb"); fwrite(bai.data(), 1, bai.size(), fb); fclose(fb);
    }
    {
        FILE* f = fopen((outdir + "/synth.json").c_str(), "w");
        fprintf(f, "{\"preset\": \"%s\", \"scale\": %g, \"n_targets\": %zu, \"genome_bases\": %" PRId64 ", \"n_records\": %" PRId64 ", \"n_spliced\": %" PRId64 ", \"n_pairs\": %" PRId64 ", \"paired\": %s, \"read_len\": [%d, %d]}\n",
                P.name.c_str(), scale, P.target_len.size(), total_len, n_rec, n_spliced, n_pairs, P.paired ? "true" : "false", P.read_lo, P.read_hi - 1);
        fclose(f);
    }
    fprintf(stderr, "pjsynth: preset %s scale %g -> %" PRId64 " records (%" PRId64 " spliced, %" PRId64 " read-junction pairs) on %zu targets, %" PRId64 " bases\n",
            P.name.c_str(), scale, n_rec, n_spliced, n_pairs, P.target_len.size(), total_len);
    return 0;
}
