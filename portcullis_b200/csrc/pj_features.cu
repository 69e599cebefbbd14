// pj_features.cu — `filt` feature extraction on the GPU: the per-junction feature vector of
// portcullis::ml::ModelFeatures::setRow (lib/src/model_features.cc:168-209) from the junction table and the packed genome that are
// already resident, plus the training of the six k-mer Markov models and two positional models it scores against
// (trainCodingPotentialModel :70-110, trainSplicingModels :112-166, lib/src/markov_model.cc).
//
// What runs where
//   * k_feat_count_coding / k_feat_count_splicing: one warp per training junction; lanes stride over the bases of the exon /
//     intron / splice-site windows, form the 5-base context of their position from the 2-bit genome plane and count
//     (context, next base) with integer atomics — deterministic.  Alphabet {A,C,G,T,N}: SeqUtils::makeClean maps every other
//     byte to N, and contexts containing N are ordinary keys of the reference's hash maps.
//   * host: probability = count / sum over the next bases of a context, in fp64, exactly the division the reference does.
//   * k_feat_score: one thread per junction; every window is scored with the reference's left-to-right product of
//     probabilities (KmerMarkovModel::getScore / PosMarkovModel::getScore), so the products are bit-identical and only the
//     final log() may differ in the last ulp (the tests allow 1e-6 relative).
//
// Window semantics follow faidx_fetch_seq (deps/htslib-1.3/faidx.c:439-476: end < beg -> beg = end, both clamped to
// [0, len-1]) and the negative-strand reverse complement of SeqUtils::reverseComplement + makeClean: A<->T, C<->G, U->A,
// everything else N.  One deliberate deviation: for a LOWER-CASE base on the negative strand the reference indexes
// REVCOMP_LOOKUP[c - 65] past the end of its 26-entry table (undefined behaviour, seq_utils.hpp:33-40,115); here the base is
// complemented like its upper-case form.
#include "pj_ctx.hpp"
#include <cmath>
#include <cstring>
#include <vector>

using namespace pjk;
using namespace pjapi;

namespace {

constexpr int KCTX = 3125;                    // 5^5 contexts of order 5
constexpr int KTAB = KCTX * 5;                // (context, next base)
constexpr int PMAX = 32;                      // positions of a positional model (splice-site windows have 23 / 24 bases)
enum { M_EXON = 0, M_INTRON, M_DONOR_T, M_DONOR_F, M_ACC_T, M_ACC_F, N_KMER };

// symbol of SeqUtils::makeClean: A C G T -> 0..3, anything else -> 4 ('N')
__device__ __forceinline__ int clean_sym(uint8_t ch) { return ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4; }
// same after SeqUtils::reverseComplement: A<->T, C<->G, U->A, the rest of the table (N, IUPAC, holes) is not A/C/G/T -> N
__device__ __forceinline__ int clean_sym_rc(uint8_t ch) { return ch == 'A' ? 3 : ch == 'C' ? 2 : ch == 'G' ? 1 : (ch == 'T' || ch == 'U') ? 0 : 4; }

// A window of the genome as the reference's string: fetchBases(name, beg, end) [+ reverseComplement] + makeClean.
struct Window {
    uint64_t gbase; int32_t beg, len; bool neg;
    __device__ int sym(const Genome& G, int32_t i) const {          // i-th character of the (possibly reverse-complemented) string
        const int32_t p = neg ? beg + len - 1 - i : beg + i;
        const uint8_t ch = genome_char(G, gbase + (uint64_t)(uint32_t)p);
        return neg ? clean_sym_rc(ch) : clean_sym(ch);
    }
};
__device__ __forceinline__ Window make_window(const Genome& G, int32_t tid, int32_t beg, int32_t end, bool neg) {
    Window w; w.gbase = G.goff[tid]; w.neg = neg;
    const int64_t glen = G.glen[tid];
    if (glen <= 0) { w.beg = 0; w.len = 0; return w; }               // sequence not loaded: faidx returns no bases
    if (end < beg) beg = end;                                        // faidx.c:455-459
    if (beg < 0) beg = 0; else if (glen <= beg) beg = (int32_t)glen - 1;
    if (end < 0) end = 0; else if (glen <= end) end = (int32_t)glen - 1;
    w.beg = beg; w.len = end - beg + 1;
    return w;
}

// KmerMarkovModel::train (markov_model.cc:31-55) over one window: a warp counts the (context, next) pairs of positions 5 .. len-1.
__device__ void count_kmer_window(const Genome& G, const Window& w, unsigned long long* __restrict__ tab, int lane) {
    if (w.len <= 6) return;                                          // `if (s.size() > order + 1)`
    for (int32_t i = 5 + lane; i < w.len; i += 32) {
        int ctx = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) ctx = ctx * 5 + w.sym(G, i - 5 + k);
        atomicAdd(tab + ctx * 5 + w.sym(G, i), 1ull);
    }
}
// PosMarkovModel::train (markov_model.cc:84-104), order 1: position i >= 1 counts its own base
__device__ void count_pos_window(const Genome& G, const Window& w, unsigned long long* __restrict__ tab, int lane) {
    for (int32_t i = 1 + lane; i < w.len && i < PMAX; i += 32) atomicAdd(tab + i * 5 + w.sym(G, i), 1ull);
}

struct JuncKey { int32_t tid, start, end; uint8_t neg, pad[3]; };

__global__ void __launch_bounds__(256) k_feat_count_coding(int64_t n, const JuncKey* __restrict__ j, Genome G,
                                                            unsigned long long* __restrict__ exon, unsigned long long* __restrict__ intron) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const JuncKey k = j[w];
    const bool neg = k.neg != 0;
    count_kmer_window(G, make_window(G, k.tid, k.start - 202, k.start - 2, neg), exon, lane);      // model_features.cc:74
    count_kmer_window(G, make_window(G, k.tid, k.start, k.end, neg), intron, lane);                 // :92
    count_kmer_window(G, make_window(G, k.tid, k.end + 1, k.end + 201, neg), exon, lane);           // :97
}
__global__ void __launch_bounds__(256) k_feat_count_splicing(int64_t n, const JuncKey* __restrict__ j, Genome G,
                                                              unsigned long long* __restrict__ donor_k, unsigned long long* __restrict__ acc_k,
                                                              unsigned long long* __restrict__ donor_p, unsigned long long* __restrict__ acc_p) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const JuncKey k = j[w];
    const bool neg = k.neg != 0;
    const Window left = make_window(G, k.tid, k.start - 3, k.start + 20, neg), right = make_window(G, k.tid, k.end - 20, k.end + 2, neg);
    const Window& donor = neg ? right : left; const Window& acc = neg ? left : right;               // :124-131
    count_kmer_window(G, donor, donor_k, lane); count_kmer_window(G, acc, acc_k, lane);
    if (donor_p) { count_pos_window(G, donor, donor_p, lane); count_pos_window(G, acc, acc_p, lane); }
}

// KmerMarkovModel::getScore (markov_model.cc:58-82); *touched: the reference's operator[] inserts every key it looks up
__device__ double kmer_score(const Genome& G, const Window& w, const double* __restrict__ prob) {
    double score = 1.0; uint32_t no_count = 0;
    int ctx = 0;
    for (int32_t i = 0; i < w.len; i++) {
        const int s = w.sym(G, i);
        if (i >= 5) {
            const double m = prob[ctx * 5 + s];
            if (m != 0.0) score *= m; else no_count++;
        }
        ctx = (ctx % 625) * 5 + s;                                   // keep the last five symbols
    }
    if (score == 0.0) return -100.0;
    if (no_count > 2) score /= ((double)no_count * 0.5);
    return log(score);
}
// PosMarkovModel::getScore (markov_model.cc:107-117)
__device__ double pos_score(const Genome& G, const Window& w, const double* __restrict__ prob) {
    double score = 1.0;
    for (int32_t i = 1; i < w.len; i++) score *= (i < PMAX ? prob[i * 5 + w.sym(G, i)] : 0.0);
    if (score == 0.0) return -300.0;
    return log(score);
}

struct FeatIn {                       // what setRow reads of a junction besides its coordinates
    uint32_t nb_raw, nb_ms, nb_dist, nb_rel, max_min_anc, maxmmes, hamming5p, hamming3p;
    double entropy, rel2raw, mean_mismatches, mean_readlen;
    uint32_t jad[PJ_NB_JAD];
};

__global__ void __launch_bounds__(128) k_feat_score(int64_t n, const JuncKey* __restrict__ jk, const FeatIn* __restrict__ fin, Genome G,
                                                     const double* __restrict__ kmer /* [N_KMER][KTAB] */, const double* __restrict__ posm /* [2][PMAX*5] */,
                                                     int coding_empty, uint32_t l95, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const JuncKey k = jk[i]; const FeatIn f = fin[i];
    const bool neg = k.neg != 0;
    double* o = out + i * PJ_NB_FEATURES;
    // Junction::calcSplicingScores (junction.cc:1360-1382)
    const Window left = make_window(G, k.tid, k.start - 3, k.start + 20, neg), right = make_window(G, k.tid, k.end - 20, k.end + 2, neg);
    const Window& donor = neg ? right : left; const Window& acc = neg ? left : right;
    const double pws = pos_score(G, donor, posm) + pos_score(G, acc, posm + PMAX * 5);
    const double ss = (kmer_score(G, donor, kmer + (size_t)M_DONOR_T * KTAB) - kmer_score(G, donor, kmer + (size_t)M_DONOR_F * KTAB))
                    + (kmer_score(G, acc, kmer + (size_t)M_ACC_T * KTAB) - kmer_score(G, acc, kmer + (size_t)M_ACC_F * KTAB));
    o[0] = 0.0;                                                      // Genuine: a label, not a measurement
    o[1] = (double)(f.nb_raw - f.nb_ms); o[2] = (double)f.nb_dist; o[3] = (double)f.nb_rel; o[4] = f.entropy;
    o[5] = (double)f.nb_rel / (double)f.nb_raw;                     // Junction::getReliable2RawAlignmentRatio (junction.hpp:551-553): formed from the counts
    o[6] = (double)f.max_min_anc; o[7] = (double)f.maxmmes; o[8] = f.mean_mismatches;
    {   // Junction::calcIntronScore (junction.cc:953-956)
        const uint32_t size = (uint32_t)(k.end - k.start + 1);
        o[9] = (l95 == 0 || size <= l95) ? 0.0 : log((double)(size - l95));
    }
    o[10] = (double)min(f.hamming5p, f.hamming3p);
    if (coding_empty) o[11] = 0.0;
    else {   // Junction::calcCodingPotential (junction.cc:1328-1358)
        const double* ex = kmer + (size_t)M_EXON * KTAB; const double* in = kmer + (size_t)M_INTRON * KTAB;
        const Window le = make_window(G, k.tid, k.start - 82, k.start - 2, neg), li = make_window(G, k.tid, k.start, k.start + 80, neg);
        const Window ri = make_window(G, k.tid, k.end - 80, k.end, neg), re = make_window(G, k.tid, k.end + 1, k.end + 81, neg);
        o[11] = (kmer_score(G, le, ex) - kmer_score(G, le, in)) + (kmer_score(G, li, in) - kmer_score(G, li, ex))
              + (kmer_score(G, ri, in) - kmer_score(G, ri, ex)) + (kmer_score(G, re, ex) - kmer_score(G, re, in));
    }
    // the positional models are "empty" only until calcSplicingScores has looked a key up (operator[] inserts it), i.e. never
    // for a window of two or more bases (model_features.cc:169,203-207)
    const bool pw_empty = donor.len <= 1 && acc.len <= 1;
    o[12] = pw_empty ? 0.0 : pws; o[13] = pw_empty ? 0.0 : ss;
    // Junction::calcJunctionAnchorDepthLogDeviation (junction.cc:1384-1391); meanReadLength is the uint32 getter's value
    for (int q = 0; q < PJ_NB_JAD; q++) {
        double Ni = (double)f.jad[q]; if (Ni == 0.0) Ni = 0.000000000001;
        const double Pi = 1.0 - ((double)q / (f.mean_readlen / 2.0));
        const double Ei = (double)f.nb_raw * Pi;
        o[14 + q] = log2(Ni / Ei);
    }
}

} // namespace

struct pj_feat_models {
    pj_ctx* ctx = nullptr;
    unsigned long long* d_counts = nullptr;     // [N_KMER][KTAB] + [2][PMAX*5]
    double* d_prob = nullptr;                   // same layout, probabilities
    std::vector<unsigned long long> h_counts;
    bool coding_trained = false; int coding_empty = 1;
    static constexpr size_t NTAB = (size_t)N_KMER * KTAB + 2 * PMAX * 5;
};

namespace {

int gather_inputs(const pj_junction* rows, int64_t n, const uint8_t* subset, std::vector<JuncKey>& keys, std::vector<FeatIn>* fin) {
    keys.clear(); if (fin) fin->clear();
    for (int64_t i = 0; i < n; i++) {
        if (subset && !subset[i]) continue;
        const pj_junction& r = rows[i];
        JuncKey k; memset(&k, 0, sizeof k);
        k.tid = r.tid; k.start = r.start; k.end = r.end; k.neg = r.consensus_strand == PJ_STRAND_NEG ? 1 : 0;
        keys.push_back(k);
        if (fin) {
            FeatIn f; memset(&f, 0, sizeof f);
            f.nb_raw = r.nb_raw_aln; f.nb_ms = r.nb_ms_aln; f.nb_dist = r.nb_dist_aln; f.nb_rel = r.nb_rel_aln; f.max_min_anc = r.max_min_anc; f.maxmmes = r.maxmmes;
            f.hamming5p = r.hamming5p; f.hamming3p = r.hamming3p; f.entropy = r.entropy; f.rel2raw = r.rel2raw; f.mean_mismatches = r.mean_mismatches;
            f.mean_readlen = r.mean_readlen;
            for (int q = 0; q < PJ_NB_JAD; q++) f.jad[q] = r.jad[q];
            fin->push_back(f);
        }
    }
    return PJ_OK;
}

int check_rows(pj_ctx* c, const pj_junction* rows, int64_t n) {
    for (int64_t i = 0; i < n; i++) if (rows[i].tid < 0 || rows[i].tid >= c->n_targets) return fail(c, PJ_EINVAL, "pj_features: junction %lld refers to target %d", (long long)i, rows[i].tid);
    return PJ_OK;
}

Genome genome_of(pj_ctx* c) { return Genome{c->d_g2, c->d_gx, c->d_gsum, c->d_goff, c->d_glen, c->d_exc_pos, c->d_exc_byte, c->n_exc, c->n_exc_x, c->any_gx}; }

// counts -> probabilities: model[ctx][next] = count / sum over next (markov_model.cc:45-54, 94-103), fp64
void normalise(const unsigned long long* cnt, double* prob, int n_ctx) {
    for (int x = 0; x < n_ctx; x++) {
        double sum = 0; for (int s = 0; s < 5; s++) sum += (double)cnt[x * 5 + s];
        for (int s = 0; s < 5; s++) prob[x * 5 + s] = cnt[x * 5 + s] ? (double)cnt[x * 5 + s] / sum : 0.0;
    }
}

int upload_probabilities(pj_feat_models* m) {
    pj_ctx* c = m->ctx;
    std::vector<double> prob(pj_feat_models::NTAB, 0.0);
    for (int k = 0; k < N_KMER; k++) normalise(m->h_counts.data() + (size_t)k * KTAB, prob.data() + (size_t)k * KTAB, KCTX);
    for (int k = 0; k < 2; k++) normalise(m->h_counts.data() + (size_t)N_KMER * KTAB + (size_t)k * PMAX * 5, prob.data() + (size_t)N_KMER * KTAB + (size_t)k * PMAX * 5, PMAX);
    CU(c, cudaMemcpy(m->d_prob, prob.data(), prob.size() * sizeof(double), cudaMemcpyHostToDevice));
    auto any = [&](int model) { const unsigned long long* p = m->h_counts.data() + (size_t)model * KTAB; for (int i = 0; i < KTAB; i++) if (p[i]) return true; return false; };
    m->coding_empty = (any(M_EXON) && any(M_INTRON)) ? 0 : 1;       // ModelFeatures::isCodingPotentialModelEmpty
    return PJ_OK;
}

} // namespace

extern "C" {

int pj_features_create(pj_ctx* c, pj_feat_models** out) {
    if (!c || !out) return fail(c, PJ_EINVAL, "pj_features_create: null argument");
    if (!c->n_targets) return fail(c, PJ_ESTATE, "pj_features_create: call pj_targets_set and load the genome first");
    *out = nullptr;
    CU(c, cudaSetDevice(c->device));
    int rc = finish_genome(c); if (rc) return rc;
    pj_feat_models* m = new pj_feat_models();
    m->ctx = c; m->h_counts.assign(pj_feat_models::NTAB, 0ull);
    if (cudaMalloc(&m->d_counts, pj_feat_models::NTAB * 8) != cudaSuccess || cudaMalloc(&m->d_prob, pj_feat_models::NTAB * 8) != cudaSuccess) {
        cudaFree(m->d_counts); delete m; return fail(c, PJ_ECUDA, "pj_features_create: out of device memory");
    }
    CU(c, cudaMemset(m->d_prob, 0, pj_feat_models::NTAB * 8));
    *out = m;
    return PJ_OK;
}

void pj_features_destroy(pj_feat_models* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaFree(m->d_counts); cudaFree(m->d_prob);
    delete m;
}

// ModelFeatures::trainCodingPotentialModel over the junctions with subset[i] != 0 (all when subset is NULL)
int pj_features_train_coding(pj_feat_models* m, const pj_junction* rows, int64_t n, const uint8_t* subset) {
    if (!m || (n && !rows) || n < 0) return fail(m ? m->ctx : nullptr, PJ_EINVAL, "pj_features_train_coding: bad argument");
    pj_ctx* c = m->ctx;
    CU(c, cudaSetDevice(c->device));
    int rc = check_rows(c, rows, n); if (rc) return rc;
    if ((rc = finish_genome(c))) return rc;
    std::vector<JuncKey> keys; gather_inputs(rows, n, subset, keys, nullptr);
    const size_t off_e = (size_t)M_EXON * KTAB, off_i = (size_t)M_INTRON * KTAB;
    std::fill(m->h_counts.begin() + off_e, m->h_counts.begin() + off_e + KTAB, 0ull);           // train() replaces the model
    std::fill(m->h_counts.begin() + off_i, m->h_counts.begin() + off_i + KTAB, 0ull);
    if (!keys.empty()) {
        JuncKey* dk = nullptr;
        CU(c, cudaMalloc(&dk, keys.size() * sizeof(JuncKey)));
        cudaMemcpy(dk, keys.data(), keys.size() * sizeof(JuncKey), cudaMemcpyHostToDevice);
        cudaMemset(m->d_counts + off_e, 0, KTAB * 8); cudaMemset(m->d_counts + off_i, 0, KTAB * 8);
        const uint64_t threads = (uint64_t)keys.size() * 32;
        k_feat_count_coding<<<(unsigned)((threads + 255) / 256), 256, 0, c->compute_stream>>>((int64_t)keys.size(), dk, genome_of(c), m->d_counts + off_e, m->d_counts + off_i);
        cudaError_t e = cudaStreamSynchronize(c->compute_stream);
        cudaMemcpy(m->h_counts.data() + off_e, m->d_counts + off_e, KTAB * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(m->h_counts.data() + off_i, m->d_counts + off_i, KTAB * 8, cudaMemcpyDeviceToHost);
        cudaFree(dk);
        if (e != cudaSuccess) return fail(c, PJ_ECUDA, "pj_features_train_coding: %s", cudaGetErrorString(e));
    }
    return upload_probabilities(m);
}

// ModelFeatures::trainSplicingModels: donor / acceptor k-mer and positional models from `pass`, k-mer models from `fail`
int pj_features_train_splicing(pj_feat_models* m, const pj_junction* rows, int64_t n, const uint8_t* pass, const uint8_t* failset) {
    if (!m || (n && !rows) || n < 0 || (n && (!pass || !failset))) return fail(m ? m->ctx : nullptr, PJ_EINVAL, "pj_features_train_splicing: bad argument");
    pj_ctx* c = m->ctx;
    CU(c, cudaSetDevice(c->device));
    int rc = check_rows(c, rows, n); if (rc) return rc;
    if ((rc = finish_genome(c))) return rc;
    const size_t off_p = (size_t)N_KMER * KTAB;
    for (int pass_no = 0; pass_no < 2; pass_no++) {
        std::vector<JuncKey> keys; gather_inputs(rows, n, pass_no == 0 ? pass : failset, keys, nullptr);
        const size_t off_d = (size_t)(pass_no == 0 ? M_DONOR_T : M_DONOR_F) * KTAB, off_a = (size_t)(pass_no == 0 ? M_ACC_T : M_ACC_F) * KTAB;
        std::fill(m->h_counts.begin() + off_d, m->h_counts.begin() + off_d + KTAB, 0ull);
        std::fill(m->h_counts.begin() + off_a, m->h_counts.begin() + off_a + KTAB, 0ull);
        if (pass_no == 0) std::fill(m->h_counts.begin() + off_p, m->h_counts.end(), 0ull);
        if (keys.empty()) continue;
        JuncKey* dk = nullptr;
        CU(c, cudaMalloc(&dk, keys.size() * sizeof(JuncKey)));
        cudaMemcpy(dk, keys.data(), keys.size() * sizeof(JuncKey), cudaMemcpyHostToDevice);
        cudaMemset(m->d_counts + off_d, 0, KTAB * 8); cudaMemset(m->d_counts + off_a, 0, KTAB * 8);
        if (pass_no == 0) cudaMemset(m->d_counts + off_p, 0, 2 * PMAX * 5 * 8);
        const uint64_t threads = (uint64_t)keys.size() * 32;
        k_feat_count_splicing<<<(unsigned)((threads + 255) / 256), 256, 0, c->compute_stream>>>((int64_t)keys.size(), dk, genome_of(c), m->d_counts + off_d, m->d_counts + off_a,
            pass_no == 0 ? m->d_counts + off_p : nullptr, pass_no == 0 ? m->d_counts + off_p + PMAX * 5 : nullptr);
        cudaError_t e = cudaStreamSynchronize(c->compute_stream);
        cudaMemcpy(m->h_counts.data() + off_d, m->d_counts + off_d, KTAB * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(m->h_counts.data() + off_a, m->d_counts + off_a, KTAB * 8, cudaMemcpyDeviceToHost);
        if (pass_no == 0) cudaMemcpy(m->h_counts.data() + off_p, m->d_counts + off_p, 2 * PMAX * 5 * 8, cudaMemcpyDeviceToHost);
        cudaFree(dk);
        if (e != cudaSuccess) return fail(c, PJ_ECUDA, "pj_features_train_splicing: %s", cudaGetErrorString(e));
    }
    return upload_probabilities(m);
}

// ModelFeatures::calcIntronThreshold: the intron size at the 95th percentile of the subset (host; needs no device)
uint32_t pj_features_intron_threshold(const pj_junction* rows, int64_t n, const uint8_t* subset) {
    std::vector<uint32_t> sizes;
    for (int64_t i = 0; i < n; i++) if (!subset || subset[i]) sizes.push_back((uint32_t)(rows[i].end - rows[i].start + 1));
    if (sizes.empty()) return 0;
    std::sort(sizes.begin(), sizes.end());
    return sizes[(size_t)((double)sizes.size() * 0.95)];
}

// ModelFeatures::juncs2FeatureVectors: out[n][PJ_NB_FEATURES], the columns of setRow
int pj_features_run(pj_feat_models* m, const pj_junction* rows, int64_t n, uint32_t l95, double* out, float* device_ms) {
    if (!m || n < 0 || (n && (!rows || !out))) return fail(m ? m->ctx : nullptr, PJ_EINVAL, "pj_features_run: bad argument");
    pj_ctx* c = m->ctx;
    if (n == 0) return PJ_OK;
    CU(c, cudaSetDevice(c->device));
    int rc = check_rows(c, rows, n); if (rc) return rc;
    if ((rc = finish_genome(c))) return rc;
    std::vector<JuncKey> keys; std::vector<FeatIn> fin; gather_inputs(rows, n, nullptr, keys, &fin);
    JuncKey* dk = nullptr; FeatIn* df = nullptr; double* dout = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() { cudaFree(dk); cudaFree(df); cudaFree(dout); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
    if (cudaMalloc(&dk, (size_t)n * sizeof(JuncKey)) != cudaSuccess || cudaMalloc(&df, (size_t)n * sizeof(FeatIn)) != cudaSuccess ||
        cudaMalloc(&dout, (size_t)n * PJ_NB_FEATURES * 8) != cudaSuccess) { cleanup(); return fail(c, PJ_ECUDA, "pj_features_run: out of device memory"); }
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaMemcpyAsync(dk, keys.data(), (size_t)n * sizeof(JuncKey), cudaMemcpyHostToDevice, c->compute_stream);
    cudaMemcpyAsync(df, fin.data(), (size_t)n * sizeof(FeatIn), cudaMemcpyHostToDevice, c->compute_stream);
    cudaEventRecord(e0, c->compute_stream);
    k_feat_score<<<(unsigned)((n + 127) / 128), 128, 0, c->compute_stream>>>(n, dk, df, genome_of(c), m->d_prob, m->d_prob + (size_t)N_KMER * KTAB, m->coding_empty, l95, dout);
    cudaEventRecord(e1, c->compute_stream);
    cudaMemcpyAsync(out, dout, (size_t)n * PJ_NB_FEATURES * 8, cudaMemcpyDeviceToHost, c->compute_stream);
    cudaError_t e = cudaStreamSynchronize(c->compute_stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (device_ms) { *device_ms = 0; if (e == cudaSuccess) cudaEventElapsedTime(device_ms, e0, e1); }
    cleanup();
    if (e != cudaSuccess) return fail(c, PJ_ECUDA, "pj_features_run: %s", cudaGetErrorString(e));
    return PJ_OK;
}

} // extern "C"
