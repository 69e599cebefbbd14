// junc_driver.cpp — the `junc` stage driver: the B200-native counterpart of JunctionBuilder
// (src/junction_builder.cc:63-150, 228-291, 359-454).  Host work only: validate the prep directory, plan and
// run the parallel BGZF decode into pinned columnar staging buffers, shard targets over the GPUs (longest
// processing time first on index record counts; shards are independent, no collective), gather the junction
// rows, A12/A13 finalize and write the output files.  All junction arithmetic happens in the CUDA library.
#include "../../include/portcullis_junc_host.h"
#include "bam_io.hpp"
#include "fasta_io.hpp"
#include "junc_host.hpp"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <sstream>
#include <thread>
#if defined(__linux__)
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif
#include <sys/ioctl.h>
#include <unistd.h>

namespace fs = std::filesystem;
using pjio::BamFile; using pjio::ColumnarChunk; using pjio::DecodeTask; using pjio::FastaFile;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// PreparedFiles naming contract (src/prepare.hpp:114-140)
struct PrepPaths {
    std::string dir, bam, bai, csi, fasta, fai;
    explicit PrepPaths(const std::string& d) : dir(d) {
        bam = d + "/portcullis.sorted.alignments.bam"; bai = bam + ".bai"; csi = bam + ".csi";
        fasta = d + "/portcullis.genome.fa"; fai = fasta + ".fai";
    }
};
bool present(const std::string& p) { std::error_code ec; return fs::exists(p, ec) || fs::is_symlink(fs::symlink_status(p, ec)); }

// Ordered parallel pipeline: `produce(k)` runs on a pool of `threads` workers (out of order), `consume(k, payload)` runs
// on the calling thread strictly in task order.  At most `window` tasks are produced-but-not-consumed at any time.
template <typename Payload>
int ordered_pipeline(size_t n, int threads, size_t window, const std::function<int(size_t, Payload&)>& produce,
                     const std::function<int(size_t, Payload&)>& consume) {
    if (n == 0) return PJ_OK;
    threads = std::max(1, std::min<int>(threads, (int)n));
    std::vector<std::unique_ptr<Payload>> done(n);
    std::mutex mu; std::condition_variable cv;
    std::atomic<size_t> next{0}; size_t consumed = 0; bool failed = false; std::string fail_msg; int fail_code = PJ_OK;
    auto worker = [&]() {
#if defined(__linux__)
        // The in-order consumer is the critical path of a run (it was busy 3.3-4.4 s of a 4.5 s c3 run while the workers kept every core
        // occupied): the CPU-bound decode workers give way to it — and to the genome upload thread — when they compete for a core.
        (void)setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), 5);
#endif
        for (;;) {
            size_t k = next.fetch_add(1);
            if (k >= n) return;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || k < consumed + window; });
                if (failed) return;
            }
            auto p = std::make_unique<Payload>();
            int rc = PJ_OK; std::string msg;
            try { rc = produce(k, *p); if (rc) msg = g_err; }
            catch (const pjio::DataError& e) { rc = PJ_EDATA; msg = e.what(); }
            catch (const std::exception& e) { rc = PJ_EIO; msg = e.what(); }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (rc) { if (!failed) { failed = true; fail_code = rc; fail_msg = msg; } }
                else done[k] = std::move(p);
            }
            cv.notify_all();
            if (rc) return;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(worker);
    int rc = PJ_OK;
    static const bool trace = getenv("PJ_TRACE") != nullptr;
    double t_wait = 0, t_busy = 0;
    for (size_t k = 0; k < n; k++) {
        std::unique_ptr<Payload> p;
        const double tw0 = trace ? now_s() : 0;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return failed || done[k]; });
            if (failed) { rc = fail(fail_code, fail_msg); break; }
            p = std::move(done[k]);
        }
        const double tw1 = trace ? now_s() : 0;
        rc = consume(k, *p);
        if (trace) { t_wait += tw1 - tw0; t_busy += now_s() - tw1; }
        { std::lock_guard<std::mutex> lk(mu); consumed = k + 1; if (rc) failed = true; }
        cv.notify_all();
        if (rc) break;
    }
    { std::lock_guard<std::mutex> lk(mu); if (rc) failed = true; consumed = n; }
    cv.notify_all();
    for (auto& t : pool) t.join();
    if (trace) fprintf(stderr, "[pj pipeline] %zu tasks on %d workers: the in-order consumer was busy %.3f s and waited %.3f s for the workers\n", n, threads, t_busy, t_wait);
    return rc;
}

void chunk_view(const ColumnarChunk& c, pj_batch* b) {
    memset(b, 0, sizeof *b);
    b->n_records = c.n(); b->tid = c.tid.data(); b->pos = c.pos.data(); b->flag = c.flag.data(); b->mapq = c.mapq.data(); b->xs = c.xs.data();
    b->l_qseq = c.l_qseq.data(); b->mtid = c.mtid.data(); b->mpos = c.mpos.data(); b->cigar_off = c.cigar_off.data(); b->cigar = c.cigar.data();
    b->seq_off = c.seq_off.data(); b->seq4 = c.seq4.data();
    b->name_code = (c.with_names && (int64_t)c.name_code.size() == c.n()) ? c.name_code.data() : nullptr;
}

} // namespace

// ------------------------------------------------------------------------------------------------
struct pjh_prep {
    PrepPaths paths; BamFile bam; FastaFile fasta; bool indexed = false;
    ColumnarChunk decoded; std::string genome; int32_t genome_tid = -1;
    bool want_names = false;
    explicit pjh_prep(const std::string& d) : paths(d) {}
};

extern "C" {

const char* pjh_last_error(void) { return g_err.c_str(); }

void pjh_options_default(pjh_options* o) {
    memset(o, 0, sizeof *o);
    o->output_prefix = "portcullis_junc/portcullis"; o->threads = 1; o->n_gpus = 1; o->orientation = PJ_ORIENT_UNKNOWN;
    o->strandedness = PJ_STRANDED_UNKNOWN; o->source = "portcullis"; o->version = "1.2.4";
}

int pjh_prep_open(const char* prep_dir, int use_csi, pjh_prep** out) {
    if (!prep_dir || !out) return fail(PJ_EINVAL, "pjh_prep_open: null argument");
    *out = nullptr;
    auto p = std::make_unique<pjh_prep>(prep_dir);
    // PreparedFiles::valid (src/prepare.cc:57-75)
    if (!present(p->paths.bam)) return fail(PJ_EIO, "Could not find prepared BAM file at: " + p->paths.bam);
    if (!present(use_csi ? p->paths.csi : p->paths.bai) || !present(p->paths.fasta) || !present(p->paths.fai))
        return fail(PJ_EIO, "Prepared data is not complete: " + p->paths.dir);
    try {
        p->bam.open(p->paths.bam);
        p->indexed = use_csi ? p->bam.load_csi(p->paths.csi) : p->bam.load_bai(p->paths.bai);
        p->fasta.open(p->paths.fasta, p->paths.fai);
    } catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    *out = p.release();
    return PJ_OK;
}
void pjh_prep_close(pjh_prep* p) { delete p; }
int32_t pjh_prep_n_targets(const pjh_prep* p) { return p ? (int32_t)p->bam.header().names.size() : 0; }
const char* pjh_prep_target_name(const pjh_prep* p, int32_t t) { return (p && t >= 0 && t < pjh_prep_n_targets(p)) ? p->bam.header().names[t].c_str() : nullptr; }
int32_t pjh_prep_target_len(const pjh_prep* p, int32_t t) { return (p && t >= 0 && t < pjh_prep_n_targets(p)) ? p->bam.header().lens[t] : -1; }
int64_t pjh_prep_target_records(const pjh_prep* p, int32_t t) {
    if (!p || !p->indexed || t < 0 || (size_t)t >= p->bam.index().size() || !p->bam.index()[t].has_counts) return -1;
    return (int64_t)(p->bam.index()[t].n_mapped + p->bam.index()[t].n_unmapped);
}

static uint64_t target_weight(const pjh_prep* p, int32_t t) {
    const int64_t n = pjh_prep_target_records(p, t);
    if (n > 0) return (uint64_t)n;
    std::vector<DecodeTask> tasks; p->bam.plan_target(t, 4u << 20, tasks);
    uint64_t bytes = 0; for (auto& k : tasks) bytes += k.approx_bytes;
    return bytes / 64;
}

int pjh_inflate_selftest(int32_t n_cases) { return pjio::inflate_selftest(n_cases); }
int pjh_format_selftest(int32_t n_cases) { return pjhost::format_selftest(n_cases); }

int pjh_plan_shards(const pjh_prep* p, int32_t n_gpus, int32_t* gpu_of_target) {
    if (!p || n_gpus < 1 || !gpu_of_target) return fail(PJ_EINVAL, "pjh_plan_shards: bad argument");
    const int32_t T = pjh_prep_n_targets(p);
    std::vector<uint64_t> weight((size_t)T, 0);
    if (p->indexed) for (int32_t t = 0; t < T; t++) weight[t] = target_weight(p, t);
    std::vector<int32_t> ord((size_t)T); for (int32_t t = 0; t < T; t++) ord[t] = t;
    std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return weight[a] > weight[b]; });
    std::vector<uint64_t> load((size_t)n_gpus, 0);
    for (int32_t t : ord) { const size_t g = (size_t)(std::min_element(load.begin(), load.end()) - load.begin()); gpu_of_target[t] = (int32_t)g; load[g] += weight[t] + 1; }
    return PJ_OK;
}

int pjh_prep_decode(pjh_prep* p, int32_t tid, int32_t threads, pj_batch* out) {
    if (!p || !out) return fail(PJ_EINVAL, "pjh_prep_decode: null argument");
    std::vector<DecodeTask> tasks;
    const int32_t T = pjh_prep_n_targets(p);
    if (p->indexed) { for (int32_t t = (tid < 0 ? 0 : tid); t < (tid < 0 ? T : tid + 1); t++) p->bam.plan_target(t, 2u << 20, tasks); }
    else { DecodeTask w = p->bam.whole_file_task(); if (tid >= 0) { w.tid = tid; } tasks.push_back(w); }
    p->decoded.clear(); p->decoded.with_names = p->want_names; p->decoded.lean = false;
    int rc = ordered_pipeline<ColumnarChunk>(tasks.size(), threads, (size_t)threads * 3 + 2,
        [&](size_t k, ColumnarChunk& c) { c.with_names = p->want_names; p->bam.decode(tasks[k], c); return PJ_OK; },
        [&](size_t, ColumnarChunk& c) { p->decoded.append(c); return PJ_OK; });
    if (rc) return rc;
    chunk_view(p->decoded, out);
    return PJ_OK;
}

void pjh_prep_want_names(pjh_prep* p, int32_t on) { if (p) p->want_names = on != 0; }

int pjh_prep_genome(pjh_prep* p, int32_t tid, const char** bases, int64_t* n_bases) {
    if (!p || !bases || !n_bases || tid < 0 || tid >= pjh_prep_n_targets(p)) return fail(PJ_EINVAL, "pjh_prep_genome: bad argument");
    if (p->genome_tid != tid) {
        const pjio::FaiEntry* e = p->fasta.find(p->bam.header().names[tid]);
        if (!e) return fail(PJ_EDATA, "The sequence \"" + p->bam.header().names[tid] + "\" not found in the genome index");
        p->fasta.fetch_all(*e, p->genome); p->genome_tid = tid;
    }
    *bases = p->genome.data(); *n_bases = (int64_t)p->genome.size();
    return PJ_OK;
}

int pjh_separate_bams(const char* prep_dir, const char* output_prefix, int32_t use_csi, int32_t threads, uint64_t* counts) {
    if (!prep_dir || !output_prefix) return fail(PJ_EINVAL, "pjh_separate_bams: null argument");
    pjh_prep* prep = nullptr;
    int rc = pjh_prep_open(prep_dir, use_csi, &prep);
    if (rc) return rc;
    std::unique_ptr<pjh_prep> guard(prep);
    const std::string pre(output_prefix);
    fs::path parent = fs::path(pre).parent_path();
    if (!parent.empty()) { std::error_code ec; fs::create_directories(parent, ec); }
    pjio::SeparateCounts sc;
    try { pjio::separate_bams(prep->bam, pre + ".spliced.bam", pre + ".unspliced.bam", pre + ".unmapped.bam", use_csi != 0, std::max(1, threads), sc); }
    catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    if (counts) { counts[0] = sc.spliced; counts[1] = sc.unspliced; counts[2] = sc.unmapped; }
    return PJ_OK;
}

int pjh_write_outputs(const char* output_prefix, const pj_junction* rows, int64_t n_rows, int32_t n_targets, const char* const* names,
                      const int32_t* lens, const char* source, const char* version, int32_t exon_gff, int32_t intron_gff) {
    return pjh_write_outputs_extra(output_prefix, rows, nullptr, n_rows, n_targets, names, lens, source, version, exon_gff, intron_gff);
}

int pjh_write_outputs_extra(const char* output_prefix, const pj_junction* rows, const pj_junction_extra* extra, int64_t n_rows, int32_t n_targets,
                            const char* const* names, const int32_t* lens, const char* source, const char* version, int32_t exon_gff, int32_t intron_gff) {
    if (!output_prefix || (n_rows && !rows) || !names || !lens) return fail(PJ_EINVAL, "pjh_write_outputs: null argument");
    try {
        std::vector<pjhost::TargetInfo> t((size_t)n_targets);
        for (int32_t i = 0; i < n_targets; i++) { t[i].name = names[i]; t[i].length = lens[i]; }
        const std::string pre(output_prefix), src(source ? source : "portcullis"), ver(version ? version : "");
        fs::path parent = fs::path(pre).parent_path();
        if (!parent.empty()) { std::error_code ec; fs::create_directories(parent, ec); }
        // the files are independent: write them concurrently (each writer also formats its rows on a few threads)
        std::string werr; std::mutex wmu;
        auto guarded = [&](const std::function<void()>& f) { try { f(); } catch (const std::exception& e) { std::lock_guard<std::mutex> lk(wmu); werr = e.what(); } };
        std::vector<std::thread> wt;
        wt.emplace_back([&]() { guarded([&]() { pjhost::write_bed(pre + ".junctions.bed", rows, n_rows, t, src, ver); }); });
        if (exon_gff) wt.emplace_back([&]() { guarded([&]() { pjhost::write_exon_gff(pre + ".junctions.exon.gff3", rows, n_rows, t, src); }); });
        if (intron_gff) wt.emplace_back([&]() { guarded([&]() { pjhost::write_intron_gff(pre + ".junctions.intron.gff3", rows, n_rows, t, src); }); });
        guarded([&]() { pjhost::write_tab(pre + ".junctions.tab", rows, n_rows, t, extra); });
        for (auto& x : wt) x.join();
        if (!werr.empty()) return fail(PJ_EIO, werr);
    } catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    return PJ_OK;
}

int pj_genome_load_fasta(pj_ctx* ctx, const char* fasta_path, const char* fai_path, int32_t n_targets, const char* const* names) {
    if (!ctx || !fasta_path || !fai_path || !names) return fail(PJ_EINVAL, "pj_genome_load_fasta: null argument");
    try {
        FastaFile fa; fa.open(fasta_path, fai_path);
        std::string seq;
        for (int32_t t = 0; t < n_targets; t++) {
            const pjio::FaiEntry* e = fa.find(names[t]);
            if (!e) continue;                        // like the reference, a missing sequence only matters if a junction needs it
            fa.fetch_all(*e, seq);
            int rc = pj_genome_set_target(ctx, t, seq.data(), (int64_t)seq.size());
            if (rc) return fail(rc, pj_last_error(ctx));
        }
    } catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    return PJ_OK;
}

// ------------------------------------------------------------------------------------------------
// Work plan.  A *segment* is a list of decode tasks that one GPU context runs as one shard (pj_shard_begin ... pj_shard_fetch); a
// *part* is the list of segments one GPU owns.  Two plans:
//   * ranges (default): the BAM-ordered list of decode tasks, weighted by the index's record counts, is cut into n_parts contiguous
//     ranges of equal weight, and every range into segments of at most `seg_budget` records.  A cut that falls inside a target is
//     moved to the next record no spliced read spans (BamFile::find_gap_cut), so no junction has reads on both sides: segments are
//     independent, their rows concatenate in (tid, start, end) order, and device memory is bounded by the segment size instead of
//     the whole shard (the reference bounds memory the same way, by flushing junctions the scan has passed, junction_builder.cc:324-331).
//     An oversized target (or a genome with fewer targets than GPUs) no longer pins the balance to whole targets.
//   * whole targets (`--extra`, or PJ_WHOLE_TARGETS=1): LPT of whole targets on record counts, one segment per part; the extra
//     metrics need every unspliced record of a target on one device.
// ------------------------------------------------------------------------------------------------
namespace {

struct Segment { std::vector<DecodeTask> tasks; std::vector<int32_t> targets; uint64_t weight = 0; };
typedef std::vector<Segment> Part;

struct PlanItem { int32_t tid; size_t k; uint64_t w; };        // task k of target tid, estimated records

uint64_t env_u64(const char* name, uint64_t dflt) { const char* e = getenv(name); return (e && *e) ? strtoull(e, nullptr, 10) : dflt; }

void seg_add_task(Segment& sg, const DecodeTask& t, uint64_t w) {
    sg.tasks.push_back(t); sg.weight += w;
    if (sg.targets.empty() || sg.targets.back() != t.tid) sg.targets.push_back(t.tid);
}

// returns PJ_OK; parts.size() == n_parts (parts may be empty when there is less work than parts)
int plan_parts(const pjh_prep* prep, int n_parts, bool whole_targets, uint64_t seg_budget, std::vector<Part>& parts, int* n_gap_cuts) {
    const int32_t T = (int32_t)prep->bam.header().names.size();
    parts.assign((size_t)n_parts, Part());
    if (n_gap_cuts) *n_gap_cuts = 0;
    if (!prep->indexed) {                                     // no index: one whole-file task on part 0
        Segment sg; sg.tasks.push_back(prep->bam.whole_file_task()); for (int32_t t = 0; t < T; t++) sg.targets.push_back(t);
        parts[0].push_back(std::move(sg));
        return PJ_OK;
    }
    std::vector<std::vector<DecodeTask>> ttasks((size_t)T);
    std::vector<uint64_t> weight((size_t)T, 0);
    // decode tasks of about 4 MB of BGZF data; smaller for small files so that every part still gets tasks to balance with
    const uint64_t task_bytes = std::min<uint64_t>(4u << 20, std::max<uint64_t>(64u << 10, prep->bam.file().size() / ((uint64_t)n_parts * 64)));
    for (int32_t t = 0; t < T; t++) { prep->bam.plan_target(t, task_bytes, ttasks[t]); weight[t] = ttasks[t].empty() ? 0 : target_weight(prep, t); }
    if (whole_targets) {
        std::vector<int32_t> owner((size_t)T, 0);
        int rc = pjh_plan_shards(prep, n_parts, owner.data());
        if (rc) return rc;
        for (int g = 0; g < n_parts; g++) {
            Segment sg;
            for (int32_t t = 0; t < T; t++) if (owner[t] == g) {                 // ascending tid = BAM order inside a shard
                if (ttasks[t].empty()) { sg.targets.push_back(t); continue; }
                uint64_t bytes = 0; for (auto& k : ttasks[t]) bytes += k.approx_bytes;
                for (auto& k : ttasks[t]) seg_add_task(sg, k, bytes ? weight[t] * k.approx_bytes / bytes : 0);
            }
            if (!sg.tasks.empty() || !sg.targets.empty()) parts[(size_t)g].push_back(std::move(sg));
        }
        return PJ_OK;
    }
    // ---- flatten, weigh, place the boundaries ----
    std::vector<PlanItem> items;
    uint64_t W = 0;
    for (int32_t t = 0; t < T; t++) {
        uint64_t bytes = 0; for (auto& k : ttasks[t]) bytes += std::max<uint64_t>(k.approx_bytes, 1);
        for (size_t k = 0; k < ttasks[t].size(); k++) {
            const uint64_t w = std::max<uint64_t>(1, (uint64_t)((double)weight[t] * (double)std::max<uint64_t>(ttasks[t][k].approx_bytes, 1) / (double)bytes));
            items.push_back(PlanItem{t, k, w}); W += w;
        }
    }
    const size_t n = items.size();
    if (n == 0) return PJ_OK;
    std::vector<uint64_t> cum(n + 1, 0);
    for (size_t i = 0; i < n; i++) cum[i + 1] = cum[i] + items[i].w;
    // boundary = (item index where the next segment starts, part that segment belongs to)
    struct Boundary { size_t at; int part; };
    std::vector<Boundary> bounds;
    auto first_at_or_after = [&](double target) { return (size_t)(std::lower_bound(cum.begin(), cum.end(), (uint64_t)std::ceil(target)) - cum.begin()); };
    size_t prev = 0;
    for (int g = 0; g < n_parts; g++) {
        size_t end = g + 1 == n_parts ? n : std::min(n, std::max(prev, first_at_or_after((double)W * (g + 1) / n_parts)));
        // index `end` in cum[] means items [prev, end) are in the part; choose the nearer of the two candidate cuts
        if (end > prev && end < n && g + 1 < n_parts) { const double tgt = (double)W * (g + 1) / n_parts; if (tgt - (double)cum[end - 1] < (double)cum[end] - tgt && end - 1 > prev) end--; }
        const uint64_t wp = cum[end] - cum[prev];
        const int nseg = (int)std::max<uint64_t>(1, (wp + seg_budget - 1) / std::max<uint64_t>(seg_budget, 1));
        size_t sp = prev;
        for (int q = 0; q < nseg; q++) {
            bounds.push_back(Boundary{sp, g});
            size_t se = q + 1 == nseg ? end : std::min(end, std::max(sp, first_at_or_after((double)cum[prev] + (double)wp * (q + 1) / nseg)));
            sp = se;
        }
        prev = end;
    }
    // ---- build the segments; intra-target boundaries become gap cuts ----
    size_t cur = 0;                                            // next item not yet handed out
    bool have_head = false; DecodeTask head{}; uint64_t head_w = 0;
    for (size_t b = 0; b < bounds.size(); b++) {
        const size_t stop = b + 1 < bounds.size() ? std::max(bounds[b + 1].at, cur) : n;   // items [cur, stop) belong to this segment
        Segment sg;
        if (have_head) { seg_add_task(sg, head, head_w); have_head = false; }
        for (; cur < stop; cur++) seg_add_task(sg, ttasks[items[cur].tid][items[cur].k], items[cur].w);
        if (cur < n && !sg.tasks.empty() && sg.tasks.back().tid == items[cur].tid) {
            // the next segment would start inside target t: look for the gap cut from task (t, k) on
            const int32_t t = items[cur].tid; const size_t k = items[cur].k;
            const DecodeTask& at = ttasks[t][k];
            uint64_t cut_voff = 0; int32_t cut_pos = 0;
            if (prep->bam.find_gap_cut(t, at.pos_lo, at.voff, &cut_voff, &cut_pos)) {
                if (n_gap_cuts) (*n_gap_cuts)++;
                DecodeTask tail = at; tail.pos_hi = INT32_MAX; tail.end_voff = cut_voff; tail.approx_bytes = 1;
                seg_add_task(sg, tail, 0);
                size_t k2 = k; while (k2 + 1 < ttasks[t].size() && ttasks[t][k2].pos_hi <= cut_pos) k2++;
                head = ttasks[t][k2]; head.voff = cut_voff; head.pos_lo = cut_pos; head.end_voff = 0; have_head = true;
                head_w = 0;
                for (; cur < n && items[cur].tid == t && items[cur].k <= k2; cur++) head_w += items[cur].w;   // these items are covered by tail + head
            } else {
                for (; cur < n && items[cur].tid == t; cur++) seg_add_task(sg, ttasks[t][items[cur].k], items[cur].w);   // no gap before the target ends
            }
        }
        if (!sg.tasks.empty()) parts[(size_t)bounds[b].part].push_back(std::move(sg));
    }
    if (have_head) {                                           // a cut after the last boundary cannot happen, but never drop work
        Segment sg; seg_add_task(sg, head, head_w);
        for (; cur < n; cur++) seg_add_task(sg, ttasks[items[cur].tid][items[cur].k], items[cur].w);
        parts.back().push_back(std::move(sg));
    }
    return PJ_OK;
}

} // namespace

// rows + per-target scalars of one part (one GPU), before A12/A13
struct pjh_partial {
    std::vector<pj_junction> rows; std::vector<pj_target_stats> stats; pjh_report rep;
    std::thread teardown;                  // pj_destroy of the part's context, running while the caller gathers the rows
    ~pjh_partial() { if (teardown.joinable()) teardown.join(); }
};

// ------------------------------------------------------------------------------------------------
// JunctionBuilder::process equivalent
// ------------------------------------------------------------------------------------------------
namespace {

struct GpuOut {
    pj_ctx* ctx = nullptr; std::vector<pj_junction_extra> extra; std::vector<pj_junction> rows; std::vector<pj_target_stats> stats;
    float gpu_ms = 0; int launches = 0; double genome_s = 0, decode_s = 0, init_s = 0, run_s = 0, teardown_s = 0; int n_segments = 0;
    int rc = PJ_OK; std::string err;
};

void merge_stats(std::vector<pj_target_stats>& into, const pj_target_stats* from, int32_t T) {
    for (int32_t t = 0; t < T; t++) {
        pj_target_stats& a = into[(size_t)t]; const pj_target_stats& b = from[t];
        a.spliced_count += b.spliced_count; a.unspliced_count += b.unspliced_count; a.sum_query_lengths += b.sum_query_lengths;
        a.min_query_length = std::min(a.min_query_length, b.min_query_length); a.max_query_length = std::max(a.max_query_length, b.max_query_length);
    }
}

// One GPU: CUDA context + library context + genome of the part's targets, then the part's segments one after the other
// (decode workers -> pinned staging -> H2D -> pj_shard_run -> rows appended).  `device` is the CUDA ordinal.
void run_part(const pjh_options* o, pjh_prep* prep, const Part& part, int device, int threads, bool extra, GpuOut& out) {
    const pjio::BamHeader& H = prep->bam.header();
    const int32_t T = (int32_t)H.names.size();
    auto bail = [&](int code, const std::string& m) { out.rc = code; out.err = m; };
    out.stats.assign((size_t)T, pj_target_stats{0, 0, 0, INT32_MAX, 0});
    if (part.empty()) return;                                          // less work than GPUs
    uint64_t hint = 0; for (auto& sg : part) hint = std::max(hint, sg.weight);
    std::vector<int32_t> my_targets;
    for (auto& sg : part) for (int32_t t : sg.targets) if (my_targets.empty() || my_targets.back() != t) my_targets.push_back(t);
    // ---- GPU side on its own host thread: CUDA start-up takes about a second on a B200, so BGZF decode starts right away and
    // only the staging copies wait for the context. ----
    pj_ctx* ctx = nullptr;
    std::mutex gm; std::condition_variable gcv; bool ready = false; int gpu_rc = PJ_OK; std::string gpu_err;
    int genome_rc = PJ_OK; std::string genome_err;
    std::vector<char> genome_done((size_t)T, 0); bool genome_finished = false;   // guarded by gm: a segment only waits for the genome of ITS targets
    const double ti = now_s();
    std::thread gpu_thread([&]() {
        pj_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.device = device; cfg.orientation = o->orientation;
        cfg.reserved[2] = 4;                                           // pinned staging buffers (filled by one thread, drained by the copy engine)
        cfg.extra_metrics = extra ? 1 : 0;
        int r = pj_create(&cfg, &ctx);
        std::string em;
        if (r) em = pj_global_last_error();
        if (!r && (r = pj_targets_set(ctx, T, H.lens.data()))) em = pj_last_error(ctx);
        if (!r && (r = pj_shard_begin(ctx, (int64_t)hint + 1024, (int64_t)hint * 4 + 1024, (int64_t)hint * 56 + 1024))) em = pj_last_error(ctx);
        out.init_s = now_s() - ti;
        { std::lock_guard<std::mutex> lk(gm); ready = true; gpu_rc = r; gpu_err = em; }
        gcv.notify_all();
        if (r) return;
        // genome: only this part's targets become resident on this GPU (own CUDA stream; overlaps the batch submission).  Reading a
        // target out of the FASTA file (line by line, faidx semantics) is host work of about 1 s per Gb on one thread, slower than the
        // decode workers cover the genome on a human-scale run: two parser threads read ahead, this thread uploads in target order.
        const double tg = now_s();
        auto done = [&](int32_t t, int code, const std::string& msg) {
            { std::lock_guard<std::mutex> lk(gm); if (t >= 0) genome_done[(size_t)t] = 1; if (code) { genome_rc = code; genome_err = msg; } if (t < 0 || code) genome_finished = true; }
            gcv.notify_all();
        };
        const size_t nt = my_targets.size();
        std::vector<std::string> seqs(nt); std::vector<char> state(nt, 0);         // 0 not parsed, 1 parsed, 2 nothing to load, 3 failed
        std::vector<std::string> perr(nt);
        std::mutex pm; std::condition_variable pcv; std::atomic<size_t> pnext{0}; size_t uploaded = 0; bool pstop = false;
        const int n_parsers = (int)std::min<size_t>(nt, (size_t)std::max(1, std::min(2, threads / 4)));
        std::vector<std::thread> parsers;
        for (int w = 0; w < n_parsers; w++) parsers.emplace_back([&]() {
            for (;;) {
                const size_t i = pnext.fetch_add(1);
                if (i >= nt) return;
                { std::unique_lock<std::mutex> lk(pm); pcv.wait(lk, [&] { return pstop || i < uploaded + 4; }); if (pstop) return; }   // at most four targets in host memory
                const int32_t t = my_targets[i];
                char st = 1; std::string msg;
                const pjio::FaiEntry* e = nullptr;
                if (prep->indexed && prep->bam.index()[(size_t)t].first_voff == 0) st = 2;       // no records -> no junctions -> no genome needed
                else if (!(e = prep->fasta.find(H.names[t]))) st = 2;
                else { try { prep->fasta.fetch_all(*e, seqs[i]); } catch (const std::exception& ex) { st = 3; msg = ex.what(); } }
                { std::lock_guard<std::mutex> lk(pm); state[i] = st; perr[i] = msg; }
                pcv.notify_all();
            }
        });
        struct JoinParsers { std::vector<std::thread>& th; std::mutex& m; std::condition_variable& cv; bool& stop;
                             ~JoinParsers() { { std::lock_guard<std::mutex> lk(m); stop = true; } cv.notify_all(); for (auto& t : th) if (t.joinable()) t.join(); } } join_parsers{parsers, pm, pcv, pstop};
        for (size_t i = 0; i < nt; i++) {                                  // ascending = the order the segments need them
            const int32_t t = my_targets[i];
            char st;
            { std::unique_lock<std::mutex> lk(pm); pcv.wait(lk, [&] { return state[i] != 0; }); st = state[i]; }
            if (st == 3) { done(t, PJ_EIO, perr[i]); return; }
            if (st == 1) {
                int q = pj_genome_set_target(ctx, t, seqs[i].data(), (int64_t)seqs[i].size());
                if (q) { done(t, q, pj_last_error(ctx)); return; }
                std::string().swap(seqs[i]);
            }
            { std::lock_guard<std::mutex> lk(pm); uploaded = i + 1; }
            pcv.notify_all();
            done(t, PJ_OK, "");
        }
        out.genome_s = now_s() - tg;
        done(-1, PJ_OK, "");
    });
    struct Cleanup {
        std::thread& t; pj_ctx*& c; double* td;
        ~Cleanup() { if (t.joinable()) t.join(); if (c) { const double a = now_s(); pj_destroy(c); *td = now_s() - a; } }
    } cleanup{gpu_thread, ctx, &out.teardown_s};
    auto wait_ready = [&]() -> int { std::unique_lock<std::mutex> lk(gm); gcv.wait(lk, [&] { return ready; }); return gpu_rc; };
    // copy one decoded chunk into a pinned staging buffer of the context
    const bool lean = prep->indexed;                                   // a decode task of an indexed BAM lies on one target: the lean batch form applies
    const bool keep_mate = o->orientation == PJ_ORIENT_FR || o->orientation == PJ_ORIENT_RF || o->orientation == PJ_ORIENT_FF;
    // The copy of a decoded chunk into pinned staging stays on this thread.  (Splitting it over short-lived helper threads was measured
    // on the 16-core box: with every core busy decoding, waiting for the helpers to be scheduled cost more than the copy, +1 s on c3.)
    auto add_copy = [&](const void* dst, const void* src, size_t n) { if (n) memcpy(const_cast<void*>(dst), src, n); };
    auto run_copies = [&]() {};
    auto stage = [&](const ColumnarChunk& ch, pj_batch& st) -> int {
        if (ch.lean) {
            // about half the bytes of the classic form cross PCIe: no tid / cigar_off / seq_off columns (formed on the device), SEQ at
            // 2 bits per base, mate columns only when the orientation needs them
            if (ch.runs.size() != 1) return fail(PJ_EINVAL, "internal: a lean chunk must lie on one target");
            int q = pj_staging_acquire_lean(ctx, ch.n(), (int64_t)ch.cigar.size(), (int64_t)ch.seq2.size(), (int64_t)ch.seqx_pos.size(), &st);
            if (q) return fail(q, pj_last_error(ctx));
            const size_t n = (size_t)ch.n();
            add_copy(st.pos, ch.pos.data(), n * 4); add_copy(st.flag, ch.flag.data(), n * 2); add_copy(st.mapq, ch.mapq.data(), n);
            add_copy(st.xs, ch.xs.data(), n); add_copy(st.l_qseq, ch.l_qseq.data(), n * 4); add_copy(st.n_cigar, ch.n_cigar.data(), n * 2);
            if (keep_mate) { add_copy(st.mtid, ch.mtid.data(), n * 4); add_copy(st.mpos, ch.mpos.data(), n * 4); } else { st.mtid = nullptr; st.mpos = nullptr; }
            add_copy(st.cigar, ch.cigar.data(), ch.cigar.size() * 4); add_copy(st.seq2, ch.seq2.data(), ch.seq2.size());
            if (!ch.seqx_pos.empty()) { add_copy(st.seqx_pos, ch.seqx_pos.data(), ch.seqx_pos.size() * 8); add_copy(st.seqx_code, ch.seqx_code.data(), ch.seqx_code.size()); }
            if (extra) add_copy(st.name_code, ch.name_code.data(), n * 8);
            run_copies();
            st.n_records = ch.n(); st.const_tid = ch.runs[0].tid; st.n_cigar_total = (int64_t)ch.cigar.size(); st.n_seq2_bytes = (int64_t)ch.seq2.size();
            st.n_seqx = (int64_t)ch.seqx_pos.size();
            return PJ_OK;
        }
        int q = pj_staging_acquire(ctx, ch.n(), (int64_t)ch.cigar.size(), (int64_t)ch.seq4.size(), &st);
        if (q) return fail(q, pj_last_error(ctx));
        const size_t n = (size_t)ch.n();
        memcpy((void*)st.tid, ch.tid.data(), n * 4); memcpy((void*)st.pos, ch.pos.data(), n * 4); memcpy((void*)st.flag, ch.flag.data(), n * 2);
        memcpy((void*)st.mapq, ch.mapq.data(), n); memcpy((void*)st.xs, ch.xs.data(), n); memcpy((void*)st.l_qseq, ch.l_qseq.data(), n * 4);
        memcpy((void*)st.mtid, ch.mtid.data(), n * 4); memcpy((void*)st.mpos, ch.mpos.data(), n * 4);
        memcpy((void*)st.cigar_off, ch.cigar_off.data(), (n + 1) * 4); memcpy((void*)st.cigar, ch.cigar.data(), ch.cigar.size() * 4);
        memcpy((void*)st.seq_off, ch.seq_off.data(), (n + 1) * 8); memcpy((void*)st.seq4, ch.seq4.data(), ch.seq4.size());
        if (extra) memcpy((void*)st.name_code, ch.name_code.data(), n * 8);
        st.n_records = ch.n();
        return PJ_OK;
    };
    // alignments: workers inflate + parse a task into a pooled pageable chunk (recycled, so its vectors keep their capacity and
    // their faulted-in pages); this thread takes the chunks in BAM order, copies each into one of a few pinned staging buffers
    // and enqueues the host-to-device copies.  Growing a large pinned pool costs far more (cudaMallocHost, about 0.6 ms per MB,
    // serialised) than this one extra memcpy, and the decode can run ahead of a context that is still starting.
    struct ChunkPool {
        std::mutex mu; std::vector<std::unique_ptr<ColumnarChunk>> free_list;
        std::unique_ptr<ColumnarChunk> get() {
            { std::lock_guard<std::mutex> lk(mu); if (!free_list.empty()) { auto c = std::move(free_list.back()); free_list.pop_back(); return c; } }
            return std::make_unique<ColumnarChunk>();
        }
        void put(std::unique_ptr<ColumnarChunk> c) { c->clear(); std::lock_guard<std::mutex> lk(mu); free_list.push_back(std::move(c)); }
    } pool;
    struct Payload { std::unique_ptr<ColumnarChunk> chunk; };
    // decoded-but-unsubmitted chunks (pageable): enough for the decode to keep going while a CUDA context is still starting
    const size_t window = std::max<size_t>((size_t)threads * 4 + 2, 192);
    bool first = true;
    {   // address space for the rows of all segments up front (untouched pages cost nothing): growing the vector segment by segment
        // copied hundreds of MB of 256-byte rows around on a human-scale run
        uint64_t recs = 0; for (auto& sg : part) recs += sg.weight;
        out.rows.reserve((size_t)std::min<uint64_t>(recs / 8 + (1u << 20), 16u << 20));
    }
    // ONE ordered pipeline over the decode tasks of all segments: when the last task of a segment has been submitted the consumer runs
    // the shard (a few ms of GPU time + the row fetch) and opens the next one, while the workers keep decoding ahead into the window.
    // (One pipeline per segment left the workers idle through every run + fetch and through the ramp-up and tail of every segment.)
    std::vector<const DecodeTask*> all; std::vector<size_t> seg_end;
    for (const Segment& sg : part) { for (const DecodeTask& t : sg.tasks) all.push_back(&t); seg_end.push_back(all.size()); }
    size_t cur = 0;                                                    // segment being filled
    auto finish_segment = [&](const Segment& sg) -> int {
        int r;
        if (first) { if ((r = wait_ready())) return fail(r, gpu_err); first = false; }
        {   // the genome of this segment's targets must be resident; later targets keep uploading under the next segments' decode
            std::unique_lock<std::mutex> lk(gm);
            gcv.wait(lk, [&] { if (genome_rc || genome_finished) return true; for (int32_t t : sg.targets) if (!genome_done[(size_t)t]) return false; return true; });
            if (genome_rc) return fail(genome_rc, genome_err);
        }
        const double tr = now_s();
        if ((r = pj_shard_run(ctx))) return fail(r, pj_last_error(ctx));
        const int64_t J = pj_shard_num_junctions(ctx);
        const size_t base = out.rows.size();
        out.rows.resize(base + (size_t)J);
        std::vector<pj_target_stats> st((size_t)T);
        if ((r = pj_shard_fetch(ctx, out.rows.data() + base, J, st.data(), T))) return fail(r, pj_last_error(ctx));
        merge_stats(out.stats, st.data(), T);                          // a target may be spread over several segments
        out.run_s += now_s() - tr;
        float ms = 0; int32_t nl = 0; pj_shard_timing(ctx, &ms, &nl); out.gpu_ms += ms; out.launches += nl; out.n_segments++;
        return PJ_OK;
    };
    auto open_segment = [&](const Segment& sg) -> int {                // the first shard was opened by the GPU thread
        int r = pj_shard_begin(ctx, (int64_t)sg.weight + 1024, (int64_t)sg.weight * 4 + 1024, (int64_t)sg.weight * 56 + 1024);
        return r ? fail(r, pj_last_error(ctx)) : PJ_OK;
    };
    // segments without a decode task at the front of the part
    auto drain_empty = [&]() -> int {
        while (cur < part.size() && seg_end[cur] == (cur ? seg_end[cur - 1] : 0)) {
            int r;
            if ((r = wait_ready())) return fail(r, gpu_err);
            if (cur > 0 && (r = open_segment(part[cur]))) return r;
            if ((r = finish_segment(part[cur]))) return r;
            cur++;
        }
        return PJ_OK;
    };
    const double td = now_s();
    int r = drain_empty();
    if (r) { wait_ready(); return bail(r, g_err); }
    bool opened = true;                                                // is the shard of segment `cur` open?  (segment 0: by the GPU thread)
    if (cur > 0) opened = false;
    r = ordered_pipeline<Payload>(all.size(), threads, window,
        [&](size_t k, Payload& p) -> int {
            p.chunk = pool.get();
            p.chunk->with_names = extra; p.chunk->lean = lean; p.chunk->keep_mate = keep_mate;
            prep->bam.decode(*all[k], *p.chunk);
            return PJ_OK;
        },
        [&](size_t k, Payload& p) -> int {
            int q = PJ_OK;
            if ((q = wait_ready())) return fail(q, gpu_err);
            if (!opened) { if ((q = open_segment(part[cur]))) return q; opened = true; }
            if (p.chunk->n() > 0) {
                pj_batch st; memset(&st, 0, sizeof st);
                if ((q = stage(*p.chunk, st))) return q;
                if ((q = pj_batch_submit(ctx, &st))) return fail(q, pj_last_error(ctx));
            }
            pool.put(std::move(p.chunk));
            if (k + 1 == seg_end[cur]) {                               // the segment is complete
                if ((q = finish_segment(part[cur]))) return q;
                cur++; opened = false;
                // the following segments that have no task at all
                while (cur < part.size() && seg_end[cur] == seg_end[cur - 1]) {
                    if ((q = open_segment(part[cur]))) return q;
                    if ((q = finish_segment(part[cur]))) return q;
                    cur++;
                }
            }
            return q;
        });
    if (r) { wait_ready(); return bail(r, g_err); }
    out.decode_s = now_s() - td - out.run_s;
    // The context goes back to the caller: the extra metrics need every shard's context, and without them the caller destroys it on a
    // helper thread while the rows are finalized and written (freeing ~20 GB of device and pinned memory takes 0.3 s on a human-scale run).
    out.ctx = ctx; ctx = nullptr;
}


// A12/A13 + writers + strand report over the gathered rows (junction_builder.cc:249-290, junction_system.cc:250-383, 455-560)
int finish_rows(const pjh_options* o, const pjio::BamHeader& H, std::vector<pj_junction>& rows, std::vector<pj_junction_extra>& xrows,
                const std::vector<pj_target_stats>& stats, pjh_report& R) {
    const bool say = !o->quiet; const bool extra = o->extra != 0 && xrows.size() == rows.size() && !rows.empty();
    const int32_t T = (int32_t)H.names.size();
    const std::string prefix = (o->output_prefix && *o->output_prefix) ? o->output_prefix : "portcullis";
    const double tf = now_s();
    uint64_t spliced = 0, unspliced = 0, sumq = 0; int32_t minq = INT32_MAX, maxq = 0;
    if (say) std::cout << " - All shards completed.\n - Combining results.\n\n" << std::left << std::setw(12) << "Sequence" << "\t" << std::right << std::setw(12) << "unspliced"
                       << "\t" << std::setw(12) << "spliced" << "\t" << std::setw(12) << "total" << std::endl;
    for (int32_t t = 0; t < T; t++) {
        spliced += stats[t].spliced_count; unspliced += stats[t].unspliced_count; sumq += stats[t].sum_query_lengths;
        minq = std::min(minq, stats[t].min_query_length); maxq = std::max(maxq, stats[t].max_query_length);
        if (say) std::cout << std::left << std::setw(12) << H.names[t] << "\t" << std::right << std::setw(12) << stats[t].unspliced_count << "\t"
                           << std::setw(12) << stats[t].spliced_count << "\t" << std::setw(12) << stats[t].spliced_count + stats[t].unspliced_count << std::endl;
    }
    const uint64_t total = spliced + unspliced;
    const double mean_q = (double)sumq / (double)total;
    bool sorted = true;
    for (size_t i = 1; i < rows.size() && sorted; i++) {
        const pj_junction &x = rows[i - 1], &y = rows[i];
        sorted = x.tid != y.tid ? x.tid < y.tid : x.start != y.start ? x.start < y.start : x.end <= y.end;
    }
    if (!sorted || extra) {     // bring rows (and their extra columns) into the final (tid, start, end) order together
        std::vector<size_t> ord(rows.size()); for (size_t i = 0; i < ord.size(); i++) ord[i] = i;
        if (!sorted) std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
            const pj_junction &x = rows[a], &y = rows[b];
            return x.tid != y.tid ? x.tid < y.tid : x.start != y.start ? x.start < y.start : x.end < y.end; });
        std::vector<pj_junction> r2(rows.size()); for (size_t i = 0; i < ord.size(); i++) r2[i] = rows[ord[i]];
        rows.swap(r2);
        if (extra) { std::vector<pj_junction_extra> x2(rows.size()); for (size_t i = 0; i < ord.size(); i++) x2[i] = xrows[ord[i]]; xrows.swap(x2); pj_extra_finalize(xrows.data(), (int64_t)xrows.size()); }
    }
    int rc;
    if ((rc = pj_junctions_finalize(rows.data(), (int64_t)rows.size(), mean_q))) return fail(rc, "finalize failed");
    R.t_finalize_s = now_s() - tf;
    if (say) {
        std::cout << "\nFinal stats:\n - Processed " << total << " alignments.\n - Alignment query length statistics: min: " << minq << "; mean: " << mean_q
                  << "; max: " << maxq << ";\n - Found " << rows.size() << " junctions from " << spliced << " spliced alignments.\n - Found " << unspliced
                  << " unspliced alignments.\n\nSaving junctions: " << std::endl;
    }
    const double tw = now_s();
    {
        std::vector<const char*> names((size_t)T); for (int32_t t = 0; t < T; t++) names[t] = H.names[t].c_str();
        rc = pjh_write_outputs_extra(prefix.c_str(), rows.data(), extra ? xrows.data() : nullptr, (int64_t)rows.size(), T, names.data(), H.lens.data(),
                                     o->source ? o->source : "portcullis", o->version ? o->version : "1.2.4", o->exon_gff, o->intron_gff);
        if (rc) return rc;
    }
    R.t_write_s = now_s() - tw;
    if (say) {
        // JunctionSystem::determineStrandedness(true) (junction_system.cc:455-560): report only
        uint32_t c[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
        for (const auto& j : rows) {
            if (j.ss_strand == PJ_STRAND_UNKNOWN) continue;
            uint32_t* k = c[j.ss_strand == PJ_STRAND_POS ? 0 : 1];
            k[0] += j.nb_r1_pos; k[1] += j.nb_r1_neg; k[2] += j.nb_r2_pos; k[3] += j.nb_r2_neg;
        }
        auto ratio = [](uint32_t a, uint32_t b) { return ((double)((int32_t)a - (int32_t)b)) / ((double)(a + b)); };
        const double posr1 = ratio(c[0][0], c[0][1]), negr1 = ratio(c[1][1], c[1][0]), posr2 = ratio(c[0][2], c[0][3]), negr2 = ratio(c[1][3], c[1][2]);
        const uint32_t totalr1 = c[0][0] + c[0][1] + c[1][0] + c[1][1], totalr2 = c[0][2] + c[0][3] + c[1][2] + c[1][3];
        std::cout << "Strand Analysis\n---------------\n\nTotal Alignments:\n - R1:" << totalr1 << "\n - R2:" << totalr2
                  << "\nAlignment counts when splice site suggests +ve strand:\n - R1+: " << c[0][0] << "\n - R1-: " << c[0][1] << "\n - R2+: " << c[0][2] << "\n - R2-: " << c[0][3]
                  << "\nAlignment counts when splice site suggests -ve strand:\n - R1+: " << c[1][0] << "\n - R1-: " << c[1][1] << "\n - R2+: " << c[1][2] << "\n - R2-: " << c[1][3]
                  << "\nCorrelation of read strand to splice site strand (1.0 = complete agreement, -1.0 = complete disagreement):\n - R1+: " << posr1
                  << "\n - R1-: " << negr1 << "\n - R2+: " << posr2 << "\n - R2-: " << negr2 << "\n" << std::endl;
        int s = PJ_STRANDED_UNKNOWN; const char* ori = "Unknown";
        if (totalr1 == 0 && totalr2 == 0) {}
        else if (totalr2 == 0) { ori = "Single-End (SE)"; if (posr1 > 0.5 && negr1 > 0.5) s = PJ_STRANDED_SECONDSTRAND; else if (posr1 < -0.5 && negr1 < -0.5) s = PJ_STRANDED_FIRSTSTRAND; }
        else {
            ori = "Paired-End (FR): Forward Reverse (-> <-)";
            if (posr1 > 0.5 && negr1 > 0.5 && posr2 < -0.5 && negr2 < -0.5) s = PJ_STRANDED_SECONDSTRAND;
            else if (posr1 < -0.5 && negr1 < -0.5 && posr2 > 0.5 && negr2 > 0.5) s = PJ_STRANDED_FIRSTSTRAND;
            else if (posr1 > 0.5 && negr1 > 0.5 && posr2 > 0.5 && negr2 > 0.5) { s = PJ_STRANDED_SECONDSTRAND; ori = "Paired-End (FF): Forward Forward (-> ->)"; }
            else if (posr1 < -0.5 && negr1 < -0.5 && posr2 < -0.5 && negr2 < -0.5) { s = PJ_STRANDED_FIRSTSTRAND; ori = "Paired-End (FF): Forward Forward (-> ->)"; }
        }
        if (std::abs(posr1) <= 0.5 && std::abs(negr1) <= 0.5 && std::abs(posr2) <= 0.5 && std::abs(negr2) <= 0.5) s = PJ_STRANDED_UNSTRANDED;
        static const char* LONG[] = {"Unstranded - can't determine transcript strand from read strand", "Firststrand - R1 is not on transcript strand",
                                     "Secondstrand - R1 is on transcript strand", "Unknown strand protocol"};
        std::cout << "Determined sequence orientation to be: " << ori << "\nDetermined RNAseq strandedness to be: " << LONG[s] << "\n" << std::endl;
        if (o->strandedness != PJ_STRANDED_UNKNOWN && o->strandedness != s)
            std::cerr << "Warning!  User input and portcullis disagree about the strandedness of the dataset\n" << std::endl;
    }
    R.n_junctions = (int64_t)rows.size(); R.n_spliced = spliced; R.n_unspliced = unspliced; R.mean_query_length = mean_q;
    R.min_query_length = minq; R.max_query_length = maxq;
    return PJ_OK;
}

// part < 0: the whole stage in this process (o->n_gpus GPUs, one host thread each).  part >= 0: only part `part` of `n_parts`
// (one process per GPU, e.g. under torchrun): rows and per-target scalars go to `partial`, nothing is finalized or written.
int junc_core(const pjh_options* o, int part, int n_parts_in, pjh_partial* partial, pjh_report* rep) {
    if (!o || !o->prep_dir) return fail(PJ_EINVAL, "pjh_junc_run: no prep directory given");
    pjh_report R; memset(&R, 0, sizeof R);
    const double t0 = now_s();
    const bool rank_mode = part >= 0;
    const bool say = !o->quiet && !rank_mode;
    const bool extra = o->extra != 0;      // junction_builder.cc:113-117 turns --separate on for --extra; here the metrics come from the records in HBM instead
    if (rank_mode && extra) return fail(PJ_EINVAL, "pjh_junc_run_part: the --extra metrics exchange read names between contexts; run them in one process (pjh_junc_run)");
    const std::string prefix = (o->output_prefix && *o->output_prefix) ? o->output_prefix : "portcullis";
    if (!rank_mode) {   // output directory (junction_builder.cc:65-66, 86-91)
        fs::path parent = fs::path(prefix).parent_path();
        std::error_code ec;
        if (!parent.empty() && !fs::exists(parent, ec) && !fs::create_directories(parent, ec))
            return fail(PJ_EIO, "Could not create output directory at: " + parent.string());
    }
    pjh_prep* prep = nullptr;
    int rc = pjh_prep_open(o->prep_dir, o->use_csi, &prep);
    if (rc) return rc;
    std::unique_ptr<pjh_prep> prep_guard(prep);
    const pjio::BamHeader& H = prep->bam.header();
    const int32_t T = (int32_t)H.names.size();
    if (T == 0) return fail(PJ_EDATA, "BAM header declares no target sequences");
    int n_parts = rank_mode ? std::max(1, n_parts_in) : std::max(1, o->n_gpus);
    if (rank_mode && part >= n_parts) return fail(PJ_EINVAL, "pjh_junc_run_part: part out of range");
    int threads = std::max(1, o->threads);
    static const char* ORI[] = {"SE", "FR", "RF", "FF", "UNKNOWN"};
    static const char* STR[] = {"UNSTRANDED", "FIRSTSTRAND", "SECONDSTRAND", "UNKNOWN"};
    if (say) {
        std::cout << "Settings:\n - BAM Strandedness: " << STR[std::min(std::max(o->strandedness, 0), 3)]
                  << "\n - BAM Read Orientation: " << ORI[std::min(std::max(o->orientation, 0), 4)]
                  << "\n - BAM Indexing mode: " << (o->use_csi ? "CSI" : "BAI")
                  << "\n - Host decode threads: " << threads << "\n - GPUs: " << n_parts << "\n - Separate BAMs: " << (o->separate ? "true" : "false") << "\n\n";
    }
    if (o->separate && !rank_mode) {
        // JunctionBuilder::separateBams (junction_builder.cc:152-226): host I/O only, before the junction pass like the reference
        const double ts = now_s();
        pjio::SeparateCounts sc;
        const std::string un = prefix + ".unspliced.bam", sp = prefix + ".spliced.bam", um = prefix + ".unmapped.bam";
        if (say) std::cout << "Splitting BAM:\n - Saving unspliced alignments to: \"" << un << "\"\n - Saving spliced alignments to: \"" << sp
                           << "\"\n - Saving unmapped reads to: \"" << um << "\"\n - Processing BAM ..." << std::flush;
        try { pjio::separate_bams(prep->bam, sp, un, um, o->use_csi != 0, threads, sc); }
        catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
        R.t_separate_s = now_s() - ts;
        if (say) std::cout << " done.\n - Found " << sc.spliced << " spliced alignments.\n - Found " << sc.unspliced << " unspliced alignments.\n - Found "
                           << sc.unmapped << " unmapped reads.\n - Indexed the unspliced and spliced alignments (" << (o->use_csi ? "CSI" : "BAI") << ").\n = Wall time taken: "
                           << std::fixed << std::setprecision(1) << R.t_separate_s << "s\n" << std::defaultfloat << std::setprecision(6) << std::endl;
    }
    // ---- work plan: contiguous record-balanced ranges cut at gaps (whole targets by LPT for --extra) ----
    if (!prep->indexed) n_parts = rank_mode ? n_parts : 1;
    const bool whole_targets = extra || env_u64("PJ_WHOLE_TARGETS", 0) != 0;
    if (whole_targets && !rank_mode && n_parts > T) n_parts = T;
    std::vector<Part> parts; int n_cuts = 0;
    try { if ((rc = plan_parts(prep, n_parts, whole_targets, env_u64("PJ_SEG_RECORDS", 32u << 20), parts, &n_cuts))) return rc; }
    catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    R.t_open_s = now_s() - t0;
    R.n_gpus_used = n_parts;
    if (say) {
        size_t nseg = 0; for (auto& p : parts) nseg += p.size();
        std::cout << "Finding junctions and calculating basic metrics:\n - Sharding " << T << " target sequences over " << n_parts << " GPU(s)";
        if (!whole_targets) std::cout << " in " << nseg << " segment(s), " << n_cuts << " cut(s) inside a target";
        std::cout << std::endl;
    }
    if (rank_mode) {
        GpuOut out;
        run_part(o, prep, parts[(size_t)part], o->gpu_ids ? o->gpu_ids[0] : 0, threads, false, out);
        if (out.rc) { if (out.ctx) pj_destroy(out.ctx); return fail(out.rc, out.err); }
        if (out.ctx) { pj_ctx* c = out.ctx; out.ctx = nullptr; partial->teardown = std::thread([c]() { pj_destroy(c); }); }
        partial->rows.swap(out.rows); partial->stats.swap(out.stats);
        R.t_gpu_ms = out.gpu_ms; R.n_kernel_launches = out.launches; R.t_genome_s = out.genome_s; R.t_decode_s = out.decode_s; R.t_init_s = out.init_s;
        R.t_run_s = out.run_s; R.t_teardown_s = out.teardown_s; R.n_junctions = (int64_t)partial->rows.size(); R.n_segments = out.n_segments; R.n_gap_cuts = n_cuts;
        for (auto& st : partial->stats) { R.n_spliced += st.spliced_count; R.n_unspliced += st.unspliced_count; }
        R.t_total_s = now_s() - t0;
        partial->rep = R;
        if (rep) *rep = R;
        return PJ_OK;
    }
    std::vector<GpuOut> outs((size_t)n_parts);
    const int threads_per_gpu = std::max(1, threads / n_parts);
    {
        std::vector<std::thread> th;
        auto one = [&](int g) { run_part(o, prep, parts[(size_t)g], o->gpu_ids ? o->gpu_ids[g] : g, threads_per_gpu, extra, outs[(size_t)g]); };
        for (int g = 1; g < n_parts; g++) th.emplace_back(one, g);
        one(0);
        for (auto& t : th) t.join();
    }
    const int n_gpus = n_parts;
    struct CtxGuard { std::vector<GpuOut>& o; ~CtxGuard() { for (auto& x : o) if (x.ctx) { pj_destroy(x.ctx); x.ctx = nullptr; } } } ctx_guard{outs};
    for (auto& out : outs) if (out.rc) return fail(out.rc, out.err);
    struct Teardown { std::vector<std::thread> th; void join() { for (auto& t : th) if (t.joinable()) t.join(); th.clear(); } ~Teardown() { join(); } } teardown;
    if (!extra) for (auto& x : outs) if (x.ctx) { pj_ctx* c = x.ctx; x.ctx = nullptr; teardown.th.emplace_back([c]() { pj_destroy(c); }); }
    if (extra) {
        // ---- calcExtraMetrics (junction_builder.cc:293-312) over the records resident on the GPUs ----
        const double tx = now_s();
        if (say) std::cout << "Calculating extra junction metrics:" << std::endl;
        std::vector<int32_t> owner((size_t)T, 0);
        for (int g = 0; g < n_gpus; g++) for (auto& sg : parts[(size_t)g]) for (int32_t t : sg.targets) owner[t] = g;
        int32_t maxq_all = 0;
        for (int g = 0; g < n_gpus; g++) for (auto& st : outs[g].stats) maxq_all = std::max(maxq_all, st.max_query_length);
        for (int g = 0; g < n_gpus; g++) if (!outs[g].ctx) return fail(PJ_ESTATE, "--extra: a GPU has no records to work on; use fewer GPUs");
        // spliced read names of the whole file on every GPU (the reference's map spans the BAM, junction_builder.cc:179-186)
        if (n_gpus > 1) {
            std::vector<std::vector<uint64_t>> names((size_t)n_gpus);
            for (int g = 0; g < n_gpus; g++) {
                names[g].resize((size_t)std::max<int64_t>(pj_extra_num_spliced_names(outs[g].ctx), 0));
                if ((rc = pj_extra_export_names(outs[g].ctx, names[g].data(), (int64_t)names[g].size()))) return fail(rc, pj_last_error(outs[g].ctx));
            }
            for (int g = 0; g < n_gpus; g++) for (int h = 0; h < n_gpus; h++)
                if (h != g && (rc = pj_extra_import_names(outs[g].ctx, names[h].data(), (int64_t)names[h].size()))) return fail(rc, pj_last_error(outs[g].ctx));
        }
        {
            std::vector<std::thread> th; std::vector<int> xr((size_t)n_gpus, PJ_OK);
            auto one = [&](int g) { outs[g].extra.resize(outs[g].rows.size()); xr[g] = pj_extra_run(outs[g].ctx, maxq_all, outs[g].extra.data(), (int64_t)outs[g].extra.size()); };
            for (int g = 1; g < n_gpus; g++) th.emplace_back(one, g);
            one(0);
            for (auto& t : th) t.join();
            for (int g = 0; g < n_gpus; g++) if (xr[g]) return fail(xr[g], pj_last_error(outs[g].ctx));
        }
        // coverage: depth vector of the previous covered target (Q14), wherever that target lives
        std::vector<uint8_t> covered((size_t)T, 0); std::vector<int32_t> src((size_t)T, -1);
        for (int32_t t = 0; t < T; t++) {
            int32_t cv = 0; uint32_t live = 0;
            if ((rc = pj_extra_target_pileup(outs[owner[t]].ctx, t, &cv, &live))) return fail(rc, pj_last_error(outs[owner[t]].ctx));
            covered[t] = (uint8_t)cv;
            if (live >= 8000 && o->verbose) std::cerr << " - " << H.names[t] << ": up to " << live << " unspliced alignments on one position; htslib's 8000-read pileup cap is replayed there\n";
        }
        pj_extra_coverage_source(T, covered.data(), src.data());
        {   // one batched query per GPU that owns depth vectors (not one launch + sync per target: a 3 000-target genome went through the host)
            struct Q { std::vector<int32_t> dt, st, en; std::vector<std::pair<int, size_t>> where; };
            std::vector<Q> q((size_t)n_gpus);
            for (int g = 0; g < n_gpus; g++) {
                auto& rws = outs[g].rows;
                for (size_t k = 0; k < rws.size(); k++) {
                    const int32_t d = src[rws[k].tid];
                    if (d < 0) continue;
                    Q& x = q[(size_t)owner[d]];
                    x.dt.push_back(d); x.st.push_back(rws[k].start); x.en.push_back(rws[k].end); x.where.emplace_back(g, k);
                }
            }
            for (int g = 0; g < n_gpus; g++) {
                Q& x = q[(size_t)g];
                if (x.dt.empty()) continue;
                std::vector<uint32_t> sums(x.dt.size() * 4);
                if ((rc = pj_extra_coverage_batch(outs[g].ctx, (int64_t)x.dt.size(), x.dt.data(), x.st.data(), x.en.data(), sums.data()))) return fail(rc, pj_last_error(outs[g].ctx));
                for (size_t k = 0; k < x.where.size(); k++) memcpy(outs[(size_t)x.where[k].first].extra[x.where[k].second].cov_sum, &sums[k * 4], 16);
            }
        }
        for (auto& x : outs) if (x.ctx) { pj_destroy(x.ctx); x.ctx = nullptr; }
        R.t_extra_s = now_s() - tx;
    }
    // ---- gather (junction_builder.cc:249-283) ----
    std::vector<pj_junction> rows; std::vector<pj_junction_extra> xrows;
    std::vector<pj_target_stats> stats((size_t)T, pj_target_stats{0, 0, 0, INT32_MAX, 0});
    for (int g = 0; g < n_gpus; g++) {
        if (n_gpus == 1) { rows.swap(outs[g].rows); xrows.swap(outs[g].extra); }       // one part: its rows are the table, no second copy
        else {
            rows.insert(rows.end(), outs[g].rows.begin(), outs[g].rows.end());
            xrows.insert(xrows.end(), outs[g].extra.begin(), outs[g].extra.end());
        }
        merge_stats(stats, outs[g].stats.data(), T);
        R.t_gpu_ms = std::max<double>(R.t_gpu_ms, outs[g].gpu_ms); R.n_kernel_launches += outs[g].launches; R.n_segments += outs[g].n_segments;
        R.t_genome_s = std::max(R.t_genome_s, outs[g].genome_s); R.t_decode_s = std::max(R.t_decode_s, outs[g].decode_s);
        R.t_init_s = std::max(R.t_init_s, outs[g].init_s); R.t_run_s = std::max(R.t_run_s, outs[g].run_s); R.t_teardown_s = std::max(R.t_teardown_s, outs[g].teardown_s);
    }
    R.n_gap_cuts = n_cuts;
    if ((rc = finish_rows(o, H, rows, xrows, stats, R))) return rc;
    { const double tj = now_s(); teardown.join(); R.t_teardown_s += now_s() - tj; }     // what is left of the contexts' teardown after the writers
    R.t_total_s = now_s() - t0;
    if (rep) *rep = R;
    return PJ_OK;
}

} // namespace

int pjh_junc_run(const pjh_options* o, pjh_report* rep) { return junc_core(o, -1, 0, nullptr, rep); }

int pjh_plan_describe(const pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t* segments_per_part, int32_t* n_gap_cuts) {
    if (!p || n_parts < 1 || !segments_per_part) return fail(PJ_EINVAL, "pjh_plan_describe: bad argument");
    std::vector<Part> parts; int cuts = 0;
    try { int rc = plan_parts(p, n_parts, whole_targets != 0, seg_records > 0 ? (uint64_t)seg_records : (32u << 20), parts, &cuts); if (rc) return rc; }
    catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    for (int32_t g = 0; g < n_parts; g++) segments_per_part[g] = (int32_t)parts[(size_t)g].size();
    if (n_gap_cuts) *n_gap_cuts = cuts;
    return PJ_OK;
}

int pjh_plan_decode(pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t part, int32_t segment, int32_t threads, pj_batch* out) {
    if (!p || n_parts < 1 || part < 0 || part >= n_parts || !out) return fail(PJ_EINVAL, "pjh_plan_decode: bad argument");
    std::vector<Part> parts;
    try { int rc = plan_parts(p, n_parts, whole_targets != 0, seg_records > 0 ? (uint64_t)seg_records : (32u << 20), parts, nullptr); if (rc) return rc; }
    catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    if (segment < 0 || (size_t)segment >= parts[(size_t)part].size()) return fail(PJ_EINVAL, "pjh_plan_decode: segment out of range");
    const Segment& sg = parts[(size_t)part][(size_t)segment];
    p->decoded.clear(); p->decoded.with_names = p->want_names;
    int rc = ordered_pipeline<ColumnarChunk>(sg.tasks.size(), std::max(1, threads), (size_t)std::max(1, threads) * 3 + 2,
        [&](size_t k, ColumnarChunk& c) { c.with_names = p->want_names; p->bam.decode(sg.tasks[k], c); return PJ_OK; },
        [&](size_t, ColumnarChunk& c) { p->decoded.append(c); return PJ_OK; });
    if (rc) return rc;
    chunk_view(p->decoded, out);
    return PJ_OK;
}

int pjh_plan_decode_lean(pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t part, int32_t segment, int32_t threads,
                         int32_t keep_mate, pj_batch* out, const pjh_lean_run** runs, int32_t* n_runs) {
    if (!p || n_parts < 1 || part < 0 || part >= n_parts || !out || !runs || !n_runs) return fail(PJ_EINVAL, "pjh_plan_decode_lean: bad argument");
    if (!p->indexed) return fail(PJ_EINVAL, "pjh_plan_decode_lean: the lean form needs an indexed BAM (one target per decode task)");
    std::vector<Part> parts;
    try { int rc = plan_parts(p, n_parts, whole_targets != 0, seg_records > 0 ? (uint64_t)seg_records : (32u << 20), parts, nullptr); if (rc) return rc; }
    catch (const std::exception& e) { return fail(PJ_EIO, e.what()); }
    if (segment < 0 || (size_t)segment >= parts[(size_t)part].size()) return fail(PJ_EINVAL, "pjh_plan_decode_lean: segment out of range");
    const Segment& sg = parts[(size_t)part][(size_t)segment];
    p->decoded.clear(); p->decoded.with_names = p->want_names; p->decoded.lean = true; p->decoded.keep_mate = keep_mate != 0;
    int rc = ordered_pipeline<ColumnarChunk>(sg.tasks.size(), std::max(1, threads), (size_t)std::max(1, threads) * 3 + 2,
        [&](size_t k, ColumnarChunk& c) { c.with_names = p->want_names; c.lean = true; c.keep_mate = keep_mate != 0; p->bam.decode(sg.tasks[k], c); return PJ_OK; },
        [&](size_t, ColumnarChunk& c) { p->decoded.append(c); return PJ_OK; });
    p->decoded.lean = false;                                        // later classic decodes reuse the object
    if (rc) return rc;
    const ColumnarChunk& c = p->decoded;
    memset(out, 0, sizeof *out);
    out->n_records = c.n(); out->pos = c.pos.data(); out->flag = c.flag.data(); out->mapq = c.mapq.data(); out->xs = c.xs.data(); out->l_qseq = c.l_qseq.data();
    out->mtid = keep_mate ? c.mtid.data() : nullptr; out->mpos = keep_mate ? c.mpos.data() : nullptr;
    out->cigar = c.cigar.data(); out->lean = 1; out->const_tid = -1; out->n_cigar = c.n_cigar.data(); out->n_cigar_total = (int64_t)c.cigar.size();
    out->seq2 = c.seq2.data(); out->n_seq2_bytes = (int64_t)c.seq2.size(); out->seqx_pos = c.seqx_pos.data(); out->seqx_code = c.seqx_code.data(); out->n_seqx = (int64_t)c.seqx_pos.size();
    out->name_code = (c.with_names && (int64_t)c.name_code.size() == c.n()) ? c.name_code.data() : nullptr;
    static_assert(sizeof(pjh_lean_run) == sizeof(ColumnarChunk::Run), "pjh_lean_run mirrors ColumnarChunk::Run");
    *runs = reinterpret_cast<const pjh_lean_run*>(c.runs.data()); *n_runs = (int32_t)c.runs.size();
    return PJ_OK;
}

int pjh_junc_run_part(const pjh_options* o, int32_t part, int32_t n_parts, pjh_partial** out, pjh_report* rep) {
    if (!out || part < 0 || n_parts < 1) return fail(PJ_EINVAL, "pjh_junc_run_part: bad argument");
    *out = nullptr;
    auto p = std::make_unique<pjh_partial>();
    int rc = junc_core(o, part, n_parts, p.get(), rep);
    if (rc) return rc;
    *out = p.release();
    return PJ_OK;
}
int64_t pjh_partial_rows(const pjh_partial* p, const pj_junction** rows) { if (!p) return -1; if (rows) *rows = p->rows.data(); return (int64_t)p->rows.size(); }
int32_t pjh_partial_stats(const pjh_partial* p, const pj_target_stats** stats) { if (!p) return -1; if (stats) *stats = p->stats.data(); return (int32_t)p->stats.size(); }
void pjh_partial_free(pjh_partial* p) { delete p; }

int pjh_junc_finish(const pjh_options* o, pj_junction* rows_in, int64_t n_rows, const pj_target_stats* stats_in, int32_t n_targets, pjh_report* rep) {
    if (!o || !o->prep_dir || (n_rows && !rows_in) || !stats_in) return fail(PJ_EINVAL, "pjh_junc_finish: bad argument");
    const double t0 = now_s();
    pjh_prep* prep = nullptr;
    int rc = pjh_prep_open(o->prep_dir, o->use_csi, &prep);
    if (rc) return rc;
    std::unique_ptr<pjh_prep> guard(prep);
    const pjio::BamHeader& H = prep->bam.header();
    if ((int32_t)H.names.size() != n_targets) return fail(PJ_EINVAL, "pjh_junc_finish: stats must have one entry per target of the BAM header");
    {
        const std::string prefix = (o->output_prefix && *o->output_prefix) ? o->output_prefix : "portcullis";
        fs::path parent = fs::path(prefix).parent_path(); std::error_code ec;
        if (!parent.empty() && !fs::exists(parent, ec) && !fs::create_directories(parent, ec)) return fail(PJ_EIO, "Could not create output directory at: " + parent.string());
    }
    std::vector<pj_junction> rows(rows_in, rows_in + n_rows); std::vector<pj_junction_extra> xrows;
    std::vector<pj_target_stats> stats(stats_in, stats_in + n_targets);
    pjh_report R; memset(&R, 0, sizeof R);
    if ((rc = finish_rows(o, H, rows, xrows, stats, R))) return rc;
    if (n_rows) memcpy(rows_in, rows.data(), (size_t)n_rows * sizeof(pj_junction));
    R.t_total_s = now_s() - t0;
    if (rep) *rep = R;
    return PJ_OK;
}

// ------------------------------------------------------------------------------------------------
// `portcullis junc` command line (JunctionBuilder::main, junction_builder.cc:359-454)
// ------------------------------------------------------------------------------------------------
static void junc_usage() {
    std::cout << "Portcullis Junction Builder Mode Help\n\n"
                 "Analyses all potential junctions found in the input BAM file.\n"
                 "Run \"portcullis prep ...\" to generate data suitable for junction finding\n"
                 "before running \"portcullis junc ...\"\n\n"
                 "Usage: portcullis junc [options] <prep_data_dir>\n"
                 "System options:\n"
                 "  -t [ --threads ] arg (=1)         The number of host threads used to decode the BAM file.\n"
                 "  --gpus arg (=1)                   The number of GPUs to shard target sequences over.\n"
                 "  --separate                        Separate spliced from unspliced reads.\n"
                 "  --extra                           Calculate the additional metrics mm_score, coverage, up_aln and down_aln (from the\n"
                 "                                    records on the GPU; no separated BAM files are written).\n"
                 "  --orientation arg (=UNKNOWN)      The orientation of the reads that produced the BAM alignments: \"SE\", \"FR\", \"RF\", \"FF\", \"UNKNOWN\".\n"
                 "  --strandedness arg (=UNKNOWN)     \"unstranded\", \"firststrand\", \"secondstrand\" or \"UNKNOWN\".\n"
                 "  -c [ --use_csi ]                  Whether to use CSI indexing rather than BAI indexing.\n"
                 "  -v [ --verbose ]                  Print extra information\n"
                 "  --help                            Produce help message\n\n"
                 "Output options:\n"
                 "  -o [ --output ] arg (=portcullis_junc/portcullis)\n"
                 "                                    Output prefix for files generated by this program.\n"
                 "  --exon_gff                        Output exon-based junctions in GFF format.\n"
                 "  --intron_gff                      Output intron-based junctions in GFF format.\n"
                 "  --source arg (=portcullis)        The value to enter into the \"source\" field in GFF files.\n" << std::endl;
}

static bool ieq(const std::string& a, const char* b) {
    if (a.size() != strlen(b)) return false;
    for (size_t i = 0; i < a.size(); i++) if (tolower((unsigned char)a[i]) != tolower((unsigned char)b[i])) return false;
    return true;
}

int pjh_junc_main(int argc, char** argv) {
    pjh_options o; pjh_options_default(&o);
    std::string prep, output = "portcullis_junc/portcullis", source = "portcullis", orient = "UNKNOWN", strand = "UNKNOWN";
    bool help = false;
    auto need = [&](int& i, const std::string& name) -> const char* {
        if (i + 1 >= argc) { std::cerr << "Error: the required argument for option '--" << name << "' is missing" << std::endl; return nullptr; }
        return argv[++i];
    };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i]; std::string val; bool has_val = false;
        if (a.rfind("--", 0) == 0) { size_t eq = a.find('='); if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; } }
        auto value = [&](const std::string& name) -> const char* { if (has_val) return val.c_str(); return need(i, name); };
        const char* v = nullptr;
        if (a == "-t" || a == "--threads") { if (!(v = value("threads"))) return 1; o.threads = atoi(v); }
        else if (a == "--gpus") { if (!(v = value("gpus"))) return 1; o.n_gpus = atoi(v); }
        else if (a == "--separate") o.separate = 1;
        else if (a == "--extra") o.extra = 1;
        else if (a == "--orientation") { if (!(v = value("orientation"))) return 1; orient = v; }
        else if (a == "--strandedness") { if (!(v = value("strandedness"))) return 1; strand = v; }
        else if (a == "-c" || a == "--use_csi") o.use_csi = 1;
        else if (a == "-v" || a == "--verbose") o.verbose = 1;
        else if (a == "--help") help = true;
        else if (a == "-o" || a == "--output") { if (!(v = value("output"))) return 1; output = v; }
        else if (a == "--exon_gff") o.exon_gff = 1;
        else if (a == "--intron_gff") o.intron_gff = 1;
        else if (a == "--source") { if (!(v = value("source"))) return 1; source = v; }
        else if (a == "-i" || a == "--prep_data_dir") { if (!(v = value("prep_data_dir"))) return 1; prep = v; }
        else if (!a.empty() && a[0] == '-' && a.size() > 1) { std::cerr << "Error: unrecognised option '" << a << "'" << std::endl; return 1; }
        else { if (!prep.empty()) { std::cerr << "Error: too many positional options have been specified on the command line" << std::endl; return 1; } prep = a; }
    }
    if (help || argc <= 1) { junc_usage(); return 1; }
    if (ieq(orient, "SE")) o.orientation = PJ_ORIENT_SE; else if (ieq(orient, "FR")) o.orientation = PJ_ORIENT_FR; else if (ieq(orient, "RF")) o.orientation = PJ_ORIENT_RF;
    else if (ieq(orient, "FF")) o.orientation = PJ_ORIENT_FF; else if (ieq(orient, "UNKNOWN")) o.orientation = PJ_ORIENT_UNKNOWN;
    else { std::cerr << "Error: Unknown orientation: " << orient << std::endl; return 4; }
    if (ieq(strand, "UNSTRANDED")) o.strandedness = PJ_STRANDED_UNSTRANDED; else if (ieq(strand, "FIRSTSTRAND")) o.strandedness = PJ_STRANDED_FIRSTSTRAND;
    else if (ieq(strand, "SECONDSTRAND")) o.strandedness = PJ_STRANDED_SECONDSTRAND; else if (ieq(strand, "UNKNOWN")) o.strandedness = PJ_STRANDED_UNKNOWN;
    else { std::cerr << "Error: Unknown strandedness: " << strand << std::endl; return 4; }
    o.prep_dir = prep.c_str(); o.output_prefix = output.c_str(); o.source = source.c_str();
    std::cout << "Running portcullis in junction builder mode\n------------------------------------------\n" << std::endl;
    pjh_report rep;
    // CUDA start-up initialises every visible device (seconds on an 8-GPU box).  The command-line tool only ever uses
    // devices 0..gpus-1, so unless the user chose the devices already, hide the rest before the first CUDA call.
    if (!getenv("CUDA_VISIBLE_DEVICES") && o.n_gpus >= 1) {
        std::string vis; for (int g = 0; g < o.n_gpus; g++) vis += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 0);
    }
    const int rc = pjh_junc_run(&o, &rep);
    if (rc) { std::cerr << "Error: " << pjh_last_error() << std::endl; return rc == PJ_EINVAL ? 1 : 4; }
    std::cout << std::fixed << std::setprecision(1) << "\nPortcullis junc completed.\nTotal runtime: " << rep.t_total_s << "s"
              << "  (open " << rep.t_open_s << "s, GPU init " << rep.t_init_s << "s, genome " << rep.t_genome_s << "s [overlapped], decode+H2D " << rep.t_decode_s << "s, run+fetch "
              << rep.t_run_s << "s, teardown " << rep.t_teardown_s << "s, GPU pipeline "
              << std::setprecision(3) << rep.t_gpu_ms << " ms on " << rep.n_gpus_used << " GPU(s), " << std::setprecision(1) << "finalize " << rep.t_finalize_s << "s, write " << rep.t_write_s << "s)\n" << std::endl;
    return 0;
}

} // extern "C"
