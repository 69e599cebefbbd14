// pj_api.cu — the extern "C" layer of include/portcullis_junc.h: context, genome residency, double-buffered
// pinned staging, shard arena in HBM and the kernel pipeline driver.  One context = one B200.
#include "pj_ctx.hpp"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <climits>
#include <mutex>
#include <thread>
#include <chrono>

using namespace pjk;

namespace {
std::mutex g_trim_mu; std::thread g_trim_thread;          // background release of the memory pool of a destroyed context
struct TrimJoin { ~TrimJoin() { std::lock_guard<std::mutex> lk(g_trim_mu); if (g_trim_thread.joinable()) g_trim_thread.join(); } } g_trim_join;
}

namespace pjapi {

thread_local std::string g_last_error;

int fail(pj_ctx* c, int code, const char* fmt, ...) {
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf;
    g_last_error = buf;
    return code;
}

} // namespace pjapi

using namespace pjapi;

namespace {

void free_slot(StagingSlot& s) {
    if (s.block) cudaFreeHost(s.block);
    s.block = nullptr; s.block_bytes = 0;
    s.tid = s.pos = s.l_qseq = s.mtid = s.mpos = nullptr; s.flag = nullptr; s.mapq = s.xs = s.seq4 = nullptr;
    s.cigar_off = s.cigar = nullptr; s.seq_off = nullptr; s.name_code = nullptr; s.cap_rec = s.cap_cig = s.cap_seq = s.cap_seqx = 0;
    s.n_cigar = nullptr; s.seq2 = nullptr; s.seqx_pos = nullptr; s.seqx_code = nullptr;
}

// One pinned block carved into the columns of a batch.  Classic slots hold the 13 columns of the original pj_batch; lean
// slots hold what a lean batch ships (no tid / cigar_off / seq_off, n_cigar and seq2 + exceptions instead).
int alloc_slot(pj_ctx* c, StagingSlot& s, bool lean, int64_t cr, int64_t cc, int64_t cs, int64_t cx) {
    if (s.lean == lean && cr <= s.cap_rec && cc <= s.cap_cig && cs <= s.cap_seq && cx <= s.cap_seqx && s.block) return PJ_OK;
    // Growing a slot is a cudaMallocHost (tens of ms, on the submitting thread): start with room for a typical 4 MB decode task and grow
    // by half.  (A full-size c3 run re-allocated its four slots 24 times = 0.7 s before this.)
    // Every column has a floor, the exception list included (its size differs from batch to batch, and a slot that outgrew it by a few
    // entries was re-allocated as a whole: 15 allocations = 0.85 s on a full-size c3 run).  What one slot had to grow to is the floor of the
    // context's other slots.
    if (s.lean == lean) { cr = std::max(cr + cr / 2, s.cap_rec); cc = std::max(cc + cc / 2, s.cap_cig); cs = std::max(cs + cs / 2, s.cap_seq); cx = std::max(cx * 2, s.cap_seqx); }
    cr = std::max<int64_t>(cr, 384 << 10); cc = std::max<int64_t>(cc, 3 << 19); cs = std::max<int64_t>(cs, lean ? (12 << 20) : (24 << 20)); cx = std::max<int64_t>(cx, 512 << 10);   // ~34 MB per lean slot: room for the 2-3x larger-than-average decode tasks of a dense region
    int64_t* hw = c->slot_floor[lean ? 1 : 0];
    cr = hw[0] = std::max(cr, hw[0]); cc = hw[1] = std::max(cc, hw[1]); cs = hw[2] = std::max(cs, hw[2]); cx = hw[3] = std::max(cx, hw[3]);
    free_slot(s);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t r = (size_t)cr;
    size_t sz[16]; int nsz = 0;
    if (!lean) {
        const size_t q[13] = {up(r * 4), up(r * 4), up(r * 4), up(r * 4), up(r * 4), up(r * 2), up(r), up(r),
                              up((r + 1) * 4), up((size_t)cc * 4), up((r + 1) * 8), up((size_t)cs + 16), c->extra ? up(r * 8) : 0};
        for (size_t v : q) sz[nsz++] = v;
    } else {
        const size_t q[12] = {up(r * 4), up(r * 4), up(r * 4), up(r * 4), up(r * 2), up(r), up(r), up(r * 2),
                              up((size_t)cc * 4), up((size_t)cs + 16), up((size_t)cx * 8 + 8) + up((size_t)cx + 8), c->extra ? up(r * 8) : 0};
        for (size_t v : q) sz[nsz++] = v;
    }
    size_t total = 0; for (int k = 0; k < nsz; k++) total += sz[k];
    {
        const auto t0 = std::chrono::steady_clock::now();
        CU(c, cudaMallocHost((void**)&s.block, total));
        c->t_pinned_alloc_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); c->pinned_alloc_bytes += total; c->n_pinned_allocs++;
    }
    s.block_bytes = total; s.lean = lean;
    uint8_t* p = s.block; size_t k = 0;
    if (!lean) {
        s.tid = (int32_t*)p; p += sz[k++]; s.pos = (int32_t*)p; p += sz[k++]; s.l_qseq = (int32_t*)p; p += sz[k++]; s.mtid = (int32_t*)p; p += sz[k++];
        s.mpos = (int32_t*)p; p += sz[k++]; s.flag = (uint16_t*)p; p += sz[k++]; s.mapq = p; p += sz[k++]; s.xs = p; p += sz[k++];
        s.cigar_off = (uint32_t*)p; p += sz[k++]; s.cigar = (uint32_t*)p; p += sz[k++]; s.seq_off = (uint64_t*)p; p += sz[k++]; s.seq4 = p; p += sz[k++];
        s.name_code = c->extra ? (uint64_t*)p : nullptr;
        s.cigar_off[0] = 0; s.seq_off[0] = 0;
    } else {
        s.pos = (int32_t*)p; p += sz[k++]; s.l_qseq = (int32_t*)p; p += sz[k++]; s.mtid = (int32_t*)p; p += sz[k++]; s.mpos = (int32_t*)p; p += sz[k++];
        s.flag = (uint16_t*)p; p += sz[k++]; s.mapq = p; p += sz[k++]; s.xs = p; p += sz[k++]; s.n_cigar = (uint16_t*)p; p += sz[k++];
        s.cigar = (uint32_t*)p; p += sz[k++]; s.seq2 = p; p += sz[k++];
        s.seqx_pos = (uint64_t*)p; s.seqx_code = p + up((size_t)cx * 8 + 8); p += sz[k++];
        s.name_code = c->extra ? (uint64_t*)p : nullptr;
    }
    s.cap_rec = cr; s.cap_cig = cc; s.cap_seq = cs; s.cap_seqx = cx;
    return PJ_OK;
}

int bit_length(uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

void mark(pj_ctx* c, const char* name) {
    if (c->n_stage >= c->stages.size()) { StageTime s{name, nullptr}; cudaEventCreate(&s.ev); c->stages.push_back(s); }
    c->stages[c->n_stage].name = name;
    cudaEventRecord(c->stages[c->n_stage].ev, c->compute_stream);
    c->n_stage++;
}

} // namespace

// sort + upload the "other exception byte" table after genome uploads
int pjapi::finish_genome(pj_ctx* c) {
    std::lock_guard<std::mutex> glk(c->genome_mu);
    if (!c->genome_dirty) return PJ_OK;
    CU(c, cudaStreamSynchronize(c->genome_stream));
    uint32_t cnt2[2] = {0, 0};
    CU(c, cudaMemcpy(cnt2, c->d_exc_count, sizeof cnt2, cudaMemcpyDeviceToHost));
    const uint32_t cnt = cnt2[0];
    c->any_gx = cnt2[1] ? 1 : 0;
    if (cnt > c->exc_cap)
        return fail(c, PJ_EDATA, "genome holds %u bytes outside ACGTN after upper-casing (limit %u): not a nucleotide FASTA?", cnt, c->exc_cap);
    std::vector<uint64_t> pos(cnt); std::vector<uint8_t> byt(cnt);
    if (cnt) {
        CU(c, cudaMemcpy(pos.data(), c->d_exc_pos, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        CU(c, cudaMemcpy(byt.data(), c->d_exc_byte, cnt, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> ord(cnt); for (uint32_t i = 0; i < cnt; i++) ord[i] = i;
        std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return pos[a] < pos[b]; });
        std::vector<uint64_t> p2(cnt); std::vector<uint8_t> b2(cnt);
        for (uint32_t i = 0; i < cnt; i++) { p2[i] = pos[ord[i]]; b2[i] = byt[ord[i]]; }
        CU(c, cudaMemcpy(c->d_exc_pos, p2.data(), cnt * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CU(c, cudaMemcpy(c->d_exc_byte, b2.data(), cnt, cudaMemcpyHostToDevice));
        c->n_exc_x = (int32_t)std::count(b2.begin(), b2.end(), (uint8_t)'X');
    } else c->n_exc_x = 0;
    c->n_exc = (int32_t)cnt;
    c->genome_dirty = false;
    return PJ_OK;
}

namespace {
// Stream-ordered temporaries of one pj_shard_run: freed (cudaFreeAsync on the same stream) on EVERY exit path, also the early error
// returns — the pool's release threshold is "never", so a leaked temporary would stay allocated for the life of the process.
// the temporaries of one pj_shard_run: handed out by the context's arena, all of them given back when the run ends (also on an error return)
struct StreamTemps {
    pjapi::TempArena& A; size_t high = 0;
    explicit StreamTemps(pjapi::TempArena& a) : A(a) { A.reset(); }
    template <typename T> cudaError_t alloc(T** p, size_t bytes) { void* q = nullptr; const cudaError_t e = A.alloc(&q, bytes); if (e == cudaSuccess) *p = (T*)q; high = std::max(high, A.need); return e; }
    void release(void* p, size_t bytes) { A.release(p, bytes); }
    ~StreamTemps() { A.reset(); }
};
} // namespace

extern "C" {

static_assert(sizeof(pj_junction) == 256, "pj_junction layout changed: update the bindings");
int pj_abi_version(void) { return PJ_ABI_VERSION; }
int pj_junction_size(void) { return (int)sizeof(pj_junction); }
const char* pj_global_last_error(void) { return g_last_error.c_str(); }
const char* pj_last_error(const pj_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int pj_create(const pj_config* cfg, pj_ctx** out) {
    if (!cfg || !out) return fail(nullptr, PJ_EINVAL, "pj_create: null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PJ_ECUDA, "pj_create: no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, PJ_EINVAL, "pj_create: device %d out of range (0..%d)", cfg->device, ndev - 1);
    if (cfg->orientation < PJ_ORIENT_SE || cfg->orientation > PJ_ORIENT_UNKNOWN) return fail(nullptr, PJ_EINVAL, "pj_create: bad orientation %d", cfg->orientation);
    { std::lock_guard<std::mutex> lk(g_trim_mu); if (g_trim_thread.joinable()) g_trim_thread.join(); }
    pj_ctx* c = new pj_ctx();
    c->device = cfg->device; c->orientation = cfg->orientation;
    c->match_group = cfg->reserved[0];                 // 0 = choose from the data; 1..32 forces the lanes-per-pair of k_match (tuning / tests)
    if (const char* e = getenv("PJ_MATCH_GROUP")) c->match_group = atoi(e);
    c->legacy_sort = cfg->reserved[1];                 // 1 = multi-kernel histogram/scan/scatter sort (kept for > 2^30 pairs and as a cross-check)
    if (const char* e = getenv("PJ_LEGACY_SORT")) c->legacy_sort = atoi(e);
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, c->device));
    CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&c->compute_stream, cudaStreamNonBlocking));
    CU(c, cudaEventCreateWithFlags(&c->copies_done, cudaEventDisableTiming));
    CU(c, cudaStreamCreateWithFlags(&c->genome_stream, cudaStreamNonBlocking));
    for (int s = 0; s < 2; s++) CU(c, cudaEventCreateWithFlags(&c->graw_ev[s], cudaEventDisableTiming));
    c->max_slots = cfg->reserved[2] > 0 ? std::max(2, cfg->reserved[2]) : 4;
    c->extra = cfg->extra_metrics != 0;
    CU(c, cudaMalloc(&c->d_scalars, 16 * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->d_shard_acc, 4 * sizeof(unsigned long long)));
    CU(c, cudaMallocHost((void**)&c->h_scalars, 16 * sizeof(uint32_t)));
    // keep stream-ordered allocations cached between shards
    cudaMemPool_t pool; CU(c, cudaDeviceGetDefaultMemPool(&pool, c->device));
    uint64_t thr = UINT64_MAX; CU(c, cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = c;
    return PJ_OK;
}

void pj_destroy(pj_ctx* c) {
    if (!c) return;
    const bool trace = getenv("PJ_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    auto lap = [&](const char* what) { if (trace) { const double t = now(); fprintf(stderr, "[pj_destroy] %-22s %.3f s\n", what, t - t0); t0 = t; } };
    if (c->prewarm_thread.joinable()) c->prewarm_thread.join();
    if (trace) fprintf(stderr, "[pj_destroy] pinned staging: %d allocations, %.1f MB, %.3f s inside cudaMallocHost\n", c->n_pinned_allocs, c->pinned_alloc_bytes / 1e6, c->t_pinned_alloc_s);
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    lap("sync");
    c->tid.free_(); c->pos.free_(); c->l_qseq.free_(); c->mtid.free_(); c->mpos.free_(); c->flag.free_(); c->mapq.free_(); c->xs.free_();
    c->seq2.free_(); c->cigar_off.free_(); c->cigar.free_(); c->seq_off.free_(); c->name_code.free_();
    c->seqx_pos.free_(); c->seqx_code.free_(); c->tmp_seq4.free_(); c->tmp_off4.free_(); c->tmp_xcount.free_(); c->tmp_xoff.free_(); c->tmp_scan.free_(); c->tmp_ncig.free_(); c->tmp_fs.free_();
    c->arena.free_all();
    extra_reset(c);
    lap("arena");
    for (StagingSlot* sl : c->slots) { free_slot(*sl); if (sl->done) cudaEventDestroy(sl->done); delete sl; }
    lap("pinned staging");
    if (c->genome_stream) cudaStreamDestroy(c->genome_stream);
    for (int s = 0; s < 2; s++) { if (c->graw_ev[s]) cudaEventDestroy(c->graw_ev[s]);
                                  if (c->h_graw[s]) cudaFreeHost(c->h_graw[s]); if (c->d_graw[s]) cudaFree(c->d_graw[s]); }
    lap("genome staging");
    cudaFree(c->d_tlen); cudaFree(c->d_toff); cudaFree(c->d_goff); cudaFree(c->d_glen); cudaFree(c->d_g2); cudaFree(c->d_gx); cudaFree(c->d_gsum);
    cudaFree(c->d_exc_pos); cudaFree(c->d_exc_byte); cudaFree(c->d_exc_count);
    cudaFree(c->d_spliced); cudaFree(c->d_unspliced); cudaFree(c->d_sumq); cudaFree(c->d_minq); cudaFree(c->d_maxq);
    cudaFree(c->d_scalars); cudaFree(c->d_shard_acc); cudaFreeHost(c->h_scalars); cudaFree(c->d_rows);
    lap("genome + misc");
    for (auto& s : c->stages) cudaEventDestroy(s.ev);
    if (c->copies_done) cudaEventDestroy(c->copies_done);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->compute_stream) cudaStreamDestroy(c->compute_stream);
    {   // Give the stream-ordered pool back (the `--extra` path allocates from it, and it was told to keep everything while the context lived).  Unmapping tens of GB takes
        // about a second, so it runs on its own thread, under the caller's finalize + writers; the next pj_create / pj_destroy (or
        // the end of the process) waits for it.
        const int dev = c->device;
        std::lock_guard<std::mutex> lk(g_trim_mu);
        if (g_trim_thread.joinable()) g_trim_thread.join();
        g_trim_thread = std::thread([dev]() { cudaSetDevice(dev); cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); });
    }
    lap("streams + pool");
    delete c;
}

// Coordinate order of n records on a GPU (no context needed): order[k] = index of the record that comes k-th.
int pj_coordinate_order(int32_t device, int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, uint32_t* order) {
    if (n < 0 || (n && (!tid || !pos || !flag || !order))) return fail(nullptr, PJ_EINVAL, "pj_coordinate_order: bad arguments");
    if (n >= (1ll << 30)) return fail(nullptr, PJ_EINVAL, "pj_coordinate_order: more than 2^30 records");
    if (n == 0) return PJ_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(nullptr, PJ_ECUDA, "pj_coordinate_order: no CUDA device available; this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, PJ_EINVAL, "pj_coordinate_order: device %d out of range", device);
    pj_ctx* c = nullptr;
    CU(c, cudaSetDevice(device));
    int n_sm = 148; CU(c, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t st; CU(c, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int32_t *d_tid = nullptr, *d_pos = nullptr; uint16_t* d_flag = nullptr; uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr, *scratch = nullptr;
    const size_t N = (size_t)n;
    int rc = PJ_OK;
    auto run = [&]() -> int {
        CU(c, cudaMalloc(&d_tid, N * 4)); CU(c, cudaMalloc(&d_pos, N * 4)); CU(c, cudaMalloc(&d_flag, N * 2));
        CU(c, cudaMalloc(&ka, N * 8)); CU(c, cudaMalloc(&kb, N * 8)); CU(c, cudaMalloc(&va, N * 4)); CU(c, cudaMalloc(&vb, N * 4));
        CU(c, cudaMalloc(&scratch, os_scratch_words((uint32_t)n, 64) * sizeof(uint32_t)));
        CU(c, cudaMemcpyAsync(d_tid, tid, N * 4, cudaMemcpyHostToDevice, st)); CU(c, cudaMemcpyAsync(d_pos, pos, N * 4, cudaMemcpyHostToDevice, st));
        CU(c, cudaMemcpyAsync(d_flag, flag, N * 2, cudaMemcpyHostToDevice, st));
        launch_coord_keys(n, d_tid, d_pos, d_flag, ka, va, st);
        int launches = 0;
        const int which = launch_onesweep_sort(ka, va, kb, vb, (uint32_t)n, 64, scratch, n_sm, st, &launches);
        CU(c, cudaMemcpyAsync(order, which ? vb : va, N * 4, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st)); CU(c, cudaGetLastError());
        return PJ_OK;
    };
    rc = run();
    for (void* p : {(void*)d_tid, (void*)d_pos, (void*)d_flag, (void*)ka, (void*)kb, (void*)va, (void*)vb, (void*)scratch}) cudaFree(p);
    cudaStreamDestroy(st);
    return rc;
}

int pj_targets_set(pj_ctx* c, int32_t n_targets, const int32_t* target_len) {
    if (!c || n_targets <= 0 || !target_len) return fail(c, PJ_EINVAL, "pj_targets_set: bad arguments");
    if (c->n_targets) return fail(c, PJ_ESTATE, "pj_targets_set: targets already set");
    CU(c, cudaSetDevice(c->device));
    c->n_targets = n_targets;
    c->h_tlen.assign(target_len, target_len + n_targets);
    c->h_toff.assign(n_targets + 1, 0); c->h_goff.assign(n_targets, 0); c->h_glen.assign(n_targets, -1);
    uint64_t g = 64;                                   // lead pad: k_match may look one base before a target start
    for (int32_t t = 0; t < n_targets; t++) {
        if (target_len[t] < 0) return fail(c, PJ_EINVAL, "pj_targets_set: negative target length");
        c->h_toff[t + 1] = c->h_toff[t] + (uint64_t)target_len[t];
        c->h_goff[t] = g; g += ((uint64_t)target_len[t] + 63) / 64 * 64 + 64;
    }
    c->g_total_bases = g;
    CU(c, cudaMalloc(&c->d_tlen, n_targets * sizeof(int32_t))); CU(c, cudaMalloc(&c->d_toff, (n_targets + 1) * sizeof(uint64_t)));
    CU(c, cudaMalloc(&c->d_goff, n_targets * sizeof(uint64_t))); CU(c, cudaMalloc(&c->d_glen, n_targets * sizeof(int64_t)));
    CU(c, cudaMemcpy(c->d_tlen, c->h_tlen.data(), n_targets * sizeof(int32_t), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_toff, c->h_toff.data(), (n_targets + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_goff, c->h_goff.data(), n_targets * sizeof(uint64_t), cudaMemcpyHostToDevice));
    CU(c, cudaMemcpy(c->d_glen, c->h_glen.data(), n_targets * sizeof(int64_t), cudaMemcpyHostToDevice));
    CU(c, cudaMalloc(&c->d_g2, g / 32 * sizeof(uint64_t) + 64)); CU(c, cudaMalloc(&c->d_gx, g / 64 * sizeof(uint64_t) + 64));
    CU(c, cudaMemset(c->d_g2, 0, g / 32 * sizeof(uint64_t) + 64)); CU(c, cudaMemset(c->d_gx, 0, g / 64 * sizeof(uint64_t) + 64));
    CU(c, cudaMalloc(&c->d_gsum, (g / 32768 + 2) * sizeof(uint32_t))); CU(c, cudaMemset(c->d_gsum, 0, (g / 32768 + 2) * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->d_exc_pos, c->exc_cap * sizeof(uint64_t))); CU(c, cudaMalloc(&c->d_exc_byte, c->exc_cap));
    CU(c, cudaMalloc(&c->d_exc_count, 2 * sizeof(uint32_t))); CU(c, cudaMemset(c->d_exc_count, 0, 2 * sizeof(uint32_t)));
    CU(c, cudaMalloc(&c->d_spliced, n_targets * 8)); CU(c, cudaMalloc(&c->d_unspliced, n_targets * 8)); CU(c, cudaMalloc(&c->d_sumq, n_targets * 8));
    CU(c, cudaMalloc(&c->d_minq, n_targets * 4)); CU(c, cudaMalloc(&c->d_maxq, n_targets * 4));
    return PJ_OK;
}

int pj_genome_set_target(pj_ctx* c, int32_t tid, const char* bases, int64_t n_bases) {
    if (!c || !c->n_targets) return fail(c, PJ_ESTATE, "pj_genome_set_target: call pj_targets_set first");
    if (tid < 0 || tid >= c->n_targets || n_bases < 0 || (n_bases && !bases)) return fail(c, PJ_EINVAL, "pj_genome_set_target: bad arguments");
    CU(c, cudaSetDevice(c->device));
    const int64_t n = std::min<int64_t>(n_bases, c->h_tlen[tid]);   // windows never reach past the BAM header length
    for (int s = 0; s < 2; s++) if (!c->h_graw[s]) { CU(c, cudaMallocHost((void**)&c->h_graw[s], pj_ctx::GRAW_CHUNK)); CU(c, cudaMalloc(&c->d_graw[s], pj_ctx::GRAW_CHUNK)); }
    int s = 0;
    for (int64_t o = 0; o < n; o += (int64_t)pj_ctx::GRAW_CHUNK, s ^= 1) {
        const int64_t k = std::min<int64_t>((int64_t)pj_ctx::GRAW_CHUNK, n - o);
        CU(c, cudaEventSynchronize(c->graw_ev[s]));                  // slot free again?
        memcpy(c->h_graw[s], bases + o, (size_t)k);
        std::lock_guard<std::mutex> glk(c->genome_mu);               // per chunk: finish_genome (another thread) must not sort the exception table under a pack kernel
        CU(c, cudaMemcpyAsync(c->d_graw[s], c->h_graw[s], (size_t)k, cudaMemcpyHostToDevice, c->genome_stream));
        launch_pack_genome(c->d_graw[s], k, c->h_goff[tid] + (uint64_t)o, c->d_g2, c->d_gx, c->d_gsum, c->d_exc_pos, c->d_exc_byte, c->d_exc_count, c->exc_cap, c->genome_stream);
        CU(c, cudaEventRecord(c->graw_ev[s], c->genome_stream));
    }
    std::lock_guard<std::mutex> glk(c->genome_mu);
    c->h_glen[tid] = n_bases < c->h_tlen[tid] ? n_bases : (int64_t)c->h_tlen[tid];
    CU(c, cudaMemcpyAsync(c->d_glen + tid, &c->h_glen[tid], sizeof(int64_t), cudaMemcpyHostToDevice, c->genome_stream));
    CU(c, cudaGetLastError());
    c->genome_dirty = true;
    return PJ_OK;
}

int pj_shard_begin(pj_ctx* c, int64_t n_records_hint, int64_t n_cigar_hint, int64_t n_seq_bytes_hint) {
    if (!c || !c->n_targets) return fail(c, PJ_ESTATE, "pj_shard_begin: call pj_targets_set first");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->compute_stream));
    extra_reset(c);
    c->n_rec = 0; c->n_cig = 0; c->n_seq = 16; c->have_result = false;   // SEQ stream: 16-byte lead pad (k_match may look back up to 31 bases)
    c->n_junc = 0; c->n_pairs = 0; c->n_seqx = 0;
    cudaStream_t st = c->copy_stream;
    const size_t r = (size_t)std::max<int64_t>(n_records_hint, 1024);
    int rc;
    if ((rc = ensure(c, c->tid, r, 0, st)) || (rc = ensure(c, c->pos, r, 0, st)) || (rc = ensure(c, c->l_qseq, r, 0, st)) ||
        (rc = ensure(c, c->mtid, r, 0, st)) || (rc = ensure(c, c->mpos, r, 0, st)) || (rc = ensure(c, c->flag, r, 0, st)) ||
        (rc = ensure(c, c->mapq, r, 0, st)) || (rc = ensure(c, c->xs, r, 0, st)) || (rc = ensure(c, c->cigar_off, r + 1, 0, st)) ||
        (rc = ensure(c, c->seq_off, r + 1, 0, st)) || (rc = ensure(c, c->cigar, (size_t)std::max<int64_t>(n_cigar_hint, 1024) + 64, 0, st)) ||
        (rc = ensure(c, c->seq2, (size_t)std::max<int64_t>(n_seq_bytes_hint / 2, 1024) + 256, 0, st))) return rc;   // the hint counts 4-bit bytes; the stream holds 2 bits per base
    if (c->extra && (rc = ensure(c, c->name_code, r, 0, st))) return rc;
    CU(c, cudaMemsetAsync(c->cigar_off.p, 0, sizeof(uint32_t), st));
    CU(c, cudaMemsetAsync(c->d_shard_acc, 0, 4 * sizeof(unsigned long long), st));
    { static const uint64_t lead = 16; CU(c, cudaMemcpyAsync(c->seq_off.p, &lead, sizeof(uint64_t), cudaMemcpyHostToDevice, st)); }
    CU(c, cudaMemsetAsync(c->seq2.p, 0, 16, st));                       // the lead pad is read (and masked out) by k_match: keep it defined
    // Size the arena pj_shard_run takes its temporaries from (about 12 B per record and 110 B per read-junction pair) on a helper
    // thread: cudaMalloc of a multi-GB block takes tens of milliseconds, and this way it overlaps the caller's decode instead of
    // sitting in front of the first kernel.  An estimate that turns out too small costs one extra chunk in the first run.
    if (c->prewarm_thread.joinable()) c->prewarm_thread.join();
    {
        const size_t est = (size_t)std::max<int64_t>(n_records_hint, 0) * (12 + 110) + (64u << 20);
        const int dev = c->device;
        if (n_records_hint > (1 << 20) && c->arena.capacity() < est) c->prewarm_thread = std::thread([c, est, dev]() {
            cudaSetDevice(dev);
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || est > (free_b + c->arena.capacity()) / 2) return;
            c->arena.reserve(est);
        });
    }
    c->shard_open = true;
    return PJ_OK;
}

static int staging_acquire(pj_ctx* c, bool lean, int64_t cap_records, int64_t cap_cigar, int64_t cap_seq_bytes, int64_t cap_seqx, pj_batch* out) {
    if (!c || !out || cap_records < 0 || cap_cigar < 0 || cap_seq_bytes < 0 || cap_seqx < 0) return fail(c, PJ_EINVAL, "pj_staging_acquire: bad arguments");
    CU(c, cudaSetDevice(c->device));
    std::lock_guard<std::mutex> lk(c->staging_mu);
    StagingSlot* pick = nullptr;
    for (;;) {
        StagingSlot* oldest = nullptr;
        for (StagingSlot* sl : c->slots) {
            if (sl->state == 2 && cudaEventQuery(sl->done) == cudaSuccess) sl->state = 0;
            if (sl->state == 0 && (!pick || (sl->lean == lean && pick->lean != lean) || (sl->lean == pick->lean && sl->cap_seq > pick->cap_seq))) pick = sl;   // prefer a slot of the right kind, then the roomiest
            if (sl->state == 2 && (!oldest || sl->seq_no < oldest->seq_no)) oldest = sl;
        }
        if (pick) break;
        if ((int)c->slots.size() < c->max_slots) {
            pick = new StagingSlot();
            cudaError_t e = cudaEventCreateWithFlags(&pick->done, cudaEventDisableTiming);
            if (e != cudaSuccess) { delete pick; return fail(c, PJ_ECUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); }
            c->slots.push_back(pick);
            break;
        }
        if (!oldest) return fail(c, PJ_ESTATE, "pj_staging_acquire: all %d staging buffers are handed out and none was submitted", c->max_slots);
        CU(c, cudaEventSynchronize(oldest->done));
        oldest->state = 0;
    }
    int rc = alloc_slot(c, *pick, lean, std::max<int64_t>(cap_records, 1), std::max<int64_t>(cap_cigar, 1), std::max<int64_t>(cap_seq_bytes, 1), std::max<int64_t>(cap_seqx, 1));
    if (rc) return rc;
    pick->state = 1;
    StagingSlot& s = *pick;
    memset(out, 0, sizeof *out);
    out->n_records = 0; out->tid = s.tid; out->pos = s.pos; out->flag = s.flag; out->mapq = s.mapq; out->xs = s.xs; out->l_qseq = s.l_qseq;
    out->mtid = s.mtid; out->mpos = s.mpos; out->cigar_off = s.cigar_off; out->cigar = s.cigar; out->seq_off = s.seq_off; out->seq4 = s.seq4;
    out->name_code = s.name_code;
    out->lean = lean ? 1 : 0; out->n_cigar = s.n_cigar; out->seq2 = s.seq2; out->seqx_pos = s.seqx_pos; out->seqx_code = s.seqx_code;
    return PJ_OK;
}
int pj_staging_acquire(pj_ctx* c, int64_t cap_records, int64_t cap_cigar, int64_t cap_seq_bytes, pj_batch* out) {
    return staging_acquire(c, false, cap_records, cap_cigar, cap_seq_bytes, 0, out);
}
int pj_staging_acquire_lean(pj_ctx* c, int64_t cap_records, int64_t cap_cigar, int64_t cap_seq2_bytes, int64_t cap_seqx, pj_batch* out) {
    return staging_acquire(c, true, cap_records, cap_cigar, cap_seq2_bytes, cap_seqx, out);
}

int pj_batch_submit(pj_ctx* c, const pj_batch* b) {
    if (!c || !b) return fail(c, PJ_EINVAL, "pj_batch_submit: null argument");
    if (!c->shard_open) return fail(c, PJ_ESTATE, "pj_batch_submit: no open shard");
    const int64_t n = b->n_records;
    if (n < 0) return fail(c, PJ_EINVAL, "pj_batch_submit: negative record count");
    if (n == 0) return PJ_OK;
    if (n >= (1ll << 31)) return fail(c, PJ_EINVAL, "pj_batch_submit: more than 2^31 records in one batch");
    const bool lean = b->lean != 0;
    const bool oriented = c->orientation == PJ_ORIENT_FR || c->orientation == PJ_ORIENT_RF || c->orientation == PJ_ORIENT_FF;
    if (!b->pos || !b->flag || !b->mapq || !b->xs || !b->l_qseq) return fail(c, PJ_EINVAL, "pj_batch_submit: null column");
    if ((!lean || oriented) && (!b->mtid || !b->mpos)) return fail(c, PJ_EINVAL, "pj_batch_submit: mtid / mpos are required (the orientation enables the proper-pair rule)");
    if (!lean && (!b->tid || !b->cigar_off || !b->seq_off)) return fail(c, PJ_EINVAL, "pj_batch_submit: null column");
    if (lean && (!b->n_cigar || b->n_cigar_total < 0 || b->n_seq2_bytes < 0 || b->n_seqx < 0 || (b->n_seqx && (!b->seqx_pos || !b->seqx_code)) ||
                 b->const_tid < 0 || b->const_tid >= c->n_targets))
        return fail(c, PJ_EINVAL, "pj_batch_submit: malformed lean batch");
    if (c->extra && !b->name_code) return fail(c, PJ_EINVAL, "pj_batch_submit: the context computes the extra metrics, so batches must carry name_code");
    CU(c, cudaSetDevice(c->device));
    uint64_t cb = 0, ncig, sb = 0, nseq4 = 0, nseq2;
    std::vector<uint64_t> off2;                                   // classic batches: offsets of the records in the 2-bit stream, formed here
    if (!lean) {
        cb = b->cigar_off[0]; const uint64_t ce = b->cigar_off[n];
        sb = b->seq_off[0]; const uint64_t se = b->seq_off[n];
        if (ce < cb || se < sb) return fail(c, PJ_EINVAL, "pj_batch_submit: offsets not monotone");
        ncig = ce - cb; nseq4 = se - sb;
        // a record's SEQ enters the stream when it is complete ((l_qseq + 1) / 2 bytes); anything shorter counts as missing
        off2.resize((size_t)n + 1); off2[0] = c->n_seq;
        for (int64_t i = 0; i < n; i++) {
            const uint64_t a = b->seq_off[i], e = b->seq_off[i + 1]; const int32_t l = b->l_qseq[i];
            const uint64_t n4 = e >= a ? e - a : 0;
            off2[(size_t)i + 1] = off2[(size_t)i] + ((n4 > 0 && l > 0 && n4 >= (uint64_t)((l + 1) >> 1)) ? (uint64_t)((l + 3) >> 2) : 0ull);
        }
        nseq2 = off2[(size_t)n] - c->n_seq;
    } else { ncig = (uint64_t)b->n_cigar_total; nseq2 = (uint64_t)b->n_seq2_bytes; }
    if ((ncig && !b->cigar) || (nseq4 && !b->seq4) || (lean && nseq2 && !b->seq2)) return fail(c, PJ_EINVAL, "pj_batch_submit: null cigar/seq column");
    if (c->n_cig + ncig > 0xfffffff0ull) return fail(c, PJ_EINVAL, "pj_batch_submit: more than 2^32 CIGAR operations in one shard");
    cudaStream_t st = c->copy_stream;
    const size_t R = (size_t)c->n_rec, need = R + (size_t)n;
    int rc;
    if ((rc = ensure(c, c->tid, need, R, st)) || (rc = ensure(c, c->pos, need, R, st)) || (rc = ensure(c, c->l_qseq, need, R, st)) ||
        (rc = ensure(c, c->mtid, need, R, st)) || (rc = ensure(c, c->mpos, need, R, st)) || (rc = ensure(c, c->flag, need, R, st)) ||
        (rc = ensure(c, c->mapq, need, R, st)) || (rc = ensure(c, c->xs, need, R, st)) || (rc = ensure(c, c->cigar_off, need + 1, R + 1, st)) ||
        (rc = ensure(c, c->seq_off, need + 1, R + 1, st)) || (rc = ensure(c, c->cigar, (size_t)(c->n_cig + ncig) + 64, (size_t)c->n_cig, st)) ||
        (rc = ensure(c, c->seq2, (size_t)(c->n_seq + nseq2) + 256, (size_t)c->n_seq, st))) return rc;
    if (c->extra && (rc = ensure(c, c->name_code, need, R, st))) return rc;
    const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
    CU(c, cudaMemcpyAsync(c->pos.p + R, b->pos, n * 4, H2D, st)); CU(c, cudaMemcpyAsync(c->l_qseq.p + R, b->l_qseq, n * 4, H2D, st));
    CU(c, cudaMemcpyAsync(c->flag.p + R, b->flag, n * 2, H2D, st));
    CU(c, cudaMemcpyAsync(c->mapq.p + R, b->mapq, n, H2D, st)); CU(c, cudaMemcpyAsync(c->xs.p + R, b->xs, n, H2D, st));
    if (b->mtid && b->mpos) { CU(c, cudaMemcpyAsync(c->mtid.p + R, b->mtid, n * 4, H2D, st)); CU(c, cudaMemcpyAsync(c->mpos.p + R, b->mpos, n * 4, H2D, st)); }
    else { CU(c, cudaMemsetAsync(c->mtid.p + R, 0xff, n * 4, st)); CU(c, cudaMemsetAsync(c->mpos.p + R, 0xff, n * 4, st)); }   // never read without an orientation
    if (c->extra) CU(c, cudaMemcpyAsync(c->name_code.p + R, b->name_code, n * 8, H2D, st));
    if (ncig) {
        CU(c, cudaMemcpyAsync(c->cigar.p + c->n_cig, b->cigar + cb, ncig * 4, H2D, st));
        launch_prescan_cigar(c->cigar.p + c->n_cig, ncig, reinterpret_cast<uint32_t*>(c->d_shard_acc), c->d_shard_acc + 1, c->n_sm, st);
    }
    int64_t new_seqx = 0;
    if (!lean) {
        CU(c, cudaMemcpyAsync(c->tid.p + R, b->tid, n * 4, H2D, st));
        CU(c, cudaMemcpyAsync(c->cigar_off.p + R + 1, b->cigar_off + 1, n * 4, H2D, st));
        launch_rebase_u32(c->cigar_off.p + R + 1, n, (uint32_t)c->n_cig - (uint32_t)cb, st);
        // 4-bit SEQ -> 2-bit stream on the device (the prefix offsets of both forms come from the host)
        if ((rc = ensure(c, c->tmp_off4, (size_t)n + 1, 0, st)) || (rc = ensure(c, c->tmp_seq4, (size_t)nseq4 + 16, 0, st)) ||
            (rc = ensure(c, c->tmp_xcount, (size_t)n, 0, st)) || (rc = ensure(c, c->tmp_xoff, (size_t)n, 0, st)) ||
            (rc = ensure(c, c->tmp_scan, (size_t)scan_tmp_elems((uint64_t)n) + 4, 0, st))) return rc;
        CU(c, cudaMemcpyAsync(c->seq_off.p + R + 1, off2.data() + 1, n * 8, H2D, st));
        CU(c, cudaMemcpyAsync(c->tmp_off4.p, b->seq_off, (n + 1) * 8, H2D, st));
        if (nseq4) CU(c, cudaMemcpyAsync(c->tmp_seq4.p, b->seq4 + sb, nseq4, H2D, st));
        // the prefix columns of the batch must stay inside what was copied (a malformed batch must not send the kernels out of the arena)
        launch_check_offsets(0, n, c->cigar_off.p + R, c->tmp_off4.p, c->n_cig, c->n_cig + ncig, sb, sb + nseq4, c->d_shard_acc + 2, st);
        unsigned long long bad = 0;
        CU(c, cudaMemcpyAsync(&bad, c->d_shard_acc + 2, sizeof bad, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
        if (bad) return fail(c, PJ_EINVAL, "pj_batch_submit: cigar_off / seq_off are not non-decreasing prefix offsets within the batch's cigar / seq4 arrays");
        launch_seq4_to_2(n, c->tmp_seq4.p, c->tmp_off4.p, c->l_qseq.p + R, c->seq_off.p + R, c->seq2.p, c->flag.p + R, c->tmp_xcount.p, st);
        launch_exclusive_scan(c->tmp_xcount.p, c->tmp_xoff.p, (uint64_t)n, c->tmp_scan.p, c->tmp_scan.p + scan_tmp_elems((uint64_t)n), st);
        uint32_t nx = 0;
        CU(c, cudaMemcpyAsync(&nx, c->tmp_scan.p + scan_tmp_elems((uint64_t)n), sizeof nx, cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
        if (nx) {
            if ((rc = ensure(c, c->seqx_pos, (size_t)c->n_seqx + nx, (size_t)c->n_seqx, st)) || (rc = ensure(c, c->seqx_code, (size_t)c->n_seqx + nx, (size_t)c->n_seqx, st))) return rc;
            launch_seq4_exceptions(n, c->tmp_seq4.p, c->tmp_off4.p, c->l_qseq.p + R, c->seq_off.p + R, c->tmp_xcount.p, c->tmp_xoff.p,
                                   c->seqx_pos.p + c->n_seqx, c->seqx_code.p + c->n_seqx, st);
            new_seqx = nx;
        }
    } else {
        launch_fill_i32(c->tid.p + R, n, b->const_tid, st);
        if ((rc = ensure(c, c->tmp_ncig, (size_t)n, 0, st)) || (rc = ensure(c, c->tmp_fs, fs_scratch_bytes((uint32_t)n) / 8 + 2, 0, st))) return rc;
        CU(c, cudaMemcpyAsync(c->tmp_ncig.p, b->n_cigar, n * 2, H2D, st));
        launch_cigar_off((uint32_t)n, c->tmp_ncig.p, c->cigar_off.p + R, (uint32_t)c->n_cig, (uint32_t)ncig, c->d_shard_acc + 2, c->tmp_fs.p, st);
        if (nseq2) CU(c, cudaMemcpyAsync(c->seq2.p + c->n_seq, b->seq2, nseq2, H2D, st));
        launch_seq_off((uint32_t)n, c->cigar_off.p + R, c->cigar.p, c->l_qseq.p + R, c->seq_off.p + R, c->n_seq, nseq2, c->d_shard_acc + 2, c->tmp_fs.p, st);
        if (b->n_seqx) {
            const size_t nx = (size_t)b->n_seqx;
            if ((rc = ensure(c, c->seqx_pos, (size_t)c->n_seqx + nx, (size_t)c->n_seqx, st)) || (rc = ensure(c, c->seqx_code, (size_t)c->n_seqx + nx, (size_t)c->n_seqx, st))) return rc;
            CU(c, cudaMemcpyAsync(c->seqx_pos.p + c->n_seqx, b->seqx_pos, nx * 8, H2D, st));
            CU(c, cudaMemcpyAsync(c->seqx_code.p + c->n_seqx, b->seqx_code, nx, H2D, st));
            launch_rebase_u64(c->seqx_pos.p + c->n_seqx, (int64_t)nx, c->n_seq * 4ull, st);
            new_seqx = (int64_t)nx;
        }
    }
    CU(c, cudaGetLastError());
    {
        std::lock_guard<std::mutex> lk(c->staging_mu);
        for (StagingSlot* sl : c->slots) if (b->pos == sl->pos) { CU(c, cudaEventRecord(sl->done, st)); sl->state = 2; sl->seq_no = ++c->submit_seq; }
    }
    c->n_rec += n; c->n_cig += ncig; c->n_seq += nseq2; c->n_seqx += new_seqx;
    return PJ_OK;
}

int pj_shard_run(pj_ctx* c) {
    if (!c || !c->shard_open) return fail(c, PJ_ESTATE, "pj_shard_run: no open shard");
    CU(c, cudaSetDevice(c->device));
    if (c->prewarm_thread.joinable()) c->prewarm_thread.join();
    int rc = finish_genome(c); if (rc) return rc;
    cudaStream_t st = c->compute_stream;
    StreamTemps tmp(c->arena);
    const auto t_run0 = std::chrono::steady_clock::now();
    CU(c, cudaMemsetAsync(c->seq2.p + c->n_seq, 0, 16, c->copy_stream));  // tail slack of the SEQ stream: read in 8-byte words, masked out
    CU(c, cudaEventRecord(c->copies_done, c->copy_stream));
    CU(c, cudaStreamWaitEvent(st, c->copies_done, 0));
    c->n_stage = 0; c->n_launches = 0; c->have_result = false;
    const int64_t R = c->n_rec; const int32_t T = c->n_targets;
    if (R >= (int64_t)0xfffffff0ll) return fail(c, PJ_EINVAL, "pj_shard_run: more than 2^32 records in one shard");
    mark(c, "begin");
    // ---- per-target accumulators ----
    CU(c, cudaMemsetAsync(c->d_spliced, 0, T * 8, st)); CU(c, cudaMemsetAsync(c->d_unspliced, 0, T * 8, st)); CU(c, cudaMemsetAsync(c->d_sumq, 0, T * 8, st));
    CU(c, cudaMemsetAsync(c->d_maxq, 0, T * 4, st));
    launch_fill_i32(c->d_minq, T, INT32_MAX, st); c->n_launches++;
    CU(c, cudaMemsetAsync(c->d_scalars, 0, 16 * sizeof(uint32_t), st));
    Reads Rd{R, c->tid.p, c->pos.p, c->flag.p, c->mapq.p, c->xs.p, c->l_qseq.p, c->mtid.p, c->mpos.p, c->cigar_off.p, c->cigar.p, c->seq_off.p, c->seq2.p,
             c->seqx_pos.p, c->seqx_code.p, c->n_seqx};
    TargetAcc TA{c->d_spliced, c->d_unspliced, c->d_sumq, c->d_minq, c->d_maxq};
    uint32_t* d_err = c->d_scalars + 0; uint32_t* d_P = c->d_scalars + 2;
    uint32_t* d_J = c->d_scalars + 3; uint32_t* d_E = c->d_scalars + 4; uint32_t* d_tmp_total = c->d_scalars + 5;

    uint64_t *keys_a = nullptr, *keys_b = nullptr; uint32_t *vals_a = nullptr, *vals_b = nullptr, *counts = nullptr, *scan_tmp2 = nullptr;
    PairRec* pr = nullptr; unsigned long long* se_status = nullptr;
    uint32_t P = 0; int len_bits = 1, key_bits = 2;
    const int gbits = std::max(1, bit_length(c->h_toff[T]));
    auto alloc_pairs = [&](uint32_t cap) -> int {
        const size_t n = std::max<size_t>(cap, 1);
        const uint32_t nb = rs_num_blocks((uint32_t)n);
        CU(c, tmp.alloc(&keys_a, n * 8)); CU(c, tmp.alloc(&keys_b, n * 8));
        CU(c, tmp.alloc(&vals_a, n * 4)); CU(c, tmp.alloc(&vals_b, n * 4));
        CU(c, tmp.alloc(&counts, (size_t)256 * nb * 4));
        CU(c, tmp.alloc(&scan_tmp2, scan_tmp_elems(std::max<uint64_t>((uint64_t)256 * nb, n)) * 4));
        CU(c, tmp.alloc(&pr, n * sizeof(PairRec)));
        return PJ_OK;
    };
    {
        // front end: the longest N op and the number of N ops were accumulated while the batches were copied in
        unsigned long long acc[3] = {0, 0, 0};
        CU(c, cudaStreamSynchronize(c->copy_stream));
        CU(c, cudaMemcpy(acc, c->d_shard_acc, sizeof acc, cudaMemcpyDeviceToHost));
        if (acc[2]) return fail(c, PJ_EINVAL, "pj_shard_run: a submitted batch is inconsistent (prefix offsets outside its arrays, or the n_cigar / seq2 totals of a lean batch do not match its records)");
        const uint32_t maxN = (uint32_t)(acc[0] & 0xffffffffull);
        if (acc[1] >= 0xfffffff0ull) return fail(c, PJ_EINVAL, "pj_shard_run: more than 2^32 read-junction pairs in one shard");
        len_bits = std::max(1, bit_length(maxN)); key_bits = len_bits + gbits;
        if (key_bits > 64) return fail(c, PJ_EINVAL, "pj_shard_run: junction key needs %d bits; split the shard into fewer targets", key_bits);
        if ((rc = alloc_pairs((uint32_t)acc[1]))) return rc;
        CU(c, tmp.alloc(&se_status, (size_t)std::max<uint32_t>(se_num_tiles(R), 1) * sizeof(unsigned long long)));
        launch_scan_emit(Rd, c->d_tlen, T, c->d_toff, reinterpret_cast<const uint32_t*>(c->d_shard_acc), c->orientation, TA, keys_a, pr,
                         se_status, c->d_scalars + 6, d_P, (uint32_t)acc[1], d_err, st); c->n_launches++;
        mark(c, "scan_emit");
        CU(c, cudaMemcpyAsync(c->h_scalars, c->d_scalars, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(c, cudaStreamSynchronize(st));
        CU(c, cudaGetLastError());
        P = c->h_scalars[2];
    }
    c->n_pairs = P;

    uint32_t J = 0;
    if (P > 0) {
        int which;
        if (P < (1u << 30) && !c->legacy_sort) {
            uint32_t* os_scratch = nullptr;
            const size_t os_bytes = os_scratch_words(P, key_bits) * sizeof(uint32_t);
            CU(c, tmp.alloc(&os_scratch, os_bytes));
            which = launch_onesweep_sort(keys_a, vals_a, keys_b, vals_b, P, key_bits, os_scratch, c->n_sm, st, &c->n_launches);
            tmp.release(os_scratch, os_bytes);                         // stream order keeps the sort ahead of the next user of these bytes
        } else {
            which = launch_radix_sort(keys_a, vals_a, keys_b, vals_b, P, key_bits, counts, scan_tmp2, d_tmp_total, st, &c->n_launches);
        }
        const uint64_t* keys = which ? keys_b : keys_a; const uint32_t* vals = which ? vals_b : vals_a;
        uint32_t* spare_u32 = which ? vals_a : vals_b;        // free again: reused for the head flags / entropy flags
        mark(c, "radix_sort");
        uint32_t *jid = nullptr, *seg_start = nullptr; unsigned long long* fs_scratch = nullptr;
        CU(c, tmp.alloc(&jid, (size_t)P * 4));
        CU(c, tmp.alloc(&fs_scratch, ((size_t)fs_num_tiles(P) + 1) * 8));
        if (!c->legacy_sort) {
            CU(c, tmp.alloc(&seg_start, ((size_t)P + 1) * 4));        // J <= P is only known after the pass
            launch_segment(keys, P, jid, seg_start, d_J, fs_scratch, st); c->n_launches++;
            CU(c, cudaMemcpyAsync(c->h_scalars, c->d_scalars, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(c, cudaStreamSynchronize(st));
            J = c->h_scalars[3];
        } else {
            launch_seg_heads(keys, P, spare_u32, st); c->n_launches++;
            launch_exclusive_scan(spare_u32, spare_u32, P, scan_tmp2, d_J, st); c->n_launches += 3;
            CU(c, cudaMemcpyAsync(c->h_scalars, c->d_scalars, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            CU(c, cudaStreamSynchronize(st));
            J = c->h_scalars[3];
            CU(c, tmp.alloc(&seg_start, ((size_t)J + 1) * 4));
            launch_seg_ids(keys, P, spare_u32, jid, seg_start, J, st); c->n_launches++;
        }
        mark(c, "segments");
        // ---- per-junction accumulators: one zero-filled block of uint32 columns ----
        const size_t NCOL = 24; uint32_t* acc = nullptr; uint32_t* jadhist = nullptr; double* entropy = nullptr;
        CU(c, tmp.alloc(&acc, NCOL * (size_t)J * 4)); CU(c, cudaMemsetAsync(acc, 0, NCOL * (size_t)J * 4, st));
        CU(c, tmp.alloc(&jadhist, (size_t)J * (PJ_NB_JAD + 1) * 4)); CU(c, cudaMemsetAsync(jadhist, 0, (size_t)J * (PJ_NB_JAD + 1) * 4, st));
        CU(c, tmp.alloc(&entropy, (size_t)J * 8));
        auto col = [&](size_t k) { return acc + k * (size_t)J; };
        JuncAcc A{(int32_t*)col(0), (int32_t*)col(1), (int32_t*)col(2), (int32_t*)col(3), (int32_t*)col(4),
                  col(5), col(6), col(7), col(8), col(9), col(10), col(11), col(12), col(13), col(14), col(15), col(16), col(17), col(18), col(19),
                  col(20), col(21), col(22), col(23), jadhist};
        CU(c, cudaMemsetAsync(A.firstmm, 0xff, (size_t)J * 4, st));
        launch_junc_init(J, seg_start, keys, vals, pr, c->tid.p, len_bits, A, st); c->n_launches++;
        launch_reduce1(P, vals, jid, pr, (c->orientation == PJ_ORIENT_FR || c->orientation == PJ_ORIENT_RF || c->orientation == PJ_ORIENT_FF) ? 1 : 0,
                       A, spare_u32, st); c->n_launches++;
        mark(c, "reduce1");
        uint64_t* free_keys = which ? keys_a : keys_b;          // the non-result key buffer: 8 bytes per pair, reused below
        uint32_t* eoff = reinterpret_cast<uint32_t*>(free_keys);
        uint32_t* epos = eoff + P;
        if (!c->legacy_sort) { launch_entropy_index(P, spare_u32, eoff, epos, d_E, fs_scratch, st); c->n_launches++; }
        else {
            launch_exclusive_scan(spare_u32, eoff, P, scan_tmp2, d_E, st); c->n_launches += 3;
            launch_entropy_compact(P, spare_u32, eoff, epos, st); c->n_launches++;
        }
        launch_entropy_sum(J, seg_start, eoff, epos, entropy, st); c->n_launches++;
        mark(c, "entropy");
        Genome G{c->d_g2, c->d_gx, c->d_gsum, c->d_goff, c->d_glen, c->d_exc_pos, c->d_exc_byte, c->n_exc, c->n_exc_x, c->any_gx};
        uint4* pm = nullptr; CU(c, tmp.alloc(&pm, (size_t)P * sizeof(uint4)));
        {   // lanes per (read, junction) pair: one 16-base word per lane and step; long reads get wider groups
            int group = c->match_group;
            if (group <= 0) {
                // one lane per pair wins on every preset measured (c5, 1-10 kb reads: G = 1 / 2 / 4 / 8 -> 7.5 / 8.2 / 10.2 / 13.3 ms); wider groups
                // only for extreme anchors
                const double bases_per_pair = 4.0 * (double)c->n_seq / (double)P;      // read bases available per pair (lower bound on read length)
                group = bases_per_pair <= 4000 ? 1 : 8;
            }
            static const int match_ctas = [] { const char* e = getenv("PJ_MATCH_CTAS"); return e ? atoi(e) : 0; }();
            launch_match(P, group, match_ctas, vals, jid, pr, Rd, G, A, pm, d_err, st); c->n_launches++;
        }
        mark(c, "match");
        launch_reduce2(P, jid, pm, A, st); c->n_launches++;
        mark(c, "reduce2");
        if (c->rows_cap < J) { if (c->d_rows) { CU(c, cudaStreamSynchronize(st)); cudaFree(c->d_rows); } c->rows_cap = (size_t)J + J / 4 + 16; CU(c, cudaMalloc(&c->d_rows, c->rows_cap * sizeof(pj_junction))); }
        launch_finalize(J, seg_start, A, G, entropy, c->d_rows, d_err, st); c->n_launches++;
        mark(c, "finalize");
        if (c->extra && (rc = extra_keep_pairs(c, P, vals, jid, pr, st))) return rc;
    }
    if (c->extra && (rc = extra_classify(c, st))) return rc;
    mark(c, "end");
    CU(c, cudaMemcpyAsync(c->h_scalars, c->d_scalars, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    CU(c, cudaGetLastError());
    c->arena.reset(); c->arena.consolidate(tmp.high);      // everything that used the temporaries has completed
    // timings
    c->stage_ms.clear(); c->stage_names.clear();
    for (size_t k = 1; k < c->n_stage; k++) {
        float ms = 0; cudaEventElapsedTime(&ms, c->stages[k - 1].ev, c->stages[k].ev);
        c->stage_ms.push_back(ms); c->stage_names.push_back(c->stages[k].name);
    }
    cudaEventElapsedTime(&c->total_ms, c->stages[0].ev, c->stages[c->n_stage - 1].ev);
    static const bool trace_host = getenv("PJ_TRACE_HOST") != nullptr;     // debugging aid: report runs that took 20 % longer than the best one
    if (trace_host) {
        static float best = 1e30f; best = std::min(best, c->total_ms);
        if (c->total_ms > 1.2f * best) {
            std::string m;
            for (size_t k = 0; k < c->stage_ms.size(); k++) { char b[64]; snprintf(b, sizeof b, " %s=%.2f", c->stage_names[k], c->stage_ms[k]); m += b; }
            fprintf(stderr, "[pj trace] slow run: host %.2f ms, device %.2f ms (best %.2f):%s\n",
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_run0).count(), c->total_ms, best, m.c_str());
        }
    }
    const uint32_t err = c->h_scalars[0];
    if (err) {
        std::string m = "input rejected (the reference aborts or is undefined on it):";
        if (err & ERR_ANCHOR_ORDER) m += " intron not enclosed by its anchors;";
        if (err & ERR_START_RANGE) m += " intron start outside the target (donor site cannot be fetched);";
        if (err & ERR_KEY_OVERFLOW) m += " internal key overflow;";
        if (err & ERR_NO_PRESENCE) m += " alignment has no presence in the requested region;";
        if (err & ERR_QUERY_RANGE) m += " CIGAR consumes more query bases than SEQ holds;";
        if (err & ERR_EMPTY_ANCHOR) m += " empty anchor window;";
        if (err & ERR_GENOME_RANGE) m += " junction window leaves the genome sequence (or the target has no sequence loaded);";
        if (err & ERR_SEQ_MISSING) m += " spliced read without SEQ bytes;";
        if (err & ERR_ZERO_LEN) m += " zero-length window op;";
        return fail(c, (err & ERR_KEY_OVERFLOW) ? PJ_EINVAL : PJ_EDATA, "%s", m.c_str());
    }
    c->n_junc = J; c->have_result = true;
    return PJ_OK;
}

int64_t pj_shard_num_junctions(const pj_ctx* c) { return (c && c->have_result) ? c->n_junc : -1; }

int pj_shard_fetch(pj_ctx* c, pj_junction* rows, int64_t cap_rows, pj_target_stats* stats, int32_t cap_targets) {
    if (!c || !c->have_result) return fail(c, PJ_ESTATE, "pj_shard_fetch: no result (run pj_shard_run first)");
    if (cap_rows < c->n_junc || (c->n_junc && !rows)) return fail(c, PJ_EINVAL, "pj_shard_fetch: rows capacity %lld < %lld", (long long)cap_rows, (long long)c->n_junc);
    CU(c, cudaSetDevice(c->device));
    if (c->n_junc) CU(c, cudaMemcpyAsync(rows, c->d_rows, (size_t)c->n_junc * sizeof(pj_junction), cudaMemcpyDeviceToHost, c->compute_stream));
    if (stats) {
        if (cap_targets < c->n_targets) return fail(c, PJ_EINVAL, "pj_shard_fetch: stats capacity too small");
        const int32_t T = c->n_targets;
        std::vector<unsigned long long> a(T), b(T), s(T); std::vector<int32_t> mn(T), mx(T);
        CU(c, cudaMemcpyAsync(a.data(), c->d_spliced, T * 8, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(c, cudaMemcpyAsync(b.data(), c->d_unspliced, T * 8, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(c, cudaMemcpyAsync(s.data(), c->d_sumq, T * 8, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(c, cudaMemcpyAsync(mn.data(), c->d_minq, T * 4, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(c, cudaMemcpyAsync(mx.data(), c->d_maxq, T * 4, cudaMemcpyDeviceToHost, c->compute_stream));
        CU(c, cudaStreamSynchronize(c->compute_stream));
        for (int32_t t = 0; t < T; t++) stats[t] = pj_target_stats{a[t], b[t], s[t], mn[t], mx[t]};
    }
    CU(c, cudaStreamSynchronize(c->compute_stream));
    return PJ_OK;
}

int pj_shard_timing(const pj_ctx* c, float* total_ms, int32_t* n_launches) {
    if (!c) return PJ_EINVAL;
    if (total_ms) *total_ms = c->total_ms;
    if (n_launches) *n_launches = c->n_launches;
    return PJ_OK;
}

int pj_shard_kernel_times(const pj_ctx* c, int32_t cap, float* kernel_ms, const char** kernel_names, int32_t* n) {
    if (!c || !n) return PJ_EINVAL;
    const int32_t k = (int32_t)c->stage_ms.size();
    *n = k;
    for (int32_t i = 0; i < k && i < cap; i++) { if (kernel_ms) kernel_ms[i] = c->stage_ms[i]; if (kernel_names) kernel_names[i] = c->stage_names[i]; }
    return PJ_OK;
}

} // extern "C"
