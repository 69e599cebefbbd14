// pj_ctx.hpp — the context object behind include/portcullis_junc.h, shared by pj_api.cu (the junction pipeline) and
// pj_extra.cu (the `--extra` metrics).  Internal to the library.
#pragma once
#include "junc_launch.hpp"
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace pjapi {

template <typename T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    void free_() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct StagingSlot {
    // one pinned host block carved into the columns of a pj_batch
    uint8_t* block = nullptr; size_t block_bytes = 0;
    int32_t *tid = nullptr, *pos = nullptr, *l_qseq = nullptr, *mtid = nullptr, *mpos = nullptr;
    uint16_t* flag = nullptr; uint8_t *mapq = nullptr, *xs = nullptr;
    uint32_t *cigar_off = nullptr, *cigar = nullptr; uint64_t* seq_off = nullptr; uint8_t* seq4 = nullptr;
    uint64_t* name_code = nullptr;         // only when the context computes the `--extra` metrics
    uint16_t* n_cigar = nullptr; uint8_t* seq2 = nullptr; uint64_t* seqx_pos = nullptr; uint8_t* seqx_code = nullptr;   // lean slots
    bool lean = false;
    int64_t cap_rec = 0, cap_cig = 0, cap_seq = 0, cap_seqx = 0;
    cudaEvent_t done = nullptr;
    int state = 0;                         // 0 free, 1 handed out (being filled / waiting for submit), 2 copy in flight
    uint64_t seq_no = 0;                   // submit order, to find the oldest in-flight slot
};

struct StageTime { const char* name; cudaEvent_t ev; };

// The temporaries of one pj_shard_run (sort keys, pair records, per-junction accumulators ...) come out of this per-context arena: a
// bump pointer over one cudaMalloc'ed block.  The stream-ordered allocator they used before (cudaMallocAsync on the default pool) took
// 2-450 ms for single allocations during the first dozen runs of every context (profiles/r2_history.md, "allocator stalls").
// A run that outgrows the block spills into extra chunks; consolidate() at the end of that run replaces them by one block of the total.
struct TempArena {
    struct Chunk { char* p; size_t cap, used; };
    std::vector<Chunk> chunks; size_t need = 0;
    static size_t align(size_t b) { return (b + 255) & ~(size_t)255; }
    size_t capacity() const { size_t t = 0; for (const Chunk& k : chunks) t += k.cap; return t; }
    cudaError_t alloc(void** out, size_t bytes) {
        const size_t a = align(bytes ? bytes : 1); need += a;
        for (Chunk& k : chunks) if (k.used + a <= k.cap) { *out = k.p + k.used; k.used += a; return cudaSuccess; }
        Chunk k{nullptr, std::max(a, (size_t)64 << 20), a};
        const cudaError_t e = cudaMalloc((void**)&k.p, k.cap);
        if (e != cudaSuccess) return e;
        chunks.push_back(k); *out = k.p;
        return cudaSuccess;
    }
    void release(void* p, size_t bytes) {              // only the most recent allocation of a chunk can be handed back
        const size_t a = align(bytes ? bytes : 1);
        for (Chunk& k : chunks) if (k.used >= a && k.p + (k.used - a) == (char*)p) { k.used -= a; need -= a; return; }
    }
    void reset() { for (Chunk& k : chunks) k.used = 0; need = 0; }
    void free_all() { for (Chunk& k : chunks) cudaFree(k.p); chunks.clear(); need = 0; }
    // all work that used the arena has completed (the caller synchronised the stream)
    void consolidate(size_t high_water) {
        if (chunks.size() <= 1) return;
        free_all();
        Chunk k{nullptr, high_water + high_water / 16 + ((size_t)1 << 20), 0};
        if (cudaMalloc((void**)&k.p, k.cap) == cudaSuccess) chunks.push_back(k); else cudaGetLastError();
    }
    void reserve(size_t bytes) {                       // nothing of the arena is in use
        if (capacity() >= bytes && chunks.size() == 1) return;
        free_all();
        Chunk k{nullptr, bytes, 0};
        if (cudaMalloc((void**)&k.p, k.cap) == cudaSuccess) chunks.push_back(k); else cudaGetLastError();
    }
};

} // namespace pjapi

struct pj_ctx {
    int device = 0; int orientation = PJ_ORIENT_UNKNOWN; int match_group = 0; int n_sm = 148; int legacy_sort = 0;
    cudaStream_t copy_stream = nullptr, compute_stream = nullptr;
    std::string err;
    // targets / genome
    int32_t n_targets = 0;
    std::vector<int32_t> h_tlen; std::vector<uint64_t> h_toff, h_goff; std::vector<int64_t> h_glen;
    int32_t* d_tlen = nullptr; uint64_t* d_toff = nullptr; uint64_t* d_goff = nullptr; int64_t* d_glen = nullptr;
    uint64_t* d_g2 = nullptr; uint64_t* d_gx = nullptr; uint32_t* d_gsum = nullptr; uint64_t g_total_bases = 0;
    uint64_t* d_exc_pos = nullptr; uint8_t* d_exc_byte = nullptr; uint32_t* d_exc_count = nullptr; uint32_t exc_cap = 1u << 20;
    int32_t n_exc = 0, n_exc_x = 0, any_gx = 0; bool genome_dirty = false;
    std::mutex genome_mu;               // pj_genome_set_target (its own host thread) against finish_genome (inside pj_shard_run / pj_features_*): a target may still be
                                        // uploading while a shard that does not need it runs
    uint8_t* h_graw[2] = {nullptr, nullptr}; uint8_t* d_graw[2] = {nullptr, nullptr}; cudaEvent_t graw_ev[2] = {nullptr, nullptr};
    static constexpr size_t GRAW_CHUNK = 64u << 20;
    // shard arena
    bool shard_open = false;
    int64_t n_rec = 0; uint64_t n_cig = 0, n_seq = 0;
    pjapi::DevBuf<int32_t> tid, pos, l_qseq, mtid, mpos; pjapi::DevBuf<uint16_t> flag; pjapi::DevBuf<uint8_t> mapq, xs, seq2;
    pjapi::DevBuf<uint32_t> cigar_off, cigar; pjapi::DevBuf<uint64_t> seq_off;
    // non-ACGT read bases of the shard (sorted by base index) + per-batch scratch of pj_batch_submit
    pjapi::DevBuf<uint64_t> seqx_pos; pjapi::DevBuf<uint8_t> seqx_code; int64_t n_seqx = 0;
    pjapi::DevBuf<uint8_t> tmp_seq4; pjapi::DevBuf<uint64_t> tmp_off4; pjapi::DevBuf<uint32_t> tmp_xcount, tmp_xoff, tmp_scan; pjapi::DevBuf<uint16_t> tmp_ncig;
    pjapi::DevBuf<unsigned long long> tmp_fs;
    std::vector<pjapi::StagingSlot*> slots; int max_slots = 4; uint64_t submit_seq = 0;
    int64_t slot_floor[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // [classic | lean][records, CIGAR words, SEQ bytes, exceptions]: largest capacities any slot was given
    pjapi::TempArena arena;             // temporaries of pj_shard_run
    std::thread prewarm_thread;         // sizes the arena while the caller is still decoding
    double t_pinned_alloc_s = 0; size_t pinned_alloc_bytes = 0; int n_pinned_allocs = 0;   // PJ_TRACE: cost of growing the staging pool
    std::mutex staging_mu;              // pj_staging_acquire / pj_batch_submit may be called from several host threads
    cudaStream_t genome_stream = nullptr;
    cudaEvent_t copies_done = nullptr;
    // per-target accumulators + misc device scalars
    unsigned long long *d_spliced = nullptr, *d_unspliced = nullptr, *d_sumq = nullptr; int32_t *d_minq = nullptr, *d_maxq = nullptr;
    uint32_t* d_scalars = nullptr;    // [0]=err [1]=max_nlen [2]=P [3]=J [4]=E [5]=scratch total [6]=tile ticket
    unsigned long long* d_shard_acc = nullptr;   // [0] (low 32 bits) longest N op, [1] number of N ops of the shard, [2] != 0: a batch had malformed prefix offsets / inconsistent lean totals; filled while batches are copied in
    uint32_t* h_scalars = nullptr;    // pinned mirror
    // results
    pj_junction* d_rows = nullptr; size_t rows_cap = 0; int64_t n_junc = 0; uint64_t n_pairs = 0;
    bool have_result = false;
    // `--extra` metrics (pj_extra.cu): state kept between pj_shard_run and the pj_extra_* calls
    bool extra = false;
    pjapi::DevBuf<uint64_t> name_code;                                  // per record of the arena
    uint32_t* x_pair_rid = nullptr; uint32_t* x_pair_jid = nullptr;     // record / junction of every pair, in sorted-pair order
    uint64_t* x_names = nullptr; int64_t x_n_spliced = -1;              // name codes of the shard's spliced records
    uint32_t* x_uflag = nullptr; int32_t* x_alen = nullptr;             // per record: unspliced + mapped flag, reference span (until pj_extra_run)
    std::vector<uint64_t> x_imported;                                   // spliced names of other contexts (pj_extra_import_names)
    uint32_t* x_depth = nullptr; std::vector<uint64_t> x_doff;          // unspliced pileup per target: x_doff[t] .. + tlen + 1 slots
    std::vector<uint8_t> x_covered; std::vector<uint32_t> x_maxlive; bool x_ready = false;
    std::vector<float> x_stage_ms; std::vector<const char*> x_stage_names; float x_total_ms = 0; int x_launches = 0;
    // timing
    std::vector<pjapi::StageTime> stages; size_t n_stage = 0; float total_ms = 0; int n_launches = 0;
    std::vector<float> stage_ms; std::vector<const char*> stage_names;
};


namespace pjapi {

extern thread_local std::string g_last_error;
int fail(pj_ctx* c, int code, const char* fmt, ...);

#define CU(c, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return pjapi::fail((c), PJ_ECUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); } while (0)

template <typename T> int ensure(pj_ctx* c, DevBuf<T>& b, size_t need, size_t keep, cudaStream_t st) {
    if (need <= b.cap) return PJ_OK;
    size_t ncap = std::max(need, b.cap + b.cap / 2 + 1024);
    T* np = nullptr;
    CU(c, cudaMalloc(&np, ncap * sizeof(T)));
    if (b.p && keep) CU(c, cudaMemcpyAsync(np, b.p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
    if (b.p) { CU(c, cudaStreamSynchronize(st)); cudaFree(b.p); }
    b.p = np; b.cap = ncap;
    return PJ_OK;
}

int finish_genome(pj_ctx* c);           // pj_api.cu: sorts + uploads the genome's exception table once the uploads are done

// pj_extra.cu: called at the end of pj_shard_run / from pj_shard_begin / pj_destroy
int extra_keep_pairs(pj_ctx* c, uint32_t n_pairs, const uint32_t* vals, const uint32_t* jid, const pjk::PairRec* pr, cudaStream_t st);
int extra_classify(pj_ctx* c, cudaStream_t st);
void extra_reset(pj_ctx* c);

} // namespace pjapi
