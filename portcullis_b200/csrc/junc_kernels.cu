// junc_kernels.cu — hand-written sm_100a kernels of the junc hot path.
//
// Pipeline over one shard (all alignment columns resident in HBM):
//   k_prescan_cigar   at submit : longest N op and N-op count of the shard (key width, pair capacity)
//   k_scan_emit       per read : N-op count, reference span, per-target scalars (A1); pair slots by look-back scan;
//                     CIGAR walk, one (key, PairA..PairD) record per N op (A2, A3, A6 flags)
//   k_os_*            one-sweep stable LSD radix sort of (key, pair index): junction order, BAM order inside a junction
//   k_flag_scan       junction ids and segment starts (single pass)
//   k_reduce1         per pair : warp-shuffle segmented reduce of the integer metrics + entropy run boundaries
//   k_entropy_*       run-length terms -> fp64 entropy per junction (A5, quirk Q1)
//   k_match           warp per pair : junction-wide anchor window walk against the packed genome (A10)
//   k_reduce2         per pair : segmented reduce of mmes / mismatches / minMatch / JAD histogram (A11)
//   k_finalize        per junction : strands, motif, Hamming distances, JAD suffix sums, row assembly (A4, A8, A9)
#include "junc_kernels.cuh"
#include "junc_launch.hpp"
#include <cstdio>
#include <cstdlib>
#include <algorithm>

namespace pjk {

#define FULL 0xffffffffu

// ================================================================================================
// generic exclusive scan (uint32), three phases; tile = 512 threads x 8 items
// ================================================================================================
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, o); if (lane >= o) v += t; }
    return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = warp_incl_scan(v, lane);
    __syncthreads();                                  // protect wsum reuse across calls
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < nw ? wsum[lane] : 0;
        uint32_t si = warp_incl_scan(s, lane);
        wsum[lane] = si - s;
        if (lane == 31) *total = si;                  // lanes >= nw contribute 0, so lane 31 holds the block sum
    }
    __syncthreads();
    return inc - v + wsum[w];
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ bsum) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + (uint64_t)k * SCAN_THREADS + threadIdx.x; if (i < n) s += in[i]; }
    __shared__ uint32_t tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_bsums(uint32_t* __restrict__ bsum, uint32_t nb, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t tot;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        uint32_t i = b0 + threadIdx.x;
        uint32_t v = i < nb ? bsum[i] : 0;
        uint32_t ex = block_excl_scan(v, &tot);
        if (i < nb) bsum[i] = ex + carry;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n,
                                                              const uint32_t* __restrict__ bsum) {
    // blocked arrangement: thread t owns items [t*ITEMS, t*ITEMS+ITEMS) of the tile
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS]; uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + k; v[k] = i < n ? in[i] : 0; s += v[k]; }
    __shared__ uint32_t tot;
    uint32_t ex = block_excl_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { uint64_t i = base + k; if (i < n) out[i] = ex; ex += v[k]; }
}

void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* bsum_tmp, uint32_t* total_dev, cudaStream_t st) {
    const uint32_t nb = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
    if (n == 0) { cudaMemsetAsync(total_dev, 0, sizeof(uint32_t), st); return; }
    k_scan_partials<<<nb, SCAN_THREADS, 0, st>>>(in, n, bsum_tmp);
    k_scan_bsums<<<1, 1024, 0, st>>>(bsum_tmp, nb, total_dev);
    k_scan_apply<<<nb, SCAN_THREADS, 0, st>>>(in, out, n, bsum_tmp);
}
uint64_t scan_tmp_elems(uint64_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

// ================================================================================================
// small utility kernels
// ================================================================================================
__global__ void k_rebase_u32(uint32_t* __restrict__ a, int64_t n, uint32_t add) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] += add;
}
__global__ void k_rebase_u64(uint64_t* __restrict__ a, int64_t n, uint64_t add) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] += add;
}
__global__ void k_fill_i32(int32_t* __restrict__ a, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = v;
}
// samtools' coordinate order (bam_sort.c bam1_lt, samtools 1.x): key = tid << 32 | (pos + 1) << 1 | reverse-strand; unplaced
// reads (tid = -1) compare as the largest target.  Used by `prep` (csrc/prep_driver.cpp); the LSD radix sort is stable, so
// records with equal keys keep their input order, like samtools' merge sort.
__global__ void __launch_bounds__(256) k_coord_keys(int64_t n, const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, const uint16_t* __restrict__ flag,
                                                     uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lo = ((uint32_t)(pos[i] + 1) << 1) | ((flag[i] & 0x10u) ? 1u : 0u);
    keys[i] = ((uint64_t)(uint32_t)tid[i] << 32) | lo;
    vals[i] = (uint32_t)i;
}
void launch_coord_keys(int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, uint64_t* keys, uint32_t* vals, cudaStream_t st) {
    if (n > 0) k_coord_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, tid, pos, flag, keys, vals);
}

void launch_fill_i32(int32_t* a, int64_t n, int32_t v, cudaStream_t st) { if (n > 0) k_fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n, v); }
void launch_rebase_u32(uint32_t* a, int64_t n, uint32_t add, cudaStream_t st) { if (n > 0 && add) k_rebase_u32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n, add); }
// Offsets of a submitted batch (already rebased into the shard arena): both prefix columns must be non-decreasing and stay inside
// what has been copied, or the CIGAR / SEQ walks of the pipeline would leave the arena.  Sets *bad on the first violation.
__global__ void __launch_bounds__(256) k_check_offsets(int64_t first, int64_t n, const uint32_t* __restrict__ cigar_off, const uint64_t* __restrict__ seq_off,
                                                        uint64_t cig_lo, uint64_t cig_hi, uint64_t seq_lo, uint64_t seq_hi, unsigned long long* __restrict__ bad) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t i = first + k;
    const uint64_t c0 = cigar_off[i], c1 = cigar_off[i + 1], s0 = seq_off[i], s1 = seq_off[i + 1];
    if (c0 > c1 || c0 < cig_lo || c1 > cig_hi || s0 > s1 || s0 < seq_lo || s1 > seq_hi) *bad = 1ull;
}
void launch_check_offsets(int64_t first, int64_t n, const uint32_t* cigar_off, const uint64_t* seq_off, uint64_t cig_lo, uint64_t cig_hi,
                          uint64_t seq_lo, uint64_t seq_hi, unsigned long long* bad, cudaStream_t st) {
    if (n > 0) k_check_offsets<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(first, n, cigar_off, seq_off, cig_lo, cig_hi, seq_lo, seq_hi, bad);
}
void launch_rebase_u64(uint64_t* a, int64_t n, uint64_t add, cudaStream_t st) { if (n > 0 && add) k_rebase_u64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, n, add); }

// ================================================================================================
// genome packing: raw FASTA bytes (already unwrapped) -> 2-bit plane + exception bitmask
// one thread packs 64 bases: four 128-bit loads, two g2 words, one gx word
// ================================================================================================
__global__ void __launch_bounds__(256) k_pack_genome(const uint8_t* __restrict__ raw, int64_t n, uint64_t base_index /* multiple of 64 */,
                                                      uint64_t* __restrict__ g2, uint64_t* __restrict__ gx, uint32_t* __restrict__ gsum,
                                                      uint64_t* __restrict__ exc_pos, uint8_t* __restrict__ exc_byte,
                                                      uint32_t* __restrict__ exc_count /* [0] side-table entries, [1] != 0: some base is not A/C/G/T */, uint32_t exc_cap) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b0 = t * 64;
    if (b0 >= n) return;
    uint8_t c[64];
    if (b0 + 64 <= n && ((reinterpret_cast<uintptr_t>(raw + b0) & 15) == 0)) {
        const uint4* p = reinterpret_cast<const uint4*>(raw + b0);
#pragma unroll
        for (int k = 0; k < 4; k++) { uint4 v = __ldg(p + k); memcpy(c + 16 * k, &v, 16); }
    } else {
#pragma unroll 8
        for (int k = 0; k < 64; k++) c[k] = (b0 + k < n) ? raw[b0 + k] : (uint8_t)'A';
    }
    uint64_t w0 = 0, w1 = 0, x = 0;
#pragma unroll
    for (int k = 0; k < 64; k++) {
        uint8_t ch = c[k];
        if (ch >= 'a' && ch <= 'z') ch -= 32;                        // boost::to_upper
        uint32_t code; bool exc = false;
        switch (ch) { case 'A': code = 0; break; case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break;
                      default: exc = true; code = (ch == 'N') ? 0u : 1u; }
        if (b0 + k >= n) { exc = false; code = 0; }
        if (exc) {
            x |= 1ull << k;
            if (code == 1u) {
                uint32_t slot = atomicAdd(exc_count, 1u);
                if (slot < exc_cap) { exc_pos[slot] = base_index + (uint64_t)(b0 + k); exc_byte[slot] = ch; }
            }
        }
        if (k < 32) w0 |= (uint64_t)code << (2 * k); else w1 |= (uint64_t)code << (2 * (k - 32));
    }
    const uint64_t gi = base_index + (uint64_t)b0;
    g2[gi >> 5] = w0; g2[(gi >> 5) + 1] = w1; gx[gi >> 6] = x;
    if (x) { exc_count[1] = 1u; atomicOr(gsum + (gi >> 15), 1u << ((gi >> 10) & 31)); }   // 64 | 1024: the thread's bases lie in one stretch
}

void launch_pack_genome(const uint8_t* raw, int64_t n, uint64_t base_index, uint64_t* g2, uint64_t* gx, uint32_t* gsum,
                        uint64_t* exc_pos, uint8_t* exc_byte, uint32_t* exc_count, uint32_t exc_cap, cudaStream_t st) {
    if (n <= 0) return;
    const int64_t threads = (n + 63) / 64;
    k_pack_genome<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(raw, n, base_index, g2, gx, gsum, exc_pos, exc_byte, exc_count, exc_cap);
}

// ================================================================================================
// k_scan_emit: k_scan_reads + the prefix scan of the N-op counts + k_emit_pairs in ONE pass over the records.
// A tile of 1024 records is scanned (N-op counts, reference spans, per-target scalars), its pair slots come from a
// block scan plus a decoupled look-back over per-tile status words (tiles take their index from an atomic ticket),
// and the pair records are written straight away while the tile's CIGAR words are still in L1.  The record columns are
// read once and the npairs / pair_off / read_end arrays disappear.  The width of the size field of the key comes from
// the longest N op of the shard, which pj_batch_submit tracks while the batches are copied in (k_prescan_cigar).
// ================================================================================================
constexpr int SE_THREADS = 256;
constexpr unsigned long long SE_PREFIX = 1ull << 63, SE_AGG = 1ull << 62, SE_MASK = (1ull << 62) - 1ull;

// Decoupled look-back by a whole warp: lane l polls the status word of tile (first - l), 32 predecessors per probe, and the warp
// adds the aggregates down to the nearest inclusive prefix.  (One polling thread pays a dependent global round trip per
// predecessor; with a few hundred tiles in flight that chain was most of the time of the single-pass scans.)
// Must be called by all 32 lanes of one warp; every lane returns the exclusive prefix of `tile`.
__device__ __forceinline__ unsigned long long lookback_warp(const unsigned long long* status, int64_t tile, int lane) {
    unsigned long long excl = 0;
    for (int64_t first = tile - 1; first >= 0; first -= 32) {
        const int64_t t = first - lane;
        unsigned long long v = SE_PREFIX;                              // before tile 0: an inclusive prefix of 0
        if (t >= 0) { const volatile unsigned long long* sp = status + t; do { v = *sp; } while ((v >> 62) == 0ull); }
        const uint32_t pm = __ballot_sync(FULL, (v & SE_PREFIX) != 0ull);
        const int stop = pm ? __ffs(pm) - 1 : 31;                      // nearest tile that already knows its inclusive prefix
        unsigned long long add = lane <= stop ? (v & SE_MASK) : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) add += __shfl_xor_sync(FULL, add, o);
        excl += add;
        if (pm) break;
    }
    return excl;
}

__global__ void __launch_bounds__(256) k_prescan_cigar(const uint32_t* __restrict__ cigar, uint64_t n, uint32_t* __restrict__ max_nlen,
                                                        unsigned long long* __restrict__ n_nops) {
    uint32_t m = 0, k = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = __ldg(cigar + i);
        if (cig_op(w) == OP_N) { m = max(m, (uint32_t)cig_len(w)); k++; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m = max(m, __shfl_xor_sync(FULL, m, o)); k += __shfl_xor_sync(FULL, k, o); }
    if ((threadIdx.x & 31) == 0 && k) { atomicMax(max_nlen, m); atomicAdd(n_nops, (unsigned long long)k); }
}
// Longest N op and number of N ops of a batch's CIGAR words (upper bound of its read-junction pairs), accumulated per shard.
void launch_prescan_cigar(const uint32_t* cigar, uint64_t n, uint32_t* max_nlen, unsigned long long* n_nops, int n_sm, cudaStream_t st) {
    if (!n) return;
    const unsigned blocks = (unsigned)std::min<uint64_t>((uint64_t)n_sm * 8, (n + 255) / 256);
    k_prescan_cigar<<<blocks, 256, 0, st>>>(cigar, n, max_nlen, n_nops);
}

template <int SE_ITEMS>
__global__ void __launch_bounds__(SE_THREADS, 5) k_scan_emit(Reads R, const int32_t* __restrict__ tlen, int32_t n_targets, const uint64_t* __restrict__ toff,
                                                           const uint32_t* __restrict__ max_nlen, int32_t orientation, TargetAcc T,
                                                           uint64_t* __restrict__ keys, PairRec* __restrict__ pr,
                                                           unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket,
                                                           uint32_t* __restrict__ total_pairs, uint32_t pair_cap, uint32_t* __restrict__ err) {
    __shared__ uint32_t s_tile, s_tot;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long b_sp[4], b_us[4], b_sum[4];
    __shared__ int32_t b_mn[4], b_mx[4];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    if (threadIdx.x < 4) { b_sp[threadIdx.x] = 0; b_us[threadIdx.x] = 0; b_sum[threadIdx.x] = 0; b_mn[threadIdx.x] = INT32_MAX; b_mx[threadIdx.x] = 0; }
    __syncthreads();
    constexpr int SE_TILE = SE_THREADS * SE_ITEMS;
    const uint32_t tile = s_tile;
    const int64_t base = (int64_t)tile * SE_TILE;
    const int32_t tid_first = R.tid[min(base, R.n - 1)];
    const int32_t len_bits = max(1, 32 - __clz((int)__ldg(max_nlen)));
    // ---- phase A: scan the tile's records.  Loads and CIGAR walks of the thread's records come first (independent, so
    // their latencies overlap); the warp collectives for the per-target scalars follow in a second loop. ----
    uint32_t cnt[SE_ITEMS]; int32_t rend[SE_ITEMS]; bool zeron[SE_ITEMS];
    int32_t tids[SE_ITEMS], lqs[SE_ITEMS]; bool viss[SE_ITEMS];
#pragma unroll
    for (int r = 0; r < SE_ITEMS; r++) {
        const int64_t i = base + r * SE_THREADS + threadIdx.x;
        bool vis = false; int32_t tid = -2, lq = 0; uint32_t nN = 0; rend[r] = 0; zeron[r] = false;
        if (i < R.n) {
            tid = R.tid[i];
            const int32_t pos = R.pos[i];
            const uint32_t c0 = R.cigar_off[i], c1 = R.cigar_off[i + 1];
            const uint32_t flag = R.flag[i];
            lq = R.l_qseq[i];
            int64_t rlen = 0;
            // the first four ops with independent loads (one round trip covers most short-read CIGARs; a zero word is a
            // neutral 0M), the rest one by one
            uint32_t w4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) w4[k] = (c0 + k < c1) ? __ldg(R.cigar + c0 + k) : 0u;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t w = w4[k], op = cig_op(w);
                if (op_ref(op)) rlen += cig_len(w);
                if (op == OP_N) { nN++; if (cig_len(w) == 0) zeron[r] = true; }
            }
            for (uint32_t c = c0 + 4; c < c1; c++) {
                const uint32_t w = __ldg(R.cigar + c), op = cig_op(w);
                if (op_ref(op)) rlen += cig_len(w);
                if (op == OP_N) { nN++; if (cig_len(w) == 0) zeron[r] = true; }
            }
            rend[r] = (int32_t)(pos + rlen - 1);
            if (tid >= 0 && tid < n_targets) {
                const int64_t endpos = (!(flag & 0x4) && c1 > c0) ? (int64_t)pos + rlen : (int64_t)pos + 1;
                vis = pos < tlen[tid] && endpos > 0;
            }
            if (!vis) nN = 0;
        }
        cnt[r] = nN; tids[r] = tid; lqs[r] = lq; viss[r] = vis;
    }
#pragma unroll
    for (int r = 0; r < SE_ITEMS; r++) {
        const bool vis = viss[r]; const int32_t tid = tids[r], lq = lqs[r]; const uint32_t nN = cnt[r];
        // per-target scalars: warp partials into the block's shared slots (slot = tid - first tid of the tile)
        const uint32_t vmask = __ballot_sync(FULL, vis);
        if (vmask) {
            const int32_t t0 = __shfl_sync(FULL, tid, __ffs(vmask) - 1);
            if (__all_sync(FULL, !vis || tid == t0)) {
                const uint32_t sp = __popc(__ballot_sync(FULL, vis && nN > 0)), us = __popc(vmask) - sp;
                unsigned long long sm = vis ? (unsigned long long)(long long)lq : 0ull;
                int32_t mn = vis ? lq : INT32_MAX, mx = vis ? lq : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) { sm += __shfl_xor_sync(FULL, sm, o); mn = min(mn, __shfl_xor_sync(FULL, mn, o)); mx = max(mx, __shfl_xor_sync(FULL, mx, o)); }
                if (lane == 0) {
                    const int32_t slot = t0 - tid_first;
                    if (slot >= 0 && slot < 4) { atomicAdd(&b_sp[slot], (unsigned long long)sp); atomicAdd(&b_us[slot], (unsigned long long)us); atomicAdd(&b_sum[slot], sm); atomicMin(&b_mn[slot], mn); atomicMax(&b_mx[slot], mx); }
                    else { atomicAdd(T.spliced + t0, (unsigned long long)sp); atomicAdd(T.unspliced + t0, (unsigned long long)us); atomicAdd(T.sumq + t0, sm); atomicMin(T.minq + t0, mn); atomicMax(T.maxq + t0, mx); }
                }
            } else if (vis) {
                if (nN > 0) atomicAdd(T.spliced + tid, 1ull); else atomicAdd(T.unspliced + tid, 1ull);
                atomicAdd(T.sumq + tid, (unsigned long long)(long long)lq); atomicMin(T.minq + tid, lq); atomicMax(T.maxq + tid, lq);
            }
        }
    }
    // ---- pair slots: block scan in record order + look-back ----
    uint32_t off[SE_ITEMS]; uint32_t carry = 0;
#pragma unroll
    for (int r = 0; r < SE_ITEMS; r++) { const uint32_t ex = block_excl_scan(cnt[r], &s_tot); off[r] = carry + ex; carry += s_tot; }
    if (threadIdx.x < 32) {                                            // warp 0: publish the tile aggregate, look back, publish the prefix
        volatile unsigned long long* st = status + tile;
        unsigned long long excl = 0;
        if (tile == 0) { if (lane == 0) *st = (unsigned long long)carry | SE_PREFIX; }
        else {
            if (lane == 0) { *st = (unsigned long long)carry | SE_AGG; __threadfence(); }
            __syncwarp();
            excl = lookback_warp(status, (int64_t)tile, lane);
            if (lane == 0) *st = ((excl + carry) & SE_MASK) | SE_PREFIX;
        }
        if (lane == 0) {
            s_base = excl;
            if (base + SE_TILE >= R.n) *total_pairs = (uint32_t)min(excl + carry, 0xffffffffull);   // the last tile knows the total
        }
    }
    if (threadIdx.x < 4) {   // flush the block's per-target partials
        const int32_t t = tid_first + (int32_t)threadIdx.x;
        if (t >= 0 && t < n_targets && (b_sp[threadIdx.x] | b_us[threadIdx.x])) {
            atomicAdd(T.spliced + t, b_sp[threadIdx.x]); atomicAdd(T.unspliced + t, b_us[threadIdx.x]); atomicAdd(T.sumq + t, b_sum[threadIdx.x]);
            atomicMin(T.minq + t, b_mn[threadIdx.x]); atomicMax(T.maxq + t, b_mx[threadIdx.x]);
        }
    }
    __syncthreads();
    const unsigned long long tile_base = s_base;
    // ---- phase B: JunctionSystem::addJunctions (junction_system.cc:140-210, recursion unrolled) for the spliced records ----
    uint32_t e = 0;
#pragma unroll
    for (int r = 0; r < SE_ITEMS; r++) {
        const uint32_t nN = cnt[r];
        if (nN == 0) continue;
        const int64_t i = base + r * SE_THREADS + threadIdx.x;
        const unsigned long long slot0 = tile_base + off[r];
        if (slot0 + nN > (unsigned long long)pair_cap) { e |= ERR_KEY_OVERFLOW; continue; }
        uint32_t slot = (uint32_t)slot0;
        // every column of the record first: independent loads, one round trip instead of one per use
        const int32_t tid = R.tid[i], pos = R.pos[i];
        const uint32_t flag = R.flag[i];
        const uint32_t cig0 = R.cigar_off[i], cig1 = R.cigar_off[i + 1];
        const uint8_t xs = R.xs[i], mq = R.mapq[i];
        const bool oriented = orientation == PJ_ORIENT_FR || orientation == PJ_ORIENT_RF || orientation == PJ_ORIENT_FF;   // uniform: the mate columns are only read then
        const int32_t mtid = oriented ? R.mtid[i] : -1, mpos = oriented ? R.mpos[i] : -1, lq = R.l_qseq[i];
        const uint64_t so = R.seq_off[i], so1 = R.seq_off[i + 1];
        const int32_t refLen = tlen[tid];
        const uint64_t tbase = toff[tid];
        const uint32_t* cg = R.cigar + cig0;
        const int32_t n = (int32_t)(cig1 - cig0);
        const uint32_t wf = __ldg(cg), wl = __ldg(cg + n - 1);
        uint32_t bits = 0;
        if (flag & 0x40u) bits |= PB_R1;
        if (flag & 0x10u) bits |= PB_REV;
        if (nN > 1) bits |= PB_MS;
        if (mq >= PJ_MAP_QUALITY_THRESHOLD) bits |= PB_UM;
        if (flag & 0x2u) bits |= PB_BPP;
        if (portcullis_proper_pair(flag, tid, mtid, pos, mpos, orientation)) bits |= PB_PPP;
        if (xs == '+') bits |= PB_XSP; else if (xs == '-') bits |= PB_XSN;
        if (flag & FLAG_SEQX) bits |= PB_SEQX;
        // getQuerySeqAfterClipping (bam_alignment.cc:256-264), quirk Q3: only a FIRST / LAST op of type S clips
        int32_t ds = cig_op(wf) == OP_S ? cig_len(wf) : 0; const int32_t de = cig_op(wl) == OP_S ? cig_len(wl) : 0;
        if (ds > lq) ds = lq;
        int64_t qs = (int64_t)lq - ds - de + 1; if (qs > lq - ds) qs = lq - ds; if (qs < 0) qs = 0;
        if (lq > 1 && (int64_t)(so1 - so) < (int64_t)((lq + 3) >> 2)) e |= ERR_SEQ_MISSING;
        int32_t lStart = pos, lEndExc = pos;
        int32_t p = pos, qpos = 0;            // plain reference / query position at the start of op c (calcAlignmentStats / getPadded* walks)
        uint32_t kN = 0, a_eq = 0;            // N ops seen so far; how many of them end exactly at p
        for (int32_t c = 0; c < n; c++) {
            const uint32_t w = __ldg(cg + c), op = cig_op(w); const int32_t L = cig_len(w);
            if (op == OP_N) {
                int32_t rStart = lEndExc + L, rEndExc = rStart;
                int32_t j = c + 1;
                while (j < n && rEndExc <= refLen) {
                    const uint32_t w2 = __ldg(cg + j); if (cig_op(w2) == OP_N) break;
                    if (op_ref(cig_op(w2))) rEndExc += cig_len(w2);
                    j++;
                }
                bool clamped = false;
                if (rStart - 1 >= refLen) { rStart = refLen - 1; clamped = true; }
                if (rEndExc - 1 >= refLen) { rEndExc = refLen; clamped = true; }
                const int32_t start = lEndExc, end = rStart - 1, rendj = rEndExc - 1;
                if (lStart > start || rendj < end) e |= ERR_ANCHOR_ORDER;
                if (start < 0 || start > refLen - 2 || end < start - 1) e |= ERR_START_RANGE;
                const uint64_t sz = (uint64_t)(uint32_t)(end - start + 1);
                if (sz >> len_bits) e |= ERR_KEY_OVERFLOW;
                // nbUpstreamJunctions / nbDownstreamJunctions contribution of this read (junction.cc:795-812): N ops whose end lies
                // before the intron start / beyond its end + 1.  Ends are non-decreasing along the read, so with positive-length
                // N ops and no clamping the counts follow from the op's rank; degenerate CIGARs take the literal loop.
                uint32_t up, down;
                if (zeron[r] || clamped || start != p) {
                    up = 0; down = 0; int32_t pp = pos;
                    for (int32_t k = 0; k < n; k++) {
                        const uint32_t w3 = __ldg(cg + k), o3 = cig_op(w3);
                        if (op_ref(o3)) pp += cig_len(w3);
                        if (o3 == OP_N) { if (pp < start) up++; else if (pp > end + 1) down++; }
                    }
                } else { up = kN - a_eq; down = nN - 1u - kN; }
                keys[slot] = ((tbase + (uint64_t)(uint32_t)start) << len_bits) | sz;
                PairRec* const o = pr + slot;
                o->a = PairA{(uint32_t)i, lStart, rendj, pos};
                o->b = PairB{rend[r], bits, (up << 16) | (down & 0xffffu), start};
                o->c = PairC{so * 4 + (uint64_t)ds, cig0 + (uint32_t)c, qpos};
                o->d = PairD{(int32_t)qs, lq, (uint32_t)c | ((uint32_t)(n - 1 - c) << 16), 0u};
                slot++;
                kN++; p += L; a_eq = L > 0 ? 1u : a_eq + 1u;
                if (j < n) { lStart = rStart; lEndExc = rStart; } else break;
            } else {
                if (op_ref(op)) { lEndExc += L; if (L > 0) { p += L; a_eq = 0; } }
                if (op_query(op)) qpos += L;
            }
        }
    }
    if (e) atomicOr(err, e);
}

static int se_items() { static int v = [] { const char* e = getenv("PJ_SE_ITEMS"); int k = e ? atoi(e) : 4; return (k == 1 || k == 2 || k == 4) ? k : 4; }(); return v; }
uint32_t se_num_tiles(int64_t n) { const int64_t tile = (int64_t)SE_THREADS * se_items(); return (uint32_t)((n + tile - 1) / tile); }
void launch_scan_emit(const Reads& R, const int32_t* tlen, int32_t n_targets, const uint64_t* toff, const uint32_t* max_nlen, int32_t orientation,
                      const TargetAcc& T, uint64_t* keys, PairRec* pr, unsigned long long* status, uint32_t* ticket,
                      uint32_t* total_pairs, uint32_t pair_cap, uint32_t* err, cudaStream_t st) {
    if (R.n <= 0) return;
    const uint32_t nt = se_num_tiles(R.n);
    cudaMemsetAsync(status, 0, (size_t)nt * sizeof(unsigned long long), st);
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
    switch (se_items()) {
    case 1: k_scan_emit<1><<<nt, SE_THREADS, 0, st>>>(R, tlen, n_targets, toff, max_nlen, orientation, T, keys, pr, status, ticket, total_pairs, pair_cap, err); break;
    case 4: k_scan_emit<4><<<nt, SE_THREADS, 0, st>>>(R, tlen, n_targets, toff, max_nlen, orientation, T, keys, pr, status, ticket, total_pairs, pair_cap, err); break;
    default: k_scan_emit<2><<<nt, SE_THREADS, 0, st>>>(R, tlen, n_targets, toff, max_nlen, orientation, T, keys, pr, status, ticket, total_pairs, pair_cap, err); break;
    }
}

// ================================================================================================
// stable LSD radix sort of (key64, val32), 8-bit digits.
// Per pass: block digit histograms -> exclusive scan over [digit][block] -> stable scatter.
// ================================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;       // 4096 keys per block
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t nblocks,
                                                         uint32_t* __restrict__ counts /* [256][nblocks] */) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int k = 0; k < RS_ITEMS; k++) {
        const uint32_t i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[(uint32_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in /* null: iota */,
                                                            uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                            uint32_t n, int shift, uint32_t nblocks, const uint32_t* __restrict__ offs /* scanned counts */) {
    __shared__ uint32_t wcnt[RS_WARPS][256];
    __shared__ uint32_t dbase[256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&wcnt[0][0])[k] = 0;
    __syncthreads();
    // warp w owns keys [base + w*512, base + (w+1)*512), visited in rounds of 32 consecutive keys => stable
    const uint32_t wbase = blockIdx.x * RS_TILE + w * (RS_ITEMS * 32);
    uint64_t key[RS_ITEMS]; uint32_t loc[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[i] : ~0ull;
        const uint32_t d = ok ? ((uint32_t)(key[r] >> shift) & 255u) : 256u;   // 256: padding lanes group together, never counted
        const uint32_t m = __match_any_sync(FULL, d);
        uint32_t prev = 0;
        if (ok) prev = wcnt[w][d];
        __syncwarp();
        if (ok && (m & lt) == 0) wcnt[w][d] = prev + __popc(m);
        __syncwarp();
        loc[r] = prev + __popc(m & lt);
    }
    __syncthreads();
    {   // exclusive prefix over warps for digit = threadIdx.x, plus the global base of this block's digit run
        const uint32_t d = threadIdx.x; uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ww++) { uint32_t t = wcnt[ww][d]; wcnt[ww][d] = run; run += t; }
        dbase[d] = offs[d * nblocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & 255u;
            const uint32_t dst = dbase[d] + wcnt[w][d] + loc[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in ? vals_in[i] : i;
        }
    }
}

uint32_t rs_num_blocks(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts n (key,val) pairs on bits [0, key_bits).  Result ends in (*keys_a, *vals_a) or the alternates; returns 0 if the
// sorted data is in the `a` buffers and 1 if it is in the `b` buffers.
int launch_radix_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int key_bits,
                      uint32_t* counts /* 256*nblocks */, uint32_t* scan_tmp, uint32_t* total_tmp, cudaStream_t st, int* n_launches) {
    if (n == 0) return 0;
    const uint32_t nb = rs_num_blocks(n);
    int cur = 0; bool first = true;
    for (int shift = 0; shift < key_bits; shift += 8) {
        const uint64_t* kin = cur ? keys_b : keys_a; const uint32_t* vin = cur ? vals_b : vals_a;
        uint64_t* kout = cur ? keys_a : keys_b; uint32_t* vout = cur ? vals_a : vals_b;
        k_rs_hist<<<nb, RS_THREADS, 0, st>>>(kin, n, shift, nb, counts);
        launch_exclusive_scan(counts, counts, (uint64_t)256 * nb, scan_tmp, total_tmp, st);
        k_rs_scatter<<<nb, RS_THREADS, 0, st>>>(kin, first ? nullptr : vin, kout, vout, n, shift, nb, counts);
        if (n_launches) *n_launches += 5;
        cur ^= 1; first = false;
    }
    if (first) {   // key_bits == 0: nothing to sort, but vals must still be the identity
        k_rs_hist<<<nb, RS_THREADS, 0, st>>>(keys_a, n, 0, nb, counts);
        launch_exclusive_scan(counts, counts, (uint64_t)256 * nb, scan_tmp, total_tmp, st);
        k_rs_scatter<<<nb, RS_THREADS, 0, st>>>(keys_a, nullptr, keys_b, vals_b, n, 0, nb, counts);
        if (n_launches) *n_launches += 5;
        cur = 1;
    }
    return cur;
}

// ================================================================================================
// One-sweep radix sort: stable LSD, up to 10-bit digits, ONE kernel per digit.
// A single upfront kernel builds the global digit histograms of every pass (the multiset of keys does not change between
// passes).  Each pass then ranks a tile of 8192 keys in shared memory (warp-synchronous match_any ranking, stable) and
// obtains the tile's base offsets with a decoupled look-back over per-tile status words, so keys are read once and
// written once per pass and no separate histogram / scan launches are needed.  Tiles take their index from an atomic
// ticket, which guarantees that every tile a block waits on has already started.
// ================================================================================================
constexpr int OS_THREADS = 512;
#ifndef PJ_OS_ITEMS
#define PJ_OS_ITEMS 16
#endif
constexpr int OS_ITEMS = PJ_OS_ITEMS;
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;       // 8192 keys per tile
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_MAX_BITS = 10;
constexpr int OS_MAX_BINS = 1 << OS_MAX_BITS;
constexpr int OS_MAX_PASSES = 7;
constexpr uint32_t OS_FLAG_AGG = 1u << 30, OS_FLAG_PREFIX = 2u << 30, OS_VALUE_MASK = (1u << 30) - 1u;

struct OsPlan { int npass; int shift[OS_MAX_PASSES]; int bits[OS_MAX_PASSES]; };

__global__ void __launch_bounds__(512) k_os_hist(const uint64_t* __restrict__ keys, uint32_t n, OsPlan plan, uint32_t* __restrict__ ghist /* [npass][1024] */) {
    extern __shared__ uint32_t sh[];                 // [npass][1024]
    for (int k = threadIdx.x; k < plan.npass * OS_MAX_BINS; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = keys[i];
#pragma unroll 4
        for (int p = 0; p < plan.npass; p++) atomicAdd(&sh[p * OS_MAX_BINS + ((uint32_t)(key >> plan.shift[p]) & ((1u << plan.bits[p]) - 1u))], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < plan.npass * OS_MAX_BINS; k += blockDim.x) { const uint32_t v = sh[k]; if (v) atomicAdd(&ghist[k], v); }
}

// exclusive scan of each pass's histogram in place (one block per pass, 1024 threads)
__global__ void __launch_bounds__(1024) k_os_scan_hist(uint32_t* __restrict__ ghist) {
    __shared__ uint32_t tot;
    uint32_t* h = ghist + (size_t)blockIdx.x * OS_MAX_BINS;
    const uint32_t v = h[threadIdx.x];
    const uint32_t ex = block_excl_scan(v, &tot);
    h[threadIdx.x] = ex;
}

__global__ void __launch_bounds__(OS_THREADS, 2) k_os_pass(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in /* null: iota */,
                                                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint32_t n,
                                                         int shift, int bits, const uint32_t* __restrict__ gbase /* [1024] exclusive digit offsets */,
                                                         uint32_t* __restrict__ status /* [ntiles][1024] */, uint32_t* __restrict__ ticket) {
    __shared__ uint16_t wcnt[OS_WARPS][OS_MAX_BINS];
    __shared__ uint32_t dbase[OS_MAX_BINS];
    __shared__ uint32_t s_tile;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t nbins = 1u << bits, dmask = nbins - 1u;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int k = threadIdx.x; k < OS_WARPS * OS_MAX_BINS / 2; k += OS_THREADS) reinterpret_cast<uint32_t*>(&wcnt[0][0])[k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    // warp w owns keys [tile*8192 + w*512, +512), visited in rounds of 32 consecutive keys => stable
    const uint32_t wbase = tile * OS_TILE + w * (OS_ITEMS * 32);
    uint32_t dl[OS_ITEMS];                           // digit << 16 | rank of the key among equal digits of its warp
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = ok ? ((uint32_t)(keys_in[i] >> shift) & dmask) : 0xffffu;   // padding lanes group together, never counted
        const uint32_t m = __match_any_sync(FULL, d);
        uint32_t prev = 0;
        if (ok) prev = wcnt[w][d];
        __syncwarp();
        if (ok && (m & lt) == 0) wcnt[w][d] = (uint16_t)(prev + __popc(m));
        __syncwarp();
        dl[r] = (d << 16) | (prev + __popc(m & lt));
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, tile total, publish + look back
    for (uint32_t d = threadIdx.x; d < nbins; d += OS_THREADS) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < OS_WARPS; ww++) { const uint32_t t = wcnt[ww][d]; wcnt[ww][d] = (uint16_t)run; run += t; }
        volatile uint32_t* st = status + (size_t)tile * OS_MAX_BINS + d;
        if (tile == 0) { *st = run | OS_FLAG_PREFIX; dbase[d] = gbase[d]; }
        else {
            *st = run | OS_FLAG_AGG;
            __threadfence();
            uint32_t excl = 0;
            for (int32_t t = (int32_t)tile - 1; t >= 0; t--) {
                volatile const uint32_t* sp = status + (size_t)t * OS_MAX_BINS + d;
                uint32_t v;
                do { v = *sp; } while ((v >> 30) == 0u);
                excl += v & OS_VALUE_MASK;
                if ((v >> 30) == 2u) break;
            }
            *st = ((excl + run) & OS_VALUE_MASK) | OS_FLAG_PREFIX;
            dbase[d] = gbase[d] + excl;
        }
    }
    __syncthreads();
    // the tile (64 KB of keys) was just read: the second read below is served by L2, and not holding 16 keys in
    // registers across the ranking doubles the resident warps
#pragma unroll
    for (int r = 0; r < OS_ITEMS; r++) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = dl[r] >> 16;
            const uint32_t dst = dbase[d] + wcnt[w][d] + (dl[r] & 0xffffu);
            keys_out[dst] = keys_in[i];
            vals_out[dst] = vals_in ? vals_in[i] : i;
        }
    }
}

uint32_t os_num_tiles(uint32_t n) { return (n + OS_TILE - 1) / OS_TILE; }
// scratch: [npass][1024] histograms + [npass] tickets + [npass][ntiles][1024] status words
size_t os_scratch_words(uint32_t n, int key_bits) {
    const int npass = key_bits <= 0 ? 1 : (key_bits + OS_MAX_BITS - 1) / OS_MAX_BITS;
    return (size_t)npass * OS_MAX_BINS + 64 + (size_t)npass * os_num_tiles(n) * OS_MAX_BINS;
}

// Sorts n (key,val) pairs on bits [0, key_bits), n < 2^30.  Returns 0 if the sorted data ends in the `a` buffers, 1 for `b`.
int launch_onesweep_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int key_bits,
                         uint32_t* scratch, int n_sm, cudaStream_t st, int* n_launches) {
    if (n == 0) return 0;
    OsPlan plan;
    plan.npass = key_bits <= 0 ? 1 : (key_bits + OS_MAX_BITS - 1) / OS_MAX_BITS;
    const int per = key_bits <= 0 ? 1 : (key_bits + plan.npass - 1) / plan.npass;
    for (int p = 0, sh = 0; p < plan.npass; p++, sh += per) { plan.shift[p] = sh; plan.bits[p] = std::max(1, std::min(per, key_bits - sh)); }
    const uint32_t ntiles = os_num_tiles(n);
    uint32_t* ghist = scratch; uint32_t* tickets = scratch + (size_t)plan.npass * OS_MAX_BINS; uint32_t* status = tickets + 64;
    cudaMemsetAsync(scratch, 0, os_scratch_words(n, key_bits) * sizeof(uint32_t), st);
    const int hist_blocks = (int)std::min<uint64_t>((uint64_t)n_sm * 4, ((uint64_t)n + 511) / 512);
    k_os_hist<<<hist_blocks, 512, (size_t)plan.npass * OS_MAX_BINS * sizeof(uint32_t), st>>>(keys_a, n, plan, ghist);
    k_os_scan_hist<<<plan.npass, 1024, 0, st>>>(ghist);
    int cur = 0;
    for (int p = 0; p < plan.npass; p++) {
        const uint64_t* kin = cur ? keys_b : keys_a; const uint32_t* vin = cur ? vals_b : vals_a;
        uint64_t* kout = cur ? keys_a : keys_b; uint32_t* vout = cur ? vals_a : vals_b;
        k_os_pass<<<ntiles, OS_THREADS, 0, st>>>(kin, p == 0 ? nullptr : vin, kout, vout, n, plan.shift[p], plan.bits[p],
                                                 ghist + (size_t)p * OS_MAX_BINS, status + (size_t)p * ntiles * OS_MAX_BINS, tickets + p);
        cur ^= 1;
    }
    if (n_launches) *n_launches += 2 + plan.npass;
    return cur;
}

// ================================================================================================
// junction segmentation
// ================================================================================================
__global__ void __launch_bounds__(256) k_seg_heads(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ head) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
// head[] holds the exclusive scan on entry to this kernel; jid = excl + is_head - 1
__global__ void __launch_bounds__(256) k_seg_ids(const uint64_t* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ excl,
                                                  uint32_t* __restrict__ jid, uint32_t* __restrict__ seg_start, uint32_t n_junc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool h = (i == 0 || keys[i] != keys[i - 1]);
    const uint32_t j = excl[i] + (h ? 1u : 0u) - 1u;
    jid[i] = j;
    if (h) seg_start[j] = i;
    if (i == n - 1) seg_start[n_junc] = n;
}
void launch_seg_heads(const uint64_t* keys, uint32_t n, uint32_t* head, cudaStream_t st) { if (n) k_seg_heads<<<(n + 255) / 256, 256, 0, st>>>(keys, n, head); }
void launch_seg_ids(const uint64_t* keys, uint32_t n, const uint32_t* excl, uint32_t* jid, uint32_t* seg_start, uint32_t n_junc, cudaStream_t st) {
    if (n) k_seg_ids<<<(n + 255) / 256, 256, 0, st>>>(keys, n, excl, jid, seg_start, n_junc);
}

// ================================================================================================
// single-pass flag scan (decoupled look-back, ticketed tiles of 2048 items): the exclusive prefix of a 0/1 flag per
// item is consumed right where it is produced, so segmentation and the entropy compaction are one kernel each instead of
// flag kernel + three scan kernels + consumer kernel.
// ================================================================================================
constexpr int FS_THREADS = 512;
#ifndef PJ_FS_ITEMS
#define PJ_FS_ITEMS 4
#endif
constexpr int FS_ITEMS = PJ_FS_ITEMS;
constexpr int FS_TILE = FS_THREADS * FS_ITEMS;

template <typename F>
__global__ void __launch_bounds__(FS_THREADS) k_flag_scan(uint32_t n, F f, unsigned long long* __restrict__ status, uint32_t* __restrict__ ticket,
                                                          uint32_t* __restrict__ total_out) {
    __shared__ uint32_t s_tile, s_tot, s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t i0 = tile * FS_TILE + threadIdx.x * FS_ITEMS;          // blocked: thread t owns 4 consecutive items
    uint32_t fl[FS_ITEMS]; uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < FS_ITEMS; k++) { fl[k] = (i0 + k < n) ? f.flag(i0 + k) : 0u; sum += fl[k]; }
    uint32_t ex = block_excl_scan(sum, &s_tot);
    if (threadIdx.x < 32) {                                            // warp 0: publish the tile aggregate, look back, publish the prefix
        const int lane = threadIdx.x;
        const uint32_t tot = s_tot;
        volatile unsigned long long* st = status + tile;
        unsigned long long excl = 0;
        if (tile == 0) { if (lane == 0) *st = (unsigned long long)tot | SE_PREFIX; }
        else {
            if (lane == 0) { *st = (unsigned long long)tot | SE_AGG; __threadfence(); }
            __syncwarp();
            excl = lookback_warp(status, (int64_t)tile, lane);
            if (lane == 0) *st = ((excl + tot) & SE_MASK) | SE_PREFIX;
        }
        if (lane == 0) {
            s_base = (uint32_t)excl;
            if ((uint64_t)(tile + 1) * FS_TILE >= n) { *total_out = (uint32_t)(excl + tot); f.finish(n, (uint32_t)(excl + tot)); }
        }
    }
    __syncthreads();
    ex += s_base;
#pragma unroll
    for (int k = 0; k < FS_ITEMS; k++) { if (i0 + k < n) f.emit(i0 + k, ex, fl[k]); ex += fl[k]; }
}

// junction segmentation: flag = first pair of a junction; jid = number of heads up to and including the pair, minus one
struct SegmentOp {
    const uint64_t* keys; uint32_t* jid; uint32_t* seg_start;
    __device__ uint32_t flag(uint32_t i) const { return (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u; }
    __device__ void emit(uint32_t i, uint32_t excl, uint32_t fl) const { const uint32_t j = excl + fl - 1u; jid[i] = j; if (fl) seg_start[j] = i; }
    __device__ void finish(uint32_t n, uint32_t total) const { seg_start[total] = n; }
};
// entropy: flag = emission point of the loop in Junction::calcEntropy; epos = compacted emission indices
struct EntropyIndexOp {
    const uint32_t* eflag; uint32_t* eoff; uint32_t* epos;
    __device__ uint32_t flag(uint32_t i) const { return eflag[i]; }
    __device__ void emit(uint32_t i, uint32_t excl, uint32_t fl) const { eoff[i] = excl; if (fl) epos[excl] = i; }
    __device__ void finish(uint32_t, uint32_t) const {}
};
uint32_t fs_num_tiles(uint32_t n) { return (n + FS_TILE - 1) / FS_TILE; }
// scratch: ntiles status words (8 B each) followed by one 4-byte ticket; zeroed here
void launch_segment(const uint64_t* keys, uint32_t n, uint32_t* jid, uint32_t* seg_start /* n+1 */, uint32_t* n_junc_dev, unsigned long long* scratch, cudaStream_t st) {
    if (!n) return;
    const uint32_t nt = fs_num_tiles(n);
    cudaMemsetAsync(scratch, 0, ((size_t)nt + 1) * 8, st);
    k_flag_scan<<<nt, FS_THREADS, 0, st>>>(n, SegmentOp{keys, jid, seg_start}, scratch, reinterpret_cast<uint32_t*>(scratch + nt), n_junc_dev);
}
void launch_entropy_index(uint32_t n, const uint32_t* eflag, uint32_t* eoff, uint32_t* epos, uint32_t* total_dev, unsigned long long* scratch, cudaStream_t st) {
    if (!n) return;
    const uint32_t nt = fs_num_tiles(n);
    cudaMemsetAsync(scratch, 0, ((size_t)nt + 1) * 8, st);
    k_flag_scan<<<nt, FS_THREADS, 0, st>>>(n, EntropyIndexOp{eflag, eoff, epos}, scratch, reinterpret_cast<uint32_t*>(scratch + nt), total_dev);
}

// ---- batch expansion at submit: prefix columns derived on the device (lean batches), 4-bit SEQ repacked (classic batches) ----
// cigar_off from the per-record op counts
struct CigarOffOp {
    const uint16_t* ncig; uint32_t* out /* cigar_off + R: out[0] is the base already */; uint32_t base; uint32_t expect; unsigned long long* bad;
    __device__ uint32_t flag(uint32_t i) const { return ncig[i]; }
    __device__ void emit(uint32_t i, uint32_t excl, uint32_t fl) const { out[i + 1] = base + excl + fl; }
    __device__ void finish(uint32_t, uint32_t total) const { if (total != expect) *bad = 1ull; }
};
// seq_off of a lean batch: a record has SEQ bytes in the 2-bit stream iff it has an N op and l_qseq > 0 (what the host decoder keeps)
struct SeqOffOp {
    const uint32_t* cigar_off /* + R */; const uint32_t* cigar; const int32_t* lq /* + R */; uint64_t* out /* seq_off + R */; uint64_t base; uint64_t expect; unsigned long long* bad;
    __device__ uint32_t flag(uint32_t i) const {
        const int32_t l = lq[i]; if (l <= 0) return 0u;
        bool spliced = false;
        for (uint32_t c = cigar_off[i]; c < cigar_off[i + 1]; c++) if (cig_op(__ldg(cigar + c)) == OP_N) { spliced = true; break; }
        return spliced ? (uint32_t)((l + 3) >> 2) : 0u;
    }
    __device__ void emit(uint32_t i, uint32_t excl, uint32_t fl) const { out[i + 1] = base + (uint64_t)excl + fl; }
    __device__ void finish(uint32_t, uint32_t total) const { if ((uint64_t)total != expect) *bad = 1ull; }
};
void launch_cigar_off(uint32_t n, const uint16_t* ncig, uint32_t* cigar_off_at_R, uint32_t base, uint32_t expect, unsigned long long* bad, unsigned long long* scratch, cudaStream_t st) {
    if (!n) return;
    const uint32_t nt = fs_num_tiles(n);
    cudaMemsetAsync(scratch, 0, ((size_t)nt + 2) * 8, st);
    k_flag_scan<<<nt, FS_THREADS, 0, st>>>(n, CigarOffOp{ncig, cigar_off_at_R, base, expect, bad}, scratch, reinterpret_cast<uint32_t*>(scratch + nt), reinterpret_cast<uint32_t*>(scratch + nt) + 1);
}
void launch_seq_off(uint32_t n, const uint32_t* cigar_off_at_R, const uint32_t* cigar, const int32_t* lq_at_R, uint64_t* seq_off_at_R, uint64_t base, uint64_t expect,
                    unsigned long long* bad, unsigned long long* scratch, cudaStream_t st) {
    if (!n) return;
    const uint32_t nt = fs_num_tiles(n);
    cudaMemsetAsync(scratch, 0, ((size_t)nt + 2) * 8, st);
    k_flag_scan<<<nt, FS_THREADS, 0, st>>>(n, SeqOffOp{cigar_off_at_R, cigar, lq_at_R, seq_off_at_R, base, expect, bad}, scratch, reinterpret_cast<uint32_t*>(scratch + nt), reinterpret_cast<uint32_t*>(scratch + nt) + 1);
}
size_t fs_scratch_bytes(uint32_t n) { return ((size_t)fs_num_tiles(n) + 2) * 8; }

// Classic batches carry BAM's 4-bit SEQ: repack into the 2-bit stream, one thread per record.  Bases that are not A/C/G/T are
// written as 0, counted per record (pass 1) and listed in record order at the scanned offsets (pass 2).
__device__ __forceinline__ uint32_t nib_code2(uint32_t nib) { return nib == 2u ? 1u : nib == 4u ? 2u : nib == 8u ? 3u : 0u; }   // A(1)->0 C(2)->1 G(4)->2 T(8)->3
__global__ void __launch_bounds__(256) k_seq4_to_2(int64_t n, const uint8_t* __restrict__ s4, const uint64_t* __restrict__ off4 /* n+1, relative to s4 + off4[0] */,
                                                    const int32_t* __restrict__ lq, const uint64_t* __restrict__ seq_off /* arena, + R */, uint8_t* __restrict__ seq2,
                                                    uint16_t* __restrict__ flag /* + R */, uint32_t* __restrict__ xcount) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t o2 = seq_off[i], nb2 = seq_off[i + 1] - o2;
    uint32_t nx = 0;
    if (nb2) {
        const uint8_t* src = s4 + (off4[i] - off4[0]);
        const int32_t l = lq[i];
        for (uint64_t b = 0; b < nb2; b++) {                         // 4 bases = 2 source bytes -> 1 byte
            uint32_t out = 0;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int32_t base = (int32_t)(4 * b) + q;
                if (base < l) {
                    const uint32_t byte = src[base >> 1], nib = (base & 1) ? (byte & 15u) : (byte >> 4);
                    if (nib != 1u && nib != 2u && nib != 4u && nib != 8u) nx++;
                    out |= nib_code2(nib) << (2 * q);
                }
            }
            seq2[o2 + b] = (uint8_t)out;
        }
    }
    xcount[i] = nx;
    if (nx) flag[i] = (uint16_t)(flag[i] | FLAG_SEQX);
}
__global__ void __launch_bounds__(256) k_seq4_exceptions(int64_t n, const uint8_t* __restrict__ s4, const uint64_t* __restrict__ off4, const int32_t* __restrict__ lq,
                                                          const uint64_t* __restrict__ seq_off, const uint32_t* __restrict__ xcount, const uint32_t* __restrict__ xoff,
                                                          uint64_t* __restrict__ xpos /* + first free slot */, uint8_t* __restrict__ xcode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || xcount[i] == 0u) return;
    const uint8_t* src = s4 + (off4[i] - off4[0]);
    const int32_t l = lq[i];
    uint32_t w = xoff[i];
    for (int32_t base = 0; base < l; base++) {
        const uint32_t byte = src[base >> 1], nib = (base & 1) ? (byte & 15u) : (byte >> 4);
        if (nib != 1u && nib != 2u && nib != 4u && nib != 8u) { xpos[w] = seq_off[i] * 4ull + (uint64_t)base; xcode[w] = (uint8_t)nib; w++; }
    }
}
void launch_seq4_to_2(int64_t n, const uint8_t* s4, const uint64_t* off4, const int32_t* lq, const uint64_t* seq_off, uint8_t* seq2, uint16_t* flag, uint32_t* xcount, cudaStream_t st) {
    if (n > 0) k_seq4_to_2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, s4, off4, lq, seq_off, seq2, flag, xcount);
}
void launch_seq4_exceptions(int64_t n, const uint8_t* s4, const uint64_t* off4, const int32_t* lq, const uint64_t* seq_off, const uint32_t* xcount, const uint32_t* xoff,
                            uint64_t* xpos, uint8_t* xcode, cudaStream_t st) {
    if (n > 0) k_seq4_exceptions<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, s4, off4, lq, seq_off, xcount, xoff, xpos, xcode);
}

// ================================================================================================
// warp-shuffle segmented reduction helpers.  Lanes hold nondecreasing junction ids; after the inclusive
// segmented scan the LAST lane of every run holds the reduction of the run.
// ================================================================================================
struct SegCtx { uint32_t same[5]; bool tail; };   // same[s]: lane-2^s exists and is in my segment

__device__ __forceinline__ SegCtx seg_ctx(uint32_t j, int lane) {
    SegCtx c;
#pragma unroll
    for (int s = 0; s < 5; s++) { const uint32_t o = __shfl_up_sync(FULL, j, 1 << s); c.same[s] = (lane >= (1 << s)) && (o == j); }
    const uint32_t nx = __shfl_down_sync(FULL, j, 1);
    c.tail = (lane == 31) || (nx != j);
    return c;
}
template <typename T, typename Op>
__device__ __forceinline__ T seg_reduce(T v, const SegCtx& c, Op op) {
#pragma unroll
    for (int s = 0; s < 5; s++) { const T o = __shfl_up_sync(FULL, v, 1 << s); if (c.same[s]) v = op(v, o); }
    return v;
}
struct OpAdd { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMin { template <typename T> __device__ T operator()(T a, T b) const { return a < b ? a : b; } };
struct OpMax { template <typename T> __device__ T operator()(T a, T b) const { return a > b ? a : b; } };

// ================================================================================================
// k_reduce1: stage-1 reductions per junction (SURVEY §8 "reduction algebra" stage 1)
// ================================================================================================
__global__ void __launch_bounds__(512) k_reduce1(uint32_t n, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ jid,
                                                  const PairRec* __restrict__ pr, int32_t ppcheck,
                                                  JuncAcc A, uint32_t* __restrict__ eflag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool ok = i < n;
    uint32_t j = 0xffffffffu;
    uint32_t c0 = 0, c1 = 0, c2 = 0;                 // packed 6-bit counters
    int32_t lmin = INT32_MAX, rmax = INT32_MIN; uint32_t anc = 0, up = 0, down = 0;
    PairA a = PairA{0u, 0, 0, 0}; PairB b = PairB{0, 0u, 0u, 0};
    uint32_t idx = 0;
    if (ok) { j = jid[i]; idx = vals[i]; a = pr[idx].a; b = pr[idx].b; }
    // (pos, read_end) of the previous pair in sorted order: from the neighbouring lane; only lane 0 has to gather it
    const uint32_t jprev = __shfl_up_sync(FULL, j, 1);
    int32_t ppos = __shfl_up_sync(FULL, a.pos, 1), pend = __shfl_up_sync(FULL, b.read_end, 1);
    if (ok) {
        bool has_prev;
        if (lane == 0) {
            has_prev = i > 0 && jid[i - 1] == j;
            if (has_prev) { const uint32_t ip = vals[i - 1]; ppos = pr[ip].a.pos; pend = pr[ip].b.read_end; }
        } else has_prev = jprev == j;
        const bool last = (i + 1 == n) || (jid[i + 1] != j);
        bool dist = true, newpos = false;
        if (has_prev) {
            dist = (a.pos != ppos) || (b.read_end != pend);
            newpos = a.pos != ppos;
        }
        eflag[i] = (newpos || last) ? 1u : 0u;       // emission points of the entropy loop (quirk Q1)
        const uint32_t bits = b.bits;
        const bool r1 = bits & PB_R1, rev = bits & PB_REV, um = bits & PB_UM, ppp = bits & PB_PPP;
        const bool rel = um && (!ppcheck || ppp);
        c0 = (uint32_t)(r1 && !rev) | ((uint32_t)(r1 && rev) << 6) | ((uint32_t)(!r1 && !rev) << 12) | ((uint32_t)(!r1 && rev) << 18) |
             ((uint32_t)((bits & PB_MS) != 0) << 24);
        c1 = (uint32_t)um | ((uint32_t)((bits & PB_BPP) != 0) << 6) | ((uint32_t)(ppcheck && ppp) << 12) | ((uint32_t)rel << 18) |
             ((uint32_t)((bits & PB_XSP) != 0) << 24);
        c2 = (uint32_t)((bits & PB_XSN) != 0) | ((uint32_t)dist << 6);
        lmin = a.lstart; rmax = a.rend;
        // Intron::minAnchorLength (intron.cc:84-86); the intron end comes from the junction table (k_junc_init)
        anc = (uint32_t)min(b.start - a.lstart, a.rend - A.end[j]);
        up = b.updown >> 16; down = b.updown & 0xffffu;
    }
    const SegCtx sc = seg_ctx(j, lane);
    c0 = seg_reduce(c0, sc, OpAdd()); c1 = seg_reduce(c1, sc, OpAdd()); c2 = seg_reduce(c2, sc, OpAdd());
    lmin = seg_reduce(lmin, sc, OpMin()); rmax = seg_reduce(rmax, sc, OpMax());
    anc = seg_reduce(anc, sc, OpMax()); up = seg_reduce(up, sc, OpMax()); down = seg_reduce(down, sc, OpMax());
    // A block whose pairs all belong to ONE junction (deep junctions span thousands of blocks) first combines its warps in
    // shared memory, so the junction's counters see one atomic per block and field instead of one per warp.
    __shared__ uint32_t part[32][17];
    __shared__ uint32_t s_j0;
    if (threadIdx.x == 0) s_j0 = j;
    __syncthreads();
    const bool uniform = __syncthreads_and(!ok || j == s_j0) != 0;
    if (uniform) {
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        const uint32_t okm = __ballot_sync(FULL, ok);
        if (okm ? (ok && sc.tail) : lane == 31) {      // exactly one writer per warp (an all-invalid warp publishes neutral values)
            uint32_t* p = part[w];
            p[0] = c0 & 63u; p[1] = (c0 >> 6) & 63u; p[2] = (c0 >> 12) & 63u; p[3] = (c0 >> 18) & 63u; p[4] = (c0 >> 24) & 63u;
            p[5] = c1 & 63u; p[6] = (c1 >> 6) & 63u; p[7] = (c1 >> 12) & 63u; p[8] = (c1 >> 18) & 63u; p[9] = (c1 >> 24) & 63u;
            p[10] = c2 & 63u; p[11] = (c2 >> 6) & 63u;
            p[12] = (uint32_t)lmin; p[13] = (uint32_t)rmax; p[14] = anc; p[15] = up; p[16] = down;
        }
        __syncthreads();
        if (w == 0) {
            const bool has = lane < nw;
            uint32_t v[17];
#pragma unroll
            for (int f = 0; f < 17; f++) v[f] = has ? part[lane][f] : (f == 12 ? (uint32_t)INT32_MAX : f == 13 ? (uint32_t)INT32_MIN : 0u);
#pragma unroll
            for (int o = 16; o; o >>= 1) {
#pragma unroll
                for (int f = 0; f < 12; f++) v[f] += __shfl_xor_sync(FULL, v[f], o);
                v[12] = (uint32_t)min((int32_t)v[12], (int32_t)__shfl_xor_sync(FULL, v[12], o));
                v[13] = (uint32_t)max((int32_t)v[13], (int32_t)__shfl_xor_sync(FULL, v[13], o));
                v[14] = max(v[14], __shfl_xor_sync(FULL, v[14], o)); v[15] = max(v[15], __shfl_xor_sync(FULL, v[15], o)); v[16] = max(v[16], __shfl_xor_sync(FULL, v[16], o));
            }
            const uint32_t jj = s_j0;
            if (lane == 0 && jj != 0xffffffffu) {
                uint32_t* const sums[12] = {A.r1p, A.r1n, A.r2p, A.r2n, A.ms, A.um, A.bpp, A.ppp, A.rel, A.xsp, A.xsn, A.dist};
#pragma unroll
                for (int f = 0; f < 12; f++) if (v[f]) atomicAdd(sums[f] + jj, v[f]);
                atomicMin(A.left + jj, (int32_t)v[12]); atomicMax(A.right + jj, (int32_t)v[13]);
                atomicMax(A.anc + jj, v[14]); atomicMax(A.up + jj, v[15]); atomicMax(A.down + jj, v[16]);
            }
        }
        return;
    }
    if (ok && sc.tail) {
        uint32_t v;
        if ((v = c0 & 63u)) atomicAdd(A.r1p + j, v);
        if ((v = (c0 >> 6) & 63u)) atomicAdd(A.r1n + j, v);
        if ((v = (c0 >> 12) & 63u)) atomicAdd(A.r2p + j, v);
        if ((v = (c0 >> 18) & 63u)) atomicAdd(A.r2n + j, v);
        if ((v = (c0 >> 24) & 63u)) atomicAdd(A.ms + j, v);
        if ((v = c1 & 63u)) atomicAdd(A.um + j, v);
        if ((v = (c1 >> 6) & 63u)) atomicAdd(A.bpp + j, v);
        if ((v = (c1 >> 12) & 63u)) atomicAdd(A.ppp + j, v);
        if ((v = (c1 >> 18) & 63u)) atomicAdd(A.rel + j, v);
        if ((v = (c1 >> 24) & 63u)) atomicAdd(A.xsp + j, v);
        if ((v = c2 & 63u)) atomicAdd(A.xsn + j, v);
        if ((v = (c2 >> 6) & 63u)) atomicAdd(A.dist + j, v);
        atomicMin(A.left + j, lmin); atomicMax(A.right + j, rmax);
        atomicMax(A.anc + j, anc); atomicMax(A.up + j, up); atomicMax(A.down + j, down);
    }
}

// Decodes the junction coordinates from the key of each segment head and initialises the accumulators.
__global__ void __launch_bounds__(256) k_junc_init(uint32_t n_junc, const uint32_t* __restrict__ seg_start, const uint64_t* __restrict__ keys,
                                                    const uint32_t* __restrict__ vals, const PairRec* __restrict__ pr,
                                                    const int32_t* __restrict__ read_tid, int32_t len_bits, JuncAcc A) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_junc) return;
    const uint32_t s = seg_start[j];
    const uint64_t key = keys[s];
    const uint32_t idx = vals[s];
    const int32_t start = pr[idx].b.start;
    const int32_t size = (int32_t)(key & ((1ull << len_bits) - 1ull));
    A.tid[j] = read_tid[pr[idx].a.rid];
    A.start[j] = start; A.end[j] = start + size - 1;
    A.left[j] = INT32_MAX; A.right[j] = INT32_MIN;
}

void launch_junc_init(uint32_t n_junc, const uint32_t* seg_start, const uint64_t* keys, const uint32_t* vals, const PairRec* pr,
                      const int32_t* read_tid, int32_t len_bits, const JuncAcc& A, cudaStream_t st) {
    if (n_junc) k_junc_init<<<(n_junc + 255) / 256, 256, 0, st>>>(n_junc, seg_start, keys, vals, pr, read_tid, len_bits, A);
}
void launch_reduce1(uint32_t n, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, int32_t ppcheck,
                    const JuncAcc& A, uint32_t* eflag, cudaStream_t st) {
    if (n) k_reduce1<<<(n + 511) / 512, 512, 0, st>>>(n, vals, jid, pr, ppcheck, A, eflag);
}

// ================================================================================================
// entropy (Junction::calcEntropy, junction.cc:718-749).  eflag marks the loop's emission points; their
// compacted indices give the run-length terms; one thread per junction adds the terms in loop order.
// ================================================================================================
__global__ void __launch_bounds__(256) k_entropy_compact(uint32_t n, const uint32_t* __restrict__ eflag, const uint32_t* __restrict__ eoff,
                                                          uint32_t* __restrict__ epos) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && eflag[i]) epos[eoff[i]] = i;
}
// One warp per junction: lane l evaluates the terms of emission points k0 + 32c + l in parallel, and the warp then adds the 32
// terms of the round ONE BY ONE in loop order (every lane runs the same chain of fp64 additions on shuffled values), so the
// sum is formed exactly as Junction::calcEntropy forms it: sum = (...((0 + t0) + t1) + ...).  A junction has at most one
// emission point per distinct read start, so the sequential part is short (entropy text must not depend on a tree order).
__global__ void __launch_bounds__(256) k_entropy_sum(uint32_t n_junc, const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ eoff,
                                                      const uint32_t* __restrict__ epos, double* __restrict__ entropy) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= n_junc) return;
    const uint32_t s = seg_start[j], e = seg_start[j + 1], n = e - s;
    double sum = 0.0;
    if (n > 1) {
        const uint32_t k0 = eoff[s], k1 = eoff[e - 1];       // the last element of a segment always emits
        for (uint32_t kb = k0; kb <= k1; kb += 32) {
            const uint32_t k = kb + lane;
            double t = 0.0;
            if (k <= k1) {
                const uint32_t i = epos[k];
                const uint32_t term = (k == k0) ? (i - s + 1) : (i - epos[k - 1]);   // elements since the previous emission (quirk Q1)
                const double p = (double)term / (double)n;      // a true division: term == n must give exactly 1.0 (entropy 0)
                t = p * log2(p);
            }
            const int cnt = (int)min(32u, k1 - kb + 1u);
            for (int l = 0; l < cnt; l++) sum += __shfl_sync(FULL, t, l);
        }
    }
    if (lane == 0) entropy[j] = fabs(sum);
}
void launch_entropy_compact(uint32_t n, const uint32_t* eflag, const uint32_t* eoff, uint32_t* epos, cudaStream_t st) {
    if (n) k_entropy_compact<<<(n + 255) / 256, 256, 0, st>>>(n, eflag, eoff, epos);
}
void launch_entropy_sum(uint32_t n_junc, const uint32_t* seg_start, const uint32_t* eoff, const uint32_t* epos, double* entropy, cudaStream_t st) {
    if (n_junc) k_entropy_sum<<<(unsigned)(((uint64_t)n_junc * 32 + 255) / 256), 256, 0, st>>>(n_junc, seg_start, eoff, epos, entropy);
}

// ================================================================================================
// k_match: AlignmentInfo::calcMatchStats (junction.cc:147-240) = getPaddedQuerySeq / getPaddedGenomeSeq
// (bam_alignment.cc:341-462) + hammingDistance + getNbMatchesFromEnd/Start, fused into one walk that never
// materialises the strings.  Quirks Q3-Q6 are kept.
//
// A group of G lanes (G = 1, 2, 4, 8, 16 or 32; 1 on every preset measured) owns one (read, junction) pair.  The CIGAR
// walk is uniform inside the group; the columns of every M/=/X block are compared 32 bases per step: 64 bits of the read's
// 2-bit stream (one unaligned window) XOR the aligned word of the genome's 2-bit plane.
// ================================================================================================

#ifndef PJ_MATCH_CTAS
#define PJ_MATCH_CTAS 5          // resident CTAs per SM the register budget of k_match is set for (measured on B200 with the 2-bit compare: 4 / 5 -> 0.445 / 0.439 ms on c2, 4.96 / 4.74 ms on c5)
#endif
constexpr uint64_t EVEN2 = 0x5555555555555555ull;

// bit k of v -> bit 2k (positions of a 1-bit-per-base mask in the 2-bit-per-base layout)
__device__ __forceinline__ uint64_t spread32(uint32_t v) {
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & EVEN2;
    return x;
}
// ---- per-lane queue of compare blocks (shared memory, one column per thread: conflict-free) ----
// Lanes of a warp reach their M/=/X blocks in different CIGAR iterations; comparing inside the walk would make the
// warp pay the longest block in EVERY iteration.  Instead the walk only queues (read base index, genome index, length,
// string offset, side) and drain() then runs ONE flat loop over 32-base chunks in which every lane is busy.
constexpr int MQ = 4;              // queued blocks per lane before a drain
struct MatchQueue {
    uint64_t qn[MQ][256];          // index of the block's first read base in the shard's SEQ stream
    uint64_t gi[MQ][256];          // global genome base index of its first column
    int32_t  len[MQ][256];
    int32_t  sb[MQ][256];          // (string offset of column 0) << 1 | side
};

struct PairStats { uint32_t mism_l, mism_r; int32_t last_left; int32_t first_right; };

// Character of read base `rb` (base index in the SEQ stream) as the reference's padded query string holds it (bam_alignment.cc:
// 341-462 prints SEQ through "=ACMGRSVTWYHKDBN"): from the exception list when the base is not A/C/G/T.  *ex is a cursor into the
// sorted list that only moves forward (columns are visited in increasing order).
__device__ __forceinline__ uint8_t read_char_exact(const Reads& R, uint64_t rb, int64_t* ex) {
    int64_t e = *ex;
    while (e < R.n_seqx && R.seqx_pos[e] < rb) e++;
    *ex = e;
    if (e < R.n_seqx && R.seqx_pos[e] == rb) return (uint8_t)("=ACMGRSVTWYHKDBN"[R.seqx_code[e] & 15]);
    return (uint8_t)("ACGT"[(R.seq2[rb >> 2] >> (2 * (rb & 3))) & 3]);
}

// Flat loop over the 32-base chunks of all queued blocks.  Chunks are aligned to the genome's 32-base words (one aligned
// 64-bit load); the read bits are extracted unaligned (the stream has a 16-byte lead pad and tail slack).  The left anchor
// tracks its LAST mismatch, the right anchor its FIRST: that is all getNbMatchesFromEnd / getNbMatchesFromStart need.
// A read base stored in the 2-bit stream is A/C/G/T, so it equals the genome character iff the 2-bit codes are equal and the
// genome base is not an exception (gx); reads that do carry other codes (`exact`) are compared character by character.
template <int G>
__device__ __forceinline__ void drain(const MatchQueue& Q, int nq, const Genome& Gn, const Reads& R, bool exact, int gl, PairStats& r) {
    const int col = threadIdx.x;
    if (exact) {                                                                      // rare: the read has N / IUPAC bases
        for (int bi = 0; bi < nq; bi++) {
            const uint64_t gi0 = Q.gi[bi][col], qn0 = Q.qn[bi][col];
            const int32_t len = Q.len[bi][col], sbv = Q.sb[bi][col], sbase = sbv >> 1, side = sbv & 1;
            int64_t ex = 0;
            {   // first exception at or after the block's first base
                int64_t lo = 0, hi = R.n_seqx;
                while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (R.seqx_pos[mid] < qn0) lo = mid + 1; else hi = mid; }
                ex = lo;
            }
            for (int32_t cc = gl; cc < len; cc += G) {
                const uint8_t q = read_char_exact(R, qn0 + (uint64_t)cc, &ex), g = genome_char(Gn, gi0 + (uint64_t)cc);
                if (q != g) {
                    if (side == 0) { r.mism_l++; r.last_left = max(r.last_left, sbase + cc); }
                    else           { r.mism_r++; r.first_right = min(r.first_right, sbase + cc); }
                }
            }
        }
        return;
    }
    int bi = -1; int32_t k = 0, nchunk = 0, a0 = 0, sbase = 0, side = 0;
    uint32_t sh = 0; int32_t tl = 0;
    uint32_t Bl = 0, Bh = 0;                                                          // read word k (G == 1: carried between chunks)
    uint64_t gw0 = 0;                                                                 // index of the block's first genome word (32-base units)
    bool blk_gx = false;                                                              // the block's stretch of the genome has exception bases: consult gx
    const uint2* qw = nullptr;
    for (;;) {
        if (k >= nchunk) {                                                            // next block of this lane
            if (++bi >= nq) break;
            const uint64_t gi0 = Q.gi[bi][col];
            const int32_t len = Q.len[bi][col];
            const int32_t sbv = Q.sb[bi][col]; sbase = sbv >> 1; side = sbv & 1;
            a0 = (int32_t)(gi0 & 31);                                                 // chunks are aligned to the genome words
            nchunk = (a0 + len + 31) >> 5;
            gw0 = (gi0 - (uint64_t)a0) >> 5;
            // the read bases of chunk k are the 64 bits at bit offset sh of the little-endian words qw[k], qw[k + 1]: the offset
            // is the same for every chunk of the block
            const uint64_t qb0 = Q.qn[bi][col] - (uint64_t)a0;
            qw = reinterpret_cast<const uint2*>(R.seq2) + (qb0 >> 5);
            sh = (uint32_t)(qb0 & 31) * 2;
            tl = a0 + len;                                                            // end column of the block in chunk coordinates
            blk_gx = false;
            if (Gn.any_gx) {                                                          // does any 1024-base stretch under the block hold a non-ACGT base?
                for (uint64_t sidx = gi0 >> 10; sidx <= (gi0 + (uint64_t)len - 1) >> 10; sidx++) blk_gx |= ((__ldg(Gn.gsum + (sidx >> 5)) >> (sidx & 31)) & 1u) != 0u;
            }
            k = gl;
            if (k >= nchunk) continue;
            if (G == 1) { const uint2 v = __ldg(qw); Bl = v.x; Bh = v.y; }
        }
        if (G != 1) { const uint2 v = __ldg(qw + k); Bl = v.x; Bh = v.y; }
        const uint2 cv = __ldg(qw + k + 1);                                           // consecutive chunks share a read word: one load per chunk when G == 1
        // 64 bits at bit offset sh of Bl:Bh:Cl:Ch (little-endian), as two 32-bit funnel shifts
        const bool up = (sh & 32u) != 0; const uint32_t s5 = sh & 31u;
        const uint32_t w0 = up ? Bh : Bl, w1 = up ? cv.x : Bh, w2 = up ? cv.y : cv.x;
        const uint64_t x = ((uint64_t)__funnelshift_r(w1, w2, s5) << 32) | (uint64_t)__funnelshift_r(w0, w1, s5);
        if (G == 1) { Bl = cv.x; Bh = cv.y; }
        const uint64_t g = __ldg(Gn.g2 + gw0 + k);
        const int32_t tk = tl - 32 * k;                                               // valid columns from this chunk's start (>= 32: all)
        uint64_t V = EVEN2;
        if (k == 0) V &= ~0ull << (2 * a0);                                           // a0 columns before the block
        if (tk < 32) V &= ~(~0ull << (2 * tk));
        const uint64_t d = x ^ g;
        uint64_t m = (d | (d >> 1)) & V;                                              // one bit (2c) per mismatching column c
        if (blk_gx) {                                                                 // genome bases that are not A/C/G/T never equal a read base stored here
            const uint64_t wi = gw0 + (uint64_t)k;
            const uint32_t e32 = (uint32_t)(__ldg(Gn.gx + (wi >> 1)) >> ((wi & 1) * 32));
            if (e32) m |= spread32(e32) & V;
        }
        // count and extreme positions with predicated arithmetic, not a branch: with 25 lanes and 0.5 % substitutions some lane of
        // the warp has a mismatch in most chunks.  max / min, not assignment: insertions and deletions update the same fields in the walk.
        const uint32_t cnt = (uint32_t)__popcll(m);
        const int32_t cb = sbase + 32 * k - a0;                                       // string offset of this chunk's column 0
        const int32_t lastc = m ? cb + ((63 - __clzll((long long)m)) >> 1) : -1;
        const int32_t firstc = m ? cb + ((__ffsll((long long)m) - 1) >> 1) : INT32_MAX;
        if (side == 0) { r.mism_l += cnt; r.last_left = max(r.last_left, lastc); }
        else           { r.mism_r += cnt; r.first_right = min(r.first_right, firstc); }
        k += G;
    }
}

// One pass over the CIGAR serves both anchor windows: the left walk of the reference stops at the first op it rejects,
// and every op the right walk accepts starts at or after rightStart > leftEnd, so the two walks touch disjoint ops while
// rPos / qPos accumulate identically (bam_alignment.cc:349-399).
template <int G>
__device__ __forceinline__ PairStats walk_pair(MatchQueue& Q, const Genome& Gn, const Reads& R, bool exact, uint64_t gbase, int64_t glen,
                                               const uint32_t* __restrict__ cgn /* this junction's N op */, int32_t ops_before, int32_t ops_after,
                                               int32_t start, int32_t qpos_n, uint64_t seq_b0, int32_t qsize,
                                               int32_t left, int32_t leftEnd, int32_t rightStart, int32_t right, int gl, uint32_t& err,
                                               uint32_t& cols_l, uint32_t& cols_r) {
    auto cigw = [&](int32_t rel) -> uint32_t { return __ldg(cgn + rel); };              // CIGAR word at offset `rel` from this junction's N op
    PairStats r{0u, 0u, -1, INT32_MAX};
    const int col = threadIdx.x;
    // The reference walks the CIGAR from its first op and skips, op by op, everything that starts before the window
    // (bam_alignment.cc:353-357).  Reference positions never decrease along the CIGAR, so the first op it does NOT skip is found
    // by stepping BACK from this junction's N op while the op still starts at or after `left`; the forward walk below then
    // starts there with the same rPos / qPos the reference would have.  A long read is no longer re-walked for each of its introns.
    int32_t rPos = start, qPos = qpos_n, back = 0;
    while (back < ops_before) {
        const uint32_t w = cigw(-(back + 1)), op = cig_op(w); const int32_t L = cig_len(w);
        const int32_t rs = rPos - (op_ref(op) ? L : 0);
        if (rs < left) break;
        rPos = rs; if (op_query(op)) qPos -= L;
        back++;
    }
    const int32_t n_cig = back + 1 + ops_after;
    int side = 0; int32_t wstart = left, wend = leftEnd;
    uint32_t cols = 0; int nq = 0;
    cols_l = 0; cols_r = 0;
    for (int32_t k = 0; k < n_cig; k++) {
        const uint32_t w = cigw(k - back), op = cig_op(w); const int32_t L = cig_len(w);
        const bool cr = op_ref(op), cq = op_query(op);
        if (side == 0 && rPos >= wstart && ((rPos > wend && op != OP_I) || (op == OP_N && rPos + L > wend))) {   // left walk ends here
            side = 1; wstart = rightStart; wend = right; cols_l = cols; cols = 0;
        }
        if (rPos < wstart) { if (cr) rPos += L; if (cq) qPos += L; continue; }          // Q4: op-granular skip
        if ((rPos > wend && op != OP_I) || (op == OP_N && rPos + L > wend)) break;      // Q5 (only reachable with side == 1)
        if (cq) {
            const int32_t len = (rPos + L > wend && op != OP_I) ? wend - rPos + 1 : L;
            if (len == 0) { err |= ERR_ZERO_LEN; break; }
            if (qPos + len > qsize) { err |= ERR_QUERY_RANGE; break; }
            if (op == OP_I) {
                // query bases against 'X' padding in the genome string: never equal (no 'X' in the BAM alphabet)
                if (gl == 0) {
                    if (side == 0) { r.mism_l += (uint32_t)len; r.last_left = max(r.last_left, (int32_t)cols + len - 1); }
                    else           { r.mism_r += (uint32_t)len; r.first_right = min(r.first_right, (int32_t)cols); }
                }
            } else {
                if ((int64_t)rPos + len > glen) { err |= ERR_GENOME_RANGE; break; }
                Q.qn[nq][col] = seq_b0 + (uint64_t)qPos; Q.gi[nq][col] = gbase + (uint64_t)(uint32_t)rPos;
                Q.len[nq][col] = len; Q.sb[nq][col] = (int32_t)(cols << 1) | side;
                if (++nq == MQ) { drain<G>(Q, nq, Gn, R, exact, gl, r); nq = 0; }
            }
            cols += (uint32_t)len;
        } else if (cr) {                                                                   // D or N inside the window: 'X' vs genome
            const int32_t len = rPos + L > wend ? wend - rPos + 1 : L;
            if ((int64_t)rPos + len > glen) { err |= ERR_GENOME_RANGE; break; }
            if (Gn.n_exc_x == 0) {
                if (gl == 0 && len > 0) {
                    if (side == 0) { r.mism_l += (uint32_t)len; r.last_left = max(r.last_left, (int32_t)cols + len - 1); }
                    else           { r.mism_r += (uint32_t)len; r.first_right = min(r.first_right, (int32_t)cols); }
                }
            } else {
                for (int32_t c = gl; c < len; c += G) {
                    if (genome_char(Gn, gbase + (uint64_t)(uint32_t)(rPos + c)) != (uint8_t)'X') {
                        if (side == 0) { r.mism_l++; r.last_left = max(r.last_left, (int32_t)cols + c); }
                        else           { r.mism_r++; r.first_right = min(r.first_right, (int32_t)cols + c); }
                    }
                }
            }
            cols += (uint32_t)len;
        }
        if (cr) rPos += L;
        if (cq) qPos += L;
    }
    drain<G>(Q, nq, Gn, R, exact, gl, r);
    if (side == 0) cols_l = cols; else cols_r = cols;
    if (G > 1) {   // combine the lanes of the group
        const int lane = threadIdx.x & 31;
        const uint32_t gmask = (G == 32) ? FULL : (((1u << G) - 1u) << (lane & ~(G - 1)));
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            r.mism_l += __shfl_xor_sync(gmask, r.mism_l, o); r.mism_r += __shfl_xor_sync(gmask, r.mism_r, o);
            r.last_left = max(r.last_left, __shfl_xor_sync(gmask, r.last_left, o));
            r.first_right = min(r.first_right, __shfl_xor_sync(gmask, r.first_right, o));
        }
    }
    return r;
}

template <int G, int CT /* resident CTAs per SM the register budget is set for */>
__global__ void __launch_bounds__(256, CT) k_match(uint32_t n, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ jid,
                                                const PairRec* __restrict__ pr,
                                                Reads R, Genome Gn, JuncAcc A, uint4* __restrict__ pm, uint32_t* __restrict__ errw) {
    __shared__ MatchQueue Q;
    // Pairs are visited in junction (sorted) order: the lanes of a warp then share the junction-wide window, so their walks have the same
    // shape.  (Emit order — coalesced records, sequential SEQ — was measured: +13 % on c2, +38 % on c5; profiles/r2_history.md.)
    const uint32_t i = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G);
    const int gl = threadIdx.x % G;
    if (i >= n) return;                                              // whole groups leave together
    const uint32_t idx = vals[i];                                    // emit slot of this pair
    const PairA a = pr[idx].a; const PairB b = pr[idx].b; const PairC c = pr[idx].c; const PairD d = pr[idx].d;
    const uint32_t j = jid[i];
    // The walk below is a chain of dependent loads (CIGAR ops -> SEQ words / genome words).  The lines it will need are
    // known already: ask for them now so that they arrive while the junction's window is being loaded.  (Copying the SEQ span and
    // the CIGAR window into shared memory with cp.async instead was measured on B200 and is slower in every variant — the wait
    // replaces the same first-touch latency and the extra shared memory costs resident warps; profiles/r2_history.md.)
    if (G == 1) {
        const uint8_t* sp = R.seq2 + (c.seq_b0 >> 2);
        prefetch_l1(R.cigar + c.cig_abs);
        prefetch_l1(sp);
        if (d.qsize > 400) prefetch_l1(sp + 100);
    }
    const int32_t start = b.start, end = A.end[j], left = A.left[j], right = A.right[j];
    const int32_t tid = A.tid[j];
    if (G == 1) {   // genome words under the two anchors of this read (16 bases per 8-byte word)
        const uint64_t gb = Gn.goff[tid];
        prefetch_l1(Gn.g2 + ((gb + (uint64_t)(uint32_t)max(left, a.pos)) >> 5));
        prefetch_l1(Gn.g2 + ((gb + (uint64_t)(uint32_t)(end + 1)) >> 5));
    }
    const int32_t lq = d.lq;
    uint32_t err = 0, mmes, minMatch, nbMism;
    const int32_t leftEnd = start - 1, rightStart = end + 1;
    if (lq <= 1) {                                                    // junction.cc:168-185
        const uint32_t um = (uint32_t)(leftEnd - left + 1), dm = (uint32_t)(right - rightStart + 1);
        nbMism = 0; minMatch = 0; mmes = min(um, dm);
    } else {
        const int64_t glen = Gn.glen[tid];
        if (left > b.read_end || leftEnd < a.pos || rightStart > b.read_end || right < a.pos) err |= ERR_NO_PRESENCE;
        if (glen < 0) err |= ERR_GENOME_RANGE;
        PairStats St{0u, 0u, -1, INT32_MAX}; uint32_t cols_l = 0, cols_r = 0;
        if (!err) {
            St = walk_pair<G>(Q, Gn, R, (b.bits & PB_SEQX) != 0u, Gn.goff[tid], glen, R.cigar + c.cig_abs, (int32_t)(d.nops & 0xffffu), (int32_t)(d.nops >> 16), start, c.qpos_n,
                             c.seq_b0, d.qsize, left, leftEnd, rightStart, right, gl, err, cols_l, cols_r);
            if (cols_l == 0 || cols_r == 0) err |= ERR_EMPTY_ANCHOR;
        }
        const uint32_t upMatches = cols_l - St.mism_l, downMatches = cols_r - St.mism_r;
        nbMism = St.mism_l + St.mism_r;
        const uint32_t us = St.last_left < 0 ? cols_l : (cols_l - 1u - (uint32_t)St.last_left);            // getNbMatchesFromEnd
        const uint32_t dsm = St.first_right == INT32_MAX ? cols_r : (uint32_t)St.first_right;               // getNbMatchesFromStart
        minMatch = min(us, dsm);
        mmes = min(upMatches, downMatches);
    }
    if (gl == 0) {
        pm[i] = make_uint4(mmes, minMatch, nbMism, 0u);
        if (err) atomicOr(errw, err);
    }
}

template <int G, int CT>
static void launch_match_gc(uint32_t n, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, const Reads& R, const Genome& Gn,
                            const JuncAcc& A, uint4* pm, uint32_t* err, cudaStream_t st) {
    const uint64_t threads = (uint64_t)n * G;
    k_match<G, CT><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, vals, jid, pr, R, Gn, A, pm, err);
}
// ctas: register budget (resident CTAs per SM) of the G == 1 kernel, 0 = default; a tuning knob (PJ_MATCH_CTAS).
void launch_match(uint32_t n, int group, int ctas, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, const Reads& R, const Genome& Gn,
                  const JuncAcc& A, uint4* pm, uint32_t* err, cudaStream_t st) {
    if (!n) return;
    switch (group) {
    case 1:
        if (ctas == 4) launch_match_gc<1, 4>(n, vals, jid, pr, R, Gn, A, pm, err, st);
        else if (ctas == 3) launch_match_gc<1, 3>(n, vals, jid, pr, R, Gn, A, pm, err, st);
        else if (ctas == 6) launch_match_gc<1, 6>(n, vals, jid, pr, R, Gn, A, pm, err, st);
        else launch_match_gc<1, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st);
        break;
    case 2: launch_match_gc<2, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st); break;
    case 4: launch_match_gc<4, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st); break;
    case 8: launch_match_gc<8, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st); break;
    case 16: launch_match_gc<16, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st); break;
    default: launch_match_gc<32, PJ_MATCH_CTAS>(n, vals, jid, pr, R, Gn, A, pm, err, st); break;
    }
}

// ================================================================================================
// k_reduce2: Junction::calcMismatchStats (junction.cc:862-909) as a segmented reduction
// ================================================================================================
__global__ void __launch_bounds__(512) k_reduce2(uint32_t n, const uint32_t* __restrict__ jid, const uint4* __restrict__ pm, JuncAcc A) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool ok = i < n;
    uint32_t j = 0xffffffffu, mmes = 0, mism = 0, firstmm = 0xffffffffu, maxmin = 0, bin = 31;
    if (ok) {
        j = jid[i];
        const uint4 v = pm[i];
        mmes = v.x; mism = v.z; maxmin = v.y;
        if (v.y > 0) firstmm = v.y;
        bin = min(v.y, (uint32_t)PJ_NB_JAD);
    }
    const SegCtx sc = seg_ctx(j, lane);
    mmes = seg_reduce(mmes, sc, OpMax()); mism = seg_reduce(mism, sc, OpAdd());
    firstmm = seg_reduce(firstmm, sc, OpMin()); maxmin = seg_reduce(maxmin, sc, OpMax());
    // single-junction blocks (deep junctions) combine in shared memory first: one atomic per block and field / histogram bin
    __shared__ uint32_t part[32][4];
    __shared__ uint32_t hist[PJ_NB_JAD + 1];
    __shared__ uint32_t s_j0;
    if (threadIdx.x == 0) s_j0 = j;
    if (threadIdx.x <= PJ_NB_JAD) hist[threadIdx.x] = 0;
    __syncthreads();
    const bool uniform = __syncthreads_and(!ok || j == s_j0) != 0;
    if (uniform) {
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        if (ok) atomicAdd(&hist[bin], 1u);
        const uint32_t okm = __ballot_sync(FULL, ok);
        if (okm ? (ok && sc.tail) : lane == 31) { part[w][0] = mmes; part[w][1] = mism; part[w][2] = firstmm; part[w][3] = maxmin; }
        __syncthreads();
        const uint32_t jj = s_j0;
        if (jj == 0xffffffffu) return;
        if (threadIdx.x <= PJ_NB_JAD && hist[threadIdx.x]) atomicAdd(A.jadhist + (size_t)jj * (PJ_NB_JAD + 1) + threadIdx.x, hist[threadIdx.x]);
        if (w == 1) {
            const bool has = lane < nw;
            uint32_t a = has ? part[lane][0] : 0u, b = has ? part[lane][1] : 0u, c = has ? part[lane][2] : 0xffffffffu, d = has ? part[lane][3] : 0u;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                a = max(a, __shfl_xor_sync(FULL, a, o)); b += __shfl_xor_sync(FULL, b, o);
                c = min(c, __shfl_xor_sync(FULL, c, o)); d = max(d, __shfl_xor_sync(FULL, d, o));
            }
            if (lane == 0) { atomicMax(A.maxmmes + jj, a); if (b) atomicAdd(A.mism + jj, b); atomicMin(A.firstmm + jj, c); atomicMax(A.maxminmatch + jj, d); }
        }
        return;
    }
    // JAD histogram: one atomic per distinct (junction, bin) in the warp
    const uint32_t hk = ok ? (j * 32u + bin) : 0xffffffffu;
    const uint32_t m = __match_any_sync(FULL, hk);
    if (ok && (m & ((1u << lane) - 1u)) == 0) atomicAdd(A.jadhist + (size_t)j * (PJ_NB_JAD + 1) + bin, (uint32_t)__popc(m));
    if (ok && sc.tail) {
        atomicMax(A.maxmmes + j, mmes);
        if (mism) atomicAdd(A.mism + j, mism);
        atomicMin(A.firstmm + j, firstmm);
        atomicMax(A.maxminmatch + j, maxmin);
    }
}
void launch_reduce2(uint32_t n, const uint32_t* jid, const uint4* pm, const JuncAcc& A, cudaStream_t st) {
    if (n) k_reduce2<<<(n + 511) / 512, 512, 0, st>>>(n, jid, pm, A);
}

// ================================================================================================
// k_finalize: per junction strands (junction.cc:531-559), motif (:504-516, 289-326), Hamming (:823-857),
// suspicious flag (:897-908), JAD suffix sums, and the output row.
// ================================================================================================
__global__ void __launch_bounds__(128) k_finalize(uint32_t n_junc, const uint32_t* __restrict__ seg_start, JuncAcc A, Genome G,
                                                   const double* __restrict__ entropy, pj_junction* __restrict__ rows, uint32_t* __restrict__ errw) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_junc) return;
    pj_junction o;
    memset(&o, 0, sizeof o);
    const int32_t tid = A.tid[j], s = A.start[j], e = A.end[j], left = A.left[j], right = A.right[j];
    const uint32_t n = seg_start[j + 1] - seg_start[j];
    o.tid = tid; o.start = s; o.end = e; o.left = left; o.right = right;
    o.nb_raw_aln = n; o.nb_dist_aln = A.dist[j]; o.nb_ms_aln = A.ms[j]; o.nb_um_aln = A.um[j]; o.nb_bpp_aln = A.bpp[j];
    o.nb_ppp_aln = A.ppp[j]; o.nb_rel_aln = A.rel[j];
    o.nb_r1_pos = A.r1p[j]; o.nb_r1_neg = A.r1n[j]; o.nb_r2_pos = A.r2p[j]; o.nb_r2_neg = A.r2n[j];
    o.nb_xs_pos = A.xsp[j]; o.nb_xs_neg = A.xsn[j];
    o.max_min_anc = A.anc[j]; o.maxmmes = A.maxmmes[j]; o.nb_mismatches = A.mism[j];
    o.nb_up_juncs = A.up[j]; o.nb_down_juncs = A.down[j];
    o.entropy = entropy[j];
    // determineStrandFromReads: 0.95 rule in fp64 exactly as the reference evaluates it
    const double tot = (double)n;
    uint8_t rs = PJ_STRAND_UNKNOWN;
    if ((double)o.nb_xs_pos / tot >= 0.95) rs = PJ_STRAND_POS; else if ((double)o.nb_xs_neg / tot >= 0.95) rs = PJ_STRAND_NEG;
    o.read_strand = rs;
    // JAD_k = #reads with minMatch >= k  (suffix sums of the minMatch histogram)
    {
        const uint32_t* h = A.jadhist + (size_t)j * (PJ_NB_JAD + 1);
        uint32_t run = h[PJ_NB_JAD];
        for (int k = PJ_NB_JAD - 1; k >= 0; k--) { o.jad[k] = run; run += h[k]; }
    }
    {
        const uint32_t fm = A.firstmm[j] == 0xffffffffu ? 100000000u : A.firstmm[j];
        o.suspicious = (o.nb_mismatches > 0 && fm < 20 && !(A.maxminmatch[j] > fm)) ? 1 : 0;
    }
    uint32_t err = 0;
    const int64_t glen = G.glen[tid];
    o.hamming5p = 10; o.hamming3p = 10; o.canonical_ss = 'N'; o.ss_strand = PJ_STRAND_UNKNOWN; o.consensus_strand = rs;
    if (glen < 0 || s < 0 || (int64_t)s + 9 >= glen || e - 9 < 0 || e >= glen || left < 0 || right >= glen || left >= s || right <= e) {
        err |= ERR_GENOME_RANGE;                      // the reference throws in processJunctionWindow (junction.cc:573-633)
    } else {
        const uint64_t gb = G.goff[tid];
        const uint8_t d0 = genome_char(G, gb + s), d1 = genome_char(G, gb + s + 1);
        const uint8_t a0 = genome_char(G, gb + e - 1), a1 = genome_char(G, gb + e);
        uint8_t css = 'N', ss = PJ_STRAND_UNKNOWN;
        const uint32_t m = ((uint32_t)d0 << 24) | ((uint32_t)d1 << 16) | ((uint32_t)a0 << 8) | a1;
        switch (m) {
        case 0x47544147u: css = 'C'; ss = PJ_STRAND_POS; break;     // GTAG
        case 0x43544143u: css = 'C'; ss = PJ_STRAND_NEG; break;     // CTAC
        case 0x41544143u: css = 'S'; ss = PJ_STRAND_POS; break;     // ATAC
        case 0x47434147u: css = 'S'; ss = PJ_STRAND_POS; break;     // GCAG
        case 0x47544154u: css = 'S'; ss = PJ_STRAND_NEG; break;     // GTAT
        case 0x43544743u: css = 'S'; ss = PJ_STRAND_NEG; break;     // CTGC
        default: break;
        }
        o.canonical_ss = css; o.ss_strand = ss;
        const uint8_t cs = rs == ss ? rs : rs == PJ_STRAND_UNKNOWN ? ss : ss == PJ_STRAND_UNKNOWN ? rs : (uint8_t)PJ_STRAND_UNKNOWN;
        o.consensus_strand = cs;
        if (cs == PJ_STRAND_NEG) { o.ss1[0] = revcomp_char(a1); o.ss1[1] = revcomp_char(a0); o.ss2[0] = revcomp_char(d1); o.ss2[1] = revcomp_char(d0); }
        else { o.ss1[0] = d0; o.ss1[1] = d1; o.ss2[0] = a0; o.ss2[1] = a1; }
        // Hamming: la = last <=10 bases of the left anchor vs ri = first |la| bases of [e-9,e];
        //          ra = first <=10 bases of the right anchor vs li = first |ra| bases of [s,s+9]   (quirk Q7)
        const int32_t lan = min(10, s - left), ran = min(10, right - e);
        uint32_t hl = 0, hr = 0;
        for (int32_t k = 0; k < lan; k++) {
            uint8_t x = genome_char(G, gb + (s - lan + k)), y = genome_char(G, gb + (e - 9 + k));
            if (cs == PJ_STRAND_NEG) { x = revcomp_char(x); y = revcomp_char(y); }
            hl += x != y;
        }
        for (int32_t k = 0; k < ran; k++) {
            uint8_t x = genome_char(G, gb + (e + 1 + k)), y = genome_char(G, gb + (s + k));
            if (cs == PJ_STRAND_NEG) { x = revcomp_char(x); y = revcomp_char(y); }
            hr += x != y;
        }
        if (cs == PJ_STRAND_NEG) { o.hamming5p = hr; o.hamming3p = hl; } else { o.hamming5p = hl; o.hamming3p = hr; }
    }
    rows[j] = o;
    if (err) atomicOr(errw, err);
}
void launch_finalize(uint32_t n_junc, const uint32_t* seg_start, const JuncAcc& A, const Genome& G, const double* entropy,
                     pj_junction* rows, uint32_t* err, cudaStream_t st) {
    if (n_junc) k_finalize<<<(n_junc + 127) / 128, 128, 0, st>>>(n_junc, seg_start, A, G, entropy, rows, err);
}

} // namespace pjk
