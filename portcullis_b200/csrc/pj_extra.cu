// pj_extra.cu — the hidden `--extra` metrics of `junc` on the columnar records already in HBM.
//
// Reference: JunctionBuilder::separateBams + calcExtraMetrics (/root/reference/src/junction_builder.cc:152-226, 293-312)
//   mm_score          Junction::calcMultipleMappingScore   lib/src/junction.cc:914-921   (name map: junction_builder.cc:179-186)
//   up_aln / down_aln Junction::processJunctionVicinity     lib/src/junction.cc:651-677   (region query on unspliced.bam)
//   coverage          Junction::calcCoverage                lib/src/junction.cc:923-951   (DepthParser, lib/src/depth_parser.cc:112-167)
//
// The reference re-reads two BAM files it has just written; here the same record classes are views of the shard arena:
//   spliced   = any N op (bam_alignment.cc:294-301); unspliced = not spliced and mapped (junction_builder.cc:188-191).
// Kernels (all HBM-bound integer work, grids sized from the data):
//   k_x_classify    one thread per record: class, reference span, spliced name codes compacted with warp-aggregated atomics
//   k_x_scatter     unspliced records compacted in BAM order (offsets from the library's exclusive scan)
//   k_x_names_add   open-addressing table name code -> number of spliced alignments (atomicCAS + atomicAdd)
//   k_x_mm          one thread per (read, junction) pair: table lookup, integer atomics into mm_n / mm_m
//   k_x_flank       one warp per junction: binary searches into the unspliced start positions, lanes stride the window
//   k_x_live/k_x_depth  +1/-1 difference arrays per target, turned into depth vectors by the exclusive scan
//   k_x_max         per-target maximum of the live-read vector (htslib's 8000-read cap binds iff it reaches 8000)
//   k_x_coverage    one thread per (junction, window): the four read-count sums of Junction::calcCoverage
#include "pj_ctx.hpp"
#include <climits>
#include <cstring>

using namespace pjk;
using namespace pjapi;

namespace {

constexpr uint64_t X_EMPTY = ~0ull;
constexpr uint32_t XERR_NOCIGAR = 1u;      // mapped unspliced record without CIGAR: htslib's pileup asserts (sam.c:1537)
constexpr uint32_t XERR_UNSORTED = 2u;     // unspliced records not in (tid, pos) order

__device__ __forceinline__ uint64_t x_mix(uint64_t h) { h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33; return h; }
__device__ __forceinline__ uint64_t x_key(uint64_t code) { return code == X_EMPTY ? X_EMPTY - 1 : code; }

__global__ void __launch_bounds__(256) k_x_classify(int64_t n, const uint32_t* __restrict__ cigar_off, const uint32_t* __restrict__ cigar,
                                                     const uint16_t* __restrict__ flag, const int32_t* __restrict__ tid, const uint64_t* __restrict__ name_code,
                                                     int32_t n_targets, uint32_t* __restrict__ uflag, int32_t* __restrict__ alen_out,
                                                     uint64_t* __restrict__ names_out, unsigned long long* __restrict__ n_spliced, uint32_t* __restrict__ err) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool spliced = false; bool uns = false; int32_t alen = 0;
    if (r < n) {
        const uint32_t c0 = cigar_off[r], c1 = cigar_off[r + 1];
        for (uint32_t k = c0; k < c1; k++) {
            const uint32_t w = __ldg(cigar + k); const uint32_t op = w & 15u;
            if (op == 3u) spliced = true;
            if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) alen += (int32_t)(w >> 4);      // bam_alignment.cc:78-88
        }
        const int32_t t = tid[r];
        uns = !spliced && !(flag[r] & 0x4u) && t >= 0 && t < n_targets;
        if (uns && c1 == c0) { atomicOr(err, XERR_NOCIGAR); uns = false; }
        uflag[r] = uns ? 1u : 0u;
        alen_out[r] = alen;
    }
    const unsigned m = __ballot_sync(0xffffffffu, spliced);
    if (spliced) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(n_spliced, (unsigned long long)__popc(m));
        base = __shfl_sync(m, base, leader);
        names_out[base + __popc(m & ((1u << lane) - 1u))] = name_code[r];
    }
}

__global__ void __launch_bounds__(256) k_x_scatter(int64_t n, const uint32_t* __restrict__ uflag, const uint32_t* __restrict__ uoff,
                                                    const int32_t* __restrict__ tid, const int32_t* __restrict__ pos, const int32_t* __restrict__ alen,
                                                    int32_t* __restrict__ u_tid, int32_t* __restrict__ u_pos, int32_t* __restrict__ u_alen, uint32_t* __restrict__ u_rid,
                                                    int32_t* __restrict__ maxspan, uint8_t* __restrict__ covered) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n || !uflag[r]) return;
    const uint32_t i = uoff[r];
    const int32_t t = tid[r], a = alen[r];
    u_tid[i] = t; u_pos[i] = pos[r]; u_alen[i] = a; u_rid[i] = (uint32_t)r;
    if (a > maxspan[t]) atomicMax(maxspan + t, a);      // racy pre-check only skips atomics that cannot raise the maximum
    if (a > 0) covered[t] = 1;
}

// u_toff[t] = first unspliced index of target t (T + 1 entries); also checks the (tid, pos) order the searches rely on
__global__ void __launch_bounds__(256) k_x_toff(uint32_t U, const int32_t* __restrict__ u_tid, const int32_t* __restrict__ u_pos, int32_t T,
                                                 uint32_t* __restrict__ u_toff, uint32_t* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > U) return;
    const int32_t prev = i == 0 ? -1 : u_tid[i - 1];
    const int32_t cur = i == U ? T : u_tid[i];
    if (cur < prev || (i > 0 && i < U && cur == prev && u_pos[i] < u_pos[i - 1])) { atomicOr(err, XERR_UNSORTED); return; }
    for (int32_t t = prev + 1; t <= cur; t++) u_toff[t] = i;
}

__global__ void __launch_bounds__(256) k_x_names_add(int64_t n, const uint64_t* __restrict__ codes, uint64_t* __restrict__ keys, uint32_t* __restrict__ counts, uint64_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = x_key(codes[i]);
    uint64_t s = x_mix(k) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(keys + s), (unsigned long long)X_EMPTY, (unsigned long long)k);
        if (old == X_EMPTY || old == k) { atomicAdd(counts + s, 1u); return; }
        s = (s + 1) & mask;
    }
}

__global__ void __launch_bounds__(256) k_x_mm(uint32_t P, const uint32_t* __restrict__ pair_rid, const uint32_t* __restrict__ pair_jid, const uint64_t* __restrict__ name_code,
                                               const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t mask, pj_junction_extra* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint64_t k = x_key(name_code[pair_rid[i]]);
    uint64_t s = x_mix(k) & mask;
    uint32_t cnt = 0;
    for (;;) {
        const uint64_t cur = keys[s];
        if (cur == k) { cnt = counts[s] & 0xffffu; break; }       // map values are uint16_t (junction.hpp:38)
        if (cur == X_EMPTY) break;                                 // operator[] on a missing name inserts 0
        s = (s + 1) & mask;
    }
    // pairs are sorted by junction: lanes of one junction combine first, so a junction with 100 000 alignments costs
    // thousands of atomics on its row, not hundreds of thousands
    const uint32_t j = pair_jid[i];
    const unsigned peers = __match_any_sync(__activemask(), j);
    const uint32_t total = __reduce_add_sync(peers, cnt);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) {
        pj_junction_extra* o = out + j;
        atomicAdd(&o->mm_n, (uint32_t)__popc(peers));
        atomicAdd(&o->mm_m, total);                                // uint32 sum, wraps like `uint32_t M` (junction.cc:916)
    }
}

__device__ __forceinline__ uint32_t x_lower(const int32_t* __restrict__ a, uint32_t lo, uint32_t hi, int64_t v) {     // first index with a[i] >= v
    while (lo < hi) { const uint32_t m = lo + ((hi - lo) >> 1); if ((int64_t)a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

// Junction::processJunctionVicinity (junction.cc:651-677): one warp per junction.  The region query is htslib's
// iterator (hts.c:1944-1960): records of the target with pos < regionEnd and bam_endpos > regionStart.
__global__ void __launch_bounds__(256) k_x_flank(uint32_t J, const pj_junction* __restrict__ rows, const int32_t* __restrict__ tlen, const uint32_t* __restrict__ u_toff,
                                                  const int32_t* __restrict__ u_pos, const int32_t* __restrict__ u_alen, const int32_t* __restrict__ maxspan,
                                                  int32_t max_query_length, pj_junction_extra* __restrict__ out) {
    const uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
    if (j >= J) return;
    const pj_junction* q = rows + j;
    const int32_t t = q->tid, start = q->start, end = q->end, left = q->left, right = q->right;
    const int32_t refLength = tlen[t];
    int32_t regionStart = left - max_query_length - 1; regionStart = regionStart < 0 ? 0 : regionStart;
    int32_t regionEnd = right + max_query_length + 1; regionEnd = regionEnd >= refLength ? refLength - 1 : regionEnd;
    uint32_t nl = 0, nr = 0;
    if (regionEnd > regionStart) {
        const uint32_t lo = u_toff[t], hi = u_toff[t + 1];
        const uint32_t stop = x_lower(u_pos, lo, hi, regionEnd);                       // pos >= regionEnd ends the iteration
        // left flank: pos < intron start and getEnd() >= leftAncStart; no read further left than the longest span can reach it
        uint32_t a = x_lower(u_pos, lo, stop, (int64_t)left - (int64_t)maxspan[t]);
        uint32_t z = x_lower(u_pos, a, stop, start);
        for (uint32_t i = a + lane; i < z; i += 32) {
            const int32_t pos = u_pos[i], al = u_alen[i];
            const int64_t endpos = (int64_t)pos + al;                                  // bam_endpos: mapped, n_cigar > 0
            const int32_t getEnd = pos + al - 1;                                       // bam_alignment.hpp:221-223
            if (endpos > regionStart && start > pos && left <= getEnd) nl++;
        }
        // right flank: intron end < pos <= rightAncEnd
        a = x_lower(u_pos, lo, stop, (int64_t)end + 1);
        z = x_lower(u_pos, a, stop, (int64_t)right + 1);
        for (uint32_t i = a + lane; i < z; i += 32) {
            const int32_t pos = u_pos[i];
            const int64_t endpos = (int64_t)pos + u_alen[i];
            if (endpos > regionStart && right >= pos && end < pos) nr++;
        }
    }
    nl = __reduce_add_sync(0xffffffffu, nl); nr = __reduce_add_sync(0xffffffffu, nr);
    if (lane == 0) { out[j].up_aln = nl; out[j].down_aln = nr; }
}

// live-read vector: +1 at pos, -1 after the last base a node stays allocated for (bam_plp_next frees a node once the
// iterator has passed its end, so a read is "live" on [pos, endpos] inclusive)
__global__ void __launch_bounds__(256) k_x_live(uint32_t U, const int32_t* __restrict__ u_tid, const int32_t* __restrict__ u_pos, const int32_t* __restrict__ u_alen,
                                                 const int32_t* __restrict__ tlen, const uint64_t* __restrict__ doff, uint32_t* __restrict__ diff) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    const int32_t t = u_tid[i]; const int64_t L = tlen[t]; const int64_t p = u_pos[i];
    if (p < 0 || p >= L) return;
    uint32_t* d = diff + doff[t];
    int64_t e = p + (int64_t)u_alen[i] + 1; if (e > L) e = L;
    // reads are in position order: the lanes of a warp mostly share their start (and, on a pile-up, their end) — one atomic per
    // distinct address instead of one per read
    const unsigned act = __activemask();
    const unsigned ps = __match_any_sync(act, (unsigned long long)(uintptr_t)(d + p));
    if ((int)(threadIdx.x & 31) == __ffs(ps) - 1) atomicAdd(d + p, (uint32_t)__popc(ps));
    const unsigned es = __match_any_sync(act, (unsigned long long)(uintptr_t)(d + e));
    if ((int)(threadIdx.x & 31) == __ffs(es) - 1) atomicAdd(d + e, 0u - (uint32_t)__popc(es));
}

// depth vector of DepthParser (depth_parser.cc:121-156): reads count on M/=/X columns only (is_del / is_refskip are
// subtracted), stored with the reference's `rpos = pos + 1` shift — the exclusive scan supplies the shift
__global__ void __launch_bounds__(256) k_x_depth(uint32_t U, const int32_t* __restrict__ u_tid, const int32_t* __restrict__ u_pos, const uint32_t* __restrict__ u_rid,
                                                  const uint32_t* __restrict__ cigar_off, const uint32_t* __restrict__ cigar, const uint8_t* __restrict__ accepted,
                                                  const int32_t* __restrict__ tlen, const uint64_t* __restrict__ doff, uint32_t* __restrict__ diff) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    if (accepted && !accepted[i]) return;               // dropped by the pileup's read cap (k_x_cap)
    const int32_t t = u_tid[i]; const int64_t L = tlen[t];
    uint32_t* d = diff + doff[t];
    const uint32_t r = u_rid[i];
    int64_t x = u_pos[i];
    for (uint32_t k = cigar_off[r]; k < cigar_off[r + 1]; k++) {
        const uint32_t w = __ldg(cigar + k); const uint32_t op = w & 15u; const int64_t n = w >> 4;
        if (op == 0u || op == 7u || op == 8u) {
            int64_t a = x < 0 ? 0 : x, b = x + n; if (b > L) b = L;
            if (a < b) { atomicAdd(d + a, 1u); atomicAdd(d + b, 0xffffffffu); }
            x += n;
        } else if (op == 2u || op == 3u) x += n;
    }
}

constexpr int32_t PLP_MAXCNT = 8000;       // bam_plp_init (htslib-1.3 sam.c:1622)

// hot[i] = 1 when read i starts on a column whose live-read count (reads with pos <= p <= endpos, cap ignored) reaches the
// cap: only there can bam_plp_push drop reads (sam.c:1906).  `live` is the exclusive scan of k_x_live: L(p) = live[p + 1].
__global__ void __launch_bounds__(256) k_x_hot(uint32_t U, const int32_t* __restrict__ u_tid, const int32_t* __restrict__ u_pos, const int32_t* __restrict__ tlen,
                                                const uint64_t* __restrict__ doff, const uint32_t* __restrict__ live, uint32_t* __restrict__ hot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    const int32_t t = u_tid[i]; const int64_t p = u_pos[i];
    hot[i] = (p >= 0 && p < tlen[t] && live[doff[t] + p + 1] >= (uint32_t)PLP_MAXCNT) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_x_compact(uint32_t U, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ off, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < U && flag[i]) out[off[i]] = i;
}

// htslib's pileup buffer cap, replayed (bam_plp_push / bam_plp_next, sam.c:1838-1936).  A read that starts on the column
// the iterator stands on is dropped when the node pool already holds more than 8000 nodes: pool = live reads + the spare
// tail + the dummy node (sam.c:1619-1620).  With `live` = accepted reads whose end is >= P when the first read of column
// P has just been pushed, the k-th read of the column (k >= 1) is kept iff 2 + live + k <= 8000, the first one always.
// The process is sequential along a target, but it can only drop reads on "hot" columns (k_x_hot), and its state equals
// the uncapped one again once every read that started on a hot column has ended.  One warp per target walks those
// regions column by column: `endcnt` (the target's slice of the zeroed depth buffer, used as scratch and left zeroed)
// counts accepted reads per end position, lanes share the per-column work.
__global__ void __launch_bounds__(32) k_x_cap(int32_t T, const uint32_t* __restrict__ u_toff, const int32_t* __restrict__ u_pos, const int32_t* __restrict__ u_alen,
                                               const int32_t* __restrict__ maxspan_t, const int32_t* __restrict__ tlen, const uint64_t* __restrict__ doff,
                                               const uint32_t* __restrict__ H, uint32_t nH, uint32_t* __restrict__ scratch, uint8_t* __restrict__ accepted) {
    const int32_t t = blockIdx.x; const int lane = threadIdx.x;
    if (t >= T) return;
    const uint32_t lo = u_toff[t], hi = u_toff[t + 1];
    uint32_t hlo = 0, hhi = nH;
    { uint32_t a = 0, z = nH; while (a < z) { const uint32_t m = a + ((z - a) >> 1); if (H[m] < lo) a = m + 1; else z = m; } hlo = a; }
    { uint32_t a = hlo, z = nH; while (a < z) { const uint32_t m = a + ((z - a) >> 1); if (H[m] < hi) a = m + 1; else z = m; } hhi = a; }
    if (hlo == hhi) return;
    uint32_t* endcnt = scratch + doff[t];
    const int64_t L = tlen[t]; const int64_t span = maxspan_t[t];
    uint32_t h = hlo;
    while (h < hhi) {
        const uint32_t start = H[h];
        const int64_t P0 = u_pos[start];
        // state at the start of the region: every earlier read is accepted (no hot column within reach)
        int64_t live = 0, maxE = P0;
        {
            const uint32_t a0 = x_lower(u_pos, lo, start, P0 - span);
            uint32_t c = 0; int64_t me = P0;
            for (uint32_t j = a0 + lane; j < start; j += 32) {
                const int64_t E = (int64_t)u_pos[j] + u_alen[j];
                if (E >= P0) { c++; atomicAdd(endcnt + (E < L ? E : L), 1u); if (E > me) me = E; }
            }
            live = __reduce_add_sync(0xffffffffu, c);
            maxE = __reduce_max_sync(0xffffffffu, (unsigned)(me > 0x7fffffff ? 0x7fffffff : me));
        }
        __syncwarp();
        uint32_t i = start; int64_t prevP = P0, lastHot = P0;
        while (i < hi) {
            const int64_t P = u_pos[i];
            if (P > lastHot + span) break;
            if (P > prevP) {                                   // nodes whose end has been passed are freed (end <= P - 1)
                uint32_t s = 0;
                for (int64_t e = prevP + lane; e < P; e += 32) if (e <= L) s += __ldcg(endcnt + e);       // L2 read: the counts are built with atomics
                live -= __reduce_add_sync(0xffffffffu, s);
            }
            uint32_t n = 0;                                    // reads on this column
            for (;;) {
                const uint32_t j = i + n + lane;
                const unsigned same = __ballot_sync(0xffffffffu, j < hi && u_pos[j] == P);
                if (same == 0xffffffffu) { n += 32; continue; }
                n += __ffs(~same) - 1; break;
            }
            const bool is_hot = h < hhi && H[h] == i;
            uint32_t keep = n;
            if (is_hot) { const int64_t room = (int64_t)(PLP_MAXCNT - 1) - live; keep = (uint32_t)(room < 1 ? 1 : room > (int64_t)n ? (int64_t)n : room); h += n; }
            bool zero_len = false;
            if (is_hot) { for (uint32_t k = lane; k < n; k += 32) zero_len |= u_alen[i + k] == 0; zero_len = __any_sync(0xffffffffu, zero_len); }
            if (is_hot && zero_len) {                          // a read without reference span gets no node unless it opens the column: replay one by one
                if (lane == 0) {
                    int64_t lv = live;
                    for (uint32_t k = 0; k < n; k++) {
                        const int64_t E = P + u_alen[i + k];
                        const bool ok = k == 0 || 2 + lv <= PLP_MAXCNT;
                        accepted[i + k] = ok ? 1 : 0;
                        if (ok && (k == 0 ? E >= P : E > P)) { atomicAdd(endcnt + (E < L ? E : L), 1u); lv++; if (E > maxE) maxE = E; }
                    }
                    live = lv;
                }
                live = __shfl_sync(0xffffffffu, live, 0); maxE = __shfl_sync(0xffffffffu, maxE, 0);
            } else {
                uint32_t c = 0; int64_t me = maxE;
                for (uint32_t k = lane; k < n; k += 32) {
                    if (k >= keep) { accepted[i + k] = 0; continue; }
                    const int64_t E = P + u_alen[i + k];
                    if (k == 0 ? E >= P : E > P) { c++; atomicAdd(endcnt + (E < L ? E : L), 1u); if (E > me) me = E; }
                }
                live += __reduce_add_sync(0xffffffffu, c);
                maxE = __reduce_max_sync(0xffffffffu, (unsigned)(me > 0x7fffffff ? 0x7fffffff : me));
            }
            __syncwarp();
            i += n; prevP = P; if (is_hot) lastHot = P;
        }
        // leave the scratch zeroed for the next region and for the depth pass
        { const int64_t a = P0 < 0 ? 0 : P0, z = maxE < L ? maxE : L; for (int64_t e = a + lane; e <= z; e += 32) endcnt[e] = 0; }
        __syncwarp();
        while (h < hhi && H[h] < i) h++;
    }
}

// Region starts among the hot reads: a hot read opens a new region when no earlier hot read of its target can still be alive
// at its position (the sequential walk of k_x_cap stops at `P > lastHot + span` for the same reason).
__global__ void __launch_bounds__(256) k_x_region_flag(uint32_t nH, const uint32_t* __restrict__ H, const int32_t* __restrict__ u_tid, const int32_t* __restrict__ u_pos,
                                                        const int32_t* __restrict__ maxspan_t, uint32_t* __restrict__ flag) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nH) return;
    bool st = h == 0;
    if (!st) {
        const uint32_t a = H[h - 1], b = H[h];
        st = u_tid[a] != u_tid[b] || (int64_t)u_pos[b] > (int64_t)u_pos[a] + maxspan_t[u_tid[b]];
    }
    flag[h] = st ? 1u : 0u;
}

// k_x_cap with the state in shared memory: one CTA per REGION (regions are independent), end counts in a ring indexed by
// position (valid while span + 2 <= XR_RING).  Same arithmetic as k_x_cap, which remains the path for targets whose reads span
// more than the ring.  The columns of a region are walked one after the other (htslib's rule is sequential), but everything
// inside a column — freeing the ends that were passed, counting the reads that start on it, inserting the accepted ones — is
// spread over the 256 threads: a hot locus puts tens of thousands of reads on ONE column, and with one warp per region those
// passed through 32 lanes (37 ms on the hot-locus preset c4).
constexpr int XR_RING = 4096;
constexpr int XC_THREADS = 256;
__device__ __forceinline__ uint32_t xc_block_sum(uint32_t v, uint32_t* s_red) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < XC_THREADS / 32; w++) t += s_red[w];
    __syncthreads();
    return t;
}
__global__ void __launch_bounds__(XC_THREADS) k_x_cap_ring(uint32_t nR, const uint32_t* __restrict__ RS, const uint32_t* __restrict__ H, uint32_t nH,
                                                            const int32_t* __restrict__ u_tid, const uint32_t* __restrict__ u_toff, const int32_t* __restrict__ u_pos,
                                                            const int32_t* __restrict__ u_alen, const int32_t* __restrict__ maxspan_t, uint8_t* __restrict__ accepted) {
    __shared__ uint32_t ring[XR_RING];
    __shared__ uint32_t s_red[XC_THREADS / 32];
    __shared__ long long s_live;
    const uint32_t r = blockIdx.x; const int tid = threadIdx.x;
    if (r >= nR) return;
    constexpr uint32_t M = XR_RING - 1;
    uint32_t h = RS[r]; const uint32_t hEnd = r + 1 < nR ? RS[r + 1] : nH;
    const uint32_t start = H[h];
    const int32_t t = u_tid[start];
    const uint32_t lo = u_toff[t], hi = u_toff[t + 1];
    const int64_t span = maxspan_t[t];
    const int64_t P0 = u_pos[start];
    for (int k = tid; k < XR_RING; k += XC_THREADS) ring[k] = 0;
    __syncthreads();
    int64_t live = 0;
    {   // every earlier read still alive at P0 is accepted (no hot column within reach)
        const uint32_t a0 = x_lower(u_pos, lo, start, P0 - span);
        uint32_t c = 0;
        for (uint32_t j = a0 + tid; j < start; j += XC_THREADS) {
            const int64_t E = (int64_t)u_pos[j] + u_alen[j];
            if (E >= P0) { c++; atomicAdd(&ring[(uint32_t)E & M], 1u); }
        }
        live = xc_block_sum(c, s_red);
    }
    uint32_t i = start; int64_t prevP = P0, lastHot = P0;
    while (i < hi) {                                       // uniform over the block: every quantity below is the same in all threads
        const int64_t P = u_pos[i];
        if (P > lastHot + span) break;
        if (P > prevP) {                                   // nodes whose end has been passed are freed (end <= P - 1); their slots recycle
            uint32_t s2 = 0;
            for (int64_t e = prevP + tid; e < P; e += XC_THREADS) { s2 += ring[(uint32_t)e & M]; ring[(uint32_t)e & M] = 0; }
            live -= xc_block_sum(s2, s_red);
        }
        const uint32_t n = x_lower(u_pos, i, hi, P + 1) - i;   // reads on this column (positions are sorted)
        const bool is_hot = h < hEnd && H[h] == i;
        uint32_t keep = n;
        if (is_hot) { const int64_t room = (int64_t)(PLP_MAXCNT - 1) - live; keep = (uint32_t)(room < 1 ? 1 : room > (int64_t)n ? (int64_t)n : room); h += n; }
        bool zero_len = false;
        if (is_hot) { uint32_t z = 0; for (uint32_t k = tid; k < n; k += XC_THREADS) z |= u_alen[i + k] == 0 ? 1u : 0u; zero_len = xc_block_sum(z, s_red) != 0u; }
        if (is_hot && zero_len) {                          // a read without reference span gets no node unless it opens the column: replay one by one
            if (tid == 0) {
                int64_t lv = live;
                for (uint32_t k = 0; k < n; k++) {
                    const int64_t E = P + u_alen[i + k];
                    const bool ok = k == 0 || 2 + lv <= PLP_MAXCNT;
                    accepted[i + k] = ok ? 1 : 0;
                    if (ok && (k == 0 ? E >= P : E > P)) { ring[(uint32_t)E & M] += 1; lv++; }
                }
                s_live = lv;
            }
            __syncthreads();
            live = s_live;
            __syncthreads();
        } else {
            uint32_t c = 0;
            for (uint32_t k = tid; k < n; k += XC_THREADS) {
                if (k >= keep) { accepted[i + k] = 0; continue; }
                const int64_t E = P + u_alen[i + k];
                if (k == 0 ? E >= P : E > P) { c++; atomicAdd(&ring[(uint32_t)E & M], 1u); }
            }
            live += xc_block_sum(c, s_red);
        }
        i += n; prevP = P; if (is_hot) lastHot = P;
    }
}

// out[t] = max of target t's slice of v; blockIdx.y walks the targets [t0, t0 + gridDim.y).  16-byte loads once the slice
// is aligned (slices start at arbitrary word offsets), several of them in flight per thread.
__global__ void __launch_bounds__(256) k_x_max(const uint32_t* __restrict__ v, const uint64_t* __restrict__ doff, int32_t t0, uint32_t* __restrict__ out) {
    const int32_t t = t0 + (int32_t)blockIdx.y;
    const uint32_t* p = v + doff[t]; const uint64_t n = doff[t + 1] - doff[t];
    uint64_t head = (4 - (((uint64_t)(uintptr_t)p >> 2) & 3)) & 3; if (head > n) head = n;
    const uint64_t nv = (n - head) >> 2;
    const uint4* q = reinterpret_cast<const uint4*>(p + head);
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t m = 0;
    uint64_t i = tid;
    for (; i + 3 * stride < nv; i += 4 * stride) {
        const uint4 a = q[i], b = q[i + stride], c = q[i + 2 * stride], d = q[i + 3 * stride];
        m = max(m, max(max(max(a.x, a.y), max(a.z, a.w)), max(max(b.x, b.y), max(b.z, b.w))));
        m = max(m, max(max(max(c.x, c.y), max(c.z, c.w)), max(max(d.x, d.y), max(d.z, d.w))));
    }
    for (; i < nv; i += stride) { const uint4 a = q[i]; m = max(m, max(max(a.x, a.y), max(a.z, a.w))); }
    if (tid < head) m = max(m, p[tid]);
    for (uint64_t k = head + (nv << 2) + tid; k < n; k += stride) m = max(m, p[k]);
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out + t, m);
}

// Junction::calcCoverage(a, b, levels) (junction.cc:923-933) for the four windows of :935-951
__global__ void __launch_bounds__(256) k_x_coverage(int64_t n, const int32_t* __restrict__ start, const int32_t* __restrict__ end, const uint32_t* __restrict__ depth,
                                                     int64_t size, uint32_t* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * 4) return;
    const int64_t j = g >> 2; const int w = (int)(g & 3);
    const int32_t R = 10;
    const int32_t s = start[j], e = end[j];
    int32_t a, b;
    if (w == 0) { a = s - 2 * R; b = s - R - 1; } else if (w == 1) { a = s - R; b = s; }
    else if (w == 2) { a = e + R; b = e + 2 * R; } else { a = e; b = e + R - 1; }
    uint32_t sum = 0;
    for (int64_t i = a; i <= b; i++) if (i >= 0 && i < size) sum += depth[i];
    out[g] = sum;
}

// The same for junctions of many targets at once: dtid[j] names the target whose depth vector junction j is scored against
// (the Q14 mapping, -1: none -> zeros), doff / tlen locate that vector.
__global__ void __launch_bounds__(256) k_x_coverage_batch(int64_t n, const int32_t* __restrict__ start, const int32_t* __restrict__ end, const int32_t* __restrict__ dtid,
                                                           const uint32_t* __restrict__ depth, const uint64_t* __restrict__ doff, const int32_t* __restrict__ tlen,
                                                           uint32_t* __restrict__ out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n * 4) return;
    const int64_t j = g >> 2; const int w = (int)(g & 3);
    const int32_t dt = dtid[j];
    if (dt < 0) { out[g] = 0; return; }
    const uint32_t* d = depth + doff[dt]; const int64_t size = tlen[dt];
    const int32_t R = 10;
    const int32_t s = start[j], e = end[j];
    int32_t a, b;
    if (w == 0) { a = s - 2 * R; b = s - R - 1; } else if (w == 1) { a = s - R; b = s; }
    else if (w == 2) { a = e + R; b = e + 2 * R; } else { a = e; b = e + R - 1; }
    uint32_t sum = 0;
    for (int64_t i = a; i <= b; i++) if (i >= 0 && i < size) sum += d[i];
    out[g] = sum;
}

__global__ void __launch_bounds__(256) k_x_keep(uint32_t P, const uint32_t* __restrict__ vals, const PairRec* __restrict__ pr, uint32_t* __restrict__ rid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P) rid[i] = pr[vals[i]].a.rid;
}

// CUDA-event stage marks of pj_extra_run (same idea as the marks of pj_shard_run)
struct XMarks {
    cudaStream_t st; std::vector<std::pair<const char*, cudaEvent_t>> ev;
    explicit XMarks(cudaStream_t s) : st(s) {}
    void mark(const char* name) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); ev.emplace_back(name, e); }
    void collect(pj_ctx* c) {
        c->x_stage_ms.clear(); c->x_stage_names.clear();
        for (size_t k = 1; k < ev.size(); k++) { float ms = 0; cudaEventElapsedTime(&ms, ev[k - 1].second, ev[k].second); c->x_stage_ms.push_back(ms); c->x_stage_names.push_back(ev[k].first); }
        c->x_total_ms = 0; if (ev.size() > 1) cudaEventElapsedTime(&c->x_total_ms, ev.front().second, ev.back().second);
    }
    ~XMarks() { for (auto& e : ev) cudaEventDestroy(e.second); }
};

inline uint32_t blocks_for(uint64_t n, uint32_t per) { return (uint32_t)((n + per - 1) / per); }

} // namespace

namespace pjapi {

// All of this state comes from the stream-ordered pool (the pool keeps the memory between shards, so a steady-state
// step pays no allocation) and is returned on the compute stream.
void extra_reset(pj_ctx* c) {
    if (!c->compute_stream) return;
    for (void* p : {(void*)c->x_pair_rid, (void*)c->x_pair_jid, (void*)c->x_names, (void*)c->x_uflag, (void*)c->x_alen, (void*)c->x_depth})
        if (p) cudaFreeAsync(p, c->compute_stream);
    c->x_pair_rid = c->x_pair_jid = nullptr; c->x_names = nullptr; c->x_uflag = nullptr; c->x_alen = nullptr; c->x_depth = nullptr;
    c->x_n_spliced = -1; c->x_imported.clear(); c->x_doff.clear(); c->x_covered.clear(); c->x_maxlive.clear(); c->x_ready = false;
}

// end of pj_shard_run: remember, for every (read, junction) pair in sorted order, its record and its junction
int extra_keep_pairs(pj_ctx* c, uint32_t P, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, cudaStream_t st) {
    CU(c, cudaMallocAsync(&c->x_pair_rid, (size_t)std::max<uint32_t>(P, 1) * 4, st)); CU(c, cudaMallocAsync(&c->x_pair_jid, (size_t)std::max<uint32_t>(P, 1) * 4, st));
    if (P) {
        k_x_keep<<<blocks_for(P, 256), 256, 0, st>>>(P, vals, pr, c->x_pair_rid); c->n_launches++;
        CU(c, cudaMemcpyAsync(c->x_pair_jid, jid, (size_t)P * 4, cudaMemcpyDeviceToDevice, st));
    }
    return PJ_OK;
}

// end of pj_shard_run: classify the shard's records (spliced / unspliced + mapped / neither) and compact the name codes
// of the spliced ones, so that pj_extra_export_names can hand them to the other contexts before pj_extra_run
int extra_classify(pj_ctx* c, cudaStream_t st) {
    const int64_t R = c->n_rec;
    unsigned long long* d_n = nullptr;
    CU(c, cudaMallocAsync(&c->x_names, (size_t)std::max<int64_t>(R, 1) * 8, st));
    CU(c, cudaMallocAsync(&c->x_uflag, (size_t)std::max<int64_t>(R, 1) * 4, st)); CU(c, cudaMallocAsync(&c->x_alen, (size_t)std::max<int64_t>(R, 1) * 4, st));
    uint32_t* uflag = c->x_uflag; int32_t* alen = c->x_alen;
    CU(c, cudaMallocAsync(&d_n, 8, st)); CU(c, cudaMemsetAsync(d_n, 0, 8, st));
    CU(c, cudaMemsetAsync(c->d_scalars + 8, 0, 4 * sizeof(uint32_t), st));
    if (R) k_x_classify<<<blocks_for((uint64_t)R, 256), 256, 0, st>>>(R, c->cigar_off.p, c->cigar.p, c->flag.p, c->tid.p, c->name_code.p, c->n_targets,
                                                                       uflag, alen, c->x_names, d_n, c->d_scalars + 8);
    c->n_launches++;
    unsigned long long n = 0;
    CU(c, cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st)); CU(c, cudaGetLastError());
    CU(c, cudaFreeAsync(d_n, st));
    c->x_n_spliced = (int64_t)n;
    return PJ_OK;
}

} // namespace pjapi

extern "C" {

static_assert(sizeof(pj_junction_extra) == 48, "pj_junction_extra layout changed: update the bindings");

int64_t pj_extra_num_spliced_names(const pj_ctx* c) { return (c && c->extra && c->have_result) ? c->x_n_spliced : -1; }

int pj_extra_export_names(pj_ctx* c, uint64_t* codes, int64_t cap) {
    if (!c || !c->extra || !c->have_result || c->x_n_spliced < 0) return fail(c, PJ_ESTATE, "pj_extra_export_names: needs a context created with extra_metrics and a finished pj_shard_run");
    if (cap < c->x_n_spliced || (c->x_n_spliced && !codes)) return fail(c, PJ_EINVAL, "pj_extra_export_names: capacity %lld < %lld", (long long)cap, (long long)c->x_n_spliced);
    CU(c, cudaSetDevice(c->device));
    if (c->x_n_spliced) CU(c, cudaMemcpy(codes, c->x_names, (size_t)c->x_n_spliced * 8, cudaMemcpyDeviceToHost));
    return PJ_OK;
}

int pj_extra_import_names(pj_ctx* c, const uint64_t* codes, int64_t n) {
    if (!c || !c->extra || !c->have_result) return fail(c, PJ_ESTATE, "pj_extra_import_names: needs a context created with extra_metrics and a finished pj_shard_run");
    if (n < 0 || (n && !codes)) return fail(c, PJ_EINVAL, "pj_extra_import_names: bad arguments");
    c->x_imported.insert(c->x_imported.end(), codes, codes + n);
    return PJ_OK;
}

int pj_extra_run(pj_ctx* c, int32_t max_query_length, pj_junction_extra* out, int64_t cap_rows) {
    if (!c || !c->extra || !c->have_result) return fail(c, PJ_ESTATE, "pj_extra_run: needs a context created with extra_metrics and a finished pj_shard_run");
    if (cap_rows < c->n_junc || (c->n_junc && !out)) return fail(c, PJ_EINVAL, "pj_extra_run: rows capacity %lld < %lld", (long long)cap_rows, (long long)c->n_junc);
    if (c->x_ready) return fail(c, PJ_ESTATE, "pj_extra_run: already run for this shard");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->compute_stream;
    const int64_t R = c->n_rec; const int32_t T = c->n_targets; const uint32_t J = (uint32_t)c->n_junc; const uint32_t P = (uint32_t)c->n_pairs;
    if (c->x_n_spliced < 0 || !c->x_uflag) return fail(c, PJ_ESTATE, "pj_extra_run: internal state");
    uint32_t* uflag = c->x_uflag; int32_t* alen = c->x_alen;
    uint32_t err = 0;
    CU(c, cudaMemcpy(&err, c->d_scalars + 8, 4, cudaMemcpyDeviceToHost));
    if (err & XERR_NOCIGAR) return fail(c, PJ_EDATA, "input rejected (the reference aborts on it): mapped unspliced record without CIGAR (htslib pileup, sam.c:1537)");

    XMarks marks(st); marks.mark("begin"); int launches = 0;
    // ---- unspliced records, compacted in BAM order ----
    uint32_t* uoff = nullptr; uint32_t* scan_tmp = nullptr;
    CU(c, cudaMallocAsync(&uoff, (size_t)std::max<int64_t>(R, 1) * 4, st));
    CU(c, cudaMallocAsync(&scan_tmp, scan_tmp_elems((uint64_t)std::max<int64_t>(R, 1)) * 4, st));
    launch_exclusive_scan(uflag, uoff, (uint64_t)R, scan_tmp, c->d_scalars + 9, st);
    uint32_t U = 0;
    CU(c, cudaMemcpyAsync(&U, c->d_scalars + 9, 4, cudaMemcpyDeviceToHost, st)); CU(c, cudaStreamSynchronize(st));
    int32_t *u_tid = nullptr, *u_pos = nullptr, *u_alen = nullptr, *maxspan = nullptr; uint32_t *u_rid = nullptr, *u_toff = nullptr; uint8_t* covered = nullptr;
    const size_t Un = std::max<uint32_t>(U, 1);
    CU(c, cudaMallocAsync(&u_tid, Un * 4, st)); CU(c, cudaMallocAsync(&u_pos, Un * 4, st)); CU(c, cudaMallocAsync(&u_alen, Un * 4, st)); CU(c, cudaMallocAsync(&u_rid, Un * 4, st));
    CU(c, cudaMallocAsync(&maxspan, (size_t)T * 4, st)); CU(c, cudaMemsetAsync(maxspan, 0, (size_t)T * 4, st));
    CU(c, cudaMallocAsync(&covered, (size_t)T, st)); CU(c, cudaMemsetAsync(covered, 0, (size_t)T, st));
    CU(c, cudaMallocAsync(&u_toff, ((size_t)T + 1) * 4, st)); CU(c, cudaMemsetAsync(u_toff, 0, ((size_t)T + 1) * 4, st));
    if (R) k_x_scatter<<<blocks_for((uint64_t)R, 256), 256, 0, st>>>(R, uflag, uoff, c->tid.p, c->pos.p, alen, u_tid, u_pos, u_alen, u_rid, maxspan, covered);
    k_x_toff<<<blocks_for((uint64_t)U + 1, 256), 256, 0, st>>>(U, u_tid, u_pos, T, u_toff, c->d_scalars + 8);
    CU(c, cudaFreeAsync(uoff, st));
    launches += 5; marks.mark("x_unspliced");

    // ---- multiple-mapping score: name table over the spliced records of the whole file ----
    pj_junction_extra* d_out = nullptr;
    CU(c, cudaMallocAsync(&d_out, (size_t)std::max<uint32_t>(J, 1) * sizeof(pj_junction_extra), st));
    CU(c, cudaMemsetAsync(d_out, 0, (size_t)std::max<uint32_t>(J, 1) * sizeof(pj_junction_extra), st));
    {
        const uint64_t n_names = (uint64_t)c->x_n_spliced + c->x_imported.size();
        uint64_t cap = 1024; while (cap < 2 * n_names + 2) cap <<= 1;
        uint64_t* keys = nullptr; uint32_t* counts = nullptr; uint64_t* d_imp = nullptr;
        CU(c, cudaMallocAsync(&keys, cap * 8, st)); CU(c, cudaMemsetAsync(keys, 0xff, cap * 8, st));
        CU(c, cudaMallocAsync(&counts, cap * 4, st)); CU(c, cudaMemsetAsync(counts, 0, cap * 4, st));
        if (c->x_n_spliced) k_x_names_add<<<blocks_for((uint64_t)c->x_n_spliced, 256), 256, 0, st>>>(c->x_n_spliced, c->x_names, keys, counts, cap - 1);
        if (!c->x_imported.empty()) {
            CU(c, cudaMallocAsync(&d_imp, c->x_imported.size() * 8, st));
            CU(c, cudaMemcpyAsync(d_imp, c->x_imported.data(), c->x_imported.size() * 8, cudaMemcpyHostToDevice, st));
            k_x_names_add<<<blocks_for(c->x_imported.size(), 256), 256, 0, st>>>((int64_t)c->x_imported.size(), d_imp, keys, counts, cap - 1);
        }
        if (P) k_x_mm<<<blocks_for(P, 256), 256, 0, st>>>(P, c->x_pair_rid, c->x_pair_jid, c->name_code.p, keys, counts, cap - 1, d_out);
        CU(c, cudaFreeAsync(keys, st)); CU(c, cudaFreeAsync(counts, st)); if (d_imp) { CU(c, cudaStreamSynchronize(st)); CU(c, cudaFreeAsync(d_imp, st)); }
    }

    launches += 3; marks.mark("x_mm_score");
    // ---- flanking alignments ----
    if (J) k_x_flank<<<blocks_for((uint64_t)J * 32, 256), 256, 0, st>>>(J, c->d_rows, c->d_tlen, u_toff, u_pos, u_alen, maxspan, max_query_length, d_out);
    if (J) CU(c, cudaMemcpyAsync(out, d_out, (size_t)J * sizeof(pj_junction_extra), cudaMemcpyDeviceToHost, st));

    launches += 1; marks.mark("x_flank");
    // ---- unspliced pileup: live-read maximum (htslib cap check), then the depth vectors kept for pj_extra_coverage ----
    c->x_doff.assign((size_t)T + 1, 0);
    for (int32_t t = 0; t < T; t++) c->x_doff[t + 1] = c->x_doff[t] + (uint64_t)c->h_tlen[t] + 1;
    const uint64_t D = c->x_doff[T];
    uint64_t* d_doff = nullptr; uint32_t* d_max = nullptr; uint32_t* dscan_tmp = nullptr;
    CU(c, cudaMallocAsync(&c->x_depth, D * 4, st));
    CU(c, cudaMallocAsync(&d_doff, ((size_t)T + 1) * 8, st)); CU(c, cudaMemcpyAsync(d_doff, c->x_doff.data(), ((size_t)T + 1) * 8, cudaMemcpyHostToDevice, st));
    CU(c, cudaMallocAsync(&d_max, (size_t)T * 4, st)); CU(c, cudaMemsetAsync(d_max, 0, (size_t)T * 4, st));
    CU(c, cudaMallocAsync(&dscan_tmp, scan_tmp_elems(std::max<uint64_t>(D, U)) * 4, st));      // also scans the U hot-read flags
    CU(c, cudaMemsetAsync(c->x_depth, 0, D * 4, st));
    if (U) k_x_live<<<blocks_for(U, 256), 256, 0, st>>>(U, u_tid, u_pos, u_alen, c->d_tlen, d_doff, c->x_depth);
    launch_exclusive_scan(c->x_depth, c->x_depth, D, dscan_tmp, c->d_scalars + 10, st);
    if (U) {
        int32_t longest = 1; for (int32_t t = 0; t < T; t++) longest = std::max(longest, c->h_tlen[t]);
        const uint32_t gx = std::max<uint32_t>(1, std::min<uint32_t>(blocks_for((uint64_t)longest + 1, 256 * 8), (uint32_t)(148 * 8 / std::min<int32_t>(T, 148) + 1)));
        for (int32_t t0 = 0; t0 < T; t0 += 65535) k_x_max<<<dim3(gx, (uint32_t)std::min<int32_t>(65535, T - t0)), 256, 0, st>>>(c->x_depth, d_doff, t0, d_max);
    }
    launches += 5 + (T - 1) / 65535; marks.mark("x_live");
    c->x_covered.assign((size_t)T, 0); c->x_maxlive.assign((size_t)T, 0);
    CU(c, cudaMemcpyAsync(c->x_covered.data(), covered, (size_t)T, cudaMemcpyDeviceToHost, st));
    CU(c, cudaMemcpyAsync(c->x_maxlive.data(), d_max, (size_t)T * 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st));
    bool capped = false; for (int32_t t = 0; t < T; t++) capped |= c->x_maxlive[t] >= (uint32_t)PLP_MAXCNT;
    uint32_t *hot = nullptr, *hoff = nullptr, *H = nullptr; uint8_t* accepted = nullptr;
    uint32_t nH = 0;
    if (capped && U) {      // some column holds >= 8000 live reads: replay htslib's read cap there
        CU(c, cudaMallocAsync(&hot, (size_t)U * 4, st)); CU(c, cudaMallocAsync(&hoff, (size_t)U * 4, st));
        k_x_hot<<<blocks_for(U, 256), 256, 0, st>>>(U, u_tid, u_pos, c->d_tlen, d_doff, c->x_depth, hot);
        launch_exclusive_scan(hot, hoff, U, dscan_tmp, c->d_scalars + 11, st);
        CU(c, cudaMemcpyAsync(&nH, c->d_scalars + 11, 4, cudaMemcpyDeviceToHost, st)); CU(c, cudaStreamSynchronize(st));
    }
    CU(c, cudaMemsetAsync(c->x_depth, 0, D * 4, st));
    if (nH) {
        CU(c, cudaMallocAsync(&H, (size_t)nH * 4, st)); CU(c, cudaMallocAsync(&accepted, (size_t)U, st)); CU(c, cudaMemsetAsync(accepted, 1, (size_t)U, st));
        k_x_compact<<<blocks_for(U, 256), 256, 0, st>>>(U, hot, hoff, H);
        // regions are independent: one warp each with its state in shared memory when every capped target's longest read span fits
        // the ring, else the general kernel (one warp per target, counters in the zeroed depth buffer)
        std::vector<int32_t> h_span((size_t)T);
        CU(c, cudaMemcpyAsync(h_span.data(), maxspan, (size_t)T * 4, cudaMemcpyDeviceToHost, st)); CU(c, cudaStreamSynchronize(st));
        bool fits = true; for (int32_t t = 0; t < T; t++) if (c->x_maxlive[t] >= (uint32_t)PLP_MAXCNT && (int64_t)h_span[t] + 2 > XR_RING) fits = false;
        if (getenv("PJ_CAP_GENERAL")) fits = false;            // tests: force the general kernel
        if (fits) {
            uint32_t *rflag = nullptr, *roff = nullptr, *RS = nullptr; uint32_t nR = 0;
            CU(c, cudaMallocAsync(&rflag, (size_t)nH * 4, st)); CU(c, cudaMallocAsync(&roff, (size_t)nH * 4, st));
            k_x_region_flag<<<blocks_for(nH, 256), 256, 0, st>>>(nH, H, u_tid, u_pos, maxspan, rflag);
            launch_exclusive_scan(rflag, roff, nH, dscan_tmp, c->d_scalars + 11, st);
            CU(c, cudaMemcpyAsync(&nR, c->d_scalars + 11, 4, cudaMemcpyDeviceToHost, st)); CU(c, cudaStreamSynchronize(st));
            CU(c, cudaMallocAsync(&RS, (size_t)std::max<uint32_t>(nR, 1) * 4, st));
            k_x_compact<<<blocks_for(nH, 256), 256, 0, st>>>(nH, rflag, roff, RS);
            if (nR) k_x_cap_ring<<<nR, XC_THREADS, 0, st>>>(nR, RS, H, nH, u_tid, u_toff, u_pos, u_alen, maxspan, accepted);
            CU(c, cudaFreeAsync(rflag, st)); CU(c, cudaFreeAsync(roff, st)); CU(c, cudaFreeAsync(RS, st));
            launches += 6;
        } else {
            k_x_cap<<<(uint32_t)T, 32, 0, st>>>(T, u_toff, u_pos, u_alen, maxspan, c->d_tlen, d_doff, H, nH, c->x_depth, accepted);
        }
    }
    if (nH) { launches += 6; marks.mark("x_cap"); }
    if (U) k_x_depth<<<blocks_for(U, 256), 256, 0, st>>>(U, u_tid, u_pos, u_rid, c->cigar_off.p, c->cigar.p, accepted, c->d_tlen, d_doff, c->x_depth);
    launch_exclusive_scan(c->x_depth, c->x_depth, D, dscan_tmp, c->d_scalars + 10, st);
    launches += 4; marks.mark("x_depth");
    for (void* p : {(void*)hot, (void*)hoff, (void*)H, (void*)accepted}) if (p) CU(c, cudaFreeAsync(p, st));
    CU(c, cudaMemcpyAsync(&err, c->d_scalars + 8, 4, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st)); CU(c, cudaGetLastError());
    for (void* p : {(void*)u_tid, (void*)u_pos, (void*)u_alen, (void*)u_rid, (void*)maxspan, (void*)covered, (void*)u_toff, (void*)scan_tmp, (void*)d_out,
                    (void*)d_doff, (void*)d_max, (void*)dscan_tmp}) CU(c, cudaFreeAsync(p, st));
    marks.collect(c); c->x_launches = launches;
    if (err & XERR_UNSORTED) return fail(c, PJ_EINVAL, "pj_extra_run: records are not in (tid, pos) order");
    cudaFreeAsync(c->x_uflag, st); cudaFreeAsync(c->x_alen, st); c->x_uflag = nullptr; c->x_alen = nullptr;
    c->x_ready = true;
    return PJ_OK;
}

int pj_extra_timing(const pj_ctx* c, float* total_ms, int32_t* n_launches) {
    if (!c || !c->x_ready) return PJ_ESTATE;
    if (total_ms) *total_ms = c->x_total_ms;
    if (n_launches) *n_launches = c->x_launches;
    return PJ_OK;
}

int pj_extra_kernel_times(const pj_ctx* c, int32_t cap, float* kernel_ms, const char** kernel_names, int32_t* n) {
    if (!c || !n) return PJ_EINVAL;
    *n = (int32_t)c->x_stage_ms.size();
    for (int32_t k = 0; k < cap && k < *n; k++) { if (kernel_ms) kernel_ms[k] = c->x_stage_ms[k]; if (kernel_names) kernel_names[k] = c->x_stage_names[k]; }
    return PJ_OK;
}

int pj_extra_target_pileup(pj_ctx* c, int32_t tid, int32_t* covered, uint32_t* max_depth) {
    if (!c || !c->x_ready) return fail(c, PJ_ESTATE, "pj_extra_target_pileup: call pj_extra_run first");
    if (tid < 0 || tid >= c->n_targets) return fail(c, PJ_EINVAL, "pj_extra_target_pileup: bad target");
    if (covered) *covered = c->x_covered[tid];
    if (max_depth) *max_depth = c->x_maxlive[tid];
    return PJ_OK;
}

int pj_extra_coverage(pj_ctx* c, int32_t depth_tid, int64_t n, const int32_t* intron_start, const int32_t* intron_end, uint32_t* cov_sum4) {
    if (!c || !c->x_ready) return fail(c, PJ_ESTATE, "pj_extra_coverage: call pj_extra_run first");
    if (depth_tid < 0 || depth_tid >= c->n_targets || n < 0 || (n && (!intron_start || !intron_end || !cov_sum4))) return fail(c, PJ_EINVAL, "pj_extra_coverage: bad arguments");
    if (n == 0) return PJ_OK;
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->compute_stream;
    int32_t *ds = nullptr, *de = nullptr; uint32_t* dout = nullptr;
    CU(c, cudaMallocAsync(&ds, (size_t)n * 4, st)); CU(c, cudaMallocAsync(&de, (size_t)n * 4, st)); CU(c, cudaMallocAsync(&dout, (size_t)n * 16, st));
    CU(c, cudaMemcpyAsync(ds, intron_start, (size_t)n * 4, cudaMemcpyHostToDevice, st)); CU(c, cudaMemcpyAsync(de, intron_end, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    k_x_coverage<<<blocks_for((uint64_t)n * 4, 256), 256, 0, st>>>(n, ds, de, c->x_depth + c->x_doff[depth_tid], (int64_t)c->h_tlen[depth_tid], dout);
    CU(c, cudaMemcpyAsync(cov_sum4, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st)); CU(c, cudaGetLastError());
    CU(c, cudaFreeAsync(ds, st)); CU(c, cudaFreeAsync(de, st)); CU(c, cudaFreeAsync(dout, st));
    return PJ_OK;
}

// One launch for the junctions of any number of targets: depth_tid[j] is the target whose depth vector junction j uses
// (pj_extra_coverage_source of the junction's own target), or -1.  cov_sum4[j][4] as pj_extra_coverage.
int pj_extra_coverage_batch(pj_ctx* c, int64_t n, const int32_t* depth_tid, const int32_t* intron_start, const int32_t* intron_end, uint32_t* cov_sum4) {
    if (!c || !c->x_ready) return fail(c, PJ_ESTATE, "pj_extra_coverage_batch: call pj_extra_run first");
    if (n < 0 || (n && (!depth_tid || !intron_start || !intron_end || !cov_sum4))) return fail(c, PJ_EINVAL, "pj_extra_coverage_batch: bad arguments");
    if (n == 0) return PJ_OK;
    for (int64_t j = 0; j < n; j++) if (depth_tid[j] >= c->n_targets) return fail(c, PJ_EINVAL, "pj_extra_coverage_batch: bad target");
    CU(c, cudaSetDevice(c->device));
    cudaStream_t st = c->compute_stream;
    int32_t *ds = nullptr, *de = nullptr, *dt = nullptr; uint32_t* dout = nullptr; uint64_t* ddoff = nullptr;
    const size_t T = (size_t)c->n_targets;
    CU(c, cudaMallocAsync(&ds, (size_t)n * 4, st)); CU(c, cudaMallocAsync(&de, (size_t)n * 4, st)); CU(c, cudaMallocAsync(&dt, (size_t)n * 4, st));
    CU(c, cudaMallocAsync(&dout, (size_t)n * 16, st)); CU(c, cudaMallocAsync(&ddoff, (T + 1) * 8, st));
    CU(c, cudaMemcpyAsync(ds, intron_start, (size_t)n * 4, cudaMemcpyHostToDevice, st)); CU(c, cudaMemcpyAsync(de, intron_end, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(dt, depth_tid, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(c, cudaMemcpyAsync(ddoff, c->x_doff.data(), std::min(c->x_doff.size(), T + 1) * 8, cudaMemcpyHostToDevice, st));
    k_x_coverage_batch<<<blocks_for((uint64_t)n * 4, 256), 256, 0, st>>>(n, ds, de, dt, c->x_depth, ddoff, c->d_tlen, dout);
    CU(c, cudaMemcpyAsync(cov_sum4, dout, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    CU(c, cudaStreamSynchronize(st)); CU(c, cudaGetLastError());
    for (void* p : {(void*)ds, (void*)de, (void*)dt, (void*)dout, (void*)ddoff}) CU(c, cudaFreeAsync(p, st));
    return PJ_OK;
}

} // extern "C"
