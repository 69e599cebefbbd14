// junc_kernels.cuh — device-side data layout and per-item device functions of the junc pipeline.
//
// Everything here is written for sm_100a (B200).  No tensor-core work exists on this path (every step
// is integer compare / scan / sort / reduce), so the design rules that matter are coalesced columnar
// loads, 64/128-bit vector accesses, warp-shuffle reductions and grids sized from the SM count.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/portcullis_junc.h"

namespace pjk {

// ---- CIGAR (htslib sam.h:75-81; CigarOp::opConsumes*, bam_alignment.hpp:75-99) ----
enum : uint32_t { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
__host__ __device__ __forceinline__ uint32_t cig_op(uint32_t c) { return c & 0xfu; }
__host__ __device__ __forceinline__ int32_t  cig_len(uint32_t c) { return (int32_t)(c >> 4); }
// bit i set <=> op i consumes the reference (M,D,N,=,X) / the query excluding soft clips (M,I,=,X)
__host__ __device__ __forceinline__ bool op_ref(uint32_t op)   { return (0x18Du >> op) & 1u; }
__host__ __device__ __forceinline__ bool op_query(uint32_t op) { return (0x183u >> op) & 1u; }

// ---- error bits raised by kernels (mapped to PJ_EDATA / PJ_EINVAL on the host) ----
enum : uint32_t {
    ERR_ANCHOR_ORDER   = 1u << 0,   // intron not enclosed by its anchors (Intron::minAnchorLength throws, intron.cc:68-85)
    ERR_START_RANGE    = 1u << 1,   // intron start outside [0, target_len-2]: reference aborts fetching the donor site
    ERR_KEY_OVERFLOW   = 1u << 2,   // key does not fit 64 bits for this shard
    ERR_NO_PRESENCE    = 1u << 3,   // "alignment does not have a presence in the requested region" (bam_alignment.cc:342)
    ERR_QUERY_RANGE    = 1u << 4,   // "Can't extract cigar op sequence from query string" (bam_alignment.cc:376)
    ERR_EMPTY_ANCHOR   = 1u << 5,   // empty / mismatched anchor strings: undefined behaviour in the reference (Q6)
    ERR_GENOME_RANGE   = 1u << 6,   // splice-site / intron / anchor window leaves the genome sequence (junction.cc:573-633)
    ERR_SEQ_MISSING    = 1u << 7,   // spliced read without SEQ bytes in the batch
    ERR_ZERO_LEN       = 1u << 8,   // "length has been calculated as 0" (bam_alignment.cc:363)
};

// ---- packed genome resident in HBM ----
// g2 : 2 bits per base (A=0,C=1,G=2,T=3), 32 bases per 64-bit word, little-endian within the word (base k at bits 2k, 2k+1).
// gx : 1 bit per base, set when the upper-cased byte is not A/C/G/T; 64 bases per 64-bit word.
//      For an exception base the 2-bit field holds a sub-code: 0 = 'N', 1 = any other byte, whose exact value lives
//      in the sorted side table (exc_pos, exc_byte).  Soft-masking never reaches the device: the reference upper-cases
//      every fetched window (junction.cc:586-587, 635-638).
// The read side uses the same 2-bit code (Reads::seq2), so the anchor comparison of k_match is a 64-bit XOR of 32 bases; the
// motif and Hamming windows of k_finalize read the same plane.  0.375 B per base: 1.2 GB for a 3.1 Gb genome.
struct Genome {
    const uint64_t* g2;
    const uint64_t* gx;
    const uint32_t* gsum;       // 1 bit per 1024 bases: set when the stretch holds a base that is not A/C/G/T.  378 KB for a human genome, so it
                                // stays in L1/L2: the compare loop consults it once per block and reads the gx plane only where it says so
    const uint64_t* goff;       // [n_targets] first base index of each target (multiple of 64)
    const int64_t*  glen;       // [n_targets] sequence length from the FASTA (-1 when the target was not loaded)
    const uint64_t* exc_pos;    // sorted global base indices of "other" exception bytes
    const uint8_t*  exc_byte;
    int32_t n_exc;
    int32_t n_exc_x;            // how many of them are 'X'
    int32_t any_gx;             // != 0: some base of the genome is not A/C/G/T (the compare loop then also reads gx)
};

__device__ __forceinline__ uint8_t genome_exc_lookup(const Genome& g, uint64_t gi) {
    int lo = 0, hi = g.n_exc;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (g.exc_pos[mid] < gi) lo = mid + 1; else hi = mid; }
    return (lo < g.n_exc && g.exc_pos[lo] == gi) ? g.exc_byte[lo] : (uint8_t)'N';
}

// Upper-cased genome byte at global base index gi.
__device__ __forceinline__ uint8_t genome_char(const Genome& g, uint64_t gi) {
    const uint64_t w2 = __ldg(g.g2 + (gi >> 5));
    const uint64_t wx = __ldg(g.gx + (gi >> 6));
    const uint32_t code = (uint32_t)(w2 >> ((gi & 31) * 2)) & 3u;
    if (!((wx >> (gi & 63)) & 1ull)) return (uint8_t)("ACGT"[code]);
    return code == 0 ? (uint8_t)'N' : genome_exc_lookup(g, gi);
}

// SeqUtils::reverseComplement lookup (seq_utils.hpp:33-40), index c-'A'; 0 for holes and out-of-range bytes.
__device__ __forceinline__ uint8_t revcomp_char(uint8_t c) {
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    case 'N': return 'N'; case 'D': return 'H'; case 'H': return 'D'; case 'M': return 'K';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'W'; case 'W': return 'S';
    case 'U': return 'A'; case 'V': return 'B'; case 'X': return 'X';
    default: return 0;
    }
}

// ---- alignment columns of one shard, resident in HBM ----
struct Reads {
    int64_t n;
    const int32_t*  tid;
    const int32_t*  pos;
    const uint16_t* flag;
    const uint8_t*  mapq;
    const uint8_t*  xs;
    const int32_t*  l_qseq;
    const int32_t*  mtid;
    const int32_t*  mpos;
    const uint32_t* cigar_off;
    const uint32_t* cigar;
    const uint64_t* seq_off;    // n+1 byte offsets into seq2
    const uint8_t*  seq2;       // read bases, 2 bits each (A=0,C=1,G=2,T=3; base q of a record at bits 2(q&3) of its byte q>>2), every record on a
                                // byte boundary; the stream has a 16-byte lead pad and tail slack.  A base that is not A/C/G/T is stored as 0 and
                                // listed in (seqx_pos, seqx_code); its record carries FLAG_SEQX in the flag column.
    const uint64_t* seqx_pos;   // sorted base indices (byte offset * 4 + q) of the non-ACGT read bases
    const uint8_t*  seqx_code;  // their BAM nibbles (index into "=ACMGRSVTWYHKDBN")
    int64_t n_seqx;
};
constexpr uint32_t FLAG_SEQX = 0x8000u;   // flag column, bit 15 (unused by the SAM spec): the record has non-ACGT read bases

// Per-(read, N-op) record, written once by the emit kernel in BAM order and gathered once per later stage.
// Two 16-byte halves so that each is one 128-bit load.
struct __align__(16) PairA { uint32_t rid; int32_t lstart; int32_t rend; int32_t pos; };
struct __align__(16) PairB { int32_t read_end; uint32_t bits; uint32_t updown; int32_t start; };
// Junction-local view of the read for the anchor comparison (k_match): lets it start at the N op instead of re-walking the
// whole CIGAR (long reads have dozens of ops and a pair per N op) and spares it the per-read column gathers.
struct __align__(16) PairC { uint64_t seq_b0; uint32_t cig_abs; int32_t qpos_n; };     // first base of the clipped query in the SEQ stream (base index);
                                                                                       // index of this N op in the CIGAR stream; query offset at it
struct __align__(16) PairD { int32_t qsize; int32_t lq; uint32_t nops; uint32_t pad; }; // clipped query length (Q3); l_qseq; ops before | ops after << 16
// The four 16-byte parts of one pair live in ONE 64-byte, 64-byte-aligned record: a gather touches two full 32-byte sectors of
// one line instead of four half-used sectors in four arrays; the stage-1 reduction needs only the first sector (a, b).
struct __align__(64) PairRec { PairA a; PairB b; PairC c; PairD d; };
enum : uint32_t { PB_R1 = 1u << 0, PB_REV = 1u << 1, PB_MS = 1u << 2, PB_UM = 1u << 3, PB_BPP = 1u << 4, PB_PPP = 1u << 5,
                  PB_XSP = 1u << 6, PB_XSN = 1u << 7, PB_SEQX = 1u << 8 /* the read has non-ACGT bases: exact per-base compare */ };

// BamAlignment::calcIfProperPair (bam_alignment.cc:271-292)
__host__ __device__ __forceinline__ bool portcullis_proper_pair(uint32_t flag, int32_t tid, int32_t mtid, int32_t pos, int32_t mpos, int orientation) {
    if (!(flag & 0x1u) || (flag & 0x8u)) return false;
    if (tid != mtid) return false;
    const bool rev = (flag & 0x10u) != 0, mrev = (flag & 0x20u) != 0;
    const bool diff = rev != mrev;
    const bool gap = !rev ? pos < mpos : pos > mpos;
    if (orientation == PJ_ORIENT_FR) return diff && gap;
    if (orientation == PJ_ORIENT_RF) return diff && !gap;
    if (orientation == PJ_ORIENT_FF) return !diff && gap;
    return false;
}

} // namespace pjk
