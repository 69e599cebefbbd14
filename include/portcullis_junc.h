/*
 * portcullis_junc.h — C ABI of the B200-native `junc` hot path.
 *
 * This is the drop-in boundary for Portcullis' junction-building pass.  The reference has no
 * plugin / FFI interface for this path; its inner seam is
 *
 *     void JunctionBuilder::findJuncs(BamReader&, GenomeMapper&, int32_t seq)
 *         -> results[seq] (RegionResult)            /root/reference/src/junction_builder.cc:314-357
 *                                                    /root/reference/src/junction_builder.hpp:72-80
 *
 * followed by the merge / sort / index / group statistics of
 *
 *     JunctionBuilder::findJunctions()               /root/reference/src/junction_builder.cc:228-291
 *     JunctionSystem::calcJunctionStats()            /root/reference/lib/src/junction_system.cc:250-320
 *
 * The entry points below are what a maintainer of the reference would bind in place of that seam
 * (see INTEGRATION.md for the C++ stub).  Plain pointers and sizes only; no exceptions cross the
 * boundary; every call returns 0 on success and a negative PJ_E* code on failure, with a message
 * available from pj_last_error().
 *
 * Threading: a pj_ctx is bound to ONE CUDA device.  Create one context per GPU (targets are independent,
 * so shards need no exchange).  Calls on one context must be serialised by the caller, with two exceptions
 * that exist for the decode pipeline: pj_staging_acquire / pj_batch_submit are internally locked (decode
 * workers may acquire and fill staging buffers concurrently while one thread submits them in BAM order), and
 * pj_genome_set_target / pj_genome_load_fasta use their own stream and may run on a second host thread
 * while batches are being staged and submitted.  pj_shard_run must not overlap any other call.
 */
#ifndef PORTCULLIS_JUNC_H
#define PORTCULLIS_JUNC_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PJ_ABI_VERSION 2

/* status codes */
#define PJ_OK            0
#define PJ_EINVAL       -1   /* bad argument / malformed input                         */
#define PJ_ECUDA        -2   /* CUDA runtime error (fatal: there is no CPU fallback)   */
#define PJ_ENOMEM       -3
#define PJ_ESTATE       -4   /* call sequence violated                                 */
#define PJ_EDATA        -5   /* input the reference itself aborts on (see Q6/Q12)      */
#define PJ_EIO          -6

/* Orientation, same order as portcullis::bam::Orientation
 * (/root/reference/lib/include/portcullis/bam/bam_master.hpp:141-147) */
enum { PJ_ORIENT_SE = 0, PJ_ORIENT_FR = 1, PJ_ORIENT_RF = 2, PJ_ORIENT_FF = 3, PJ_ORIENT_UNKNOWN = 4 };

/* Strand, same order as portcullis::bam::Strand (bam_master.hpp:50-54) */
enum { PJ_STRAND_POS = 0, PJ_STRAND_NEG = 1, PJ_STRAND_UNKNOWN = 2 };

/* MAPQ threshold for "uniquely mapped" (/root/reference/lib/include/portcullis/junction.hpp:65) */
#define PJ_MAP_QUALITY_THRESHOLD 30
#define PJ_NB_JAD 20

typedef struct pj_ctx pj_ctx;

typedef struct pj_config {
    int32_t device;        /* CUDA device ordinal                                              */
    int32_t orientation;   /* PJ_ORIENT_*; only FR/RF/FF enable the portcullis proper-pair rule */
    int32_t reserved[6];   /* tuning knobs, 0 = default. [0]: lanes per (read, junction) pair in the match kernel;
                              [1]: 1 = multi-kernel radix sort instead of the one-sweep sort;
                              [2]: number of pinned staging buffers pj_staging_acquire may create (default 4) */
    int32_t extra_metrics; /* 1 = also keep what the hidden `--extra` metrics need (pj_extra_* below); batches must
                              then carry name_code */
    int32_t pad;
} pj_config;

/*
 * A columnar batch of alignment records, in BAM (coordinate) order.  This replaces the stream of
 * bam1_t records that BamReader::next() hands to findJuncs (/root/reference/lib/src/bam_reader.cc:134-138,
 * record layout /root/reference/deps/htslib-1.3/htslib/sam.h:148-181).  Only the fields the path
 * reads are carried.  Records of several targets may share a batch; tid must be non-decreasing and
 * pos non-decreasing within a tid across the whole shard.
 *
 *   xs        : 0 when the record has no XS tag, otherwise the XS:A character ('+','-','?','.').
 *   cigar     : raw BAM CIGAR words (len<<4 | op), op index into "MIDNSHP=XB".
 *   cigar_off : n_records+1 prefix offsets into cigar[] (in words).
 *   seq4      : BAM 4-bit packed SEQ (high nibble first), each record starting on a byte boundary.
 *               Records without an N operation may omit their SEQ (seq_off[i+1]==seq_off[i]).
 *   seq_off   : n_records+1 prefix offsets into seq4[] (in bytes).
 *   name_code : optional (NULL unless pj_config.extra_metrics): a 64-bit code of BamAlignment::deriveName()
 *               (bam_alignment.cc:233-242: QNAME plus "_R1"/"_R2"/"_R?" for paired reads).  Only equality of codes
 *               matters — it plays the role of std::hash<string> in junction.hpp:158 / junction_builder.cc:182-185.
 */
typedef struct pj_batch {
    int64_t         n_records;
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
    const uint64_t* seq_off;
    const uint8_t*  seq4;
    const uint64_t* name_code;
    /* ---- lean form (ABI 2): what the host decoder ships, about half the bytes of the classic form over PCIe ----
     * A batch is lean when lean != 0.  Then tid / cigar_off / seq_off / seq4 are ignored and may be NULL:
     *   const_tid     : every record of the batch lies on this target (a decode task never spans targets);
     *   n_cigar       : ops per record (BAM's own n_cigar_op); cigar[] holds the n_cigar_total words back to back and the
     *                   prefix offsets are formed on the device (scan);
     *   seq2          : read bases at 2 bits each (A=0, C=1, G=2, T=3; base q of a record in bits 2(q&3)..2(q&3)+1 of its byte
     *                   q>>2), (l_qseq+3)/4 bytes for every record that has an N op and l_qseq > 0, nothing for the others,
     *                   back to back (n_seq2_bytes in total); the offsets are formed on the device from cigar + l_qseq;
     *   seqx_pos/code : the read bases that are not A/C/G/T (stored as 0 in seq2): base index within this batch's seq2
     *                   (byte offset * 4 + q), ascending, and the BAM nibble (index into "=ACMGRSVTWYHKDBN"); records that own
     *                   one must have bit 15 (0x8000, unused by the SAM spec) set in flag;
     *   mtid / mpos   : may be NULL when the context's orientation is SE or UNKNOWN (nothing reads them then). */
    int32_t         lean;
    int32_t         const_tid;
    const uint16_t* n_cigar;
    int64_t         n_cigar_total;
    const uint8_t*  seq2;
    int64_t         n_seq2_bytes;
    const uint64_t* seqx_pos;
    const uint8_t*  seqx_code;
    int64_t         n_seqx;
} pj_batch;

/* Per-target scalars: RegionResult minus the junction system
 * (/root/reference/src/junction_builder.hpp:72-80, filled at junction_builder.cc:352-356). */
typedef struct pj_target_stats {
    uint64_t spliced_count;
    uint64_t unspliced_count;
    uint64_t sum_query_lengths;
    int32_t  min_query_length;    /* INT32_MAX when the target has no records */
    int32_t  max_query_length;
} pj_target_stats;

/*
 * One junction row.  The first block is produced on the GPU; the block marked "host finalize" is
 * filled by pj_junctions_finalize() over the merged, sorted list (A12/A13: it needs every shard).
 * Field names follow the junctions.tab columns (/root/reference/lib/src/junction.cc:50-115).
 */
typedef struct pj_junction {
    int32_t  tid;
    int32_t  start;              /* intron start, 0-based inclusive */
    int32_t  end;                /* intron end,   0-based inclusive */
    int32_t  left;               /* leftAncStart  */
    int32_t  right;              /* rightAncEnd   */
    uint32_t nb_raw_aln;
    uint32_t nb_dist_aln;
    uint32_t nb_ms_aln;
    uint32_t nb_um_aln;
    uint32_t nb_bpp_aln;
    uint32_t nb_ppp_aln;
    uint32_t nb_rel_aln;
    uint32_t nb_r1_pos;
    uint32_t nb_r1_neg;
    uint32_t nb_r2_pos;
    uint32_t nb_r2_neg;
    uint32_t nb_xs_pos;          /* XS votes, inputs of the 0.95 read-strand rule */
    uint32_t nb_xs_neg;
    uint32_t max_min_anc;
    uint32_t maxmmes;
    uint32_t nb_mismatches;      /* uint32 sum, wraps like the reference */
    uint32_t hamming5p;
    uint32_t hamming3p;
    uint32_t nb_up_juncs;
    uint32_t nb_down_juncs;
    uint32_t jad[PJ_NB_JAD];
    uint32_t pad_a;              /* keeps entropy 8-byte aligned; sizeof(pj_junction) == 256 */
    double   entropy;
    uint8_t  read_strand;        /* PJ_STRAND_* */
    uint8_t  ss_strand;
    uint8_t  consensus_strand;
    uint8_t  canonical_ss;       /* 'C', 'S' or 'N' */
    uint8_t  suspicious;
    char     ss1[2];             /* da1 / da2 as printed in the ss1 / ss2 columns */
    char     ss2[2];
    uint8_t  pad0[7];
    /* ---- host finalize (pj_junctions_finalize) ---- */
    uint32_t index;
    uint32_t dist_2_up_junc;
    uint32_t dist_2_down_junc;
    uint32_t dist_nearest_junc;
    uint8_t  uniq_junc;
    uint8_t  primary_junc;
    uint8_t  pfp;
    uint8_t  pad1[5];
    double   mean_readlen;       /* (uint32) truncation of the mean query length, or 0 when J<=1 */
    double   rel2raw;
    double   mean_mismatches;
} pj_junction;

/* ---- lifecycle ---------------------------------------------------------------------------- */

int         pj_abi_version(void);
/* sizeof(pj_junction) as compiled into the library (256): lets bindings verify their mirror of the row layout */
int         pj_junction_size(void);
/* message of the last failure in this thread when no context exists (e.g. pj_create failed) */
const char* pj_global_last_error(void);

int         pj_create(const pj_config* cfg, pj_ctx** out);
void        pj_destroy(pj_ctx* ctx);
const char* pj_last_error(const pj_ctx* ctx);

/* ---- genome (replaces GenomeMapper / faidx_fetch_seq, genome_mapper.cc:111-118) ------------ */

/* Declare the target list exactly as the BAM header gives it (bam_reader.cc:103-110). */
int pj_targets_set(pj_ctx* ctx, int32_t n_targets, const int32_t* target_len);

/*
 * Upload one target's sequence: `bases` are the bytes faidx would return for the whole sequence
 * (only isgraph() bytes, any case).  The library upper-cases, packs to 2 bits per base plus an
 * exception bitmask on the device and keeps it resident in HBM.  `n_bases` may differ from the BAM
 * header length (the reference clamps fetches to the .fai length).
 */
int pj_genome_set_target(pj_ctx* ctx, int32_t tid, const char* bases, int64_t n_bases);

/*
 * Convenience: load every target straight from FASTA + .fai on disk.  `names` are the BAM header
 * target names in tid order; raw FASTA bytes are staged through pinned memory and unwrapped +
 * packed on the device.
 */
int pj_genome_load_fasta(pj_ctx* ctx, const char* fasta_path, const char* fai_path,
                         int32_t n_targets, const char* const* names);

/* ---- alignment shard ----------------------------------------------------------------------- */

/* Start a shard (a set of whole targets owned by this GPU).  Hints size the device arena. */
int pj_shard_begin(pj_ctx* ctx, int64_t n_records_hint, int64_t n_cigar_hint, int64_t n_seq_bytes_hint);

/*
 * Pinned staging pool: returns writable columnar buffers of at least the requested capacities (the pointers in
 * `out` are NON-const views of pinned host memory owned by the context; n_records is set to 0).  Fill them, set
 * n_records, then pj_batch_submit().  A buffer returns to the pool when its host-to-device copies have completed;
 * acquiring blocks only while every buffer of the pool is in flight.
 */
int pj_staging_acquire(pj_ctx* ctx, int64_t cap_records, int64_t cap_cigar, int64_t cap_seq_bytes, pj_batch* out);
/* Same for a lean batch (out->lean = 1): pos, flag, mapq, xs, l_qseq, mtid, mpos, n_cigar, cigar, seq2, seqx_pos, seqx_code
 * (and name_code) point into the pinned buffer; the caller fills them and sets n_records, const_tid, n_cigar_total,
 * n_seq2_bytes and n_seqx before pj_batch_submit. */
int pj_staging_acquire_lean(pj_ctx* ctx, int64_t cap_records, int64_t cap_cigar, int64_t cap_seq2_bytes, int64_t cap_seqx, pj_batch* out);

/* Enqueue host->device copies of a batch (cudaMemcpyAsync on the context's copy stream) and
 * append it to the shard.  `b` may be a staging view or any caller-owned host buffers (those are
 * copied synchronously unless they are pinned). */
int pj_batch_submit(pj_ctx* ctx, const pj_batch* b);

/* Run the device pipeline over everything submitted since pj_shard_begin. */
int pj_shard_run(pj_ctx* ctx);

/* Number of junctions found by the last pj_shard_run. */
int64_t pj_shard_num_junctions(const pj_ctx* ctx);

/* Copy results to caller-allocated arrays: rows[J] in (tid,start,end) order, stats[n_targets]. */
int pj_shard_fetch(pj_ctx* ctx, pj_junction* rows, int64_t cap_rows, pj_target_stats* stats, int32_t cap_targets);

/* Device time (ms, CUDA events on the compute stream) of the last pj_shard_run, plus the number of
 * kernel launches it made.  kernel_ms/kernel_names expose the per-kernel breakdown (n entries). */
int pj_shard_timing(const pj_ctx* ctx, float* total_ms, int32_t* n_launches);
int pj_shard_kernel_times(const pj_ctx* ctx, int32_t cap, float* kernel_ms, const char** kernel_names, int32_t* n);

/* ---- `--extra` metrics (JunctionBuilder::calcExtraMetrics, junction_builder.cc:293-312) ---- */

/*
 * The four junctions.tab columns plain `junc` leaves 0.  The reference computes them from the separated BAMs
 * (spliced / unspliced); here the same record classes are taken from the columnar batches already in HBM:
 *   spliced   = any N op in the CIGAR (BamAlignment::isSplicedRead, bam_alignment.cc:294-301)
 *   unspliced = not spliced and mapped (junction_builder.cc:188-191)
 * Integer columns are produced on the GPU; the two doubles are filled by pj_extra_finalize() on the host.
 */
typedef struct pj_junction_extra {
    uint32_t up_aln;             /* nbUpstreamFlankingAlignments   (Junction::processJunctionVicinity, junction.cc:651-677) */
    uint32_t down_aln;           /* nbDownstreamFlankingAlignments                                                       */
    uint32_t mm_n;               /* N of calcMultipleMappingScore (junction.cc:914-921): alignments of the junction        */
    uint32_t mm_m;               /* M: sum over them of the number of spliced alignments with the same name (uint16 map
                                    values, uint32 sum — both wrap like the reference)                                     */
    uint32_t cov_sum[4];         /* read counts of Junction::calcCoverage (junction.cc:935-951), in its call order:
                                    [start-20,start-11], [start-10,start], [end+10,end+20], [end,end+9]                  */
    double   mm_score;           /* (double)N / (double)M                                                                  */
    double   coverage;
} pj_junction_extra;

/* 64-bit codes of the shard's spliced records (device -> host), so that a multi-GPU caller can give every context the
 * names of the whole file: the reference's map is built over the whole BAM (junction_builder.cc:179-186). */
int64_t pj_extra_num_spliced_names(const pj_ctx* ctx);
int pj_extra_export_names(pj_ctx* ctx, uint64_t* codes, int64_t cap);
/* Add spliced-read name codes that live on OTHER contexts (host -> device). */
int pj_extra_import_names(pj_ctx* ctx, const uint64_t* codes, int64_t n);

/*
 * After pj_shard_run: up_aln / down_aln / mm_n / mm_m for the shard's junctions, in pj_shard_fetch order
 * (cov_sum, mm_score and coverage are left 0).  max_query_length is the maximum over ALL targets of the file
 * (JunctionSystem::setQueryLengthStats, junction_builder.cc:270): it sizes the region the reference queries.
 */
int pj_extra_run(pj_ctx* ctx, int32_t max_query_length, pj_junction_extra* out, int64_t cap_rows);

/* Device time (CUDA events) and kernel launches of the last pj_extra_run, and its per-stage breakdown (x_unspliced,
 * x_mm_score, x_flank, x_live, [x_cap,] x_depth) — same conventions as pj_shard_timing / pj_shard_kernel_times. */
int pj_extra_timing(const pj_ctx* ctx, float* total_ms, int32_t* n_launches);
int pj_extra_kernel_times(const pj_ctx* ctx, int32_t cap, float* kernel_ms, const char** kernel_names, int32_t* n);

/*
 * Unspliced pileup of one target held by this context (DepthParser, depth_parser.cc:112-167): *covered = 1 when the
 * pileup reports at least one position, *max_depth = the largest number of reads alive on one position (reads with
 * pos <= p <= endpos, before any are dropped).  htslib stops accepting reads that start on a position whose node pool
 * already holds 8000 (sam.c:1622, 1906); where max_depth reaches 8000 pj_extra_run replays that rule on the device, so the
 * depth vectors equal the reference's there too.  Informational.
 */
int pj_extra_target_pileup(pj_ctx* ctx, int32_t tid, int32_t* covered, uint32_t* max_depth);

/*
 * cov_sum[4] of n junction coordinates against the depth vector of target depth_tid (which must belong to this
 * context's shard).  The reference pairs each depth vector with the junctions of the NEXT covered target
 * (depth_parser.hpp:94-96 returns the index of the target the parser has just moved on to; quirk Q14) — the caller
 * chooses depth_tid accordingly, see pj_extra_coverage_source().
 */
int pj_extra_coverage(pj_ctx* ctx, int32_t depth_tid, int64_t n, const int32_t* intron_start, const int32_t* intron_end,
                      uint32_t* cov_sum4);
/* The same for junctions of any number of targets in ONE launch: depth_tid[j] = the target whose depth vector junction j is
 * scored against (pj_extra_coverage_source of the junction's own target; -1: none, the sums are 0).  The depth vectors must
 * live in this context. */
int pj_extra_coverage_batch(pj_ctx* ctx, int64_t n, const int32_t* depth_tid, const int32_t* intron_start, const int32_t* intron_end, uint32_t* cov_sum4);

/* Q14 as a function: covered[t] per target -> depth_src[t] = target whose depth vector the reference applies to the
 * junctions of t (-1: coverage stays 0). */
void pj_extra_coverage_source(int32_t n_targets, const uint8_t* covered, int32_t* depth_src);

/* mm_score and coverage from the integer columns, with the reference's operation order. */
void pj_extra_finalize(pj_junction_extra* x, int64_t n);

/* ---- coordinate sort for `prep` (replaces `samtools sort`, src/prepare.cc:202-236) ---------------- */

/*
 * order[k] = index of the record that comes k-th in samtools' coordinate order: key = tid << 32 | (pos + 1) << 1 |
 * reverse-strand flag (bam_sort.c bam1_lt), unplaced reads (tid = -1) last, equal keys in input order (stable).  One-sweep
 * LSD radix sort of 64-bit keys on the given device; n < 2^30.  No context needed.
 */
int pj_coordinate_order(int32_t device, int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, uint32_t* order);

/* ---- junction-set membership for `bamfilt` (BamFilter, src/bam_filter.cc:75-245) ------------------- */

typedef struct pj_jset pj_jset;
/* A junction set resident on `device`: n introns sorted by (tid, start, end), coordinates as in junctions.tab columns
 * refid / start / end (JunctionSystem::load + Intron key, junction_system.cc:424-444). */
int  pj_jset_create(int32_t device, int64_t n, const int32_t* tid, const int32_t* start, const int32_t* end, pj_jset** out);
void pj_jset_destroy(pj_jset* s);
/* keep[i] = 1 when record i has no N op, or at least one of its N ops is in the set (containsJunctionInSystem,
 * bam_filter.cc:75-99 — the condition all three clip modes reduce to, see csrc/pj_bamfilt.cu); n_nops[i] (optional) = its
 * number of N ops, saturated at 255.  cigar_off has n_records + 1 entries starting at 0. */
int  pj_jset_filter(pj_jset* s, int64_t n_records, const int32_t* tid, const int32_t* pos, const uint32_t* cigar_off,
                    const uint32_t* cigar, uint8_t* keep, uint8_t* n_nops);

/* ---- host finalize (A12/A13) --------------------------------------------------------------- */

/*
 * rows: all junctions of all shards, concatenated.  Sorts by (tid,start,end), assigns index, and —
 * only when n_rows > 1, like junction_builder.cc:285 — group / distance / mean_readlen / pfp
 * statistics.  Always fills rel2raw and mean_mismatches.
 */
int pj_junctions_finalize(pj_junction* rows, int64_t n_rows, double mean_query_length);

/* ---- `filt` feature extraction (portcullis::ml::ModelFeatures, lib/src/model_features.cc) ----
 * The per-junction feature vector the reference's filter trains its random forest on, computed on the device from junction
 * rows and the genome already resident in the context (pj_targets_set + pj_genome_set_target / pj_genome_load_fasta):
 *   pj_features_train_coding   = ModelFeatures::trainCodingPotentialModel  (model_features.cc:70-110): exon / intron 5th-order
 *                                k-mer Markov models from the junctions with subset[i] != 0 (NULL: all)
 *   pj_features_train_splicing = ModelFeatures::trainSplicingModels        (:112-166): donor / acceptor k-mer and positional
 *                                models from pass[i] != 0, the "false" k-mer models from fail[i] != 0
 *   pj_features_intron_threshold = ModelFeatures::calcIntronThreshold      (:61-68), host only
 *   pj_features_run            = ModelFeatures::juncs2FeatureVectors / setRow (:168-228): out[n_rows][PJ_NB_FEATURES] =
 *       Genuine(0), rna_usrs, rna_dist, rna_rel, rna_entropy, rna_rel2raw, rna_maxminanc, rna_maxmmes, rna_missmatch,
 *       rna_intron (Junction::calcIntronScore, junction.cc:953-956), dna_minhamm, dna_coding (calcCodingPotential :1328-1358),
 *       dna_pws, dna_ss (calcSplicingScores :1360-1382), JAD01..JAD20 log deviations (calcJunctionAnchorDepthLogDeviation :1384-1391).
 * Counts are integer atomics and every probability product is formed left to right like the reference, so the products are
 * bit-identical; only the final log() may differ in the last ulp.  rows must carry the host-finalized fields (mean_readlen,
 * rel2raw, mean_mismatches).  device_ms (may be NULL) receives the kernel time of pj_features_run. */
#define PJ_NB_FEATURES 34
typedef struct pj_feat_models pj_feat_models;
int      pj_features_create(pj_ctx* ctx, pj_feat_models** out);
void     pj_features_destroy(pj_feat_models* m);
int      pj_features_train_coding(pj_feat_models* m, const pj_junction* rows, int64_t n_rows, const uint8_t* subset);
int      pj_features_train_splicing(pj_feat_models* m, const pj_junction* rows, int64_t n_rows, const uint8_t* pass, const uint8_t* fail);
uint32_t pj_features_intron_threshold(const pj_junction* rows, int64_t n_rows, const uint8_t* subset);
int      pj_features_run(pj_feat_models* m, const pj_junction* rows, int64_t n_rows, uint32_t l95, double* out, float* device_ms);

#ifdef __cplusplus
}
#endif
#endif /* PORTCULLIS_JUNC_H */
