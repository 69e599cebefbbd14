/*
 * junc_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded restatement of the reference `junc` hot path (rows A1-A13 of SURVEY.md §8),
 * operating on the same columnar input as the CUDA library (include/portcullis_junc.h) and producing
 * the same pj_junction rows.  It exists to CHECK the CUDA path; nothing under portcullis_b200/ may
 * link, import or call it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it.
 *
 * Parity pinning: validated against outputs of the UNMODIFIED reference binary (oracle/_ref/portcullis_ref,
 * built from /root/reference by oracle/Makefile) — see tests/test_oracle_vs_reference.py and the
 * committed fixtures under tests/golden/ — and against the reference's own unit-test vectors
 * (tests/bam_tests.cpp:181-248, tests/junction_tests.cpp:48-107, tests/intron_tests.cpp, tests/seq_utils_tests.cpp).
 */
#ifndef JUNC_ORACLE_H
#define JUNC_ORACLE_H
#include "../include/portcullis_junc.h"

#ifdef __cplusplus
extern "C" {
#endif

/*
 * genome_cat / genome_off: per-target sequence bytes exactly as faidx_fetch_seq would return them for the
 * whole sequence (isgraph bytes, original case), concatenated; genome_off has n_targets+1 entries.
 * On success *rows_out is a malloc'd array of *n_rows_out rows sorted by (tid,start,end) with the host-finalize
 * block zeroed (call oj_finalize for A12/A13); free with oj_free.  stats must have n_targets entries.
 * Returns 0, or a negative PJ_E* code with a message in oj_last_error().
 */
int oj_run(const pj_batch* b, int32_t n_targets, const int32_t* target_len,
           const char* genome_cat, const int64_t* genome_off, int32_t orientation,
           pj_junction** rows_out, int64_t* n_rows_out, pj_target_stats* stats);
void oj_free(void* p);
const char* oj_last_error(void);

/* A12/A13 restated: sort, index, groups, neighbour distances, mean_readlen, pfp, rel2raw, mean_mismatches. */
int oj_finalize(pj_junction* rows, int64_t n_rows, double mean_query_length);

/* `--extra` metrics (junction_builder.cc:152-226, 293-312) restated over the same batch (which must
 * carry name_code and hold the whole file in BAM order).  rows: the finalized rows of oj_run for that batch.
 * *n_capped_reads = reads htslib's pileup dropped at its 8000-read cap (sam.c:1906), which this restatement models. */
int oj_extra(const pj_batch* b, int32_t n_targets, const int32_t* target_len, const pj_junction* rows, int64_t n_rows,
             int32_t max_query_length, pj_junction_extra* out, int64_t* n_capped_reads);

/* ---- unit-level entry points used to replay the reference's own known-answer tests ---- */

/* BamAlignment::getPaddedQuerySeq (bam_alignment.cc:341-403), include_soft_clips=false.
 * query: full read as characters. Returns string length or <0; out must hold >= query_len + window + 8 bytes. */
int oj_padded_query(int32_t pos, const uint32_t* cigar, int32_t n_cigar, const char* query, int32_t query_len,
                    int32_t start, int32_t end, char* out, int32_t* actual_start, int32_t* actual_end);
/* BamAlignment::getPaddedGenomeSeq (bam_alignment.cc:405-462). genome: bases of [start,end]. */
int oj_padded_genome(int32_t pos, const uint32_t* cigar, int32_t n_cigar, const char* genome, int32_t genome_len,
                     int32_t start, int32_t end, int32_t q_start, int32_t q_end, char* out);
/* Junction::calcEntropy(vector<int32_t>) (junction.cc:730-749); positions must be sorted. */
double oj_entropy(const int32_t* positions, int64_t n);
/* SeqUtils::hammingDistance (seq_utils.hpp:62-77) and reverseComplement (:111-118). */
int oj_hamming(const char* a, const char* b, int32_t n);
void oj_revcomp(const char* in, int32_t n, char* out);
/* Junction::hasCanonicalSpliceSites / predictedStrandFromSpliceSites (junction.cc:289-326):
 * returns 'C','S','N'; *ss_strand gets PJ_STRAND_*. */
int oj_splice_motif(const char* donor2, const char* acceptor2, int32_t* ss_strand);
/* Intron::minAnchorLength (intron.cc:67-87). */
int32_t oj_min_anchor(int32_t start, int32_t end, int32_t left, int32_t right);

#ifdef __cplusplus
}
#endif
#endif
