/*
 * portcullis_junc_host.h — host-side C entry points that sit above the GPU C ABI (portcullis_junc.h):
 * the `junc` stage driver (the JunctionBuilder equivalent), prep-directory access and the output writers.
 *
 * Reference interfaces mirrored here:
 *   JunctionBuilder(prepDir, output) + setters + process()   /root/reference/src/junction_builder.cc:63-150
 *   PreparedFiles (prep-dir naming contract)                  /root/reference/src/prepare.hpp:114-140
 *   JunctionSystem::saveAll                                   /root/reference/lib/src/junction_system.cc:336-383
 */
#ifndef PORTCULLIS_JUNC_HOST_H
#define PORTCULLIS_JUNC_HOST_H

#include "portcullis_junc.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Strandedness, same order as portcullis::bam::Strandedness (bam_master.hpp:99-104) */
enum { PJ_STRANDED_UNSTRANDED = 0, PJ_STRANDED_FIRSTSTRAND = 1, PJ_STRANDED_SECONDSTRAND = 2, PJ_STRANDED_UNKNOWN = 3 };

typedef struct pjh_options {
    const char* prep_dir;        /* positional <prep_data_dir>                                   */
    const char* output_prefix;   /* -o, default "portcullis_junc/portcullis"                      */
    int32_t threads;             /* -t: host decode threads (>=1)                                 */
    int32_t n_gpus;              /* --gpus: number of B200s to shard targets over (>=1)           */
    const int32_t* gpu_ids;      /* optional explicit device ordinals (n_gpus entries) or NULL    */
    int32_t orientation;         /* PJ_ORIENT_*                                                   */
    int32_t strandedness;        /* PJ_STRANDED_* (only compared with the detected protocol)      */
    int32_t use_csi;             /* -c                                                            */
    int32_t exon_gff;            /* --exon_gff                                                    */
    int32_t intron_gff;          /* --intron_gff                                                  */
    const char* source;          /* --source, default "portcullis"                                */
    int32_t verbose;             /* -v                                                            */
    int32_t separate;            /* --separate: also write <prefix>.spliced/.unspliced/.unmapped.bam (+ indices) */
    int32_t extra;               /* --extra: mm_score, coverage, up_aln, down_aln from the records in HBM   */
    int32_t quiet;               /* suppress the progress text on stdout                          */
    const char* version;         /* string for the BED track line; NULL -> "1.2.4"                */
} pjh_options;

typedef struct pjh_report {
    int64_t n_junctions;
    uint64_t n_spliced, n_unspliced;
    double  mean_query_length;
    int32_t min_query_length, max_query_length;
    double  t_open_s;            /* open BAM/BAI/FASTA, plan                                      */
    double  t_genome_s;          /* FASTA unwrap + upload + pack                                  */
    double  t_decode_s;          /* BGZF inflate + columnar packing + H2D (overlapped)            */
    double  t_gpu_ms;            /* max over GPUs of the device pipeline time (CUDA events)       */
    double  t_finalize_s;        /* host merge + A12/A13                                          */
    double  t_write_s;           /* tab/bed/gff writers                                           */
    double  t_total_s;
    int32_t n_gpus_used;
    int32_t n_kernel_launches;
    double  t_init_s;            /* CUDA context + pj_create + pj_targets_set (max over GPUs)     */
    double  t_run_s;             /* wall time of pj_shard_run + pj_shard_fetch (max over GPUs)    */
    double  t_teardown_s;        /* pj_destroy: what is left of it after the writers (it runs on a helper thread under them) */
    double  t_extra_s;           /* --extra: name exchange + pj_extra_run + coverage (0 otherwise) */
    double  t_separate_s;        /* --separate: splitting + indexing the BAM on the host (0 otherwise) */
    int32_t n_segments;          /* shards run (pj_shard_begin ... pj_shard_fetch), summed over GPUs    */
    int32_t n_gap_cuts;          /* segment boundaries placed inside a target (at a position no spliced read spans) */
} pjh_report;

void pjh_options_default(pjh_options* o);
/* Runs the whole `junc` stage. Returns 0 or a PJ_E* code; message in pjh_last_error(). */
int pjh_junc_run(const pjh_options* opt, pjh_report* report);
const char* pjh_last_error(void);

/* ---- one process per GPU (torchrun / MPI style launch; the reference's counterpart is one task per target on its thread
 * pool, src/junction_builder.cc:109-112, 459-542) ----
 * Every process computes the same work plan: the BAM-ordered decode tasks, weighted by the index's record counts, are cut
 * into n_parts contiguous ranges; a cut inside a target is moved to the next record no spliced read spans, so no junction
 * (key: refid, start, end — lib/include/portcullis/intron.hpp:69-73) has reads in two parts.  pjh_junc_run_part decodes and
 * runs part `part` on device opt->gpu_ids[0] (default 0) with opt->threads decode threads and keeps the rows (already in
 * (tid, start, end) order; parts concatenate in part order) and the per-target scalars.  No collective is involved: the
 * caller moves the rows to one process (shared memory, a file, MPI_Gatherv ...) and calls pjh_junc_finish there, which
 * merges, runs A12/A13 (pj_junctions_finalize) and writes the output files exactly like pjh_junc_run. */
typedef struct pjh_partial pjh_partial;
int     pjh_junc_run_part(const pjh_options* opt, int32_t part, int32_t n_parts, pjh_partial** out, pjh_report* report);
int64_t pjh_partial_rows(const pjh_partial* p, const pj_junction** rows);          /* returns the row count               */
int32_t pjh_partial_stats(const pjh_partial* p, const pj_target_stats** stats);    /* returns n_targets; one entry each   */
void    pjh_partial_free(pjh_partial* p);                                            /* must be called: it also waits for the part's GPU context, which
                                                                                    * is torn down on a helper thread while the caller gathers the rows */
/* rows[n_rows]: the parts' rows concatenated in part order (finalized in place); stats[n_targets]: per-target scalars summed
 * over the parts (counts and sums added, min/max combined; a part that saw no record of a target reports min = INT32_MAX). */
int     pjh_junc_finish(const pjh_options* opt, pj_junction* rows, int64_t n_rows, const pj_target_stats* stats, int32_t n_targets,
                        pjh_report* report);

/* `portcullis junc ...` command line (argv[0] is the mode word).  Returns the process exit code. */
int pjh_junc_main(int argc, char** argv);

/* ---- prep directory access (used by tests and the benchmark harness) ---- */
typedef struct pjh_prep pjh_prep;
int         pjh_prep_open(const char* prep_dir, int use_csi, pjh_prep** out);
void        pjh_prep_close(pjh_prep* p);
int32_t     pjh_prep_n_targets(const pjh_prep* p);
const char* pjh_prep_target_name(const pjh_prep* p, int32_t tid);
int32_t     pjh_prep_target_len(const pjh_prep* p, int32_t tid);
/* records the index reports for a target (mapped + placed unmapped), -1 when unknown */
int64_t     pjh_prep_target_records(const pjh_prep* p, int32_t tid);
/* Decode one target (tid >= 0) or every target (tid = -1) with `threads` workers into columnar arrays owned by
 * `p` (valid until the next decode or close). */
int         pjh_prep_decode(pjh_prep* p, int32_t tid, int32_t threads, pj_batch* out);
/* on != 0: later pjh_prep_decode calls also fill pj_batch.name_code (needed by the `--extra` metrics) */
void        pjh_prep_want_names(pjh_prep* p, int32_t on);
/* Unwrapped FASTA bytes of a target (owned by `p`, valid until the next call for another target or close). */
int         pjh_prep_genome(pjh_prep* p, int32_t tid, const char** bases, int64_t* n_bases);

/* Shard plan: gpu_of_target[tid] in [0, n_gpus) — whole targets, longest processing time first on the index's record
 * counts (compressed bytes when the index has no counts).  Deterministic, so every rank of a multi-process launch
 * computes the same plan. */
int         pjh_plan_shards(const pjh_prep* p, int32_t n_gpus, int32_t* gpu_of_target);

/* The range plan pjh_junc_run / pjh_junc_run_part execute (see pjh_junc_run_part): n_parts contiguous record-balanced ranges,
 * each cut into segments of at most seg_records records (<= 0: the default, 32 Mi); whole_targets != 0 gives the `--extra`
 * plan instead (LPT of whole targets, one segment per part).  segments_per_part[n_parts] receives the number of segments of
 * every part; n_gap_cuts the number of boundaries that fell inside a target.  pjh_plan_decode decodes one segment into
 * columnar arrays owned by `p` (valid until the next decode).  Used by the tests: the segments partition the records of the
 * BAM in file order, and no spliced record of an earlier segment reaches the first position of a later one. */
int         pjh_plan_describe(const pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t* segments_per_part, int32_t* n_gap_cuts);
int         pjh_plan_decode(pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t part, int32_t segment, int32_t threads, pj_batch* out);

/* Same segment in the LEAN batch form (pj_batch.lean): what the junc driver ships over PCIe.  `out` describes the whole segment
 * (const_tid = -1); a lean batch must lie on one target, so runs[n_runs] lists the stretches of the segment — submit
 * records [rec0, next rec0) of each with const_tid = tid, the CIGAR words from cig0, the seq2 bytes from seq0 and the
 * exceptions from seqx0 (their positions made relative to seq0 * 4).  keep_mate = 0 omits mtid / mpos. */
typedef struct pjh_lean_run { int32_t tid; int32_t pad; int64_t rec0, cig0, seq0, seqx0; } pjh_lean_run;
int         pjh_plan_decode_lean(pjh_prep* p, int32_t n_parts, int32_t whole_targets, int64_t seg_records, int32_t part, int32_t segment, int32_t threads,
                                 int32_t keep_mate, pj_batch* out, const pjh_lean_run** runs, int32_t* n_runs);

/* Checks the built-in fast DEFLATE decoder (BGZF blocks) against zlib on n_cases synthetic streams; returns the number of
 * mismatches (0 = pass).  BGZF blocks the fast decoder rejects are decoded by zlib, so it can only be an accelerator. */
int         pjh_inflate_selftest(int32_t n_cases);
/* Self-test of the writers' number formatting (fast "%g" / integer paths) against printf; returns the number of differences. */
int         pjh_format_selftest(int32_t n_cases);

/* ---- `--separate` (JunctionBuilder::separateBams, src/junction_builder.cc:152-226) ----
 * Splits the prepared BAM into <output_prefix>.spliced.bam (any N op), .unspliced.bam (mapped, no N) and .unmapped.bam and
 * indexes the first two (BAI, or CSI with use_csi).  Host only: pjh_junc_run calls it first when options.separate is set.
 * counts[3] = records written to spliced / unspliced / unmapped. */
int pjh_separate_bams(const char* prep_dir, const char* output_prefix, int32_t use_csi, int32_t threads, uint64_t* counts);

/* ---- `prep` without samtools (Prepare::prepare, src/prepare.cc:262-344) ----
 * Lays out <output_dir>/portcullis.genome.fa[.fai] and portcullis.sorted.alignments.bam[.bai|.csi].  A single BAM whose header
 * says SO:coordinate is symlinked (or copied); anything else — an unsorted BAM, several BAMs, or --force — is sorted / merged
 * in process: records in host memory, coordinate order from pj_coordinate_order on `device`. */
typedef struct pjh_prep_options {
    const char* genome_file;
    const char* const* bam_files; int32_t n_bam_files;
    const char* output_dir;      /* -o, default "portcullis_prep"                                  */
    int32_t force;               /* --force: clean the directory first, always re-sort             */
    int32_t copy;                /* --copy: copy instead of symlink                                 */
    int32_t use_csi;             /* -c                                                              */
    int32_t threads;             /* -t: host threads for BGZF inflate / deflate                     */
    int32_t verbose, quiet;
    int32_t device;              /* GPU used for the coordinate order                               */
} pjh_prep_options;
typedef struct pjh_prep_report {
    int64_t n_records;           /* records written by the in-process sort (0: input was linked)    */
    int32_t sorted_in_process;
    double  t_sort_s, t_sort_gpu_s, t_total_s;
} pjh_prep_report;
void pjh_prep_options_default(pjh_prep_options* o);
int  pjh_prep_run(const pjh_prep_options* opt, pjh_prep_report* report);
const char* pjh_prep_last_error(void);
/* `portcullis prep ...` command line (argv[0] is the mode word).  Returns the process exit code. */
int  pjh_prep_main(int argc, char** argv);

/* ---- `bamfilt` (BamFilter::filter, src/bam_filter.cc:152-245) ----
 * Keeps every unspliced alignment and every spliced alignment with at least one junction in junction_file (a junctions.tab);
 * writes output_bam (+ .bai / .csi) and, with save_msrs, <output>.mod.bam / .unmod.bam.  The decision runs on `device`
 * (pj_jset_filter); the files are laid out like htslib's writer lays them out. */
enum { PJ_CLIP_HARD = 0, PJ_CLIP_SOFT = 1, PJ_CLIP_COMPLETE = 2 };
typedef struct pjh_bamfilt_options {
    const char* junction_file; const char* bam_file;
    const char* output_bam;      /* -o, default "filtered.bam"                                      */
    int32_t clip_mode;           /* PJ_CLIP_*: changes the Modified count only, like the reference  */
    int32_t save_msrs, use_csi, threads, verbose, quiet, device;
} pjh_bamfilt_options;
typedef struct pjh_bamfilt_report { int64_t n_junctions; uint64_t n_in, n_out, n_modified; double t_device_s, t_total_s; } pjh_bamfilt_report;
void pjh_bamfilt_options_default(pjh_bamfilt_options* o);
int  pjh_bamfilt_run(const pjh_bamfilt_options* opt, pjh_bamfilt_report* report);
const char* pjh_bamfilt_last_error(void);
int  pjh_bamfilt_main(int argc, char** argv);

/* ---- writers (A14) ---- */
int pjh_write_outputs(const char* output_prefix, const pj_junction* rows, int64_t n_rows,
                      int32_t n_targets, const char* const* names, const int32_t* lens,
                      const char* source, const char* version, int32_t exon_gff, int32_t intron_gff);

/* Same, with the `--extra` columns (mm_score, coverage, up_aln, down_aln) taken from extra[n_rows] (NULL: printed as 0). */
int pjh_write_outputs_extra(const char* output_prefix, const pj_junction* rows, const pj_junction_extra* extra, int64_t n_rows,
                            int32_t n_targets, const char* const* names, const int32_t* lens,
                            const char* source, const char* version, int32_t exon_gff, int32_t intron_gff);

#ifdef __cplusplus
}
#endif
#endif
