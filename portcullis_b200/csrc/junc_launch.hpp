// junc_launch.hpp — host-callable launchers of the kernels in junc_kernels.cu.
#pragma once
#include "junc_kernels.cuh"

namespace pjk {

// per-target accumulators of k_scan_reads (RegionResult scalars)
struct TargetAcc {
    unsigned long long* spliced; unsigned long long* unspliced; unsigned long long* sumq;
    int32_t* minq; int32_t* maxq;
};

// per-junction accumulators, struct of arrays (every field is an integer reduction -> atomics are deterministic)
struct JuncAcc {
    int32_t *tid, *start, *end, *left, *right;
    uint32_t *r1p, *r1n, *r2p, *r2n, *ms, *um, *bpp, *ppp, *rel, *xsp, *xsn, *dist, *anc, *up, *down;
    uint32_t *maxmmes, *mism, *firstmm, *maxminmatch;
    uint32_t *jadhist;          // [n_junc][PJ_NB_JAD+1] histogram of min(minMatch, 20)
};

void launch_exclusive_scan(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* bsum_tmp, uint32_t* total_dev, cudaStream_t st);
uint64_t scan_tmp_elems(uint64_t n);
void launch_coord_keys(int64_t n, const int32_t* tid, const int32_t* pos, const uint16_t* flag, uint64_t* keys, uint32_t* vals, cudaStream_t st);
void launch_fill_i32(int32_t* a, int64_t n, int32_t v, cudaStream_t st);
void launch_rebase_u32(uint32_t* a, int64_t n, uint32_t add, cudaStream_t st);
void launch_check_offsets(int64_t first, int64_t n, const uint32_t* cigar_off, const uint64_t* seq_off, uint64_t cig_lo, uint64_t cig_hi,
                          uint64_t seq_lo, uint64_t seq_hi, unsigned long long* bad, cudaStream_t st);
void launch_rebase_u64(uint64_t* a, int64_t n, uint64_t add, cudaStream_t st);
void launch_pack_genome(const uint8_t* raw, int64_t n, uint64_t base_index, uint64_t* g2, uint64_t* gx, uint32_t* gsum,
                        uint64_t* exc_pos, uint8_t* exc_byte, uint32_t* exc_count, uint32_t exc_cap, cudaStream_t st);
void launch_prescan_cigar(const uint32_t* cigar, uint64_t n, uint32_t* max_nlen, unsigned long long* n_nops, int n_sm, cudaStream_t st);
uint32_t se_num_tiles(int64_t n);
void launch_scan_emit(const Reads& R, const int32_t* tlen, int32_t n_targets, const uint64_t* toff, const uint32_t* max_nlen, int32_t orientation,
                      const TargetAcc& T, uint64_t* keys, PairRec* pr, unsigned long long* status, uint32_t* ticket,
                      uint32_t* total_pairs, uint32_t pair_cap, uint32_t* err, cudaStream_t st);
uint32_t rs_num_blocks(uint32_t n);
int launch_radix_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int key_bits,
                      uint32_t* counts, uint32_t* scan_tmp, uint32_t* total_tmp, cudaStream_t st, int* n_launches);
size_t os_scratch_words(uint32_t n, int key_bits);
int launch_onesweep_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, uint32_t n, int key_bits,
                         uint32_t* scratch, int n_sm, cudaStream_t st, int* n_launches);
uint32_t fs_num_tiles(uint32_t n);
size_t fs_scratch_bytes(uint32_t n);
void launch_cigar_off(uint32_t n, const uint16_t* ncig, uint32_t* cigar_off_at_R, uint32_t base, uint32_t expect, unsigned long long* bad, unsigned long long* scratch, cudaStream_t st);
void launch_seq_off(uint32_t n, const uint32_t* cigar_off_at_R, const uint32_t* cigar, const int32_t* lq_at_R, uint64_t* seq_off_at_R, uint64_t base, uint64_t expect,
                    unsigned long long* bad, unsigned long long* scratch, cudaStream_t st);
void launch_seq4_to_2(int64_t n, const uint8_t* s4, const uint64_t* off4, const int32_t* lq, const uint64_t* seq_off, uint8_t* seq2, uint16_t* flag, uint32_t* xcount, cudaStream_t st);
void launch_seq4_exceptions(int64_t n, const uint8_t* s4, const uint64_t* off4, const int32_t* lq, const uint64_t* seq_off, const uint32_t* xcount, const uint32_t* xoff,
                            uint64_t* xpos, uint8_t* xcode, cudaStream_t st);
void launch_segment(const uint64_t* keys, uint32_t n, uint32_t* jid, uint32_t* seg_start, uint32_t* n_junc_dev, unsigned long long* scratch, cudaStream_t st);
void launch_entropy_index(uint32_t n, const uint32_t* eflag, uint32_t* eoff, uint32_t* epos, uint32_t* total_dev, unsigned long long* scratch, cudaStream_t st);
void launch_seg_heads(const uint64_t* keys, uint32_t n, uint32_t* head, cudaStream_t st);
void launch_seg_ids(const uint64_t* keys, uint32_t n, const uint32_t* excl, uint32_t* jid, uint32_t* seg_start, uint32_t n_junc, cudaStream_t st);
void launch_junc_init(uint32_t n_junc, const uint32_t* seg_start, const uint64_t* keys, const uint32_t* vals, const PairRec* pr,
                      const int32_t* read_tid, int32_t len_bits, const JuncAcc& A, cudaStream_t st);
void launch_reduce1(uint32_t n, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, int32_t ppcheck,
                    const JuncAcc& A, uint32_t* eflag, cudaStream_t st);
void launch_entropy_compact(uint32_t n, const uint32_t* eflag, const uint32_t* eoff, uint32_t* epos, cudaStream_t st);
void launch_entropy_sum(uint32_t n_junc, const uint32_t* seg_start, const uint32_t* eoff, const uint32_t* epos, double* entropy, cudaStream_t st);
void launch_match(uint32_t n, int group, int ctas, const uint32_t* vals, const uint32_t* jid, const PairRec* pr, const Reads& R, const Genome& G,
                  const JuncAcc& A, uint4* pm, uint32_t* err, cudaStream_t st);
void launch_reduce2(uint32_t n, const uint32_t* jid, const uint4* pm, const JuncAcc& A, cudaStream_t st);
void launch_finalize(uint32_t n_junc, const uint32_t* seg_start, const JuncAcc& A, const Genome& G, const double* entropy,
                     pj_junction* rows, uint32_t* err, cudaStream_t st);

} // namespace pjk
