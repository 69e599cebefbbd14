/*
 * junc_oracle.c — TEST INFRASTRUCTURE ONLY (see junc_oracle.h).
 *
 * A deliberately literal, single-threaded C restatement of the reference `junc` path.  It keeps the
 * reference's structure (per-junction read lists, string building for the anchor windows) so that each
 * function can be read side by side with the file:line it cites under /root/reference.  It is NOT the
 * product and is never reachable from portcullis_b200/.
 */
#include "junc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <ctype.h>
#include <limits.h>

static char g_err[1024];
const char* oj_last_error(void) { return g_err; }
void oj_free(void* p) { free(p); }
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)

/* ---- CIGAR helpers: htslib sam.h:75-81 (BAM_CIGAR_STR "MIDNSHP=XB"), CigarOp bam_alignment.hpp:75-99 ---- */
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8, OP_B = 9 };
static int op_type(uint32_t c) { return (int)(c & 0xf); }
static int32_t op_len(uint32_t c) { return (int32_t)(c >> 4); }
static int consumes_query(int t) { return t == OP_M || t == OP_I || t == OP_S || t == OP_EQ || t == OP_X; }
static int consumes_ref(int t) { return t == OP_M || t == OP_D || t == OP_N || t == OP_EQ || t == OP_X; }

static const char NT16[] = "=ACMGRSVTWYHKDBN";   /* htslib hts.c:82 seq_nt16_str */

/* seq_utils.hpp:33-40, index c-'A' */
static const char REVCOMP[26] = { 'T', 0, 'G', 'H', 0, 0, 'C', 'D', 0, 0, 0, 0, 'K', 'N', 0, 0, 0, 'Y', 'W', 'A', 'A', 'B', 'S', 'X', 'R', 0 };

void oj_revcomp(const char* in, int32_t n, char* out) {           /* seq_utils.hpp:111-118 */
    for (int32_t i = 0; i < n; i++) {
        int idx = (int)in[n - 1 - i] - 65;
        out[i] = (idx >= 0 && idx < 26) ? REVCOMP[idx] : 0;
    }
}

int oj_hamming(const char* a, const char* b, int32_t n) {          /* seq_utils.hpp:62-77 */
    int s = 0;
    for (int32_t i = 0; i < n; i++)
        if (toupper((unsigned char)a[i]) != toupper((unsigned char)b[i])) s++;
    return s;
}

int oj_splice_motif(const char* d, const char* a, int32_t* ss_strand) {   /* junction.cc:289-326, junction.hpp:73-79 */
    char m[5] = { d[0], d[1], a[0], a[1], 0 };
    int css = 'N'; int32_t st = PJ_STRAND_UNKNOWN;
    if (!strcmp(m, "GTAG")) { css = 'C'; st = PJ_STRAND_POS; }
    else if (!strcmp(m, "CTAC")) { css = 'C'; st = PJ_STRAND_NEG; }
    else if (!strcmp(m, "ATAC") || !strcmp(m, "GCAG")) { css = 'S'; st = PJ_STRAND_POS; }
    else if (!strcmp(m, "GTAT") || !strcmp(m, "CTGC")) { css = 'S'; st = PJ_STRAND_NEG; }
    if (ss_strand) *ss_strand = st;
    return css;
}

int32_t oj_min_anchor(int32_t start, int32_t end, int32_t left, int32_t right) {   /* intron.cc:67-87 */
    int32_t l = start - left, r = right - end;
    return l < r ? l : r;
}

double oj_entropy(const int32_t* p, int64_t n) {                    /* junction.cc:730-749, quirk Q1 kept */
    if (n <= 1) return 0.0;
    double sum = 0.0;
    int32_t lastOffset = p[0];
    uint32_t readsAtOffset = 0;
    for (int64_t i = 0; i < n; i++) {
        int32_t pos = p[i];
        readsAtOffset++;
        if (pos != lastOffset || i == n - 1) {
            double pI = (double)readsAtOffset / (double)n;
            sum += pI * log2(pI);
            lastOffset = pos;
            readsAtOffset = 0;
        }
    }
    return fabs(sum);
}

/* ---- padded window strings: bam_alignment.cc:256-264, 341-462 ---- */

/* getQuerySeqAfterClipping (bam_alignment.cc:256-264): offset/length of the clipped view into the full read */
static void clipped_view(const uint32_t* cg, int32_t n, int32_t qlen, int32_t* off, int32_t* len) {
    int32_t ds = (n > 0 && op_type(cg[0]) == OP_S) ? op_len(cg[0]) : 0;
    int32_t de = (n > 0 && op_type(cg[n - 1]) == OP_S) ? op_len(cg[n - 1]) : 0;
    if (ds > qlen) ds = qlen;                                   /* std::string::substr would throw; never valid */
    int64_t l = (int64_t)qlen - ds - de + 1;
    if (l > qlen - ds) l = qlen - ds;
    if (l < 0) l = 0;
    *off = ds; *len = (int32_t)l;
}

static int32_t aligned_length(const uint32_t* cg, int32_t n) {     /* bam_alignment.cc:78-88 */
    int32_t a = 0;
    for (int32_t i = 0; i < n; i++) if (consumes_ref(op_type(cg[i]))) a += op_len(cg[i]);
    return a;
}

int oj_padded_query(int32_t pos, const uint32_t* cg, int32_t n, const char* query_full, int32_t query_len,
                    int32_t start, int32_t end, char* out, int32_t* actual_start, int32_t* actual_end) {
    int32_t getEnd = pos + aligned_length(cg, n) - 1;
    if (start > getEnd || end < pos) FAIL(PJ_EDATA, "Found an alignment that does not have a presence in the requested region");
    int32_t qoff, qsize; clipped_view(cg, n, query_len, &qoff, &qsize);
    const char* query = query_full + qoff;
    int32_t qPos = 0, rPos = pos, o = 0;
    for (int32_t k = 0; k < n; k++) {
        int t = op_type(cg[k]); int32_t L = op_len(cg[k]);
        int cr = consumes_ref(t);
        int cq = consumes_query(t) && t != OP_S;                   /* include_soft_clips == false */
        if (rPos < start) { if (cr) rPos += L; if (cq) qPos += L; continue; }
        if ((rPos > end && t != OP_I) || (t == OP_N && rPos + L > end)) break;
        if (cq) {
            int32_t len = (rPos + L > end && t != OP_I) ? end - rPos + 1 : L;
            if (len == 0) FAIL(PJ_EDATA, "Can't extract cigar op sequence from query string when length has been calculated as 0.");
            if (qPos < 0 || qPos + len > qsize) FAIL(PJ_EDATA, "Can't extract cigar op sequence from query string.");
            memcpy(out + o, query + qPos, (size_t)len); o += len;
        }
        else if (cr) {
            int32_t len = rPos + L > end ? end - rPos + 1 : L;
            memset(out + o, 'X', (size_t)len); o += len;
        }
        if (cr) rPos += L;
        if (cq) qPos += L;
    }
    out[o] = 0;
    if (actual_start) *actual_start = pos > start ? pos : start;
    if (actual_end) *actual_end = rPos <= end ? rPos - 1 : end;
    return o;
}

int oj_padded_genome(int32_t pos, const uint32_t* cg, int32_t n, const char* genome, int32_t genome_len,
                     int32_t start, int32_t end, int32_t q_start, int32_t q_end, char* out) {
    int32_t getEnd = pos + aligned_length(cg, n) - 1;
    if (start > getEnd || end < pos) FAIL(PJ_EDATA, "Found an alignment that does not have a presence in the requested region");
    if (q_start - start < 0) FAIL(PJ_EDATA, "Query start position was before genomic region start position.");
    if (end - q_end < 0) FAIL(PJ_EDATA, "Query end position was beyond genomic region end position.");
    int32_t rPos = pos, o = 0;
    for (int32_t k = 0; k < n; k++) {
        int t = op_type(cg[k]); int32_t L = op_len(cg[k]);
        int cr = consumes_ref(t);
        int cq = consumes_query(t) && t != OP_S;
        if (rPos < q_start) { if (cr) rPos += L; continue; }
        if (rPos > q_end && t != OP_I) break;
        if (cr) {
            int32_t seqOffset = rPos - start;
            int32_t len = rPos + L > q_end ? q_end - rPos + 1 : L;
            if (seqOffset < 0 || seqOffset + len > genome_len) FAIL(PJ_EDATA, "Can't extract cigar op sequence from extracted genome region.");
            memcpy(out + o, genome + seqOffset, (size_t)len); o += len;
        }
        else if (cq) {
            memset(out + o, 'X', (size_t)L); o += L;
        }
        if (cr) rPos += L;
    }
    out[o] = 0;
    return o;
}

/* ---- junction store ---- */

typedef struct {
    int32_t tid, start, end, left, right;
    uint32_t maxMinAnchor;
    uint32_t r1p, r1n, r2p, r2n, ms;
    int64_t* reads; int64_t n, cap;         /* record indices, insertion (BAM) order */
} Junc;

typedef struct {
    Junc* j; int64_t n, cap;
    int64_t* table; int64_t tcap;           /* open addressing: index+1, 0 = empty */
} JuncSet;

static uint64_t key_hash(int32_t tid, int32_t s, int32_t e) {
    uint64_t h = (uint64_t)(uint32_t)tid * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)s << 32 | (uint32_t)e) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}

static int64_t set_find(JuncSet* S, int32_t tid, int32_t s, int32_t e, int64_t** slot_out) {
    uint64_t m = (uint64_t)S->tcap - 1, i = key_hash(tid, s, e) & m;
    for (;;) {
        int64_t v = S->table[i];
        if (v == 0) { *slot_out = &S->table[i]; return -1; }
        Junc* q = &S->j[v - 1];
        if (q->tid == tid && q->start == s && q->end == e) return v - 1;
        i = (i + 1) & m;
    }
}

static int set_grow(JuncSet* S) {
    if (S->n + 1 > S->cap) {
        int64_t nc = S->cap ? S->cap * 2 : 1024;
        Junc* nj = (Junc*)realloc(S->j, (size_t)nc * sizeof(Junc)); if (!nj) return -1;
        S->j = nj; S->cap = nc;
    }
    if ((S->n + 1) * 2 > S->tcap) {
        int64_t nt = S->tcap ? S->tcap * 2 : 4096;
        int64_t* t = (int64_t*)calloc((size_t)nt, sizeof(int64_t)); if (!t) return -1;
        free(S->table); S->table = t; S->tcap = nt;
        for (int64_t k = 0; k < S->n; k++) {
            uint64_t m = (uint64_t)nt - 1, i = key_hash(S->j[k].tid, S->j[k].start, S->j[k].end) & m;
            while (t[i]) i = (i + 1) & m;
            t[i] = k + 1;
        }
    }
    return 0;
}

static int junc_push_read(Junc* q, int64_t r) {
    if (q->n + 1 > q->cap) {
        int64_t nc = q->cap ? q->cap * 2 : 4;
        int64_t* nr = (int64_t*)realloc(q->reads, (size_t)nc * sizeof(int64_t)); if (!nr) return -1;
        q->reads = nr; q->cap = nc;
    }
    q->reads[q->n++] = r;
    return 0;
}

/* Junction::addJunctionAlignment (junction.cc:477-502) */
static int add_alignment(Junc* q, const pj_batch* b, int64_t r, int nbJunctionsInRead) {
    if (junc_push_read(q, r)) return -1;
    uint16_t f = b->flag[r];
    if (f & 0x40) { if (!(f & 0x10)) q->r1p++; else q->r1n++; }
    else          { if (!(f & 0x10)) q->r2p++; else q->r2n++; }
    if (nbJunctionsInRead > 1) q->ms++;
    return 0;
}

/* JunctionSystem::addJunctions (junction_system.cc:140-210), recursion kept */
static int add_junctions(JuncSet* S, const pj_batch* b, int64_t r, int32_t refLength, int nbN,
                         int32_t startOp, int32_t offset, int* found) {
    const uint32_t* cg = b->cigar + b->cigar_off[r];
    int32_t nbOps = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
    int32_t tid = b->tid[r];
    int32_t lStart = offset, lEndExc = lStart, rStart, rEndExc;
    for (int32_t i = startOp; i < nbOps; i++) {
        int t = op_type(cg[i]); int32_t L = op_len(cg[i]);
        if (t == OP_N) {
            *found = 1;
            rStart = lEndExc + L;
            rEndExc = rStart;
            int32_t j = i + 1;
            while (j < nbOps && rEndExc <= refLength && op_type(cg[j]) != OP_N) {
                if (consumes_ref(op_type(cg[j]))) rEndExc += op_len(cg[j]);
                j++;
            }
            if (rStart - 1 >= refLength) rStart = refLength - 1;
            if (rEndExc - 1 >= refLength) rEndExc = refLength;
            int32_t is = lEndExc, ie = rStart - 1;
            if (set_grow(S)) return PJ_ENOMEM;
            int64_t* slot = NULL; int64_t idx = set_find(S, tid, is, ie, &slot);
            if (idx < 0) {
                /* Junction ctor (junction.cc:328-387): maxMinAnchor = minAnchorLength(lStart, rEndExc-1) which throws
                 * when the anchors do not enclose the intron (intron.cc:68-85) */
                if (lStart > is || rEndExc - 1 < ie) FAIL(PJ_EDATA, "intron not enclosed by its anchors (record %lld)", (long long)r);
                Junc* q = &S->j[S->n]; memset(q, 0, sizeof *q);
                q->tid = tid; q->start = is; q->end = ie; q->left = lStart; q->right = rEndExc - 1;
                q->maxMinAnchor = (uint32_t)oj_min_anchor(is, ie, lStart, rEndExc - 1);
                if (add_alignment(q, b, r, nbN)) return PJ_ENOMEM;
                *slot = ++S->n;
            }
            else {
                Junc* q = &S->j[idx];
                if (add_alignment(q, b, r, nbN)) return PJ_ENOMEM;
                /* extendAnchors (junction.cc:524-529) */
                if (lStart > is || rEndExc - 1 < ie) FAIL(PJ_EDATA, "intron not enclosed by its anchors (record %lld)", (long long)r);
                if (lStart < q->left) q->left = lStart;
                if (rEndExc - 1 > q->right) q->right = rEndExc - 1;
                uint32_t o = (uint32_t)oj_min_anchor(is, ie, lStart, rEndExc - 1);
                if (o > q->maxMinAnchor) q->maxMinAnchor = o;
            }
            if (j < nbOps) {
                int dummy = 0;
                int rc = add_junctions(S, b, r, refLength, nbN, i + 1, rStart, &dummy);
                if (rc) return rc;
                break;
            }
        }
        else if (consumes_ref(t)) lEndExc += L;
    }
    return 0;
}

/* faidx_fetch_seq clamping (htslib faidx.c:455-459) + to_upper (junction.cc:586-587, 635-638).
 * Returns malloc'd upper-cased string, length in *len. */
static char* fetch_bases(const char* g, int64_t glen, int32_t beg, int32_t end, int32_t* len) {
    int64_t b = beg, e = end;
    if (glen <= 0) { *len = 0; char* s = (char*)malloc(1); if (s) s[0] = 0; return s; }
    if (e < b) b = e;
    if (b < 0) b = 0; else if (glen <= b) b = glen - 1;
    if (e < 0) e = 0; else if (glen <= e) e = glen - 1;
    int64_t l = e - b + 1;
    char* s = (char*)malloc((size_t)l + 1); if (!s) { *len = -1; return NULL; }
    for (int64_t k = 0; k < l; k++) s[k] = (char)toupper((unsigned char)g[b + k]);
    s[l] = 0; *len = (int32_t)l;
    return s;
}

/* BamAlignment::calcIfProperPair (bam_alignment.cc:271-292) */
static int calc_if_proper_pair(const pj_batch* b, int64_t r, int orientation) {
    uint16_t f = b->flag[r];
    if (!(f & 0x1) || (f & 0x8)) return 0;
    if (b->tid[r] != b->mtid[r]) return 0;
    int rev = (f & 0x10) != 0, mrev = (f & 0x20) != 0;
    int diffStrand = rev != mrev;
    int posGap = !rev ? b->pos[r] < b->mpos[r] : b->pos[r] > b->mpos[r];
    if (orientation == PJ_ORIENT_FR) return diffStrand && posGap;
    if (orientation == PJ_ORIENT_RF) return diffStrand && !posGap;
    if (orientation == PJ_ORIENT_FF) return !diffStrand && posGap;
    return 0;
}

static int strand_of_xs(uint8_t xs) {      /* bam_alignment.cc:93-99, 226-231; calcStrand() is UNKNOWN inside junc (A4) */
    if (xs == '+') return PJ_STRAND_POS;
    if (xs == '-') return PJ_STRAND_NEG;
    return PJ_STRAND_UNKNOWN;
}

static uint32_t matches_from_start(const char* q, const char* a, int32_t n) {   /* junction.cc:263-270 */
    for (int32_t i = 0; i < n; i++) if (q[i] != a[i]) return (uint32_t)i;
    return (uint32_t)n;
}
static uint32_t matches_from_end(const char* q, const char* a, int32_t n) {     /* junction.cc:272-280 */
    for (int32_t j = n; j > 0; j--) { int32_t i = j - 1; if (q[i] != a[i]) return (uint32_t)(n - i - 1); }
    return (uint32_t)n;
}

typedef struct { uint32_t mmes, minMatch, nbMismatches; } MatchStats;

/* AlignmentInfo::calcMatchStats (junction.cc:147-240) */
static int calc_match_stats(const pj_batch* b, int64_t r, const Junc* q, const char* ancLeft, int32_t ancLeftLen,
                            const char* ancRight, int32_t ancRightLen, MatchStats* ms) {
    int32_t leftStart = q->left, rightEnd = q->right;
    int32_t leftEnd = q->start - 1, rightStart = q->end + 1;
    int32_t lq = b->l_qseq[r];
    memset(ms, 0, sizeof *ms);
    if (lq <= 1) {                                               /* junction.cc:168-185 */
        uint32_t totalUpMatches = (uint32_t)(leftEnd - leftStart + 1);
        uint32_t totalDownMatches = (uint32_t)(rightEnd - rightStart + 1);
        ms->nbMismatches = 0; ms->minMatch = 0;
        ms->mmes = totalUpMatches < totalDownMatches ? totalUpMatches : totalDownMatches;
        return 0;
    }
    const uint32_t* cg = b->cigar + b->cigar_off[r];
    int32_t n = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
    if ((int64_t)(b->seq_off[r + 1] - b->seq_off[r]) < (lq + 1) / 2) FAIL(PJ_EINVAL, "record %lld: SEQ missing for a spliced read", (long long)r);
    const uint8_t* s4 = b->seq4 + b->seq_off[r];
    char* query = (char*)malloc((size_t)lq + 1);                  /* getQuerySeq, bam_alignment.cc:244-250 */
    int64_t span = (int64_t)rightEnd - leftStart + 16 + lq;
    char* qL = (char*)malloc((size_t)span); char* gL = (char*)malloc((size_t)span);
    char* qR = (char*)malloc((size_t)span); char* gR = (char*)malloc((size_t)span);
    if (!query || !qL || !gL || !qR || !gR) { free(query); free(qL); free(gL); free(qR); free(gR); FAIL(PJ_ENOMEM, "oom"); }
    for (int32_t i = 0; i < lq; i++) query[i] = NT16[(s4[i >> 1] >> ((~i & 1) << 2)) & 0xf];
    query[lq] = 0;
    int rc = 0;
    int32_t qLS, qLE, qRS, qRE;
    int nqL = oj_padded_query(b->pos[r], cg, n, query, lq, leftStart, leftEnd, qL, &qLS, &qLE);
    int nqR = nqL < 0 ? -1 : oj_padded_query(b->pos[r], cg, n, query, lq, rightStart, rightEnd, qR, &qRS, &qRE);
    int ngL = nqR < 0 ? -1 : oj_padded_genome(b->pos[r], cg, n, ancLeft, ancLeftLen, leftStart, leftEnd, qLS, qLE, gL);
    int ngR = ngL < 0 ? -1 : oj_padded_genome(b->pos[r], cg, n, ancRight, ancRightLen, rightStart, rightEnd, qRS, qRE, gR);
    if (nqL < 0 || nqR < 0 || ngL < 0 || ngR < 0) rc = PJ_EDATA;
    else if (nqL != ngL || nqL == 0 || nqR != ngR || nqR == 0) {
        /* junction.cc:192-223: the reference prints a warning, leaves the stats at 0 and then reads an empty
         * vector<bool> in calcMismatchStats (:881) — undefined behaviour.  Not valid input. */
        snprintf(g_err, sizeof g_err, "record %lld: anchor region for query and genome are not the same size (reference behaviour undefined)", (long long)r);
        rc = PJ_EDATA;
    }
    else {
        uint32_t upMism = (uint32_t)oj_hamming(qL, gL, nqL), downMism = (uint32_t)oj_hamming(qR, gR, nqR);
        uint32_t upMatches = (uint32_t)nqL - upMism, downMatches = (uint32_t)nqR - downMism;
        ms->nbMismatches = upMism + downMism;
        uint32_t us = matches_from_end(qL, gL, nqL), ds = matches_from_start(qR, gR, nqR);
        ms->minMatch = us < ds ? us : ds;
        ms->mmes = upMatches < downMatches ? upMatches : downMatches;
    }
    free(query); free(qL); free(gL); free(qR); free(gR);
    return rc;
}

static int cmp_i32(const void* a, const void* b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x > y) - (x < y); }

/* Junction::calcMetrics (junction.cc:683-687) + processJunctionWindow (:561-649) for one junction */
static int finish_junction(const pj_batch* b, const Junc* q, const char* g, int64_t glen, int orientation, pj_junction* o) {
    memset(o, 0, sizeof *o);
    o->tid = q->tid; o->start = q->start; o->end = q->end; o->left = q->left; o->right = q->right;
    o->nb_raw_aln = (uint32_t)q->n; o->nb_ms_aln = q->ms;
    o->nb_r1_pos = q->r1p; o->nb_r1_neg = q->r1n; o->nb_r2_pos = q->r2p; o->nb_r2_neg = q->r2n;
    o->max_min_anc = q->maxMinAnchor;
    o->hamming5p = 10; o->hamming3p = 10;

    /* determineStrandFromReads (junction.cc:531-559) */
    uint32_t nb_pos = 0, nb_neg = 0, nb_unk = 0;
    for (int64_t k = 0; k < q->n; k++) {
        int s = strand_of_xs(b->xs[q->reads[k]]);
        if (s == PJ_STRAND_POS) nb_pos++; else if (s == PJ_STRAND_NEG) nb_neg++; else nb_unk++;
    }
    uint32_t total = nb_pos + nb_neg + nb_unk;
    o->nb_xs_pos = nb_pos; o->nb_xs_neg = nb_neg;
    if ((double)nb_pos / (double)total >= 0.95) o->read_strand = PJ_STRAND_POS;
    else if ((double)nb_neg / (double)total >= 0.95) o->read_strand = PJ_STRAND_NEG;
    else o->read_strand = PJ_STRAND_UNKNOWN;

    /* calcEntropy (junction.cc:718-728) */
    int32_t* ps = (int32_t*)malloc((size_t)(q->n ? q->n : 1) * sizeof(int32_t)); if (!ps) FAIL(PJ_ENOMEM, "oom");
    for (int64_t k = 0; k < q->n; k++) ps[k] = b->pos[q->reads[k]];
    qsort(ps, (size_t)q->n, sizeof(int32_t), cmp_i32);
    o->entropy = oj_entropy(ps, q->n);
    free(ps);

    /* calcAlignmentStats (junction.cc:755-814) */
    int32_t lastStart = -1, lastEnd = -1;
    int ppcheck = orientation == PJ_ORIENT_FR || orientation == PJ_ORIENT_FF || orientation == PJ_ORIENT_RF;
    for (int64_t k = 0; k < q->n; k++) {
        int64_t r = q->reads[k];
        const uint32_t* cg = b->cigar + b->cigar_off[r];
        int32_t n = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
        int32_t start = b->pos[r], end = start + aligned_length(cg, n) - 1;
        if (start != lastStart || end != lastEnd) { o->nb_dist_aln++; lastStart = start; lastEnd = end; }
        int reliable = 1;
        if (b->mapq[r] >= PJ_MAP_QUALITY_THRESHOLD) o->nb_um_aln++; else reliable = 0;
        if (b->flag[r] & 0x2) o->nb_bpp_aln++;
        if (ppcheck) { if (calc_if_proper_pair(b, r, orientation)) o->nb_ppp_aln++; else reliable = 0; }
        if (reliable) o->nb_rel_aln++;
        uint32_t up = 0, down = 0; int32_t pos = start;
        for (int32_t c = 0; c < n; c++) {
            int t = op_type(cg[c]);
            if (consumes_ref(t)) pos += op_len(cg[c]);
            if (t == OP_N) { if (pos < q->start) up++; else if (pos > q->end + 1) down++; }
        }
        if (up > o->nb_up_juncs) o->nb_up_juncs = up;
        if (down > o->nb_down_juncs) o->nb_down_juncs = down;
    }

    /* processJunctionWindow (junction.cc:561-649) */
    int32_t dl, al, lal, ral, lil, ril;
    char* donor = fetch_bases(g, glen, q->start, q->start + 1, &dl);
    char* acceptor = fetch_bases(g, glen, q->end - 1, q->end, &al);
    char* leftAnc = fetch_bases(g, glen, q->left, q->start - 1, &lal);
    char* rightAnc = fetch_bases(g, glen, q->end + 1, q->right, &ral);
    char* leftInt = fetch_bases(g, glen, q->start, q->start + 9, &lil);
    char* rightInt = fetch_bases(g, glen, q->end - 9, q->end, &ril);
    int rc = 0;
    if (!donor || !acceptor || !leftAnc || !rightAnc || !leftInt || !rightInt) { snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; }
    else if (dl != 2 || al != 2) { snprintf(g_err, sizeof g_err, "splice site sequence of junction %d:%d-%d is not the expected length", q->tid, q->start, q->end); rc = PJ_EDATA; }
    else if ((lal != q->start - q->left && q->start - q->left > 0) || (ral != q->right - q->end && q->right - q->end > 0)) {
        snprintf(g_err, sizeof g_err, "anchor sequence of junction %d:%d-%d is not the expected length", q->tid, q->start, q->end); rc = PJ_EDATA; }
    else if (lil != 10 || ril != 10) { snprintf(g_err, sizeof g_err, "intron region of junction %d:%d-%d is not the expected length", q->tid, q->start, q->end); rc = PJ_EDATA; }
    if (!rc) {
        /* setDonorAndAcceptorMotif (junction.cc:504-516) */
        int32_t ss; o->canonical_ss = (uint8_t)oj_splice_motif(donor, acceptor, &ss);
        o->ss_strand = (uint8_t)ss;
        o->consensus_strand = o->read_strand == o->ss_strand ? o->read_strand :
                              o->read_strand == PJ_STRAND_UNKNOWN ? o->ss_strand :
                              o->ss_strand == PJ_STRAND_UNKNOWN ? o->read_strand : PJ_STRAND_UNKNOWN;
        if (o->consensus_strand == PJ_STRAND_NEG) { oj_revcomp(acceptor, 2, o->ss1); oj_revcomp(donor, 2, o->ss2); }
        else { memcpy(o->ss1, donor, 2); memcpy(o->ss2, acceptor, 2); }

        /* junction.cc:639-641 then calcHammingScores (:823-857) */
        const char* la10 = lal < 10 ? leftAnc : leftAnc + (lal - 10); int32_t la10n = lal < 10 ? lal : 10;
        const char* ra10 = rightAnc;                                   int32_t ra10n = ral < 10 ? ral : 10;
        int32_t leftDelta = la10n - ril, leftOffset = leftDelta <= 0 ? 0 : leftDelta;
        int32_t leftLen = la10n < ril ? la10n : ril, rightLen = lil < ra10n ? lil : ra10n;
        char la[16], li[16], ri[16], ra[16], t1[16], t2[16];
        int32_t lan, lin, rin, ran;
        if (la10n > leftLen) { memcpy(la, la10 + leftOffset, (size_t)leftLen); lan = leftLen; } else { memcpy(la, la10, (size_t)la10n); lan = la10n; }
        if (lil > rightLen)  { memcpy(li, leftInt, (size_t)rightLen); lin = rightLen; } else { memcpy(li, leftInt, (size_t)lil); lin = lil; }
        if (ril > leftLen)   { memcpy(ri, rightInt + leftOffset, (size_t)leftLen); rin = leftLen; } else { memcpy(ri, rightInt, (size_t)ril); rin = ril; }
        if (ra10n > rightLen){ memcpy(ra, ra10, (size_t)rightLen); ran = rightLen; } else { memcpy(ra, ra10, (size_t)ra10n); ran = ra10n; }
        if (lan != rin || ran != lin) { snprintf(g_err, sizeof g_err, "hamming strings differ in length"); rc = PJ_EDATA; }
        else if (o->consensus_strand == PJ_STRAND_NEG) {
            oj_revcomp(ra, ran, t1); oj_revcomp(li, lin, t2); o->hamming5p = (uint32_t)oj_hamming(t1, t2, ran);   /* anchor5p=rc(ra) vs intron3p=rc(li) */
            oj_revcomp(la, lan, t1); oj_revcomp(ri, rin, t2); o->hamming3p = (uint32_t)oj_hamming(t1, t2, lan);   /* anchor3p=rc(la) vs intron5p=rc(ri) */
        }
        else {
            o->hamming5p = (uint32_t)oj_hamming(la, ri, lan);
            o->hamming3p = (uint32_t)oj_hamming(ra, li, ran);
        }
    }
    if (!rc) {
        /* per-read match stats then calcMismatchStats (junction.cc:862-909) */
        uint32_t nbMismatches = 0, firstMismatch = 100000000u, maxMinMatch = 0;
        for (int64_t k = 0; k < q->n && !rc; k++) {
            MatchStats ms;
            rc = calc_match_stats(b, q->reads[k], q, leftAnc, lal, rightAnc, ral, &ms);
            if (rc) break;
            if (ms.mmes > o->maxmmes) o->maxmmes = ms.mmes;
            nbMismatches += ms.nbMismatches;
            if (ms.minMatch > 0 && ms.minMatch < firstMismatch) firstMismatch = ms.minMatch;
            for (uint32_t i = 0; i < PJ_NB_JAD && i < ms.minMatch; i++) o->jad[i]++;
            if (ms.minMatch > maxMinMatch) maxMinMatch = ms.minMatch;
        }
        o->nb_mismatches = nbMismatches;
        if (nbMismatches > 0 && firstMismatch < 20 && !(maxMinMatch > firstMismatch)) o->suspicious = 1;
    }
    free(donor); free(acceptor); free(leftAnc); free(rightAnc); free(leftInt); free(rightInt);
    return rc;
}

static int cmp_rows(const void* a, const void* b) {            /* IntronComparator, intron.cc:112-127 */
    const pj_junction* x = (const pj_junction*)a; const pj_junction* y = (const pj_junction*)b;
    if (x->tid != y->tid) return x->tid < y->tid ? -1 : 1;
    if (x->start != y->start) return x->start < y->start ? -1 : 1;
    if (x->end != y->end) return x->end < y->end ? -1 : 1;
    return 0;
}

/* findJuncs (junction_builder.cc:314-357).  Record visibility (Q13, htslib hts.c:1951-1953): tid==target,
 * pos < target_len, endpos > 0.  The flush at :324-331 only bounds memory; results equal a batch pass. */
static int build_junction_set(const pj_batch* b, int32_t n_targets, const int32_t* target_len, JuncSet* S, pj_target_stats* stats) {
    int rc = 0;
    for (int64_t r = 0; r < b->n_records && !rc; r++) {
        int32_t tid = b->tid[r];
        if (tid < 0 || tid >= n_targets) continue;
        const uint32_t* cg = b->cigar + b->cigar_off[r];
        int32_t n = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
        int32_t pos = b->pos[r];
        int32_t rlen = aligned_length(cg, n);
        int64_t endpos = (!(b->flag[r] & 0x4) && n > 0) ? (int64_t)pos + rlen : (int64_t)pos + 1;   /* sam.c:336-342 bam_endpos */
        if (!(pos < target_len[tid] && endpos > 0)) continue;
        int32_t len = b->l_qseq[r];
        if (stats) {
            pj_target_stats* st = &stats[tid];
            if (len < st->min_query_length) st->min_query_length = len;
            if (len > st->max_query_length) st->max_query_length = len;
            st->sum_query_lengths += (uint64_t)(int64_t)len;
        }
        int nbN = 0;
        for (int32_t c = 0; c < n; c++) if (op_type(cg[c]) == OP_N) nbN++;
        int found = 0;
        rc = add_junctions(S, b, r, target_len[tid], nbN, 0, pos, &found);
        if (stats) { if (found) stats[tid].spliced_count++; else stats[tid].unspliced_count++; }
    }
    return rc;
}

int oj_run(const pj_batch* b, int32_t n_targets, const int32_t* target_len,
           const char* genome_cat, const int64_t* genome_off, int32_t orientation,
           pj_junction** rows_out, int64_t* n_rows_out, pj_target_stats* stats) {
    JuncSet S; memset(&S, 0, sizeof S);
    for (int32_t t = 0; t < n_targets; t++) {
        memset(&stats[t], 0, sizeof stats[t]);
        stats[t].min_query_length = INT32_MAX;
    }
    int rc = build_junction_set(b, n_targets, target_len, &S, stats);
    pj_junction* rows = NULL;
    if (!rc) {
        rows = (pj_junction*)calloc((size_t)(S.n ? S.n : 1), sizeof(pj_junction));
        if (!rows) { snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; }
    }
    for (int64_t k = 0; k < S.n && !rc; k++) {
        int32_t t = S.j[k].tid;
        rc = finish_junction(b, &S.j[k], genome_cat + genome_off[t], genome_off[t + 1] - genome_off[t], orientation, &rows[k]);
    }
    if (!rc) qsort(rows, (size_t)S.n, sizeof(pj_junction), cmp_rows);
    for (int64_t k = 0; k < S.n; k++) free(S.j[k].reads);
    int64_t n = S.n;
    free(S.j); free(S.table);
    if (rc) { free(rows); return rc; }
    *rows_out = rows; *n_rows_out = n;
    return 0;
}

/* junction_builder.cc:258-290 + junction_system.cc:55-70, 250-330 */
int oj_finalize(pj_junction* rows, int64_t n, double meanQueryLength) {
    qsort(rows, (size_t)n, sizeof(pj_junction), cmp_rows);
    for (int64_t i = 0; i < n; i++) {
        pj_junction* j = &rows[i];
        j->index = (uint32_t)i;
        j->rel2raw = (double)j->nb_rel_aln / (double)j->nb_raw_aln;                 /* junction.hpp:551-553 */
        j->mean_mismatches = (double)j->nb_mismatches / (double)j->nb_raw_aln;     /* junction.cc:893 */
        j->uniq_junc = 0; j->primary_junc = 0; j->pfp = 0; j->mean_readlen = 0;
        j->dist_2_up_junc = 0; j->dist_2_down_junc = 0; j->dist_nearest_junc = 0;
    }
    if (n <= 1) return 0;                                                            /* junction_builder.cc:285 */
    for (int64_t i = 0; i < n; i++) {
        /* createJunctionGroup (junction_system.cc:55-70) */
        int64_t first = i, last = i;
        for (int64_t j = i + 1; j < n; j++) {
            const pj_junction* a = &rows[last]; const pj_junction* c = &rows[j];
            if (a->tid == c->tid && (a->start == c->start || a->end == c->end)) last = j; else break;
        }
        uint32_t maxReads = 0; int64_t maxIndex = first;
        int uniq = (last == first);
        for (int64_t j = first; j <= last; j++) {
            if (maxReads < rows[j].nb_raw_aln) { maxReads = rows[j].nb_raw_aln; maxIndex = j; }
            rows[j].uniq_junc = (uint8_t)uniq;
        }
        rows[maxIndex].primary_junc = 1;
        i = last;
    }
    int64_t i = 0; int lastdiffseq = 0;
    while (i < n - 1) {
        pj_junction* first = &rows[i]; pj_junction* second = &rows[i + 1];
        int32_t diff = second->start - first->end; if (diff < 0) diff = 0;
        if (first->tid != second->tid) {
            first->dist_2_up_junc = (uint32_t)-1; second->dist_2_down_junc = (uint32_t)-1;
            if (i == 0 || lastdiffseq) first->dist_2_down_junc = (uint32_t)-1;
            if (i == n - 2) second->dist_2_up_junc = (uint32_t)-1;
            lastdiffseq = 1;
        }
        else if (i == 0) { first->dist_2_down_junc = (uint32_t)-1; first->dist_2_up_junc = (uint32_t)diff; second->dist_2_down_junc = (uint32_t)diff; lastdiffseq = 0; }
        else if (i == n - 2) { first->dist_2_up_junc = (uint32_t)diff; second->dist_2_down_junc = (uint32_t)diff; second->dist_2_up_junc = (uint32_t)-1; lastdiffseq = 0; }
        else { first->dist_2_up_junc = (uint32_t)diff; second->dist_2_down_junc = (uint32_t)diff; lastdiffseq = 0; }
        i++;
    }
    for (int64_t k = 0; k < n; k++) {
        pj_junction* j = &rows[k];
        int32_t down = (int32_t)j->dist_2_down_junc, up = (int32_t)j->dist_2_up_junc;
        int32_t nearest = (down == -1 || up == -1) ? (down > up ? down : up) : (down < up ? down : up);
        j->dist_nearest_junc = (uint32_t)nearest;
        j->mean_readlen = (double)(uint32_t)meanQueryLength;                        /* junction.hpp:928-930 */
        if (j->suspicious) {
            double prob = 1.0 - pow((double)j->maxmmes / (meanQueryLength / 2.0), (double)j->nb_raw_aln);
            if (prob > 0.99) j->pfp = 1;
        }
    }
    return 0;
}


/* ================================================================================================================
 * `--extra` metrics: JunctionBuilder::separateBams + calcExtraMetrics restated on the
 * columnar batch (junction_builder.cc:152-226, 293-312).  The batch plays the role of the whole sorted BAM.
 * ================================================================================================================ */

static int is_spliced(const uint32_t* cg, int32_t n) {                /* bam_alignment.cc:294-301 */
    for (int32_t i = 0; i < n; i++) if (op_type(cg[i]) == OP_N) return 1;
    return 0;
}

/* SplicedAlignmentMap = unordered_map<size_t, uint16_t> (junction.hpp:38) */
typedef struct { uint64_t* key; uint16_t* val; uint8_t* used; uint64_t cap; } NameMap;
static uint64_t mix64(uint64_t h) { h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33; return h; }
static uint16_t* namemap_slot(NameMap* m, uint64_t k) {
    uint64_t i = mix64(k) & (m->cap - 1);
    while (m->used[i] && m->key[i] != k) i = (i + 1) & (m->cap - 1);
    if (!m->used[i]) { m->used[i] = 1; m->key[i] = k; m->val[i] = 0; }
    return &m->val[i];
}

typedef struct { int64_t* v; int64_t n, cap; } Heap;                  /* min-heap of end positions */
static int heap_push(Heap* h, int64_t x) {
    if (h->n == h->cap) { int64_t nc = h->cap ? h->cap * 2 : 1024; int64_t* nv = (int64_t*)realloc(h->v, (size_t)nc * 8); if (!nv) return -1; h->v = nv; h->cap = nc; }
    int64_t i = h->n++;
    while (i > 0 && h->v[(i - 1) / 2] > x) { h->v[i] = h->v[(i - 1) / 2]; i = (i - 1) / 2; }
    h->v[i] = x;
    return 0;
}
static void heap_pop(Heap* h) {
    int64_t x = h->v[--h->n], i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && h->v[c + 1] < h->v[c]) c++;
        if (h->v[c] >= x) break;
        h->v[i] = h->v[c]; i = c;
    }
    if (h->n) h->v[i] = x;
}

#ifndef PLP_MAXCNT
#define PLP_MAXCNT 8000                                               /* bam_plp_init, sam.c:1651 */
#endif

/* DepthParser::loadNextBatch over one target of the unspliced BAM (depth_parser.cc:112-167): bam_plp_auto /
 * bam_plp_push / resolve_cigar2 of htslib-1.3 sam.c:1515-1590, 1838-1936.  u[]: the target's unspliced mapped
 * records in file order.  depth[] has target_len entries and is indexed with the reference's +1 shift
 * (`rpos = pos + 1`, depth_parser.cc:155).  Returns 1 when the pileup reports at least one column. */
static int pileup_target(const pj_batch* b, const int64_t* u, int64_t nu, uint32_t* depth, int64_t size, int64_t* n_capped, Heap* live) {
    int covered = 0;
    live->n = 0;
    int64_t last_pos = -1;
    for (int64_t k = 0; k < nu; k++) {
        int64_t r = u[k];
        const uint32_t* cg = b->cigar + b->cigar_off[r];
        int32_t n = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
        int64_t P = b->pos[r];
        int64_t end = n > 0 ? P + aligned_length(cg, n) : P + 1;      /* bam_endpos; FUNMAP is clear here */
        int first_here = (P != last_pos);
        last_pos = P;
        /* bam_plp_next frees nodes with end <= every column it has visited; before a second read at P is pushed
         * the iterator stands on P, having visited up to P-1 */
        while (live->n && live->v[0] <= P - 1) heap_pop(live);
        /* mempool count = live nodes + the spare tail + the dummy node (bam_plp_init, sam.c:1619-1620) */
        if (!first_here && 2 + live->n > PLP_MAXCNT) { (*n_capped)++; continue; }      /* sam.c:1906-1910 */
        int64_t iter_pos = first_here ? P - 1 : P;
        if (end > iter_pos) { if (heap_push(live, end)) return -1; }                   /* sam.c:1930-1933 */
        if (end <= P) continue;                                                        /* never piled up */
        covered = 1;
        int64_t x = P;
        for (int32_t c = 0; c < n; c++) {
            int t = op_type(cg[c]); int64_t L = op_len(cg[c]);
            if (t == OP_M || t == OP_EQ || t == OP_X) {
                for (int64_t q = x; q < x + L; q++) if (q + 1 >= 0 && q + 1 < size) depth[q + 1]++;
                x += L;
            }
            else if (t == OP_D || t == OP_N) x += L;                   /* is_del / is_refskip: not counted (:121-124) */
        }
    }
    return covered;
}

/* Junction::calcCoverage(a, b, levels) (junction.cc:923-933) */
static double coverage_window(int32_t a, int32_t b, const uint32_t* lev, int64_t size, uint32_t* sum) {
    double multiplier = 1.0 / (b - a);
    uint32_t readCount = 0;
    for (int32_t i = a; i <= b; i++) if (i >= 0 && i < size) readCount += lev[i];
    *sum = readCount;
    return multiplier * (double)readCount;
}

static int64_t lower_row(const pj_junction* rows, int64_t n, int32_t tid) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (rows[m].tid < tid) lo = m + 1; else hi = m; }
    return lo;
}

int oj_extra(const pj_batch* b, int32_t n_targets, const int32_t* target_len, const pj_junction* rows, int64_t n_rows,
             int32_t max_query_length, pj_junction_extra* out, int64_t* n_capped_reads) {
    int rc = 0;
    memset(out, 0, (size_t)n_rows * sizeof *out);
    *n_capped_reads = 0;
    if (!b->name_code && b->n_records) FAIL(PJ_EINVAL, "oj_extra needs name_code");

    /* ---- separateBams (junction_builder.cc:176-198): name map over every spliced record of the file ---- */
    NameMap nm; memset(&nm, 0, sizeof nm);
    nm.cap = 1024; while (nm.cap < (uint64_t)b->n_records * 2 + 2) nm.cap <<= 1;
    nm.key = (uint64_t*)malloc(nm.cap * 8); nm.val = (uint16_t*)malloc(nm.cap * 2); nm.used = (uint8_t*)calloc(nm.cap, 1);
    int64_t* uns = (int64_t*)malloc((size_t)(b->n_records + 1) * 8);            /* unspliced mapped, file order */
    int64_t* uoff = (int64_t*)calloc((size_t)n_targets + 1, 8);
    JuncSet S; memset(&S, 0, sizeof S);
    int64_t* pmax = NULL; uint32_t* depth = NULL; Heap live; memset(&live, 0, sizeof live);
    if (!nm.key || !nm.val || !nm.used || !uns || !uoff) { snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; goto done; }
    int64_t nuns = 0;
    {
        int32_t cur = 0;
        for (int64_t r = 0; r < b->n_records; r++) {
            const uint32_t* cg = b->cigar + b->cigar_off[r];
            int32_t n = (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]);
            if (is_spliced(cg, n)) { (*namemap_slot(&nm, b->name_code[r]))++; continue; }
            if (b->flag[r] & 0x4) continue;                                        /* unmapped.bam */
            int32_t tid = b->tid[r];
            if (tid < 0 || tid >= n_targets) continue;                            /* no region query can reach it */
            if (tid < cur) { snprintf(g_err, sizeof g_err, "batch not sorted by tid"); rc = PJ_EINVAL; goto done; }
            if (n == 0) { snprintf(g_err, sizeof g_err, "mapped record %lld has no CIGAR (htslib pileup asserts, sam.c:1537)", (long long)r); rc = PJ_EDATA; goto done; }
            while (cur < tid) uoff[++cur] = nuns;
            uns[nuns++] = r;
        }
        while (cur < n_targets) uoff[++cur] = nuns;
    }

    /* ---- calcMultipleMappingScore (junction.cc:914-921) ---- */
    rc = build_junction_set(b, n_targets, target_len, &S, NULL);
    if (rc) goto done;
    for (int64_t k = 0; k < S.n; k++) {
        const Junc* q = &S.j[k];
        pj_junction key; key.tid = q->tid; key.start = q->start; key.end = q->end;
        const pj_junction* hit = (const pj_junction*)bsearch(&key, rows, (size_t)n_rows, sizeof *rows, cmp_rows);
        if (!hit) { snprintf(g_err, sizeof g_err, "rows do not belong to this batch"); rc = PJ_EINVAL; goto done; }
        pj_junction_extra* x = &out[hit - rows];
        uint32_t M = 0;
        for (int64_t i = 0; i < q->n; i++) M += *namemap_slot(&nm, b->name_code[q->reads[i]]);
        x->mm_n = (uint32_t)q->n; x->mm_m = M;
        x->mm_score = (double)(size_t)q->n / (double)M;
    }

    /* ---- findFlankingAlignments -> processJunctionVicinity (junction_system.cc:212-229, junction.cc:651-677) ---- */
    pmax = (int64_t*)malloc((size_t)(nuns + 1) * 8);
    if (!pmax) { snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; goto done; }
    for (int32_t t = 0; t < n_targets; t++) {
        int64_t m = INT64_MIN;
        for (int64_t k = uoff[t]; k < uoff[t + 1]; k++) {
            int64_t r = uns[k];
            int64_t e = (int64_t)b->pos[r] + aligned_length(b->cigar + b->cigar_off[r], (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]));
            if (e > m) m = e;
            pmax[k] = m;
        }
    }
    for (int64_t j = 0; j < n_rows; j++) {
        const pj_junction* q = &rows[j];
        if (q->tid < 0 || q->tid >= n_targets) continue;
        int32_t refLength = target_len[q->tid];
        int32_t regionStart = q->left - max_query_length - 1;
        regionStart = regionStart < 0 ? 0 : regionStart;
        int32_t regionEnd = q->right + max_query_length + 1;
        regionEnd = regionEnd >= refLength ? refLength - 1 : regionEnd;
        uint32_t nl = 0, nr = 0;
        if (regionEnd > regionStart) {                                             /* hts_itr_query: end < beg -> no iterator; reg2bins: beg >= end -> none */
            int64_t lo = uoff[q->tid], hi = uoff[q->tid + 1];
            int64_t a = lo, z = hi;                                                /* first record whose running max end exceeds regionStart */
            while (a < z) { int64_t m = (a + z) / 2; if (pmax[m] > regionStart) z = m; else a = m + 1; }
            for (int64_t k = a; k < hi; k++) {
                int64_t r = uns[k];
                int32_t pos = b->pos[r];
                if (pos >= regionEnd) break;                                       /* hts_itr_next: beg >= iter->end -> finished */
                int32_t alen = aligned_length(b->cigar + b->cigar_off[r], (int32_t)(b->cigar_off[r + 1] - b->cigar_off[r]));
                int64_t endpos = (int64_t)pos + alen;                              /* bam_endpos (mapped, n_cigar > 0) */
                if (!(endpos > regionStart)) continue;
                int32_t getEnd = pos + alen - 1;                                   /* bam_alignment.hpp:221-223 */
                if (q->start > pos && q->left <= getEnd) nl++;
                if (q->right >= pos && q->end < pos) nr++;
            }
        }
        out[j].up_aln = nl; out[j].down_aln = nr;
    }

    /* ---- calcCoverage (junction_system.cc:231-243) with DepthParser's batch/target pairing (Q14) ---- */
    {
        int32_t prev_cov = -1;           /* previous covered target, whose batch has been built but is applied to the next one */
        int32_t maxlen = 0;
        for (int32_t t = 0; t < n_targets; t++) if (target_len[t] > maxlen) maxlen = target_len[t];
        depth = (uint32_t*)malloc(((size_t)maxlen + 1) * 4);
        uint32_t* depth_prev = (uint32_t*)malloc(((size_t)maxlen + 1) * 4);
        if (!depth || !depth_prev) { free(depth_prev); snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; goto done; }
        int32_t last_cov = -1;
        for (int32_t t = 0; t < n_targets; t++) {
            if (uoff[t + 1] == uoff[t]) continue;
            memset(depth, 0, ((size_t)target_len[t] + 1) * 4);
            int c = pileup_target(b, uns + uoff[t], uoff[t + 1] - uoff[t], depth, target_len[t], n_capped_reads, &live);
            if (c < 0) { free(depth_prev); snprintf(g_err, sizeof g_err, "oom"); rc = PJ_ENOMEM; goto done; }
            if (!c) continue;
            /* the batch of prev_cov is handed to the junctions of t (getCurrentRefIndex() == last.ref == t) */
            if (prev_cov >= 0) {
                for (int64_t j = lower_row(rows, n_rows, t); j < n_rows && rows[j].tid == t; j++) {
                    const int32_t R = 10;                                          /* junction.cc:936 */
                    int32_t ds = rows[j].start - 2 * R, dm = rows[j].start - R, de = rows[j].start;
                    int32_t as = rows[j].end, am = rows[j].end + R, ae = rows[j].end + 2 * R;
                    uint32_t* cs = out[j].cov_sum;
                    double donor = coverage_window(ds, dm - 1, depth_prev, target_len[prev_cov], &cs[0]) - coverage_window(dm, de, depth_prev, target_len[prev_cov], &cs[1]);
                    double acceptor = coverage_window(am, ae, depth_prev, target_len[prev_cov], &cs[2]) - coverage_window(as, am - 1, depth_prev, target_len[prev_cov], &cs[3]);
                    out[j].coverage = donor + acceptor;
                }
            }
            uint32_t* sw = depth_prev; depth_prev = depth; depth = sw;
            prev_cov = t; last_cov = t;
        }
        if (last_cov >= 0) {             /* final batch: res == 0, last.ref still names the target the batch belongs to */
            int32_t t = last_cov;
            for (int64_t j = lower_row(rows, n_rows, t); j < n_rows && rows[j].tid == t; j++) {
                const int32_t R = 10;
                int32_t ds = rows[j].start - 2 * R, dm = rows[j].start - R, de = rows[j].start;
                int32_t as = rows[j].end, am = rows[j].end + R, ae = rows[j].end + 2 * R;
                uint32_t* cs = out[j].cov_sum;
                double donor = coverage_window(ds, dm - 1, depth_prev, target_len[t], &cs[0]) - coverage_window(dm, de, depth_prev, target_len[t], &cs[1]);
                double acceptor = coverage_window(am, ae, depth_prev, target_len[t], &cs[2]) - coverage_window(as, am - 1, depth_prev, target_len[t], &cs[3]);
                out[j].coverage = donor + acceptor;
            }
        }
        free(depth_prev);
    }
done:
    for (int64_t k = 0; k < S.n; k++) free(S.j[k].reads);
    free(S.j); free(S.table);
    free(nm.key); free(nm.val); free(nm.used); free(uns); free(uoff); free(pmax); free(depth); free(live.v);
    return rc;
}
