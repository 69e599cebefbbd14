/* Test infrastructure only: SAM text -> BAM (+ .bai) using the reference's vendored htslib-1.3
 * (sam_open/sam_write1/sam_index_build, deps/htslib-1.3/sam.c). Used to make fixtures that the
 * reference binary (oracle/_ref/portcullis_ref) can read, and to cross-check our own BAM/BAI writer.
 *   bamtool sam2bam in.sam out.bam     (input must already be coordinate sorted)
 *   bamtool index in.bam
 *   bamtool index_csi in.bam           (writes in.bam.csi)
 *   bamtool view in.bam                (SAM text to stdout)
 *   bamtool sort in.bam out.bam | merge out.bam in1.bam in2.bam ...   (restated samtools 1.3 coordinate sort / merge)
 *   bamtool query in.bam idx tid:beg-end ...   (record count + checksum per region through a chosen index file)
 */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "htslib/sam.h"
#include "htslib/kstring.h"

int main(int argc, char** argv) {
    if (argc >= 4 && strcmp(argv[1], "sam2bam") == 0) {
        samFile* in = sam_open(argv[2], "r");
        if (!in) { fprintf(stderr, "cannot open %s\n", argv[2]); return 2; }
        bam_hdr_t* h = sam_hdr_read(in);
        samFile* out = sam_open(argv[3], "wb");
        if (!out || !h) { fprintf(stderr, "cannot open output/header\n"); return 2; }
        if (sam_hdr_write(out, h) != 0) return 2;
        bam1_t* b = bam_init1();
        long n = 0; int r;
        while ((r = sam_read1(in, h, b)) >= 0) { if (sam_write1(out, h, b) < 0) return 2; n++; }
        bam_destroy1(b); bam_hdr_destroy(h); sam_close(in); sam_close(out);
        if (r < -1) { fprintf(stderr, "truncated/invalid SAM\n"); return 2; }
        if (argc >= 5 && strcmp(argv[4], "noindex") == 0) return 0;          /* unsorted input for the prep tests */
        if (sam_index_build(argv[3], 0) != 0) { fprintf(stderr, "index build failed\n"); return 3; }
        fprintf(stderr, "wrote %ld records\n", n);
        return 0;
    }
    if (argc >= 3 && strcmp(argv[1], "index") == 0) {
        return sam_index_build(argv[2], 0) == 0 ? 0 : 3;
    }
    if (argc >= 3 && strcmp(argv[1], "index_csi") == 0) {
        return sam_index_build(argv[2], 14) == 0 ? 0 : 3;          /* min_shift 14 -> .csi */
    }
    if (argc >= 3 && strcmp(argv[1], "view") == 0) {
        samFile* in = sam_open(argv[2], "r");
        if (!in) return 2;
        bam_hdr_t* h = sam_hdr_read(in);
        bam1_t* b = bam_init1(); kstring_t s = {0, 0, 0};
        fputs(h->text, stdout);
        while (sam_read1(in, h, b) >= 0) { sam_format1(h, b, &s); puts(s.s); }
        return 0;
    }
    if (argc >= 4 && (strcmp(argv[1], "sort") == 0 || strcmp(argv[1], "merge") == 0)) {
        /* bamtool sort in.bam out.bam | bamtool merge out.bam in1.bam in2.bam ...
         * Stand-in for `samtools sort` / `samtools merge` (samtools 1.3, bam_sort.c — NOT part of /root/reference, which only
         * shells out to it: src/prepare.cc:182, 217).  Published behaviour restated: coordinate order compares
         * (uint64_t)tid << 32 | (pos + 1) << 1 | is_reverse (bam1_lt), the sort is a stable merge sort, merge breaks ties by
         * input file order, the header is the first input's with @HD ... SO:coordinate (change_SO).  Records and header are
         * read and written by htslib-1.3 itself. */
        const int is_merge = argv[1][0] == 'm';
        const char* outfn = is_merge ? argv[2] : argv[3];
        int n_in = is_merge ? argc - 3 : 1; char** infn = is_merge ? argv + 3 : argv + 2;
        bam1_t** recs = NULL; size_t n = 0, cap = 0; bam_hdr_t* h0 = NULL;
        for (int f = 0; f < n_in; f++) {
            samFile* in = sam_open(infn[f], "r");
            if (!in) { fprintf(stderr, "cannot open %s\n", infn[f]); return 2; }
            bam_hdr_t* h = sam_hdr_read(in);
            if (!h0) h0 = h;
            bam1_t* b = bam_init1();
            while (sam_read1(in, h, b) >= 0) {
                if (n == cap) { cap = cap ? cap * 2 : 1 << 16; recs = (bam1_t**)realloc(recs, cap * sizeof *recs); }
                recs[n++] = bam_dup1(b);
            }
            bam_destroy1(b); sam_close(in);
        }
        /* stable merge sort on the key (bottom-up, with a scratch array) */
        uint64_t* key = (uint64_t*)malloc((n + 1) * 8);
        for (size_t i = 0; i < n; i++) key[i] = (uint64_t)recs[i]->core.tid << 32 | (uint32_t)((recs[i]->core.pos + 1) << 1 | bam_is_rev(recs[i]));
        size_t* idx = (size_t*)malloc((n + 1) * sizeof(size_t)); size_t* tmp = (size_t*)malloc((n + 1) * sizeof(size_t));
        for (size_t i = 0; i < n; i++) idx[i] = i;
        for (size_t w = 1; w < n; w *= 2) {
            for (size_t lo = 0; lo < n; lo += 2 * w) {
                size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, a = lo, b2 = mid, o = lo;
                while (a < mid && b2 < hi) tmp[o++] = key[idx[b2]] < key[idx[a]] ? idx[b2++] : idx[a++];
                while (a < mid) tmp[o++] = idx[a++];
                while (b2 < hi) tmp[o++] = idx[b2++];
            }
            size_t* t = idx; idx = tmp; tmp = t;
        }
        /* header: @HD with SO:coordinate */
        {
            kstring_t t = {0, 0, 0};
            const char* text = h0->text ? h0->text : "";
            if (strncmp(text, "@HD", 3) == 0) {
                const char* e = strchr(text, '\n'); size_t hl = e ? (size_t)(e - text) : strlen(text);
                char* hd = (char*)malloc(hl + 1); memcpy(hd, text, hl); hd[hl] = 0;
                char* so = strstr(hd, "\tSO:");
                if (so) { char* q = strchr(so + 1, '\t'); kputsn(hd, so - hd, &t); kputs("\tSO:coordinate", &t); if (q) kputs(q, &t); }
                else { kputs(hd, &t); kputs("\tSO:coordinate", &t); }
                kputs(e ? e : "\n", &t);
                free(hd);
            } else { kputs("@HD\tVN:1.3\tSO:coordinate\n", &t); kputs(text, &t); }
            free(h0->text); h0->text = t.s; h0->l_text = (uint32_t)t.l;
        }
        samFile* out = sam_open(outfn, "wb");
        if (!out || sam_hdr_write(out, h0) != 0) return 3;
        for (size_t i = 0; i < n; i++) if (sam_write1(out, h0, recs[idx[i]]) < 0) return 3;
        sam_close(out);
        return 0;
    }
    if (argc >= 5 && strcmp(argv[1], "query") == 0) {
        /* bamtool query in.bam index_file tid:beg-end ... : htslib-1.3's own iterator (sam_itr_queryi, the call behind
         * BamReader::setRegion) over an index file of the caller's choice; prints, per region, the number of records and a
         * checksum of their (pos, flag, l_qseq, first 8 name bytes), so two index files can be compared on the same BAM. */
        samFile* in = sam_open(argv[2], "r");
        if (!in) return 2;
        bam_hdr_t* h = sam_hdr_read(in);
        hts_idx_t* idx = sam_index_load2(in, argv[2], argv[3]);
        if (!idx) { fprintf(stderr, "cannot load index %s\n", argv[3]); return 3; }
        bam1_t* b = bam_init1();
        for (int a = 4; a < argc; a++) {
            int tid, beg, end;
            if (sscanf(argv[a], "%d:%d-%d", &tid, &beg, &end) != 3) return 4;
            hts_itr_t* it = sam_itr_queryi(idx, tid, beg, end);
            unsigned long long n = 0, sum = 1469598103934665603ull;
            while (it && sam_itr_next(in, it, b) >= 0) {
                unsigned long long v = ((unsigned long long)(unsigned)b->core.pos << 32) ^ ((unsigned long long)b->core.flag << 16) ^ (unsigned)b->core.l_qseq;
                const char* q = bam_get_qname(b);
                for (int k = 0; k < 8 && q[k]; k++) v = v * 131 + (unsigned char)q[k];
                sum = (sum ^ v) * 1099511628211ull; n++;
            }
            if (it) hts_itr_destroy(it);
            printf("%s\t%llu\t%016llx\n", argv[a], n, sum);
        }
        (void)h;
        return 0;
    }
    fprintf(stderr, "usage: bamtool sam2bam in.sam out.bam | index in.bam | index_csi in.bam | view in.bam | query in.bam index tid:beg-end ...\n");
    return 1;
}
