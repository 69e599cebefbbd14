/* Test infrastructure only: SAM text -> BAM (+ .bai) using the reference's vendored htslib-1.3
 * (sam_open/sam_write1/sam_index_build, deps/htslib-1.3/sam.c). Used to make fixtures that the
 * reference binary (oracle/_ref/portcullis_ref) can read, and to cross-check our own BAM/BAI writer.
 *   bamtool sam2bam in.sam out.bam     (input must already be coordinate sorted)
 *   bamtool index in.bam
 *   bamtool index_csi in.bam           (writes in.bam.csi)
 *   bamtool view in.bam                (SAM text to stdout)
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
