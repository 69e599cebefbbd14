/* Test infrastructure only: SAM text -> BAM (+ .bai) using the reference's vendored htslib-1.3
 * (sam_open/sam_write1/sam_index_build, deps/htslib-1.3/sam.c). Used to make fixtures that the
 * reference binary (oracle/_ref/portcullis_ref) can read, and to cross-check our own BAM/BAI writer.
 *   bamtool sam2bam in.sam out.bam     (input must already be coordinate sorted)
 *   bamtool index in.bam
 *   bamtool index_csi in.bam           (writes in.bam.csi)
 *   bamtool view in.bam                (SAM text to stdout)
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
    fprintf(stderr, "usage: bamtool sam2bam in.sam out.bam | index in.bam | view in.bam\n");
    return 1;
}
