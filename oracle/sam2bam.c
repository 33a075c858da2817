/* TEST INFRASTRUCTURE ONLY.  SAM (with @SQ header) -> coordinate-sorted-as-given BAM + .bai, through the
 * reference's bundled samtools-0.1.19 API (libbam.a, compiled where it lies by oracle/Makefile).
 * Used by tests/test_e2e_synthetic.py to turn synthetic reads into a BAM the reference program can
 * fetch from; there is no samtools binary in the image.
 *   sam2bam in.sam out.bam */
#include <stdio.h>
#include <stdlib.h>

#include "sam.h"

int main(int argc, char** argv)
{
    if (argc != 3) { fprintf(stderr, "usage: sam2bam in.sam out.bam\n"); return 2; }
    samfile_t* in = samopen(argv[1], "r", NULL);
    if (in == NULL || in->header == NULL) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    samfile_t* out = samopen(argv[2], "wb", in->header);
    if (out == NULL) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    bam1_t* b = bam_init1();
    long n = 0;
    while (samread(in, b) >= 0) { samwrite(out, b); n++; }
    bam_destroy1(b);
    samclose(out);
    samclose(in);
    if (bam_index_build(argv[2]) != 0) { fprintf(stderr, "indexing %s failed\n", argv[2]); return 1; }
    fprintf(stderr, "sam2bam: %ld records\n", n);
    return 0;
}
