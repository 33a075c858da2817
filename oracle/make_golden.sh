#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/ from the reference itself
# (run in the build container, where /root/reference exists):
#   testdata_trace.tsv.gz   every local_align / attempt_pe_alignment call the reference makes on
#                           its own test_data (-i indelminer.config), inputs + outputs, recorded by
#                           oracle/_ref/indelminer_trace (reference sources + oracle/ref_shim.c)
#   testdata_reference.fa   the test_data contig (data fixture, needed to replay the PE lines)
#   testdata_refrun.vcf     VCF printed by the unmodified reference built here
#   testdata_refrun_noconfig.vcf  the same without -i (BASELINE config 1: IL / RC estimated from the BAM)
#   testdata_expected.vcf   the reference's own golden VCF (differs from refrun in one BF token)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${REF:-/root/reference}
OUT=$HERE/../tests/golden
make -s -C "$HERE" ref refprog REF="$REF"
mkdir -p "$OUT"
T=$(mktemp -d)
REFSHIM_TRACE_FILE=$T/trace.tsv "$HERE/_ref/indelminer_trace" -i "$REF/test_data/indelminer.config" \
    "$REF/test_data/reference.fa" sample="$REF/test_data/alignments.bam" > "$T/trace.vcf" 2> "$T/trace.err"
"$HERE/_ref/indelminer_ref" -i "$REF/test_data/indelminer.config" \
    "$REF/test_data/reference.fa" sample="$REF/test_data/alignments.bam" > "$OUT/testdata_refrun.vcf" 2> "$T/ref.err"
cmp "$T/trace.vcf" "$OUT/testdata_refrun.vcf"
# BASELINE config 1 as written: no -i, the insert ranges and coverage are estimated first (bamoperations.c:62-147)
"$HERE/_ref/indelminer_ref" "$REF/test_data/reference.fa" sample="$REF/test_data/alignments.bam" \
    > "$OUT/testdata_refrun_noconfig.vcf" 2> "$T/ref_noconfig.err"
gzip -9 -n -c "$T/trace.tsv" > "$OUT/testdata_trace.tsv.gz"
cp "$REF/test_data/reference.fa" "$OUT/testdata_reference.fa"
cp "$REF/test_data/indelminer.expected.vcf" "$OUT/testdata_expected.vcf"
rm -rf "$T"
ls -la "$OUT"
