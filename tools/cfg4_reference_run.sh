#!/bin/sh
# TEST INFRASTRUCTURE.  BASELINE config 4 at its named depth (tumor / normal, 30x / 30x, `-q 0 -a -e 1`): generate a
# tumor and a normal BAM with oracle/_ref/synth_bam (same genome; the normal carries half of the planted indels and
# its own reads), call the tumor with the UNMODIFIED reference program, annotate that VCF with the normal's BAM --
# again the unmodified reference -- and write the md5 of both VCFs to tests/golden/cfg4_reference.json.
# tests/test_e2e_configs.py regenerates the same BAMs on the GPU box and requires the same hashes.
#   tools/cfg4_reference_run.sh [WORKDIR] [LENGTH] [DEPTH] [OUT.json]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/cfg4}; LEN=${2:-16000000}; DEPTH=${3:-30}; OUT=${4:-$HERE/tests/golden/cfg4_reference.json}
G=$HERE/oracle/_ref
mkdir -p "$W"; cd "$W"
"$G/synth_bam" tumor --length "$LEN" --depth "$DEPTH" --seed 11 > tumor.json
"$G/synth_bam" normal --length "$LEN" --depth "$DEPTH" --seed 11 --keep 0.5 --readseed 99 > normal.json
s=$(date +%s); nice "$G/indelminer_ref" -i tumor.config tumor.fa tumor=tumor.bam > tumor.vcf 2> tumor.err; t1=$(( $(date +%s) - s ))
s=$(date +%s); nice "$G/indelminer_ref" -q 0 -a -e 1 -i normal.config normal.fa tumor.vcf normal=normal.bam > annotated.vcf 2> annot.err; t2=$(( $(date +%s) - s ))
python3 - "$W" "$LEN" "$DEPTH" "$t1" "$t2" "$OUT" <<'PY'
import hashlib, json, sys
w, length, depth, t1, t2, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()
body = lambda p: [l for l in open(p) if not l.startswith("#")]
json.dump({"generator": [f"oracle/_ref/synth_bam tumor --length {length} --depth {depth} --seed 11",
                         f"oracle/_ref/synth_bam normal --length {length} --depth {depth} --seed 11 --keep 0.5 --readseed 99"],
           "tumor": json.load(open(f"{w}/tumor.json")), "normal": json.load(open(f"{w}/normal.json")),
           "tumor_bam_md5": md5(f"{w}/tumor.bam"), "normal_bam_md5": md5(f"{w}/normal.bam"),
           "commands": ["indelminer_ref -i tumor.config tumor.fa tumor=tumor.bam > tumor.vcf",
                        "indelminer_ref -q 0 -a -e 1 -i normal.config normal.fa tumor.vcf normal=normal.bam > annotated.vcf"],
           "tumor_vcf_md5": md5(f"{w}/tumor.vcf"), "tumor_records": len(body(f"{w}/tumor.vcf")), "tumor_reference_seconds": t1,
           "annotated_vcf_md5": md5(f"{w}/annotated.vcf"), "annotated_records": len(body(f"{w}/annotated.vcf")),
           "annotated_tagged": sum(l.rstrip().endswith(";normal") for l in body(f"{w}/annotated.vcf")), "annotate_reference_seconds": t2},
          open(out, "w"), indent=1)
print(open(out).read())
PY
