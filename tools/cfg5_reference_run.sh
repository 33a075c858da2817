#!/bin/sh
# TEST INFRASTRUCTURE.  BASELINE config 5's SHAPE at a size the unmodified reference can run: 24 contigs (sized
# like a scaled-down GRCh38 chr1..22, X, Y), 30x, 2x150 bp, planted indels; the whole 3.1 Gb / 620 M-read BAM of the
# config cannot be generated or read by the single-threaded reference in any budget available here.
# The unmodified reference runs once over all contigs and once per `-c` contig for two of them (the reference's own
# region sharding, indelminer.c:536-542); md5s go to tests/golden/cfg5_reference.json.  The GPU test regenerates the
# BAM and runs the same commands, region-sharded one process per contig group.
#   tools/cfg5_reference_run.sh [WORKDIR] [SCALE_DIVISOR] [DEPTH] [OUT.json]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/cfg5}; DIV=${2:-100}; DEPTH=${3:-30}; OUT=${4:-$HERE/tests/golden/cfg5_reference.json}
G=$HERE/oracle/_ref
mkdir -p "$W"; cd "$W"
# GRCh38 chromosome lengths (Mb, rounded), divided by DIV
LENS=$(python3 -c "
mb=[248,242,198,190,182,171,159,145,138,134,135,133,114,107,102,90,83,80,59,64,47,51,156,57]
print(','.join(str(int(x*1000000//$DIV)) for x in mb))")
"$G/synth_bam" g --contigs 24 --lengths "$LENS" --depth "$DEPTH" --seed 5 > gen.json
s=$(date +%s); nice "$G/indelminer_ref" -i g.config g.fa s=g.bam > all.vcf 2> all.err; t=$(( $(date +%s) - s ))
nice "$G/indelminer_ref" -i g.config -c chr7 g.fa s=g.bam > chr7.vcf 2> chr7.err
nice "$G/indelminer_ref" -i g.config -c chr21 g.fa s=g.bam > chr21.vcf 2> chr21.err
python3 - "$W" "$LENS" "$DEPTH" "$t" "$OUT" <<'PY'
import hashlib, json, sys
w, lens, depth, t, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()
body = lambda p: [l for l in open(p) if not l.startswith("#")]
json.dump({"generator": f"oracle/_ref/synth_bam g --contigs 24 --lengths {lens} --depth {depth} --seed 5",
           "generated": json.load(open(f"{w}/gen.json")), "bam_md5": md5(f"{w}/g.bam"),
           "runs": {"all": {"args": [], "vcf_md5": md5(f"{w}/all.vcf"), "records": len(body(f"{w}/all.vcf")), "reference_seconds": t},
                    "chr7": {"args": ["-c", "chr7"], "vcf_md5": md5(f"{w}/chr7.vcf"), "records": len(body(f"{w}/chr7.vcf"))},
                    "chr21": {"args": ["-c", "chr21"], "vcf_md5": md5(f"{w}/chr21.vcf"), "records": len(body(f"{w}/chr21.vcf"))}}},
          open(out, "w"), indent=1)
print(open(out).read())
PY
