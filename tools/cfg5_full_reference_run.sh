#!/bin/sh
# TEST INFRASTRUCTURE.  BASELINE config 5 at its named REFERENCE size, end to end: 24 contigs with the lengths of
# GRCh38 chr1..22, X, Y (3.1 Gb), 2x150 bp pairs with planted indels -- at 0.05x coverage instead of 30x, because the
# unmodified reference spends ~20 ms per candidate read in strlen() of a 248 Mb contig (alignment.c:771): the
# ~80 k calls of this data set already cost it about an hour, 30x would cost it a week.  `-e 1` (minimum support 1)
# so that the thin coverage still yields thousands of calls.  Writes tests/golden/cfg5_full_reference.json.
#   tools/cfg5_full_reference_run.sh [WORKDIR] [DEPTH] [OUT.json]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/cfg5full}; DEPTH=${2:-0.05}; OUT=${3:-$HERE/tests/golden/cfg5_full_reference.json}
G=$HERE/oracle/_ref
mkdir -p "$W"; cd "$W"
LENS=$(python3 -c "
mb=[248,242,198,190,182,171,159,145,138,134,135,133,114,107,102,90,83,80,59,64,47,51,156,57]
print(','.join(str(x*1000000) for x in mb))")
[ -f g.bam.bai ] || "$G/synth_bam" g --contigs 24 --lengths "$LENS" --depth "$DEPTH" --seed 55 > gen.json
s=$(date +%s); nice "$G/indelminer_ref" -e 1 -i g.config g.fa s=g.bam > all.vcf 2> all.err; t=$(( $(date +%s) - s ))
python3 - "$W" "$LENS" "$DEPTH" "$t" "$OUT" <<'PY'
import hashlib, json, sys
w, lens, depth, t, out = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
def md5(p):
    h = hashlib.md5()
    with open(p, "rb") as f:
        for b in iter(lambda: f.read(1 << 22), b""):
            h.update(b)
    return h.hexdigest()
body = [l for l in open(f"{w}/all.vcf") if not l.startswith("#")]
json.dump({"generator": f"oracle/_ref/synth_bam g --contigs 24 --lengths {lens} --depth {depth} --seed 55",
           "generated": json.load(open(f"{w}/gen.json")), "bam_md5": md5(f"{w}/g.bam"),
           "command": "indelminer_ref -e 1 -i g.config g.fa s=g.bam   (unmodified reference, one process, all 24 contigs)",
           "vcf_md5": md5(f"{w}/all.vcf"), "records": len(body), "reference_seconds": t}, open(out, "w"), indent=1)
print(open(out).read())
PY
