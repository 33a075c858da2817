#!/bin/sh
# TEST INFRASTRUCTURE.  BASELINE config 3 at its named size (64 Mb contig, 30x, 2x150 bp, planted 1-50 bp
# indels): generate the data set with oracle/_ref/synth_bam (deterministic), run the UNMODIFIED reference
# program (oracle/_ref/indelminer_ref) on it as NREG `-c` region instances (indelminer.c:536-542, 711-713) in
# parallel -- the reference is single-threaded and spends ~6 ms per candidate in strlen(contig)
# (alignment.c:771), so one instance would take ~100 minutes -- and write the md5 + record count of every
# region's VCF to tests/golden/cfg3_reference.json.  Run in the build container (CPU time is not metered);
# tests/test_e2e_configs.py regenerates the same BAM on the GPU box and requires the same hashes.
#   tools/cfg_reference_run.sh [WORKDIR] [LENGTH] [DEPTH] [NREG] [OUT.json]
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
W=${1:-/tmp/cfg3}; LEN=${2:-64000000}; DEPTH=${3:-30}; NREG=${4:-8}; OUT=${5:-$HERE/tests/golden/cfg3_reference.json}
G=$HERE/oracle/_ref
mkdir -p "$W"; cd "$W"
[ -f d.bam.bai ] || "$G/synth_bam" d --length "$LEN" --depth "$DEPTH" --seed 20261018 > gen.json
STEP=$((LEN / NREG))
i=0
while [ $i -lt $NREG ]; do
    A=$((i * STEP + 1)); B=$(((i + 1) * STEP))
    ( s=$(date +%s); nice "$G/indelminer_ref" -i d.config -c "chr1:$A-$B" d.fa s=d.bam > ref_$i.vcf 2> ref_$i.err; echo $(( $(date +%s) - s )) > ref_$i.secs ) &
    i=$((i + 1))
done
wait
python3 - "$W" "$LEN" "$DEPTH" "$NREG" "$OUT" <<'PY'
import hashlib, json, sys
w, length, depth, nreg, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()
step = length // nreg
regs = []
for i in range(nreg):
    body = [l for l in open(f"{w}/ref_{i}.vcf") if not l.startswith("#")]
    regs.append({"region": f"chr1:{i * step + 1}-{(i + 1) * step}", "vcf_md5": md5(f"{w}/ref_{i}.vcf"), "records": len(body),
                 "reference_seconds": int(open(f"{w}/ref_{i}.secs").read())})
json.dump({"generator": f"oracle/_ref/synth_bam d --length {length} --depth {depth} --seed 20261018",
           "generated": json.load(open(f"{w}/gen.json")), "bam_md5": md5(f"{w}/d.bam"), "fa_md5": md5(f"{w}/d.fa"),
           "command": "indelminer_ref -i d.config -c REGION d.fa s=d.bam   (unmodified reference, oracle/_ref/indelminer_ref)",
           "note": "reference_seconds: wall time of each region instance, all instances running in parallel in the 8-vCPU build container",
           "regions": regs}, open(out, "w"), indent=1)
print(open(out).read())
PY
