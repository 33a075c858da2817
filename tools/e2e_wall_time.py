#!/usr/bin/env python
"""End-to-end VCF wall time on a synthetic contig (BASELINE.json: "end-to-end VCF wall time"): the
unmodified reference program against the same program with the GPU alignment path (batched record /
replay as two runs, auto = both concurrently in one command, and per-read), same command line, same BAM; VCFs must be identical.
  python tools/e2e_wall_time.py [--length 4000000] [--depth 20] [--out profiles/r01_e2e_wall_time.json]
Needs oracle/_ref/{indelminer_ref,indelminer_gpu,sam2bam} (built where /root/reference exists)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run(exe, cwd, flags, env=None):
    cmd = [os.path.join(REFDIR, exe)] + flags + ["-i", "d.config", "d.fa", "sample=d.bam"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise SystemExit(f"{exe} failed: {r.stderr[-1500:]}")
    return r.stdout, dt, r.stderr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=4_000_000)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--per-read", action="store_true", help="also time the per-read glue mode")
    ap.add_argument("--only-default", action="store_true", help="default flags only (skip -g 4)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from tests.synth_bam import make_dataset
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        info = make_dataset(os.path.join(d, "d"), length=a.length, depth=a.depth, seed=11)
        subprocess.check_call([os.path.join(REFDIR, "sam2bam"), "d.sam", "d.bam"], cwd=d, stderr=subprocess.DEVNULL)
        t_gen = time.perf_counter() - t0
        res = dict(dataset=dict(length=a.length, depth=a.depth, **info), generate_s=t_gen, runs=[])
        for flags in ([[]] if a.only_default else [[], ["-g", "4"]]):
            ref_vcf, t_ref, _ = run("indelminer_ref", d, flags)
            env = dict(INDELGPU_REPLAY_FILE=os.path.join(d, "replay.bin"))
            _v, t_rec, err = run("indelminer_gpu", d, flags, dict(env, INDELGPU_MODE="record"))
            vcf, t_rep, _ = run("indelminer_gpu", d, flags, dict(env, INDELGPU_MODE="replay"))
            row = dict(flags=flags, variants=sum(1 for ln in ref_vcf.splitlines() if not ln.startswith("#")),
                       candidates=next((int(w) for ln in err.splitlines() if "candidate reads realigned" in ln for w in ln.split() if w.isdigit()), None),
                       identical_vcf=(vcf == ref_vcf),
                       wall_s=dict(reference=t_ref, gpu_record=t_rec, gpu_replay=t_rep, gpu_total=t_rec + t_rep))
            v3, t_auto, _ = run("indelminer_gpu", d, flags, dict(INDELGPU_MODE="auto"))     # both passes at once, one command
            row["wall_s"]["gpu_auto"] = t_auto
            row["identical_vcf_auto"] = (v3 == ref_vcf)
            if a.per_read:
                v2, t_dir, _ = run("indelminer_gpu", d, flags)
                row["wall_s"]["gpu_per_read"] = t_dir
                row["identical_vcf_per_read"] = (v2 == ref_vcf)
            res["runs"].append(row)
            print(json.dumps(row), flush=True)
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
