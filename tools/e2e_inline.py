#!/usr/bin/env python
"""End-to-end VCF wall time (the third part of BASELINE.json's metric): the unmodified reference PROGRAM
against the same program with the GPU alignment path in its one-pass batched mode (INDELGPU_MODE=inline,
host/indelgpu_inline.c), same command line, same BAM; the VCFs must be byte-identical.

  python tools/e2e_inline.py [--length 4000000] [--depth 30] [--modes inline,auto] [--no-reference]
                             [--regions 8] [--out profiles/...json]

The data set comes from oracle/_ref/synth_bam (deterministic).  bench.py imports vcf_wall_time() for the
`vcf_wall_time` object of its JSON line.  Needs oracle/_ref/{indelminer_ref,indelminer_gpu,synth_bam}
(built where /root/reference exists; they travel to the GPU box as built files)."""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def have_programs():
    return all(os.path.exists(os.path.join(REFDIR, e)) for e in ("indelminer_ref", "indelminer_gpu", "synth_bam"))


def generate(workdir, length, depth, seed=20261018, name="d", extra=()):
    t0 = time.perf_counter()
    out = subprocess.run([os.path.join(REFDIR, "synth_bam"), name, "--length", str(length), "--depth", str(depth),
                          "--seed", str(seed)] + list(extra), cwd=workdir, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("synth_bam failed: " + out.stderr[-500:])
    info = json.loads(out.stdout)
    info["generate_s"] = time.perf_counter() - t0
    return info


def run_program(exe, workdir, args, env=None, stdout_path=None):
    cmd = [os.path.join(REFDIR, exe)] + args
    t0 = time.perf_counter()
    if stdout_path:
        with open(stdout_path, "w") as f:
            r = subprocess.run(cmd, cwd=workdir, stdout=f, stderr=subprocess.PIPE, text=True, env=dict(os.environ, **(env or {})))
        out = None
    else:
        r = subprocess.run(cmd, cwd=workdir, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        out = r.stdout
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{exe} failed ({r.returncode}): {r.stderr[-1500:]}")
    return out, dt, r.stderr


def vcf_wall_time(length=4_000_000, depth=30, modes=("inline",), reference=True, workdir=None):
    """One data set, the reference once (1 thread: it has none), the GPU program once per mode.
    Returns {reference_s, gpu_s, records, identical, ...}; gpu_s is the first mode's wall time."""
    own = workdir is None
    tmp = tempfile.TemporaryDirectory() if own else None
    d = tmp.name if own else workdir
    try:
        info = generate(d, length, depth)
        args = ["-i", "d.config", "d.fa", "sample=d.bam"]
        res = {"dataset": f"synth_bam: {length} bp contig, {depth}x 2x150 bp, planted 1-50 bp indels every ~2 kb",
               "bam_records": info["records"], "generate_s": round(info["generate_s"], 2),
               "command": "indelminer -i d.config d.fa sample=d.bam"}
        ref_vcf = None
        if reference:
            ref_vcf, t_ref, _ = run_program("indelminer_ref", d, args)
            res["reference_s"] = round(t_ref, 3)
            res["reference_threads"] = 1
            res["records"] = sum(1 for ln in ref_vcf.splitlines() if not ln.startswith("#"))
        for k, mode in enumerate(modes):
            vcf, t, err = run_program("indelminer_gpu", d, args, dict(INDELGPU_MODE=mode, INDELGPU_VERBOSE="1"))
            key = "gpu_s" if k == 0 else f"gpu_{mode}_s"
            res[key] = round(t, 3)
            if k == 0:
                res["gpu_mode"] = f"INDELGPU_MODE={mode}"
                res["records"] = sum(1 for ln in vcf.splitlines() if not ln.startswith("#"))
                res["vcf_md5"] = hashlib.md5(vcf.encode()).hexdigest()
                res["gpu_log"] = [ln.strip() for ln in err.splitlines() if ln.startswith("libindelgpu:")]
            if ref_vcf is not None:
                res["identical" if k == 0 else f"identical_{mode}"] = (vcf == ref_vcf)
        return res
    finally:
        if own:
            tmp.cleanup()


def regions_run(workdir, length, nreg, exe="indelminer_gpu", env=None, ndev=1, concurrent=True):
    """the reference's own scale-out: one process per `-c` region (indelminer.c:536-542), one GPU per process"""
    step = length // nreg
    regs = [f"chr1:{i * step + 1}-{(i + 1) * step}" for i in range(nreg)]
    t0 = time.perf_counter()
    procs = []
    for i, r in enumerate(regs):
        e = dict(os.environ, **(env or {}), INDELGPU_DEVICE=str(i % ndev))
        f = open(os.path.join(workdir, f"gpu_{i}.vcf"), "w")
        p = subprocess.Popen([os.path.join(REFDIR, exe), "-i", "d.config", "-c", r, "d.fa", "s=d.bam"], cwd=workdir, stdout=f, stderr=subprocess.DEVNULL, env=e)
        procs.append((p, f))
        if not concurrent:
            p.wait()
    for p, f in procs:
        p.wait()
        f.close()
        if p.returncode != 0:
            raise RuntimeError("a region process failed")
    dt = time.perf_counter() - t0
    out = []
    for i, r in enumerate(regs):
        data = open(os.path.join(workdir, f"gpu_{i}.vcf"), "rb").read()
        out.append({"region": r, "vcf_md5": hashlib.md5(data).hexdigest(),
                    "records": sum(1 for ln in data.decode().splitlines() if not ln.startswith("#"))})
    return out, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=4_000_000)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--modes", default="inline,auto")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    res = vcf_wall_time(a.length, a.depth, tuple(a.modes.split(",")), reference=not a.no_reference)
    print(json.dumps(res), flush=True)
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    main()
