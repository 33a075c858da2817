#!/usr/bin/env python
"""End-to-end wall time of the GPU-linked reference program run the way the reference scales out: one
process per region (-c), one GPU per process (INDELGPU_DEVICE), INDELGPU_MODE=auto -- against the same
program run once over the whole contig.  Synthetic contig as in tools/e2e_wall_time.py.
  python tools/e2e_regions.py [--length 16000000] [--depth 20] [--procs 2] [--out gpurun_out/e2e_regions.json]
Needs oracle/_ref/{indelminer_gpu,sam2bam}."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=16_000_000)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--procs", type=int, default=2)
    ap.add_argument("--visible", action="store_true", help="pick the GPU with CUDA_VISIBLE_DEVICES instead of INDELGPU_DEVICE")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    ndev = max(1, torch.cuda.device_count())
    from tests.synth_bam import make_dataset
    with tempfile.TemporaryDirectory() as d:
        info = make_dataset(os.path.join(d, "d"), length=a.length, depth=a.depth, seed=11)
        subprocess.check_call([os.path.join(REFDIR, "sam2bam"), "d.sam", "d.bam"], cwd=d, stderr=subprocess.DEVNULL)
        exe = os.path.join(REFDIR, "indelminer_gpu")
        base = [exe, "-i", "d.config"]
        env = dict(os.environ, INDELGPU_MODE="auto")
        t0 = time.perf_counter()
        whole = subprocess.run(base + ["d.fa", "sample=d.bam"], cwd=d, capture_output=True, text=True, env=env)
        t_whole = time.perf_counter() - t0
        assert whole.returncode == 0, whole.stderr[-1000:]
        step = a.length // a.procs
        regions = [f"chrS:{k * step + 1}-{a.length if k == a.procs - 1 else (k + 1) * step}" for k in range(a.procs)]
        # one region alone (what a single process of the sharded run costs)
        t0 = time.perf_counter()
        one = subprocess.run(base + ["-c", regions[0], "d.fa", "sample=d.bam"], cwd=d, capture_output=True, text=True, env=env)
        t_one = time.perf_counter() - t0
        assert one.returncode == 0, one.stderr[-1000:]
        t0 = time.perf_counter()
        files = [open(os.path.join(d, f"region{k}.vcf"), "w") for k in range(a.procs)]
        procs = [subprocess.Popen(base + ["-c", r, "d.fa", "sample=d.bam"], cwd=d, stdout=files[k], stderr=subprocess.DEVNULL,
                                  env=dict(env, CUDA_VISIBLE_DEVICES=str(k % ndev)) if a.visible else dict(env, INDELGPU_DEVICE=str(k % ndev)))
                 for k, r in enumerate(regions)]
        for p in procs:
            p.wait()
        t_regions = time.perf_counter() - t0
        for f in files:
            f.close()
        assert all(p.returncode == 0 for p in procs)
        outs = [open(os.path.join(d, f"region{k}.vcf")).read() for k in range(a.procs)]
        body = lambda v: [ln for ln in v.splitlines() if not ln.startswith("#")]   # noqa: E731
        res = dict(dataset=dict(length=a.length, depth=a.depth, **info), gpus=ndev, procs=a.procs, regions=regions, device_choice="CUDA_VISIBLE_DEVICES" if a.visible else "INDELGPU_DEVICE",
                   variants_whole=len(body(whole.stdout)), variants_regions=sum(len(body(o)) for o in outs),
                   wall_s=dict(whole=t_whole, one_region_alone=t_one, regions_concurrent=t_regions))
        print(json.dumps(res), flush=True)
        if a.out:
            with open(os.path.join(ROOT, a.out), "w") as f:
                json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
