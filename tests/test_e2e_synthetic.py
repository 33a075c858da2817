"""End-to-end VCF parity on SYNTHETIC data with planted insertions and deletions (the reference's own
test_data has no insertion): tests/synth_bam.py writes a SAM, oracle/_ref/sam2bam (bundled samtools API)
turns it into an indexed BAM, and the same command line is run through
  * oracle/_ref/indelminer_ref   the unmodified reference program (built by oracle/Makefile), and
  * oracle/_ref/indelminer_gpu   the same program with host/indelgpu_attempt.c + libindelgpu.so,
                                 in per-read mode and in batched record / replay mode.
The VCFs must be byte-identical.  Wall times go to gpurun_out/e2e_synthetic.json when that directory exists."""
import json
import os
import subprocess
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for exe in ("indelminer_ref", "indelminer_gpu", "sam2bam"):
        if not os.path.exists(os.path.join(REFDIR, exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (needs /root/reference: make -C oracle refprog gpuprog tools)")
    from indelminer_b200 import build
    build.build()
    from tests.synth_bam import make_dataset
    d = tmp_path_factory.mktemp("synth")
    prefix = str(d / "d")
    info = make_dataset(prefix, length=400_000, depth=15, seed=11)
    subprocess.check_call([os.path.join(REFDIR, "sam2bam"), prefix + ".sam", prefix + ".bam"], stderr=subprocess.DEVNULL)
    return dict(dir=str(d), info=info)


def run(exe, dataset, flags, env=None):
    cmd = [os.path.join(REFDIR, exe)] + flags + ["-i", "d.config", "d.fa", "sample=d.bam"]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=dataset["dir"], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, dt


@pytest.mark.parametrize("flags", [[], ["-g", "4"], ["-k", "8"]])
def test_synthetic_vcf_identical(dataset, flags):
    ref_vcf, t_ref = run("indelminer_ref", dataset, flags)
    body = [ln for ln in ref_vcf.splitlines() if not ln.startswith("#")]
    assert sum("INSERTION" in ln for ln in body) > 30 and sum("DELETION" in ln for ln in body) > 30
    replay = os.path.join(dataset["dir"], "replay.bin")
    _junk, t_rec = run("indelminer_gpu", dataset, flags, dict(INDELGPU_MODE="record", INDELGPU_REPLAY_FILE=replay))
    batched_vcf, t_rep = run("indelminer_gpu", dataset, flags, dict(INDELGPU_MODE="replay", INDELGPU_REPLAY_FILE=replay))
    assert batched_vcf == ref_vcf
    direct_vcf, t_dir = run("indelminer_gpu", dataset, flags)
    assert direct_vcf == ref_vcf
    auto_vcf, t_auto = run("indelminer_gpu", dataset, flags, dict(INDELGPU_MODE="auto"))
    assert auto_vcf == ref_vcf
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "e2e_synthetic.json"), "a") as f:
            f.write(json.dumps(dict(flags=flags, dataset=dataset["info"], variants=len(body),
                                    wall_s=dict(reference=t_ref, gpu_record=t_rec, gpu_replay=t_rep, gpu_per_read=t_dir, gpu_auto=t_auto))) + "\n")


def test_region_runs_match_the_reference(dataset):
    """the reference's own scale-out: one process per region (-c, indelminer.c:711); the GPU-linked program
    run the same way (INDELGPU_MODE=auto, INDELGPU_DEVICE chosen per process, as on a multi-GPU box) must
    print the reference's VCF for every region"""
    import torch
    ndev = torch.cuda.device_count()
    total = 0
    for r, region in enumerate(["chrS:1-200000", "chrS:200001-400000"]):
        ref_vcf, _ = run("indelminer_ref", dataset, ["-c", region])
        gpu_vcf, _ = run("indelminer_gpu", dataset, ["-c", region], dict(INDELGPU_MODE="auto", INDELGPU_DEVICE=str(r % ndev)))
        assert gpu_vcf == ref_vcf, region
        total += sum(1 for ln in ref_vcf.splitlines() if not ln.startswith("#"))
    assert total > 150
