"""End-to-end parity in ANNOTATE mode (BASELINE config 4: `-q 0 -a -e 1 ref.fa tumor.vcf normal=normal.bam`)
on a synthetic tumor / normal pair: the tumor's VCF is annotated with the normal's BAM by
  * oracle/_ref/indelminer_ref        the unmodified reference program, and
  * oracle/_ref/indelminer_gpu_annot  the same program with host/indelgpu_attempt.c (attempt_pe_alignment)
                                      and host/indelgpu_support.c (realign_with_indel)
                                      on libindelgpu.so -- per call, record / replay, and auto.
The annotated VCFs must be byte-identical."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for exe in ("indelminer_ref", "indelminer_gpu_annot", "sam2bam"):
        if not os.path.exists(os.path.join(REFDIR, exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (needs /root/reference: make -C oracle refprog gpuprog_annotate tools)")
    from indelminer_b200 import build
    build.build()
    from tests.synth_bam import make_dataset
    d = str(tmp_path_factory.mktemp("tn"))
    tumor = make_dataset(os.path.join(d, "tumor"), length=400_000, depth=15, seed=11)
    normal = make_dataset(os.path.join(d, "normal"), length=400_000, depth=15, seed=11, keep_frac=0.5, read_seed=99)
    for n in ("tumor", "normal"):
        subprocess.check_call([os.path.join(REFDIR, "sam2bam"), os.path.join(d, n + ".sam"), os.path.join(d, n + ".bam")],
                              stderr=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(REFDIR, "indelminer_ref"), "-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"],
                       cwd=d, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(os.path.join(d, "tumor.vcf"), "w") as f:
        f.write(r.stdout)
    return dict(dir=d, tumor=tumor, normal=normal)


def annotate(exe, pair, flags, env=None):
    cmd = [os.path.join(REFDIR, exe), "-q", "0", "-a", "-e", "1"] + flags + ["-i", "normal.config", "normal.fa", "tumor.vcf", "normal=normal.bam"]
    r = subprocess.run(cmd, cwd=pair["dir"], capture_output=True, text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


@pytest.mark.parametrize("flags", [[], ["-g", "4"]])
def test_annotated_vcf_identical(pair, flags):
    ref_vcf, _ = annotate("indelminer_ref", pair, flags)
    body = [ln for ln in ref_vcf.splitlines() if not ln.startswith("#")]
    tagged = sum(ln.endswith(";normal") for ln in body)
    assert len(body) > 150 and 40 < tagged < len(body) - 40          # some variants are in the normal, some are not
    direct, _ = annotate("indelminer_gpu_annot", pair, flags)
    assert direct == ref_vcf
    replay = os.path.join(pair["dir"], "replay.bin")
    _junk, log = annotate("indelminer_gpu_annot", pair, flags, dict(INDELGPU_MODE="record", INDELGPU_REPLAY_FILE=replay))
    m = re.search(r"(\d+) \(variant, read\) pairs scored in one batch", log)
    assert m and int(m.group(1)) > 100                                # the support check really ran, batched
    batched, _ = annotate("indelminer_gpu_annot", pair, flags, dict(INDELGPU_MODE="replay", INDELGPU_REPLAY_FILE=replay))
    assert batched == ref_vcf
    auto, _ = annotate("indelminer_gpu_annot", pair, flags, dict(INDELGPU_MODE="auto"))
    assert auto == ref_vcf
    # one pass, no fork: realignment prefetched per block, the support check batched per known variant (two-pass loop)
    inline, log = annotate("indelminer_gpu_annot", pair, flags, dict(INDELGPU_MODE="inline"))
    assert inline == ref_vcf
    m = re.search(r"(\d+) known variants checked, (\d+) \(variant, read\) pairs scored in (\d+) batches", log)
    assert m and int(m.group(2)) > 100
