"""End-to-end parity: the reference PROGRAM (its unchanged indelminer.c, evidence.c, graph.c, variant.c,
shared.c, bundled samtools) with its alignment path replaced by host/indelgpu_attempt.c +
libindelgpu.so must print the same VCF as the unmodified reference on the reference's own
test_data.  oracle/_ref/indelminer_gpu is built here by `make -C oracle gpuprog` (it needs the
reference sources) and travels to the GPU box as a built file; the golden VCFs were printed by
oracle/_ref/indelminer_ref in the build container (oracle/make_golden.sh)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
PROG = os.path.join(ROOT, "oracle", "_ref", "indelminer_gpu")

pytestmark = pytest.mark.gpu


def run_gpu_program(extra):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(PROG):
        pytest.skip("oracle/_ref/indelminer_gpu not built (needs /root/reference; make -C oracle gpuprog)")
    from indelminer_b200 import build
    build.build()
    cmd = [PROG] + extra + ["-i", "indelminer.config", "testdata_reference.fa", "sample=alignments.bam"]
    r = subprocess.run(cmd, cwd=GOLD, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("flags,golden", [([], "testdata_refrun.vcf"),
                                          (["-g", "4"], "testdata_refrun_g4.vcf"),
                                          (["-g", "16"], "testdata_refrun_g16.vcf")])
def test_vcf_identical_to_reference(flags, golden):
    out = run_gpu_program(flags)
    with open(os.path.join(GOLD, golden)) as f:
        want = f.read()
    assert out == want


def test_vcf_vs_upstream_expected_differs_only_in_the_known_token():
    """the reference's own golden file differs from the reference built here in one BF token
    (libc qsort tie order in the evidence code); the GPU build inherits exactly that."""
    out = run_gpu_program([]).splitlines()
    with open(os.path.join(GOLD, "testdata_expected.vcf")) as f:
        exp = f.read().splitlines()
    assert len(out) == len(exp)
    diff = [(a, b) for a, b in zip(out, exp) if a != b]
    assert len(diff) <= 1
    for a, b in diff:
        assert a.replace("BF=52,48", "BF=48,52") == b


@pytest.mark.parametrize("flags,golden", [([], "testdata_refrun.vcf"), (["-g", "4"], "testdata_refrun_g4.vcf")])
def test_record_replay_batched_mode_prints_the_same_vcf(tmp_path, flags, golden):
    """the batched integration with an UNCHANGED caller (host/indelgpu_attempt.c): run 1 records every
    attempt_pe_alignment call and realigns them with one indelgpu_realign_batch per contig at exit,
    run 2 replays the results in call order and prints the VCF"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(PROG):
        pytest.skip("oracle/_ref/indelminer_gpu not built")
    from indelminer_b200 import build
    build.build()
    env = dict(os.environ, INDELGPU_REPLAY_FILE=str(tmp_path / "replay.bin"))
    cmd = [PROG] + flags + ["-i", "indelminer.config", "testdata_reference.fa", "sample=alignments.bam"]
    r1 = subprocess.run(cmd, cwd=GOLD, capture_output=True, text=True, timeout=600, env=dict(env, INDELGPU_MODE="record"))
    assert r1.returncode == 0, r1.stderr[-2000:]
    assert "697 candidate reads realigned in batches" in r1.stderr
    r2 = subprocess.run(cmd, cwd=GOLD, capture_output=True, text=True, timeout=600, env=dict(env, INDELGPU_MODE="replay"))
    assert r2.returncode == 0, r2.stderr[-2000:]
    with open(os.path.join(GOLD, golden)) as f:
        assert r2.stdout == f.read()
    assert r1.stdout != r2.stdout          # the recording run answered NULL everywhere: its VCF is not the result


@pytest.mark.parametrize("flags,golden", [([], "testdata_refrun.vcf"), (["-g", "16"], "testdata_refrun_g16.vcf")])
def test_auto_mode_single_run_batched(flags, golden):
    """INDELGPU_MODE=auto: one command; the process forks its own recording run at the first call (which
    gives itself private descriptions of the open BAM and streams its batches through a pipe) and
    continues, concurrently, as the replay run"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(PROG):
        pytest.skip("oracle/_ref/indelminer_gpu not built")
    from indelminer_b200 import build
    build.build()
    cmd = [PROG] + flags + ["-i", "indelminer.config", "testdata_reference.fa", "sample=alignments.bam"]
    r = subprocess.run(cmd, cwd=GOLD, capture_output=True, text=True, timeout=600, env=dict(os.environ, INDELGPU_MODE="auto"))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "697 candidate reads realigned in batches" in r.stderr
    with open(os.path.join(GOLD, golden)) as f:
        assert r.stdout == f.read()


@pytest.mark.parametrize("args,golden", [(["-i", "indelminer.config"], "testdata_refrun.vcf"),
                                         ([], "testdata_refrun_noconfig.vcf"),
                                         (["-g", "4", "-i", "indelminer.config"], "testdata_refrun_g4.vcf"),
                                         (["-g", "16", "-i", "indelminer.config"], "testdata_refrun_g16.vcf")])
def test_inline_mode_one_pass_batched(args, golden):
    """INDELGPU_MODE=inline (host/indelgpu_inline.c): one pass over the BAM, no fork; a
    prefetching thread realigns each block of records in one indelgpu_realign_batch while fetch_func -- unchanged --
    consumes the previous block.  The second case is BASELINE config 1 as written (no -i: insert ranges and coverage
    are estimated from the BAM first, bamoperations.c:62-147)."""
    import re
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(PROG):
        pytest.skip("oracle/_ref/indelminer_gpu not built")
    from indelminer_b200 import build
    build.build()
    cmd = [PROG] + args + ["testdata_reference.fa", "sample=alignments.bam"]
    r = subprocess.run(cmd, cwd=GOLD, capture_output=True, text=True, timeout=600, env=dict(os.environ, INDELGPU_MODE="inline"))
    assert r.returncode == 0, r.stderr[-2000:]
    with open(os.path.join(GOLD, golden)) as f:
        assert r.stdout == f.read()
    m = re.search(r"(\d+) calls answered from (\d+) prefetched batches \((\d+) reads realigned in them\), (\d+) computed per read", r.stderr)
    assert m and int(m.group(1)) + int(m.group(4)) == 697 and int(m.group(1)) > 400


def test_config1_without_config_file_per_read():
    """BASELINE config 1 through the per-read path as well"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(PROG):
        pytest.skip("oracle/_ref/indelminer_gpu not built")
    r = subprocess.run([PROG, "testdata_reference.fa", "sample=alignments.bam"], cwd=GOLD, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(os.path.join(GOLD, "testdata_refrun_noconfig.vcf")) as f:
        assert r.stdout == f.read()
