"""CPU check of the synthetic end-to-end fixture chain (no GPU): tests/synth_bam.py -> oracle/_ref/sam2bam
-> the unmodified reference program.  Guards the generator's SAM records (flags, CIGARs, mate fields,
coordinate order) against the reference's own BAM reader and classifier."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def test_reference_calls_the_planted_indels(tmp_path):
    for exe in ("indelminer_ref", "sam2bam"):
        if not os.path.exists(os.path.join(REFDIR, exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (needs /root/reference: make -C oracle refprog tools)")
    from tests.synth_bam import make_dataset
    prefix = str(tmp_path / "d")
    info = make_dataset(prefix, length=120_000, depth=15, seed=3)
    assert info["nins"] > 10 and info["ndel"] > 10
    # coordinate order and mate symmetry of the SAM itself
    last, seen = -1, {}
    for ln in open(prefix + ".sam"):
        if ln.startswith("@"):
            continue
        t = ln.rstrip("\n").split("\t")
        pos, flag = int(t[3]), int(t[1])
        assert pos >= last
        last = pos
        assert flag & 0x1 and not (flag & 0x4 and flag & 0x8)
        seen.setdefault(t[0], []).append((flag, pos, int(t[7])))
    for name, ends in seen.items():
        assert len(ends) == 2, name
        (f1, p1, n1), (f2, p2, n2) = ends
        assert (f1 & 0xC0) != (f2 & 0xC0)                      # one first, one second in pair
        assert bool(f1 & 0x4) == bool(f2 & 0x8) and bool(f2 & 0x4) == bool(f1 & 0x8)
        if not (f1 & 0x4) and not (f2 & 0x4):
            assert n1 == p2 and n2 == p1
    subprocess.check_call([os.path.join(REFDIR, "sam2bam"), prefix + ".sam", prefix + ".bam"], stderr=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(REFDIR, "indelminer_ref"), "-i", "d.config", "d.fa", "sample=d.bam"],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    body = [ln for ln in r.stdout.splitlines() if not ln.startswith("#")]
    nsites = info["nsites"]
    assert len(body) >= 0.9 * nsites                           # nearly every planted site is called
    assert sum("INSERTION" in ln for ln in body) >= 0.8 * info["nins"]
    assert sum("DELETION" in ln for ln in body) >= 0.8 * info["ndel"]
