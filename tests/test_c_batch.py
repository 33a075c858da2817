"""The host side in C: host/realign_tsv (C driver) -> host/indelgpu_batch.c (pinned SoA batch builder)
-> C ABI of libindelgpu.so -> GPU.  No Python between the input files and the segment lists; the
output is compared with the CPU oracle read by read."""
import os
import subprocess

import pytest

from tests.util import make_rng, rseq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def driver():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from indelminer_b200 import build
    build.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    return os.path.join(ROOT, "host", "realign_tsv")


@pytest.mark.parametrize("k,g,n", [(6, 0, 9000), (6, 3, 1500), (8, 0, 1200)])
def test_c_driver_matches_oracle(driver, oracle, tmp_path, k, g, n):
    rng = make_rng(31 * k + g)
    contigs = [rseq(rng, rng.randrange(4000, 12000), "ACGT" if c % 2 else "ACGTN") for c in range(5)]
    cases = []
    for _ in range(n):
        t = rng.randrange(len(contigs))
        ref = contigs[t]
        M = rng.randrange(40, 151)
        start = rng.randrange(0, len(ref) - M - 400)
        mode = rng.random()
        if mode < 0.45:
            dl, cut = rng.randrange(1, 300), rng.randrange(5, M - 5)
            read = ref[start:start + cut] + ref[start + cut + dl:start + dl + M]
        elif mode < 0.75:
            il, cut = rng.randrange(1, 40), rng.randrange(5, M - 5)
            read = (ref[start:start + cut] + rseq(rng, il) + ref[start + cut:start + M])[:M]
        elif mode < 0.9:
            read = ref[start:start + M]
        else:
            read = rseq(rng, M)
        position = max(0, min(len(ref) - 1, start + rng.randrange(-500, 500)))
        cases.append((t, position, rng.randrange(300, 800), read))
    cpath, tpath = tmp_path / "contigs.txt", tmp_path / "cand.tsv"
    cpath.write_text("".join(c + "\n" for c in contigs))
    tpath.write_text("".join(f"{t}\t{p}\t{r}\t{read}\n" for t, p, r, read in cases))
    out = subprocess.run([driver, "-k", str(k), "-g", str(g), str(cpath), str(tpath)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == n
    p = oracle.default_params(k, g)
    nsplit = 0
    for i, (t, position, range1, read) in enumerate(cases):
        o = oracle.realign_read(p, contigs[t], position, range1, read)
        st, rs, words = lines[i].split("\t")
        segs = o.segments()
        assert int(st) == o.status, (i, lines[i])
        assert [int(w) for w in words.split(",") if w] == [(ln << 4) | op for op, ln, _s, _e in segs], i
        if segs:
            assert int(rs) == segs[0][2], i
        nsplit += o.status == 6
    assert nsplit > n // 10


def test_c_driver_region_sharded_workers(driver, tmp_path):
    """-G 3: three host worker threads, one context each (devices taken round-robin; on a one-GPU box they
    share it), 2 kb regions dealt round-robin, results merged into input order: the output must be the
    single-worker output byte for byte (row (e) of the scope table, and the library's thread safety)"""
    rng = make_rng(5)
    contigs = [rseq(rng, rng.randrange(20000, 40000), "ACGT") for _ in range(3)]
    lines = []
    for _ in range(6000):
        t = rng.randrange(len(contigs))
        ref = contigs[t]
        M = rng.randrange(60, 151)
        start = rng.randrange(0, len(ref) - M - 400)
        if rng.random() < 0.6:
            dl, cut = rng.randrange(1, 200), rng.randrange(5, M - 5)
            read = ref[start:start + cut] + ref[start + cut + dl:start + dl + M]
        else:
            read = ref[start:start + M]
        lines.append(f"{t}\t{max(0, start + rng.randrange(-400, 400))}\t{rng.randrange(300, 800)}\t{read}\n")
    cpath, tpath = tmp_path / "contigs.txt", tmp_path / "cand.tsv"
    cpath.write_text("".join(c + "\n" for c in contigs))
    tpath.write_text("".join(lines))
    one = subprocess.run([driver, str(cpath), str(tpath)], capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stderr[-2000:]
    many = subprocess.run([driver, "-G", "3", "-R", "2000", str(cpath), str(tpath)], capture_output=True, text=True, timeout=600)
    assert many.returncode == 0, many.stderr[-2000:]
    assert many.stdout == one.stdout
    assert len(one.stdout.splitlines()) == len(lines)
    counts = [int(x) for x in __import__("re").findall(r"worker \d+ \(device \d+\): (\d+)", many.stderr)]
    assert len(counts) == 3 and sum(counts) == len(lines) and min(counts) > len(lines) // 6
