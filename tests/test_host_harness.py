"""The serial pieces of the device source (banded sweeps, explicit-stack divide and conquer,
script -> CIGAR) compiled for the host and compared with the oracle: catches control-flow bugs
without GPU time.  The product never runs these on the CPU; this is a test build only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import make_rng, mutate, rseq

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "harness.cu")
SO = os.path.join(HERE, "host_harness", "libharness.so")


@pytest.fixture(scope="module")
def harness():
    deps = [SRC] + [os.path.join(HERE, "..", "indelminer_b200", "csrc", f) for f in ("kernels.cuh", "band_dp.cuh", "indel_support.cuh", "indel_support_pack.cuh", "task_kernels.cuh", "warp_vote.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hh_band_align.restype = C.c_int
    L.hh_band_align_interleaved.restype = C.c_int
    return L


def run(L, read, win, low, up, prm=(6, 0, 1000, 10, 1, -10, 10, 10), lane=None):
    read, win = read.encode(), win.encode()
    M, N = len(read), len(win)
    out = (C.c_int * 10)()
    cig = (C.c_uint32 * (2 * M + N + 8))()
    script = (C.c_int * (2 * M + N + 8))()
    if lane is None:
        rc = L.hh_band_align((C.c_int * 8)(*prm), read, M, win, N, low, up, out, cig, script, len(script))
    else:
        rc = L.hh_band_align_interleaved((C.c_int * 8)(*prm), read, M, win, N, low, up, out, cig, script,
                                         len(script), lane)
    assert rc == 0
    return list(out), list(cig[:out[5]]), list(script[:out[9]])


def test_serial_banded_path_matches_oracle(harness, oracle):
    rng = make_rng(4242)
    p = oracle.default_params()
    n = npos = 0
    for _ in range(4000):
        alpha = rng.choice(["AC", "ACGT", "ACGTN", "AAC"])
        N = rng.randrange(8, 400)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 150)
        if rng.random() < 0.8 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha, sub=rng.choice([0, 0.02, 0.1]),
                          nindel=rng.randrange(0, 3), maxindel=20)
            d = off + rng.randrange(-3, 4)
        else:
            read = rseq(rng, M, alpha)
            d = rng.randrange(-M + 1, N)
        M = len(read)
        w = rng.choice([2, 3, 5, 8, 17, 33, 65, 129])
        low = d - w // 2
        up = low + w - 1
        if min(N, up) - max(-M, low) + 1 < 2:
            continue
        cells = oracle.Cells()
        score, ends, script = oracle.local_align(p, read, ref, low, up, cells=cells)
        out, cig, scr = run(harness, read, ref, low, up)
        ctx = (read, ref, low, up)
        assert out[0] == score, ctx
        assert (out[6], out[7], out[8]) == (cells.fwd, cells.rev, cells.glob), ctx
        n += 1
        if score > 0:
            npos += 1
            assert tuple(out[1:5]) == ends, ctx
            assert scr == script, ctx
            (_c, exp_cig, _s) = oracle.attempt_band_alignment(p, ref, 0, N, read, 0, M, low, up)
            assert cig == exp_cig, ctx
    assert npos > n // 2


def test_interleaved_view_matches_oracle_incl_band_1(harness, oracle):
    """the thread-per-alignment kernel runs this exact source on lane-interleaved scratch (IArr<32>),
    also for bands of one diagonal"""
    rng = make_rng(777)
    p = oracle.default_params()
    npos = 0
    for it in range(1500):
        alpha = rng.choice(["AC", "ACGT", "ACGTN"])
        N = rng.randrange(8, 300)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 120)
        if rng.random() < 0.8 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha, sub=rng.choice([0, 0.02, 0.1]),
                          nindel=rng.randrange(0, 3), maxindel=12)
            d = off + rng.randrange(-2, 3)
        else:
            read = rseq(rng, M, alpha)
            d = rng.randrange(-M + 1, N)
        M = len(read)
        w = rng.choice([1, 1, 2, 5, 9, 17, 40])
        low = d - w // 2
        up = low + w - 1
        if min(N, up) - max(-M, low) + 1 < 1:
            continue
        cells = oracle.Cells()
        score, ends, script = oracle.local_align(p, read, ref, low, up, cells=cells)
        out, cig, scr = run(harness, read, ref, low, up, lane=it % 32)
        ctx = (read, ref, low, up)
        assert out[0] == score, ctx
        if score > 0:
            npos += 1
            assert tuple(out[1:5]) == ends, ctx
            assert scr == script, ctx
            (_c, exp_cig, _s) = oracle.attempt_band_alignment(p, ref, 0, N, read, 0, M, low, up)
            assert cig == exp_cig, ctx
            assert (out[6], out[7], out[8]) == (cells.fwd, cells.rev, cells.glob), ctx
    assert npos > 500


@pytest.mark.parametrize("scoring", [(1, -10, 10, 10), (1, -3, 2, 1), (2, -1, 4, 1), (1, -1, 0, 1), (3, -2, 1, 0)])
def test_unique_diagonal_shortcut_is_exact_under_other_scorings(harness, oracle, scoring):
    """align_banded_serial skips ALIGN's sweeps when the ungapped path is provably the unique optimum.
    With cheap gaps the proof condition is tight (few mismatches allowed), so near-ungapped reads over
    two- and four-letter alphabets probe both sides of it; the oracle always runs the full divide and
    conquer.  Scripts, CIGARs and the three cell counters must agree."""
    match, mismatch, G, H = scoring
    rng = make_rng(9000 + 7 * match - mismatch + 3 * G + H)
    p = oracle.default_params()
    p.match, p.mismatch, p.gapopen, p.gapextend = match, mismatch, G, H
    prm = (6, 0, 1000, 10, match, mismatch, G, H)
    n = 0
    for it in range(1200):
        alpha = rng.choice(["AC", "ACGT"])
        N = rng.randrange(20, 200)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(4, min(100, N - 2))
        off = rng.randrange(0, N - M)
        read = mutate(rng, ref[off:off + M], alpha, sub=rng.choice([0, 0.01, 0.03, 0.06]),
                      nindel=rng.choice([0, 0, 0, 1]), maxindel=3)
        M = len(read)
        w = rng.choice([2, 3, 5, 9, 17])
        low = off - rng.randrange(0, w)
        up = low + w - 1
        if min(N, up) - max(-M, low) + 1 < 1:
            continue
        cells = oracle.Cells()
        score, ends, script = oracle.local_align(p, read, ref, low, up, cells=cells)
        out, cig, scr = run(harness, read, ref, low, up, prm=prm, lane=it % 32)
        ctx = (scoring, read, ref, low, up)
        assert out[0] == score, ctx
        if score > 0:
            n += 1
            assert tuple(out[1:5]) == ends, ctx
            assert scr == script, ctx
            assert (out[6], out[7], out[8]) == (cells.fwd, cells.rev, cells.glob), ctx
    assert n > 600


def _harness_support():
    deps = [SRC] + [os.path.join(HERE, "..", "indelminer_b200", "csrc", f) for f in ("kernels.cuh", "band_dp.cuh", "indel_support.cuh", "indel_support_pack.cuh", "task_kernels.cuh", "warp_vote.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.hh_support_pack.restype = C.c_int
    return L


def test_packed_support_check_matches_oracle(oracle):
    """indel_support_pack.cuh (16-bit halves, shifted score domain, direction bits + walk back): the per-lane row step
    and the walk back are the device source compiled for the host, one segment of lanes stepped in lockstep.  Two
    pairs share every register; compared with the oracle's plain DP on reference-shaped cases and on edge shapes."""
    from tests.util import indel_support_cases
    L = _harness_support()
    rng = make_rng(91)
    pairs = []
    for c in indel_support_cases(rng, 1500):
        ref, rstart, rstop, read, qstart, qstop, vtype, vstart, vstop, alt = c
        pairs.append((oracle.indel_target(ref, rstart, rstop, vtype, vstart, vstop, alt), read[qstart:qstop].encode()))
    for _ in range(600):                                   # unrelated / repetitive / N-rich / tiny pairs
        alpha = rng.choice(["ACGT", "AC", "A", "ACGTN", "acgtACGT"])
        pairs.append((rseq(rng, rng.randrange(0, 300), alpha).encode(), rseq(rng, rng.randrange(0, 200), alpha).encode()))
    for _ in range(60):                                    # the longest shapes the kernel takes: 512 x 500, related and not
        t = rseq(rng, rng.randrange(400, 513), "ACGT")
        q = mutate(rng, t[rng.randrange(0, 12):], "ACGT", sub=0.03, nindel=rng.randrange(0, 4), maxindel=30)[:500] if rng.random() < 0.7 else rseq(rng, 500, "AC")
        pairs.append((t.encode(), q.encode()))
    pairs += [(b"", b"ACGT"), (b"ACGT", b""), (b"", b""), (b"A", b"A"), (b"A", b"C"), (b"ACGT" * 128, b"ACGT" * 50), (b"acgt" * 20, b"ACGT" * 20)]
    want = [oracle.indel_support_dp(t, q) for t, q in pairs]
    out = (C.c_int * 6)()
    n = 0
    for k in range(0, len(pairs) - 1):
        (ta, qa), (tb, qb) = pairs[k], pairs[(k * 7 + 3) % len(pairs)]
        for seg in (8, 16, 32):
            if max(len(ta), len(tb)) > 16 * seg:
                continue
            rc = L.hh_support_pack(seg, ta, len(ta), qa, len(qa), tb, len(tb), qb, len(qb), out)
            assert rc > 0
            assert tuple(out[0:3]) == want[k], (k, seg, rc, len(ta), len(qa))
            assert tuple(out[3:6]) == want[(k * 7 + 3) % len(pairs)], (k, seg, rc, "B")
            n += 1
    assert n > 4000


def test_serial_one_diagonal_alignment_matches_oracle(oracle):
    """align_diag1_serial (the thread-per-task kernel of one-diagonal bands): score, end points, cell counts and CIGAR
    against the oracle's local_align / attempt_band_alignment with low == up, under several scorings"""
    L = _harness_support()
    L.hh_diag1.restype = C.c_int
    rng = make_rng(606)
    n = npos = 0
    for it in range(6000):
        prm = rng.choice([(6, 0, 1000, 10, 1, -10, 10, 10), (6, 0, 1000, 10, 2, -1, 4, 1), (6, 0, 1000, 10, 5, -4, 0, 0), (6, 0, 1000, 10, 1, -1, 1, 1)])
        p = oracle.default_params()
        p.match, p.mismatch, p.gapopen, p.gapextend = prm[4], prm[5], prm[6], prm[7]
        alpha = rng.choice(["AC", "ACGT", "ACGTN", "A"])
        N = rng.randrange(1, 400)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 160) if it % 5 else rng.randrange(150, 390)       # more than 160 rows: the loop without the register mask
        if rng.random() < 0.8 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha, sub=rng.choice([0, 0.02, 0.1, 0.4]), nindel=0)
            d = off + rng.randrange(-1, 2)
        else:
            read = rseq(rng, M, alpha)
            d = rng.randrange(-M - 2, N + 2)
        M = len(read)
        low = up = d
        if min(N, up) - max(-M, low) + 1 != 1:
            continue
        cells = oracle.Cells()
        score, ends, _script = oracle.local_align(p, read, ref, low, up, cells=cells)
        out = (C.c_int * 9)()
        cig = (C.c_uint32 * (M + 8))()
        assert L.hh_diag1((C.c_int * 8)(*prm), read.encode(), M, ref.encode(), N, low, up, out, cig, M + 8) == 0
        ctx = (prm, read, ref, d)
        assert out[0] == score, ctx
        assert (out[6], out[7], out[8]) == (cells.fwd, cells.rev, cells.glob), ctx
        n += 1
        if score > 0:
            npos += 1
            assert tuple(out[1:5]) == ends, ctx
            (_c, exp_cig, _s) = oracle.attempt_band_alignment(p, ref, 0, N, read, 0, M, low, up)
            assert list(cig[:out[5]]) == exp_cig, ctx
    assert npos > n // 3 and n > 3000
