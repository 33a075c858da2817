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
    deps = [SRC] + [os.path.join(HERE, "..", "indelminer_b200", "csrc", f) for f in ("kernels.cuh", "band_dp.cuh")]
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
