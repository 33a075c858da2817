"""GPU parity for row f1 of SURVEY.md section 8: the known-indel support check, realign_with_indel
(variant.c:1246-1424), through indelgpu_indel_support_batch of the C ABI, against the reference's
committed outputs (tests/golden/indel_support.tsv.gz) and against the oracle on seeded cases.
Bit-exact: three integer counters per task."""
import os

import numpy as np
import pytest

from tests.util import indel_support_cases, load_indel_support_golden, make_rng, rseq

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["two-pass", "by-size"], autouse=True)
def kernel_choice(request):
    """every test runs twice: with the two-pass 16-bit kernels forced for any batch size, and with the library's own
    choice (batches under 4 096 pairs take the single-launch wavefront kernel)"""
    old = os.environ.get("INDELGPU_SUPPORT_PACK_MIN")
    if request.param == "two-pass":
        os.environ["INDELGPU_SUPPORT_PACK_MIN"] = "0"
    else:
        os.environ.pop("INDELGPU_SUPPORT_PACK_MIN", None)
    yield request.param
    if old is None:
        os.environ.pop("INDELGPU_SUPPORT_PACK_MIN", None)
    else:
        os.environ["INDELGPU_SUPPORT_PACK_MIN"] = old


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from indelminer_b200 import build
    build.build()
    import indelminer_b200
    return indelminer_b200


def split(oracle, c):
    ref, rstart, rstop, read, qstart, qstop, vtype, vstart, vstop, alt = c
    return oracle.indel_target(ref, rstart, rstop, vtype, vstart, vstop, alt), read[qstart:qstop].encode()


def test_golden_reference_outputs(gpu, oracle):
    g = load_indel_support_golden()
    pairs = [split(oracle, c) for c, _ in g]
    R = gpu.Realigner()
    out = R.indel_support_batch([t for t, _ in pairs], [q for _, q in pairs])
    for i, (_c, want) in enumerate(g):
        assert (out["subs"][i], out["indels"][i], out["aligned"][i]) == want, i
    assert out["cells"] == sum(len(t) * len(q) for t, q in pairs)
    R.close()


def test_seeded_cases_against_oracle(gpu, oracle):
    cases = indel_support_cases(make_rng(77), 5000)
    pairs = [split(oracle, c) for c in cases]
    R = gpu.Realigner()
    out = R.indel_support_batch([t for t, _ in pairs], [q for _, q in pairs])
    for i, (t, q) in enumerate(pairs):
        assert (out["subs"][i], out["indels"][i], out["aligned"][i]) == oracle.indel_support_dp(t, q), i
    R.close()


def test_edge_shapes(gpu, oracle):
    """empty and one-base sequences, nothing in common (score 0: no traceback), N and lower case,
    ragged lengths in one batch, a long pair"""
    rng = make_rng(3)
    T = [b"", b"A", b"", b"ACGT", b"AAAA", b"acgtnACGTN", rseq(rng, 700, "ACGT").encode(), b"ACGTACGTAC" * 30]
    Q = [b"", b"", b"C", b"ACGT", b"CCCC", b"ACGTNacgtn", rseq(rng, 250, "ACGT").encode(), b"ACGTACGTAC" * 12]
    long_t = rseq(rng, 3000, "ACGT")
    T.append(long_t.encode())
    Q.append((long_t[500:900] + long_t[930:1500]).encode())
    R = gpu.Realigner()
    out = R.indel_support_batch(T, Q)
    for i, (t, q) in enumerate(zip(T, Q)):
        assert (out["subs"][i], out["indels"][i], out["aligned"][i]) == oracle.indel_support_dp(t, q), i
    assert tuple(int(out[k][4]) for k in ("subs", "indels", "aligned")) == (0, 0, 1)     # the NUL column only
    with pytest.raises(gpu.IndelGpuError):
        R.indel_support_batch([b"A" * 8001], [b"A"])
    out = R.indel_support_batch([], [])
    assert len(out["subs"]) == 0
    R.close()


def test_batch_larger_than_resident_threads(gpu, oracle):
    """more tasks than threads in flight: the grid-stride loop reuses each thread's scratch"""
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    n = 200_000
    T, Q = [], []
    base = acgt[rng.integers(0, 4, size=4096)]
    for i in range(n):
        s = int(rng.integers(0, 3900))
        t = base[s:s + int(rng.integers(20, 60))]
        q = t[int(rng.integers(0, 5)):].copy()
        if len(q) > 8 and i % 3 == 0:
            q = np.delete(q, 5)
        T.append(t.tobytes())
        Q.append(q.tobytes())
    R = gpu.Realigner()
    out = R.indel_support_batch(T, Q)
    for i in range(0, n, 97):
        assert (out["subs"][i], out["indels"][i], out["aligned"][i]) == oracle.indel_support_dp(T[i], Q[i]), i
    assert out["cells"] == sum(len(t) * len(q) for t, q in zip(T, Q))
    R.close()


def test_every_columns_per_lane_variant(gpu, oracle):
    """target lengths 1..512 exercise every register-tile width of the wavefront kernel (2..16 columns per
    lane, both instantiations); queries up to the 500-base limit; lengths just past the limits take the
    thread-per-pair kernel in the same batch"""
    rng = make_rng(19)
    T, Q = [], []
    for len1 in list(range(1, 513, 5)) + [255, 256, 257, 511, 512, 513, 520]:
        t = rseq(rng, len1, "ACGT")
        lo = rng.randrange(0, max(1, len1 // 3))
        q = t[lo:lo + rng.randrange(1, 500)]
        if len(q) > 30:
            cut = rng.randrange(10, len(q) - 10)
            q = q[:cut] + (rseq(rng, rng.randrange(1, 12), "ACGT") if rng.random() < 0.5 else "") + q[cut + rng.randrange(0, 8):]
        q = "".join(rng.choice("ACGT") if rng.random() < 0.04 else ch for ch in q)[:500]
        T.append(t.encode())
        Q.append(q.encode())
    T += [rseq(rng, 300, "ACGT").encode(), rseq(rng, 512, "AC").encode()]
    Q += [rseq(rng, 501, "ACGT").encode(), rseq(rng, 500, "AC").encode()]
    for small_first in (True, False):                      # MAXCPL = 8 and 16 kernels
        sel = [i for i in range(len(T)) if (len(T[i]) <= 256) == small_first or not small_first]
        R = gpu.Realigner()
        out = R.indel_support_batch([T[i] for i in sel], [Q[i] for i in sel])
        for k, i in enumerate(sel):
            assert (out["subs"][k], out["indels"][k], out["aligned"][k]) == oracle.indel_support_dp(T[i], Q[i]), (i, len(T[i]), len(Q[i]))
        R.close()


def test_scratch_ring_and_chunked_launches(gpu, oracle):
    """enough long pairs that the direction bits of one class exceed the 2 GB a launch may use: the class is cut into
    several wavefront launches, more than the three scratch slots, so slots are reused after a merged walk launch;
    mixed with short pairs of the other classes in the same batch"""
    rng = np.random.default_rng(23)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = acgt[rng.integers(0, 4, size=1 << 16)]
    T, Q = [], []
    for i in range(66_000):
        s = int(rng.integers(0, (1 << 16) - 600))
        if i % 11 == 0:
            t = base[s:s + int(rng.integers(30, 250))]
            q = t[int(rng.integers(0, 20)):][:150].copy()
        else:
            t = base[s:s + int(rng.integers(480, 513))]
            q = np.concatenate([t[3:200], t[200 + int(rng.integers(0, 30)):]])[:500].copy()
        if len(q) > 40:
            q[int(rng.integers(0, len(q)))] = acgt[int(rng.integers(0, 4))]
        T.append(t.tobytes())
        Q.append(q.tobytes())
    R = gpu.Realigner()
    out = R.indel_support_batch(T, Q)
    assert out["launches"] >= 7                              # wavefront launches of three classes, walks after every third (66 000 pairs: two-pass either way)
    for i in list(range(0, len(T), 1009)) + [len(T) - 1, len(T) - 2]:
        assert (out["subs"][i], out["indels"][i], out["aligned"][i]) == oracle.indel_support_dp(T[i], Q[i]), i
    again = R.indel_support_batch(T, Q)                       # same context, warm buffers: identical
    for k in ("subs", "indels", "aligned"):
        assert np.array_equal(out[k], again[k])
    R.close()
