"""GPU parity: the CUDA path (through the C ABI of libindelgpu.so) against the CPU oracle and
against the golden vectors traced from the reference.  Bit-exact: integer scores, coordinates,
CIGAR words, segment lists."""
import numpy as np
import pytest

from tests.util import load_reference_contig, load_trace, make_rng, mutate, rseq, split_read_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from indelminer_b200 import build
    build.build()
    import indelminer_b200
    return indelminer_b200


def oracle_cigar_for_la(O, p, read, window, low, up):
    (r1, r2, q1, q2), cig, score = O.attempt_band_alignment(p, window, 0, len(window), read, 0, len(read), low, up)
    return score, (r1, r2, q1, q2), cig


# ----------------------------------------------------------------------------- golden vectors
def test_golden_local_align_calls(gpu, oracle):
    """the 1255 local_align calls the reference makes on its test_data"""
    la, _ = load_trace()
    R = gpu.Realigner()
    out = R.band_align_batch([r["read"] for r in la], [r["window"] for r in la],
                             [r["low"] for r in la], [r["up"] for r in la], want_script=True)
    p = oracle.default_params()
    for i, r in enumerate(la):
        assert out["score"][i] == r["score"], (i, r)
        q1, r1, q2, r2 = (int(x) for x in out["ends"][i])
        assert (q1, r1, q2, r2) == (r["si"], r["sj"], r["ei"], r["ej"]), (i, r)
        if r["score"] > 0:
            _s, _c, cig = oracle_cigar_for_la(oracle, p, r["read"], r["window"], r["low"], r["up"])
            assert list(out["cigar"][i][:out["ncigar"][i]]) == cig
            assert list(out["script"][i][:q2 - q1 + 1]) == [0] * (q2 - q1 + 1)
    R.close()


def test_golden_attempt_pe_alignment_calls(gpu):
    """the 697 attempt_pe_alignment calls of test_data: final segment lists and evidence"""
    _, pe = load_trace()
    contig = load_reference_contig()
    R = gpu.Realigner()
    R.set_reference([contig])
    res = R.attempt_pe_alignment_batch([r["read"] for r in pe], [r["tid"] for r in pe],
                                       [r["position"] for r in pe], [r["range1"] for r in pe], detail=True)
    assert res.launches >= 1
    nev = 0
    for i, r in enumerate(pe):
        assert res.segments(i) == r["segments"], (i, r)
        assert len(res.evidence(i)) == r["nev"]
        nev += r["nev"] > 0
    assert nev == 443
    R.close()


# ----------------------------------------------------------------------------- oracle, two rounds
@pytest.mark.parametrize("k,g", [(6, 0), (8, 0), (4, 0), (11, 0), (15, 0), (6, 3), (5, 8), (6, 32), (6, 47)])
def test_realign_vs_oracle(gpu, oracle, k, g):
    rng = make_rng(1000 + 17 * k + g)
    p = oracle.default_params(k, g)
    n = 400 if g == 0 else 150
    contigs, reads, tids, poss, rngs = [], [], [], [], []
    for c in range(8):
        L = rng.randrange(3000, 9000)
        alpha = "ACGT" if c % 3 else "ACGTN"
        contigs.append(rseq(rng, L, alpha))
    cases = []
    for i in range(n):
        t = rng.randrange(len(contigs))
        ref = contigs[t]
        _ref, position, range1, read = split_read_case(rng, L=len(ref))
        # re-plant the read into this contig
        M = len(read)
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
        if rng.random() < 0.3:
            read = "".join(rng.choice("ACGTN") if rng.random() < 0.02 else ch for ch in read)
        position = max(0, min(len(ref) - 1, start + rng.randrange(-500, 500)))
        cases.append((t, position, range1, read))
    R = gpu.Realigner(klength=k, numgaps=g)
    R.set_reference(contigs)
    res = R.attempt_pe_alignment_batch([c[3] for c in cases], [c[0] for c in cases],
                                       [c[1] for c in cases], [c[2] for c in cases], detail=True, cigars=True)
    seen = set()
    tot = [0, 0, 0]
    for i, (t, position, range1, read) in enumerate(cases):
        cells = oracle.Cells()
        o = oracle.realign_read(p, contigs[t], position, range1, read, cells=cells)
        d = res.detail[i]
        ctx = (k, g, i, t, position, range1, read)
        assert res.status[i] == o.status, ctx
        assert (d["low1"], d["up1"]) == (o.low1, o.up1), ctx
        assert (d["r1"], d["r2"], d["q1"], d["q2"], d["score1"]) == (o.r1, o.r2, o.q1, o.q2, o.score1), ctx
        assert list(res.cigar1[i][:d["n1"]]) == list(o.cigar1[:o.n1]), ctx
        if o.status in (3, 5, 6):
            assert (d["low2"], d["up2"]) == (o.low2, o.up2), ctx
            assert (d["r3"], d["r4"], d["q3"], d["q4"], d["score2"]) == (o.r3, o.r4, o.q3, o.q4, o.score2), ctx
        if o.status in (5, 6):
            assert list(res.cigar2[i][:d["n2"]]) == list(o.cigar2[:o.n2]), ctx
        assert res.segments(i) == o.segments(), ctx
        assert len(res.evidence(i)) == o.nevidence, ctx
        if o.status == 6:
            assert d["index"] == o.index, ctx
        assert (d["cells_fwd"], d["cells_rev"], d["cells_glob"]) == (cells.fwd, cells.rev, cells.glob), ctx
        tot[0] += cells.fwd; tot[1] += cells.rev; tot[2] += cells.glob
        seen.add(int(o.status))
    assert {1, 3, 4, 6} <= seen
    assert res.cells == tuple(tot)
    R.close()


# ----------------------------------------------------------------------------- kernels in isolation
@pytest.mark.parametrize("k,g", [(6, 0), (6, 4), (4, 0), (8, 3), (2, 0), (11, 1), (15, 0)])
def test_find_best_band_vs_oracle(gpu, oracle, k, g):
    rng = make_rng(500 + k * 16 + g)
    p = oracle.default_params(k, g)
    reads, wins, anchors = [], [], []
    for _ in range(600):
        alpha = rng.choice(["AC", "ACGT", "ACGTN"])
        N = rng.randrange(30, 1500)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 30) if rng.random() < 0.1 else rng.randrange(10, 160)
        if rng.random() < 0.7 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha)
        else:
            read = rseq(rng, M, alpha)
        if N + len(read) - 2 * (k - 1) <= g:
            continue
        reads.append(read); wins.append(ref)
        anchors.append(rng.choice([rng.randrange(-50, N + 50), -(10 ** 6), 10 ** 6, 0, N // 2]))
    R = gpu.Realigner(klength=k, numgaps=g)
    low, up = R.find_best_band_batch(reads, wins, anchors)
    for i in range(len(reads)):
        exp = oracle.find_best_band(p, wins[i], 0, len(wins[i]), anchors[i] & 0xFFFFFFFF, reads[i], 0, len(reads[i]))
        assert (int(low[i]), int(up[i])) == exp, (k, g, i, reads[i], wins[i], anchors[i])
    R.close()


def test_band_align_vs_oracle_all_bands(gpu, oracle):
    """local_align + ALIGN (divide-and-conquer script) + fetch_cigar for band widths 1..129"""
    rng = make_rng(77)
    p = oracle.default_params()
    reads, wins, lows, ups = [], [], [], []
    for _ in range(1500):
        alpha = rng.choice(["AC", "ACGT", "ACGTN", "AAC"])
        N = rng.randrange(8, 400)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 150)
        if rng.random() < 0.75 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha, sub=rng.choice([0, 0.02, 0.1]),
                          nindel=rng.randrange(0, 3), maxindel=20)
            d = off + rng.randrange(-3, 4)
        else:
            read = rseq(rng, M, alpha)
            d = rng.randrange(-M + 1, N)
        M = len(read)
        w = rng.choice([1, 1, 2, 3, 5, 8, 17, 33, 65, 129])
        low = d - w // 2
        up = low + w - 1
        if max(-M, low) > min(N, up):
            continue
        reads.append(read); wins.append(ref); lows.append(low); ups.append(up)
    R = gpu.Realigner()
    out = R.band_align_batch(reads, wins, lows, ups, want_script=True)
    tot = [0, 0, 0]
    npos = 0
    for i in range(len(reads)):
        cells = oracle.Cells()
        score, ends, script = oracle.local_align(p, reads[i], wins[i], lows[i], ups[i], cells=cells)
        ctx = (i, reads[i], wins[i], lows[i], ups[i])
        assert int(out["score"][i]) == score, ctx
        tot[0] += cells.fwd; tot[1] += cells.rev; tot[2] += cells.glob
        if score > 0:
            npos += 1
            q1, r1, q2, r2 = (int(x) for x in out["ends"][i])
            assert (q1, r1, q2, r2) == ends, ctx
            assert list(out["script"][i][:len(script)]) == script, ctx
            _s, _c, cig = oracle_cigar_for_la(oracle, p, reads[i], wins[i], lows[i], ups[i])
            assert list(out["cigar"][i][:out["ncigar"][i]]) == cig, ctx
    assert npos > len(reads) // 2
    assert tuple(int(x) for x in out["cells"]) == tuple(tot)
    R.close()


# ----------------------------------------------------------------------------- reference prototypes
def test_reference_prototype_symbols(gpu, oracle):
    """local_align / ALIGN / fetch_cigar called through the exported C symbols"""
    rng = make_rng(5)
    p = oracle.default_params()
    for _ in range(60):
        alpha = rng.choice(["AC", "ACGT"])
        N = rng.randrange(20, 200)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(5, 80)
        off = rng.randrange(0, max(1, N - M))
        read = mutate(rng, ref[off:off + M], alpha, nindel=rng.randrange(0, 2))
        w = rng.choice([1, 3, 9, 33])
        low = off - w // 2
        up = low + w - 1
        if max(-len(read), low) > min(N, up):
            continue
        exp = oracle.local_align(p, read, ref, low, up)
        got = gpu.local_align(read, ref, low, up)
        assert got == exp, (read, ref, low, up)
        if exp[0] > 0:
            si, sj, ei, ej = exp[1]
            A, B = read[si - 1:ei], ref[sj - 1:ej]
            lo2, up2 = low - (sj - si), up - (sj - si)
            assert gpu.ALIGN(A, B, lo2, up2) == oracle.global_align(p, A, B, lo2, up2)
            mm, words = gpu.fetch_cigar(A, B, exp[2], si, len(read))
            _s, _c, cig = oracle_cigar_for_la(oracle, p, read, ref, low, up)
            assert words == cig
            assert mm == sum(w_ >> 4 for w_ in cig if (w_ & 15) == 8)
    # degenerate ALIGN exits (globalalign.c:350-365)
    assert gpu.ALIGN("ACGT", "ACGT", 0, 0) == (4, [0, 0, 0, 0])
    assert gpu.ALIGN("ACGT", "ACTT", -1, 1) == oracle.global_align(p, "ACGT", "ACTT", -1, 1)


# ----------------------------------------------------------------------------- edge cases
def test_edge_cases(gpu, oracle):
    p = oracle.default_params()
    rng = make_rng(9)
    contig = rseq(rng, 5000)
    R = gpu.Realigner()
    R.set_reference([contig, "ACGT" * 30, "N" * 500])
    # empty batch
    res = R.attempt_pe_alignment_batch([], [], [], [])
    assert res.n == 0 and res.seg_count == 0
    cases = [
        (0, 0, 705, contig[0:100]),                       # window clipped at the contig start
        (0, 4999, 705, contig[4890:4990]),                # window clipped at the contig end
        (0, 2500, 705, "ACG"),                            # M < k (alignment.c:408-412)
        (0, 2500, 705, "ACGTA"),
        (0, 2500, 705, "N" * 100),                        # no ACGT at all
        (0, 2500, 705, contig[2400:2450] + contig[2700:2750]),   # clean 250 bp deletion
        (0, 2500, 705, contig[2400:2450] + "TTTTTTTTTT" + contig[2450:2490]),
        (1, 60, 705, "ACGT" * 10),                        # tiny repetitive contig
        (2, 250, 200, "N" * 50),                          # all-N contig: N matches N (localalign.c:61-67)
        (0, 2500, 1, contig[2480:2580]),                  # tiny range
        (0, 2500, 705, contig[2400:2420]),                # short read
        (0, 2500, 705, contig[2400:2700]),                # ragged: long read among short ones
    ]
    res = R.attempt_pe_alignment_batch([c[3] for c in cases], [c[0] for c in cases],
                                       [c[1] for c in cases], [c[2] for c in cases], detail=True)
    refs = [contig, "ACGT" * 30, "N" * 500]
    for i, (t, pos, rg, read) in enumerate(cases):
        o = oracle.realign_read(p, refs[t], pos, rg, read)
        assert res.status[i] == o.status, (i, cases[i])
        assert res.segments(i) == o.segments(), (i, cases[i])
    # inputs on which the reference itself aborts are reported, not guessed
    with pytest.raises(gpu.IndelGpuError):
        R.attempt_pe_alignment_batch(["ACGTACGTAC"], [7], [10], [705])       # unknown contig
    with pytest.raises(gpu.IndelGpuError):
        R.attempt_pe_alignment_batch(["ACGTACGTAC"], [0], [10 ** 7], [705])  # anchor beyond the contig (:550)
    R.close()
    R2 = gpu.Realigner()
    with pytest.raises(gpu.IndelGpuError):
        R2.attempt_pe_alignment_batch(["ACGT"], [0], [1], [10])              # no reference uploaded
    R2.close()
    with pytest.raises(gpu.IndelGpuError):
        gpu.Realigner(klength=16)                                            # indelminer.c:1028


def test_large_batch_properties(gpu, oracle):
    """full-size style checks that do not need the oracle on every read: determinism across
    launches / batch splits, and structural invariants of every segment list"""
    rng = np.random.default_rng(3)
    L = 2_000_000
    contig = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    n, M = 60000, 150
    starts = rng.integers(1000, L - 3000, size=n)
    reads = np.empty((n, M), dtype=np.uint8)
    dl = rng.integers(1, 50, size=n)
    cut = rng.integers(20, M - 20, size=n)
    kind = rng.random(n)
    for i in range(n):
        s = starts[i]
        if kind[i] < 0.5:
            reads[i, :cut[i]] = contig[s:s + cut[i]]
            reads[i, cut[i]:] = contig[s + cut[i] + dl[i]:s + dl[i] + M]
        elif kind[i] < 0.8:
            ins = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=dl[i])
            reads[i] = np.concatenate([contig[s:s + cut[i]], ins, contig[s + cut[i]:s + M]])[:M]
        else:
            reads[i] = contig[s:s + M]
    data = np.ascontiguousarray(reads.reshape(-1))
    off = np.arange(n + 1, dtype=np.int64) * M
    tid = np.zeros(n, dtype=np.int32)
    pos = (starts + rng.integers(-300, 300, size=n)).astype(np.int32)
    rg = np.full(n, 700, dtype=np.int32)
    R = gpu.Realigner()
    R.set_reference([contig.tobytes()])
    a = R.attempt_pe_alignment_batch(None, tid, pos, rg, packed=(data, off))
    b = R.attempt_pe_alignment_batch(None, tid, pos, rg, packed=(data, off))
    assert np.array_equal(a.status, b.status) and np.array_equal(a.nseg, b.nseg) and np.array_equal(a.rstart, b.rstart)
    assert a.seg_count == b.seg_count == int(a.nseg.sum())
    # split batches give the same per-read answers
    h = n // 3
    c = R.attempt_pe_alignment_batch(None, tid[:h], pos[:h], rg[:h], packed=(data[:h * M], off[:h + 1]))
    assert np.array_equal(c.status, a.status[:h]) and np.array_equal(c.nseg, a.nseg[:h])
    for i in list(range(0, n, 997)) + list(range(0, h, 499)):
        assert list(a.words(i)) == list(b.words(i))
        if i < h:
            assert list(a.words(i)) == list(c.words(i))
    # invariants: segments consume exactly the read; deletions found where planted
    for i in range(0, n, 211):
        segs = a.segments(i)
        if segs:
            consumed = sum(ln for op, ln, _s, _e in segs if op != 2)
            assert consumed == M
            for (op, ln, s, e), (_o2, _l2, s2, _e2) in zip(segs, segs[1:]):
                assert e == s2
    p = oracle.default_params()
    cs = contig.tobytes().decode()
    for i in range(0, n, 1500):
        o = oracle.realign_read(p, cs, int(pos[i]), 700, reads[i].tobytes().decode())
        assert a.segments(i) == o.segments()
    ndel = int(((a.status == 6)).sum())
    assert ndel > n // 4
    R.close()


def test_chunked_host_path_equals_single_launch(gpu, monkeypatch):
    """indelgpu_realign_batch cuts large batches into chunks whose copies overlap the kernels;
    the results must not depend on the chunking."""
    from indelminer_b200 import synth
    ref = synth.make_reference(400_000, seed=21)
    w = synth.make_candidates(ref, 6000, seed=22)
    R = gpu.Realigner()
    R.set_reference([ref.tobytes()])
    args = (None, w["tid"], w["position"], w["range1"])
    a = R.attempt_pe_alignment_batch(*args, packed=(w["read_bases"], w["read_off"]))
    monkeypatch.setenv("INDELGPU_CHUNK_READS", "1000")
    b = R.attempt_pe_alignment_batch(*args, packed=(w["read_bases"], w["read_off"]))
    assert b.launches == 11 and a.launches == 1         # 125 + 250 + 500 reads (ramp-up), five chunks of 850, then 500 + 250 + 125
    assert np.array_equal(a.status, b.status) and np.array_equal(a.nseg, b.nseg) and np.array_equal(a.rstart, b.rstart)
    assert a.seg_count == b.seg_count == int(a.nseg.sum())
    assert a.cells == b.cells and a.alg_bytes == b.alg_bytes
    for i in range(6000):
        assert list(a.words(i)) == list(b.words(i)), i
    R.close()


@pytest.mark.parametrize("k,g", [(6, 0), (9, 0), (6, 2)])
def test_long_reads_use_the_wide_counter_path(gpu, oracle, k, g):
    """reads longer than 255 + k bases switch the vote to 16-bit table entries / histogram counters and
    32-bit hit-list entries (HB = 16); large range1 values widen the windows beyond 65535 diagonals"""
    rng = make_rng(4000 + k + g)
    p = oracle.default_params(k, g)
    contigs = [rseq(rng, 90000, "ACGT"), rseq(rng, 60000, "ACGTN")]
    cases = []
    for i in range(120):
        t = rng.randrange(2)
        ref = contigs[t]
        M = rng.randrange(300, 1200)
        start = rng.randrange(2000, len(ref) - M - 3000)
        mode = rng.random()
        if mode < 0.5:
            dl, cut = rng.randrange(1, 600), rng.randrange(20, M - 20)
            read = ref[start:start + cut] + ref[start + cut + dl:start + dl + M]
        elif mode < 0.8:
            il, cut = rng.randrange(1, 80), rng.randrange(20, M - 20)
            read = (ref[start:start + cut] + rseq(rng, il) + ref[start + cut:start + M])[:M]
        else:
            read = ref[start:start + M]
        read = "".join(rng.choice("ACGT") if rng.random() < 0.01 else ch for ch in read)
        position = max(0, min(len(ref) - 1, start + rng.randrange(-800, 800)))
        range1 = rng.choice([700, 1500, (40000 if k == 6 else 12000) if i % 7 == 0 else 900])   # hash tables (k > 6) leave less room
        cases.append((t, position, range1, read))
    R = gpu.Realigner(klength=k, numgaps=g)
    R.set_reference(contigs)
    res = R.attempt_pe_alignment_batch([c[3] for c in cases], [c[0] for c in cases],
                                       [c[1] for c in cases], [c[2] for c in cases], detail=True)
    seen = set()
    for i, (t, position, range1, read) in enumerate(cases):
        o = oracle.realign_read(p, contigs[t], position, range1, read)
        d = res.detail[i]
        ctx = (k, g, i, t, position, range1, len(read))
        assert res.status[i] == o.status, ctx
        assert (d["low1"], d["r1"], d["r2"], d["q1"], d["q2"], d["score1"]) == (o.low1, o.r1, o.r2, o.q1, o.q2, o.score1), ctx
        assert res.segments(i) == o.segments(), ctx
        seen.add(int(o.status))
    assert 6 in seen
    R.close()


def test_synthetic_cfg3_batch_exact_vs_oracle(gpu, oracle):
    """20 000 candidates of the benchmark's own generator (deletions, insertions, clean, chimeric, all
    anchored by a mate position), every read compared with the oracle: status, start, segment words"""
    from indelminer_b200 import synth
    ref = synth.make_reference(3_000_000, seed=1, n_frac=0.0005)
    w = synth.make_candidates(ref, 20000, seed=99)
    R = gpu.Realigner()
    R.set_reference([ref.tobytes()])
    res = R.attempt_pe_alignment_batch(None, w["tid"], w["position"], w["range1"],
                                       packed=(w["read_bases"], w["read_off"]))
    p = oracle.default_params()
    cs = ref.tobytes()
    M = w["read_len"]
    hist = {}
    for i in range(20000):
        o = oracle.realign_read(p, cs, int(w["position"][i]), int(w["range1"][i]),
                                w["read_bases"][i * M:(i + 1) * M].tobytes())
        assert int(res.status[i]) == o.status, i
        assert res.segments(i) == o.segments(), i
        hist[o.status] = hist.get(o.status, 0) + 1
    assert hist.get(6, 0) > 5000 and hist.get(1, 0) > 200
    R.close()


def test_chunked_host_path_with_pinned_buffers(gpu, monkeypatch):
    """With pinned host buffers every copy of the chunked path is asynchronous: kernel ch+1 runs while chunk
    ch's results are copied back, and keeps allocating segment words.  The per-chunk segment count must be the
    one snapshot between the two kernels (capi.cu), or words allocated but not yet written reach the host.
    Many small chunks, repeated, segment words compared one by one with the unchunked pageable run."""
    import ctypes as C
    from indelminer_b200 import lib as _lib, synth
    L = _lib.load()
    ref = synth.make_reference(400_000, seed=31)
    n = 30000
    w = synth.make_candidates(ref, n, seed=32)
    R = gpu.Realigner()
    R.set_reference([ref.tobytes()])
    a = R.attempt_pe_alignment_batch(None, w["tid"], w["position"], w["range1"], packed=(w["read_bases"], w["read_off"]))
    want = [list(a.words(i)) for i in range(n)]

    def pinned(arr):
        p = L.indelgpu_host_alloc(arr.nbytes)
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.nbytes,)).view(arr.dtype).reshape(arr.shape)
        out[...] = arr
        return out, p

    cap = int(L.indelgpu_seg_bound(n, int(w["read_off"][-1])))
    bufs = {k: pinned(w[k]) for k in ("read_bases", "read_off", "tid", "position", "range1")}
    outs = {"status": pinned(np.zeros(n, np.int32)), "nseg": pinned(np.zeros(n, np.int32)), "rstart": pinned(np.zeros(n, np.int32)),
            "seg_off": pinned(np.zeros(n, np.int64)), "segs": pinned(np.zeros(cap, np.uint32))}
    hb = _lib.Batch(n, *[bufs[k][0].ctypes.data for k in ("read_bases", "read_off", "tid", "position", "range1")])
    hr = _lib.Result(outs["status"][0].ctypes.data, outs["nseg"][0].ctypes.data, outs["rstart"][0].ctypes.data,
                     outs["seg_off"][0].ctypes.data, outs["segs"][0].ctypes.data, cap, 0, None, None, None, 0)
    monkeypatch.setenv("INDELGPU_CHUNK_READS", "512")
    for rep in range(6):
        outs["segs"][0][...] = 0xFFFFFFFF
        assert L.indelgpu_realign_batch(R._ctx, C.byref(hb), C.byref(hr)) == 0, _lib.last_error()
        assert L.indelgpu_last_launch_count(R._ctx) > 50
        assert int(hr.seg_count) == a.seg_count
        assert np.array_equal(outs["status"][0], a.status) and np.array_equal(outs["nseg"][0], a.nseg)
        so, ns, sg = outs["seg_off"][0], outs["nseg"][0], outs["segs"][0]
        for i in range(n):
            assert list(sg[so[i]:so[i] + ns[i]]) == want[i], (rep, i)
    for v, p in list(bufs.values()) + list(outs.values()):
        L.indelgpu_host_free(p)
    R.close()


def test_packed_4bit_batches_equal_ascii_batches(gpu, monkeypatch):
    """indelgpu_realign_batch4: reads in the BAM's own 4-bit form (0.6 bytes per base on the wire), unpacked -- and,
    where flagged, reverse-complemented -- on the device.  Same answers as the ASCII batch, word for word: equal-length
    reads, ragged reads with odd lengths, N bases, flagged reads, one launch and the chunked host path."""
    from indelminer_b200 import api, synth
    ref = synth.make_reference(500_000, seed=41, n_frac=0.002)
    w = synth.make_candidates(ref, 9000, seed=42)
    R = gpu.Realigner()
    R.set_reference([ref.tobytes()])
    a = R.attempt_pe_alignment_batch(None, w["tid"], w["position"], w["range1"], packed=(w["read_bases"], w["read_off"]))
    rng = np.random.default_rng(43)
    rc = rng.random(9000) < 0.5
    for flags in (None, rc):
        seq4, boff, lens, fl = api.pack4(w["read_bases"], w["read_off"], flags)
        assert len(seq4) == 9000 * 75
        b = R.attempt_pe_alignment_batch4(seq4, boff, lens, fl, w["tid"], w["position"], w["range1"])
        assert np.array_equal(a.status, b.status) and np.array_equal(a.nseg, b.nseg) and np.array_equal(a.rstart, b.rstart)
        assert a.seg_count == b.seg_count and a.cells == b.cells
        for i in range(9000):
            assert list(a.words(i)) == list(b.words(i)), i
    monkeypatch.setenv("INDELGPU_CHUNK_READS", "1000")
    seq4, boff, lens, fl = api.pack4(w["read_bases"], w["read_off"], rc)
    b = R.attempt_pe_alignment_batch4(seq4, boff, lens, fl, w["tid"], w["position"], w["range1"])
    assert b.launches > 10 and np.array_equal(a.status, b.status)
    for i in range(9000):
        assert list(a.words(i)) == list(b.words(i)), i
    monkeypatch.delenv("INDELGPU_CHUNK_READS")
    # ragged: odd lengths, reads cut at random
    M = w["read_len"]
    keep = rng.integers(31, M + 1, size=3000)
    reads = [w["read_bases"][i * M:i * M + keep[i]].tobytes() for i in range(3000)]
    data, off = api.pack_sequences(reads)
    a2 = R.attempt_pe_alignment_batch(None, w["tid"][:3000], w["position"][:3000], w["range1"][:3000], packed=(data, off))
    seq4, boff, lens, fl = api.pack4(data, off, rc[:3000])
    b2 = R.attempt_pe_alignment_batch4(seq4, boff, lens, fl, w["tid"][:3000], w["position"][:3000], w["range1"][:3000])
    assert np.array_equal(a2.status, b2.status)
    for i in range(3000):
        assert list(a2.words(i)) == list(b2.words(i)), i
    # a code bit2char stops the program on (readaln.c:13-15) is an error, not an N
    bad = seq4.copy()
    bad[boff[5]] = 0x31                     # M (3) A (1)
    with pytest.raises(gpu.IndelGpuError):
        R.attempt_pe_alignment_batch4(bad, boff, lens, fl, w["tid"][:3000], w["position"][:3000], w["range1"][:3000])
    R.close()


def test_gpu_against_the_reference_objects_directly(gpu, oracle):
    """No oracle in between: the GPU against the reference's own object code (oracle/_ref/libref_dp.so,
    libref_align.so -- they travel to the GPU box as built files).  (1) local_align + ALIGN at bands 33 ... 160 on
    windows up to 1 500 bases: score, end points, divide-and-conquer script.  (2) attempt_diagonal_alignments +
    update_readsegs for wide -g (16, 32, 47): the final segment lists."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = make_rng(321)
    reads, wins, lows, ups = [], [], [], []
    for _ in range(260):
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "AC"])
        N = rng.randrange(300, 1500)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(40, 151)
        off = rng.randrange(0, N - M - 60)
        read = mutate(rng, ref[off:off + M + 55], alpha, sub=0.01, nindel=rng.randrange(0, 3), maxindel=50)[:M]
        w = rng.choice([33, 41, 65, 129, 160])
        low = off - w // 2 + rng.randrange(-4, 5)
        if max(-len(read), low) > min(N, low + w - 1):
            continue
        reads.append(read); wins.append(ref); lows.append(low); ups.append(low + w - 1)
    R = gpu.Realigner()
    out = R.band_align_batch(reads, wins, lows, ups, want_script=True)
    npos = 0
    for i in range(len(reads)):
        score, ends, script = oracle.ref_local_align(reads[i], wins[i], lows[i], ups[i])
        ctx = (i, lows[i], ups[i], len(reads[i]), len(wins[i]))
        assert int(out["score"][i]) == score, ctx
        if score > 0:
            npos += 1
            assert tuple(int(x) for x in out["ends"][i]) == ends, ctx
            assert list(out["script"][i][:len(script)]) == script, ctx
    assert npos > 200
    R.close()
    for g in (16, 32, 47):
        oracle.ref_set_params(6, g, 1000, 10)
        cases = [split_read_case(rng) for _ in range(120)]
        Rg = gpu.Realigner(numgaps=g)
        seen = set()
        aborts = 0
        for ref, position, range1, read in cases:          # one contig per case
            Rg.set_reference([ref])
            try:
                res = Rg.attempt_pe_alignment_batch([read], [0], [position], [range1])
            except gpu.IndelGpuError as e:
                # numdiagonals <= numgaps in round 2 (short slice, wide -g): the reference's forceassert (alignment.c:405)
                # would stop THIS process, so it cannot be asked; the library reports the read as rejected (status 7),
                # and the oracle names exactly these reads (ORC_ST_ASSERT)
                assert "rejected" in str(e), e
                assert oracle.realign_read(oracle.default_params(6, g), ref, position, range1, read).status == 7
                aborts += 1
                continue
            assert oracle.realign_read(oracle.default_params(6, g), ref, position, range1, read).status != 7
            want_segs, want_nev = oracle.ref_realign(ref, position, range1, read)
            assert res.segments(0) == want_segs, (g, position, range1, read)
            seen.add(int(res.status[0]))
        assert 6 in seen and aborts <= 12, (g, aborts)
        Rg.close()
    oracle.ref_set_params()
