"""Differential pinning of the oracle (oracle/indel_oracle.c) against the reference's own
object code (oracle/_ref/*.so, built from /root/reference by oracle/Makefile).  Skipped where
the reference objects are absent (they are git-ignored; the GPU box receives the prebuilt files)."""
import pytest

from tests.util import make_rng, mutate, rseq, split_read_case


@pytest.fixture(scope="module")
def O(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle


def test_local_align_small(O):
    rng = make_rng(1)
    p = O.default_params()
    n = pos = 0
    for _ in range(6000):
        alpha = rng.choice(["AC", "ACGT", "ACGTN", "A"])
        N = rng.randrange(8, 120)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(1, 60)
        if rng.random() < 0.7 and N > M + 2:
            off = rng.randrange(0, N - M)
            read = mutate(rng, ref[off:off + M], alpha)
            d = off + rng.randrange(-3, 4)
        else:
            read = rseq(rng, M, alpha)
            d = rng.randrange(-M + 1, N)
        M = len(read)
        w = rng.choice([1, 1, 2, 3, 5, 8, 13, 33])
        low = d - w // 2
        up = low + w - 1
        if max(-M, low) > min(N, up):
            continue
        a = O.local_align(p, read, ref, low, up)
        b = O.ref_local_align(read, ref, low, up)
        assert a == b, (read, ref, low, up)
        n += 1
        pos += a[0] > 0
    assert pos > n // 2


def test_local_align_at_the_sizes_the_gpu_is_tested_on(O):
    """local_align's forward / reverse sweeps at the shapes of the D1 band sweep and of the GPU parity tests: bands
    65, 129, 160 (the warp max-scan rewrite, local_sweeps_warp) and 33 / 41 (the register sweeps' upper end), windows
    up to 1 500 bases, reads up to 150 -- score, end points AND the script ALIGN returns for the sub-rectangle"""
    rng = make_rng(21)
    p = O.default_params()
    pos = 0
    for it in range(700):
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "AC"])
        N = rng.randrange(300, 1500)
        ref = rseq(rng, N, alpha)
        M = rng.randrange(40, 151)
        off = rng.randrange(0, N - M - 60)
        read = mutate(rng, ref[off:off + M + 55], alpha, sub=0.01, nindel=rng.randrange(0, 3), maxindel=50)[:M]
        w = rng.choice([33, 41, 65, 129, 160])
        low = off - w // 2 + rng.randrange(-4, 5)
        up = low + w - 1
        if max(-len(read), low) > min(N, up):
            continue
        a = O.local_align(p, read, ref, low, up)
        b = O.ref_local_align(read, ref, low, up)
        assert a == b, (it, w, read, off, low, up)
        pos += a[0] > 0
    assert pos > 500


def test_global_align_dc_script(O):
    """ALIGN incl. the divide-and-conquer traceback: scores AND scripts (tie-heavy alphabets)."""
    rng = make_rng(7)
    p = O.default_params()
    for _ in range(4000):
        alpha = rng.choice(["AC", "ACGT", "ACGTN", "AAC"])
        M = rng.randrange(1, 150)
        A = rseq(rng, M, alpha)
        B = mutate(rng, A, alpha, sub=rng.choice([0, 0.02, 0.2]), nindel=rng.randrange(0, 4), maxindel=30)
        w = rng.choice([1, 2, 3, 4, 5, 9, 17, 33, 65, 129, 200])
        low = -rng.randrange(0, w)
        up = low + w - 1
        assert O.global_align(p, A, B, low, up) == O.ref_ALIGN(A, B, low, up), (A, B, low, up)


@pytest.mark.parametrize("k,g", [(6, 0), (6, 4), (4, 0), (8, 3), (2, 0), (11, 1)])
def test_find_best_band(O, k, g):
    rng = make_rng(100 + k * 16 + g)
    O.ref_set_params(k, g, 1000, 10)
    p = O.default_params(k, g)
    for _ in range(700):
        alpha = rng.choice(["AC", "ACGT", "ACGTN"])
        L = rng.randrange(60, 600)
        ref = rseq(rng, L, alpha)
        zs1 = rng.randrange(0, 20)
        e1 = rng.randrange(zs1 + 30, L)
        M = rng.randrange(1, 50) if rng.random() < 0.1 else rng.randrange(10, 120)
        if rng.random() < 0.7 and e1 - zs1 > M + 2:
            off = rng.randrange(zs1, e1 - M)
            read = mutate(rng, ref[off:off + M], alpha)
        else:
            read = rseq(rng, M, alpha)
        rd = rseq(rng, 5, alpha) + read + rseq(rng, 5, alpha)
        zs2 = rng.randrange(0, 6)
        e2 = len(rd) - rng.randrange(0, 6)
        if (e1 - zs1) + (e2 - zs2) - 2 * (k - 1) <= g:
            continue
        anchor = rng.randrange(0, L)
        assert (O.find_best_band(p, ref, zs1, e1, anchor, rd, zs2, e2)
                == O.ref_find_best_band(ref, zs1, e1, anchor, rd, zs2, e2))
    O.ref_set_params()


@pytest.mark.parametrize("k,g", [(6, 0), (6, 3), (8, 0), (5, 8), (6, 16), (6, 32), (6, 47)])
def test_realign_two_rounds(O, k, g):
    """attempt_diagonal_alignments + update_readsegs: final segment lists and evidence counts."""
    rng = make_rng(11 + k + 31 * g)
    O.ref_set_params(k, g, 1000, 10)
    p = O.default_params(k, g)
    seen = set()
    stops = 0
    for _ in range(500):
        ref, position, range1, read = split_read_case(rng)
        a = O.realign_read(p, ref, position, range1, read)
        if a.status == 7:           # numdiagonals <= numgaps in round 2 (short slice, wide -g): the reference's forceassert
            stops += 1              # (alignment.c:405) would end THIS process, so it cannot be asked
            continue
        b = O.ref_realign(ref, position, range1, read)
        assert (a.segments(), a.nevidence) == b, (k, g, position, range1, read)
        seen.add(a.status)
    assert {1, 2, 3, 4, 5, 6} <= seen and stops <= 40, stops
    O.ref_set_params()


# ---------------------------------------------------------------- row f1: realign_with_indel (variant.c:1246-1424)
def test_realign_with_indel_against_reference(oracle):
    if not oracle.have_ref_variant():
        pytest.skip("oracle/_ref/libref_variant.so not built")
    from tests.util import indel_support_cases
    n = 0
    for c in indel_support_cases(make_rng(5), 2500):
        assert oracle.realign_with_indel(*c) == oracle.ref_realign_with_indel(*c), c
        n += 1
    assert n == 2500


def test_realign_with_indel_golden(oracle):
    """the committed outputs of the reference (oracle/make_golden_support.py); also checks that the split
    the GPU entry point uses (target built on the host, DP on the slices) is the same function"""
    from tests.util import load_indel_support_golden
    g = load_indel_support_golden()
    assert len(g) == 600
    seen = set()
    for c, want in g:
        assert oracle.realign_with_indel(*c) == want
        ref, rstart, rstop, read, qstart, qstop, vtype, vstart, vstop, alt = c
        target = oracle.indel_target(ref, rstart, rstop, vtype, vstart, vstop, alt)
        assert oracle.indel_support_dp(target, read[qstart:qstop]) == want
        seen.add((vtype, want[1] > 0))
    assert len(seen) == 4            # insertions and deletions, with and without gap columns
