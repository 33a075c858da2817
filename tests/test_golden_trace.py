"""Oracle vs the golden vectors traced from the reference run on its own test_data
(README.md:58-67 of the reference is the only upstream test; the trace pins the 1255
local_align and 697 attempt_pe_alignment calls behind indelminer.expected.vcf)."""
from tests.util import load_reference_contig, load_trace


def test_trace_shape():
    la, pe = load_trace()
    assert len(la) == 1255 and len(pe) == 697
    assert all(r["up"] - r["low"] + 1 == 1 for r in la)  # default -g 0: one diagonal
    assert sum(1 for r in pe if r["nev"] > 0) == 443


def test_oracle_local_align_matches_trace(oracle):
    la, _ = load_trace()
    p = oracle.default_params()
    for r in la:
        score, ends, script = oracle.local_align(p, r["read"], r["window"], r["low"], r["up"])
        assert score == r["score"], r
        assert ends == (r["si"], r["sj"], r["ei"], r["ej"]), r
        if score > 0:
            assert script == [0] * (r["ei"] - r["si"] + 1)


def test_oracle_realign_matches_trace(oracle):
    _, pe = load_trace()
    contig = load_reference_contig()
    p = oracle.default_params()
    for r in pe:
        assert r["tid"] == 0
        out = oracle.realign_read(p, contig, r["position"], r["range1"], r["read"])
        assert out.segments() == r["segments"], r
        assert out.nevidence == r["nev"], r
