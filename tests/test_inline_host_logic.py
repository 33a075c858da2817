"""Host logic of the in-process batched driver (host/indelgpu_inline.c,
host/indelgpu_bamcache.c, the two-pass loop of host/indelgpu_support.c), checked WITHOUT a GPU.

oracle/_ref/indelminer_fakegpu[_annot] is the reference program + this repo's C glue, linked against
tests/fake_gpu/fake_indelgpu.c -- the C ABI of include/indelgpu.h answered by the CPU oracle -- instead of
libindelgpu.so (oracle/Makefile target `fakeprog`; test infrastructure, never shipped).  What is under
test is everything around the library call: the producer / consumer blocks, the candidate
classification from flags and CIGAR, the 4-bit decode + reverse complement, the prefetch-cache key, the
learning of range[1], READCHUNK-independent ordering, the cached BAM handle.  The VCF must be
byte-identical to the unmodified reference's.  The same programs run against the real library in the
`-m gpu` tests (tests/test_e2e_vcf.py, tests/test_e2e_annotate.py, tests/test_e2e_configs.py)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def need(*exes):
    for e in exes:
        if not os.path.exists(os.path.join(REFDIR, e)):
            pytest.skip(f"oracle/_ref/{e} not built (needs /root/reference: make -C oracle fakeprog refprog tools)")


def run(exe, args, cwd, env=None):
    r = subprocess.run([os.path.join(REFDIR, exe)] + args, cwd=cwd, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout, r.stderr


@pytest.mark.parametrize("args,golden", [
    (["-i", "indelminer.config"], "testdata_refrun.vcf"),
    ([], "testdata_refrun_noconfig.vcf"),                      # BASELINE config 1 as written: IL / RC estimated first
    (["-g", "4", "-i", "indelminer.config"], "testdata_refrun_g4.vcf"),
])
@pytest.mark.parametrize("block", ["64", "65536"])
def test_inline_mode_vcf_identical_on_test_data(args, golden, block):
    need("indelminer_fakegpu")
    out, err = run("indelminer_fakegpu", args + ["testdata_reference.fa", "sample=alignments.bam"], GOLD,
                   dict(INDELGPU_MODE="inline", INDELGPU_INLINE_BLOCK=block))
    with open(os.path.join(GOLD, golden)) as f:
        assert out == f.read()
    m = re.search(r"inline mode: (\d+) BAM records, (\d+) calls answered from (\d+) prefetched batches \((\d+) reads realigned in them\), (\d+) computed per read", err)
    assert m, err[-500:]
    records, hits, batches, prefetched, direct = map(int, m.groups())
    assert hits + direct == 697                  # every attempt_pe_alignment call of the reference's own run (golden trace)
    assert hits > 400 and batches >= 1
    assert prefetched == hits                    # the producer foresaw exactly the calls that were made: no wasted GPU work
    # row f3: calculate_cov_params' per-variant region fetches are answered from the records still in memory
    m = re.search(r"(\d+) per-variant region fetches served from the retained records, (\d+) from the BAM", err)
    assert m and int(m.group(1)) + int(m.group(2)) >= 10
    if block == "65536":
        assert int(m.group(2)) == 0                 # with 64-record blocks the four kept blocks rarely cover a region: samtools' fetch then


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    need("indelminer_ref", "indelminer_fakegpu", "indelminer_fakegpu_annot", "synth_bam")
    d = str(tmp_path_factory.mktemp("tn"))
    gen = os.path.join(REFDIR, "synth_bam")
    subprocess.check_call([gen, "tumor", "--length", "200000", "--depth", "20", "--seed", "11"], cwd=d, stdout=subprocess.DEVNULL)
    subprocess.check_call([gen, "normal", "--length", "200000", "--depth", "20", "--seed", "11", "--keep", "0.5", "--readseed", "99"],
                          cwd=d, stdout=subprocess.DEVNULL)
    vcf, _ = run("indelminer_ref", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], d)
    with open(os.path.join(d, "tumor.vcf"), "w") as f:
        f.write(vcf)
    return d


def test_synth_bam_is_a_pure_function_of_its_arguments(pair, tmp_path):
    """tests/golden/cfg3_reference.json compares a VCF made in one container with a run in another: the
    generator must write the same bytes every time"""
    gen = os.path.join(REFDIR, "synth_bam")
    subprocess.check_call([gen, "tumor", "--length", "200000", "--depth", "20", "--seed", "11"], cwd=str(tmp_path), stdout=subprocess.DEVNULL)
    for ext in (".bam", ".bam.bai", ".fa", ".config"):
        with open(os.path.join(pair, "tumor" + ext), "rb") as a, open(os.path.join(str(tmp_path), "tumor" + ext), "rb") as b:
            assert a.read() == b.read(), ext


def test_inline_mode_on_synthetic_insertions_and_deletions(pair):
    want, _ = run("indelminer_ref", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], pair)
    body = [ln for ln in want.splitlines() if not ln.startswith("#")]
    assert len(body) > 80 and any(len(ln.split("\t")[4]) > 1 for ln in body) and any(len(ln.split("\t")[3]) > 1 for ln in body)
    for env in (dict(INDELGPU_MODE="inline"), dict(INDELGPU_MODE="inline", INDELGPU_NO_BAM_CACHE="1"), {}):
        out, _ = run("indelminer_fakegpu", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], pair, env)
        assert out == want, env


def test_inline_annotate_mode_two_pass_support(pair):
    """BASELINE config 4's command line on a small tumor / normal pair: attempt_pe_alignment prefetched per
    block, realign_with_indel batched per known variant by showing check_for_indel the records twice"""
    args = ["-q", "0", "-a", "-e", "1", "-i", "normal.config", "normal.fa", "tumor.vcf", "normal=normal.bam"]
    want, _ = run("indelminer_ref", args, pair)
    tagged = sum(ln.endswith(";normal") for ln in want.splitlines())
    assert tagged > 20
    out, err = run("indelminer_fakegpu_annot", args, pair, dict(INDELGPU_MODE="inline"))
    assert out == want
    m = re.search(r"(\d+) known variants checked, (\d+) \(variant, read\) pairs scored in (\d+) batches", err)
    assert m and int(m.group(2)) > 100 and int(m.group(3)) <= int(m.group(1))
    direct, _ = run("indelminer_fakegpu_annot", args, pair)
    assert direct == want


def test_stable_sort_flavour_is_libc_independent_and_agrees_here(pair):
    """row f4: slsort bound to a stable merge sort (host/indelgpu_stablesort.c, -Dqsort=... on the unchanged slinklist.c).
    On this C library its VCFs equal the default build's -- test_data (incl. the BF token that differs from upstream's
    golden file) and a synthetic set with many equal-position evidence records"""
    need("indelminer_fakegpu_stable")
    out, _ = run("indelminer_fakegpu_stable", ["-i", "indelminer.config", "testdata_reference.fa", "sample=alignments.bam"], GOLD,
                 dict(INDELGPU_MODE="inline"))
    with open(os.path.join(GOLD, "testdata_refrun.vcf")) as f:
        assert out == f.read()
    want, _ = run("indelminer_ref", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], pair)
    out, _ = run("indelminer_fakegpu_stable", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], pair, dict(INDELGPU_MODE="inline"))
    assert out == want


def test_inline_mode_learns_the_range_of_every_read_group(tmp_path):
    """three read groups with different insert ranges (RG:Z tags, one IL line each): range[1] lives in a hashtable
    private to indelminer.c, so the prefetcher learns it per group from the first call that passes it; after that the
    calls of all three groups are answered from the prefetched batches"""
    need("indelminer_ref", "indelminer_fakegpu", "synth_bam")
    d = str(tmp_path)
    subprocess.check_call([os.path.join(REFDIR, "synth_bam"), "c", "--length", "300000", "--depth", "20", "--seed", "12", "--rg", "3"],
                          cwd=d, stdout=subprocess.DEVNULL)
    with open(os.path.join(d, "c.config")) as f:
        assert sum(ln.startswith("IL rg") for ln in f) == 3
    want, _ = run("indelminer_ref", ["-i", "c.config", "c.fa", "s=c.bam"], d)
    out, err = run("indelminer_fakegpu", ["-i", "c.config", "c.fa", "s=c.bam"], d, dict(INDELGPU_MODE="inline"))
    assert out == want
    m = re.search(r"(\d+) calls answered from (\d+) prefetched batches \((\d+) reads realigned in them\), (\d+) computed per read", err)
    hits, _b, prefetched, direct = map(int, m.groups())
    assert hits > 10 * direct and prefetched == hits


def test_paired_end_evidence_and_the_indexed_graph(tmp_path):
    """1.5-3 kb deletions: pairs spanning them are flagged improper and become PAIRED_READ evidence (indelminer.c:516-612),
    the reads crossing them go through attempt_pe_alignment.  Exercises the indexed add_node (host/indelgpu_graph.c, row f4)
    on both kinds of node, the mate look-up's nested bam_fetch, and per-variant fetches wider than a read.  The VCF must
    equal the unmodified reference's with the index on and off."""
    need("indelminer_ref", "indelminer_fakegpu", "synth_bam")
    d = str(tmp_path)
    subprocess.check_call([os.path.join(REFDIR, "synth_bam"), "pe", "--length", "600000", "--depth", "25", "--seed", "13", "--bigdel", "12"],
                          cwd=d, stdout=subprocess.DEVNULL)
    want, _ = run("indelminer_ref", ["-i", "pe.config", "pe.fa", "s=pe.bam"], d)
    assert sum("PAIRED_READ" in ln for ln in want.splitlines() if not ln.startswith("#")) >= 4
    for env in (dict(INDELGPU_MODE="inline"), dict(INDELGPU_MODE="inline", INDELGPU_NO_GRAPH_INDEX="1"), dict(INDELGPU_NO_GRAPH_INDEX="1"), {}):
        out, _ = run("indelminer_fakegpu", ["-i", "pe.config", "pe.fa", "s=pe.bam"], d, env)
        assert out == want, env
    # region runs: the mate of an improper pair may lie outside the region (find_mate_rln, indelminer.c:255-280)
    for reg in ("chr1:1-300000", "chr1:300001-600000"):
        want_r, _ = run("indelminer_ref", ["-i", "pe.config", "-c", reg, "pe.fa", "s=pe.bam"], d)
        out_r, _ = run("indelminer_fakegpu", ["-i", "pe.config", "-c", reg, "pe.fa", "s=pe.bam"], d, dict(INDELGPU_MODE="inline"))
        assert out_r == want_r, reg
