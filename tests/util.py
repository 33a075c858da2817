"""Shared helpers for the test-suite: golden-vector loader and seeded generators."""
import gzip
import os
import random

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_reference_contig():
    """tests/golden/testdata_reference.fa -> upper-cased contig (shared.c:46-82 upper-cases too)."""
    seq = []
    with open(os.path.join(GOLDEN, "testdata_reference.fa")) as f:
        for line in f:
            if not line.startswith(">"):
                seq.append(line.strip().upper())
    return "".join(seq)


def load_trace():
    """Golden vectors recorded from the reference on its test_data (oracle/make_golden.sh).

    Returns (la, pe):
      la: list of dict(M,N,low,up,score,si,sj,ei,ej,read,window)      -- 1255 local_align calls
      pe: list of dict(tid,position,range0,range1,read,nev,segments)  -- 697 attempt_pe_alignment calls
    """
    la, pe = [], []
    with gzip.open(os.path.join(GOLDEN, "testdata_trace.tsv.gz"), "rt") as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            if t[0] == "LA":
                M, N, low, up, score, si, sj, ei, ej = map(int, t[1:10])
                la.append(dict(M=M, N=N, low=low, up=up, score=score, si=si, sj=sj, ei=ei, ej=ej,
                               read=t[10], window=t[11]))
            elif t[0] == "PE":
                segs = []
                if t[8] != ".":
                    segs = [tuple(map(int, s.split(":"))) for s in t[8].split(",")]
                pe.append(dict(tid=int(t[1]), position=int(t[2]), range0=int(t[3]), range1=int(t[4]),
                               read=t[5], nev=int(t[6]), nseg=int(t[7]), segments=segs))
    return la, pe


def rseq(rng, n, alpha="ACGT"):
    return "".join(rng.choice(alpha) for _ in range(n))


def mutate(rng, s, alpha, sub=0.02, nindel=1, maxindel=8):
    s = list(s)
    for i in range(len(s)):
        if rng.random() < sub:
            s[i] = rng.choice(alpha)
    for _ in range(nindel):
        if len(s) > 10:
            p = rng.randrange(1, len(s) - 1)
            ln = rng.randrange(1, maxindel)
            if rng.random() < 0.5:
                del s[p:p + ln]
            else:
                s[p:p] = list(rseq(rng, ln, alpha))
    return "".join(s) or "A"


def split_read_case(rng, L=None, alpha=None):
    """One synthetic candidate: (contig, position, range1, read) with a planted indel."""
    L = L or rng.randrange(3000, 8000)
    alpha = alpha or ("ACGT" if rng.random() < 0.8 else "ACGTN")
    ref = rseq(rng, L, alpha)
    M = rng.choice([100, 150, 36, 75])
    start = rng.randrange(0, L - M - 400)
    mode = rng.random()
    if mode < 0.45:
        dl = rng.randrange(1, 300)
        cut = rng.randrange(5, M - 5)
        read = ref[start:start + cut] + ref[start + cut + dl:start + dl + M]
    elif mode < 0.75:
        il = rng.randrange(1, 40)
        cut = rng.randrange(5, M - 5)
        read = (ref[start:start + cut] + rseq(rng, il) + ref[start + cut:start + M])[:M]
    elif mode < 0.9:
        read = ref[start:start + M]
    else:
        read = rseq(rng, M)
    read = "".join(rng.choice("ACGT") if rng.random() < 0.01 else c for c in read)
    position = max(0, min(L - 1, start + rng.randrange(-500, 500)))
    range1 = rng.choice([705, 300, 450])
    return ref, position, range1, read


def make_rng(seed):
    return random.Random(seed)


def indel_support_cases(rng, n, lower_frac=0.2):
    """Seeded (reference, rstart, rstop, read, qstart, qstop, vtype, vstart, vstop, alternate) tuples in the
    shape check_for_indel hands to realign_with_indel (variant.c:1520-1548): a known 1-40 bp insertion
    (vtype 0) or deletion (1) in VCF convention (the alternate / reference allele starts with the base at
    vstart), a read that carries it or not, 3 % substitutions, some reads in lower case, trimmed ends."""
    out = []
    while len(out) < n:
        L = rng.randrange(400, 900)
        ref = "".join(rng.choice("ACGT") for _ in range(L))
        vtype = rng.randrange(2)
        vstart = rng.randrange(120, L - 220)
        if vtype == 1:
            vstop = vstart + rng.randrange(1, 40) + 1
            alt = ref[vstart]
            size = vstop - vstart - 1
        else:
            vstop = vstart + 1
            alt = ref[vstart] + "".join(rng.choice("ACGT") for _ in range(rng.randrange(1, 40)))
            size = len(alt) - 1
        M = rng.randrange(40, 120)
        s = max(0, vstart - rng.randrange(5, M - 5))
        if rng.random() < 0.5:
            mut = ref[:vstart + 1] + (alt[1:] if vtype == 0 else "") + ref[(vstop - 1 if vtype == 1 else vstart + 1):]
            read = mut[s:s + M]
        else:
            read = ref[s:s + M]
        read = "".join(rng.choice("ACGT") if rng.random() < 0.03 else c for c in read)
        if rng.random() < lower_frac:
            read = read.lower()
        rstart = max(0, s - size)
        rstop = min(L, s + M + size + rng.randrange(0, 5))
        if not (rstart <= vstart and vstop < rstop):
            continue
        qstart = rng.randrange(0, 5)
        qstop = len(read) - rng.randrange(0, 5)
        out.append((ref, rstart, rstop, read, qstart, qstop, vtype, vstart, vstop, alt))
    return out


def load_indel_support_golden():
    """tests/golden/indel_support.tsv.gz (oracle/make_golden_support.py): the reference's own
    realign_with_indel on seeded cases -> list of (case tuple, (subs, indels, aligned))."""
    out = []
    with gzip.open(os.path.join(GOLDEN, "indel_support.tsv.gz"), "rt") as f:
        for line in f:
            t = line.rstrip("\n").split("\t")
            case = (t[0], int(t[1]), int(t[2]), t[3], int(t[4]), int(t[5]), int(t[6]), int(t[7]), int(t[8]), t[9])
            out.append((case, (int(t[10]), int(t[11]), int(t[12]))))
    return out
