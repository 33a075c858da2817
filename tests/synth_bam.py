"""Synthetic paired-end SAM for end-to-end runs of the indelMINER program ("D2" at a size a
test can afford): one contig of uniform ACGT, homozygous 1-50 bp insertions and deletions planted
every ~2 kb, 2x150 bp pairs drawn from the mutated genome with 1 % substitutions.  There is no read
aligner in the image, so every read gets the record an aligner could have produced: reads that do not
touch an indel are plain 150M; reads across an indel get, with fixed probabilities, the true I/D CIGAR,
a soft clip at the indel, or flag 0x4 with the mate mapped.  MAPQ 60, MQ:i:60, no RG tag (read group
"generic", indelminer.c:370).  Records are written in coordinate order, as bam_fetch needs them.

Test infrastructure only; the conversion to BAM is oracle/_ref/sam2bam (bundled samtools API).
"""
import numpy as np

COMP = bytes.maketrans(b"ACGTN", b"TGCAN")


def _cigar_from_refcoords(rc):
    """rc: reference coordinate of every read base (-1 = inserted base).  Returns (pos0, ops) with ops a
    list of (len, op) over M / I / D, or (None, None) if no base maps."""
    mapped = np.nonzero(rc >= 0)[0]
    if len(mapped) == 0:
        return None, None
    ops = []

    def push(n, op):
        if n <= 0:
            return
        if ops and ops[-1][1] == op:
            ops[-1] = (ops[-1][0] + n, op)
        else:
            ops.append((n, op))

    prev = None
    for i in range(len(rc)):
        if rc[i] < 0:
            push(1, "I")
        else:
            if prev is not None and rc[i] > prev + 1:
                push(int(rc[i] - prev - 1), "D")
            push(1, "M")
            prev = rc[i]
    return int(rc[mapped[0]]), ops


def make_dataset(path_prefix, length=1_000_000, depth=20, read_len=150, insert_mean=500.0, insert_sd=50.0,
                 spacing=2000, max_indel=50, sub_rate=0.01, seed=7, contig="chrS"):
    """Writes <prefix>.fa, <prefix>.sam, <prefix>.config; returns dict(sites=..., npairs=..., nrecords=...)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = acgt[rng.integers(0, 4, size=length)]
    # ---- mutated genome + reference coordinate of each of its bases
    nsites = (length - 4000) // spacing
    spos = 2000 + np.arange(nsites) * spacing + rng.integers(0, spacing // 2, size=nsites)
    slen = rng.integers(1, max_indel + 1, size=nsites)
    sdel = rng.random(nsites) < 0.5
    seq_parts, rc_parts, prev = [], [], 0
    for p, ln, isdel in zip(spos, slen, sdel):
        seq_parts.append(ref[prev:p]); rc_parts.append(np.arange(prev, p, dtype=np.int64))
        if isdel:
            prev = p + ln
        else:
            seq_parts.append(acgt[rng.integers(0, 4, size=ln)]); rc_parts.append(np.full(ln, -1, dtype=np.int64))
            prev = p
    seq_parts.append(ref[prev:]); rc_parts.append(np.arange(prev, length, dtype=np.int64))
    samp = np.concatenate(seq_parts); rcmap = np.concatenate(rc_parts)
    S, M = len(samp), read_len

    npairs = int(depth * length / (2 * M))
    fstart = rng.integers(0, S - 1000, size=npairs)
    isz = np.clip(rng.normal(insert_mean, insert_sd, size=npairs), 2 * M + 10, insert_mean + 4 * insert_sd).astype(np.int64)
    records = []          # (sortpos, line)
    nrec = 0
    for k in range(npairs):
        f, ins = int(fstart[k]), int(isz[k])
        if f + ins >= S:
            continue
        name = f"p{k}"
        ends = []
        for which, s in ((0, f), (1, f + ins - M)):
            bases = samp[s:s + M].copy()
            sub = rng.random(M) < sub_rate
            if sub.any():
                bases[sub] = acgt[rng.integers(0, 4, size=int(sub.sum()))]
            rc = rcmap[s:s + M]
            pos0, ops = _cigar_from_refcoords(rc)
            ends.append(dict(which=which, bases=bases.tobytes(), pos0=pos0, ops=ops, rev=(which == 1)))
        first_is_left = rng.random() < 0.5
        # representation of reads that touch an indel
        for e in ends:
            e["unmapped"] = e["ops"] is None
            if e["unmapped"]:
                continue
            ops = e["ops"]
            # inserted bases at a read end are what an aligner soft-clips
            if ops[0][1] == "I":
                ops[0] = (ops[0][0], "S")
            if ops[-1][1] == "I":
                ops[-1] = (ops[-1][0], "S")
            if any(o in "ID" for _n, o in ops):
                u = rng.random()
                if u < 0.4:
                    pass                                   # the aligner found the indel
                elif u < 0.8:                              # soft clip at the (first) indel, keeping the longer side
                    idx = next(i for i, (_n, o) in enumerate(ops) if o in "ID")
                    left = sum(n for n, o in ops[:idx] if o in "MIS")
                    right = sum(n for n, o in ops[idx + 1:] if o in "MIS") + (ops[idx][0] if ops[idx][1] == "I" else 0)
                    if left >= right:
                        e["ops"] = ops[:idx] + [(M - left, "S")]
                    else:
                        consumed_ref = sum(n for n, o in ops[:idx + 1] if o in "MD")
                        keep = ops[idx + 1:]
                        e["pos0"] = e["pos0"] + consumed_ref
                        e["ops"] = [(M - sum(n for n, o in keep if o in "MIS"), "S")] + keep
                else:
                    e["unmapped"] = True
        a, b = ends
        if a["unmapped"] and b["unmapped"]:
            continue
        for e, mate in ((a, b), (b, a)):
            flag = 0x1 | (0x40 if (e["which"] == 0) == first_is_left else 0x80)
            if e["rev"]:
                flag |= 0x10
            if mate["rev"]:
                flag |= 0x20
            if e["unmapped"]:
                flag |= 0x4
            if mate["unmapped"]:
                flag |= 0x8
            if not e["unmapped"] and not mate["unmapped"]:
                flag |= 0x2
            if e["unmapped"]:
                pos, cigar, mapq = mate["pos0"], "*", 0
            else:
                pos, cigar, mapq = e["pos0"], "".join(f"{n}{o}" for n, o in e["ops"]), 60
            pnext = mate["pos0"] if not mate["unmapped"] else pos
            tlen = 0
            if not e["unmapped"] and not mate["unmapped"]:
                def ref_end(x):
                    return x["pos0"] + sum(n for n, o in x["ops"] if o in "MD")
                lo = min(e["pos0"], mate["pos0"]); hi = max(ref_end(e), ref_end(mate))
                tlen = (hi - lo) if e["pos0"] <= mate["pos0"] else -(hi - lo)
            seq = e["bases"].decode()
            line = (f"{name}\t{flag}\t{contig}\t{pos + 1}\t{mapq}\t{cigar}\t=\t{pnext + 1}\t{tlen}\t{seq}\t{'I' * M}\tMQ:i:60")
            records.append((pos, nrec, line))
            nrec += 1
    records.sort()
    with open(path_prefix + ".fa", "w") as f:
        f.write(f">{contig}\n")
        s = ref.tobytes().decode()
        for i in range(0, length, 60):
            f.write(s[i:i + 60] + "\n")
    with open(path_prefix + ".sam", "w") as f:
        f.write(f"@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:{contig}\tLN:{length}\n")
        for _p, _i, line in records:
            f.write(line + "\n")
    with open(path_prefix + ".config", "w") as f:
        f.write(f"IL generic {int(insert_mean - 6 * insert_sd)} {int(insert_mean + 4 * insert_sd)}\nRC {contig} {depth}\n")
    return dict(nsites=int(nsites), ndel=int(sdel.sum()), nins=int((~sdel).sum()), npairs=npairs, nrecords=nrec)
