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
                 spacing=2000, max_indel=50, sub_rate=0.01, seed=7, contig="chrS", keep_frac=1.0, read_seed=None):
    """Writes <prefix>.fa, <prefix>.sam, <prefix>.config; returns dict(sites=..., npairs=..., nrecords=...).
    keep_frac < 1 plants only that fraction of the sites (same reference, same site list: the "normal" of a
    tumor/normal pair); read_seed draws different reads from the same genome."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    ref = acgt[rng.integers(0, 4, size=length)]
    # ---- mutated genome + reference coordinate of each of its bases
    nsites = (length - 4000) // spacing
    spos = 2000 + np.arange(nsites) * spacing + rng.integers(0, spacing // 2, size=nsites)
    slen = rng.integers(1, max_indel + 1, size=nsites)
    sdel = rng.random(nsites) < 0.5
    ins_all = [acgt[rng.integers(0, 4, size=ln)] for ln in slen]      # drawn for every site so that subsets agree
    site_kept = np.random.default_rng(seed + 1000).random(nsites) < keep_frac
    if read_seed is not None:
        rng = np.random.default_rng(read_seed)
    seq_parts, rc_parts, prev = [], [], 0
    for p, ln, isdel, ins, keep in zip(spos, slen, sdel, ins_all, site_kept):
        if not keep:
            continue
        seq_parts.append(ref[prev:p]); rc_parts.append(np.arange(prev, p, dtype=np.int64))
        if isdel:
            prev = p + ln
        else:
            seq_parts.append(ins); rc_parts.append(np.full(ln, -1, dtype=np.int64))
            prev = p
    seq_parts.append(ref[prev:]); rc_parts.append(np.arange(prev, length, dtype=np.int64))
    samp = np.concatenate(seq_parts); rcmap = np.concatenate(rc_parts)
    S, M = len(samp), read_len

    npairs = int(depth * length / (2 * M))
    fstart = rng.integers(0, S - 1000, size=npairs)
    isz = np.clip(rng.normal(insert_mean, insert_sd, size=npairs), 2 * M + 10, insert_mean + 4 * insert_sd).astype(np.int64)
    keep = fstart + isz < S
    fstart, isz = fstart[keep], isz[keep]
    npairs = len(fstart)
    first_left = rng.random(npairs) < 0.5
    records = []          # (sortpos, serial, line)
    nrec = 0
    qual = "I" * M
    cols = np.arange(M, dtype=np.int64)[None, :]
    for c0 in range(0, npairs, 50000):                      # chunks keep the [pairs, M] arrays small
        c1 = min(npairs, c0 + 50000)
        starts = np.stack([fstart[c0:c1], fstart[c0:c1] + isz[c0:c1] - M], axis=1)        # [m, 2] sample offsets
        idx = starts[:, :, None] + cols[None, :, :]                                          # [m, 2, M]
        bases = samp[idx]
        sub = rng.random(bases.shape) < sub_rate
        bases = np.where(sub, acgt[rng.integers(0, 4, size=bases.shape)], bases)
        rc_first, rc_last = rcmap[starts], rcmap[starts + M - 1]
        clean = (rc_first >= 0) & (rc_last - rc_first == M - 1)                              # no indel inside the read
        for r in range(c1 - c0):
            k = c0 + r
            name = f"p{k}"
            ends = []
            for which in (0, 1):
                if clean[r, which]:
                    ends.append(dict(which=which, bases=bases[r, which].tobytes(), pos0=int(rc_first[r, which]),
                                     ops=[(M, "M")], rev=(which == 1), unmapped=False))
                    continue
                s0 = int(starts[r, which])
                pos0, ops = _cigar_from_refcoords(rcmap[s0:s0 + M])
                e = dict(which=which, bases=bases[r, which].tobytes(), pos0=pos0, ops=ops, rev=(which == 1),
                         unmapped=ops is None)
                if not e["unmapped"]:
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
                            i0 = next(i for i, (_n, o) in enumerate(ops) if o in "ID")
                            left = sum(n for n, o in ops[:i0] if o in "MIS")
                            right = sum(n for n, o in ops[i0 + 1:] if o in "MIS") + (ops[i0][0] if ops[i0][1] == "I" else 0)
                            if left >= right:
                                e["ops"] = ops[:i0] + [(M - left, "S")]
                            else:
                                consumed_ref = sum(n for n, o in ops[:i0 + 1] if o in "MD")
                                kept = ops[i0 + 1:]
                                e["pos0"] = e["pos0"] + consumed_ref
                                e["ops"] = [(M - sum(n for n, o in kept if o in "MIS"), "S")] + kept
                        else:
                            e["unmapped"] = True
                ends.append(e)
            a, b = ends
            if a["unmapped"] and b["unmapped"]:
                continue
            for e, mate in ((a, b), (b, a)):
                flag = 0x1 | (0x40 if (e["which"] == 0) == bool(first_left[k]) else 0x80)
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
                    e_end = e["pos0"] + sum(n for n, o in e["ops"] if o in "MD")
                    m_end = mate["pos0"] + sum(n for n, o in mate["ops"] if o in "MD")
                    lo = min(e["pos0"], mate["pos0"]); hi = max(e_end, m_end)
                    tlen = (hi - lo) if e["pos0"] <= mate["pos0"] else -(hi - lo)
                line = (f"{name}\t{flag}\t{contig}\t{pos + 1}\t{mapq}\t{cigar}\t=\t{pnext + 1}\t{tlen}\t"
                        f"{e['bases'].decode()}\t{qual}\tMQ:i:60")
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
    return dict(nsites=int(site_kept.sum()), ndel=int((sdel & site_kept).sum()), nins=int((~sdel & site_kept).sum()), npairs=npairs, nrecords=nrec)
