"""Seeded synthetic workloads for the benchmark and the large parity tests.

D2 ("cfg3"): a chr20-sized contig of uniform ACGT, 1-50 bp indels planted every ~2 kb (50/50
insertion/deletion, homozygous), 2x150 bp pairs with insert ~ N(500, 50) and 1 % substitutions.
Only the reads that reach the hot path are generated -- the candidates fetch_func hands to
attempt_pe_alignment (indelminer.c:411,486): reads overlapping a planted indel (they carry S/I/D
in their CIGAR or are unmapped with a mapped mate), plus a few clean and chimeric ones.  Reads are
produced in reference orientation, as BAM stores them / as fetch_func's reverse-complement rule
(indelminer.c:404-409,479-484) leaves them, with the mate position as the anchor.

D1: band-sweep tasks (read, window, low, up) for the banded DP kernels.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def make_reference(length, seed=1, n_frac=0.0):
    rng = np.random.default_rng(seed)
    ref = ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]
    if n_frac > 0:
        ref[rng.random(length) < n_frac] = ord("N")
    return ref


def make_candidates(ref, n, seed=20261018, read_len=150, indel_spacing=2000, max_indel=50,
                    insert_mean=500.0, insert_sd=50.0, sub_rate=0.01, region=None, chunk=1 << 16):
    """Returns dict(read_bases uint8[n*read_len], read_off int64[n+1], tid, position, range1 int32[n],
    kind uint8[n]) for n candidate reads drawn from `region` = (lo, hi) of the contig."""
    rng = np.random.default_rng(seed)
    L = len(ref)
    lo, hi = region if region is not None else (0, L)
    lo = max(lo, 2000)
    hi = min(hi, L - 2000)
    nsites = max(1, (hi - lo) // indel_spacing)
    site_pos = lo + (np.arange(nsites, dtype=np.int64) * indel_spacing) + rng.integers(0, indel_spacing // 2, size=nsites)
    site_len = rng.integers(1, max_indel + 1, size=nsites)
    site_is_del = rng.random(nsites) < 0.5
    ins_pool = ACGT[rng.integers(0, 4, size=(nsites, max_indel), dtype=np.uint8)]
    M = read_len
    range1 = int(insert_mean + 4 * insert_sd)        # max proper insert of the (single) read group

    reads = np.empty((n, M), dtype=np.uint8)
    position = np.empty(n, dtype=np.int32)
    kind = np.empty(n, dtype=np.uint8)               # 0 deletion, 1 insertion, 2 clean, 3 chimeric tail
    cols = np.arange(M, dtype=np.int64)[None, :]
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = c1 - c0
        site = rng.integers(0, nsites, size=m)
        sp, sl, sd = site_pos[site], site_len[site], site_is_del[site]
        u = rng.random(m)
        k = np.where(u < 0.90, np.where(sd, 0, 1), np.where(u < 0.95, 2, 3)).astype(np.uint8)
        cut = rng.integers(12, M - 12, size=m)       # read offset of the breakpoint
        start = sp - cut                              # reference position of the read's first base
        # reference index of every read base: deletion skips sl bases after the cut; insertion
        # takes sl bases from the pool at the cut
        idx = start[:, None] + cols
        isdel = (k == 0)[:, None]
        isins = (k == 1)[:, None]
        after = cols >= cut[:, None]
        idx = np.where(isdel & after, idx + sl[:, None], idx)
        ins_zone = isins & after & (cols < (cut + sl)[:, None])
        idx = np.where(isins & (cols >= (cut + sl)[:, None]), idx - sl[:, None], idx)
        blk = ref[np.clip(idx, 0, L - 1)]
        if ins_zone.any():
            pool_col = np.clip(cols - cut[:, None], 0, max_indel - 1)
            blk = np.where(ins_zone, ins_pool[site][np.arange(m)[:, None], pool_col], blk)
        chim = k == 3
        if chim.any():                                # random tail of 30-60 bases
            tl = rng.integers(30, 61, size=m)
            tail = (cols >= (M - tl)[:, None]) & chim[:, None]
            blk = np.where(tail, ACGT[rng.integers(0, 4, size=(m, M), dtype=np.uint8)], blk)
        subs = rng.random((m, M)) < sub_rate
        blk = np.where(subs, ACGT[rng.integers(0, 4, size=(m, M), dtype=np.uint8)], blk)
        reads[c0:c1] = blk
        ins = np.clip(rng.normal(insert_mean, insert_sd, size=m), 2 * M - 100, range1).astype(np.int64)
        fwd = rng.random(m) < 0.5                     # mate downstream or upstream of the read
        mate = np.where(fwd, start + ins - M, start - (ins - M))
        position[c0:c1] = np.clip(mate, 0, L - 1).astype(np.int32)
        kind[c0:c1] = k
    order = np.argsort(position, kind="stable")       # BAM order: by coordinate
    reads, position, kind = reads[order], position[order], kind[order]
    return dict(read_bases=np.ascontiguousarray(reads.reshape(-1)),
                read_off=np.arange(n + 1, dtype=np.int64) * M,
                tid=np.zeros(n, dtype=np.int32), position=position,
                range1=np.full(n, range1, dtype=np.int32), kind=kind, read_len=M)


def make_band_tasks(n, band, seed=20261018, read_len=150, win_len=1410, ref_seed=1):
    """D1: n alignments of a read_len read against a win_len window on a band of `band` diagonals
    centred on the true diagonal; 1 % substitutions, half of the reads carry one 1-50 bp indel."""
    rng = np.random.default_rng(seed + band)
    L = 1 << 22
    ref = make_reference(L, seed=ref_seed, n_frac=0.001)
    M, N = read_len, win_len
    wstart = rng.integers(0, L - N - 200, size=n).astype(np.int32)
    off = rng.integers(0, N - M - 60, size=n).astype(np.int32)         # true offset of the read in its window
    has = rng.random(n) < 0.5
    isdel = rng.random(n) < 0.5
    sl = rng.integers(1, 51, size=n).astype(np.int32)
    cut = rng.integers(12, M - 12, size=n).astype(np.int32)
    reads = np.empty((n, M), dtype=np.uint8)
    wins = np.empty((n, N), dtype=np.uint8)
    cols = np.arange(M, dtype=np.int32)[None, :]
    wcols = np.arange(N, dtype=np.int32)[None, :]
    for c0 in range(0, n, 1 << 15):                                     # chunks keep the index arrays small at n = 2^20
        c = slice(c0, min(n, c0 + (1 << 15)))
        m = c.stop - c.start
        idx = (wstart[c] + off[c])[:, None] + cols
        after = cols >= cut[c][:, None]
        idx = np.where((has[c] & isdel[c])[:, None] & after, idx + sl[c][:, None], idx)
        insz = (has[c] & ~isdel[c])[:, None] & after & (cols < (cut[c] + sl[c])[:, None])
        idx = np.where((has[c] & ~isdel[c])[:, None] & (cols >= (cut[c] + sl[c])[:, None]), idx - sl[c][:, None], idx)
        blk = ref[np.clip(idx, 0, L - 1)]
        blk = np.where(insz, ACGT[rng.integers(0, 4, size=(m, M), dtype=np.uint8)], blk)
        blk = np.where(rng.random((m, M), dtype=np.float32) < 0.01, ACGT[rng.integers(0, 4, size=(m, M), dtype=np.uint8)], blk)
        reads[c] = blk
        wins[c] = ref[wstart[c][:, None] + wcols]
    low = (off - band // 2).astype(np.int32)
    up = (low + band - 1).astype(np.int32)
    return dict(reads=np.ascontiguousarray(reads.reshape(-1)), read_off=np.arange(n + 1, dtype=np.int64) * M,
                wins=np.ascontiguousarray(wins.reshape(-1)), win_off=np.arange(n + 1, dtype=np.int64) * N,
                low=low, up=up, true_off=off)


def make_support_tasks(n, seed=20261018, read_len=150, max_indel=50, ref_len=1 << 22, sub_rate=0.01):
    """Row f1 (annotate mode): n (target, query) pairs in the shape check_for_indel builds them
    (variant.c:1520-1548): a known 1-50 bp insertion or deletion, a 150 bp read overlapping it -- half of
    the reads carry the variant, half are reference reads -- and the target = the reference interval
    [read start - size, read end + size) with the variant spliced in (variant.c:1259-1275)."""
    rng = np.random.default_rng(seed)
    ref = make_reference(ref_len, seed=3)
    M = read_len
    vstart = rng.integers(1000, ref_len - 2000, size=n)
    size = rng.integers(1, max_indel + 1, size=n)
    isdel = rng.random(n) < 0.5
    carries = rng.random(n) < 0.5
    back = rng.integers(5, M - 5, size=n)              # read start = vstart - back
    ins_bases = ACGT[rng.integers(0, 4, size=(n, max_indel), dtype=np.uint8)]
    subs = rng.random((n, M)) < sub_rate
    sub_bases = ACGT[rng.integers(0, 4, size=(n, M), dtype=np.uint8)]
    targets, queries = [], []
    for k in range(n):
        v, sz, s = int(vstart[k]), int(size[k]), int(vstart[k] - back[k])
        if isdel[k]:
            mut = np.concatenate([ref[s - sz:v + 1], ref[v + 1 + sz:s + M + 2 * sz + 1]])
        else:
            mut = np.concatenate([ref[s - sz:v + 1], ins_bases[k, :sz], ref[v + 1:s + M + sz]])
        targets.append(mut)
        q = (mut[sz:sz + M] if carries[k] else ref[s:s + M]).copy()
        q[subs[k]] = sub_bases[k][subs[k]]
        queries.append(q)
    toff = np.zeros(n + 1, dtype=np.int64)
    toff[1:] = np.cumsum([len(t) for t in targets])
    qoff = np.arange(n + 1, dtype=np.int64) * M
    return dict(targets=np.ascontiguousarray(np.concatenate(targets)), target_off=toff,
                queries=np.ascontiguousarray(np.concatenate(queries)), query_off=qoff)
