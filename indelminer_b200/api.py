"""Host-side mirror of the reference's interface for the realignment path, on top of the
C ABI of libindelgpu.so (include/indelgpu.h).

Names and argument meaning follow ratan-lab/indelMINER:
  attempt_pe_alignment   src/alignment.h:21-25      -> Realigner.attempt_pe_alignment[_batch]
  find_best_band         src/alignment.c:393-447    -> Realigner.find_best_band[_batch]
  attempt_band_alignment src/alignment.c:343-391    -> Realigner.attempt_band_alignment[_batch]
  local_align            src/localalign.h:15-25     -> local_align
  ALIGN                  src/globalalign.h:19-28    -> ALIGN
  fetch_cigar            src/globalalign.h:39-48    -> fetch_cigar
Every result is computed by the CUDA kernels; there is no CPU implementation here.
"""
import ctypes as C

import numpy as np

from . import lib as _lib

BAM_CINS, BAM_CDEL, BAM_CSOFT_CLIP, BAM_CPMATCH, BAM_CMMATCH = 1, 2, 4, 7, 8
ST_UNALIGNED, ST_WHOLE, ST_SHORT, ST_R2FAIL, ST_NOBRANCH, ST_NOCOMBINE, ST_SPLIT, ST_ASSERT = range(8)
INSERTION, DELETION = 0, 1          # varianttype, src/evidence.h:14-18


class IndelGpuError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise IndelGpuError(f"libindelgpu error {rc}: {_lib.last_error()}")


def _as_bytes(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode()


def pack_sequences(seqs):
    """list of str/bytes -> (uint8 concatenation, int64 offsets[n+1])."""
    bs = [_as_bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    data = np.frombuffer(b"".join(bs), dtype=np.uint8) if bs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(data), off


_NT16 = np.zeros(256, dtype=np.uint8)            # bam_nt16_table restricted to what bit2char accepts (readaln.c:4-17)
for _ch, _code in ((b"A", 1), (b"C", 2), (b"G", 4), (b"T", 8), (b"N", 15)):
    _NT16[_ch[0]] = _code
_COMP = np.arange(256, dtype=np.uint8)
for _a, _b in ((b"A", b"T"), (b"C", b"G"), (b"G", b"C"), (b"T", b"A")):
    _COMP[_a[0]] = _b[0]


def pack4(read_bases, read_off, revcomp=None):
    """ASCII reads -> the BAM's 4-bit form (seq4, byte_off, len, flags) for attempt_pe_alignment_batch4.  With
    `revcomp` (bool per read) a read is stored as the BAM would hold its reverse complement and flagged, so that the
    device undoes it: the alignment input is the same ASCII read either way.  Equal-length reads are vectorised."""
    n = len(read_off) - 1
    lens = (read_off[1:] - read_off[:-1]).astype(np.int32)
    nb = (lens.astype(np.int64) + 1) // 2
    byte_off = np.zeros(n + 1, dtype=np.int64)
    byte_off[1:] = np.cumsum(nb)
    flags = np.zeros(n, dtype=np.uint8) if revcomp is None else np.ascontiguousarray(revcomp, dtype=np.uint8)
    seq4 = np.zeros(int(byte_off[-1]), dtype=np.uint8)
    if n and (lens == lens[0]).all() and int(read_off[0]) == 0:
        M = int(lens[0])
        mat = read_bases[:n * M].reshape(n, M)
        if revcomp is not None:
            mat = np.where(flags[:, None].astype(bool), _COMP[mat[:, ::-1]], mat)
        codes = _NT16[mat]
        if M % 2:
            codes = np.concatenate([codes, np.zeros((n, 1), dtype=np.uint8)], axis=1)
        seq4[:] = ((codes[:, 0::2] << 4) | codes[:, 1::2]).reshape(-1)
    else:
        for i in range(n):
            r = read_bases[read_off[i]:read_off[i + 1]]
            if flags[i]:
                r = _COMP[r[::-1]]
            codes = _NT16[r]
            if len(codes) % 2:
                codes = np.concatenate([codes, np.zeros(1, dtype=np.uint8)])
            seq4[byte_off[i]:byte_off[i + 1]] = (codes[0::2] << 4) | codes[1::2]
    return seq4, byte_off, lens, flags


def walk_segments(rstart, words):
    """new_readseg's coordinate bookkeeping (readaln.c:24-99): words -> [(op, len, start, end)]."""
    out = []
    refindx = int(rstart)
    for w in words:
        op, ln = int(w) & 15, int(w) >> 4
        start = refindx
        if op in (BAM_CPMATCH, BAM_CMMATCH, BAM_CDEL):
            refindx += ln
        out.append((op, ln, start, refindx))
    return out


class BatchResult:
    """Result of Realigner.attempt_pe_alignment_batch (SoA, numpy)."""

    def __init__(self, n, seg_capacity, detail, cigar_stride):
        self.n = n
        self.status = np.zeros(n, dtype=np.int32)
        self.nseg = np.zeros(n, dtype=np.int32)
        self.rstart = np.zeros(n, dtype=np.int32)
        self.seg_off = np.zeros(n, dtype=np.int64)
        self.segs = np.zeros(max(seg_capacity, 1), dtype=np.uint32)
        self.seg_count = 0
        self.detail = np.zeros(n, dtype=np.dtype([(f, np.int32) for f, _ in _lib.Detail._fields_])) if detail else None
        self.cigar_stride = cigar_stride
        self.cigar1 = np.zeros((n, cigar_stride), dtype=np.uint32) if cigar_stride else None
        self.cigar2 = np.zeros((n, cigar_stride), dtype=np.uint32) if cigar_stride else None
        self.cells = (0, 0, 0)
        self.alg_bytes = 0
        self.launches = 0

    def words(self, i):
        o = int(self.seg_off[i])
        return self.segs[o:o + int(self.nseg[i])]

    def segments(self, i):
        """[(op, oplen, start, end)] of the readseg list update_readsegs builds (readaln.c:348-458)."""
        return walk_segments(self.rstart[i], self.words(i))

    def evidence(self, i):
        """[(variantclass, b1, b2)] in the reference's list order: add_evidence_from_segment
        (alignment.c:449-476) prepends, so the last D/I segment comes first."""
        ev = []
        for op, _ln, start, end in self.segments(i):
            if op == BAM_CDEL:
                ev.insert(0, (DELETION, start, end))
            elif op == BAM_CINS:
                ev.insert(0, (INSERTION, start, end))
        return ev


class Realigner:
    """One GPU context (indelgpu_ctx): the globals of alignment.c:3-9 plus the resident reference."""

    def __init__(self, device=0, klength=6, numgaps=0, maxdelsize=1000, ethreshold=10,
                 match=1, mismatch=-10, gapopen=10, gapextend=10):
        self._L = _lib.load()
        p = _lib.Params(klength, numgaps, maxdelsize, ethreshold, match, mismatch, gapopen, gapextend)
        self.params = p
        self._ctx = self._L.indelgpu_create(device, C.byref(p))
        if not self._ctx:
            raise IndelGpuError(f"indelgpu_create failed: {_lib.last_error()}")
        self.contig_lengths = []

    def close(self):
        if getattr(self, "_ctx", None):
            self._L.indelgpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def sm_count(self):
        return self._L.indelgpu_sm_count(self._ctx)

    # ------------------------------------------------------------------ reference
    def set_reference(self, sequences):
        """`char** sequences` of attempt_pe_alignment: one upper-cased string per contig."""
        # numpy uint8 arrays are passed by address (a 3.1 Gb genome is not copied into bytes objects first)
        bs = [s if isinstance(s, np.ndarray) else _as_bytes(s) for s in sequences]
        n = len(bs)
        ptrs = (C.c_void_p * n)(*[b.ctypes.data if isinstance(b, np.ndarray) else C.cast(C.c_char_p(b), C.c_void_p).value for b in bs])
        lens = (C.c_int64 * n)(*[len(b) for b in bs])
        _check(self._L.indelgpu_set_reference(self._ctx, n, ptrs, lens))
        self.contig_lengths = [len(b) for b in bs]

    # ------------------------------------------------------------------ batched attempt_pe_alignment
    def attempt_pe_alignment_batch(self, reads, tid, position, range1, detail=False, cigars=False,
                                   packed=None):
        """reads: list of str/bytes (or pass packed=(uint8 data, int64 offsets)).
        tid/position/range1: per-read int arrays (range1 = range[1] of the read group)."""
        data, off = packed if packed is not None else pack_sequences(reads)
        n = len(off) - 1
        tid = np.ascontiguousarray(tid, dtype=np.int32)
        position = np.ascontiguousarray(position, dtype=np.int32)
        range1 = np.ascontiguousarray(range1, dtype=np.int32)
        assert len(tid) == n and len(position) == n and len(range1) == n
        cap = int(self._L.indelgpu_seg_bound(n, int(off[-1])))
        stride = int((off[1:] - off[:-1]).max()) + 4 if (cigars and n) else 0
        res = BatchResult(n, cap, detail, stride)
        b = _lib.Batch(n, data.ctypes.data, off.ctypes.data, tid.ctypes.data, position.ctypes.data,
                       range1.ctypes.data)
        r = _lib.Result(res.status.ctypes.data, res.nseg.ctypes.data, res.rstart.ctypes.data,
                        res.seg_off.ctypes.data, res.segs.ctypes.data, cap, 0,
                        res.detail.ctypes.data if detail else None,
                        res.cigar1.ctypes.data if stride else None,
                        res.cigar2.ctypes.data if stride else None, stride)
        _check(self._L.indelgpu_realign_batch(self._ctx, C.byref(b), C.byref(r)))
        res.seg_count = int(r.seg_count)
        res.launches = self._L.indelgpu_last_launch_count(self._ctx)
        res.cells, res.alg_bytes = self.last_counters()
        return res

    def attempt_pe_alignment_batch4(self, seq4, byte_off, length, flags, tid, position, range1):
        """The same batch with the reads as the BAM holds them (indelgpu_realign_batch4): 4 bits per base, high
        nibble first, every read on a byte boundary; flags bit 0 = reverse-complement on the device."""
        n = len(byte_off) - 1
        seq4 = np.ascontiguousarray(seq4, dtype=np.uint8)
        byte_off = np.ascontiguousarray(byte_off, dtype=np.int64)
        length = np.ascontiguousarray(length, dtype=np.int32)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        tid = np.ascontiguousarray(tid, dtype=np.int32)
        position = np.ascontiguousarray(position, dtype=np.int32)
        range1 = np.ascontiguousarray(range1, dtype=np.int32)
        cap = int(self._L.indelgpu_seg_bound(n, 2 * int(byte_off[-1])))
        res = BatchResult(n, cap, False, 0)
        b = _lib.Batch4(n, seq4.ctypes.data, byte_off.ctypes.data, length.ctypes.data, flags.ctypes.data,
                        tid.ctypes.data, position.ctypes.data, range1.ctypes.data)
        r = _lib.Result(res.status.ctypes.data, res.nseg.ctypes.data, res.rstart.ctypes.data,
                        res.seg_off.ctypes.data, res.segs.ctypes.data, cap, 0, None, None, None, 0)
        _check(self._L.indelgpu_realign_batch4(self._ctx, C.byref(b), C.byref(r)))
        res.seg_count = int(r.seg_count)
        res.launches = self._L.indelgpu_last_launch_count(self._ctx)
        res.cells, res.alg_bytes = self.last_counters()
        return res

    def last_counters(self):
        """((fwd, rev, glob) DP cells, algorithmic bytes) of the last batch call."""
        out = (C.c_int64 * 4)()
        _check(self._L.indelgpu_last_counters(self._ctx, out))
        return (int(out[0]), int(out[1]), int(out[2])), int(out[3])

    def attempt_pe_alignment(self, tid, position, range_, read):
        """One read, reference argument meaning (alignment.h:21-25; `sequences` is the resident
        reference).  Returns the evidence list [(variantclass, b1, b2)] or None."""
        assert range_[0] <= range_[1]                      # alignment.c:773
        res = self.attempt_pe_alignment_batch([read], [tid], [position], [range_[1]])
        if res.nseg[0] == 0:
            return None
        ev = res.evidence(0)
        return ev if ev else None

    # ------------------------------------------------------------------ kernel-level tasks
    def find_best_band_batch(self, reads, windows, anchor_rel):
        """find_best_band on independent (read, window) pairs; anchor_rel = anchor - zstart1."""
        rd, roff = pack_sequences(reads)
        wd, woff = pack_sequences(windows)
        n = len(roff) - 1
        anchor_rel = np.ascontiguousarray(anchor_rel, dtype=np.int32)
        low = np.zeros(n, dtype=np.int32)
        up = np.zeros(n, dtype=np.int32)
        _check(self._L.indelgpu_find_best_band_batch(self._ctx, n, rd.ctypes.data, roff.ctypes.data,
                                                     wd.ctypes.data, woff.ctypes.data,
                                                     anchor_rel.ctypes.data, low.ctypes.data, up.ctypes.data))
        return low, up

    def find_best_band(self, refseq, zstart1, end1, anchor, readseq, zstart2, end2):
        """alignment.c:393-447, same arguments; returns (low, up)."""
        a = np.array([anchor - zstart1], dtype=np.int64).astype(np.uint32).astype(np.int32)   # uint wrap, :431
        low, up = self.find_best_band_batch([_as_bytes(readseq)[zstart2:end2]],
                                            [_as_bytes(refseq)[zstart1:end1]], a)
        return int(low[0]), int(up[0])

    def band_align_batch(self, reads, windows, low, up, want_script=False, packed=None, want_cigar=True):
        """local_align (+ALIGN) + fetch_cigar on independent tasks.
        Returns dict(score, ends[n,4] = (si, sj, ei, ej), ncigar, cigar[n,stride], script, cells, shortcut_cells).
        want_cigar=False leaves the CIGAR words on the device (band-sweep micro-bench: scores, end points and
        CIGAR lengths still come back)."""
        if packed is not None:
            rd, roff, wd, woff = packed
        else:
            rd, roff = pack_sequences(reads)
            wd, woff = pack_sequences(windows)
        n = len(roff) - 1
        low = np.ascontiguousarray(low, dtype=np.int32)
        up = np.ascontiguousarray(up, dtype=np.int32)
        maxm = int((roff[1:] - roff[:-1]).max()) if n else 0
        maxband = int((up - low).max()) + 1 if n else 1
        cstride = (2 * maxm + maxband + 4) if maxband > 1 else maxm + 4   # ops <= M + N' + 2, N' <= M + band
        sstride = (2 * maxm + 2 * maxband + 4) if want_script else 0
        score = np.zeros(n, dtype=np.int32)
        ends = np.zeros((n, 4), dtype=np.int32)
        ncig = np.zeros(n, dtype=np.int32)
        cig = np.zeros((n, cstride), dtype=np.uint32) if want_cigar else None
        script = np.zeros((n, sstride), dtype=np.int32) if want_script else None
        cells = np.zeros(3, dtype=np.int64)
        _check(self._L.indelgpu_band_align_batch(
            self._ctx, n, rd.ctypes.data, roff.ctypes.data, wd.ctypes.data, woff.ctypes.data,
            low.ctypes.data, up.ctypes.data, score.ctypes.data, ends.ctypes.data, ncig.ctypes.data,
            cig.ctypes.data if want_cigar else None, cstride if want_cigar else 0, script.ctypes.data if want_script else None, sstride,
            cells.ctypes.data))
        skipped = C.c_int64(0)
        _check(self._L.indelgpu_last_shortcut_cells(self._ctx, C.byref(skipped)))
        return dict(score=score, ends=ends, ncigar=ncig, cigar=cig, script=script, cells=cells, shortcut_cells=int(skipped.value))

    def indel_support_batch(self, targets, queries, packed=None, out=None):
        """Batched realign_with_indel (variant.c:1246-1424) on already built targets (the reference interval
        with the variant spliced in, variant.c:1259-1275) and query slices (read[qstart:qstop], :1278-1283).
        `packed` = (target bytes, target_off, query bytes, query_off) skips the packing; `out` = three int32[n]
        arrays to receive the counters (pinned host memory makes both directions asynchronous copies).
        Returns dict(subs, indels, aligned: int32[n]; cells: sum of len1 * len2)."""
        if packed is not None:
            tb, toff, qb, qoff = packed
            n = len(toff) - 1
        else:
            n = len(targets)
            assert len(queries) == n
            tb, toff = pack_sequences(targets)
            qb, qoff = pack_sequences(queries)
        tb = tb if tb.size else np.zeros(1, dtype=np.uint8)
        qb = qb if qb.size else np.zeros(1, dtype=np.uint8)
        subs, indels, aligned = out if out is not None else (np.zeros(n, dtype=np.int32) for _ in range(3))
        cells = C.c_int64(0)
        _check(self._L.indelgpu_indel_support_batch(
            self._ctx, n, tb.ctypes.data, toff.ctypes.data, qb.ctypes.data, qoff.ctypes.data,
            subs.ctypes.data, indels.ctypes.data, aligned.ctypes.data, C.addressof(cells)))
        return dict(subs=subs, indels=indels, aligned=aligned, cells=cells.value,
                    launches=int(self._L.indelgpu_last_launch_count(self._ctx)))

    def attempt_band_alignment(self, refseq, zstart1, end1, readseq, zstart2, end2, low, up):
        """alignment.c:343-391, same arguments; returns ((r1, r2, q1, q2), cigar words)."""
        r = self.band_align_batch([_as_bytes(readseq)[zstart2:end2]], [_as_bytes(refseq)[zstart1:end1]],
                                  [low], [up])
        if r["score"][0] <= 0:
            return (0, 0, 0, 0), []
        q1, r1, q2, r2 = (int(x) for x in r["ends"][0])
        return ((r1 + zstart1 - 1, r2 + zstart1, q1 + zstart2 - 1, q2 + zstart2),
                [int(x) for x in r["cigar"][0][:r["ncigar"][0]]])


# ---------------------------------------------------------------------- reference prototypes
def _script_len(S, M, N):
    i = j = k = 0
    while i < M or j < N:
        op = S[k]
        k += 1
        if op == 0:
            i += 1
            j += 1
        elif op > 0:
            j += op
        else:
            i -= op
    return k


def local_align(seq1, seq2, indx1, indx2):
    """localalign.h:15-25 through the C symbol `local_align` of libindelgpu.so.
    Returns (score, (si, sj, ei, ej), script); score 0 => nothing else is defined."""
    L = _lib.load()
    seq1, seq2 = _as_bytes(seq1), _as_bytes(seq2)
    M, N = len(seq1), len(seq2)
    S = (C.c_int * (M + N + 2))()
    si, sj, ei, ej = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    b1, b2 = C.create_string_buffer(seq1), C.create_string_buffer(seq2)
    score = L.local_align(C.addressof(b1), M, C.addressof(b2), N, indx1, indx2,
                          C.addressof(si), C.addressof(sj), C.addressof(ei), C.addressof(ej), C.addressof(S))
    if score <= 0:
        return 0, (0, 0, 0, 0), []
    n = _script_len(S, ei.value - si.value + 1, ej.value - sj.value + 1)
    return score, (si.value, sj.value, ei.value, ej.value), list(S[:n])


def ALIGN(A, B, low, up, W=None, G=10, H=10):
    """globalalign.h:19-28 through the C symbol `ALIGN`; A, B are the sequences themselves (the
    1-based pointer convention is applied here).  Returns (score, script)."""
    L = _lib.load()
    A, B = _as_bytes(A), _as_bytes(B)
    M, N = len(A), len(B)
    if W is None:
        W = np.full((128, 128), -10, dtype=np.int32)
        np.fill_diagonal(W, 1)
    W = np.ascontiguousarray(W, dtype=np.int32)
    b1, b2 = C.create_string_buffer(b"\0" + A), C.create_string_buffer(b"\0" + B)
    S = (C.c_int * (M + N + 2))()
    score = L.ALIGN(C.addressof(b1), C.addressof(b2), M, N, low, up, W.ctypes.data, G, H, C.addressof(S))
    return score, list(S[:_script_len(S, M, N)])


def fetch_cigar(A, B, S, AP, readlength):
    """globalalign.h:39-48 through the C symbol `fetch_cigar`; A, B are the aligned sub-sequences.
    Returns (mismatches, cigar words)."""
    L = _lib.load()
    A, B = _as_bytes(A), _as_bytes(B)
    M, N = len(A), len(B)
    b1, b2 = C.create_string_buffer(b"\0" + A), C.create_string_buffer(b"\0" + B)
    Sa = (C.c_int * (len(S) + 1))(*S)
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.free.argtypes = [C.c_void_p]
    pc = C.c_void_p(libc.malloc(4))                      # caller-allocated 1 word (alignment.c:562)
    nops = C.c_int()
    mm = L.fetch_cigar(C.addressof(b1), C.addressof(b2), M, N, C.addressof(Sa), AP, 0, readlength,
                       C.addressof(nops), C.addressof(pc))
    words = list((C.c_uint32 * nops.value).from_address(pc.value)) if nops.value else []
    libc.free(pc)
    return mm, words
