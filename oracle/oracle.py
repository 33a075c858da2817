"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle (oracle/liboracle.so)
and, when built, the reference's own objects (oracle/_ref/libref_dp.so,
libref_align.so -- compiled by oracle/Makefile from /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (indelminer_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_DP_SO = os.path.join(HERE, "_ref", "libref_dp.so")
REF_ALIGN_SO = os.path.join(HERE, "_ref", "libref_align.so")
REF_VARIANT_SO = os.path.join(HERE, "_ref", "libref_variant.so")
ORC_MAXSEG = 1024


def build(ref=True):
    """Compile the oracle (and the reference objects when /root/reference exists)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(os.environ.get("INDEL_REF", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("klength", "numgaps", "maxdelsize", "ethreshold",
                 "match", "mismatch", "gapopen", "gapextend")]


class Cells(C.Structure):
    _fields_ = [("fwd", C.c_longlong), ("rev", C.c_longlong), ("glob", C.c_longlong)]


class Result(C.Structure):
    _fields_ = ([("status", C.c_int), ("nseg", C.c_int), ("nevidence", C.c_int)]
                + [(n, C.c_int * ORC_MAXSEG) for n in ("seg_op", "seg_len", "seg_start", "seg_end")]
                + [(n, C.c_int) for n in ("low1", "up1", "r1", "r2", "q1", "q2", "n1", "score1",
                                          "low2", "up2", "r3", "r4", "q3", "q4", "n2", "score2",
                                          "index")]
                + [("cigar1", C.c_uint32 * ORC_MAXSEG), ("cigar2", C.c_uint32 * ORC_MAXSEG)])

    def segments(self):
        return [(self.seg_op[i], self.seg_len[i], self.seg_start[i], self.seg_end[i])
                for i in range(self.nseg)]


def default_params(k=6, g=0, maxdel=1000, ethr=10):
    p = Params()
    lib().orc_default_params(C.byref(p))
    p.klength, p.numgaps, p.maxdelsize, p.ethreshold = k, g, maxdel, ethr
    return p


_lib = None
_ref_dp = None
_ref_align = None
_ref_variant = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _lib = C.CDLL(ORACLE_SO)
        _lib.orc_local_align.restype = C.c_int
        _lib.orc_global_align.restype = C.c_int
        _lib.orc_fetch_cigar.restype = C.c_int
        _lib.orc_attempt_band_alignment.restype = C.c_int
        _lib.orc_indel_target.restype = C.c_void_p
    return _lib


def have_ref():
    return os.path.exists(REF_DP_SO) and os.path.exists(REF_ALIGN_SO)


def ref_dp():
    global _ref_dp
    if _ref_dp is None:
        _ref_dp = C.CDLL(REF_DP_SO)
    return _ref_dp


def ref_align():
    global _ref_align
    if _ref_align is None:
        _ref_align = C.CDLL(REF_ALIGN_SO)
    return _ref_align


def _b(s):
    return s if isinstance(s, bytes) else s.encode()


# ------------------------------------------------------------------ oracle calls
def find_best_band(p, ref, zs1, e1, anchor, read, zs2, e2):
    low, up = C.c_int(), C.c_int()
    lib().orc_find_best_band(C.byref(p), _b(ref), C.c_uint32(zs1), C.c_uint32(e1),
                             C.c_uint32(anchor & 0xFFFFFFFF), _b(read), C.c_uint32(zs2),
                             C.c_uint32(e2), C.byref(low), C.byref(up))
    return low.value, up.value


def local_align(p, seq1, seq2, low, up, cells=None):
    seq1, seq2 = _b(seq1), _b(seq2)
    M, N = len(seq1), len(seq2)
    S = (C.c_int * (M + N + 2))()
    nS = C.c_int()
    si, sj, ei, ej = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    score = lib().orc_local_align(C.byref(p), seq1, M, seq2, N, low, up,
                                  C.byref(si), C.byref(sj), C.byref(ei), C.byref(ej),
                                  S, C.byref(nS), C.byref(cells) if cells is not None else None)
    if score <= 0:
        return 0, (0, 0, 0, 0), []
    return score, (si.value, sj.value, ei.value, ej.value), list(S[:nS.value])


def global_align(p, A, B, low, up, cells=None):
    A, B = _b(A), _b(B)
    M, N = len(A), len(B)
    S = (C.c_int * (M + N + 2))()
    nS = C.c_int()
    score = lib().orc_global_align(C.byref(p), A, B, M, N, low, up, S, C.byref(nS),
                                   C.byref(cells) if cells is not None else None)
    return score, list(S[:nS.value])


def attempt_band_alignment(p, ref, zs1, e1, read, zs2, e2, low, up, cells=None):
    r1, r2, q1, q2, sc = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    cig = (C.c_uint32 * (e2 - zs2 + 8))()
    n = lib().orc_attempt_band_alignment(C.byref(p), _b(ref), C.c_uint32(zs1), C.c_uint32(e1),
                                         _b(read), C.c_uint32(zs2), C.c_uint32(e2), low, up,
                                         C.byref(r1), C.byref(r2), C.byref(q1), C.byref(q2),
                                         cig, C.byref(sc),
                                         C.byref(cells) if cells is not None else None)
    return (r1.value, r2.value, q1.value, q2.value), list(cig[:n]), sc.value


def realign_read(p, ref, position, range1, read, reflength=None, cells=None):
    ref, read = _b(ref), _b(read)
    out = Result()
    lib().orc_realign_read(C.byref(p), ref, len(ref) if reflength is None else reflength,
                           position, range1, read, len(read), C.byref(out),
                           C.byref(cells) if cells is not None else None)
    return out


# ------------------------------------------------------------------ reference calls
def ref_set_params(k=6, g=0, maxdel=1000, ethr=10):
    ref_align().refshim_set_params(C.c_uint(k), C.c_uint(g), C.c_uint(maxdel), C.c_uint(ethr))


def ref_find_best_band(ref, zs1, e1, anchor, read, zs2, e2):
    low, up = C.c_int(), C.c_int()
    ref_align().refshim_find_best_band(_b(ref), C.c_uint(zs1), C.c_uint(e1),
                                       C.c_uint(anchor & 0xFFFFFFFF), _b(read), C.c_uint(zs2),
                                       C.c_uint(e2), C.byref(low), C.byref(up))
    return low.value, up.value


def ref_local_align(seq1, seq2, low, up):
    """reference local_align (localalign.h:15-25); returns (score, (si,sj,ei,ej), script)."""
    seq1, seq2 = _b(seq1), _b(seq2)
    M, N = len(seq1), len(seq2)
    # one guard byte in front: the reference forms A = seq1 - 1
    b1 = C.create_string_buffer(b"\0" + seq1 + b"\0")
    b2 = C.create_string_buffer(b"\0" + seq2 + b"\0")
    S = (C.c_int * (M + N + 2))()
    si, sj, ei, ej = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    p1 = C.cast(C.addressof(b1) + 1, C.c_char_p)
    p2 = C.cast(C.addressof(b2) + 1, C.c_char_p)
    f = ref_dp().local_align
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    score = f(C.addressof(b1) + 1, M, C.addressof(b2) + 1, N, low, up,
              C.addressof(si), C.addressof(sj), C.addressof(ei), C.addressof(ej), C.addressof(S))
    del p1, p2
    if score <= 0:
        return 0, (0, 0, 0, 0), []
    # script length: walk until both sequences are consumed
    m, n = ei.value - si.value + 1, ej.value - sj.value + 1
    i = j = k = 0
    while i < m or j < n:
        op = S[k]
        k += 1
        if op == 0:
            i += 1
            j += 1
        elif op > 0:
            j += op
        else:
            i -= op
    return score, (si.value, sj.value, ei.value, ej.value), list(S[:k])


_W = None


def _wtable():
    global _W
    if _W is None:
        w = np.full((128, 128), -10, dtype=np.int32)
        np.fill_diagonal(w, 1)
        _W = np.ascontiguousarray(w)
    return _W


def ref_ALIGN(A, B, low, up, G=10, H=10):
    """reference ALIGN (globalalign.h:19-28); returns (score, script)."""
    A, B = _b(A), _b(B)
    M, N = len(A), len(B)
    b1 = C.create_string_buffer(b"\0" + A + b"\0")
    b2 = C.create_string_buffer(b"\0" + B + b"\0")
    S = (C.c_int * (M + N + 2))()
    f = ref_dp().ALIGN
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                  C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    W = _wtable()
    score = f(C.addressof(b1), C.addressof(b2), M, N, low, up, W.ctypes.data, G, H, C.addressof(S))
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
    return score, list(S[:k])


def ref_attempt_band_alignment(ref, zs1, e1, read, zs2, e2, low, up):
    r1, r2, q1, q2 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    cig = (C.c_uint32 * (e2 - zs2 + 8))()
    f = ref_align().refshim_attempt_band_alignment
    f.restype = C.c_int
    n = f(_b(ref), C.c_uint(zs1), C.c_uint(e1), _b(read), C.c_uint(zs2), C.c_uint(e2),
          low, up, C.byref(r1), C.byref(r2), C.byref(q1), C.byref(q2), cig, len(cig))
    return (r1.value, r2.value, q1.value, q2.value), list(cig[:n])


def ref_realign(ref, position, range1, read, reflength=None):
    """reference attempt_diagonal_alignments + update_readsegs; returns (segments, nevidence)."""
    ref, read = _b(ref), _b(read)
    mx = ORC_MAXSEG
    op, ln, st, en = ((C.c_int * mx)() for _ in range(4))
    nev = C.c_int()
    f = ref_align().refshim_realign
    f.restype = C.c_int
    n = f(ref, len(ref) if reflength is None else reflength, position, range1, read,
          op, ln, st, en, mx, C.byref(nev))
    return [(op[i], ln[i], st[i], en[i]) for i in range(n)], nev.value


# ------------------------------------------------------------------ row f1: realign_with_indel (variant.c:1246-1424)
INSERTION, DELETION = 0, 1          # varianttype, evidence.h:14-18


def have_ref_variant():
    return os.path.exists(REF_VARIANT_SO)


def ref_variant():
    global _ref_variant
    if _ref_variant is None:
        _ref_variant = C.CDLL(REF_VARIANT_SO)
    return _ref_variant


def indel_target(reference, rstart, rstop, vtype, vstart, vstop, alternate):
    """the target string variant.c:1259-1275 builds (reference interval with the variant spliced in)"""
    L = lib()
    p = L.orc_indel_target(_b(reference), rstart, rstop, vtype, vstart, vstop, _b(alternate))
    out = C.string_at(p)
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(p)
    return out


def indel_support_dp(target, query, cells=None):
    """(subs, indels, aligned) of variant.c:1277-1423 for an already built target and query slice"""
    t, q = _b(target), _b(query)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    cc = C.c_longlong(0)
    lib().orc_indel_support_dp(t, len(t), q, len(q), C.byref(a), C.byref(b), C.byref(c), C.byref(cc))
    if cells is not None:
        cells[0] += cc.value
    return a.value, b.value, c.value


def realign_with_indel(reference, rstart, rstop, query, qstart, qstop, vtype, vstart, vstop, alternate):
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().orc_realign_with_indel(_b(reference), rstart, rstop, _b(query), qstart, qstop, vtype, vstart, vstop,
                                 _b(alternate), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def ref_realign_with_indel(reference, rstart, rstop, query, qstart, qstop, vtype, vstart, vstop, alternate):
    """the reference's own static function, through oracle/ref_shim_variant.c"""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    ref_variant().refshim_realign_with_indel(_b(reference), rstart, rstop, _b(query), qstart, qstop, vtype, vstart,
                                             vstop, _b(alternate), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value
