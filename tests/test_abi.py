"""CPU-side checks of the drop-in boundary: libindelgpu.so builds, loads and exports every symbol
include/indelgpu.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from indelminer_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "indelgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src))
    return {n for n in names if n not in ("defined",)}


def test_exports_every_declared_symbol(libpath):
    from indelminer_b200 import lib
    L = C.CDLL(libpath)
    decl = declared_symbols()
    assert {"local_align", "ALIGN", "DISPLAY", "fetch_cigar", "indelgpu_realign_batch"} <= decl
    assert decl == set(lib.EXPORTS)
    for name in decl:
        assert hasattr(L, name), name


def test_reference_prototypes_are_kept():
    """the four entry points keep the reference's argument lists (localalign.h:15-25, globalalign.h:19-48)"""
    src = open(os.path.join(ROOT, "include", "indelgpu.h")).read()
    flat = re.sub(r"\s+", " ", src)
    assert ("int local_align(char* seq1, const int seq1len, char* seq2, const int seq2len, const int indx1, "
            "const int indx2, int* const psi, int* const psj, int* const pei, int* const pej, int* const S);") in flat
    assert "int ALIGN(char* A, char* B, int M, int N, int low, int up, int W[][128], int G, int H, int* S);" in flat
    assert ("int fetch_cigar(char* A, char* B, int M, int N, int* S, int AP, int BP, const int readlength, "
            "int* const pnumops, uint32_t** pcigar);") in flat
    assert "int DISPLAY(FILE* F, char* A, char* B, int M, int N, int* S, int AP, int BP);" in flat


def test_no_gpu_means_loud_failure(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from indelminer_b200 import IndelGpuError, Realigner
    with pytest.raises(IndelGpuError):
        Realigner()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "indelminer_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu", ""), os.path.join(dirpath, f)


def test_display_matches_reference_byte_for_byte(tmp_path):
    """DISPLAY (globalalign.h:30-37) is pure host formatting, so it can be checked without a GPU:
    same script through libindelgpu.so and through the reference object (oracle/_ref/libref_dp.so)."""
    import ctypes as C
    import random
    ref_so = os.path.join(ROOT, "oracle", "_ref", "libref_dp.so")
    if not os.path.exists(ref_so):
        import pytest
        pytest.skip("oracle/_ref not built")
    from indelminer_b200 import lib as _lib
    ours, ref = _lib.load(), C.CDLL(ref_so)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    rng = random.Random(5)
    for case in range(40):
        # a random script over random sequences: 0 = pair, +k = k reference symbols, -k = k read symbols
        S, M, N = [], 0, 0
        for _ in range(rng.randrange(1, 40)):
            r = rng.random()
            if r < 0.7:
                n = rng.randrange(1, 30); S += [0] * n; M += n; N += n
            elif r < 0.85:
                n = rng.randrange(1, 70); S.append(n); N += n
            else:
                n = rng.randrange(1, 70); S.append(-n); M += n
        A = b"\0" + bytes(rng.choice(b"ACGT") for _ in range(M))
        B = b"\0" + bytes(rng.choice(b"ACGT") for _ in range(N))
        outs = []
        for tag, L in (("ours", ours), ("ref", ref)):
            path = str(tmp_path / f"{tag}{case}.txt").encode()
            fp = libc.fopen(path, b"w")
            Sa = (C.c_int * (len(S) + 1))(*S)
            L.DISPLAY.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
            L.DISPLAY(fp, A, B, M, N, Sa, 7, 1001)
            libc.fclose(fp)
            outs.append(open(path, "rb").read())
        assert outs[0] == outs[1], case
