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
