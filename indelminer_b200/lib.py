"""ctypes loader for libindelgpu.so (the C ABI of include/indelgpu.h).

There is deliberately no fallback: if the shared library is missing the import of the
product API fails loudly, and if no GPU is usable every call returns an error.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libindelgpu.so")

# every symbol include/indelgpu.h declares
EXPORTS = [
    "indelgpu_version", "indelgpu_default_params", "indelgpu_last_error",
    "indelgpu_create", "indelgpu_destroy", "indelgpu_device", "indelgpu_sm_count",
    "indelgpu_host_alloc", "indelgpu_host_free", "indelgpu_set_reference",
    "indelgpu_seg_bound", "indelgpu_realign_batch", "indelgpu_realign_batch4", "indelgpu_realign_batch_device",
    "indelgpu_last_counters", "indelgpu_last_error_flag", "indelgpu_last_shortcut_cells", "indelgpu_last_launch_count", "indelgpu_last_kernel_ms", "indelgpu_int32_peak",
    "indelgpu_find_best_band_batch", "indelgpu_band_align_batch", "indelgpu_indel_support_batch", "indelgpu_device_count",
    "local_align", "ALIGN", "DISPLAY", "fetch_cigar",
]


class Params(C.Structure):
    """indelgpu_params: alignment.c:3-9 globals + localalign.c:10-13 scoring."""
    _fields_ = [(n, C.c_int32) for n in
                ("klength", "numgaps", "maxdelsize", "ethreshold",
                 "match", "mismatch", "gapopen", "gapextend")]


class Detail(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("low1", "up1", "r1", "r2", "q1", "q2", "n1", "score1",
                 "low2", "up2", "r3", "r4", "q3", "q4", "n2", "score2",
                 "index", "cells_fwd", "cells_rev", "cells_glob")]


class Batch(C.Structure):
    _fields_ = [("n", C.c_int32), ("read_bases", C.c_void_p), ("read_off", C.c_void_p),
                ("tid", C.c_void_p), ("position", C.c_void_p), ("range1", C.c_void_p)]


class Batch4(C.Structure):
    """indelgpu_batch4: reads in the BAM's own 4-bit form"""
    _fields_ = [("n", C.c_int32), ("seq4", C.c_void_p), ("byte_off", C.c_void_p), ("len", C.c_void_p), ("flags", C.c_void_p),
                ("tid", C.c_void_p), ("position", C.c_void_p), ("range1", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("status", C.c_void_p), ("nseg", C.c_void_p), ("rstart", C.c_void_p),
                ("seg_off", C.c_void_p), ("segs", C.c_void_p), ("seg_capacity", C.c_int64),
                ("seg_count", C.c_int64), ("detail", C.c_void_p), ("cigar1", C.c_void_p),
                ("cigar2", C.c_void_p), ("cigar_stride", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m indelminer_b200.build` "
            "(nvcc, sm_100a). indelminer_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.indelgpu_version.restype = C.c_int
    L.indelgpu_last_error.restype = C.c_char_p
    L.indelgpu_create.restype = C.c_void_p
    L.indelgpu_create.argtypes = [C.c_int, C.POINTER(Params)]
    L.indelgpu_destroy.argtypes = [C.c_void_p]
    L.indelgpu_device.argtypes = [C.c_void_p]
    L.indelgpu_sm_count.argtypes = [C.c_void_p]
    L.indelgpu_host_alloc.restype = C.c_void_p
    L.indelgpu_host_alloc.argtypes = [C.c_size_t]
    L.indelgpu_host_free.argtypes = [C.c_void_p]
    L.indelgpu_set_reference.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    L.indelgpu_seg_bound.restype = C.c_int64
    L.indelgpu_seg_bound.argtypes = [C.c_int32, C.c_int64]
    L.indelgpu_realign_batch.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(Result)]
    L.indelgpu_realign_batch4.argtypes = [C.c_void_p, C.POINTER(Batch4), C.POINTER(Result)]
    L.indelgpu_realign_batch_device.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_int32, C.c_int32,
                                                C.POINTER(Result), C.c_void_p, C.c_void_p]
    L.indelgpu_last_counters.argtypes = [C.c_void_p, C.c_void_p]
    L.indelgpu_last_launch_count.argtypes = [C.c_void_p]
    L.indelgpu_last_error_flag.argtypes = [C.c_void_p]
    L.indelgpu_last_shortcut_cells.argtypes = [C.c_void_p, C.c_void_p]
    L.indelgpu_last_kernel_ms.restype = C.c_double
    L.indelgpu_last_kernel_ms.argtypes = [C.c_void_p]
    L.indelgpu_int32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.indelgpu_find_best_band_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 7
    L.indelgpu_band_align_batch.argtypes = ([C.c_void_p, C.c_int32] + [C.c_void_p] * 10
                                            + [C.c_int32, C.c_void_p, C.c_int32, C.c_void_p])
    L.indelgpu_indel_support_batch.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
    L.local_align.restype = C.c_int
    L.local_align.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 5
    L.ALIGN.restype = C.c_int
    L.ALIGN.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                        C.c_int, C.c_int, C.c_void_p]
    L.fetch_cigar.restype = C.c_int
    L.fetch_cigar.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                              C.c_int, C.c_void_p, C.c_void_p]
    _lib = L
    return L


def last_error():
    return load().indelgpu_last_error().decode()
