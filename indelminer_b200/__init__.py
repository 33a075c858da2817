"""indelminer_b200: B200-native (sm_100a) split-read realignment for indelMINER.

The product is libindelgpu.so (include/indelgpu.h, indelminer_b200/csrc); this package is the
thin host mirror of the reference's interface used by the tests and the benchmark.
"""
from .api import (ALIGN, BatchResult, IndelGpuError, Realigner, fetch_cigar, local_align,  # noqa: F401
                  pack_sequences, walk_segments)

__all__ = ["ALIGN", "BatchResult", "IndelGpuError", "Realigner", "fetch_cigar", "local_align",
           "pack_sequences", "walk_segments"]
