"""Writes tests/golden/indel_support.tsv.gz: seeded realign_with_indel cases (tests/util.py) with the
outputs of the REFERENCE's own static function (variant.c:1246-1424), called through
oracle/_ref/libref_variant.so (oracle/ref_shim_variant.c includes variant.c where it lies).
Run in the container that has /root/reference:  make -C oracle ref && python oracle/make_golden_support.py"""
import gzip
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path = [os.path.dirname(HERE)] + [p for p in sys.path if os.path.abspath(p or ".") != HERE]
from oracle import oracle as O                      # noqa: E402
from tests.util import GOLDEN, indel_support_cases, make_rng   # noqa: E402

cases = indel_support_cases(make_rng(20261018), 600)
with gzip.open(os.path.join(GOLDEN, "indel_support.tsv.gz"), "wt") as f:
    for c in cases:
        r = O.ref_realign_with_indel(*c)
        f.write("\t".join(str(x) for x in (*c, *r)) + "\n")
print(len(cases), "cases")
