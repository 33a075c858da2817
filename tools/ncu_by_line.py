#!/usr/bin/env python
"""Per-source-line totals of an ncu capture: joins `ncu --page source --csv` (SASS rows with
executed-instruction counts and stall samples) with `nvdisasm -g` line info of the same cubin.

  python tools/ncu_by_line.py <prof.ncu-rep> <lib.so> <kernel-mangled-substring> [top]

Prints, per (file, line): warp instructions executed, share, stall samples, share.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    table = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, active = None, False
        for ln in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                active = kernel_sub in m.group(1)
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m and cur:
                table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, lib, ksub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    table = line_table(lib, ksub)
    for b in blocks[:1]:
        hdr = b["rows"][0]
        ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        base = None
        inst, samp = defaultdict(int), defaultdict(int)
        for r in b["rows"][1:]:
            if len(r) <= isamp:
                continue
            addr = int(r[ia], 16)
            if base is None:
                base = addr
            key = table.get(addr - base, ("?", 0))
            inst[key] += int(r[ii] or 0)
            samp[key] += int(r[isamp] or 0)
        ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
        print(f"# {b['name']}: {ti} warp instructions, {ts} samples")
        print(f"{'file:line':38s} {'inst':>14s} {'%':>6s} {'samples':>9s} {'%':>6s}")
        for key in sorted(inst, key=lambda k: -samp[k])[:top]:
            print(f"{key[0] + ':' + str(key[1]):38s} {inst[key]:14d} {100 * inst[key] / ti:6.2f} {samp[key]:9d} {100 * samp[key] / ts:6.2f}")


if __name__ == "__main__":
    main()
