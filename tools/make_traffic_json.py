#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of one realign_kernel launch of the default bench batch:
DRAM bytes of the launch, its duration and instruction count, and the hash of csrc/ they were measured on
(bench.py reports `roofline.traffic` only while that hash equals the build's).

  python tools/make_traffic_json.py <capture.ncu-rep> [out.json] [raw.csv]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "traffic.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    row = [r for r in rows[2:] if "realign_kernel" in r[hdr.index("Kernel Name")]][0]
    d, u = dict(zip(hdr, row)), dict(zip(hdr, units))

    def num(key, want):
        v = float(d[key].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1,
                 "nsecond": 1e-6, "second": 1e3, "inst": 1, "": 1}[u[key]]
        return v * scale

    from bench import csrc_sha
    res = {"dram_bytes_per_launch": int(num("dram__bytes_read.sum", "byte") + num("dram__bytes_write.sum", "byte")),
           "csrc_sha": csrc_sha(), "kernel": d["Kernel Name"],
           "source": "ncu --set full --clock-control none, one launch of the default bench batch (%s)" % os.path.basename(rep),
           "gpu_time_ms": num("gpu__time_duration.sum", "ms"),
           "warp_instructions": int(num("smsp__inst_executed.sum", "inst"))}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
