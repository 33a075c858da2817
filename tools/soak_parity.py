#!/usr/bin/env python
"""One-off soak test: EVERY read of large synthetic batches compared with the oracle (status, reference start, segment
words), the oracle spread over the host cores.  Not part of the test suite (it needs the GPU and a minute of 16 cores);
its output is committed under profiles/.
  python tools/soak_parity.py [--reads 1048576] [--numgaps 0] [--seeds 3] [--out profiles/r02_soak.json]"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_S = {}


def worker(args):
    lo, hi, g = args
    from oracle import oracle as O
    from indelminer_b200.api import walk_segments
    p = O.default_params(6, g)
    w, cs, res = _S["w"], _S["cs"], _S["res"]
    M = w["read_len"]
    bad = []
    hist = {}
    for i in range(lo, hi):
        o = O.realign_read(p, cs, int(w["position"][i]), int(w["range1"][i]), w["read_bases"][i * M:(i + 1) * M].tobytes())
        ns = int(res["nseg"][i])
        got = walk_segments(int(res["rstart"][i]), res["segs"][res["seg_off"][i]:res["seg_off"][i] + ns]) if ns else []
        if int(res["status"][i]) != o.status or got != o.segments():
            bad.append(i)
        hist[o.status] = hist.get(o.status, 0) + 1
    return bad, hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=1 << 20)
    ap.add_argument("--numgaps", type=int, default=0)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import indelminer_b200
    from indelminer_b200 import synth, api
    ref = synth.make_reference(16_000_000, seed=77, n_frac=0.0005)
    R = indelminer_b200.Realigner(numgaps=a.numgaps)
    R.set_reference([ref.tobytes()])
    out = {"reads_per_batch": a.reads, "numgaps": a.numgaps, "reference": "16 Mb, 0.05 % N", "batches": []}
    cores = os.cpu_count() or 1
    for s in range(a.seeds):
        w = synth.make_candidates(ref, a.reads, seed=4000 + s)
        t0 = time.perf_counter()
        if s % 2 == 0:
            res = R.attempt_pe_alignment_batch(None, w["tid"], w["position"], w["range1"], packed=(w["read_bases"], w["read_off"]))
            entry = "indelgpu_realign_batch (ASCII)"
        else:
            rc = np.random.default_rng(s).random(a.reads) < 0.5
            seq4, boff, lens, fl = api.pack4(w["read_bases"], w["read_off"], rc)
            res = R.attempt_pe_alignment_batch4(seq4, boff, lens, fl, w["tid"], w["position"], w["range1"])
            entry = "indelgpu_realign_batch4 (4-bit, half of the reads stored reverse-complemented)"
        t_gpu = time.perf_counter() - t0
        _S.update(w=w, cs=ref.tobytes(), res=dict(status=res.status, nseg=res.nseg, rstart=res.rstart, seg_off=res.seg_off, segs=res.segs))
        per = (a.reads + cores - 1) // cores
        jobs = [(c * per, min(a.reads, (c + 1) * per), a.numgaps) for c in range(cores) if c * per < a.reads]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            parts = pool.map(worker, jobs)
        t_cpu = time.perf_counter() - t0
        bad = [i for b, _h in parts for i in b]
        hist = {}
        for _b, h in parts:
            for k, v in h.items():
                hist[k] = hist.get(k, 0) + v
        row = {"seed": 4000 + s, "entry_point": entry, "reads": a.reads, "mismatches": len(bad), "first_mismatches": bad[:5],
               "status_histogram": {str(k): v for k, v in sorted(hist.items())}, "gpu_call_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 1), "oracle_cores": len(jobs)}
        print(json.dumps(row), flush=True)
        out["batches"].append(row)
    if a.out:
        with open(os.path.join(ROOT, a.out), "w") as f:
            json.dump(out, f, indent=1)
    R.close()
    sys.exit(1 if any(b["mismatches"] for b in out["batches"]) else 0)


if __name__ == "__main__":
    main()
