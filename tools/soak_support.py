#!/usr/bin/env python
"""One-off soak test of the known-indel support check (row f1): EVERY pair of large batches compared with the oracle's
plain DP (orc_indel_support_dp), the oracle spread over the host cores.  Reference-shaped pairs (a 150 bp read against
the variant spliced into its reference interval) and low-complexity / repetitive / mixed-case pairs, where the tie
rules and the case handling decide the answer.  Not part of the test suite; its output is committed under profiles/.
  python tools/soak_support.py [--pairs 262144] [--seeds 2] [--out profiles/r02_soak_support.json]"""
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
    lo, hi = args
    from oracle import oracle as O
    t, out = _S["t"], _S["out"]
    bad = []
    for k in range(lo, hi):
        tt = t["targets"][t["target_off"][k]:t["target_off"][k + 1]].tobytes()
        q = t["queries"][t["query_off"][k]:t["query_off"][k + 1]].tobytes()
        if O.indel_support_dp(tt, q) != (int(out["subs"][k]), int(out["indels"][k]), int(out["aligned"][k])):
            bad.append(k)
    return bad


def hard_pairs(n, seed):
    """few letters, tandem repeats, runs, lower / mixed case, N: ties everywhere"""
    rng = np.random.default_rng(seed)
    T, Q = [], []
    for _ in range(n):
        kind = int(rng.integers(0, 5))
        alpha = [b"AC", b"A", b"ACGT", b"acgtACGT", b"ACGTN"][kind]
        a = np.frombuffer(alpha, dtype=np.uint8)
        unit = a[rng.integers(0, len(a), size=int(rng.integers(1, 7)))]
        L1 = int(rng.integers(1, 513))
        t = np.tile(unit, L1 // len(unit) + 1)[:L1].copy()
        noise = rng.random(L1) < 0.1
        t[noise] = a[rng.integers(0, len(a), size=int(noise.sum()))]
        s = int(rng.integers(0, max(1, L1 // 2)))
        q = t[s:s + int(rng.integers(1, 501))].copy()
        if len(q) > 12:
            cut = int(rng.integers(4, len(q) - 4))
            q = np.concatenate([q[:cut], q[cut + int(rng.integers(0, 9)):]]) if rng.random() < 0.5 else np.concatenate([q[:cut], unit, q[cut:]])[:500]
        if rng.random() < 0.3:
            q = np.frombuffer(q.tobytes().lower(), dtype=np.uint8)
        T.append(t)
        Q.append(q)
    toff = np.zeros(n + 1, dtype=np.int64); toff[1:] = np.cumsum([len(x) for x in T])
    qoff = np.zeros(n + 1, dtype=np.int64); qoff[1:] = np.cumsum([len(x) for x in Q])
    return dict(targets=np.ascontiguousarray(np.concatenate(T)), target_off=toff, queries=np.ascontiguousarray(np.concatenate(Q)), query_off=qoff)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1 << 18)
    ap.add_argument("--seeds", type=int, default=2)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import indelminer_b200
    from indelminer_b200 import synth
    R = indelminer_b200.Realigner()
    cores = os.cpu_count() or 1
    res = {"pairs_per_batch": a.pairs, "batches": []}
    for s in range(a.seeds):
        for name, t in (("reference-shaped (synth.make_support_tasks)", synth.make_support_tasks(a.pairs, seed=9000 + s)),
                        ("low-complexity / repeats / mixed case / N, lengths up to 512 x 500", hard_pairs(a.pairs // 4, 9100 + s))):
            n = len(t["target_off"]) - 1
            t0 = time.perf_counter()
            out = R.indel_support_batch(None, None, packed=(t["targets"], t["target_off"], t["queries"], t["query_off"]))
            t1 = time.perf_counter()
            _S["t"], _S["out"] = t, out
            step = (n + 8 * cores - 1) // (8 * cores)
            with mp.get_context("fork").Pool(cores) as pool:
                bad = sum(pool.map(worker, [(lo, min(n, lo + step)) for lo in range(0, n, step)]), [])
            t2 = time.perf_counter()
            line = {"seed": s, "kind": name, "pairs": n, "cells": int(out["cells"]), "mismatches": len(bad), "first_mismatches": bad[:10],
                    "gpu_call_s": round(t1 - t0, 3), "oracle_s": round(t2 - t1, 1), "oracle_cores": cores}
            print(json.dumps(line), flush=True)
            res["batches"].append(line)
    R.close()
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    sys.exit(1 if any(b["mismatches"] for b in res["batches"]) else 0)


if __name__ == "__main__":
    main()
