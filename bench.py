#!/usr/bin/env python
"""Benchmark of the split-read realignment hot path (BASELINE.json: reads/sec and GCUPS).

A step = one pass of the batched attempt_pe_alignment over one batch of synthetic candidate reads
of the cfg3 shape (64 Mb contig, planted 1-50 bp indels, 2x150 bp pairs; SURVEY.md 8d "D2").

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, libindelgpu.so)
  python bench.py --impl reference ...                            the reference's own CPU code
  python bench.py --workload band --band 33                       banded-DP micro-bench (D1): GCUPS
  python bench.py --workload support                              known-indel support check (row f1): GCUPS

Under torchrun (N > 1) every rank realigns its own region's candidates (weak scaling, no
data-path collective); NCCL carries only the tiny per-read-group insert-range all-reduce.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "split_read_realign_reads_per_sec"
UNIT = "reads/s"
INT_OPS_PER_CELL = 10          # fixed constant of the INT32 roofline


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "band", "support"])
    ap.add_argument("--reads", type=int, default=1 << 20, help="candidate reads per GPU per step")
    ap.add_argument("--ref-mb", type=int, default=64, help="contig size in Mb")
    ap.add_argument("--band", type=int, default=33)
    ap.add_argument("--numgaps", type=int, default=0)
    ap.add_argument("--maxdel", type=int, default=1000, help="-s of the reference (experiments only; the metric is quoted at 1000)")
    ap.add_argument("--tasks", type=int, default=1 << 17, help="alignments / pairs for --workload band / support (SURVEY D1 size: 1048576)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-ascii", action="store_true", help="e2e leg through indelgpu_realign_batch (1 byte per base) instead of the 4-bit entry point")
    ap.add_argument("--no-extra", action="store_true", help="skip the band sweep, the support check and the VCF wall time")
    ap.add_argument("--sweep-tasks", type=int, default=1 << 20, help="alignments per band of extra.band_sweep (SURVEY D1: 2^20)")
    ap.add_argument("--vcf-mb", type=int, default=4, help="contig size of the end-to-end VCF wall-time run")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark(self, wait_s=3.0):
        """call right before the timed region: the sampler was started earlier (nvidia-smi needs a few hundred ms before
        its first line, more than a short timed region lasts), only the lines from here on are the region's"""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < wait_s:
            time.sleep(0.01)
        self.first = len(self.lines)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        first = getattr(self, "first", 0)
        lines = self.lines[first:] or self.lines[-1:]          # a region shorter than the sampling period: the line just before it
        for ln in lines:
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU arms
_CPU_SHARED = {}        # inherited by the forked workers (no pickling of the 64 MB contig)


def cpu_worker(args):
    """one process = one core: the reference's (or the port's) loop over a slice of the sample"""
    kind, lo, hi, as_shipped = args
    from oracle import oracle as O
    ref_bytes, w = _CPU_SHARED["ref"], _CPU_SHARED["w"]
    M = w["read_len"]
    reflength = len(ref_bytes)
    reads, off = w["read_bases"][lo * M:hi * M], w["read_off"][lo:hi + 1]
    pos, rng = w["position"][lo:hi], w["range1"][lo:hi]
    n = len(pos)
    reads = np.ascontiguousarray(reads)
    off = np.ascontiguousarray(off - off[0], dtype=np.int64)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    rng = np.ascontiguousarray(rng, dtype=np.int32)
    t0 = time.perf_counter()
    if kind == "reference":
        L = O.ref_align()
        O.ref_set_params()
        f = L.refshim_realign_batch
        f.restype = C.c_long
        f(C.c_char_p(ref_bytes), reflength, n, reads.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
          pos.ctypes.data_as(C.c_void_p), rng.ctypes.data_as(C.c_void_p), int(as_shipped), None)
    else:
        L = O.lib()
        p = O.default_params()
        f = L.orc_realign_batch
        f.restype = C.c_long
        f(C.byref(p), C.c_char_p(ref_bytes), reflength, n, reads.ctypes.data_as(C.c_void_p),
          off.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), rng.ctypes.data_as(C.c_void_p), None)
    return time.perf_counter() - t0


def cpu_arm(ref, w, nsample, cores, as_shipped=False):
    """reads/s of the CPU implementation on the first `nsample` reads, `cores` processes."""
    import multiprocessing as mp
    from oracle import oracle as O
    kind = "reference" if O.have_ref() else "port"
    if kind == "port":
        O.lib()
    nsample = min(nsample, len(w["position"]))
    if _CPU_SHARED.get("refid") != id(ref):
        _CPU_SHARED.update(ref=ref.tobytes(), refid=id(ref))
    _CPU_SHARED["w"] = w
    per = (nsample + cores - 1) // cores
    jobs = []
    for c in range(cores):
        a, b = c * per, min(nsample, (c + 1) * per)
        if a >= b:
            break
        jobs.append((kind, a, b, as_shipped))
    t0 = time.perf_counter()
    if len(jobs) == 1:
        cpu_worker(jobs[0])
    else:
        with mp.get_context("fork").Pool(len(jobs)) as pool:
            pool.map(cpu_worker, jobs)
    dt = time.perf_counter() - t0
    return nsample / dt, kind, len(jobs), dt


def calibrate_cpu_sample(ref, w, cores, target_s=12.0):
    """size the sample for ~target_s seconds of wall time on `cores` cores"""
    probe = 2000
    rate, _k, _c, _dt = cpu_arm(ref, w, probe, 1)
    return int(max(probe, min(len(w["position"]), rate * cores * target_s)))


def bench_config(a):
    """the same `config` object for both arms (the driver compares them key by key)"""
    return {"workload": "cfg3: 64 Mb synthetic contig, planted 1-50 bp indels, 2x150 bp candidates, -g %d -k 6" % a.numgaps,
            "reads_per_gpu_per_step": a.reads, "ref_mb": a.ref_mb, "read_len": 150, "max_range": 700,
            "sharding": "contig region per rank, no data-path collective",
            "l2": "256 MB flush write between timed iterations (outside the event pairs)"}


def csrc_sha():
    """hash of the CUDA sources: ties a committed ncu number to the build it was measured on"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "indelminer_b200", "csrc")
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh")):
            with open(os.path.join(d, fn), "rb") as f:
                h.update(f.read())
    return h.hexdigest()[:16]


# --------------------------------------------------------------------------- main arms
def run_reference(a, rank, world):
    if rank != 0:
        return
    from indelminer_b200 import synth
    ref = synth.make_reference(a.ref_mb * 1_000_000, seed=1)
    w = synth.make_candidates(ref, a.reads, seed=20261018)
    cores = os.cpu_count() or 1
    nsample = a.cpu_sample or calibrate_cpu_sample(ref, w, cores, target_s=8.0)
    vals = []
    for s in range(a.warmup + a.steps):
        rate, kind, used, dt = cpu_arm(ref, w, nsample, cores)
        if s >= a.warmup:
            vals.append((rate, dt))
    rate = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": bench_config(a),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": f"each step = the first {nsample} of the {a.reads} candidate reads of the config's batch, function level "
                                   "(attempt_diagonal_alignments with windows precomputed; the as-shipped per-read "
                                   "strlen(contig) of alignment.c:771 is excluded), one forked process per core; the reference "
                                   "leaks its evidence records (oracle/ref_shim.c), harmless per forked step"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(a, rank, world, local_rank):
    import torch
    import indelminer_b200
    from indelminer_b200 import lib as _lib, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peaks()
    L = _lib.load()
    R = indelminer_b200.Realigner(device=local_rank, numgaps=a.numgaps, maxdelsize=a.maxdel)
    ref = synth.make_reference(a.ref_mb * 1_000_000, seed=1)
    R.set_reference([ref.tobytes()])

    if a.workload == "band":
        return run_band(a, R, L, torch, dev, peak, peak_src, local_rank)
    if a.workload == "support":
        return run_support(a, R, L, torch, local_rank)

    # region sharding: rank r owns the r-th slice of the contig (weak scaling: same count per GPU)
    Lr = len(ref)
    region = (rank * Lr // world, (rank + 1) * Lr // world)
    w = synth.make_candidates(ref, a.reads, seed=20261018 + rank, region=region)
    n, M = a.reads, w["read_len"]

    # the only collective on the path: per-read-group max proper insert (range[1]); a few bytes
    rg = torch.tensor([int(w["range1"].max())], device=dev, dtype=torch.int32)
    if dist is not None:
        dist.all_reduce(rg, op=dist.ReduceOp.MAX)
    max_range = int(rg.item())

    # ---- device-resident inputs (value) -------------------------------------------------
    d_reads = torch.from_numpy(w["read_bases"]).to(dev)
    d_off = torch.from_numpy(w["read_off"]).to(dev)
    d_tid = torch.from_numpy(w["tid"]).to(dev)
    d_pos = torch.from_numpy(w["position"]).to(dev)
    d_rng = torch.from_numpy(w["range1"]).to(dev)
    cap = int(L.indelgpu_seg_bound(n, n * M))
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    d_nseg = torch.empty(n, dtype=torch.int32, device=dev)
    d_rstart = torch.empty(n, dtype=torch.int32, device=dev)
    d_segoff = torch.empty(n, dtype=torch.int64, device=dev)
    d_segs = torch.empty(cap, dtype=torch.int32, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    b = _lib.Batch(n, d_reads.data_ptr(), d_off.data_ptr(), d_tid.data_ptr(), d_pos.data_ptr(), d_rng.data_ptr())
    r = _lib.Result(d_status.data_ptr(), d_nseg.data_ptr(), d_rstart.data_ptr(), d_segoff.data_ptr(),
                    d_segs.data_ptr(), cap, 0, None, None, None, 0)
    stream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: the library treats NULL as "use mine"
    torch.cuda.set_stream(stream)

    def step_device():
        rc = L.indelgpu_realign_batch_device(R._ctx, C.byref(b), M, max_range, C.byref(r),
                                             d_count.data_ptr(), C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(_lib.last_error())

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(a.warmup):
        flush.fill_(1)
        step_device()
    barrier()
    sampler.mark()
    evs = []
    launches = 0
    for _ in range(a.steps):
        flush.fill_(1)                       # L2 flush between timed iterations, outside the event pair
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        launches += L.indelgpu_last_launch_count(R._ctx)
        evs.append((e0, e1))
    barrier()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    clocks = sampler.stop()
    total_ms = float(sum(step_ms))
    (cf, cr, cg), alg_bytes = R.last_counters()
    nsplit = int((d_status == 6).sum().item())

    # ---- end to end through the host API (pinned host buffers in, host results out) -----
    def pinned(arr):
        p = L.indelgpu_host_alloc(arr.nbytes)
        out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(arr.nbytes,)).view(arr.dtype).reshape(arr.shape)
        out[...] = arr
        return out

    # the host side of the e2e leg holds the reads the way the BAM does (4 bits per base): indelgpu_realign_batch4
    from indelminer_b200 import api as _api
    seq4, boff, lens4, flags4 = _api.pack4(w["read_bases"], w["read_off"])
    h = {"seq4": pinned(seq4), "byte_off": pinned(boff), "len": pinned(lens4), "flags": pinned(flags4),
         "tid": pinned(w["tid"]), "position": pinned(w["position"]), "range1": pinned(w["range1"])}
    ho = {"status": pinned(np.zeros(n, np.int32)), "nseg": pinned(np.zeros(n, np.int32)),
          "rstart": pinned(np.zeros(n, np.int32)), "seg_off": pinned(np.zeros(n, np.int64)),
          "segs": pinned(np.zeros(cap, np.uint32))}
    hb = _lib.Batch4(n, h["seq4"].ctypes.data, h["byte_off"].ctypes.data, h["len"].ctypes.data, h["flags"].ctypes.data,
                     h["tid"].ctypes.data, h["position"].ctypes.data, h["range1"].ctypes.data)
    hr = _lib.Result(ho["status"].ctypes.data, ho["nseg"].ctypes.data, ho["rstart"].ctypes.data,
                     ho["seg_off"].ctypes.data, ho["segs"].ctypes.data, cap, 0, None, None, None, 0)

    if a.e2e_ascii:
        h = {k: pinned(w[k]) for k in ("read_bases", "read_off", "tid", "position", "range1")}
        hb = _lib.Batch(n, h["read_bases"].ctypes.data, h["read_off"].ctypes.data, h["tid"].ctypes.data,
                        h["position"].ctypes.data, h["range1"].ctypes.data)

    def step_host():
        rc = (L.indelgpu_realign_batch if a.e2e_ascii else L.indelgpu_realign_batch4)(R._ctx, C.byref(hb), C.byref(hr))
        if rc != 0:
            raise RuntimeError(_lib.last_error())

    for _ in range(max(1, a.warmup)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    seg_words = int(hr.seg_count)
    assert np.array_equal(ho["status"], d_status.cpu().numpy()), "host and device paths disagree"
    assert np.array_equal(ho["nseg"], d_nseg.cpu().numpy()) and np.array_equal(ho["rstart"], d_rstart.cpu().numpy())
    h2d = int(sum(v.nbytes for v in h.values()))
    d2h = int(20 * n + 4 * seg_words + 64)

    # max over ranks, whole-job aggregate
    t = torch.tensor([total_ms, e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0].item()), float(t[1].item())
    ms_per_step = total_ms / a.steps
    value = world * n * a.steps / (total_ms / 1e3)
    e2e_value = world * n * a.steps / e2e_s
    kernel_s = float(np.mean(step_ms)) / 1e3          # one launch per step: step time == kernel time
    achieved = alg_bytes / kernel_s / 1e9
    cells = cf + cr + cg

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": bench_config(a),
        "split_reads_per_step": nsplit,
        "gcups": cells / kernel_s / 1e9,
        "cells_per_step": {"forward": cf, "reverse": cr, "align": cg},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "entry_point": ("indelgpu_realign_batch: pinned host buffers, ASCII reads" if a.e2e_ascii else
                                "indelgpu_realign_batch4: pinned host buffers, reads in the BAM's 4-bit form") + ", chunked copies overlapping the kernels"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None,
                     "kernel": "realign_kernel<DIRECT, HB=8> (fused, warp per read)" if a.numgaps == 0
                               else "pipe_vote / pipe_dp / pipe_combine (5 launches per step; time = whole step)",
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_s * 1e3, "peak_source": peak_src},
    }
    assert M == 150 and max_range == 700, "bench_config() states the generator's read length and range"
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and a.numgaps == 0 and a.reads == 1 << 20:
        # DRAM bytes of one launch from the committed `ncu --set full` capture -- only if it was taken on THIS build
        try:
            with open(prof) as f:
                tj = json.load(f)
            if tj.get("csrc_sha") == csrc_sha():
                line["roofline"]["traffic"] = tj.get("dram_bytes_per_launch")
                line["roofline"]["traffic_source"] = tj.get("source")
        except Exception:
            pass
    err_flag = L.indelgpu_last_error_flag(R._ctx)
    assert err_flag == 0, f"device error flag {err_flag} after the timed region"
    if rank == 0:
        # the timed batch itself, spot-checked against the oracle after the timed region (never inside it)
        line["oracle_spot_check"] = spot_check(R, w, d_status.cpu().numpy(), d_nseg.cpu().numpy(), d_rstart.cpu().numpy(),
                                               d_segoff.cpu().numpy(), d_segs.cpu().numpy(), ref, 4096, a.numgaps)
    if rank == 0 and world == 1 and not a.no_cpu:
        cores = os.cpu_count() or 1
        nsample = a.cpu_sample or calibrate_cpu_sample(ref, w, cores)
        rate, kind, used, dt = cpu_arm(ref, w, nsample, cores)
        shipped_n = 200
        srate, _k, _u, _d = cpu_arm(ref, w, shipped_n, 1, as_shipped=True) if kind == "reference" else (None, 0, 0, 0)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": used, "kind": kind,
            "sample": f"first {nsample} of {n} candidate reads ({dt:.1f} s), function level: "
                      "attempt_diagonal_alignments with windows precomputed, one process per core",
            "as_shipped_1core": srate,
            "as_shipped_note": f"attempt_pe_alignment incl. its per-read strlen(contig) (alignment.c:771), {shipped_n} reads, 1 core",
        }
    if rank == 0 and world == 1 and not a.no_extra:
        # the other parts of BASELINE.json's metric, measured in the same process and clock-stamped:
        # banded-DP GCUPS against the INT32 issue rate (D1 sweep), the support check (row f1), end-to-end VCF wall time
        line["extra"] = {}
        try:
            line["extra"]["band_sweep"] = band_sweep(a, R, L, torch, local_rank)
        except Exception as e:                                   # noqa: BLE001  (the headline must survive)
            line["extra"]["band_sweep"] = {"error": repr(e)}
        try:
            line["extra"]["support"] = support_line(a, R, L, torch, local_rank, tasks=1 << 17, steps=100, warmup=3, cpu=not a.no_cpu)
        except Exception as e:                                   # noqa: BLE001
            line["extra"]["support"] = {"error": repr(e)}
        try:
            # the GPU-linked program is its own process with its own CUDA context: give the device back first
            del d_reads, d_off, d_tid, d_pos, d_rng, d_status, d_nseg, d_rstart, d_segoff, d_segs, d_count, flush
            R.close()
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import e2e_inline
            if e2e_inline.have_programs():
                line["vcf_wall_time"] = e2e_inline.vcf_wall_time(length=a.vcf_mb * 1_000_000, depth=30, modes=("inline",))
            else:
                line["vcf_wall_time"] = {"unavailable": "oracle/_ref programs not built (they need /root/reference at build time)"}
        except Exception as e:                                   # noqa: BLE001
            line["vcf_wall_time"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def spot_check(R, w, status, nseg, rstart, seg_off, segs, ref, count, numgaps=0):
    """`count` reads spread over the timed batch, realigned by the oracle and compared (status, start, segment words)"""
    from oracle import oracle as O
    from indelminer_b200.api import walk_segments
    p = O.default_params(6, numgaps)
    cs = ref.tobytes()
    M = w["read_len"]
    n = len(status)
    bad = 0
    idx = np.linspace(0, n - 1, count).astype(np.int64)
    for i in idx:
        o = O.realign_read(p, cs, int(w["position"][i]), int(w["range1"][i]), w["read_bases"][i * M:(i + 1) * M].tobytes())
        words = segs[seg_off[i]:seg_off[i] + nseg[i]].view(np.uint32)
        got = walk_segments(int(rstart[i]), words) if nseg[i] else []
        if int(status[i]) != o.status or got != o.segments():
            bad += 1
    return {"reads": int(count), "mismatches": int(bad), "checker": "oracle/indel_oracle.c orc_realign_read"}


def band_sweep(a, R, L, torch, dev_index):
    """D1: bands 1..129 at 2^20 alignments, every band with its own clocks record"""
    from indelminer_b200 import synth
    gops = C.c_double(0)
    if L.indelgpu_int32_peak(R._ctx, C.byref(gops)) != 0:
        raise RuntimeError(_liberr())
    out = {"tasks": a.sweep_tasks, "shape": "M=150, N=1410, 1 % substitutions, half of the reads with one 1-50 bp indel",
           "int32_peak_gops": gops.value, "int_ops_per_cell": INT_OPS_PER_CELL,
           "peak_source": "measured in this process: indelgpu_int32_peak (independent add + max chains)", "bands": []}
    t = synth.make_band_tasks(a.sweep_tasks, 0)
    packed = (t["reads"], t["read_off"], t["wins"], t["win_off"])
    for band in (1, 5, 9, 17, 33, 65, 129):
        low = (t["true_off"] - band // 2).astype(np.int32)
        up = (low + band - 1).astype(np.int32)
        sampler = ClockSampler(dev_index)
        sampler.start()
        R.band_align_batch(None, None, low, up, packed=packed, want_cigar=False)          # warm-up
        sampler.mark()
        kms, launches = [], 0
        for _ in range(3):
            res = R.band_align_batch(None, None, low, up, packed=packed, want_cigar=False)
            kms.append(L.indelgpu_last_kernel_ms(R._ctx))
            launches += 1
        clocks = sampler.stop()
        cells = int(res["cells"].sum())
        executed = cells - res["shortcut_cells"]
        k_s = float(np.mean(kms)) / 1e3
        gc = cells / k_s / 1e9
        out["bands"].append({"band": band, "kernel_ms": k_s * 1e3, "gcups": gc, "gcups_executed_cells": executed / k_s / 1e9,
                             "cells": {"forward": int(res["cells"][0]), "reverse": int(res["cells"][1]), "align": int(res["cells"][2]),
                                       "align_not_swept_shortcut": int(res["shortcut_cells"])},
                             "frac_of_int32_peak": gc * INT_OPS_PER_CELL / gops.value if gops.value else None,
                             "gpu_launches": launches, "clocks": clocks})
    return out


def run_band(a, R, L, torch, dev, peak, peak_src, dev_index=0):
    """D1: banded local_align + ALIGN + fetch_cigar on independent tasks; GCUPS of the
    kernel (CUDA events around the launch, inside the library) against the INT32 issue rate measured on
    this GPU by indelgpu_int32_peak, at INT_OPS_PER_CELL integer operations per DP cell."""
    from indelminer_b200 import synth
    t = synth.make_band_tasks(a.tasks, a.band)
    packed = (t["reads"], t["read_off"], t["wins"], t["win_off"])
    gops = C.c_double(0)
    if L.indelgpu_int32_peak(R._ctx, C.byref(gops)) != 0:
        raise RuntimeError(_liberr())
    sampler = ClockSampler(dev_index)
    sampler.start()
    for _ in range(a.warmup):
        out = R.band_align_batch(None, None, t["low"], t["up"], packed=packed, want_cigar=False)
    torch.cuda.synchronize()
    sampler.mark()
    kms, wall = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        out = R.band_align_batch(None, None, t["low"], t["up"], packed=packed, want_cigar=False)
        wall.append(time.perf_counter() - t0)
        kms.append(L.indelgpu_last_kernel_ms(R._ctx))
    clocks = sampler.stop()
    cells = int(out["cells"].sum())
    k_s = float(np.mean(kms)) / 1e3
    gcups = cells / k_s / 1e9
    line = {"metric": "banded_dp_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": k_s * 1e3, "higher_is_better": True, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"D1 band sweep: {a.tasks} alignments, M=150, N=1410, band={a.band}",
                       "timing": "CUDA events around the kernel on the library's stream"},
            "cells_per_step": {"forward": int(out["cells"][0]), "reverse": int(out["cells"][1]), "align": int(out["cells"][2]),
                               "align_not_swept_shortcut": int(out["shortcut_cells"])},
            "gcups_executed_cells": (cells - out["shortcut_cells"]) / k_s / 1e9,
            "gpu_launches": a.steps, "clocks": clocks,
            "e2e": {"value": cells / float(np.mean(wall)) / 1e9, "unit": "GCUPS",
                    "note": "pageable host buffers in (reads + windows), scores / end points / CIGAR lengths out, copies included; "
                            "the CIGAR words stay on the device"},
            "roofline": {"bound": "int32", "achieved": gcups * INT_OPS_PER_CELL, "peak": gops.value, "unit": "Gop/s",
                         "frac": gcups * INT_OPS_PER_CELL / gops.value if gops.value else None,
                         "int_ops_per_cell": INT_OPS_PER_CELL,
                         "peak_source": "measured here: indelgpu_int32_peak (independent add + max chains)"}}
    print(json.dumps(line), flush=True)




def support_line(a, R, L, torch, dev_index, tasks, steps, warmup, cpu):
    """Row f1: the known-indel support check, realign_with_indel (variant.c:1246-1424), on
    `tasks` (target, query) pairs; GCUPS of the kernel (CUDA events inside the library) against the INT32
    issue rate measured here, and the oracle's C port timed on a sample of the same pairs."""
    from indelminer_b200 import synth
    t = synth.make_support_tasks(tasks)
    pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory().numpy()      # noqa: E731  (host buffers, pinned)
    packed = (pin(t["targets"]), pin(t["target_off"]), pin(t["queries"]), pin(t["query_off"]))
    outbuf = tuple(pin(np.zeros(tasks, dtype=np.int32)) for _ in range(3))
    gops = C.c_double(0)
    if L.indelgpu_int32_peak(R._ctx, C.byref(gops)) != 0:
        raise RuntimeError(_liberr())
    sampler = ClockSampler(dev_index)
    sampler.start()
    for _ in range(warmup):
        out = R.indel_support_batch(None, None, packed=packed, out=outbuf)
    torch.cuda.synchronize()
    sampler.mark()
    kms, wall, launches = [], [], 0
    for _ in range(steps):
        t0 = time.perf_counter()
        out = R.indel_support_batch(None, None, packed=packed, out=outbuf)
        wall.append(time.perf_counter() - t0)
        kms.append(L.indelgpu_last_kernel_ms(R._ctx))
        launches += int(L.indelgpu_last_launch_count(R._ctx))
    clocks = sampler.stop()
    cells = int(out["cells"])
    k_s = float(np.mean(kms)) / 1e3
    gcups = cells / k_s / 1e9
    from oracle import oracle as O          # the checker: a sample of the timed batch compared; timed as the CPU baseline when asked
    ns = min(tasks, 4096 if cpu else 512)
    t0 = time.perf_counter()
    cc = [0]
    for k in range(ns):
        tt = t["targets"][t["target_off"][k]:t["target_off"][k + 1]].tobytes()
        q = t["queries"][t["query_off"][k]:t["query_off"][k + 1]].tobytes()
        r = O.indel_support_dp(tt, q, cells=cc)
        assert r == (int(out["subs"][k]), int(out["indels"][k]), int(out["aligned"][k])), k
    dt = time.perf_counter() - t0
    cpu_obj = {"value": cc[0] / dt / 1e9, "unit": "GCUPS", "cores": 1, "kind": "port",
               "sample": f"first {ns} pairs through oracle/indel_oracle.c orc_indel_support_dp (results compared)"} if cpu else None
    return {"metric": "indel_support_gcups", "value": gcups, "unit": "GCUPS", "n_gpus": 1, "steps": steps,
            "warmup": warmup, "ms_per_step": k_s * 1e3, "higher_is_better": True, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"f1 known-indel support: {tasks} pairs, 150 bp reads, 1-50 bp indels, target ~{int(t['target_off'][-1] / tasks)} bp",
                       "timing": "CUDA events around the kernels on the library's stream"},
            "pairs_per_s": tasks / k_s, "gpu_launches": launches, "clocks": clocks, "oracle_checked_pairs": ns,
            "e2e": {"value": cells / float(np.mean(wall)) / 1e9, "unit": "GCUPS", "pairs_per_s": tasks / float(np.mean(wall)),
                    "note": "host buffers in and out, copies included"},
            # SURVEY.md 8d's model: 10 integer operations per cell of an affine-gap recurrence, against the INT32 issue
            # rate measured in this process.  (The kernel works on 16-bit halves, two cells per instruction.)
            "roofline": {"bound": "int32", "achieved": gcups * INT_OPS_PER_CELL, "peak": gops.value, "unit": "Gop/s",
                         "frac": gcups * INT_OPS_PER_CELL / gops.value if gops.value else None,
                         "int_ops_per_cell": INT_OPS_PER_CELL,
                         "kernels": "indel_support_pack_kernel (16-bit SIMD wavefront, direction bits) + indel_support_walk_kernel",
                         "peak_source": "measured here: indelgpu_int32_peak (independent add + max chains)"},
            "cpu_baseline": cpu_obj}


def run_support(a, R, L, torch, dev_index=0):
    print(json.dumps(support_line(a, R, L, torch, dev_index, a.tasks, a.steps, a.warmup, not a.no_cpu)), flush=True)


def _liberr():
    from indelminer_b200 import lib as _lib
    return _lib.last_error()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
    else:
        run_ours(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
