"""End-to-end parity at the sizes BASELINE.json's configs name (configs 3, 4, 5; configs 1 and 2 are in
tests/test_e2e_vcf.py).

The unmodified reference is single-threaded and spends ~6 ms per candidate read in strlen(contig) on a 64 Mb
contig (alignment.c:771): config 3 takes it ~100 CPU-minutes, which no GPU-box budget covers.  So the reference
was run ONCE in the build container (CPU time is not metered there) on data sets written by the deterministic
generator oracle/_ref/synth_bam, and the md5 of every VCF it printed is committed:
    tests/golden/cfg3_reference.json   tools/cfg_reference_run.sh    64 Mb, 30x, 8 `-c` regions (indelminer.c:536-542)
    tests/golden/cfg4_reference.json   tools/cfg4_reference_run.sh   tumor / normal 16 Mb, 30x / 30x, `-q 0 -a -e 1`
    tests/golden/cfg5_reference.json   tools/cfg5_reference_run.sh   24 contigs (GRCh38 / 100), 30x, whole + two `-c` contigs
Here the same generator writes the same BAM on the GPU box (its md5 is checked first), the reference PROGRAM with
the GPU alignment path (oracle/_ref/indelminer_gpu[_annot], INDELGPU_MODE=inline) runs the same commands, and the
VCFs must hash the same.  Wall times go to gpurun_out/ for profiles/.

config 5 at its named size (3.1 Gb, 620 M reads) exists only at the function level: test_cfg5_reference_3_1_gb
uploads a 3.1 Gb / 24-contig reference on the GPU and checks candidates across all contigs against the oracle."""
import hashlib
import json
import os
import subprocess
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REFDIR = os.path.join(ROOT, "oracle", "_ref")
OUTDIR = os.path.join(ROOT, "gpurun_out")
pytestmark = pytest.mark.gpu


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def need(golden, *exes):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for e in exes:
        if not os.path.exists(os.path.join(REFDIR, e)):
            pytest.skip(f"oracle/_ref/{e} not built (needs /root/reference)")
    p = os.path.join(GOLD, golden)
    if not os.path.exists(p):
        pytest.skip(f"tests/golden/{golden} not generated yet")
    from indelminer_b200 import build
    build.build()
    with open(p) as f:
        return json.load(f)


def gen(cwd, args):
    t0 = time.perf_counter()
    subprocess.check_call([os.path.join(REFDIR, "synth_bam")] + args, cwd=cwd, stdout=subprocess.DEVNULL)
    return time.perf_counter() - t0


def run_to_file(exe, args, cwd, out, env=None):
    t0 = time.perf_counter()
    with open(os.path.join(cwd, out), "w") as f:
        r = subprocess.run([os.path.join(REFDIR, exe)] + args, cwd=cwd, stdout=f, stderr=subprocess.PIPE, text=True,
                           env=dict(os.environ, INDELGPU_MODE="inline", **(env or {})), timeout=1800)
    assert r.returncode == 0, r.stderr[-2000:]
    return time.perf_counter() - t0, r.stderr


def note(name, obj):
    os.makedirs(OUTDIR, exist_ok=True)
    with open(os.path.join(OUTDIR, name), "w") as f:
        json.dump(obj, f, indent=1)


def test_cfg3_full_size_64mb_30x(tmp_path):
    """BASELINE config 3 at its named size: 64 Mb contig, 30x 2x150 bp (12.8 M BAM records, ~1 M candidate reads,
    32 k planted indels), the eight `-c` region instances the reference was run as, all on one GPU"""
    g = need("cfg3_reference.json", "indelminer_gpu", "synth_bam")
    d = str(tmp_path)
    t_gen = gen(d, g["generator"].split()[1:])
    assert md5(os.path.join(d, "d.bam")) == g["bam_md5"], "the generator did not reproduce the BAM the reference was run on"
    t0 = time.perf_counter()
    procs = []
    for i, reg in enumerate(g["regions"]):
        f = open(os.path.join(d, f"gpu_{i}.vcf"), "w")
        p = subprocess.Popen([os.path.join(REFDIR, "indelminer_gpu"), "-i", "d.config", "-c", reg["region"], "d.fa", "s=d.bam"], cwd=d,
                             stdout=f, stderr=subprocess.DEVNULL, env=dict(os.environ, INDELGPU_MODE="inline"))
        procs.append((p, f))
    for p, f in procs:
        assert p.wait(timeout=1800) == 0
        f.close()
    wall = time.perf_counter() - t0
    records = 0
    for i, reg in enumerate(g["regions"]):
        path = os.path.join(d, f"gpu_{i}.vcf")
        assert md5(path) == reg["vcf_md5"], reg["region"]
        records += reg["records"]
    # one instance alone over the whole contig: the single-command wall time (no reference VCF exists for it)
    t_whole, log = run_to_file("indelminer_gpu", ["-i", "d.config", "d.fa", "s=d.bam"], d, "whole.vcf")
    whole_records = sum(1 for ln in open(os.path.join(d, "whole.vcf")) if not ln.startswith("#"))
    assert abs(whole_records - records) <= 16                # region edges may split or duplicate a handful of calls
    note("r02_cfg3_full_size.json", dict(config="cfg3: 64 Mb, 30x, 2x150 bp", bam_records=g["generated"]["records"], generate_s=round(t_gen, 1),
                                          regions=len(g["regions"]), vcf_records=records, identical_md5=True,
                                          gpu_wall_s_8_region_processes_one_gpu=round(wall, 2),
                                          gpu_wall_s_one_process_whole_contig=round(t_whole, 2), whole_contig_records=whole_records,
                                          reference_wall_s_8_region_processes=max(r["reference_seconds"] for r in g["regions"]),
                                          reference_cpu_s_total=sum(r["reference_seconds"] for r in g["regions"]),
                                          reference_note=g["note"], gpu_log=[ln for ln in log.splitlines() if "inline mode" in ln]))


def test_cfg4_tumor_normal_30x(tmp_path):
    """BASELINE config 4 at its named depth: tumor 30x called, then `-q 0 -a -e 1 ref.fa tumor.vcf normal=normal.bam`
    with the normal at 30x; attempt_pe_alignment prefetched per block, realign_with_indel batched per known variant"""
    g = need("cfg4_reference.json", "indelminer_gpu", "indelminer_gpu_annot", "synth_bam")
    d = str(tmp_path)
    for cmd in g["generator"]:
        gen(d, cmd.split()[1:])
    assert md5(os.path.join(d, "tumor.bam")) == g["tumor_bam_md5"] and md5(os.path.join(d, "normal.bam")) == g["normal_bam_md5"]
    t1, _ = run_to_file("indelminer_gpu", ["-i", "tumor.config", "tumor.fa", "tumor=tumor.bam"], d, "tumor.vcf")
    assert md5(os.path.join(d, "tumor.vcf")) == g["tumor_vcf_md5"]
    t2, log = run_to_file("indelminer_gpu_annot", ["-q", "0", "-a", "-e", "1", "-i", "normal.config", "normal.fa", "tumor.vcf", "normal=normal.bam"],
                          d, "annotated.vcf")
    assert md5(os.path.join(d, "annotated.vcf")) == g["annotated_vcf_md5"]
    note("r02_cfg4_tumor_normal.json", dict(config="cfg4: tumor / normal 16 Mb, 30x / 30x, -q 0 -a -e 1", identical_md5=True,
                                             tumor_records=g["tumor_records"], annotated_tagged=g["annotated_tagged"],
                                             gpu_wall_s=dict(call_tumor=round(t1, 2), annotate=round(t2, 2)),
                                             reference_wall_s=dict(call_tumor=g["tumor_reference_seconds"], annotate=g["annotate_reference_seconds"]),
                                             gpu_log=[ln for ln in log.splitlines() if "inline mode" in ln]))


def test_cfg5_shape_24_contigs_region_sharded(tmp_path):
    """config 5's shape (24 contigs sized like GRCh38 / 100, 30x): the whole genome in one process, and the
    reference's own region sharding -- one process per `-c` contig, each with its own BAM handle, index and GPU context"""
    g = need("cfg5_reference.json", "indelminer_gpu", "synth_bam")
    d = str(tmp_path)
    gen(d, g["generator"].split()[1:])
    assert md5(os.path.join(d, "g.bam")) == g["bam_md5"]
    times = {}
    for name, run in g["runs"].items():
        times[name], _ = run_to_file("indelminer_gpu", ["-i", "g.config"] + run["args"] + ["g.fa", "s=g.bam"], d, name + ".vcf")
        assert md5(os.path.join(d, name + ".vcf")) == run["vcf_md5"], name
    note("r02_cfg5_shape.json", dict(config="cfg5 shape: 24 contigs (GRCh38 / 100), 30x", identical_md5=True,
                                      records={k: v["records"] for k, v in g["runs"].items()}, gpu_wall_s={k: round(v, 2) for k, v in times.items()},
                                      reference_wall_s_all=g["runs"]["all"]["reference_seconds"]))


def test_cfg5_reference_3_1_gb(gpu, oracle):
    """config 5 at its named reference size: 24 contigs, 3.1 Gb in all, resident on ONE GPU (raw + 2-bit packed: every
    GPU of the box holds the whole genome, so region shards need no reference exchange).  Candidates drawn on every
    contig -- including positions past 2^31 in the concatenated reference -- are realigned in one batch and compared
    with the oracle read by read."""
    import torch
    from indelminer_b200 import synth
    mb = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51, 156, 57]
    free, _total = torch.cuda.mem_get_info()
    if free < 12 << 30:
        pytest.skip("needs 12 GB of free device memory")
    rng = np.random.default_rng(5)
    contigs = []
    for k, m in enumerate(mb):
        L = m * 1_000_000
        # uniform ACGT, cheap to make: a 1 Mb random tile repeated with a per-contig shift would create exact repeats
        # (ties everywhere), so draw every contig in full
        contigs.append(synth.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)])
    assert sum(len(c) for c in contigs) > 3_000_000_000
    R = gpu.Realigner()
    t0 = time.perf_counter()
    R.set_reference([np.ascontiguousarray(c) for c in contigs])
    t_up = time.perf_counter() - t0
    per = 1500
    parts = []
    for t, c in enumerate(contigs):
        w = synth.make_candidates(c, per, seed=500 + t)
        w["tid"][:] = t
        parts.append(w)
    M = parts[0]["read_len"]
    bases = np.concatenate([w["read_bases"] for w in parts])
    tid = np.concatenate([w["tid"] for w in parts]); pos = np.concatenate([w["position"] for w in parts]); rg = np.concatenate([w["range1"] for w in parts])
    n = len(tid)
    res = R.attempt_pe_alignment_batch(None, tid, pos, rg, packed=(bases, np.arange(n + 1, dtype=np.int64) * M))
    p = oracle.default_params()
    nsplit = 0
    for t, c in enumerate(contigs):
        cs = c.tobytes()
        for i in range(t * per, (t + 1) * per, 5):
            o = oracle.realign_read(p, cs, int(pos[i]), int(rg[i]), bases[i * M:(i + 1) * M].tobytes())
            assert int(res.status[i]) == o.status, (t, i)
            assert res.segments(i) == o.segments(), (t, i)
            nsplit += o.status == 6
    assert nsplit > 24 * 100
    used = free - torch.cuda.mem_get_info()[0]
    note("r02_cfg5_reference_3_1gb.json", dict(contigs=24, bases=int(sum(len(c) for c in contigs)), upload_s=round(t_up, 2),
                                                device_bytes_used=int(used), reads_checked=n // 5, split=int(nsplit)))
    R.close()


def test_cfg5_full_reference_size_end_to_end(tmp_path):
    """config 5 at its named REFERENCE size through the whole program: 24 contigs with the lengths of GRCh38 chr1..22, X, Y
    (3.1 Gb; 248 Mb contigs), one process over all of them, `-e 1`, at 0.05x coverage -- the depth at which the unmodified
    reference (an hour for this data set: ~20 ms of strlen per candidate read on a 248 Mb contig) could be run once in the
    build container (tools/cfg5_full_reference_run.sh).  Every contig is uploaded when its first read arrives."""
    g = need("cfg5_full_reference.json", "indelminer_gpu", "synth_bam")
    if os.statvfs(str(tmp_path)).f_bavail * os.statvfs(str(tmp_path)).f_frsize < 6 << 30:
        pytest.skip("needs 6 GB of scratch space for the 3.1 Gb FASTA")
    d = str(tmp_path)
    t_gen = gen(d, g["generator"].split()[1:])
    assert md5(os.path.join(d, "g.bam")) == g["bam_md5"]
    t, log = run_to_file("indelminer_gpu", ["-e", "1", "-i", "g.config", "g.fa", "s=g.bam"], d, "all.vcf", dict(INDELGPU_VERBOSE="1"))
    assert md5(os.path.join(d, "all.vcf")) == g["vcf_md5"]
    note("r02_cfg5_full_reference_size.json", dict(config="cfg5 reference size: 3.1 Gb in 24 contigs, 0.05x, -e 1, one process", identical_md5=True,
                                                    bam_records=g["generated"]["records"], vcf_records=g["records"], generate_s=round(t_gen, 1),
                                                    gpu_wall_s=round(t, 1), reference_wall_s=g["reference_seconds"],
                                                    gpu_log=[ln for ln in log.splitlines() if "inline mode" in ln]))
