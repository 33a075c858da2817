"""The JSON line bench.py prints is a contract with the driver: check the committed line of the final
build (profiles/r02_bench.json, produced by `python bench.py` on the GPU box) for every key the
contract names, the reference arm's line for its own, and the round-2 additions: the two arms'
`config` objects are equal, the band sweep and the support check are part of the default line with their own
clocks records, the end-to-end VCF wall time is in the line."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_default_line_has_every_contract_key():
    d = load("r02_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and d["dtype"] == "int32" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] >= d["steps"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert 0 < d["e2e"]["value"] < d["value"]                      # copies inside the timed region cost something
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = load("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["config"] == load("r02_bench.json")["config"]          # the driver compares the arms' configs key by key
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]


def test_round2_additions_of_the_default_line():
    d = load("r02_bench.json")
    assert d["oracle_spot_check"]["reads"] >= 4096 and d["oracle_spot_check"]["mismatches"] == 0
    bands = d["extra"]["band_sweep"]["bands"]
    assert [b["band"] for b in bands] == [1, 5, 9, 17, 33, 65, 129] and d["extra"]["band_sweep"]["tasks"] == 1 << 20
    for b in bands:
        for k in ("kernel_ms", "gcups", "gcups_executed_cells", "cells", "frac_of_int32_peak", "gpu_launches", "clocks"):
            assert k in b, (b["band"], k)
        assert b["clocks"]["sm_mhz"] and not b["clocks"]["reasons"] and b["gpu_launches"] > 0
        c = b["cells"]
        executed = c["forward"] + c["reverse"] + c["align"] - c["align_not_swept_shortcut"]
        assert 0 < b["gcups_executed_cells"] <= b["gcups"] and executed > 0
    s = d["extra"]["support"]
    assert s["clocks"]["sm_mhz"] and s["gpu_launches"] > 0 and s["oracle_checked_pairs"] >= 512 and s["roofline"]["frac"] > 0.5
    v = d["vcf_wall_time"]
    assert v["identical"] is True and v["records"] > 1000 and 0 < v["gpu_s"] < v["reference_s"]
    assert "indelgpu_realign_batch4" in d["e2e"]["entry_point"]
