"""The JSON lines bench.py prints (kept under profiles/) carry every key the measurement contract names."""
import glob
import json
import os

from conftest import ROOT

BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e"}


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    assert files, pattern
    return json.load(open(files[-1]))


def test_round2_line_carries_parity_sampler_and_stop_rule_records():
    """BASELINE.json's metric is 'LVG solves/s & walker-steps/s ...; flux rel-err': the line the driver records holds all
    three, at one GPU and at several (the sampler record exercises the NCCL all-gather there)."""
    for pattern in ("r2*_bench_n1.json", "r2*_bench_n2.json"):
        d = _latest(pattern)
        assert BASE | {"roofline", "gpu_launches", "clocks", "parity", "sampler", "stop_radex"} <= set(d), pattern
        p = d["parity"]
        assert p["models"] >= 2000 and p["tolerance"] == 1e-5
        assert p["well_posed"]["within_tolerance"] == p["classes"]["well_posed"] > 0.9 * p["models"]
        assert p["well_posed"]["max_rel_err_flux"] < 1e-5 and p["well_posed"]["max_rel_err_pops"] < 1e-5
        assert sum(p["classes"].values()) == p["models"]
        s = d["sampler"]
        assert s["walkers"] == 1 << 20 and s["steps"] >= 10 and s["burn_steps"] >= 10
        assert s["walker_steps_per_s"] > 0 and s["solves_per_s"] > s["walker_steps_per_s"]
        assert 0 <= s["allgather"]["fraction"] < 0.5
        if d["n_gpus"] > 1:
            assert "NCCL" in s["allgather"]["collective"] and s["allgather"]["fraction"] > 0
        r = d["stop_radex"]
        assert r["value"] > d["value"] and r["iters_per_solve"] < d["iters_per_solve"]
        assert len(r["error_vs_fixed_point"]["flux_rel_err"]) == 5
        assert d["roofline"]["traffic"] is None or d["roofline"]["traffic"] > d["e2e"]["d2h_bytes_per_step"]
        # size-independent properties of the whole timed 2^20 batch: converged models normalised, and the same batch in a
        # random order gives the same bits for every model (nothing of the launch schedule leaks into an answer)
        q = d["properties"]
        assert q["models"] == 1 << 20 and q["converged_models"] > 0.85 * q["models"]
        assert q["max_abs_sum_xpop_minus_1_converged"] < 1e-11 and q["min_xpop"] >= 1e-20
        for k in ("permuted_batch_models_with_identical_populations", "permuted_batch_models_with_identical_brightness",
                  "permuted_batch_identical_niter_and_status"):
            assert q[k] == q["models"], k


def test_own_arm_line():
    d = _latest("r*_bench_n1.json")
    assert BASE | {"roofline", "cpu_baseline", "gpu_launches", "clocks"} <= set(d)
    assert d["metric"] == "LVG solves/s" and d["unit"] == "solves/s" and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"]
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _latest("r*_bench_reference.json")
    assert BASE | {"impl", "cpu_baseline"} <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "LVG solves/s" and d["unit"] == "solves/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
