"""CPU: the bench lines recorded on the B200 (profiles/r01_bench_lines.jsonl) carry every key of the bench contract,
and bench.py's argument defaults / workload specs are importable without a GPU."""
import json

from conftest import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e"}


def test_recorded_bench_lines_follow_the_contract():
    lines = [json.loads(l) for l in (ROOT / "profiles" / "r01_bench_lines.jsonl").read_text().splitlines() if l.strip()]
    assert lines
    ours = [l for l in lines if l.get("impl") != "reference"]
    ref = [l for l in lines if l.get("impl") == "reference"]
    assert ours and ref
    for l in ours:
        assert REQUIRED <= set(l), REQUIRED - set(l)
        assert "workload" in l["config"] and "model" not in l["config"]
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(l["e2e"])
        assert l["e2e"]["h2d_bytes_per_step"] > 0 and l["e2e"]["d2h_bytes_per_step"] > 0
        assert l["gpu_launches"] > 0 and l["vs_baseline"] is None and l["dtype"] == "f32"
        r = l["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(l["clocks"])
        assert not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    for l in ref:
        assert l["impl"] == "reference" and l["e2e"]["h2d_bytes_per_step"] == 0 and l["e2e"]["d2h_bytes_per_step"] == 0
        assert {"value", "cores", "kind", "sample"} <= set(l["cpu_baseline"]) and l["cpu_baseline"]["kind"] == "port"
    final = [l for l in ours if "FINAL" in l.get("_note", "")][0]
    assert final["n_gpus"] == 1 and final["cpu_baseline"]["cores"] >= 1
    assert final["value"] > 1e9          # the BASELINE.json target on the 20-resource breakdown job shop


def test_bench_workloads_compile_on_cpu():
    import bench
    from manusim_b200 import compile_scenario

    for name in ("jobshop20", "simple_det", "toc_dbr", "toc_dbr_shipped"):
        spec = bench.workload_spec(name)
        for (cfg, res, prod), policy, phase, period in spec["parts"]:
            t = compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False))
            assert t.n_rings == 11 * t.P + 4 * t.R + 3
    keys = bench.rep_keys(20260101, __import__("numpy").arange(4))
    assert len(set(keys.tolist())) == 4
