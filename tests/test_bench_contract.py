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


def test_round2_bench_lines_carry_the_reproducible_roofline():
    """Round-2 lines: e2e is the same step as `value` (log rings on, all steps), the roofline's ncu-derived figures come
    from profiles/kernel_counters.json (named in the line), the CPU arm carries the port-vs-reference calibration."""
    lines = [json.loads(l) for l in (ROOT / "profiles" / "r02_bench_lines.jsonl").read_text().splitlines() if l.strip()]
    ours = [l for l in lines if l.get("impl") != "reference"]
    ref = [l for l in lines if l.get("impl") == "reference"]
    assert len(ours) >= 4 and ref
    for l in ours:
        assert REQUIRED <= set(l)
        assert l["e2e"]["steps"] == l["steps"] and "device rings" in l["e2e"]["logs"]
        assert l["e2e"]["h2d_bytes_per_step"] > 0 and l["e2e"]["d2h_bytes_per_step"] > 0 and l["gpu_launches"] > 0
        r = l["roofline"]
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["bound"] == "hbm"
        assert all("waves" in p for p in r["kernel_by_part"])
        assert not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    final = [l for l in ours if "FINAL" in l["_note"]][0]
    assert final["value"] > 6e9 and final["e2e"]["value"] > 6e9 and final["failed_replications"] == 0
    assert "kernel_counters.json" in final["roofline"]["traffic_source"]
    assert final["roofline"]["issue_slots"]["warp_instr_per_event"] is not None
    assert final["cpu_baseline"]["calibration"]["port_over_unmodified_reference"] > 1.0
    counters = json.loads((ROOT / "profiles" / "kernel_counters.json").read_text())
    parts = counters["workloads"]["jobshop20"]["parts"]
    assert len(parts) == 2 and all(p["warp_instr_per_event"] and p["dram_bytes_per_algorithmic_byte"] for p in parts)
    assert ref[0]["cpu_baseline"]["kind"] == "port" and ref[0]["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port as multiprocessing replications on the host cores) on a tiny horizon."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--run-until", "300", "--cpu-workers", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "events/s"
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert "engine" in line["cpu_baseline"] and line["e2e"]["value"] == line["value"]


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


def test_batch_is_sized_in_whole_waves_of_every_part():
    """bench.py sizes a step as whole waves of every workload part's kernel (wave quantisation)."""
    import bench

    sm = 148
    assert bench.whole_wave_batch([36 * sm, 32 * sm], 8.0) == 8 * 36 * sm == 9 * 32 * sm      # config #4 on a B200
    assert bench.whole_wave_batch([36 * sm, 32 * sm], 4.0) == 4 * 36 * sm                     # 8 / 9 waves would double a step
    assert bench.whole_wave_batch([36 * sm, 30 * sm], 4.0) == 5 * 36 * sm == 6 * 30 * sm
    assert bench.whole_wave_batch([36 * sm], 4.0) == 4 * 36 * sm                             # one part: nothing to round
    assert bench.whole_wave_batch([36 * sm, 36 * sm], 2.0) == 2 * 36 * sm
    assert bench.whole_wave_batch([36 * sm, 31 * sm], 1.0) == 36 * sm                        # common multiple too far: keep
