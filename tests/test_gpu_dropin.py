"""GPU: the drop-in surface (FactorySimulation / Logger view / metrics DataFrames / batched ExperimentRunner) and the
sharding invariants, through the public Python API."""
import json

import numpy as np
import pandas as pd
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _ring_records(g, rid):
    sel = g["ring"] == rid
    return np.stack([g["time"][sel], g["value"][sel]], axis=1)


def test_factory_simulation_replay_matches_reference_logs_and_metrics():
    from manusim_b200.factory_sim import NewSimulation

    g = load_golden("ss_cfg1")
    s = g["scenario"]
    sim = NewSimulation(s["cfg"], s["resources"], s["products"], print_mode="none")
    sim.reset_simulation(s["seed"], None, tapes={k: g["tape_" + k] for k in "ABCD"})
    elapsed = sim.run_simulation()
    assert elapsed > 0 and sim.seed == s["seed"]
    assert sim.env.now == 1000 and isinstance(sim.env.now, int)          # run_until stays an int
    assert sim.sim_events == g["steps"]
    for rid, (variable, key) in enumerate(sim.tables.ring_names()):
        want = _ring_records(g, rid)
        li = sim.logs.log_index[variable][key]
        assert li == len(want)
        assert np.array_equal(getattr(sim.logs, variable)[key][:li], want)
    m = g["metrics"]
    for fam, df in (("products", sim.products_metrics()), ("resources", sim.resources_metrics()),
                    ("general", sim.general_metrics())):
        want = pd.DataFrame(m[fam])
        assert list(df.columns) == list(want.columns)
        assert list(df["key"]) == list(want["key"])
        for col in df.columns[1:]:
            np.testing.assert_allclose(df[col].to_numpy(dtype=float), want[col].to_numpy(dtype=float), rtol=1e-6,
                                       equal_nan=True)


def test_small_memory_size_reproduces_ring_wrap(tmp_path):
    """memory_size=100 as shipped (conf/config.yaml:22): visible counts wrap, old rows go to CSV when a path is set."""
    from manusim_b200.factory_sim import NewSimulation

    g = load_golden("ss_cfg1")
    s = g["scenario"]
    cfg = dict(s["cfg"], memory_size=100)
    sim = NewSimulation(cfg, s["resources"], s["products"], print_mode="none")
    sim.reset_simulation(s["seed"], tmp_path / "logs", tapes={k: g["tape_" + k] for k in "ABCD"})
    sim.run_simulation()
    rid = [i for i, n in enumerate(sim.tables.ring_names()) if n == ("utilization", "r09")][0]
    want = _ring_records(g, rid)
    li = sim.logs.log_index["utilization"]["r09"]
    assert li == (len(want) - 1) % 100 + 1
    saved = pd.read_csv(tmp_path / "logs" / "log_utilization_r09.csv")
    assert len(saved) == len(want) - li
    sim.save_metrics(tmp_path / "run", saved_logs=True)                   # the ExperimentRunner flow (experiment.py:59)
    got = pd.read_csv(tmp_path / "run" / "metrics_resources.csv")
    ref = pd.DataFrame(g["metrics"]["resources"])
    np.testing.assert_allclose(got["utilization"].to_numpy(), ref["utilization"].to_numpy(dtype=float), rtol=1e-5)
    assert got.columns[0] == "Unnamed: 0"                                 # pandas index column kept (quirk E14)


def test_batched_experiment_runner_tree(tmp_path):
    import yaml

    from manusim_b200 import scenarios
    from manusim_b200.experiment import ExperimentRunner, generate_run_seeds
    from manusim_b200.factory_sim import NewSimulation

    sim_cfg, res, prod, exp = scenarios.load_example("simple_scheduler")
    cfg = dict(sim_cfg, run_until=300, memory_size=5000)
    sim = NewSimulation(cfg, res, prod, print_mode="none")
    runner = ExperimentRunner(sim, number_of_runs=6, save_folder_path=tmp_path, run_name="t", save_logs=True,
                              seed=exp["exp_seed"])
    metas = runner.run_experiment()
    seeds = generate_run_seeds(exp["exp_seed"], 6)
    assert [m["seed"] for m in metas] == seeds
    for i in range(1, 7):
        rf = tmp_path / f"run_{i:03d}"
        for f in ("metrics_products.csv", "metrics_resources.csv", "metrics_general.csv", "run_info.yaml"):
            assert (rf / f).exists()
        info = yaml.safe_load((rf / "run_info.yaml").read_text())
        assert set(info) == {"run_id", "seed", "elapsed_time", "simulation_end_time", "run_folder"}
        assert info["simulation_end_time"] == 300
        assert (rf / "logs" / "log_utilization_r01.csv").exists()
        lg = pd.read_csv(rf / "logs" / "log_utilization_r01.csv")
        assert list(lg.columns) == ["time", "value", "variable", "key"]
        mr = pd.read_csv(rf / "metrics_resources.csv")
        np.testing.assert_allclose(mr["utilization"][0], lg["value"].sum() / (300 - 50), rtol=1e-6)
    assert (tmp_path / "experiment_summary.yaml").exists() and (tmp_path / "experiment_metadata.csv").exists()
    st = pd.read_csv(tmp_path / "experiment_stats_device.csv")
    u = [pd.read_csv(tmp_path / f"run_{i:03d}" / "metrics_resources.csv")["utilization"][0] for i in range(1, 7)]
    row = st[(st.variable == "utilization") & (st.key == "r01")].iloc[0]
    assert row["size"] == 6
    np.testing.assert_allclose(row["mean"], np.mean(u), rtol=1e-9)
    np.testing.assert_allclose(row["std"], np.std(u, ddof=1), rtol=1e-6)


def test_results_are_shard_invariant():
    """Replication i gives identical bits whether it runs in one 96-replication launch or in shards with rep_offset
    (keys are a function of the GLOBAL index); moments add up exactly like a single-GPU run."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=400, warmup=50)
    eng = BatchEngine(compile_scenario(cfg, res, prod))
    full = eng.run(96, seed=5)
    a = eng.run(40, seed=5, rep_offset=0)
    b = eng.run(56, seed=5, rep_offset=40)
    torch.cuda.synchronize()
    for f in ("acc_sum", "acc_cnt", "n_events", "status"):
        assert torch.equal(getattr(full, f), torch.cat([getattr(a, f), getattr(b, f)])), f
    m_full = eng.reduce(full)
    m_parts = eng.reduce(a)
    eng.reduce(b, out=m_parts)
    torch.cuda.synchronize()
    np.testing.assert_allclose(m_full.cpu().numpy(), m_parts.cpu().numpy(), rtol=1e-12)
    assert m_full[0, 2].item() == 96


def test_host_buffer_path_equals_device_path():
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=300, warmup=50, mode="onDue")
    eng = BatchEngine(compile_scenario(cfg, res, prod))
    d = eng.run(33, seed=9)
    h = eng.run_host(33, seed=9)
    torch.cuda.synchronize()
    assert h.kernel_ms > 0
    assert torch.equal(d.acc_sum.cpu(), h.acc_sum) and torch.equal(d.acc_cnt.cpu(), h.acc_cnt)
    assert torch.equal(d.n_events.cpu(), h.n_events) and torch.equal(d.status.cpu(), h.status)


def test_philox_statistics_agree_with_reference_generators():
    """Philox mode vs the reference's MT19937/PCG64 generators: per-metric means over replications must agree within
    4.5 standard errors (two-sample z, Bonferroni-safe for ~60 metrics)."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios
    from manusim_b200.metrics import metric_values
    from oracle.factory_oracle import metrics_from_trace, run_oracle
    from oracle.variates import run_seeds

    sim_cfg, res, prod, _ = scenarios.load_example("simple_scheduler")
    cfg = dict(sim_cfg, run_until=400, warmup=50)
    t = compile_scenario(cfg, res, prod, policy="max_priority")
    n_cpu, n_gpu = 48, 4096
    cpu = []
    for seed in run_seeds(777, n_cpu):
        o = run_oracle(cfg, res, prod, seed, "max_priority")
        cpu.append(metrics_from_trace(o, t.P, t.R, 400, 50)[2])
    cpu = np.array(cpu)
    eng = BatchEngine(t)
    r = eng.run(n_gpu, seed=2024)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all()
    gpu = metric_values(r.acc_sum.cpu().numpy(), r.acc_cnt.cpu().numpy(), t, 350.0)
    names = t.ring_names()
    checked = 0
    for k, (variable, key) in enumerate(names):
        if variable in ("queue", "breakdown", "lostSales", "setup"):
            continue
        a, b = cpu[:, k], gpu[:, k]
        a, b = a[~np.isnan(a)], b[~np.isnan(b)]
        if len(a) < n_cpu // 2 or len(b) < n_gpu // 2:
            continue
        # under H0 both samples share one distribution: use the large (GPU) sample's variance for both terms, so that
        # rare-event metrics whose 48 CPU replications are all zero do not get a zero standard error
        v = max(a.var(ddof=1), b.var(ddof=1))
        se = np.sqrt(v / len(a) + v / len(b))
        assert abs(a.mean() - b.mean()) <= 4.5 * se + 1e-9, (variable, key, a.mean(), b.mean(), se)
        checked += 1
    assert checked > 100


def test_pool_overflow_is_retried_with_big_pools():
    """Tiny pools overflow (status 3/4); retry_overflow re-runs exactly those replications on the big-pool twin and the
    patched results equal a run that had big pools from the start."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=600, warmup=50)
    small = BatchEngine(compile_scenario(cfg, res, prod, max_production_orders=24, max_demand_orders=24))
    big = BatchEngine(compile_scenario(cfg, res, prod, max_production_orders=250, max_demand_orders=250, fifo_capacity=256))
    a = small.run(64, seed=3)
    torch.cuda.synchronize()
    n_bad = int(((a.status >= 3) & (a.status <= 5)).sum())
    assert n_bad > 0
    assert small.retry_overflow(a, seed=3) == n_bad
    b = big.run(64, seed=3)
    torch.cuda.synchronize()
    assert (a.status == 0).all()
    assert torch.equal(a.acc_sum, b.acc_sum) and torch.equal(a.n_events, b.n_events)


def test_sweep_runner_matches_per_scenario_runs():
    """config #5 in miniature: demand-rate x buffer-size grid of the DBR model; per-scenario statistics equal those
    of running every scenario on its own."""
    import torch

    from manusim_b200 import scenarios
    from manusim_b200.factory_sim import DBRSimulation
    from manusim_b200.sweep import SweepRunner

    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    sim = dict(sim, run_until=801, warmup=100)
    base = {"simulation": sim, "resources": res, "products": prod}
    sw = SweepRunner(DBRSimulation, base, {"products.*.demand.freq.params.0": [14.0, 10.0],
                                           "products.*.shipping_buffer": [60, 120]}, reps=48, seed=11,
                     sim_kwargs={"dbr_repair": True})
    df = sw.run()
    assert sw.moments.shape[0] == 4 and sw.events > 0
    assert set(df["scenario"]) == {0, 1, 2, 3}
    # scenario 3 on its own
    _, tree = sw.points[3]
    one = DBRSimulation(tree["simulation"], tree["resources"], tree["products"], print_mode="none", dbr_repair=True)
    r = one.engine.run(48, seed=11 + 1000003 * 3)
    m = one.engine.reduce(r)
    torch.cuda.synchronize()
    np.testing.assert_allclose(sw.moments[3], m.cpu().numpy(), rtol=1e-12)
    # more demand (smaller gap) => more lost sales or later deliveries, never fewer released orders
    rel = df[(df.variable == "released") & (df.key == "produto01")].set_index("scenario")["mean"]
    assert np.isfinite(rel).all()
    # a fine grid: many scenarios with few replications each, launched round-robin on 3 streams without a per-scenario
    # synchronise, tiny pools so that some replications overflow and are re-run at the end -- same moments as scenario by
    # scenario on one stream
    grid = {"products.*.demand.freq.params.0": [16.0, 13.0, 11.0, 10.0, 9.0], "products.*.shipping_buffer": [40, 90, 140]}
    kw = {"dbr_repair": True, "max_production_orders": 40, "max_demand_orders": 12}
    fine = SweepRunner(DBRSimulation, base, grid, reps=9, seed=5, sim_kwargs=kw, streams=3)
    fine.run()
    serial = SweepRunner(DBRSimulation, base, grid, reps=9, seed=5, sim_kwargs=kw, streams=1)
    serial.run()
    assert fine.launches == 15 and fine.moments.shape[0] == 15
    np.testing.assert_allclose(fine.moments, serial.moments, rtol=1e-12, atol=1e-12)
    assert (fine.moments[:, :, 2].max(axis=1) == 9).all()          # every replication of every scenario is in the statistics


def test_sweep_over_delivery_modes_launches_kernel_family_by_kernel_family():
    """A sweep whose scenarios need DIFFERENT kernels (delivery families) groups its launches by kernel and fences the
    streams between the groups; statistics equal the scenario-by-scenario runs."""
    import torch

    from manusim_b200 import scenarios
    from manusim_b200.factory_sim import FactorySimulation
    from manusim_b200.sweep import SweepRunner

    cfg, res, prod = scenarios.jobshop20(run_until=400, warmup=50, mode="asReady")
    base = {"simulation": cfg, "resources": res, "products": prod}
    grid = {"simulation.delivery_mode": ["asReady", "onDue"], "products.*.demand.freq.params.0": [10.0, 12.0, 14.0]}
    sw = SweepRunner(FactorySimulation, base, grid, reps=40, seed=3, streams=3)
    sw.run()
    assert sw.kernel_families == 2 and sw.launches == 6
    for s in (1, 4):
        _, tree = sw.points[s]
        one = FactorySimulation(tree["simulation"], tree["resources"], tree["products"], print_mode="none")
        m = one.engine.reduce(one.engine.run(40, seed=3 + 1000003 * s))
        torch.cuda.synchronize()
        np.testing.assert_allclose(sw.moments[s], m.cpu().numpy(), rtol=1e-12)
    modes = [sw.points[s][1]["simulation"]["delivery_mode"] for s in range(6)]
    assert sorted(set(modes)) == ["asReady", "onDue"]


def test_host_path_retries_pool_overflow():
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=600, warmup=50)
    small = BatchEngine(compile_scenario(cfg, res, prod, max_production_orders=24, max_demand_orders=24))
    raw = small.run_host(48, seed=3, retry_overflow=False)
    assert int(((raw.status >= 3) & (raw.status <= 5)).sum()) > 0
    fixed = small.run_host(48, seed=3)
    assert (fixed.status == 0).all()
    big = BatchEngine(compile_scenario(cfg, res, prod, max_production_orders=250, max_demand_orders=250, fifo_capacity=256))
    ref = big.run_host(48, seed=3)
    assert torch.equal(fixed.acc_sum, ref.acc_sum) and torch.equal(fixed.n_events, ref.n_events)


def test_philox_statistics_agree_with_reference_generators_dbr():
    """Same check for the DBR model (repaired variant, P2+P3) with TBF/TTR breakdowns, lots > 1 and normal setups."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios
    from manusim_b200.metrics import metric_values
    from oracle.factory_oracle import metrics_from_trace, run_oracle
    from oracle.variates import run_seeds

    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    cfg = dict(sim, run_until=4001, warmup=800)
    cfg.pop("print_mode", None)
    t = compile_scenario(cfg, res, prod, policy="dbr", dbr_repair=True)
    n_cpu, n_gpu = 24, 2048
    cpu = np.array([metrics_from_trace(run_oracle(cfg, res, prod, s, "dbr", dbr_repair=True), t.P, t.R, 4001, 800)[2]
                    for s in run_seeds(4242, n_cpu)])
    r = BatchEngine(t).run(n_gpu, seed=99)
    torch.cuda.synchronize()
    assert (r.status.cpu().numpy() == 0).all()
    gpu = metric_values(r.acc_sum.cpu().numpy(), r.acc_cnt.cpu().numpy(), t, 3201.0)
    checked = 0
    for k, (variable, key) in enumerate(t.ring_names()):
        if variable in ("queue", "breakdown", "lostSales", "deliveredLate", "tardiness"):
            continue
        a, b = cpu[:, k], gpu[:, k]
        a, b = a[~np.isnan(a)], b[~np.isnan(b)]
        if len(a) < n_cpu // 2 or len(b) < n_gpu // 2:
            continue
        v = max(a.var(ddof=1), b.var(ddof=1))
        se = np.sqrt(v / len(a) + v / len(b))
        assert abs(a.mean() - b.mean()) <= 4.5 * se + 1e-9, (variable, key, a.mean(), b.mean(), se)
        checked += 1
    assert checked > 80


def test_dbr_simulation_dropin_replays_reference_run():
    """DBRSimulation (examples/toc_dbr) through the drop-in class, replaying the reference's recorded variates."""
    from manusim_b200.factory_sim import DBRSimulation

    g = load_golden("dbr_repaired")
    s = g["scenario"]
    sim = DBRSimulation(s["cfg"], s["resources"], s["products"], print_mode="none", dbr_repair=True)
    assert sim.contraint_resource == "r09"
    sim.reset_simulation(s["seed"], None, tapes={k: g["tape_" + k] for k in "ABCD"})
    sim.run_simulation()
    assert sim.sim_events == g["steps"]
    ref = pd.DataFrame(g["metrics"]["resources"])
    got = sim.resources_metrics()
    np.testing.assert_allclose(got["utilization"].to_numpy(dtype=float), ref["utilization"].to_numpy(dtype=float), rtol=1e-6)
    np.testing.assert_allclose(got["setup"].to_numpy(dtype=float), ref["setup"].to_numpy(dtype=float), rtol=1e-6)
    refp = pd.DataFrame(g["metrics"]["products"])
    gotp = sim.products_metrics()
    for col in ("deliveredOntime", "lostSales", "flowTime", "released"):
        np.testing.assert_allclose(gotp[col].to_numpy(dtype=float), refp[col].to_numpy(dtype=float), rtol=1e-6, equal_nan=True)


def test_device_variable_moments_equal_host_computation(tmp_path):
    """msim_reduce_variables (per run: mean over the variable's keys, then sum / sum^2 / n over runs) against numpy on the
    same accumulators; the runner writes the reference-format per-variable statistics file from it."""
    import torch

    from manusim_b200 import scenarios
    from manusim_b200.experiment import ExperimentRunner
    from manusim_b200.factory_sim import NewSimulation
    from manusim_b200.metrics import metric_values

    sim_cfg, res, prod, exp = scenarios.load_example("simple_scheduler")
    cfg = dict(sim_cfg, run_until=300, memory_size=5000)
    sim = NewSimulation(cfg, res, prod, print_mode="none")
    eng, t = sim.engine, sim.tables
    r = eng.run(37, seed=4)
    mv = eng.reduce_variables(r)
    torch.cuda.synchronize()
    vals = metric_values(r.acc_sum.cpu().numpy(), r.acc_cnt.cpu().numpy(), t, 250.0)
    P, R = t.P, t.R
    groups = [slice(m * P, (m + 1) * P) for m in range(11)] + [slice(11 * P + m * R, 11 * P + (m + 1) * R) for m in range(4)] \
        + [slice(11 * P + 4 * R + g, 11 * P + 4 * R + g + 1) for g in range(3)]
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xv = np.stack([np.nanmean(vals[:, g], axis=1) for g in groups], axis=1)
    ok = ~np.isnan(xv)
    want = np.stack([np.where(ok, xv, 0).sum(0), np.where(ok, xv ** 2, 0).sum(0), ok.sum(0).astype(float)], axis=1)
    np.testing.assert_allclose(mv.cpu().numpy(), want, rtol=1e-12, atol=1e-12)
    runner = ExperimentRunner(sim, number_of_runs=5, save_folder_path=tmp_path, save_logs=False, seed=123)
    runner.run_experiment()
    st = pd.read_csv(tmp_path / "experiment_stats_device_variables.csv")
    assert list(st["variable"])[:3] == ["deliveredOntime", "deliveredLate", "lostSales"] and len(st) == 18
    assert st["size"].between(0, 5).all() and (st["size"] == 5).sum() >= 10


def test_one_engine_on_two_streams_at_once():
    """Every launch takes its own work counter: two launches of ONE handle in flight on different streams give the
    replications each of them was asked for (msim.h, msim_run)."""
    import torch

    from manusim_b200 import BatchEngine, compile_scenario, scenarios

    cfg, res, prod = scenarios.jobshop20(run_until=600, warmup=60, mode="asReady")
    eng = BatchEngine(compile_scenario(cfg, res, prod))
    n = 3 * eng.resident_replications // 2
    ref_a, ref_b = eng.run(n, seed=5), eng.run(n, seed=6)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    a = eng.run(n, seed=5, stream=sa)
    b = eng.run(n, seed=6, stream=sb)
    torch.cuda.synchronize()
    for x, y in ((a, ref_a), (b, ref_b)):
        assert (x.status == 0).all()
        assert torch.equal(x.n_events, y.n_events) and torch.equal(x.acc_cnt, y.acc_cnt)
        assert torch.equal(x.acc_sum.view(torch.int64), y.acc_sum.view(torch.int64))
    eng.close()
