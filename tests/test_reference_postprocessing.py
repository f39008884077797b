"""CPU (build container only): the UNMODIFIED reference post-processing (`manusim.metrics.ExperimentMetrics`,
/root/reference/src/manusim/metrics.py:207-935) must read the run tree our writers produce -- per-run metrics CSVs with
the pandas index column, `log_*.csv` files -- and arrive at the same per-run values and experiment statistics.
The tree is built from oracle replications (no GPU here); the writers are the product's (frames_from_values, Logger)."""
import numpy as np
import pandas as pd
import pytest

from oracle import run_reference as rr

pytestmark = pytest.mark.skipif(not rr.reference_available(), reason="/root/reference not mounted")


def test_reference_experiment_metrics_reads_our_tree(tmp_path):
    import yaml

    from manusim_b200 import compile_scenario, scenarios
    from manusim_b200.experiment import generate_run_seeds, stats_from_moments
    from manusim_b200.logs import Logger
    from manusim_b200.metrics import frames_from_values, metric_values
    from oracle.factory_oracle import metrics_from_trace, run_oracle

    ref = rr.load_reference()
    import importlib

    metrics_mod = importlib.import_module("manusim.metrics")
    sim, res, prod, exp = scenarios.load_example("simple_scheduler")
    cfg = dict(sim, run_until=250, warmup=50, memory_size=100000)
    t = compile_scenario(cfg, res, prod, policy="max_priority")
    seeds = generate_run_seeds(exp["exp_seed"], 4)
    vals = []
    for i, seed in enumerate(seeds, start=1):
        o = run_oracle(cfg, res, prod, seed, "max_priority")
        s, c, _ = metrics_from_trace(o, t.P, t.R, 250, 50)
        v = metric_values(s, c, t, 200.0)
        vals.append(v)
        rf = tmp_path / f"run_{i:03d}"
        rf.mkdir()
        for df, name in zip(frames_from_values(v, t), ("products", "resources", "general")):
            df.to_csv(rf / f"metrics_{name}.csv")
        lg = Logger.for_tables(t, rf / "logs", 100000)
        lg.load_stream(t.ring_names(), o["ring"].astype(np.uint32), o["time"], o["value"])
        lg.save_all_logs()
        (rf / "run_info.yaml").write_text(yaml.dump({"run_id": i, "seed": seed, "elapsed_time": 0.1,
                                                     "simulation_end_time": 250, "run_folder": str(rf)}))
    vals = np.array(vals)
    em = metrics_mod.ExperimentMetrics(tmp_path)
    em.read_runs_metrics()
    rm = em.runs_metrics
    assert set(rm.columns) >= {"variable", "key", "run", "value"}
    # per-run values the reference parsed == what we wrote
    names = t.ring_names()
    k = names.index(("utilization", "r09"))
    got = rm[(rm.variable == "utilization") & (rm.key == "r09")].sort_values("run")["value"].to_numpy(dtype=float)
    np.testing.assert_allclose(got, vals[:, k], rtol=1e-9)
    k2 = names.index(("flowTime", "produto03"))
    got2 = rm[(rm.variable == "flowTime") & (rm.key == "produto03")].sort_values("run")["value"].to_numpy(dtype=float)
    np.testing.assert_allclose(got2, vals[:, k2], rtol=1e-9, equal_nan=True)
    # experiment statistics: the reference's save_stats runs on the tree; our moments give the same per-key mean/std
    stats = em.save_stats(0.95, 0.05)
    assert (tmp_path / "experiment_stats.csv").exists() and (tmp_path / "runs_metrics.csv").exists()
    assert not stats.empty
    ok = ~np.isnan(vals)
    mom = np.stack([np.where(ok, vals, 0).sum(0), np.where(ok, vals ** 2, 0).sum(0), ok.sum(0).astype(float)], axis=1)
    st = stats_from_moments(mom)
    np.testing.assert_allclose(st["mean"][k], got.mean(), rtol=1e-12)
    np.testing.assert_allclose(st["std"][k], got.std(ddof=1), rtol=1e-9)
    # per-key statistics with the reference's own columns / formulas, from the moment vector
    from manusim_b200.experiment import experiment_stats_from_moments

    ours = experiment_stats_from_moments(mom, names, tmp_path.name, 0.95, 0.05)
    theirs = em.calculate_experiment_stats(0.95, 0.05, aggregate_variables=False)
    cols = ["size", "size_ideal", "mean", "std", "confidence", "precision", "t_crit", "error", "error_p", "ci_low", "ci_high"]
    assert list(ours.columns) == ["experiment", "variable", "key", *cols]
    merged = theirs.merge(ours, on=["variable", "key"], suffixes=("_ref", "_dev"))
    assert len(merged) >= 160
    for c in cols:
        np.testing.assert_allclose(merged[c + "_dev"].to_numpy(dtype=float), merged[c + "_ref"].to_numpy(dtype=float),
                                   rtol=1e-7, atol=1e-12, equal_nan=True, err_msg=c)
    # per-variable statistics (the reference's default experiment_stats.csv): per run mean over keys, then over runs
    from manusim_b200.experiment import variable_stats_from_moments

    P, R = t.P, t.R
    groups = [slice(m * P, (m + 1) * P) for m in range(11)] + [slice(11 * P + m * R, 11 * P + (m + 1) * R) for m in range(4)] \
        + [slice(11 * P + 4 * R + g, 11 * P + 4 * R + g + 1) for g in range(3)]
    with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
        __import__("warnings").simplefilter("ignore")
        xv = np.stack([np.nanmean(vals[:, g], axis=1) for g in groups], axis=1)          # [runs, 18]
    okv = ~np.isnan(xv)
    mom18 = np.stack([np.where(okv, xv, 0).sum(0), np.where(okv, xv ** 2, 0).sum(0), okv.sum(0).astype(float)], axis=1)
    ours_v = variable_stats_from_moments(mom18, tmp_path.name, 0.95, 0.05)
    theirs_v = stats[stats.variable.isin(ours_v.variable)]
    mv = theirs_v.merge(ours_v, on="variable", suffixes=("_ref", "_dev"))
    assert len(mv) == 18
    for c in cols:
        np.testing.assert_allclose(mv[c + "_dev"].to_numpy(dtype=float), mv[c + "_ref"].to_numpy(dtype=float),
                                   rtol=1e-7, atol=1e-12, equal_nan=True, err_msg=c)
    # the reference can also read the raw logs we saved
    em.read_logs(["utilization"])
    assert len(em.logs) > 0 and set(em.logs["variable"]) == {"utilization"}
