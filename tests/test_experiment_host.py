"""CPU: the host logic above the C ABI, driven end to end through the emulated library (tests/emu/torch_adapter.py puts
the CPU-fiber build of the kernel source behind the BatchEngine surface): the batched ExperimentRunner's run tree and
experiment-level files against the UNMODIFIED reference post-processing, run i == FactorySimulation(seed_i), failure
handling, log rings that wrap or exceed the memory budget, and the world_size-2 gloo path."""
import os
import shutil
import sys
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE / "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None or os.uname().machine != "x86_64",
                                reason="the SIMT emulation needs g++ on x86-64")

FILES = ("runs_metrics.csv", "experiment_metrics_products.csv", "experiment_metrics_resources.csv",
         "experiment_metrics_general.csv", "experiment_stats.csv")


def _sim(run_until=250, warmup=50, **cfg_kw):
    from torch_adapter import EmuTorchEngine

    from manusim_b200 import scenarios
    from manusim_b200.factory_sim import NewSimulation

    sim_cfg, res, prod, exp = scenarios.load_example("simple_scheduler")
    cfg = dict(sim_cfg, run_until=run_until, warmup=warmup, memory_size=100000, **cfg_kw)
    cfg.pop("print_mode", None)
    sim = NewSimulation(cfg, res, prod, print_mode="none")
    sim._engine = EmuTorchEngine(sim.tables)
    return sim, exp


def _assert_same_csv(a: Path, b: Path):
    x, y = pd.read_csv(a), pd.read_csv(b)
    assert list(x.columns) == list(y.columns) and x.shape == y.shape, (a.name, x.shape, y.shape)
    for c in x.columns:
        if x[c].dtype.kind in "fi":
            np.testing.assert_allclose(x[c].to_numpy(float), y[c].to_numpy(float), rtol=1e-9, atol=1e-12, equal_nan=True,
                                       err_msg=f"{a.name}:{c}")
        else:
            assert (x[c].astype(str) == y[c].astype(str)).all(), (a.name, c)


def test_runner_tree_and_experiment_files_equal_the_reference_postprocessing(tmp_path):
    """ExperimentRunner writes runs_metrics.csv / experiment_metrics_*.csv / experiment_stats.csv itself; the unmodified
    `ExperimentMetrics.save_stats` (metrics.py:921-935) run on the same tree's run folders must produce the same files."""
    from manusim_b200.experiment import ExperimentRunner, generate_run_seeds

    sim, exp = _sim()
    ours = tmp_path / "exp"
    runner = ExperimentRunner(sim, number_of_runs=5, save_folder_path=ours, save_logs=True, seed=exp["exp_seed"])
    metas = runner.run_experiment()
    assert [m["seed"] for m in metas] == generate_run_seeds(exp["exp_seed"], 5)
    for f in FILES + ("experiment_summary.yaml", "experiment_metadata.csv", "experiment_stats_device.csv"):
        assert (ours / f).is_file(), f
    assert (ours / "run_003" / "logs").is_dir() and (ours / "run_005" / "metrics_products.csv").is_file()
    rm = pd.read_csv(ours / "runs_metrics.csv")
    assert set(rm["run"]) == {1, 2, 3, 4, 5} and list(rm.columns) == ["experiment", "variable", "key", "run", "value"]
    from oracle import run_reference as rr

    if not rr.reference_available():
        pytest.skip("/root/reference not mounted: the reference's own files cannot be generated here")
    import importlib

    rr.load_reference()
    metrics_mod = importlib.import_module("manusim.metrics")
    theirs = tmp_path / "ref" / "exp"                       # same folder name = same `experiment` column
    theirs.mkdir(parents=True)
    for d in ours.glob("run_*"):
        shutil.copytree(d, theirs / d.name)
    metrics_mod.ExperimentMetrics(theirs).save_stats(0.95, 0.05)
    for f in FILES:
        _assert_same_csv(theirs / f, ours / f)


def test_experiment_files_do_not_need_run_folders(tmp_path):
    """materialize_runs=0: no run folder, the experiment-level files are still complete and equal to a full tree's."""
    from manusim_b200.experiment import ExperimentRunner

    sim, exp = _sim(run_until=150, warmup=30)
    a, b = tmp_path / "a" / "exp", tmp_path / "b" / "exp"
    ExperimentRunner(sim, 4, a, save_logs=False, seed=7).run_experiment()
    metas = ExperimentRunner(sim, 4, b, save_logs=False, seed=7, materialize_runs=0).run_experiment()
    assert metas == [] and not list(b.glob("run_*"))
    for f in FILES:
        _assert_same_csv(a / f, b / f)
    off = tmp_path / "c" / "exp"
    ExperimentRunner(sim, 4, off, save_logs=False, seed=7, index_column_artifact=False).run_experiment()
    assert "Unnamed: 0" not in set(pd.read_csv(off / "runs_metrics.csv")["variable"])
    assert "Unnamed: 0" in set(pd.read_csv(a / "runs_metrics.csv")["variable"])


def test_run_i_of_the_runner_is_factory_simulation_with_seed_i(tmp_path):
    """The seed in run_info.yaml reproduces the run: reset_simulation(seed) + run_simulation() (experiment.py:55-57)."""
    from manusim_b200.experiment import ExperimentRunner, generate_run_seeds

    sim, exp = _sim(run_until=150, warmup=30)
    ExperimentRunner(sim, 3, tmp_path / "exp", save_logs=False, seed=99).run_experiment()
    seeds = generate_run_seeds(99, 3)
    for i, seed in enumerate(seeds, start=1):
        sim.reset_simulation(seed, None)
        sim.run_simulation()
        want = pd.read_csv(tmp_path / "exp" / f"run_{i:03d}" / "metrics_resources.csv")
        got = sim.resources_metrics()
        np.testing.assert_allclose(got["utilization"].to_numpy(float), want["utilization"].to_numpy(float), rtol=1e-6)
        wantp = pd.read_csv(tmp_path / "exp" / f"run_{i:03d}" / "metrics_products.csv")
        np.testing.assert_allclose(sim.products_metrics()["flowTime"].to_numpy(float), wantp["flowTime"].to_numpy(float),
                                   rtol=1e-6, equal_nan=True)


def test_failed_replications_are_never_dropped_silently(tmp_path):
    """A replication that does not finish (here: the zero-delay event ring is too small) makes the experiment raise;
    allow_failed=True records it, writes nothing for it and leaves it out of every file."""
    from torch_adapter import EmuTorchEngine

    from manusim_b200 import compile_scenario, scenarios
    from manusim_b200.experiment import ExperimentRunner

    sim, exp = _sim(run_until=150, warmup=30)
    sim_cfg, res, prod, _ = scenarios.load_example("simple_scheduler")
    cfg = dict(sim_cfg, run_until=150, warmup=30)
    tiny = compile_scenario(cfg, res, prod, policy="max_priority", max_production_orders=6, max_demand_orders=40)
    sim._engine = EmuTorchEngine(tiny)
    with pytest.raises(RuntimeError, match="did not finish"):
        ExperimentRunner(sim, 3, tmp_path / "x", save_logs=False, seed=5).run_experiment()
    r = ExperimentRunner(sim, 3, tmp_path / "y", save_logs=False, seed=5, allow_failed=True)
    r.run_experiment()
    import yaml

    summ = yaml.safe_load((tmp_path / "y" / "experiment_summary.yaml").read_text())
    assert summ["experiment_info"]["failed_runs"] == len(r.failed_runs) > 0
    for run_id in r.failed_runs:
        assert not (tmp_path / "y" / f"run_{run_id:03d}").exists()
    rm = pd.read_csv(tmp_path / "y" / "runs_metrics.csv")
    assert not set(rm["run"]) & set(r.failed_runs)


def test_wrapped_or_over_budget_log_rings_are_rerun_not_truncated(tmp_path):
    """log_capacity far below the records of a run, and a memory budget of two rings per launch: every materialised run
    still gets its complete log files (re-simulated from its key with a ring that fits)."""
    from manusim_b200.experiment import ExperimentRunner

    sim, exp = _sim(run_until=150, warmup=30)
    full = tmp_path / "full" / "exp"
    ExperimentRunner(sim, 4, full, save_logs=True, seed=3, log_capacity=1 << 16).run_experiment()
    small = tmp_path / "small" / "exp"
    ExperimentRunner(sim, 4, small, save_logs=True, seed=3, log_capacity=512, log_memory_budget=2 * 16 * 512).run_experiment()
    assert any(l[2] > 512 for l in sim._engine.launches)           # a re-run with a larger ring happened
    for i in range(1, 5):
        fa = sorted(p.name for p in (full / f"run_{i:03d}" / "logs").glob("*.csv"))
        fb = sorted(p.name for p in (small / f"run_{i:03d}" / "logs").glob("*.csv"))
        assert fa == fb and fa
        for name in fa[:: max(1, len(fa) // 12)]:
            assert (full / f"run_{i:03d}" / "logs" / name).read_text() == (small / f"run_{i:03d}" / "logs" / name).read_text()


def _gloo_worker(rank, world, port, folder):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, str(HERE / "emu"))
    sys.path.insert(0, str(HERE))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from manusim_b200.experiment import ExperimentRunner

    sim, exp = _sim(run_until=150, warmup=30)
    ExperimentRunner(sim, 5, folder, save_logs=False, seed=11).run_experiment()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_write_the_same_experiment_as_one(tmp_path):
    """world_size 2 (gloo): contiguous shards by global run index, statistics allreduced, rank 0 gathers run metadata
    and per-run values -- the tree equals the single-rank one, run folders of both shards included."""
    import socket

    import torch.multiprocessing as mp

    from manusim_b200.experiment import ExperimentRunner

    import emu

    emu.build("plain")                                     # before the workers race to build it
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    two = tmp_path / "two" / "exp"
    mp.spawn(_gloo_worker, args=(2, port, str(two)), nprocs=2, join=True)
    sim, exp = _sim(run_until=150, warmup=30)
    one = tmp_path / "one" / "exp"
    ExperimentRunner(sim, 5, one, save_logs=False, seed=11).run_experiment()
    for f in FILES + ("experiment_stats_device.csv",):
        _assert_same_csv(one / f, two / f)
    assert sorted(p.name for p in two.glob("run_*")) == [f"run_{i:03d}" for i in range(1, 6)]
    meta = pd.read_csv(two / "experiment_metadata.csv")
    assert list(meta["run_id"]) == [1, 2, 3, 4, 5]
    for i in range(1, 6):
        _assert_same_csv(one / f"run_{i:03d}" / "metrics_products.csv", two / f"run_{i:03d}" / "metrics_products.csv")
