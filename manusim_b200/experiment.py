"""Batched ``ExperimentRunner``: same constructor and the same on-disk tree as the reference runner
(/root/reference/src/manusim/experiment.py:92-235), but every replication of the experiment is simulated in ONE
kernel launch per GPU instead of a Python loop / ProcessPoolExecutor.

Tree written (experiment.py:42-76, 212-235):
    <save_folder>/run_001/{metrics_products.csv, metrics_resources.csv, metrics_general.csv, run_info.yaml[, logs/log_*.csv]}
    <save_folder>/experiment_summary.yaml, experiment_metadata.csv
plus, new, ``experiment_stats_device.csv``: per (variable,key) mean / sample std / n over runs computed from the
allreduced {sum x, sum x^2, n} vector (what ExperimentMetrics derives from the per-run CSVs, metrics.py:859-919).

Run seeds are generated exactly like the reference (``random.Random(exp_seed).randint(0, 999999)`` per run,
utils.py:104-105 / experiment.py:138-139) and used as the replications' Philox keys.

Multi-GPU: when torch.distributed is initialised the runs are sharded contiguously over ranks; rank 0 writes the
per-run folders of its own shard only unless ``gather_runs=True``; the statistics vector is allreduced (NCCL on
GPUs, gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import random
import time
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np
import pandas as pd
import yaml

from .metrics import frames_from_values, metric_values


def generate_run_seeds(exp_seed, n: int) -> List[int]:
    rng = random.Random(exp_seed)
    return [rng.randint(0, 999999) for _ in range(n)]


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous replication shard [lo, hi) of rank `rank`."""
    return (n * rank) // world, (n * (rank + 1)) // world


def stats_from_moments(mom: np.ndarray) -> pd.DataFrame:
    """[K,3] {sum x, sum x^2, n} -> mean, sample std (ddof=1), n per ring."""
    s1, s2, n = mom[:, 0], mom[:, 1], mom[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(n > 0, s1 / n, np.nan)
        var = np.where(n > 1, (s2 - n * mean * mean) / (n - 1), np.nan)
    return pd.DataFrame({"mean": mean, "std": np.sqrt(np.maximum(var, 0.0)), "size": n.astype(np.int64)})


def experiment_stats_from_moments(mom: np.ndarray, ring_names, experiment: str = "", confidence: float = 0.95,
                                  precision: float = 0.05) -> pd.DataFrame:
    """Per-(variable,key) experiment statistics with the reference's columns and formulas
    (`ExperimentMetrics.calculate_experiment_stats(aggregate_variables=False)` + `_calculate_stats`, metrics.py:837-919):
    size, size_ideal, mean, sample std, t-critical value, half-width `error`, relative error, confidence interval --
    computed from the allreduced {sum x, sum x^2, n} vector instead of from one CSV per run."""
    from scipy import stats as sps

    base = stats_from_moments(mom)
    rows = []
    for (variable, key), mean, std, size in zip(ring_names, base["mean"], base["std"], base["size"]):
        size = int(size)
        if size == 0:
            mean = std = t_crit = error = error_p = ci_low = ci_high = n_size = np.nan
        elif size < 2:
            std = t_crit = error = error_p = ci_low = ci_high = n_size = np.nan
        else:
            t_crit = sps.t.ppf(1 - ((1 - confidence) / 2), size - 1)
            error = t_crit * (std / np.sqrt(size))
            error_p = (0.0 if mean == 0 else np.nan) if (mean == 0 or np.isnan(mean)) else error / mean
            ci_low, ci_high = mean - error, mean + error
            n_size = np.square(t_crit * std / (precision * mean)) if (mean != 0 and not np.isnan(mean)) else 0.0
        rows.append({"experiment": experiment, "variable": variable, "key": key, "size": size, "size_ideal": n_size,
                     "mean": mean, "std": std, "confidence": confidence, "precision": precision, "t_crit": t_crit,
                     "error": error, "error_p": error_p, "ci_low": ci_low, "ci_high": ci_high})
    return pd.DataFrame(rows)


def variable_stats_from_moments(mom18: np.ndarray, experiment: str = "", confidence: float = 0.95,
                                precision: float = 0.05) -> pd.DataFrame:
    """`experiment_stats.csv` as the reference writes it by default (one row per variable: per run the per-key values
    are first averaged over the keys, metrics.py:809-835), from the [18,3] vector of `msim_reduce_variables`."""
    from .compiler import GENERAL_METRICS, PRODUCT_METRICS, RESOURCE_METRICS

    names = [(v, "") for v in (*PRODUCT_METRICS, *RESOURCE_METRICS, *GENERAL_METRICS)]
    return experiment_stats_from_moments(mom18, names, experiment, confidence, precision).drop(columns=["key"])


class ExperimentRunner:
    def __init__(self, simulation, number_of_runs: int, save_folder_path: Path = None, run_name: str = None,
                 save_logs: bool = True, seed: int = None, n_jobs: int = 1, simulation_factory=None,
                 materialize_runs: Optional[int] = None, log_capacity: int = 1 << 18):
        if n_jobs < 1:
            raise ValueError("n_jobs must be >= 1")
        self.sim = simulation
        self.number_of_runs = number_of_runs
        self.run_name = run_name
        self.save_logs = save_logs
        self.seed = seed
        self.n_jobs = n_jobs                      # kept for signature compatibility; the batch is the parallelism
        self.simulation_factory = simulation_factory
        self.materialize_runs = number_of_runs if materialize_runs is None else materialize_runs
        self.log_capacity = log_capacity
        if save_folder_path:
            self.save_folder_path = Path(save_folder_path)
        else:
            try:
                from hydra.core.hydra_config import HydraConfig

                self.save_folder_path = Path(HydraConfig.get().runtime.output_dir)
            except Exception as exc:
                raise ValueError("save_folder_path is required outside a Hydra run") from exc
        self.save_folder_path.mkdir(parents=True, exist_ok=True)
        self.run_metas: List[Dict[str, Any]] = []
        self.moments = None

    def _generate_run_seeds(self) -> List[int]:
        return generate_run_seeds(self.seed, self.number_of_runs)

    def run_experiment(self) -> List[Dict[str, Any]]:
        import torch
        import torch.distributed as dist

        from .engine import decode_log
        from .logs import Logger

        sim = self.sim
        eng = sim.engine
        tables = sim.tables
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        seeds = self._generate_run_seeds()
        lo, hi = shard_bounds(self.number_of_runs, world, rank)
        n = hi - lo
        n_mat = max(0, min(hi, self.materialize_runs) - lo)
        n_log = n_mat if (self.save_logs and sim.enable_event_logging) else 0
        start = time.time()
        keys = np.array(seeds[lo:hi], dtype=np.uint64)
        res = eng.run(n, rep_seeds=keys, rep_offset=lo, n_log_reps=n_log, log_cap=self.log_capacity if n_log else 0)
        eng.retry_overflow(res, rep_seeds=keys)
        mom = eng.reduce(res)
        mom_var = eng.reduce_variables(res)
        if world > 1:
            dist.all_reduce(mom)
            dist.all_reduce(mom_var)
        torch.cuda.synchronize(eng.device)
        batch_time = time.time() - start
        self.moments = mom.cpu().numpy()
        self.variable_moments = mom_var.cpu().numpy()
        status = res.status.cpu().numpy()
        acc_sum = res.acc_sum[:n_mat].cpu().numpy() if n_mat else np.zeros((0, eng.K))
        acc_cnt = res.acc_cnt[:n_mat].cpu().numpy() if n_mat else np.zeros((0, eng.K), dtype=np.int32)
        span = float(sim.run_until) - float(sim.warmup)
        self.run_metas = []
        for i in range(n_mat):
            run_id = lo + i + 1
            run_folder = self.save_folder_path / f"run_{run_id:03d}"
            run_folder.mkdir(exist_ok=True)
            if status[i] in (1, 2):
                raise ValueError(f"run {run_id}: replication aborted with a SimPy ValueError condition (status {status[i]})")
            vals = metric_values(acc_sum[i], acc_cnt[i], tables, span)
            prod, resd, gen = frames_from_values(vals, tables)
            prod.to_csv(run_folder / "metrics_products.csv")
            resd.to_csv(run_folder / "metrics_resources.csv")
            gen.to_csv(run_folder / "metrics_general.csv")
            if n_log:
                count = int(res.log_count[i])
                ring, tm, val, _ = decode_log(res.log[i].cpu().numpy(), count)
                lg = Logger.for_tables(tables, run_folder / "logs", max(sim.memory_size, 1))
                lg.load_stream(tables.ring_names(), ring, tm, val)
                lg.save_all_logs()
            info = {"run_id": run_id, "seed": int(seeds[lo + i]), "elapsed_time": batch_time / max(n, 1),
                    "simulation_end_time": sim.run_until, "run_folder": str(run_folder)}
            with open(run_folder / "run_info.yaml", "w") as f:
                yaml.dump(info, f)
            self.run_metas.append(info)
        total_time = time.time() - start
        if rank == 0:
            self._save_experiment_summary(total_time, batch_time, world)
        return self.run_metas

    def _save_experiment_summary(self, total_time: float, batch_time: float, world: int):
        summary = {"experiment_info": {"number_of_runs": self.number_of_runs, "n_jobs": self.n_jobs,
                                       "total_experiment_time": total_time,
                                       "average_run_time": batch_time / max(self.number_of_runs, 1),
                                       "save_folder_path": str(self.save_folder_path),
                                       "timestamp": time.strftime("%Y-%m-%d %H:%M:%S"), "gpus": world,
                                       "batch_kernel_time": batch_time},
                   "runs_summary": self.run_metas}
        with open(self.save_folder_path / "experiment_summary.yaml", "w") as f:
            yaml.dump(summary, f)
        pd.DataFrame(self.run_metas).to_csv(self.save_folder_path / "experiment_metadata.csv", index=False)
        st = experiment_stats_from_moments(self.moments, self.sim.tables.ring_names(), self.save_folder_path.name)
        st.to_csv(self.save_folder_path / "experiment_stats_device.csv", index=False)
        variable_stats_from_moments(self.variable_moments, self.save_folder_path.name).to_csv(
            self.save_folder_path / "experiment_stats_device_variables.csv", index=False)
