"""Batched ``ExperimentRunner``: same constructor and the same on-disk tree as the reference runner
(/root/reference/src/manusim/experiment.py:92-235), but every replication of the experiment is simulated in ONE
kernel launch per GPU instead of a Python loop / ProcessPoolExecutor.

Tree written (experiment.py:42-76, 212-235):
    <save_folder>/run_001/{metrics_products.csv, metrics_resources.csv, metrics_general.csv, run_info.yaml[, logs/log_*.csv]}
    <save_folder>/experiment_summary.yaml, experiment_metadata.csv
plus the experiment-level files the reference's post-processing (`ExperimentMetrics.save_stats`, metrics.py:921-935)
would derive from the per-run CSVs -- ``runs_metrics.csv``, ``experiment_metrics_{products,resources,general}.csv``,
``experiment_stats.csv`` -- written straight from the device results (`experiment_outputs.py`), so they exist even when
only a few (or no) run folders are materialised; and ``experiment_stats_device.csv``: the per-(variable,key) statistics
(`calculate_experiment_stats(aggregate_variables=False)`) from the allreduced {sum x, sum x^2, n} vector.

Run seeds are generated exactly like the reference (``random.Random(exp_seed).randint(0, 999999)`` per run,
utils.py:104-105 / experiment.py:138-139) and used as the replications' Philox keys.

Run i of the experiment is the same trajectory as ``FactorySimulation(..., seed=seeds[i]).run_simulation()`` (the
reference's `_run_single_simulation` is exactly reset_simulation(seed) + run_simulation(), experiment.py:55-57).

Multi-GPU: when torch.distributed is initialised the runs are sharded contiguously over ranks; every rank writes the
run folders of its own shard (run ids are global), the statistics vectors are allreduced (NCCL on GPUs, gloo in the CPU
tests of the host logic), rank 0 gathers the run metadata and the per-run metric values and writes the experiment-level
files.

No result is dropped silently: after the re-run of pool overflows every replication must have status 0, otherwise the
experiment raises (``allow_failed=True`` instead records the failed runs in experiment_summary.yaml, writes no files for
them and leaves them out of every statistic); a device log ring that wrapped is re-simulated with a ring large enough.
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
                 materialize_runs: Optional[int] = None, log_capacity: int = 1 << 18,
                 log_memory_budget: int = 8 << 30, write_runs_metrics: Optional[bool] = None,
                 allow_failed: bool = False, index_column_artifact: bool = True):
        """Reference arguments (experiment.py:92-131) + batch-specific ones:
        materialize_runs   how many run folders (run_001..) to write; default all
        log_capacity       records per replication of the device log ring (a wrapped ring is re-run larger)
        log_memory_budget  bytes of HBM the log rings of one launch may take; more logged runs => extra launches
        write_runs_metrics write runs_metrics.csv (default: when number_of_runs * K values fit 2 GB)
        allow_failed       do not raise when replications end with a non-zero status (they are left out and counted)
        index_column_artifact  reproduce the `Unnamed: 0` variable the reference's reader derives from the index column"""
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
        self.log_memory_budget = log_memory_budget
        self.write_runs_metrics = write_runs_metrics
        self.allow_failed = allow_failed
        self.index_column_artifact = index_column_artifact
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
        self.failed_runs: Dict[int, int] = {}

    def _generate_run_seeds(self) -> List[int]:
        return generate_run_seeds(self.seed, self.number_of_runs)

    # ------------------------------------------------------------------ log rings of the materialised runs
    def _stream_logs(self, eng, keys: np.ndarray, first: "object", n_first: int, wanted: int):
        """Yield (local run index, ring, time, value) for local runs [0, wanted).  `first` is the main launch's result,
        whose first `n_first` replications streamed their records; runs beyond that, and runs whose ring wrapped, are
        re-simulated (same Philox key => same trajectory) in chunks that respect the log-memory budget."""
        from .engine import decode_log

        todo = []
        cap = self.log_capacity
        counts = first.log_count.cpu().numpy() if n_first else np.zeros(0, dtype=np.int64)
        for i in range(min(n_first, wanted)):
            if int(counts[i]) > cap:
                todo.append(i)
            else:
                yield (i, *decode_log(first.log[i].cpu().numpy(), int(counts[i]))[:3])
        todo += list(range(n_first, wanted))
        while todo:
            chunk_n = max(1, int(self.log_memory_budget // (16 * cap)))
            chunk, todo = todo[:chunk_n], todo[chunk_n:]
            sub = eng.run(len(chunk), rep_seeds=np.ascontiguousarray(keys[chunk]), n_log_reps=len(chunk), log_cap=cap)
            eng.retry_overflow(sub, rep_seeds=np.ascontiguousarray(keys[chunk]))
            cnt = sub.log_count.cpu().numpy()
            if int(cnt.max()) > cap:                   # wrapped: size the ring for the longest of them and go again
                cap = 1 << int(np.ceil(np.log2(int(cnt.max()) + 1)))
                todo = chunk + todo
                continue
            for j, i in enumerate(chunk):
                yield (i, *decode_log(sub.log[j].cpu().numpy(), int(cnt[j]))[:3])

    def run_experiment(self) -> List[Dict[str, Any]]:
        import torch
        import torch.distributed as dist

        from .engine import STATUS_TEXT
        from .experiment_outputs import write_experiment_outputs
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
        want_logs = n_mat if (self.save_logs and sim.enable_event_logging) else 0
        n_log = min(want_logs, max(1, int(self.log_memory_budget // (16 * self.log_capacity)))) if want_logs else 0
        start = time.time()
        keys = np.array(seeds[lo:hi], dtype=np.uint64)
        res = eng.run(n, rep_seeds=keys, rep_offset=lo, n_log_reps=n_log, log_cap=self.log_capacity if n_log else 0)
        eng.retry_overflow(res, rep_seeds=keys)
        mom = eng.reduce(res)
        mom_var = eng.reduce_variables(res)
        if world > 1:
            dist.all_reduce(mom)
            dist.all_reduce(mom_var)
        eng.synchronize()
        batch_time = time.time() - start
        self.moments = mom.cpu().numpy()
        self.variable_moments = mom_var.cpu().numpy()
        # ---- every replication of the shard must have finished (not only the materialised ones)
        status = res.status.cpu().numpy()
        bad = np.nonzero(status != 0)[0]
        self.failed_runs = {int(lo + i + 1): int(status[i]) for i in bad}
        for i in bad:
            if status[i] in (1, 2):            # what SimPy raises inside env.run(): the reference's experiment aborts too
                raise ValueError(f"run {lo + i + 1}: {STATUS_TEXT[int(status[i])]}")
        if len(bad) and not self.allow_failed:
            hist = {STATUS_TEXT.get(int(c), int(c)): int((status == c).sum()) for c in np.unique(status[bad])}
            raise RuntimeError(f"{len(bad)} of {n} replications did not finish even with the largest pools: {hist}; "
                               f"pass allow_failed=True to leave them out (they would bias the statistics)")
        write_rm = self.write_runs_metrics
        if write_rm is None:
            write_rm = self.number_of_runs * eng.K * 8 <= (2 << 30)
        values = eng.metric_values(res).cpu().numpy() if (write_rm or n_mat) else None     # [n, K], NaN rows = failed
        self.run_metas = []
        for i in range(n_mat):
            if status[i] != 0:
                continue                           # allow_failed: no files for a run that did not finish
            run_id = lo + i + 1
            run_folder = self.save_folder_path / f"run_{run_id:03d}"
            run_folder.mkdir(exist_ok=True)
            prod, resd, gen = frames_from_values(values[i], tables)
            prod.to_csv(run_folder / "metrics_products.csv")
            resd.to_csv(run_folder / "metrics_resources.csv")
            gen.to_csv(run_folder / "metrics_general.csv")
            info = {"run_id": run_id, "seed": int(seeds[lo + i]), "elapsed_time": batch_time / max(n, 1),
                    "simulation_end_time": sim.run_until, "run_folder": str(run_folder)}
            with open(run_folder / "run_info.yaml", "w") as f:
                yaml.dump(info, f)
            self.run_metas.append(info)
        if want_logs:
            for i, ring, tm, val in self._stream_logs(eng, keys, res, n_log, want_logs):
                if status[i] != 0:
                    continue
                run_folder = self.save_folder_path / f"run_{lo + i + 1:03d}"
                lg = Logger.for_tables(tables, run_folder / "logs", max(sim.memory_size, 1))
                lg.load_stream(tables.ring_names(), ring, tm, val)
                lg.save_all_logs()
        # ---- rank 0 writes the experiment-level files from everybody's results
        metas, failed, all_values = self.run_metas, dict(self.failed_runs), values if write_rm else None
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((self.run_metas, self.failed_runs), gathered, dst=0)
            if rank == 0:
                metas = [m for part in gathered for m in part[0]]
                failed = {k: v for part in gathered for k, v in part[1].items()}
            if write_rm:
                parts = [None] * world if rank == 0 else None
                dist.gather_object(values, parts, dst=0)        # per-run values of every shard, in rank (= run) order
                if rank == 0:
                    all_values = np.concatenate(parts, axis=0)
        total_time = time.time() - start
        if rank == 0:
            self._save_experiment_summary(total_time, batch_time, world, metas, failed)
            n_ok = self.number_of_runs - len(failed)
            write_experiment_outputs(self.save_folder_path, tables, self.moments, self.variable_moments, n_ok,
                                     run_values=None, index_column_artifact=self.index_column_artifact)
            if write_rm and all_values is not None:
                from .experiment_outputs import runs_metrics_frame

                ok_rows = ~np.isnan(all_values).all(axis=1) if len(failed) else np.ones(len(all_values), dtype=bool)
                step = 20000                                     # rows are written in chunks of runs
                first = True
                for a in range(0, len(all_values), step):
                    blk = slice(a, min(a + step, len(all_values)))
                    frame = runs_metrics_frame(all_values[blk], tables, self.save_folder_path.name, first_run=a + 1,
                                               index_column_artifact=self.index_column_artifact)
                    if len(failed):
                        frame = frame[frame["run"].map(lambda r: bool(ok_rows[r - 1]))]
                    frame.to_csv(self.save_folder_path / "runs_metrics.csv", index=False, mode="w" if first else "a",
                                 header=first)
                    first = False
        return self.run_metas

    def _save_experiment_summary(self, total_time: float, batch_time: float, world: int, metas=None, failed=None):
        metas = self.run_metas if metas is None else metas
        failed = self.failed_runs if failed is None else failed
        summary = {"experiment_info": {"number_of_runs": self.number_of_runs, "n_jobs": self.n_jobs,
                                       "total_experiment_time": total_time,
                                       "average_run_time": batch_time / max(self.number_of_runs, 1),
                                       "save_folder_path": str(self.save_folder_path),
                                       "timestamp": time.strftime("%Y-%m-%d %H:%M:%S"), "gpus": world,
                                       "batch_kernel_time": batch_time,
                                       "materialized_runs": len(metas), "failed_runs": len(failed),
                                       "failed_run_status": {int(k): int(v) for k, v in failed.items()}},
                   "runs_summary": metas}
        with open(self.save_folder_path / "experiment_summary.yaml", "w") as f:
            yaml.dump(summary, f)
        pd.DataFrame(metas).to_csv(self.save_folder_path / "experiment_metadata.csv", index=False)
        st = experiment_stats_from_moments(self.moments, self.sim.tables.ring_names(), self.save_folder_path.name)
        st.to_csv(self.save_folder_path / "experiment_stats_device.csv", index=False)
        variable_stats_from_moments(self.variable_moments, self.save_folder_path.name).to_csv(
            self.save_folder_path / "experiment_stats_device_variables.csv", index=False)
