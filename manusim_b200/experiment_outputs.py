"""The reference's experiment-level files, written straight from the batch results -- no per-run folders needed.

`ExperimentMetrics.save_stats` (/root/reference/src/manusim/metrics.py:921-935) walks every ``run_XXX`` folder, melts the
three per-run metrics CSVs into the long table ``runs_metrics.csv`` (columns experiment, variable, key, run, value;
metrics.py:21,497-605), averages it over runs into ``experiment_metrics_{products,resources,general}.csv``
(metrics.py:690-745) and reduces it to ``experiment_stats.csv`` (one row per variable: per run the mean over the
variable's keys, then mean / sample std / t interval over runs; metrics.py:809-919).  With 10^5..10^6 replications
per experiment the per-run folders are the bottleneck, so this module produces the same five files from what the
engine already holds: the per-replication metric values [n, K] (`msim_metric_values`), the allreduced per-ring
moments [K, 3] and per-variable moments [18, 3].

Faithful to the letter, including one artefact: the per-run CSVs are written with the pandas index column
(factory_sim.py:797-799), which the reference's reader melts into a bogus variable ``Unnamed: 0`` whose value is the
row number of the key; ``index_column_artifact=True`` (default) reproduces it so that the files are the ones the
unmodified post-processing would have written; ``False`` leaves it out.
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np
import pandas as pd

from .compiler import GENERAL_METRICS, PRODUCT_METRICS, RESOURCE_METRICS, ScenarioTables

ARTIFACT = "Unnamed: 0"
STATS_COLS = ["size", "size_ideal", "mean", "std", "confidence", "precision", "t_crit", "error", "error_p", "ci_low",
              "ci_high"]


def _stats_row(size: int, mean: float, std: float, confidence: float, precision: float) -> dict:
    """`ExperimentMetrics._calculate_stats` (metrics.py:859-919) from (n, mean, sample std)."""
    from scipy import stats as sps

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
    return {"size": size, "size_ideal": n_size, "mean": mean, "std": std, "confidence": confidence,
            "precision": precision, "t_crit": t_crit, "error": error, "error_p": error_p, "ci_low": ci_low,
            "ci_high": ci_high}


def _moment_stats(mom: np.ndarray):
    s1, s2, n = mom[:, 0], mom[:, 1], mom[:, 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(n > 0, s1 / n, np.nan)
        var = np.where(n > 1, (s2 - n * mean * mean) / (n - 1), np.nan)
    return mean, np.sqrt(np.maximum(var, 0.0)), n.astype(np.int64)


def runs_metrics_frame(values: np.ndarray, tables: ScenarioTables, experiment: str, first_run: int = 1,
                       index_column_artifact: bool = True) -> pd.DataFrame:
    """[n, K] per-replication metric values -> the long table in the reference's row order: run-major; inside a run the
    products file (variable-major in column order, keys in config order), then resources, then general."""
    n, K = values.shape
    P, R = tables.P, tables.R
    blocks = [(PRODUCT_METRICS, list(tables.product_names), 0), (RESOURCE_METRICS, list(tables.resource_names), 11 * P),
              (GENERAL_METRICS, ["general"], 11 * P + 4 * R)]
    var_col, key_col, src = [], [], []          # per-run template: variable, key, source column (-1 - i = artefact row i)
    for names, keys, base in blocks:
        nk = len(keys)
        if index_column_artifact:
            var_col += [ARTIFACT] * nk
            key_col += keys
            src += [-1 - i for i in range(nk)]
        if names is GENERAL_METRICS:
            for j, m in enumerate(names):
                var_col.append(m); key_col.append("general"); src.append(base + j)
        else:
            for j, m in enumerate(names):
                var_col += [m] * nk
                key_col += keys
                src += [base + j * nk + i for i in range(nk)]
    src = np.asarray(src)
    per_run = len(src)
    vals = np.empty((n, per_run), dtype=np.float64)
    real = src >= 0
    vals[:, real] = values[:, src[real]]
    vals[:, ~real] = (-1 - src[~real]).astype(np.float64)[None, :]
    return pd.DataFrame({"experiment": experiment,
                         "variable": np.tile(np.asarray(var_col, dtype=object), n),
                         "key": np.tile(np.asarray(key_col, dtype=object), n),
                         "run": np.repeat(np.arange(first_run, first_run + n, dtype=np.int64), per_run),
                         "value": vals.reshape(-1)})


def experiment_metrics_frames(moments: np.ndarray, tables: ScenarioTables, index_column_artifact: bool = True):
    """Mean over runs per (variable, key) (NaN runs skipped, as pandas' groupby.mean does) in the reference's three wide
    tables: key column, variables in alphabetical order (pivot_table sorts), keys sorted."""
    P, R = tables.P, tables.R
    mean, _, _ = _moment_stats(moments)
    pk, rk = list(tables.product_names), list(tables.resource_names)
    prod = pd.DataFrame({"key": pk})
    for j, m in enumerate(PRODUCT_METRICS):
        prod[m] = mean[j * P:(j + 1) * P]
    if index_column_artifact:
        # the artefact variable is unknown to the metric registry and lands in the `products` scope with ALL its keys
        prod[ARTIFACT] = np.arange(P, dtype=np.float64)
        extra = pd.DataFrame({"key": rk + ["general"], ARTIFACT: list(np.arange(R, dtype=np.float64)) + [0.0]})
        prod = pd.concat([prod, extra], ignore_index=True)
    res = pd.DataFrame({"key": rk})
    for j, m in enumerate(RESOURCE_METRICS):
        res[m] = mean[11 * P + j * R: 11 * P + (j + 1) * R]
    gen = pd.DataFrame({"key": ["general"]})
    for j, m in enumerate(GENERAL_METRICS):
        gen[m] = mean[11 * P + 4 * R + j: 11 * P + 4 * R + j + 1]
    out = []
    for df in (prod, res, gen):
        df = df.dropna(axis=1, how="all") if False else df
        # pivot_table drops variables whose every value is NaN (e.g. `queue` without log_queues)
        keep = ["key"] + sorted(c for c in df.columns if c != "key" and not df[c].isna().all())
        out.append(df.sort_values("key")[keep].reset_index(drop=True))
    return tuple(out)


def experiment_stats_frame(variable_moments: np.ndarray, tables: ScenarioTables, experiment: str, n_runs: int,
                           confidence: float = 0.95, precision: float = 0.05,
                           index_column_artifact: bool = True) -> pd.DataFrame:
    """`experiment_stats.csv` (one row per variable, metrics.py:809-835) from the [18, 3] per-variable moments."""
    mean, std, size = _moment_stats(variable_moments)
    rows = []
    for v, name in enumerate((*PRODUCT_METRICS, *RESOURCE_METRICS, *GENERAL_METRICS)):
        rows.append({"experiment": experiment, "variable": name,
                     **_stats_row(int(size[v]), float(mean[v]), float(std[v]), confidence, precision)})
    if index_column_artifact:
        # per run: mean of the row numbers over all P + R + 1 keys -- the same number in every run
        idx = np.concatenate([np.arange(tables.P), np.arange(tables.R), [0.0]])
        rows.append({"experiment": experiment, "variable": ARTIFACT,
                     **_stats_row(n_runs, float(idx.mean()), 0.0, confidence, precision)})
    return pd.DataFrame(rows)[["experiment", "variable", *STATS_COLS]]


def write_experiment_outputs(folder: Path, tables: ScenarioTables, moments: np.ndarray, variable_moments: np.ndarray,
                             n_runs: int, run_values: Optional[np.ndarray] = None, first_run: int = 1,
                             confidence: float = 0.95, precision: float = 0.05, index_column_artifact: bool = True,
                             append_runs: bool = False) -> dict:
    """Write experiment_stats.csv, experiment_metrics_{products,resources,general}.csv and (when `run_values` [n, K] is
    given) runs_metrics.csv into `folder`.  `append_runs` appends a further chunk of runs to runs_metrics.csv."""
    folder = Path(folder)
    folder.mkdir(parents=True, exist_ok=True)
    name = folder.name
    out = {}
    if run_values is not None:
        rm = runs_metrics_frame(np.asarray(run_values, dtype=np.float64), tables, name, first_run, index_column_artifact)
        rm.to_csv(folder / "runs_metrics.csv", index=False, mode="a" if append_runs else "w", header=not append_runs)
        out["runs_metrics"] = rm
    frames = experiment_metrics_frames(moments, tables, index_column_artifact)
    for df, scope in zip(frames, ("products", "resources", "general")):
        df.to_csv(folder / f"experiment_metrics_{scope}.csv", index=False)
        out[f"experiment_metrics_{scope}"] = df
    st = experiment_stats_frame(variable_moments, tables, name, n_runs, confidence, precision, index_column_artifact)
    st.to_csv(folder / "experiment_stats.csv", index=False)
    out["experiment_stats"] = st
    return out
