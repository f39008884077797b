"""Metric families and the wide-table completion contract of the reference
(/root/reference/src/manusim/metrics.py:29-205), restated for the host side of the engine.

The enums define WHAT is reduced (sum vs event-weighted mean, time-normalised or not, 0 vs NaN when empty);
the reductions themselves are fused into the simulation kernel (per-ring sum/count) and finished here.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Sequence

import numpy as np
import pandas as pd

from .compiler import GENERAL_METRICS, PRODUCT_METRICS, RESOURCE_METRICS, ScenarioTables


@dataclass(eq=False)
class MetricParams:
    agg: str          # "sum" | "mean"
    timebase: bool    # divided by (now - warmup) in resources_metrics (factory_sim.py:781)
    missing_fill: str = "zero"


def _sum(timebase=False):
    return MetricParams("sum", timebase, "zero")


def _mean():
    return MetricParams("mean", False, "nan")


class MetricProducts(Enum):                      # metrics.py:43-54
    deliveredOntime = _sum()
    deliveredLate = _sum()
    lostSales = _sum()
    flowTime = _mean()
    leadTime = _mean()
    tardiness = _mean()
    earliness = _mean()
    wip = _mean()
    finishedGoods = _mean()
    totalInventory = _mean()
    released = _mean()


class MetricResources(Enum):                     # metrics.py:57-61
    utilization = _sum(True)
    breakdown = _sum(True)
    setup = _sum(True)
    queue = _mean()


class MetricGeneral(Enum):                       # metrics.py:64-67
    wip_general = _mean()
    finishedGoods_general = _mean()
    totalInventory_general = _mean()


assert [m.name for m in MetricProducts] == PRODUCT_METRICS
assert [m.name for m in MetricResources] == RESOURCE_METRICS
assert [m.name for m in MetricGeneral] == GENERAL_METRICS

STANDARD_METRICS = (*MetricProducts, *MetricResources, *MetricGeneral)


def pivot_aggfunc_for_metric(metric: Enum) -> str:           # metrics.py:112-113
    return "sum" if metric.value.agg in ("sum", "count") else "mean"


def missing_fill_value(metric: Enum) -> float:               # metrics.py:116-127
    return np.nan if metric.value.agg == "mean" else 0.0


def complete_wide_metrics_df(df: pd.DataFrame, metrics: Sequence[Enum], keys: Sequence[str]) -> pd.DataFrame:
    """Columns ["key", *metric names]; one row per key in `keys` order; missing -> 0 for sum-like metrics, NaN for
    mean metrics (metrics.py:166-204)."""
    names = [m.name for m in metrics]
    fills = {m.name: missing_fill_value(m) for m in metrics}
    keys = list(keys)
    if not names or not keys:
        return pd.DataFrame(columns=["key", *names])
    work = df.copy()
    if work.empty:
        out = pd.DataFrame({"key": keys})
        for n in names:
            out[n] = fills[n]
        return out[["key", *names]]
    if "key" not in work.columns:
        work = work.reset_index(names="key")
    work = work.set_index("key")
    for n in names:
        if n not in work.columns:
            work[n] = fills[n]
    work = work.reindex(keys)
    for n in names:
        if not np.isnan(fills[n]):
            work[n] = work[n].fillna(fills[n])
    return work.reset_index()[["key", *names]]


def metric_values(acc_sum: np.ndarray, acc_cnt: np.ndarray, tables: ScenarioTables, span: float) -> np.ndarray:
    """Per-ring metric value from the fused accumulators: sum, sum/span or sum/count (NaN when empty).
    Works on [K] or [n, K] arrays.  Same definition as factory_sim.py:748-767 applied to all post-warm-up records."""
    P, R = tables.P, tables.R
    s = np.asarray(acc_sum, dtype=np.float64)
    c = np.asarray(acc_cnt).astype(np.float64)
    K = s.shape[-1]
    cls = np.full(K, 2)
    cls[: 3 * P] = 0
    cls[11 * P: 11 * P + 3 * R] = 1
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = np.where(c > 0, s / c, np.nan)
    return np.where(cls == 0, s, np.where(cls == 1, s / span, mean))


def frames_from_values(values: np.ndarray, tables: ScenarioTables):
    """[K] metric values -> the three wide DataFrames (products, resources, general) in the reference's layout."""
    P, R = tables.P, tables.R
    v = np.asarray(values, dtype=np.float64)
    prod = pd.DataFrame({"key": tables.product_names})
    for i, n in enumerate(PRODUCT_METRICS):
        prod[n] = v[i * P:(i + 1) * P]
    res = pd.DataFrame({"key": tables.resource_names})
    for i, n in enumerate(RESOURCE_METRICS):
        res[n] = v[11 * P + i * R: 11 * P + (i + 1) * R]
    gen = pd.DataFrame({"key": ["general"]})
    for i, n in enumerate(GENERAL_METRICS):
        gen[n] = v[11 * P + 4 * R + i: 11 * P + 4 * R + i + 1]
    return prod, res, gen
