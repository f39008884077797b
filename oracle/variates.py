"""CPU restatement of the reference's variate generator (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/src/manusim/engine/utils.py:
  * derived parameters: uniform bounds (utils.py:8-12), gamma/erlang shape+scale (utils.py:15-18);
  * scalar draws from CPython ``random.Random(seed)`` then ``max(0, np.float32(v))`` (utils.py:42-67);
  * batch sums from ``numpy.random.default_rng(SeedSequence([seed, 0x4E505F42]))`` (utils.py:36-40, 69-102);
  * ``random_int`` for ExperimentRunner run seeds (utils.py:104-105, experiment.py:138-139).

Pinned by known-answer vectors generated with the reference's own class (tests/golden/variates_kat.json,
made by oracle/gen_golden.py) -- see tests/test_oracle_variates.py.

``TapeSource`` is the replay twin: it hands back pre-recorded values call by call (SURVEY.md 8(c)
"Replay tapes"), which is how GPU-recorded Philox tapes are replayed through the oracle model.
"""
from __future__ import annotations

import math
import random

import numpy as np

DIST_KINDS = ("constant", "uniform", "gamma", "erlang", "expo", "normal")


def derived_params(dist: str, params):
    """(d0, d1) in float64, computed with the reference's exact expressions."""
    if dist == "constant":
        return params[0], 0.0
    if dist == "uniform":
        c = params[1] * 2 * np.sqrt(3)          # utils.py:9  (np.float64 result)
        return params[0] - (c / 2), params[0] + (c / 2)
    if dist in ("gamma", "erlang"):
        return params[0] ** 2 / params[1] ** 2, params[1] ** 2 / params[0]
    if dist == "expo":
        return params[0], 0.0
    if dist == "normal":
        return params[0], params[1]
    raise ValueError(f"Unknowh distribution type {dist}")


class ReferenceVariates:
    """One ``DistributionGenerator`` (utils.py:21-105); optionally records what it returns."""

    def __init__(self, seed, scalar_tape=None, batch_tape=None):
        self.rng = random.Random(seed)
        ss = np.random.SeedSequence() if seed is None else np.random.SeedSequence([int(seed), 0x4E505F42])
        self.np_rng = np.random.default_rng(ss)
        self.scalar_tape = scalar_tape
        self.batch_tape = batch_tape

    def scalar(self, dist, params):
        if dist == "constant":
            v = params[0]
        elif dist == "uniform":
            a, b = derived_params(dist, params)
            v = self.rng.uniform(a, b)
        elif dist in ("gamma", "erlang"):
            k, th = derived_params(dist, params)
            v = self.rng.gammavariate(k, th)
        elif dist == "expo":
            v = self.rng.expovariate(1 / params[0])
        elif dist == "normal":
            v = self.rng.normalvariate(params[0], params[1])
        else:
            raise ValueError(f"Unknowh distribution type {dist}")
        out = max(0, np.float32(v))             # utils.py:67 : np.float32, or the int 0
        if self.scalar_tape is not None:
            self.scalar_tape.append(float(out))
        return out

    def batch_sum(self, dist, params, count):
        if count < 0:
            raise ValueError("count must be non-negative")
        if count == 0:
            total = 0.0
        elif dist == "constant":
            total = count * params[0]
        elif dist == "uniform":
            a, b = derived_params(dist, params)
            total = self.np_rng.uniform(a, b, size=count).sum()
        elif dist in ("gamma", "erlang"):
            k, th = derived_params(dist, params)
            total = self.np_rng.gamma(k, th, size=count).sum()
        elif dist == "expo":
            total = self.np_rng.exponential(scale=params[0], size=count).sum()
        elif dist == "normal":
            s = self.np_rng.normal(params[0], params[1], size=count)
            total = np.maximum(0, np.float32(s)).sum()
        else:
            raise ValueError(f"Unknowh distribution type {dist}")
        out = float(total)
        if self.batch_tape is not None:
            self.batch_tape.append(out)
        return out

    def random_int(self, start=0, end=999999):
        return self.rng.randint(a=start, b=end)


class TapeSource:
    """Replay twin of ``ReferenceVariates``: returns recorded values in call order.

    Scalar tape values are float32; 0.0 stands for the Python ``int 0`` the reference returns when
    ``np.float32(v) <= 0`` (utils.py:67) -- ``np.float32(0.0)`` itself can never be returned.
    """

    def __init__(self, scalar_values, batch_values=()):
        self.s = np.asarray(scalar_values, dtype=np.float32)
        self.b = np.asarray(batch_values, dtype=np.float64)
        self.si = 0
        self.bi = 0

    def scalar(self, dist, params):
        v = self.s[self.si]
        self.si += 1
        return max(0, np.float32(v))

    def batch_sum(self, dist, params, count):
        v = float(self.b[self.bi])
        self.bi += 1
        return v


def run_seeds(exp_seed, n):
    """ExperimentRunner._generate_run_seeds (experiment.py:128-139)."""
    g = ReferenceVariates(exp_seed)
    return [g.random_int() for _ in range(n)]
