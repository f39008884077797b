"""Scenario-sweep front end (BASELINE config #5): a grid of config overrides, every grid point a
scenario with its own compiled tables, `reps` replications each, all on the GPU(s).

The reference only sweeps through Hydra multirun (examples/*/conf/config.yaml:25-30; experiment.py:16-26 forces
n_jobs=1 there), i.e. one OS process per grid point running replications one after another.  Here the flattened
index space (scenario, replication) is split contiguously over the ranks; each rank launches one batch per scenario
slice it owns and the per-scenario {sum x, sum x^2, n} moments are allreduced once at the end.

Overrides address the reference's config tree with dotted paths, e.g.
    {"products.*.demand.freq.params.0": [10 / 0.7, 10 / 0.85, 10.0], "products.*.shipping_buffer": [50, 100, 150]}
(`*` fans out over every key at that level).
"""
from __future__ import annotations

import copy
import itertools
from typing import Any, Dict, List, Mapping, Sequence, Tuple

import numpy as np
import pandas as pd

from .experiment import shard_bounds, stats_from_moments


def _set_path(tree, path: Sequence[str], value):
    key = path[0]
    if key == "*":
        keys = list(tree.keys()) if isinstance(tree, Mapping) else range(len(tree))
        for k in keys:
            if len(path) == 1:
                tree[k] = value
            else:
                _set_path(tree[k], path[1:], value)
        return
    if isinstance(tree, list):
        key = int(key)
    if len(path) == 1:
        tree[key] = value
    else:
        _set_path(tree[key], path[1:], value)


def expand_grid(base: Dict[str, Any], overrides: Dict[str, Sequence[Any]]) -> List[Tuple[Dict[str, Any], Dict[str, Any]]]:
    """base = {"simulation":..., "resources":..., "products":...}; returns [(point, config tree)] in itertools.product
    order of the override lists (first key slowest)."""
    names = list(overrides.keys())
    out = []
    for combo in itertools.product(*[overrides[n] for n in names]):
        tree = copy.deepcopy(base)
        for name, value in zip(names, combo):
            _set_path(tree, name.split("."), value)
        out.append((dict(zip(names, combo)), tree))
    return out


def kernel_family(sim) -> Tuple[int, int, int, int]:
    """Which simulation kernel a scenario launches: policy, delivery family, queue records, constant layout."""
    t = sim.tables
    return (int(t.policy), int(t.delivery_mode), int(bool(t.log_queues)), int(sim.engine.info.get("const_layout", 0)))


def flat_slices(n_scenarios: int, reps: int, world: int, rank: int) -> List[Tuple[int, int, int]]:
    """[(scenario, first replication, count)] owned by `rank` when the flattened (scenario, replication) space is
    split contiguously over `world` ranks."""
    lo, hi = shard_bounds(n_scenarios * reps, world, rank)
    out = []
    s = lo // reps if reps else 0
    while lo < hi:
        first = lo - s * reps
        cnt = min(reps - first, hi - lo)
        out.append((s, first, cnt))
        lo += cnt
        s += 1
    return out


class SweepRunner:
    def __init__(self, sim_cls, base: Dict[str, Any], overrides: Dict[str, Sequence[Any]], reps: int, seed: int = 0,
                 sim_kwargs: Dict[str, Any] | None = None, streams: int | None = None):
        """`streams` (default: as many as it takes to keep two waves of resident replications in flight, at most 64):
        scenario slices are launched round-robin on this many CUDA streams and nobody waits for a single
        scenario: the persistent grids of consecutive scenarios overlap, so a scenario's tail wave (and a fine grid's tiny
        launches) run next to the following scenarios' bodies.  The host synchronises once, after the last launch; pool
        overflows are re-run then (the reduction skips failed replications, a re-run adds its own contribution)."""
        self.sim_cls = sim_cls
        self.points = expand_grid(base, overrides)
        self.reps = reps
        self.seed = seed
        self.sim_kwargs = sim_kwargs or {}
        self.streams = None if streams is None else max(1, int(streams))
        self.moments = None
        self.events = 0
        self.launches = 0
        self.reruns = 0
        self.kernel_families = 1

    def _new_sim(self, s):
        _, tree = self.points[s]
        return self.sim_cls(tree["simulation"], tree["resources"], tree["products"], print_mode="none", **self.sim_kwargs)

    def run(self) -> pd.DataFrame:
        import torch
        import torch.distributed as dist

        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        slices = flat_slices(len(self.points), self.reps, world, rank)
        first_sim = self._new_sim(slices[0][0] if slices else 0)
        eng0 = first_sim.engine
        dev = eng0.device
        names = first_sim.tables.ring_names()
        mom = torch.zeros((len(self.points), eng0.K, 3), dtype=torch.float64, device=dev)
        ev_total = torch.zeros((), dtype=torch.int64, device=dev)
        main = torch.cuda.current_stream(dev)
        if self.streams is None:
            resident = eng0.info["warps_per_block"] * eng0.info["blocks_per_sm"] * eng0.info["n_sms"]
            per_slice = max(1, max((c for _, _, c in slices), default=1))
            self.streams = int(min(64, max(2, -(-2 * resident // per_slice))))
        pool = [torch.cuda.Stream(dev) for _ in range(min(self.streams, max(1, len(slices))))]
        start = torch.cuda.Event()
        start.record(main)
        for st in pool:
            st.wait_event(start)
        pending = []                                   # (sim, result, scenario, first, count): kept alive until the sync
        sims = {slices[0][0]: first_sim} if slices else {}
        for s, _, _ in slices:
            if s not in sims:
                sims[s] = self._new_sim(s)
        # Slices that run the SAME kernel share the SMs on concurrent streams; slices of DIFFERENT kernels (a sweep over
        # delivery_mode, or shapes on and off the constant layout) must not: two instruction streams per SM cost an
        # instruction-fetch-bound kernel 6-8 % (profiles/r02_waves_sweep_t.jsonl).  Launch kernel family by kernel family.
        slices = sorted(slices, key=lambda sl: (kernel_family(sims[sl[0]]), sl[0], sl[1]))
        family = None
        for i, (s, first, cnt) in enumerate(slices):
            sim = sims[s]
            if kernel_family(sim) != family:
                if family is not None:
                    fence = [torch.cuda.Event() for _ in pool]
                    for ev, q in zip(fence, pool):
                        ev.record(q)
                    for q in pool:
                        for ev in fence:
                            q.wait_event(ev)
                    self.kernel_families += 1
                family = kernel_family(sim)
            st = pool[i % len(pool)]
            # replication key = f(seed of the scenario, replication index within the scenario): shard-invariant
            with torch.cuda.stream(st):
                res = sim.engine.run(cnt, seed=self.seed + 1000003 * s, rep_offset=first, stream=st)
                sim.engine.reduce(res, out=mom[s], stream=st)
            self.launches += 1
            pending.append((sim, res, s, first, cnt))
        for st in pool:
            done = torch.cuda.Event()
            done.record(st)
            main.wait_event(done)
        torch.cuda.synchronize(dev)
        # re-run what overflowed the shared-memory pools (few replications, largest pools): the first reduction skipped
        # them, so the re-run's own result adds exactly their contribution
        for sim, res, s, first, cnt in pending:
            ctx = sim.engine.retry_launch(res, seed=self.seed + 1000003 * s, rep_offset=first)
            if ctx is not None:
                self.reruns += sim.engine.retry_collect(ctx)
                sim.engine.reduce(ctx[3], out=mom[s])
        for sim, res, s, first, cnt in pending:
            ev_total += res.n_events.sum()
        if world > 1:
            dist.all_reduce(mom)
            dist.all_reduce(ev_total)
        torch.cuda.synchronize(dev)
        self.moments = mom.cpu().numpy()
        self.events = int(ev_total.item())
        frames = []
        for s, (point, _) in enumerate(self.points):
            st = stats_from_moments(self.moments[s])
            st.insert(0, "key", [k for _, k in names])
            st.insert(0, "variable", [v for v, _ in names])
            for k, v in point.items():
                st[k] = v
            st.insert(0, "scenario", s)
            frames.append(st)
        return pd.concat(frames, ignore_index=True)
