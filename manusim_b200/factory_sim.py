"""Drop-in ``FactorySimulation``: the reference's constructor, hook surface, ``logs`` and metrics views
(/root/reference/src/manusim/factory_sim.py) in front of the batched CUDA engine.

What is kept:
  * ``FactorySimulation(config, resources, products, print_mode="metrics", seed=None)`` and the config keys/defaults
    of factory_sim.py:36-57; ``reset_simulation(seed, log_save_path)``, ``run_simulation() -> seconds``,
    ``products_metrics/resources_metrics/general_metrics``, ``save_metrics``, ``save_custom_metrics``, ``logs``,
    ``seed``, ``env.now`` -- the seven members ``ExperimentRunner._run_single_simulation`` touches
    (experiment.py:55-69);
  * the scheduler hook as a CLOSED policy set: the base class (FIFO + release-per-demand), ``NewSimulation`` of
    examples/simple_scheduler (max priority, FIFO ties) and ``DBRSimulation`` of examples/toc_dbr (rope + buffer
    dispatch) map to compile-time device policies.  Overriding ``scheduler``/``order_selection`` or any of the
    ``_custom_*`` hooks with arbitrary Python raises ``NotImplementedError`` -- Python generators over SimPy stores
    cannot run inside a CUDA kernel, and silently ignoring them would be worse.

What is new: ``run_batch(n)`` runs n replications in one launch and returns device tensors.

Variates: a single ``run_simulation()`` draws from Philox4x32-10 keyed by ``seed`` (statistically equivalent to, not
bit-identical with, CPython's MT19937 stream); pass ``tapes=`` (reference-recorded variates) to ``reset_simulation``
for a bit-exact replay of a reference run.
"""
from __future__ import annotations

from pathlib import Path
from time import time
from typing import Any, Dict, Literal, Optional

import numpy as np
import pandas as pd

from .compiler import ScenarioTables, compile_scenario
from .logs import Logger
from .metrics import (MetricGeneral, MetricProducts, MetricResources, complete_wide_metrics_df,
                      pivot_aggfunc_for_metric)

_HOOKS = ("scheduler", "order_selection", "_initiate_custom_env", "_create_custom_logs", "_register_custom_logs",
          "_custom_order_in_resource_input", "_custom_order_out_resource_input", "_custom_part_processed",
          "_custom_order_processed", "_custom_fg_reduced", "_start_custom_process", "_release_order",
          "_production_system", "_transportation", "_generate_demand_orders", "_delivery_orders", "_breakdowns")


class _Env:
    """Stand-in for ``simpy.Environment``: only ``now`` is observable through the drop-in boundary."""

    def __init__(self):
        self.now = 0


class FactorySimulation:
    policy = "fifo"          # device policy this class stands for
    dbr_repair = False

    def __init__(self, config: dict, resources: dict, products: dict,
                 print_mode: Literal["status", "metrics", "all"] = "metrics", seed: int = None, **engine_opts):
        self.config: Dict[str, Any] = config
        self.resources_config: Dict[str, dict] = resources
        self.products_config: Dict[str, dict] = products
        self.seed = seed
        self.print_mode = print_mode
        self.log_save_path = self.config.get("log_save_path", None)
        self.memory_size = self.config.get("memory_size", 10000)
        self.run_until = self.config.get("run_until", None)
        if self.run_until is None:
            raise ValueError("run_until must be specified")
        self.warmup = self.config.get("warmup", 0)
        self.monitor_warmup = self.config.get("monitor_warmup", self.warmup)
        self.monitor_interval = self.config.get("monitor_interval", int(self.run_until / 4))
        self.delivery_mode = self.config.get("delivery_mode", "asReady")
        self.log_queues = self.config.get("log_queues", False)
        self.enable_event_logging = self.config.get("enable_event_logging", True)
        self.allocate_log_storage = self.config.get("allocate_log_storage", self.enable_event_logging)
        self._check_hooks()
        self._engine_opts = engine_opts
        self._engine = None
        self._tapes = None
        self.tables: ScenarioTables = compile_scenario(config, resources, products, policy=self.policy,
                                                       dbr_repair=self.dbr_repair, **engine_opts)
        self._initiate_environment()

    def monitor_steps(self) -> int:
        """`Environment.step()` calls the reference's console monitor adds to a run (factory_sim.py:656-684): with
        print_mode in {"all", "metrics", "status"} -- "metrics" is the constructor default -- `_run_monitor` starts one more
        process: its Initialize event, the Timeout(monitor_warmup) and one Timeout(monitor_interval) per report, each
        counted while its instant lies before run_until (the URGENT stop event precedes anything scheduled at run_until).
        The monitor only prints; it logs nothing and changes no other event's order, so records and metrics are the same
        with and without it -- the engine does not print progress reports, it only counts these steps in `sim_events`."""
        if self.print_mode not in ("all", "metrics", "status"):
            return 0
        steps = 1                                              # Initialize[_monitor]
        t, dt = self.monitor_warmup, self.monitor_interval
        if not (t < self.run_until):
            return steps
        if not (dt > 0):
            raise ValueError("monitor_interval must be positive (the reference's monitor would never advance)")
        k = int((self.run_until - t) / dt)                     # number of k >= 0 with t + k * dt < run_until, exactly
        while t + k * dt >= self.run_until:
            k -= 1
        while t + (k + 1) * dt < self.run_until:
            k += 1
        return 1 + (k + 1)

    def host_only_steps(self) -> int:
        """Environment.step() calls of reference processes that touch no model state and therefore do not run on the
        device: the console monitor here, DBRSimulation adds the sleeping adjust_shipping_buffer processes."""
        return self.monitor_steps()

    # ------------------------------------------------------------------ hook policing
    def _check_hooks(self):
        for name in _HOOKS:
            impl = getattr(type(self), name, None)
            base = getattr(FactorySimulation, name, None)
            if impl is None or impl is base:
                continue
            owner = next(c for c in type(self).__mro__ if name in c.__dict__)
            if owner.__dict__.get("_device_policy_class", False):
                continue
            raise NotImplementedError(
                f"{type(self).__name__}.{name} overrides a SimPy-process hook with Python code; the CUDA engine only "
                f"runs the recognised policies (fifo, max_priority, dbr, edd): set `policy` on the class instead")

    # the reference's no-op hooks, kept so subclasses/calls resolve (factory_sim.py:91,131,135,360-370,603,731,801,816)
    def scheduler(self):
        raise NotImplementedError("runs on the device")

    def order_selection(self, resource):
        raise NotImplementedError("runs on the device")

    def _initiate_custom_env(self):
        pass

    def _create_custom_logs(self):
        return

    def _register_custom_logs(self):
        pass

    def _custom_order_in_resource_input(self, productionOrder, resource):
        pass

    def _custom_order_out_resource_input(self, productionOrder, resource):
        pass

    def _custom_part_processed(self, product, resource, process):
        pass

    def _custom_order_processed(self, productionOrder, resource):
        pass

    def _custom_fg_reduced(self, product):
        pass

    def _start_custom_process(self):
        pass

    def print_custom_metrics(self):
        pass

    def save_custom_metrics(self, save_path: Path, saved_logs=False) -> None:
        pass

    # ------------------------------------------------------------------ environment / logs
    def _initiate_environment(self):
        self.env = _Env()
        self.warmup_finished = False
        self._init_loggint()

    def _init_loggint(self) -> None:
        self.logs: Logger = Logger(logs_save_path=self.log_save_path, mem_size=self.memory_size,
                                   enabled=self.enable_event_logging, allocate_storage=self.allocate_log_storage)
        for m in MetricProducts:
            self.logs.create_log(m.name, self.products_config.keys())
        for m in MetricResources:
            self.logs.create_log(m.name, self.resources_config.keys())
        for m in MetricGeneral:
            self.logs.create_log(m.name, ["general"])

    def reset_simulation(self, seed, log_save_path, tapes: Optional[dict] = None) -> None:
        """factory_sim.py:804-808; ``tapes`` = {"A","B","C","D"} arrays switches the next run to replay mode."""
        self.seed = seed
        self.log_save_path = log_save_path
        self._tapes = tapes
        self._initiate_environment()

    @property
    def engine(self):
        if self._engine is None:
            from .engine import BatchEngine

            self._engine = BatchEngine(self.tables)
        return self._engine

    # ------------------------------------------------------------------ execution
    def run_simulation(self) -> float:
        """One replication on the GPU; returns wall seconds like factory_sim.py:810-814 and fills ``logs``."""

        from .engine import decode_log

        eng = self.engine
        start = time()
        cap = int(self.config.get("device_log_capacity", 1 << 20))
        n_log = 1 if self.enable_event_logging else 0
        seed = 0 if self.seed is None else int(self.seed)
        # the run seed IS the replication's Philox key -- the mapping ExperimentRunner uses for run i, so that a run can
        # be reproduced from the seed recorded in its run_info.yaml (experiment.py:55-57: reset_simulation + run_simulation)
        key = np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
        res = eng.run(1, rep_seeds=None if self._tapes is not None else key,
                      tapes=[self._tapes] if self._tapes is not None else None,
                      n_log_reps=n_log, log_cap=cap if n_log else 0)
        if self._tapes is None:
            eng.retry_overflow(res, rep_seeds=key)
        eng.synchronize()
        status = int(res.status[0])
        if status in (1, 2):
            from .engine import STATUS_TEXT
            raise ValueError(STATUS_TEXT[status])         # what SimPy raises inside env.run()
        if status != 0:
            from .engine import STATUS_TEXT
            raise RuntimeError(f"replication failed: {STATUS_TEXT.get(status, status)}")
        self.env.now = self.run_until
        self.warmup_finished = True
        self.sim_events = int(res.n_events[0]) + self.host_only_steps()
        self._acc = (res.acc_sum[0].cpu().numpy(), res.acc_cnt[0].cpu().numpy())
        if n_log:
            count = int(res.log_count[0])
            if count > cap:
                raise RuntimeError(f"device log ring overflowed ({count} > {cap}); raise config['device_log_capacity']")
            ring, tm, val, _ = decode_log(res.log[0].cpu().numpy(), count)
            self.logs.load_stream(self.tables.ring_names(), ring, tm, val)
        return time() - start

    def run_batch(self, n_replications: int, seed: Optional[int] = None, **kw):
        """n replications in one launch (device tensors: per-ring sums/counts, SimPy-equivalent event counts, status)."""
        s = self.seed if seed is None else seed
        return self.engine.run(n_replications, seed=0 if s is None else int(s), **kw)

    # ------------------------------------------------------------------ metrics views (factory_sim.py:734-799)
    def _pivot_metric_logs(self, metric, saved_logs: bool = False) -> pd.DataFrame:
        df = self.logs.get_variable_logs(variable=metric.name, saved_logs=saved_logs)
        if df.empty:
            return pd.DataFrame()
        return df.pivot_table(values="value", index="key", columns="variable", aggfunc=pivot_aggfunc_for_metric(metric))

    def _build_family_metrics(self, metrics, keys, saved_logs=False, time_normalize=None) -> pd.DataFrame:
        frames = []
        for metric in metrics:
            df = self._pivot_metric_logs(metric, saved_logs=saved_logs)
            if time_normalize and metric.name in time_normalize and not df.empty:
                df = df / (self.env.now - self.warmup)
            frames.append(df)
        partial = pd.concat(frames, axis=1) if frames else pd.DataFrame()
        return complete_wide_metrics_df(partial, list(metrics), list(keys))

    def products_metrics(self, saved_logs=False) -> pd.DataFrame:
        return self._build_family_metrics(MetricProducts, self.products_config.keys(), saved_logs=saved_logs)

    def resources_metrics(self, saved_logs=False) -> pd.DataFrame:
        return self._build_family_metrics(MetricResources, self.resources_config.keys(), saved_logs=saved_logs,
                                          time_normalize={"utilization", "breakdown", "setup"})

    def general_metrics(self, saved_logs=False) -> pd.DataFrame:
        return self._build_family_metrics(MetricGeneral, ["general"], saved_logs=saved_logs)

    def save_metrics(self, save_path: Path, saved_logs=False) -> None:
        save_path = Path(save_path)
        save_path.mkdir(exist_ok=True, parents=True)
        self.products_metrics(saved_logs=saved_logs).to_csv(save_path / "metrics_products.csv")    # index kept (E14)
        self.resources_metrics(saved_logs=saved_logs).to_csv(save_path / "metrics_resources.csv")
        self.general_metrics(saved_logs=saved_logs).to_csv(save_path / "metrics_general.csv")


class NewSimulation(FactorySimulation):
    """examples/simple_scheduler/simulation.py:13-44: dispatch the queued order with the highest priority (first one
    among equals; the base scheduler sets every priority to 0)."""
    policy = "max_priority"
    _device_policy_class = True

    def __init__(self, config, resources, products, print_mode="all", seed=None, **kw):
        super().__init__(config, resources, products, print_mode, seed, **kw)


class EDDSimulation(FactorySimulation):
    """Extension: earliest-due-date dispatch (ProductionOrder.duedate, orders.py:11)."""
    policy = "edd"
    _device_policy_class = True


class DBRSimulation(FactorySimulation):
    """examples/toc_dbr/simulation.py:12-412 drum-buffer-rope: periodic rope release against the constraint buffer,
    buffer-penetration dispatch, FG seeded with the shipping buffers.  ``dbr_repair=True`` applies the documented
    repair P3 (the CCR feeds ``stores.resource_finished`` so the constraint buffer drains)."""
    policy = "dbr"
    _device_policy_class = True

    def __init__(self, config, resources, products, print_mode="all", seed=None, dbr_repair=False, **kw):
        self.dbr_repair = dbr_repair
        super().__init__(config, resources, products, print_mode, seed, **kw)
        d = self.tables.dbr
        self.order_release_limit = d["release_limit"]
        self.ccr_release_limit = d["ccr_limit"]
        self.scheduler_interval = d["interval"]
        self.cb_target_level = d["cb_target"]
        self.contraint_resource = self.tables.resource_names[d["ccr"]]
        self.shipping_buffer = {p: self.products_config[p].get("shipping_buffer", 0) for p in self.products_config.keys()}
        self.sb_update = bool(self.config.get("sb_update", False))
        self.cb_update = bool(self.config.get("cb_update", False))       # read and never used by the reference (:43)

    def host_only_steps(self) -> int:
        """+ the adjust_shipping_buffer processes of sb_update=True (examples/toc_dbr/simulation.py:357-406): as shipped they
        never see a penetration sample (process_fg_reduce is not wired to the base class's _custom_fg_reduced hook), so each
        of the P processes is Initialize, Timeout(warmup) and one Timeout(20 * scheduler_interval) after another -- steps
        only, counted while the instant lies before run_until."""
        steps = super().host_only_steps()
        if self.sb_update:
            per = 1                                                   # Initialize[adjust_shipping_buffer]
            t, dt = self.warmup, self.scheduler_interval * 20
            if t < self.run_until:
                k = int((self.run_until - t) / dt)
                while t + k * dt >= self.run_until:
                    k -= 1
                while t + (k + 1) * dt < self.run_until:
                    k += 1
                per += k + 1
            steps += per * len(self.products_config)
        return steps
