"""Run the UNMODIFIED reference model (/root/reference/src/manusim) on oracle/minisimpy.py.

TEST INFRASTRUCTURE ONLY (golden-vector generation + validation of oracle/factory_oracle.py).
Works only where /root/reference is mounted (the build container); nothing executed on the
GPU box imports this module.

What it does:
  * registers ``oracle.minisimpy`` as ``simpy`` and the three import stubs in ``oracle/_stubs``;
  * imports ``manusim.factory_sim.FactorySimulation`` (factory_sim.py:22) and the example
    subclasses (examples/simple_scheduler/simulation.py:13, examples/toc_dbr/simulation.py:12);
  * wraps the three ``DistributionGenerator`` instances (factory_sim.py:103-107) so every returned
    variate is recorded on a per-stream tape:
        A = rnd_product.random_number     (factory_sim.py:155-162)   float32 (int 0 -> 0.0)
        B = rnd_process.random_number     (factory_sim.py:410)       float32
        C = rnd_process.sum_random_numbers_batch (factory_sim.py:436-438) float64
        D = rnd_breakdown.random_number   (factory_sim.py:258-260,275,288) float32
  * wraps ``Logger.log`` (engine/logs.py:44) to capture the ordered record stream
    ``(ring id, float32 time, float32 value)``;
  * reports the SimPy step count (= next(env._eid) - len(env._queue)).

Documented oracle patches; each is a subclass override, the reference is never edited:
  P1  onDue: ``_delivery_orders`` calls ``_delivery_on_duedate`` without ``env.process`` (factory_sim.py:477-478).
  P2  toc_dbr: ``_log_vars`` no-op (examples/toc_dbr/simulation.py:86-119).
  P3  toc_dbr repair: ``_custom_order_processed`` feeds ``stores.resource_finished`` for the CCR.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
from pathlib import Path

import numpy as np

REF_ROOT = Path("/root/reference")
_HERE = Path(__file__).resolve().parent

PRODUCT_METRICS = ["deliveredOntime", "deliveredLate", "lostSales", "flowTime", "leadTime", "tardiness",
                   "earliness", "wip", "finishedGoods", "totalInventory", "released"]
RESOURCE_METRICS = ["utilization", "breakdown", "setup", "queue"]
GENERAL_METRICS = ["wip_general", "finishedGoods_general", "totalInventory_general"]


def reference_available() -> bool:
    return (REF_ROOT / "src" / "manusim" / "factory_sim.py").is_file()


_loaded = {}


def load_reference():
    """Import the reference package on top of minisimpy; returns a dict of modules/classes."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError("/root/reference is not mounted; golden vectors can only be generated in the build container")
    from oracle.simpy_select import pick

    engine, _loaded["simpy_engine"] = pick()      # minisimpy unless ORACLE_SIMPY=real|auto finds the real package
    sys.modules["simpy"] = engine
    for p in (str(_HERE / "_stubs"), str(REF_ROOT / "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    fs = importlib.import_module("manusim.factory_sim")
    utils = importlib.import_module("manusim.engine.utils")
    _loaded["factory_sim"] = fs
    _loaded["utils"] = utils
    _loaded["FactorySimulation"] = fs.FactorySimulation
    for name, rel in (("simple_scheduler", "examples/simple_scheduler/simulation.py"),
                      ("toc_dbr", "examples/toc_dbr/simulation.py")):
        spec = importlib.util.spec_from_file_location(f"_ref_example_{name}", REF_ROOT / rel)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _loaded[name] = mod
    return _loaded


def ring_index(variable: str, key, product_names, resource_names) -> int:
    P, R = len(product_names), len(resource_names)
    if variable in PRODUCT_METRICS:
        return PRODUCT_METRICS.index(variable) * P + product_names.index(key)
    if variable in RESOURCE_METRICS:
        return len(PRODUCT_METRICS) * P + RESOURCE_METRICS.index(variable) * R + resource_names.index(key)
    return len(PRODUCT_METRICS) * P + len(RESOURCE_METRICS) * R + GENERAL_METRICS.index(variable)


def _make_harness(base_cls, patches):
    ref = load_reference()
    DG = ref["utils"].DistributionGenerator

    class RecordingGenerator(DG):
        def __init__(self, seed, scalar_tape, batch_tape):
            super().__init__(seed)
            self._scalar_tape = scalar_tape
            self._batch_tape = batch_tape

        def random_number(self, distribution, params):
            v = super().random_number(distribution, params)
            self._scalar_tape.append(float(v))
            return v

        def sum_random_numbers_batch(self, distribution, params, count):
            v = super().sum_random_numbers_batch(distribution, params, count)
            self._batch_tape.append(float(v))
            return v

    class Harness(base_cls):
        _patches = tuple(patches)

        def _init_random_generators(self):
            self.tapes = {"A": [], "B": [], "C": [], "D": [], "A_batch": [], "D_batch": []}
            self.rnd_product = RecordingGenerator(self.seed, self.tapes["A"], self.tapes["A_batch"])
            self.rnd_process = RecordingGenerator(self.seed, self.tapes["B"], self.tapes["C"])
            self.rnd_breakdown = RecordingGenerator(self.seed, self.tapes["D"], self.tapes["D_batch"])

        def _init_loggint(self):
            super()._init_loggint()
            self.trace = []
            inner = self.logs.log
            pn = list(self.products_config.keys())
            rn = list(self.resources_config.keys())
            standard = set(PRODUCT_METRICS + RESOURCE_METRICS + GENERAL_METRICS)

            def log(variable, key, value):
                if variable in standard and self.logs.enabled:
                    self.trace.append((ring_index(variable, key, pn, rn),
                                       np.float32(value[0]), np.float32(value[1])))
                inner(variable=variable, key=key, value=value)

            self.logs.log = log

    if "P1" in patches:
        def _delivery_orders(self, product):
            while True:
                d = yield self.stores.outbound_demand_orders[product].get()
                if d.delivery_mode == "asReady":
                    self.env.process(self._delivery_as_ready(d))
                elif d.delivery_mode == "onDue":
                    self._delivery_on_duedate(d)          # P1: no env.process around a non-generator
                elif d.delivery_mode == "instantly":
                    self.env.process(self._delivery_instantly(d))
        Harness._delivery_orders = _delivery_orders
    if "P2" in patches:
        Harness._log_vars = lambda self, *a, **k: None
    if "P3" in patches:
        def _custom_order_processed(self, po, resource):
            if resource == self.contraint_resource:
                self.stores.resource_finished[resource].put(po)
        Harness._custom_order_processed = _custom_order_processed
    return Harness


def run_reference(sim_cfg: dict, resources: dict, products: dict, seed: int, policy: str = "base",
                  patches=()) -> dict:
    """Execute one replication the way ExperimentRunner does (experiment.py:55-57):
    construct, ``reset_simulation(seed, None)``, ``run_simulation()``."""
    ref = load_reference()
    base = {"base": ref["FactorySimulation"],
            "simple_scheduler": ref["simple_scheduler"].NewSimulation,
            "dbr": ref["toc_dbr"].DBRSimulation}[policy]
    cls = _make_harness(base, patches)
    sim = cls(dict(sim_cfg), resources, products, print_mode="none", seed=seed)
    if policy == "dbr":
        # E19: fresh-object semantics; E12: ExperimentRunner always resets first, so only the t=0
        # `_update_finished_goods` puts apply.  Build a fresh object and reset it exactly once.
        sim = cls(dict(sim_cfg), resources, products, print_mode="none", seed=seed)
    sim.reset_simulation(seed, None)
    error = None
    try:
        sim.run_simulation()
    except ValueError as exc:  # SimPy "Negative delay" / Container amount<=0
        error = repr(exc)
    env = sim.env
    steps = next(env._eid) - len(env._queue)
    tr = sim.trace
    out = {
        "ring": np.array([t[0] for t in tr], dtype=np.uint16),
        "time": np.array([t[1] for t in tr], dtype=np.float32),
        "value": np.array([t[2] for t in tr], dtype=np.float32),
        "tape_A": np.array(sim.tapes["A"], dtype=np.float32),
        "tape_B": np.array(sim.tapes["B"], dtype=np.float32),
        "tape_C": np.array(sim.tapes["C"], dtype=np.float64),
        "tape_D": np.array(sim.tapes["D"], dtype=np.float32),
        "steps": int(steps),
        "steps_counted": int(getattr(env, "steps", steps)),
        "now": env.now,
        "error": error,
        "sim": sim,
    }
    assert not sim.tapes["A_batch"] and not sim.tapes["D_batch"]
    return out


def load_example(name: str):
    """Return (simulation cfg, resources, products, experiment cfg) of a shipped example as plain dicts."""
    import yaml

    conf_dir = {"simple_scheduler": "conf", "toc_dbr": "config"}[name]
    root = REF_ROOT / "examples" / name / conf_dir
    cfg = yaml.safe_load((root / "config.yaml").read_text())
    products = yaml.safe_load((root / "products" / "products.yaml").read_text())
    resources = yaml.safe_load((root / "resources" / "resources.yaml").read_text())
    return cfg["simulation"], resources, products, cfg["experiment"]
