"""manusim_b200 -- B200-native batched discrete-event engine for manusim factory replications.

Import is CPU-safe (tables, scenarios, config compiler).  Anything that simulates needs libmsim.so and a CUDA
device and raises otherwise -- there is no CPU fallback.
"""
from .compiler import ScenarioTables, compile_scenario  # noqa: F401

__all__ = ["ScenarioTables", "compile_scenario", "BatchEngine"]


def __getattr__(name):
    if name == "BatchEngine":
        from .engine import BatchEngine
        return BatchEngine
    raise AttributeError(name)
