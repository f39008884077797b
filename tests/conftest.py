import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    import json

    import numpy as np

    z = np.load(GOLDEN / f"{name}.npz")
    g = {k: z[k] for k in z.files}
    g["scenario"] = json.loads(str(g["scenario"]))
    g["metrics"] = json.loads(str(g["metrics"]))
    g["steps"] = int(g["steps"])
    g["error"] = str(g["error"])
    return g


ORACLE_POLICY = {"base": "fifo", "simple_scheduler": "max_priority", "dbr": "dbr"}

GOLDEN_NAMES = ["toy2", "toy2_instantly", "ss_cfg1", "ss_det", "ss_instantly", "ss_lots", "js20_asready",
                "js20_ondue", "js20_breakdowns", "dbr_shipped", "dbr_repaired", "dbr_typed", "ss_queues", "dbr_limits",
                "ss_ondue", "js20_lots_queues", "dbr_ondue", "dbr_tiny_setup", "ss_zero_ops", "dbr_mixed_types", "dbr_f64_order"]
ERROR_NAMES = ["err_negdelay", "err_amount"]
# pinned for the oracle only: options the kernel refuses (compile_scenario raises NotImplementedError)
ORACLE_ONLY_NAMES = ["dbr_min_lot", "dbr_min_lot_frac"]
