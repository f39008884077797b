"""CPU: host compiler, C-ABI surface, scenario builders -- no compute calls (no GPU here)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, load_golden
from manusim_b200 import compile_scenario, scenarios
from manusim_b200.compiler import DIST_KINDS


def test_compile_tables_simple_scheduler():
    sim, res, prod, exp = scenarios.load_example("simple_scheduler")
    t = compile_scenario(sim, res, prod, policy="max_priority")
    assert (t.R, t.P, int(t.route_start[-1]), t.n_rings) == (13, 10, 93, 165)
    assert t.resource_names[0] == "r01" and t.product_names[-1] == "produto10"
    assert t.delivery_mode == 0 and t.log_queues is False       # `log_queue` typo key is ignored
    d = t.prod_freq[0]
    assert d.kind == DIST_KINDS["erlang"] and (d.d0, d.d1) == (3.003643419467815, 3.3292899999999994)
    assert exp["exp_seed"] == 123
    assert t.ring_names()[0] == ("deliveredOntime", "produto01") and t.ring_names()[-1] == ("totalInventory_general", "general")


def test_compile_errors_mirror_reference():
    cfg, res, prod = scenarios.toy2()
    with pytest.raises(ValueError, match="run_until must be specified"):
        compile_scenario({}, res, prod)
    bad = {"rA": {"setup": {"dist": "weibull", "params": [1]}}, "rB": res["rB"]}
    with pytest.raises(ValueError, match="Unknowh distribution type"):
        compile_scenario(cfg, bad, prod)
    with pytest.raises(NotImplementedError):
        compile_scenario(cfg, res, prod, policy="edd_custom")
    import copy

    stuck = copy.deepcopy(prod)
    stuck["p1"]["demand"]["freq"] = {"dist": "constant", "params": [0]}
    with pytest.raises(ValueError, match="never advances"):
        compile_scenario(cfg, res, stuck)


def test_dbr_constraint_matches_reference():
    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    t = compile_scenario(sim, res, prod, policy="dbr")
    assert t.resource_names[t.dbr["ccr"]] == "r09"
    assert t.dbr["ccr_setup"] == 0.45 and t.dbr["interval"] == 48 and t.dbr["cb_target"] == 220
    assert t.dbr["ccr_pt"][6] == 0 and t.dbr["ccr_pt"][8] == 0   # produto07/09 skip the CCR


def test_jobshop20_spec():
    cfg, res, prod = scenarios.jobshop20()
    t = compile_scenario(cfg, res, prod)
    assert (t.R, t.P, int(t.route_start[-1])) == (20, 10, 100)
    visits = np.bincount(t.step_resource, minlength=20)
    assert visits.tolist() == [4, 4, 5, 5, 7, 5, 4, 5, 4, 6, 6, 4, 6, 4, 5, 5, 4, 6, 5, 6]
    assert all(t.step_resource[i] != t.step_resource[i + 1] for p in range(10)
               for i in range(t.route_start[p], t.route_start[p + 1] - 1))


def test_library_exports_every_declared_symbol():
    from manusim_b200 import _lib

    header = (ROOT / "include" / "msim.h").read_text()
    declared = set(re.findall(r"\b(msim_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.msim_version() == 201
    # struct sizes of the ctypes mirror match the header's layout (all 8-byte aligned fields)
    assert ctypes.sizeof(_lib.Dist) == 40
    assert ctypes.sizeof(_lib.RunArgs) % 8 == 0 and ctypes.sizeof(_lib.Tables) % 8 == 0


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from manusim_b200 import BatchEngine

    cfg, res, prod = scenarios.toy2()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchEngine(compile_scenario(cfg, res, prod))


def test_product_does_not_import_oracle():
    for f in (ROOT / "manusim_b200").rglob("*.py"):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, f


def test_golden_scenarios_roundtrip_through_compiler():
    for name in ("toy2", "ss_lots", "js20_ondue"):
        s = load_golden(name)["scenario"]
        t = compile_scenario(s["cfg"], s["resources"], s["products"])
        assert t.n_rings == 11 * t.P + 4 * t.R + 3


def test_unsupported_dbr_options_are_refused():
    sim, res, prod, _ = scenarios.load_example("toc_dbr")
    with pytest.raises(NotImplementedError):
        compile_scenario(dict(sim, min_lotsize=2), res, prod, policy="dbr")
    compile_scenario(dict(sim, ccr_release_limit=0.6, order_release_limit=3), res, prod, policy="dbr")
    # sb_update / cb_update: accepted -- as shipped they change no record (see compiler.py), only the step count
    t = compile_scenario(dict(sim, sb_update=True, cb_update=True), res, prod, policy="dbr")
    assert t.dbr["sb_update"] is True


def test_ctypes_mirror_matches_the_c_header(tmp_path):
    """sizeof / offsetof of every struct in include/msim.h, as gcc lays them out, equal the ctypes mirror in _lib.py."""
    import subprocess

    from manusim_b200 import _lib

    structs = {"msim_dist_t": _lib.Dist, "msim_tables_t": _lib.Tables, "msim_run_args_t": _lib.RunArgs,
               "msim_host_args_t": _lib.HostArgs, "msim_info_t": _lib.Info}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/msim.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, *_ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, *_ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_library_refuses_without_device():
    """msim_create through the real libmsim.so on a box without a GPU gets as far as the device query and fails there
    -- there is no host execution path in the library."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from manusim_b200 import _lib
    from manusim_b200.engine import make_ctables

    lib = _lib.load()
    cfg, res, prod = scenarios.toy2()
    ct, keep = make_ctables(compile_scenario(cfg, res, prod))
    h = ctypes.c_void_p()
    assert lib.msim_create(ctypes.byref(ct), ctypes.byref(h)) == -2          # MSIM_E_CUDA
    assert b"no CPU fallback" in lib.msim_last_error()


def test_product_never_loads_the_emulated_library():
    """tests/emu is test infrastructure: nothing under manusim_b200/, bench.py or __graft_entry__.py refers to it."""
    for f in list((ROOT / "manusim_b200").rglob("*.py")) + [ROOT / "bench.py", ROOT / "__graft_entry__.py"]:
        txt = f.read_text()
        assert "tests/emu" not in txt and "libmsim_emu" not in txt and "import emu" not in txt, f
