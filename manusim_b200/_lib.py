"""ctypes binding of libmsim.so (include/msim.h).  There is no CPU fallback: if the shared library is missing or
no CUDA device is present, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

PKG = Path(__file__).resolve().parent
# MSIM_LIB: load another build of the same library (A/B variants made by manusim_b200.build.build_lib(tag=...))
LIB_PATH = Path(os.environ["MSIM_LIB"]) if os.environ.get("MSIM_LIB") else PKG / "libmsim.so"

EXPORTS = ["msim_version", "msim_last_error", "msim_create", "msim_destroy", "msim_query", "msim_run",
           "msim_run_host", "msim_reduce_metrics", "msim_reduce_variables", "msim_metric_values", "msim_test_philox",
           "msim_test_variates"]


class Dist(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("raw0", C.c_double), ("raw1", C.c_double),
                ("d0", C.c_double), ("d1", C.c_double)]


class Tables(C.Structure):
    _fields_ = [
        ("n_resources", C.c_int32), ("n_products", C.c_int32), ("n_steps", C.c_int32),
        ("route_start", C.POINTER(C.c_int32)), ("step_resource", C.POINTER(C.c_int32)),
        ("step_time", C.POINTER(Dist)), ("res_setup", C.POINTER(Dist)), ("res_tbf", C.POINTER(Dist)),
        ("res_ttr", C.POINTER(Dist)), ("prod_freq", C.POINTER(Dist)), ("prod_qty", C.POINTER(Dist)),
        ("prod_due", C.POINTER(Dist)), ("prod_shipping_buffer", C.POINTER(C.c_double)),
        ("run_until", C.c_double), ("warmup", C.c_double),
        ("delivery_mode", C.c_int32), ("policy", C.c_int32), ("log_queues", C.c_int32),
        ("dbr_ccr", C.c_int32), ("dbr_repair", C.c_int32),
        ("dbr_interval", C.c_double), ("dbr_cb_target", C.c_double), ("dbr_release_limit", C.c_double),
        ("dbr_ccr_limit", C.c_double), ("dbr_min_lot", C.c_double), ("dbr_ccr_setup", C.c_double),
        ("dbr_ccr_pt", C.POINTER(C.c_double)),
        ("max_production_orders", C.c_int32), ("max_demand_orders", C.c_int32), ("fifo_capacity", C.c_int32),
    ]


class RunArgs(C.Structure):
    _fields_ = [
        ("n_reps", C.c_int64), ("rep_offset", C.c_int64), ("rng_mode", C.c_int32), ("_pad", C.c_int32),
        ("philox_seed", C.c_uint64), ("rep_seeds", C.c_void_p),
        ("tape_A", C.c_void_p), ("tape_A_off", C.c_void_p), ("tape_B", C.c_void_p), ("tape_B_off", C.c_void_p),
        ("tape_C", C.c_void_p), ("tape_C_off", C.c_void_p), ("tape_D", C.c_void_p), ("tape_D_off", C.c_void_p),
        ("rec_A", C.c_void_p), ("rec_B", C.c_void_p), ("rec_C", C.c_void_p), ("rec_D", C.c_void_p),
        ("rec_cap", C.c_int64), ("rec_count", C.c_void_p),
        ("acc_sum", C.c_void_p), ("acc_cnt", C.c_void_p), ("n_events", C.c_void_p), ("n_timeouts", C.c_void_p),
        ("status", C.c_void_p), ("log", C.c_void_p), ("log_count", C.c_void_p), ("log_cap", C.c_int64),
        ("n_log_reps", C.c_int64), ("stream", C.c_void_p),
    ]


class HostArgs(C.Structure):
    _fields_ = [
        ("n_reps", C.c_int64), ("rep_offset", C.c_int64), ("rng_mode", C.c_int32), ("_pad", C.c_int32),
        ("philox_seed", C.c_uint64), ("rep_seeds", C.c_void_p),
        ("tape_A", C.c_void_p), ("tape_A_off", C.c_void_p), ("tape_B", C.c_void_p), ("tape_B_off", C.c_void_p),
        ("tape_C", C.c_void_p), ("tape_C_off", C.c_void_p), ("tape_D", C.c_void_p), ("tape_D_off", C.c_void_p),
        ("acc_sum", C.c_void_p), ("acc_cnt", C.c_void_p), ("n_events", C.c_void_p), ("n_timeouts", C.c_void_p),
        ("status", C.c_void_p), ("log", C.c_void_p), ("log_count", C.c_void_p), ("log_cap", C.c_int64),
        ("n_log_reps", C.c_int64), ("kernel_ms", C.POINTER(C.c_double)),
        ("log_dev", C.c_void_p), ("log_count_dev", C.c_void_p), ("log_cap_dev", C.c_int64),
        ("n_log_reps_dev", C.c_int64), ("stats_dev", C.c_void_p),
    ]


class Info(C.Structure):
    _fields_ = [("n_rings", C.c_int32), ("smem_per_replication", C.c_int32), ("smem_tables", C.c_int32),
                ("warps_per_block", C.c_int32), ("blocks_per_sm", C.c_int32), ("n_sms", C.c_int32),
                ("regs_per_thread", C.c_int32), ("max_production_orders", C.c_int32),
                ("max_demand_orders", C.c_int32), ("fifo_capacity", C.c_int32), ("const_layout", C.c_int32)]


_lib = None


def load():
    """dlopen libmsim.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(manusim_b200 has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.msim_version.restype = C.c_int
    lib.msim_last_error.restype = C.c_char_p
    lib.msim_create.argtypes = [C.POINTER(Tables), C.POINTER(C.c_void_p)]
    lib.msim_destroy.argtypes = [C.c_void_p]
    lib.msim_query.argtypes = [C.c_void_p, C.POINTER(Info)]
    lib.msim_run.argtypes = [C.c_void_p, C.POINTER(RunArgs)]
    lib.msim_run_host.argtypes = [C.c_void_p, C.POINTER(HostArgs)]
    lib.msim_reduce_metrics.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                        C.c_void_p]
    lib.msim_reduce_variables.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                          C.c_void_p]
    lib.msim_metric_values.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p]
    lib.msim_test_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    lib.msim_test_variates.argtypes = [C.POINTER(Dist), C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64]
    for name in EXPORTS:
        getattr(lib, name)
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().msim_last_error().decode()
        if rc == -3:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError(f"libmsim error {rc}: {msg}")
