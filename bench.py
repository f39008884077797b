#!/usr/bin/env python
"""bench.py -- simulated events/s and replications/s of the batched factory-replication path on B200.

    python bench.py --gpus N --steps K --warmup W              # our arm (CUDA engine)
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference arm: SimPy-process model on host cores

A "step" is one pass of the hot path over one batch of synthetic replications:
  * default workload ``jobshop20`` = BASELINE.json config #4: 20 resources x 10 products, gamma
    processing times, normal setups, expo TBF / normal TTR breakdowns, run_until=10000, warmup=1000; replications
    with even global index deliver asReady, odd ones onDue; ``--reps`` replications per GPU per step (a bounded
    slice of the 1M-replication job; weak scaling: every rank simulates its own contiguous index range);
  * ``--workload simple_det`` = config #2 (examples/simple_scheduler, every distribution constant).
Per step each rank launches the simulation kernel(s), the per-ring {sum x, sum x^2, n} reduction, and (N>1) one NCCL
allreduce of that [K,3] float64 vector -- the only collective on the path.

Unit of `value`: SimPy-equivalent events per second = Environment.step() calls the reference would execute for
the same trajectories (counted by the kernel, bit-exact against the oracle in replay mode), whole job over all ranks.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BASE_SEED = 20260101


# --------------------------------------------------------------------------------------------- workloads
def workload_spec(name, run_until=None):
    from manusim_b200 import scenarios

    if name == "jobshop20":
        ru = run_until or 10000
        wu = ru // 10
        a = scenarios.jobshop20(run_until=ru, warmup=wu, mode="asReady")
        b = scenarios.jobshop20(run_until=ru, warmup=wu, mode="onDue")
        return {"name": name, "parts": [(a, "fifo", 0, 2), (b, "fifo", 1, 2)], "run_until": ru, "warmup": wu,
                "pools": [(112, 112), (112, 128)],     # (production-order, demand-order) pool sizes; onDue keeps orders until due
                                                       # (smaller pools = more resident replications; overflow is re-run)
                "desc": f"jobshop20 (BASELINE config #4: R=20,P=10, setups, TBF/TTR breakdowns, even=asReady/odd=onDue, "
                        f"run_until={ru}, warmup={wu})"}
    if name == "simple_det":
        sim, res, prod, _ = scenarios.load_example("simple_scheduler")
        res, prod = scenarios.make_deterministic(res, prod)
        ru = run_until or 1000
        cfg = dict(sim, run_until=ru, warmup=50)
        return {"name": name, "parts": [((cfg, res, prod), "max_priority", 0, 1)], "run_until": ru, "warmup": 50,
                "desc": f"simple_scheduler deterministic (BASELINE config #2, run_until={ru})"}
    if name in ("toc_dbr", "toc_dbr_shipped"):
        sim, res, prod, _ = scenarios.load_example("toc_dbr")
        ru = run_until or 200001
        wu = ru // 2
        cfg = dict(sim, run_until=ru, warmup=wu)
        cfg.pop("print_mode", None)
        repaired = name == "toc_dbr"
        return {"name": name, "parts": [((cfg, res, prod), "dbr", 0, 1)], "run_until": ru, "warmup": wu,
                "dbr_repair": repaired, "pools": [(80, 24)],   # rope batches orders; `instantly` frees demand orders at once
                "desc": f"examples/toc_dbr drum-buffer-rope (BASELINE config #3, "
                        f"{'repaired: oracle patches P2+P3' if repaired else 'as shipped: oracle patch P2'}, "
                        f"run_until={ru}, warmup={wu})"}
    raise SystemExit(f"unknown workload {name}")


def whole_wave_batch(residents, waves):
    """Replications per workload part: `waves` waves of the part with the most resident replications, rounded up to whole
    waves of EVERY part (kernel time is quantised in waves: a half-filled last wave costs 0.7 of a whole one,
    profiles/r02_wave_quantisation_x.jsonl) unless that would grow the batch -- and with it the duration of a step -- by
    more than a quarter."""
    per_part = int(round(waves * max(residents)))
    whole = math.lcm(*[int(r) for r in residents])
    rounded = -(-per_part // whole) * whole
    return rounded if 4 * rounded <= 5 * per_part else per_part


def rep_keys(seed, idx):
    from manusim_b200.engine import rep_keys as _rk

    return _rk(seed, idx)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [x.strip() for x in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU side
def _oracle_one(task):
    spec_name, run_until, part_idx, seed = task
    os.environ.setdefault("ORACLE_SIMPY", "auto")          # the real SimPy package when importable, else the restatement
    from oracle.factory_oracle import FactoryOracle

    spec = workload_spec(spec_name, run_until)
    (cfg, res, prod), policy, _, _ = spec["parts"][part_idx]
    cfg = dict(cfg, enable_event_logging=True)
    t0 = time.perf_counter()
    out = FactoryOracle(cfg, res, prod, seed, policy, dbr_repair=spec.get("dbr_repair", False)).run()
    return out["steps"], time.perf_counter() - t0, len(out["ring"])


def cpu_pool_pass(pool, spec_name, run_until, n_parts, seeds):
    """One pass of the reference-style parallel path (experiment.py:186-210): one replication per task."""
    tasks = [(spec_name, run_until, i % n_parts, int(s)) for i, s in enumerate(seeds)]
    t0 = time.perf_counter()
    outs = list(pool.map(_oracle_one, tasks))
    wall = time.perf_counter() - t0
    events = sum(o[0] for o in outs)
    return events, wall, sum(o[1] for o in outs)


def cpu_baseline_meta():
    """What the CPU arm is: the oracle PORT of the reference's model (the unmodified reference needs /root/reference and
    cannot travel to the GPU box) on the best SimPy layer importable here, plus the calibration of that port against the
    UNMODIFIED reference measured in the build container (tools/cpu_calibration.py -> profiles/cpu_calibration.json)."""
    os.environ.setdefault("ORACLE_SIMPY", "auto")
    from oracle.factory_oracle import SIMPY_ENGINE

    meta = {"kind": "port", "engine": f"oracle/factory_oracle.py on {SIMPY_ENGINE}"}
    try:
        c = json.loads((ROOT / "profiles" / "cpu_calibration.json").read_text())
        meta["calibration"] = {"port_over_unmodified_reference": c["port_over_reference"],
                               "unmodified_reference_events_per_s_per_core": c["unmodified_reference_events_per_s_per_core"],
                               "port_events_per_s_per_core": c["port_events_per_s_per_core"],
                               "source": "profiles/cpu_calibration.json (build container, one core, same seeds)"}
    except Exception:
        meta["calibration"] = None
    return meta


def reference_arm(args):
    """Times the reference's own CPU implementation of the path.  The reference is pure Python on SimPy and cannot
    travel to the GPU box, so this runs its restatement (oracle/factory_oracle.py on oracle/minisimpy.py) the way
    ExperimentRunner._run_parallel does: a ProcessPoolExecutor over replications on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    from oracle.variates import run_seeds

    cores = os.cpu_count() or 1
    workers = min(cores, args.cpu_workers) if args.cpu_workers else cores
    spec = workload_spec(args.workload, args.run_until)
    n_parts = len(spec["parts"])
    per_step = workers
    seeds = run_seeds(BASE_SEED, per_step * (args.steps + args.warmup))
    with ProcessPoolExecutor(max_workers=workers) as pool:
        for w in range(args.warmup):
            cpu_pool_pass(pool, args.workload, args.run_until, n_parts, seeds[w * per_step:(w + 1) * per_step])
        events = 0
        wall = 0.0
        for k in range(args.steps):
            i = (args.warmup + k) * per_step
            e, t, _ = cpu_pool_pass(pool, args.workload, args.run_until, n_parts, seeds[i:i + per_step])
            events += e
            wall += t
    value = events / wall
    sample = f"{per_step} replications per step (one per worker process), full-length {spec['desc']}"
    line = {"impl": "reference", "metric": "simulated_events_per_sec", "value": value, "unit": "events/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["desc"], "replications_per_step": per_step, "l2": "n/a (CPU)"},
            "replications_per_sec": per_step * args.steps / wall,
            "cpu_baseline": dict(cpu_baseline_meta(), value=value, unit="events/s", cores=workers, sample=sample),
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
COUNTERS = ROOT / "profiles" / "kernel_counters.json"


def load_counters(workload):
    """ncu-derived per-kernel counters committed under profiles/ (written by tools/summarize_profiles.py from the
    `ncu --set full` capture named inside the file).  Absent file or workload => None: the bench never invents them."""
    try:
        d = json.loads(COUNTERS.read_text())
        return d["workloads"][workload], d.get("source")
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="jobshop20")
    ap.add_argument("--reps", type=int, default=0,
                    help="replications per GPU per step (0 = `--waves` full waves of the resident replications)")
    ap.add_argument("--waves", type=float, default=4.0,
                    help="with --reps 0: replications per part = waves x (resident warps per SM x SMs) of its kernel, "
                         "the same count for every part (the largest)")
    ap.add_argument("--run-until", type=int, default=None)
    ap.add_argument("--log-cap", type=int, default=4096, help="records per replication HBM ring (0 = no log streaming)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--max-po", type=int, default=0, help="production-order pool per replication (0 = library default)")
    ap.add_argument("--max-do", type=int, default=0, help="demand-order pool per replication (0 = library default)")
    ap.add_argument("--fifo", type=int, default=0, help="zero-delay event ring capacity (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--serial-parts", action="store_true", help="(default; kept for old command lines) launch the workload "
                    "parts one after the other on one stream")
    ap.add_argument("--concurrent-parts", action="store_true", help="one stream per workload part: two DIFFERENT kernels then "
                    "share the SMs and their instruction caches, which costs 6-8 %% (profiles/r02_waves_sweep_t.jsonl)")
    ap.add_argument("--no-retry", action="store_true", help="do not re-run pool-overflow replications with big pools")
    args = ap.parse_args()
    args.serial_parts = not args.concurrent_parts
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist

    from manusim_b200 import BatchEngine, compile_scenario

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = workload_spec(args.workload, args.run_until)
    parts = []
    for i, ((cfg, res, prod), policy, phase, period) in enumerate(spec["parts"]):
        mpo, mdo = spec.get("pools", [(0, 0)] * len(spec["parts"]))[i]
        eng = BatchEngine(compile_scenario(cfg, res, prod, policy=policy, dbr_repair=spec.get("dbr_repair", False),
                                           max_production_orders=args.max_po or mpo,
                                           max_demand_orders=args.max_do or mdo, fifo_capacity=args.fifo))
        resident = eng.resident_replications
        parts.append({"eng": eng, "phase": phase, "period": period, "resident": resident})
    if args.reps > 0:
        reps = args.reps
    else:
        # every part simulates the same number of replications (the workload alternates them by index parity): at least
        # `waves` waves of the part with the most resident replications, rounded up to WHOLE waves of every part when the
        # parts' kernels hold different numbers of replications (a last wave that fills 40 % of the SMs costs a whole one)
        reps = whole_wave_batch([p["resident"] for p in parts], args.waves) * max(p["period"] for p in parts)
    for p in parts:
        p["n"] = reps // p["period"]
    K = parts[0]["eng"].K
    n_total_rank = sum(p["n"] for p in parts)
    log_cap = args.log_cap
    for p in parts:
        p["res"] = p["eng"].alloc(p["n"], p["n"] if log_cap else 0, log_cap)
        p["host"] = p["eng"].alloc_host(p["n"])
        p["seeds_host"] = torch.empty((p["n"],), dtype=torch.int64).pin_memory()
        p["seeds_dev"] = torch.empty((p["n"],), dtype=torch.int64, device=dev)
        p["ev_part"] = torch.zeros((), dtype=torch.int64, device=dev)      # events of this part over the timed steps
    stats = torch.zeros((K, 3), dtype=torch.float64, device=dev)
    stats_host = torch.zeros((K, 3), dtype=torch.float64).pin_memory()
    ev_total = torch.zeros((), dtype=torch.int64, device=dev)
    fail_total = torch.zeros((), dtype=torch.int64, device=dev)
    fail_codes = torch.zeros((9,), dtype=torch.int64, device=dev)
    rec_total = torch.zeros((), dtype=torch.int64, device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    part_streams = [stream if args.serial_parts else torch.cuda.Stream(dev) for _ in parts]
    side = [torch.cuda.Stream(dev) for _ in parts]

    def set_seeds(step):
        for p in parts:
            gidx = (np.arange(p["n"], dtype=np.int64) * p["period"] + p["phase"] + rank * reps)
            keys = rep_keys(BASE_SEED + step, gidx)
            p["seeds_host"].copy_(torch.from_numpy(keys.view(np.int64)))

    sim_events = []
    retried = [0, 0]
    launches = [0]

    def device_step(step, timed):
        set_seeds(step)
        stats.zero_()
        fork = torch.cuda.Event()
        fork.record(stream)
        for p, ps in zip(parts, part_streams):
            ps.wait_event(fork)
            with torch.cuda.stream(ps):
                p["seeds_dev"].copy_(p["seeds_host"], non_blocking=True)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record(ps)
                p["eng"].run(p["n"], rep_seeds=p["seeds_dev"], out=p["res"], stream=ps)
                e1.record(ps)
            launches[0] += timed
            if timed:
                sim_events.append((e0, e1, p, step))
        for ps in part_streams:
            if ps is not stream:
                j = torch.cuda.Event()
                j.record(ps)
                stream.wait_event(j)
        if not args.no_retry:
            # congested replications that overflowed the shared-memory pools are re-run on a big-pool twin engine;
            # the (few, long) re-runs of all parts overlap on side streams
            ctxs = [p["eng"].retry_launch(p["res"], rep_seeds=p["seeds_dev"], stream=side[i]) for i, p in enumerate(parts)]
            for p, c in zip(parts, ctxs):
                retried[0] += p["eng"].retry_collect(c)
                if timed and c is not None:
                    retried[1] += 1
        for p in parts:
            p["eng"].reduce(p["res"], out=stats, stream=stream)
            launches[0] += timed
        if world > 1:
            dist.all_reduce(stats)
        if timed:
            for p in parts:
                ev_total.add_(p["res"].n_events.sum())
                p["ev_part"].add_(p["res"].n_events.sum())
                rec_total.add_(p["res"].acc_cnt.sum(dtype=torch.int64))
                fail_total.add_((p["res"].status != 0).sum())
                fail_codes.add_(torch.bincount(p["res"].status, minlength=9)[:9])

    for w in range(args.warmup):
        device_step(w, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    step_ms = []
    torch.cuda.synchronize()
    for k in range(args.steps):
        flush_buf.fill_(k & 0xFF)          # L2 flush between timed iterations (outside the timed events)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        device_step(args.warmup + k, True)
        s1.record(stream)
        torch.cuda.synchronize()
        step_ms.append(s0.elapsed_time(s1))
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    ev = ev_total.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev)
    total_ms = float(t.item())
    events = int(ev.item())
    value = events / (total_ms * 1e-3)
    reps_total = n_total_rank * world * args.steps

    # dominant kernel: the simulation kernel(s); algorithmic bytes per SURVEY 8(d).  With one stream per part the
    # kernels of a step overlap, so the kernel time of a step is the span from the first start to the last end.
    n_launch = len(sim_events)
    kern_ms = 0.0
    for st in sorted({s for _, _, _, s in sim_events}):
        evs = [(e0, e1) for e0, e1, _, s in sim_events if s == st]
        first = evs[0][0]
        kern_ms += max(first.elapsed_time(e1) for _, e1 in evs) - min(first.elapsed_time(e0) for e0, _ in evs)
    records_rank = int(rec_total.item())
    alg_bytes = 8 * records_rank + (16 * K + 8) * n_total_rank * args.steps
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    counters, counters_src = load_counters(args.workload)
    dram_ratio = (counters or {}).get("dram_bytes_per_algorithmic_byte") if log_cap else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (dram_ratio * alg_bytes / max(n_launch, 1)) if dram_ratio else None,
                "traffic_source": (f"{COUNTERS.relative_to(ROOT)}: dram__bytes_read.sum + dram__bytes_write.sum per "
                                   f"algorithmic byte of the capture named there ({counters_src}), times this run's "
                                   f"algorithmic bytes per launch") if dram_ratio else None,
                "peak_source": peak_src,
                "kernel": "msim::sim_kernel<policy,rng,opts,minb>",
                "kernel_ms_per_launch": kern_ms / max(n_launch, 1), "launches": n_launch,
                "kernel_ms_note": "span of the overlapping per-part launches of a step / launches" if not args.serial_parts
                                  else "serial launches",
                "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "note": "path is instruction-fetch / issue bound, not HBM bound (SURVEY 8(d)); see issue_slots",
                "kernel_share_of_step": kern_ms / sum(step_ms)}

    # per workload part (e.g. the asReady and the onDue half of jobshop20): simulation-kernel time and events/s of this
    # rank over the timed steps -- reporting only, never allowed to break the line
    try:
        by_part = []
        for i, p in enumerate(parts):
            ms = sum(e0.elapsed_time(e1) for e0, e1, q, _ in sim_events if q is p)
            evp = int(p["ev_part"].item())
            recp = int(p["res"].acc_cnt.sum(dtype=torch.int64).item())      # records of the last launch of this part
            by_part.append({"part": i, "delivery_mode": spec["parts"][i][0][0].get("delivery_mode"),
                            "replications_per_launch": p["n"], "kernel_ms_total": ms,
                            "events_per_launch": evp / args.steps,
                            "algorithmic_bytes_per_launch": 8 * recp + (16 * K + 8) * p["n"],
                            "events_per_s_while_resident": evp / (ms * 1e-3) if ms > 0 else None,
                            "warps_per_sm": p["eng"].info["warps_per_block"] * p["eng"].info["blocks_per_sm"],
                            "waves": p["n"] / p["resident"],
                            "smem_per_replication": p["eng"].info["smem_per_replication"],
                            "regs_per_thread": p["eng"].info["regs_per_thread"]})
        roofline["kernel_by_part"] = by_part
        if not args.serial_parts and len(parts) > 1:
            roofline["kernel_by_part_note"] = ("the parts run concurrently (one stream each): a part's kernel_ms covers the "
                                               "time it shares the SMs with the others")
    except Exception as exc:                                   # pragma: no cover
        roofline["kernel_by_part"] = f"unavailable: {exc!r}"

    # the ceiling that actually bounds this path: warp-instruction issue slots = SMs x 4 schedulers x clock
    clocks = sampler.stop() if rank == 0 else None
    try:
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        n_sms = parts[0]["eng"].info["n_sms"]
        ceiling = n_sms * 4 * mhz * 1e6
        ipe = (counters or {}).get("warp_instr_per_event")
        roofline["issue_slots"] = {
            "ceiling_warp_instr_per_s": ceiling, "warp_instr_per_event": ipe,
            "busy_frac_estimate": ((events / world) / (kern_ms * 1e-3) * ipe / ceiling) if (ipe and kern_ms > 0) else None,
            "ncu_issue_active_frac": (counters or {}).get("issue_active_frac"),
            "ncu_lanes_per_instruction": (counters or {}).get("lanes_per_instruction"),
            "ncu_stall_no_instruction_per_issue": (counters or {}).get("stall_no_instruction_per_issue"),
            "source": (f"{COUNTERS.relative_to(ROOT)} ({counters_src})" if counters else None)}
    except Exception as exc:                                   # pragma: no cover
        roofline["issue_slots"] = f"unavailable: {exc!r}"

    # ---- end to end through the public host-buffer API, the SAME step as above: pinned H2D of the step's inputs
    # (replication keys), the simulation with its HBM log rings, the re-run of overflowed replications, the per-ring
    # reduction, the allreduce (N > 1) and the D2H of the per-replication accumulators and of the statistics vector
    e2e = None
    if not args.no_e2e:
        def e2e_step(step):
            set_seeds(step)
            stats.zero_()
            tot = 0
            for p in parts:
                out = p["eng"].run_host(p["n"], rep_seeds=p["seeds_host"], out=p["host"],
                                        device_log=p["res"] if log_cap else None, stats=stats,
                                        retry_overflow=not args.no_retry)
                tot += int(out.n_events.sum())
            if world > 1:
                dist.all_reduce(stats)
            stats_host.copy_(stats, non_blocking=True)
            torch.cuda.synchronize()
            return tot

        for w in range(max(1, min(args.warmup, 2))):
            e2e_step(w)
        torch.cuda.synchronize()
        e2e_events = 0
        e2e_s = 0.0
        for k in range(args.steps):
            flush_buf.fill_(k & 0xFF)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e2e_events += e2e_step(args.warmup + k)
            e2e_s += time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        ev = torch.tensor([e2e_events], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(ev)
        e2e = {"value": int(ev.item()) / float(t.item()), "unit": "events/s",
               "h2d_bytes_per_step": n_total_rank * 8 * world,
               "d2h_bytes_per_step": (n_total_rank * (K * 12 + 8 + 4 + 4) + K * 24) * world,
               "steps": args.steps, "logs": f"device rings of {log_cap} records per replication, as in `value`" if log_cap else "off",
               "api": "BatchEngine.run_host -> msim_run_host (pinned host buffers; log rings stay in HBM) "
                      "+ per-ring reduction" + (" + NCCL allreduce" if world > 1 else ""),
               "note": "host-buffer calls are synchronous, so the workload parts run one after the other here"}

    line = {"_cmd": "python bench.py " + " ".join(sys.argv[1:]),
            "metric": "simulated_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["desc"], "replications_per_gpu_per_step": n_total_rank,
                       "rng": "philox4x32-10", "log_ring_records_per_replication": log_cap,
                       "parts": "concurrent streams" if not args.serial_parts else "serial",
                       "l2": "flushed with a 256 MB write between timed steps"},
            "replications_per_sec": reps_total / (total_ms * 1e-3),
            "events_per_replication": events / reps_total,
            "e2e": e2e,
            "gpu_launches": launches[0] + retried[1],
            "failed_replications": int(fail_total.item()),
            "pool_overflow_replications_rerun": retried[0],
            "status_histogram": fail_codes.tolist(),
            "engine": {k: parts[0]["eng"].info[k] for k in ("smem_per_replication", "warps_per_block", "blocks_per_sm",
                                                             "regs_per_thread", "n_sms", "const_layout")},
            "roofline": roofline}
    if rank == 0:
        line["clocks"] = clocks
        if world == 1 and not args.no_cpu_baseline:
            from concurrent.futures import ProcessPoolExecutor

            from oracle.variates import run_seeds

            cores = os.cpu_count() or 1
            workers = min(cores, args.cpu_workers) if args.cpu_workers else cores
            seeds = run_seeds(BASE_SEED, workers)
            with ProcessPoolExecutor(max_workers=workers) as pool:
                e, wall, _ = cpu_pool_pass(pool, args.workload, args.run_until, len(spec["parts"]), seeds)
            line["cpu_baseline"] = dict(cpu_baseline_meta(), value=e / wall, unit="events/s", cores=workers,
                                        sample=f"{workers} full-length replications of the same workload, one per "
                                               f"worker process ({wall:.1f} s wall)",
                                        replications_per_sec=workers / wall)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
