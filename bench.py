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
                "pools": [(112, 112), (112, 176)],     # (production-order, demand-order) pool sizes; onDue keeps orders until due
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
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": workers, "kind": "port", "sample": sample,
                             "engine": "oracle/factory_oracle.py on oracle/minisimpy.py (SimPy 4.1.1 restated; "
                                       "real SimPy is not installable offline)"},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="jobshop20")
    ap.add_argument("--reps", type=int, default=16384, help="replications per GPU per step")
    ap.add_argument("--run-until", type=int, default=None)
    ap.add_argument("--log-cap", type=int, default=4096, help="records per replication HBM ring (0 = no log streaming)")
    ap.add_argument("--cpu-workers", type=int, default=0)
    ap.add_argument("--max-po", type=int, default=0, help="production-order pool per replication (0 = library default)")
    ap.add_argument("--max-do", type=int, default=0, help="demand-order pool per replication (0 = library default)")
    ap.add_argument("--fifo", type=int, default=0, help="zero-delay event ring capacity (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retry", action="store_true", help="do not re-run pool-overflow replications with big pools")
    args = ap.parse_args()
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
        n = args.reps // period
        parts.append({"eng": eng, "n": n, "phase": phase, "period": period})
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
    ev_total = torch.zeros((), dtype=torch.int64, device=dev)
    fail_total = torch.zeros((), dtype=torch.int64, device=dev)
    fail_codes = torch.zeros((9,), dtype=torch.int64, device=dev)
    rec_total = torch.zeros((), dtype=torch.int64, device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    side = [torch.cuda.Stream(dev) for _ in parts]

    def set_seeds(step):
        for p in parts:
            gidx = (np.arange(p["n"], dtype=np.int64) * p["period"] + p["phase"] + rank * args.reps)
            keys = rep_keys(BASE_SEED + step, gidx)
            p["seeds_host"].copy_(torch.from_numpy(keys.view(np.int64)))

    sim_events = []
    retried = [0, 0]

    def device_step(step, timed):
        set_seeds(step)
        stats.zero_()
        for p in parts:
            p["seeds_dev"].copy_(p["seeds_host"], non_blocking=True)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            p["eng"].run(p["n"], rep_seeds=p["seeds_dev"], out=p["res"], stream=stream)
            e1.record(stream)
            if timed:
                sim_events.append((e0, e1, p))
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
    clocks = sampler.stop() if rank == 0 else None
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

    # dominant kernel: the simulation kernel(s); algorithmic bytes per SURVEY 8(d)
    kern_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in sim_events)
    n_launch = len(sim_events)
    records_rank = int(rec_total.item())
    alg_bytes = 8 * records_rank + (16 * K + 8) * n_total_rank * args.steps
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak = float(json.loads(peaks_file.read_text())["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    # DRAM traffic per algorithmic byte measured once with `ncu --set full` (profiles/r01_ncu_full_final.csv:
    # dram read 0.109 GB + write 4.624 GB for 2.341 GB algorithmic): records are streamed as 16-byte slots, twice the
    # 8 bytes the reference's (time,value) float32 pair needs
    traffic_per_alg_byte = 2.02 if log_cap else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic_per_alg_byte * alg_bytes / max(n_launch, 1)) if traffic_per_alg_byte else None,
                "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum of one --set full capture, scaled by "
                                  "algorithmic bytes (profiles/r01_ncu_full_final.csv)",
                "peak_source": peak_src,
                "kernel": ("msim::sim_kernel<policy,rng>" if os.environ.get("MSIM_SLAB_LAYOUT") == "soa"
                           else "msim::aos::sim_kernel<policy,rng,opts>"),
                "kernel_ms_per_launch": kern_ms / max(n_launch, 1), "launches": n_launch,
                "algorithmic_bytes_per_launch": alg_bytes / max(n_launch, 1),
                "note": "path is instruction-fetch / issue bound, not HBM bound (SURVEY 8(d)): 86 warp instructions per "
                        "SimPy event, issue slots 40 % busy, stall_no_instruction dominant (ncu capture of the SoA "
                        "slab build, profiles/README.md; the AoS slab that is now the default was A/B-timed only)",
                "kernel_share_of_step": kern_ms / sum(step_ms)}

    # per workload part (e.g. the asReady and the onDue half of jobshop20): simulation-kernel time and events/s of this
    # rank over the timed steps -- reporting only, never allowed to break the line
    try:
        by_part = []
        for i, p in enumerate(parts):
            ms = sum(e0.elapsed_time(e1) for e0, e1, q in sim_events if q is p)
            evp = int(p["ev_part"].item())
            by_part.append({"part": i, "delivery_mode": spec["parts"][i][0][0].get("delivery_mode"),
                            "replications_per_launch": p["n"], "kernel_ms_total": ms,
                            "events_per_s": evp / (ms * 1e-3) if ms > 0 else None,
                            "warps_per_sm": p["eng"].info["warps_per_block"] * p["eng"].info["blocks_per_sm"],
                            "smem_per_replication": p["eng"].info["smem_per_replication"]})
        roofline["kernel_by_part"] = by_part
    except Exception as exc:                                   # pragma: no cover
        roofline["kernel_by_part"] = f"unavailable: {exc!r}"

    # the ceiling that actually bounds this path: warp-instruction issue slots = SMs x 4 schedulers x clock
    try:
        mhz = float((clocks or {}).get("sm_mhz") or 1965.0)
        n_sms = parts[0]["eng"].info["n_sms"]
        ceiling = n_sms * 4 * mhz * 1e6
        ipe = 86.0
        roofline["issue_slots"] = {
            "ceiling_warp_instr_per_s": ceiling, "warp_instr_per_event": ipe,
            "warp_instr_per_event_source": "ncu smsp__inst_executed.sum / SimPy events, SoA slab build "
                                           "(profiles/README.md); not re-measured for the AoS default",
            "busy_frac_estimate": (events / world) / (kern_ms * 1e-3) * ipe / ceiling if kern_ms > 0 else None,
            "ncu_issue_active_frac": 0.399}
    except Exception as exc:                                   # pragma: no cover
        roofline["issue_slots"] = f"unavailable: {exc!r}"

    # ---- end to end through the public host-buffer API: pinned H2D of the step's inputs + D2H of its results
    def e2e_step(step):
        set_seeds(step)
        tot = 0
        for p in parts:
            out = p["eng"].run_host(p["n"], rep_seeds=p["seeds_host"], out=p["host"])
            tot += int(out.n_events.sum())
        return tot

    e2e_step(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_events = 0
    e2e_steps = min(args.steps, 3)            # same step, fewer repetitions: keeps the default run within minutes
    for k in range(e2e_steps):
        e2e_events += e2e_step(args.warmup + k)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    ev = torch.tensor([e2e_events], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev)
    e2e_value = int(ev.item()) / float(t.item())
    h2d = n_total_rank * 8 * world
    d2h = n_total_rank * (K * 12 + 8 + 4 + 4) * world

    line = {"metric": "simulated_events_per_sec", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": spec["desc"], "replications_per_gpu_per_step": n_total_rank,
                       "rng": "philox4x32-10", "log_ring_records_per_replication": log_cap,
                       "slab_layout": os.environ.get("MSIM_SLAB_LAYOUT", "aos"),
                       "l2": "flushed with a 256 MB write between timed steps"},
            "replications_per_sec": reps_total / (total_ms * 1e-3),
            "events_per_replication": events / reps_total,
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "BatchEngine.run_host -> msim_run_host (pinned host buffers)"},
            "gpu_launches": (2 * len(parts)) * args.steps + retried[1],
            "failed_replications": int(fail_total.item()),
            "pool_overflow_replications_rerun": retried[0],
            "status_histogram": fail_codes.tolist(),
            "engine": {k: parts[0]["eng"].info[k] for k in ("smem_per_replication", "warps_per_block", "blocks_per_sm",
                                                             "regs_per_thread", "n_sms")},
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
            line["cpu_baseline"] = {"value": e / wall, "unit": "events/s", "cores": workers, "kind": "port",
                                    "sample": f"{workers} full-length replications of the same workload, one per "
                                              f"worker process ({wall:.1f} s wall)",
                                    "replications_per_sec": workers / wall}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
