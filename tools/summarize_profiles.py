"""Summaries committed under profiles/ (all run HERE, without a GPU, on files a `gpurun` call brought back):

    python tools/summarize_profiles.py launches  gpurun_out/launches_TAG.csv   profiles/TAG_launches.csv
    python tools/summarize_profiles.py full      gpurun_out/prof_TAG.ncu-rep   profiles/TAG_ncu            # one CSV per captured launch
    python tools/summarize_profiles.py counters  gpurun_out/prof_TAG_part0.ncu-rep,gpurun_out/prof_TAG_part1.ncu-rep  gpurun_out/plain_TAG.log  WORKLOAD

`counters` writes profiles/kernel_counters.json -- the ONLY place bench.py takes ncu-derived numbers from (warp
instructions per SimPy event, issue-slot utilisation, lanes per instruction, DRAM bytes per algorithmic byte): it joins
the captured simulation-kernel launches of the `--set full` report (tools/ncu_capture.sh: serial parts, one launch per
workload part) with the per-part event counts and algorithmic bytes of the plain run of the same command."""
import collections, csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi: continue
        v = float(r[vi].replace(',', '')); u = r[ui]
        v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
        agg[r[ki][:80]][0] += 1; agg[r[ki][:80]][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1]:.1f},{v[1]/tot:.5f}\n")

PREF = ('Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration', 'dram__bytes', 'gpu__dram_throughput', 'launch__',
        'sm__warps_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed', 'smsp__issue_active', 'sm__inst_executed.avg.per_cycle',
        'smsp__sass_average_branch', 'smsp__sass_branch_targets', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'l1tex__data_pipe_lsu_wavefronts_mem_shared',
        'sm__throughput', 'smsp__warps_eligible', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled', 'lts__t_bytes', 'l1tex__t_bytes')

def raw_rows(rep):
    raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(raw.split('\n')) if r]
    return rows[0], rows[1], rows[2:]

def full(rep, out_prefix):
    hdr, units, data = raw_rows(rep)
    for k, r in enumerate(data):
        with open(f"{out_prefix}_launch{k}.csv", 'w', newline='') as f:
            w = csv.writer(f); w.writerow(['metric', 'unit', 'value'])
            for i, h in enumerate(hdr):
                if any(h.startswith(p) for p in PREF): w.writerow([h, units[i], r[i]])

def num(hdr, units, r, name):
    if name not in hdr: return None
    i = hdr.index(name)
    try: v = float(r[i].replace(',', ''))
    except ValueError: return None
    if v != v: return None
    u = units[i]
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 'msecond': 1e-3, 'usecond': 1e-6, 'nsecond': 1e-9, 'second': 1.0}.get(u, 1.0)
    return v * scale

def counters(rep, plain_log, workload):
    """rep: one report holding one launch per workload part, or a comma-separated list of one-launch reports in part order"""
    line = json.loads([l for l in open(plain_log).read().splitlines() if l.startswith('{')][-1])
    parts = line["roofline"]["kernel_by_part"]
    rows = []                                  # (header, units, row) per captured launch
    for one in rep.split(","):
        h, u, d = raw_rows(one)
        rows += [(h, u, r) for r in d]
    rep = rep.split(",")[0]
    per_part = []
    for k, (hdr, units, r) in enumerate(rows):
        if k >= len(parts): break
        p = parts[k]
        ev, alg = p["events_per_launch"], p["algorithmic_bytes_per_launch"]
        inst = num(hdr, units, r, 'smsp__inst_executed.sum')
        dr, dw = num(hdr, units, r, 'dram__bytes_read.sum'), num(hdr, units, r, 'dram__bytes_write.sum')
        per_part.append({
            "part": k, "delivery_mode": p.get("delivery_mode"), "kernel": r[hdr.index('Kernel Name')],
            "grid": r[hdr.index('Grid Size')], "block": r[hdr.index('Block Size')],
            "events_per_launch": ev, "algorithmic_bytes_per_launch": alg,
            "gpu_time_s_under_ncu": num(hdr, units, r, 'gpu__time_duration.sum'),
            "warp_instr": inst, "warp_instr_per_event": inst / ev if inst else None,
            "issue_active_frac": (num(hdr, units, r, 'smsp__issue_active.avg.pct_of_peak_sustained_active') or 0) / 100 or None,
            "lanes_per_instruction": num(hdr, units, r, 'smsp__thread_inst_executed_per_inst_executed.ratio'),
            "stall_no_instruction_per_issue": num(hdr, units, r, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),
            "stall_short_scoreboard_per_issue": num(hdr, units, r, 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio'),
            "stall_wait_per_issue": num(hdr, units, r, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
            "branch_targets_uniform_frac": (num(hdr, units, r, 'smsp__sass_average_branch_targets_threads_uniform.pct') or 0) / 100 or None,
            "shared_bank_conflicts": num(hdr, units, r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'),
            "warps_active_frac": (num(hdr, units, r, 'sm__warps_active.avg.pct_of_peak_sustained_active') or 0) / 100 or None,
            "dram_bytes": (dr + dw) if dr is not None and dw is not None else None,
            "dram_bytes_per_algorithmic_byte": (dr + dw) / alg if dr is not None and dw is not None else None,
            "registers_per_thread": num(hdr, units, r, 'launch__registers_per_thread')})
    ok = [p for p in per_part if p["warp_instr"]]
    ev = sum(p["events_per_launch"] for p in ok)
    def wavg(key):
        v = [(p[key], p["events_per_launch"]) for p in ok if p.get(key) is not None]
        return sum(a * b for a, b in v) / sum(b for _, b in v) if v else None
    agg = {"warp_instr_per_event": sum(p["warp_instr"] for p in ok) / ev if ok else None,
           "issue_active_frac": wavg("issue_active_frac"), "lanes_per_instruction": wavg("lanes_per_instruction"),
           "stall_no_instruction_per_issue": wavg("stall_no_instruction_per_issue"),
           "dram_bytes_per_algorithmic_byte": (sum(p["dram_bytes"] for p in ok if p["dram_bytes"]) /
                                               sum(p["algorithmic_bytes_per_launch"] for p in ok if p["dram_bytes"])) if any(p["dram_bytes"] for p in ok) else None,
           "parts": per_part}
    path = ROOT / "profiles" / "kernel_counters.json"
    doc = json.loads(path.read_text()) if path.exists() else {"workloads": {}}
    doc["source"] = f"ncu --set full --clock-control none capture {Path(rep).name} of `{line.get('_cmd', 'python bench.py ...')}` (shortened horizon; per-event ratios)"
    doc["workloads"][workload] = agg
    doc.setdefault("captures", {})[workload] = {"report": Path(rep).name, "plain_run_config": line["config"]}
    path.write_text(json.dumps(doc, indent=1))
    print(json.dumps({k: v for k, v in agg.items() if k != "parts"}))

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "counters": counters(sys.argv[2], sys.argv[3], sys.argv[4])
    else: full(sys.argv[2], sys.argv[3])
