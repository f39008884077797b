"""Summaries committed under profiles/: (1) per-kernel totals of an ncu launch list, (2) key metrics of a --set full
report, (3) per-source-region warp instructions per SimPy event."""
import collections, csv, subprocess, sys, json

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

def full(rep, out):
    raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.split('\n')))
    hdr, units, r = rows[0], rows[1], rows[2]
    pref = ('Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration', 'dram__bytes', 'gpu__dram_throughput', 'launch__',
            'sm__warps_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed', 'smsp__issue_active', 'sm__inst_executed.avg.per_cycle',
            'smsp__sass_average_branch', 'smsp__sass_branch_targets', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'l1tex__data_pipe_lsu_wavefronts_mem_shared',
            'sm__throughput', 'smsp__warps_eligible', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'smsp__average_warps_issue_stalled', 'lts__t_bytes', 'l1tex__t_bytes')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f); w.writerow(['metric', 'unit', 'value'])
        for i, h in enumerate(hdr):
            if any(h.startswith(p) for p in pref): w.writerow([h, units[i], r[i]])

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(sys.argv[2], sys.argv[3])
    else: full(sys.argv[2], sys.argv[3])
