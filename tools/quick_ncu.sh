#!/bin/bash
# one-pass ncu counters for the simulation kernel on a short bench run (cheap: no --set full)
CMD="python bench.py --reps ${REPS:-8192} --steps 1 --run-until ${RU:-1000} --no-cpu-baseline ${EXTRA}"
$CMD > gpurun_out/q_plain.log 2>&1 && ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:sim_kernel -s 6 -c 2 --csv --log-file gpurun_out/q_ncu.csv $CMD > gpurun_out/q_ncu.log 2>&1
python - <<PY
import csv, json
rows=[r for r in csv.reader(open("gpurun_out/q_ncu.csv")) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    print(r[hdr.index("Kernel Name")][:40], r[hdr.index("Grid Size")], r[hdr.index("Metric Name")], r[hdr.index("Metric Value")])
d=json.loads(open("gpurun_out/q_plain.log").read().strip().splitlines()[-1])
print("value", d["value"], "events/rep", d["events_per_replication"], d["engine"])
PY
