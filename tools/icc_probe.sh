#!/usr/bin/env bash
# GPU: instruction-cache counters of the simulation kernel for several library builds (one cheap ncu pass each):
# usage: bash tools/icc_probe.sh TAG "name|ENV=..|bench args" ...
TAG=$1; shift
OUT=gpurun_out/icc_${TAG}.txt
: > "$OUT"
M=sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
for spec in "$@"; do
  IFS='|' read name envs extra <<< "$spec"
  env $envs timeout 600 ncu --metrics $M --clock-control none -k regex:sim_kernel -s 3 -c 1 --csv --log-file gpurun_out/icc_tmp.csv \
     python bench.py --steps 1 --warmup 3 --run-until 1000 --waves 1 --serial-parts --no-retry --no-cpu-baseline --no-e2e $extra > /dev/null 2>&1
  python - "$name" >> "$OUT" <<'PY'
import csv, sys, json
rows=[r for r in csv.reader(open("gpurun_out/icc_tmp.csv")) if len(r) > 10]
h=rows[0]; out={"variant": sys.argv[1]}
for r in rows[1:]:
    out[r[h.index("Metric Name")]] = r[h.index("Metric Value")]
print(json.dumps(out))
PY
done
cat "$OUT"
