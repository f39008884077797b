#!/usr/bin/env bash
# GPU: A/B of library builds / environment knobs on the bench workload (short horizon, parts one after the other).
# usage: bash tools/ab_bench.sh TAG "name|ENV=val ENV=val|extra bench args" ...
TAG=$1; shift
OUT=gpurun_out/ab_${TAG}.jsonl
: > "$OUT"
for spec in "$@"; do
  IFS='|' read name envs extra <<< "$spec"
  env $envs timeout 300 python bench.py --steps 2 --warmup 3 --run-until ${RU:-2000} --waves 3 --serial-parts --no-cpu-baseline --no-e2e $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'variant': '$name', 'value': d['value'], 'rerun': d['pool_overflow_replications_rerun'], 'failed': d['failed_replications'], 'parts': [{k: p[k] for k in ('delivery_mode','warps_per_sm','events_per_s_while_resident','waves','smem_per_replication','regs_per_thread')} for p in d['roofline']['kernel_by_part']]}))" >> "$OUT"
done
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print(d['variant'].ljust(14), '%.3e'%d['value'], 'rerun', d['rerun'], [(p['delivery_mode'],p['warps_per_sm'],p['regs_per_thread'],'%.3e'%p['events_per_s_while_resident']) for p in d['parts']])
PY
