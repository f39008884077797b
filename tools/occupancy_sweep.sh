#!/usr/bin/env bash
# GPU: throughput of the simulation kernel against resident CTAs per SM (MSIM_MAX_CTAS_PER_SM caps the occupancy the
# library would otherwise choose).  Tells latency-bound (throughput ~ warps) from fetch/issue-bound (saturates).
OUT=gpurun_out/occupancy_${1:-r02}.jsonl
: > "$OUT"
for c in 1 2 3 4 5; do
  MSIM_MAX_CTAS_PER_SM=$c timeout 200 python bench.py --steps 2 --warmup 3 --run-until 2000 --waves 2 --serial-parts --no-retry \
      --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'max_ctas_per_sm': $c, 'value': d['value'], 'parts': [{k: p[k] for k in ('delivery_mode','warps_per_sm','events_per_s_while_resident','waves','regs_per_thread')} for p in d['roofline']['kernel_by_part']]}))" >> "$OUT"
done
cat "$OUT"
