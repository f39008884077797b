#!/usr/bin/env bash
# GPU: throughput of the simulation kernel against its launch shape -- resident CTAs per SM (MSIM_MAX_CTAS_PER_SM), CTA
# width (MSIM_WARPS_PER_CTA) and register build (MSIM_WIDE) -- on the bench workload, parts one after the other.
# usage: bash tools/occupancy_sweep.sh TAG "w:ctas:shape w:ctas:shape ..."   (0 = library default for that knob)
OUT=gpurun_out/occupancy_${1:-r02}.jsonl
SHAPES=${2:-"8:1:-1 8:2:-1 8:3:-1 8:4:-1 8:5:-1"}
WL=${3:-jobshop20}
RU=${4:-2000}
: > "$OUT"
for s in $SHAPES; do
  IFS=: read w c shape <<< "$s"
  env $( [ "$w" != 0 ] && echo MSIM_WARPS_PER_CTA=$w ) $( [ "$c" != 0 ] && echo MSIM_MAX_CTAS_PER_SM=$c ) $( [ "$shape" != -1 ] && echo MSIM_SHAPE=$shape ) \
    timeout 200 python bench.py --workload $WL --steps 2 --warmup 3 --run-until $RU --waves 3 --serial-parts --no-retry \
      --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'shape': '$s', 'value': d['value'], 'parts': [{k: p[k] for k in ('delivery_mode','warps_per_sm','events_per_s_while_resident','waves','smem_per_replication')} for p in d['roofline']['kernel_by_part']]}))" >> "$OUT"
done
cat "$OUT"
