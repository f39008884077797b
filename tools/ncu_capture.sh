#!/usr/bin/env bash
# GPU (run under gpurun, ONE GPU): the two ncu passes of B200_PROFILING.md for the default bench workload, shortened so
# that the ~40 replays per kernel stay within a few GPU-minutes.  Each pass runs only after the same command has exited 0
# without ncu.  Outputs land in gpurun_out/ (scratch); summarise HERE afterwards into profiles/ with
#   python tools/summarize_profiles.py launches gpurun_out/launches_TAG.csv profiles/TAG_launches.csv
#   python tools/summarize_profiles.py full     gpurun_out/prof_TAG.ncu-rep profiles/TAG_ncu
#   python tools/summarize_profiles.py counters gpurun_out/prof_TAG.ncu-rep gpurun_out/plain_TAG.log jobshop20
#   python tools/ncu_lines.py gpurun_out/prof_TAG.ncu-rep EVENTS_PER_LAUNCH      # per-handler instructions / stalls
#
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r02'          # tag used in the output file names
set -u
TAG=${1:-rXX}
WL=${2:-jobshop20}
RU=${3:-1000}
OUT=gpurun_out
mkdir -p "$OUT"
# short but representative: one full wave of every workload part, 1/10 of the horizon, parts one after the other (ncu
# serialises kernels anyway), no overflow re-runs, no e2e leg
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --run-until $RU --waves 1 --serial-parts --no-retry --no-cpu-baseline --no-e2e"
KERNEL='regex:sim_kernel'          # msim::sim_kernel<policy,rng,opts,minb>

timeout 300 $CMD > "$OUT/plain_$TAG.log" 2>&1 || { echo "plain run failed, see $OUT/plain_$TAG.log"; exit 1; }
NPARTS=$(python -c "import json,sys; print(len(json.loads(open('$OUT/plain_$TAG.log').read().strip().splitlines()[-1])['roofline']['kernel_by_part']))")

# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches_$TAG.csv" $CMD > "$OUT/ncu_launches_$TAG.log" 2>&1

# the simulation kernel, full set: the launches of the timed step (after 3 warm-up steps of NPARTS launches each)
for i in $(seq 0 $((NPARTS - 1))); do
  timeout 900 ncu --set full --clock-control none --import-source on -k "$KERNEL" -s $((3 * NPARTS + i)) -c 1 \
      -o "$OUT/prof_${TAG}_part$i" -f $CMD > "$OUT/ncu_full_${TAG}_part$i.log" 2>&1
done
ls -la "$OUT" | grep "$TAG"
