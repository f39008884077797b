#!/usr/bin/env bash
# GPU (run under gpurun, ONE GPU): the two ncu passes of B200_PROFILING.md for the default bench workload, shortened so
# that the ~40 replays per kernel stay within a few GPU-minutes.  Each pass runs only after the same command has exited 0
# without ncu.  Outputs land in gpurun_out/ (scratch); copy the summaries you want judged into profiles/.
#
#   gpurun --timeout 600 -- 'bash tools/ncu_capture.sh r02'          # tag used in the output file names
#
# Read here (no GPU) with:
#   ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv | grep -E 'dram__bytes_(read|write)\.sum|smsp__inst_executed\.sum|
#        smsp__issue_active|stalled_no_instruction|smsp__thread_inst_executed_per_inst_executed|l1tex__data_bank_conflicts'
#   ncu -i gpurun_out/prof_<tag>.ncu-rep --page source --csv      # per-line instruction counts / stall reasons (-lineinfo)
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p "$OUT"
# short but representative: both jobshop20 parts, 8192 replications each, 1/10 of the horizon, no overflow re-runs
CMD="python bench.py --steps 2 --warmup 3 --run-until 1000 --reps 16384 --no-retry --no-cpu-baseline"
KERNEL='regex:sim_kernel'          # msim::aos::sim_kernel<policy,rng,opts> (default) or msim::sim_kernel<policy,rng>

timeout 300 $CMD > "$OUT/plain_$TAG.log" 2>&1 || { echo "plain run failed, see $OUT/plain_$TAG.log"; exit 1; }

# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches_$TAG.csv" $CMD > "$OUT/ncu_launches_$TAG.log" 2>&1

# the simulation kernel, full set, 2 launches (one asReady, one onDue) after the warm-up launches
timeout 900 ncu --set full --clock-control none --import-source on -k "$KERNEL" -s 6 -c 2 \
    -o "$OUT/prof_$TAG" -f $CMD > "$OUT/ncu_full_$TAG.log" 2>&1
ls -la "$OUT" | grep "$TAG"
