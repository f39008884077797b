# bench lines of the other BASELINE workloads (docs only; the default bench is jobshop20)
O=gpurun_out/other_${1:-rXX}.jsonl; : > $O
python bench.py --workload simple_det --no-cpu-baseline 2>/dev/null | tail -1 >> $O
python bench.py --workload toc_dbr --run-until 6001 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 >> $O
python bench.py --workload toc_dbr --waves 1 --steps 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 >> $O
