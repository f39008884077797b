for cfg in "--waves 2" "--waves 3" "--waves 4" "--waves 2 --serial-parts" "--waves 3 --serial-parts" "--waves 4 --serial-parts"; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $cfg 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); d['_cfg']='$cfg'; print(json.dumps(d))" >> gpurun_out/waves_r02t.jsonl
done
