for c in 100 75 50 35 25; do
  TKZ_BB_CARVEOUT=$c timeout 300 python bench.py --workload c2a --no-configs --no-e2e --no-cpu-baseline --no-strong --no-verify --steps 4 --warmup 3 > gpurun_out/co_$c.json 2> gpurun_out/co_$c.err; python - <<PY
import json; d=json.loads(open("gpurun_out/co_$c.json").read().strip().splitlines()[-1]); print("carveout $c", d["value"], d["roofline"]["stage_ms_per_step"])
PY
done
