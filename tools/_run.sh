timeout 900 python -m pytest tests -m gpu -x -q -k "padding or trunc or kats or span or properties" 2>&1 | tail -3
for f in 1 0; do
  TKZ_PAD_FILL=$f timeout 600 python bench.py --workload c3 --no-configs --no-e2e --no-cpu-baseline --no-strong --steps 5 --warmup 3 > gpurun_out/fill${f}_c3.json 2> gpurun_out/fill${f}_c3.err; python - <<PY
import json; d=json.loads(open("gpurun_out/fill${f}_c3.json").read().strip().splitlines()[-1]); print("c3 fill=$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["stage_ms_per_step"], d["parity"] and d["parity"]["ok"], d.get("hf_compat"))
PY
done
