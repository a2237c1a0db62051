set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "windowed or grid or long_words or cascades or skewed or synthesised or bytelevel or whole" 2>&1 | tail -5
for w in c2a c5a c5b; do
 for s in 1 0; do
  TKZ_BLOCK_STREAMS=$s timeout 300 python bench.py --workload $w --no-configs --no-e2e --no-cpu-baseline --no-strong --steps 5 --warmup 3 > gpurun_out/st${s}_$w.json 2> gpurun_out/st${s}_$w.err; python - <<PY
import json; d=json.loads(open("gpurun_out/st${s}_$w.json").read().strip().splitlines()[-1]); print("$w streams=$s", d["value"], d["roofline"]["stage_ms_per_step"], d["parity"] and d["parity"]["ok"])
PY
 done
done
