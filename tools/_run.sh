timeout 900 python -m pytest tests -m gpu -x -q -k "windowed or grid or long_words or cascades or skewed or synthesised or bytelevel or whole or huge or malformed" 2>&1 | tail -3
for w in c2a c5a; do
  timeout 300 python bench.py --workload $w --no-configs --no-e2e --no-cpu-baseline --no-strong --steps 5 --warmup 3 > gpurun_out/bit_$w.json 2> gpurun_out/bit_$w.err; python - <<PY
import json; d=json.loads(open("gpurun_out/bit_$w.json").read().strip().splitlines()[-1]); print("$w", d["value"], d["roofline"]["stage_ms_per_step"], d["parity"] and d["parity"]["ok"])
PY
done
