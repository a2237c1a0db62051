( time timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | grep -v "^$\|user\|sys"
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"],"e2e",d["e2e"]["value"])
for k,v in d["configs"].items(): print(k, round(v["value"],2), round(v["ms_per_step"],2), v.get("parity_checked_vs_oracle",{}) and v["parity_checked_vs_oracle"].get("ok"), (v.get("e2e") or {}).get("value"), v.get("hf_compat",{}).get("value"))
PY
