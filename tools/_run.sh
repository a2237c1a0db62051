timeout 900 python -m pytest tests -x -q -m gpu -k "chunked or compact or packed_formats or multi_context or grid_kernel_on_the_skewed" 2>&1 | tail -3
for w in c2b c5b c2a; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-verify --no-configs --no-materialise --workload $w | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$w', 'device', round(d['value'],1), 'e2e', round(e['value'],2), round(e['ms_per_step'],1))"
done
