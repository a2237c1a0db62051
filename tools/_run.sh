timeout 900 python -m pytest tests -x -q -m gpu -k "not 1gib and not at_size" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-verify --no-configs"
$B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2b o7', d['ms_per_step'], d['roofline']['stage_ms_per_step']); print({k:(v['ms_per_step'],v['stage_ms_per_step']) for k,v in d['device_variants'].items()})"
