timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-verify"
$B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2b o7', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --outputs 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2b o1', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --outputs 33 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2b o33', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --workload c3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 o7', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --workload c3 --outputs 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c3 o1', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --workload c4b | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4b o7', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B --workload c5b | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5b o7', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
