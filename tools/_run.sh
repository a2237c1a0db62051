timeout 600 python -m pytest tests -x -q -m gpu -k "not 1gib" 2>&1 | tail -4
B="python bench.py --size-mib 512 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-verify"
$B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bulk', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
TKZ_STAGE=ldg $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ldg', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
TKZ_CLASSIFY=lut $B | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bulk+lut', d['ms_per_step'], d['roofline']['stage_ms_per_step'])"
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:slice_words -s 3 -c 1 -o gpurun_out/r02_v2_passA $B > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
