B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify --no-configs"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v6_launches_c2b_1GiB.csv $B > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log
$B > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:slice_emit -s 3 -c 1 -o gpurun_out/r02_v6_passB $B > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
C="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-verify --no-configs --workload c2a --size-mib 256"
$C > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:bpe_block -s 15 -c 5 -o gpurun_out/r02_v6_block_c2a $C > gpurun_out/ncu3.log 2>&1; tail -1 gpurun_out/ncu3.log
