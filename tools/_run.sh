N=8
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_n$N.err
python tools/pcie_probe.py --gpus 1,2,4,8 --mib 256 | tee gpurun_out/r02_pcie_probe_8gpu.json
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo_8gpu.txt
