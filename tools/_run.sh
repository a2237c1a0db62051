nvidia-smi -L
timeout 300 python -m pytest tests -x -q -m gpu -k "multi_context" 2>&1 | tail -3
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err ) 2>&1 | grep real
tail -5 gpurun_out/bench_n2.err
