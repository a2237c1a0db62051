( time python bench.py --steps 10 --warmup 3 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err ) 2>&1 | grep real
tail -5 gpurun_out/bench_new.err
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
