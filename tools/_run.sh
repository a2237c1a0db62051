( time timeout 900 python tools/stress_parity.py 400 5000 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys" | tee gpurun_out/stress_parity.txt
( time timeout 900 python tools/stress_grid.py 240 7000 2>&1 | tail -3 ) 2>&1 | grep -v "^$\|user\|sys" | tee gpurun_out/stress_grid.txt
( time timeout 1500 python tools/stress_corpus.py 512 2>&1 | tail -12 ) 2>&1 | grep -v "^$\|user\|sys" | tee gpurun_out/stress_corpus.txt
