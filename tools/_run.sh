timeout 900 python -m pytest tests/test_gpu_hf_compat.py -x -q 2>&1 | tail -15
timeout 1200 python -m pytest tests -m gpu -x -q -k "not hf_compat and (truncation or padding or kats or occurrence or struct or span)" 2>&1 | tail -3
