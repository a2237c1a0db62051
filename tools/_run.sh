timeout 900 python -m pytest tests -x -q -m gpu -k "multi_context or compact or smoke or kats" 2>&1 | tail -8
