timeout 900 python -m pytest tests -x -q -m gpu -k "decode or tokens_come or example_basic or multi_context" 2>&1 | tail -5
