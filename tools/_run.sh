timeout 900 python -m pytest tests -x -q -m gpu -k "fast or span_token or truncation or bpe_random" 2>&1 | tail -15
