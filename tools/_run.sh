timeout 900 python -m pytest tests -x -q -m gpu -k "compact or packed_formats or truncation or chunked" 2>&1 | tail -12
