#!/usr/bin/env python3
"""Measures pinned host<->device copy bandwidth on this box: H2D alone, D2H alone, both at once on two streams.
The end-to-end line of bench.py is bounded by these numbers (tkz_encode_batch moves text in and the encoding out)."""
import json
import time
import torch

n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9


for _ in range(2):
    run(True, True, 1)
res = {"h2d_alone_GBps": run(True, False), "d2h_alone_GBps": run(False, True), "both_each_GBps": run(True, True)}
res["both_aggregate_GBps"] = 2 * res["both_each_GBps"]
print(json.dumps(res))
