#!/usr/bin/env python3
"""Measures pinned host<->device copy bandwidth on this box: H2D alone, D2H alone, both at once on two streams -- on ONE GPU and
on the first k GPUs AT THE SAME TIME (one thread per GPU).  The end-to-end lines of bench.py are bounded by these numbers
(the host-buffer calls move the text in and the encoding out): e2e at N GPUs cannot exceed the aggregate H2D rate with D2H
running beside it.      usage: python tools/pcie_probe.py [--gpus 1,2,4,8] [--mib 512]"""
import argparse
import json
import threading
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", default="1")
ap.add_argument("--mib", type=int, default=1024)
args = ap.parse_args()
n = args.mib << 20
ng_all = torch.cuda.device_count()


class Dev:
    def __init__(self, i):
        self.i = i
        with torch.cuda.device(i):
            self.h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            self.h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            self.d_a = torch.empty(n, dtype=torch.uint8, device=f"cuda:{i}")
            self.d_b = torch.empty(n, dtype=torch.uint8, device=f"cuda:{i}")
            self.s1, self.s2 = torch.cuda.Stream(i), torch.cuda.Stream(i)

    def issue(self, h2d, d2h, reps):
        with torch.cuda.device(self.i):
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(self.s1):
                        self.d_a.copy_(self.h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(self.s2):
                        self.h_out.copy_(self.d_b, non_blocking=True)

    def sync(self):
        torch.cuda.synchronize(self.i)


def run(devs, h2d, d2h, reps=4):
    for d in devs:
        d.sync()
    t0 = time.perf_counter()
    th = [threading.Thread(target=d.issue, args=(h2d, d2h, reps)) for d in devs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for d in devs:
        d.sync()
    dt = (time.perf_counter() - t0) / reps
    return len(devs) * n / dt / 1e9            # aggregate GB/s per direction


out = {}
devs_all = [Dev(i) for i in range(min(ng_all, max(int(x) for x in args.gpus.split(","))))]
for k in [int(x) for x in args.gpus.split(",")]:
    if k > len(devs_all):
        continue
    devs = devs_all[:k]
    run(devs, True, True, 1)
    out[str(k)] = {"h2d_alone_aggregate_GBps": run(devs, True, False), "d2h_alone_aggregate_GBps": run(devs, False, True),
                   "both_aggregate_each_direction_GBps": run(devs, True, True)}
print(json.dumps({"buffer_mib": args.mib, "gpus": out}))
