#!/usr/bin/env python
"""Where the host time of one HostExtractor.submit goes (cProfile over 3000 batches of the bench workload) — the e2e leg's
per-batch host enqueue cost (bench.py: host.enqueue_us_per_batch) is what limits eight Python ranks on one box."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xvec_b200
import bench

m = bench.synthetic_model(xvec_b200, sys.argv[1] if len(sys.argv) > 1 else "bf16").cuda().eval()
x = torch.randn(4, 256 * 300, 24).pin_memory()
lengths = [300] * 256
hx = xvec_b200.HostExtractor(m, n_slots=6)
for i in range(12):
    hx.result(hx.submit(x[i % 4], lengths))


def loop(n):
    tickets = []
    for i in range(n):
        tickets.append(hx.submit(x[i % 4], lengths))
        if len(tickets) == 6:
            hx.result(tickets.pop(0))
    while tickets:
        hx.result(tickets.pop(0))


t0 = time.perf_counter()
loop(3000)
print(f"unprofiled: {(time.perf_counter() - t0) / 3000 * 1e6:.1f} us per batch (wall, GPU-bound if > host time)")
pr = cProfile.Profile()
pr.enable()
loop(3000)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(35)
