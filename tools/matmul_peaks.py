#!/usr/bin/env python
"""cuBLAS dense peaks of this GPU the way MEASURED_PEAKS.json takes them (torch.matmul 8192^3, 2*N^3 FLOP): best of 10 (burst)
and back to back for `secs` seconds (sustained, under the power cap) — for bf16 (cross-check of the driver's file) and for TF32
(float32 operands with torch.backends.cuda.matmul.allow_tf32 = True), which MEASURED_PEAKS.json does not carry."""
import json
import sys
import time

import torch


def measure(dtype, tf32, n=8192, secs=4.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    c = torch.empty(n, n, device="cuda", dtype=dtype)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    flop = 2.0 * n ** 3
    # sustained: batches of 20 launches until `secs` have passed; the last second's average
    t_end = time.perf_counter() + secs
    rates = []
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        rates.append((time.perf_counter(), 20 * flop / (e0.elapsed_time(e1) / 1e3) / 1e12))
    last = [r for t, r in rates if t > rates[-1][0] - 1.0]
    return {"burst_tflops": flop / (best / 1e3) / 1e12, "sustained_tflops": sum(last) / len(last)}


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    out = {"bf16": measure(torch.bfloat16, False, secs=secs), "tf32": measure(torch.float32, True, secs=secs),
           "how": "torch.matmul 8192^3 (cuBLAS), best of 10 and last second of a %.0f s back-to-back run" % secs,
           "gpu": torch.cuda.get_device_name(0)}
    print(json.dumps(out))
