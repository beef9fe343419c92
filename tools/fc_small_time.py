#!/usr/bin/env python
"""Isolated launch time of the small-footprint segment kernel (xvec_linear_small) vs the tcgen05 kernel with split-K on the
segment6 shape (256 x 3000 -> 512), bf16 and float32/TF32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops

def timeit(fn, n=50):
    for _ in range(5): fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2] * 1e3

for dtype in (torch.bfloat16, torch.float32):
    x = torch.randn(256, 3000, device="cuda").to(dtype)
    W = (torch.randn(512, 3000, device="cuda") / 55).to(dtype)
    b = torch.randn(512, device="cuda")
    wp = ops.pack_weight(W.float(), 1, 3000, dtype)
    ws = ops.splitk_workspace(256, 3000, 1, 512, dtype, "cuda")
    bp = ops.pad32(b)
    t_small = timeit(lambda: ops.linear_small(x, W, b))
    t_big = timeit(lambda: ops.tdnn_layer_flat(x, wp, 512, [0], bp, None, None, relu=False, out_dtype=torch.float32, cin=3000, workspace=ws))
    print(f"{dtype}: xvec_linear_small {t_small:.1f} us   tcgen05 split-K + reduce {t_big:.1f} us", flush=True)
