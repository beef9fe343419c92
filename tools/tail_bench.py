#!/usr/bin/env python
"""Tail of a batch (after the stack kernel): pool_finalize + fc_small (two launches) vs pool_fc_kernel (one launch), CUDA-event time of
the launches back to back on one stream (L2-warm: the partials were just written, as after a stack kernel), 256 x 300 workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xvec_b200
import bench
from xvec_b200 import ops

for precision in ("bf16", "tf32"):
    m = bench.synthetic_model(xvec_b200, precision).cuda().eval()
    lengths = [300] * 256
    x = torch.randn(256 * 300, 24, device="cuda")
    pooled, pooled_lp = m.pooled_stats_flat(x, lengths)  # fills the partials
    lay = m._layout_for(lengths)
    sc = m._scratch_for(0)
    pipe = m._pipeline()
    part = sc.part[: lay.n_slots]
    W = m.segment_layer6.weight.detach().to(m.act_dtype).contiguous()
    b = m.segment_layer6.bias.detach().float().contiguous()
    ws = torch.zeros(xvec_b200._lib.load().xvec_pool_fc_workspace_bytes(256, 1500, 512, xvec_b200._lib.dtype_code(m.act_dtype)), dtype=torch.uint8, device="cuda")

    def unfused():
        p32 = ops.pool_finalize(part, lay.utt_slot_start, lay.n_pool, 1500, pipe["scale5"], pipe["shift5"], out=sc.pooled[:256], out_lp=None if sc.pooled_lp is None else sc.pooled_lp[:256])
        a = p32 if sc.pooled_lp is None else sc.pooled_lp[:256]
        return ops.linear_small(a, W, b)

    def fused():
        return ops.pool_fc_fused(part, lay.utt_slot_start, lay.n_pool, 1500, W, b, pipe["scale5"], pipe["shift5"], workspace=ws)

    for name, fn in (("finalize + fc_small", unfused), ("pool_fc (fused)", fused)):
        for _ in range(5):
            fn()
        evs = []
        for _ in range(50):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b_) for a, b_ in evs)
        print(f"{precision} {name:22s}: median {t[len(t) // 2] * 1e3:7.1f} us  best {t[0] * 1e3:7.1f} us")
    d = (unfused().double() - fused().double()).abs().max().item()
    print(f"{precision} max |fused - unfused| = {d:.3e}")
