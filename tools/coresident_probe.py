#!/usr/bin/env python
"""Do the tail kernels of batch i really run NEXT TO the persistent stack kernel of batch i+1?  (ncu cannot show it: it serialises
kernels.)  A stack kernel is launched on stream A; while it runs, pool_finalize + fc_small (segment6) are launched on stream B
with CUDA events around them.  Co-resident: their elapsed time stays tens of microseconds; not co-resident: they wait for the
stack kernel to drain (~290 us)."""
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
    m.pooled_stats_flat(x, lengths)
    lay = m._layout_for(lengths)
    sc = m._scratch_for(0)
    sc1 = m._scratch_for(1)
    sc1.ensure(lay.rows, lay.n_slots, lay.n_utts)
    pipe = m._pipeline()
    part = sc.part[: lay.n_slots]
    W = m.segment_layer6.weight.detach().to(m.act_dtype).contiguous()
    b = m.segment_layer6.bias.detach().float().contiguous()
    A, B = torch.cuda.Stream(), torch.cuda.Stream()
    pooled, pooled_lp = sc.pooled[:256], (None if sc.pooled_lp is None else sc.pooled_lp[:256])

    def tail():
        ops.pool_finalize(part, lay.utt_slot_start, lay.n_pool, 1500, pipe["scale5"], pipe["shift5"], out=pooled, out_lp=pooled_lp)
        return ops.linear_small(pooled if pooled_lp is None else pooled_lp, W, b)

    def stack():
        ops.tdnn_stack(pipe["tdnn"], pipe["n_tdnn"], m._frames_for(x, pipe), sc1.act[0], sc1.act[1], lay.row_utt, lay.blk_slot_base,
                       sc1.part[: lay.n_slots], sc1.ctrl)

    for _ in range(3):
        stack(); tail()
    torch.cuda.synchronize()
    res = {"alone": [], "under_stack": [], "stack_alone": [], "stack_with_tail": []}
    for rep in range(30):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(B):
            e[0].record(B); tail(); e[1].record(B)
        torch.cuda.synchronize()
        res["alone"].append(e[0].elapsed_time(e[1]))
        with torch.cuda.stream(A):
            e[2].record(A); stack(); e[3].record(A)
        torch.cuda.synchronize()
        res["stack_alone"].append(e[2].elapsed_time(e[3]))
        with torch.cuda.stream(A):
            e[2].record(A); stack(); e[3].record(A)
        with torch.cuda.stream(B):
            e[0].record(B); tail(); e[1].record(B)
        torch.cuda.synchronize()
        res["under_stack"].append(e[0].elapsed_time(e[1]))
        res["stack_with_tail"].append(e[2].elapsed_time(e[3]))
    med = {k: sorted(v)[len(v) // 2] * 1e3 for k, v in res.items()}
    print(f"{precision}: tail (finalize + fc_small) alone {med['alone']:.1f} us | launched while a stack kernel runs {med['under_stack']:.1f} us | "
          f"stack kernel alone {med['stack_alone']:.1f} us | with the tail next to it {med['stack_with_tail']:.1f} us "
          f"(events include ~20-30 us of Python launch overhead per op)")
