#!/usr/bin/env python
"""Times the tdnn_stack_kernel launch alone (CUDA events, back-to-back launches) on the bench workload (256 x 300 frames) for a
sweep of band heights (explicit argument) and XVEC_STACK_DBG / XVEC_L2HINT settings (debug library only).  The DBG switches need the debug library:
    python speaker-recognition-x-vectors_b200/build.py --debug
    XVEC_LIB=$PWD/speaker-recognition-x-vectors_b200/libxvec_b200_debug.so python tools/stack_bench.py --dbg 0,1,2,3,4,7
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xvec_b200
from xvec_b200 import ops
from oracle import xvector_oracle as ox

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--band", default="0")
ap.add_argument("--dbg", default="0")
ap.add_argument("--iters", type=int, default=30)
args = ap.parse_args()

m = xvec_b200.XVectorModel(precision=args.dtype)
m.load_state_dict(ox.make_state_dict(seed=0))
m = m.cuda().eval()
n_res = 24
x = torch.randn(n_res, args.batch * args.frames, 24, device="cuda")
lengths = [args.frames] * args.batch
lay = m._layout_for(lengths)
sc = m._scratch_for(0)
sc.ensure(lay.rows, lay.n_slots, lay.n_utts)
pipe = m._pipeline()
part = sc.part[: lay.n_slots]
flops = args.batch * sum(f * (args.frames - l) for f, l in zip([122880, 1572864, 1572864, 524288, 1536000], [4, 8, 14, 14, 14]))


def run(iters, band=0):
    evs = []
    for it in range(iters + 3):
        xs = x[it % n_res]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.tdnn_stack(pipe["tdnn"], pipe["n_tdnn"], xs, sc.act[0], sc.act[1], lay.row_utt, lay.blk_slot_base, part, sc.ctrl, band=band)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs[3:])
    return sum(t) / len(t), t[0]


for band in args.band.split(","):
    for dbg in args.dbg.split(","):
        os.environ["XVEC_STACK_DBG"] = dbg
        avg, best = run(args.iters, int(band))
        cnt = sc.ctrl[:256].view(torch.int32).tolist()
        extra = ("  counters(spun,polls,fw,pub,mma_full,mma_tempty)=" + str([c for c in cnt[1:7]])) if any(cnt[1:7]) else ""
        if cnt[9]:
            extra += f" sm_clock={cnt[8] * 64 / cnt[9]:.3f} GHz cta0={cnt[9] / 1e3:.1f} us cta0_cycles={cnt[8] * 64} mma_step_cycles/pair={cnt[10] * 64 // 74} mma_ring_wait/pair={cnt[11] * 64 // 74} mma_full/pair={cnt[5] * 64 // 74} mma_tempty/pair={cnt[6] * 64 // 74}"
        if cnt[13]:
            extra += (f" epilogue warp cycles/tile (wait for accumulator, busy until release): pooled {cnt[12] * 64 // 1800} {cnt[13] * 64 // 1800}"
                      f" stored {cnt[14] * 64 // 2400} {cnt[15] * 64 // 2400} after release (all tiles) {cnt[7] * 64 // 4200}")
        if cnt[9]:
            tiles = [600, 600, 600, 600, 1800]
            ideal = [2560, 12288, 12288, 4096, 4096]
            extra += " per-layer cycles/tile (total incl. waits, tempty wait, ideal MMA): " + ", ".join(
                f"L{l + 1}: {cnt[24 + l] * 16 // tiles[l]} / {cnt[16 + l] * 16 // tiles[l]} / {ideal[l]}" for l in range(5))
            extra += " | item wait, K loop, operand waits, in-step: " + ", ".join(
                f"L{l + 1}: {cnt[32 + l] * 16 // tiles[l]} {cnt[40 + l] * 16 // tiles[l]} {cnt[48 + l] * 16 // tiles[l]} {cnt[56 + l] * 16 // tiles[l]}" for l in range(5))
        print(f"band={band:>5} dbg={dbg}{extra} avg {avg * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {flops / avg / 1e9:7.1f} TFLOP/s", flush=True)
