"""Micro-benchmark: time of xvec_tdnn_layer vs K (taps=1) to separate mainloop rate from per-tile epilogue cost."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

rows = 74 * 256 * 8   # exactly 8 tiles per pair per n-tile
n = 512
for dtype in (torch.bfloat16, torch.float32):
    for out_dtype in (torch.bfloat16, torch.float32):
        for cin in (128, 256, 512, 1024, 2048, 4096):
            x = torch.randn(rows, cin, device="cuda").to(dtype)
            w = ops.pack_weight(torch.randn(n, cin, device="cuda") / cin ** 0.5, 1, cin, dtype)
            b = torch.zeros(n, device="cuda")
            out = torch.empty(rows, n, device="cuda", dtype=out_dtype)
            ms = timeit(lambda: ops.tdnn_layer_flat(x, w, n, [0], b, None, None, relu=True, out=out))
            tiles_per_pair = (rows // 256) * (n // 256) / 74
            cyc = ms * 1e-3 * 1.965e9 / tiles_per_pair
            print(f"in={str(dtype)[6:]:9s} out={str(out_dtype)[6:]:9s} K={cin:5d} ms={ms:.4f} TF={2*rows*cin*n/ms/1e9:8.1f} cycles/tile@1965={cyc:8.0f}", flush=True)
