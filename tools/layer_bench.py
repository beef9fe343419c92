"""Per-layer kernel times of the c2 batch (256 x 300 frames), each layer timed alone over back-to-back launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, xvec_b200
from xvec_b200 import ops
torch.manual_seed(0)
B, T = int(os.environ.get("B", "256")), 300
rows = B * T
def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
lay = xvec_b200.build_layout([T] * B)
res = {}
x0 = torch.randn(rows, 24, device="cuda")
h = torch.randn(rows, 512, device="cuda").bfloat16()
out = torch.empty(rows, 512, device="cuda", dtype=torch.bfloat16)
specs = [("tdnn1", x0, 24, 512, [0, 1, 2, 3, 4], torch.float32), ("tdnn2", h, 512, 512, [0, 2, 4], torch.bfloat16),
         ("tdnn3", h, 512, 512, [0, 3, 6], torch.bfloat16), ("tdnn4", h, 512, 512, [0], torch.bfloat16)]
tot = 0
for name, x, cin, n, offs, dt in specs:
    w = ops.pack_weight(torch.randn(n, cin * len(offs), device="cuda") / (cin * len(offs)) ** 0.5, len(offs), cin, dt)
    b = torch.zeros(n, device="cuda")
    us = timeit(lambda: ops.tdnn_layer_flat(x, w, n, offs, b, None, None, relu=True, out=out, cin=cin))
    fl = 2.0 * (rows) * cin * len(offs) * n
    res[name] = us; tot += us
    print(f"{name}: {us:7.1f} us  {fl / us / 1e6:7.1f} TF", flush=True)
w5 = ops.pack_weight(torch.randn(1500, 512, device="cuda") / 512 ** 0.5, 1, 512, torch.bfloat16)
b5 = torch.zeros(1500, device="cuda")
part = torch.empty(lay.n_slots, 2, 1500, device="cuda")
ru, bs = torch.from_numpy(lay.row_utt).cuda(), torch.from_numpy(lay.blk_slot_base).cuda()
us = timeit(lambda: ops.tdnn_pool_fused(h, w5, 1500, [0], b5, ru, bs, part))
tot += us
print(f"tdnn5: {us:7.1f} us  {2.0 * rows * 512 * 1500 / us / 1e6:7.1f} TF")
print(f"sum {tot:.1f} us")
