import sys, os
# the XVEC_DBG skip switches exist only in the debug library: python speaker-recognition-x-vectors_b200/build.py --debug; XVEC_LIB=.../libxvec_b200_debug.so
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
rows = 74 * 256 * 8
n = 512
for cin in (128, 512, 2048):
    x = torch.randn(rows, cin, device="cuda").bfloat16()
    w = ops.pack_weight(torch.randn(n, cin, device="cuda") / cin ** 0.5, 1, cin, torch.bfloat16)
    b = torch.zeros(n, device="cuda")
    out = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.tdnn_layer_flat(x, w, n, [0], b, None, None, relu=True, out=out))
    print(f"dbg={os.environ.get('XVEC_DBG','0')} K={cin:5d} ms={ms:.4f} cycles/tile@1965={ms*1e-3*1.965e9/16:8.0f}", flush=True)
