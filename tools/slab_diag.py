import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops
torch.manual_seed(0)
rows, cin, n = 512, 64, 256
for offs in ([0, 1], [0, 2], [0, 8], [0, 2, 4], [0, 3, 6]):
    x = torch.randn(rows, cin, device="cuda").bfloat16()
    W = torch.randn(n, cin * len(offs), device="cuda") / (cin * len(offs)) ** 0.5
    wp = ops.pack_weight(W, len(offs), cin, torch.bfloat16)
    y = ops.tdnn_layer_flat(x, wp, n, offs, None, None, None, relu=False, out_dtype=torch.float32)
    xp = torch.cat([x.float(), torch.zeros(16, cin, device="cuda")], 0)
    ref = torch.cat([xp[o:o + rows] for o in offs], 1) @ W.bfloat16().float().t()
    err = (y - ref).abs()
    bad_rows = (err.max(1).values > 0.05).nonzero().flatten()
    print(os.environ.get("XVEC_GEMM_MODE"), offs, "max err", err.max().item(), "bad rows", bad_rows.numel(), bad_rows[:24].tolist())
