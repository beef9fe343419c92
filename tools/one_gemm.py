import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops
cin = int(os.environ.get("K", "512")); n = 512; rows = 74 * 256 * 8
x = torch.randn(rows, cin, device="cuda").bfloat16()
w = ops.pack_weight(torch.randn(n, cin, device="cuda") / cin ** 0.5, 1, cin, torch.bfloat16)
b = torch.zeros(n, device="cuda")
out = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.tdnn_layer_flat(x, w, n, [0], b, None, None, relu=True, out=out)
torch.cuda.synchronize()
print("ok")
