import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops
torch.manual_seed(0)
rows = 256 * 300
which = os.environ.get("LAYER", "tdnn4")
cin, n, offs = {"tdnn4": (512, 512, [0]), "tdnn2": (512, 512, [0, 2, 4])}[which]
h = torch.randn(rows, cin, device="cuda").bfloat16()
out = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
w = ops.pack_weight(torch.randn(n, cin * len(offs), device="cuda") / (cin * len(offs)) ** 0.5, len(offs), cin, torch.bfloat16)
b = torch.zeros(n, device="cuda")
for _ in range(4):
    ops.tdnn_layer_flat(h, w, n, offs, b, None, None, relu=True, out=out, cin=cin)
torch.cuda.synchronize()
print("ok")
