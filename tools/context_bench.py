"""Context numbers on the same B200 (NOT part of the product or of the bench line):
 1. cuBLAS (torch.matmul, bf16) on the GEMM shapes of the TDNN layers — what a library GEMM reaches on a materialised unfold.
 2. The reference's op sequence in PyTorch eager on the GPU (slice+cat unfold, Linear, ReLU, eval BatchNorm, mean/std pooling),
    fp32 with TF32 matmuls allowed and under bf16 autocast — what a user of the reference gets by just moving it to this GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
torch.manual_seed(0)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
rows = 256 * 300
for (k, n) in ((1536, 512), (512, 512), (512, 1536)):
    a = torch.randn(rows, k, device="cuda").bfloat16(); w = torch.randn(n, k, device="cuda").bfloat16()
    us = timeit(lambda: a @ w.t())
    print(f"cuBLAS bf16 M={rows} K={k} N={n}: {us:7.1f} us  {2.0 * rows * k * n / us / 1e6:7.1f} TF")

class Tdnn(nn.Module):
    def __init__(s, cin, n, ctx):
        super().__init__(); s.ctx = ctx; s.linear = nn.Linear(cin * len(ctx), n); s.norm = nn.BatchNorm1d(n)
    def forward(s, x):
        c = s.ctx; T = x.shape[1]; span = c[-1] - c[0]
        x = torch.cat([x[:, cj - c[0]: T - span + (cj - c[0]), :] for cj in c], 2)
        x = F.relu(s.linear(x))
        return s.norm(x.transpose(1, 2)).transpose(1, 2)
class Net(nn.Module):
    def __init__(s):
        super().__init__()
        s.t = nn.Sequential(Tdnn(24, 512, [-2, -1, 0, 1, 2]), Tdnn(512, 512, [-2, 0, 2]), Tdnn(512, 512, [-3, 0, 3]), Tdnn(512, 512, [0]), Tdnn(512, 1500, [0]))
        s.s6 = nn.Linear(3000, 512)
    def forward(s, x):
        h = s.t(x)
        return s.s6(torch.cat((h.mean(1), h.std(1)), 1))
net = Net().cuda().eval()
x = torch.randn(256, 300, 24, device="cuda")
with torch.no_grad():
    torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
    us = timeit(lambda: net(x), 10)
    print(f"PyTorch eager fp32 (TF32 matmul) batch 256x300: {us/1e3:.3f} ms/step  {256 / us * 1e6:,.0f} utt/s")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        us = timeit(lambda: net(x), 10)
    print(f"PyTorch eager bf16 autocast       batch 256x300: {us/1e3:.3f} ms/step  {256 / us * 1e6:,.0f} utt/s")
