#!/usr/bin/env python
"""Small driver for the ncu captures of the two kernels bench.py's roofline objects quote besides the bf16 stack kernel:
stats_pool_partial_kernel on the long-form activation of the pooling roofline (64 x 5986 x 1500 float32) and the TF32 stack kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xvec_b200
import bench

m = bench.synthetic_model(xvec_b200, "tf32").cuda().eval()
a = torch.randn(64, 5986, 1500, device="cuda")
for _ in range(3):
    m.stat_pool(a)
del a
x = torch.randn(256, 300, 24, device="cuda")
for _ in range(4):
    m.extract_x_vec(x)
torch.cuda.synchronize()
print("ok")
