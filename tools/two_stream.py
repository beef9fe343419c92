import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from oracle import xvector_oracle as ox
m = xvec_b200.XVectorModel(precision="bf16"); m.load_state_dict(ox.make_state_dict(0)); m = m.cuda().eval()
B, T = 256, 300
x = ox.synth_mfcc(1024, T, seed=1).reshape(4, B * T, 24).cuda()
lengths = [T] * B
def run(nstreams, steps=40):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for i in range(8):
        with torch.cuda.stream(streams[i % nstreams]): m.extract_x_vec_flat(x[i % 4], lengths, slot=i % nstreams)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        with torch.cuda.stream(streams[i % nstreams]): m.extract_x_vec_flat(x[i % 4], lengths, slot=i % nstreams)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3
for ns in (1, 2, 3, 1, 2):
    r, ms = run(ns)
    print(f"streams {ns}: {r:,.0f} utt/s  {ms:.4f} ms/step")
