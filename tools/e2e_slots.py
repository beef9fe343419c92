import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from oracle import xvector_oracle as ox
m = xvec_b200.XVectorModel(precision="bf16"); m.load_state_dict(ox.make_state_dict(0)); m = m.cuda().eval()
B, T = 256, 300
x_host = ox.synth_mfcc(1024, T, seed=1).reshape(4, B * T, 24).pin_memory()
lengths = [T] * B
for slots in (2, 3, 4, 6, 3):
    hx = xvec_b200.HostExtractor(m, n_slots=slots)
    for i in range(2 * slots): hx.result(hx.submit(x_host[i % 4], lengths))
    torch.cuda.synchronize()
    steps = 60; tickets = []; t0 = time.perf_counter(); chk = 0.0
    for i in range(steps):
        tickets.append(hx.submit(x_host[i % 4], lengths))
        if len(tickets) == slots: chk += float(hx.result(tickets.pop(0))[0, 0])
    while tickets: chk += float(hx.result(tickets.pop(0))[0, 0])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"slots {slots}: {B * steps / dt:,.0f} utt/s  {dt / steps * 1e3:.4f} ms/step")
