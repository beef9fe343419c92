import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from oracle import xvector_oracle as ox
m = xvec_b200.XVectorModel(precision="bf16"); m.load_state_dict(ox.make_state_dict(0)); m = m.cuda().eval()
for B in (256, 64, 16):
    x = ox.synth_mfcc(B, 300, seed=1).reshape(B * 300, 24).cuda(); lengths = [300] * B
    for _ in range(5): m.extract_x_vec_flat(x, lengths)
    torch.cuda.synchronize()
    # CPU enqueue time: GPU kept busy by a long dummy kernel so launches never block on a full queue
    big = torch.empty(1 << 28, device="cuda")
    big.zero_(); 
    t0 = time.perf_counter()
    for _ in range(20): m.extract_x_vec_flat(x, lengths)
    t_cpu = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): m.extract_x_vec_flat(x, lengths)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: CPU enqueue {t_cpu * 1e6:.0f} us/step, GPU {e0.elapsed_time(e1) / 20 * 1e3:.0f} us/step")
