"""Run the TDNN2-shaped GEMM for ~2 s and sample SM clock / power through NVML to know the real clock under load."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml, xvec_b200
from xvec_b200 import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
rows, cin, n, offs = 256 * 300, 512, 512, [0, 2, 4]
x = torch.randn(rows, cin, device="cuda").bfloat16()
out = torch.empty(rows, n, device="cuda", dtype=torch.bfloat16)
w = ops.pack_weight(torch.randn(n, cin * 3, device="cuda") / (cin * 3) ** 0.5, 3, cin, torch.bfloat16)
b = torch.zeros(n, device="cuda")
samples = []
stop = False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.01)
th = threading.Thread(target=sampler); th.start()
for dur in (0.05, 0.5, 2.0):
    torch.cuda.synchronize(); samples.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); n_it = 0
    e0.record()
    while time.time() - t0 < dur:
        for _ in range(20):
            ops.tdnn_layer_flat(x, w, n, offs, b, None, None, relu=True, out=out, cin=cin)
        n_it += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n_it
    clk = sorted(s[0] for s in samples); pw = sorted(s[1] for s in samples)
    print(f"dur {dur}s: {us:.1f} us/launch  {2.0*rows*cin*3*n/us/1e6:.0f} TF  clk median {clk[len(clk)//2] if clk else None} min {clk[0] if clk else None} power median {pw[len(pw)//2] if pw else None:.0f} W  samples {len(clk)}")
stop = True; th.join()
