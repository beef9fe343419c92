import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, xvec_b200
from xvec_b200 import ops
B, S = 256, 48000
for dt in (torch.float32, torch.int16):
    wav = (torch.randn(B * S, device="cuda") * (3000 if dt == torch.int16 else 1)).to(dt)
    lens = [S] * B
    for _ in range(3): ops.mfcc(wav, lens)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.mfcc(wav, lens)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"mfcc {dt}: {ms:.3f} ms per 256 x 3 s  ({B / ms * 1e3:,.0f} utt/s, {B * 299 / ms / 1e3:,.1f} M frames/s, reads {wav.numel() * wav.element_size() / ms / 1e6:,.0f} GB/s)")
