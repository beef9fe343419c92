#!/usr/bin/env python
"""SM clock, board power and throttle reasons (NVML, every 10 ms) while tdnn_stack_kernel runs back to back for ~2 s on the bench
workload: is the kernel power-capped?"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml, torch, xvec_b200
from xvec_b200 import ops
from oracle import xvector_oracle as ox

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
m = xvec_b200.XVectorModel(precision=sys.argv[1] if len(sys.argv) > 1 else "bf16"); m.load_state_dict(ox.make_state_dict(0)); m = m.cuda().eval()
B, T, NRES = 256, 300, 24
x = torch.randn(NRES, B * T, 24, device="cuda")
lengths = [T] * B
BAND = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # m-tiles per scheduling band (0 = the library's choice)
lay = m._layout_for(lengths); sc = m._scratch_for(0); sc.ensure(lay.rows, lay.n_slots, lay.n_utts)
pipe = m._pipeline(); part = sc.part[: lay.n_slots]
flops = B * sum(f * (T - l) for f, l in zip([122880, 1572864, 1572864, 524288, 1536000], [4, 8, 14, 14, 14]))
samples, stop = [], threading.Event()

def sampler():
    while not stop.is_set():
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.01)

def run(n):
    evs = []
    for it in range(n):
        xs = x[it % NRES]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.tdnn_stack(pipe["tdnn"], pipe["n_tdnn"], xs, sc.act[0], sc.act[1], lay.row_utt, lay.blk_slot_base, part, sc.ctrl, band=BAND)
        e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]

run(20)
th = threading.Thread(target=sampler, daemon=True); th.start()
t0 = time.perf_counter()
ms = run(6000)
t1 = time.perf_counter()
stop.set(); th.join()
limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
print(f"power limit {limit:.0f} W, max SM clock {pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz")
for lo, hi in ((0, 50), (50, 200), (200, 1000), (1000, 3000), (3000, 6000)):
    seg = ms[lo:hi]; avg = sum(seg) / len(seg)
    print(f"launches {lo:5d}-{hi:5d}: {avg * 1e3:7.1f} us/launch  {flops / avg / 1e9:7.1f} TFLOP/s")
in_run = [s for s in samples if t0 <= s[0] <= t1]
for frac in (0.02, 0.1, 0.3, 0.6, 0.95):
    s = in_run[min(len(in_run) - 1, int(frac * len(in_run)))]
    reasons = [n for n, b in (("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap), ("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown),
                              ("sw_thermal", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown), ("hw_thermal", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown)) if s[3] & b]
    print(f"t = {s[0] - t0:5.2f} s: SM {s[1]} MHz, {s[2]:6.1f} W, reasons {reasons}")
