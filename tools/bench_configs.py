"""Throughput of every BASELINE.json config on one B200, both precisions (fills the table in DESIGN.md §5).

device: flat frame matrices already resident in HBM, length-bucketed batches, CUDA-event time over the whole set.
e2e:    HostExtractor.extract_flat from one flat pinned host tensor + lengths (H2D + kernels + D2H + float64 result), wall clock.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, xvec_b200
from oracle import xvector_oracle as ox

sd = ox.make_state_dict(0)
CONFIGS = {
    "c2 1024 x 3 s, batch 256": (np.full(1024, 300), 256 * 300, 256),
    "c3 4096 ragged 1-20 s": (ox.synth_lengths(4096, 100, 2000, seed=2), 1 << 17, 1 << 30),
    "c4 256 x 60 s": (np.full(256, 6000), 64 * 6000, 64),
    "c5 4874 x 4-20 s (VoxCeleb1-test sized)": (ox.synth_lengths(4874, 400, 2000, seed=3), 1 << 17, 1 << 30),
}
out = {}
for precision in ("bf16", "tf32"):
    m = xvec_b200.XVectorModel(precision=precision); m.load_state_dict(sd); m = m.cuda().eval()
    for name, (lens, max_frames, max_utts) in CONFIGS.items():
        lens = np.asarray(lens, dtype=np.int64)
        g = torch.Generator().manual_seed(1)
        flat = torch.randn(int(lens.sum()), 24, generator=g)
        utts = list(torch.split(flat, [int(v) for v in lens]))
        batches = xvec_b200.bucket_batches(lens, max_frames, max_utts)
        dev_batches = [(torch.cat([utts[i] for i in b]).cuda(), lens[b]) for b in batches]
        streams = [torch.cuda.Stream() for _ in range(2)]
        def run_dev():
            for k, (xb, lb) in enumerate(dev_batches):
                with torch.cuda.stream(streams[k % 2]):
                    m.extract_x_vec_flat(xb, lb, slot=k % 2)
        run_dev(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for s in streams: s.wait_event(e0)
        for _ in range(reps): run_dev()
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        e1.record(); torch.cuda.synchronize()
        dev_s = e0.elapsed_time(e1) / 1e3 / reps
        hx = xvec_b200.HostExtractor(m)
        flat_pinned = flat.pin_memory()
        for _ in range(2):  # warm-up: every slot has seen every batch shape (pinned staging, layouts, scratch)
            hx.extract_flat(flat_pinned, lens, max_frames, max_utts)
        t0 = time.perf_counter()
        hx.extract_flat(flat_pinned, lens, max_frames, max_utts)
        e2e_s = time.perf_counter() - t0
        fl = sum(ox.flops_per_utt(int(t)) for t in lens)
        out[f"{name} [{precision}]"] = {"utts": len(lens), "frames": int(lens.sum()), "batches": len(batches),
                                        "device_utt_s": len(lens) / dev_s, "device_frames_s": lens.sum() / dev_s, "device_tflops": fl / dev_s / 1e12,
                                        "e2e_utt_s": len(lens) / e2e_s, "e2e_frames_s": lens.sum() / e2e_s}
        print(name, precision, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in out[f"{name} [{precision}]"].items()}, flush=True)
        del dev_batches, utts, flat
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "configs.json"), "w"), indent=1)
