#!/usr/bin/env python
"""Benchmark of the x-vector extraction hot path (BASELINE.json metric: x-vectors/sec & frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype bf16|tf32] [--impl b200|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (one rank per GPU, weak scaling)

A step = one pass of the hot path (TDNN1-5 -> statistics pooling -> segment6) over ONE batch of the workload
(BASELINE.json configs[1]: 1024 fixed-length 3 s utterances = 4 batches of 256 x 300 x 24 MFCC, cycled).
  value      utterances/s, device-resident inputs (more distinct batches than fit in L2), K steps in one CUDA-event bracket,
             two batches in flight on two streams
  e2e        utterances/s through the public host API (HostExtractor): pinned host MFCCs -> H2D -> kernels -> D2H x-vectors
  roofline   tcgen05 TDNN stack kernel: algorithmic FLOPs of a launch / its CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (port of the reference's fp32 PyTorch path) on this box's host cores, bounded sample
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, FRAMES, CEPS, N_UTTS = int(os.environ.get("XVEC_BENCH_BATCH", "256")), 300, 24, 1024  # BASELINE.json configs[1]
N_UTTS = (N_UTTS // BATCH) * BATCH
FLOPS_L = [122_880, 1_572_864, 1_572_864, 524_288, 1_536_000]  # per output frame, SURVEY §8d
LOST = [4, 8, 14, 14, 14]
SEG6_FLOPS = 3_072_000
WORKLOAD = "c2: 1024 x 3 s utterances (300 x 24 MFCC), batch 256 per step and GPU, x_vec_extract_layer 6"


def flops_per_utt(t):
    return sum(f * (t - l) for f, l in zip(FLOPS_L, LOST)) + SEG6_FLOPS


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port of its fp32 PyTorch ops), all host threads."""
    if rank != 0:
        return
    import torch
    from oracle import xvector_oracle as ox
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = ox.make_state_dict(seed=0)
    sample = 64  # utterances of the 256-utterance batch per step (bounded CPU sample)
    x = ox.synth_mfcc(sample, FRAMES, seed=1234)
    for _ in range(args.warmup):
        ox.extract_x_vec_aten(sd, x, 6)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ox.extract_x_vec_aten(sd, x, 6)
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {"impl": "reference", "metric": "x-vectors/sec", "value": v, "unit": "utt/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "frames_per_sec": v * FRAMES,
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "frames": FRAMES,
                       "parallelism": "reference CPU implementation (oracle port), rank 0 only, all host threads"},
            "cpu_baseline": {"value": v, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{sample} of the {BATCH} utterances of a batch per step, oracle/xvector_oracle.extract_x_vec_aten (the reference's ATen op sequence, fp32 torch CPU)"},
            "e2e": {"value": v, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def cpu_baseline_leg(budget_s=12.0):
    import torch
    from oracle import xvector_oracle as ox
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = ox.make_state_dict(seed=0)
    x = ox.synth_mfcc(64, FRAMES, seed=1234)  # BASELINE.json configs[0]
    for _ in range(2):
        ox.extract_x_vec_aten(sd, x, 6)
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < 5 or (time.perf_counter() < t_end and len(times) < 50):
        t0 = time.perf_counter()
        ox.extract_x_vec_aten(sd, x, 6)
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": 64 / med, "unit": "utt/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"config c1 (64 x 300 x 24, batch 64), median of {len(times)} runs of oracle.extract_x_vec_aten (the reference's ATen op sequence, fp32 torch CPU)",
            "frames_per_sec": 64 * FRAMES / med, "gflops": 64 * flops_per_utt(FRAMES) / med / 1e9, "best_utt_s": 64 / times[0]}


def synthetic_model(xvec_b200, precision):
    """Random-init weights of the reference architecture (PyTorch default initialisers, as main.XVectorModel() gets them) with
    non-trivial eval-mode BatchNorm statistics (default-init BN is the identity and would make the BN fold free)."""
    import torch
    torch.manual_seed(0)
    model = xvec_b200.XVectorModel(precision=precision)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for layer in model.time_context_layers:
            n = layer.norm.num_features
            layer.norm.running_mean.copy_(0.5 * torch.randn(n, generator=g))
            layer.norm.running_var.copy_(0.3 + 1.7 * torch.rand(n, generator=g))
            layer.norm.weight.copy_(1.0 + 0.5 * torch.randn(n, generator=g))
            layer.norm.bias.copy_(0.5 * torch.randn(n, generator=g))
    return model


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import xvec_b200  # the product; nothing under oracle/ is imported on this arm except by cpu_baseline_leg()

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.dtype
    model = synthetic_model(xvec_b200, precision).to(dev).eval()

    n_batches = N_UTTS // BATCH
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(N_UTTS, FRAMES, CEPS, generator=g).reshape(n_batches, BATCH * FRAMES, CEPS).pin_memory()
    # device-resident inputs: the 1024-utterance set replicated (with a per-copy scale) to N_RESIDENT distinct batches so that
    # the inputs cycled through the timed region (N_RESIDENT x 7.4 MB) are larger than the 126 MB L2
    n_res = max(n_batches, -(-(160 << 20) // (BATCH * FRAMES * CEPS * 4)))
    x_dev = torch.empty((n_res, BATCH * FRAMES, CEPS), dtype=torch.float32, device=dev)
    for i in range(n_res):
        x_dev[i].copy_(x_host[i % n_batches])
        x_dev[i].mul_(1.0 + 0.01 * (i // n_batches))
    lengths = [FRAMES] * BATCH
    n_inflight = int(os.environ.get("XVEC_BENCH_INFLIGHT", "2"))  # batches in flight on separate streams: the tail of one batch overlaps the next batch's kernels
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_inflight)]

    def step(i):
        with torch.cuda.stream(streams[i % n_inflight]):
            return model.extract_x_vec_flat(x_dev[i % n_res], lengths, slot=i % n_inflight)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K steps, 2 in flight, one CUDA-event bracket on the device
    for i in range(max(args.warmup, 3) * n_inflight):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    t_wall = time.perf_counter()
    e_beg.record(cur)
    for st in streams:
        st.wait_event(e_beg)
    for i in range(args.steps):
        step(i)
    for st in streams:
        cur.wait_stream(st)
    e_end.record(cur)
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = e_beg.elapsed_time(e_end)

    # ---- per-kernel timing of the dominant kernel (separate instrumented pass right after the throughput leg, i.e. in the same
    #      thermal / power state; not part of the numbers above)
    stack_ms = instrumented_stack_time(model, x_dev, lengths, n_res, iters=max(10, min(args.steps, 50)))
    layer_ms = instrumented_layer_times(model, x_dev, lengths, n_res, iters=max(5, min(args.steps, 20)))

    # ---- end to end through the host API (pinned host -> H2D -> kernels -> D2H), six pipeline slots
    N_SLOTS = 6
    hx = xvec_b200.HostExtractor(model, n_slots=N_SLOTS)
    for i in range(2 * N_SLOTS):
        hx.result(hx.submit(x_host[i % n_batches], lengths))
    barrier()
    hx.h2d_bytes = hx.d2h_bytes = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e = time.perf_counter()
    e0.record()
    tickets = []
    checksum = 0.0
    for i in range(args.steps):
        tickets.append(hx.submit(x_host[i % n_batches], lengths))
        if len(tickets) == N_SLOTS:
            checksum += float(hx.result(tickets.pop(0))[0, 0])
    while tickets:
        checksum += float(hx.result(tickets.pop(0))[0, 0])
    e1.record()
    barrier()
    t_e2e = time.perf_counter() - t_e2e
    clocks = sampler.stop()


    dev_ms_t = torch.tensor([dev_ms, t_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dev_ms_t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = dev_ms_t.tolist()

    if rank == 0:
        peaks = load_peaks()
        utts = BATCH * args.steps * world
        value = utts / (dev_ms / 1e3)
        e2e = utts / (e2e_ms / 1e3)
        tdnn_flops = BATCH * sum(f * (FRAMES - l) for f, l in zip(FLOPS_L, LOST))
        achieved = tdnn_flops / (stack_ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"] if precision == "bf16" else peaks["bf16_tflops_sustained"] / 2
        per_layer = {k: {"ms": round(v, 5)} for k, v in layer_ms.items()}
        for i in range(5):
            fl = BATCH * FLOPS_L[i] * (FRAMES - LOST[i])
            per_layer[f"tdnn{i + 1}"]["tflops"] = round(fl / (layer_ms[f"tdnn{i + 1}"] / 1e3) / 1e12, 2)
        line = {
            "metric": "x-vectors/sec", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "tf32", "data": "synthetic",
            "frames_per_sec": value * FRAMES,
            "config": {"workload": WORKLOAD,
                       "global_batch": BATCH * world, "frames": FRAMES, "parallelism": f"utterance-sharded x{world}, no data-path collective",
                       "l2": f"inputs larger than L2: {n_res} distinct device-resident batches ({n_res * BATCH * FRAMES * CEPS * 4 >> 20} MiB) cycled; "
                             "2 batches in flight; the stack kernel's banded schedule deliberately keeps one band's activations (2 x 43.5 MB in bf16) L2-resident between layers",
                       "tdnn1": "TF32 math on the float32 MFCCs in both modes (window form: one K = 120 GEMM over overlapping rows)"},
            "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": hx.h2d_bytes // args.steps, "d2h_bytes_per_step": hx.d2h_bytes // args.steps,
                    "api": "HostExtractor.submit/result (pinned host MFCCs in, pinned host x-vectors out, 6 slots / streams)",
                    "frames_per_sec": e2e * FRAMES, "checksum": checksum},
            # per step: tdnn_stack_kernel, pool_finalize_kernel, fc_small_kernel (segment6)
            "gpu_launches": args.steps * 3,
            "clocks": clocks,
            "wall_ms_per_step": t_wall / args.steps * 1e3,
            "roofline": {"kernel": "tdnn_stack_kernel (1 launch/step: all tiles of TDNN1-5 from one work queue; TDNN5 epilogue = pooling partials)",
                         "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": STACK_DRAM_BYTES.get(precision),
                         "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops_sustained" + ("" if precision == "bf16" else " / 2 (TF32)"),
                         "frac_of_burst_peak": achieved / (peaks["bf16_tflops"] if precision == "bf16" else peaks["bf16_tflops"] / 2),
                         "algorithmic_flops_per_launch": tdnn_flops, "ms_per_launch": stack_ms,
                         "timing": "CUDA events around each launch on its stream, launches back to back on one stream, averaged"},
            # for comparison only: the same layers as one tdnn_gemm_kernel launch each (the XVEC_STACK=0 / XVEC_FC_SMALL=0 path;
            # segment6 here is the tcgen05 split-K GEMM + reduce, the product runs it on fc_small_kernel)
            "per_layer_launches": per_layer,
        }
        if world == 1:
            line["roofline_pool"] = pooling_roofline(model, dev, peaks)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE tdnn_stack_kernel launch on this workload
# (ncu --set full, profiles/r01_v13_ncu_full_summary.txt: 17.6 MB read + 186.9 MB written — with L2-sized bands only the final
# contents of the two 78.6 MB activation buffers and the pooling partials go back to HBM)
STACK_DRAM_BYTES = {"bf16": 204429312}


def instrumented_stack_time(model, x_dev, lengths, n_batches, iters):
    """Average CUDA-event duration of the tdnn_stack_kernel launch (incl. the 5 KB control-block memset it is enqueued with)."""
    import torch
    from xvec_b200 import ops
    from xvec_b200.tdnn_layer import _aligned_rows
    lay = model._layout_for(lengths)
    sc = model._scratch_for(0)
    sc.ensure(lay.rows, lay.n_slots, lay.n_utts)
    pipe = model._pipeline()
    part = sc.part[: lay.n_slots]
    evs = []
    for it in range(iters + 3):
        x = _aligned_rows(x_dev[it % n_batches])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.tdnn_stack(pipe["tdnn"], pipe["n_tdnn"], x, sc.act[0], sc.act[1], lay.row_utt, lay.blk_slot_base, part, sc.ctrl)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs[3:]) / iters


def instrumented_layer_times(model, x_dev, lengths, n_batches, iters):
    """CUDA-event time of every launch of one step, averaged over `iters` steps (events on the launching stream)."""
    import torch
    import xvec_b200
    from xvec_b200 import ops
    from xvec_b200.tdnn_layer import tap_offsets, _aligned_rows
    names = ["tdnn1", "tdnn2", "tdnn3", "tdnn4", "tdnn5", "pool_finalize", "segment6"]
    acc = dict.fromkeys(names, 0.0)
    layers = list(model.time_context_layers)
    lay = model._layout_for(lengths)
    sc = model._scratch_for(0)
    sc.ensure(lay.rows, lay.n_slots, lay.n_utts)
    part, pooled = sc.part[: lay.n_slots], sc.pooled[: lay.n_utts]
    pooled_lp = None if sc.pooled_lp is None else sc.pooled_lp[: lay.n_utts]
    stack, (scale5, shift5) = model._stack_params()
    all_evs = []
    for it in range(iters + 2):  # no host sync inside: the launches queue up and run back to back on the device
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        h = _aligned_rows(x_dev[it % n_batches])
        evs[0].record()
        for i, layer in enumerate(layers[:-1]):
            w, bias, offs = stack[i]
            h = ops.tdnn_layer_flat(h, w, layer.output_size, offs, bias, None, None, relu=True,
                                    out=sc.act[i & 1][: lay.rows, : layer.output_size], cin=layer.input_size)
            evs[i + 1].record()
        last = layers[-1]
        w, bias, offs = stack[-1]
        ops.tdnn_pool_fused(h, w, last.output_size, offs, bias, lay.row_utt, lay.blk_slot_base, part)
        evs[5].record()
        ops.pool_finalize(part, lay.utt_slot_start, lay.n_pool, last.output_size, scale5, shift5, out=pooled, out_lp=pooled_lp)
        evs[6].record()
        model._head(pooled, pooled_lp, 6)
        evs[7].record()
        all_evs.append(evs)
    torch.cuda.synchronize()
    for evs in all_evs[2:]:
        for k, name in enumerate(names):
            acc[name] += evs[k].elapsed_time(evs[k + 1])
    return {k: v / iters for k, v in acc.items()}


def pooling_roofline(model, dev, peaks):
    """Standalone statistics pooling (XVectorModel.stat_pool) on a long-form activation larger than L2: HBM roofline."""
    import torch
    b, t, p = 64, 5986, 1500                      # 64 of the 256 long-form (60 s) utterances of config c4: 2.3 GB fp32
    a = torch.randn(b, t, p, device=dev)
    for _ in range(3):
        model.stat_pool(a)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for e0, e1 in evs:
        e0.record()
        model.stat_pool(a)
        e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in evs)[len(evs) // 2]
    nbytes = b * t * p * 4 + b * 2 * p * 4
    gbs = nbytes / (ms / 1e3) / 1e9
    del a
    return {"kernel": "stats_pool_partial_kernel + pool_finalize_kernel (standalone stat_pool, 64 x 5986 x 1500 fp32)", "bound": "hbm",
            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": 2441932472,  # dram__bytes_read+write of stats_pool_partial_kernel per launch, profiles/r01_v5_ncu_pool_summary.txt
            "algorithmic_bytes": nbytes, "ms": ms, "peak_source": f"{peaks['source']} MEASURED_PEAKS.json hbm_gbs"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)  # ~0.3 s per leg: long enough for the 1000 W power cap to engage
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
