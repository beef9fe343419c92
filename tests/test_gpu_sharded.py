"""The utterance-sharded path on real GPUs over NCCL (BASELINE.json config 5 at test size): one process per GPU,
world_size = min(2, device_count), LPT shards, HostExtractor per rank, ONE all_gather_into_tensor of device-resident float32
blocks, centred-cosine trials on the gathered matrix.  Checked against the unsharded extraction on one GPU and the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_UTTS, N_TRIALS, N_SPK = 96, 600, 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    from oracle import xvector_oracle as ox
    lens = ox.synth_lengths(N_UTTS, 40, 700, seed=21)
    utts = ox.synth_speaker_utts(lens, N_SPK, seed=22)
    enrol, test, target = ox.synth_trials(N_UTTS, N_TRIALS, n_speakers=N_SPK, seed=23)
    return lens, utts, enrol, test, target


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import xvec_b200
    from xvec_b200 import sharding
    from oracle import xvector_oracle as ox
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lens, utts, enrol, test, target = _inputs()
        m = xvec_b200.XVectorModel(precision="tf32")
        m.load_state_dict(ox.make_state_dict(seed=0, randomize_bn=True))
        m = m.to(dev).eval()
        hx = xvec_b200.HostExtractor(m)
        # (1) the generic entry point: list of host utterances, per-rank extractor, tensor gather over NCCL
        out = sharding.extract_sharded(utts, lambda us: hx.extract_all(us, max_frames=6000))
        # (2) the GPU-native form: batch-granular shards as one flat pinned host tensor, x-vectors stay on the devices
        parts, sizes = sharding.shard_batches(lens, world, target_frames=4000)
        mine = parts[rank]
        flat = torch.cat([utts[i] for i in mine]).pin_memory()
        full_dev = sharding.extract_sharded_flat(hx, flat, lens[mine], parts, len(utts), batch_sizes=sizes[rank])
        assert full_dev.is_cuda and full_dev.shape == (N_UTTS, 512) and full_dev.dtype == torch.float32
        scores = xvec_b200.ops.cosine_trials(full_dev, torch.from_numpy(enrol).int().to(dev), torch.from_numpy(test).int().to(dev), center=True)
        q.put((rank, None if out is None else out, full_dev.cpu().numpy(), scores.cpu().numpy(), mine))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_extraction_nccl_matches_single_gpu():
    import xvec_b200
    from oracle import xvector_oracle as ox
    world = min(2, torch.cuda.device_count())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out, full, scores, mine = q.get(timeout=600)
        res[r] = (out, full, scores, mine)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    out0, full0, scores0, _ = res[0]
    assert out0 is not None and out0.dtype == np.float64 and out0.shape == (N_UTTS, 512)
    assert np.array_equal(np.sort(np.concatenate([res[r][3] for r in range(world)])), np.arange(N_UTTS))  # every utterance once
    for r in range(1, world):
        assert res[r][0] is None
        assert np.array_equal(res[r][1], full0) and np.array_equal(res[r][2], scores0)  # every rank holds the same matrix / scores
    # unsharded on one GPU (other batch boundaries: fp32 summation order of the pooling only) and the oracle
    lens, utts, enrol, test, target = _inputs()
    sd = ox.make_state_dict(seed=0, randomize_bn=True)
    m = xvec_b200.XVectorModel(precision="tf32")
    m.load_state_dict(sd)
    m = m.cuda().eval()
    single = xvec_b200.HostExtractor(m).extract_all(utts, max_frames=1 << 17)
    scale = np.abs(single).max()
    assert np.abs(out0 - single).max() < 1e-4 * scale and np.abs(full0 - single).max() < 1e-4 * scale
    sel = np.arange(0, N_UTTS, 7)
    ref = ox.extract_ragged_t(sd, [utts[i] for i in sel], 6).numpy().astype(np.float64)
    rel = np.abs(full0[sel] - ref).max(1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 1e-3, rel.max()
    # batch-granular sharding is bit-identical to the same batches run on ONE GPU (world-size independence)
    parts1, sizes1 = xvec_b200.sharding.shard_batches(lens, 1, target_frames=4000)
    hx1 = xvec_b200.HostExtractor(m)
    one = hx1.extract_flat(torch.cat(utts).pin_memory(), lens, to_host=False, batch_sizes=sizes1[0]).cpu().numpy()
    assert np.array_equal(one, full0)
    # trial decisions: identical to float64 cosine scoring of the single-GPU embeddings at its EER threshold
    s64 = ox.cosine_scores_np(single, enrol, test, center=True)
    eer, thr, margin = ox.eer_threshold_np(s64, target)
    assert np.abs(scores0 - s64).max() < margin
    assert np.array_equal(scores0 >= thr, s64 >= thr)


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_batch_sharding_is_bit_identical_for_any_world_size(precision):
    """The workers of a 1-, 2-, 4- and 8-GPU job emulated one after the other on this GPU (separate HostExtractors): with
    shard_batches the assembled (N, 512) matrices are equal bit for bit — the checksum bench.py's c5 object prints can only be
    N-independent because of this — whereas the utterance-granular LPT shards agree to rounding only."""
    import xvec_b200
    from xvec_b200 import sharding
    from oracle import xvector_oracle as ox
    lens = ox.synth_lengths(300, 30, 900, seed=5)
    utts = ox.synth_ragged(lens, seed=6)
    m = xvec_b200.XVectorModel(precision=precision)
    m.load_state_dict(ox.make_state_dict(seed=0, randomize_bn=True))
    m = m.cuda().eval()
    outs = {}
    for world in (1, 2, 4, 8):
        parts, sizes = sharding.shard_batches(lens, world, target_frames=9000)
        full = torch.empty((len(lens), 512), device="cuda")
        for r in range(world):
            hx = xvec_b200.HostExtractor(m, n_slots=3)
            flat = torch.cat([utts[i] for i in parts[r]]).pin_memory()
            local = hx.extract_flat(flat, lens[parts[r]], to_host=False, batch_sizes=sizes[r])
            full[torch.from_numpy(parts[r]).cuda()] = local
        outs[world] = full.cpu()
    for world in (2, 4, 8):
        assert torch.equal(outs[world], outs[1]), world
    assert torch.isfinite(outs[1]).all()
