"""CPU tests of the rows either side of the kernel path: the reference's x-vector .csv format (f1), the trial-file
parser, and the utterance-sharded multi-process path (world_size 2, gloo) with a stub per-rank extractor."""
import io
import os
import socket

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import xvec_b200
from xvec_b200 import io_csv, sharding


def test_csv_matches_reference_writer_and_reader(tmp_path):
    rng = np.random.default_rng(0)
    xv = rng.standard_normal((7, 512)) * np.exp(rng.standard_normal((7, 1)) * 3)
    ids = [f"id1027{k}/5r0dWxy17C8/0000{k}.wav" for k in range(7)]
    labels = list(range(10270, 10277))
    path = str(tmp_path / "x_vector_test.csv")
    xvec_b200.write_xvector_csv(path, ids, labels, xv)
    # byte-identical to what the reference's own writer produces (main.py:246-247)
    ref_records = [(i, int(l), np.array(x, dtype=np.float64)) for i, l, x in zip(ids, labels, xv)]
    buf = io.StringIO()
    pd.DataFrame(ref_records).to_csv(buf)
    assert open(path).read() == buf.getvalue()
    assert open(path).readline().strip() == ",0,1,2"
    # the reference's reader (plda_score_stat.py:16-17) recovers the vectors to print precision
    df = pd.read_csv(path)
    got = np.array([np.array(c[1:-1].split(), dtype=np.float64) for c in df.iloc[:, 3]])
    assert np.allclose(got, xv, rtol=1e-8, atol=0)
    rid, rlab, rx = xvec_b200.read_xvector_csv(path)
    assert list(rid) == ids and rlab.tolist() == labels and np.array_equal(rx, got)
    assert int(rid[0].split("/")[0][2:]) == 10270        # numeric label parse of plda_score_stat.py:73


def test_trial_file_parser():
    tgt, en, te = io_csv.parse_trial_file(["1 id10270/a/00001.wav id10270/b/00002.wav\n", "0 id10270/a/00001.wav id10300/c/00001.wav\n", "\n"])
    assert tgt.tolist() == [True, False] and en[1] == "id10270/a/00001.wav" and te[1] == "id10300/c/00001.wav"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_extract(utts):
    # stands in for HostExtractor.extract_all: any per-utterance function of the frames (float32-exact values, like x-vectors)
    return np.stack([np.concatenate([u.double().mean(0).numpy(), [float(u.shape[0])]]) for u in utts]).astype(np.float32).astype(np.float64)


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    utts = list(torch.split(torch.randn(int(sum(lens)), 24, generator=g), [int(v) for v in lens]))
    out = sharding.extract_sharded(utts, _fake_extract, dim=25)
    shard = sharding.my_shard(lens, rank, world)
    # the tensor gather itself: every rank gets every row, in the original order, bit for bit
    parts = xvec_b200.lpt_partition(lens, world)
    local = torch.from_numpy(_fake_extract([utts[i] for i in parts[rank]])).float() if len(parts[rank]) else torch.zeros((0, 25))
    full = sharding.gather_rows(local, parts, len(utts))
    # batch-granular sharding through a reusable RowGather (two calls: the buffers are reused)
    bparts, bsizes = sharding.shard_batches(lens, world, target_frames=2000)
    rg = sharding.RowGather(bparts, len(utts), 25, "cpu")
    for _ in range(2):
        if len(bparts[rank]):
            rg.local_view.copy_(torch.from_numpy(_fake_extract([utts[i] for i in bparts[rank]])).float())
        full_b = rg().clone()
    assert torch.equal(full_b, full) and sum(bsizes[rank]) == len(bparts[rank])
    q.put((rank, None if out is None else out, shard, full.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_utts", [57, 1])
def test_sharded_extraction_world2_gloo(n_utts):
    """extract_sharded / gather_rows over a real 2-process group (gloo, host tensors): fixed-shape all_gather_into_tensor of
    padded float32 blocks, no pickled objects; n_utts = 1 leaves rank 1 with an empty shard."""
    lens = np.random.default_rng(3).integers(20, 400, n_utts)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out, shard, full = q.get(timeout=120)
        res[r] = (out, shard, full)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out0, shard0, full0 = res[0]
    out1, shard1, full1 = res[1]
    assert out1 is None and out0.shape == (n_utts, 25) and out0.dtype == np.float64
    assert np.array_equal(full0, full1) and np.array_equal(full0.astype(np.float64), out0)     # every rank holds the whole matrix
    assert np.array_equal(np.sort(np.concatenate([shard0, shard1])), np.arange(n_utts))    # every utterance exactly once
    assert abs(int((lens[shard0] - 14).sum()) - int((lens[shard1] - 14).sum())) <= lens.max()  # balanced
    g = torch.Generator().manual_seed(5)
    utts = list(torch.split(torch.randn(int(lens.sum()), 24, generator=g), [int(v) for v in lens]))
    assert np.array_equal(out0, _fake_extract(utts))                                   # original order restored
    assert np.array_equal(sharding.extract_sharded(utts, _fake_extract), out0)         # world == 1 path


def test_balanced_batches_do_not_depend_on_world_size():
    """shard_batches: contiguous batches of equal frames whose composition is a function of the lengths alone; worker r of any
    world takes batches r, r + world, ... — the property that makes sharded x-vectors bit-identical for 1, 2, 4, 8 GPUs."""
    lens = np.random.default_rng(3).integers(400, 2001, 4874)
    batches = xvec_b200.balanced_batches(lens, target_frames=49152, multiple_of=8)
    assert len(batches) % 8 == 0 and np.array_equal(np.concatenate(batches), np.arange(4874))
    frames = np.asarray([lens[b].sum() for b in batches])
    assert frames.max() - frames.min() <= 2 * lens.max() and abs(frames.mean() - 49152) < 0.05 * 49152
    for world in (1, 2, 4, 8):
        parts, sizes = sharding.shard_batches(lens, world)
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(4874))
        loads = np.asarray([lens[p].sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01
        for r in range(world):  # rank r's batches are exactly batches r, r + world, ... of the world-independent list
            assert sizes[r] == [len(b) for b in batches[r::world]]
            assert np.array_equal(parts[r], np.concatenate(batches[r::world]))
    for n in (1, 2, 5, 9, 40):  # tiny sets: never an empty batch, every utterance once
        small = np.random.default_rng(n).integers(20, 400, n)
        bb = xvec_b200.balanced_batches(small, target_frames=3000)
        assert all(len(b) for b in bb) and np.array_equal(np.concatenate(bb), np.arange(n))
        parts, sizes = sharding.shard_batches(small, 4, target_frames=3000)
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(n))


def test_eer_and_min_dcf_from_first_principles():
    from xvec_b200 import scoring
    rng = np.random.default_rng(0)
    tar = rng.normal(1.0, 1.0, 4000)
    non = rng.normal(-1.0, 1.0, 6000)
    s = np.concatenate([tar, non])
    t = np.concatenate([np.ones(4000, bool), np.zeros(6000, bool)])
    e, thr = scoring.eer(s, t)
    far = (non >= thr).mean()
    frr = (tar < thr).mean()
    assert abs(far - frr) < 2e-3 and abs(e - 0.5 * (far + frr)) < 1e-12
    assert abs(e - 0.1587) < 0.01                          # Phi(-1) for unit-variance classes 2 sigma apart
    d, thr_d = scoring.min_dcf(s, t, p_target=0.5)
    assert abs(d - 0.5 * ((tar < thr_d).mean() + (non >= thr_d).mean())) < 1e-12
    assert d <= 0.5 * (far + frr) + 1e-12 and abs(d - 0.1587) < 0.01
    # brute force over every candidate threshold
    cands = np.sort(s)
    brute = min(0.5 * ((tar < c).mean() + (non >= c).mean()) for c in cands[::50])
    assert d <= brute + 1e-12
    # the oracle's EER threshold agrees on decisions up to ties
    from oracle import xvector_oracle as ox
    e2, thr2, _ = ox.eer_threshold_np(s, t)
    assert abs(e - e2) < 2e-3
    with pytest.raises(ValueError):
        scoring.eer(np.zeros(4), np.ones(4, bool))
