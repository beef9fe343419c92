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
    # stands in for HostExtractor.extract_all: any per-utterance function of the frames
    return np.stack([np.concatenate([u.double().mean(0).numpy(), [float(u.shape[0])]]) for u in utts])


def _worker(rank, world, port, lens, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    utts = list(torch.split(torch.randn(int(sum(lens)), 24, generator=g), [int(v) for v in lens]))
    out = sharding.extract_sharded(utts, _fake_extract)
    shard = sharding.my_shard(lens, rank, world)
    q.put((rank, None if out is None else out, shard))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_extraction_world2_gloo():
    lens = np.random.default_rng(3).integers(20, 400, 57)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out, shard = q.get(timeout=120)
        res[r] = (out, shard)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out0, shard0 = res[0]
    out1, shard1 = res[1]
    assert out1 is None and out0.shape == (57, 25)
    assert np.array_equal(np.sort(np.concatenate([shard0, shard1])), np.arange(57))    # every utterance exactly once
    assert abs(int((lens[shard0] - 14).sum()) - int((lens[shard1] - 14).sum())) <= lens.max()  # balanced
    g = torch.Generator().manual_seed(5)
    utts = list(torch.split(torch.randn(int(lens.sum()), 24, generator=g), [int(v) for v in lens]))
    assert np.array_equal(out0, _fake_extract(utts))                                   # original order restored
    assert np.array_equal(sharding.extract_sharded(utts, _fake_extract), out0)         # world == 1 path


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
