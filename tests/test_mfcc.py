"""MFCC front end.  Parity is UNPINNED against python_speech_features (not available offline): the CPU
tests check oracle/mfcc_oracle.py against independent identities, the GPU test checks the kernel against that oracle."""
import numpy as np
import pytest
import scipy.fftpack
import torch

from oracle import mfcc_oracle as mo


def test_oracle_building_blocks():
    assert mo.num_frames(48000) == 299 and mo.num_frames(400) == 1 and mo.num_frames(401) == 2 and mo.num_frames(100) == 1
    # orthonormal DCT-II matrix == scipy's (the routine the package calls)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5, mo.NFILT))
    assert np.allclose(x @ mo.dct_matrix().T, scipy.fftpack.dct(x, type=2, axis=1, norm="ortho")[:, : mo.NUMCEP])
    # filterbank: 28 non-decreasing integer edges from 0 to the Nyquist bin, unit peak, partition-like triangles
    b = mo.filterbank_bins()
    assert b[0] == 0 and b[-1] == 256 and (np.diff(b) >= 0).all() and len(b) == 28
    fb = mo.filterbank()
    assert fb.shape == (26, 257) and fb.min() >= 0 and np.isclose(fb.max(), 1.0)
    for j in range(26):
        if b[j + 1] < b[j + 2]:
            assert fb[j, b[j + 1]] == 1.0
    assert np.isclose(mo.lifter_weights()[0], 1.0) and np.isclose(mo.lifter_weights()[11], 12.0)


def test_oracle_mfcc_shapes_and_energy():
    rng = np.random.default_rng(1)
    sig = rng.random(48000)
    c = mo.mfcc_np(sig)
    assert c.shape == (299, 24) and np.isfinite(c).all()
    # coefficient 0 is the log of the frame's total power (Parseval on the zero-padded 512-point frame)
    y = np.append(sig[0], sig[1:] - 0.97 * sig[:-1])
    fr = y[160 * 7: 160 * 7 + 400]
    full = np.abs(np.fft.fft(fr, 512)) ** 2 / 512
    assert np.isclose(c[7, 0], np.log(full[:257].sum()))
    # silence -> the eps substitution, no -inf
    z = mo.mfcc_np(np.zeros(1000))
    assert np.isfinite(z).all() and np.isclose(z[0, 0], np.log(np.finfo(float).eps))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "int16"])
def test_mfcc_kernel_matches_oracle(dtype):
    import xvec_b200
    rng = np.random.default_rng(2)
    lens = [48000, 400, 401, 16000, 50000, 123, 31999]
    sigs = []
    for n in lens:
        t = np.arange(n) / 16000.0
        s = 0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 3100 * t + 1.0) + 0.05 * rng.standard_normal(n)
        sigs.append(np.round(s * 12000).astype(np.int16) if dtype == "int16" else s.astype(np.float32))
    wav = torch.from_numpy(np.concatenate(sigs)).cuda()
    out, nf = xvec_b200.ops.mfcc(wav, lens, normalize=True)
    assert nf.tolist() == [mo.num_frames(n) for n in lens]
    got = out.cpu().numpy().astype(np.float64)
    row = 0
    for s, n in zip(sigs, nf):
        x = s.astype(np.float64)
        x = (x - x.min()) / (x.max() - x.min())        # dataset.py:217-218 (x -= min; x /= max)
        ref = mo.mfcc_np(x)
        g = got[row: row + n]
        assert ref.shape == g.shape
        assert np.abs(g - ref).max() < 2e-3 * max(1.0, np.abs(ref).max()), (len(s), np.abs(g - ref).max())
        row += n
    # un-normalised float path
    if dtype == "float32":
        out2, _ = xvec_b200.ops.mfcc(wav, lens, normalize=False)
        ref0 = mo.mfcc_np(sigs[0].astype(np.float64))
        assert np.abs(out2[: nf[0]].cpu().numpy() - ref0).max() < 2e-3 * np.abs(ref0).max()


@pytest.mark.gpu
def test_wav_to_xvector_pipeline(state_dict):
    """Waveforms -> GPU MFCC -> x-vectors without leaving the device == oracle MFCC -> oracle extraction."""
    import xvec_b200
    from oracle import xvector_oracle as ox
    rng = np.random.default_rng(3)
    lens = [16000, 24000, 9000]
    sigs = [rng.standard_normal(n).astype(np.float32) for n in lens]
    feats, nf = xvec_b200.ops.mfcc(torch.from_numpy(np.concatenate(sigs)).cuda(), lens)
    m = xvec_b200.XVectorModel(precision="tf32")
    m.load_state_dict(state_dict)
    m = m.cuda().eval()
    got = m.extract_x_vec_flat(feats, nf).cpu().numpy()
    ref = []
    for s in sigs:
        x = s.astype(np.float64)
        x = (x - x.min()) / (x.max() - x.min())
        ref.append(ox.extract_x_vec_t(state_dict, torch.from_numpy(mo.mfcc_np(x)).float()[None], 6)[0].numpy())
    ref = np.stack(ref)
    rel = np.abs(got - ref).max(1) / np.linalg.norm(ref, axis=1)
    assert rel.max() < 2e-3, rel.max()
