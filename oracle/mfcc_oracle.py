"""CPU restatement (numpy float64) of the MFCC front end the reference calls before the extraction path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity status: **UNPINNED**.  The reference computes its features with the third-party package
``python-speech-features==0.6`` (requirements.txt:39; call site dataset.py:130:
``mfcc(sample, 16000, numcep=24, nfilt=26, nfft=512)``).  That package is neither vendored under /root/reference nor
installed, and the reference holds no MFCC golden vectors, so this file restates the package's published algorithm
(base.py: mfcc / fbank / get_filterbanks / lifter, sigproc.py: preemphasis / framesig / powspec) from its documentation:

    preemphasis 0.97  ->  25 ms frames every 10 ms (400 / 160 samples at 16 kHz, zero padded at the end, rectangular
    window)  ->  |rfft_512|^2 / 512  ->  26 triangular mel filters on integer FFT bins, 0..8000 Hz  ->  log  ->
    DCT-II (ortho), first 24  ->  sinusoidal lifter L=22  ->  coefficient 0 replaced by log(frame energy)

It is checked only against internal identities (scipy's DCT, numpy's rfft, hand-computed filter edges).
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
FRAME_LEN = 400      # 0.025 s
FRAME_STEP = 160     # 0.010 s
NFFT = 512
NFILT = 26
NUMCEP = 24
PREEMPH = 0.97
CEPLIFTER = 22
EPS = np.finfo(float).eps


def num_frames(n_samples: int) -> int:
    if n_samples <= FRAME_LEN:
        return 1
    return 1 + int(np.ceil((n_samples - FRAME_LEN) / FRAME_STEP))


def hz2mel(hz):
    return 2595.0 * np.log10(1.0 + np.asarray(hz, dtype=np.float64) / 700.0)


def mel2hz(mel):
    return 700.0 * (10.0 ** (np.asarray(mel, dtype=np.float64) / 2595.0) - 1.0)


def filterbank_bins() -> np.ndarray:
    """The nfilt+2 integer FFT-bin edges of the triangular filters."""
    mel = np.linspace(hz2mel(0.0), hz2mel(SAMPLE_RATE / 2.0), NFILT + 2)
    return np.floor((NFFT + 1) * mel2hz(mel) / SAMPLE_RATE).astype(np.int64)


def filterbank() -> np.ndarray:
    """(nfilt, nfft/2+1) triangular filters: rising over [bin_j, bin_j+1), falling over [bin_j+1, bin_j+2)."""
    b = filterbank_bins()
    fb = np.zeros((NFILT, NFFT // 2 + 1))
    for j in range(NFILT):
        for i in range(b[j], b[j + 1]):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(b[j + 1], b[j + 2]):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def dct_matrix() -> np.ndarray:
    """(numcep, nfilt) orthonormal DCT-II rows."""
    n = np.arange(NFILT)
    k = np.arange(NUMCEP)[:, None]
    m = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2.0 * NFILT))
    m[0] *= np.sqrt(1.0 / (4.0 * NFILT))
    m[1:] *= np.sqrt(1.0 / (2.0 * NFILT))
    return m


def lifter_weights() -> np.ndarray:
    n = np.arange(NUMCEP)
    return 1.0 + (CEPLIFTER / 2.0) * np.sin(np.pi * n / CEPLIFTER)


def mfcc_np(signal: np.ndarray) -> np.ndarray:
    """(n_samples,) -> (num_frames, 24) float64."""
    x = np.asarray(signal, dtype=np.float64).reshape(-1)
    y = np.append(x[0], x[1:] - PREEMPH * x[:-1])
    nf = num_frames(len(y))
    padded = np.zeros((nf - 1) * FRAME_STEP + FRAME_LEN)
    padded[: len(y)] = y
    idx = np.arange(FRAME_LEN)[None, :] + FRAME_STEP * np.arange(nf)[:, None]
    frames = padded[idx]                                   # rectangular window
    pspec = np.abs(np.fft.rfft(frames, NFFT)) ** 2 / NFFT
    energy = pspec.sum(1)
    energy = np.where(energy == 0, EPS, energy)
    feat = pspec @ filterbank().T
    feat = np.where(feat == 0, EPS, feat)
    ceps = np.log(feat) @ dct_matrix().T
    ceps = ceps * lifter_weights()[None, :]
    ceps[:, 0] = np.log(energy)
    return ceps
