"""Trial scoring on top of the extracted x-vectors.

The reference scores with SpeechBrain's PLDA over the full N x N matrix and then looks trial pairs up one by one
(plda_score_stat.py:59-87), and takes EER / minDCF(p_target=0.5) from speechbrain.utils.metric_stats (:92-97).  SpeechBrain
is not vendored, so those semantics are unpinned; here the metrics are defined from first principles:
  * score = centred cosine of the trial's two x-vectors (GPU kernel xvec_cosine_trials, BASELINE.json config 5),
  * EER: operating point where false-acceptance and false-rejection rates cross (threshold between two adjacent scores),
  * minDCF: min over thresholds of  c_miss * p_target * FRR + c_fa * (1 - p_target) * FAR.
Decisions are `score >= threshold`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .io_csv import parse_trial_file


def cosine_score_trials(xvecs, enrol_idx, test_idx, center: bool = True) -> np.ndarray:
    """Scores for (enrol, test) index pairs into `xvecs` (N, D); runs on the GPU `xvecs` lives on (or cuda:0)."""
    x = xvecs if isinstance(xvecs, torch.Tensor) else torch.as_tensor(np.asarray(xvecs, dtype=np.float32))
    if not x.is_cuda:
        x = x.cuda()
    x = x.float().contiguous()
    e = torch.as_tensor(np.asarray(enrol_idx, dtype=np.int32)).to(x.device)
    t = torch.as_tensor(np.asarray(test_idx, dtype=np.int32)).to(x.device)
    if e.numel() and (int(e.max()) >= x.shape[0] or int(t.max()) >= x.shape[0] or int(e.min()) < 0 or int(t.min()) < 0):
        raise ValueError("trial index out of range")
    return ops.cosine_trials(x, e, t, center=center).cpu().numpy()


def score_trial_file(xvecs, ids, trial_lines, center: bool = True):
    """VoxCeleb-style trial list ('<0|1> <enrol_id> <test_id>') scored against x-vectors identified by `ids`.
    Returns (scores float32, is_target bool)."""
    target, enrol, test = parse_trial_file(trial_lines)
    pos = {str(i): k for k, i in enumerate(ids)}
    try:
        e = np.asarray([pos[i] for i in enrol], dtype=np.int32)
        t = np.asarray([pos[i] for i in test], dtype=np.int32)
    except KeyError as ex:
        raise ValueError(f"trial refers to an utterance without x-vector: {ex}") from None
    return cosine_score_trials(xvecs, e, t, center), target


def _rates(scores, target):
    s = np.asarray(scores, dtype=np.float64)
    t = np.asarray(target, dtype=bool)
    if s.shape != t.shape or s.ndim != 1 or t.sum() == 0 or (~t).sum() == 0:
        raise ValueError("need 1-D scores with at least one target and one non-target trial")
    order = np.argsort(s, kind="stable")
    s, t = s[order], t[order]
    # threshold index i = "between s[i-1] and s[i]": FRR = targets below, FAR = non-targets at or above
    frr = np.concatenate(([0], np.cumsum(t))) / t.sum()
    far = 1.0 - np.concatenate(([0], np.cumsum(~t))) / (~t).sum()
    thr = np.concatenate(([s[0] - 1e-6], 0.5 * (s[1:] + s[:-1]), [s[-1] + 1e-6]))
    return frr, far, thr


def eer(scores, target):
    """(EER, threshold)."""
    frr, far, thr = _rates(scores, target)
    i = int(np.argmin(np.abs(frr - far)))
    return float(0.5 * (frr[i] + far[i])), float(thr[i])


def min_dcf(scores, target, p_target: float = 0.5, c_miss: float = 1.0, c_fa: float = 1.0):
    """(minDCF, threshold); p_target = 0.5 is what the reference passes (plda_score_stat.py:97)."""
    frr, far, thr = _rates(scores, target)
    dcf = c_miss * p_target * frr + c_fa * (1.0 - p_target) * far
    i = int(np.argmin(dcf))
    return float(dcf[i]), float(thr[i])
