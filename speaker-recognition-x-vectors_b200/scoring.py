"""Trial scoring on top of the extracted x-vectors.

The reference scores with SpeechBrain's PLDA over the full N x N matrix and then looks trial pairs up one by one
(plda_score_stat.py:59-87), and takes EER / minDCF(p_target=0.5) from speechbrain.utils.metric_stats (:92-97).  SpeechBrain
is not vendored, so those semantics are unpinned; here the metrics are defined from first principles:
  * score = centred cosine of the trial's two x-vectors (GPU kernel xvec_cosine_trials, BASELINE.json config 5), or the PLDA
    log-likelihood ratio of a given two-covariance model (PldaScorer; the published fast-scoring algebra, training out of scope),
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


class PldaScorer:
    """PLDA log-likelihood-ratio trial scoring on the GPU for a trained two-covariance model (mean (D), F (D, rank),
    Sigma (D, D)) — the attributes of the object the reference pickles (plda_classifier.py:40-50, :89-95).
    replaces: plda_classifier.plda_scores (:81-87) + the trial lookup of plda_score_stat.py:59-87; the N x N score matrix is never
    built.  Model-only quantities (Phi, Psi, the constant) are derived once on the host in float64; everything per x-vector and
    per trial runs in libxvec_b200 (two split-TF32 GEMMs on the tcgen05 kernel + three small kernels).  Parity with SpeechBrain is unpinned (oracle/plda_oracle.py)."""

    def __init__(self, mean, F, Sigma, scaling_factor: float = 1.0, device="cuda"):
        mean = np.asarray(mean, dtype=np.float64).reshape(-1)
        F = np.asarray(F, dtype=np.float64)
        Sigma = np.asarray(Sigma, dtype=np.float64)
        d = mean.shape[0]
        if F.ndim != 2 or F.shape[0] != d or Sigma.shape != (d, d):
            raise ValueError("expected mean (D), F (D, rank), Sigma (D, D)")
        inv_sigma = np.linalg.inv(Sigma)
        eye = np.eye(F.shape[1])
        K = F.T @ inv_sigma @ F
        self.cst = float(-0.5 * np.linalg.slogdet(2.0 * K + eye)[1] + np.linalg.slogdet(K + eye)[1])
        ac = F @ F.T
        tot_inv = np.linalg.inv(ac + Sigma)
        T = np.linalg.inv(ac + Sigma - ac @ tot_inv @ ac)
        phi, psi = tot_inv - T, tot_inv @ ac @ T
        dev = torch.device(device)
        f32 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
        # y = x W' + b with W = Phi' (nn.Linear convention) and b = -mean Phi centres the x-vectors inside the GEMM
        def split3(w):  # [W_hi | W_hi | W_lo], hi = the part exactly representable in TF32 (13 low mantissa bits cleared)
            w32 = np.ascontiguousarray(w, dtype=np.float32)
            hi = (w32.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
            return np.concatenate([hi, hi, w32 - hi], axis=1)

        self.w_phi, self.b_phi = ops.pack_weight(f32(split3(phi.T)), 1, 3 * d, torch.float32), ops.pad32(f32(-(mean @ phi)))
        self.w_psi, self.b_psi = ops.pack_weight(f32(split3(psi.T)), 1, 3 * d, torch.float32), ops.pad32(f32(-(mean @ psi)))
        self.mean = f32(mean)
        self.dim, self.scale, self.device = d, float(scaling_factor), dev

    def score_trials(self, xvecs, enrol_idx, test_idx) -> np.ndarray:
        x = xvecs if isinstance(xvecs, torch.Tensor) else torch.as_tensor(np.asarray(xvecs, dtype=np.float32))
        x = x.to(self.device).float().contiguous()
        if x.dim() != 2 or x.shape[1] != self.dim:
            raise ValueError(f"expected x-vectors of dimension {self.dim}")
        e = torch.as_tensor(np.asarray(enrol_idx, dtype=np.int32)).to(self.device)
        t = torch.as_tensor(np.asarray(test_idx, dtype=np.int32)).to(self.device)
        if e.numel() and (int(e.max()) >= x.shape[0] or int(t.max()) >= x.shape[0] or int(e.min()) < 0 or int(t.min()) < 0):
            raise ValueError("trial index out of range")
        return ops.plda_trials(x, self.mean, self.w_phi, self.b_phi, self.w_psi, self.b_psi, self.cst, self.scale, e, t).cpu().numpy()


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
