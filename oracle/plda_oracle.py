"""CPU restatement of PLDA log-likelihood-ratio trial scoring.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity status: UNPINNED against the reference.  The reference scores with
``speechbrain.processing.PLDA_LDA.fast_PLDA_scoring`` (speechbrain==0.5.12, requirements.txt:55; call site
plda_classifier.py:81-87: ``fast_PLDA_scoring(en_stat, te_stat, ndx, plda.mean, plda.F, plda.Sigma, p_known=0.0)``,
trial lookup plda_score_stat.py:59-87).  SpeechBrain is neither vendored under /root/reference nor installed, and the
reference's own PLDA test (extra/plda_test_online_example.py) needs absent .pkl fixtures, so there is no golden vector.
``fast_plda_scoring`` below restates the published algorithm (SIDEKIT's fast PLDA scoring, which SpeechBrain 0.5 carries):
    centre by `mean`;  K = F' Sigma^-1 F;  cst = -1/2 logdet(2K + I) + logdet(K + I)
    Sigma_ac = F F';  Sigma_tot = Sigma_ac + Sigma;  T = (Sigma_tot - Sigma_ac Sigma_tot^-1 Sigma_ac)^-1
    Phi = Sigma_tot^-1 - T;  Psi = Sigma_tot^-1 Sigma_ac T
    score[i, j] = scaling * ( 1/2 e_i' Phi e_i + 1/2 t_j' Phi t_j + e_i' Psi t_j + cst )
What IS pinned (tests/test_oracle.py): the restated formula equals, to float64 rounding, the log-likelihood ratio of the
two-covariance model it claims to score, evaluated directly from its definition (``llr_direct``).
"""
from __future__ import annotations

import numpy as np


def plda_matrices(F: np.ndarray, Sigma: np.ndarray):
    """(Phi, Psi, cst) of the fast scoring formula, float64."""
    F = np.asarray(F, dtype=np.float64)
    Sigma = np.asarray(Sigma, dtype=np.float64)
    inv_sigma = np.linalg.inv(Sigma)
    eye = np.eye(F.shape[1])
    K = F.T @ inv_sigma @ F
    cst = -0.5 * np.linalg.slogdet(2.0 * K + eye)[1] + np.linalg.slogdet(K + eye)[1]
    sigma_ac = F @ F.T
    sigma_tot = sigma_ac + Sigma
    sigma_tot_inv = np.linalg.inv(sigma_tot)
    T = np.linalg.inv(sigma_tot - sigma_ac @ sigma_tot_inv @ sigma_ac)
    return sigma_tot_inv - T, sigma_tot_inv @ sigma_ac @ T, float(cst)


def fast_plda_scoring(enroll: np.ndarray, test: np.ndarray, mean: np.ndarray, F: np.ndarray, Sigma: np.ndarray,
                      scaling_factor: float = 1.0) -> np.ndarray:
    """Score matrix (n_enroll, n_test), float64 (the scoremat the reference indexes at plda_score_stat.py:82)."""
    phi, psi, cst = plda_matrices(F, Sigma)
    e = np.asarray(enroll, dtype=np.float64) - np.asarray(mean, dtype=np.float64)
    t = np.asarray(test, dtype=np.float64) - np.asarray(mean, dtype=np.float64)
    model_part = 0.5 * np.einsum("ij,ij->i", e @ phi, e)
    seg_part = 0.5 * np.einsum("ij,ij->i", t @ phi, t)
    return scaling_factor * (model_part[:, None] + seg_part[None, :] + cst + e @ psi @ t.T)


def trial_scores(x: np.ndarray, enrol_idx, test_idx, mean, F, Sigma, scaling_factor: float = 1.0) -> np.ndarray:
    """Scores of (enrol, test) index pairs into one x-vector matrix, without building the N x N matrix."""
    phi, psi, cst = plda_matrices(F, Sigma)
    xc = np.asarray(x, dtype=np.float64) - np.asarray(mean, dtype=np.float64)
    q = 0.5 * np.einsum("ij,ij->i", xc @ phi, xc)
    p = xc @ psi
    e, t = np.asarray(enrol_idx), np.asarray(test_idx)
    return scaling_factor * (q[e] + q[t] + np.einsum("ij,ij->i", p[e], xc[t]) + cst)


def llr_direct(e: np.ndarray, t: np.ndarray, mean, F, Sigma) -> float:
    """log p(e, t | same speaker) - log p(e) - log p(t) of the model x = mean + F y + eps, y ~ N(0, I), eps ~ N(0, Sigma),
    straight from the Gaussian densities (no algebra shared with fast_plda_scoring)."""
    F = np.asarray(F, dtype=np.float64)
    Sigma = np.asarray(Sigma, dtype=np.float64)
    e = np.asarray(e, dtype=np.float64) - mean
    t = np.asarray(t, dtype=np.float64) - mean
    ac = F @ F.T
    tot = ac + Sigma

    def logpdf(v, cov):
        sign, logdet = np.linalg.slogdet(cov)
        assert sign > 0
        return -0.5 * (v @ np.linalg.solve(cov, v) + logdet + v.size * np.log(2.0 * np.pi))

    joint = np.block([[tot, ac], [ac, tot]])
    return float(logpdf(np.concatenate([e, t]), joint) - logpdf(e, tot) - logpdf(t, tot))


def synth_plda(dim: int, rank: int, seed: int = 5):
    """A seeded, well-conditioned synthetic PLDA model (mean, F (dim, rank), Sigma (dim, dim) SPD)."""
    rng = np.random.default_rng(seed)
    mean = rng.standard_normal(dim) * 0.3
    F = rng.standard_normal((dim, rank)) / np.sqrt(rank)
    A = rng.standard_normal((dim, dim)) / np.sqrt(dim)
    Sigma = 0.5 * np.eye(dim) + 0.5 * (A @ A.T)
    return mean, F, Sigma
