"""CPU restatement of the reference's x-vector extraction path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity status: PINNED —
tests/test_oracle.py checks every function here against the reference's own
known-answer fixtures and against tests/golden/*.npz, which were produced by
the unmodified reference (tests/golden/make_golden.py).

Written independently of the reference source: the time-context unfold is an
index gather (not slice+cat), the TDNN layer is gather -> matmul -> clamp ->
explicit affine BatchNorm, pooling is explicit sums.  Two arithmetic flavours:

* ``*_t``  : torch CPU float32 — the arithmetic type of the reference
             (main.py:137 ``samples.float()``); this is what CUDA results are
             compared with, and what bench.py times as the CPU baseline.
* ``*_np`` : numpy float64 — a "truth" with negligible rounding, used to put
             the fp32 oracle's own error and the CUDA error on the same scale.

Reference map
    time_context_index   tdnn_layer.py:43-60   get_time_context
    tdnn_layer_*         tdnn_layer.py:26-41   TdnnLayer.forward
    LAYER_CONTEXTS       main.py:38-44         time_context_layers
    stat_pool_*          main.py:59-63         XVectorModel.stat_pool
    extract_x_vec_*      main.py:81-94         XVectorModel.extract_x_vec
    forward_*            main.py:66-75         XVectorModel.forward
    make_state_dict      main.py:38-47 + torch default initialisers
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence

import numpy as np
import torch

# main.py:39-43 — (context, in, out) of the five frame-level layers
LAYER_CONTEXTS: List[List[int]] = [[-2, -1, 0, 1, 2], [-2, 0, 2], [-3, 0, 3], [0], [0]]
BN_EPS = 1e-5  # nn.BatchNorm1d default (tdnn_layer.py:22)
POOL_DIM = 1500
TOTAL_CONTEXT = sum(c[-1] - c[0] for c in LAYER_CONTEXTS)  # 14 frames


def layer_sizes(input_size=24, hidden_size=512):
    """(Cin, N) per TDNN layer, main.py:39-43."""
    return [(input_size, hidden_size), (hidden_size, hidden_size), (hidden_size, hidden_size),
            (hidden_size, hidden_size), (hidden_size, POOL_DIM)]


# --------------------------------------------------------------------------- unfold
def time_context_index(T: int, context: Sequence[int]) -> np.ndarray:
    """Frame indices read by every output frame: (T_out, k) int64.

    tdnn_layer.py:43-60: tap j of output frame t is input frame
    ``t + (c_j - c_0)``; T_out = T - (c_last - c_0).  The reference's slicing is
    only self-consistent for symmetric contexts (c_last == -c_0); an asymmetric
    context makes its torch.cat fail, which is mirrored here as ValueError.
    """
    c = list(context)
    if len(c) == 0:
        raise ValueError("empty context")
    if len(c) > 1 and c[-1] != -c[0]:
        raise ValueError(f"context {c} is not symmetric; the reference cannot concatenate its views")
    span = c[-1] - c[0]
    t_out = T - span
    if t_out <= 0:
        raise ValueError(f"input of {T} frames is shorter than the context span {span}+1")
    offs = np.asarray([cj - c[0] for cj in c], dtype=np.int64)
    return np.arange(t_out, dtype=np.int64)[:, None] + offs[None, :]


def unfold_np(x: np.ndarray, context: Sequence[int]) -> np.ndarray:
    """(B,T,C) -> (B,T_out,k*C), context-major like torch.cat(views, 2) at tdnn_layer.py:29."""
    idx = time_context_index(x.shape[1], context)
    g = x[:, idx, :]  # (B, T_out, k, C)
    return g.reshape(x.shape[0], idx.shape[0], -1)


def unfold_t(x: torch.Tensor, context: Sequence[int]) -> torch.Tensor:
    idx = torch.from_numpy(time_context_index(x.shape[1], context))
    g = x[:, idx, :]
    return g.reshape(x.shape[0], idx.shape[0], -1)


# --------------------------------------------------------------------------- TDNN layer
def bn_affine(gamma, beta, mean, var, eps=BN_EPS):
    """Eval-mode BatchNorm1d as y = r*s + h (tdnn_layer.py:36-39)."""
    if isinstance(gamma, torch.Tensor):
        s = gamma / torch.sqrt(var + eps)
    else:
        s = gamma / np.sqrt(var + eps)
    return s, beta - mean * s


def tdnn_layer_t(x, W, b, context, bn=None):
    """tdnn_layer.py:26-41 in eval mode: unfold -> Linear -> ReLU -> [Dropout=id] -> BN(running stats)."""
    u = unfold_t(x, context)
    y = torch.clamp_min(u @ W.t() + b, 0.0)
    if bn is not None:
        gamma, beta, mean, var = bn
        y = (y - mean) / torch.sqrt(var + BN_EPS) * gamma + beta
    return y


def tdnn_layer_np(x, W, b, context, bn=None):
    u = unfold_np(x, context)
    y = np.maximum(u @ W.T + b, 0.0)
    if bn is not None:
        gamma, beta, mean, var = bn
        y = (y - mean) / np.sqrt(var + BN_EPS) * gamma + beta
    return y


# --------------------------------------------------------------------------- pooling
def stat_pool_t(x: torch.Tensor) -> torch.Tensor:
    """main.py:59-63: [mean over time || unbiased std over time]; n-1 == 0 gives NaN like torch.std."""
    n = x.shape[1]
    mean = x.sum(1) / n
    d = x - mean[:, None, :]
    var = (d * d).sum(1) / (n - 1) if n > 1 else torch.full_like(mean, float("nan"))
    return torch.cat((mean, torch.sqrt(var)), 1)


def stat_pool_np(x: np.ndarray) -> np.ndarray:
    n = x.shape[1]
    mean = x.mean(1)
    with np.errstate(invalid="ignore", divide="ignore"):
        var = ((x - mean[:, None, :]) ** 2).sum(1) / (n - 1) if n > 1 else np.full_like(mean, np.nan)
    return np.concatenate((mean, np.sqrt(var)), 1)


# --------------------------------------------------------------------------- parameters
def _bn_of(sd, i, lib):
    p = f"time_context_layers.{i}.norm."
    if p + "weight" not in sd:
        return None
    return tuple(lib(sd[p + k]) for k in ("weight", "bias", "running_mean", "running_var"))


def _as_t(v):
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _as_np64(v):
    return (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)).astype(np.float64)


def make_state_dict(seed: int = 0, input_size=24, hidden_size=512, num_classes=1211, x_vector_size=512,
                    batch_norm=True, randomize_bn: bool = True, bn_seed: int = 7) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's key names (SURVEY §8a parameter inventory).

    Modules are created in the order main.py:38-47 / tdnn_layer.py:19-22 creates
    them, so with ``torch.manual_seed(seed)`` the tensors are bit-identical to
    ``main.XVectorModel()`` built after the same seed (checked by the golden
    hash).  BatchNorm statistics are then randomised — default-init
    BN in eval mode is the identity up to eps and would hide fold errors.
    """
    import torch.nn as nn

    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for i, ((cin, n), ctx) in enumerate(zip(layer_sizes(input_size, hidden_size), LAYER_CONTEXTS)):
        lin = nn.Linear(cin * len(ctx), n)
        sd[f"time_context_layers.{i}.linear.weight"] = lin.weight.detach().clone()
        sd[f"time_context_layers.{i}.linear.bias"] = lin.bias.detach().clone()
        if batch_norm:
            sd[f"time_context_layers.{i}.norm.weight"] = torch.ones(n)
            sd[f"time_context_layers.{i}.norm.bias"] = torch.zeros(n)
            sd[f"time_context_layers.{i}.norm.running_mean"] = torch.zeros(n)
            sd[f"time_context_layers.{i}.norm.running_var"] = torch.ones(n)
            sd[f"time_context_layers.{i}.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for name, (fin, fout) in (("segment_layer6", (2 * POOL_DIM, x_vector_size)),
                              ("segment_layer7", (x_vector_size, x_vector_size)),
                              ("output", (x_vector_size, num_classes))):
        lin = nn.Linear(fin, fout)
        sd[name + ".weight"] = lin.weight.detach().clone()
        sd[name + ".bias"] = lin.bias.detach().clone()
    if batch_norm and randomize_bn:
        randomize_bn_stats(sd, bn_seed)
    return sd


def randomize_bn_stats(sd: Dict[str, torch.Tensor], seed: int = 7) -> None:
    """In place: running_mean~N(0,.5), running_var~U(.3,2), gamma~N(1,.5) (some negative), beta~N(0,.5)."""
    g = torch.Generator().manual_seed(seed)
    for i in range(5):
        p = f"time_context_layers.{i}.norm."
        n = sd[p + "weight"].numel()
        sd[p + "running_mean"] = torch.randn(n, generator=g) * 0.5
        sd[p + "running_var"] = torch.rand(n, generator=g) * 1.7 + 0.3
        sd[p + "weight"] = 1.0 + torch.randn(n, generator=g) * 0.5
        sd[p + "bias"] = torch.randn(n, generator=g) * 0.5


def state_dict_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------- model
def tdnn_stack_t(sd, x: torch.Tensor) -> torch.Tensor:
    """main.py:38-44,82 — five TdnnLayers; (B,T,24) -> (B,T-14,1500)."""
    out = x
    for i, ctx in enumerate(LAYER_CONTEXTS):
        out = tdnn_layer_t(out, _as_t(sd[f"time_context_layers.{i}.linear.weight"]),
                           _as_t(sd[f"time_context_layers.{i}.linear.bias"]), ctx, _bn_of(sd, i, _as_t))
    return out


def tdnn_stack_np(sd, x: np.ndarray) -> np.ndarray:
    out = x.astype(np.float64)
    for i, ctx in enumerate(LAYER_CONTEXTS):
        out = tdnn_layer_np(out, _as_np64(sd[f"time_context_layers.{i}.linear.weight"]),
                            _as_np64(sd[f"time_context_layers.{i}.linear.bias"]), ctx, _bn_of(sd, i, _as_np64))
    return out


def _head_t(sd, pooled, layer):
    w6, b6 = _as_t(sd["segment_layer6.weight"]), _as_t(sd["segment_layer6.bias"])
    s6 = pooled @ w6.t() + b6
    if layer == 7:  # main.py:88-90; every other value behaves as 6 (main.py:86-87,91-92)
        w7, b7 = _as_t(sd["segment_layer7.weight"]), _as_t(sd["segment_layer7.bias"])
        return torch.clamp_min(s6, 0.0) @ w7.t() + b7
    return s6


def _head_np(sd, pooled, layer):
    s6 = pooled @ _as_np64(sd["segment_layer6.weight"]).T + _as_np64(sd["segment_layer6.bias"])
    if layer == 7:
        return np.maximum(s6, 0.0) @ _as_np64(sd["segment_layer7.weight"]).T + _as_np64(sd["segment_layer7.bias"])
    return s6


@torch.no_grad()
def extract_x_vec_t(sd, x: torch.Tensor, x_vec_extract_layer: int = 6) -> torch.Tensor:
    """main.py:81-94 — (B,T,24) float32 -> (B,512) float32, pre-ReLU affine output of layer 6 or 7."""
    return _head_t(sd, stat_pool_t(tdnn_stack_t(sd, x)), x_vec_extract_layer)


def extract_x_vec_np(sd, x: np.ndarray, x_vec_extract_layer: int = 6) -> np.ndarray:
    return _head_np(sd, stat_pool_np(tdnn_stack_np(sd, x)), x_vec_extract_layer)


@torch.no_grad()
def extract_x_vec_aten(sd, x: torch.Tensor, x_vec_extract_layer: int = 6) -> torch.Tensor:
    """Same result as extract_x_vec_t, but issuing the ATen operator sequence the reference issues on this path — k shifted
    views + cat (tdnn_layer.py:28-29), addmm (:30), clamp_min (:31), native_batch_norm in eval mode on the transposed view
    (:36-39), mean + std + cat (main.py:60-62), addmm (main.py:87-90) — so that its wall time is the reference's CPU time.
    This is what bench.py times as the CPU baseline / `--impl reference` arm."""
    import torch.nn.functional as F
    h = x
    for i, ctx in enumerate(LAYER_CONTEXTS):
        span, t_in = ctx[-1] - ctx[0], h.shape[1]
        views = [h[:, c - ctx[0]: t_in - span + (c - ctx[0]), :] for c in ctx]
        h = F.relu(F.linear(torch.cat(views, 2), _as_t(sd[f"time_context_layers.{i}.linear.weight"]),
                            _as_t(sd[f"time_context_layers.{i}.linear.bias"])))
        bn = _bn_of(sd, i, _as_t)
        if bn is not None:
            gamma, beta, mean, var = bn
            h = F.batch_norm(h.transpose(1, 2), mean, var, gamma, beta, False, 0.1, BN_EPS).transpose(1, 2)
    pooled = torch.cat((torch.mean(h, 1), torch.std(h, 1)), 1)
    s6 = F.linear(pooled, _as_t(sd["segment_layer6.weight"]), _as_t(sd["segment_layer6.bias"]))
    if x_vec_extract_layer == 7:
        return F.linear(F.relu(s6), _as_t(sd["segment_layer7.weight"]), _as_t(sd["segment_layer7.bias"]))
    return s6


@torch.no_grad()
def forward_t(sd, x: torch.Tensor) -> torch.Tensor:
    """main.py:66-75 — classifier logits (B,num_classes): relu(seg6) -> relu(seg7) -> output."""
    p = stat_pool_t(tdnn_stack_t(sd, x))
    h = torch.clamp_min(p @ _as_t(sd["segment_layer6.weight"]).t() + _as_t(sd["segment_layer6.bias"]), 0.0)
    h = torch.clamp_min(h @ _as_t(sd["segment_layer7.weight"]).t() + _as_t(sd["segment_layer7.bias"]), 0.0)
    return h @ _as_t(sd["output.weight"]).t() + _as_t(sd["output.bias"])


@torch.no_grad()
def extract_ragged_t(sd, utts: Sequence[torch.Tensor], x_vec_extract_layer: int = 6, chunk: int = 64) -> torch.Tensor:
    """Ragged oracle = the reference run on each utterance alone at its true length
    (SURVEY §5 long-context row: batched == per-utterance in eval mode).  Equal-length
    utterances are batched together purely to save time; results are per-utterance."""
    out = torch.empty(len(utts), _as_t(sd["segment_layer6.weight"]).shape[0] if x_vec_extract_layer != 7
                      else _as_t(sd["segment_layer7.weight"]).shape[0])
    by_len: Dict[int, List[int]] = {}
    for i, u in enumerate(utts):
        by_len.setdefault(int(u.shape[0]), []).append(i)
    for _, idxs in by_len.items():
        for s in range(0, len(idxs), chunk):
            sel = idxs[s:s + chunk]
            out[sel] = extract_x_vec_t(sd, torch.stack([utts[i] for i in sel]).float(), x_vec_extract_layer)
    return out


# --------------------------------------------------------------------------- synthetic inputs
def synth_mfcc(n_utts: int, n_frames: int, n_ceps: int = 24, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n_utts, n_frames, n_ceps, generator=g)


def synth_lengths(n_utts: int, lo: int, hi: int, seed: int) -> np.ndarray:
    """Uniform integer frame counts in [lo, hi]."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(lo, hi + 1, (n_utts,), generator=g).numpy().astype(np.int64)


def synth_ragged(lengths: Sequence[int], n_ceps: int = 24, seed: int = 1234) -> List[torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    flat = torch.randn(int(np.sum(lengths)), n_ceps, generator=g)
    return list(torch.split(flat, [int(v) for v in lengths]))


def synth_trials(n_utts: int, n_trials: int, n_speakers: int = 40, seed: int = 4):
    """Balanced synthetic trial list: (enrol_idx, test_idx, is_target) with speakers assigned round-robin."""
    rng = np.random.default_rng(seed)
    spk = np.arange(n_utts) % n_speakers
    n_tar = n_trials // 2
    e_t = rng.integers(0, n_utts, n_tar)
    # a target partner: another utterance of the same speaker
    per_spk = [np.nonzero(spk == s)[0] for s in range(n_speakers)]
    t_t = np.array([rng.choice(per_spk[spk[e]]) for e in e_t])
    n_non = n_trials - n_tar
    e_n = rng.integers(0, n_utts, n_non)
    t_n = rng.integers(0, n_utts, n_non)
    clash = spk[e_n] == spk[t_n]
    t_n[clash] = (t_n[clash] + 1) % n_utts  # round-robin speakers: the next utterance is another speaker
    enrol = np.concatenate([e_t, e_n])
    test = np.concatenate([t_t, t_n])
    target = np.concatenate([np.ones(n_tar, bool), np.zeros(n_non, bool)])
    perm = rng.permutation(n_trials)
    return enrol[perm], test[perm], target[perm]


def synth_speaker_utts(lengths: Sequence[int], n_speakers: int, n_ceps: int = 24, seed: int = 55, spk_sigma: float = 1.0,
                       sess_sigma: float = 0.1) -> List[torch.Tensor]:
    """Ragged synthetic MFCCs with speaker structure: frames ~ N(0,1) + a per-speaker offset (round-robin speakers,
    matching synth_trials) + a small per-utterance session offset, so cosine trials are meaningful."""
    g = torch.Generator().manual_seed(seed)
    spk = torch.randn(n_speakers, n_ceps, generator=g) * spk_sigma
    out = []
    for i, l in enumerate(lengths):
        sess = torch.randn(n_ceps, generator=g) * sess_sigma
        out.append(torch.randn(int(l), n_ceps, generator=g) + spk[i % n_speakers] + sess)
    return out


def cosine_scores_np(xvecs: np.ndarray, enrol: np.ndarray, test: np.ndarray, center: bool = False) -> np.ndarray:
    """fp64 cosine score per trial (BASELINE.json config 5; not present in the reference).  center=True subtracts the
    mean x-vector of the set first (random-init embeddings share a large common component)."""
    x = np.asarray(xvecs, dtype=np.float64)
    if center:
        x = x - x.mean(0, keepdims=True)
    x = x / np.linalg.norm(x, axis=1, keepdims=True)
    return np.einsum("ij,ij->i", x[enrol], x[test])


def eer_threshold_np(scores: np.ndarray, target: np.ndarray):
    """Equal-error-rate operating point from first principles: returns (eer, threshold, margin) where the
    threshold is placed in the middle of the widest score gap among the candidates closest to FAR==FRR,
    and margin is half that gap (so decisions are robust to score perturbations < margin)."""
    order = np.argsort(scores)
    s = scores[order]
    t = target[order]
    n_tar, n_non = t.sum(), (~t).sum()
    # threshold between s[i-1] and s[i]: FRR = targets below, FAR = non-targets at/above
    frr = np.concatenate([[0], np.cumsum(t)]) / n_tar
    far = 1.0 - np.concatenate([[0], np.cumsum(~t)]) / n_non
    diff = np.abs(frr - far)
    cand = np.nonzero(diff <= diff.min() + 2.0 / min(n_tar, n_non))[0]
    cand = cand[(cand > 0) & (cand < len(s))]
    gaps = s[cand] - s[cand - 1]
    j = cand[int(np.argmax(gaps))]
    thr = 0.5 * (s[j] + s[j - 1])
    return float(0.5 * (frr[j] + far[j])), float(thr), float(0.5 * (s[j] - s[j - 1]))


def flops_per_utt(T: int, layer: int = 6) -> int:
    """Algorithmic FLOPs of one utterance of T input frames."""
    f = 122_880 * (T - 4) + 1_572_864 * (T - 8) + 3_633_152 * (T - 14) + 3_072_000
    if layer == 7:
        f += 524_288
    return f
