"""x-vector .csv in the reference's format.

Writer: main.py:246-247 does ``pd.DataFrame(x_vector).to_csv(path)`` on a list of ``(id, label, float64 ndarray)``
tuples, i.e. header ``,0,1,2`` and rows ``idx,id,label,"[ v0 v1 ...]"`` where the vector is ``str(ndarray)``
(whitespace separated, numpy print precision, wrapped over several lines inside one quoted cell).
Reader: main.py:276-279 / plda_score_stat.py:16-17 split the bracketed string: ``np.array(cell[1:-1].split(), float64)``.
"""
from __future__ import annotations

from typing import Iterable, Sequence, Tuple

import numpy as np
import pandas as pd


def to_records(ids: Sequence[str], labels: Sequence[int], xvecs) -> list:
    """The list test_epoch_end builds (main.py:141-146): (id, int label, float64 vector) per utterance."""
    x = np.asarray(xvecs, dtype=np.float64)
    if x.ndim != 2 or len(ids) != x.shape[0] or len(labels) != x.shape[0]:
        raise ValueError("ids, labels and xvecs must describe the same number of utterances")
    return [(str(i), int(l), x[k]) for k, (i, l) in enumerate(zip(ids, labels))]


def write_xvector_csv(path: str, ids: Sequence[str], labels: Sequence[int], xvecs) -> None:
    pd.DataFrame(to_records(ids, labels, xvecs)).to_csv(path)


def read_xvector_csv(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Returns (ids object array, labels int64, xvecs float64 (N, D)) exactly like the reference's consumers parse it."""
    df = pd.read_csv(path)
    df.columns = ["index", "id", "label", "xvector"]  # main.py:318
    ids = np.array(df.iloc[:, 1])
    labels = np.array(df.iloc[:, 2], dtype=np.int64)
    xv = np.array([np.array(cell[1:-1].split(), dtype=np.float64) for cell in df.iloc[:, 3]])
    return ids, labels, xv


def parse_trial_file(lines: Iterable[str]):
    """VoxCeleb veri_test2.txt syntax '<0|1> <enrol_id> <test_id>' (parsed at plda_score_stat.py:64-67)."""
    target, enrol, test = [], [], []
    for pair in lines:
        if not pair.strip():
            continue
        parts = pair.split(" ")
        target.append(bool(int(parts[0].rstrip().split(".")[0].strip())))
        enrol.append(parts[1].strip())
        test.append(parts[2].strip())
    return np.asarray(target, dtype=bool), enrol, test
