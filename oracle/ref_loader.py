"""Import the UNMODIFIED reference modules (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may call this; it is used by
``tests/golden/make_golden.py`` (fixture generation) and by the CPU tests that
re-validate the restatement whenever the reference happens to be reachable.

``tdnn_layer.py`` only needs torch.  ``main.py`` imports pytorch_lightning,
torchmetrics, speechbrain, resampy, python_speech_features, matplotlib and
seaborn at module level (main.py:1-20, dataset.py, plda_*.py); none are
installed and none are touched by ``extract_x_vec`` / ``forward``, so they are
replaced by inert stub modules before the import.  The ``__main__`` guard
(main.py:176) keeps the script body from running.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# $XVEC_REF_DIR, the build container's read-only mount, and where a driver-side `pip install --target baseline/_ref` would put it
_DEFAULT_DIRS = ("/root/reference", os.path.join(_REPO, "baseline", "_ref"))


def reference_dir() -> str | None:
    for d in (os.environ.get("XVEC_REF_DIR"),) + _DEFAULT_DIRS:
        if d and os.path.isfile(os.path.join(d, "tdnn_layer.py")) and os.path.isfile(os.path.join(d, "main.py")):
            return d
    return None


def available() -> bool:
    return reference_dir() is not None


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__xvec_stub__"] = True
        sys.modules[name] = mod
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_stub(parent), leaf, mod)
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def _install_stubs() -> None:
    import torch.nn as nn

    class LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class _Inert(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    class _Any:
        def __init__(self, *a, **k):
            pass

    def _needs(name):
        try:
            importlib.import_module(name)
            return False
        except Exception:
            return True

    if _needs("pytorch_lightning"):
        _stub("pytorch_lightning", LightningModule=LightningModule, Trainer=_Any)
        _stub("pytorch_lightning.loggers", TensorBoardLogger=_Any)
        _stub("pytorch_lightning.callbacks", ModelCheckpoint=_Any)
        _stub("pytorch_lightning.callbacks.early_stopping", EarlyStopping=_Any)
    if _needs("torchmetrics"):
        _stub("torchmetrics", Accuracy=_Inert)
    if _needs("torch.utils.tensorboard"):
        _stub("torch.utils.tensorboard", SummaryWriter=_Any)
    if _needs("resampy"):
        _stub("resampy", resample=lambda *a, **k: None)
    if _needs("python_speech_features"):
        _stub("python_speech_features", mfcc=lambda *a, **k: None)
    if _needs("speechbrain"):
        _stub("speechbrain")
        _stub("speechbrain.processing")
        m = _stub("speechbrain.processing.PLDA_LDA", StatObject_SB=_Any, PLDA=_Any, Ndx=_Any, LDA=_Any,
                  fast_PLDA_scoring=lambda *a, **k: None)
        m.__all__ = ["StatObject_SB", "PLDA", "Ndx", "LDA", "fast_PLDA_scoring"]
        _stub("speechbrain.utils")
        _stub("speechbrain.utils.metric_stats", EER=lambda *a, **k: None, minDCF=lambda *a, **k: None)
    if _needs("matplotlib"):
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
    if _needs("seaborn"):
        _stub("seaborn")


def load():
    """Return (tdnn_layer_module, main_module) of the real reference."""
    d = reference_dir()
    if d is None:
        raise RuntimeError("reference sources not reachable (set XVEC_REF_DIR)")
    _install_stubs()
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        tdnn_layer = importlib.import_module("tdnn_layer")
        main = importlib.import_module("main")
    finally:
        # keep sys.path clean for the rest of the test session; modules stay cached
        if d in sys.path:
            sys.path.remove(d)
    return tdnn_layer, main
