"""ctgcn_b200 — B200-native (sm_100a) implementation of CTGCN's forward hot path.

    from ctgcn_b200.layers import CoreDiffusion, MLP
    from ctgcn_b200.models import CDN, CGCN, CTGCN

mirror the reference's ``layers`` / ``models`` modules (same signatures, same state_dict keys) and run on
hand-written CUDA kernels behind the C-ABI of include/ctgcn_b200.h.  ``install_as_reference_modules()``
registers them under the names the reference's train.py / embedding.py import.
"""
from __future__ import annotations

import sys

from . import _lib  # raises ImportError if libctgcn_b200.so has not been built — no fallback
from . import dist, layers, loss, models, ops, plan
from .layers import CoreDiffusion, MLP
from .models import CDN, CGCN, CTGCN

__all__ = ["CoreDiffusion", "MLP", "CDN", "CGCN", "CTGCN", "layers", "models", "ops", "plan", "dist", "loss",
           "install_as_reference_modules"]


def install_as_reference_modules(ref_models=None, ref_metrics=None):
    """Make ``from layers import CoreDiffusion, MLP`` / ``from models import CGCN, CTGCN`` (reference
    models.py:4, train.py:101) resolve to this package.  If the reference's own ``models`` module is
    already imported (it also defines the classifier heads that stay reference code), pass it and only the
    hot-path classes are rebound.  Optional: pass the reference's ``metrics`` module BEFORE
    ``import train`` (train.py:8 binds the name at import) to rebind ``NegativeSamplingLoss`` to the device-side one."""
    if ref_metrics is not None:
        ref_metrics.NegativeSamplingLoss = loss.NegativeSamplingLoss
    sys.modules["layers"] = layers
    if ref_models is None:
        sys.modules["models"] = models
        return
    for name in ("CoreDiffusion", "MLP"):
        setattr(ref_models, name, getattr(layers, name))
    for name in ("CDN", "CGCN", "CTGCN"):
        setattr(ref_models, name, getattr(models, name))
