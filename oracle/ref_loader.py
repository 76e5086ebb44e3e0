"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

The reference (TUT-SLP-lab/MultimodalReactionGeneration) needs pytorch_lightning,
omegaconf and torchmetrics, none of which exist in this image.  This loader installs
three tiny in-memory stand-ins and registers ``mr_gen`` / ``mr_gen.utils`` as bare
namespace modules so the heavy ``__init__`` files (mediapipe, dfcon, ...) are skipped
(SURVEY.md Appendix E).  No reference source is copied: the model files are executed
from where they lie.

Only ``oracle/make_golden.py`` (run in the build container) uses this.  /root/reference
does not exist on the GPU box, so nothing under tests/, bench.py or smoke() imports it.
"""
from __future__ import annotations

import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("MRG_REFERENCE_ROOT", "/root/reference")


class AttrDict(dict):
    """Minimal omegaconf.DictConfig stand-in: attribute access + .get."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:  # pragma: no cover
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = value


def _install_stubs() -> None:
    if "pytorch_lightning" in sys.modules and getattr(
        sys.modules["pytorch_lightning"], "_mrg_stub", False
    ):
        return

    # ---- pytorch_lightning ------------------------------------------------------
    pl = types.ModuleType("pytorch_lightning")
    pl._mrg_stub = True

    class LightningModule(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self.current_epoch = 0

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, *a, **k):
            return None

        def log_dict(self, *a, **k):
            return None

    class LightningDataModule:
        def __init__(self, *a, **k):
            pass

    class Trainer:
        def __init__(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = LightningDataModule
    pl.Trainer = Trainer
    util = types.ModuleType("pytorch_lightning.utilities")
    util_types = types.ModuleType("pytorch_lightning.utilities.types")
    util_types.STEP_OUTPUT = object
    util_types.EVAL_DATALOADERS = object
    util_types.TRAIN_DATALOADERS = object
    util.types = util_types
    pl.utilities = util
    sys.modules["pytorch_lightning"] = pl
    sys.modules["pytorch_lightning.utilities"] = util
    sys.modules["pytorch_lightning.utilities.types"] = util_types

    # ---- omegaconf --------------------------------------------------------------
    oc = types.ModuleType("omegaconf")
    oc.DictConfig = AttrDict
    sys.modules["omegaconf"] = oc

    # ---- torchmetrics (logging only; never touches loss/grad) -------------------
    tm = types.ModuleType("torchmetrics")

    class Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            return None

        def update(self, *a, **k):
            return None

    class MeanSquaredError(Metric):
        pass

    class MetricCollection(nn.Module):
        def __init__(self, metrics=None, *a, **k):
            super().__init__()

        def forward(self, *a, **k):
            return None

    tm.Metric = Metric
    tm.MeanSquaredError = MeanSquaredError
    tm.MetricCollection = MetricCollection
    sys.modules["torchmetrics"] = tm

    # ---- mr_gen namespace shells ------------------------------------------------
    root = os.path.join(REFERENCE_ROOT, "mr_gen")
    for name, path in (
        ("mr_gen", root),
        ("mr_gen.utils", os.path.join(root, "utils")),
    ):
        mod = types.ModuleType(name)
        mod.__path__ = [path]
        sys.modules[name] = mod
    db = types.ModuleType("mr_gen.databuild")
    db.DataBuilder = object
    db.DataBuilderNX = object
    sys.modules["mr_gen.databuild"] = db
    pp = types.ModuleType("mr_gen.utils.preprocess")
    pp.AudioPreprocessor = object
    pp.MotionPreprocessor = object
    pp.MotionPreprocessorNX = object
    sys.modules["mr_gen.utils.preprocess"] = pp


def load_reference():
    """Return a namespace with the reference's model classes (unmodified code)."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    ns = types.SimpleNamespace()
    from mr_gen.model.utils.lstm_block import LSTMModule, LSTMBlock, LSTMLayerd
    from mr_gen.model.utils.lstm_sampler import LSTMSampler
    from mr_gen.model.utils.residual_connection import ResidualConnection
    from mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM

    ns.LSTMModule, ns.LSTMBlock, ns.LSTMLayerd = LSTMModule, LSTMBlock, LSTMLayerd
    ns.LSTMSampler, ns.ResidualConnection = LSTMSampler, ResidualConnection
    ns.LSTMwithSample, ns.SimpleLSTM = LSTMwithSample, SimpleLSTM
    try:
        from mr_gen.model.utils.mixer_block import (
            LSTMMixer,
            LSTMMixerBlock,
            LSTMMixerLayerd,
        )

        ns.LSTMMixer, ns.LSTMMixerBlock, ns.LSTMMixerLayerd = (
            LSTMMixer,
            LSTMMixerBlock,
            LSTMMixerLayerd,
        )
    except Exception as exc:  # pragma: no cover
        ns.mixer_import_error = exc
    ns.AttrDict = AttrDict
    return ns
