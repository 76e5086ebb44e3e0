import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """-> (state_dict, inputs, outputs, grads, meta) as dicts of torch tensors / numpy."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    parts = {"sd": {}, "in": {}, "out": {}, "grad": {}, "meta": {}}
    for k in z.files:
        head, rest = k.split("/", 1)
        parts[head][rest] = z[k] if head == "meta" else torch.from_numpy(z[k])
    return parts["sd"], parts["in"], parts["out"], parts["grad"], parts["meta"]


def golden(name):
    """The raw npz of a fixture (numpy arrays keyed "in/..", "out/..")."""
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rel_err(a, b):
    """norm-relative error  ||a-b||_inf / ||b||_inf  (SURVEY.md §8c recommended metric)."""
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
