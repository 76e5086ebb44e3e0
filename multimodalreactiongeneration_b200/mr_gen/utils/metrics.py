"""Per-target-slice MSE bookkeeping (reference: mr_gen/utils/metrics/multi_modal_metrics.py:6-56,
built on torchmetrics).  Logging only — it never touches loss or gradients."""
from typing import Dict, Tuple

import torch
from torch import nn


def gen_target_dict(cfg) -> Dict[str, Tuple[int, int]]:
    """Slices of the 6-d (x3 with deltas) pose vector: centroid / angle / their deltas
    (simple_lstm.py:15-45, lstm_with_sample.py:25-56)."""
    names = [n for n, on in (("centroid", cfg.use_centroid), ("angle", cfg.use_angle)) if on]
    out, pos = {}, 0
    for order in range(cfg.delta_order + 1):
        for n in names:
            out[n if order == 0 else f"delta{order}-{n}"] = (pos, pos + 3)
            pos += 3
    return out


class MultiTargetMetrics(nn.Module):
    def __init__(self, target_range: Dict[str, Tuple[int, int]], prefix: str = ""):
        super().__init__()
        self.target_range = dict(target_range)
        self.prefix = prefix
        self.last: Dict[str, float] = {}

    @torch.no_grad()
    def forward(self, pred: torch.Tensor, target: torch.Tensor):
        vals = {self.prefix + k: torch.mean((pred[..., a:b] - target[..., a:b]) ** 2)
                for k, (a, b) in self.target_range.items()}
        self.last = vals
        return vals
