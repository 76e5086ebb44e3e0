"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the GRU arithmetic of lstmformer's GRU mixer.

The reference constructs the third-party ``torch.nn.GRU`` at mr_gen/model/utils/mixer_block.py:194 and calls it at
:207 (un-pinned torch; 2.11.0 in this image).  Restated from torch's published definition,
``torch/nn/modules/rnn.py`` (class GRU):

    r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)          gate order in the packed weights: r, z, n
    z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)          (weight_ih_l{k} is [3H, I])
    n = tanh   (W_in x + b_in + r * (W_hn h + b_hn))
    h' = (1 - z) * n + z * h

Pinned against ``torch.nn.GRU`` executed on CPU in fp64 (tests/test_oracle_cpu.py) and, one level up, by the
``gru_mixer_layerd`` fixture generated from the unmodified reference (oracle/make_golden.py).  Only tests/ may
import this module; the product path never does.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_layer_forward(x, w_ih, w_hh, b_ih, b_hh, h0=None):
    """One direction of one layer, x time-major [T, B, I] -> y [T, B, H] and the cache for the backward."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = np.zeros((B, H), x.dtype) if h0 is None else h0.astype(x.dtype)
    y = np.zeros((T, B, H), x.dtype)
    cache = []
    for t in range(T):
        gx = x[t] @ w_ih.T + (b_ih if b_ih is not None else 0.0)
        gh = h @ w_hh.T + (b_hh if b_hh is not None else 0.0)
        r = _sigmoid(gx[:, :H] + gh[:, :H])
        z = _sigmoid(gx[:, H:2 * H] + gh[:, H:2 * H])
        hn = gh[:, 2 * H:]
        n = np.tanh(gx[:, 2 * H:] + r * hn)
        cache.append((h, r, z, n, hn))
        h = (1.0 - z) * n + z * h
        y[t] = h
    return y, cache


def gru_layer_backward(dy, x, cache, w_ih, w_hh):
    """-> dx, dw_ih, dw_hh, db_ih, db_hh, dh0 for the loss whose gradient w.r.t. y is ``dy`` (h_n = y[-1])."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    dx = np.zeros_like(x)
    dw_ih, dw_hh = np.zeros_like(w_ih), np.zeros_like(w_hh)
    db_ih, db_hh = np.zeros(3 * H, x.dtype), np.zeros(3 * H, x.dtype)
    dh = np.zeros((B, H), x.dtype)
    for t in range(T - 1, -1, -1):
        hp, r, z, n, hn = cache[t]
        d = dh + dy[t]
        dpn = d * (1.0 - z) * (1.0 - n * n)
        dpz = d * (hp - n) * z * (1.0 - z)
        dpr = dpn * hn * r * (1.0 - r)
        dgx = np.concatenate([dpr, dpz, dpn], axis=1)         # seen from the input side
        dgh = np.concatenate([dpr, dpz, dpn * r], axis=1)     # seen from the hidden side (b_hn sits inside r * (...))
        dx[t] = dgx @ w_ih
        dw_ih += dgx.T @ x[t]
        dw_hh += dgh.T @ hp
        db_ih += dgx.sum(0)
        db_hh += dgh.sum(0)
        dh = d * z + dgh @ w_hh
    return dx, dw_ih, dw_hh, db_ih, db_hh, dh
