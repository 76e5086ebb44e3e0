"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy) of the LSTM arithmetic on the hot path.

The reference (TUT-SLP-lab/MultimodalReactionGeneration) delegates every LSTM call to the
third-party ``torch.nn.LSTM`` (un-pinned in requirements.txt; Docker base
nvcr.io/nvidia/pytorch:23.04-py3; torch 2.11.0 in this image).  Call sites:
``mr_gen/model/utils/lstm_block.py:21,41``, ``mr_gen/model/utils/lstm_sampler.py:16,29``,
``mr_gen/model/utils/mixer_block.py:237,251``.  The algorithm restated here is the one
published in torch's own documentation, ``torch/nn/modules/rnn.py:842-847``:

    i = sigmoid(W_ii x + b_ii + W_hi h + b_hi)      gate order in the packed weights: i, f, g, o
    f = sigmoid(W_if x + b_if + W_hf h + b_hf)      (rnn.py:935-956 -> weight_ih_l{k} is [4H, I])
    g = tanh   (W_ig x + b_ig + W_hg h + b_hg)
    o = sigmoid(W_io x + b_io + W_ho h + b_ho)
    c' = f*c + i*g ;  h' = o*tanh(c')

Parity pinning: the reference holds NO tests / golden vectors for this path,
so this file is pinned against (a) ``torch.nn.LSTM`` itself executed on CPU
(tests/test_oracle_cpu.py) and (b) fixtures produced by importing the unmodified reference
in the build container (oracle/make_golden.py -> tests/golden/*.npz).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product path (multimodalreactiongeneration_b200) never does.
"""
from __future__ import annotations

import numpy as np


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_layer_forward(x, w_ih, w_hh, b_ih, b_hh, h0=None, c0=None, reverse=False):
    """One direction of one layer.  x is time-major [T, B, I].

    Returns y [T, B, H], (h_n, c_n) and the cache needed by ``lstm_layer_backward``:
    post-activation gates [T, B, 4H] (i,f,g,o) and cell states [T, B, H].
    """
    T, B, _ = x.shape
    H = w_hh.shape[1]
    dt = x.dtype
    h = np.zeros((B, H), dt) if h0 is None else h0.astype(dt)
    c = np.zeros((B, H), dt) if c0 is None else c0.astype(dt)
    bias = b_ih + b_hh if b_ih is not None else np.zeros(4 * H, dt)
    y = np.zeros((T, B, H), dt)
    gates = np.zeros((T, B, 4 * H), dt)
    cells = np.zeros((T, B, H), dt)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        pre = x[t] @ w_ih.T + h @ w_hh.T + bias
        i = _sigmoid(pre[:, 0 * H:1 * H])
        f = _sigmoid(pre[:, 1 * H:2 * H])
        g = np.tanh(pre[:, 2 * H:3 * H])
        o = _sigmoid(pre[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * np.tanh(c)
        y[t] = h
        cells[t] = c
        gates[t] = np.concatenate([i, f, g, o], axis=1)
    return y, (h, c), (gates, cells)


def lstm_layer_backward(dy, dh_n, dc_n, x, y, cache, w_ih, w_hh, h0=None, c0=None,
                        reverse=False):
    """BPTT for ``lstm_layer_forward``.  Returns dx, dw_ih, dw_hh, db, dh0, dc0.

    ``db`` is the gradient of both bias_ih and bias_hh (they enter as a sum).
    """
    gates, cells = cache
    T, B, _ = x.shape
    H = w_hh.shape[1]
    dt = x.dtype
    h0 = np.zeros((B, H), dt) if h0 is None else h0
    c0 = np.zeros((B, H), dt) if c0 is None else c0
    dh = np.zeros((B, H), dt) if dh_n is None else dh_n.astype(dt).copy()
    dc = np.zeros((B, H), dt) if dc_n is None else dc_n.astype(dt).copy()
    dx = np.zeros_like(x)
    dw_ih = np.zeros_like(w_ih)
    dw_hh = np.zeros_like(w_hh)
    db = np.zeros(4 * H, dt)
    order = list(range(T - 1, -1, -1) if reverse else range(T))
    for idx in range(T - 1, -1, -1):
        t = order[idx]
        t_prev = order[idx - 1] if idx > 0 else None
        i = gates[t][:, 0 * H:1 * H]
        f = gates[t][:, 1 * H:2 * H]
        g = gates[t][:, 2 * H:3 * H]
        o = gates[t][:, 3 * H:4 * H]
        c = cells[t]
        c_prev = cells[t_prev] if t_prev is not None else c0
        h_prev = y[t_prev] if t_prev is not None else h0
        dh_t = dh + dy[t]
        tc = np.tanh(c)
        do = dh_t * tc
        dc_t = dc + dh_t * o * (1.0 - tc * tc)
        di = dc_t * g
        dg = dc_t * i
        df = dc_t * c_prev
        dpre = np.concatenate(
            [di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1
        )
        dx[t] = dpre @ w_ih
        dw_ih += dpre.T @ x[t]
        dw_hh += dpre.T @ h_prev
        db += dpre.sum(axis=0)
        dh = dpre @ w_hh
        dc = dc_t * f
    return dx, dw_ih, dw_hh, db, dh, dc


def lstm_forward(x_bf, params, num_layers=1, bidirectional=False, hx=None):
    """Multi-layer / bidirectional stack, batch_first like every reference call site.

    ``params`` maps torch names (``weight_ih_l0``, ``weight_hh_l0_reverse``, ...) to arrays.
    Returns y [B, T, D*H], (h_n [L*D, B, H], c_n [L*D, B, H]).
    """
    x = np.swapaxes(x_bf, 0, 1)
    D = 2 if bidirectional else 1
    hs, cs = [], []
    for layer in range(num_layers):
        outs = []
        for d in range(D):
            sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
            k = layer * D + d
            h0 = None if hx is None else hx[0][k]
            c0 = None if hx is None else hx[1][k]
            y, (h, c), _ = lstm_layer_forward(
                x,
                params["weight_ih" + sfx],
                params["weight_hh" + sfx],
                params.get("bias_ih" + sfx),
                params.get("bias_hh" + sfx),
                h0,
                c0,
                reverse=(d == 1),
            )
            outs.append(y)
            hs.append(h)
            cs.append(c)
        x = np.concatenate(outs, axis=2) if D == 2 else outs[0]
    return np.swapaxes(x, 0, 1), (np.stack(hs), np.stack(cs))


def layer_norm(x, weight, bias, eps=1e-5):
    mu = x.mean(axis=-1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * weight + bias
