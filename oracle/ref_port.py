"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch, functional) of the reference's model code.

Where ``lstm_numpy.py`` restates the LSTM cell, this file restates the *callers* of the hot
path as plain functions over a reference-layout ``state_dict``, with the
LSTM arithmetic delegated to the same third-party ``torch.nn.LSTM`` (CPU / oneDNN) that the
reference itself calls.  Each function cites the reference lines it follows.  Quirks Q1-Q7 of
SURVEY.md Appendix C are reproduced on purpose.

It is pinned by tests/golden/*.npz, produced by oracle/make_golden.py from the *unmodified*
reference imported in the build container.  It is the checker for the CUDA path and the timed
CPU baseline ("port") in bench.py; the product path never imports it.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn
from torch.func import functional_call

PADDING_VALUE = -100  # mr_gen/model/utils/values.py:2

_LSTM_CACHE: Dict[tuple, nn.LSTM] = {}


def _sub(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def lstm(sd, prefix, x, hx=None):
    """``nn.LSTM(..., batch_first=True)(x, hx)`` with weights taken from ``sd[prefix + name]``.

    Shape facts are recovered from the tensors: weight_hh_l0 is [4H, H], weight_ih_l0 [4H, I]
    (torch/nn/modules/rnn.py:935-956)."""
    p = _sub(sd, prefix)
    H = p["weight_hh_l0"].shape[1]
    I = p["weight_ih_l0"].shape[1]
    bi = "weight_ih_l0_reverse" in p
    L = 1 + max(int(k.split("_l")[1].split("_")[0]) for k in p if k.startswith("weight_ih_l"))
    key = (I, H, L, bi, p["weight_ih_l0"].dtype)
    if key not in _LSTM_CACHE:
        _LSTM_CACHE[key] = nn.LSTM(I, H, L, batch_first=True, bidirectional=bi).to(key[-1])
    return functional_call(_LSTM_CACHE[key], p, (x, hx))


def linear(sd, prefix, x):
    return F.linear(x, sd[prefix + "weight"], sd.get(prefix + "bias"))


def residual_ln(sd, prefix, y, x):
    """ResidualConnection tail: LN(y + x)  (mr_gen/model/utils/residual_connection.py:29-32)."""
    y = y + x
    if prefix + "layer_norm.weight" in sd:
        y = F.layer_norm(y, (y.shape[-1],), sd[prefix + "layer_norm.weight"],
                         sd[prefix + "layer_norm.bias"])
    return y


# --------------------------------------------------------------------------------------
# lstm_block.py
# --------------------------------------------------------------------------------------
def lstm_module(sd, prefix, x, hx=None):
    """LSTMModule.forward (lstm_block.py:38-46): LSTM then optional ``mixer`` Linear."""
    hs, hx = lstm(sd, prefix + "lstm_module.", x, hx)
    if prefix + "mixer.weight" in sd:
        hs = linear(sd, prefix + "mixer.", hs)
    return hs, hx


def lstm_block(sd, prefix, x, hx=None):
    """LSTMBlock.forward (lstm_block.py:101-107), residual form when the ckpt has `.module.`."""
    residual = any(k.startswith(prefix + "lstm_module.module.") for k in sd)
    if residual:
        y, hx = lstm_module(sd, prefix + "lstm_module.module.", x, hx)
        y = residual_ln(sd, prefix + "lstm_module.", y, x)
    else:
        y, hx = lstm_module(sd, prefix + "lstm_module.", x, hx)
    ff = prefix + "feed_forward_module."
    if any(k.startswith(ff) for k in sd):
        inner = ff + "module." if any(k.startswith(ff + "module.") for k in sd) else ff
        z = linear(sd, inner + "input.", y)
        z = torch.relu(z)  # use_relu=True at every reference call site
        z = linear(sd, inner + "mapping.", z)
        y = residual_ln(sd, ff, z, y) if inner != ff else z
    return y, hx


def lstm_layerd(sd, prefix, x, hxs=None):
    """LSTMLayerd.forward (lstm_block.py:159-169).  Returns the INPUT ``hxs`` (quirk Q2)."""
    i = 0
    while any(k.startswith(f"{prefix}lstm_layered.{i}.") for k in sd):
        hx = None if hxs is None else hxs[i]
        x, _ = lstm_block(sd, f"{prefix}lstm_layered.{i}.", x, hx)
        i += 1
    return x, hxs


def lstm_sampler(sd, prefix, x, ratio, hx=None):
    """LSTMSampler.forward (lstm_sampler.py:26-34): keep every ``ratio``-th output."""
    h, hx = lstm(sd, prefix + "sampler.", x, hx)
    return h[:, ratio - 1::ratio, :].contiguous(), hx


# --------------------------------------------------------------------------------------
# lstm_with_sample.py
# --------------------------------------------------------------------------------------
def lws_forward(sd, ratio, acoustic, motion_p, motion_s, lead_a, lead_p, lead_s, cell_state=None):
    """LSTMwithSample.forward (lstm_with_sample.py:151-232), tensors only (lengths unused)."""
    hx_sampler, hxs = (None, None) if cell_state is None else cell_state
    a = torch.cat([lead_a, acoustic], dim=1)
    mp = torch.cat([lead_p, motion_p], dim=1)
    ms = torch.cat([lead_s, motion_s], dim=1)
    a = linear(sd, "acoustic_projection.", a)
    a, hx_sampler = lstm_sampler(sd, "sampling_lstm.", a, ratio, hx_sampler)
    if not (a.shape[1] == mp.shape[1] == ms.shape[1]):
        raise RuntimeError("length mismatch")  # :203-209
    feat = linear(sd, "feature_projection.", torch.cat([a, mp, ms], dim=-1))
    h, hxs = lstm_layerd(sd, "layerd_lstm.", feat, hxs)
    y = linear(sd, "feed_forward.mapping.", torch.relu(linear(sd, "feed_forward.input.", h)))
    return y, lead_p.shape[1], (hx_sampler, hxs)


def masked_loss(y, target, loss_type="huber", delta=1.0, beta=1.0):
    """training_step tail (lstm_with_sample.py:288-296) with delta_loss_scale == 1 (Q7)."""
    m = (target != PADDING_VALUE).int()
    y, target = y * m, target * m
    if loss_type == "mse":
        return F.mse_loss(y, target)
    if loss_type == "mae":
        return F.l1_loss(y, target)
    if loss_type == "huber":
        return F.huber_loss(y, target, delta=delta)
    if loss_type == "smoothl1":
        return F.smooth_l1_loss(y, target, beta=beta)
    raise ValueError("invalid loss type")


def lws_training_step(sd, ratio, batch, loss_type="huber"):
    """Teacher-forced branch of training_step (lstm_with_sample.py:283-296)."""
    y, lead_len, _ = lws_forward(sd, ratio, *batch[:6])
    return masked_loss(y[:, lead_len:], batch[6], loss_type)


def lws_rollout(sd, ratio, batch, mask):
    """prediction() (lstm_with_sample.py:339-408) for a GIVEN mask.

    ``mask`` is bool [T] (reference shape, Q4) or [T, B] (per-sample extension).  True feeds the
    prediction back, False feeds ``motion_s[step]`` (one-frame lag, Q5).  The predictor LSTMs
    restart from zero state every step (Q2); only the sampler state is carried."""
    acoustic, motion_p, motion_s, lead_a, lead_p, lead_s = batch[:6]
    B, T, _ = motion_p.shape
    fb = acoustic.view(B, T, ratio, acoustic.shape[-1]).transpose(0, 1)
    empty = lambda t: t.new_empty((t.shape[0], 0, t.shape[2]))
    # warm-up on the leading segment (:374-377)
    _, _, state = lws_forward(sd, ratio, empty(acoustic), empty(motion_p), empty(motion_s),
                              lead_a, lead_p, lead_s, None)
    y = motion_s[:, 0:1]
    preds = []
    for step in range(T):
        y, _, state = lws_forward(
            sd, ratio, fb[step], motion_p[:, step:step + 1], y,
            empty(lead_a), empty(lead_p), empty(lead_s), state)
        preds.append(y)
        m = mask[step]
        gt = motion_s[:, step:step + 1]
        if torch.is_tensor(m) and m.dim() > 0:
            y = torch.where(m.view(B, 1, 1), y, gt)
        else:
            y = y if bool(m) else gt
    return torch.cat(preds, dim=1)


# --------------------------------------------------------------------------------------
# simple_lstm.py (with the documented Q1 unwrap)
# --------------------------------------------------------------------------------------
def mha_block(sd, prefix, q, kv, heads):
    """MultimodalAttentionBlock (multi_modal_att.py:22-31,58-59) — residual + LN variant."""
    p = prefix + "att_module.module."
    att, _ = F.multi_head_attention_forward(
        q.transpose(0, 1), kv.transpose(0, 1), kv.transpose(0, 1), q.shape[-1], heads,
        sd[p + "cross_modal_att.in_proj_weight"], sd[p + "cross_modal_att.in_proj_bias"],
        None, None, False, 0.0,
        sd[p + "cross_modal_att.out_proj.weight"], sd[p + "cross_modal_att.out_proj.bias"],
        training=False, need_weights=False)
    y = linear(sd, p + "projection.", att.transpose(0, 1))
    return residual_ln(sd, prefix + "att_module.", y, q)


def simple_lstm_forward(sd, acoustic, motion, heads=8):
    """SimpleLSTM.forward (simple_lstm.py:181-188) with LSTMLayerd outputs unwrapped (Q1)."""
    a, _ = lstm_layerd(sd, "acoustic_encoder.acostic_lstm.",
                       linear(sd, "acoustic_encoder.embed_layer.", acoustic))
    m, _ = lstm_layerd(sd, "motion_encoder.motion_lstm.",
                       linear(sd, "motion_encoder.embed_layer.", motion))
    i = 0
    while any(k.startswith(f"multimodal_att.att_layers.{i}.") for k in sd):
        m = mha_block(sd, f"multimodal_att.att_layers.{i}.", m, a, heads)
        i += 1
    d, _ = lstm_layerd(sd, "motion_decoder.decoder_lstm.", m)
    d = d[:, -1:, :]  # seq_reshape :127-138
    return linear(sd, "motion_decoder.mapping.output.",
                  torch.relu(linear(sd, "motion_decoder.mapping.input.", d)))


def simple_lstm_training_step(sd, batch, heads=8):
    """training_step (simple_lstm.py:239-255) with all_static False, delta_loss_scale 1."""
    acoustic, motion, target = batch
    return F.mse_loss(simple_lstm_forward(sd, acoustic, motion, heads), target)
