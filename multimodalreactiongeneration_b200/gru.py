"""``B200GRU`` — the drop-in for ``torch.nn.GRU`` at lstmformer's GRU mixer (reference: mr_gen/model/utils/
mixer_block.py:194,207 ``GRUMixer.mixer``, selected by ``emb_mixers: gru`` in mr_gen/model/lstmformer/config_gru.yaml).

Same constructor, parameter names / shapes / init order and ``state_dict`` as ``nn.GRU`` (``weight_ih_l0`` [3H, I],
gate order r, z, n); ``forward`` is replaced: the time-parallel ``x W_ih^T + b_ih`` and the four backward
contractions (dX, dW_ih, dW_hh and the two bias column sums) run on the library's tcgen05 3xTF32 GEMM / column-sum
kernels.  The recurrence runs
* for H in {128, 256} and T > 1 on the CLUSTER-RESIDENT recurrent kernels of the LSTM path (``rec_fwd2 / rec_bwd2_kernel<..,
  gru>``, flag ``MRG_F_GRU``): the GRU's three gates are laid out in the LSTM's four gate slots — W_ih rows (r, z, n, 0),
  W_hh rows (r, z, 0, n), bias (b_ir + b_hr, b_iz + b_hz, b_in, b_hn) — built here with differentiable torch ops, so one
  four-slot d(pre-activation) vector serves dX, both weight gradients and both bias gradients and autograd routes every
  block to the right ``nn.GRU`` parameter (the zero blocks receive the unused products);
* otherwise on the generic ``mrg_gru_forward`` / ``mrg_gru_backward`` (csrc/mrg_gru.cu: any hidden size, W_hh streamed
  from L2), composed per layer and direction below (a reverse direction runs on the time-flipped sequence).
No cuDNN, no CPU path."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import _cabi
from .linear import _colsum, _gemm, fused_grad_target
from .lstm import _F_INFER, _LSTMLayerFn, _default_flags


def _four_gate_weights(w_ih, w_hh, b_ih, b_hh, H):
    """nn.GRU parameters of one direction -> the four-gate form of MRG_F_GRU (include/mrg_lstm.h)."""
    w_ih4 = torch.cat([w_ih, w_ih.new_zeros(H, w_ih.shape[1])], dim=0)                 # (r, z, n, 0)
    w_hh4 = torch.cat([w_hh[:2 * H], w_hh.new_zeros(H, H), w_hh[2 * H:]], dim=0)       # (r, z, 0, n)
    if b_ih is None:
        return w_ih4, w_hh4, None, None
    b4 = torch.cat([b_ih[:2 * H] + b_hh[:2 * H], b_ih[2 * H:], b_hh[2 * H:]])          # (b_r, b_z, b_in, b_hn)
    return w_ih4, w_hh4, b4, None


class _GRULayerFn(torch.autograd.Function):
    """One direction of one layer, time-major: x [T, B, I], h0 [B, H] or None -> y [T, B, H] (h_n = y[-1])."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, h0):
        T, B, I = x.shape
        H = w_hh.shape[1]
        dev = x.device
        flags = _default_flags()
        x2 = x.contiguous().view(T * B, I)
        gx = torch.empty((T * B, 3 * H), dtype=torch.float32, device=dev)
        _gemm(x2, I, 1, w_ih.contiguous(), 1, I, b_ih, gx, T * B, 3 * H, I, flags)
        y_ext = torch.empty((T + 1, B, H), dtype=torch.float32, device=dev)
        if h0 is None:
            y_ext[0].zero_()
        else:
            y_ext[0].copy_(h0)
        train = any(ctx.needs_input_grad)   # grad mode is off inside Function.forward; this is the reliable signal
        reserve = torch.empty((T, B, 4, H), dtype=torch.float32, device=dev) if train else None
        whh = w_hh.contiguous()
        whh_t = whh.t().contiguous() if H % 4 == 0 else None   # [H, 3H]: the forward kernel streams its columns
        with torch.cuda.device(dev):
            st = _cabi.lib().mrg_gru_forward(gx.data_ptr(), whh.data_ptr(), _cabi.ptr(whh_t), _cabi.ptr(b_hh), y_ext.data_ptr(),
                                             _cabi.ptr(reserve), T, B, H, 1 if train else 0,
                                             torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "mrg_gru_forward")
        ctx.save_for_backward(x2, w_ih, whh, y_ext, reserve)
        ctx.params = (w_ih, w_hh, b_ih, b_hh)
        ctx.dims = (T, B, I, H, flags, h0 is not None)
        return y_ext[1:]

    @staticmethod
    def backward(ctx, dy):
        x2, w_ih, whh, y_ext, reserve = ctx.saved_tensors
        T, B, I, H, flags, has_h0 = ctx.dims
        dev = dy.device
        dy = dy.contiguous()
        dgx = torch.empty((T * B, 3 * H), dtype=torch.float32, device=dev)
        dgh = torch.empty((T * B, 3 * H), dtype=torch.float32, device=dev)
        dh0 = torch.empty((B, H), dtype=torch.float32, device=dev) if has_h0 else None
        with torch.cuda.device(dev):
            st = _cabi.lib().mrg_gru_backward(dy.data_ptr(), None, reserve.data_ptr(), y_ext.data_ptr(), whh.data_ptr(),
                                              dgx.data_ptr(), dgh.data_ptr(), _cabi.ptr(dh0), T, B, H,
                                              torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "mrg_gru_backward")
        M, G = T * B, 3 * H
        wp_ih, wp_hh, bp_ih, bp_hh = ctx.params
        dx = dw_ih = dw_hh = db_ih = db_hh = None
        if ctx.needs_input_grad[0]:   # dX[M, I] = dgx[M, 3H] . W_ih[3H, I]
            dx = torch.empty((M, I), dtype=torch.float32, device=dev)
            _gemm(dgx, G, 1, w_ih.contiguous(), I, 1, None, dx, M, I, G, flags)
            dx = dx.view(T, B, I)
        hprev = y_ext[:T].view(M, H)

        def wgrad(dg, act, K, param):   # dW[3H, K] = dg^T[3H, M] . act[M, K]
            tgt = fused_grad_target(param)
            if tgt is not None:
                _gemm(dg, 1, G, act, K, 1, None, tgt, G, K, M, flags, accumulate=1)
                return None
            dw = torch.empty((G, K), dtype=torch.float32, device=dev)
            _gemm(dg, 1, G, act, K, 1, None, dw, G, K, M, flags)
            return dw

        if ctx.needs_input_grad[1]:
            dw_ih = wgrad(dgx, x2, I, wp_ih)
        if ctx.needs_input_grad[2]:
            dw_hh = wgrad(dgh, hprev, H, wp_hh)
        if bp_ih is not None and ctx.needs_input_grad[3]:
            db_ih = _colsum(dgx, into=fused_grad_target(bp_ih))
        if bp_hh is not None and ctx.needs_input_grad[4]:
            db_hh = _colsum(dgh, into=fused_grad_target(bp_hh))
        return dx, dw_ih, dw_hh, db_ih, db_hh, dh0


class B200GRU(nn.GRU):
    def forward(self, input, hx: Optional[torch.Tensor] = None):  # noqa: A002 - nn.GRU's argument name
        if not input.is_cuda:
            raise RuntimeError("B200GRU has no CPU path: inputs must live on a B200 (sm_100a) device")
        if input.dtype != torch.float32:
            raise TypeError("B200GRU computes in fp32")
        if input.dim() != 3:
            raise ValueError(f"B200GRU: expected a 3-D input (batched sequences), got {input.dim()}-D")
        x = input.transpose(0, 1) if self.batch_first else input          # time-major [T, B, F]
        T, B = x.shape[0], x.shape[1]
        D = 2 if self.bidirectional else 1
        if hx is not None and tuple(hx.shape) != (self.num_layers * D, B, self.hidden_size):
            raise RuntimeError(f"Expected hidden size {(self.num_layers * D, B, self.hidden_size)}, got {list(hx.shape)}")
        if T == 0:
            raise RuntimeError("B200GRU: empty sequence")
        h_n = []
        Hs = self.hidden_size
        flags = _default_flags()
        cluster = Hs in (128, 256) and T > 1 and not flags & _cabi.F_GENERIC_REC
        infer = 0 if torch.is_grad_enabled() else _F_INFER   # no reserve for a backward under torch.no_grad()
        for layer in range(self.num_layers):
            if cluster:   # both directions in one launch on the cluster-resident kernels
                ws = []
                for d in range(D):
                    sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
                    ws += _four_gate_weights(getattr(self, "weight_ih" + sfx), getattr(self, "weight_hh" + sfx),
                                             getattr(self, "bias_ih" + sfx) if self.bias else None,
                                             getattr(self, "bias_hh" + sfx) if self.bias else None, Hs)
                h0 = None if hx is None else hx[layer * D:(layer + 1) * D]
                x, hl, _ = _LSTMLayerFn.apply(x, h0, None, D, Hs, (flags & ~_cabi.F_BF16) | _cabi.F_GRU | infer, *ws)
                h_n += [hl[d] for d in range(D)]
                if self.dropout > 0.0 and self.training and layer + 1 < self.num_layers:
                    x = F.dropout(x, self.dropout, True)
                continue
            outs = []
            for d in range(D):
                sfx = f"_l{layer}" + ("_reverse" if d == 1 else "")
                w_ih, w_hh = getattr(self, "weight_ih" + sfx), getattr(self, "weight_hh" + sfx)
                b_ih = getattr(self, "bias_ih" + sfx) if self.bias else None
                b_hh = getattr(self, "bias_hh" + sfx) if self.bias else None
                h0 = None if hx is None else hx[layer * D + d]
                xin = x.flip(0) if d == 1 else x
                y = _GRULayerFn.apply(xin, w_ih, w_hh, b_ih, b_hh, h0)
                h_n.append(y[-1])
                outs.append(y.flip(0) if d == 1 else y)
            x = outs[0] if D == 1 else torch.cat(outs, dim=-1)
            if self.dropout > 0.0 and self.training and layer + 1 < self.num_layers:
                x = F.dropout(x, self.dropout, True)
        out = x.transpose(0, 1) if self.batch_first else x
        return out, torch.stack(h_n, dim=0)
