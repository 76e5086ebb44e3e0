"""``rollout`` — the autoregressive predictor of ``LSTMwithSample`` as ONE persistent kernel per direction
(``mrg_rollout_forward`` / ``mrg_rollout_backward``, csrc/mrg_rollout.cu).

Reference: the Python time loop mr_gen/model/lstm_with_sampling/lstm_with_sample.py:379-408
(``head_motion_generation``) calling :410-433 (``generate_one_step`` -> ``forward`` with one frame).  Per step the
reference runs feature_projection -> ``LSTMLayerd`` from zero state (quirk Q2) -> bottleneck FFN and selects the next
step's "previous pose" between its own prediction and the ground truth (Q5 one-frame lag; Q6 no detach).  Here the
whole T loop stays on the device; autograd sees one node whose backward is the BPTT kernel followed by the
time-parallel weight-gradient GEMMs / column sums on the library's kernels."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import torch

from . import _cabi
from .linear import _colsum, _gemm, fused_grad_target
from .lstm import _default_flags


def supported(H: int, L: int, P: int, FB: int) -> bool:
    return bool(_cabi.lib().mrg_rollout_supported(int(H), int(L), int(P), int(FB)))


def _weights_struct(H, L, P, FB, relu, eps, w_prev, layers, w1, b1, w2, b2):
    w = _cabi.RolloutWeights()
    w.H, w.L, w.P, w.FB, w.relu, w.ln_eps = H, L, P, FB, int(relu), float(eps)
    w.w_prev, w.w_prev_ld = w_prev.data_ptr(), w_prev.stride(0)
    for l, (w_ih, _w_hh, b_ih, b_hh, g, b) in enumerate(layers):
        w.w_ih[l], w.b_ih[l], w.b_hh[l] = w_ih.data_ptr(), _cabi.ptr(b_ih), _cabi.ptr(b_hh)
        w.ln_g[l], w.ln_b[l] = g.data_ptr(), b.data_ptr()
    w.w1, w.b1, w.w2, w.b2 = w1.data_ptr(), _cabi.ptr(b1), w2.data_ptr(), _cabi.ptr(b2)
    return w


def _check(t: Optional[torch.Tensor], name: str, dev) -> None:
    if t is None:
        return
    if t.dtype != torch.float32 or t.device != dev:
        raise ValueError(f"rollout: {name} must be an fp32 tensor on {dev}")


def _into_or_return(grad: torch.Tensor, param: torch.Tensor) -> Optional[torch.Tensor]:
    """Add ``grad`` into the trainer's flat bucket when it owns ``param.grad`` (nothing for autograd to add then)."""
    tgt = fused_grad_target(param)
    if tgt is None:
        return grad
    tgt.add_(grad)
    return None


class _RolloutFn(torch.autograd.Function):
    """pred[T,B,P] = rollout(base[T,B,H], gt_prev[T,B,P], mask[T,B] u8 | None; weights).  Inputs of the node:
    base, gt_prev, w_prev (a view of feature_projection.weight[:, -P:]), w1, b1, w2, b2, then per layer
    (w_ih, w_hh, b_ih, b_hh, ln_weight, ln_bias).  ``w_hh`` takes no part in the arithmetic (zero state, quirk Q2) but is
    a node input so that it receives the exact-zero gradient the reference's graph gives it."""

    @staticmethod
    def forward(ctx, base, gt_prev, mask, relu, eps, w_prev, w1, b1, w2, b2, *layer_params):
        if not base.is_cuda:
            raise RuntimeError("rollout has no CPU path: tensors must live on a B200 (sm_100a) device")
        L = len(layer_params) // 6
        layers = [layer_params[6 * l:6 * l + 6] for l in range(L)]
        T, B, H = base.shape
        P, FB = w2.shape
        dev = base.device
        base, gt_prev = base.contiguous(), gt_prev.contiguous()
        for name, t in (("base", base), ("gt_prev", gt_prev), ("w1", w1), ("w2", w2), ("b1", b1), ("b2", b2)):
            _check(t, name, dev)
        if w_prev.stride(1) != 1:
            w_prev = w_prev.contiguous()
        w1c, w2c = w1.contiguous(), w2.contiguous()
        layers_c = [tuple(None if t is None else t.contiguous() for t in lay) for lay in layers]
        if mask is not None:
            mask = mask.to(device=dev, dtype=torch.uint8).contiguous()
            if tuple(mask.shape) != (T, B):
                raise ValueError(f"rollout: mask must be [T, B] = {(T, B)}, got {tuple(mask.shape)}")
        need_grad = any(ctx.needs_input_grad)
        opts = dict(dtype=torch.float32, device=dev)
        pred = torch.empty((T, B, P), **opts)
        reserve = None
        keep = None
        if need_grad:
            keep = dict(xs=torch.empty((L + 1, T, B, H), **opts), gates=torch.empty((L, T, B, 3, H), **opts),
                        xhat=torch.empty((L, T, B, H), **opts), rstd=torch.empty((L, T, B), **opts),
                        fact=torch.empty((T, B, FB), **opts), prev=torch.empty((T, B, P), **opts))
            reserve = _cabi.RolloutReserve(*(keep[k].data_ptr() for k in ("xs", "gates", "xhat", "rstd", "fact", "prev")))
        w = _weights_struct(H, L, P, FB, relu, eps, w_prev, layers_c, w1c, b1, w2c, b2)
        with torch.cuda.device(dev):
            st = _cabi.lib().mrg_rollout_forward(base.data_ptr(), gt_prev.data_ptr(), _cabi.ptr(mask), ctypes.byref(w),
                                                 pred.data_ptr(), None if reserve is None else ctypes.byref(reserve),
                                                 T, B, torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "mrg_rollout_forward")
        if need_grad:
            ctx.keep = keep
            ctx.mask = mask
            ctx.cfg = (T, B, H, L, P, FB, relu, eps)
            ctx.tensors = (w_prev, w1c, b1, w2c, b2, layers_c)
            ctx.params = (w1, b1, w2, b2, layers)   # the python objects (fused gradient targets hang off them)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        T, B, H, L, P, FB, relu, eps = ctx.cfg
        keep, mask = ctx.keep, ctx.mask
        w_prev, w1c, b1, w2c, b2, layers_c = ctx.tensors
        w1, b1p, w2, b2p, layers = ctx.params
        dev = dpred.device
        opts = dict(dtype=torch.float32, device=dev)
        dpred = dpred.contiguous()
        g = dict(dy=torch.empty((T, B, P), **opts), df=torch.empty((T, B, FB), **opts),
                 dpre=torch.empty((L, T, B, 4 * H), **opts), dbase=torch.empty((T, B, H), **opts),
                 dprev=torch.empty((T, B, P), **opts), dln_g=torch.empty((L, B, H), **opts),
                 dln_b=torch.empty((L, B, H), **opts))
        grads = _cabi.RolloutGrads(*(g[k].data_ptr() for k in ("dy", "df", "dpre", "dbase", "dprev", "dln_g", "dln_b")))
        reserve = _cabi.RolloutReserve(*(keep[k].data_ptr() for k in ("xs", "gates", "xhat", "rstd", "fact", "prev")))
        w = _weights_struct(H, L, P, FB, relu, eps, w_prev, layers_c, w1c, b1, w2c, b2)
        with torch.cuda.device(dev):
            st = _cabi.lib().mrg_rollout_backward(dpred.data_ptr(), _cabi.ptr(mask), ctypes.byref(w),
                                                  ctypes.byref(reserve), ctypes.byref(grads), T, B,
                                                  torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "mrg_rollout_backward")
        flags = _default_flags()
        M = T * B

        def wgrad(dout2, x2, param):
            """d(param)[N, K] = dout2[M, N]^T x2[M, K]; straight into the trainer's flat bucket when it owns .grad."""
            N, K = dout2.shape[1], x2.shape[1]
            tgt = fused_grad_target(param) if param is not None else None
            if tgt is not None:
                _gemm(dout2, 1, N, x2, K, 1, None, tgt, N, K, M, flags, accumulate=1)
                return None
            out = torch.empty((N, K), **opts)
            _gemm(dout2, 1, N, x2, K, 1, None, out, N, K, M, flags)
            return out

        def bgrad(dout2, param):
            if param is None:
                return None
            tgt = fused_grad_target(param)
            return _colsum(dout2, into=tgt)

        ni = ctx.needs_input_grad
        d_base = g["dbase"] if ni[0] else None
        d_gt = None
        if ni[1]:   # ground truth entered step t (> 0) only where step t-1 was NOT fed back
            d_gt = g["dprev"].clone()
            if mask is not None and T > 1:
                d_gt[1:] = d_gt[1:] * (mask[:-1] == 0).unsqueeze(-1)
        d_wprev = wgrad(g["dbase"].view(M, H), keep["prev"].view(M, P), None) if ni[5] else None
        dy2, df2 = g["dy"].view(M, P), g["df"].view(M, FB)
        d_w1 = wgrad(df2, keep["xs"][L].view(M, H), w1) if ni[6] else None
        d_b1 = bgrad(df2, b1p) if (b1p is not None and ni[7]) else None
        d_w2 = wgrad(dy2, keep["fact"].view(M, FB), w2) if ni[8] else None
        d_b2 = bgrad(dy2, b2p) if (b2p is not None and ni[9]) else None
        out: List[Optional[torch.Tensor]] = []
        for l in range(L):
            w_ih, w_hh, b_ih, b_hh, ln_g, ln_b = layers[l]
            dpre2 = g["dpre"][l].view(M, 4 * H)
            k = 10 + 6 * l
            d_wih = wgrad(dpre2, keep["xs"][l].view(M, H), w_ih) if ni[k] else None
            d_whh = None   # h_0 = 0 at every step: exactly zero (the flat bucket already holds zeros)
            if w_hh is not None and ni[k + 1] and fused_grad_target(w_hh) is None:
                d_whh = torch.zeros_like(w_hh)
            db = None
            if (b_ih is not None and ni[k + 2]) or (b_hh is not None and ni[k + 3]):
                db = _colsum(dpre2)
            d_bih = d_bhh = None
            if b_ih is not None and ni[k + 2]:
                d_bih = _into_or_return(db, b_ih)
            if b_hh is not None and ni[k + 3]:
                d_bhh = _into_or_return(db, b_hh)
            d_g = bgrad(g["dln_g"][l], ln_g) if ni[k + 4] else None
            d_b = bgrad(g["dln_b"][l], ln_b) if ni[k + 5] else None
            out += [d_wih, d_whh, d_bih, d_bhh, d_g, d_b]
        ctx.keep = None
        return (d_base, d_gt, None, None, None, d_wprev, d_w1, d_b1, d_w2, d_b2, *out)


def rollout(base: torch.Tensor, gt_prev: torch.Tensor, mask: Optional[torch.Tensor], w_prev: torch.Tensor,
            layers: Sequence[Sequence[Optional[torch.Tensor]]], w1, b1, w2, b2, relu: bool = True,
            eps: float = 1e-5) -> torch.Tensor:
    """Time-major rollout: base [T,B,H], gt_prev [T,B,P], mask [T,B] (bool / uint8; mask[t] feeds pred[t] to step
    t+1) or None, layers = [(w_ih [4H,H], w_hh | None, b_ih, b_hh, ln_weight, ln_bias), ...] -> pred [T,B,P]."""
    flat = [t for lay in layers for t in lay]
    return _RolloutFn.apply(base, gt_prev, mask, bool(relu), float(eps), w_prev, w1, b1, w2, b2, *flat)
