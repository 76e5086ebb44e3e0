"""``B200Linear`` — ``nn.Linear`` whose three GEMMs (y = xWᵀ+b, dx = dy·W, dW = dyᵀ·x) run on the library's
tcgen05 3xTF32 kernel (fp32-grade accuracy) instead of torch's fp32 SIMT sgemm.

Used for the Linear layers that sit directly on the ``[B, T, 256]`` stream next to the LSTMs (SURVEY.md §8a
rows a1/a2/a5: ``LSTMModule.mixer``, the bottleneck FFN of ``LSTMBlock``, ``embed_layer`` /
``acoustic_projection`` / ``feature_projection`` / ``feed_forward``).  Same parameters, init and
``state_dict`` keys as ``nn.Linear``; no CPU path."""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi
from .lstm import _default_flags, _workspace


def _gemm(a, a_sm, a_sk, b, b_sk, b_sn, bias, c, M, N, K, flags, accumulate=0):
    L = _cabi.lib()
    dev = c.device
    ws = _workspace(dev, L.mrg_gemm_workspace_bytes(M, N, K))
    with torch.cuda.device(dev):
        st = L.mrg_gemm_strided(a.data_ptr(), a_sm, a_sk, b.data_ptr(), b_sk, b_sn, _cabi.ptr(bias), c.data_ptr(),
                                N, M, N, K, accumulate, 0, ws.data_ptr(), ws.numel(), flags,
                                torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(st, "mrg_gemm_strided")


def _split_ok(M, N, K, a_sm, a_sk, b_sk, b_sn, flags):
    """Does the persistent pre-split-weight kernel (csrc/mrg_gemm_tc4.cu) cover this GEMM?"""
    if flags & _cabi.F_SIMT_GEMM:
        return False
    return bool(_cabi.lib().mrg_gemm_split_supported(M, N, K, a_sm, a_sk, b_sk, b_sn, N))


def _split_weight(w):
    """[2, *w.shape]: tf32 hi part and lo = w - hi of a weight matrix (one small launch per forward; the planes are
    reused by the backward's dX GEMM)."""
    L = _cabi.lib()
    hl = torch.empty((2,) + tuple(w.shape), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        st = L.mrg_split_tf32(w.data_ptr(), hl[0].data_ptr(), hl[1].data_ptr(), w.numel(),
                              torch.cuda.current_stream(w.device).cuda_stream)
    _cabi.check(st, "mrg_split_tf32")
    return hl


def _gemm_split(a, a_sm, a_sk, hl, b_sk, b_sn, bias, c, M, N, K, flags):
    L = _cabi.lib()
    dev = c.device
    with torch.cuda.device(dev):
        st = L.mrg_gemm_strided_split(a.data_ptr(), a_sm, a_sk, hl[0].data_ptr(), hl[1].data_ptr(), b_sk, b_sn,
                                      _cabi.ptr(bias), c.data_ptr(), N, M, N, K, 0, None, 0, flags,
                                      torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(st, "mrg_gemm_strided_split")


def _colsum(x2, into=None):
    """Column sums of a contiguous [M, N] fp32 matrix (bias gradient) through the C-ABI (``mrg_colsum``).
    ``into``: accumulate into this [N] tensor instead of returning a new one."""
    M, N = x2.shape
    L = _cabi.lib()
    dev = x2.device
    out = torch.empty(N, dtype=torch.float32, device=dev) if into is None else into
    ws = _workspace(dev, L.mrg_colsum_workspace_bytes(M, N))
    with torch.cuda.device(dev):
        st = L.mrg_colsum(x2.data_ptr(), out.data_ptr(), M, N, 0 if into is None else 1, ws.data_ptr(), ws.numel(),
                          torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(st, "mrg_colsum")
    return out if into is None else None


def fused_grad_target(p):
    """The trainer's flat gradient bucket owns ``p.grad`` and is cleared by the optimizer kernel: weight-gradient
    kernels may then ADD straight into it (and return no gradient to autograd) instead of writing a temporary that
    autograd adds with one more launch per parameter."""
    if (getattr(p, "_mrg_grad_fused", False) and p.requires_grad and p.grad is not None
            and p.grad.is_contiguous()):
        # autograd never sees this write, so it cannot order it: remember the stream it is queued on and let
        # ``lstm.join_wgrad_streams`` (the trainer, before the all-reduce / optimizer) wait for it explicitly
        from .lstm import note_fused_write
        note_fused_write(p.grad.device)
        return p.grad
    return None


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        if not x.is_cuda:
            raise RuntimeError("B200Linear has no CPU path: inputs must live on a B200 (sm_100a) device")
        if x.dtype != torch.float32 or weight.dtype != torch.float32:
            raise TypeError("B200Linear computes in fp32")
        N, K = weight.shape
        x2 = (_cabi.contiguous3(x) if x.dim() == 3 else x).reshape(-1, K).contiguous()
        M = x2.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        flags = _default_flags()
        w = weight.contiguous()
        aligned = w.data_ptr() % 16 == 0 and x2.data_ptr() % 16 == 0 and (bias is None or bias.data_ptr() % 16 == 0)
        fwd_split = aligned and M > 0 and _split_ok(M, N, K, K, 1, 1, K, flags)
        bwd_split = aligned and M > 0 and ctx.needs_input_grad[0] and _split_ok(M, K, N, N, 1, K, 1, flags)
        hl = _split_weight(w) if (fwd_split or bwd_split) else None
        if fwd_split:
            _gemm_split(x2, K, 1, hl, 1, K, bias, y, M, N, K, flags)
        elif M > 0:
            _gemm(x2, K, 1, w, 1, K, bias, y, M, N, K, flags)
        ctx.split = hl if bwd_split else None
        ctx.save_for_backward(x2, weight)
        ctx.params = (weight, bias)   # the python objects (saved_tensors may hand back fresh wrappers)
        ctx.has_bias = bias is not None
        ctx.flags = flags
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        N, K = weight.shape
        M = x2.shape[0]
        dy2 = (_cabi.contiguous3(dy) if dy.dim() == 3 else dy).reshape(-1, N).contiguous()
        dx = dw = db = None
        w = weight.contiguous()
        if ctx.needs_input_grad[0]:
            dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
            if M > 0 and ctx.split is not None and dy2.data_ptr() % 16 == 0:   # dx[M,K] = dy[M,N] · W[N,K]
                _gemm_split(dy2, N, 1, ctx.split, K, 1, None, dx, M, K, N, ctx.flags)
            elif M > 0:
                _gemm(dy2, N, 1, w, K, 1, None, dx, M, K, N, ctx.flags)
            dx = dx.view(*dy.shape[:-1], K)
        wp, bp = ctx.params
        # Only dX feeds the layer below.  Inside the trainer's backward (lstm.wgrad_overlap) the weight / bias gradients
        # that go straight into the flat bucket are queued on the side stream of the weight-gradient GEMMs, where they run
        # beside the next kernels of the chain instead of delaying them; the trainer joins that stream before the
        # all-reduce / optimizer (lstm.join_wgrad_streams).
        from . import lstm as _lstm
        w_tgt = fused_grad_target(wp) if ctx.needs_input_grad[1] else None
        b_tgt = fused_grad_target(bp) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        if (_lstm._WGRAD["on"] and M > 0 and w_tgt is not None
                and (b_tgt is not None or not (ctx.has_bias and ctx.needs_input_grad[2]))):
            dev = dy.device
            main = torch.cuda.current_stream(dev)
            side = _lstm._wgrad_stream(dev)
            side.wait_stream(main)
            with torch.cuda.device(dev), torch.cuda.stream(side):
                _gemm(dy2, 1, N, x2, K, 1, None, w_tgt, N, K, M, ctx.flags, accumulate=1)
                if b_tgt is not None:
                    _colsum(dy2, into=b_tgt)
            for t in (dy2, x2):   # the allocator must not hand these out again before the side stream is done
                t.record_stream(side)
            # ... and autograd must not ACCUMULATE IN PLACE into the gradient this view aliases (it does when the storage
            # has no other owner — e.g. the residual branch adding its gradient to the same tensor): keep the view alive
            # until the trainer has joined the side stream
            _lstm._WGRAD["keepalive"].append(dy2)
            if all(side != s_ for s_ in _lstm._WGRAD["pending"]):
                _lstm._WGRAD["pending"].append(side)
            return dx, None, None
        if ctx.needs_input_grad[1]:
            tgt = w_tgt
            if tgt is not None:   # accumulate into the flat bucket, nothing for autograd to add
                if M > 0:
                    _gemm(dy2, 1, N, x2, K, 1, None, tgt, N, K, M, ctx.flags, accumulate=1)
            else:
                dw = torch.zeros((N, K), dtype=torch.float32, device=dy.device) if M == 0 else \
                    torch.empty((N, K), dtype=torch.float32, device=dy.device)
                if M > 0:   # dW[N,K] = dyᵀ[N,M] · x[M,K]
                    _gemm(dy2, 1, N, x2, K, 1, None, dw, N, K, M, ctx.flags)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy2, into=b_tgt)
        return dx, dw, db


class B200Linear(nn.Linear):
    def forward(self, input):  # noqa: A002 - nn.Linear's argument name
        return _LinearFn.apply(input, self.weight, self.bias)
