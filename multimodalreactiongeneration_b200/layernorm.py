"""Fused ``LayerNorm(y + x)`` (the tail of the reference's ResidualConnection,
mr_gen/model/utils/residual_connection.py:29-32) through ``mrg_residual_layernorm_forward/backward``.
Reads the time-major LSTM output and the batch-first block input in place (per-tensor row strides)."""
from __future__ import annotations

import torch

from . import _cabi
from .lstm import _workspace

SUPPORTED_H = (128, 256, 512)


def _rows(t: torch.Tensor):
    """3-D tensor [n0, n1, H] with a contiguous last dimension -> (s0, s1) element strides."""
    return t.stride(0), t.stride(1)


def supported(y: torch.Tensor, x: torch.Tensor) -> bool:
    return (y.is_cuda and y.dtype == torch.float32 and x.dtype == torch.float32 and y.dim() == 3
            and y.shape == x.shape and y.shape[-1] in SUPPORTED_H and y.numel() > 0)


def _dense3(t: torch.Tensor) -> torch.Tensor:
    ok = t.stride(-1) == 1 and t.stride(0) % 4 == 0 and t.stride(1) % 4 == 0 and t.data_ptr() % 16 == 0
    if ok and (t.is_contiguous() or t.transpose(0, 1).is_contiguous()):
        return t
    return _cabi.contiguous3(t)


class _ResidualLNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, x, gamma, beta, eps):
        L = _cabi.lib()
        y, x = _dense3(y), _dense3(x)
        n0, n1, H = y.shape
        out = torch.empty_like(y)  # keeps y's (possibly time-major) memory order
        mean = torch.empty(n0 * n1, dtype=torch.float32, device=y.device)
        rstd = torch.empty_like(mean)
        stream = torch.cuda.current_stream(y.device).cuda_stream
        with torch.cuda.device(y.device):
            st = L.mrg_residual_layernorm_forward(y.data_ptr(), *_rows(y), x.data_ptr(), *_rows(x), gamma.data_ptr(),
                                                  beta.data_ptr(), out.data_ptr(), *_rows(out), mean.data_ptr(),
                                                  rstd.data_ptr(), n0, n1, H, float(eps), stream)
        _cabi.check(st, "mrg_residual_layernorm_forward")
        ctx.save_for_backward(y, x, gamma, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _cabi.lib()
        y, x, gamma, mean, rstd = ctx.saved_tensors
        n0, n1, H = y.shape
        dout = _dense3(dout)
        dsum = torch.empty_like(y)
        dgamma = torch.empty_like(gamma)
        dbeta = torch.empty_like(gamma)
        ws = _workspace(y.device, L.mrg_layernorm_workspace_bytes(H))
        stream = torch.cuda.current_stream(y.device).cuda_stream
        with torch.cuda.device(y.device):
            st = L.mrg_residual_layernorm_backward(dout.data_ptr(), *_rows(dout), y.data_ptr(), *_rows(y),
                                                   x.data_ptr(), *_rows(x), gamma.data_ptr(), mean.data_ptr(),
                                                   rstd.data_ptr(), dsum.data_ptr(), *_rows(dsum), dgamma.data_ptr(),
                                                   dbeta.data_ptr(), ws.data_ptr(), ws.numel(), n0, n1, H, stream)
        _cabi.check(st, "mrg_residual_layernorm_backward")
        return dsum, dsum, dgamma, dbeta, None


def residual_layer_norm(y, x, gamma, beta, eps=1e-5):
    return _ResidualLNFn.apply(y, x, gamma, beta, eps)
