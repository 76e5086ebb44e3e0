"""``B200MultiheadAttention`` — ``nn.MultiheadAttention`` on the library's kernels: the q / k / v / output
PROJECTIONS (the dense ``[B*T, E] x [E, E]`` contractions) run on the tcgen05 3xTF32 GEMM, and
``softmax(QK^T / sqrt(d)) V`` runs on the fused fp32 attention kernels (``mrg_attention_forward/backward``,
csrc/mrg_attention.cu) — scores never reach HBM, heads are addressed in place in the
projection outputs (no transposes), and lstmformer's causal-rectangular + padding mask is evaluated as a function
(``AttentionMaskSpec``) instead of being read from a ``[B*heads, L, S]`` tensor.

Reference call sites: mr_gen/model/utils/multi_modal_att.py:12-31 (``batch_first=True``, ``need_weights=False``,
no masks) and mr_gen/model/utils/for_sequential.py:25-50 (lstmformer's integrators: ``batch_first=True``, a bool
``attn_mask`` of shape ``[B*heads, L, S]`` from ``gen_attention_mask``, True = masked out).  Parameters, their
names / shapes / init order and ``state_dict`` keys are ``nn.MultiheadAttention``'s (``in_proj_weight`` or
``q_/k_/v_proj_weight``, ``in_proj_bias``, ``out_proj.*``).

Which path runs: ONE — the fused kernels.  Head dims other than 32 / 64 are zero-padded per head to the next built
width (zero columns change neither the scores nor the outputs; the scale stays 1/sqrt(true head_dim)), mask = None
or an ``AttentionMaskSpec``.  There is no dispatch to a torch library kernel: an explicit mask TENSOR, attention
dropout in training mode, head_dim > 64, and the argument combinations the reference does not use on this path
(attention weights requested, key padding masks, ``bias_k`` / ``add_zero_attn``, time-major layout, CPU tensors)
raise ``NotImplementedError`` / ``RuntimeError``."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from . import _cabi
from .linear import _LinearFn


class AttentionMaskSpec:
    """The mask of mr_gen/model/utils/multi_modal_metaformer.py:32-79 as a rule: query i of ``L`` frames may see key
    j of ``S`` frames iff ``j // rate <= i`` (``mode`` 1, S = rate*L) or ``j <= i // rate`` (``mode`` 2, L = rate*S),
    and not (query frame padded AND key frame padded).  ``pad_q`` [B, L] / ``pad_k`` [B, S]: uint8, 1 = padded."""

    def __init__(self, mode: int, rate: int, pad_q: torch.Tensor, pad_k: torch.Tensor):
        self.mode, self.rate, self.pad_q, self.pad_k = mode, rate, pad_q, pad_k

    def materialize(self, head_num: int) -> torch.Tensor:
        """bool [B, head_num, L, S] (broadcast view), True = masked — what ``gen_attention_mask`` returns."""
        L, S = self.pad_q.shape[1], self.pad_k.shape[1]
        dev = self.pad_q.device
        q = torch.arange(L, device=dev).view(L, 1)
        k = torch.arange(S, device=dev).view(1, S)
        if self.mode == 1:
            causal = torch.div(k, self.rate, rounding_mode="floor") > q
        else:
            causal = k > torch.div(q, self.rate, rounding_mode="floor")
        both = self.pad_q.bool().unsqueeze(-1) & self.pad_k.bool().unsqueeze(1)
        return (causal.unsqueeze(0) | both).unsqueeze(1).expand(self.pad_q.shape[0], head_num, L, S)


def _rows(t: torch.Tensor):
    """[B, T, C] view whose rows are contiguous and whose batch stride is T * row stride -> (tensor, row stride)."""
    if t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1) or t.stride(1) % 4 or t.data_ptr() % 16:
        t = _cabi.contiguous3(t) if t.dim() == 3 else t.contiguous()
    return t, t.stride(1)


class _AttentionFn(torch.autograd.Function):
    """o[B, Tq, E] = softmax(scale * q k^T + mask) v per head; ``v is None``: ``k`` is a fused [B, Tk, 2E] k|v tensor
    (one projection GEMM) and its gradient comes back fused as well."""

    @staticmethod
    def forward(ctx, q, k, v, nh, spec, scale=None):
        B, Tq, E = q.shape
        fused = v is None
        q, ldq = _rows(q)
        kv, ldk = _rows(k)
        if fused:
            kk, vv, ldv = kv, kv[..., E:], ldk
        else:
            kk = kv
            vv, ldv = _rows(v)
        Tk = kk.shape[1]
        hd = E // nh
        o = torch.empty((B, Tq, E), dtype=torch.float32, device=q.device)
        lse = torch.empty((B, nh, Tq), dtype=torch.float32, device=q.device)
        mode, rate, pq, pk = (0, 1, None, None) if spec is None else (spec.mode, spec.rate, spec.pad_q, spec.pad_k)
        if scale is None:
            scale = float(hd) ** -0.5
        with torch.cuda.device(q.device):
            st = _cabi.lib().mrg_attention_forward(
                q.data_ptr(), ldq, kk.data_ptr(), ldk, vv.data_ptr(), ldv, o.data_ptr(), E, lse.data_ptr(), B, nh, Tq,
                Tk, hd, scale, mode, rate, _cabi.ptr(pq), _cabi.ptr(pk), torch.cuda.current_stream(q.device).cuda_stream)
        _cabi.check(st, "mrg_attention_forward")
        ctx.save_for_backward(q, kk, vv, o, lse, pq, pk)
        ctx.meta = (nh, hd, scale, mode, rate, fused, ldq, ldk, ldv)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kk, vv, o, lse, pq, pk = ctx.saved_tensors
        nh, hd, scale, mode, rate, fused, ldq, ldk, ldv = ctx.meta
        B, Tq, E = o.shape
        Tk = kk.shape[1]
        do, lddo = _rows(do)
        dq = torch.empty((B, Tq, E), dtype=torch.float32, device=o.device)
        if fused:
            dkv = torch.empty((B, Tk, 2 * E), dtype=torch.float32, device=o.device)
            dk, dv, ldd = dkv, dkv[..., E:], 2 * E
        else:
            dk = torch.empty((B, Tk, E), dtype=torch.float32, device=o.device)
            dv = torch.empty((B, Tk, E), dtype=torch.float32, device=o.device)
            ldd = E
        dvec = torch.empty((B, nh, Tq), dtype=torch.float32, device=o.device)
        with torch.cuda.device(o.device):
            st = _cabi.lib().mrg_attention_backward(
                q.data_ptr(), ldq, kk.data_ptr(), ldk, vv.data_ptr(), ldv, o.data_ptr(), E, lse.data_ptr(),
                do.data_ptr(), lddo, dq.data_ptr(), E, dk.data_ptr(), ldd, dv.data_ptr(), ldd, dvec.data_ptr(), B, nh,
                Tq, Tk, hd, scale, mode, rate, _cabi.ptr(pq), _cabi.ptr(pk),
                torch.cuda.current_stream(o.device).cuda_stream)
        _cabi.check(st, "mrg_attention_backward")
        return (dq, dkv, None, None, None, None) if fused else (dq, dk, dv, None, None, None)


def fused_attention(q, k, v, num_heads: int, mask: AttentionMaskSpec = None):
    """q [B, Tq, E], k / v [B, Tk, E] (or ``k`` = fused [B, Tk, 2E] and ``v`` = None) -> [B, Tq, E]; fp32 CUDA only."""
    if not (q.is_cuda and q.dtype == torch.float32):
        raise RuntimeError("fused_attention: fp32 CUDA tensors only (no CPU path)")
    return _AttentionFn.apply(q, k, v, num_heads, mask, None)


def _padded_head_dim(hd: int) -> int:
    if hd <= 32:
        return 32
    if hd <= 64:
        return 64
    raise NotImplementedError(f"B200MultiheadAttention: head_dim {hd} > 64 is not built")


def _pad_heads(t: torch.Tensor, nh: int, hd: int, hdp: int) -> torch.Tensor:
    """[B, T, nh*hd] -> [B, T, nh*hdp] with zero columns appended to every head."""
    B, T, _ = t.shape
    return F.pad(t.reshape(B, T, nh, hd), (0, hdp - hd)).reshape(B, T, nh * hdp)


class B200MultiheadAttention(nn.MultiheadAttention):
    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True, is_causal=False):
        if not (query.is_cuda and query.dtype == torch.float32):
            raise RuntimeError("B200MultiheadAttention has no CPU path: fp32 CUDA tensors only")
        if attn_mask is not None and not isinstance(attn_mask, AttentionMaskSpec):
            raise NotImplementedError("B200MultiheadAttention: pass the mask as an AttentionMaskSpec (the rule of "
                                      "multi_modal_metaformer.py:32-79); mask tensors are not read")
        if not (self.batch_first and not need_weights and key_padding_mask is None and not is_causal
                and self.bias_k is None and self.bias_v is None and not self.add_zero_attn and query.dim() == 3):
            raise NotImplementedError("B200MultiheadAttention: only the reference's call form is built (batch_first, "
                                      "need_weights=False, no key_padding_mask / bias_kv / zero_attn / is_causal)")
        if self.training and self.dropout > 0.0:
            raise NotImplementedError("B200MultiheadAttention: attention dropout is not built (the reference "
                                      "configurations use dropout 0)")
        E, nh = self.embed_dim, self.num_heads
        hd = E // nh
        hdp = _padded_head_dim(hd)
        b = self.in_proj_bias
        bq, bk, bv = (None, None, None) if b is None else (b[:E], b[E:2 * E], b[2 * E:])
        kv = None
        if self._qkv_same_embed_dim:
            w = self.in_proj_weight
            q = _LinearFn.apply(query, w[:E], bq)
            if key is value:  # one GEMM for k and v (N = 2E)
                kv = _LinearFn.apply(key, w[E:], None if b is None else b[E:])
                k, v = kv[..., :E], kv[..., E:]
            else:
                k = _LinearFn.apply(key, w[E:2 * E], bk)
                v = _LinearFn.apply(value, w[2 * E:], bv)
        else:
            q = _LinearFn.apply(query, self.q_proj_weight, bq)
            k = _LinearFn.apply(key, self.k_proj_weight, bk)
            v = _LinearFn.apply(value, self.v_proj_weight, bv)
        if hdp != hd:
            q, k, v = (_pad_heads(t, nh, hd, hdp) for t in (q, k, v))
            att = _AttentionFn.apply(q, k, v, nh, attn_mask, float(hd) ** -0.5)
            B, Tq = att.shape[:2]
            att = att.reshape(B, Tq, nh, hdp)[..., :hd].reshape(B, Tq, E)
        elif kv is not None:
            att = _AttentionFn.apply(q, kv, None, nh, attn_mask, None)
        else:
            att = _AttentionFn.apply(q, k, v, nh, attn_mask, None)
        return _LinearFn.apply(att, self.out_proj.weight, self.out_proj.bias), None
