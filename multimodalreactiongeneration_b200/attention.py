"""``B200MultiheadAttention`` — ``nn.MultiheadAttention`` whose q / k / v / output PROJECTIONS (the dense
``[B*T, E] x [E, E]`` contractions, 2/3 of the module's FLOPs at the reference's shapes) run on the library's
tcgen05 3xTF32 GEMM instead of torch's fp32 SIMT sgemm.  ``softmax(QK^T / sqrt(d)) V`` itself stays
``F.scaled_dot_product_attention`` (library kernel; SURVEY.md §8(f) item 3 lists a fused attention as "next").

Reference call sites: mr_gen/model/utils/multi_modal_att.py:12-31 (``batch_first=True``, ``need_weights=False``,
no masks) and mr_gen/model/utils/for_sequential.py:25-50 (lstmformer's integrators: ``batch_first=True``, a bool
``attn_mask`` of shape ``[B*heads, L, S]`` from ``gen_attention_mask``, True = masked out).  Parameters, their
names / shapes / init order and ``state_dict`` keys are ``nn.MultiheadAttention``'s
(``in_proj_weight`` or ``q_/k_/v_proj_weight``, ``in_proj_bias``, ``out_proj.*``).  Argument combinations the
reference does not use on this path (attention weights requested, key padding masks, ``bias_k`` /
``add_zero_attn``, time-major layout) are delegated to ``nn.MultiheadAttention.forward`` unchanged."""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .linear import _LinearFn


class B200MultiheadAttention(nn.MultiheadAttention):
    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True, is_causal=False):
        plain = (self.batch_first and not need_weights and key_padding_mask is None
                 and (attn_mask is None or attn_mask.dim() in (2, 3, 4)) and not is_causal
                 and self.bias_k is None and self.bias_v is None and not self.add_zero_attn
                 and query.dim() == 3 and query.is_cuda and query.dtype == torch.float32)
        if not plain:
            return super().forward(query, key, value, key_padding_mask=key_padding_mask, need_weights=need_weights,
                                   attn_mask=attn_mask, average_attn_weights=average_attn_weights,
                                   is_causal=is_causal)
        E, nh = self.embed_dim, self.num_heads
        hd = E // nh
        b = self.in_proj_bias
        bq, bk, bv = (None, None, None) if b is None else (b[:E], b[E:2 * E], b[2 * E:])
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
        B, Tq, Tk = query.shape[0], query.shape[1], key.shape[1]
        q = q.reshape(B, Tq, nh, hd).transpose(1, 2)
        k = k.reshape(B, Tk, nh, hd).transpose(1, 2)
        v = v.reshape(B, Tk, nh, hd).transpose(1, 2)
        sdpa_mask = None
        if attn_mask is not None:
            # [L,S] | [B*heads,L,S] (nn.MultiheadAttention's forms) | [B, 1 or heads, L, S] (broadcast view: no
            # per-head copy is made)
            m = attn_mask.reshape(B, nh, Tq, Tk) if attn_mask.dim() == 3 else attn_mask
            # nn.MultiheadAttention: bool True = "not allowed", float = additive; SDPA: bool True = "takes part"
            sdpa_mask = ~m if m.dtype == torch.bool else m.to(q.dtype)
        att = F.scaled_dot_product_attention(q, k, v, attn_mask=sdpa_mask,
                                             dropout_p=self.dropout if self.training else 0.0)
        att = att.transpose(1, 2).reshape(B, Tq, E)
        return _LinearFn.apply(att, self.out_proj.weight, self.out_proj.bias), None
