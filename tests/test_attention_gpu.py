"""GPU parity of the fused fp32 attention kernels (C-ABI ``mrg_attention_forward/backward``) against an fp64
restatement of what the reference computes with nn.MultiheadAttention (mr_gen/model/utils/multi_modal_att.py:12-31,
for_sequential.py:25-50) and the mask of multi_modal_metaformer.py:32-79.  Tolerances: outputs 1e-5, gradients
1e-4 (norm-relative), the bounds of the LSTM path."""
import math

import pytest
import torch

from conftest import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _reference(q, k, v, nh, mask):
    """fp64 softmax attention; q [B,Tq,E], k/v [B,Tk,E]; mask bool [B,nh,Tq,Tk], True = masked."""
    B, Tq, E = q.shape
    Tk, hd = k.shape[1], E // nh
    qh = q.double().reshape(B, Tq, nh, hd).transpose(1, 2)
    kh = k.double().reshape(B, Tk, nh, hd).transpose(1, 2)
    vh = v.double().reshape(B, Tk, nh, hd).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(hd)
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return (torch.softmax(s, dim=-1) @ vh).transpose(1, 2).reshape(B, Tq, E)


CASES = [  # B, nh, hd, Tq, Tk, mode, rate, padded, fused_kv
    (2, 8, 32, 300, 300, 0, 1, False, True),     # SimpleLSTM cross-modal attention (no mask), fused k|v projection
    (3, 4, 64, 330, 330, 1, 1, True, True),      # lstmformer own x partner motion: causal + padding
    (2, 4, 64, 33, 264, 1, 8, True, False),      # own motion x audio at the reference's ratio 8
    (2, 2, 32, 130, 65, 2, 2, True, False),      # queries run faster than keys
    (1, 1, 32, 1, 1, 0, 1, False, False),        # a single frame (autoregressive step)
    (2, 4, 32, 70, 5, 0, 1, False, True),        # ragged tiles
]


@pytest.mark.parametrize("B,nh,hd,Tq,Tk,mode,rate,padded,fused_kv", CASES)
def test_fused_attention_matches_fp64_reference(B, nh, hd, Tq, Tk, mode, rate, padded, fused_kv):
    from multimodalreactiongeneration_b200.attention import AttentionMaskSpec, fused_attention
    g = torch.Generator().manual_seed(Tq * 1000 + Tk)
    E = nh * hd
    q = torch.randn(B, Tq, E, generator=g).cuda().requires_grad_(True)
    kv = torch.randn(B, Tk, 2 * E, generator=g).cuda().requires_grad_(True)
    w = torch.randn(B, Tq, E, generator=g).cuda()
    spec, mask = None, None
    if mode:
        pad_q = torch.zeros(B, Tq, dtype=torch.uint8)
        pad_k = torch.zeros(B, Tk, dtype=torch.uint8)
        if padded:   # trailing padding on one sequence, like the collate function produces
            pad_q[0, Tq - max(1, Tq // 5):] = 1
            pad_k[0, Tk - max(1, Tk // 5):] = 1
        spec = AttentionMaskSpec(mode, rate, pad_q.cuda(), pad_k.cuda())
        mask = spec.materialize(nh)
    if fused_kv:
        out = fused_attention(q, kv, None, nh, spec)
    else:
        out = fused_attention(q, kv[..., :E], kv[..., E:], nh, spec)
    (out * w).sum().backward()
    got = (out.detach().cpu(), q.grad.cpu(), kv.grad.cpu())

    q64 = q.detach().double().requires_grad_(True)
    kv64 = kv.detach().double().requires_grad_(True)
    ref = _reference(q64, kv64[..., :E], kv64[..., E:], nh, mask)
    (ref * w.double()).sum().backward()
    assert rel_err(got[0], ref.detach().cpu()) <= 1e-5
    if Tk == 1:   # softmax over one key: dq and dk are exactly zero in exact arithmetic
        assert float(got[1].abs().max()) <= 1e-6 and float(got[2][..., :E].abs().max()) <= 1e-6
        assert rel_l2(got[2][..., E:], kv64.grad[..., E:].cpu()) <= 1e-4
        return
    assert rel_l2(got[1], q64.grad.cpu()) <= 1e-4
    assert rel_l2(got[2][..., :E], kv64.grad[..., :E].cpu()) <= 1e-4
    assert rel_l2(got[2][..., E:], kv64.grad[..., E:].cpu()) <= 1e-4


@pytest.mark.parametrize("mode,out_tol,grad_tol", [(1, 5e-3, 2e-2), (2, 1e-5, 1e-4)])
def test_attention_modes(mode, out_tol, grad_tol):
    """mrg_attention_set_mode: 1 = one tf32 pass (the tf32 / bf16 precision modes; stated bound of those modes) — must
    differ from the 3xTF32 default, i.e. the switch is honoured; 2 = the CUDA-core fp32 kernels (cross-check path) at the
    fp32-grade tolerance.  Causal + padded lstmformer shape."""
    from multimodalreactiongeneration_b200 import _cabi
    from multimodalreactiongeneration_b200.attention import AttentionMaskSpec, fused_attention
    B, nh, hd, T = 2, 4, 64, 200
    E = nh * hd
    g = torch.Generator().manual_seed(17)
    q0 = torch.randn(B, T, E, generator=g).cuda()
    kv0 = torch.randn(B, T, 2 * E, generator=g).cuda()
    w = torch.randn(B, T, E, generator=g).cuda()
    pad = torch.zeros(B, T, dtype=torch.uint8)
    pad[0, T - 30:] = 1
    spec = AttentionMaskSpec(1, 1, pad.cuda(), pad.cuda())

    def run():
        q, kv = q0.clone().requires_grad_(True), kv0.clone().requires_grad_(True)
        out = fused_attention(q, kv, None, nh, spec)
        (out * w).sum().backward()
        return out.detach().cpu(), q.grad.cpu(), kv.grad.cpu()

    base = run()
    try:
        _cabi.check(_cabi.lib().mrg_attention_set_mode(mode), "mrg_attention_set_mode")
        got = run()
    finally:
        _cabi.lib().mrg_attention_set_mode(0)
    q64, kv64 = q0.double().cpu().requires_grad_(True), kv0.double().cpu().requires_grad_(True)
    ref = _reference(q64, kv64[..., :E], kv64[..., E:], nh, spec.materialize(nh).cpu())
    (ref * w.double().cpu()).sum().backward()
    assert rel_err(got[0], ref.detach()) <= out_tol
    assert rel_l2(got[1], q64.grad) <= grad_tol and rel_l2(got[2], kv64.grad) <= grad_tol
    if mode == 1:
        assert rel_err(got[0], base[0]) > 1e-5   # really one pass


def test_mask_rule_equals_the_reference_mask_tensor():
    """AttentionMaskSpec.materialize == the tile / transpose construction of the reference (restated here)."""
    from multimodalreactiongeneration_b200.attention import AttentionMaskSpec
    for L, S in ((6, 6), (4, 12), (12, 4)):
        pad_q = torch.zeros(2, L, dtype=torch.uint8)
        pad_k = torch.zeros(2, S, dtype=torch.uint8)
        pad_q[1, L - 2:] = 1
        pad_k[1, S - 1:] = 1
        if S % L == 0:
            rate = S // L
            tri = torch.triu(torch.ones(L, L, dtype=torch.bool), diagonal=1)
            want = torch.tile(tri, (1, rate)).view(L, rate, L).transpose(1, 2).contiguous().view(L, S)
            spec = AttentionMaskSpec(1, rate, pad_q, pad_k)
        else:
            rate = L // S
            tri = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)
            want = torch.tile(tri, (rate, 1)).view(rate, S, S).transpose(1, 0).contiguous().view(L, S)
            spec = AttentionMaskSpec(2, rate, pad_q, pad_k)
        both = torch.matmul(pad_q.float().unsqueeze(-1), pad_k.float().unsqueeze(1)).bool()
        want = (want.view(1, 1, L, S) + both.unsqueeze(1)).expand(2, 3, L, S)
        assert torch.equal(spec.materialize(3), want)


def test_multihead_attention_module_uses_the_fused_kernels_and_matches_torch():
    from multimodalreactiongeneration_b200 import _cabi
    from multimodalreactiongeneration_b200.attention import B200MultiheadAttention
    torch.manual_seed(3)
    mine = B200MultiheadAttention(256, 8, batch_first=True).cuda()
    ref = torch.nn.MultiheadAttention(256, 8, batch_first=True).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in mine.state_dict().items()})
    x = torch.randn(4, 150, 256).cuda().requires_grad_(True)
    y = torch.randn(4, 90, 256).cuda().requires_grad_(True)
    n0 = _cabi.launch_count()
    out, _ = mine(x, y, y, need_weights=False)
    out.square().sum().backward()
    assert _cabi.launch_count() - n0 >= 3 + 6   # 3 attention kernels + the projection GEMMs
    x64, y64 = x.detach().double().cpu().requires_grad_(True), y.detach().double().cpu().requires_grad_(True)
    want, _ = ref(x64, y64, y64, need_weights=False)
    want.square().sum().backward()
    assert rel_err(out.detach().cpu(), want.detach()) <= 1e-5
    assert rel_l2(x.grad.cpu(), x64.grad) <= 1e-4 and rel_l2(y.grad.cpu(), y64.grad) <= 1e-4
    for (n, p), (_, r) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel_l2(p.grad.cpu(), r.grad) <= 1e-4, n
