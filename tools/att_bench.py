"""Kernel-level time of the cross-modal attention core: fused fp32 kernels vs torch SDPA (fp32 -> sm_80 mem-efficient
kernels), SimpleLSTM shape (B=64, 8 heads x 32, T=300) and lstmformer shape (B=256, 4 heads x 64, T=330, causal)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodalreactiongeneration_b200.attention import AttentionMaskSpec, fused_attention


def timed(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, B, nh, hd, T, causal in (("simple_lstm", 64, 8, 32, 300, False), ("lstmformer", 256, 4, 64, 330, True)):
    E = nh * hd
    q = torch.randn(B, T, E, device="cuda", requires_grad=True)
    kv = torch.randn(B, T, 2 * E, device="cuda", requires_grad=True)
    w = torch.randn(B, T, E, device="cuda")
    spec = AttentionMaskSpec(1, 1, torch.zeros(B, T, dtype=torch.uint8, device="cuda"),
                             torch.zeros(B, T, dtype=torch.uint8, device="cuda")) if causal else None
    mask = None if spec is None else ~spec.materialize(nh)

    def sdpa():
        qh = q.reshape(B, T, nh, hd).transpose(1, 2)
        kh = kv[..., :E].reshape(B, T, nh, hd).transpose(1, 2)
        vh = kv[..., E:].reshape(B, T, nh, hd).transpose(1, 2)
        return F.scaled_dot_product_attention(qh, kh, vh, attn_mask=mask).transpose(1, 2).reshape(B, T, E)

    flops = 4.0 * B * nh * T * T * hd * (0.5 if causal else 1.0)
    for tag, fwd in (("fused", lambda: fused_attention(q, kv, None, nh, spec)), ("sdpa ", sdpa)):
        t_f = timed(lambda: fwd())
        out = fwd()
        t_b = timed(lambda: torch.autograd.grad(out, (q, kv), w, retain_graph=True))
        print(f"{name:12s} {tag}: fwd {t_f:7.1f} us ({flops/t_f*1e-6:5.1f} TFLOP/s)  bwd {t_b:7.1f} us "
              f"({3.5*flops/t_b*1e-6:5.1f} TFLOP/s incl. recompute)", flush=True)
