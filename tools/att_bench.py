"""fp32 attention at the bench shape (B=64, heads=8, T=300, d=32): torch SDPA backends vs explicit matmul/softmax."""
import torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
B, h, T, d = 64, 8, 300, 32
q, k, v = (torch.randn(B, h, T, d, device="cuda", requires_grad=True) for _ in range(3))
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
def run_sdpa(backend):
    def f():
        with sdpa_kernel(backend):
            o = F.scaled_dot_product_attention(q, k, v)
        o.sum().backward()
    return f
def run_math_explicit():
    s = torch.matmul(q, k.transpose(-1, -2)) * (d ** -0.5)
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v)
    o.sum().backward()
for name, be in (("efficient", SDPBackend.EFFICIENT_ATTENTION), ("math", SDPBackend.MATH)):
    try:
        print(name, f"{timeit(run_sdpa(be)):.1f} us fwd+bwd", flush=True)
    except Exception as e:
        print(name, "failed", e)
print("explicit", f"{timeit(run_math_explicit):.1f} us fwd+bwd", flush=True)
