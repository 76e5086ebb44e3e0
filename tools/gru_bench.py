"""Kernel-level time of the GRU recurrence (generic kernels) at lstmformer's mixer shapes, beside torch.nn.GRU
(cuDNN) on the same device.  Developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import B200GRU, _cabi


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in (64, 256):
    T, H = 300, 256
    mine = B200GRU(H, H, 1, batch_first=True).cuda()
    ref = torch.nn.GRU(H, H, 1, batch_first=True).cuda()
    x = torch.randn(B, T, H, device="cuda", requires_grad=True)
    _cabi.profile_enable(True)
    for _ in range(3):
        mine(x)[0].sum().backward()
    torch.cuda.synchronize()
    p = _cabi.profile_read()
    _cabi.profile_enable(False)
    f, b = p["rec_fwd"][0] / p["rec_fwd"][1], p["rec_bwd"][0] / p["rec_bwd"][1]
    t_mine = timed(lambda: mine(x)[0].sum().backward())
    t_ref = timed(lambda: ref(x)[0].sum().backward())
    print(f"GRU B={B} T={T} H={H}: recurrence fwd {f*1e3:.0f} us ({f*1e3/T:.2f}/step) bwd {b*1e3:.0f} us ({b*1e3/T:.2f}/step); "
          f"layer fwd+bwd {t_mine:.2f} ms (torch.nn.GRU / cuDNN {t_ref:.2f} ms)", flush=True)
