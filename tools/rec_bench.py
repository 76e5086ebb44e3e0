"""Kernel-level time of the recurrent fwd / BPTT kernels (library CUDA-event hooks) for a few shapes,
developer tool, not the bench."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import lstm_layer, _cabi


def run(B, T, H, D, flags, iters=5):
    I = H
    k = 1.0 / H ** 0.5
    ws = []
    for _ in range(D):
        ws += [torch.empty(4 * H, I, device="cuda").uniform_(-k, k).requires_grad_(True),
               torch.empty(4 * H, H, device="cuda").uniform_(-k, k).requires_grad_(True),
               torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True),
               torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True)]
    x = torch.randn(T, B, I, device="cuda", requires_grad=True)
    for it in range(iters + 2):
        if it == 2:
            torch.cuda.synchronize()
            _cabi.profile_enable(True)
        y, h, c = lstm_layer(x, ws, H, D, flags=flags)
        y.sum().backward()
    torch.cuda.synchronize()
    p = _cabi.profile_read()
    _cabi.profile_enable(False)
    f = p["rec_fwd"][0] / max(1, p["rec_fwd"][1])
    b = p["rec_bwd"][0] / max(1, p["rec_bwd"][1])
    return f, b


if __name__ == "__main__":
    shapes = [(64, 300, 256, 1), (64, 300, 256, 2), (8, 300, 256, 1), (256, 300, 256, 1), (64, 300, 128, 1),
              (64, 300, 128, 2), (1024, 30, 256, 1), (16, 300, 256, 1), (32, 300, 256, 1)]
    if os.environ.get("REDUCED", "0") == "1":
        shapes = [(16, 300, 256, 1), (32, 300, 256, 1), (64, 300, 256, 1), (96, 300, 256, 1), (64, 300, 256, 2), (128, 300, 256, 1), (240, 300, 256, 1), (256, 300, 256, 1), (256, 300, 256, 2)]
    for (B, T, H, D) in shapes:
        row = f"B={B:5d} T={T} H={H} D={D}:"
        variants = (("cluster", 0),) if os.environ.get("REDUCED", "0") != "1" else (("fp32", 0), ("tf32-mode", _cabi.F_TF32))
        for name, fl in variants:
            try:
                f, b = run(B, T, H, D, fl)
                row += f"  {name} fwd {f*1e3:8.1f} us ({f*1e3/T:5.2f}/step) bwd {b*1e3:8.1f} us ({b*1e3/T:5.2f}/step)"
            except Exception as e:  # noqa
                row += f"  {name} FAILED {e}"
        print(row, flush=True)
