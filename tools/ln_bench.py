"""Residual + LayerNorm forward / backward and the bias-gradient column sum at the cfg 2 / cfg 4 row counts: time per call
and achieved HBM bandwidth (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.layernorm import residual_layer_norm, _ResidualLNFn
from multimodalreactiongeneration_b200.linear import _colsum


def timeit(fn, n=20):
    """n calls captured in one CUDA graph (the calls are ~10 us of host work each: an eager loop would time Python)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for rows in (19200, 76800):
    H = 256
    B = rows // 300
    # several buffer sets so that consecutive calls do not hit L2
    sets = [(torch.randn(B, 300, H, device="cuda"), torch.randn(B, 300, H, device="cuda", requires_grad=True),
             torch.randn(B, 300, H, device="cuda")) for _ in range(4)]
    gamma = torch.ones(H, device="cuda", requires_grad=True)
    beta = torch.zeros(H, device="cuda", requires_grad=True)
    i = [0]

    def fwd():
        y, x, _ = sets[i[0] % 4]
        i[0] += 1
        with torch.no_grad():
            residual_layer_norm(y, x, gamma, beta)

    class Ctx:   # the Function's static methods called directly: the C-ABI launches without the autograd engine
        def save_for_backward(self, *t):
            self.saved_tensors = t

    outs = []
    for y, x, d in sets:
        c = Ctx()
        _ResidualLNFn.forward(c, y, x.detach(), gamma.detach(), beta.detach(), 1e-5)
        outs.append((c, d))

    def bwd():
        c, d = outs[i[0] % 4]
        i[0] += 1
        _ResidualLNFn.backward(c, d)

    def cs():
        y = sets[i[0] % 4][0]
        i[0] += 1
        _colsum(y.view(-1, H))

    mb = rows * H * 4 / 1e6
    tf, tb, tc = timeit(fwd), timeit(bwd), timeit(cs)
    print(f"rows {rows}: ln fwd {tf:6.1f} us ({3 * mb / tf:5.2f} TB/s)  ln bwd {tb:6.1f} us ({4 * mb / tb:5.2f} TB/s)"
          f"  colsum {tc:6.1f} us ({mb / tc:5.2f} TB/s)")
