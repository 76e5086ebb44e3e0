"""Smallest reproducer runs for the cluster kernels (developer tool): args B T H D mode(fwd|bwd)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import lstm_layer
B, T, H, D = (int(v) for v in sys.argv[1:5])
mode = sys.argv[5]
k = 1.0 / H ** 0.5
torch.manual_seed(0)
ws = []
for _ in range(D):
    ws += [torch.empty(4 * H, H, device="cuda").uniform_(-k, k).requires_grad_(True),
           torch.empty(4 * H, H, device="cuda").uniform_(-k, k).requires_grad_(True),
           torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True),
           torch.empty(4 * H, device="cuda").uniform_(-k, k).requires_grad_(True)]
x = torch.randn(T, B, H, device="cuda", requires_grad=(mode == "bwd"))
if mode == "fwd":
    with torch.no_grad():
        y, h, c = lstm_layer(x, ws, H, D)
else:
    y, h, c = lstm_layer(x, ws, H, D)
    y.sum().backward()
torch.cuda.synchronize()
print("ok", B, T, H, D, mode, float(y.abs().sum()))
