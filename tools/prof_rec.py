"""One fwd+bwd of a single B200LSTM layer at the headline shape, for ncu (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import B200LSTM

B, T, H = int(os.environ.get("B", 64)), int(os.environ.get("T", 300)), int(os.environ.get("H", 256))
m = B200LSTM(H, H, 1, batch_first=True).cuda()
x = torch.randn(B, T, H, device="cuda", requires_grad=True)
for _ in range(int(os.environ.get("ITERS", 3))):
    y, _ = m(x)
    y.sum().backward()
torch.cuda.synchronize()
print("ok")
