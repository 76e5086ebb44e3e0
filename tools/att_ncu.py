"""One forward + backward of the fused attention at the SimpleLSTM shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.attention import fused_attention
B, nh, hd, T = 64, 8, int(os.environ.get("HD", 32)), 300
E = nh * hd
q = torch.randn(B, T, E, device="cuda", requires_grad=True)
kv = torch.randn(B, T, 2 * E, device="cuda", requires_grad=True)
for _ in range(2):
    out = fused_attention(q, kv, None, nh, None)
    out.sum().backward()
torch.cuda.synchronize()
