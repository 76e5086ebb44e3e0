import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from multimodalreactiongeneration_b200 import B200LSTM

def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

torch.manual_seed(0)
B, T, H = 6, 24, 256
ref = torch.nn.LSTM(H, H, 1, batch_first=True).double()
m = B200LSTM(H, H, 1, batch_first=True)
m.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
m = m.cuda()
x = torch.randn(B, T, H)
lnw, lnb = torch.randn(H), torch.randn(H)

def tail(kind, y, xx, w, b):
    if kind == "plain": return y
    if kind == "add": return y + xx
    if kind == "ln": return F.layer_norm(y, (H,), w, b)
    if kind == "addln": return F.layer_norm(y + xx, (H,), w, b)
    if kind == "clone_addln": return F.layer_norm(y.clone() + xx, (H,), w, b)
    if kind == "mul": return y * 3.0
    if kind == "randw": return y * xx

for kind in ["plain", "mul", "randw", "add", "ln", "addln", "clone_addln"]:
    ref.zero_grad(); m.zero_grad()
    yr, _ = ref(x.double())
    yr.retain_grad()
    tail(kind, yr, x.double(), lnw.double(), lnb.double()).square().sum().backward()
    captured = {}
    y, _ = m(x.cuda())
    y.register_hook(lambda g: captured.__setitem__("dy", g.detach().clone()))
    tail(kind, y, x.cuda(), lnw.cuda(), lnb.cuda()).square().sum().backward()
    errs = [rel(pm.grad, pr.grad) for pr, pm in zip(ref.parameters(), m.parameters())]
    print(f"{kind:12s} fwd {rel(y, yr):.1e} dy {rel(captured['dy'], yr.grad):.1e} grads " + " ".join(f"{e:.1e}" for e in errs),
          "dy strides", captured['dy'].stride(), "max|dy|", float(captured['dy'].abs().max()), flush=True)
    # replay my backward and torch's with the captured dy on fresh graphs
    dy = captured["dy"]
    ref.zero_grad(); m.zero_grad()
    yr, _ = ref(x.double()); yr.backward(dy.cpu().double())
    y, _ = m(x.cuda()); y.backward(dy)
    errs = [rel(pm.grad, pr.grad) for pr, pm in zip(ref.parameters(), m.parameters())]
    print(f"{'':12s} replay with captured dy: grads " + " ".join(f"{e:.1e}" for e in errs), flush=True)
