import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.mr_gen.model.utils.lstm_block import LSTMLayerd
from multimodalreactiongeneration_b200 import B200LSTM
from oracle import ref_port

def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

def run_layerd(B, T, nl, xgrad):
    torch.manual_seed(0)
    net = LSTMLayerd(input_size=256, lstm_hidden_size=256, affine_hidden_size=256, num_layers=nl,
                     bidirectional=False, use_mixing=False, use_feed_forward=False)
    sd = {k: v.detach().clone().double().requires_grad_(True) for k, v in net.state_dict().items()}
    x = torch.randn(B, T, 256)
    xr = x.double().requires_grad_(xgrad)
    want, _ = ref_port.lstm_layerd(sd, "", xr)
    want.square().sum().backward()
    net = net.cuda()
    xc = x.cuda().requires_grad_(xgrad)
    got, _ = net(xc)
    got.square().sum().backward()
    print(f"layerd B={B} T={T} nl={nl} xgrad={xgrad}: fwd {rel(got, want):.2e}", flush=True)
    for name, p in net.named_parameters():
        print(f"    {name}: {rel(p.grad, sd[name].grad):.2e}")

def run_lstm(B, T, xgrad, env=None):
    torch.manual_seed(0)
    ref = torch.nn.LSTM(256, 256, 1, batch_first=True).double()
    m = B200LSTM(256, 256, 1, batch_first=True)
    m.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    m = m.cuda()
    x = torch.randn(B, T, 256)
    xr = x.double().requires_grad_(xgrad)
    yr, _ = ref(xr); yr.square().sum().backward()
    xc = x.cuda().requires_grad_(xgrad)
    y, _ = m(xc); y.square().sum().backward()
    print(f"lstm B={B} T={T} xgrad={xgrad}: fwd {rel(y, yr):.2e}", flush=True)
    for (n, pr), pm in zip(ref.named_parameters(), m.parameters()):
        print(f"    {n}: {rel(pm.grad, pr.grad):.2e}")

run_lstm(6, 24, False)
run_lstm(6, 24, True)
run_lstm(7, 23, False)
run_layerd(6, 24, 1, False)
run_layerd(6, 24, 2, False)
run_layerd(6, 24, 2, True)
os.environ["MRG_GENERIC_REC"] = "1"
print("generic rec")
run_lstm(6, 24, False)
os.environ["MRG_SIMT_GEMM"] = "1"
