"""Diagnostic: gradient errors of one bench-shape SimpleLSTM training step vs the fp64 CPU oracle, for the CPU fp32
oracle (conditioning band) and for the CUDA path with the weight-gradient overlap / two-stream encoders on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import ref_port
from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer


def model_batch():
    torch.manual_seed(0)
    m = SimpleLSTM(*simple_lstm_cfg(bench.HIDDEN, bench.LAYERS, False, bench.ACOUSTIC, bench.POSE))
    return m, bench.synthetic_batch(1234, bench.B_PER_GPU, pin=False)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def oracle(dtype):
    m, batch = model_batch()
    sd = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in m.state_dict().items()}
    loss = ref_port.simple_lstm_training_step(sd, tuple(t.to(dtype) for t in batch))
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items()}


def report(tag, loss, grads, want_loss, want):
    errs = sorted(((rel(g, want[k]), k) for k, g in grads.items()), reverse=True)
    print(f"== {tag}: loss rel err {abs(loss - want_loss) / abs(want_loss):.2e}; worst gradients:")
    for e, k in errs[:6]:
        print(f"   {e:.3e}  {k}  |g|={float(want[k].norm()):.3e}")
    print(f"   median {errs[len(errs) // 2][0]:.3e}", flush=True)


if __name__ == "__main__":
    l64, g64 = oracle(torch.float64)
    l32, g32 = oracle(torch.float32)
    report("CPU fp32 oracle", l32, g32, l64, g64)
    for ov, ts in (("1", "1"), ("0", "1"), ("1", "0"), ("0", "0")):
        os.environ["MRG_WGRAD_OVERLAP"], os.environ["MRG_TWO_STREAMS"] = ov, ts
        m, batch = model_batch()
        m = m.cuda()
        tr = Trainer(m)
        loss = tr.forward_backward(tuple(t.cuda() for t in batch))
        torch.cuda.synchronize()
        report(f"CUDA overlap={ov} two_streams={ts}", float(loss), {k: p.grad.cpu() for k, p in m.named_parameters()},
               l64, g64)
    os.environ["MRG_SIMT_GEMM"] = "1"
    m, batch = model_batch()
    m = m.cuda()
    tr = Trainer(m)
    loss = tr.forward_backward(tuple(t.cuda() for t in batch))
    torch.cuda.synchronize()
    report("CUDA SIMT GEMM overlap=0 two_streams=0", float(loss), {k: p.grad.cpu() for k, p in m.named_parameters()},
           l64, g64)
