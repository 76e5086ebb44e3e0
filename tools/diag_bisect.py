"""Diagnostic: where do the tensor-core and the SIMT GEMM paths part ways in the backward of the bench step?
Runs the bench-shape SimpleLSTM step twice on the GPU (MRG_SIMT_GEMM=1 as the reference) and compares the gradient
that arrives at / leaves every sub-module, in backward order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM

os.environ["MRG_TWO_STREAMS"] = "0"


def run(simt):
    os.environ["MRG_SIMT_GEMM"] = "1" if simt else "0"
    torch.manual_seed(0)
    m = SimpleLSTM(*simple_lstm_cfg(bench.HIDDEN, bench.LAYERS, False, bench.ACOUSTIC, bench.POSE)).cuda()
    batch = tuple(t.cuda() for t in bench.synthetic_batch(1234, bench.B_PER_GPU, pin=False))
    rec = []

    def hook(name):
        def fn(mod, gin, gout):
            go = [g.detach().clone() for g in gout if g is not None]
            gi = [g.detach().clone() for g in gin if g is not None]
            rec.append((name, go, gi))
        return fn

    for name, mod in m.named_modules():
        if name and len(list(mod.children())) == 0 or name.endswith("lstm_module.module") :
            mod.register_full_backward_hook(hook(name))
    loss = m.training_step(batch)["loss"]
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    return rec, grads


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


if __name__ == "__main__":
    ref, gref = run(True)
    got, ggot = run(False)
    assert [r[0] for r in ref] == [r[0] for r in got]
    for (name, go_r, gi_r), (_, go_g, gi_g) in zip(ref, got):
        eo = max([rel(a, b) for a, b in zip(go_g, go_r)] or [0.0])
        ei = max([rel(a, b) for a, b in zip(gi_g, gi_r)] or [0.0])
        mark = "  <<<<" if ei > 1e-4 and eo <= 1e-4 else ""
        shp = [tuple(t.shape) for t in go_r]
        print(f"{name:90s} grad_out err {eo:.2e}  grad_in err {ei:.2e} {shp}{mark}", flush=True)
    errs = sorted(((rel(ggot[k], gref[k]), k) for k in gref), reverse=True)
    for e, k in errs[:8]:
        print(f"param {e:.2e} {k}")
