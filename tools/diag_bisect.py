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

    fwd = []

    relu_state = {}

    def fhook(name):
        def fn(mod, inp, out):
            o = out[0] if isinstance(out, (tuple, list)) else out
            if torch.is_tensor(o):
                fwd.append((name, o.detach().clone()))
                if name.endswith("relu") and o.requires_grad:
                    fwd_copy = o.detach().clone()

                    def at_backward(g, o=o, fwd_copy=fwd_copy, name=name):
                        now = o.detach()
                        relu_state[name] = (fwd_copy, now.clone(), g.detach().clone())
                    o.register_hook(at_backward)
        return fn

    for name, mod in m.named_modules():
        if name and len(list(mod.children())) == 0 or name.endswith("lstm_module.module") :
            mod.register_full_backward_hook(hook(name))
            mod.register_forward_hook(fhook(name))
    loss = m.training_step(batch)["loss"]
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    return rec, grads, fwd, relu_state


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


if __name__ == "__main__":
    ref, gref, fref, rref = run(True)
    got, ggot, fgot, rgot = run(False)
    for name in rgot:
        f0, now, g = rgot[name]
        f0r, nowr, gr = rref[name]
        changed = int((f0 != now).sum())
        flips = ((now > 0) != (nowr > 0))
        nflip = int(flips.sum())
        gnorm = float(g.double().norm())
        gflip = float((g.double() * flips).norm())
        print(f"RELU {name}: saved output changed since forward in {changed} entries (tc run); mask flips vs simt run {nflip}; "
              f"|grad| {gnorm:.3e}, |grad on flipped| {gflip:.3e}; |out| at flips max "
              f"{float((now.abs() * flips).max()):.3e} out max {float(now.abs().max()):.3e}", flush=True)
    for (name, a), (_, b) in zip(fgot, fref):
        d = (a.double() - b.double()).abs()
        scale = float(b.double().abs().max())
        nbad = int((d > 1e-4 * scale).sum())
        extra = ""
        if nbad and a.dim() == 3:
            rows = (d > 1e-4 * scale).any(dim=-1).nonzero()
            extra = f" first bad (b,t) {rows[:4].tolist()} last {rows[-2:].tolist()}"
        print(f"FWD {name:86s} max err {float(d.max()) / scale:.2e} bad {nbad}/{a.numel()}{extra}", flush=True)
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
