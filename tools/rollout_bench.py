"""cfg 3: lstm_with_sampling (T=900, B=64): scheduled-sampling training step at several sampling rates and the three
generation modes, for the persistent rollout kernel and the wavefront schedule; kernel-level times of the rollout
kernels from the library's CUDA-event hooks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer

B, T, lead = int(os.environ.get("B", 64)), int(os.environ.get("T", 900)), 30
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).cuda()
batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None), (r(B, lead, 6), None),
         (r(B, lead, 6), None), (r(B, T, 6), None)]


def timed(fn, n):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


for mode in os.environ.get("MODES", "kernel,wavefront").split(","):
    for epoch in (50, 100) if mode != "stepwise" else (50,):
        torch.manual_seed(0)
        m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=100, seed=7)).cuda()
        m.rollout = mode
        m.current_epoch = epoch
        tr = Trainer(m)
        slow = mode == "stepwise" or (mode == "wavefront" and epoch == 100)
        for _ in range(0 if slow else 2):
            tr.train_step(batch)
        dt, loss = timed(lambda: tr.train_step(batch), 1 if slow else 5)
        print(f"{mode}: scheduled-sampling train step, rate {epoch / 100:.1f}: {dt*1e3:.1f} ms/step  {B*T/dt:.0f} frames/s  "
              f"loss {float(loss):.5f}", flush=True)
        if mode == "kernel":
            _cabi.profile_enable(True)
            tr.train_step(batch)
            torch.cuda.synchronize()
            p = _cabi.profile_read()
            _cabi.profile_enable(False)
            print(f"   kernels in that step: rollout fwd {p['rollout_fwd'][0]:.2f} ms ({1e3 * p['rollout_fwd'][0] / T:.2f} us/frame), "
                  f"rollout bwd {p['rollout_bwd'][0]:.2f} ms ({1e3 * p['rollout_bwd'][0] / T:.2f} us/frame); sampler recurrence fwd "
                  f"{p['rec_fwd'][0]:.2f} ms in {p['rec_fwd'][1]} launches, bwd {p['rec_bwd'][0]:.2f} ms in {p['rec_bwd'][1]}; "
                  f"GEMM {p['gemm'][0]:.2f} ms in {p['gemm'][1]}", flush=True)
    if mode == "stepwise":
        continue
    m.eval()
    with torch.no_grad():
        for kind, kw in (("step-wise teacher-forced generation", {}), ("free-running generation", {"full_generation": True})):
            m.prediction(batch, **kw)
            dt, _ = timed(lambda: m.prediction(batch, **kw), 3)
            print(f"  {mode}: {kind}: {dt*1e3:.2f} ms  ({dt*1e6/T:.2f} us per frame, B={B})", flush=True)
        if mode == "kernel":
            _cabi.profile_enable(True)
            m.prediction(batch, full_generation=True)
            torch.cuda.synchronize()
            p = _cabi.profile_read()
            _cabi.profile_enable(False)
            print(f"   free-running generation: rollout kernel {p['rollout_fwd'][0]:.2f} ms ({1e3 * p['rollout_fwd'][0] / T:.2f} us/frame), "
                  f"sampler recurrence {p['rec_fwd'][0]:.2f} ms in {p['rec_fwd'][1]} launches", flush=True)
