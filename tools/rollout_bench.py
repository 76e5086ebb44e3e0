"""cfg 3: lstm_with_sampling scheduled-sampling training step (T=900, B=64), wavefront vs stepwise."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer

B, T, lead = int(os.environ.get("B", 64)), int(os.environ.get("T", 900)), 30
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).cuda()
batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None), (r(B, lead, 6), None),
         (r(B, lead, 6), None), (r(B, T, 6), None)]
for mode in os.environ.get("MODES", "wavefront,stepwise").split(","):
    torch.manual_seed(0)
    m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=100, seed=7)).cuda()
    m.rollout = mode
    m.current_epoch = 50
    tr = Trainer(m)
    n = 5 if mode == "wavefront" else 1
    for _ in range(2 if mode == "wavefront" else 0):
        tr.train_step(batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n):
        loss = tr.train_step(batch)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"{mode}: {dt*1e3:.1f} ms/step  {B*T/dt:.0f} frames/s  loss {float(loss):.5f}", flush=True)
    if mode == "wavefront":
        m.eval()
        with torch.no_grad():
            for kind, kw in (("teacher-forced generation", {}), ("free-running generation", {"full_generation": True})):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                m.prediction(batch, **kw)
                torch.cuda.synchronize(); print(f"  {kind}: {(time.perf_counter()-t0)*1e3:.1f} ms", flush=True)
