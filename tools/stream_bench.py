"""BASELINE configs[4]: streaming generation, 1024 concurrent dyads, one frame per call; p50/p99 latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.streaming import StreamingGenerator

B, frames = int(os.environ.get("B", 1024)), int(os.environ.get("FRAMES", 300))
torch.manual_seed(0)
model = LSTMwithSample(*lstm_with_sampling_cfg(scheduled=False)).cuda()
g = torch.Generator().manual_seed(3)
audio = torch.randn(frames, B, 1, 80, generator=g).pin_memory()
partner = torch.randn(frames, B, 6, generator=g).pin_memory()
out_host = torch.empty(B, 6).pin_memory()
for graph in (True, False):
    gen = StreamingGenerator(model, B, use_cuda_graph=graph)
    gen.reset()
    for f in range(5):
        gen.step(audio[f], partner[f])
    torch.cuda.synchronize()
    lat = []
    for f in range(frames):
        t0 = time.perf_counter()
        y = gen.step(audio[f], partner[f])          # H2D of the frame's inputs + the frame
        out_host.copy_(y, non_blocking=True)        # D2H of the 1024 x 6 poses
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e6)
    lat.sort()
    print(f"B={B} cuda_graph={graph}: per-frame latency p50 {lat[len(lat)//2]:.1f} us  p99 {lat[int(len(lat)*0.99)]:.1f} us  "
          f"max {lat[-1]:.1f} us  (budget 33333 us at 30 fps)  -> {B*1e6/lat[len(lat)//2]:.0f} dyad-frames/s", flush=True)
# graph vs eager must agree
g1, g2 = StreamingGenerator(model, B, True), StreamingGenerator(model, B, False)
g1.reset(); g2.reset()
for f in range(20):
    y1 = g1.step(audio[f], partner[f]).clone(); y2 = g2.step(audio[f], partner[f]).clone()
print("graph vs eager max abs diff after 20 frames:", float((y1 - y2).abs().max()))
