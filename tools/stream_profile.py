"""Kernel list of ONE streaming frame (cfg 5: 1024 dyads) in launch order with durations (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.streaming import StreamingGenerator

B = 1024
torch.manual_seed(0)
model = LSTMwithSample(*lstm_with_sampling_cfg(scheduled=False)).cuda()
gen = StreamingGenerator(model, B, use_cuda_graph=False)
gen.reset()
a = torch.randn(B, 1, 80, device="cuda"); p = torch.randn(B, 6, device="cuda")
for _ in range(3):
    gen.step(a, p)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gen.step(a, p)
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
tot = 0.0
for e in ev:
    d = e.time_range.end - e.time_range.start
    tot += d
    print(f"{e.time_range.start - t0:8.1f} us  +{d:6.1f} us  {e.name[:100]}")
print(f"{len(ev)} kernels, {tot:.1f} us of kernel time; span {ev[-1].time_range.end - t0:.1f} us (eager)")
