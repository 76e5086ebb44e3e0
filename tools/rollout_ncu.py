"""One free-running generation and one scheduled-sampling training step at cfg 3 (T, B from the environment) for an
ncu capture of the rollout kernels:
  ncu --set full --import-source on --clock-control none -k regex:rollout -o gpurun_out/rollout python tools/rollout_ncu.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample

B, T, lead = int(os.environ.get("B", 64)), int(os.environ.get("T", 300)), 30
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).cuda()
batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None), (r(B, lead, 6), None),
         (r(B, lead, 6), None), (r(B, T, 6), None)]
torch.manual_seed(0)
m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=100, seed=7)).cuda()
m.current_epoch = 50
loss = m.training_step(batch)["loss"]
loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
