"""torch.profiler kernel table of the cfg 4 step (lstmformer, B=256 x T=300+30), developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from multimodalreactiongeneration_b200 import set_precision
from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer

set_precision(os.environ.get("PRECISION", "bf16"))
torch.manual_seed(0)
m = Metaformer(*metaformer_cfg()).cuda()
tr = Trainer(m)
B = int(os.environ.get("B", 256))
batch = [(t.cuda(), None) for t in bench.nx_batch(1, B, 300, pin=False)]
for _ in range(3):
    tr.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.train_step(batch)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.self_device_time_total for k in ka)
print(f"GPU busy {tot/2e3:.2f} ms/step")
rows = sorted(ka, key=lambda k: -k.self_device_time_total)[:40]
for k in rows:
    print(f"{k.self_device_time_total/2e3:9.3f} ms/step  {k.count//2:5d} x {k.self_device_time_total/max(1,k.count):9.1f} us  {k.key[:110]}")
