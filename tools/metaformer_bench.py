"""cfg 4: lstmformer (Metaformer) teacher-forced training step, B x 300 frames + 30 lead frames (15 LSTM mixers of
H=256, 10 masked cross-modal attentions).  B=256/GPU is BASELINE's shape; env B/T/STEPS/MODE override."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi
from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer

B, T, lead = int(os.environ.get("B", 256)), int(os.environ.get("T", 300)), 30
steps = int(os.environ.get("STEPS", 5))
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).cuda()
batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None), (r(B, lead, 6), None),
         (r(B, lead, 6), None), (r(B, T, 6), None)]
torch.manual_seed(0)
m = Metaformer(*metaformer_cfg()).cuda()
tr = Trainer(m)
L = _cabi.lib()
for _ in range(2):
    loss = tr.train_step(list(batch))
torch.cuda.synchronize()
n0 = L.mrg_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.train_step(list(batch))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"metaformer train step B={B} T={T}+{lead}: {ms:.1f} ms/step  {B*T/ms*1e3:.0f} frames/s  loss {float(loss):.5f}  "
      f"library launches/step {(L.mrg_launch_count()-n0)//steps}  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB",
      flush=True)
if os.environ.get("PROFILE"):
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_step(list(batch)); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
