"""torch.profiler summary of the bench step: GPU busy time vs wall time, top kernels (developer tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer
import bench

torch.manual_seed(0)
model = SimpleLSTM(*simple_lstm_cfg()).cuda()
tr = Trainer(model)
batch = tuple(t.cuda() for t in bench.synthetic_batch(1, 64, False))
for _ in range(5):
    tr.train_step(batch)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    tr.train_step(batch)
t_cpu = time.perf_counter() - t0          # host time to ENQUEUE 10 steps
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"10 steps: host enqueue {t_cpu*100:.2f} ms/step, wall {t_all*100:.2f} ms/step", flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_step(batch)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.self_device_time_total for k in ka)
print(f"GPU busy {tot/3e3:.2f} ms/step")
print(ka.table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=70))
