"""Timeline of the graph-replayed bench step (cfg 2): time the GPU runs nothing at all, time only ONE kernel runs, and the
longest stretches by kernel — where the 11 ms go that the per-kernel sums do not show (developer tool)."""
import os, sys, collections
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
tr.enable_cuda_graph(batch)
for _ in range(5):
    tr.train_step_graphed(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step_graphed(batch)
    torch.cuda.synchronize()
ev = [(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort()
t0, t1 = ev[0][0], max(e[1] for e in ev)
# sweep: number of kernels running at each instant
pts = []
for s, e, n in ev:
    pts.append((s, 1)); pts.append((e, -1))
pts.sort()
busy = {0: 0.0, 1: 0.0}
multi = 0.0
cur, last = 0, pts[0][0]
for t, d in pts:
    dt = t - last
    if cur == 0: busy[0] += dt
    elif cur == 1: busy[1] += dt
    else: multi += dt
    cur += d; last = t
print(f"step span {(t1 - t0) / 1e3:.3f} ms over {len(ev)} kernels: idle {busy[0] / 1e3:.3f} ms, exactly one kernel {busy[1] / 1e3:.3f} ms, "
      f"two or more {multi / 1e3:.3f} ms")
# which kernels run ALONE (the serial part), by name
alone = collections.Counter()
active = []
idx = 0
pts2 = sorted([(s, 0, i) for i, (s, e, n) in enumerate(ev)] + [(e, 1, i) for i, (s, e, n) in enumerate(ev)])
running = set(); last = pts2[0][0]
for t, kind, i in pts2:
    if len(running) == 1:
        alone[ev[next(iter(running))][2][:70]] += t - last
    if kind == 0: running.add(i)
    else: running.discard(i)
    last = t
for n, v in alone.most_common(14):
    print(f"  alone {v / 1e3:7.3f} ms  {n}")
# gaps (idle) histogram
gaps = []
end = ev[0][1]
for s, e, n in ev[1:]:
    if s > end: gaps.append(s - end)
    end = max(end, e)
gaps.sort(reverse=True)
print(f"idle gaps: {len(gaps)}, largest {[round(g, 1) for g in gaps[:8]]} us, median {gaps[len(gaps) // 2]:.1f} us")
