"""How the bench step scales with T and B (CUDA-graph replay): the T -> 0 intercept is the fixed cost of the ~570 kernels
of a step (launch / dependency latency, pipeline fill of tiny kernels), the slope is the work.  Developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer

for B, T in ((64, 300), (64, 150), (64, 75), (64, 20), (16, 300), (16, 75)):
    torch.manual_seed(0)
    model = SimpleLSTM(*simple_lstm_cfg()).cuda()
    tr = Trainer(model)
    g = torch.Generator().manual_seed(1)
    batch = (torch.randn(B, T, 80, generator=g).cuda(), torch.randn(B, T, 6, generator=g).cuda(),
             torch.randn(B, 1, 6, generator=g).cuda())
    tr.train_step(batch)
    tr.enable_cuda_graph(batch)
    for _ in range(3):
        tr.train_step_graphed(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.train_step_graphed(batch)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B:3d} T={T:3d}: {e0.elapsed_time(e1)/10:.2f} ms/step", flush=True)
    tr.release_cuda_graph()
    del tr, model
