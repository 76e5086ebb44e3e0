"""Phase timing of the forward rollout kernel (developer tool): build the library with -DMRG_RO_TRACE
(MRG_EXTRA_NVCC_FLAGS=-DMRG_RO_TRACE python -m multimodalreactiongeneration_b200._build --force), run a free-running
generation at cfg 3 and print the clocks thread 0 of CTA 0 spends per phase of a step."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi
from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample

B, T, lead = int(os.environ.get("B", 64)), int(os.environ.get("T", 900)), 30
g = torch.Generator().manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g).cuda()
batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None), (r(B, lead, 6), None),
         (r(B, lead, 6), None), (r(B, T, 6), None)]
torch.manual_seed(0)
m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=100, seed=7)).cuda().eval()
L = _cabi.lib()
buf = (ctypes.c_ulonglong * 64)()
with torch.no_grad():
    m.prediction(batch, full_generation=True)
    torch.cuda.synchronize()
    L.mrg_debug_rollout_trace(buf, 1)
    m.prediction(batch, full_generation=True)
    torch.cuda.synchronize()
    L.mrg_debug_rollout_trace(buf, 1)
B_NAMES = {0: "0: dy + prefetch + sync", 1: "1: FFN hidden grad + sync", 2: "2: dx_L + cp.async wait + sync"}
for l in (0, 1):
    B_NAMES.update({3 + 6 * l: f"3: LN^T sums + send (layer slot {l})", 4 + 6 * l: "   wait sums", 5 + 6 * l: "4: LN^T + cell^T + sync",
                    6 + 6 * l: "5: mat-vec^T + sync", 7 + 6 * l: "   reduce + send", 8 + 6 * l: "   wait partials"})
B_NAMES.update({15: "6: dx + sync + prefetch issue", 16: "7: W_prev^T + send", 17: "   wait prev", 18: "   d(prev) + sync"})
names = {0: "top: cp.async wait + sync", 1: "A: x0 + prefetch issue", 2: "B0: mat-vec", 3: "sync", 4: "C0: gates + send",
         5: "wait window 0", 6: "D0: LayerNorm + sync", 7: "B1: mat-vec", 8: "sync", 9: "C1: gates + send",
         10: "wait window 1", 11: "D1: LayerNorm + sync", 12: "E: FFN hidden + sync", 13: "send f", 14: "wait f window",
         15: "F: pose + state", 20: "  A.1 x0 compute", 21: "  A.2 sync", 22: "  A.3 feedback prefetch"}
tot = sum(buf[i] for i in range(32))
for i in sorted(names):
    print(f"{names[i]:28s} {buf[i] / T:9.1f} clk/step  {100 * buf[i] / max(tot, 1):5.1f} %")
print(f"total {tot / T:.1f} clk/step = {tot / T / 1.965e3:.2f} us at 1965 MHz")
if os.environ.get("BWD", "1") == "1":
    loss = m.train().training_step(batch)["loss"] if False else None
    m.train()
    m.current_epoch = 50
    L.mrg_debug_rollout_trace(buf, 1)
    m.training_step(batch)["loss"].backward()
    torch.cuda.synchronize()
    L.mrg_debug_rollout_trace(buf, 1)
    tot = sum(buf[32 + i] for i in range(32))
    print("backward kernel (layer slot 1 = last block, processed first):")
    for i in sorted(B_NAMES):
        print(f"{B_NAMES[i]:40s} {buf[32 + i] / T:9.1f} clk/step  {100 * buf[32 + i] / max(tot, 1):5.1f} %")
    print(f"total {tot / T:.1f} clk/step = {tot / T / 1.965e3:.2f} us at 1965 MHz")
