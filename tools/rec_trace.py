"""Event trace of the forward recurrent kernel (needs a -DMRG_REC_TRACE build): prints, for CTA 0, the
clock of every event of the first steps relative to the first event (developer tool)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import lstm_layer, _cabi
B, T, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
L = _cabi.lib()
buf = torch.zeros(16 * 1024 * 2, dtype=torch.int64, device="cuda")
k = 1.0 / H ** 0.5
ws = [torch.empty(4 * H, H, device="cuda").uniform_(-k, k), torch.empty(4 * H, H, device="cuda").uniform_(-k, k),
      torch.empty(4 * H, device="cuda").uniform_(-k, k), torch.empty(4 * H, device="cuda").uniform_(-k, k)]
x = torch.randn(T, B, H, device="cuda")
FLAGS = int(os.environ.get("FLAGS", "0"))   # 32 = reduced-precision mode (tensor-core recurrence at >= 8 rows per cluster)
with torch.no_grad():
    lstm_layer(x, ws, H, 1, flags=FLAGS)          # warm-up
    torch.cuda.synchronize()
    L.mrg_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    lstm_layer(x, ws, H, 1, flags=FLAGS)
    torch.cuda.synchronize()
d = buf.cpu().numpy().reshape(-1, 2)
d = d[d[:, 0] != 0]
ev = sorted((int(c), int(m) >> 48, (int(m) >> 32) & 0xffff, (int(m) >> 16) & 0xffff, int(m) & 0xffff) for c, m in d)
t0 = ev[0][0]
names = {1: "ffma:wait_h", 2: "ffma:h_ready", 3: "ffma:partials_out", 10: "tail:wait_p", 11: "tail:p_ready", 12: "tail:sent"}
lo, hi = int(os.environ.get("STEP_LO", 8)), int(os.environ.get("STEP_HI", 10))
for c, w, e, ch, st in ev:
    if lo <= st <= hi and (w in (0, 5) or w in (8, 9, 15)):
        print(f"{c - t0:8d}  warp {w:2d}  step {st:2d} chunk {ch}  {names.get(e, e)}")
