"""Event trace of one CTA of the projection GEMM (needs a -DMRG_REC_TRACE build: MRG_EXTRA_NVCC_FLAGS=-DMRG_REC_TRACE
python -m multimodalreactiongeneration_b200._build --force): clock of every pipeline event relative to the CTA's start
(developer tool).  usage: gemm_trace.py M N K [flags]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi
M, N, K = (int(v) for v in sys.argv[1:4])
flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
split = len(sys.argv) > 5 and sys.argv[5] == "split"
L = _cabi.lib()
buf = torch.zeros(16 * 1024 * 2, dtype=torch.int64, device="cuda")
a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda"); c = torch.empty(M, N, device="cuda")
ws = torch.empty(L.mrg_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
hl = torch.empty(2, N, K, device="cuda")
L.mrg_split_tf32(b.data_ptr(), hl[0].data_ptr(), hl[1].data_ptr(), b.numel(), st)
def run():
    if split:
        _cabi.check(L.mrg_gemm_strided_split(a.data_ptr(), K, 1, hl[0].data_ptr(), hl[1].data_ptr(), 1, K, None, c.data_ptr(), N,
                                             M, N, K, 0, None, 0, flags, st), "split gemm")
        return
    _cabi.check(L.mrg_gemm_nt(a.data_ptr(), b.data_ptr(), None, c.data_ptr(), M, N, K, ws.data_ptr(), ws.numel(), flags, st), "gemm")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"M={M} N={N} K={K} flags={flags}: {e0.elapsed_time(e1) * 100:.1f} us per launch (trace build)")
L.mrg_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
L.mrg_debug_set_trace(None)
d = buf.cpu().numpy().reshape(-1, 2)
d = d[d[:, 0] != 0]
ev = sorted((int(c_), int(m) >> 48, (int(m) >> 32) & 0xffff, int(m) & 0xffffffff) for c_, m in d)
names = {15: "tmaB:slot_free->issue", 19: "mma:acc_free", 52: "epi:acc_released",
         1: "start", 2: "setup_done", 10: "tma:slot_free->issue", 20: "mma:operands_ready", 21: "mma:issued+commit",
         30: "cvtA:tile_landed", 31: "cvtA:done", 40: "cvtB:tile_landed", 41: "cvtB:done", 50: "epi:enter",
         51: "epi:acc_ready", 52: "epi:staged", 53: "epi:stored"}
t0 = ev[0][0]
for c_, w, e, i in ev:
    if w in (0, 1, 3, 4, 8, 12) and (i < 24 or e >= 50):
        print(f"{c_ - t0:8d}  warp {w:2d}  kb {i:3d}  {names.get(e, e)}")
