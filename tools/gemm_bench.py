"""Times the tcgen05 3xTF32 GEMM on the shapes of the bench step (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi

L = _cabi.lib()
MB = 19200
shapes = [  # name, M, N, K, a_mn, b_mn
    ("xproj  X.WihT      ", MB, 1024, 256, 0, 0),
    ("dX     dG.Wih      ", MB, 256, 1024, 0, 1),
    ("dWih   dGT.X       ", 1024, 256, MB, 1, 1),
    ("linear 256->256 fwd", MB, 256, 256, 0, 0),
    ("linear 256->256 dW ", 256, 256, MB, 1, 1),
    ("linear 256->64 fwd ", MB, 64, 256, 0, 0),
    ("linear 64->256 fwd ", MB, 256, 64, 0, 0),
    ("linear 80->256 fwd ", MB, 256, 80, 0, 0),
]
for name, M, N, K, a_mn, b_mn in shapes:
    a = torch.randn((K, M) if a_mn else (M, K), device="cuda")
    b = torch.randn((K, N) if b_mn else (N, K), device="cuda")
    c = torch.empty(M, N, device="cuda")
    a_sm, a_sk = (1, M) if a_mn else (K, 1)
    b_sk, b_sn = (N, 1) if b_mn else (1, K)
    ws = torch.empty(L.mrg_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    for flags, tag in ((_cabi.F_GEMM_V3, "tc3 "), (_cabi.F_GEMM_V2, "tc2 "), (_cabi.F_GEMM_V3 | _cabi.F_TF32, "tc3 single-pass tf32")):
        def run():
            st = L.mrg_gemm_strided(a.data_ptr(), a_sm, a_sk, b.data_ptr(), b_sk, b_sn, None, c.data_ptr(), N, M, N, K,
                                    0, 0, ws.data_ptr(), ws.numel(), flags, torch.cuda.current_stream().cuda_stream)
            _cabi.check(st, "gemm")
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{name} {tag} M={M} N={N} K={K}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s-equiv", flush=True)
    if a_mn == 0 and b_mn == 0:
        def run_t(): torch.matmul(a, b.t(), out=c)
        for _ in range(3): run_t()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): run_t()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        print(f"{name} torch fp32 matmul: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TFLOP/s", flush=True)
