"""Pre-split-weight GEMM (mrg_gemm_strided_split, csrc/mrg_gemm_tc4.cu) against fp64 and against the 128 x 128 kernel:
error and time per launch on the shapes of the bench step (developer tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi

L = _cabi.lib()
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


def case(M, N, K, b_mn, bias, flags=0):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = torch.randn(K, N, generator=g)          # B(k, n)
    bias_t = torch.randn(N, generator=g).cuda() if bias else None
    ref = A.double().cpu() @ W.double() + (bias_t.double().cpu() if bias else 0)
    Wd = (W.contiguous() if b_mn else W.t().contiguous()).cuda()
    b_sk, b_sn = (N, 1) if b_mn else (1, K)
    hl = torch.empty((2,) + tuple(Wd.shape), device="cuda")
    _cabi.check(L.mrg_split_tf32(Wd.data_ptr(), hl[0].data_ptr(), hl[1].data_ptr(), Wd.numel(), st()), "split")
    ws = torch.empty(L.mrg_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    c4, c2 = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")

    def new():
        _cabi.check(L.mrg_gemm_strided_split(A.data_ptr(), K, 1, hl[0].data_ptr(), hl[1].data_ptr(), b_sk, b_sn,
                                             _cabi.ptr(bias_t), c4.data_ptr(), N, M, N, K, 0, None, 0, flags, st()), "split gemm")

    def old():
        _cabi.check(L.mrg_gemm_strided(A.data_ptr(), K, 1, Wd.data_ptr(), b_sk, b_sn, _cabi.ptr(bias_t), c2.data_ptr(), N,
                                       M, N, K, 0, 0, ws.data_ptr(), ws.numel(), flags, st()), "gemm")
    new(); old(); torch.cuda.synchronize()
    e4 = float((c4.double().cpu() - ref).abs().max() / ref.abs().max())
    e2 = float((c2.double().cpu() - ref).abs().max() / ref.abs().max())
    same = bool(torch.equal(c4, c2))
    print(f"M={M:6d} N={N:5d} K={K:5d} b_mn={b_mn} bias={int(bias)} flags={flags}: split {timeit(new):7.1f} us (err {e4:.1e})   "
          f"128x128 {timeit(old):7.1f} us (err {e2:.1e})  bit-identical={same}", flush=True)


if __name__ == "__main__":
    for args in [(19200, 1024, 256, 0, True), (19200, 256, 1024, 1, False), (19200, 256, 256, 0, True),
                 (19200, 256, 256, 1, False), (19200, 512, 128, 0, True), (19200, 1024, 80, 0, True),
                 (5000, 320, 200, 0, True), (4100, 260, 36, 1, False), (76800, 1024, 256, 0, True),
                 (19200, 1024, 256, 0, True, 32), (19200, 256, 256, 0, True, 32), (19200, 256, 256, 1, False, 32),
                 (76800, 1024, 256, 0, True, 32), (76800, 256, 256, 0, True, 32), (76800, 256, 256, 1, False, 32),
                 (76800, 256, 1024, 1, False, 32)]:
        if os.environ.get("ONLY_ONE_PASS", "0") == "1" and len(args) < 6:
            continue
        case(*args)
