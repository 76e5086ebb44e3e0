"""Diagnostic: the tcgen05 GEMM vs fp64 on the exact shapes / operand majors of the Linear layers at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import _cabi

L = _cabi.lib()


def run(M, N, K, a_mn, b_mn, acc, flags, seed=0, sparse_rows=None):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    if sparse_rows is not None:
        keep = torch.zeros(M, 1)
        keep[sparse_rows - 1::sparse_rows] = 1.0
        A = A * keep
    Bm = torch.randn(K, N, generator=g)
    c0 = torch.randn(M, N, generator=g)
    ref = A.double() @ Bm.double() + (c0.double() if acc else 0)
    a_dev = (A.t().contiguous() if a_mn else A.contiguous()).cuda()
    b_dev = (Bm.contiguous() if b_mn else Bm.t().contiguous()).cuda()
    a_sm, a_sk = (1, M) if a_mn else (K, 1)
    b_sk, b_sn = (N, 1) if b_mn else (1, K)
    ws = torch.empty(L.mrg_gemm_workspace_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    c = c0.clone().cuda()
    st = L.mrg_gemm_strided(a_dev.data_ptr(), a_sm, a_sk, b_dev.data_ptr(), b_sk, b_sn, None, c.data_ptr(), N, M, N, K,
                            acc, 0, ws.data_ptr(), ws.numel(), flags, torch.cuda.current_stream().cuda_stream)
    _cabi.check(st, "gemm")
    torch.cuda.synchronize()
    d = (c.cpu().double() - ref)
    err = float(d.abs().max() / ref.abs().max())
    # where are the bad entries?
    bad = (d.abs() > 1e-4 * ref.abs().max())
    info = ""
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        info = f" bad rows {rows.numel()} [{int(rows.min())}..{int(rows.max())}] cols {cols.numel()} [{int(cols.min())}..{int(cols.max())}]"
    return err, info


if __name__ == "__main__":
    M = 19200
    cases = []
    for (kl, nl) in [(256, 64), (64, 256), (256, 256), (80, 256), (64, 6)]:
        cases.append(("fwd", M, nl, kl, 0, 0, 0))
        cases.append(("dx ", M, kl, nl, 0, 1, 0))
        cases.append(("dW ", nl, kl, M, 1, 1, 1))
    cases += [("dx64", 64, 256, 64, 0, 1, 0), ("dx64", 64, 64, 256, 0, 1, 0), ("fwd", 64, 64, 256, 0, 0, 0),
              ("dW64", 64, 256, 64, 1, 1, 1), ("dx ", 300, 64, 256, 0, 1, 0), ("dx ", 1024, 64, 256, 0, 1, 0),
              ("dx ", 1024, 64, 128, 0, 1, 0), ("dx ", 1024, 96, 256, 0, 1, 0), ("dx ", 1024, 32, 256, 0, 1, 0),
              ("fwd", 1024, 64, 256, 0, 0, 0), ("x", 1024, 64, 256, 1, 0, 0), ("x", 1024, 64, 256, 1, 1, 0)]
    for tag, m, n, k, a_mn, b_mn, acc in cases:
        e_tc, info = run(m, n, k, a_mn, b_mn, acc, 0)
        e_si, _ = run(m, n, k, a_mn, b_mn, acc, _cabi.F_SIMT_GEMM)
        flag = "  <<<<<< BAD" if e_tc > 1e-5 else ""
        print(f"{tag} M={m:6d} N={n:4d} K={k:6d} a_mn={a_mn} b_mn={b_mn} acc={acc}: tc {e_tc:.2e} simt {e_si:.2e}{flag}{info}", flush=True)
    # sparse dy (only the last frame of every sequence carries gradient), as in the decoder's last block
    e_tc, info = run(M, 64, 256, 0, 1, 0, 0, sparse_rows=300)
    print(f"sparse dx M={M} N=64 K=256: tc {e_tc:.2e}{info}")
