"""Times one B200LSTM layer stack fwd / fwd+bwd with CUDA events (developer tool, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalreactiongeneration_b200 import B200LSTM

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

def main():
    import ctypes
    from multimodalreactiongeneration_b200 import _cabi
    L = _cabi.lib()
    v = [ctypes.c_int() for _ in range(5)]
    L.mrg_device_info(*[ctypes.byref(x) for x in v])
    print("sm_count, max_clusters_h256, max_clusters_h128, cc =", [x.value for x in v], flush=True)
    for (B, T, I, H, L, bi) in [(64, 300, 256, 256, 1, False), (64, 300, 256, 256, 2, False), (64, 300, 128, 128, 2, False),
                                (256, 300, 256, 256, 1, False), (64, 300, 256, 128, 1, True), (8, 300, 256, 256, 1, False)]:
        m = B200LSTM(I, H, L, batch_first=True, bidirectional=bi).cuda()
        x = torch.randn(B, T, I, device="cuda", requires_grad=True)
        def fwd():
            with torch.no_grad(): m(x)
        def fb():
            y, _ = m(x); y.sum().backward()
        tf, tfb = timeit(fwd), timeit(fb)
        print(f"B={B} T={T} I={I} H={H} L={L} bi={bi}: fwd {tf:.3f} ms ({tf*1e3/(T*L):.2f} us/step/layer)  fwd+bwd {tfb:.3f} ms", flush=True)
        if os.environ.get("MRG_COMPARE_CUDNN") == "1":
            r = torch.nn.LSTM(I, H, L, batch_first=True, bidirectional=bi).cuda()
            def rf():
                with torch.no_grad(): r(x)
            def rfb():
                y, _ = r(x); y.sum().backward()
            print(f"    torch/cuDNN: fwd {timeit(rf):.3f} ms  fwd+bwd {timeit(rfb):.3f} ms", flush=True)

if __name__ == "__main__":
    main()
