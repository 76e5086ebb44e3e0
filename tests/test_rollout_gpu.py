"""GPU parity of the persistent rollout kernels (mrg_rollout_forward / _backward, csrc/mrg_rollout.cu) against an
fp64 restatement of one step of the reference's ``generate_one_step`` chain
(mr_gen/model/lstm_with_sampling/lstm_with_sample.py:379-433: feedback select -> feature projection -> zero-state
LSTM blocks + residual LayerNorm (lstm_block.py:101-107) -> bottleneck FFN), forward and every gradient.
Model-level parity against fixtures of the unmodified reference is in test_models_gpu.py (those models now run through
this kernel) and against the oracle's step-by-step loop at H=256 in test_headline_shapes_gpu.py."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, rel_l2

pytestmark = pytest.mark.gpu


def _ref_rollout(base, gt_prev, mask, w_prev, layers, w1, b1, w2, b2, relu, eps):
    T, B, H = base.shape
    preds, prev = [], gt_prev[0]
    for t in range(T):
        x = base[t] + prev @ w_prev.T
        for (w_ih, _w_hh, b_ih, b_hh, g, b) in layers:
            pre = x @ w_ih.T + b_ih + b_hh
            i, _, gg, o = pre.chunk(4, dim=-1)          # zero state: the forget gate multiplies c_0 = 0
            h = torch.sigmoid(o) * torch.tanh(torch.sigmoid(i) * torch.tanh(gg))
            x = F.layer_norm(h + x, (H,), g, b, eps)
        f = x @ w1.T + b1
        if relu:
            f = torch.relu(f)
        y = f @ w2.T + b2
        preds.append(y)
        if t + 1 < T:
            prev = gt_prev[t + 1] if mask is None else torch.where(mask[t].bool().unsqueeze(-1), y, gt_prev[t + 1])
    return torch.stack(preds)


def _make(H, L, P, FB, B, T, seed, mask_mode):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g, dtype=torch.double) * scale)
    k = H ** -0.5
    ins = dict(base=r(T, B, H), gt_prev=r(T, B, P), w_prev=r(H, P, scale=0.3), w1=r(FB, H, scale=k), b1=r(FB, scale=0.1),
               w2=r(P, FB, scale=FB ** -0.5), b2=r(P, scale=0.1))
    layers = [(r(4 * H, H, scale=k), r(4 * H, H, scale=k), r(4 * H, scale=k), r(4 * H, scale=k), 1.0 + r(H, scale=0.2),
               r(H, scale=0.2))
              for _ in range(L)]
    if mask_mode == "none":
        mask = None
    elif mask_mode == "all":
        mask = torch.ones(T, B, dtype=torch.bool)
    else:
        mask = torch.rand(T, B, generator=g) < 0.6
    return ins, layers, mask, r(T, B, P)


CASES = [
    # H, L, P, FB, B, T, mask
    (32, 2, 6, 8, 3, 7, "rand"),       # golden-fixture sizes: one CTA per cluster
    (32, 1, 6, 8, 5, 5, "all"),
    (64, 2, 6, 16, 4, 6, "rand"),      # clusters of 2
    (128, 2, 18, 32, 9, 9, "rand"),    # clusters of 4, 18-d pose (reference default with deltas)
    (256, 2, 6, 64, 5, 12, "rand"),    # cfg 3 widths: clusters of 8
    (256, 2, 6, 64, 64, 6, "rand"),    # cfg 3 batch: 4-5 rows per cluster
    (256, 2, 6, 64, 130, 4, "rand"),   # 9 rows per cluster: two passes
    (256, 1, 18, 64, 7, 5, "none"),    # step-wise teacher forcing (validation's generation phase)
    (256, 2, 6, 64, 3, 40, "all"),     # free running: 40 dependent steps
]


@pytest.mark.parametrize("H,L,P,FB,B,T,mask_mode", CASES)
def test_rollout_kernels_match_fp64_restatement(H, L, P, FB, B, T, mask_mode):
    from multimodalreactiongeneration_b200 import rollout as ro
    assert ro.supported(H, L, P, FB)
    ins, layers, mask, wgt = _make(H, L, P, FB, B, T, 7, mask_mode)
    relu, eps = True, 1e-5
    # fp64 truth
    r_ins = {k: v.clone().requires_grad_(True) for k, v in ins.items()}
    r_layers = [tuple(t.clone().requires_grad_(True) for t in lay) for lay in layers]
    want = _ref_rollout(r_ins["base"], r_ins["gt_prev"], mask, r_ins["w_prev"], r_layers, r_ins["w1"], r_ins["b1"],
                        r_ins["w2"], r_ins["b2"], relu, eps)
    (want * wgt).sum().backward()
    # kernels
    c_ins = {k: v.float().cuda().requires_grad_(True) for k, v in ins.items()}
    c_layers = [tuple(t.float().cuda().requires_grad_(True) for t in lay) for lay in layers]
    got = ro.rollout(c_ins["base"], c_ins["gt_prev"], None if mask is None else mask.cuda(), c_ins["w_prev"], c_layers,
                     c_ins["w1"], c_ins["b1"], c_ins["w2"], c_ins["b2"], relu=relu, eps=eps)
    (got * wgt.float().cuda()).sum().backward()
    torch.cuda.synchronize()
    tol = 1e-5 if mask_mode != "all" else 1e-5 * max(1.0, T / 8)   # free running compounds rounding per step
    assert rel_err(got.detach().cpu(), want.detach()) <= tol
    for k in c_ins:
        assert rel_l2(c_ins[k].grad.cpu(), r_ins[k].grad) <= 1e-4, k
    names = ("w_ih", "w_hh", "b_ih", "b_hh", "ln_weight", "ln_bias")
    for l in range(L):
        for n, c, r in zip(names, c_layers[l], r_layers[l]):
            if n == "w_hh":   # inert (zero state): the node hands back exact zeros, as the reference's graph does
                assert r.grad is None and float(c.grad.abs().max()) == 0.0
                continue
            assert rel_l2(c.grad.cpu(), r.grad) <= 1e-4, (l, n)


def test_rollout_inference_path_and_no_relu():
    """No reserve is written without gradients; relu=False (use_relu: False) is honoured."""
    from multimodalreactiongeneration_b200 import rollout as ro
    H, L, P, FB, B, T = 128, 2, 6, 32, 6, 10
    ins, layers, mask, _ = _make(H, L, P, FB, B, T, 11, "rand")
    want = _ref_rollout(ins["base"], ins["gt_prev"], mask, ins["w_prev"], layers, ins["w1"], ins["b1"], ins["w2"],
                        ins["b2"], False, 1e-5)
    with torch.no_grad():
        got = ro.rollout(ins["base"].float().cuda(), ins["gt_prev"].float().cuda(), mask.cuda(),
                         ins["w_prev"].float().cuda(), [tuple(t.float().cuda() for t in lay) for lay in layers],
                         ins["w1"].float().cuda(), ins["b1"].float().cuda(), ins["w2"].float().cuda(),
                         ins["b2"].float().cuda(), relu=False, eps=1e-5)
    assert rel_err(got.cpu(), want) <= 1e-5


def test_rollout_rejects_unbuilt_shapes():
    from multimodalreactiongeneration_b200 import rollout as ro
    assert not ro.supported(256, 3, 6, 64)     # more blocks than the kernel keeps on chip
    assert not ro.supported(96, 2, 6, 64)      # hidden size
    assert not ro.supported(256, 2, 6, 48)     # bottleneck not a multiple of H/8
    assert ro.supported(256, 2, 18, 64)
