"""GPU parity of the mirrored model code against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz): same state_dict keys, outputs, losses and gradients."""
import pytest
import torch

from conftest import load_golden, rel_err, rel_l2

pytestmark = pytest.mark.gpu

OUT_TOL, GRAD_TOL = 1e-5, 1e-4


def _cuda(sd):
    return {k: v.cuda() for k, v in sd.items()}


def _check_grads(model, grads, prefix=""):
    for name, p in model.named_parameters():
        want = grads[prefix + name]
        assert p.grad is not None, name
        assert rel_l2(p.grad.cpu(), want) <= GRAD_TOL, name


def _layerd(bi):
    from multimodalreactiongeneration_b200.mr_gen.model.utils.lstm_block import LSTMLayerd
    if bi:
        return LSTMLayerd(input_size=32, lstm_hidden_size=16, affine_hidden_size=32, bottleneck_size=8,
                          num_layers=2, num_layers_per_block=1, output_size=32, dropout=0.0, bidirectional=True,
                          use_layer_norm=True, use_relu=True, use_mixing=True, use_residual=True)
    return LSTMLayerd(input_size=32, lstm_hidden_size=32, affine_hidden_size=32, bottleneck_size=8,
                      num_layers=2, num_layers_per_block=1, output_size=32, dropout=0.0, bidirectional=False,
                      use_layer_norm=True, use_mixing=False, use_residual=True, use_feed_forward=False)


@pytest.mark.parametrize("name,bi", [("lstm_layerd_uni", False), ("lstm_layerd_bi_mix_ffn", True)])
def test_lstm_layerd_matches_reference(name, bi):
    sd, ins, outs, grads, _ = load_golden(name)
    m = _layerd(bi)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    x = ins["x"].cuda().requires_grad_(True)
    y, hxs = m(x)
    assert hxs is None  # quirk Q2: the input states are handed back
    (y * ins["w"].cuda()).sum().backward()
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    assert rel_l2(x.grad.cpu(), grads["x"]) <= GRAD_TOL
    _check_grads(m, grads)


def test_layerd_last_step_equals_the_slice_of_the_full_forward():
    """LSTMLayerd.forward_last_step (what MotionDecoder consumes) against forward(x)[0][:, -1:]: output and every gradient;
    with an active dropout it must fall back to the full forward (same RNG stream as the reference)."""
    from multimodalreactiongeneration_b200.mr_gen.model.utils.lstm_block import LSTMLayerd
    torch.manual_seed(3)
    m = LSTMLayerd(input_size=256, lstm_hidden_size=256, affine_hidden_size=256, bottleneck_size=64, num_layers=2,
                   output_size=256, bidirectional=False, use_mixing=True).cuda()
    x = torch.randn(5, 37, 256, device="cuda")
    w = torch.randn(5, 1, 256, device="cuda")
    xa = x.clone().requires_grad_(True)
    (m(xa)[0][:, -1:] * w).sum().backward()
    ga = {n: p.grad.clone() for n, p in m.named_parameters()}
    ya = m(xa)[0][:, -1:].detach()
    for p in m.parameters():
        p.grad = None
    xb = x.clone().requires_grad_(True)
    yb = m.forward_last_step(xb)
    (yb * w).sum().backward()
    assert rel_err(yb.detach().cpu(), ya.cpu().double()) <= OUT_TOL
    assert rel_l2(xb.grad.cpu(), xa.grad.cpu().double()) <= GRAD_TOL
    for n, p in m.named_parameters():
        assert rel_l2(p.grad.cpu(), ga[n].cpu().double()) <= GRAD_TOL, n
    md = LSTMLayerd(input_size=256, lstm_hidden_size=256, affine_hidden_size=256, num_layers=1, output_size=256,
                    bidirectional=False, use_mixing=True, dropout=0.5).cuda().train()
    assert not md.lstm_layered[0].dropout_inactive() and md.eval().lstm_layered[0].dropout_inactive()


def test_lstm_sampler_matches_reference():
    from multimodalreactiongeneration_b200.mr_gen.model.utils import LSTMSampler
    sd, ins, outs, _, _ = load_golden("lstm_sampler")
    m = LSTMSampler(16, 2, 0.0, 4)
    m.load_state_dict(sd)
    m = m.cuda()
    x = ins["x"].cuda()
    with torch.no_grad():
        y1, hx = m(x[:, :8])
        y2, hx2 = m(x[:, 8:], hx)
    assert rel_err(y1.cpu(), outs["y1"]) <= OUT_TOL
    assert rel_err(y2.cpu(), outs["y2"]) <= OUT_TOL
    assert rel_err(hx2[0].cpu(), outs["h"]) <= OUT_TOL
    assert rel_err(hx2[1].cpu(), outs["c"]) <= OUT_TOL


def _lws(scheduled=False):
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from multimodalreactiongeneration_b200.mr_gen.utils.config import DictConfig
    model = DictConfig(
        nmels=9, delta_order=0, use_centroid=True, use_angle=True, sampler_hidden_size=16,
        sampler_num_layers=2, sampler_dropout_rate=0, sampling_rate=16000, shift=160, fps=25, pred_fps=50.0,
        hidden_size=32, bottleneck_size=8, num_layers=2, num_lstm=1, dropout_rate=0.0, use_layer_norm=True,
        use_relu=True, use_mixing=False, use_residual=True, delta_loss_scale=1, loss_type="huber",
        loss_reduction="mean", huber_delta=1.0, smoothl1_beta=1.0, use_scheduled_sampling=scheduled, max_epochs=6)
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100,
                       momentum=0.9)
    metrics = DictConfig(use_centroid=True, use_angle=True, delta_order=0)
    return LSTMwithSample(model, optim, metrics)


def _lws_batch(ins):
    names = ["acoustic", "motion_p", "motion_s", "lead_a", "lead_p", "lead_s", "target"]
    return [(ins[n].cuda(), None) for n in names]


def test_lstm_with_sample_forward_and_teacher_forced_step():
    sd, ins, outs, grads, meta = load_golden("lstm_with_sample")
    m = _lws()
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    assert m.ratio == int(meta["ratio"])
    batch = _lws_batch(ins)
    y, (lead_len, _, _), (hx_s, hxs) = m.forward(*batch[:-1])
    assert hxs is None and lead_len == int(meta["lead_len"])
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    assert rel_err(hx_s[0].cpu(), outs["hs"]) <= OUT_TOL
    assert rel_err(hx_s[1].cpu(), outs["cs"]) <= OUT_TOL
    loss = m.training_step(batch)["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-4 * abs(float(outs["loss"]))
    _check_grads(m, grads)


def test_lstm_with_sample_rollout_modes():
    sd, ins, outs, _, _ = load_golden("lstm_with_sample")
    m = _lws()
    m.load_state_dict(sd)
    m = m.cuda()
    batch = _lws_batch(ins)
    with torch.no_grad():
        tf, _ = m.prediction(batch)
        free, _ = m.prediction(batch, full_generation=True)
        ss, _ = m.prediction(batch, use_scheduled_sampling=True, sampling_mask=ins["mask_ss"].bool())
    assert rel_err(tf.cpu(), outs["pred_tf"]) <= OUT_TOL
    assert rel_err(free.cpu(), outs["pred_free"]) <= 5e-5  # 7 free-running steps compound fp32 rounding
    assert rel_err(ss.cpu(), outs["pred_ss"]) <= 5e-5


def test_lstm_with_sample_scheduled_sampling_training_step():
    """Gradient through the fed-back predictions (Q6) with the reference's own mask draw (Q4)."""
    sd, ins, outs, grads, _ = load_golden("lstm_with_sample")
    m = _lws(scheduled=True)
    m.load_state_dict(sd)
    m = m.cuda()
    m.current_epoch = 3
    torch.manual_seed(77)  # same CPU draw as the fixture: torch.rand(T) < 3/6
    loss = m.training_step(_lws_batch(ins))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss_ss"])) <= 1e-4 * abs(float(outs["loss_ss"]))
    _check_grads(m, grads, prefix="ss/")


def test_simple_lstm_matches_reference():
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.utils.config import DictConfig
    sd, ins, outs, grads, meta = load_golden("simple_lstm")
    cfg = DictConfig(
        acostic_feat_size=10, motion_feat_size=6, motion_num_lstm=1, acostic_num_lstm=1, acostic_num_layers=2,
        motion_num_layers=2, acostic_lstm_size=32, motion_lstm_size=32, acostic_affine_size=32,
        motion_affine_size=32, acostic_output_size=32, motion_output_size=32, att_heads=int(meta["heads"]),
        att_num_layers=2, att_use_residual=True, att_use_layer_norm=True, dropout_rate=0, output_size=6,
        bidirectional=False, use_layer_norm=True, use_relu=True, use_mixing=True, use_residual=True,
        decoder_num_layers=2, decoder_num_lstm=1, decoder_lstm_size=32, decoder_affine_size=32,
        decoder_bottleneck_size=8, decoder_output_size=32, decoder_mapping_size=8, decoder_bidirectional=False,
        decoder_use_layer_norm=True, decoder_use_relu=True, decoder_use_mixing=True, decoder_use_residual=True,
        delta_loss_scale=1, all_static=False)
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100,
                       momentum=0.9)
    metrics = DictConfig(use_centroid=True, use_angle=True, delta_order=0)
    m = SimpleLSTM(cfg, optim, metrics)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    a, mo, tgt = ins["acoustic"].cuda(), ins["motion"].cuda(), ins["target"].cuda()
    y = m(a, mo)
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    loss = m.training_step((a, mo, tgt))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-4 * abs(float(outs["loss"]))
    _check_grads(m, grads)


def test_lstm_mixer_layerd_matches_reference():
    """lstmformer's token mixer (mixer_block.py:762-843): 2 residual LSTM blocks + single-Linear FFN."""
    from multimodalreactiongeneration_b200.mr_gen.model.utils.mixer_block import LSTMMixerLayerd
    sd, ins, outs, grads, meta = load_golden("lstm_mixer_layerd")
    m = LSTMMixerLayerd(hidden_size=32, num_layerd=2, residual=True, residual_layer_norm=True,
                        nonlinearity="none", device=torch.device("cpu"))
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    x = ins["x"].cuda().requires_grad_(True)
    y, hx, other = m(x)
    assert hx is None and bool(meta["hx_is_none"]) and other == (None,)  # quirk Q3
    (y * ins["w"].cuda()).sum().backward()
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    assert rel_l2(x.grad.cpu(), grads["x"]) <= GRAD_TOL
    _check_grads(m, grads)


def test_wavefront_rollout_equals_stepwise_loop_with_philox_masks():
    """The re-scheduled device-resident rollout vs the reference's step-by-step loop, at the real hidden
    sizes (cluster kernels), with per-sample Philox masks; forward, loss and every gradient."""
    from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import (
        LSTMwithSample, philox_sampling_mask)
    from oracle import philox
    B, T, lead, ratio = 5, 12, 3, 2
    g = torch.Generator().manual_seed(9)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    batch = [(r(B, T * ratio, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead * ratio, 80), None),
             (r(B, lead, 6), None), (r(B, lead, 6), None), (r(B, T, 6), None)]
    results = {}
    for mode in ("stepwise", "wavefront"):
        torch.manual_seed(0)
        m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=10, ratio=ratio, seed=4321)).cuda()
        m.rollout = mode
        m.current_epoch = 5
        mask = philox_sampling_mask(4321, 0, 0.5, T, B, "cuda")
        assert (mask.cpu().numpy() == philox.sampling_mask(4321, 0, 0.5, T, B)).all()  # bit-exact masks
        assert 0 < int(mask.sum()) < T * B
        loss = m.training_step(batch)["loss"]          # draws the same Philox mask internally (seed, offset 0)
        assert m.sampling_offset == T
        loss.backward()
        with torch.no_grad():
            pred, _ = m.prediction(batch, use_scheduled_sampling=True, sampling_mask=mask)
            free, _ = m.prediction(batch, full_generation=True)
        results[mode] = (loss.detach(), pred, free, {k: p.grad.clone() for k, p in m.named_parameters()})
    (l0, p0, f0, g0), (l1, p1, f1, g1) = results["stepwise"], results["wavefront"]
    assert abs(float(l0) - float(l1)) <= 1e-5 * abs(float(l0))
    assert rel_err(p1, p0) <= OUT_TOL
    assert rel_err(f1, f0) <= 5e-5
    for k in g0:
        assert rel_l2(g1[k], g0[k]) <= GRAD_TOL, k


def test_streaming_generator_matches_free_running_prediction():
    """Frame-by-frame generation with carried sampler state == prediction(full_generation=True)."""
    from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.streaming import StreamingGenerator
    B, T, lead, ratio = 6, 9, 2, 2
    torch.manual_seed(1)
    m = LSTMwithSample(*lstm_with_sampling_cfg(scheduled=False, ratio=ratio)).cuda().eval()
    g = torch.Generator().manual_seed(2)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    batch = [(r(B, T * ratio, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead * ratio, 80), None),
             (r(B, lead, 6), None), (r(B, lead, 6), None), (r(B, T, 6), None)]
    with torch.no_grad():
        want, _ = m.prediction(batch, full_generation=True)
    for graph in (False, True):
        gen = StreamingGenerator(m, B, use_cuda_graph=graph)
        gen.reset(batch[3][0])
        got = []
        for t in range(T):
            prev = batch[2][0][:, 0] if t == 0 else None       # the rollout starts from motion_s[0]
            got.append(gen.step(batch[0][0][:, t * ratio:(t + 1) * ratio], batch[1][0][:, t], prev).clone())
        got = torch.stack(got, dim=1)
        assert rel_err(got, want) <= 5e-5, graph


def test_flat_adamw_matches_torch_adamw():
    """Trainer's fused optimizer step (mrg_adamw_flat over the flat buckets) vs torch.optim.AdamW, 5 steps,
    including a learning-rate change between steps (what CosineAnnealingLR does to param_groups[0]["lr"])."""
    import copy
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import FlatAdamW, FlatGradBucket
    torch.manual_seed(7)
    net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.Tanh(), torch.nn.Linear(19, 5)).cuda()
    ref = copy.deepcopy(net)
    bucket = FlatGradBucket(net)
    assert bucket.flat_params is not None and all(p.data_ptr() % 256 == 0 for p in net.parameters())
    opt = torch.optim.AdamW(net.parameters(), lr=3e-3, weight_decay=1e-2)
    assert FlatAdamW.applicable(bucket, opt)
    flat = FlatAdamW(bucket, opt)
    ropt = torch.optim.AdamW(ref.parameters(), lr=3e-3, weight_decay=1e-2)
    for step in range(5):
        x = torch.randn(11, 37, device="cuda")
        if step == 3:
            opt.param_groups[0]["lr"] = ropt.param_groups[0]["lr"] = 1e-3
        net(x).square().sum().backward()
        flat.step(grad_scale=0.5, zero_grad=True)       # grads were "summed over 2 ranks"
        ropt.zero_grad()
        (0.5 * ref(x).square().sum()).backward()
        ropt.step()
        assert float(bucket.flat.abs().max()) == 0.0     # cleared by the step
    for a, b in zip(net.parameters(), ref.parameters()):
        assert rel_err(a, b) <= 2e-6


def test_trainer_uses_fused_optimizer_and_keeps_state_dict_keys():
    from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer
    torch.manual_seed(0)
    model = SimpleLSTM(*simple_lstm_cfg()).cuda()
    keys = list(model.state_dict().keys())
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tr = Trainer(model)
    assert tr.flat_opt is not None
    assert list(model.state_dict().keys()) == keys
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k])                 # re-homing the parameters does not change them
    g = torch.Generator().manual_seed(3)
    batch = (torch.randn(4, 20, 80, generator=g).cuda(), torch.randn(4, 20, 6, generator=g).cuda(),
             torch.randn(4, 1, 6, generator=g).cuda())
    tr.optimizer.param_groups[0]["lr"] = 1e-3
    l0 = float(tr.train_step(batch))
    for _ in range(20):
        l1 = float(tr.train_step(batch))
    assert l1 < l0


def test_checkpoint_resume_restores_fused_optimizer_state(tmp_path):
    """save_checkpoint(trainer=...) / load_checkpoint(trainer=...): a resumed trainer (fresh model object, fresh fused
    AdamW) continues the trajectory of the original one — parameters after the next step are identical."""
    from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer, load_checkpoint, save_checkpoint
    g = torch.Generator().manual_seed(3)
    batch = (torch.randn(4, 20, 80, generator=g).cuda(), torch.randn(4, 20, 6, generator=g).cuda(),
             torch.randn(4, 1, 6, generator=g).cuda())
    torch.manual_seed(0)
    model = SimpleLSTM(*simple_lstm_cfg()).cuda()
    tr = Trainer(model)
    tr.optimizer.param_groups[0]["lr"] = 1e-3
    for _ in range(3):
        tr.train_step(batch)
    path = str(tmp_path / "ckpts" / "simple_lstm" / "last.ckpt")
    save_checkpoint(model, path, epoch=0, global_step=tr.global_step, trainer=tr)
    tr.train_step(batch)
    want = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(123)   # a different initialisation: everything must come from the file
    model2 = SimpleLSTM(*simple_lstm_cfg()).cuda()
    tr2 = Trainer(model2)
    ck = load_checkpoint(model2, path, trainer=tr2)
    assert "optimizer_states" in ck and "flat_adamw" in ck
    assert tr2.optimizer.param_groups[0]["lr"] == 1e-3
    tr2.train_step(batch)
    for k, v in model2.state_dict().items():
        assert torch.equal(v, want[k]) or rel_err(v, want[k]) <= 1e-6, k


@pytest.mark.parametrize("kdim,same_kv", [(None, True), (None, False), (96, True)])
def test_b200_multihead_attention_matches_torch(kdim, same_kv):
    """Projections on the tcgen05 GEMM, SDPA from torch: forward / input / parameter gradients vs nn.MultiheadAttention
    (the reference's cross-modal call: batch_first, need_weights=False, key is value)."""
    from multimodalreactiongeneration_b200 import B200MultiheadAttention
    torch.manual_seed(11)
    E, nh, B, Tq, Tk = 128, 2, 3, 37, 50
    kw = dict(embed_dim=E, num_heads=nh, batch_first=True, kdim=kdim, vdim=kdim)
    ref = torch.nn.MultiheadAttention(**kw).double()
    mine = B200MultiheadAttention(**kw)
    mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mine = mine.cuda()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    q = torch.randn(B, Tq, E, dtype=torch.double)
    k = torch.randn(B, Tk, kdim or E, dtype=torch.double)
    v = k if same_kv else torch.randn(B, Tk, kdim or E, dtype=torch.double)
    wq = torch.randn(B, Tq, E, dtype=torch.double)
    qr, kr = q.clone().requires_grad_(True), k.clone().requires_grad_(True)
    vr = kr if same_kv else v.clone().requires_grad_(True)
    (ref(qr, kr, vr, need_weights=False)[0] * wq).sum().backward()
    qm, km = q.float().cuda().requires_grad_(True), k.float().cuda().requires_grad_(True)
    vm = km if same_kv else v.float().cuda().requires_grad_(True)
    out, w = mine(qm, km, vm, need_weights=False)
    assert w is None
    (out * wq.float().cuda()).sum().backward()
    with torch.no_grad():
        want = ref(q, k, v, need_weights=False)[0]
    assert rel_err(out.detach().cpu(), want) <= 2e-5
    assert rel_l2(qm.grad.cpu(), qr.grad) <= 1e-4
    assert rel_l2(km.grad.cpu(), kr.grad) <= 1e-4
    for (name, pr), pm in zip(ref.named_parameters(), mine.parameters()):
        assert rel_l2(pm.grad.cpu(), pr.grad) <= 1e-4, name


def test_colsum_matches_torch_sum():
    from multimodalreactiongeneration_b200.linear import _colsum
    torch.manual_seed(5)
    for M, N in [(19200, 256), (300, 64), (1, 8), (1000, 1028)]:
        x = torch.randn(M, N, device="cuda")
        assert rel_err(_colsum(x), x.double().sum(0)) <= 1e-5


def test_fused_weight_gradients_equal_autograd_accumulation(monkeypatch):
    """Weight-gradient kernels adding straight into the trainer's flat bucket (MRG_FUSED_WGRAD, default on) give the
    same bucket as autograd's own accumulation of returned gradients."""
    from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import FlatGradBucket
    g = torch.Generator().manual_seed(4)
    batch = (torch.randn(5, 24, 80, generator=g).cuda(), torch.randn(5, 24, 6, generator=g).cuda(),
             torch.randn(5, 1, 6, generator=g).cuda())
    flats = []
    for fused in ("1", "0"):
        monkeypatch.setenv("MRG_FUSED_WGRAD", fused)
        torch.manual_seed(0)
        model = SimpleLSTM(*simple_lstm_cfg()).cuda()
        bucket = FlatGradBucket(model)
        assert all(getattr(p, "_mrg_grad_fused", None) == (fused == "1") for p in bucket.params)
        for _ in range(2):   # two backward passes: accumulation across calls must also agree
            model.training_step(batch)["loss"].backward()
        flats.append(bucket.flat.clone())
    assert float(flats[0].abs().max()) > 0
    assert rel_l2(flats[0], flats[1]) <= 1e-6


# ---------------------------------------------------------------------------------------------------------
# lstmformer (Metaformer): 3 LSTM-mixer stacks + masked cross-modal attention, fixture from the unmodified reference
# ---------------------------------------------------------------------------------------------------------
def _metaformer(scheduled=False, **kw):
    from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
    return Metaformer(*metaformer_cfg(hidden=32, blocks=2, encoder_layers=2, bottleneck=8, heads=4, acoustic=10,
                                      ratio=2, max_epochs=6, scheduled=scheduled, **kw))


def _leaves(o):
    if isinstance(o, dict):
        o = list(o.values())
    if isinstance(o, (list, tuple)):
        return [leaf for v in o for leaf in _leaves(v)]
    return [o]


def test_metaformer_forward_and_teacher_forced_step():
    sd, ins, outs, grads, meta = load_golden("metaformer")
    m = _metaformer()
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    assert m.ratio == int(meta["ratio"])
    y, hxs = m.forward(*_lws_batch(ins)[:-1])
    assert len(hxs) == 2 and all(leaf is None for leaf in _leaves(hxs))   # Q3: no state comes back
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    loss = m.training_step(_lws_batch(ins))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-4 * abs(float(outs["loss"]))
    _check_grads(m, grads)


@pytest.mark.parametrize("rollout", ["wavefront", "stepwise"])
def test_metaformer_rollout_modes(rollout):
    sd, ins, outs, _, _ = load_golden("metaformer")
    m = _metaformer()
    m.load_state_dict(sd)
    m = m.cuda()
    m.rollout = rollout
    with torch.no_grad():
        tf, target = m.prediction(_lws_batch(ins))
        free, _ = m.prediction(_lws_batch(ins), full_generation=True)
        ss, _ = m.prediction(_lws_batch(ins), use_scheduled_sampling=True, sampling_mask=ins["mask_ss"].bool())
    assert target.shape == outs["target_tf"].shape and torch.equal(target.cpu(), outs["target_tf"])  # [T,B,T,P] sic
    assert rel_err(tf.cpu(), outs["pred_tf"]) <= OUT_TOL
    assert rel_err(free.cpu(), outs["pred_free"]) <= 5e-5   # 7 free-running steps compound fp32 rounding
    assert rel_err(ss.cpu(), outs["pred_ss"]) <= 5e-5


def test_metaformer_scheduled_sampling_training_step():
    sd, ins, outs, grads, _ = load_golden("metaformer")
    m = _metaformer(scheduled=True)
    m.load_state_dict(sd)
    m = m.cuda()
    m.current_epoch = 3
    torch.manual_seed(77)   # the fixture's draw: torch.rand(T) < 3/6
    loss = m.training_step(_lws_batch(ins))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss_ss"])) <= 1e-4 * abs(float(outs["loss_ss"]))
    for name, p in m.named_parameters():
        if "ss/" + name in grads:
            assert rel_l2(p.grad.cpu(), grads["ss/" + name]) <= GRAD_TOL, name


def test_metaformer_headline_shape_trains_through_the_trainer():
    """cfg 4 shape at reduced batch (hidden 256, 5 blocks, 5-layer encoders = 15 LSTM mixers, T=300): one Trainer
    step runs, the loss is finite and every parameter receives a gradient through the fused optimizer."""
    from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer
    torch.manual_seed(0)
    m = Metaformer(*metaformer_cfg()).cuda()
    tr = Trainer(m)
    g = torch.Generator().manual_seed(1)
    B, T, lead = 8, 300, 30
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None),
             (r(B, lead, 6), None), (r(B, lead, 6), None), (r(B, T, 6), None)]
    before = tr.bucket.flat_params.clone()
    loss = tr.train_step(batch)
    assert torch.isfinite(loss)
    moved = (tr.bucket.flat_params != before)
    for p, o in zip(tr.bucket.params, tr.bucket.offsets):
        assert bool(moved[o:o + p.numel()].any())


# ---------------------------------------------------------------------------------------------------------
# GRU mixer (lstmformer with emb_mixers: gru): B200GRU against torch.nn.GRU fp64, the mixer stack against the reference
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("I,H,L,bi,B,T,with_hx", [
    (256, 256, 1, False, 64, 20, False),   # lstmformer mixer shape -> cluster-resident kernels (MRG_F_GRU)
    (256, 256, 2, False, 7, 23, True),     # cluster kernels: two layers, carried state, ragged row group
    (128, 256, 1, True, 31, 9, True),      # cluster kernels: both directions in one launch
    (128, 128, 2, False, 16, 33, True),    # cluster kernels, H = 128
    (256, 256, 1, False, 64, 300, False),  # cluster kernels at the benchmark length
    (32, 32, 2, False, 3, 9, True),        # two layers, carried state, ragged row group
    (20, 16, 1, True, 5, 7, True),         # bidirectional
    (256, 256, 1, False, 1, 1, True),      # a single frame (autoregressive step)
    (64, 48, 1, False, 130, 5, False),
])
def test_gru_matches_torch_gru(I, H, L, bi, B, T, with_hx):
    from multimodalreactiongeneration_b200 import B200GRU
    torch.manual_seed(9)
    ref = torch.nn.GRU(I, H, L, batch_first=True, bidirectional=bi).double()
    mine = B200GRU(I, H, L, batch_first=True, bidirectional=bi)
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mine = mine.cuda()
    D = 2 if bi else 1
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, T, I, generator=g, dtype=torch.double)
    h0 = torch.randn(L * D, B, H, generator=g, dtype=torch.double) * 0.5 if with_hx else None
    wy = torch.randn(B, T, D * H, generator=g, dtype=torch.double)
    wh = torch.randn(L * D, B, H, generator=g, dtype=torch.double)
    xr = x.clone().requires_grad_(True)
    hr = None if h0 is None else h0.clone().requires_grad_(True)
    yr, hnr = ref(xr, hr)
    ((yr * wy).sum() + (hnr * wh).sum()).backward()
    xm = x.float().cuda().requires_grad_(True)
    hm = None if h0 is None else h0.float().cuda().requires_grad_(True)
    ym, hnm = mine(xm, hm)
    ((ym * wy.float().cuda()).sum() + (hnm * wh.float().cuda()).sum()).backward()
    assert rel_err(ym.cpu(), yr) <= OUT_TOL and rel_err(hnm.cpu(), hnr) <= OUT_TOL
    assert rel_l2(xm.grad.cpu(), xr.grad) <= GRAD_TOL
    for (name, pr), pm in zip(ref.named_parameters(), mine.parameters()):
        assert rel_l2(pm.grad.cpu(), pr.grad) <= GRAD_TOL, name
    if h0 is not None:
        assert rel_l2(hm.grad.cpu(), hr.grad) <= GRAD_TOL


@pytest.mark.parametrize("mode,B,T,bi", [("tf32", 256, 30, False), ("bf16", 150, 17, True), ("tf32", 128, 40, False),
                                         ("tf32", 20, 15, True), ("bf16", 64, 300, False)])
def test_gru_reduced_precision_tensor_core_recurrence(mode, B, T, bi):
    """GRU variant of the tensor-core recurrent kernels (rec_fwd3 / rec_bwd3_kernel<256, gru>) in the reduced-precision
    modes: fp64 nn.GRU, the modes' stated bound (states 2e-2, gradients 5e-2), carried state, both directions."""
    import multimodalreactiongeneration_b200 as pkg
    from multimodalreactiongeneration_b200 import B200GRU, _cabi
    H = 256
    torch.manual_seed(19)
    ref = torch.nn.GRU(H, H, 1, batch_first=True, bidirectional=bi).double()
    mine = B200GRU(H, H, 1, batch_first=True, bidirectional=bi)
    mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mine = mine.cuda()
    D = 2 if bi else 1
    g = torch.Generator().manual_seed(20)
    x = torch.randn(B, T, H, generator=g, dtype=torch.double)
    h0 = torch.randn(D, B, H, generator=g, dtype=torch.double) * 0.5
    wy = torch.randn(B, T, D * H, generator=g, dtype=torch.double)
    wh = torch.randn(D, B, H, generator=g, dtype=torch.double)
    xr, hr = x.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    yr, hnr = ref(xr, hr)
    ((yr * wy).sum() + (hnr * wh).sum()).backward()
    _cabi.profile_enable(True)
    try:
        pkg.set_precision(mode)
        xm, hm = x.float().cuda().requires_grad_(True), h0.float().cuda().requires_grad_(True)
        ym, hnm = mine(xm, hm)
        ((ym * wy.float().cuda()).sum() + (hnm * wh.float().cuda()).sum()).backward()
        torch.cuda.synchronize()
        names = _cabi.profile_kernel_name("rec_fwd"), _cabi.profile_kernel_name("rec_bwd")
    finally:
        pkg.set_precision("fp32")
        _cabi.profile_read()
        _cabi.profile_enable(False)
    assert "rec_fwd3" in names[0] and "gru" in names[0], names
    assert "rec_bwd3" in names[1] and "gru" in names[1], names
    assert rel_err(ym.cpu(), yr) <= 2e-2 and rel_err(hnm.cpu(), hnr) <= 2e-2
    assert rel_l2(xm.grad.cpu(), xr.grad) <= 5e-2 and rel_l2(hm.grad.cpu(), hr.grad) <= 5e-2
    for (name, pr), pm in zip(ref.named_parameters(), mine.parameters()):
        assert rel_l2(pm.grad.cpu(), pr.grad) <= 5e-2, name


def test_gru_inference_under_no_grad_matches_torch():
    """B200GRU on the cluster kernels under torch.no_grad() (inference flag: no reserve for a backward): fp64 nn.GRU."""
    from multimodalreactiongeneration_b200 import B200GRU
    torch.manual_seed(23)
    ref = torch.nn.GRU(256, 256, 2, batch_first=True).double()
    mine = B200GRU(256, 256, 2, batch_first=True)
    mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mine = mine.cuda()
    x = torch.randn(9, 31, 256, dtype=torch.double)
    h0 = torch.randn(2, 9, 256, dtype=torch.double) * 0.5
    with torch.no_grad():
        yr, hr = ref(x, h0)
        ym, hm = mine(x.float().cuda(), h0.float().cuda())
    assert rel_err(ym.cpu(), yr) <= OUT_TOL and rel_err(hm.cpu(), hr) <= OUT_TOL


def test_gru_mixer_layerd_matches_reference():
    from multimodalreactiongeneration_b200.mr_gen.model.utils.mixer_block import GRUMixerLayerd
    sd, ins, outs, grads, meta = load_golden("gru_mixer_layerd")
    m = GRUMixerLayerd(hidden_size=32, num_layerd=2, residual=True, residual_layer_norm=True, nonlinearity="none",
                       device=torch.device("cpu"))
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m = m.cuda()
    x = ins["x"].cuda().requires_grad_(True)
    y, hx, _ = m(x)
    assert hx is None and bool(meta["hx_is_none"])
    (y * ins["w"].cuda()).sum().backward()
    assert rel_err(y.cpu(), outs["y"]) <= OUT_TOL
    assert rel_l2(x.grad.cpu(), grads["x"]) <= GRAD_TOL
    _check_grads(m, grads)


def test_metaformer_with_gru_mixers_runs_a_training_step():
    """config_gru.yaml's mixer choice: every embedding stack on B200GRU, one Trainer step, all parameters move."""
    from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer
    torch.manual_seed(0)
    m = Metaformer(*metaformer_cfg(hidden=64, blocks=2, encoder_layers=2, bottleneck=16, heads=2,
                                   mixers=("gru", "gru", "gru"))).cuda()
    tr = Trainer(m)
    g = torch.Generator().manual_seed(1)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    B, T, lead = 4, 12, 3
    batch = [(r(B, T, 80), None), (r(B, T, 6), None), (r(B, T, 6), None), (r(B, lead, 80), None),
             (r(B, lead, 6), None), (r(B, lead, 6), None), (r(B, T, 6), None)]
    before = tr.bucket.flat_params.clone()
    assert torch.isfinite(tr.train_step(batch))
    moved = tr.bucket.flat_params != before
    for p, o in zip(tr.bucket.params, tr.bucket.offsets):
        assert bool(moved[o:o + p.numel()].any())
