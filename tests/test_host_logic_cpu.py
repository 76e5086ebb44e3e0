"""CPU tests of the HOST logic of the mirrored models (state plumbing, masks, padding, wavefront re-scheduling of the
rollouts, loss assembly) against the fixtures generated from the unmodified reference.

The product modules have no CPU path (B200LSTM / B200GRU / B200Linear raise on CPU tensors — see
test_host_cpu.py::test_no_cpu_fallback).  Here, and only here, their ``forward`` is monkeypatched to the torch parent
class's CPU implementation so that everything AROUND the kernels can be exercised without a GPU; the kernels themselves
are checked on the B200 (tests/test_*_gpu.py)."""
import pytest
import torch
from torch import nn

from conftest import load_golden, rel_err, rel_l2


@pytest.fixture()
def torch_arithmetic(monkeypatch):
    from multimodalreactiongeneration_b200.gru import B200GRU
    from multimodalreactiongeneration_b200.linear import B200Linear
    from multimodalreactiongeneration_b200.lstm import B200LSTM
    monkeypatch.setattr(B200LSTM, "forward", nn.LSTM.forward)
    monkeypatch.setattr(B200GRU, "forward", nn.GRU.forward)
    monkeypatch.setattr(B200Linear, "forward", nn.Linear.forward)
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling import lstm_with_sample as lws
    monkeypatch.setattr(lws, "_linear", nn.functional.linear)
    from multimodalreactiongeneration_b200.attention import B200MultiheadAttention
    monkeypatch.setattr(B200MultiheadAttention, "forward", nn.MultiheadAttention.forward)


NAMES = ["acoustic", "motion_p", "motion_s", "lead_a", "lead_p", "lead_s", "target"]


def _batch(ins):
    return [(ins[n].clone(), None) for n in NAMES]


def _grads_match(model, grads, prefix="", tol=1e-5):
    for name, p in model.named_parameters():
        if prefix + name in grads:
            assert rel_l2(p.grad, grads[prefix + name]) <= tol, name


def test_metaformer_host_logic_matches_reference(torch_arithmetic):
    from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
    sd, ins, outs, grads, _ = load_golden("metaformer")
    m = Metaformer(*metaformer_cfg(hidden=32, blocks=2, encoder_layers=2, bottleneck=8, heads=4, acoustic=10, ratio=2,
                                   max_epochs=6))
    m.load_state_dict(sd)
    y, hxs = m.forward(*_batch(ins)[:-1])
    assert rel_err(y, outs["y"]) <= 1e-6 and len(hxs) == 2
    loss = m.training_step(_batch(ins))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-6 * abs(float(outs["loss"]))
    _grads_match(m, grads)
    with torch.no_grad():
        for mode in ("wavefront", "stepwise"):     # the re-scheduled rollout must equal the reference's time loop
            m.rollout = mode
            tf, target = m.prediction(_batch(ins))
            free, _ = m.prediction(_batch(ins), full_generation=True)
            ss, _ = m.prediction(_batch(ins), sampling_mask=ins["mask_ss"].bool())
            assert torch.equal(target, outs["target_tf"])
            assert rel_err(tf, outs["pred_tf"]) <= 1e-6 and rel_err(free, outs["pred_free"]) <= 1e-6
            assert rel_err(ss, outs["pred_ss"]) <= 1e-6
    m.zero_grad()
    m.use_scheduled_sampling, m.current_epoch, m.rollout = True, 3, "wavefront"
    torch.manual_seed(77)
    loss_ss = m.training_step(_batch(ins))["loss"]
    loss_ss.backward()
    assert abs(float(loss_ss) - float(outs["loss_ss"])) <= 1e-6 * abs(float(outs["loss_ss"]))
    _grads_match(m, grads, prefix="ss/")


def test_lstm_with_sample_host_logic_matches_reference(torch_arithmetic):
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from multimodalreactiongeneration_b200.mr_gen.utils.config import DictConfig
    sd, ins, outs, grads, meta = load_golden("lstm_with_sample")
    model = DictConfig(
        nmels=9, delta_order=0, use_centroid=True, use_angle=True, sampler_hidden_size=16, sampler_num_layers=2,
        sampler_dropout_rate=0, sampling_rate=16000, shift=160, fps=25, pred_fps=50.0, hidden_size=32,
        bottleneck_size=8, num_layers=2, num_lstm=1, dropout_rate=0.0, use_layer_norm=True, use_relu=True,
        use_mixing=False, use_residual=True, delta_loss_scale=1, loss_type="huber", loss_reduction="mean",
        huber_delta=1.0, smoothl1_beta=1.0, use_scheduled_sampling=False, max_epochs=6)
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100, momentum=0.9)
    m = LSTMwithSample(model, optim, DictConfig(use_centroid=True, use_angle=True, delta_order=0))
    m.use_device = "cpu"
    m.load_state_dict(sd)
    y, (lead_len, _, _), (hx_s, hxs) = m.forward(*_batch(ins)[:-1])
    assert hxs is None and lead_len == int(meta["lead_len"]) and rel_err(y, outs["y"]) <= 1e-6
    loss = m.training_step(_batch(ins))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-6 * abs(float(outs["loss"]))
    _grads_match(m, grads)
    with torch.no_grad():
        for mode in ("wavefront", "stepwise"):
            m.rollout = mode
            tf, _ = m.prediction(_batch(ins))
            free, _ = m.prediction(_batch(ins), full_generation=True)
            ss, _ = m.prediction(_batch(ins), use_scheduled_sampling=True, sampling_mask=ins["mask_ss"].bool())
            assert rel_err(tf, outs["pred_tf"]) <= 1e-6 and rel_err(free, outs["pred_free"]) <= 1e-6
            assert rel_err(ss, outs["pred_ss"]) <= 1e-6
        # per-sample masks (the Philox extension): a [T, B] mask whose columns are equal must reproduce the shared one
        T, B = ins["mask_ss"].numel(), ins["motion_p"].shape[0]
        per_sample = ins["mask_ss"].bool().view(T, 1).expand(T, B).contiguous()
        for mode in ("wavefront", "stepwise"):
            m.rollout = mode
            ss2, _ = m.prediction(_batch(ins), use_scheduled_sampling=True, sampling_mask=per_sample)
            assert rel_err(ss2, outs["pred_ss"]) <= 1e-6


def test_simple_lstm_and_mixer_stacks_host_logic_match_reference(torch_arithmetic):
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.model.utils.mixer_block import GRUMixerLayerd, LSTMMixerLayerd
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
    optim = DictConfig(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True, max_epochs=100, momentum=0.9)
    m = SimpleLSTM(cfg, optim, DictConfig(use_centroid=True, use_angle=True, delta_order=0))
    m.load_state_dict(sd)
    y = m(ins["acoustic"], ins["motion"])           # the documented Q1 unwrap makes this run at all
    assert rel_err(y, outs["y"]) <= 1e-6
    loss = m.training_step((ins["acoustic"], ins["motion"], ins["target"]))["loss"]
    loss.backward()
    assert abs(float(loss) - float(outs["loss"])) <= 1e-6 * abs(float(outs["loss"]))
    _grads_match(m, grads)
    for name, cls in (("gru_mixer_layerd", GRUMixerLayerd), ("lstm_mixer_layerd", LSTMMixerLayerd)):
        sd, ins, outs, grads, _ = load_golden(name)
        m = cls(hidden_size=32, num_layerd=2, residual=True, residual_layer_norm=True, nonlinearity="none",
                device=torch.device("cpu"))
        m.load_state_dict(sd)
        x = ins["x"].clone().requires_grad_(True)
        y, hx, _ = m(x)
        (y * ins["w"]).sum().backward()
        assert hx is None and rel_err(y, outs["y"]) <= 1e-6 and rel_l2(x.grad, grads["x"]) <= 1e-5
        _grads_match(m, grads)


@pytest.mark.parametrize("bi", [False, True])
def test_layerd_last_step_host_logic(torch_arithmetic, bi):
    """LSTMLayerd.forward_last_step (MotionDecoder's path) against the slice of the full forward, host logic on CPU: output,
    input gradient and every parameter gradient; one direction takes the kept frame from h_n, two directions from the slice;
    an active dropout falls back to the full forward (same RNG stream: identical draws under a fixed seed)."""
    from multimodalreactiongeneration_b200.mr_gen.model.utils.lstm_block import LSTMLayerd
    torch.manual_seed(5)
    m = LSTMLayerd(input_size=16, lstm_hidden_size=8 if bi else 16, affine_hidden_size=16, bottleneck_size=4, num_layers=2,
                   output_size=16, bidirectional=bi, use_mixing=True).double()
    x = torch.randn(3, 11, 16, dtype=torch.double)
    w = torch.randn(3, 1, 16, dtype=torch.double)
    xa = x.clone().requires_grad_(True)
    ya = m(xa)[0][:, -1:]
    (ya * w).sum().backward()
    ga = {n: p.grad.clone() for n, p in m.named_parameters()}
    for p in m.parameters():
        p.grad = None
    xb = x.clone().requires_grad_(True)
    yb = m.forward_last_step(xb)
    (yb * w).sum().backward()
    assert yb.shape == ya.shape and rel_err(yb.detach(), ya.detach()) <= 1e-12
    assert rel_l2(xb.grad, xa.grad) <= 1e-12
    for n, p in m.named_parameters():
        assert rel_l2(p.grad, ga[n]) <= 1e-12, n
    md = LSTMLayerd(input_size=16, lstm_hidden_size=16, affine_hidden_size=16, bottleneck_size=4, num_layers=2,
                    output_size=16, bidirectional=False, use_mixing=True, dropout=0.3).double().train()
    torch.manual_seed(9)
    full = md(x)[0][:, -1:]
    torch.manual_seed(9)
    assert torch.equal(md.forward_last_step(x), full)
