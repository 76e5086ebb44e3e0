"""CPU tests: pin the oracle (oracle/) against torch.nn.LSTM, Random123 KATs and the golden
fixtures generated from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import golden, load_golden, rel_err
from oracle import lstm_numpy, philox, ref_port


@pytest.mark.parametrize("bi,layers", [(False, 1), (False, 2), (True, 1), (True, 2)])
def test_numpy_lstm_matches_torch_forward(bi, layers):
    torch.manual_seed(0)
    m = torch.nn.LSTM(5, 7, layers, batch_first=True, bidirectional=bi).double()
    x = torch.randn(3, 6, 5, dtype=torch.double)
    D = 2 if bi else 1
    hx = (torch.randn(layers * D, 3, 7, dtype=torch.double),
          torch.randn(layers * D, 3, 7, dtype=torch.double))
    y, (h, c) = m(x, hx)
    p = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    yn, (hn, cn) = lstm_numpy.lstm_forward(x.numpy(), p, layers, bi, (hx[0].numpy(), hx[1].numpy()))
    assert np.abs(yn - y.detach().numpy()).max() < 1e-12
    assert np.abs(hn - h.detach().numpy()).max() < 1e-12
    assert np.abs(cn - c.detach().numpy()).max() < 1e-12


@pytest.mark.parametrize("reverse", [False, True])
def test_numpy_lstm_backward_matches_autograd(reverse):
    torch.manual_seed(1)
    T, B, I, H = 5, 3, 4, 6
    m = torch.nn.LSTM(I, H, 1, bidirectional=reverse).double()
    x = torch.randn(T, B, I, dtype=torch.double, requires_grad=True)
    h0 = torch.randn(1, B, H, dtype=torch.double, requires_grad=True)
    c0 = torch.randn(1, B, H, dtype=torch.double, requires_grad=True)
    sfx = "_l0_reverse" if reverse else "_l0"
    w_ih, w_hh = getattr(m, "weight_ih" + sfx), getattr(m, "weight_hh" + sfx)
    b_ih, b_hh = getattr(m, "bias_ih" + sfx), getattr(m, "bias_hh" + sfx)
    # drive one direction through torch by running the bidirectional module and keeping one half
    if reverse:
        hx = (torch.cat([torch.zeros_like(h0), h0]), torch.cat([torch.zeros_like(c0), c0]))
        y_all, (hn_all, cn_all) = m(x, hx)
        y, hn, cn = y_all[..., H:], hn_all[1], cn_all[1]
    else:
        y, (hn_all, cn_all) = m(x, (h0, c0))
        hn, cn = hn_all[0], cn_all[0]
    wy, wh, wc = torch.randn_like(y), torch.randn_like(hn), torch.randn_like(cn)
    ((y * wy).sum() + (hn * wh).sum() + (cn * wc).sum()).backward()
    n = lambda t: t.detach().numpy()
    yn, _, cache = lstm_numpy.lstm_layer_forward(n(x), n(w_ih), n(w_hh), n(b_ih), n(b_hh),
                                                 n(h0[0]), n(c0[0]), reverse)
    dx, dwi, dwh, db, dh0, dc0 = lstm_numpy.lstm_layer_backward(
        n(wy), n(wh), n(wc), n(x), yn, cache, n(w_ih), n(w_hh), n(h0[0]), n(c0[0]), reverse)
    for got, ref in ((dx, x.grad), (dwi, w_ih.grad), (dwh, w_hh.grad), (db, b_ih.grad),
                     (db, b_hh.grad), (dh0, h0.grad[0]), (dc0, c0.grad[0])):
        assert np.abs(got - n(ref)).max() < 1e-10


def test_philox_known_answers():
    # Random123 kat_vectors, "philox4x32 10"
    assert philox.philox4x32_10((0, 0, 0, 0), (0, 0)) == (
        0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    f = 0xFFFFFFFF
    assert philox.philox4x32_10((f, f, f, f), (f, f)) == (
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344),
                                (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_sampling_mask_properties():
    m = philox.sampling_mask(1234, 0, 0.5, 64, 8)
    assert m.shape == (64, 8) and 0.35 < m.mean() < 0.65
    assert not philox.sampling_mask(1234, 0, 0.0, 16, 4).any()
    assert philox.sampling_mask(1234, 0, 1.0, 16, 4).all()
    # offset shifts the stream; shared mode repeats column 0
    assert (philox.sampling_mask(1234, 5, 0.5, 8, 4) == m[5:13, :4]).all()
    s = philox.sampling_mask(1234, 0, 0.5, 16, 4, shared=True)
    assert (s == s[:, :1]).all() and (s[:, 0] == m[:16, 0]).all()


@pytest.mark.parametrize("name", ["lstm_layerd_uni", "lstm_layerd_bi_mix_ffn"])
def test_port_lstm_layerd_matches_reference(name):
    sd, ins, outs, grads, _ = load_golden(name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = ins["x"].clone().requires_grad_(True)
    y, hxs = ref_port.lstm_layerd(sd, "", x)
    assert hxs is None
    assert rel_err(y, outs["y"]) < 1e-6
    (y * ins["w"]).sum().backward()
    assert rel_err(x.grad, grads["x"]) < 1e-5
    for k, v in sd.items():
        assert rel_err(v.grad, grads[k]) < 1e-5, k


def test_port_sampler_matches_reference():
    sd, ins, outs, _, _ = load_golden("lstm_sampler")
    y1, hx = ref_port.lstm_sampler(sd, "", ins["x"][:, :8], 4)
    y2, hx2 = ref_port.lstm_sampler(sd, "", ins["x"][:, 8:], 4, hx)
    assert y1.shape == (3, 2, 16) and y2.shape == (3, 1, 16)
    for got, key in ((y1, "y1"), (y2, "y2"), (hx2[0], "h"), (hx2[1], "c")):
        assert rel_err(got, outs[key]) < 1e-6


def _lws_batch(ins):
    return [ins[k] for k in ("acoustic", "motion_p", "motion_s", "lead_a", "lead_p", "lead_s",
                             "target")]


def test_port_lstm_with_sample_matches_reference():
    sd, ins, outs, grads, meta = load_golden("lstm_with_sample")
    ratio = int(meta["ratio"])
    batch = _lws_batch(ins)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y, lead_len, (hx_s, hxs) = ref_port.lws_forward(sdg, ratio, *batch[:6])
    assert lead_len == int(meta["lead_len"]) and hxs is None
    assert rel_err(y, outs["y"]) < 1e-6
    assert rel_err(hx_s[0], outs["hs"]) < 1e-6 and rel_err(hx_s[1], outs["cs"]) < 1e-6
    loss = ref_port.lws_training_step(sdg, ratio, batch)
    assert abs(float(loss) - float(outs["loss"])) <= 1e-6 * abs(float(outs["loss"]))
    loss.backward()
    for k, v in sdg.items():
        assert rel_err(v.grad, grads[k]) < 1e-5, k
    T = ins["motion_p"].shape[1]
    with torch.no_grad():
        for key, mask in (("pred_tf", torch.zeros(T, dtype=torch.bool)),
                          ("pred_free", torch.ones(T, dtype=torch.bool)),
                          ("pred_ss", ins["mask_ss"])):
            pred = ref_port.lws_rollout(sd, ratio, batch, mask)
            assert rel_err(pred, outs[key]) < 1e-6, key
    assert 0 < int(ins["mask_ss"].sum()) < T  # the golden mask exercises both branches
    # scheduled-sampling training step: gradient flows through the feedback path (Q6)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    pred = ref_port.lws_rollout(sdg, ratio, batch, ins["mask_ss"])
    loss = ref_port.masked_loss(pred, ins["target"])
    assert abs(float(loss) - float(outs["loss_ss"])) <= 1e-6 * abs(float(outs["loss_ss"]))
    loss.backward()
    for k, v in sdg.items():
        assert rel_err(v.grad, grads["ss/" + k]) < 1e-5, k


def test_port_simple_lstm_matches_reference():
    sd, ins, outs, grads, meta = load_golden("simple_lstm")
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    heads = int(meta["heads"])
    y = ref_port.simple_lstm_forward(sdg, ins["acoustic"], ins["motion"], heads)
    assert y.shape == (3, 1, 6) and rel_err(y, outs["y"]) < 1e-6
    loss = ref_port.simple_lstm_training_step(sdg, (ins["acoustic"], ins["motion"], ins["target"]),
                                              heads)
    assert abs(float(loss) - float(outs["loss"])) <= 1e-6 * abs(float(outs["loss"]))
    loss.backward()
    for k, v in sdg.items():
        assert rel_err(v.grad, grads[k]) < 2e-5, k


def test_numpy_gru_matches_torch_forward_and_autograd():
    """oracle/gru_numpy.py pinned to torch.nn.GRU (fp64, CPU): outputs and every gradient."""
    from oracle import gru_numpy
    torch.manual_seed(11)
    T, B, I, H = 6, 3, 5, 7
    m = torch.nn.GRU(I, H, 1).double()
    x = torch.randn(T, B, I, dtype=torch.double, requires_grad=True)
    h0 = torch.randn(1, B, H, dtype=torch.double, requires_grad=True)
    y, hn = m(x, h0)
    wy = torch.randn_like(y)
    (y * wy).sum().backward()
    n = lambda t: t.detach().numpy()
    yn, cache = gru_numpy.gru_layer_forward(n(x), n(m.weight_ih_l0), n(m.weight_hh_l0), n(m.bias_ih_l0),
                                            n(m.bias_hh_l0), n(h0[0]))
    assert np.abs(yn - n(y)).max() < 1e-12 and np.abs(yn[-1] - n(hn[0])).max() < 1e-12
    grads = gru_numpy.gru_layer_backward(n(wy), n(x), cache, n(m.weight_ih_l0), n(m.weight_hh_l0))
    for got, ref in zip(grads, (x.grad, m.weight_ih_l0.grad, m.weight_hh_l0.grad, m.bias_ih_l0.grad,
                                m.bias_hh_l0.grad, h0.grad[0])):
        assert np.abs(got - n(ref)).max() < 1e-10


def test_audio_oracle_matches_reference_fixture_and_product_filterbank():
    """oracle/audio_ref.py (fp64 numpy) against the output of the reference's AudioPreprocessor (fixture made by
    oracle/make_golden.py from the unmodified class); the product's mel filterbank / DFT basis (host constants of the GPU
    front-end) against the oracle's filterbank and numpy's rfft."""
    from oracle.audio_ref import audio_features, mel_filterbank
    from multimodalreactiongeneration_b200.mr_gen.utils.preprocess import audio as fe
    d = golden("audio_features")
    for order in (0, 1, 2):
        ref = d[f"out/features_order{order}"]
        mine = audio_features(d["in/wave"], 400, 160, 26, 16000, order)
        assert mine.shape == ref.shape
        assert np.abs(mine - ref).max() <= 5e-5
    assert np.abs(fe.mel_filterbank(201, 26, 16000) - mel_filterbank(201, 26, 16000)).max() <= 1e-12
    basis = fe.dft_basis(400, 404)
    x = np.random.default_rng(0).standard_normal(400)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    spec = np.fft.rfft(x * w)
    assert np.abs(basis[:201] @ x - spec.real).max() <= 1e-9
    assert np.abs(basis[201:402] @ x - spec.imag).max() <= 1e-9
    assert np.abs(basis[402:]).max() == 0.0
