"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden            # needs /root/reference; writes tests/golden/

Every fixture holds the reference module's ``state_dict`` (keys = checkpoint layout, SURVEY.md
Appendix B), the seeded inputs, and what the reference computed from them on CPU: outputs,
loss and parameter gradients.  Sizes are small (H=32) so the files stay a few hundred KB; the
full-size configurations are checked on the GPU against torch.nn.LSTM directly.

The only deviation from the reference is the documented Q1 unwrap for SimpleLSTM (its
``forward`` raises at HEAD because ``LSTMLayerd`` returns a tuple): a forward hook on the three
``LSTMLayerd`` instances returns element ``[0]``.  No reference file is modified or copied.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.ref_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, sd, inputs, outputs, grads=None, meta=None):
    blob = {}
    for k, v in sd.items():
        blob["sd/" + k] = v.detach().numpy()
    for k, v in inputs.items():
        blob["in/" + k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in outputs.items():
        blob["out/" + k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in (grads or {}).items():
        blob["grad/" + k] = v.detach().numpy()
    for k, v in (meta or {}).items():
        blob["meta/" + k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KB, {len(blob)} arrays")


def lws_cfg(ns, scheduled=False):
    model = ns.AttrDict(
        nmels=9, delta_order=0, use_centroid=True, use_angle=True,
        sampler_hidden_size=16, sampler_num_layers=2, sampler_dropout_rate=0,
        sampling_rate=16000, shift=160, fps=25, pred_fps=50.0,  # ratio = 100/50 = 2
        hidden_size=32, bottleneck_size=8, num_layers=2, num_lstm=1, dropout_rate=0.0,
        use_layer_norm=True, use_relu=True, use_mixing=False, use_residual=True,
        delta_loss_scale=1, loss_type="huber", loss_reduction="mean", huber_delta=1.0,
        smoothl1_beta=1.0, use_scheduled_sampling=scheduled, max_epochs=6,
    )
    optim = ns.AttrDict(use_optimizer="adam", lr=5e-6, weight_decay=1e-2, use_lr_sched=True,
                        max_epochs=100, momentum=0.9)
    metrics = ns.AttrDict(use_centroid=True, use_angle=True, delta_order=0)
    return model, optim, metrics


def lws_batch(g, B=3, T=7, lead=2, ratio=2, A=10, P=6, pad_rows=1):
    r = lambda *s: torch.randn(*s, generator=g)
    acoustic, mp, ms = r(B, T * ratio, A), r(B, T, P), r(B, T, P)
    la, lp, ls = r(B, lead * ratio, A), r(B, lead, P), r(B, lead, P)
    target = r(B, T, P)
    lens = torch.full((B,), T, dtype=torch.long)
    for b in range(pad_rows):  # trailing padding exactly like collate_fn (dataloader.py:114-121)
        cut = T - 2
        acoustic[b, cut * ratio:], mp[b, cut:], ms[b, cut:], target[b, cut:] = -100, -100, -100, -100
        lens[b] = cut
    return [(acoustic, lens * ratio), (mp, lens), (ms, lens), (la, None), (lp, None), (ls, None),
            (target, lens)]


def metaformer_cfg(ns, scheduled=False):
    """reference lstmformer/config.yaml model block at test size (hidden 32, 2 blocks, 2-layer encoders)"""
    model = ns.AttrDict(
        main_modal_idx=2, hidden_size=32, dropout=0.0, num_block=2, num_layerd=1, encoder_num_layer=2,
        num_internal_layer=1, residual=True, residual_layer_norm=True, bias=True,
        emb_mixers=["lstm", "lstm", "lstm"], bottleneck_size=8, nonlinearity="none", ffn_nonlinearity="relu",
        proj_size=0, num_heads=4, add_bias_kv=False, add_zero_attn=False, max_context_len=10,
        repeat_with_encoder=False, interlayer_residual=False, interlayer_residual_norm=True,
        sampling_rate=16000, shift=160, pred_fps=50.0,  # ratio = 100/50 = 2
        modalities=["audio", "motion", "motion"], use_centroid=True, use_angle=True, nmels=9, delta_order=0,
        loss_type="huber", loss_reduction="mean", huber_delta=1.0, smoothl1_beta=1.0, delta_loss_scale=1,
        use_scheduled_sampling=scheduled, max_epochs=6)
    _, optim, metrics = lws_cfg(ns)
    return model, optim, metrics


def metaformer_fixture(ns):
    """7. Metaformer (lstmformer): forward, teacher-forced step, attention masks, rollout in 3 modes, scheduled-
    sampling training step — all from the unmodified reference on CPU."""
    from mr_gen.model.lstmformer.lstmformer import Metaformer
    from mr_gen.model.utils.multi_modal_metaformer import gen_attention_mask

    torch.manual_seed(7)
    m = Metaformer(*metaformer_cfg(ns))
    g = torch.Generator().manual_seed(707)
    batch = lws_batch(g)
    names = ["acoustic", "motion_p", "motion_s", "lead_a", "lead_p", "lead_s", "target"]
    ins = {n: batch[i][0].clone() for i, n in enumerate(names)}
    y, hxs = m.forward(*batch[:-1])
    leaves = []

    def walk(o):
        if isinstance(o, dict):
            for v in o.values():
                walk(v)
        elif isinstance(o, (list, tuple)):
            for v in o:
                walk(v)
        else:
            leaves.append(o)
    walk(hxs)
    assert len(hxs) == 2 and all(leaf is None for leaf in leaves)   # Q3
    outs = {"y": y}
    own = torch.cat([batch[5][0], batch[2][0]], dim=1)
    audio = torch.cat([batch[3][0], batch[0][0]], dim=1)
    outs["mask_own_audio"] = gen_attention_mask(own, audio, 4)
    outs["mask_own_own"] = gen_attention_mask(own, own, 4)
    outs["mask_audio_own"] = gen_attention_mask(audio, own, 4)
    tf_batch = [(t.clone(), l) for t, l in batch]
    loss = m.training_step(tf_batch)["loss"]
    loss.backward()
    outs["loss"] = loss
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    with torch.no_grad():
        outs["pred_tf"], outs["target_tf"] = m.prediction([(t.clone(), l) for t, l in batch])
        outs["pred_free"], _ = m.prediction([(t.clone(), l) for t, l in batch], full_generation=True)
        m.current_epoch = 3
        torch.manual_seed(77)
        ins["mask_ss"] = torch.rand(batch[1][0].shape[1]) < (3 / 6)
        torch.manual_seed(77)
        outs["pred_ss"], _ = m.prediction([(t.clone(), l) for t, l in batch], use_scheduled_sampling=True)
    m.zero_grad()
    m.use_scheduled_sampling = True
    torch.manual_seed(77)
    loss_ss = m.training_step([(t.clone(), l) for t, l in batch])["loss"]
    loss_ss.backward()
    outs["loss_ss"] = loss_ss
    grads.update({"ss/" + k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    _save("metaformer", m.state_dict(), ins, outs, grads, {"ratio": m.ratio, "lead_len": batch[4][0].shape[1]})


def gru_mixer_fixture(ns):
    """8. GRUMixerLayerd (lstmformer with emb_mixers: gru) from the unmodified reference."""
    from mr_gen.model.utils.mixer_block import GRUMixerLayerd
    torch.manual_seed(8)
    m = GRUMixerLayerd(hidden_size=32, num_layerd=2, residual=True, residual_layer_norm=True, nonlinearity="none",
                       device=torch.device("cpu"))
    x = torch.randn(3, 7, 32, requires_grad=True)
    y, hx, _ = m(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    _save("gru_mixer_layerd", m.state_dict(), {"x": x, "w": w}, {"y": y},
          {**{k: p.grad for k, p in m.named_parameters()}, "x": x.grad}, {"hx_is_none": hx is None})


def audio_fixture():
    """AudioPreprocessor of the reference (mr_gen/utils/preprocess/audio.py) on a seeded waveform.  The class is loaded
    from its own file; the only stub is ``torchaudio._backend.soundfile_backend`` (absent from this torchaudio; it is
    the file reader, which is bypassed: the steps of ``__call__`` after the read are executed here verbatim)."""
    import importlib.util
    import sys
    import types
    from oracle.ref_loader import REFERENCE_ROOT
    if "torchaudio._backend.soundfile_backend" not in sys.modules:
        be = types.ModuleType("torchaudio._backend")
        sf = types.ModuleType("torchaudio._backend.soundfile_backend")
        be.soundfile_backend = sf
        sys.modules["torchaudio._backend"] = be
        sys.modules["torchaudio._backend.soundfile_backend"] = sf
    spec = importlib.util.spec_from_file_location(
        "_ref_audio", os.path.join(REFERENCE_ROOT, "mr_gen", "utils", "preprocess", "audio.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = types.SimpleNamespace(nfft=400, shift=160, nmels=26, sample_rate=16000, delta_order=2)
    g = torch.Generator().manual_seed(77)
    n = 16000 + 240                                   # one second and a bit: 100 frames -> 98 after the deltas
    t = torch.arange(n) / 16000.0
    wave = 0.3 * torch.sin(2 * torch.pi * 220.0 * t) + 0.1 * torch.sin(2 * torch.pi * 3100.0 * t) \
        + 0.05 * torch.randn(n, generator=g)
    wave[4000:5200] = 0.0                             # a silent stretch: the 1e-6 / 1e-10 clamps are exercised
    blob = {"in/wave": wave.numpy()}
    for order in (0, 1, 2):
        cfg.delta_order = order
        pre = mod.AudioPreprocessor(cfg)
        fbank = pre.fbank(wave)                       # audio.py:30-37, verbatim
        fbank = pre.log(torch.clamp(fbank, 1e-10))
        power = pre.compute_log_power(wave)
        fbank = torch.cat([fbank, power.unsqueeze(0)], dim=0).T.to(torch.float32)
        blob[f"out/features_order{order}"] = pre.compute_delta(fbank).numpy()
    path = os.path.join(OUT, "audio_features.npz")
    np.savez_compressed(path, **blob)
    print(f"audio_features: {os.path.getsize(path) / 1024:.0f} KB")


def main():
    import sys
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    if "--only-audio" in sys.argv:
        return audio_fixture()
    ns = load_reference()
    if "--only-metaformer" in sys.argv:
        return metaformer_fixture(ns)
    if "--only-gru" in sys.argv:
        return gru_mixer_fixture(ns)

    # ---- 1. LSTMLayerd as used by lstm_with_sampling (uni, residual+LN, no FFN) -----------
    torch.manual_seed(1)
    m = ns.LSTMLayerd(input_size=32, lstm_hidden_size=32, affine_hidden_size=32, bottleneck_size=8,
                      num_layers=2, num_layers_per_block=1, output_size=32, dropout=0.0,
                      bidirectional=False, use_layer_norm=True, use_mixing=False,
                      use_residual=True, use_feed_forward=False)
    x = torch.randn(4, 9, 32, requires_grad=True)
    y, hxs = m(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    _save("lstm_layerd_uni", m.state_dict(), {"x": x, "w": w}, {"y": y},
          {**{k: p.grad for k, p in m.named_parameters()}, "x": x.grad})

    # ---- 2. LSTMLayerd as used by simple_lstm (bi, mixing, FFN) ---------------------------
    torch.manual_seed(2)
    m = ns.LSTMLayerd(input_size=32, lstm_hidden_size=16, affine_hidden_size=32, bottleneck_size=8,
                      num_layers=2, num_layers_per_block=1, output_size=32, dropout=0.0,
                      bidirectional=True, use_layer_norm=True, use_relu=True, use_mixing=True,
                      use_residual=True)
    x = torch.randn(3, 6, 32, requires_grad=True)
    y, _ = m(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    _save("lstm_layerd_bi_mix_ffn", m.state_dict(), {"x": x, "w": w}, {"y": y},
          {**{k: p.grad for k, p in m.named_parameters()}, "x": x.grad})

    # ---- 3. LSTMSampler (2 layers, stride 4, carried state) -------------------------------
    torch.manual_seed(3)
    m = ns.LSTMSampler(16, 2, 0.0, 4)
    x = torch.randn(3, 12, 16)
    y1, hx = m(x[:, :8])
    y2, hx2 = m(x[:, 8:], hx)
    _save("lstm_sampler", m.state_dict(), {"x": x}, {"y1": y1, "y2": y2, "h": hx2[0], "c": hx2[1]})

    # ---- 4. LSTMwithSample: forward, teacher-forced step, rollout in 3 modes --------------
    torch.manual_seed(4)
    model, optim, metrics = lws_cfg(ns)
    m = ns.LSTMwithSample(model, optim, metrics)
    m.use_device = "cpu"
    g = torch.Generator().manual_seed(44)
    batch = lws_batch(g)
    y, (lead_len, motion_len, _), (hx_s, hxs) = m.forward(*batch[:-1])
    assert hxs is None  # Q2
    loss = m.training_step(batch)["loss"]
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters()}
    ins = {n: batch[i][0] for i, n in enumerate(
        ["acoustic", "motion_p", "motion_s", "lead_a", "lead_p", "lead_s", "target"])}
    outs = {"y": y, "loss": loss, "hs": hx_s[0], "cs": hx_s[1]}
    with torch.no_grad():
        outs["pred_tf"], _ = m.prediction(batch)                       # mask all False
        outs["pred_free"], _ = m.prediction(batch, full_generation=True)  # mask all True
        m.current_epoch = 3                                             # rate 3/6
        torch.manual_seed(77)
        mask = torch.rand(batch[1][0].shape[1]) < (3 / 6)              # same draw as :389
        torch.manual_seed(77)
        outs["pred_ss"], _ = m.prediction(batch, use_scheduled_sampling=True)
        ins["mask_ss"] = mask
    # scheduled-sampling TRAINING step: gradient through the feedback path (Q6)
    m.zero_grad()
    m.use_scheduled_sampling = True
    torch.manual_seed(77)
    loss_ss = m.training_step(batch)["loss"]
    loss_ss.backward()
    outs["loss_ss"] = loss_ss
    grads_ss = {"ss/" + k: p.grad.clone() for k, p in m.named_parameters()}
    _save("lstm_with_sample", m.state_dict(), ins, outs, {**grads, **grads_ss},
          {"ratio": m.ratio, "lead_len": lead_len})

    # ---- 5. SimpleLSTM (Q1 unwrap by forward hook) ----------------------------------------
    torch.manual_seed(5)
    cfg = ns.AttrDict(
        acostic_feat_size=10, motion_feat_size=6, motion_num_lstm=1, acostic_num_lstm=1,
        acostic_num_layers=2, motion_num_layers=2, acostic_lstm_size=32, motion_lstm_size=32,
        acostic_affine_size=32, motion_affine_size=32, acostic_output_size=32,
        motion_output_size=32, att_heads=4, att_num_layers=2, att_use_residual=True,
        att_use_layer_norm=True, dropout_rate=0, output_size=6, bidirectional=False,
        use_layer_norm=True, use_relu=True, use_mixing=True, use_residual=True,
        decoder_num_layers=2, decoder_num_lstm=1, decoder_lstm_size=32, decoder_affine_size=32,
        decoder_bottleneck_size=8, decoder_output_size=32, decoder_mapping_size=8,
        decoder_bidirectional=False, decoder_use_layer_norm=True, decoder_use_relu=True,
        decoder_use_mixing=True, decoder_use_residual=True, delta_loss_scale=1, all_static=False)
    _, optim, metrics = lws_cfg(ns)
    m = ns.SimpleLSTM(cfg, optim, metrics)
    unwrap = lambda mod, args, out: out[0]
    for layerd in (m.acoustic_encoder.acostic_lstm, m.motion_encoder.motion_lstm,
                   m.motion_decoder.decoder_lstm):
        layerd.register_forward_hook(unwrap)
    g = torch.Generator().manual_seed(55)
    a = torch.randn(3, 8, 10, generator=g)
    mo = torch.randn(3, 8, 6, generator=g)
    tgt = torch.randn(3, 1, 6, generator=g)
    y = m(a, mo)
    loss = m.training_step((a, mo, tgt))["loss"]
    loss.backward()
    _save("simple_lstm", m.state_dict(), {"acoustic": a, "motion": mo, "target": tgt},
          {"y": y, "loss": loss}, {k: p.grad for k, p in m.named_parameters()}, {"heads": 4})

    # ---- 6. LSTMMixerLayerd (lstmformer token mixer) --------------------------------------
    if hasattr(ns, "LSTMMixerLayerd"):
        torch.manual_seed(6)
        m = ns.LSTMMixerLayerd(hidden_size=32, num_layerd=2, residual=True,
                               residual_layer_norm=True, nonlinearity="none",
                               device=torch.device("cpu"))
        x = torch.randn(3, 7, 32, requires_grad=True)
        y, hx, _ = m(x)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        _save("lstm_mixer_layerd", m.state_dict(), {"x": x, "w": w}, {"y": y},
              {**{k: p.grad for k, p in m.named_parameters()}, "x": x.grad},
              {"hx_is_none": hx is None})

    metaformer_fixture(ns)
    gru_mixer_fixture(ns)
    audio_fixture()


if __name__ == "__main__":
    main()
