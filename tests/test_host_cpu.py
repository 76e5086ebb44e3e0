"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares, the
drop-in module keeps nn.LSTM's parameter layout and refuses to run without a GPU, the mirrored models keep
the reference's checkpoint keys, the config loader, the checkpoint round trip, and the data-parallel
gradient bucket under gloo with world_size 2."""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden


def _header_functions():
    text = open(os.path.join(ROOT, "include", "mrg_lstm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrg_[a-z0-9_]+)\s*\(", text)))


def test_cabi_library_exports_every_declared_symbol():
    import ctypes
    from multimodalreactiongeneration_b200 import _build, _cabi
    _build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.EXPORTS) == names  # the binding covers the whole header
    assert _cabi.lib().mrg_version() == 100


def test_b200lstm_is_parameter_compatible_with_nn_lstm():
    from multimodalreactiongeneration_b200 import B200LSTM
    for kw in (dict(num_layers=1), dict(num_layers=2, bidirectional=True), dict(num_layers=2, bias=False)):
        torch.manual_seed(0)
        a = B200LSTM(12, 20, batch_first=True, **kw)
        torch.manual_seed(0)
        b = torch.nn.LSTM(12, 20, batch_first=True, **kw)
        assert list(a.state_dict().keys()) == list(b.state_dict().keys())
        assert all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
        b.load_state_dict(a.state_dict())  # interchangeable checkpoints


def test_no_cpu_fallback():
    from multimodalreactiongeneration_b200 import B200LSTM
    m = B200LSTM(8, 16, batch_first=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(2, 3, 8))
    with pytest.raises(NotImplementedError):
        B200LSTM(8, 16, proj_size=4)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multimodalreactiongeneration_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_mirrored_models_keep_reference_checkpoint_keys(tmp_path):
    from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg, simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import load_checkpoint, save_checkpoint
    m = LSTMwithSample(*lstm_with_sampling_cfg())
    keys = list(m.state_dict().keys())
    # SURVEY.md Appendix B
    assert "sampling_lstm.sampler.weight_ih_l1" in keys
    assert "layerd_lstm.lstm_layered.1.lstm_module.module.lstm_module.weight_hh_l0" in keys
    assert "layerd_lstm.lstm_layered.0.lstm_module.layer_norm.weight" in keys
    assert "feed_forward.mapping.bias" in keys and len(keys) == 28
    s = SimpleLSTM(*simple_lstm_cfg(bidirectional=True))
    sk = list(s.state_dict().keys())
    assert "acoustic_encoder.acostic_lstm.lstm_layered.0.lstm_module.module.lstm_module.weight_ih_l0_reverse" in sk
    assert "motion_decoder.decoder_lstm.lstm_layered.1.feed_forward_module.module.mapping.weight" in sk
    assert "multimodal_att.att_layers.2.att_module.module.cross_modal_att.in_proj_weight" in sk
    path = str(tmp_path / "ckpts" / "lstm_with_sampling" / "last.ckpt")
    save_checkpoint(m, path, epoch=3, global_step=7)
    m2 = LSTMwithSample(*lstm_with_sampling_cfg())
    ck = load_checkpoint(m2, path)
    assert ck["epoch"] == 3 and set(ck["state_dict"].keys()) == set(keys)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


def test_error_behaviour_mirrors_reference():
    from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    from multimodalreactiongeneration_b200.mr_gen.model.utils.lstm_block import LSTMBlock, LSTMModule
    from multimodalreactiongeneration_b200.mr_gen.model.utils.residual_connection import ResidualConnection
    with pytest.raises(ValueError):  # lstm_block.py:33-36
        LSTMModule(input_size=8, hidden_size=8, output_size=32, bidirectional=False, use_mixing=False)
    with pytest.raises(ValueError):  # lstm_block.py:67-70
        LSTMBlock(input_size=8, hidden_size=8, lstm_out_size=16, output_size=8, bidirectional=False)
    with pytest.raises(ValueError):  # residual_connection.py:10-13
        ResidualConnection(torch.nn.Identity(), use_layer_norm=True)
    model, optim, metrics = lstm_with_sampling_cfg()
    model["loss_type"] = "hinge"
    with pytest.raises(ValueError, match="invalid loss type"):  # lstm_with_sample.py:71-72
        LSTMwithSample(model, optim, metrics)


def test_config_loader_interpolation_and_overrides(tmp_path):
    from multimodalreactiongeneration_b200.mr_gen.utils.config import load_config
    p = tmp_path / "config.yaml"
    p.write_text("hidden_size: 256\nlr: 5e-6\nname: run\nmodel:\n  hidden_size: ${hidden_size}\n"
                 "  tag: ${name}-${hidden_size}\noptim:\n  lr: ${lr}\n")
    cfg = load_config(str(p), ["hidden_size=128", "model.extra=true"])
    assert cfg.model.hidden_size == 128 and cfg.model.tag == "run-128"
    assert cfg.model.extra is True and float(cfg.optim.lr) == 5e-6


def test_scheduled_sampling_mask_modes_match_reference_draw():
    from multimodalreactiongeneration_b200.mr_gen.configs import lstm_with_sampling_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstm_with_sampling.lstm_with_sample import LSTMwithSample
    m = LSTMwithSample(*lstm_with_sampling_cfg(max_epochs=6))
    m.current_epoch = 3
    torch.manual_seed(77)
    mine = m.draw_sampling_mask(7, 3)
    torch.manual_seed(77)
    ref = torch.rand(7) < (3 / 6)  # lstm_with_sample.py:389
    assert torch.equal(mine, ref) and mine.shape == (7,)
    assert torch.equal(load_golden("lstm_with_sample")[1]["mask_ss"].bool(), ref)


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import FlatGradBucket, broadcast_parameters, init_distributed
world = init_distributed("gloo")
rank = dist.get_rank()
torch.manual_seed(100 + rank)                      # ranks start from different weights ...
net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
broadcast_parameters(net)                          # ... and agree after the broadcast
bucket = FlatGradBucket(net)
opt = torch.optim.SGD(net.parameters(), lr=0.1)
torch.manual_seed(7)
x_all, y_all = torch.randn(8, 5), torch.randn(8, 3)
shard = slice(rank * 4, rank * 4 + 4)              # batch sharded by sequence, equal shards
for _ in range(3):
    bucket.zero()
    torch.nn.functional.mse_loss(net(x_all[shard]), y_all[shard]).backward()
    bucket.all_reduce_mean()
    opt.step()
flat = torch.cat([p.detach().flatten() for p in net.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
if rank == 0:
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for _ in range(3):
        ropt.zero_grad()
        torch.nn.functional.mse_loss(ref(x_all), y_all).backward()   # single-process full batch
        ropt.step()
    rflat = torch.cat([p.detach().flatten() for p in ref.parameters()])
    assert torch.allclose(gathered[0], gathered[1]), "ranks diverged"
    assert torch.allclose(gathered[0], rflat, atol=1e-6), float((gathered[0] - rflat).abs().max())
    print("DDP_OK")
dist.destroy_process_group()
"""


def test_flat_gradient_bucket_data_parallel_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29531", str(script), ROOT],
        capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DDP_OK" in out.stdout


_WORKER_OVERLAP = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer, init_distributed


class Toy(torch.nn.Module):
    # `head` finishes its backward first, then `mid`; `stem` is left for the trailing all-reduce
    ddp_overlap_children = ("head", "mid")

    def __init__(self):
        super().__init__()
        self.stem = torch.nn.Linear(5, 7)
        self.mid = torch.nn.Sequential(torch.nn.Tanh(), torch.nn.Linear(7, 6))
        self.head = torch.nn.Linear(6, 3)

    def forward(self, x):
        return self.head(torch.tanh(self.mid(self.stem(x))))

    def training_step(self, batch):
        x, y = batch
        return {"loss": torch.nn.functional.mse_loss(self(x), y)}

    def configure_optimizers(self):
        return {"optimizer": torch.optim.SGD(self.parameters(), lr=0.1)}


world = init_distributed("gloo")
rank = dist.get_rank()
torch.manual_seed(100 + rank)
net = Toy()
tr = Trainer(net)                                   # broadcasts rank 0's weights, registers the early hooks
assert len(tr._early) == 2, tr._early
torch.manual_seed(7)
x_all, y_all = torch.randn(8, 5), torch.randn(8, 3)
shard = slice(rank * 4, rank * 4 + 4)
for _ in range(3):
    tr.train_step((x_all[shard], y_all[shard]))
assert tr.n_early_all_reduces == 6, tr.n_early_all_reduces   # both early slices went out in every step
flat = torch.cat([p.detach().flatten() for p in net.parameters()])
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
if rank == 0:
    torch.manual_seed(100)
    ref = Toy()
    ropt = torch.optim.SGD(ref.parameters(), lr=0.1)
    for _ in range(3):
        ropt.zero_grad()
        ref.training_step((x_all, y_all))["loss"].backward()
        ropt.step()
    rflat = torch.cat([p.detach().flatten() for p in ref.parameters()])
    assert torch.allclose(gathered[0], gathered[1]), "ranks diverged"
    assert torch.allclose(gathered[0], rflat, atol=1e-6), float((gathered[0] - rflat).abs().max())
    print("DDP_OVERLAP_OK")
dist.destroy_process_group()
"""


def test_trainer_overlapped_bucket_all_reduce_gloo_world2(tmp_path):
    """Early (per sub-module) + trailing all-reduce of the flat bucket must give the single-process trajectory."""
    script = tmp_path / "worker_overlap.py"
    script.write_text(_WORKER_OVERLAP)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
         "127.0.0.1", "--master-port", "29533", str(script), ROOT],
        capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "DDP_OVERLAP_OK" in out.stdout


def test_metaformer_mirror_keeps_reference_keys_masks_and_arguments():
    """lstmformer mirror: checkpoint keys of the unmodified reference (golden fixture), the causal-rectangular
    attention masks (index arithmetic here vs the reference's tile/transpose construction), argument selection."""
    from multimodalreactiongeneration_b200.mr_gen.configs import metaformer_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.lstmformer.lstmformer import Metaformer
    from multimodalreactiongeneration_b200.mr_gen.model.utils.argparser import mixer_layerd_argments_select
    from multimodalreactiongeneration_b200.mr_gen.model.utils.mixer_block import MixerLayerdFactory
    from multimodalreactiongeneration_b200.mr_gen.model.utils.multi_modal_metaformer import gen_attention_mask
    sd, ins, outs, _, _ = load_golden("metaformer")
    m = Metaformer(*metaformer_cfg(hidden=32, blocks=2, encoder_layers=2, bottleneck=8, heads=4, acoustic=10,
                                   ratio=2, max_epochs=6))
    mine = m.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    assert all(mine[k].shape == sd[k].shape for k in sd)
    m.load_state_dict(sd)
    own = torch.cat([ins["lead_s"], ins["motion_s"]], 1)
    audio = torch.cat([ins["lead_a"], ins["acoustic"]], 1)
    for q, k, name in ((own, audio, "mask_own_audio"), (own, own, "mask_own_own"), (audio, own, "mask_audio_own")):
        got = gen_attention_mask(q, k, 4)
        assert got.shape == outs[name].shape and torch.equal(got, outs[name].bool()), name
    with pytest.raises(ValueError):
        gen_attention_mask(torch.zeros(1, 3, 2), torch.zeros(1, 5, 2), 2)
    lstm_kw = mixer_layerd_argments_select("lstm", hidden_size=8, num_heads=4, proj_size=0, kdim=8)
    assert "num_heads" not in lstm_kw and "kdim" not in lstm_kw and lstm_kw["proj_size"] == 0
    assert mixer_layerd_argments_select("conv", hidden_size=8) is None
    gru = MixerLayerdFactory().build("gru", mixer_layerd_argments_select("gru", hidden_size=8, num_layerd=2,
                                                                        residual=True, residual_layer_norm=True))
    assert type(gru.mixer[0].mixer.module.mixer).__name__ == "B200GRU"   # never nn.GRU / cuDNN
    with pytest.raises(RuntimeError):
        gru(torch.zeros(1, 2, 8))                                         # and no CPU path
    with pytest.raises(ValueError):
        MixerLayerdFactory().build("conv", {})


def test_attention_mask_rule_properties():
    """AttentionMaskSpec (the rule the fused attention evaluates instead of reading the reference's mask tensor):
    causal in both rate directions, every query sees at least the keys of its own frame, padding masks only
    padded x padded pairs, and the [B, heads, L, S] view is a broadcast (no per-head copy)."""
    from multimodalreactiongeneration_b200.attention import AttentionMaskSpec
    for L, S, mode, rate in ((5, 15, 1, 3), (12, 4, 2, 3), (7, 7, 1, 1)):
        pad_q = torch.zeros(2, L, dtype=torch.uint8)
        pad_k = torch.zeros(2, S, dtype=torch.uint8)
        pad_q[1, L - 2:] = 1
        pad_k[1, S - 1:] = 1
        m = AttentionMaskSpec(mode, rate, pad_q, pad_k).materialize(4)
        assert m.shape == (2, 4, L, S) and m.stride(1) == 0                 # heads are a broadcast view
        for i in range(L):
            for j in range(S):
                frame_q, frame_k = (i, j // rate) if mode == 1 else (i // rate, j)
                want = frame_k > frame_q
                assert bool(m[0, 0, i, j]) == want                          # unpadded sample: pure causality
                assert bool(m[1, 0, i, j]) == (want or (bool(pad_q[1, i]) and bool(pad_k[1, j])))
            assert not bool(m[0, 0, i].all())                               # no query is left without a key
