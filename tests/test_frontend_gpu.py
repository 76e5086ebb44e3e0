"""On-GPU audio front-end (mr_gen/utils/preprocess/audio.py mirror) against the reference's own output (golden fixture made
by running the unmodified reference class) and against the fp64 oracle on batched / ragged input."""
import types

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

TOL = 2e-4   # absolute, on log-energies of magnitude up to 23: fp32 DFT-by-GEMM (3xTF32) + fp32 mel sums


def _cfg(order):
    return types.SimpleNamespace(nfft=400, shift=160, nmels=26, sample_rate=16000, delta_order=order)


@pytest.mark.parametrize("order", [0, 1, 2])
def test_audio_features_match_reference_fixture(order):
    from multimodalreactiongeneration_b200.mr_gen.utils.preprocess import AudioPreprocessor
    d = golden("audio_features")
    pre = AudioPreprocessor(_cfg(order))
    out = pre.features(torch.from_numpy(d["in/wave"]).cuda())
    ref = d[f"out/features_order{order}"]
    assert tuple(out.shape) == ref.shape
    assert float(np.abs(out.cpu().numpy() - ref).max()) <= TOL


def test_audio_features_batched_ragged_length_vs_oracle():
    """A batch of sequences whose length is not a multiple of the hop (padded internally), large enough for the
    persistent pre-split GEMM (M >= 4096 frames): every sequence must equal the fp64 oracle on its own waveform."""
    from multimodalreactiongeneration_b200.mr_gen.utils.preprocess import AudioPreprocessor
    from oracle.audio_ref import audio_features
    g = torch.Generator().manual_seed(5)
    B, S = 6, 16000 * 8 + 37
    wave = torch.randn(B, S, generator=g) * torch.linspace(0.01, 1.0, B)[:, None]
    pre = AudioPreprocessor(_cfg(2))
    out = pre.features(wave.cuda()).cpu().numpy()
    frames = (S - 400) // 160 + 1
    assert out.shape == (B, frames - 2, 81)
    for b in range(B):
        ref = audio_features(wave[b].numpy(), 400, 160, 26, 16000, 2)
        assert float(np.abs(out[b] - ref).max()) <= TOL, b


def test_audio_features_reject_cpu_and_short_input():
    from multimodalreactiongeneration_b200.mr_gen.utils.preprocess import AudioPreprocessor
    pre = AudioPreprocessor(_cfg(2))
    with pytest.raises(RuntimeError):
        pre.features(torch.zeros(16000))
    with pytest.raises(ValueError):
        pre.features(torch.zeros(400 + 160, device="cuda"))   # 2 frames, order 2 needs 3
