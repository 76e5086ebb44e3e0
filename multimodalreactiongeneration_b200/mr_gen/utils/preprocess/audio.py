"""``AudioPreprocessor`` on the GPU — mirror of mr_gen/utils/preprocess/audio.py:6-67.

Same constructor (``cfg.nfft / shift / nmels / sample_rate / delta_order``) and the same result for a waveform: log-mel
filterbank energies (torchaudio ``MelSpectrogram(center=False)`` semantics: periodic Hann window, power spectrum, HTK mel
scale, no normalisation) + the log-power of the raw frame, then delta / delta-delta by first differences with the leading
frames cut (81-d at the reference's 26 mels, order 2).  What changes is where it runs:

* the windowed real DFT of all frames of all sequences is ONE tensor-core GEMM of the library over the waveform viewed as
  overlapping frames (row stride = ``shift``; nothing is unfolded) against a constant cos | -sin basis that is pre-split
  into tf32 hi / lo planes once (``mrg_gemm_strided_split``, fp32-grade 3xTF32);
* power spectrum -> mel -> log, the raw-frame log-power (a Python per-frame loop in the reference, audio.py:41-53) and
  the deltas are two small kernels (``mrg_audio_features``).

``features(waveform)`` takes a device tensor ``[S]`` or ``[B, S]`` and returns ``[frames - delta_order, 81]`` (or
``[B, ..]``) on the device, so streaming generation can go waveform -> pose without a host hop.  ``__call__(wavepath,
start, end)`` keeps the reference's signature; reading the file is the reference's own host-side step (soundfile) and
is only attempted when that package is present.  No CPU fallback for the arithmetic."""
from __future__ import annotations

import math

import numpy as np
import torch

from .... import _cabi
from ....linear import _gemm, _gemm_split, _split_ok, _split_weight


def mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int) -> np.ndarray:
    """[n_freqs, n_mels] triangular filters, HTK mel scale, f_min 0, f_max sample_rate / 2, no normalisation
    (torchaudio.functional.melscale_fbanks defaults as used by MelSpectrogram in audio.py:14-20)."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    hz2mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    mel2hz = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_pts = mel2hz(np.linspace(hz2mel(0.0), hz2mel(float(sample_rate // 2)), n_mels + 2))
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def dft_basis(nfft: int, ld: int) -> np.ndarray:
    """[ld, nfft]: row k < bins = w[n] cos(2 pi k n / nfft), row bins + k = -w[n] sin(..), periodic Hann window w."""
    bins = nfft // 2 + 1
    n = np.arange(nfft)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / nfft)
    ang = 2.0 * np.pi * np.outer(np.arange(bins), n) / nfft
    out = np.zeros((ld, nfft))
    out[:bins] = np.cos(ang) * w
    out[bins:2 * bins] = -np.sin(ang) * w
    return out


class AudioPreprocessor:
    def __init__(self, cfg, device="cuda"):
        self.cfg = cfg
        self.nfft = cfg.nfft
        self.shift = cfg.shift
        self.nmels = cfg.nmels
        self.sample_rate = cfg.sample_rate
        self.delta_order = cfg.delta_order
        if self.delta_order not in (0, 1, 2):
            raise ValueError("delta_order must be 0, 1 or 2")
        if self.shift % 4 != 0:
            raise ValueError("AudioPreprocessor (sm_100a): shift must be a multiple of 4 samples (16-byte frame stride)")
        self.device = torch.device(device)
        self.bins = self.nfft // 2 + 1
        self.ld = (2 * self.bins + 3) // 4 * 4
        self._basis = None   # device tensors, built on first use (so that construction works without a GPU)

    def _constants(self):
        if self._basis is None:
            basis = torch.from_numpy(dft_basis(self.nfft, self.ld)).to(torch.float32).to(self.device)
            self._basis = basis
            self._basis_split = _split_weight(basis)
            self._fb = torch.from_numpy(mel_filterbank(self.bins, self.nmels, self.sample_rate)).to(torch.float32) \
                .contiguous().to(self.device)
        return self._basis, self._basis_split, self._fb

    @torch.no_grad()
    def features(self, waveform: torch.Tensor) -> torch.Tensor:
        if not waveform.is_cuda:
            raise RuntimeError("AudioPreprocessor.features has no CPU path: the waveform must live on a B200 device")
        single = waveform.dim() == 1
        wave = (waveform[None] if single else waveform).to(torch.float32)
        B, S = wave.shape
        frames = (S - self.nfft) // self.shift + 1 if S >= self.nfft else 0
        if frames <= self.delta_order:
            raise ValueError(f"waveform too short: {S} samples give {max(frames, 0)} frames")
        pad = (-S) % self.shift       # one uniform frame stride over the whole batch needs S % shift == 0
        if pad:
            wave = torch.nn.functional.pad(wave, (0, pad))
        wave = wave.contiguous()
        Sp = S + pad
        basis, basis_split, fb = self._constants()
        dev = wave.device
        M = (B * Sp - self.nfft) // self.shift + 1          # rows of the strided frame view over the flattened batch
        spec = torch.empty((M, self.ld), dtype=torch.float32, device=dev)
        flags = 0
        if wave.data_ptr() % 16 == 0 and _split_ok(M, self.ld, self.nfft, self.shift, 1, 1, self.nfft, flags):
            _gemm_split(wave, self.shift, 1, basis_split, 1, self.nfft, None, spec, M, self.ld, self.nfft, flags)
        else:
            _gemm(wave, self.shift, 1, basis, 1, self.nfft, None, spec, M, self.ld, self.nfft, flags)
        nf = self.nmels + 1
        feat = torch.empty((B, frames, nf), dtype=torch.float32, device=dev)
        out = torch.empty((B, frames - self.delta_order, nf * (self.delta_order + 1)), dtype=torch.float32, device=dev)
        L = _cabi.lib()
        with torch.cuda.device(dev):
            st = L.mrg_audio_features(spec.data_ptr(), self.ld, wave.data_ptr(), fb.data_ptr(), feat.data_ptr(),
                                      out.data_ptr(), B, frames, Sp // self.shift, Sp, self.shift, self.nfft, self.nmels,
                                      self.delta_order, torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "mrg_audio_features")
        return out[0] if single else out

    def __call__(self, wavepath: str, start: int, end: int) -> torch.Tensor:
        try:
            import soundfile
        except ImportError as e:   # reading audio files is the reference's host-side step, not part of this library
            raise ImportError("AudioPreprocessor.__call__ reads the file with soundfile (as the reference does through "
                              "torchaudio's soundfile backend); use features(waveform) with a waveform tensor") from e
        length = -1 if end == -1 else end - start
        data, sample_rate = soundfile.read(wavepath, start=start, frames=length, dtype="float32", always_2d=True)
        if sample_rate != self.sample_rate:
            raise ValueError("sample_rate must be same as --sample-rate")
        out = self.features(torch.from_numpy(np.ascontiguousarray(data[:, 0])).to(self.device))
        assert len(out) != 0, f"start: {start}, end: {end}, stride: {1}"
        return out
