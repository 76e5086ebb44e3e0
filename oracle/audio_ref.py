"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, fp64) of the reference's audio features,
mr_gen/utils/preprocess/audio.py:24-67: MelSpectrogram(center=False) -> log(clamp 1e-6) (:14-22,30-31), raw-frame
log-power with clamp 1e-10 (:41-53), concatenation to nmels + 1 (:33), delta / delta-delta (:55-67).
Pinned by tests/golden/audio_features.npz, which oracle/make_golden.py generates by running the UNMODIFIED reference
class (its MelSpectrogram, compute_log_power and compute_delta) on a seeded waveform; only the file read is bypassed."""
import numpy as np


def mel_filterbank(n_freqs, n_mels, sample_rate):
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_max = 2595.0 * np.log10(1.0 + (sample_rate // 2) / 700.0)
    f_pts = 700.0 * (10.0 ** (np.linspace(0.0, m_max, n_mels + 2) / 2595.0) - 1.0)
    f_diff = np.diff(f_pts)
    slopes = f_pts[None, :] - all_freqs[:, None]
    return np.maximum(0.0, np.minimum(-slopes[:, :-2] / f_diff[:-1], slopes[:, 2:] / f_diff[1:]))


def audio_features(wave, nfft, shift, nmels, sample_rate, delta_order):
    wave = np.asarray(wave, dtype=np.float64)
    frames = (len(wave) - nfft) // shift + 1
    idx = np.arange(nfft)[None, :] + shift * np.arange(frames)[:, None]
    x = wave[idx]                                                   # [frames, nfft]
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(nfft) / nfft)   # periodic Hann (torch.hann_window default)
    power = np.abs(np.fft.rfft(x * window, axis=1)) ** 2
    mel = power @ mel_filterbank(nfft // 2 + 1, nmels, sample_rate)
    logmel = np.log(np.maximum(mel, 1e-6))
    logpow = np.log(np.maximum((x ** 2).sum(axis=1), 1e-10))
    fbank = np.concatenate([logmel, logpow[:, None]], axis=1)
    if delta_order == 0:
        return fbank
    d1 = fbank[1:] - fbank[:-1]
    if delta_order == 1:
        return np.concatenate([fbank[1:], d1], axis=1)
    d2 = d1[1:] - d1[:-1]
    if delta_order == 2:
        return np.concatenate([fbank[2:], d1[1:], d2], axis=1)
    raise ValueError("delta_order must be 0, 1 or 2")
