"""Frozen CPU definition of the TTS-side 24 kHz log-mel (BASELINE.json config 5).

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED: the reference repository has no audio->mel code
(its TTS service is text->wav Sambert-HiFiGAN, R:tts-service/app/services/tts_service.py:36-38,164-177), so there
is nothing of the reference's to compare against.  SURVEY.md section 8(c) freezes the HiFi-GAN / BigVGAN `mel_spectrogram`
convention used by zero-shot TTS reference-audio encoders:

  sr 24000, n_fft = win = 1024, hop 256, periodic Hann, reflect pad (n_fft-hop)/2 = 384 per side, center=False
  => frames = N // 256;   magnitude sqrt(re^2 + im^2 + 1e-9);
  80 Slaney-scale, Slaney-normalised triangular filters over 0..12 kHz on 513 bins;
  log(clamp(mel, min=1e-5));   layout [80, frames] (mel-major).

This file restates that definition with numpy; tests/golden/make_golden.py checks it against torch.stft +
torchaudio.functional.melscale_fbanks in the build container.
"""
from __future__ import annotations

import numpy as np


def hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_filters(n_freqs: int = 513, f_min: float = 0.0, f_max: float = 12000.0, n_mels: int = 80,
                       sample_rate: int = 24000) -> np.ndarray:
    """[n_mels, n_freqs] triangular filters in the Hz domain, area-normalised (norm='slaney')."""
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_slaney(f_min), hz_to_mel_slaney(f_max), n_mels + 2)
    f_pts = mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]          # [n_freqs, n_mels+2]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))            # [n_freqs, n_mels]
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm[None, :]
    return fb.T.astype(np.float32)


def num_frames(n_samples: int, hop: int = 256) -> int:
    return n_samples // hop


def tts_log_mel(wave: np.ndarray, *, sample_rate: int = 24000, n_fft: int = 1024, hop: int = 256, n_mels: int = 80,
                f_min: float = 0.0, f_max: float = 12000.0, dtype=np.float32) -> np.ndarray:
    """[N] in [-1, 1] -> [n_mels, N // hop] log-mel."""
    x = np.asarray(wave, dtype=dtype).reshape(-1)
    pad = (n_fft - hop) // 2
    if x.shape[0] <= pad:
        raise AssertionError("reflect padding needs more than (n_fft-hop)/2 samples")
    xp = np.pad(x, (pad, pad), mode="reflect")
    t = 1 + (xp.shape[0] - n_fft) // hop if xp.shape[0] >= n_fft else 0
    idx = np.arange(t)[:, None] * hop + np.arange(n_fft)[None, :]
    n = np.arange(n_fft, dtype=np.float64)
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft)).astype(dtype)   # periodic Hann
    spec = np.fft.rfft(xp[idx] * win[None, :], axis=1)
    mag = np.sqrt((spec.real.astype(dtype) ** 2 + spec.imag.astype(dtype) ** 2 + dtype(1e-9))).astype(dtype)
    fb = slaney_mel_filters(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate).astype(dtype)
    mel = (mag @ fb.T).astype(dtype)
    return np.log(np.maximum(mel, dtype(1e-5))).astype(dtype).T
