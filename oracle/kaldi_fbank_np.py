"""numpy restatement of the Kaldi-compatible fbank the reference computes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
torchaudio.compliance.kaldi (TA = site-packages/torchaudio/compliance/kaldi.py):

  frame_count / frames      TA:44-83   (_get_strided)
  window_function           TA:86-113  (_feature_window_function)
  window_properties         TA:125-151 (_get_waveform_and_window_properties)
  preprocess_frames         TA:154-217 (_get_window)
  mel_scale / mel_banks     TA:318-331, TA:436-511
  fbank                     TA:514-645

`dtype` selects the arithmetic type: np.float32 reproduces the reference's
float32 pipeline (same op order, so the only differences are rfft/mm reduction
order); np.float64 is the "truth" used for tolerance studies.
"""
from __future__ import annotations

import math

import numpy as np

EPSILON_F32 = float(np.finfo(np.float32).eps)  # 1.1920929e-07, TA:20


def next_power_of_2(x: int) -> int:
    """TA:39-41."""
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


def window_properties(num_samples: int, sample_frequency: float, frame_shift_ms: float,
                      frame_length_ms: float, round_to_power_of_two: bool = True):
    """TA:125-151: returns (window_shift, window_size, padded_window_size) in samples."""
    window_shift = int(sample_frequency * frame_shift_ms * 0.001)
    window_size = int(sample_frequency * frame_length_ms * 0.001)
    padded = next_power_of_2(window_size) if round_to_power_of_two else window_size
    if not (2 <= window_size <= num_samples):
        raise AssertionError(f"choose a window size {window_size} that is [2, {num_samples}]")
    if window_shift <= 0:
        raise AssertionError("`window_shift` must be greater than 0")
    if padded % 2 != 0:
        raise AssertionError("the padded `window_size` must be divisible by two.")
    return window_shift, window_size, padded


def frame_count(num_samples: int, window_size: int, window_shift: int, snip_edges: bool = True) -> int:
    """TA:65-70."""
    if snip_edges:
        if num_samples < window_size:
            return 0
        return 1 + (num_samples - window_size) // window_shift
    return (num_samples + (window_shift // 2)) // window_shift


def frames(wave: np.ndarray, window_size: int, window_shift: int, snip_edges: bool = True) -> np.ndarray:
    """TA:44-83: [N] -> [T, window_size] (a copy, not a strided view)."""
    n = wave.shape[0]
    m = frame_count(n, window_size, window_shift, snip_edges)
    if snip_edges:
        if m == 0:
            return np.zeros((0, 0), dtype=wave.dtype)
        src = wave
    else:
        rev = wave[::-1]
        pad = window_size // 2 - window_shift // 2
        if pad > 0:
            src = np.concatenate((rev[-pad:], wave, rev))
        else:
            src = np.concatenate((wave[-pad:], rev))
    idx = np.arange(m)[:, None] * window_shift + np.arange(window_size)[None, :]
    return src[idx]


def window_function(kind: str, size: int, dtype=np.float32, blackman_coeff: float = 0.42) -> np.ndarray:
    """TA:86-113.  periodic=False windows, denominators (size-1)."""
    n = np.arange(size, dtype=np.float64)
    a = 2.0 * math.pi / (size - 1)
    if kind == "hanning":
        w = 0.5 - 0.5 * np.cos(a * n)
    elif kind == "hamming":
        w = 0.54 - 0.46 * np.cos(a * n)
    elif kind == "povey":
        # hann rounded to the working dtype first, then pow(0.85) in that dtype
        w = np.power((0.5 - 0.5 * np.cos(a * n)).astype(dtype), dtype(0.85))
    elif kind == "rectangular":
        w = np.ones(size)
    elif kind == "blackman":
        w = blackman_coeff - 0.5 * np.cos(a * n) + (0.5 - blackman_coeff) * np.cos(2 * a * n)
    else:
        raise Exception("Invalid window type " + kind)
    return w.astype(dtype)


def preprocess_frames(wave: np.ndarray, padded: int, window_size: int, window_shift: int, window_type: str,
                      *, snip_edges=True, dither=0.0, remove_dc_offset=True, preemphasis=0.97,
                      blackman_coeff=0.42, rng: np.random.Generator | None = None) -> np.ndarray:
    """TA:154-217 without the (unused here) log-energy: [N] -> [T, padded]."""
    dtype = wave.dtype.type
    x = frames(wave, window_size, window_shift, snip_edges)
    if x.shape[0] == 0:
        return np.zeros((0, padded), dtype=dtype)
    if dither != 0.0:
        rng = rng or np.random.default_rng(0)
        # one independent Gaussian per (frame, sample): overlapping frames get different noise (TA:179-181)
        x = x + rng.standard_normal(x.shape).astype(dtype) * dtype(dither)
    if remove_dc_offset:
        x = x - x.mean(axis=1, keepdims=True, dtype=dtype)
    if preemphasis != 0.0:
        prev = np.concatenate((x[:, :1], x[:, :-1]), axis=1)  # replicate pad on the left (TA:193-198)
        x = x - dtype(preemphasis) * prev
    x = x * window_function(window_type, window_size, dtype, blackman_coeff)[None, :]
    if padded != window_size:
        x = np.concatenate((x, np.zeros((x.shape[0], padded - window_size), dtype=dtype)), axis=1)
    return x


def mel_scale(freq):
    """TA:326-331: 1127 ln(1 + f/700)."""
    return 1127.0 * np.log(1.0 + freq / 700.0)


def mel_banks(num_bins: int, padded: int, sample_freq: float, low_freq: float = 20.0,
              high_freq: float = 0.0, dtype=np.float32) -> np.ndarray:
    """TA:436-511 (vtln_warp == 1.0 only): [num_bins, padded/2] triangular weights, triangles in the mel domain.

    The scalar edges are Python floats (float64), the per-bin quantities are `dtype`, as in the reference where
    `torch.arange(..)` tensors are float32 and `mel_scale_scalar` is math.log.
    """
    assert num_bins > 3, "Must have at least 3 mel bins"
    assert padded % 2 == 0
    num_fft_bins = padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist
    assert (0.0 <= low_freq < nyquist) and (0.0 < high_freq <= nyquist) and (low_freq < high_freq)
    fft_bin_width = sample_freq / padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = np.arange(num_bins, dtype=dtype)[:, None]
    left = (dtype(mel_low) + b * dtype(delta)).astype(dtype)
    center = (dtype(mel_low) + (b + dtype(1.0)) * dtype(delta)).astype(dtype)
    right = (dtype(mel_low) + (b + dtype(2.0)) * dtype(delta)).astype(dtype)
    mel = mel_scale((dtype(fft_bin_width) * np.arange(num_fft_bins, dtype=dtype)).astype(dtype)).astype(dtype)[None, :]
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return np.maximum(dtype(0.0), np.minimum(up, down)).astype(dtype)


def fbank(wave: np.ndarray, *, num_mel_bins: int = 23, frame_length: float = 25.0, frame_shift: float = 10.0,
          dither: float = 0.0, energy_floor: float = 1.0, window_type: str = "povey",
          sample_frequency: float = 16000.0, snip_edges: bool = True, preemphasis_coefficient: float = 0.97,
          remove_dc_offset: bool = True, round_to_power_of_two: bool = True, low_freq: float = 20.0,
          high_freq: float = 0.0, use_power: bool = True, use_log_fbank: bool = True,
          blackman_coeff: float = 0.42, dtype=np.float32, rng=None) -> np.ndarray:
    """TA:514-645 for use_energy=False, subtract_mean=False, vtln_warp=1: [N] -> [T, num_mel_bins]."""
    del energy_floor  # only enters the (unused) energy column
    wave = np.asarray(wave, dtype=dtype).reshape(-1)
    shift, size, padded = window_properties(wave.shape[0], sample_frequency, frame_shift, frame_length,
                                            round_to_power_of_two)
    x = preprocess_frames(wave, padded, size, shift, window_type, snip_edges=snip_edges, dither=dither,
                          remove_dc_offset=remove_dc_offset, preemphasis=preemphasis_coefficient,
                          blackman_coeff=blackman_coeff, rng=rng)
    spec = np.abs(np.fft.rfft(x, axis=1)).astype(dtype)  # [T, padded/2+1]
    if use_power:
        spec = spec * spec
    bank = mel_banks(num_mel_bins, padded, sample_frequency, low_freq, high_freq, dtype)
    bank = np.concatenate((bank, np.zeros((num_mel_bins, 1), dtype=dtype)), axis=1)  # Nyquist column = 0 (TA:627)
    mel = (spec @ bank.T).astype(dtype)
    if use_log_fbank:
        mel = np.log(np.maximum(mel, dtype(EPSILON_F32))).astype(dtype)
    return mel
