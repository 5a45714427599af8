"""Seeded, counter-based synthetic PCM shared by tests, smoke() and bench.py.

x[u][n] = amp * (2*U01(mix(seed, u, n)) - 1) with a splitmix64 finaliser; the CUDA kernel
`synth_uniform_kernel` (csrc/aux_kernels.cuh) produces bit-identical values, so the GPU can synthesise a corpus in
place and the CPU oracle can regenerate any utterance of it."""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def _mix(seed: int, u, n):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) * _M1 + np.asarray(u, dtype=np.uint64) * _M2 + np.asarray(n, dtype=np.uint64)
        z ^= z >> np.uint64(30)
        z *= _M2
        z ^= z >> np.uint64(27)
        z *= _M3
        z ^= z >> np.uint64(31)
    return z


def uniform_pcm(seed: int, utt: int, n_samples: int, amp: float = 0.3) -> np.ndarray:
    """float32 [n_samples] in [-amp, amp)."""
    k = (_mix(seed, utt, np.arange(n_samples, dtype=np.uint64)) >> np.uint64(40)).astype(np.int64)
    c = ((k - (1 << 23)).astype(np.float32)) * np.float32(1.0 / 8388608.0)
    return (np.float32(amp) * c).astype(np.float32)


def utterance_lengths(seed: int, batch: int, lo: int = 16000, hi: int = 480000) -> np.ndarray:
    """Lengths uniform in [lo, hi] (1-30 s at 16 kHz by default), int64 [batch]."""
    span = np.uint64(hi - lo + 1)
    z = _mix(seed ^ 0x5EED, np.arange(batch, dtype=np.uint64), np.uint64(0xFFFFFFFF))
    return (np.int64(lo) + (z % span).astype(np.int64)).astype(np.int64)


def packed_offsets(lengths: np.ndarray, align: int = 4) -> tuple[np.ndarray, int]:
    """Start of each utterance in a length-packed buffer, every start aligned to `align` samples (16 bytes)."""
    padded = (np.asarray(lengths, dtype=np.int64) + align - 1) // align * align
    ends = np.cumsum(padded)
    offs = np.concatenate(([0], ends[:-1])).astype(np.int64) if len(padded) else np.zeros(0, dtype=np.int64)
    return offs, int(ends[-1]) if len(padded) else 0
