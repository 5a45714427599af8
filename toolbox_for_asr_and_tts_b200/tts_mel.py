"""TTS-side log-mel feature extraction (BASELINE.json configs[4]): 24 kHz, n_fft = win = 1024, hop 256, 80 mels.

The reference has no audio->mel code (its TTS service is text->wav); the definition is the HiFi-GAN / BigVGAN
`mel_spectrogram` convention frozen in oracle/tts_mel_np.py: periodic Hann, reflect padding (n_fft-hop)/2 per side
(frames = N // hop), magnitude sqrt(re^2+im^2+1e-9), Slaney filters 0..f_max, log(clamp(min=1e-5)), layout
[B, n_mels, frames]."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import _native
from .frontend import _as_length_tensor


class TtsLogMel(nn.Module):
    def __init__(self, sample_rate: int = 24000, n_fft: int = 1024, hop_length: int = 256, n_mels: int = 80,
                 f_min: float = 0.0, f_max: float = 12000.0):
        super().__init__()
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self.f_min, self.f_max = f_min, f_max
        self._ops = None
        self._hs = {}      # one native handle per CUDA device (its tables live there)

    def _handle(self, device):
        dev = torch.device(device).index
        if dev is None:
            dev = torch.cuda.current_device()
        if dev not in self._hs:
            self._ops = _native.ops()
            with torch.cuda.device(dev):
                self._hs[dev] = self._ops.tts_create(self.sample_rate, self.n_fft, self.hop_length, self.n_mels,
                                                     float(self.f_min), float(self.f_max))
        return self._hs[dev]

    def __del__(self):
        try:
            for h in self._hs.values():
                self._ops.tts_destroy(h)
            self._hs = {}
        except Exception:
            pass

    def num_frames(self, n_samples: int) -> int:
        return int(n_samples) // self.hop_length

    def forward(self, waveform: torch.Tensor, lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        """waveform: CUDA float32 [B, Nmax] in [-1, 1]; returns (mel [B, n_mels, max_frames], frames int64 [B])."""
        if not isinstance(waveform, torch.Tensor) or not waveform.is_cuda:
            raise RuntimeError("waveform must be a CUDA tensor: the B200 front-end has no CPU fallback")
        h = self._handle(waveform.device)
        return self._ops.tts_forward(h, waveform.to(torch.float32), None, _as_length_tensor(lengths), self.hop_length,
                                     self.n_mels)
