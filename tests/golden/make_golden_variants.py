"""Generates tests/golden/variants_golden.npz from the ACTUAL third-party code (torchaudio.compliance.kaldi.fbank and
vLLM's verbatim funasr WavFrontend) for the configurations beyond the Paraformer one: the FSMN-VAD front-end (LFR 5/1),
Kaldi option variants (frame length / shift, mel count, windows, no pre-emphasis / DC removal, band limits) and
subtract_mean (the CAM++ features).

    python tests/golden/make_golden_variants.py

Inputs are regenerated from seeds by toolbox_for_asr_and_tts_b200.synth; only outputs are stored.  Committed together
with its output; the GPU box never runs this."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_thirdparty as ref  # noqa: E402
from toolbox_for_asr_and_tts_b200 import synth  # noqa: E402

# name -> torchaudio.compliance.kaldi.fbank keyword arguments (besides dither=0, energy_floor=0, 16 kHz)
VARIANTS = {
    "povey_32ms_40mel": dict(frame_length=32.0, frame_shift=10.0, num_mel_bins=40, window_type="povey"),
    "hanning_20ms_64mel_raw": dict(frame_length=20.0, frame_shift=10.0, num_mel_bins=64, window_type="hanning",
                                   preemphasis_coefficient=0.0, remove_dc_offset=False),
    "blackman_band_100_7800": dict(frame_length=25.0, frame_shift=10.0, num_mel_bins=80, window_type="blackman",
                                   low_freq=100.0, high_freq=-200.0),
    "rectangular_shift5": dict(frame_length=25.0, frame_shift=5.0, num_mel_bins=80, window_type="rectangular"),
    "hamming_shift20_24mel": dict(frame_length=25.0, frame_shift=20.0, num_mel_bins=24, window_type="hamming"),
    "subtract_mean_80mel": dict(frame_length=25.0, frame_shift=10.0, num_mel_bins=80, window_type="hamming",
                                subtract_mean=True),
}
VARIANT_SEED, VARIANT_LENS = 55, (16000, 4001)


def vad_cmvn() -> np.ndarray:
    rng = np.random.default_rng(5)
    return np.stack([rng.normal(-8.0, 1.0, 400), rng.uniform(0.2, 0.5, 400)]).astype(np.float32)


def main():
    out = {}
    for name, kw in VARIANTS.items():
        for i, n in enumerate(VARIANT_LENS):
            x = torch.from_numpy(synth.uniform_pcm(VARIANT_SEED, i, n))[None] * 32768.0
            y = kaldi.fbank(x, dither=0.0, energy_floor=0.0, sample_frequency=16000.0, **kw)
            out[f"{name}_{n}"] = y.numpy().astype(np.float32)
    # FSMN-VAD front-end: verbatim funasr WavFrontend, LFR 5/1, its own CMVN
    cm = vad_cmvn()
    w = synth.uniform_pcm(63, 0, 48000)
    feats, lens, impl = ref.reference_forward([w], [48000], cmvn=cm, fs=16000, window="hamming", n_mels=80, frame_length=25,
                                              frame_shift=10, lfr_m=5, lfr_n=1)
    out["vad_5_1_cmvn"] = cm
    out["vad_5_1_feats"] = feats[0]
    out["impl"] = np.array([impl])
    path = Path(__file__).with_name("variants_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
