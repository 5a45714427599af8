"""Generates tests/golden/resample_golden.npz with scipy.signal.resample itself - the resampler the reference calls when
scipy is installed (R:voice-service/app/services/voice_interface.py:1022-1027) - on seeded wire PCM, following the
reference's steps around it (:1008-1019 width normalisation and channel mean in float64, :1026 output length, :1045
float32 cast).

    python tests/golden/make_golden_resample.py

Inputs are regenerated from seeds by the tests; only the float32 outputs are stored."""
from pathlib import Path

import numpy as np
from scipy import signal

CASES = {   # name: (dtype, channels, source rate, sample frames) - 240 / 400 ms client chunks and odd lengths
    "s16_mono_48000_240ms": ("int16", 1, 48000, 11520),
    "s16_stereo_48000_400ms": ("int16", 2, 48000, 19200),
    "s16_mono_44100_240ms": ("int16", 1, 44100, 10584),
    "s16_mono_8000_400ms": ("int16", 1, 8000, 3200),
    "s16_mono_22050_odd": ("int16", 1, 22050, 5513),
    "u8_mono_11025": ("uint8", 1, 11025, 2757),
    "s32_stereo_32000": ("int32", 2, 32000, 7681),
}


def wire_pcm(name):
    dtype, ch, rate, n = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    if dtype == "uint8":
        return rng.integers(0, 256, size=n * ch, dtype=np.uint8)
    if dtype == "int16":
        return rng.integers(-20000, 20000, size=n * ch, dtype=np.int16)
    return rng.integers(-2 ** 30, 2 ** 30, size=n * ch, dtype=np.int32)


def reference(name):
    dtype, ch, rate, n = CASES[name]
    raw = wire_pcm(name)
    if dtype == "uint8":
        audio = (raw - 128) / 128.0
    elif dtype == "int16":
        audio = raw / 32768.0
    else:
        audio = raw / 2147483648.0
    if ch > 1:
        audio = np.mean(audio.reshape(-1, ch), axis=1)
    num = int(len(audio) * 16000 / rate)
    return signal.resample(audio, num).astype(np.float32)


if __name__ == "__main__":
    out = {k: reference(k) for k in CASES}
    path = Path(__file__).with_name("resample_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})
