"""Error report (run on the GPU box): fbank-only log-mel of the CUDA path against (a) the committed float32 golden of
the reference and (b) the float64 oracle, with the reference's own float32-vs-float64 error beside it.

    python tests/parity_report.py            # prints one line per case
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import kaldi_fbank_np as kf  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, synth  # noqa: E402

PARAFORMER = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6)


def main():
    golden = dict(np.load(ROOT / "tests" / "golden" / "frontend_golden.npz"))
    fe = WavFrontend(cmvn=None, **dict(PARAFORMER, dither=0.0))
    cases = [("uniform_16000", synth.uniform_pcm(1234, 16000, 16000)),
             ("uniform_160000", synth.uniform_pcm(1234, 160000, 160000))]
    rng = np.random.default_rng(7)
    cases.append(("gauss0.1_48000", np.clip(0.1 * rng.standard_normal(48000), -1, 1).astype(np.float32)))
    t = np.arange(48000) / 16000.0
    cases.append(("speechlike_48000", (0.3 * np.sin(2 * np.pi * 220 * t) * (1 + 0.5 * np.sin(2 * np.pi * 3 * t))
                                       + 0.01 * rng.standard_normal(48000)).astype(np.float32)))
    for name, x in cases:
        n = len(x)
        got = fe.forward_fbank(torch.from_numpy(x)[None].cuda(), [n])[0][0].cpu().numpy().astype(np.float64)
        kw = dict(num_mel_bins=80, frame_length=25.0, frame_shift=10.0, dither=0.0, energy_floor=0.0,
                  window_type="hamming", sample_frequency=16000.0)
        r32 = kf.fbank(x * np.float32(32768.0), dtype=np.float32, **kw).astype(np.float64)
        r64 = kf.fbank(x.astype(np.float64) * 32768.0, dtype=np.float64, **kw)
        e_gpu64 = np.abs(got - r64)
        e_gpu32 = np.abs(got - r32)
        e_ref = np.abs(r32 - r64)
        i = np.unravel_index(np.argmax(e_gpu64), e_gpu64.shape)
        line = (f"{name:18s} gpu-vs-f64 max {e_gpu64.max():.2e} mean {e_gpu64.mean():.2e} at {i} (value {r64[i]:.2f}, row median "
                f"{np.median(r64[i[0]]):.2f}) | gpu-vs-f32oracle max {e_gpu32.max():.2e} | f32oracle-vs-f64 max {e_ref.max():.2e} mean {e_ref.mean():.2e}")
        if name == "uniform_16000":
            line += f" | gpu-vs-golden max {np.abs(got - golden['fbank_16000']).max():.2e}"
        print(line)


if __name__ == "__main__":
    main()
