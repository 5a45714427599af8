"""Generates tests/golden/frontend_golden.npz from the ACTUAL third-party code the reference runs
(torchaudio.compliance.kaldi.fbank + vLLM's verbatim funasr WavFrontend), in the build container.

    python tests/golden/make_golden.py

The inputs are regenerated from seeds by toolbox_for_asr_and_tts_b200.synth, so only outputs are stored.
Committed together with its output so the fixtures can be re-derived; the GPU box never runs this."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ref_thirdparty as ref  # noqa: E402
from toolbox_for_asr_and_tts_b200 import synth  # noqa: E402

SEED = 1234
LENGTHS = [399, 400, 401, 559, 560, 1000, 16000, 160000, 480000]
PARAFORMER = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6)


def synthetic_cmvn(dim: int, seed: int = 7) -> np.ndarray:
    u = (synth.uniform_pcm(seed, 1000, dim, amp=1.0) + 1.0) * 0.5
    v = (synth.uniform_pcm(seed, 1001, dim, amp=1.0) + 1.0) * 0.5
    return np.stack([-(8.0 + 4.0 * u), 0.2 + 0.3 * v]).astype(np.float32)


def main():
    out = {}
    cmvn = synthetic_cmvn(560)
    out["cmvn"] = cmvn
    impl_used = set()
    for n in LENGTHS:
        x = synth.uniform_pcm(SEED, n, n)            # utterance id = its length
        feats, lens, impl = ref.reference_forward([x], [n], cmvn=cmvn, **PARAFORMER)
        impl_used.add(impl)
        out[f"paraformer_{n}"] = feats[0]
        # the same through the plain torchaudio loop: must agree bit for bit with the verbatim class
        f2, l2, _ = ref.reference_forward([x], [n], cmvn=cmvn, prefer_vllm=False, **PARAFORMER)
        assert np.array_equal(feats, f2) and np.array_equal(lens, l2), n
    for n in (400, 16000):
        x = synth.uniform_pcm(SEED, n, n)
        conf = dict(PARAFORMER, window="povey")
        feats, _, _ = ref.reference_forward([x], [n], cmvn=cmvn, **conf)
        out[f"povey_{n}"] = feats[0]
    x = synth.uniform_pcm(SEED, 16000, 16000)
    feats, _, _ = ref.reference_forward([x], [16000], cmvn=None, fs=16000, window="hamming", n_mels=80, frame_length=25,
                                        frame_shift=10, lfr_m=1, lfr_n=1)
    out["fbank_16000"] = feats[0]
    # a small ragged batch through the verbatim class (pad_sequence, lengths)
    lens = [16000, 4000, 48000, 399, 8000]
    waves = [synth.uniform_pcm(SEED + 1, i, n) for i, n in enumerate(lens)]
    feats, flens, _ = ref.reference_forward(waves, lens, cmvn=cmvn, **PARAFORMER)
    out["batch_feats"] = feats
    out["batch_lens"] = flens
    out["batch_input_lens"] = np.array(lens, dtype=np.int64)
    # a second distribution: 0.1*N(0,1) clipped, with a DC offset (exercises remove_dc / pre-emphasis edges)
    rng = np.random.default_rng(5)
    g = np.clip(0.1 * rng.standard_normal(24000) + 0.05, -1, 1).astype(np.float32)
    feats, _, _ = ref.reference_forward([g], [24000], cmvn=cmvn, **PARAFORMER)
    out["gauss_input"] = g
    out["gauss_feats"] = feats[0]
    out["impl"] = np.array(sorted(impl_used))
    path = Path(__file__).with_name("frontend_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
