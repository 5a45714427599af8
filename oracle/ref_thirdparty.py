"""Runs the ACTUAL third-party code the reference executes for this path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference repository holds no front-end arithmetic; its
`funasr.AutoModel(...).generate(...)` calls (R:voice-service/app/services/voice_interface.py:422,457,1370,1585,
2049) run upstream `WavFrontend.forward` over `torchaudio.compliance.kaldi.fbank`.  Both are importable in this
image (torchaudio directly; WavFrontend as vLLM's verbatim copy, VF), so this module

  * `reference_forward`: the reference's CPU front-end, per-utterance Python loop as upstream (VF:137).  Prefers the
    verbatim class (`impl="vllm-funasr"`); falls back to the same loop written here around the real
    `kaldi.fbank` (`impl="torchaudio"`).  This is what bench.py times as `cpu_baseline.kind == "reference"`.
  * is what tests/golden/make_golden.py uses to generate the committed golden vectors.

Nothing here is imported by the product package.
"""
from __future__ import annotations

import os
import tempfile

import numpy as np


def have_torchaudio() -> bool:
    try:
        import torchaudio.compliance.kaldi  # noqa: F401
        return True
    except Exception:
        return False


_VLLM_FRONTEND = None


def vllm_wavfrontend_cls():
    """The verbatim upstream WavFrontend class (VF:89-218), or None when vllm does not import."""
    global _VLLM_FRONTEND
    if _VLLM_FRONTEND is None:
        try:
            from vllm.transformers_utils.processors.funasr import WavFrontend  # type: ignore
            _VLLM_FRONTEND = WavFrontend
        except Exception:
            _VLLM_FRONTEND = False
    return _VLLM_FRONTEND or None


def _torch_apply_lfr(mat, lfr_m, lfr_n):
    import torch
    t = mat.shape[0]
    rows = -(-t // lfr_n)
    left = (lfr_m - 1) // 2
    idx = (torch.arange(rows)[:, None] * lfr_n + torch.arange(lfr_m)[None, :] - left).clamp_(0, t - 1)
    return mat[idx].reshape(rows, -1).to(torch.float32)


class TorchaudioFrontend:
    """WavFrontend.forward semantics (VF:128-168) around the real torchaudio kaldi.fbank."""

    impl = "torchaudio"

    def __init__(self, cmvn: np.ndarray | None, **conf):
        import torch
        self.conf = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=1, lfr_n=1,
                         dither=0.0, snip_edges=True, upsacle_samples=True)
        self.conf.update(conf)
        self.cmvn = None if cmvn is None else torch.as_tensor(np.asarray(cmvn, dtype=np.float32))

    def __call__(self, waves, lengths):
        import torch
        import torchaudio.compliance.kaldi as kaldi
        from torch.nn.utils.rnn import pad_sequence
        c = self.conf
        feats = []
        with torch.no_grad():
            for i in range(len(lengths)):
                n = int(lengths[i])
                w = torch.as_tensor(waves[i])[:n]
                if c["upsacle_samples"]:
                    w = w * (1 << 15)
                mat = kaldi.fbank(w.unsqueeze(0), num_mel_bins=c["n_mels"],
                                  frame_length=min(c["frame_length"], n / c["fs"] * 1000),
                                  frame_shift=c["frame_shift"], dither=c["dither"], energy_floor=0.0,
                                  window_type=c["window"], sample_frequency=c["fs"], snip_edges=c["snip_edges"])
                if c["lfr_m"] != 1 or c["lfr_n"] != 1:
                    mat = _torch_apply_lfr(mat, c["lfr_m"], c["lfr_n"])
                if self.cmvn is not None:
                    mat = (mat + self.cmvn[0:1, :mat.shape[1]]) * self.cmvn[1:2, :mat.shape[1]]
                feats.append(mat)
        lens = torch.as_tensor([m.shape[0] for m in feats])
        return pad_sequence(feats, batch_first=True, padding_value=0.0), lens


class VllmFunasrFrontend:
    """The verbatim upstream class; cmvn goes through a temporary am.mvn so that load_cmvn (VF:63-86) is exercised."""

    impl = "vllm-funasr"

    def __init__(self, cmvn: np.ndarray | None, **conf):
        from .wav_frontend_np import write_cmvn
        cls = vllm_wavfrontend_cls()
        if cls is None:
            raise ImportError("vllm funasr processor is not importable")
        conf = dict(conf)
        conf.setdefault("dither", 0.0)
        path = None
        if cmvn is not None:
            fd, path = tempfile.mkstemp(suffix=".mvn")
            os.close(fd)
            write_cmvn(path, cmvn[0], cmvn[1])
        try:
            self.fe = cls(cmvn_file=path, **conf)
        finally:
            if path:
                os.unlink(path)

    def __call__(self, waves, lengths):
        import torch
        with torch.no_grad():
            if isinstance(waves, (list, tuple)):
                nmax = max(int(n) for n in lengths)
                buf = torch.zeros(len(lengths), nmax)
                for i, n in enumerate(lengths):
                    buf[i, :int(n)] = torch.as_tensor(waves[i])[:int(n)]
                waves = buf
            feats, lens = self.fe(torch.as_tensor(waves), [int(n) for n in lengths])
        return feats, lens


def make_reference_frontend(cmvn=None, prefer_vllm: bool = True, **conf):
    if prefer_vllm and vllm_wavfrontend_cls() is not None:
        return VllmFunasrFrontend(cmvn, **conf)
    return TorchaudioFrontend(cmvn, **conf)


def reference_forward(waves, lengths, cmvn=None, prefer_vllm: bool = True, **conf):
    fe = make_reference_frontend(cmvn, prefer_vllm, **conf)
    feats, lens = fe(waves, lengths)
    return feats.numpy(), lens.numpy().astype(np.int64), fe.impl
