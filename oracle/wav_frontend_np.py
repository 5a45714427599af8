"""numpy restatement of the FunASR front-end the reference reaches through funasr.AutoModel.

TEST INFRASTRUCTURE (see oracle/__init__.py).  VF = vllm/transformers_utils/processors/funasr.py, the verbatim
copy of upstream funasr/frontends/wav_frontend.py that is on disk in this image.

  lfr_num_rows / apply_lfr     VF:40-60
  apply_cmvn                   VF:23-37
  load_cmvn / write_cmvn       VF:63-86  (Kaldi-nnet text: <AddShift>, <Rescale>)
  frontend_forward             VF:128-168 (WavFrontend.forward), forward_fbank VF:170-196
  OnlineFrontend               UPSTREAM-RECALLED WavFrontendOnline (funasr wav_frontend.py; not on disk):
                               input_cache sample carry + lfr splice carry.  Pinned by the invariant
                               concat(stream outputs) == offline output; the final flush always emits
                               ceil(T/lfr_n) rows in total (superset of upstream's tiny-last-chunk quirk,
                               SURVEY.md section 5.7).
  cmvn_stats / stats_to_cmvn   UPSTREAM-RECALLED funasr/bin/compute_audio_cmvn.py (sum, sum of squares, count
                               over LFR features without CMVN -> AddShift=-mean, Rescale=1/std).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

from . import kaldi_fbank_np as kf


# --------------------------------------------------------------------------- LFR / CMVN
def lfr_num_rows(num_frames: int, lfr_n: int) -> int:
    """VF:43: T_lfr = ceil(T / lfr_n)."""
    return int(math.ceil(num_frames / lfr_n)) if num_frames > 0 else 0


def apply_lfr(feats: np.ndarray, lfr_m: int, lfr_n: int) -> np.ndarray:
    """VF:40-60 in closed form: out[i, j*D + d] = feats[clamp(lfr_n*i + j - (lfr_m-1)//2, 0, T-1), d].

    (left pad = (lfr_m-1)//2 copies of frame 0, right pad = copies of the last frame.)
    """
    t, d = feats.shape
    rows = lfr_num_rows(t, lfr_n)
    left = (lfr_m - 1) // 2
    idx = np.arange(rows)[:, None] * lfr_n + np.arange(lfr_m)[None, :] - left
    idx = np.clip(idx, 0, t - 1)
    return feats[idx].reshape(rows, lfr_m * d).astype(np.float32)


def apply_lfr_literal(feats: np.ndarray, lfr_m: int, lfr_n: int) -> np.ndarray:
    """VF:40-60 step by step (explicit padding then strided windows); cross-checks the closed form."""
    t = feats.shape[0]
    rows = int(np.ceil(t / lfr_n))
    left = (lfr_m - 1) // 2
    padded = np.concatenate([np.repeat(feats[:1], left, axis=0), feats], axis=0)
    need = (rows - 1) * lfr_n + lfr_m
    if need > padded.shape[0]:
        padded = np.concatenate([padded, np.repeat(padded[-1:], need - padded.shape[0], axis=0)], axis=0)
    out = np.stack([padded[i * lfr_n:i * lfr_n + lfr_m].reshape(-1) for i in range(rows)], axis=0)
    return out.astype(np.float32)


def apply_cmvn(feats: np.ndarray, cmvn: np.ndarray) -> np.ndarray:
    """VF:23-37: (x + cmvn[0]) * cmvn[1], float32, in that order."""
    dim = feats.shape[1]
    out = feats.astype(np.float32) + cmvn[0:1, :dim].astype(np.float32)
    out = out * cmvn[1:2, :dim].astype(np.float32)
    return out.astype(np.float32)


def load_cmvn(path: str) -> np.ndarray:
    """VF:63-86: the line after <AddShift>/<Rescale> is '<LearnRateCoef> 0 [ v ... ]'; tokens [3:-1] are the values."""
    with open(path, encoding="utf-8") as f:
        lines = f.readlines()
    shift, scale = [], []
    for i, line in enumerate(lines):
        tok = line.split()
        if not tok:
            continue
        if tok[0] in ("<AddShift>", "<Rescale>"):
            nxt = lines[i + 1].split()
            if nxt[0] == "<LearnRateCoef>":
                vals = nxt[3:len(nxt) - 1]
                if tok[0] == "<AddShift>":
                    shift = vals
                else:
                    scale = vals
    return np.array([np.array(shift).astype(np.float32), np.array(scale).astype(np.float32)], dtype=np.float32)


def write_cmvn(path: str, shift: np.ndarray, scale: np.ndarray) -> None:
    """Write a Kaldi-nnet text am.mvn that load_cmvn (and VF:63-86) parses back."""
    d = len(shift)

    def row(v):
        return " ".join(repr(float(np.float32(x))) for x in v)

    with open(path, "w", encoding="utf-8") as f:
        f.write("<Nnet> \n")
        f.write(f"<Splice> {d} {d}\n[ 0 ]\n")
        f.write(f"<AddShift> {d} {d} \n")
        f.write(f"<LearnRateCoef> 0 [ {row(shift)} ]\n")
        f.write(f"<Rescale> {d} {d}\n")
        f.write(f"<LearnRateCoef> 0 [ {row(scale)} ]\n")
        f.write("</Nnet> \n")


# --------------------------------------------------------------------------- offline front-end
def frontend_forward(waves: Sequence[np.ndarray] | np.ndarray, lengths: Sequence[int], *, cmvn: np.ndarray | None = None,
                     fs: int = 16000, window: str = "hamming", n_mels: int = 80, frame_length: float = 25,
                     frame_shift: float = 10, lfr_m: int = 1, lfr_n: int = 1, dither: float = 0.0,
                     snip_edges: bool = True, upscale: bool = True, dtype=np.float32, rng=None):
    """VF:128-168: per-utterance fbank -> LFR -> CMVN, zero-padded to the longest, lengths int64."""
    feats = []
    for i, n in enumerate(lengths):
        n = int(n)
        w = np.asarray(waves[i][:n], dtype=dtype)
        if upscale:
            w = w * dtype(1 << 15)
        mat = kf.fbank(w, num_mel_bins=n_mels, frame_length=min(frame_length, n / fs * 1000),
                       frame_shift=frame_shift, dither=dither, energy_floor=0.0, window_type=window,
                       sample_frequency=fs, snip_edges=snip_edges, dtype=dtype, rng=rng)
        if lfr_m != 1 or lfr_n != 1:
            mat = apply_lfr(mat, lfr_m, lfr_n) if dtype == np.float32 else _lfr_keep_dtype(mat, lfr_m, lfr_n)
        if cmvn is not None:
            mat = apply_cmvn(mat, cmvn) if dtype == np.float32 else (mat + cmvn[0:1]) * cmvn[1:2]
        feats.append(mat)
    lens = np.array([m.shape[0] for m in feats], dtype=np.int64)
    dim = n_mels * lfr_m
    out = np.zeros((len(feats), int(lens.max()) if len(feats) else 0, dim), dtype=feats[0].dtype if feats else dtype)
    for i, m in enumerate(feats):
        out[i, :m.shape[0]] = m
    return out, lens


def _lfr_keep_dtype(feats, lfr_m, lfr_n):
    t, d = feats.shape
    rows = lfr_num_rows(t, lfr_n)
    idx = np.clip(np.arange(rows)[:, None] * lfr_n + np.arange(lfr_m)[None, :] - (lfr_m - 1) // 2, 0, t - 1)
    return feats[idx].reshape(rows, lfr_m * d)


# --------------------------------------------------------------------------- streaming front-end
class OnlineFrontend:
    """Chunked front-end with sample carry and LFR splice carry (one stream).

    State: `carry` = samples after the last hop of the last full frame (upstream `input_cache`), `frames` = log-mel
    frames not yet fully consumed by an emitted LFR row (upstream `lfr_splice_cache`, without materialising the
    left-pad copies), `t_seen` frames so far, `rows_out` rows so far.  Row i is emitted as soon as frame
    lfr_n*i + lfr_m-1 - (lfr_m-1)//2 exists; on `is_final` the remaining rows up to ceil(T/lfr_n) are emitted with
    the last frame replicated.
    """

    def __init__(self, *, cmvn=None, fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10,
                 lfr_m=1, lfr_n=1, upscale=True, dtype=np.float32):
        self.cmvn, self.fs, self.window, self.n_mels = cmvn, fs, window, n_mels
        self.frame_length, self.frame_shift, self.lfr_m, self.lfr_n = frame_length, frame_shift, lfr_m, lfr_n
        self.upscale, self.dtype = upscale, dtype
        self.win = int(frame_length * fs / 1000)
        self.hop = int(frame_shift * fs / 1000)
        self.reset()

    def reset(self):
        self.carry = np.zeros(0, dtype=self.dtype)
        self.frames = np.zeros((0, self.n_mels), dtype=self.dtype)
        self.base = 0      # absolute index of self.frames[0]
        self.t_seen = 0
        self.rows_out = 0

    def push(self, chunk: np.ndarray, is_final: bool = False) -> np.ndarray:
        x = np.concatenate((self.carry, np.asarray(chunk, dtype=self.dtype)))
        nf = (x.shape[0] - self.win) // self.hop + 1 if x.shape[0] >= self.win else 0
        if nf > 0:
            w = x[:(nf - 1) * self.hop + self.win]
            if self.upscale:
                w = w * self.dtype(1 << 15)
            new = kf.fbank(w, num_mel_bins=self.n_mels, frame_length=self.frame_length, frame_shift=self.frame_shift,
                           dither=0.0, energy_floor=0.0, window_type=self.window, sample_frequency=self.fs,
                           dtype=self.dtype)
            self.frames = np.concatenate((self.frames, new), axis=0)
            self.t_seen += nf
        self.carry = x[nf * self.hop:]
        left = (self.lfr_m - 1) // 2
        t = self.t_seen
        if is_final:
            rows_total = lfr_num_rows(t, self.lfr_n)
        else:
            last_needed = self.lfr_m - 1 - left  # offset of the newest frame row i needs
            rows_total = 0 if t - 1 < last_needed else (t - 1 - last_needed) // self.lfr_n + 1
            rows_total = min(rows_total, lfr_num_rows(t, self.lfr_n))
        out = []
        for i in range(self.rows_out, rows_total):
            idx = np.clip(np.arange(self.lfr_m) + i * self.lfr_n - left, 0, t - 1) - self.base
            out.append(self.frames[idx].reshape(-1))
        self.rows_out = max(self.rows_out, rows_total)
        keep_from = max(self.rows_out * self.lfr_n - left, 0)
        keep_from = min(keep_from, max(t - 1, 0))  # always keep the newest frame for right replication
        if keep_from > self.base:
            self.frames = self.frames[keep_from - self.base:]
            self.base = keep_from
        dim = self.n_mels * self.lfr_m
        rows = np.stack(out, axis=0).astype(np.float32) if out else np.zeros((0, dim), dtype=np.float32)
        if self.cmvn is not None and rows.shape[0]:
            rows = apply_cmvn(rows, self.cmvn)
        if is_final:
            self.reset()
        return rows


# --------------------------------------------------------------------------- global CMVN statistics
def cmvn_stats(feature_mats: Sequence[np.ndarray]):
    """sum, sum of squares (float64) and frame count over un-normalised LFR features."""
    dim = feature_mats[0].shape[1]
    s = np.zeros(dim, dtype=np.float64)
    s2 = np.zeros(dim, dtype=np.float64)
    n = 0
    for m in feature_mats:
        m64 = m.astype(np.float64)
        s += m64.sum(axis=0)
        s2 += (m64 * m64).sum(axis=0)
        n += m.shape[0]
    return s, s2, n


def stats_to_cmvn(s: np.ndarray, s2: np.ndarray, n: int) -> np.ndarray:
    """AddShift = -mean, Rescale = 1/sqrt(E[x^2] - mean^2)  ->  float32 [2, dim]."""
    mean = s / n
    var = s2 / n - mean * mean
    return np.stack([-mean, 1.0 / np.sqrt(var)]).astype(np.float32)


# --------------------------------------------------------------------------- host-side by-products of the reference
def audio_statistics(audio: np.ndarray, clip_level: float = 0.999) -> np.ndarray:
    """R:voice-service/app/services/voice_interface.py:873-939 (_log_audio_statistics), the fields a consumer gates on:
    [max, min, mean |x| (:1298, :1569), rms (:897, without its +1e-10), clipping ratio (:900-901), max |x| (:894)]."""
    a = np.asarray(audio)
    if a.size == 0:
        return np.zeros(6)
    a64 = a.astype(np.float64)
    return np.array([a64.max(), a64.min(), np.abs(a64).mean(), np.sqrt((a64 ** 2).mean()),
                     float((np.abs(a) >= np.float32(clip_level)).sum()) / a.size, np.abs(a64).max()])


def subtract_column_mean(feats: np.ndarray) -> np.ndarray:
    """TA:642-644 (_subtract_column_mean): mean over frames, per mel bin."""
    return (feats - feats.mean(axis=0, keepdims=True)).astype(feats.dtype)


def resample_fourier(x: np.ndarray, num: int) -> np.ndarray:
    """scipy.signal.resample(x, num) for real 1-D input, window=None, restated with numpy's FFT (scipy 1.18,
    scipy/signal/_signaltools.py::resample): keep rfft bins 0 .. min(n, num)//2, the unpaired middle bin x2 when
    shortening / x0.5 when lengthening, irfft(X / (n / num), n=num).  float64 in, float64 out."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    m = min(num, n)
    X = np.fft.rfft(x)[: m // 2 + 1].copy()
    if m % 2 == 0 and num != n:
        X[m // 2] *= 2 if num < n else 0.5
    return np.fft.irfft(X / (n / num), n=num)


def ingest_pcm(raw: np.ndarray, channels: int, src_rate: int, dst_rate: int = 16000, method: str = "interp") -> np.ndarray:
    """R:voice-service/app/services/voice_interface.py:1004-1045 (base64_to_audio_np after the WAV header).  raw: uint8 /
    int16 / int32 interleaved samples.  method "scipy": the scipy.signal.resample branch (:1022-1027, what runs when
    scipy is installed); "interp": the numpy fallback (:1028-1034)."""
    if raw.dtype == np.uint8:
        audio = (raw - 128) / 128.0          # :1007, uint8 arithmetic wraps exactly as upstream's does
    elif raw.dtype == np.int16:
        audio = raw / 32768.0                # :1010
    elif raw.dtype == np.int32:
        audio = raw / 2147483648.0           # :1013
    else:
        raise ValueError(f"unsupported sample width: {raw.dtype}")
    if channels > 1:
        audio = np.mean(audio.reshape(-1, channels), axis=1)     # :1019
    if src_rate != dst_rate:
        old_length = len(audio)
        new_length = int(old_length * dst_rate / src_rate)        # :1026 / :1030
        if method == "scipy":
            audio = resample_fourier(audio, new_length)           # :1027
        else:
            old_indices = np.linspace(0, old_length - 1, old_length)
            new_indices = np.linspace(0, old_length - 1, new_length)
            audio = np.interp(new_indices, old_indices, audio)    # :1031-1034
    return audio.astype(np.float32)                                # :1045
