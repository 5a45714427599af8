"""`WavFrontend`: the reference's front-end interface (upstream funasr WavFrontend, verbatim copy at VF:89-218) on
the B200 kernels.  Same constructor arguments (including upstream's `upsacle_samples` spelling), same
`forward(input, input_lengths) -> (feats_pad, feats_lens)`, `forward_fbank`, `forward_lfr_cmvn`, `output_size`.

Differences, all deliberate:
  * tensors must live on a CUDA device - there is no CPU path (the reference's CPU path is the baseline);
  * the whole batch is one fused kernel pass instead of a Python loop over utterances (VF:137);
  * dither (default 1.0 upstream, 0.0 in the shipped Paraformer configs) uses a counter-based generator, so it
    is reproducible from `dither_seed` but not bit-equal to torch.randn (TA:179-181).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _native
from .cmvn import load_cmvn

WINDOW_TYPES = {"hamming": 0, "hanning": 1, "povey": 2, "rectangular": 3, "blackman": 4}


def _as_length_tensor(input_lengths) -> torch.Tensor:
    if isinstance(input_lengths, torch.Tensor):
        return input_lengths.detach().to("cpu", torch.int64).reshape(-1)
    return torch.as_tensor([int(v) for v in input_lengths], dtype=torch.int64)


class _Handle:
    """Owns one native b200fe handle."""

    def __init__(self, **kw):
        self.ops = _native.ops()
        self.h = self.ops.create(**kw)

    def __del__(self):
        try:
            if getattr(self, "h", 0):
                self.ops.destroy(self.h)
                self.h = 0
        except Exception:
            pass


class WavFrontend(nn.Module):
    """Conventional frontend structure for ASR (drop-in for funasr.frontends.wav_frontend.WavFrontend)."""

    def __init__(
        self,
        cmvn_file: Optional[str] = None,
        fs: int = 16000,
        window: str = "hamming",
        n_mels: int = 80,
        frame_length: int = 25,
        frame_shift: int = 10,
        filter_length_min: int = -1,
        filter_length_max: int = -1,
        lfr_m: int = 1,
        lfr_n: int = 1,
        dither: float = 1.0,
        snip_edges: bool = True,
        upsacle_samples: bool = True,
        cmvn: Optional[torch.Tensor] = None,
        dither_seed: int = 0,
        **kwargs,
    ):
        super().__init__()
        if window not in WINDOW_TYPES:
            raise Exception("Invalid window type " + str(window))   # TA:113
        self.fs = fs
        self.window = window
        self.n_mels = n_mels
        self.frame_length = frame_length
        self.frame_shift = frame_shift
        self.filter_length_min = filter_length_min
        self.filter_length_max = filter_length_max
        self.lfr_m = lfr_m
        self.lfr_n = lfr_n
        self.cmvn_file = cmvn_file
        self.dither = dither
        self.snip_edges = snip_edges
        self.upsacle_samples = upsacle_samples
        self.dither_seed = dither_seed
        # Kaldi fbank options the reference leaves at their defaults (TA:514-541)
        self.preemphasis_coefficient = float(kwargs.pop("preemphasis_coefficient", 0.97))
        self.remove_dc_offset = bool(kwargs.pop("remove_dc_offset", True))
        self.low_freq = float(kwargs.pop("low_freq", 20.0))
        self.high_freq = float(kwargs.pop("high_freq", 0.0))
        self.blackman_coeff = float(kwargs.pop("blackman_coeff", 0.42))
        # Kaldi subtract_mean (TA:539, 642-644): utterance mean normalisation of the fbank, as the CAM++ speaker
        # verification front-end applies it (R:voice_interface.py:2430,2520,2558); only without LFR / CMVN
        self.subtract_mean = bool(kwargs.pop("subtract_mean", False))
        if self.subtract_mean and (lfr_m != 1 or lfr_n != 1 or cmvn is not None or cmvn_file not in (None, "null")):
            raise NotImplementedError("subtract_mean is implemented for plain fbank features (lfr_m = lfr_n = 1, no CMVN)")
        if cmvn is not None:
            self.cmvn = torch.as_tensor(cmvn, dtype=torch.float32)
        elif cmvn_file is not None and cmvn_file != "null":
            self.cmvn = load_cmvn(cmvn_file)
        else:
            self.cmvn = None
        self._handles = {}
        self._calls = 0

    def output_size(self) -> int:
        return self.n_mels * self.lfr_m

    # ------------------------------------------------------------------ native handles
    def _handle(self, lfr: bool, cmvn: bool, fbank_only_cfg: bool = False, device=None) -> _Handle:
        """One native handle per (configuration, CUDA device): its tables live on the device it was created on."""
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: the B200 front-end has no CPU fallback")
        dev = torch.cuda.current_device() if device is None else torch.device(device).index
        if dev is None:
            dev = torch.cuda.current_device()
        key = (lfr, cmvn, fbank_only_cfg, dev)
        if key not in self._handles:
            with torch.cuda.device(dev):
                self._handles[key] = self._new_handle(lfr, cmvn, fbank_only_cfg)
            code = getattr(self, "_kernel_code", 0)
            if code:
                self._handles[key].ops.select_kernel(self._handles[key].h, code)
        return self._handles[key]

    def _new_handle(self, lfr: bool, cmvn: bool, fbank_only_cfg: bool) -> _Handle:
        return _Handle(
            fs=int(self.fs), frame_length=float(self.frame_length), frame_shift=float(self.frame_shift),
            n_mels=int(self.n_mels), window_type=WINDOW_TYPES[self.window],
            lfr_m=int(self.lfr_m) if lfr else 1, lfr_n=int(self.lfr_n) if lfr else 1,
            dither=float(self.dither), snip_edges=bool(self.snip_edges),
            upscale=True if fbank_only_cfg else bool(self.upsacle_samples),
            preemph=self.preemphasis_coefficient, remove_dc=self.remove_dc_offset, low_freq=self.low_freq,
            high_freq=self.high_freq, blackman_coeff=self.blackman_coeff,
            cmvn=self.cmvn if (cmvn and self.cmvn is not None) else None)

    def launch_count(self) -> int:
        return sum(int(h.ops.launch_count(h.h)) for h in self._handles.values())

    def select_kernel(self, which: str) -> None:
        """'auto' (warp-autonomous kernel whenever it applies) or 'tile' (the kernel that also does the statistics
        pass): for A/B measurements and tests; results are identical."""
        self._kernel_code = {"auto": 0, "tile": 1}[which]
        for h in self._handles.values():
            h.ops.select_kernel(h.h, self._kernel_code)

    def profile(self, every) -> None:
        """Bracket every `every`-th launch of the fused kernel with CUDA events on its stream (bench.py roofline);
        True / 1 = every launch, False / 0 = off."""
        h = self._handle(lfr=True, cmvn=True)
        h.ops.profile_enable(h.h, int(every))

    def profile_collect(self):
        """(summed kernel milliseconds, launches) since the last collect."""
        h = self._handle(lfr=True, cmvn=True)
        ms, n = h.ops.profile_collect(h.h)
        return float(ms), int(n)

    @staticmethod
    def _check_cuda(t: torch.Tensor, what: str):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise RuntimeError(f"{what} must be a CUDA tensor: the B200 front-end has no CPU fallback")

    # ------------------------------------------------------------------ reference interface
    def forward(self, input: torch.Tensor, input_lengths, **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """VF:128-168.  input: float32 [B, Nmax] in [-1, 1]; returns ([B, max T_lfr, n_mels*lfr_m], int64 [B])."""
        self._check_cuda(input, "input")
        lens = _as_length_tensor(input_lengths)
        h = self._handle(lfr=True, cmvn=True, device=input.device)
        self._calls += 1
        feats, feat_lens = h.ops.forward(h.h, self._pcm(input), None, lens, 0, kwargs.get("stats"),
                                         int(self.dither_seed + self._calls))
        if self.subtract_mean:
            h.ops.subtract_column_mean(feats, feat_lens)
        return feats, self._host_lengths(h, lens)

    @staticmethod
    def _host_lengths(h: "_Handle", lens: torch.Tensor) -> torch.Tensor:
        """feats_lens as the reference returns them: an int64 CPU tensor (VF:160, torch.as_tensor of a Python list).
        The counts are a function of the input lengths alone (TA:65-70, VF:43), so they come from the host-side plan:
        no device read-back, no synchronisation.  (`forward_packed` returns the kernel's device-side copy instead.)"""
        return h.ops.plan(h.h, lens)[1]

    @staticmethod
    def _pcm(x: torch.Tensor) -> torch.Tensor:
        """float32 in [-1, 1] (the reference's input) or int16 PCM still in its wire format (value = s / 32768, the
        conversion of R:voice_interface.py:1008-1013, fused into the kernel's loads: same result, half the bytes)."""
        return x if x.dtype == torch.int16 else x.to(torch.float32)

    def forward_packed(self, wave: torch.Tensor, offsets, lengths, stats: Optional[torch.Tensor] = None,
                       rows_cap: int = 0, pad: bool = True):
        """Length-packed batch: `wave` is one flat CUDA buffer, utterance i occupies wave[offsets[i] : +lengths[i]].
        Same outputs as `forward` (feature lengths on the device).  `stats` (CUDA float64 [2*D+1]) accumulates global
        CMVN statistics of the un-normalised LFR features.
        pad=False: ROWS-PACKED output for consumers that batch by length themselves - returns (feats [sum of rows, D],
        feature lengths on the device, row_offsets int64 [B+1] on the CPU); utterance i is feats[row_offsets[i] :
        row_offsets[i+1]].  No padding rows exist, so none are written."""
        self._check_cuda(wave, "wave")
        h = self._handle(lfr=True, cmvn=True, device=wave.device)
        self._calls += 1
        lens = _as_length_tensor(lengths)
        if pad:
            return h.ops.forward(h.h, self._pcm(wave), _as_length_tensor(offsets), lens, int(rows_cap), stats,
                                 int(self.dither_seed + self._calls))
        feats, feat_lens = h.ops.forward(h.h, self._pcm(wave), _as_length_tensor(offsets), lens, -1, stats,
                                         int(self.dither_seed + self._calls))
        rows = self._host_lengths(h, lens)
        return feats, feat_lens, torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(rows, 0)))

    def forward_fbank(self, input: torch.Tensor, input_lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        """VF:170-196: Kaldi fbank only (always upscaled by 2^15 upstream), zero-padded, lengths int64."""
        self._check_cuda(input, "input")
        h = self._handle(lfr=False, cmvn=False, fbank_only_cfg=True, device=input.device)
        self._calls += 1
        lens = _as_length_tensor(input_lengths)
        feats, feat_lens = h.ops.forward(h.h, self._pcm(input), None, lens, 0, None, int(self.dither_seed + self._calls))
        if self.subtract_mean:
            h.ops.subtract_column_mean(feats, feat_lens)
        return feats, self._host_lengths(h, lens)

    def forward_lfr_cmvn(self, input: torch.Tensor, input_lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        """VF:198-218: LFR + CMVN of given [B, T, n_mels] features."""
        self._check_cuda(input, "input")
        lens = _as_length_tensor(input_lengths)
        h = self._handle(lfr=True, cmvn=True, device=input.device)
        feats, feat_lens = h.ops.lfr_cmvn(h.h, input.to(torch.float32), lens)
        rows = int(-(-int(lens.max()) // self.lfr_n)) if lens.numel() else 0
        return feats[:, :rows], -(-lens // self.lfr_n)      # int64 on the CPU, as VF:218

    @staticmethod
    def audio_statistics(wave: torch.Tensor, lengths, offsets=None, clip_level: float = 0.999) -> torch.Tensor:
        """Per-utterance statistics the reference logs / gates on around its funasr calls (_log_audio_statistics,
        R:voice_interface.py:873-939; energy gate :1298-1300, :1569-1570), for a whole ragged batch on the GPU.
        Returns float64 [B, 6]: max, min, mean |x|, rms, clipping ratio (|x| >= clip_level), max |x|."""
        WavFrontend._check_cuda(wave, "wave")
        from . import _native
        return _native.ops().audio_stats(wave.to(torch.float32), None if offsets is None else _as_length_tensor(offsets),
                                         _as_length_tensor(lengths), float(clip_level))

    @staticmethod
    def ingest_pcm(pcm: torch.Tensor, channels: int = 1, src_rate: int = 16000, dst_rate: int = 16000,
                   method: str = "scipy") -> torch.Tensor:
        """Wire PCM (CUDA uint8 / int16 / int32, interleaved channels) -> float32 mono at dst_rate, exactly as the
        reference's base64_to_audio_np does after the WAV header (R:voice_interface.py:1004-1045): width normalisation,
        channel mean, resampling, float32.  method = "scipy": scipy.signal.resample, the Fourier method the reference
        uses whenever scipy is installed (:1022-1027; float64 DFT sums on the GPU, equal to scipy's result up to float64
        round-off); method = "interp": its numpy fallback np.interp (:1028-1034), bit-identical to numpy."""
        WavFrontend._check_cuda(pcm, "pcm")
        from . import _native
        if method not in ("scipy", "interp"):
            raise ValueError("method must be 'scipy' or 'interp'")
        op = _native.ops().ingest_pcm_fft if method == "scipy" else _native.ops().ingest_pcm
        return op(pcm, int(channels), int(src_rate), int(dst_rate))

    # ------------------------------------------------------------------ extras used by tests / tools
    def frame_counts(self, input_lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        """(frames, LFR rows) per utterance, as the reference would produce (TA:65-70, VF:43)."""
        h = self._handle(lfr=True, cmvn=True)
        return h.ops.plan(h.h, _as_length_tensor(input_lengths))

    def tables(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """(window [frame_samples], mel bank [n_mels, n_fft/2]) as built by the native library."""
        h = self._handle(lfr=True, cmvn=True)
        win, mel = h.ops.get_tables(h.h)
        return win, mel[: self.n_mels]
