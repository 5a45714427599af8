"""Streaming front-end on device-resident state.

`StreamPool` is the B200-native form: S concurrent streams share one HBM state slab (sample carry, LFR splice frames,
counters per stream) and every tick is ONE kernel launch over all streams that received a chunk.  It replaces both the
upstream per-connection `WavFrontendOnline` cache dict and the reference's host-side np.concatenate accumulation
(R:voice-service/app/services/voice_interface.py:1304-1311,1688-1746).

`WavFrontendOnline` keeps upstream's single-stream call shape - forward(input[1, n], input_lengths, cache=dict,
is_final=bool) -> (feats[1, k, D], feats_lengths) - on top of a one-stream pool whose state tensor lives in `cache`.
Semantics: concat(stream outputs) == offline output; the final flush always completes ceil(T/lfr_n) rows.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .frontend import WavFrontend, _as_length_tensor


class StreamPool:
    def __init__(self, frontend: WavFrontend, n_streams: int, max_chunk_samples: int, device="cuda"):
        self.frontend = frontend
        self.n_streams = int(n_streams)
        self.max_chunk = int(max_chunk_samples)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("StreamPool state must live on a CUDA device (no CPU fallback)")
        self._h = frontend._handle(lfr=True, cmvn=True, device=self.device)
        ops = self._h.ops
        self.state = ops.stream_state(self._h.h, self.n_streams, self.max_chunk, self.device)
        self.rows_cap = int(ops.stream_max_rows(self._h.h, self.max_chunk))
        self.reset()

    def reset(self, stream_ids: Optional[torch.Tensor] = None) -> None:
        self._h.ops.stream_reset(self._h.h, self.state, self.n_streams, self.max_chunk, stream_ids)

    def push(self, chunks: torch.Tensor, chunk_lens: torch.Tensor, stream_ids: torch.Tensor,
             is_final: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """chunks: CUDA float32 [n, <=max_chunk]; returns (feats [n, rows_cap, D], rows int32 [n]) on the device,
        without synchronising.  stream_ids must be distinct within one call."""
        if not chunks.is_cuda:
            raise RuntimeError("chunks must be a CUDA tensor: the B200 front-end has no CPU fallback")
        return self._h.ops.stream_push(self._h.h, self.state, self.n_streams, self.max_chunk, chunks, chunk_lens,
                                       stream_ids, is_final)

    # the reference's energy gate constants (R:voice-service/app/services/voice_interface.py:656-658)
    VAD_ENERGY_THRESHOLD, VAD_MAX_THRESHOLD = 0.03, 0.17

    def push_with_speech_flags(self, chunks: torch.Tensor, chunk_lens: torch.Tensor, stream_ids: torch.Tensor,
                               is_final: Optional[torch.Tensor] = None, energy_threshold: float = VAD_ENERGY_THRESHOLD,
                               max_threshold: float = VAD_MAX_THRESHOLD, use_and_logic: bool = True):
        """`push` that also returns what `StreamingASRSession.process_chunk` computes on the host for every chunk
        (R:voice_interface.py:1569-1578): is_speech = mean|x| > 0.03 AND (or OR) max|x| > 0.17, as a by-product of the
        same kernel launch.  Returns (feats, rows, is_speech bool [n], stats float32 [n, 2] = mean|x|, max|x|)."""
        if not chunks.is_cuda:
            raise RuntimeError("chunks must be a CUDA tensor: the B200 front-end has no CPU fallback")
        feats, rows, stats = self._h.ops.stream_push_stats(self._h.h, self.state, self.n_streams, self.max_chunk, chunks,
                                                           chunk_lens, stream_ids, is_final)
        a, b = stats[:, 0] > energy_threshold, stats[:, 1] > max_threshold
        return feats, rows, (a & b) if use_and_logic else (a | b), stats

    def snapshot(self) -> torch.Tensor:
        """Checkpoint of every stream (one memcpy of the slab)."""
        return self.state.clone()

    def restore(self, snap: torch.Tensor) -> None:
        self.state.copy_(snap)


class AudioRing:
    """Sliding audio windows for many sessions, resident in HBM: what the reference keeps per session with
    `buf = np.concatenate([buf, chunk])[-target:]` (KWS window 1.6 s, pre-speech guard 0.4 s;
    R:voice_interface.py:1304-1311, 1742-1746).  `window(ids)` returns what the reference would hand to the model after
    its slice: the newest min(total, capacity) samples per stream, oldest first, as a dense [k, capacity] tensor with a
    zero tail, plus their counts - ready for `WavFrontend.forward(window, lens)`."""

    def __init__(self, n_streams: int, capacity_samples: int, device="cuda"):
        from . import _native
        self.n_streams, self.capacity = int(n_streams), int(capacity_samples)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AudioRing state must live on a CUDA device (no CPU fallback)")
        self._ops = _native.ops()
        self.state = self._ops.ring_state(self.n_streams, self.capacity, self.device)

    def reset(self, stream_ids) -> None:
        self._ops.ring_reset(self.state, self.n_streams, self.capacity, torch.as_tensor(stream_ids, dtype=torch.int32))

    def push(self, chunks: torch.Tensor, chunk_lens, stream_ids) -> None:
        """chunks: CUDA float32 [k, <=max_len]; stream ids must be distinct within one call."""
        if not chunks.is_cuda:
            raise RuntimeError("chunks must be a CUDA tensor: the B200 front-end has no CPU fallback")
        self._ops.ring_push(self.state, self.n_streams, self.capacity, chunks, torch.as_tensor(chunk_lens, dtype=torch.int32),
                            torch.as_tensor(stream_ids, dtype=torch.int32))

    def window(self, stream_ids) -> Tuple[torch.Tensor, torch.Tensor]:
        return self._ops.ring_window(self.state, self.n_streams, self.capacity, torch.as_tensor(stream_ids, dtype=torch.int32))


class WavFrontendOnline(WavFrontend):
    """Chunked front-end with upstream's call shape (batch size 1, state in the caller's `cache` dict).

    forward(input[1, n], input_lengths, cache=dict, is_final=bool) -> (feats[1, k, D], feats_lengths[1]); k may be 0
    (then feats is an empty [1, 0, D] tensor).  Like upstream's `WavFrontendOnline`, every call also leaves the raw
    samples that belong to the rows it returned in `cache["waveforms"]` ([1, (k'-1)*shift + frame] samples, k' = lfr_n*(k-1)+1
    frames starting at the first returned row's centre frame) and keeps the not yet consumed tail in
    `cache["reserve_waveforms"]`: FSMN-VAD reads its per-frame energies from there (the consumer call is
    R:voice-service/app/services/voice_interface.py:1585-1590, R:voice-service/app/api/voice.py:465-470).

    The stream counters are deterministic functions of the chunk lengths, so the host mirrors them (`cache["_host"]`)
    and knows how many rows a push returns without reading anything back: no synchronisation, no per-call allocation
    besides the output."""

    def __init__(self, *args, max_chunk_samples: int = 16000, **kwargs):
        super().__init__(*args, **kwargs)
        self.max_chunk_samples = int(max_chunk_samples)
        self.frame_sample_length = int(self.frame_length * self.fs / 1000)
        self.frame_shift_sample_length = int(self.frame_shift * self.fs / 1000)

    def init_cache(self, cache: dict, device="cuda") -> dict:
        dev = torch.device(device)
        cache["pool"] = StreamPool(self, 1, self.max_chunk_samples, dev)
        cache["ids"] = torch.zeros(1, dtype=torch.int32, device=dev)
        cache["_len"] = torch.zeros(1, dtype=torch.int32, device=dev)
        cache["_fin"] = (torch.zeros(1, dtype=torch.uint8, device=dev), torch.ones(1, dtype=torch.uint8, device=dev))
        cache["_host"] = dict(carry=0, frames=0, rows=0)
        cache["reserve_waveforms"] = torch.zeros(1, 0, dtype=torch.float32, device=dev)
        cache["waveforms"] = torch.zeros(1, 0, dtype=torch.float32, device=dev)
        return cache

    def _rows_after_push(self, st: dict, m: int, final: bool) -> int:
        """Host mirror of the stream counters (csrc/stream_kernel.cuh): rows one push of m samples returns."""
        L, S = self.frame_sample_length, self.frame_shift_sample_length
        total = st["carry"] + m
        nf = (total - L) // S + 1 if total >= L else 0
        st["carry"] = total - nf * S
        st["frames"] += nf
        t = st["frames"]
        rows_all = -(-t // self.lfr_n) if t > 0 else 0
        need = self.lfr_m - 1 - (self.lfr_m - 1) // 2
        rows_total = rows_all if final else ((t - 1 - need) // self.lfr_n + 1 if t - 1 >= need else 0)
        rows_total = max(min(rows_total, rows_all), st["rows"])
        k = rows_total - st["rows"]
        st["rows"] = rows_total
        if final:
            st["carry"] = st["frames"] = st["rows"] = 0
        return k

    def forward(self, input: torch.Tensor, input_lengths, cache: Optional[dict] = None, is_final: bool = False,
                **kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        self._check_cuda(input, "input")
        assert input.shape[0] == 1, "we support to extract feature online only when the batch size is equal to 1 now"
        if cache is None:
            cache = {}
        if "pool" not in cache:
            self.init_cache(cache, input.device)
        pool: StreamPool = cache["pool"]
        st = cache["_host"]
        n = int(_as_length_tensor(input_lengths)[0])
        x = input[:, :n].to(torch.float32)
        L, S = self.frame_sample_length, self.frame_shift_sample_length
        outs, k_total, pos = [], 0, 0
        while True:
            m = min(self.max_chunk_samples, n - pos)
            last = pos + m >= n
            fin = bool(is_final and last)
            cache["_len"].fill_(m)
            k = self._rows_after_push(st, m, fin)
            chunk = x[:, pos:pos + m] if m > 0 else x.new_zeros(1, 1)
            feats, _ = pool.push(chunk, cache["_len"], cache["ids"], cache["_fin"][1 if fin else 0])
            if k:
                outs.append(feats[0, :k])
            k_total += k
            pos += m
            if last:
                break
        # raw samples of the returned rows (upstream: cache["waveforms"]) and of the rows still to come
        # (cache["reserve_waveforms"]): reserve starts at the centre frame of the first row not returned yet
        wav = torch.cat((cache["reserve_waveforms"], x), dim=1)
        if k_total:
            cache["waveforms"] = wav[:, :min(wav.shape[1], self.lfr_n * (k_total - 1) * S + L)]
            wav = wav[:, min(wav.shape[1], self.lfr_n * k_total * S):]
        else:
            cache["waveforms"] = wav[:, :0]
        cache["reserve_waveforms"] = wav[:, :0] if is_final else wav
        if not outs:
            return input.new_zeros((1, 0, self.output_size()), dtype=torch.float32), torch.zeros(1, dtype=torch.int64)
        out = (outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)).unsqueeze(0)
        return out, torch.as_tensor([k_total], dtype=torch.int64)
