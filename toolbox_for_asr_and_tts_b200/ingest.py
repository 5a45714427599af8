"""Host-side entry of the front-end: from the utterances a reference caller holds in HOST memory to features in HBM.

The reference hands funasr a list of np.float32 arrays (R:voice-service/app/services/voice_interface.py:2049-2053,
1370-1374; R:voice-service/full_voice_demo.py:327); funasr pads them into one [B, Nmax] CPU tensor, runs the CPU
front-end and only then moves the features to the GPU.  `HostIngest` replaces the host part of that: the utterances are
gathered by several host threads into a length-packed pinned staging buffer, group by group, each group crossing PCIe
while the next one is gathered (libb200fe: b200fe_host_ingest), and the fused kernels run on the packed device buffer.
Staging and device buffers are double-buffered, so call k+1 gathers while call k's copy and kernels are still running.
int16 PCM (the wire format, value = s / 32768) takes the same route at half the bytes.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native
from .frontend import WavFrontend


class HostIngest:
    def __init__(self, frontend: WavFrontend, capacity_samples: int, dtype: torch.dtype = torch.float32, device="cuda",
                 depth: int = 2, groups: int = 8, threads: int = 0):
        if dtype not in (torch.float32, torch.int16):
            raise ValueError("HostIngest takes float32 or int16 PCM")
        self.frontend, self.dtype = frontend, dtype
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostIngest feeds a CUDA device (no CPU fallback)")
        self.capacity, self.depth, self.groups, self.threads = int(capacity_samples), int(depth), int(groups), int(threads)
        self.align = 4 if dtype == torch.float32 else 8           # 16-byte aligned utterance starts
        self._ops = _native.ops()
        self._stage = [torch.empty(self.capacity, dtype=dtype).pin_memory() for _ in range(self.depth)]
        self._dev = [torch.empty(self.capacity, dtype=dtype, device=self.device) for _ in range(self.depth)]
        self._copied = [torch.cuda.Event() for _ in range(self.depth)]
        self._done = [torch.cuda.Event() for _ in range(self.depth)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._k = 0
        self.last_h2d_bytes = 0

    def _as_tensors(self, waves) -> Sequence[torch.Tensor]:
        if isinstance(waves, torch.Tensor):            # dense [B, Nmax] host tensor: its rows
            return list(waves.unbind(0))
        return [w if isinstance(w, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(w)) for w in waves]

    def forward(self, waves: Union[torch.Tensor, Sequence], lengths: Optional[Sequence[int]] = None,
                rows_cap: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        """waves: list of 1-D host arrays / tensors, or a dense [B, Nmax] host tensor with `lengths`.  Returns
        (feats [B, max rows, D] on the device, feature lengths int64 on the device) without synchronising."""
        ws = self._as_tensors(waves)
        lens = np.asarray([int(w.numel()) for w in ws] if lengths is None else [int(n) for n in lengths], dtype=np.int64)
        padded = (lens + self.align - 1) // self.align * self.align
        offs = np.concatenate(([0], np.cumsum(padded)[:-1])).astype(np.int64) if len(lens) else np.zeros(0, np.int64)
        total = int(padded.sum())
        if total > self.capacity:
            raise ValueError(f"batch of {total} samples exceeds the ingest capacity {self.capacity}")
        b = self._k % self.depth
        self._k += 1
        lens_t, offs_t = torch.from_numpy(lens), torch.from_numpy(offs)
        self._copied[b].synchronize()                   # the staging buffer's previous copy has left the host
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._done[b])  # the device buffer's previous kernels are done
            self._ops.host_ingest(ws, lens_t, offs_t, self._stage[b], self._dev[b], self.groups, self.threads)
            self._copied[b].record(self._copy_stream)
        cur.wait_event(self._copied[b])
        feats, flens = self.frontend.forward_packed(self._dev[b][:max(total, 1)], offs_t, lens_t, rows_cap=rows_cap)
        self._done[b].record(cur)
        self.last_h2d_bytes = int(lens.sum()) * self._dev[b].element_size()
        return feats, flens
