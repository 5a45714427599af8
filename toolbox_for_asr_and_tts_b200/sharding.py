"""One process per GPU: partitioning of utterances / streams across ranks and the single collective of the path.

Utterances and streams are independent, so ranks never exchange audio or features; the only collective is the
all-reduce of the global CMVN statistics (sum, sum of squares, count: 2*D+1 float64 values) at the end of a corpus
pass (upstream funasr/bin/compute_audio_cmvn.py computes the same statistics in one process)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch


def partition_utterances(lengths: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Longest-first greedy assignment balancing the number of SAMPLES per rank (not the number of utterances).
    Deterministic; returns, per rank, the sorted utterance indices it owns."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owned = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        owned[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.array(sorted(o), dtype=np.int64) for o in owned]


def stream_owner(stream_id: int, world_size: int) -> int:
    """Sticky stream placement: a stream's state never migrates."""
    return int(stream_id) % int(world_size)


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of the float64 [2*D+1] statistics over the process group (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
