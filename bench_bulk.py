#!/usr/bin/env python
"""Side benchmark (not the driver's contract): BASELINE.json configs[3], offline bulk extraction with global CMVN
statistics.

    python bench_bulk.py --hours 10                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_bulk.py --hours 1000

Synthetic utterances (1-30 s) are generated on the device slab by slab (256 utterances per slab), sharded by slab over
the ranks; every rank accumulates sum / sum of squares / row count of the un-normalised LFR features in float64 inside
the fused kernel, and ONE all-reduce (2*560+1 float64 values, NCCL) at the end yields the global am.mvn table.
Rank 0 prints one JSON line; the row count is checked exactly against the host-side frame arithmetic."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from bench import CONF  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native, sharding, stats_to_cmvn, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=10.0)
    ap.add_argument("--slab", type=int, default=256)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops = _native.ops()
    fe = WavFrontend(cmvn=None, dither=0.0, **CONF)
    mean_s = (16000 + 480000) / 2 / 16000.0
    n_slabs = max(world, int(round(a.hours * 3600.0 / (a.slab * mean_s))))
    my_slabs = list(range(rank, n_slabs, world))
    stats = torch.zeros(2 * 560 + 1, dtype=torch.float64, device=dev)
    cap = a.slab * 480000 + 64
    wave = torch.empty(cap, dtype=torch.float32, device=dev)
    rows_expected, audio_s = 0, 0.0
    meta = []
    for s in my_slabs:                     # slab metadata is host work outside the timed region
        lens = synth.utterance_lengths(1000 + s, a.slab)
        offs, total = synth.packed_offsets(lens)
        meta.append((s, lens, offs, total, torch.from_numpy(offs), torch.from_numpy(lens)))
    sharding.allreduce_stats(torch.zeros(4, dtype=torch.float64, device=dev))   # communicator warm-up
    fe.forward_packed(torch.zeros(16008, device=dev), [0], [16000], stats=torch.zeros_like(stats))   # handle / kernel warm-up
    if meta:                                   # allocator warm-up: one slab-sized pass outside the timed region
        s0, lens0, offs0, total0, offs0_t, lens0_t = meta[0]
        ops.synth_uniform(wave, offs0_t, lens0_t, 1000 + s0, 0.3)
        fe.forward_packed(wave[: total0 + 8], offs0_t, lens0_t, stats=torch.zeros_like(stats), rows_cap=500)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s, lens, offs, total, offs_t, lens_t in meta:
        ops.synth_uniform(wave, offs_t, lens_t, 1000 + s, 0.3)
        fe.forward_packed(wave[: total + 8], offs_t, lens_t, stats=stats, rows_cap=500)   # one output shape: allocator reuse
        t = 1 + (lens - 400) // 160
        rows_expected += int((-(-t // 6)).sum())
        audio_s += float(lens.sum()) / 16000.0
    sharding.allreduce_stats(stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tot = torch.tensor([ms, audio_s, float(rows_expected)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        ms = float(mx[0])
    if rank == 0:
        assert float(stats[-1]) == float(tot[2]), (float(stats[-1]), float(tot[2]))   # row count exact
        table = stats_to_cmvn(stats)
        print(json.dumps({"metric": "audio_seconds_per_second", "value": float(tot[1]) / (ms * 1e-3), "unit": "audio-s/s",
                          "n_gpus": world, "hours": float(tot[1]) / 3600.0, "slabs": n_slabs, "seconds": ms * 1e-3,
                          "rows": int(stats[-1]), "includes": "on-device synthesis of the audio + fused front-end + stats + 1 all-reduce",
                          "cmvn_mean_range": [float(-table[0].max()), float(-table[0].min())],
                          "cmvn_std_range": [float((1 / table[1]).min()), float((1 / table[1]).max())]}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
