#!/usr/bin/env python
"""Side benchmark (not the driver's contract): BASELINE.json configs[2], chunked streaming.

    python bench_streaming.py [--streams 512] [--ticks 200] [--chunk 9600]

S concurrent streams on one GPU, one 600 ms chunk per stream and tick, state resident in HBM, one kernel launch per
tick.  Prints one JSON line: audio-seconds per second and ticks per second, device-resident chunks (CUDA events)."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from bench import CONF, synthetic_cmvn  # noqa: E402
from toolbox_for_asr_and_tts_b200 import StreamPool, WavFrontend, _native  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=512)
    ap.add_argument("--ticks", type=int, default=200)
    ap.add_argument("--chunk", type=int, default=9600)
    ap.add_argument("--warmup", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
    pool = StreamPool(fe, a.streams, a.chunk, dev)
    n_buf = 8                                   # rotate through 8 different chunk sets
    wave = torch.zeros(n_buf * a.streams * a.chunk, device=dev)
    offs = torch.arange(n_buf * a.streams, dtype=torch.int64) * a.chunk
    _native.ops().synth_uniform(wave, offs, torch.full((n_buf * a.streams,), a.chunk, dtype=torch.int64), 3, 0.3)
    wave = wave.view(n_buf, a.streams, a.chunk)
    ids = torch.arange(a.streams, dtype=torch.int32, device=dev)
    lens = torch.full((a.streams,), a.chunk, dtype=torch.int32, device=dev)
    for t in range(a.warmup):
        pool.push(wave[t % n_buf], lens, ids, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(a.ticks):
        feats, rows = pool.push(wave[t % n_buf], lens, ids, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    audio_s = a.streams * a.chunk / 16000.0 * a.ticks
    rows_per_tick = int(rows[0])
    alg = a.streams * (4 * a.chunk + 2240 * rows_per_tick)
    print(json.dumps({"metric": "audio_seconds_per_second", "value": audio_s / (ms * 1e-3), "unit": "audio-s/s",
                      "workload": f"{a.streams} streams x {a.chunk}-sample chunks, {a.ticks} ticks, state in HBM",
                      "ms_per_tick": ms / a.ticks, "ticks_per_second": a.ticks / (ms * 1e-3),
                      "rows_per_stream_tick": rows_per_tick,
                      "algorithmic_GBps": alg * a.ticks / (ms * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
