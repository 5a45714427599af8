#!/usr/bin/env python
"""Side benchmark (not the driver's contract): BASELINE.json configs[4], TTS reference-audio log-mel, on 1 / 8 GPUs.

    python bench_tts.py [--batch 512] [--steps 50]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_tts.py     # 512 clips per GPU, no collective

512 synthetic clips of 3-10 s at 24 kHz, n_fft 1024 / hop 256 / 80 mels, clips resident in HBM; CUDA events.
Algorithmic bytes per clip: 4 N in + 4*80*(N//256) out = 126 000 B per audio-second."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from toolbox_for_asr_and_tts_b200 import TtsLogMel, _native, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu", action="store_true", help="also time the frozen numpy definition on 16 clips")
    a = ap.parse_args()
    import os
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lens = synth.utterance_lengths(5 + rank, a.batch, lo=72000, hi=240000)
    nmax = (int(lens.max()) + 3) // 4 * 4          # rows of the dense batch start on 16-byte boundaries
    wave = torch.zeros(a.batch, nmax, device=dev)
    offs = torch.arange(a.batch, dtype=torch.int64) * nmax
    _native.ops().synth_uniform(wave, offs, torch.from_numpy(lens), 5 + rank, 0.3)
    fe = TtsLogMel()
    lens_t = torch.from_numpy(lens)
    for _ in range(a.warmup):
        mel, fr = fe(wave, lens_t)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        mel, fr = fe(wave, lens_t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    audio_s = float(lens.sum()) / 24000.0
    alg = int((4 * lens + 4 * 80 * (lens // 256)).sum())
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sm = torch.tensor([audio_s], dtype=torch.float64, device=dev)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, total_audio = float(tt[0]), float(sm[0])
        dist.destroy_process_group()
        if rank != 0:
            return
    else:
        total_audio = audio_s
    out = {"metric": "audio_seconds_per_second", "value": total_audio / (ms * 1e-3), "unit": "audio-s/s", "n_gpus": world,
           "per_gpu_value": total_audio / world / (ms * 1e-3), "parity": "unpinned (no reference code for this config, DESIGN.md)",
           "workload": f"{a.batch} clips per GPU, 3-10 s @24 kHz, n_fft 1024, hop 256, 80 mels", "ms_per_step": ms,
           "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_6544.7": alg / (ms * 1e-3) / 1e9 / 6544.7}
    if a.cpu:
        from oracle import tts_mel_np as tm
        host = wave[:16].cpu().numpy()
        t0 = time.perf_counter()
        for i in range(16):
            tm.tts_log_mel(host[i, :int(lens[i])])
        dt = time.perf_counter() - t0
        out["cpu_numpy_definition_audio_s_per_s"] = float(lens[:16].sum()) / 24000.0 / dt
    print(json.dumps(out))


if __name__ == "__main__":
    main()
