#!/usr/bin/env python
"""BASELINE.json configs[3]: offline bulk extraction of 1 000 h of synthetic audio, sharded by utterance over 8 GPUs,
with the global CMVN statistics all-reduce.

    python bench_bulk.py --hours 125                                                         # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_bulk.py --hours 1000

The corpus is `n = hours*3600 / 15.5 s` utterances with seeded lengths (1-30 s); utterance g is
synth.uniform_pcm(seed, g, len_g).  Sharding: sharding.partition_utterances - longest first to the least loaded rank
(SURVEY.md 8(e)); every rank synthesises ITS utterances on the device (by id, bit-identical to the single-process
corpus) slab by slab (256 utterances), runs the fused front-end with the float64 statistics accumulation
(sum, sum of squares, row count of the un-normalised LFR features) and ONE NCCL all-reduce of the 2*560+1 float64 values
at the end yields the global am.mvn table.  Timed with CUDA events, barrier on both sides, max over ranks; the all-reduce
is also timed on its own.  The row count is checked exactly, and the statistics of a 1 h subsample against the float64
numpy oracle (--check-hours).  Rank 0 prints one JSON line."""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from bench import CONF  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native, sharding, stats_to_cmvn, synth  # noqa: E402

SEED = 4242


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hours", type=float, default=125.0)
    ap.add_argument("--slab", type=int, default=256)
    ap.add_argument("--check-hours", type=float, default=1.0, help="subsample checked against the float64 oracle (rank 0)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ops = _native.ops()
    fe = WavFrontend(cmvn=None, dither=0.0, **CONF)
    n_utts = max(world, int(round(a.hours * 3600.0 / 15.5)))
    lens_all = synth.utterance_lengths(SEED, n_utts)                 # the corpus: every rank derives the same lengths
    t0 = time.perf_counter()
    mine = sharding.partition_utterances(lens_all, world)[rank]      # longest-first, balances samples
    t_part = time.perf_counter() - t0
    stats = torch.zeros(2 * 560 + 1, dtype=torch.float64, device=dev)
    wave = torch.empty(a.slab * 480000 + 64, dtype=torch.float32, device=dev)
    meta, rows_expected, audio_s = [], 0, 0.0
    for s in range(0, len(mine), a.slab):                            # slab metadata is host work outside the timed region
        ids = mine[s:s + a.slab]
        lens = lens_all[ids]
        offs, total = synth.packed_offsets(lens)
        meta.append((torch.from_numpy(ids).to(dev), torch.from_numpy(offs), torch.from_numpy(lens), total))
        t = 1 + (lens - 400) // 160
        rows_expected += int((-(-t // 6)).sum())
        audio_s += float(lens.sum()) / 16000.0
    # warm-up outside the timed region: communicator (same tensor size as the real all-reduce), kernels, allocator
    sharding.allreduce_stats(torch.zeros_like(stats))
    sharding.allreduce_stats(torch.zeros_like(stats))
    ids0, offs0, lens0, total0 = meta[0]
    ops.synth_uniform_ids(wave, offs0, lens0, ids0, SEED, 0.3)
    fe.forward_packed(wave[: total0 + 8], offs0, lens0, stats=torch.zeros_like(stats), rows_cap=500)
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for ids, offs_t, lens_t, total in meta:
        ops.synth_uniform_ids(wave, offs_t, lens_t, ids, SEED, 0.3)
        fe.forward_packed(wave[: total + 8], offs_t, lens_t, stats=stats, rows_cap=500)   # one output shape: allocator reuse
    e1.record()
    sharding.allreduce_stats(stats)
    e2.record()
    barrier()
    ms, ms_ar = e0.elapsed_time(e2), e1.elapsed_time(e2)
    tot = torch.tensor([ms, ms_ar, audio_s, float(rows_expected)], dtype=torch.float64, device=dev)
    if world > 1:
        mx, mn = tot.clone(), tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    else:
        mx, mn = tot.clone(), tot.clone()
    if rank == 0:
        assert float(stats[-1]) == float(tot[3]), (float(stats[-1]), float(tot[3]))   # row count exact
        table = stats_to_cmvn(stats)
        check = None
        if a.check_hours > 0:
            # statistics of a subsample against the float64 oracle: the first k utterances of the corpus
            from oracle import wav_frontend_np as wf
            k = max(1, min(n_utts, int(round(a.check_hours * 3600.0 / 15.5))))
            ids = np.arange(k, dtype=np.int64)
            lens = lens_all[ids]
            offs, total = synth.packed_offsets(lens)
            st = torch.zeros_like(stats)
            w = torch.empty(total + 8, dtype=torch.float32, device=dev)
            ops.synth_uniform_ids(w, torch.from_numpy(offs), torch.from_numpy(lens), torch.from_numpy(ids).to(dev), SEED, 0.3)
            fe.forward_packed(w, torch.from_numpy(offs), torch.from_numpy(lens), stats=st)
            mats = []
            for g, n in zip(ids, lens):
                f, _ = wf.frontend_forward([synth.uniform_pcm(SEED, int(g), int(n))], [int(n)], cmvn=None, dtype=np.float64, **CONF)
                mats.append(f[0])
            s1, s2, cnt = wf.cmvn_stats(mats)
            got = st.cpu().numpy()
            assert int(got[-1]) == cnt
            mean_err = float(np.abs(got[:560] / cnt - s1 / cnt).max())
            var_g, var_o = got[560:1120] / cnt - (got[:560] / cnt) ** 2, s2 / cnt - (s1 / cnt) ** 2
            check = {"utterances": int(k), "hours": float(lens.sum() / 16000.0 / 3600.0), "rows": int(cnt),
                     "max_abs_err_mean": mean_err, "max_rel_err_variance": float(np.abs(var_g / var_o - 1).max())}
            assert mean_err < 1e-4 and check["max_rel_err_variance"] < 1e-4, check
        print(json.dumps({"metric": "audio_seconds_per_second", "value": float(tot[2]) / (float(mx[0]) * 1e-3), "unit": "audio-s/s",
                          "n_gpus": world, "scaling": "weak" if world == 1 else "strong-per-corpus", "higher_is_better": True,
                          "per_gpu_value": float(tot[2]) / world / (float(mx[0]) * 1e-3),
                          "workload": "offline bulk extraction + global CMVN statistics (BASELINE.json configs[3])",
                          "hours": float(tot[2]) / 3600.0, "utterances": int(n_utts), "seconds": float(mx[0]) * 1e-3,
                          "rows": int(stats[-1]), "sharding": "partition_utterances (longest first to the least loaded rank)",
                          "partition_host_seconds": t_part,
                          "rank_audio_seconds_min_max": [float(mn[2]), float(mx[2])],
                          "rank_ms_min_max": [float(mn[0]), float(mx[0])],
                          "allreduce_ms_max": float(mx[1]), "allreduce_elements": 1121, "allreduce_dtype": "float64",
                          "includes": "on-device synthesis of the audio + fused front-end with float64 statistics + 1 all-reduce",
                          "cmvn_mean_range": [float(-table[0].max()), float(-table[0].min())],
                          "cmvn_std_range": [float((1 / table[1]).min()), float((1 / table[1]).max())],
                          "oracle_check": check}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
