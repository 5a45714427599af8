"""Latency of ONE utterance through the public API (BASELINE configs[0]: 10 s at 16 kHz, Paraformer front-end):
host float32 PCM -> H2D -> WavFrontend.forward -> feature lengths back on the host, synchronous, per call.

    python tools/latency_single.py [--seconds 10] [--calls 300]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import CONF, synthetic_cmvn  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--calls", type=int, default=300)
    a = ap.parse_args()
    n = int(a.seconds * 16000)
    dev = torch.device("cuda", 0)
    fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
    host = torch.from_numpy(synth.uniform_pcm(1234, 0, n)).pin_memory()[None]
    res = {}
    for name, resident in (("host_to_host_lengths", False), ("device_resident", True)):
        xd = host.to(dev)
        ts = []
        for k in range(a.calls + 20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x = xd if resident else host.to(dev, non_blocking=True)
            feats, lens = fe(x, [n])
            _ = lens.cpu()      # synchronises: the caller reads the row count
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts[20:]) * 1e6
        res[name] = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)),
                     "min_us": float(ts.min())}
    # GPU time alone (events), device-resident
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.calls):
        fe(xd, [n])
    e1.record()
    torch.cuda.synchronize()
    res["back_to_back_us_per_call"] = e0.elapsed_time(e1) * 1e3 / a.calls
    print(json.dumps({"workload": f"1 utterance x {a.seconds:g} s @16 kHz", "rows": int(feats.shape[1]), **res}))


if __name__ == "__main__":
    main()
