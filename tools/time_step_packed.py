"""tools/time_step.py for the rows-packed output (forward_packed(pad=False)): step and fused-kernel time."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import CONF, batch_layout, synthetic_cmvn  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native  # noqa: E402

dev = torch.device("cuda", 0)
lens, offs, total = batch_layout(0)
wave = torch.zeros(total + 8, device=dev)
_native.ops().synth_uniform(wave, torch.from_numpy(offs), torch.from_numpy(lens), 1234, 0.3)
fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
lt, ot = torch.from_numpy(lens), torch.from_numpy(offs)
for pad in (True, False):
    for _ in range(5):
        out = fe.forward_packed(wave, ot, lt, pad=pad)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        out = fe.forward_packed(wave, ot, lt, pad=pad)
    e1.record()
    host_ms = (time.perf_counter() - t0) / 200 * 1e3
    torch.cuda.synchronize()
    fe.profile(1)
    for _ in range(20):
        fe.forward_packed(wave, ot, lt, pad=pad)
    torch.cuda.synchronize()
    ms, n = fe.profile_collect()
    fe.profile(0)
    print(f"pad={pad}: step {e0.elapsed_time(e1) / 200:.4f} ms (host enqueue {host_ms:.4f} ms/step) | fused kernel {ms / n:.4f} ms | "
          f"checksum {float(out[0].double().sum()):.6e}")
