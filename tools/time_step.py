"""Quick timing of the fused step on BASELINE configs[1] (device-resident): step ms and fused-kernel ms (CUDA events on
the kernel's stream through b200fe_profile_enable), plus a bit-exact check against the tile kernel's row counts.

    python tools/time_step.py [steps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import CONF, algorithmic_bytes, batch_layout, synthetic_cmvn  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
lens, offs, total = batch_layout(0)
wave = torch.zeros(total + 8, device=dev)
_native.ops().synth_uniform(wave, torch.from_numpy(offs), torch.from_numpy(lens), 1234, 0.3)
fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
lt, ot = torch.from_numpy(lens), torch.from_numpy(offs)
for _ in range(5):
    feats, fl = fe.forward_packed(wave, ot, lt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    feats, fl = fe.forward_packed(wave, ot, lt)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / steps
fe.profile(1)
for _ in range(20):
    fe.forward_packed(wave, ot, lt)
torch.cuda.synchronize()
ms, n = fe.profile_collect()
fe.profile(0)
alg = algorithmic_bytes(lens)
print(f"step {step_ms:.4f} ms | fused kernel {ms / n:.4f} ms ({n} launches) | {alg / (ms / n * 1e-3) / 1e9:.0f} GB/s algorithmic "
      f"= {alg / (ms / n * 1e-3) / 1e9 / 6544.7:.3f} of 6544.7 | checksum {float(feats.double().sum()):.6e}")
