"""Minimal program for ncu: BASELINE configs[1] (the batch bench.py times), device-resident, N steps of
forward_packed and nothing else.

    python tools/profile_step.py [steps]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bench import CONF, batch_layout, synthetic_cmvn  # noqa: E402
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
lens, offs, total = batch_layout(0)
wave = torch.zeros(total + 8, device=dev)
_native.ops().synth_uniform(wave, torch.from_numpy(offs), torch.from_numpy(lens), 1234, 0.3)
fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
lt, ot = torch.from_numpy(lens), torch.from_numpy(offs)
for _ in range(steps):
    feats, fl = fe.forward_packed(wave, ot, lt)
torch.cuda.synchronize()
print("ok", tuple(feats.shape), int(fl.sum()))
