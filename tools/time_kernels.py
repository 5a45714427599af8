import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from bench import batch_layout, synthetic_cmvn, CONF
from toolbox_for_asr_and_tts_b200 import WavFrontend, _native
dev = torch.device('cuda', 0)
lens, offs, total = batch_layout(0)
wave = torch.zeros(total + 8, device=dev)
_native.ops().synth_uniform(wave, torch.from_numpy(offs), torch.from_numpy(lens), 1234, 0.3)
"""Times the fused kernels on BASELINE configs[1] (device-resident batch): warp vs tile kernel, with / without the
statistics pass, and the dither = 1.0 side run (the reference's default; Paraformer inference sets 0)."""
for which in ('auto', 'tile'):
    fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=0.0, **CONF)
    fe.select_kernel(which)
    lt, ot = torch.from_numpy(lens), torch.from_numpy(offs)
    for st in (None, torch.zeros(1121, dtype=torch.float64, device=dev)):
        for _ in range(5): fe.forward_packed(wave, ot, lt, stats=st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fe.forward_packed(wave, ot, lt, stats=st)
        e1.record(); torch.cuda.synchronize()
        print(which, 'stats' if st is not None else 'nostats', e0.elapsed_time(e1) / 50, 'ms/step')

fe = WavFrontend(cmvn=torch.from_numpy(synthetic_cmvn()), dither=1.0, **CONF)
lt, ot = torch.from_numpy(lens), torch.from_numpy(offs)
for _ in range(3): fe.forward_packed(wave, ot, lt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fe.forward_packed(wave, ot, lt)
e1.record(); torch.cuda.synchronize()
print('auto dither=1.0', e0.elapsed_time(e1) / 10, 'ms/step')
