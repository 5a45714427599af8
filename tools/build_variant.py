"""Experiment builds: variants/<name>.so = libb200fe.so with only the kernels bench.py launches instantiated
(-DB200FE_BENCH_ONLY, ~4x faster to compile) plus any extra -D flags, for same-box A/B timing:

    python tools/build_variant.py NAME [-DFOO=1 ...]
    # on the GPU box:  for v in variants/*.so; do cp $v toolbox_for_asr_and_tts_b200/libb200fe.so; python tools/time_step.py; done
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
name, extra = sys.argv[1], sys.argv[2:]
out = ROOT / "variants" / f"{name}.so"
out.parent.mkdir(exist_ok=True)
cmd = ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--shared", "-Xcompiler", "-fPIC",
       "-Xptxas", "-v", "-DB200FE_BENCH_ONLY", *extra, "-o", str(out), str(ROOT / "toolbox_for_asr_and_tts_b200" / "csrc" / "b200fe.cu")]
res = subprocess.run(cmd, capture_output=True, text=True)
(out.parent / f"{name}.log").write_text(res.stdout + res.stderr)
if res.returncode:
    sys.stderr.write(res.stderr[-4000:])
    sys.exit(1)
lines = (res.stdout + res.stderr).splitlines()
for i, ln in enumerate(lines):
    if "fbank_warp_kernel" in ln and "Compiling" in ln:
        print(lines[i + 1].strip(), "|", lines[i + 2].strip())
print("built", out)
