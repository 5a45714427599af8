"""In-tree build of the native code: libb200fe.so (CUDA kernels + C ABI) and _b200fe_torch.so (torch extension).

`python -m toolbox_for_asr_and_tts_b200._build` or `__graft_entry__.build()`.  nvcc cross-compiles sm_100a without a
GPU; the .so files are git-ignored but travel with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libb200fe.so"
EXT = PKG / "_b200fe_torch.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, log_name):
    log = PKG.parent / "build" / log_name
    log.parent.mkdir(exist_ok=True)
    res = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    log.write_text(" ".join(str(c) for c in cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout[-4000:] + res.stderr[-8000:])
        raise RuntimeError(f"build step failed: {cmd[0]} (see {log})")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def build_lib(force: bool = False, verbose_ptxas: bool = True) -> Path:
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inl")) + \
        [PKG.parent / "include" / "b200fe.h"]
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = [nvcc_path(), "-std=c++17", "-O3", *ARCH, "-lineinfo", "--shared", "-Xcompiler", "-fPIC",
           "-o", LIB, CSRC / "b200fe.cu"]
    if verbose_ptxas:
        cmd[3:3] = ["-Xptxas", "-v"]
    _run(cmd, "nvcc_libb200fe.log")
    return LIB


def build_torch_ext(force: bool = False) -> Path:
    import torch
    from torch.utils import cpp_extension as ce
    src = CSRC / "torch_binding.cpp"
    if not force and _newer(EXT, [src, PKG.parent / "include" / "b200fe.h", LIB]):
        return EXT
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-I", p]
    cuda_home = Path(nvcc_path()).resolve().parent.parent
    inc += ["-I", cuda_home / "include"]
    torch_lib = Path(torch.__file__).resolve().parent / "lib"
    abi = int(getattr(torch._C, "_GLIBCXX_USE_CXX11_ABI", True))
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
           "-DTORCH_EXTENSION_NAME=_b200fe_torch", *inc, src, "-o", EXT,
           "-L", torch_lib, "-L", PKG, "-L", cuda_home / "lib64",
           "-lb200fe", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-lcudart",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}", f"-Wl,-rpath,{cuda_home / 'lib64'}"]
    _run(cmd, "gxx_torch_ext.log")
    return EXT


def build_all(force: bool = False):
    return build_lib(force), build_torch_ext(force)


if __name__ == "__main__":
    print(*build_all("--force" in sys.argv), sep="\n")
