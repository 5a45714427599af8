"""Loading of the native code.  There is NO CPU / eager fallback: if the libraries are missing or do not load, every
front-end call raises (build them with `python -m toolbox_for_asr_and_tts_b200._build`)."""
from __future__ import annotations

import ctypes
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libb200fe.so"
EXT_PATH = PKG / "_b200fe_torch.so"

_ops = None
_cdll = None


class NativeLibraryError(RuntimeError):
    pass


def cdll() -> ctypes.CDLL:
    """The raw C ABI (include/b200fe.h) through ctypes - used for symbol checks and C-ABI-level tests."""
    global _cdll
    if _cdll is None:
        if not LIB_PATH.exists():
            raise NativeLibraryError(f"{LIB_PATH} is missing: run `python -m toolbox_for_asr_and_tts_b200._build` "
                                     "(the B200 front-end has no CPU fallback)")
        _cdll = ctypes.CDLL(str(LIB_PATH), mode=ctypes.RTLD_GLOBAL)
    return _cdll


def ops():
    """`torch.ops.b200fe` after loading the torch extension (which links libb200fe.so)."""
    global _ops
    if _ops is None:
        import torch
        if not EXT_PATH.exists() or not LIB_PATH.exists():
            raise NativeLibraryError(f"{EXT_PATH.name} / {LIB_PATH.name} are missing: run "
                                     "`python -m toolbox_for_asr_and_tts_b200._build` (no CPU fallback exists)")
        cdll()
        torch.ops.load_library(str(EXT_PATH))
        _ops = torch.ops.b200fe
    return _ops


def declared_symbols() -> list[str]:
    """Entry points declared in include/b200fe.h (parsed from the header so the test cannot drift)."""
    import re
    text = (PKG.parent / "include" / "b200fe.h").read_text()
    return sorted(set(re.findall(r"\b(b200fe_[a-z_0-9]+)\s*\(", text)))
