"""B200-native speech front-end (Kaldi fbank -> LFR -> CMVN) behind the reference's WavFrontend interface.

Public surface (mirrors upstream funasr.frontends.wav_frontend):
    WavFrontend, WavFrontendOnline, load_cmvn
plus: StreamPool (many concurrent streams per tick), cmvn statistics helpers, sharding helpers, synthetic PCM.
The native code (libb200fe.so + _b200fe_torch.so) is loaded lazily; there is no CPU fallback.
"""
from .cmvn import load_cmvn, stats_to_cmvn, write_cmvn  # noqa: F401
from .frontend import WavFrontend  # noqa: F401
from .online import AudioRing, StreamPool, WavFrontendOnline  # noqa: F401
from .tts_mel import TtsLogMel  # noqa: F401
from . import presets  # noqa: F401
from .ingest import HostIngest  # noqa: F401

__all__ = ["WavFrontend", "WavFrontendOnline", "StreamPool", "AudioRing", "HostIngest", "TtsLogMel", "presets", "load_cmvn",
           "write_cmvn", "stats_to_cmvn"]
