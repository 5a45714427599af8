"""CPU suite, part 2: host logic and the C-ABI library surface (no compute calls - there is no GPU here)."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from toolbox_for_asr_and_tts_b200 import _native, sharding, synth

ROOT = Path(__file__).resolve().parents[1]


def test_library_loads_and_exports_every_declared_symbol():
    lib = _native.cdll()
    syms = _native.declared_symbols()
    assert len(syms) >= 18
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_default_config_matches_reference_defaults():
    """VF:92-107 defaults, through the C struct."""
    lib = _native.cdll()

    class Cfg(ctypes.Structure):
        _fields_ = [("struct_size", ctypes.c_int32), ("sample_rate", ctypes.c_int32), ("frame_length_ms", ctypes.c_float),
                    ("frame_shift_ms", ctypes.c_float), ("n_mels", ctypes.c_int32), ("window_type", ctypes.c_int32),
                    ("lfr_m", ctypes.c_int32), ("lfr_n", ctypes.c_int32), ("dither", ctypes.c_float),
                    ("snip_edges", ctypes.c_int32), ("upscale_samples", ctypes.c_int32), ("preemphasis", ctypes.c_float),
                    ("remove_dc_offset", ctypes.c_int32), ("low_freq", ctypes.c_float), ("high_freq", ctypes.c_float),
                    ("blackman_coeff", ctypes.c_float), ("log_floor", ctypes.c_float), ("reserved", ctypes.c_int32 * 7)]
    c = Cfg()
    lib.b200fe_default_config(ctypes.byref(c))
    assert c.struct_size == ctypes.sizeof(Cfg)
    assert (c.sample_rate, c.n_mels, c.window_type, c.lfr_m, c.lfr_n) == (16000, 80, 0, 1, 1)
    assert (c.frame_length_ms, c.frame_shift_ms, c.dither) == (25.0, 10.0, 1.0)
    assert c.snip_edges == 1 and c.upscale_samples == 1 and c.remove_dc_offset == 1
    assert abs(c.preemphasis - 0.97) < 1e-7 and c.low_freq == 20.0 and c.high_freq == 0.0
    assert c.log_floor == np.finfo(np.float32).eps


def test_create_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from toolbox_for_asr_and_tts_b200 import WavFrontend
    fe = WavFrontend(lfr_m=7, lfr_n=6, dither=0.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fe(torch.zeros(1, 16000), [16000])


def test_invalid_window_raises_like_reference():
    from toolbox_for_asr_and_tts_b200 import WavFrontend
    with pytest.raises(Exception, match="Invalid window type"):
        WavFrontend(window="kaiser")


def test_synth_is_deterministic_and_bounded():
    a = synth.uniform_pcm(1234, 7, 5000)
    b = synth.uniform_pcm(1234, 7, 5000)
    assert np.array_equal(a, b) and a.dtype == np.float32
    assert np.abs(a).max() <= 0.3 and abs(float(a.mean())) < 0.02 and 0.16 < float(a.std()) < 0.18
    assert not np.array_equal(a, synth.uniform_pcm(1234, 8, 5000))
    assert np.array_equal(a[:100], synth.uniform_pcm(1234, 7, 100))   # counter based: prefix-stable
    lens = synth.utterance_lengths(0, 256)
    assert lens.min() >= 16000 and lens.max() <= 480000
    offs, total = synth.packed_offsets(lens)
    assert (offs % 4 == 0).all() and total >= lens.sum() and (np.diff(offs) >= lens[:-1]).all()


def test_partition_is_balanced_and_complete():
    lens = synth.utterance_lengths(3, 257)
    parts = sharding.partition_utterances(lens, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(257))
    loads = np.array([lens[p].sum() for p in parts])
    assert loads.max() - loads.min() <= lens.max()
    assert sharding.stream_owner(4095, 8) == 7 and sharding.stream_owner(8, 8) == 0


def test_stats_to_cmvn():
    from toolbox_for_asr_and_tts_b200 import stats_to_cmvn
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 6)) * 2 + 3
    stats = torch.tensor(np.concatenate([x.sum(0), (x * x).sum(0), [1000.0]]), dtype=torch.float64)
    tab = stats_to_cmvn(stats).numpy()
    assert np.allclose(tab[0], -x.mean(0), atol=1e-5) and np.allclose(tab[1], 1 / x.std(0), atol=1e-5)


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO_ROOT"])
from toolbox_for_asr_and_tts_b200 import sharding, synth
from oracle import wav_frontend_np as wf
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
lens = synth.utterance_lengths(11, 12, lo=2000, hi=9000)
mine = sharding.partition_utterances(lens, world)[rank]
conf = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6)
mats = []
for u in mine:
    f, l = wf.frontend_forward([synth.uniform_pcm(11, int(u), int(lens[u]))], [int(lens[u])], cmvn=None, **conf)
    mats.append(f[0])
s, s2, n = wf.cmvn_stats(mats)
stats = torch.tensor(np.concatenate([s, s2, [float(n)]]), dtype=torch.float64)
sharding.allreduce_stats(stats)
if rank == 0:
    allm = []
    for u in range(len(lens)):
        f, l = wf.frontend_forward([synth.uniform_pcm(11, u, int(lens[u]))], [int(lens[u])], cmvn=None, **conf)
        allm.append(f[0])
    S, S2, N = wf.cmvn_stats(allm)
    ref = np.concatenate([S, S2, [float(N)]])
    got = stats.numpy()
    assert got[-1] == ref[-1], (got[-1], ref[-1])
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-12
    print("GLOO_OK", int(got[-1]))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding_and_stats_allreduce(tmp_path):
    """world_size 2 on CPU: utterances are sharded without any data-path collective; the one collective (CMVN statistics)
    reproduces the single-process float64 sums."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", WORLD_SIZE="2",
               OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0]


@pytest.mark.parametrize("m,n", [(7, 6), (5, 1), (1, 1), (3, 2), (8, 3), (4, 6), (2, 5)])
def test_lfr_target_planner_reproduces_apply_lfr(m, n):
    """The quad list of the fused kernel routes every fbank frame to its (row, slot) pairs with b200fe_lfr_targets
    (host code of the C ABI).  Scattering frames with those targets, plus the generic clamp rule for the frames it
    flags, must rebuild apply_lfr (VF:40-60) exactly, for every utterance length."""
    from oracle import wav_frontend_np as wf
    lib = _native.cdll()
    lib.b200fe_lfr_targets.restype = ctypes.c_int
    M = 4
    D = m * M
    left = (m - 1) // 2
    out2 = (ctypes.c_uint32 * 2)()
    for T in list(range(1, 40)) + [97, 98, 99, 100, 101, 102, 103, 600]:
        feats = np.arange(T * M, dtype=np.float32).reshape(T, M) + 1.0
        ref = wf.apply_lfr(feats, m, n)
        rows = ref.shape[0]
        got = np.zeros_like(ref)
        hits = np.zeros(ref.shape, dtype=np.int32)
        for f in range(T):
            rc = lib.b200fe_lfr_targets(f, T, rows, m, n, M, out2)
            assert rc in (0, 1)
            if rc == 1:      # generic path of the kernel: every (row, slot) whose clamped source frame is f
                assert out2[0] == 0xFFFFFFFF and out2[1] == 0xFFFFFFFF
                for i in range(rows):
                    for jj in range(m):
                        if min(max(n * i + jj - left, 0), T - 1) == f:
                            got[i, jj * M:(jj + 1) * M] = feats[f]
                            hits[i, jj * M:(jj + 1) * M] += 1
            else:
                for code in out2:
                    if code == 0xFFFFFFFF:
                        continue
                    jj, off = code >> 27, code & ((1 << 27) - 1)
                    assert off % M == 0 and (off % D) == jj * M
                    got.reshape(-1)[off:off + M] = feats[f]
                    hits.reshape(-1)[off:off + M] += 1
        assert np.array_equal(got, ref), (m, n, T)
        assert (hits == 1).all(), (m, n, T)       # every output element is written exactly once


def test_online_frontend_host_mirror_counts_rows_like_the_oracle():
    """WavFrontendOnline knows how many rows a push returns without reading the device counters: its host mirror of
    the stream counters must agree with the online oracle for every chunking (incl. chunks shorter than one frame and
    a tiny final chunk), for the Paraformer (7/6) and the FSMN-VAD (5/1) LFR settings."""
    from oracle import wav_frontend_np as wf
    from toolbox_for_asr_and_tts_b200 import WavFrontendOnline
    rng = np.random.default_rng(5)
    for lfr_m, lfr_n in ((7, 6), (5, 1), (1, 1)):
        fe = WavFrontendOnline(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=lfr_m,
                               lfr_n=lfr_n, dither=0.0)
        for trial in range(6):
            n = int(rng.integers(100, 40000))
            w = (0.1 * rng.standard_normal(n)).astype(np.float32)
            orc = wf.OnlineFrontend(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=lfr_m,
                                    lfr_n=lfr_n)
            st = dict(carry=0, frames=0, rows=0)
            pos = 0
            while pos < n:
                m = int(min(n - pos, rng.choice([9600, 6400, 3840, 960, 300, 37])))
                fin = pos + m >= n
                want = orc.push(w[pos:pos + m], is_final=fin).shape[0]
                assert fe._rows_after_push(st, m, fin) == want, (lfr_m, lfr_n, n, pos, m)
                pos += m
            assert st == dict(carry=0, frames=0, rows=0)


def test_numpy_model_of_the_packed_fft_data_flow():
    """tools/model_s2.py restates quad_stage1 / quad_stage2 (fbank_tile.cuh) in numpy - the rotated second group, the
    17-column transpose, the column-0 FFT across the lanes of a group - and compares it with np.fft.rfft."""
    import runpy
    from pathlib import Path
    runpy.run_path(str(Path(__file__).resolve().parents[1] / "tools" / "model_s2.py"), run_name="__main__")


def test_compact_mel_weight_rule_against_the_reference_bank():
    """The fixed-shape mel stage stores one weight u per (bin, interval) and derives the other slope as
    min(1/4 - u, u * 2^100) (fbank_tile.cuh, B200FE_MEL_COMPACT).  Check the rule on the reference's own bank
    (TA:436-511): in float32, for every bin that feeds two filters, up + down == 1 to rounding, so the derived weight
    is within 2^-24 of the stored one and the mel energies move by less than 1e-6 relative."""
    import numpy as np
    import torch
    from torchaudio.compliance import kaldi
    bank, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)   # [80, 256]
    bank = bank.numpy().astype(np.float32)
    scale = np.float32(0.25)
    derived = np.zeros_like(bank)
    for k in range(1, bank.shape[1]):
        nz = np.nonzero(bank[:, k])[0]
        assert len(nz) <= 2 and (len(nz) < 2 or nz[1] == nz[0] + 1)
        if len(nz) == 2:                       # interval nz[1]: up-slope of nz[1], down-slope of nz[0]
            u = scale * bank[nz[1], k]
            d = np.minimum(scale - u, u * np.float32(2.0 ** 100))
            derived[nz[1], k] = u
            derived[nz[0], k] = d
            assert abs(float(d) - float(scale * bank[nz[0], k])) <= 2.0 ** -24
        elif len(nz) == 1:
            derived[nz[0], k] = scale * bank[nz[0], k]
    rng = np.random.default_rng(0)
    power = rng.random((64, 256)).astype(np.float64) * 1e6
    e_ref = power @ (0.25 * bank.astype(np.float64)).T
    e_new = power @ derived.astype(np.float64).T
    assert np.max(np.abs(e_new - e_ref) / e_ref) < 1e-6
    # padding slots: u == 0 -> both weights 0
    assert np.minimum(scale - np.float32(0), np.float32(0) * np.float32(2.0 ** 100)) == 0


def test_host_affinity_bind_and_restore_round_trip():
    """bench.py binds a rank to its GPU's NUMA node before the pinned allocations and gives every thread its original
    CPU set back before the CPU baseline runs (hostaffinity.restore_affinity walks /proc/self/task)."""
    import os
    import threading
    from toolbox_for_asr_and_tts_b200 import hostaffinity
    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no sched_getaffinity on this platform")
    before = os.sched_getaffinity(0)
    one = {min(before)}
    seen = {}
    started, release = threading.Event(), threading.Event()

    def worker():
        seen["tid"] = threading.get_native_id()
        started.set()
        release.wait(10)

    os.sched_setaffinity(0, one)
    try:
        t = threading.Thread(target=worker)
        t.start()                       # inherits the narrow mask, like a worker pool created while bound
        started.wait(10)
        assert os.sched_getaffinity(seen["tid"]) == one
        hostaffinity.restore_affinity(before)
        assert os.sched_getaffinity(0) == before and os.sched_getaffinity(seen["tid"]) == before
    finally:
        release.set()
        t.join()
        os.sched_setaffinity(0, before)
    assert hostaffinity.ingest_threads(1) >= 2


def test_work_claim_of_the_fused_kernel_is_not_warp_aggregated():
    """ptxas turns an atomic add on a provably warp-uniform address into a warp-aggregated sequence (vote, leader
    atomic, SHFL of the result); the shuffle waits for the atomic where it is issued, which costs the fused kernel
    1.3 % (profiles/r2_kernel_ab.txt).  fbank_warp.cuh routes the counter's address through a value the compiler cannot
    prove uniform; this guards the shipped SASS: one ATOMG per float32 instantiation, no vote / shuffle behind it, and
    the bulk copy (UBLKCP) and packed arithmetic (FFMA2) the design rests on are there."""
    import re
    import shutil
    from toolbox_for_asr_and_tts_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not _build.LIB.exists():
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_build.LIB)], capture_output=True, text=True, check=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and "fbank_warp_kernel" in name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            kernels.setdefault(name, []).append(re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", line).split(";")[0].strip())
    shipped = [k for k in kernels if "MelShapeFixedILi3ELi2ELi5ELi8EEELi10EfLb0" in k and "ILi25ELb1ELb0E" in k]
    assert len(shipped) == 1, sorted(kernels)[:4]
    ins = kernels[shipped[0]]
    ops = [re.sub(r"^@!?U?P\w+\s+", "", i).split()[0] for i in ins]
    atom = [i for i, o in enumerate(ops) if o.startswith("ATOMG")]
    assert len(atom) == 1
    behind = ops[atom[0] + 1:atom[0] + 8]
    assert not any(o.startswith(("SHFL", "VOTE", "POPC")) for o in behind), behind
    assert not any(o.startswith(("VOTE", "POPC", "FLO")) for o in ops[max(0, atom[0] - 8):atom[0]])
    assert sum(o.startswith("UBLKCP") for o in ops) >= 1 and sum(o.startswith("FFMA2") for o in ops) > 300
    assert not any(o.startswith(("LDL", "STL")) for o in ops)      # no spills
