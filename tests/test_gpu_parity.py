"""GPU suite (-m gpu): the CUDA path against the oracle, the committed golden vectors and size-independent
properties, all through the reference-shaped WavFrontend (torch extension -> C ABI -> sm_100a kernels) or through
the raw C ABI with ctypes.  Nothing here reads /root/reference."""
import ctypes

import numpy as np
import pytest
import torch
from conftest import LOGMEL_ATOL, PARAFORMER, VARIANT_CONFS, VARIANT_LENS, VARIANT_SEED, assert_logmel_close

from oracle import kaldi_fbank_np as kf
from oracle import wav_frontend_np as wf
from toolbox_for_asr_and_tts_b200 import StreamPool, WavFrontend, WavFrontendOnline, _native, stats_to_cmvn, synth

pytestmark = pytest.mark.gpu
SEED = 1234
DEV = "cuda:0"


def cmvn_atol(cmvn):
    return LOGMEL_ATOL * float(np.abs(cmvn[1]).max())


def assert_feats_close(got, ref, cmvn):
    """LFR+CMVN features [..., 560]: undo the CMVN and compare in the log-mel domain (conftest.assert_logmel_close)."""
    got = got.cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = ref.cpu().numpy() if isinstance(ref, torch.Tensor) else np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if got.size == 0:
        return
    sh, sc = cmvn[0].astype(np.float64), cmvn[1].astype(np.float64)
    m = cmvn.shape[1] // 80
    lg = (got.astype(np.float64) / sc - sh).reshape(got.shape[:-1] + (m, 80))
    lr = (ref.astype(np.float64) / sc - sh).reshape(ref.shape[:-1] + (m, 80))
    assert_logmel_close(lg, lr)


def make_fe(cmvn=None, **over):
    conf = dict(PARAFORMER, dither=0.0)
    conf.update(over)
    return WavFrontend(cmvn=None if cmvn is None else torch.from_numpy(cmvn), **conf)


def dense_batch(waves):
    nmax = max(len(w) for w in waves)
    buf = torch.zeros(len(waves), nmax)
    for i, w in enumerate(waves):
        buf[i, :len(w)] = torch.from_numpy(w)
    return buf.to(DEV)


def test_native_library_is_the_one_running():
    ops = _native.ops()
    assert torch.cuda.is_available()
    fe = make_fe()
    x = torch.zeros(1, 16000, device=DEV)
    before = fe.launch_count()
    fe(x, [16000])
    assert fe.launch_count() > before


def test_tables_match_reference_tables():
    fe = make_fe()
    win, mel = fe.tables()
    assert np.abs(win.numpy() - kf.window_function("hamming", 400)).max() <= 5e-7
    assert np.abs(mel.numpy() - kf.mel_banks(80, 512, 16000.0, dtype=np.float64)).max() <= 1e-7
    try:
        import torchaudio.compliance.kaldi as kaldi
    except Exception:
        return
    w = kaldi._feature_window_function("hamming", 400, 0.42, torch.device("cpu"), torch.float32)
    bank, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    assert (win - w).abs().max() <= 5e-7 and (mel - bank).abs().max() <= 3e-5


@pytest.mark.parametrize("n", [399, 400, 401, 559, 560, 1000, 16000, 160000, 480000])
def test_paraformer_against_golden(golden, cmvn, n):
    fe = make_fe(cmvn)
    x = torch.from_numpy(synth.uniform_pcm(SEED, n, n))[None].to(DEV)
    feats, lens = fe(x, [n])
    g = golden[f"paraformer_{n}"]
    assert feats.dtype == torch.float32 and lens.dtype == torch.int64
    assert tuple(feats.shape) == (1,) + g.shape and int(lens[0]) == g.shape[0]
    assert_feats_close(feats[0], g, cmvn)


@pytest.mark.parametrize("n", [400, 16000])
def test_povey_against_golden(golden, cmvn, n):
    fe = make_fe(cmvn, window="povey")
    x = torch.from_numpy(synth.uniform_pcm(SEED, n, n))[None].to(DEV)
    feats, _ = fe(x, [n])
    assert_feats_close(feats[0], golden[f"povey_{n}"], cmvn)


def test_forward_fbank_against_golden(golden):
    fe = make_fe()
    x = torch.from_numpy(synth.uniform_pcm(SEED, 16000, 16000))[None].to(DEV)
    feats, lens = fe.forward_fbank(x, [16000])
    assert tuple(feats.shape) == (1, 98, 80) and int(lens[0]) == 98
    assert_logmel_close(feats[0].cpu().numpy(), golden["fbank_16000"])


def test_ragged_batch_against_golden(golden, cmvn):
    fe = make_fe(cmvn)
    lens = golden["batch_input_lens"]
    waves = [synth.uniform_pcm(SEED + 1, i, int(n)) for i, n in enumerate(lens)]
    feats, flens = fe(dense_batch(waves), lens.tolist())
    assert np.array_equal(flens.cpu().numpy(), golden["batch_lens"])
    f = feats.cpu().numpy()
    assert f.shape == golden["batch_feats"].shape
    assert_feats_close(f, golden["batch_feats"], cmvn)
    for i, k in enumerate(golden["batch_lens"]):
        assert not f[i, k:].any()          # pad_sequence zeros, bit-exact


def test_gaussian_with_dc_against_golden(golden, cmvn):
    fe = make_fe(cmvn)
    feats, _ = fe(torch.from_numpy(golden["gauss_input"])[None].to(DEV), [24000])
    assert_feats_close(feats[0], golden["gauss_feats"], cmvn)


def test_counts_are_bit_exact_for_many_lengths():
    fe = make_fe()
    ns = list(range(400, 2400, 7)) + [15999, 16000, 16001, 479999, 480000]
    nf, nr = fe.frame_counts(ns)
    for n, a, b in zip(ns, nf.tolist(), nr.tolist()):
        t = 1 + (n - 400) // 160
        assert a == t and b == -(-t // 6), n


def test_random_ragged_batch_against_oracle(cmvn):
    """32 utterances, 0.2-6 s, against the oracle; log-mel error statistics are printed for DESIGN.md."""
    fe = make_fe(cmvn)
    lens = synth.utterance_lengths(21, 32, lo=3200, hi=96000)
    waves = [synth.uniform_pcm(21, i, int(n)) for i, n in enumerate(lens)]
    feats, flens = fe(dense_batch(waves), lens.tolist())
    ref, rlens = wf.frontend_forward(waves, lens, cmvn=cmvn, **PARAFORMER)
    assert np.array_equal(flens.cpu().numpy(), rlens)
    d = np.abs(feats.cpu().numpy() - ref)
    assert_feats_close(feats, ref, cmvn)
    print("ragged batch: max-abs", d.max(), "mean-abs", d.mean())


def test_packed_unaligned_offsets_equal_dense_bitwise(cmvn):
    """Length-packed input with arbitrary (unaligned) offsets must give the dense-layout result bit for bit."""
    fe = make_fe(cmvn)
    lens = np.array([4001, 16003, 401, 7777, 32000, 1601], dtype=np.int64)
    waves = [synth.uniform_pcm(31, i, int(n)) for i, n in enumerate(lens)]
    dense, dl = fe(dense_batch(waves), lens.tolist())
    for align, lead in ((1, 1), (1, 3), (4, 0), (2, 2)):
        offs, total = synth.packed_offsets(lens, align=align)
        offs = offs + lead
        flat = torch.zeros(int(total) + lead + 8)
        for o, w in zip(offs, waves):
            flat[o:o + len(w)] = torch.from_numpy(w)
        packed, pl = fe.forward_packed(flat.to(DEV), offs, lens)
        assert torch.equal(pl.cpu(), dl) and torch.equal(packed, dense), (align, lead)


def test_short_utterances_follow_shrunken_frame_rule(cmvn):
    """VF:147: frame_length = min(25, len/fs*1000): one frame, smaller window / FFT / mel bank."""
    fe = make_fe(cmvn)
    for n in (399, 300, 257, 256, 129, 64, 33):
        x = synth.uniform_pcm(41, n, n)
        feats, lens = fe(torch.from_numpy(x)[None].to(DEV), [n])
        ref, rl = wf.frontend_forward([x], [n], cmvn=cmvn, **PARAFORMER)
        assert int(lens[0]) == int(rl[0]) and tuple(feats.shape) == ref.shape
        assert np.abs(feats.cpu().numpy() - ref).max() <= cmvn_atol(cmvn), n


def test_forward_lfr_cmvn_is_bit_exact(cmvn):
    fe = make_fe(cmvn)
    rng = np.random.default_rng(3)
    lens = [1, 5, 6, 7, 98, 333]
    x = np.zeros((len(lens), max(lens), 80), dtype=np.float32)
    for i, t in enumerate(lens):
        x[i, :t] = rng.standard_normal((t, 80)).astype(np.float32) * 4 + 10
    out, ol = fe.forward_lfr_cmvn(torch.from_numpy(x).to(DEV), lens)
    for i, t in enumerate(lens):
        ref = wf.apply_cmvn(wf.apply_lfr(x[i, :t], 7, 6), cmvn)
        assert int(ol[i]) == ref.shape[0]
        assert np.array_equal(out[i, :ref.shape[0]].cpu().numpy(), ref)
        assert not out[i, ref.shape[0]:].any()


def test_linearity_property_at_full_size():
    """Size-independent property at BASELINE.json's full utterance size (30 s): the power spectrum is quadratic, so
    scaling the input by 2 shifts every log-mel value by log(4) (well above the floor)."""
    fe = make_fe(None, lfr_m=1, lfr_n=1)
    x = torch.from_numpy(synth.uniform_pcm(51, 0, 480000))[None].to(DEV)
    a, la = fe(x, [480000])
    b, lb = fe(x * 2, [480000])
    assert int(la[0]) == 2998 and torch.equal(la, lb)
    assert (b - a - float(np.log(4.0))).abs().max() < 2e-5


def test_time_shift_property():
    """Dropping exactly one hop (160 samples) from the front shifts the frame sequence by one, bit for bit up to the
    pairing of frames inside one FFT (tolerance covers that)."""
    fe = make_fe(None, lfr_m=1, lfr_n=1)
    w = synth.uniform_pcm(52, 0, 64000)
    a, _ = fe(torch.from_numpy(w)[None].to(DEV), [64000])
    b, _ = fe(torch.from_numpy(w[160:])[None].to(DEV), [64000 - 160])
    assert (a[0, 1:] - b[0]).abs().max() < 5e-4   # low-energy bins amplify the fp32 rounding differences


@pytest.mark.parametrize("chunk", [9600, 3840, 6400, 960, 300])
def test_streaming_concat_equals_offline(cmvn, chunk):
    fe = make_fe(cmvn)
    n_streams = 5
    lens = [48000, 31999, 9600 * 3, 2000, 16000]
    waves = [synth.uniform_pcm(61, i, n) for i, n in enumerate(lens)]
    off, off_lens = fe(dense_batch(waves), lens)
    pool = StreamPool(fe, n_streams=8, max_chunk_samples=9600, device=DEV)
    got = [[] for _ in range(n_streams)]
    pos = [0] * n_streams
    while any(p < n for p, n in zip(pos, lens)):
        ids, chunks, clens, fins = [], [], [], []
        for s in range(n_streams):
            if pos[s] < lens[s]:
                m = min(chunk, lens[s] - pos[s])
                c = np.zeros(9600, dtype=np.float32)
                c[:m] = waves[s][pos[s]:pos[s] + m]
                pos[s] += m
                ids.append(s); chunks.append(c); clens.append(m); fins.append(1 if pos[s] >= lens[s] else 0)
        feats, rows = pool.push(torch.from_numpy(np.stack(chunks)).to(DEV), torch.tensor(clens, dtype=torch.int32),
                                torch.tensor(ids, dtype=torch.int32), torch.tensor(fins, dtype=torch.uint8))
        rows = rows.cpu().tolist()
        for k, s in enumerate(ids):
            if rows[k]:
                got[s].append(feats[k, :rows[k]].cpu())
    for s in range(n_streams):
        cat = torch.cat(got[s], dim=0) if got[s] else torch.zeros(0, 560)
        assert cat.shape[0] == int(off_lens[s]), (s, chunk)
        # same kernels, but frames pair up differently inside the packed FFT: fp32 rounding noise only
        assert_feats_close(cat, off[s, :cat.shape[0]], cmvn)


def test_streaming_600ms_row_schedule_and_reference_shaped_api(cmvn):
    """10 rows per 600 ms chunk from the first chunk on, 7 on the final flush, through the
    upstream-shaped WavFrontendOnline.forward(input, lengths, cache=..., is_final=...)."""
    fe = WavFrontendOnline(cmvn=torch.from_numpy(cmvn), max_chunk_samples=9600, **dict(PARAFORMER, dither=0.0))
    off = make_fe(cmvn)
    w = synth.uniform_pcm(62, 0, 160000)
    full, fl = off(torch.from_numpy(w)[None].to(DEV), [160000])
    cache, outs, per = {}, [], []
    for s in range(0, 160000, 9600):
        c = torch.from_numpy(w[s:s + 9600])[None].to(DEV)
        f, l = fe(c, [c.shape[1]], cache=cache, is_final=(s + 9600 >= 160000))
        per.append(int(l[0]))
        if f.numel():
            outs.append(f[0])
    assert per[:3] == [10, 10, 10] and per[-1] == 7 and sum(per) == 167 == int(fl[0])
    assert_feats_close(torch.cat(outs), full[0], cmvn)


def test_global_cmvn_statistics():
    """sum / sum of squares / count accumulated by the kernel equal float64 sums over the un-normalised features."""
    fe = make_fe(None)
    lens = synth.utterance_lengths(71, 24, lo=8000, hi=80000)
    waves = [synth.uniform_pcm(71, i, int(n)) for i, n in enumerate(lens)]
    stats = torch.zeros(2 * 560 + 1, dtype=torch.float64, device=DEV)
    feats, fl = fe(dense_batch(waves), lens.tolist(), stats=stats)
    rows = torch.cat([feats[i, :int(k)] for i, k in enumerate(fl)]).double()
    ref = torch.cat([rows.sum(0), (rows * rows).sum(0), torch.tensor([float(rows.shape[0])], device=DEV, dtype=torch.float64)])
    assert float(stats[-1]) == float(ref[-1])
    rel = ((stats - ref).abs() / ref.abs().clamp(min=1.0)).max()
    assert float(rel) < 1e-9, float(rel)
    tab = stats_to_cmvn(stats)
    normed = (rows.float().cpu() + tab[0]) * tab[1]
    assert normed.mean(0).abs().max() < 1e-3 and (normed.std(0) - 1).abs().max() < 1e-2
    # against the oracle's float64 statistics of its own features: equal within the feature tolerance
    mats = [wf.frontend_forward([w], [len(w)], cmvn=None, **PARAFORMER)[0][0] for w in waves]
    s, s2, n = wf.cmvn_stats(mats)
    assert n == int(stats[-1])
    assert np.abs(stats[:560].cpu().numpy() / n - s / n).max() < 1e-4


def test_dither_is_statistically_kaldi_dither():
    """dither != 0 cannot be bit-matched (TA:179-181 draws torch.randn per (frame, sample)); compare statistics on a
    silent input, where the features are entirely the dither."""
    fe = make_fe(None, lfr_m=1, lfr_n=1, dither=1.0)
    n = 960000      # 5998 frames: the standard error of a per-bin mean of log chi-square noise is ~0.015
    x = torch.zeros(1, n, device=DEV)
    a, _ = fe(x, [n])
    b, _ = fe(x, [n])
    assert not torch.equal(a, b)                     # new noise on every call, like torch.randn
    rng = np.random.default_rng(0)
    ref = kf.fbank(np.zeros(n, dtype=np.float32), num_mel_bins=80, dither=1.0, energy_floor=0.0, window_type="hamming",
                   rng=rng)
    got = a[0].cpu().numpy()
    assert np.abs(got.mean(0) - ref.mean(0)).max() < 0.1
    assert np.abs(got.std(0) - ref.std(0)).max() < 0.1
    fe2 = make_fe(None, lfr_m=1, lfr_n=1, dither=1.0, dither_seed=5)
    fe3 = make_fe(None, lfr_m=1, lfr_n=1, dither=1.0, dither_seed=5)
    assert torch.equal(fe2(x, [n])[0], fe3(x, [n])[0])   # reproducible from the seed


def test_cpu_tensor_is_rejected_and_unsupported_options_raise():
    fe = make_fe()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fe(torch.zeros(1, 16000), [16000])
    with pytest.raises(NotImplementedError):
        make_fe(None, snip_edges=False)(torch.zeros(1, 16000, device=DEV), [16000])
    with pytest.raises(RuntimeError, match="window size"):
        fe(torch.zeros(1, 16, device=DEV), [1])


def ctypes_handle(cmvn):
    """(lib, handle) through the plain C ABI: Paraformer configuration, dither 0."""
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
    c.lfr_m, c.lfr_n, c.dither = 7, 6, 0.0
    h = ctypes.c_void_p()
    cm = np.ascontiguousarray(cmvn, dtype=np.float32)
    lib.b200fe_last_error.restype = ctypes.c_char_p
    rc = lib.b200fe_create(ctypes.byref(c), cm.ctypes.data_as(ctypes.c_void_p), ctypes.byref(h))
    assert rc == 0, lib.b200fe_last_error(None)
    return lib, h


def test_raw_c_abi_with_ctypes(cmvn):
    """The C ABI as a foreign caller would use it: plain pointers and sizes, no torch types in the signatures."""
    lib, h = ctypes_handle(cmvn)
    try:
        n = 16000
        lens = (ctypes.c_int64 * 1)(n)
        nf, nr, mx = (ctypes.c_int64 * 1)(), (ctypes.c_int64 * 1)(), ctypes.c_int64()
        ws = ctypes.c_size_t()
        assert lib.b200fe_plan(h, lens, 1, nf, nr, ctypes.byref(mx), ctypes.byref(ws)) == 0
        assert (nf[0], nr[0], mx.value) == (98, 17, 17)
        x = torch.from_numpy(synth.uniform_pcm(SEED, n, n)).to(DEV)
        feats = torch.empty(1, 17, 560, device=DEV)
        flen = torch.empty(1, dtype=torch.int64, device=DEV)
        work = torch.empty(max(ws.value, 256), dtype=torch.uint8, device=DEV)
        rc = lib.b200fe_forward(h, ctypes.c_void_p(x.data_ptr()), ctypes.c_int64(n), None, ctypes.c_int64(n), lens, 1,
                                ctypes.c_void_p(feats.data_ptr()), ctypes.c_int64(17), ctypes.c_void_p(flen.data_ptr()),
                                None, ctypes.c_uint64(0), ctypes.c_void_p(work.data_ptr()), ctypes.c_size_t(work.numel()),
                                None)
        assert rc == 0, lib.b200fe_last_error(h)
        torch.cuda.synchronize()
        g = dict(np.load(__import__("conftest").GOLDEN))["paraformer_16000"]
        assert int(flen[0]) == 17
        assert_feats_close(feats[0], g, cmvn)
    finally:
        lib.b200fe_destroy(h)


def test_streaming_512_streams_per_tick_match_single_stream_bitwise(cmvn):
    """BASELINE.json configs[2] at one GPU's share: 512 concurrent streams, 600 ms chunks, one launch per tick.
    Every stream must produce exactly what it produces when it is the only stream in the pool (state isolation), and
    the row schedule must be 10 rows per tick from the first tick on."""
    fe = make_fe(cmvn)
    n_streams, chunk, ticks = 512, 9600, 6
    pool = StreamPool(fe, n_streams=n_streams, max_chunk_samples=chunk, device=DEV)
    ids = torch.arange(n_streams, dtype=torch.int32, device=DEV)
    lens = torch.full((n_streams,), chunk, dtype=torch.int32, device=DEV)
    offs = torch.arange(n_streams, dtype=torch.int64) * (chunk * ticks)
    wave = torch.zeros(n_streams * chunk * ticks, device=DEV)
    _native.ops().synth_uniform(wave, offs, torch.full((n_streams,), chunk * ticks, dtype=torch.int64), 77, 0.3)
    wave = wave.view(n_streams, ticks, chunk)
    outs = []
    for t in range(ticks):
        fin = torch.full((n_streams,), 1 if t == ticks - 1 else 0, dtype=torch.uint8, device=DEV)
        feats, rows = pool.push(wave[:, t].contiguous(), lens, ids, fin)
        expect = 10 if t < ticks - 1 else -(-(1 + (chunk * ticks - 400) // 160) // 6) - 10 * (ticks - 1)
        assert bool((rows == expect).all()), (t, rows[:4])
        outs.append(feats[:, :expect].clone())
    allrows = torch.cat(outs, dim=1)                      # [512, 60, 560]
    solo = StreamPool(fe, n_streams=1, max_chunk_samples=chunk, device=DEV)
    one = torch.zeros(1, dtype=torch.int32, device=DEV)
    for s in (0, 1, 255, 511):
        got = []
        for t in range(ticks):
            fin = torch.tensor([1 if t == ticks - 1 else 0], dtype=torch.uint8, device=DEV)
            f, r = solo.push(wave[s:s + 1, t].contiguous(), lens[:1], one, fin)
            got.append(f[0, :int(r[0])])
        assert torch.equal(torch.cat(got), allrows[s]), s
    # and against the offline front-end on the concatenated audio
    off, ol = fe(wave[:4].reshape(4, -1), [chunk * ticks] * 4)
    assert int(ol[0]) == allrows.shape[1]
    assert_feats_close(off, allrows[:4], cmvn)


def test_sharded_bulk_extraction_equals_single_pass(cmvn):
    """BASELINE.json configs[3] in miniature: utterances sharded over 8 'ranks' (longest-first), each shard processed
    on its own, CMVN statistics summed.  Features of shard k must equal the single-pass features bit for bit and the
    reduced statistics must equal the single-pass statistics (float64 sums in a different order)."""
    from toolbox_for_asr_and_tts_b200 import sharding
    fe = make_fe(None)
    lens = synth.utterance_lengths(81, 40, lo=4000, hi=64000)
    waves = [synth.uniform_pcm(81, i, int(n)) for i, n in enumerate(lens)]
    stats_all = torch.zeros(1121, dtype=torch.float64, device=DEV)
    full, fl = fe(dense_batch(waves), lens.tolist(), stats=stats_all)
    stats_sum = torch.zeros(1121, dtype=torch.float64, device=DEV)
    for part in sharding.partition_utterances(lens, 8):
        st = torch.zeros(1121, dtype=torch.float64, device=DEV)
        f, l = fe(dense_batch([waves[i] for i in part]), [int(lens[i]) for i in part], stats=st)
        for k, i in enumerate(part):
            assert int(l[k]) == int(fl[i]) and torch.equal(f[k, :int(l[k])], full[i, :int(fl[i])])
        stats_sum += st            # what the NCCL all-reduce does across ranks
    assert float(stats_sum[-1]) == float(stats_all[-1])
    rel = ((stats_sum - stats_all).abs() / stats_all.abs().clamp(min=1.0)).max()
    assert float(rel) < 1e-12


def test_tts_log_mel_against_frozen_definition():
    """BASELINE.json configs[4]: 24 kHz, n_fft 1024, hop 256, 80 mels.  PARITY UNPINNED by the reference (it has no
    audio->mel code); the CUDA path is compared with the frozen numpy definition (oracle/tts_mel_np.py)."""
    from oracle import tts_mel_np as tm
    from toolbox_for_asr_and_tts_b200 import TtsLogMel
    fe = TtsLogMel()
    lens = [24000, 72000, 1280, 512, 1025, 385 + 256, 100003]
    waves = [synth.uniform_pcm(91, i, n) for i, n in enumerate(lens)]
    mel, frames = fe(dense_batch(waves), lens)
    assert mel.dtype == torch.float32 and tuple(mel.shape) == (len(lens), 80, max(lens) // 256)
    for i, n in enumerate(lens):
        ref = tm.tts_log_mel(waves[i])
        assert int(frames[i]) == n // 256 == ref.shape[1]
        got = mel[i, :, :ref.shape[1]].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-3, (n, np.abs(got - ref).max())
        assert not mel[i, :, ref.shape[1]:].any()           # padded frames are zero
    with pytest.raises(RuntimeError):
        fe(torch.zeros(1, 24000), [24000])                  # CPU tensors are rejected


def test_tts_log_mel_large_batch_and_other_geometry():
    """The TTS kernel's work list is built on the device (prefix sums of frame pairs over the batch, 1 024 clips per
    scan step): a batch of 1 100 short clips crosses that step.  And a second geometry (16 kHz, hop 320, 64 mels up to
    7.6 kHz) takes the run-time mel shape and a different per-warp sample buffer.  Both against the frozen numpy definition."""
    from oracle import tts_mel_np as tm
    from toolbox_for_asr_and_tts_b200 import TtsLogMel
    rng = np.random.default_rng(5)
    lens = rng.integers(385, 1400, 1100)
    waves = [synth.uniform_pcm(95, i, int(n)) for i, n in enumerate(lens)]
    fe = TtsLogMel()
    mel, frames = fe(dense_batch(waves), lens.tolist())
    assert frames.cpu().tolist() == (lens // 256).tolist()
    for i in [0, 1, 511, 1023, 1024, 1025, 1099] + rng.integers(0, 1100, 8).tolist():
        ref = tm.tts_log_mel(waves[i])
        assert np.abs(mel[i, :, :ref.shape[1]].cpu().numpy() - ref).max() <= 1e-3, i
        assert not mel[i, :, ref.shape[1]:].any()
    conf = dict(sample_rate=16000, n_fft=1024, hop_length=320, n_mels=64, f_min=50.0, f_max=7600.0)
    fe2 = TtsLogMel(**conf)
    lens2 = [16000, 48001, 353, 5000, 640]
    waves2 = [synth.uniform_pcm(96, i, n) for i, n in enumerate(lens2)]
    mel2, frames2 = fe2(dense_batch(waves2), lens2)
    assert tuple(mel2.shape) == (5, 64, max(lens2) // 320)
    for i, n in enumerate(lens2):
        ref = tm.tts_log_mel(waves2[i], sample_rate=16000, n_fft=1024, hop=320, n_mels=64, f_min=50.0, f_max=7600.0)
        assert int(frames2[i]) == n // 320 == ref.shape[1]
        assert np.abs(mel2[i, :, :ref.shape[1]].cpu().numpy() - ref).max() <= 1e-3, (n, np.abs(mel2[i, :, :ref.shape[1]].cpu().numpy() - ref).max())
        assert not mel2[i, :, ref.shape[1]:].any()


@pytest.mark.parametrize("m,n", [(7, 6), (5, 1), (1, 1), (3, 2), (9, 4)])
def test_warp_kernel_equals_tile_kernel(m, n):
    """The warp-autonomous kernel (quad list, scattered LFR rows) and the tile kernel (row-major LFR pass) run the
    same arithmetic up to which half of a warp a frame lands in (the two halves use differently rotated sample rows
    and twiddle tables): same shapes, lengths and zero padding, values equal to float32 rounding noise, whatever the
    frame count modulo 4 / lfr_n, including utterances of 1..3 frames and the LFR settings of the reference's other
    models (FSMN-VAD uses 5/1).  Both are then checked against the oracle."""
    rng = np.random.default_rng(m * 10 + n)
    cm = np.stack([rng.normal(-8.0, 1.0, m * 80), rng.uniform(0.2, 0.5, m * 80)]).astype(np.float32)
    conf = dict(PARAFORMER, dither=0.0, lfr_m=m, lfr_n=n)
    lens = np.array([400, 559, 560, 720, 880, 1040, 1200, 4001, 16003, 7777, 32000, 1601, 48017], dtype=np.int64)
    waves = [synth.uniform_pcm(77, i, int(k)) for i, k in enumerate(lens)]
    x = dense_batch(waves)
    fe_w = WavFrontend(cmvn=torch.from_numpy(cm), **conf)
    fe_t = WavFrontend(cmvn=torch.from_numpy(cm), **conf)
    fe_t.select_kernel("tile")
    a, la = fe_w(x, lens.tolist())
    b, lb = fe_t(x, lens.tolist())
    assert torch.equal(la, lb) and a.shape == b.shape
    assert torch.equal(a == 0, b == 0)
    assert_feats_close(a, b, cm)
    ref, rl = wf.frontend_forward(waves, lens, cmvn=cm, **dict(PARAFORMER, lfr_m=m, lfr_n=n))
    assert np.array_equal(la.cpu().numpy(), rl)
    got = a.cpu().numpy()
    sh, sc = cm[0].astype(np.float64), cm[1].astype(np.float64)
    lg = (got.astype(np.float64) / sc - sh).reshape(got.shape[:-1] + (m, 80))
    lr = (ref.astype(np.float64) / sc - sh).reshape(ref.shape[:-1] + (m, 80))
    for i, k in enumerate(rl):
        assert_logmel_close(lg[i, :k], lr[i, :k])
        assert not got[i, k:].any()


def test_fsmn_vad_frontend_configuration():
    """the FSMN-VAD front-end the reference runs per chunk (R:voice_interface.py:1585-1590) is the same
    fbank with LFR 5/1; offline and streaming in 600 ms chunks against the oracle."""
    rng = np.random.default_rng(5)
    cm = np.stack([rng.normal(-8.0, 1.0, 400), rng.uniform(0.2, 0.5, 400)]).astype(np.float32)
    conf = dict(PARAFORMER, lfr_m=5, lfr_n=1)
    w = synth.uniform_pcm(63, 0, 48000)
    fe = WavFrontend(cmvn=torch.from_numpy(cm), dither=0.0, **conf)
    feats, fl = fe(torch.from_numpy(w)[None].to(DEV), [48000])
    ref, rl = wf.frontend_forward([w], [48000], cmvn=cm, **conf)
    assert int(fl[0]) == int(rl[0]) == 298
    assert_feats_close(feats[0], ref[0], cm)
    on = WavFrontendOnline(cmvn=torch.from_numpy(cm), max_chunk_samples=9600, dither=0.0, **conf)
    cache, outs = {}, []
    for s in range(0, 48000, 9600):
        c = torch.from_numpy(w[s:s + 9600])[None].to(DEV)
        f, l = on(c, [9600], cache=cache, is_final=(s + 9600 >= 48000))
        if f.numel():
            outs.append(f[0])
    cat = torch.cat(outs)
    assert cat.shape[0] == 298
    assert_feats_close(cat, feats[0], cm)


def test_pcm16_input_is_bit_identical_to_converted_float(cmvn):
    """int16 PCM in its wire format, converted inside the kernel's loads with the reference's rule
    s / 32768 (R:voice_interface.py:1008-1013), must give exactly what the float path gives for the converted buffer:
    dense and length-packed with odd offsets, including an utterance shorter than one frame."""
    fe = make_fe(cmvn)
    rng = np.random.default_rng(16)
    lens = np.array([4001, 16003, 399, 7777, 32000, 401, 1601], dtype=np.int64)
    ints = [rng.integers(-20000, 20000, size=int(n), dtype=np.int16) for n in lens]
    floats = [w.astype(np.float32) / np.float32(32768.0) for w in ints]
    nmax = int(lens.max())
    xi = torch.zeros(len(lens), nmax, dtype=torch.int16)
    for i, w in enumerate(ints):
        xi[i, :len(w)] = torch.from_numpy(w)
    ref_f, ref_l = fe(dense_batch(floats), lens.tolist())
    got, gl = fe(xi.to(DEV), lens.tolist())
    assert got.dtype == torch.float32 and torch.equal(gl, ref_l) and torch.equal(got, ref_f)
    for lead in (0, 1, 3, 5):
        offs, total = synth.packed_offsets(lens, align=1)
        offs = offs + lead
        flat = torch.zeros(int(total) + lead + 16, dtype=torch.int16)
        for o, w in zip(offs, ints):
            flat[o:o + len(w)] = torch.from_numpy(w)
        p, pl = fe.forward_packed(flat.to(DEV), offs, lens)
        assert torch.equal(pl.cpu(), ref_l) and torch.equal(p, ref_f), lead
    oracle, ol = wf.frontend_forward(floats, lens, cmvn=cmvn, **PARAFORMER)
    assert np.array_equal(gl.cpu().numpy(), ol)
    for i, k in enumerate(ol):
        assert_feats_close(got[i, :k], oracle[i, :k], cmvn)
    with pytest.raises(RuntimeError, match="statistics pass takes float32"):
        fe.forward_packed(flat.to(DEV), offs, lens, stats=torch.zeros(1121, dtype=torch.float64, device=DEV))


def test_audio_statistics_match_reference_formulas():
    """max / min / mean|x| / rms / clipping ratio of a ragged batch (R:voice_interface.py:873-939)."""
    rng = np.random.default_rng(9)
    lens = np.array([1, 400, 16000, 48017, 0, 7777], dtype=np.int64)
    waves = [np.clip(0.5 * rng.standard_normal(int(n)), -1, 1).astype(np.float32) for n in lens]
    waves[3][:100] = 1.0
    waves[3][100:150] = -0.9995
    ref = np.stack([wf.audio_statistics(w) for w in waves])
    nmax = int(lens.max())
    buf = torch.zeros(len(lens), nmax)
    for i, w in enumerate(waves):
        buf[i, :len(w)] = torch.from_numpy(w)
    got = WavFrontend.audio_statistics(buf.to(DEV), lens).cpu().numpy()
    assert got.shape == (6, 6)
    assert np.array_equal(got[:, [0, 1, 5]], ref[:, [0, 1, 5]])            # max, min, max |x|: exact
    assert np.array_equal(got[:, 4], ref[:, 4])                            # clipping ratio: exact counts
    assert np.allclose(got[:, [2, 3]], ref[:, [2, 3]], rtol=1e-12, atol=0)   # float64 sums
    offs, total = synth.packed_offsets(lens, align=1)
    flat = torch.zeros(int(total) + 4)
    for o, w in zip(offs, waves):
        flat[o:o + len(w)] = torch.from_numpy(w)
    got2 = WavFrontend.audio_statistics(flat.to(DEV), lens, offsets=offs).cpu().numpy()
    assert np.array_equal(got2[:, [0, 1, 4, 5]], got[:, [0, 1, 4, 5]]) and np.allclose(got2, got, rtol=1e-12)


def test_subtract_mean_speaker_verification_features():
    """the CAM++ front-end = 80-mel Kaldi fbank + utterance mean normalisation (Kaldi subtract_mean,
    TA:642-644), against torchaudio itself when present and the oracle."""
    lens = [16000, 4001, 32000]
    waves = [synth.uniform_pcm(91, i, n) for i, n in enumerate(lens)]
    fe = WavFrontend(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, dither=0.0, subtract_mean=True)
    feats, fl = fe(dense_batch(waves), lens)
    for i, w in enumerate(waves):
        fb = kf.fbank(w * np.float32(32768.0), num_mel_bins=80, frame_length=25.0, frame_shift=10.0, dither=0.0,
                      energy_floor=0.0, window_type="hamming", sample_frequency=16000.0, dtype=np.float32)
        ref = wf.subtract_column_mean(fb)
        k = int(fl[i])
        assert k == ref.shape[0]
        got = feats[i, :k].cpu().numpy()
        assert np.abs(got - ref).mean() <= 2e-5 and np.abs(got - ref).max() <= 3e-3
        assert np.abs(got.mean(axis=0)).max() <= 1e-4            # columns are centred
        assert not feats[i, k:].any()
    try:
        import torchaudio.compliance.kaldi as kaldi
    except Exception:
        return
    t = kaldi.fbank(torch.from_numpy(waves[0])[None] * 32768.0, num_mel_bins=80, dither=0.0, energy_floor=0.0,
                    window_type="hamming", sample_frequency=16000.0, subtract_mean=True)
    assert (feats[0, :int(fl[0])].cpu() - t).abs().max() <= 3e-3
    with pytest.raises(NotImplementedError):
        WavFrontend(lfr_m=7, lfr_n=6, subtract_mean=True)


@pytest.mark.parametrize("opts", [
    dict(frame_length=32, frame_shift=10, n_mels=40, window="povey"),
    dict(frame_length=20, frame_shift=10, n_mels=64, window="hanning", preemphasis_coefficient=0.0, remove_dc_offset=False),
    dict(frame_length=25, frame_shift=10, n_mels=80, window="blackman", low_freq=100.0, high_freq=-200.0),
    dict(frame_length=25, frame_shift=5, n_mels=80, window="rectangular"),
    dict(frame_length=25, frame_shift=20, n_mels=24, window="hamming"),
])
@pytest.mark.parametrize("kernel", ["auto", "tile"])
def test_kaldi_option_variants_against_oracle(opts, kernel):
    """The Kaldi fbank options surface (TA:514-541) away from the Paraformer values: other frame lengths / shifts
    (generic 32-row FFT path, frames that are not whole 16-sample rows), mel counts, windows, no pre-emphasis, no DC
    removal, band limits - through both fused kernels, fbank only."""
    lens = [16000, 4001, 1603]
    waves = [synth.uniform_pcm(55, i, n) for i, n in enumerate(lens)]
    fe = WavFrontend(fs=16000, dither=0.0, **opts)
    fe.select_kernel(kernel)
    feats, fl = fe.forward_fbank(dense_batch(waves), lens)
    kw = dict(num_mel_bins=opts["n_mels"], frame_length=float(opts["frame_length"]), frame_shift=float(opts["frame_shift"]),
              dither=0.0, energy_floor=0.0, window_type=opts["window"], sample_frequency=16000.0,
              preemphasis_coefficient=opts.get("preemphasis_coefficient", 0.97),
              remove_dc_offset=opts.get("remove_dc_offset", True), low_freq=opts.get("low_freq", 20.0),
              high_freq=opts.get("high_freq", 0.0), dtype=np.float32)
    for i, w in enumerate(waves):
        ref = kf.fbank(w * np.float32(32768.0), **kw)
        k = int(fl[i])
        assert k == ref.shape[0], (k, ref.shape)
        assert_logmel_close(feats[i, :k].cpu().numpy(), ref)
        assert not feats[i, k:].any()


@pytest.mark.parametrize("name", sorted(VARIANT_CONFS))
def test_kaldi_variants_against_torchaudio_golden(variants_golden, name):
    """The CUDA path against torchaudio's own output (committed fixtures) for the option sets of
    tests/golden/make_golden_variants.py, incl. subtract_mean."""
    fe = WavFrontend(fs=16000, dither=0.0, **VARIANT_CONFS[name])
    waves = [synth.uniform_pcm(VARIANT_SEED, i, n) for i, n in enumerate(VARIANT_LENS)]
    feats, fl = fe.forward_fbank(dense_batch(waves), list(VARIANT_LENS))
    for i, n in enumerate(VARIANT_LENS):
        g = variants_golden[f"{name}_{n}"]
        assert int(fl[i]) == g.shape[0]
        assert_logmel_close(feats[i, :g.shape[0]].cpu().numpy(), g)


def test_fsmn_vad_frontend_against_funasr_golden(variants_golden):
    cm = variants_golden["vad_5_1_cmvn"]
    fe = WavFrontend(cmvn=torch.from_numpy(cm), dither=0.0, **dict(PARAFORMER, lfr_m=5, lfr_n=1))
    w = synth.uniform_pcm(63, 0, 48000)
    feats, fl = fe(torch.from_numpy(w)[None].to(DEV), [48000])
    g = variants_golden["vad_5_1_feats"]
    assert int(fl[0]) == g.shape[0]
    assert_feats_close(feats[0], g, cm)


@pytest.mark.parametrize("dtype,channels,src", [(np.int16, 1, 16000), (np.int16, 2, 48000), (np.int16, 1, 8000),
                                                (np.int16, 1, 44100), (np.int32, 2, 22050), (np.uint8, 1, 16000),
                                                (np.uint8, 3, 32000)])
def test_ingest_pcm_is_bit_identical_to_numpy(dtype, channels, src):
    """wire PCM -> float32 mono 16 kHz (R:voice_interface.py:1004-1034, numpy resampling branch)."""
    rng = np.random.default_rng(int(src) + channels)
    info = np.iinfo(dtype)
    for n in (1, 2, 7, 1600, 44101):
        raw = rng.integers(info.min, info.max, size=n * channels, dtype=dtype, endpoint=True)
        ref = wf.ingest_pcm(raw, channels, src)
        t = torch.from_numpy(raw.view(np.uint8) if dtype == np.uint8 else raw).to(DEV)
        got = WavFrontend.ingest_pcm(t, channels=channels, src_rate=src, method="interp").cpu().numpy()
        assert got.dtype == np.float32 and got.shape == ref.shape, (n, got.shape, ref.shape)
        assert np.array_equal(got, ref), (n, np.abs(got - ref).max() if got.size else 0)


def test_handles_follow_the_tensor_device(cmvn):
    """One front-end object used from two GPUs of a box: the native tables are per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fe = make_fe(cmvn)
    x = torch.from_numpy(synth.uniform_pcm(SEED, 16000, 16000))[None]
    a, la = fe(x.to("cuda:0"), [16000])
    b, lb = fe(x.to("cuda:1"), [16000])
    assert b.device.index == 1 and lb.device.type == "cpu" and lb.dtype == torch.int64     # lengths: CPU int64, VF:160
    assert torch.equal(a.cpu(), b.cpu()) and torch.equal(la.cpu(), lb.cpu())


def test_ten_minute_utterance_matches_oracle_on_excerpts(cmvn):
    """Large indices: one 10-minute utterance (59 998 frames, 10 000 rows) next to a short one.  Frames depend on local
    samples only and LFR rows on 7 neighbouring frames, so an excerpt that starts on a multiple of 6 frames reproduces
    the interior rows of the long utterance: compared with the oracle on three excerpts (start, middle, end)."""
    fe = make_fe(cmvn)
    n = 9_600_000
    w = synth.uniform_pcm(101, 0, n)
    feats, fl = fe(dense_batch([w, w[:16000]]), [n, 16000])
    T = 1 + (n - 400) // 160
    assert int(fl[0]) == -(-T // 6) == 10000 and int(fl[1]) == 17
    assert not feats[1, 17:].any()
    for row0 in (0, 5000, 10000 - 40):
        f0 = 6 * row0                                   # first frame of the excerpt = a multiple of lfr_n
        a = f0 * 160
        ex = w[a:a + 6 * 40 * 160 + 400 + 6 * 160]      # 46 rows worth of frames
        ref, rl = wf.frontend_forward([ex], [len(ex)], cmvn=cmvn, **PARAFORMER)
        lo = 1 if row0 > 0 else 0                       # the excerpt's first row replicates ITS first frame
        hi = 38                                         # ... and its last rows replicate its last frame
        if row0 + hi > 10000:
            hi = 10000 - row0
        assert_feats_close(feats[0, row0 + lo:row0 + hi], ref[0, lo:hi], cmvn)
    # the very last rows of the long utterance (right replication) against an excerpt that ends where it ends
    tail_frames = 6 * 30 + (T - 1) % 6 + 1
    a = (T - tail_frames) * 160
    assert (T - tail_frames) % 6 == 0
    ref, rl = wf.frontend_forward([w[a:]], [n - a], cmvn=cmvn, **PARAFORMER)
    k = int(rl[0])
    assert_feats_close(feats[0, 10000 - k + 1:10000], ref[0, 1:k], cmvn)


def test_empty_and_degenerate_batches(cmvn):
    fe = make_fe(cmvn)
    with pytest.raises(RuntimeError, match="empty list of sequences"):      # pad_sequence's error upstream (VF:163)
        fe(torch.zeros(0, 16000, device=DEV), [])
    # a batch made only of utterances shorter than one frame, and one with a single frame exactly
    lens = [399, 320, 400]
    waves = [synth.uniform_pcm(7, i, k) for i, k in enumerate(lens)]
    got, gl = fe(dense_batch(waves), lens)
    ref, rl = wf.frontend_forward(waves, lens, cmvn=cmvn, **PARAFORMER)
    assert np.array_equal(gl.cpu().numpy(), rl) and got.shape == ref.shape
    assert_feats_close(got, ref, cmvn)


def test_digital_silence_and_level_changes(cmvn):
    """Real streams contain digital silence and abrupt level changes: all-zero frames must hit the log floor exactly
    (log(FLT_EPSILON), TA:633), frames that straddle silence / signal boundaries and very quiet segments must match the
    oracle, and nothing may turn into NaN / Inf."""
    rng = np.random.default_rng(3)
    t = np.arange(8000) / 16000.0
    parts = [np.zeros(8000), 0.001 * rng.standard_normal(8000), 0.5 * np.sin(2 * np.pi * (200 + 3000 * t) * t),
             np.zeros(4000), (rng.integers(-3, 4, 8000) / 32768.0), np.zeros(8000)]
    x = np.concatenate(parts).astype(np.float32)
    fe = make_fe(cmvn)
    feats, fl = fe(torch.from_numpy(x)[None].to(DEV), [len(x)])
    ref, rl = wf.frontend_forward([x], [len(x)], cmvn=cmvn, **PARAFORMER)
    got = feats[0].cpu().numpy()
    assert int(fl[0]) == int(rl[0]) and np.isfinite(got).all()
    sh, sc = cmvn[0].astype(np.float64), cmvn[1].astype(np.float64)
    lg = (got.astype(np.float64) / sc - sh).reshape(-1, 7, 80)
    lr = (ref[0].astype(np.float64) / sc - sh).reshape(-1, 7, 80)
    floor = float(np.log(np.finfo(np.float32).eps))
    silent = lr <= floor + 1e-6
    assert silent.any() and np.abs(lg[silent] - floor).max() <= 1e-5      # exact zeros -> the floor on both sides
    # everything else: the stated tolerance, except bins within a hair of the floor where log is ill-conditioned
    ok = ~silent & (lr > floor + 3.0)
    err = np.abs(lg - lr)
    assert err[ok].max() <= 3e-3 and err[ok].mean() <= 2e-5


def test_audio_ring_reproduces_the_reference_sliding_window(cmvn):
    """`buf = np.concatenate([buf, chunk])[-target:]` (R:voice_interface.py:1304-1311) for several
    sessions at once on the device, then the KWS-style front-end call on the window."""
    from toolbox_for_asr_and_tts_b200 import AudioRing
    cap, n_streams = 25600, 5                     # 1.6 s at 16 kHz
    ring = AudioRing(n_streams, cap, DEV)
    rng = np.random.default_rng(11)
    bufs = [np.zeros(0, dtype=np.float32) for _ in range(n_streams)]
    fe = make_fe(cmvn)
    for tick in range(9):
        ids = [s for s in range(n_streams) if (tick + s) % 3 != 0]            # not every session speaks every tick
        lens = [int(rng.integers(1, 6401)) if tick != 4 else 30000 for _ in ids]   # tick 4: a chunk longer than the window
        maxlen = max(lens)
        chunks = np.zeros((len(ids), maxlen), dtype=np.float32)
        for k, (s, n) in enumerate(zip(ids, lens)):
            c = (0.3 * rng.standard_normal(n)).astype(np.float32)
            chunks[k, :n] = c
            bufs[s] = np.concatenate([bufs[s], c])[-cap:]                       # the reference's update
        ring.push(torch.from_numpy(chunks).to(DEV), lens, ids)
        win, wl = ring.window(list(range(n_streams)))
        win, wl = win.cpu().numpy(), wl.cpu().numpy()
        for s in range(n_streams):
            assert wl[s] == len(bufs[s])
            assert np.array_equal(win[s, :wl[s]], bufs[s]) and not win[s, wl[s]:].any()
    ring.reset([1])
    win, wl = ring.window([1, 2])
    assert int(wl[0]) == 0 and not win[0].any() and int(wl[1]) == len(bufs[2])
    # the window goes straight into the front-end, like kws_model.generate(input=self.kws_audio_buffer) at :1370-1374
    win, wl = ring.window([0, 2, 3])
    feats, fl = fe(win, wl)
    ref, rl = wf.frontend_forward([bufs[0], bufs[2], bufs[3]], [len(bufs[0]), len(bufs[2]), len(bufs[3])], cmvn=cmvn, **PARAFORMER)
    assert np.array_equal(fl.cpu().numpy(), rl)
    for i, k in enumerate(rl):
        assert_feats_close(feats[i, :k], ref[i, :k], cmvn)


def test_one_frontend_from_several_host_threads_on_their_own_streams(cmvn):
    """SURVEY 8(b) threading clause: one handle may be shared by host threads that use distinct streams / workspaces.
    Four threads push different ragged batches through ONE WavFrontend concurrently; every result must be bit-identical
    to the same batch run alone."""
    import threading
    fe = make_fe(cmvn)
    rng = np.random.default_rng(77)
    jobs = []
    for t in range(4):
        lens = [int(v) for v in rng.integers(400, 60000, size=6)]
        waves = [synth.uniform_pcm(SEED + t, i, n) for i, n in enumerate(lens)]
        jobs.append((dense_batch(waves), lens))
    alone = []
    for x, lens in jobs:
        f, l = fe(x, lens)
        alone.append((f.clone(), l.clone()))
    torch.cuda.synchronize()
    errors = []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream(device=DEV)
            x, lens = jobs[t]
            st.wait_stream(torch.cuda.default_stream(torch.device(DEV)))
            with torch.cuda.stream(st):
                for _ in range(25):
                    f, l = fe(x, lens)
                    st.synchronize()
                    if not (torch.equal(f, alone[t][0]) and torch.equal(l, alone[t][1])):
                        errors.append(f"thread {t}: result differs from the serial run")
                        return
        except Exception as e:   # surfaced below: an exception in a thread must fail the test
            errors.append(f"thread {t}: {type(e).__name__}: {e}")

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_batches_beyond_the_parameter_table_match_small_batches_bitwise(cmvn):
    """Up to 256 utterances travel in the prep launch's parameters, larger batches through a pinned-buffer copy: the two
    routes must give the same rows.  300 short utterances in one call against the same utterances in calls of 100."""
    fe = make_fe(cmvn)
    rng = np.random.default_rng(5)
    lens = [int(v) for v in rng.integers(300, 9000, size=300)]
    waves = [synth.uniform_pcm(SEED + 9, i, n) for i, n in enumerate(lens)]
    big, big_l = fe(dense_batch(waves), lens)
    for a in range(0, 300, 100):
        part, part_l = fe(dense_batch(waves[a:a + 100]), lens[a:a + 100])
        assert torch.equal(part_l, big_l[a:a + 100])
        k = part.shape[1]
        assert torch.equal(part, big[a:a + 100, :k])
        assert not big[a:a + 100, k:].any()
    ref, rl = wf.frontend_forward(waves[:8], lens[:8], cmvn=cmvn, **dict(PARAFORMER, dither=0.0))
    assert np.array_equal(big_l[:8].cpu().numpy(), rl)
    for i in range(8):
        assert_feats_close(big[i, :rl[i]], ref[i, :rl[i]], cmvn)


def test_forward_is_capturable_in_a_cuda_graph(cmvn):
    """A fixed-shape serving loop captures the call once and replays it (the batch's utterance table travels inside the
    launch parameters, so the capture holds no host -> device copy): replays on new PCM equal the eager call bitwise."""
    fe = make_fe(cmvn)
    lens = [16000, 52000, 9000, 31000, 400, 48000, 7777, 23456]

    def batch(seed):
        return dense_batch([synth.uniform_pcm(seed, i, n) for i, n in enumerate(lens)])

    static_in = batch(1)
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fe(static_in, lens)                      # warm-up outside the capture (handle, kernel attributes)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out, out_lens = fe(static_in, lens)
    for seed in (2, 3):
        x = batch(seed)
        static_in.copy_(x)
        graph.replay()
        torch.cuda.synchronize()
        ref, ref_lens = fe(x, lens)
        assert torch.equal(out, ref) and torch.equal(out_lens, ref_lens)


def test_stream_tick_is_capturable_in_a_cuda_graph(cmvn):
    """The per-tick push of a StreamPool takes device tensors only: captured once, replayed every tick (chunk contents,
    chunk lengths and final flags change in place) it must equal an eagerly driven pool bitwise."""
    fe = make_fe(cmvn)
    n_streams, chunk, ticks = 16, 9600, 5
    rng = np.random.default_rng(3)
    eager = StreamPool(fe, n_streams=n_streams, max_chunk_samples=chunk, device=DEV)
    graphed = StreamPool(fe, n_streams=n_streams, max_chunk_samples=chunk, device=DEV)
    ids = torch.arange(n_streams, dtype=torch.int32, device=DEV)
    s_chunks = torch.zeros(n_streams, chunk, device=DEV)
    s_lens = torch.full((n_streams,), chunk, dtype=torch.int32, device=DEV)
    s_fin = torch.zeros(n_streams, dtype=torch.uint8, device=DEV)
    side = torch.cuda.Stream(device=DEV)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        warm = StreamPool(fe, n_streams=n_streams, max_chunk_samples=chunk, device=DEV)
        warm.push(s_chunks, s_lens, ids, s_fin)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        g_feats, g_rows = graphed.push(s_chunks, s_lens, ids, s_fin)
    for t in range(ticks):
        x = torch.from_numpy(rng.uniform(-0.3, 0.3, (n_streams, chunk)).astype(np.float32)).to(DEV)
        lens = torch.from_numpy(rng.integers(1, chunk + 1, n_streams).astype(np.int32)).to(DEV)
        fin = torch.full((n_streams,), 1 if t == ticks - 1 else 0, dtype=torch.uint8, device=DEV)
        s_chunks.copy_(x); s_lens.copy_(lens); s_fin.copy_(fin)
        graph.replay()
        e_feats, e_rows = eager.push(x, lens, ids, fin)
        torch.cuda.synchronize()
        assert torch.equal(g_rows, e_rows), t
        for s in range(n_streams):
            k = int(e_rows[s])
            assert torch.equal(g_feats[s, :k], e_feats[s, :k]), (t, s)


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: parity on the benchmarked configuration itself, oracle-driven streaming, state snapshots, buffer edges.

def test_bench_batch_every_valid_row_against_the_reference(cmvn):
    """The exact batch bench.py times (BASELINE.json configs[1]: bench.batch_layout(0), 256 utterances of 1-30 s,
    length-packed, synthesised on the device by the same kernel) through forward_packed, every valid row against the
    reference run live on the same samples (torchaudio kaldi.fbank + apply_lfr + apply_cmvn when importable, else the
    numpy restatement) - 412 k frames.  Bins more than 12 nepers below their frame's peak are judged against the float64
    oracle relative to the reference's own float32 noise (conftest.assert_logmel_close)."""
    import bench
    from oracle import ref_thirdparty as ref
    lens, offs, total = bench.batch_layout(0)
    cm = bench.synthetic_cmvn()
    wave = torch.zeros(total + 8, dtype=torch.float32, device=DEV)
    _native.ops().synth_uniform(wave, torch.from_numpy(offs), torch.from_numpy(lens), 0 * 1000 + 1234, 0.3)
    fe = WavFrontend(cmvn=torch.from_numpy(cm), dither=0.0, **bench.CONF)
    feats, flens = fe.forward_packed(wave, torch.from_numpy(offs), torch.from_numpy(lens))
    host = wave.cpu().numpy()
    assert np.array_equal(host[int(offs[3]):int(offs[3]) + 1000], synth.uniform_pcm(1234, 3, 1000))   # same samples
    waves = [host[int(o):int(o) + int(n)] for o, n in zip(offs, lens)]
    t = 1 + (lens - 400) // 160
    assert np.array_equal(flens.cpu().numpy(), -(-t // 6))
    assert tuple(feats.shape) == (256, int((-(-t // 6)).max()), 560)
    sh, sc = cm[0].astype(np.float64), cm[1].astype(np.float64)
    use_ta = ref.have_torchaudio()
    rfe = ref.make_reference_frontend(cm, prefer_vllm=False, **bench.CONF) if use_ta else None
    worst, deep_n, deep_got, deep_ref = 0.0, 0, 0.0, 0.0
    kw = dict(num_mel_bins=80, frame_length=25.0, frame_shift=10.0, dither=0.0, energy_floor=0.0, window_type="hamming",
              sample_frequency=16000.0)
    for u0 in range(0, 256, 16):
        us = list(range(u0, u0 + 16))
        wl = [int(lens[u]) for u in us]
        if use_ta:
            r, rl = rfe([waves[u] for u in us], wl)
            r, rl = r.numpy(), rl.numpy()
        else:
            r, rl = wf.frontend_forward([waves[u] for u in us], wl, cmvn=cm, **bench.CONF)
        got = feats[u0:u0 + 16].cpu().numpy()
        for k, u in enumerate(us):
            n = int(rl[k])
            assert n == int(flens[u]) and not got[k, n:].any()          # pad_sequence zeros, bit-exact
            lg = (got[k, :n].astype(np.float64) / sc - sh).reshape(n, 7, 80)
            lr = (r[k, :n].astype(np.float64) / sc - sh).reshape(n, 7, 80)
            f64 = kf.fbank(waves[u].astype(np.float64) * 32768.0, dtype=np.float64, **kw)
            l64 = wf._lfr_keep_dtype(f64, 7, 6).reshape(n, 7, 80)
            assert_logmel_close(lg, lr, ref64=l64)
            well = (lr.max(axis=-1, keepdims=True) - lr) <= 12.0
            worst = max(worst, float(np.abs(lg - lr)[well].max()))
            if (~well).any():
                deep_n += int((~well).sum())
                deep_got = max(deep_got, float(np.abs(lg - l64)[~well].max()))
                deep_ref = max(deep_ref, float(np.abs(lr - l64)[~well].max()))
    print(f"bench batch, 412 k frames: worst |log-mel - reference| over bins within 12 nepers of their frame's peak {worst:.2e}; "
          f"{deep_n} deeper bins: worst vs float64 {deep_got:.2e} (the float32 reference itself: {deep_ref:.2e})")


@pytest.mark.parametrize("chunk", [9600, 6400, 3840, 960, 300])
def test_streaming_against_the_online_oracle_chunk_by_chunk(cmvn, chunk):
    """oracle.OnlineFrontend (sample carry + LFR splice carry, numpy) and the CUDA StreamPool are driven with the same
    chunks, chunk by chunk: the number of rows every push returns must be equal and so must the rows.  The last chunk
    is shorter than one frame (the final flush then only completes the padded rows)."""
    n = 3 * 9600 + 250          # ends with a chunk (or remainder) shorter than one frame
    w = synth.uniform_pcm(93, chunk, n)
    fe = make_fe(cmvn)
    pool = StreamPool(fe, n_streams=3, max_chunk_samples=9600, device=DEV)
    orc = wf.OnlineFrontend(cmvn=cmvn, **PARAFORMER)
    sid = torch.tensor([1], dtype=torch.int32)
    pos, rows_seen, per_push = 0, 0, []
    while pos < n:
        m = min(chunk, n - pos)
        fin = pos + m >= n
        c = np.zeros((1, 9600), dtype=np.float32)
        c[0, :m] = w[pos:pos + m]
        feats, rows = pool.push(torch.from_numpy(c).to(DEV), torch.tensor([m], dtype=torch.int32), sid,
                                torch.tensor([1 if fin else 0], dtype=torch.uint8))
        want = orc.push(w[pos:pos + m], is_final=fin)
        k = int(rows[0])
        assert k == want.shape[0], (chunk, pos, k, want.shape)
        if k:
            assert_feats_close(feats[0, :k], want, cmvn)
        per_push.append(k)
        rows_seen += k
        pos += m
    t = 1 + (n - 400) // 160
    assert rows_seen == -(-t // 6)
    if chunk == 9600:
        assert per_push == [10, 10, 10, 0]      # 58 + 60 + 60 frames = 30 rows; the 250-sample tail adds a frame? no: 180 frames in all


def test_streaming_fsmn_vad_configuration_against_the_online_oracle(variants_golden):
    """The FSMN-VAD front-end (LFR 5/1, its own CMVN) in streaming form, 400 ms client chunks
    (R:voice-service/app/services/voice_interface.py:648, 1585-1590), against the online oracle per push."""
    cm = variants_golden["vad_5_1_cmvn"]
    conf = dict(PARAFORMER, lfr_m=5, lfr_n=1)
    fe = WavFrontend(cmvn=torch.from_numpy(cm), dither=0.0, **conf)
    pool = StreamPool(fe, n_streams=1, max_chunk_samples=6400, device=DEV)
    orc = wf.OnlineFrontend(cmvn=cm, **conf)
    w = synth.uniform_pcm(94, 0, 6400 * 4 + 123)
    sid = torch.zeros(1, dtype=torch.int32)
    for pos in range(0, len(w), 6400):
        m = min(6400, len(w) - pos)
        fin = pos + m >= len(w)
        c = np.zeros((1, 6400), dtype=np.float32)
        c[0, :m] = w[pos:pos + m]
        feats, rows = pool.push(torch.from_numpy(c).to(DEV), torch.tensor([m], dtype=torch.int32), sid,
                                torch.tensor([1 if fin else 0], dtype=torch.uint8))
        want = orc.push(w[pos:pos + m], is_final=fin)
        assert int(rows[0]) == want.shape[0]
        sh, sc = cm[0].astype(np.float64), cm[1].astype(np.float64)
        lg = (feats[0, :want.shape[0]].cpu().numpy().astype(np.float64) / sc - sh).reshape(-1, 5, 80)
        assert_logmel_close(lg, (want.astype(np.float64) / sc - sh).reshape(-1, 5, 80))


def test_stream_state_snapshot_and_restore(cmvn):
    """a checkpoint of the stream slab (StreamPool.snapshot / restore) taken mid-stream must make the
    continuation reproduce, bit for bit, what the uninterrupted streams produce - including streams that were reset or
    had advanced further in between."""
    fe = make_fe(cmvn)
    pool = StreamPool(fe, n_streams=4, max_chunk_samples=9600, device=DEV)
    ids = torch.arange(4, dtype=torch.int32)
    w = np.stack([synth.uniform_pcm(95, s, 9600 * 4) for s in range(4)])
    lens = torch.full((4,), 9600, dtype=torch.int32)

    def tick(t, final=False):
        fin = torch.full((4,), 1 if final else 0, dtype=torch.uint8)
        f, r = pool.push(torch.from_numpy(w[:, t * 9600:(t + 1) * 9600].copy()).to(DEV), lens, ids, fin)
        return f.clone(), r.clone()

    tick(0)
    tick(1)
    snap = pool.snapshot()
    a2, ra2 = tick(2)
    a3, ra3 = tick(3, final=True)
    pool.reset(torch.tensor([0, 2], dtype=torch.int32))   # disturb the state ...
    tick(0)
    pool.restore(snap)                                     # ... and roll it back
    b2, rb2 = tick(2)
    b3, rb3 = tick(3, final=True)
    assert torch.equal(ra2, rb2) and torch.equal(ra3, rb3)
    for s in range(4):
        assert torch.equal(a2[s, :int(ra2[s])], b2[s, :int(rb2[s])])
        assert torch.equal(a3[s, :int(ra3[s])], b3[s, :int(rb3[s])])


def test_stream_push_with_too_few_output_rows_fails_loudly(cmvn):
    """A push whose output buffer holds fewer rows than one push can emit (b200fe_stream_max_rows) is refused by the C
    ABI with B200FE_E_INVALID and a message - rows are never dropped silently."""
    lib, h = ctypes_handle(cmvn)
    try:
        need = lib.b200fe_stream_max_rows(h, 9600)
        assert need >= 11
        sb = ctypes.c_size_t()
        assert lib.b200fe_stream_state_bytes(h, 1, 9600, ctypes.byref(sb)) == 0
        state = torch.zeros(sb.value, dtype=torch.uint8, device=DEV)
        assert lib.b200fe_stream_reset(h, ctypes.c_void_p(state.data_ptr()), 1, 9600, None, 1, None) == 0
        chunk = torch.from_numpy(synth.uniform_pcm(96, 0, 9600)).to(DEV)
        clen = torch.tensor([9600], dtype=torch.int32, device=DEV)
        sid = torch.zeros(1, dtype=torch.int32, device=DEV)
        rows = torch.full((1,), -1, dtype=torch.int32, device=DEV)

        def push(rows_cap):
            feats = torch.zeros(1, rows_cap, 560, device=DEV)
            rc = lib.b200fe_stream_push(h, ctypes.c_void_p(state.data_ptr()), 1, 9600, ctypes.c_void_p(chunk.data_ptr()),
                                        ctypes.c_int64(9600), ctypes.c_void_p(clen.data_ptr()), ctypes.c_void_p(sid.data_ptr()),
                                        None, 1, ctypes.c_void_p(feats.data_ptr()), ctypes.c_int64(rows_cap),
                                        ctypes.c_void_p(rows.data_ptr()), None)
            torch.cuda.synchronize()
            return rc

        assert push(need - 1) < 0 and b"rows_cap" in lib.b200fe_last_error(h)
        assert int(rows[0]) == -1                       # nothing ran, the state is untouched
        assert push(need) == 0 and int(rows[0]) == 10   # 58 frames -> 10 rows
    finally:
        lib.b200fe_destroy(h)


def test_pcm16_quad_that_fills_the_warp_buffer():
    """with frame_shift 10.25 ms (S = 164, L = 400: 3 S + L = 892) an int16 quad starts up to 7 samples before
    its first frame and stores whole 8-sample vectors, which no longer fits the 896-float warp buffer: int16 input must
    then be refused (never silently overflow), float32 input (3 samples of slack) still takes the warp kernel, and
    both agree with the oracle."""
    conf = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10.25, lfr_m=1, lfr_n=1)
    fe = WavFrontend(dither=0.0, **conf)
    rng = np.random.default_rng(164)
    n = 8000
    ints = rng.integers(-20000, 20000, size=n + 7, dtype=np.int16)
    for lead in (0, 1, 7):
        xi = torch.from_numpy(ints[lead:lead + n].copy())
        xf = xi.to(torch.float32) / 32768.0
        flat = torch.zeros(n + 16, dtype=torch.float32)
        flat[lead:lead + n] = xf
        got, gl = fe.forward_packed(flat.to(DEV), [lead], [n])
        ref = kf.fbank(xf.numpy() * np.float32(32768.0), num_mel_bins=80, frame_length=25.0, frame_shift=10.25, dither=0.0,
                       energy_floor=0.0, window_type="hamming", sample_frequency=16000.0, dtype=np.float32)
        assert int(gl[0]) == ref.shape[0] == 1 + (n - 400) // 164
        assert_logmel_close(got[0, :ref.shape[0]].cpu().numpy(), ref)
        flat16 = torch.zeros(n + 16, dtype=torch.int16)
        flat16[lead:lead + n] = xi
        with pytest.raises(RuntimeError, match="int16 input"):
            fe.forward_packed(flat16.to(DEV), [lead], [n])


def test_online_frontend_hands_back_frame_aligned_waveforms(variants_golden):
    """like upstream's WavFrontendOnline, every call leaves in cache["waveforms"] the raw samples of
    the rows it returned ((k-1)*shift + frame samples for the FSMN-VAD setting lfr_n = 1, starting at the first returned
    row's frame) and keeps the unconsumed tail in cache["reserve_waveforms"]; an empty call returns an empty
    [1, 0, D] feature tensor; feature lengths are int64 on the CPU; nothing is read back from the device."""
    cm = variants_golden["vad_5_1_cmvn"]
    conf = dict(PARAFORMER, lfr_m=5, lfr_n=1)
    fe = WavFrontendOnline(cmvn=torch.from_numpy(cm), max_chunk_samples=6400, dither=0.0, **conf)
    off = WavFrontend(cmvn=torch.from_numpy(cm), dither=0.0, **conf)
    n = 6400 * 3 + 1000
    w = synth.uniform_pcm(97, 0, n)
    full, fl = off(torch.from_numpy(w)[None].to(DEV), [n])
    cache, rows, outs = {}, 0, []
    first = torch.from_numpy(w[:300])[None].to(DEV)                      # shorter than one frame: nothing yet
    f, l = fe(first, [300], cache=cache, is_final=False)
    assert tuple(f.shape) == (1, 0, 400) and f.is_cuda and l.dtype == torch.int64 and l.device.type == "cpu" and int(l[0]) == 0
    assert cache["waveforms"].shape[1] == 0 and cache["reserve_waveforms"].shape[1] == 300
    pos = 300
    while pos < n:
        m = min(6400, n - pos)
        fin = pos + m >= n
        f, l = fe(torch.from_numpy(w[pos:pos + m])[None].to(DEV), [m], cache=cache, is_final=fin)
        k = int(l[0])
        assert f.shape[1] == k
        wav = cache["waveforms"]
        if k:
            # rows [rows, rows + k) <-> frames rows .. rows + k - 1 <-> samples [rows*160, (rows + k - 1)*160 + 400)
            a, b = rows * 160, min(n, (rows + k - 1) * 160 + 400)
            assert wav.shape == (1, b - a) and np.array_equal(wav[0].cpu().numpy(), w[a:b])
            outs.append(f[0])
        rows += k
        pos += m
        if not fin:
            assert np.array_equal(cache["reserve_waveforms"][0].cpu().numpy(), w[rows * 160:pos])
    assert rows == int(fl[0]) and cache["reserve_waveforms"].shape[1] == 0
    got = torch.cat(outs)
    sh, sc = cm[0].astype(np.float64), cm[1].astype(np.float64)
    assert_logmel_close((got.cpu().numpy().astype(np.float64) / sc - sh).reshape(-1, 5, 80),
                        (full[0].cpu().numpy().astype(np.float64) / sc - sh).reshape(-1, 5, 80))


def test_host_ingest_from_numpy_list_and_dense_tensor_equals_device_resident_forward(cmvn):
    """HostIngest: the utterances as a reference caller holds them (a list of separate np.float32 arrays, or the dense
    [B, Nmax] host tensor funasr pads them into) -> multi-threaded gather into pinned staging -> H2D -> fused kernels.
    Must equal, bit for bit, forward() on the same data already on the device; double-buffering must survive many calls
    with changing batches; int16 PCM takes the same route."""
    from toolbox_for_asr_and_tts_b200.ingest import HostIngest
    fe = make_fe(cmvn)
    ing = HostIngest(fe, capacity_samples=400000, device=DEV, groups=3, threads=4)
    rng = np.random.default_rng(12)
    for it in range(5):
        lens = [int(n) for n in rng.integers(300, 60000, size=int(rng.integers(1, 7)))]
        waves = [synth.uniform_pcm(200 + it, i, n) for i, n in enumerate(lens)]
        ref, rl = fe(dense_batch(waves), lens)
        got, gl = ing.forward(waves)
        assert torch.equal(gl.cpu(), rl) and torch.equal(got, ref), it
        dense = torch.zeros(len(lens), max(lens))
        for i, w in enumerate(waves):
            dense[i, :len(w)] = torch.from_numpy(w)
        got2, gl2 = ing.forward(dense, lens)
        assert torch.equal(gl2.cpu(), rl) and torch.equal(got2, ref), it
    ing16 = HostIngest(fe, capacity_samples=400000, dtype=torch.int16, device=DEV)
    ints = [rng.integers(-20000, 20000, size=n, dtype=np.int16) for n in (16000, 401, 7777)]
    ref, rl = fe(dense_batch([w.astype(np.float32) / np.float32(32768.0) for w in ints]), [16000, 401, 7777])
    got, gl = ing16.forward(ints)
    assert torch.equal(gl.cpu(), rl) and torch.equal(got, ref)
    with pytest.raises(ValueError, match="capacity"):
        ing.forward([np.zeros(500000, dtype=np.float32)])


def test_stream_tick_returns_the_reference_energy_gate(cmvn):
    """the per-chunk energy gate of StreamingASRSession.process_chunk (R:voice_interface.py:1569-1578:
    is_speech = mean|x| > 0.03 and max|x| > 0.17, numpy on the host) comes back from the same launch as the rows."""
    fe = make_fe(cmvn)
    pool = StreamPool(fe, n_streams=6, max_chunk_samples=6400, device=DEV)
    rng = np.random.default_rng(41)
    lens = [6400, 3840, 6400, 100, 0, 6400]
    amps = [0.3, 0.02, 0.05, 0.5, 0.0, 0.001]
    chunks = np.zeros((6, 6400), dtype=np.float32)
    for i, (n, a) in enumerate(zip(lens, amps)):
        chunks[i, :n] = a * rng.standard_normal(n).clip(-3, 3).astype(np.float32)
    chunks[2, 17] = 0.9                       # a loud click in a quiet chunk
    ids = torch.arange(6, dtype=torch.int32)
    plain_f, plain_r = pool.push(torch.from_numpy(chunks).to(DEV), torch.tensor(lens, dtype=torch.int32), ids)
    pool.reset()
    feats, rows, flags, stats = pool.push_with_speech_flags(torch.from_numpy(chunks).to(DEV),
                                                            torch.tensor(lens, dtype=torch.int32), ids)
    assert torch.equal(rows, plain_r)
    for i in range(6):                                     # rows beyond a stream's count are not written
        assert torch.equal(feats[i, :int(rows[i])], plain_f[i, :int(rows[i])]), i
    st = stats.cpu().numpy()
    for i, n in enumerate(lens):
        x = chunks[i, :n]
        energy = float(np.mean(np.abs(x))) if n else 0.0            # :1569
        peak = float(np.max(np.abs(x))) if n else 0.0               # :1570
        assert abs(st[i, 0] - energy) <= 1e-5 * max(energy, 1e-3) and st[i, 1] == np.float32(peak), i
        assert bool(flags[i]) == (energy > 0.03 and peak > 0.17), i  # STREAMING_VAD_USE_AND_LOGIC = True (:658)
    _, _, or_flags, _ = pool.push_with_speech_flags(torch.from_numpy(chunks).to(DEV), torch.tensor(lens, dtype=torch.int32),
                                                    ids, use_and_logic=False)
    assert or_flags.cpu().tolist() == [True, False, True, True, False, False]


@pytest.mark.parametrize("chunk", [9600, 3840, 300])
def test_quad_level_stream_kernel_matches_the_shipped_tick(cmvn, chunk, monkeypatch):
    """The quad-level alternative of the streaming tick (B200FE_STREAM_KERNEL=quad: descriptors, work items = (chunk, quad)
    on persistent warps that scatter their frames straight into the rows, state update) is kept as a measured
    alternative.  Same rows, counts, energy gate and state as the shipped one-CTA-per-stream tick,
    push by push, incl. a stream that starts later, a final flush and chunks shorter than one frame."""
    fe = make_fe(cmvn)
    n_streams, max_chunk = 5, 9600
    pools = {}
    for kind in ("cta", "quad"):
        monkeypatch.setenv("B200FE_STREAM_KERNEL", kind)
        pools[kind] = StreamPool(fe, n_streams=n_streams, max_chunk_samples=max_chunk, device=DEV)
    n = 3 * 9600 + 250
    waves = [synth.uniform_pcm(97, i, n) for i in range(3)]
    ids = torch.tensor([4, 0, 2], dtype=torch.int32)
    pos, tick = 0, 0
    while pos < n:
        m = min(chunk, n - pos)
        fin = pos + m >= n
        c = np.zeros((3, max_chunk), dtype=np.float32)
        lens = [m, m if tick >= 1 else 0, max(m - 7, 0)]            # stream 0 starts one tick late, stream 2 runs ragged
        for i in range(3):
            c[i, :lens[i]] = waves[i][pos:pos + lens[i]]
        out = {}
        for kind in ("cta", "quad"):
            monkeypatch.setenv("B200FE_STREAM_KERNEL", kind)
            out[kind] = pools[kind].push_with_speech_flags(torch.from_numpy(c).to(DEV), torch.tensor(lens, dtype=torch.int32), ids,
                                                           torch.tensor([1 if fin else 0] * 3, dtype=torch.uint8))
            torch.cuda.synchronize()
        (fa, ra, ga, sa), (fb, rb, gb, sb) = out["cta"], out["quad"]
        assert torch.equal(ra, rb), (pos, ra, rb)
        assert torch.equal(ga, gb) and torch.allclose(sa, sb, rtol=1e-5, atol=0) and torch.equal(sa[:, 1], sb[:, 1])
        for i in range(3):
            k = int(ra[i])
            assert torch.allclose(fa[i, :k], fb[i, :k], rtol=0, atol=2e-4), (pos, i, float((fa[i, :k] - fb[i, :k]).abs().max()))
        pos += m
        tick += 1
    # the state slabs agree where they carry state: counters, sample carry (bit-exact), splice frames (rounding noise)
    a, b = pools["cta"].state, pools["quad"].state
    nbytes_counters = 4 * n_streams * 4
    assert torch.equal(a[:nbytes_counters], b[:nbytes_counters])


def test_ingest_with_scipy_fourier_resampling_against_the_scipy_golden():
    """the resampling branch the reference takes when scipy is installed (scipy.signal.resample,
    R:voice_interface.py:1022-1027) on the GPU, against outputs of scipy itself (tests/golden/resample_golden.npz):
    240 / 400 ms chunks at 48 / 44.1 / 8 kHz, odd lengths, stereo, all three sample widths.  The float64 DFT sums agree
    with pocketfft to round-off, so after the float32 cast at most a handful of samples may differ, by one ulp."""
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    spec = importlib.util.spec_from_file_location("make_golden_resample", root / "tests" / "golden" / "make_golden_resample.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = dict(np.load(root / "tests" / "golden" / "resample_golden.npz"))
    for name, (dtype, ch, rate, n) in mod.CASES.items():
        raw = mod.wire_pcm(name)
        t = torch.from_numpy(raw).to(DEV)
        got = WavFrontend.ingest_pcm(t, channels=ch, src_rate=rate, dst_rate=16000, method="scipy").cpu().numpy()
        ref = g[name]
        assert got.dtype == np.float32 and got.shape == ref.shape, name
        ulp = np.spacing(np.abs(ref).astype(np.float32))
        assert (np.abs(got - ref) <= ulp).all(), (name, float(np.abs(got - ref).max()))
        assert (got != ref).mean() <= 1e-3, (name, float((got != ref).mean()))
    # equal rates: no resampling on either branch
    raw = np.arange(-50, 50, dtype=np.int16)
    a = WavFrontend.ingest_pcm(torch.from_numpy(raw).to(DEV), 1, 16000, 16000, method="scipy").cpu().numpy()
    assert np.array_equal(a, (raw / 32768.0).astype(np.float32))


def test_kws_sliding_window_through_the_charctc_preset():
    """the wake-word path.  The reference keeps the newest 1.6 s of audio per session
    (kws_audio_buffer, R:voice_interface.py:1304-1311) and hands it to the CharCTC-KWS model on every chunk (:1370-1374),
    whose front-end stacks 5 frames every 3 (presets.CHARCTC_KWS).  Here: device-resident AudioRing windows of several
    sessions -> one batched forward, against numpy's sliding window + the oracle."""
    from toolbox_for_asr_and_tts_b200 import AudioRing, presets
    conf = presets.CHARCTC_KWS
    fe = WavFrontend(**conf)
    ring = AudioRing(n_streams=4, capacity_samples=presets.KWS_WINDOW_SAMPLES, device=DEV)
    rng = np.random.default_rng(77)
    host = [np.zeros(0, dtype=np.float32) for _ in range(4)]
    ids = [0, 1, 3]
    ring.reset([0, 1, 2, 3])
    for tick in range(9):                                   # 240 ms chunks (R:voice_interface.py:648): 2.16 s in all
        chunks = np.stack([0.2 * rng.standard_normal(3840).astype(np.float32) for _ in ids])
        ring.push(torch.from_numpy(chunks).to(DEV), [3840] * len(ids), ids)
        for k, s in enumerate(ids):
            host[s] = np.concatenate([host[s], chunks[k]])[-presets.KWS_WINDOW_SAMPLES:]      # :1304-1311
        win, lens = ring.window(ids)
        assert lens.cpu().tolist() == [len(host[s]) for s in ids]
        feats, fl = fe(win, lens)
        oconf = {k: v for k, v in conf.items() if k != "dither"}
        ref, rl = wf.frontend_forward([host[s] for s in ids], [len(host[s]) for s in ids], cmvn=None, **oconf)
        assert np.array_equal(fl.numpy(), rl) and feats.shape[-1] == 400
        for k in range(len(ids)):
            n = int(rl[k])
            assert_logmel_close(feats[k, :n].cpu().numpy().reshape(n, 5, 80), ref[k, :n].reshape(n, 5, 80))
            assert not feats[k, n:].any()


def test_rows_packed_output_equals_the_padded_output(cmvn):
    """forward_packed(..., pad=False): [sum of rows, D] + row offsets instead of the zero-padded [B, max rows, D]
    (B200FE_ROWS_PACKED in the C ABI).  Same rows, bit for bit, including an utterance shorter than one frame; options the
    warp kernel does not cover refuse the layout."""
    fe = make_fe(cmvn)
    lens = np.array([4001, 16003, 399, 7777, 480000, 401, 1601], dtype=np.int64)
    waves = [synth.uniform_pcm(33, i, int(n)) for i, n in enumerate(lens)]
    offs, total = synth.packed_offsets(lens, align=4)
    flat = torch.zeros(int(total) + 8)
    for o, w in zip(offs, waves):
        flat[o:o + len(w)] = torch.from_numpy(w)
    padded, pl = fe.forward_packed(flat.to(DEV), offs, lens)
    rows, rl, ro = fe.forward_packed(flat.to(DEV), offs, lens, pad=False)
    assert torch.equal(rl, pl) and ro.dtype == torch.int64 and ro.device.type == "cpu"
    assert rows.shape == (int(pl.sum()), 560) and int(ro[-1]) == rows.shape[0]
    for i in range(len(lens)):
        assert torch.equal(rows[int(ro[i]):int(ro[i + 1])], padded[i, :int(pl[i])]), i
    with pytest.raises(RuntimeError, match="rows-packed"):
        fe.forward_packed(flat.to(DEV), offs, lens, pad=False, stats=torch.zeros(1121, dtype=torch.float64, device=DEV))
