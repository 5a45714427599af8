"""CPU suite, part 1: the oracle itself.  It is pinned against (a) the committed golden vectors, which were produced
by the actual third-party code of the reference's path (tests/golden/make_golden.py), and (b) that code run live when
torchaudio is importable."""
import math

import numpy as np
import pytest
from conftest import LOGMEL_ATOL, PARAFORMER, VARIANT_CONFS, VARIANT_LENS, VARIANT_SEED, assert_logmel_close
from hypothesis import given, settings
from hypothesis import strategies as st

from pathlib import Path

from oracle import kaldi_fbank_np as kf
from oracle import ref_thirdparty as ref
from oracle import wav_frontend_np as wf
from toolbox_for_asr_and_tts_b200 import synth

ROOT = Path(__file__).resolve().parents[1]

SEED = 1234


def cmvn_atol(cmvn):
    return LOGMEL_ATOL * float(np.abs(cmvn[1]).max())


@pytest.mark.parametrize("n", [399, 400, 401, 559, 560, 1000, 16000, 160000, 480000])
def test_oracle_matches_golden_paraformer(golden, cmvn, n):
    x = synth.uniform_pcm(SEED, n, n)
    feats, lens = wf.frontend_forward([x], [n], cmvn=cmvn, **PARAFORMER)
    g = golden[f"paraformer_{n}"]
    assert feats.shape[1:] == g.shape and int(lens[0]) == g.shape[0]
    assert np.abs(feats[0] - g).max() <= cmvn_atol(cmvn)


@pytest.mark.parametrize("n", [400, 16000])
def test_oracle_matches_golden_povey(golden, cmvn, n):
    x = synth.uniform_pcm(SEED, n, n)
    feats, _ = wf.frontend_forward([x], [n], cmvn=cmvn, **dict(PARAFORMER, window="povey"))
    assert np.abs(feats[0] - golden[f"povey_{n}"]).max() <= cmvn_atol(cmvn)


def test_oracle_matches_golden_fbank_and_batch(golden, cmvn):
    x = synth.uniform_pcm(SEED, 16000, 16000)
    feats, _ = wf.frontend_forward([x], [16000], cmvn=None, fs=16000, window="hamming", n_mels=80)
    assert np.abs(feats[0] - golden["fbank_16000"]).max() <= LOGMEL_ATOL
    lens = golden["batch_input_lens"]
    waves = [synth.uniform_pcm(SEED + 1, i, int(n)) for i, n in enumerate(lens)]
    feats, flens = wf.frontend_forward(waves, lens, cmvn=cmvn, **PARAFORMER)
    assert np.array_equal(flens, golden["batch_lens"]) and flens.dtype == np.int64
    assert feats.shape == golden["batch_feats"].shape
    assert np.abs(feats - golden["batch_feats"]).max() <= cmvn_atol(cmvn)
    # pad_sequence zero padding, bit-exact
    for i, k in enumerate(flens):
        assert not feats[i, k:].any() and not golden["batch_feats"][i, k:].any()


def test_oracle_matches_golden_gaussian_with_dc(golden, cmvn):
    feats, _ = wf.frontend_forward([golden["gauss_input"]], [24000], cmvn=cmvn, **PARAFORMER)
    assert np.abs(feats[0] - golden["gauss_feats"]).max() <= cmvn_atol(cmvn)


@pytest.mark.skipif(not ref.have_torchaudio(), reason="torchaudio not importable")
@pytest.mark.parametrize("window", ["hamming", "povey", "hanning", "rectangular", "blackman"])
def test_oracle_tables_match_torchaudio(window):
    import torch
    import torchaudio.compliance.kaldi as kaldi
    w = kaldi._feature_window_function(window, 400, 0.42, torch.device("cpu"), torch.float32).numpy()
    assert np.abs(kf.window_function(window, 400) - w).max() <= 5e-7  # torch builds the table in float32
    bank, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    mine = kf.mel_banks(80, 512, 16000.0)
    assert mine.shape == tuple(bank.shape)
    # the reference builds the bank in float32 (mel values ~2e3, ulp 1.2e-4): its own noise vs float64 is ~1.4e-5
    assert np.abs(mine - bank.numpy()).max() <= 3e-5
    assert np.abs(kf.mel_banks(80, 512, 16000.0, dtype=np.float64) - bank.numpy()).max() <= 3e-5
    flips = (mine > 0) != (bank.numpy() > 0)
    assert not flips.any() or np.abs(mine - bank.numpy())[flips].max() < 3e-5


@pytest.mark.skipif(not ref.have_torchaudio(), reason="torchaudio not importable")
def test_oracle_live_against_torchaudio_and_float64(cmvn):
    """float32 oracle vs the live reference, and both vs the float64 oracle (the reference's own noise floor)."""
    n = 48000
    x = synth.uniform_pcm(99, 3, n)
    feats_ref, lens_ref, _ = ref.reference_forward([x], [n], cmvn=cmvn, prefer_vllm=False, **PARAFORMER)
    f32, lens = wf.frontend_forward([x], [n], cmvn=cmvn, **PARAFORMER)
    f64, _ = wf.frontend_forward([x], [n], cmvn=cmvn.astype(np.float64), dtype=np.float64, **PARAFORMER)
    assert np.array_equal(lens, lens_ref)
    assert np.abs(f32 - feats_ref).max() <= cmvn_atol(cmvn)
    assert np.abs(feats_ref - f64).max() <= cmvn_atol(cmvn)
    assert np.abs(f32 - f64).max() <= cmvn_atol(cmvn)


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=400, max_value=500000))
def test_frame_and_row_counts(n):
    t = kf.frame_count(n, 400, 160)
    assert t == 1 + (n - 400) // 160
    assert wf.lfr_num_rows(t, 6) == math.ceil(t / 6)


@pytest.mark.parametrize("m,n", [(7, 6), (5, 1), (1, 1), (3, 2), (8, 3)])
def test_lfr_closed_form_equals_literal(m, n):
    rng = np.random.default_rng(0)
    for t in list(range(1, 40)) + [97, 200, 399]:
        x = rng.standard_normal((t, 4)).astype(np.float32)
        assert np.array_equal(wf.apply_lfr(x, m, n), wf.apply_lfr_literal(x, m, n)), (t, m, n)


@pytest.mark.skipif(ref.vllm_wavfrontend_cls() is None, reason="vllm's funasr copy not importable")
def test_lfr_and_cmvn_equal_upstream_functions(cmvn):
    import torch
    from vllm.transformers_utils.processors.funasr import apply_cmvn, apply_lfr
    rng = np.random.default_rng(1)
    for t in (1, 2, 5, 6, 7, 13, 98, 998):
        x = rng.standard_normal((t, 80)).astype(np.float32)
        up = apply_lfr(torch.from_numpy(x), 7, 6).numpy()
        assert np.array_equal(wf.apply_lfr(x, 7, 6), up)
        up2 = apply_cmvn(torch.from_numpy(up.copy()), torch.from_numpy(cmvn)).numpy()
        assert np.array_equal(wf.apply_cmvn(up, cmvn), up2)


def test_cmvn_file_roundtrip(tmp_path, cmvn):
    p = tmp_path / "am.mvn"
    wf.write_cmvn(str(p), cmvn[0], cmvn[1])
    assert np.array_equal(wf.load_cmvn(str(p)), cmvn)
    from toolbox_for_asr_and_tts_b200 import load_cmvn, write_cmvn
    assert np.array_equal(load_cmvn(str(p)).numpy(), cmvn)
    q = tmp_path / "am2.mvn"
    write_cmvn(str(q), cmvn[0], cmvn[1])
    assert np.array_equal(wf.load_cmvn(str(q)), cmvn)
    cls = ref.vllm_wavfrontend_cls()
    if cls is not None:
        from vllm.transformers_utils.processors.funasr import load_cmvn as up_load
        assert np.array_equal(up_load(str(p)).numpy(), cmvn)


@pytest.mark.parametrize("n,chunk", [(160000, 9600), (100003, 9600), (50000, 3840), (48000, 6400), (31999, 9600),
                                     (9600, 9600), (2000, 9600), (700, 300), (16000, 960)])
def test_streaming_oracle_concat_equals_offline(cmvn, n, chunk):
    x = synth.uniform_pcm(5, n, n)
    off, lens = wf.frontend_forward([x], [n], cmvn=cmvn, **PARAFORMER)
    fe = wf.OnlineFrontend(cmvn=cmvn, **PARAFORMER)
    outs, per_chunk = [], []
    for s in range(0, n, chunk):
        r = fe.push(x[s:s + chunk], is_final=(s + chunk >= n))
        outs.append(r)
        per_chunk.append(r.shape[0])
    cat = np.concatenate(outs, axis=0)
    assert cat.shape[0] == int(lens[0])
    assert np.abs(cat - off[0]).max() <= 2e-5
    if (n, chunk) == (160000, 9600):
        assert per_chunk[:3] == [10, 10, 10] and per_chunk[-1] == 7 and sum(per_chunk) == 167


def test_cmvn_stats_recover_table():
    rng = np.random.default_rng(2)
    mats = [rng.standard_normal((k, 12)).astype(np.float32) * 3 + 1.5 for k in (5, 40, 13)]
    s, s2, n = wf.cmvn_stats(mats)
    allm = np.concatenate(mats).astype(np.float64)
    assert n == allm.shape[0]
    tab = wf.stats_to_cmvn(s, s2, n)
    assert np.allclose(tab[0], -allm.mean(0), atol=1e-6) and np.allclose(tab[1], 1 / allm.std(0), atol=1e-6)


def test_tts_mel_definition_against_torch():
    """The frozen TTS definition (parity unpinned by the reference) agrees with torch.stft + torchaudio filters."""
    torch = pytest.importorskip("torch")
    from oracle import tts_mel_np as tm
    try:
        from torchaudio.functional import melscale_fbanks
    except Exception:
        pytest.skip("torchaudio not importable")
    for n in (512, 1025, 1280, 24000):
        x = synth.uniform_pcm(3, n, n)
        mine = tm.tts_log_mel(x)
        assert mine.shape == (80, n // 256)
        xt = torch.from_numpy(x)
        xp = torch.nn.functional.pad(xt[None, None], (384, 384), mode="reflect")[0, 0]
        spec = torch.stft(xp, 1024, hop_length=256, win_length=1024, window=torch.hann_window(1024), center=False,
                          return_complex=True)
        mag = torch.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)
        fb = melscale_fbanks(513, 0.0, 12000.0, 80, 24000, norm="slaney", mel_scale="slaney")
        ref_mel = torch.log(torch.clamp(fb.T @ mag, min=1e-5)).numpy()
        assert np.abs(mine - ref_mel).max() <= 2e-4


@pytest.mark.parametrize("name", sorted(VARIANT_CONFS))
def test_oracle_matches_golden_kaldi_variants(variants_golden, name):
    """The restatement against torchaudio's own output for option sets away from the Paraformer values
    (tests/golden/make_golden_variants.py): frame lengths / shifts, mel counts, windows, no pre-emphasis / DC removal,
    band limits, subtract_mean."""
    o = VARIANT_CONFS[name]
    kw = dict(num_mel_bins=o["n_mels"], frame_length=float(o["frame_length"]), frame_shift=float(o["frame_shift"]),
              dither=0.0, energy_floor=0.0, window_type=o["window"], sample_frequency=16000.0,
              preemphasis_coefficient=o.get("preemphasis_coefficient", 0.97), remove_dc_offset=o.get("remove_dc_offset", True),
              low_freq=o.get("low_freq", 20.0), high_freq=o.get("high_freq", 0.0), dtype=np.float32)
    for i, n in enumerate(VARIANT_LENS):
        x = synth.uniform_pcm(VARIANT_SEED, i, n)
        got = kf.fbank(x * np.float32(32768.0), **kw)
        if o.get("subtract_mean"):
            got = wf.subtract_column_mean(got)
        assert_logmel_close(got, variants_golden[f"{name}_{n}"])


def test_oracle_matches_golden_fsmn_vad_frontend(variants_golden):
    """LFR 5/1 + CMVN (the FSMN-VAD front-end) against the verbatim funasr WavFrontend's output."""
    cm = variants_golden["vad_5_1_cmvn"]
    w = synth.uniform_pcm(63, 0, 48000)
    feats, lens = wf.frontend_forward([w], [48000], cmvn=cm, **dict(PARAFORMER, lfr_m=5, lfr_n=1))
    g = variants_golden["vad_5_1_feats"]
    assert int(lens[0]) == g.shape[0] == 298 and feats.shape[1:] == g.shape
    assert np.abs(feats[0] - g).max() <= cmvn_atol(cm)


def test_oracle_fourier_resampling_matches_the_scipy_golden():
    """The oracle's restatement of scipy.signal.resample (the reference's resampler when scipy is installed,
    R:voice_interface.py:1022-1027) against outputs of scipy itself (tests/golden/make_golden_resample.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_resample", ROOT / "tests" / "golden" / "make_golden_resample.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = dict(np.load(ROOT / "tests" / "golden" / "resample_golden.npz"))
    for name, (dtype, ch, rate, n) in mod.CASES.items():
        got = wf.ingest_pcm(mod.wire_pcm(name), ch, rate, 16000, method="scipy")
        assert got.dtype == np.float32 and got.shape == g[name].shape, name
        assert np.abs(got.astype(np.float64) - g[name]).max() <= 2e-7, name       # float64 round-off before the cast
        assert (got != g[name]).mean() <= 1e-3, name
