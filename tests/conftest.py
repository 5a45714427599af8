import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden" / "frontend_golden.npz"
VARIANTS_GOLDEN = ROOT / "tests" / "golden" / "variants_golden.npz"
# the Kaldi option sets of tests/golden/make_golden_variants.py, in WavFrontend keywords
VARIANT_CONFS = {
    "povey_32ms_40mel": dict(frame_length=32, frame_shift=10, n_mels=40, window="povey"),
    "hanning_20ms_64mel_raw": dict(frame_length=20, frame_shift=10, n_mels=64, window="hanning",
                                   preemphasis_coefficient=0.0, remove_dc_offset=False),
    "blackman_band_100_7800": dict(frame_length=25, frame_shift=10, n_mels=80, window="blackman", low_freq=100.0,
                                   high_freq=-200.0),
    "rectangular_shift5": dict(frame_length=25, frame_shift=5, n_mels=80, window="rectangular"),
    "hamming_shift20_24mel": dict(frame_length=25, frame_shift=20, n_mels=24, window="hamming"),
    "subtract_mean_80mel": dict(frame_length=25, frame_shift=10, n_mels=80, window="hamming", subtract_mean=True),
}
VARIANT_SEED, VARIANT_LENS = 55, (16000, 4001)
PARAFORMER = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6)

# Stated tolerances.
#   * log-mel vs the float32 reference: max-abs <= 1e-3, flat, for every bin that is not ill-conditioned; mean <= 2e-5.
#   * ill-conditioned bins ("pure tones, near-silence; oracle fp32 vs fp64 differs up to 2.7e-2"):
#     bins more than 12 nepers (52 dB) below the frame's strongest mel bin.  There a float32 FFT - the reference's own
#     included - only carries the bin to its float32 noise floor, so a flat bound cannot be the statement.  Two forms:
#       - broadband inputs (all golden vectors): a flat 3e-3 cap, nothing else;
#       - inputs with deep spectral nulls (tones, silence next to signal, the 412 k-frame bench batch whose rarest bins
#         reach 20 nepers): pass the float64 oracle as `ref64`; the CUDA path's error against float64 over those bins
#         must stay within a small multiple of the float32 REFERENCE's own error against float64 over the same bins
#         (rms <= 2x + 1e-4, max <= 4x + 3e-3): "as exact as the reference is determined", no depth formula.
#     Bins within 1e-3 of the log floor log(FLT_EPSILON) are excluded, as BASELINE.md states.
LOGMEL_ATOL = 1e-3
LOGMEL_ATOL_DEEP = 3e-3
LOGMEL_MEAN_ATOL = 2e-5
DEEP_BIN_NEPERS = 12.0
LOG_FLOOR = float(np.log(np.finfo(np.float32).eps))


def assert_logmel_close(got, ref, ref64=None):
    """got/ref: [..., n_mels] natural-log mel energies (no CMVN); ref = float32 reference, ref64 = float64 oracle."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape
    err = np.abs(got - ref)
    depth = ref.max(axis=-1, keepdims=True) - ref
    floor = ref <= LOG_FLOOR + 1e-3
    deep = (depth > DEEP_BIN_NEPERS) & ~floor
    well = ~deep & ~floor
    if well.any():
        assert err[well].max() <= LOGMEL_ATOL, float(err[well].max())
        assert err[well].mean() <= LOGMEL_MEAN_ATOL, float(err[well].mean())
    if floor.any():    # excluded from the comparison; sanity only: nothing far above the floor
        assert (got[floor] >= LOG_FLOOR - 1e-6).all() and (got[floor] <= LOG_FLOOR + 3.0).all()
    if not deep.any():
        return
    if ref64 is None:
        assert err[deep].max() <= LOGMEL_ATOL_DEEP, float(err[deep].max())
        return
    ref64 = np.asarray(ref64, dtype=np.float64)
    e_got, e_ref = np.abs(got - ref64)[deep], np.abs(ref - ref64)[deep]
    rms_got, rms_ref = float(np.sqrt((e_got ** 2).mean())), float(np.sqrt((e_ref ** 2).mean()))
    assert rms_got <= 2.0 * rms_ref + 1e-4, (rms_got, rms_ref)
    assert e_got.max() <= 4.0 * e_ref.max() + LOGMEL_ATOL_DEEP, (float(e_got.max()), float(e_ref.max()))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN))


@pytest.fixture(scope="session")
def variants_golden():
    return dict(np.load(VARIANTS_GOLDEN))


@pytest.fixture(scope="session")
def cmvn(golden):
    return golden["cmvn"]
