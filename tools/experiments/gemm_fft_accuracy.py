"""Accuracy of the tensor-core formulation of the 512-point frame DFT, emulated exactly on the CPU.

The two-level GEMM-FFT a tcgen05 kernel would run (n = 16 i + j;  stage A: 32-point real DFT over i as a GEMM
[16 (j) x 32 (i)] x [32 x 32 (cos | sin of k1 = 0..16)], twiddle W512^(j k1), stage B: 16-point complex DFT over j as a
GEMM [17 (k1) x 32 (re | im of j)] x [32 x 32]) with the operand splittings that make fp16 / bf16 tensor cores usable for
float32 data - x = hi + lo (+ lo2), products hi*hi + hi*lo + lo*hi (+ ...) - and float32 accumulation.  Tensor cores
multiply exactly and accumulate in float32, so rounding the split operands with numpy and accumulating every K = 16
slice in float64 before a float32 rounding reproduces their arithmetic up to the accumulation order.

Prints the max / mean |log-mel error| against the float64 oracle next to the float32 FFT the CUDA-core kernel runs, on
the inputs of tests/parity_report.py.

    python tools/experiments/gemm_fft_accuracy.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import kaldi_fbank_np as kf  # noqa: E402
from toolbox_for_asr_and_tts_b200 import synth  # noqa: E402

f32, f64 = np.float32, np.float64


def to_bf16(x):
    u = np.asarray(x, dtype=f32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000            # round to nearest even
    return u.astype(np.uint32).view(f32)


def split(x, kind):
    """x (float32 / float64) -> list of low-precision terms (as float64 arrays holding exactly representable values)."""
    x = np.asarray(x, dtype=f64)
    terms, rest = [], x.copy()
    n_terms = 2 if kind == "fp16x2" else 3
    for _ in range(n_terms):
        t = rest.astype(np.float16).astype(f64) if kind == "fp16x2" else to_bf16(rest.astype(f32)).astype(f64)
        terms.append(t)
        rest = rest - t
    return terms


def mma(a_terms, b_terms, max_order):
    """sum over term pairs (p, q) with p + q <= max_order of A_p @ B_q, float32 accumulation per K = 16 slice."""
    m, k = a_terms[0].shape[-2:]
    acc = np.zeros(a_terms[0].shape[:-1] + (b_terms[0].shape[-1],), dtype=f32)
    for p, a in enumerate(a_terms):
        for q, b in enumerate(b_terms):
            if p + q > max_order:
                continue
            for k0 in range(0, k, 16):
                acc = (acc.astype(f64) + a[..., k0:k0 + 16] @ b[k0:k0 + 16]).astype(f32)
    return acc


def gemm_fft_power(z, kind):
    """z: [T, 512] float32 windowed frames -> |X[k]|^2 for k = 0..255 through the split-precision two-level GEMM-FFT."""
    T = z.shape[0]
    order = 1 if kind == "fp16x2" else 2
    # fp16 has no exponent range to spare: the 2^15 upscale is taken out before the GEMMs and stage B runs on 2^-5 of
    # stage A's result (powers of two: exact), put back at the end
    pre, mid = (2.0 ** -15, 2.0 ** -5) if kind == "fp16x2" else (1.0, 1.0)
    x = (z.astype(f64) * pre).astype(f32).reshape(T, 32, 16).transpose(0, 2, 1)       # [T, j, i]
    i = np.arange(32)
    cols = []
    for k1 in range(17):                                                # 32 real columns: cos k1 = 0..16, sin k1 = 1..15
        cols.append(np.cos(2 * np.pi * i * k1 / 32))
    for k1 in range(1, 16):
        cols.append(-np.sin(2 * np.pi * i * k1 / 32))
    C1 = np.stack(cols, axis=1)                                         # [32, 32]
    Y = mma(split(x, kind), split(C1, kind), order)                     # [T, j, 32] float32
    Yr = Y[..., :17].astype(f32)
    Yi = np.zeros_like(Yr)
    Yi[..., 1:16] = Y[..., 17:]
    j = np.arange(16)[:, None]
    k1 = np.arange(17)[None, :]
    tw = np.exp(-2j * np.pi * j * k1 / 512)
    twr, twi = tw.real.astype(f32), tw.imag.astype(f32)
    Ar = (Yr * twr - Yi * twi).astype(f32)                              # float32 CUDA-core twiddle
    Ai = (Yr * twi + Yi * twr).astype(f32)
    A = np.concatenate([Ar.transpose(0, 2, 1), Ai.transpose(0, 2, 1)], axis=-1) * f32(mid)    # [T, k1, 32 = (re j | im j)]
    jj = np.arange(16)[:, None]
    k2 = np.arange(16)[None, :]
    W = np.exp(-2j * np.pi * jj * k2 / 16)
    C2 = np.block([[W.real, W.imag], [-W.imag, W.real]])               # [32, 32]: (re | im) of k2
    Z = mma(split(A, kind), split(C2, kind), order)                     # [T, k1, 32]
    Xr, Xi = Z[..., :16].astype(f64) / (pre * mid), Z[..., 16:].astype(f64) / (pre * mid)           # X[k1 + 32 k2]
    P = np.zeros((T, 257))
    for a in range(17):
        for b in range(16):
            k = a + 32 * b
            if k <= 256:
                P[:, k] = Xr[:, a, b] ** 2 + Xi[:, a, b] ** 2
            kk = 512 - k
            if 0 < kk <= 256:
                P[:, kk] = Xr[:, a, b] ** 2 + Xi[:, a, b] ** 2
    return P[:, :256]


def main():
    bank = kf.mel_banks(80, 512, 16000.0, 20.0, 0.0, f64)
    for name, x in [("uniform_16000", synth.uniform_pcm(1234, 16000, 16000)),
                    ("gauss0.1_16000", np.clip(0.1 * np.random.default_rng(7).standard_normal(16000), -1, 1).astype(f32)),
                    ("quiet 1e-3 x uniform", (synth.uniform_pcm(1234, 16000, 16000) * f32(1e-3 / 0.3)).astype(f32))]:
        xs = x.astype(f64) * 32768.0
        z64 = kf.preprocess_frames(xs, 512, 400, 160, "hamming")
        truth = np.log(np.maximum((np.abs(np.fft.rfft(z64, axis=1))[:, :256] ** 2) @ bank.T, 1.19e-7))
        z32 = kf.preprocess_frames((x * f32(32768.0)).astype(f32), 512, 400, 160, "hamming").astype(f32)
        rows = []
        p32 = (np.abs(np.fft.rfft(z32.astype(f32), axis=1).astype(np.complex64))[:, :256].astype(f64)) ** 2
        rows.append(("float32 FFT (reference arithmetic)", np.log(np.maximum(p32 @ bank.T, 1.19e-7))))
        for kind in ("fp16x2", "bf16x3"):
            rows.append((f"GEMM-FFT {kind}", np.log(np.maximum(gemm_fft_power(z32, kind) @ bank.T, 1.19e-7))))
        for label, lm in rows:
            e = np.abs(lm - truth)
            print(f"{name:22s} {label:36s} max {e.max():.2e}  mean {e.mean():.2e}  at {np.unravel_index(np.argmax(e), e.shape)}")


if __name__ == "__main__":
    main()
