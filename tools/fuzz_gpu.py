"""GPU soak test (not part of the pytest suites): random ragged batches through the warp kernel, the tile kernel and the
int16 path, compared with each other (float32 rounding noise only) - looks for rare scheduling / indexing bugs.

    python tools/fuzz_gpu.py [--seconds 120] [--seed 0]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from toolbox_for_asr_and_tts_b200 import WavFrontend, synth  # noqa: E402


def close(a, b, cm, m):
    """Comparison in the log-mel domain with the depth-aware tolerance of tests/conftest.py: returns the largest error
    among bins within 12 nepers of the frame's peak, the largest excess over max(3e-3, 2e-6 e^(d/2)) below that, and the
    mean error."""
    sh, sc = cm[0].astype(np.float64), cm[1].astype(np.float64)
    la = (a.astype(np.float64) / sc - sh).reshape(a.shape[:-1] + (m, 80))
    lb = (b.astype(np.float64) / sc - sh).reshape(b.shape[:-1] + (m, 80))
    err = np.abs(la - lb)
    depth = lb.max(axis=-1, keepdims=True) - lb
    deep = depth > 12.0
    tol = np.maximum(3e-3, 2e-6 * np.exp(np.minimum(depth, 60.0) / 2.0))
    excess = float((err - tol)[deep].max()) if deep.any() else -1.0
    return float(err[~deep].max()), excess, float(err.mean())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--only", type=int, default=-1, help="replay the random draws and run just this iteration, with diagnostics")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    dev = "cuda:0"
    confs = [(7, 6), (5, 1), (1, 1), (3, 2)]
    fes = {}
    for m, n in confs:
        cm = np.stack([rng.normal(-8.0, 1.0, m * 80), rng.uniform(0.2, 0.5, m * 80)]).astype(np.float32)
        fw = WavFrontend(cmvn=torch.from_numpy(cm), fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10,
                         lfr_m=m, lfr_n=n, dither=0.0)
        ft = WavFrontend(cmvn=torch.from_numpy(cm), fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10,
                         lfr_m=m, lfr_n=n, dither=0.0)
        ft.select_kernel("tile")
        fes[(m, n)] = (fw, ft, cm)
    t0, it, worst, worst_deep = time.time(), 0, 0.0, -1.0
    while time.time() - t0 < a.seconds:
        m, n = confs[int(rng.integers(len(confs)))]
        fw, ft, cm = fes[(m, n)]
        B = int(rng.integers(1, 48))
        kind = int(rng.integers(3))
        lens = (rng.integers(400, 4000, B) if kind == 0 else rng.integers(400, 120000, B) if kind == 1
                else rng.integers(300, 900, B)).astype(np.int64)
        align = int(rng.choice([1, 2, 4]))
        lead = int(rng.integers(0, 7))
        offs, total = synth.packed_offsets(lens, align=align)
        offs = offs + lead
        ints = rng.integers(-12000, 12000, int(total) + lead + 16, dtype=np.int16)
        if a.only >= 0 and it < a.only:
            it += 1
            continue
        flat_i = torch.from_numpy(ints).to(dev)
        flat_f = (flat_i.to(torch.float32) / 32768.0)
        a1, l1 = fw.forward_packed(flat_f, offs, lens)
        a2, l2 = ft.forward_packed(flat_f, offs, lens)
        a3, l3 = fw.forward_packed(flat_i, offs, lens)
        assert torch.equal(l1, l2) and torch.equal(l1, l3), (it, "lengths")
        assert torch.equal(a1, a3), (it, "int16 path differs from the float path")
        if a.only >= 0:
            print("iteration", it, "lfr", (m, n), "B", B, "kind", kind, "align", align, "lead", lead, "lens", lens.tolist())
            bad = torch.nonzero((a1 == 0) != (a2 == 0))
            print("finite:", bool(torch.isfinite(a1).all()), bool(torch.isfinite(a2).all()), "zero-pattern mismatches:", bad.shape[0])
            for u, r, c in bad[:12].tolist():
                print("  utt", u, "len", int(lens[u]), "rows", int(l1[u]), "row", r, "col", c, "warp", float(a1[u, r, c]), "tile", float(a2[u, r, c]))
            d = (a1 - a2).abs()
            w = torch.nonzero(d == d.max())[0].tolist()
            print("max |warp - tile|", float(d.max()), "at", w, float(a1[tuple(w)]), float(a2[tuple(w)]))
        assert torch.isfinite(a1).all() and torch.isfinite(a2).all(), (it, "finiteness")
        rows = torch.arange(a1.shape[1], device=a1.device)[None, :, None]
        pad = rows >= l1.to(a1.device)[:, None, None]
        assert not (a1 * pad).any() and not (a2 * pad).any(), (it, "padding rows must be zero")
        mx, mx_deep, mean = close(a1.cpu().numpy(), a2.cpu().numpy(), cm, m)
        worst, worst_deep = max(worst, mx), max(worst_deep, mx_deep)
        assert mx <= 2e-3 and mx_deep <= 0.0 and mean <= 2e-5, (it, m, n, B, mx, mx_deep, mean)
        it += 1
        if a.only >= 0:
            break
    print(f"fuzz ok: {it} random batches in {time.time() - t0:.0f} s, worst warp-vs-tile log-mel difference {worst:.2e} "
          f"(largest excess over the depth-aware bound below 12 nepers: {worst_deep:.2e}, must be <= 0)")


if __name__ == "__main__":
    main()
