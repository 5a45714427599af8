"""CMVN tables: parsing the Kaldi-nnet `am.mvn` text file the reference's models ship (VF:63-86, `load_cmvn`),
writing one, and turning global statistics into a table (upstream funasr/bin/compute_audio_cmvn.py semantics)."""
from __future__ import annotations

import numpy as np
import torch


def load_cmvn(cmvn_file: str) -> torch.Tensor:
    """Same result as upstream `load_cmvn` (VF:63-86): float32 [2, dim]; row 0 = <AddShift>, row 1 = <Rescale>."""
    rows = {}
    with open(cmvn_file, encoding="utf-8") as f:
        lines = [ln.split() for ln in f]
    for i, tok in enumerate(lines):
        if tok and tok[0] in ("<AddShift>", "<Rescale>") and i + 1 < len(lines):
            nxt = lines[i + 1]
            if nxt and nxt[0] == "<LearnRateCoef>":
                rows[tok[0]] = np.array(nxt[3:len(nxt) - 1]).astype(np.float32)
    cmvn = np.array([rows.get("<AddShift>", np.zeros(0, np.float32)), rows.get("<Rescale>", np.zeros(0, np.float32))])
    return torch.as_tensor(cmvn, dtype=torch.float32)


def write_cmvn(path: str, shift, scale) -> None:
    shift = np.asarray(shift, dtype=np.float32)
    scale = np.asarray(scale, dtype=np.float32)
    d = shift.shape[0]
    with open(path, "w", encoding="utf-8") as f:
        f.write("<Nnet> \n")
        f.write(f"<Splice> {d} {d}\n[ 0 ]\n")
        f.write(f"<AddShift> {d} {d} \n")
        f.write("<LearnRateCoef> 0 [ " + " ".join(repr(float(v)) for v in shift) + " ]\n")
        f.write(f"<Rescale> {d} {d}\n")
        f.write("<LearnRateCoef> 0 [ " + " ".join(repr(float(v)) for v in scale) + " ]\n")
        f.write("</Nnet> \n")


def stats_to_cmvn(stats: torch.Tensor) -> torch.Tensor:
    """stats = float64 [2*D+1] (sum, sum of squares, row count) -> float32 [2, D] (AddShift=-mean, Rescale=1/std)."""
    stats = stats.detach().to("cpu", torch.float64)
    d = (stats.numel() - 1) // 2
    n = stats[2 * d]
    mean = stats[:d] / n
    var = stats[d:2 * d] / n - mean * mean
    return torch.stack([-mean, 1.0 / torch.sqrt(var)]).to(torch.float32)
