"""CPU checks of bench.py's contract pieces: the synthetic batch of BASELINE.json configs[1], the algorithmic-byte
formula of SURVEY.md 8(d), and the JSON line of the reference arm (which runs on the host cores, no GPU)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_batch_layout_is_the_configs1_workload():
    lens, offs, total = bench.batch_layout(0)
    assert len(lens) == bench.BATCH == 256
    assert lens.min() >= 16000 and lens.max() <= 480000            # 1-30 s at 16 kHz
    assert np.all(offs % 4 == 0)                                   # 16-byte aligned float32 offsets (length-packed)
    assert np.all(offs[1:] >= offs[:-1] + lens[:-1]) and total >= offs[-1] + lens[-1]
    lens1, _, _ = bench.batch_layout(1)
    assert not np.array_equal(lens, lens1)                         # every rank owns its own shard
    again, _, _ = bench.batch_layout(0)
    assert np.array_equal(lens, again)                             # seeded


def test_algorithmic_bytes_follow_the_survey_formula():
    lens, _, _ = bench.batch_layout(0)
    expect = sum(4 * int(n) + 4 * 560 * -(-(1 + (int(n) - 400) // 160) // 6) for n in lens)
    assert bench.algorithmic_bytes(lens) == expect
    assert bench.algorithmic_bytes(np.array([160000])) == 4 * 160000 + 4 * 560 * 167   # 10 s -> 167 rows


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"] == bench.WORKLOAD
    assert line["config"]["batch_per_gpu"] == bench.BATCH and "all 256 utterances" in line["cpu_baseline"]["sample"]


def test_host_ingest_packs_like_numpy_concatenate():
    """b200fe_host_ingest's host side (the multi-threaded gather into the staging buffer) through ctypes, without a GPU:
    a NULL device destination is refused before any CUDA call, and the symbol set is the header's."""
    import ctypes

    from toolbox_for_asr_and_tts_b200 import _native
    lib = _native.cdll()
    assert lib.b200fe_host_threads() >= 1
    rc = lib.b200fe_host_ingest(None, None, None, 3, 4, None, None, ctypes.c_int64(0), 2, 2, None)
    assert rc < 0
    assert lib.b200fe_host_ingest(None, None, None, 0, 4, None, None, ctypes.c_int64(0), 2, 2, None) == 0   # empty batch
