"""CPU oracle for the speech front-end hot path (fbank -> LFR -> CMVN).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`toolbox_for_asr_and_tts_b200/`) imports this directory; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it, and only as the checker / the baseline that is timed.

Contents
  kaldi_fbank_np.py   numpy restatement of torchaudio.compliance.kaldi.fbank
                      (TA:44-217, TA:318-331, TA:436-645), float32 or float64.
  wav_frontend_np.py  numpy restatement of FunASR WavFrontend / apply_lfr /
                      apply_cmvn / load_cmvn (VF:23-168), the streaming
                      WavFrontendOnline semantics (UPSTREAM-RECALLED, pinned by the
                      stream-concat == offline invariant) and compute_audio_cmvn.
  tts_mel_np.py       frozen definition of the 24 kHz TTS log-mel (no reference
                      code exists for it: parity unpinned).
  ref_thirdparty.py   runs the ACTUAL third-party code the reference calls
                      (torchaudio.compliance.kaldi.fbank and, when importable,
                      vLLM's verbatim copy of funasr WavFrontend).  Used to pin the
                      restatement, to generate tests/golden/*.npz and as the
                      `cpu_baseline.kind == "reference"` arm of bench.py.

TA = site-packages/torchaudio/compliance/kaldi.py
VF = site-packages/vllm/transformers_utils/processors/funasr.py

Parity pin: the reference repository holds no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the third-party code
itself, run in the build container by tests/golden/make_golden.py (committed)
and re-checked wherever torchaudio is importable.
"""
