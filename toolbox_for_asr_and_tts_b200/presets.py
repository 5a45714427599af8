"""Front-end configurations of the models the reference loads through funasr / ModelScope, as WavFrontend keywords.

The reference never spells these out: funasr builds each model's front-end from the `frontend_conf` of the model's own
config.yaml (R:voice-service/app/services/voice_interface.py:422-429, 686-700, 832-841), and the model directories are
not in the tree (git-ignored weights).  The values below are the published configurations of those models
(UPSTREAM-RECALLED: not verifiable offline) and every one of them is just a point in the option space the kernels are
parity-tested over (tests/test_gpu_parity.py), so a differing config.yaml only changes the keywords.

    fe = WavFrontend(cmvn_file="<model dir>/am.mvn", **presets.PARAFORMER_ZH)
"""

# damo/speech_paraformer-large_asr_nat-zh-cn-16k (offline + streaming ASR; BASELINE.json configs[0..3])
PARAFORMER_ZH = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=7, lfr_n=6, dither=0.0)

# damo/speech_fsmn_vad_zh-cn-16k-common: WavFrontendOnline, 5 stacked frames, no subsampling, its own vad.mvn
# (per-chunk calls at R:voice_interface.py:1585-1590, R:voice-service/app/api/voice.py:465-470)
FSMN_VAD = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=5, lfr_n=1, dither=0.0)

# iic/speech_charctc_kws_phone-xiaoyun (wake word, R:voice_interface.py:422-429, called on the 1.6 s sliding window at
# :1370-1374): FSMN keyword spotter on 80-mel fbank, 5 frames stacked every 3 (input dimension 400)
CHARCTC_KWS = dict(fs=16000, window="hamming", n_mels=80, frame_length=25, frame_shift=10, lfr_m=5, lfr_n=3, dither=0.0)

# damo/speech_campplus_sv_zh-cn_16k-common (speaker verification, R:voice_interface.py:2430, 2520, 2558): Kaldi fbank with
# utterance mean normalisation (torchaudio.compliance.kaldi.fbank(..., num_mel_bins=80) - mean over frames)
CAMPP_SV = dict(fs=16000, window="povey", n_mels=80, frame_length=25, frame_shift=10, lfr_m=1, lfr_n=1, dither=0.0,
                subtract_mean=True)

KWS_WINDOW_SAMPLES = 25600      # the reference keeps the newest 1.6 s for wake-word detection (R:voice_interface.py:1126, 1308-1311)
PRE_SPEECH_SAMPLES = 6400       # 400 ms pre-speech guard (:1115-1116, 1742-1746)
