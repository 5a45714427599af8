/*
 * b200fe.h - C ABI of the B200-native speech front-end (libb200fe.so).
 *
 * This is the drop-in boundary for ONE path of terrense/toolbox-for-ASR-and-TTS: the speech front-end that
 * `funasr.AutoModel(...).generate(input=np.float32[N])` runs on the CPU before the acoustic model
 * (reference call sites R:voice-service/app/services/voice_interface.py:422-429,457,1370-1374,1585-1590,
 * 2049-2053; R:voice-service/full_voice_demo.py:327,425,501).  The arithmetic lives in third-party code:
 *   VF = vllm/transformers_utils/processors/funasr.py   (verbatim upstream funasr WavFrontend)
 *   TA = torchaudio/compliance/kaldi.py
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; every `*_dev` / device pointer is CUDA device memory owned by the caller;
 *   - every call is asynchronous on the `cudaStream_t` passed as `void* stream` (NULL = default stream),
 *     except b200fe_plan / b200fe_create / b200fe_destroy which are host-side;
 *   - return value: 0 = OK, negative = error (B200FE_E_*); b200fe_last_error(handle) gives the text;
 *   - no exceptions cross the ABI; there is NO CPU fallback: without a CUDA device create() fails;
 *   - one handle may be used from several host threads only with distinct streams AND distinct workspaces;
 *   - batches of more than 65535 utterances per call are taken by b200fe_forward / _pcm16 on the warp-kernel path
 *     only; the statistics pass, b200fe_lfr_cmvn, b200fe_audio_stats, b200fe_tts_forward and
 *     b200fe_subtract_column_mean put the batch index in grid.y and refuse (or fail to launch) beyond that.
 */
#ifndef B200FE_H_
#define B200FE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FE_VERSION 1

enum {
  B200FE_OK = 0,
  B200FE_E_INVALID = -1,      /* bad argument / option combination                                     */
  B200FE_E_UNSUPPORTED = -2,  /* valid Kaldi option that the CUDA path does not implement               */
  B200FE_E_CUDA = -3,         /* CUDA runtime error (text in b200fe_last_error)                          */
  B200FE_E_WORKSPACE = -4,    /* workspace too small for the planned batch                              */
  B200FE_E_SHORT = -5         /* an utterance is shorter than 2 samples (TA:142 assertion)              */
};

/* window_type, TA:86-113 (_feature_window_function) */
enum { B200FE_WIN_HAMMING = 0, B200FE_WIN_HANNING = 1, B200FE_WIN_POVEY = 2, B200FE_WIN_RECTANGULAR = 3,
       B200FE_WIN_BLACKMAN = 4 };

/* Options of WavFrontend.__init__ (VF:92-123) + the Kaldi fbank options it forwards (TA:514-541). */
typedef struct b200fe_config {
  int32_t struct_size;        /* sizeof(b200fe_config), for ABI evolution                                */
  int32_t sample_rate;        /* fs                      (VF:95, TA sample_frequency)  default 16000     */
  float   frame_length_ms;    /* frame_length            (VF:98)                       default 25        */
  float   frame_shift_ms;     /* frame_shift             (VF:99)                       default 10        */
  int32_t n_mels;             /* n_mels / num_mel_bins   (VF:97)                       default 80        */
  int32_t window_type;        /* window                  (VF:96) B200FE_WIN_*          default HAMMING   */
  int32_t lfr_m;              /* lfr_m                   (VF:102)                      default 1         */
  int32_t lfr_n;              /* lfr_n                   (VF:103)                      default 1         */
  float   dither;             /* dither                  (VF:104)  0 = off                               */
  int32_t snip_edges;         /* snip_edges              (VF:105)  only 1 is implemented                 */
  int32_t upscale_samples;    /* upsacle_samples         (VF:106)  multiply input by 2^15                */
  float   preemphasis;        /* preemphasis_coefficient (TA:527)                      default 0.97      */
  int32_t remove_dc_offset;   /* remove_dc_offset        (TA:529)                      default 1         */
  float   low_freq;           /* low_freq                (TA:524)                      default 20        */
  float   high_freq;          /* high_freq (<=0: offset from Nyquist) (TA:522)         default 0         */
  float   blackman_coeff;     /* blackman_coeff          (TA:516)                      default 0.42      */
  float   log_floor;          /* floor applied to mel energies before log (TA:633): FLT_EPSILON          */
  int32_t reserved[7];
} b200fe_config;

typedef struct b200fe_handle b200fe_handle;

/* Fill `cfg` with the WavFrontend defaults above (Paraformer-zh would then set lfr_m=7, lfr_n=6, dither=0). */
void b200fe_default_config(b200fe_config* cfg);

/* Replaces WavFrontend.__init__ + load_cmvn's result (VF:92-123, VF:63-86).
 * cmvn_host: NULL (no CMVN) or host float32 [2, n_mels*lfr_m]: row 0 = AddShift, row 1 = Rescale (VF:31-35).
 * Builds the window (TA:86-113) and the mel filterbank (TA:436-511) once, on the host, and uploads them -
 * the reference rebuilds both on every call (TA:201, TA:621). */
int b200fe_create(const b200fe_config* cfg, const float* cmvn_host, b200fe_handle** out);
void b200fe_destroy(b200fe_handle* h);
const char* b200fe_last_error(const b200fe_handle* h);   /* h may be NULL: error of the last failed create() */

/* WavFrontend.output_size() (VF:125-126) = n_mels*lfr_m; and the derived sample counts (TA:137-140). */
int b200fe_output_dim(const b200fe_handle* h);
int b200fe_frame_samples(const b200fe_handle* h);        /* window_size   */
int b200fe_shift_samples(const b200fe_handle* h);        /* window_shift  */
int b200fe_fft_size(const b200fe_handle* h);             /* padded_window_size */

/* Read back the tables built by create(), for parity tests against TA:86-113 / TA:436-511.
 * window_out: [frame_samples] (without the 2^15 upscale), mel_out: [n_mels, fft_size/2] row-major. */
int b200fe_get_tables(const b200fe_handle* h, float* window_out_host, float* mel_out_host);

/* Frame / LFR-row counts for a batch, bit-exact with TA:65-70 (_get_strided) and VF:43 (apply_lfr), and the
 * device workspace the batch needs.  lengths_host[i] = number of samples of utterance i.
 * n_frames_out / n_rows_out may be NULL.  Utterances shorter than one frame follow VF:147's
 * frame_length=min(frame_length, len/fs*1000) rule and yield one frame. */
int b200fe_plan(b200fe_handle* h, const int64_t* lengths_host, int batch,
                int64_t* n_frames_out, int64_t* n_rows_out, int64_t* max_rows_out, size_t* workspace_bytes);

/* rows_cap value that selects the ROWS-PACKED output of b200fe_forward / b200fe_forward_pcm16: feats_dev is
 * [sum_u n_rows[u], output_dim] and utterance u starts at row sum_{v<u} n_rows[v] (b200fe_plan gives n_rows); no padding
 * rows exist, so none are written (133 MB of zeros per call for BASELINE configs[1]).  For consumers that batch by
 * length themselves; the reference's own layout (pad_sequence, VF:163-166) is rows_cap >= max n_rows. */
#define B200FE_ROWS_PACKED (-1)

/* Replaces WavFrontend.forward (VF:128-168): fbank (TA:514-645) -> apply_lfr (VF:40-60) -> apply_cmvn (VF:23-37)
 * -> pad_sequence, for a whole batch in one pass.
 *   wave_dev      float32 PCM in [-1,1] (or already upscaled when upscale_samples=0)
 *   offsets_host  [batch] start of utterance i inside wave_dev, in samples (length-packed buffers), or NULL
 *                 for a dense [batch, row_stride] layout
 *   lengths_host  [batch] samples per utterance (the reference takes lengths on the host too, VF:138)
 *   wave_total    number of float32 elements addressable behind wave_dev (bounds the vector loads)
 *   feats_dev     out, float32 [batch, rows_cap, output_dim]; rows >= n_rows[i] are zero-filled (pad_sequence)
 *   feat_lens_dev out, int64 [batch] (VF:160), may be NULL
 *   stats_dev     NULL, or float64 [2*output_dim + 1]: sum, sum of squares, row count of the un-normalised LFR
 *                 features are ADDED to it (global CMVN statistics, funasr compute_audio_cmvn)
 *   workspace_dev at least the bytes b200fe_plan reported for these lengths
 */
int b200fe_forward(b200fe_handle* h, const float* wave_dev, int64_t wave_total,
                   const int64_t* offsets_host, int64_t row_stride, const int64_t* lengths_host, int batch,
                   float* feats_dev, int64_t rows_cap, int64_t* feat_lens_dev, double* stats_dev,
                   uint64_t dither_seed, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same as b200fe_forward for int16 PCM that is still in its wire format: sample value = s / 32768, the conversion the
 * reference applies on the host before it reaches the front-end (base64_to_audio_np, R:voice-service/app/services/
 * voice_interface.py:1008-1013).  Fused into the kernel's sample loads: half the HBM and PCIe bytes per audio-second
 * (69 333 B instead of 101 333 B algorithmic), results bit-identical to converting first and calling b200fe_forward.
 * wave_total counts int16 elements.  No statistics pass (use the float entry point for that). */
int b200fe_forward_pcm16(b200fe_handle* h, const int16_t* wave_dev, int64_t wave_total,
                         const int64_t* offsets_host, int64_t row_stride, const int64_t* lengths_host, int batch,
                         float* feats_dev, int64_t rows_cap, int64_t* feat_lens_dev, uint64_t dither_seed,
                         void* workspace_dev, size_t workspace_bytes, void* stream);

/* Replaces WavFrontend.forward_lfr_cmvn (VF:198-218): LFR + CMVN of given [batch, frames_cap, n_mels] features. */
int b200fe_lfr_cmvn(b200fe_handle* h, const float* fbank_dev, int64_t frames_cap, const int64_t* n_frames_host,
                    int batch, float* feats_dev, int64_t rows_cap, int64_t* feat_lens_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream);

/* ---- streaming: replaces WavFrontendOnline.forward(input, lengths, cache=..., is_final=...) (upstream funasr
 * wav_frontend.py, reached through vad_model.generate(cache=...) at R:voice-service/app/services/
 * voice_interface.py:1585-1590) and the per-session np.concatenate accumulation at :1304-1311,1688-1746.
 * The state slab keeps, per stream, the sample carry (input_cache), a log-mel buffer whose first rows are the LFR splice
 * frames carried from tick to tick (lfr_splice_cache; the frames of the current tick are appended behind them, which
 * is why its size depends on max_chunk_samples), counters and the per-tick scratch of the kernels, resident in HBM:
 * allocate b200fe_stream_state_bytes, zero it or call b200fe_stream_reset, and treat it as opaque (a byte copy is a
 * valid checkpoint). */
int b200fe_stream_state_bytes(const b200fe_handle* h, int n_streams, int max_chunk_samples, size_t* bytes);
int b200fe_stream_reset(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples,
                        const int32_t* stream_ids_dev_or_null, int n, void* stream);
/* chunks_dev [n, chunk_stride] float32; chunk_lens_dev [n]; stream_ids_dev [n] (distinct); is_final_dev [n] or
 * NULL; feats_dev out [n, rows_cap, output_dim]; rows_out_dev out [n]. */
int b200fe_stream_push(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples,
                       const float* chunks_dev, int64_t chunk_stride, const int32_t* chunk_lens_dev,
                       const int32_t* stream_ids_dev, const uint8_t* is_final_dev, int n,
                       float* feats_dev, int64_t rows_cap, int32_t* rows_out_dev, void* stream);
/* The same tick that also returns the reference's per-chunk energy gate inputs (R:voice-service/app/services/
 * voice_interface.py:1569-1578, 1298-1300: is_speech = mean|x| > 0.03 and max|x| > 0.17): chunk_stats_dev out [n, 2]
 * = {mean |x|, max |x|} of each pushed chunk (NULL = plain b200fe_stream_push). */
int b200fe_stream_push_stats(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples,
                             const float* chunks_dev, int64_t chunk_stride, const int32_t* chunk_lens_dev,
                             const int32_t* stream_ids_dev, const uint8_t* is_final_dev, int n,
                             float* feats_dev, int64_t rows_cap, int32_t* rows_out_dev, float* chunk_stats_dev,
                             void* stream);
/* Upper bound of rows one push can emit for a chunk of max_chunk_samples (sizes rows_cap). */
int b200fe_stream_max_rows(const b200fe_handle* h, int max_chunk_samples);

/* ---- TTS-side log-mel (BASELINE.json configs[4]; no reference code exists for it, the definition is frozen in
 * oracle/tts_mel_np.py): n_fft = win = 1024, periodic Hann, reflect padding (n_fft-hop)/2 per side so that
 * frames = n_samples / hop, magnitude sqrt(re^2+im^2+1e-9), Slaney-scale Slaney-normalised filters, log(clamp 1e-5).
 * mel_dev: float32 [batch, n_mels, frames_cap] (mel-major); frames beyond an utterance are zero.
 * offsets_dev / lengths_dev: int64 device arrays; max_frames = max_i lengths[i] / hop sizes the launch.  batch <= 65535.
 * Clips whose first sample lies on an 8-byte boundary of wave_dev take the fast path (bulk copies of the samples).
 * One forward at a time per handle (calls are serialised inside: the handle owns the launch's work list). */
typedef struct b200fe_tts b200fe_tts;
int b200fe_tts_create(int sample_rate, int n_fft, int hop, int n_mels, float f_min, float f_max, b200fe_tts** out);
void b200fe_tts_destroy(b200fe_tts* t);
int b200fe_tts_forward(b200fe_tts* t, const float* wave_dev, int64_t wave_total, const int64_t* offsets_dev,
                       const int64_t* lengths_dev, int batch, int64_t max_frames, float* mel_dev, int64_t frames_cap,
                       int64_t* mel_lens_dev, void* stream);

/* ---- host ingest: the gather funasr performs on the HOST in front of the front-end call - the list of np.float32
 * utterances the reference hands to it (R:voice-service/app/services/voice_interface.py:2049-2053, 1370-1374) padded into
 * one [B, Nmax] tensor (pad_sequence in funasr's extract_fbank, UPSTREAM-RECALLED).  Here: utterance u (host pointer
 * utterances_host[u], lengths[u] elements of elem_size 4 = float32 or 2 = int16 PCM) is copied by `threads` host threads
 * (0 = all cores, at most 32) to staging_pinned + dst_offsets[u] (a caller-owned cudaMallocHost buffer) and from there to
 * wave_dev + dst_offsets[u] with cudaMemcpyAsync on `stream`, in `groups` groups of consecutive utterances so that the
 * gather of group g+1 overlaps the PCIe copy of group g.  dst_offsets must ascend and not overlap; the result is the
 * length-packed buffer b200fe_forward takes.  The call returns when the last copy is ENQUEUED; staging_pinned may be
 * reused once that copy has completed (record an event on `stream`). */
int b200fe_host_threads(void);
int b200fe_host_ingest(const void* const* utterances_host, const int64_t* lengths, const int64_t* dst_offsets, int batch,
                       int elem_size, void* staging_pinned, void* wave_dev, int64_t capacity_elems, int groups, int threads,
                       void* stream);

/* ---- by-products the reference computes on the host around its funasr calls.
 * b200fe_audio_stats replaces _log_audio_statistics and the per-chunk energy gate (R:voice-service/app/services/
 * voice_interface.py:873-939, 1298-1300, 1569-1570) for a whole ragged batch: out_dev is float64 [batch, 6] =
 * {max, min, mean |x|, rms = sqrt(mean x^2), clipping ratio = fraction with |x| >= clip_level (0.999 upstream), max |x|}.
 * offsets_dev / lengths_dev are int64 DEVICE arrays (offsets_dev NULL = dense [batch, row_stride]); max_length sizes the
 * launch; the workspace holds the per-utterance accumulators. */
size_t b200fe_audio_stats_workspace(int batch);
int b200fe_audio_stats(const float* wave_dev, const int64_t* offsets_dev, int64_t row_stride, const int64_t* lengths_dev,
                       int batch, int64_t max_length, float clip_level, double* out_dev, void* workspace_dev,
                       size_t workspace_bytes, void* stream);
/* base64_to_audio_np after the WAV header (R:voice_interface.py:1004-1034): wire PCM (sample_width 1 = uint8, 2 = int16,
 * 4 = int32; interleaved channels) -> float32 mono at dst_rate: width normalisation, channel mean and - when the rates
 * differ - the linear-interpolation resampling of its numpy branch (np.interp over np.linspace), in float64 and the
 * reference's operation order, so the float32 output is bit-identical to numpy's.  b200fe_ingest_pcm_fft below is the
 * branch the reference takes when scipy is installed.  n_frames_in counts sample FRAMES (all channels);
 * b200fe_ingest_length gives the output length, int(n * dst_rate / src_rate). */
int64_t b200fe_ingest_length(int64_t n_frames_in, int src_rate, int dst_rate);
int b200fe_ingest_pcm(const void* pcm_dev, int sample_width, int channels, int64_t n_frames_in, int src_rate, int dst_rate,
                      float* out_dev, int64_t out_capacity, void* stream);
/* The same ingest with scipy.signal.resample (Fourier method) as the resampler: the branch the reference takes when scipy
 * is importable (R:voice_interface.py:1022-1027): X = rfft(x), bins 0..min(n, num)/2 kept (the unpaired middle bin x2 when
 * shortening, x0.5 when lengthening), y = irfft(X * num / n, num), float64, cast to float32 (:1045).  Arbitrary lengths:
 * evaluated as the two DFT sums in float64 (see csrc/resample_fft.cuh); agrees with scipy to float64 round-off, i.e. to
 * <= 1 float32 ulp after the cast.  workspace: b200fe_resample_fft_workspace bytes of device memory. */
size_t b200fe_resample_fft_workspace(int64_t n_frames_in, int src_rate, int dst_rate);
int b200fe_ingest_pcm_fft(const void* pcm_dev, int sample_width, int channels, int64_t n_frames_in, int src_rate,
                          int dst_rate, float* out_dev, int64_t out_capacity, void* workspace_dev, size_t workspace_bytes,
                          void* stream);
/* Sliding audio windows per stream, resident in HBM: replaces `buf = np.concatenate([buf, chunk])[-target:]` of the KWS
 * window (1.6 s) and the pre-speech guard (0.4 s) (R:voice_interface.py:1304-1311, 1742-1746).  One slab holds a circular
 * buffer of capacity_samples per stream.  push appends chunk b ([n, chunk_stride] float32, chunk_lens_dev[b] samples) to
 * stream stream_ids_dev[b] (distinct ids within one call); window writes the newest min(total, capacity) samples of each
 * requested stream, oldest first, to out_dev [n, capacity_samples] (zero tail) and their count to lens_dev [n] - the
 * array the reference would hand to the model after its slice. */
size_t b200fe_ring_state_bytes(int n_streams, int capacity_samples);
int b200fe_ring_reset(void* state_dev, int n_streams, int capacity_samples, const int32_t* stream_ids_dev_or_null, int n,
                      void* stream);
int b200fe_ring_push(void* state_dev, int n_streams, int capacity_samples, const float* chunks_dev, int64_t chunk_stride,
                     const int32_t* chunk_lens_dev, const int32_t* stream_ids_dev, int n, int max_chunk_samples,
                     void* stream);
int b200fe_ring_window(const void* state_dev, int n_streams, int capacity_samples, const int32_t* stream_ids_dev, int n,
                       float* out_dev, int64_t* lens_dev, void* stream);
/* Kaldi subtract_mean (TA:642-644, _subtract_column_mean), i.e. the utterance mean normalisation of the CAM++
 * speaker-verification features (R:voice_interface.py:2430,2520,2558), in place on [batch, rows_cap, dim] features:
 * feats[u, t, :] -= mean over t < n_rows[u]. */
int b200fe_subtract_column_mean(float* feats_dev, int64_t rows_cap, int dim, const int64_t* n_rows_dev, int batch,
                                void* stream);

/* ---- bench / test support: counter-based synthetic PCM, identical to synth.py on the host.
 * x[u][n] = amp * (2*U01(hash(seed,u,n)) - 1) written at wave_dev[offsets_dev[u] + n], n < lengths_dev[u]. */
int b200fe_synth_uniform(float* wave_dev, const int64_t* offsets_dev, const int64_t* lengths_dev, int batch,
                         uint64_t seed, float amp, void* stream);
/* The same with an explicit generator id per batch entry (utt_ids_dev, int64 device array, NULL = 0..batch-1): a rank
 * synthesises its shard of a corpus (longest-first partition) bit-identically to the single-process
 * corpus. */
int b200fe_synth_uniform_ids(float* wave_dev, const int64_t* offsets_dev, const int64_t* lengths_dev,
                             const int64_t* utt_ids_dev_or_null, int batch, uint64_t seed, float amp, void* stream);

/* Kernel launches issued by this handle since create() (bench.py's gpu_launches). */
int64_t b200fe_launch_count(const b200fe_handle* h);

/* Where one fbank frame lands in the LFR-stacked output (VF:40-60, closed form: frame f is slot jj of row i wherever
 * clamp(lfr_n*i + jj - (lfr_m-1)/2, 0, n_frames-1) == f).  Host-side, no device needed: this is the planner the quad
 * list of the fused kernel is built with, exported so that it can be tested against apply_lfr on the CPU.
 * targets_out[k] = (jj << 27) | (row * lfr_m*n_mels + jj * n_mels), or 0xFFFFFFFF when unused.
 * Returns 1 when the frame takes the generic path instead (first / last frame of the utterance, or more than two
 * slots per frame), 0 otherwise, negative on bad arguments. */
int b200fe_lfr_targets(int frame, int n_frames, int n_rows, int lfr_m, int lfr_n, int n_mels, uint32_t targets_out[2]);

/* Kernel selection for A/B measurements and tests: 0 = automatic (warp-autonomous kernel whenever it applies),
 * 1 = always the tile kernel (the one that also accumulates CMVN statistics).  Results are identical. */
int b200fe_select_kernel(b200fe_handle* h, int which);

/* Roofline support: while enabled (every > 0), every `every`-th launch of the dominant (fused) kernel is bracketed by
 * CUDA events recorded on the launching stream (1 = all of them; an event pair costs a few microseconds of stream
 * time and keeps that launch from overlapping the tail of the kernel in front of it, so a throughput loop samples).
 * b200fe_profile_collect synchronises those events, returns the summed kernel time and the number of TIMED launches
 * since the last collect, and clears the list. */
int b200fe_profile_enable(b200fe_handle* h, int every);
int b200fe_profile_collect(b200fe_handle* h, double* total_ms, int64_t* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* B200FE_H_ */
