// torch C++ extension over the C ABI of libb200fe.so (include/b200fe.h).
//
// Registers `torch.ops.b200fe.*`.  at::Tensor stops here: below this file only raw device pointers, sizes and the
// current CUDA stream cross into the library.  PyTorch is used for device memory (caching allocator) and streams.
// No CPU fallback: every op checks that its tensors are CUDA tensors and raises otherwise.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <tuple>
#include <vector>

#include "../../include/b200fe.h"

namespace {

b200fe_handle* H(int64_t h) { return reinterpret_cast<b200fe_handle*>(static_cast<intptr_t>(h)); }

void check(int rc, b200fe_handle* h, const char* what) {
  if (rc == B200FE_OK) return;
  const char* msg = b200fe_last_error(h);
  const std::string text = std::string(what) + ": " + (msg ? msg : "");
  TORCH_CHECK_NOT_IMPLEMENTED(rc != B200FE_E_UNSUPPORTED, text);
  TORCH_CHECK(false, text);
}

void* cur_stream() { return static_cast<void*>(at::cuda::getCurrentCUDAStream().stream()); }

int64_t fe_create(int64_t fs, double frame_length, double frame_shift, int64_t n_mels, int64_t window_type, int64_t lfr_m,
                  int64_t lfr_n, double dither, bool snip_edges, bool upscale, double preemph, bool remove_dc,
                  double low_freq, double high_freq, double blackman_coeff, const c10::optional<at::Tensor>& cmvn) {
  b200fe_config c;
  b200fe_default_config(&c);
  c.sample_rate = (int32_t)fs;
  c.frame_length_ms = (float)frame_length;
  c.frame_shift_ms = (float)frame_shift;
  c.n_mels = (int32_t)n_mels;
  c.window_type = (int32_t)window_type;
  c.lfr_m = (int32_t)lfr_m;
  c.lfr_n = (int32_t)lfr_n;
  c.dither = (float)dither;
  c.snip_edges = snip_edges;
  c.upscale_samples = upscale;
  c.preemphasis = (float)preemph;
  c.remove_dc_offset = remove_dc;
  c.low_freq = (float)low_freq;
  c.high_freq = (float)high_freq;
  c.blackman_coeff = (float)blackman_coeff;
  const float* cm = nullptr;
  at::Tensor cm_host;
  if (cmvn.has_value() && cmvn->defined()) {
    cm_host = cmvn->to(at::kCPU, at::kFloat).contiguous();
    TORCH_CHECK(cm_host.dim() == 2 && cm_host.size(0) == 2 && cm_host.size(1) >= n_mels * lfr_m,
                "cmvn must be [2, >= n_mels*lfr_m]");
    if (cm_host.size(1) != n_mels * lfr_m) cm_host = cm_host.slice(1, 0, n_mels * lfr_m).contiguous();
    cm = cm_host.data_ptr<float>();
  }
  b200fe_handle* h = nullptr;
  check(b200fe_create(&c, cm, &h), nullptr, "b200fe_create");
  return static_cast<int64_t>(reinterpret_cast<intptr_t>(h));
}

void fe_destroy(int64_t h) { b200fe_destroy(H(h)); }

std::tuple<at::Tensor, at::Tensor> fe_plan(int64_t h, const at::Tensor& lengths) {
  auto len = lengths.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)len.numel();
  auto nf = at::empty({b}, at::kLong), nr = at::empty({b}, at::kLong);
  int64_t max_rows = 0;
  size_t ws = 0;
  check(b200fe_plan(H(h), len.data_ptr<int64_t>(), b, nf.data_ptr<int64_t>(), nr.data_ptr<int64_t>(), &max_rows, &ws), H(h),
        "b200fe_plan");
  return {nf, nr};
}

// wave: CUDA float32 (or int16 PCM, value = s / 32768), either [B, Nmax] (offsets undefined) or a flat length-packed
// buffer with host offsets.
std::tuple<at::Tensor, at::Tensor> fe_forward(int64_t h, const at::Tensor& wave, const c10::optional<at::Tensor>& offsets,
                                              const at::Tensor& lengths, int64_t rows_cap,
                                              const c10::optional<at::Tensor>& stats, int64_t seed) {
  TORCH_CHECK(wave.is_cuda(), "b200fe.forward: waveform must be a CUDA tensor (there is no CPU fallback)");
  const bool pcm16 = wave.scalar_type() == at::kShort;
  TORCH_CHECK(pcm16 || wave.scalar_type() == at::kFloat, "b200fe.forward: waveform must be float32 or int16");
  c10::cuda::CUDAGuard guard(wave.device());
  auto w = wave.contiguous();
  auto len = lengths.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)len.numel();
  TORCH_CHECK(b > 0, "received an empty list of sequences");   // what pad_sequence raises upstream (VF:163)
  at::Tensor off;
  const int64_t* off_ptr = nullptr;
  int64_t row_stride = 0;
  if (offsets.has_value() && offsets->defined()) {
    off = offsets->to(at::kCPU, at::kLong).contiguous();
    TORCH_CHECK(off.numel() == b, "offsets and lengths differ in size");
    off_ptr = off.data_ptr<int64_t>();
  } else {
    TORCH_CHECK(w.dim() == 2 && w.size(0) == b, "waveform must be [B, Nmax] when no offsets are given");
    row_stride = w.size(1);
    // upstream slices input[i][:len] (VF:138): a length beyond the row would read the next utterance here
    for (int u = 0; u < b; ++u)
      TORCH_CHECK(len.data_ptr<int64_t>()[u] >= 0 && len.data_ptr<int64_t>()[u] <= row_stride,
                  "input_lengths[", u, "] = ", len.data_ptr<int64_t>()[u], " is outside [0, ", row_stride, "]");
  }
  int64_t max_rows = 0;
  size_t ws = 0;
  std::vector<int64_t> n_rows(b);
  check(b200fe_plan(H(h), len.data_ptr<int64_t>(), b, nullptr, n_rows.data(), &max_rows, &ws), H(h), "b200fe_plan");
  const bool rows_packed = rows_cap == B200FE_ROWS_PACKED;     // [sum of rows, D] instead of the padded [B, rows_cap, D]
  if (!rows_packed && rows_cap <= 0) rows_cap = max_rows;
  const int64_t d = b200fe_output_dim(H(h));
  auto opts = w.options().dtype(at::kFloat);
  int64_t total_rows = 0;
  for (int u = 0; u < b; ++u) total_rows += n_rows[u];
  auto feats = rows_packed ? at::empty({total_rows, d}, opts) : at::empty({b, rows_cap, d}, opts);
  auto lens = at::empty({b}, opts.dtype(at::kLong));
  auto work = at::empty({(int64_t)(ws ? ws : 256)}, opts.dtype(at::kByte));
  double* st = nullptr;
  if (stats.has_value() && stats->defined()) {
    TORCH_CHECK(stats->is_cuda() && stats->scalar_type() == at::kDouble && stats->is_contiguous() &&
                    stats->numel() == 2 * d + 1, "stats must be a contiguous CUDA float64 [2*D+1] tensor");
    st = stats->data_ptr<double>();
  }
  if (pcm16) {
    TORCH_CHECK(st == nullptr, "b200fe.forward: the statistics pass takes float32 input");
    check(b200fe_forward_pcm16(H(h), w.data_ptr<int16_t>(), w.numel(), off_ptr, row_stride, len.data_ptr<int64_t>(), b,
                               feats.data_ptr<float>(), rows_cap, lens.data_ptr<int64_t>(), (uint64_t)seed, work.data_ptr(),
                               (size_t)work.numel(), cur_stream()),
          H(h), "b200fe_forward_pcm16");
    return {feats, lens};
  }
  check(b200fe_forward(H(h), w.data_ptr<float>(), w.numel(), off_ptr, row_stride, len.data_ptr<int64_t>(), b,
                       feats.data_ptr<float>(), rows_cap, lens.data_ptr<int64_t>(), st, (uint64_t)seed, work.data_ptr(),
                       (size_t)work.numel(), cur_stream()),
        H(h), "b200fe_forward");
  return {feats, lens};
}

std::tuple<at::Tensor, at::Tensor> fe_lfr_cmvn(int64_t h, const at::Tensor& fbank, const at::Tensor& n_frames) {
  TORCH_CHECK(fbank.is_cuda() && fbank.scalar_type() == at::kFloat && fbank.dim() == 3,
              "b200fe.lfr_cmvn: features must be a CUDA float32 [B, T, n_mels] tensor");
  c10::cuda::CUDAGuard guard(fbank.device());
  auto x = fbank.contiguous();
  auto nf = n_frames.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)x.size(0);
  TORCH_CHECK(nf.numel() == b, "n_frames size mismatch");
  int64_t max_rows = 0;
  // rows per utterance = ceil(T / lfr_n): size the output from the largest frame count
  const int64_t d = b200fe_output_dim(H(h));
  const int64_t m = x.size(2);
  TORCH_CHECK(d % m == 0, "feature dim does not match the front-end");
  int64_t tmax = 0;
  for (int i = 0; i < b; ++i) tmax = std::max<int64_t>(tmax, nf.data_ptr<int64_t>()[i]);
  // lfr_n is not exported; recover rows via plan-free arithmetic in the library: ask with rows_cap = tmax (upper bound)
  max_rows = tmax;
  auto feats_full = at::empty({b, max_rows, d}, x.options());
  auto lens = at::empty({b}, x.options().dtype(at::kLong));
  auto work = at::empty({(int64_t)(1024 + 128 * (int64_t)b)}, x.options().dtype(at::kByte));
  check(b200fe_lfr_cmvn(H(h), x.data_ptr<float>(), x.size(1), nf.data_ptr<int64_t>(), b, feats_full.data_ptr<float>(), max_rows,
                        lens.data_ptr<int64_t>(), work.data_ptr(), (size_t)work.numel(), cur_stream()),
        H(h), "b200fe_lfr_cmvn");
  return {feats_full, lens};
}

std::tuple<at::Tensor, at::Tensor> fe_get_tables(int64_t h) {
  const int L = b200fe_frame_samples(H(h)), nfft = b200fe_fft_size(H(h));
  const int d = b200fe_output_dim(H(h));
  (void)d;
  auto win = at::empty({L}, at::kFloat);
  // n_mels is recovered from the caller; the library writes n_mels * nfft/2 floats
  auto mel = at::zeros({128, nfft / 2}, at::kFloat);
  check(b200fe_get_tables(H(h), win.data_ptr<float>(), mel.data_ptr<float>()), H(h), "b200fe_get_tables");
  return {win, mel};
}

std::vector<int64_t> fe_geometry(int64_t h) {
  return {b200fe_frame_samples(H(h)), b200fe_shift_samples(H(h)), b200fe_fft_size(H(h)), b200fe_output_dim(H(h))};
}

at::Tensor fe_stream_state(int64_t h, int64_t n_streams, int64_t max_chunk, c10::Device device) {
  size_t bytes = 0;
  check(b200fe_stream_state_bytes(H(h), (int)n_streams, (int)max_chunk, &bytes), H(h), "b200fe_stream_state_bytes");
  TORCH_CHECK(device.is_cuda(), "stream state must live on a CUDA device");
  c10::cuda::CUDAGuard guard(device);
  auto st = at::zeros({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte).device(device));
  return st;
}

int64_t fe_stream_max_rows(int64_t h, int64_t max_chunk) { return b200fe_stream_max_rows(H(h), (int)max_chunk); }

void fe_stream_reset(int64_t h, at::Tensor state, int64_t n_streams, int64_t max_chunk, const c10::optional<at::Tensor>& ids) {
  TORCH_CHECK(state.is_cuda(), "state must be CUDA");
  c10::cuda::CUDAGuard guard(state.device());
  const int32_t* idp = nullptr;
  int n = 0;
  at::Tensor idt;
  if (ids.has_value() && ids->defined()) {
    idt = ids->to(state.device(), at::kInt).contiguous();
    idp = idt.data_ptr<int32_t>();
    n = (int)idt.numel();
  }
  check(b200fe_stream_reset(H(h), state.data_ptr(), (int)n_streams, (int)max_chunk, idp, n, cur_stream()), H(h),
        "b200fe_stream_reset");
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> fe_stream_push_impl(int64_t h, at::Tensor state, int64_t n_streams,
                                                                   int64_t max_chunk, const at::Tensor& chunks,
                                                                   const at::Tensor& chunk_lens, const at::Tensor& stream_ids,
                                                                   const c10::optional<at::Tensor>& is_final, bool want_stats) {
  TORCH_CHECK(state.is_cuda() && chunks.is_cuda(), "b200fe.stream_push: state and chunks must be CUDA tensors");
  TORCH_CHECK(chunks.scalar_type() == at::kFloat && chunks.dim() == 2, "chunks must be float32 [n, chunk_max]");
  c10::cuda::CUDAGuard guard(chunks.device());
  auto c = chunks.contiguous();
  const int n = (int)c.size(0);
  TORCH_CHECK(c.size(1) <= max_chunk, "chunk wider than max_chunk_samples");
  auto cl = chunk_lens.to(c.device(), at::kInt).contiguous();
  auto ids = stream_ids.to(c.device(), at::kInt).contiguous();
  TORCH_CHECK(cl.numel() == n && ids.numel() == n, "chunk_lens / stream_ids size mismatch");
  at::Tensor fin;
  const uint8_t* finp = nullptr;
  if (is_final.has_value() && is_final->defined()) {
    fin = is_final->to(c.device(), at::kByte).contiguous();
    TORCH_CHECK(fin.numel() == n, "is_final size mismatch");
    finp = fin.data_ptr<uint8_t>();
  }
  const int64_t rows_cap = b200fe_stream_max_rows(H(h), (int)max_chunk);
  const int64_t d = b200fe_output_dim(H(h));
  auto feats = at::empty({n, rows_cap, d}, c.options());
  auto rows = at::empty({n}, c.options().dtype(at::kInt));
  at::Tensor stats;
  if (want_stats) stats = at::empty({n, 2}, c.options());
  check(b200fe_stream_push_stats(H(h), state.data_ptr(), (int)n_streams, (int)max_chunk, c.data_ptr<float>(), c.size(1),
                                 cl.data_ptr<int32_t>(), ids.data_ptr<int32_t>(), finp, n, feats.data_ptr<float>(), rows_cap,
                                 rows.data_ptr<int32_t>(), want_stats ? stats.data_ptr<float>() : nullptr, cur_stream()),
        H(h), "b200fe_stream_push");
  return {feats, rows, stats};
}

std::tuple<at::Tensor, at::Tensor> fe_stream_push(int64_t h, at::Tensor state, int64_t n_streams, int64_t max_chunk,
                                                  const at::Tensor& chunks, const at::Tensor& chunk_lens,
                                                  const at::Tensor& stream_ids, const c10::optional<at::Tensor>& is_final) {
  auto r = fe_stream_push_impl(h, state, n_streams, max_chunk, chunks, chunk_lens, stream_ids, is_final, false);
  return {std::get<0>(r), std::get<1>(r)};
}

// + [n, 2] = {mean |x|, max |x|} of every chunk: the reference's energy gate (R:voice_interface.py:1569-1578)
std::tuple<at::Tensor, at::Tensor, at::Tensor> fe_stream_push_stats(int64_t h, at::Tensor state, int64_t n_streams,
                                                                    int64_t max_chunk, const at::Tensor& chunks,
                                                                    const at::Tensor& chunk_lens, const at::Tensor& stream_ids,
                                                                    const c10::optional<at::Tensor>& is_final) {
  return fe_stream_push_impl(h, state, n_streams, max_chunk, chunks, chunk_lens, stream_ids, is_final, true);
}

void fe_synth_uniform_ids(at::Tensor wave, const at::Tensor& offsets, const at::Tensor& lengths,
                          const c10::optional<at::Tensor>& utt_ids, int64_t seed, double amp) {
  TORCH_CHECK(wave.is_cuda() && wave.scalar_type() == at::kFloat && wave.is_contiguous(), "wave must be contiguous CUDA float32");
  c10::cuda::CUDAGuard guard(wave.device());
  auto off = offsets.to(wave.device(), at::kLong).contiguous();
  auto len = lengths.to(wave.device(), at::kLong).contiguous();
  at::Tensor ids;
  const int64_t* idp = nullptr;
  if (utt_ids.has_value() && utt_ids->defined()) {
    ids = utt_ids->to(wave.device(), at::kLong).contiguous();
    TORCH_CHECK(ids.numel() == len.numel(), "utt_ids and lengths differ in size");
    idp = ids.data_ptr<int64_t>();
  }
  int rc = b200fe_synth_uniform_ids(wave.data_ptr<float>(), off.data_ptr<int64_t>(), len.data_ptr<int64_t>(), idp,
                                    (int)len.numel(), (uint64_t)seed, (float)amp, cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_synth_uniform failed");
}

void fe_synth_uniform(at::Tensor wave, const at::Tensor& offsets, const at::Tensor& lengths, int64_t seed, double amp) {
  fe_synth_uniform_ids(wave, offsets, lengths, c10::nullopt, seed, amp);
}

int64_t tts_create(int64_t sample_rate, int64_t n_fft, int64_t hop, int64_t n_mels, double f_min, double f_max) {
  b200fe_tts* t = nullptr;
  check(b200fe_tts_create((int)sample_rate, (int)n_fft, (int)hop, (int)n_mels, (float)f_min, (float)f_max, &t), nullptr,
        "b200fe_tts_create");
  return static_cast<int64_t>(reinterpret_cast<intptr_t>(t));
}

void tts_destroy(int64_t t) { b200fe_tts_destroy(reinterpret_cast<b200fe_tts*>(static_cast<intptr_t>(t))); }

// wave: CUDA float32 [B, Nmax] (or flat with offsets); returns (mel [B, n_mels, max_frames], frames int64 [B])
std::tuple<at::Tensor, at::Tensor> tts_forward(int64_t t, const at::Tensor& wave, const c10::optional<at::Tensor>& offsets,
                                               const at::Tensor& lengths, int64_t hop, int64_t n_mels) {
  TORCH_CHECK(wave.is_cuda() && wave.scalar_type() == at::kFloat, "b200fe.tts_forward: waveform must be CUDA float32 (no CPU fallback)");
  c10::cuda::CUDAGuard guard(wave.device());
  auto w = wave.contiguous();
  auto len_host = lengths.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)len_host.numel();
  at::Tensor off_host;
  if (offsets.has_value() && offsets->defined()) {
    off_host = offsets->to(at::kCPU, at::kLong).contiguous();
    TORCH_CHECK(off_host.numel() == b, "offsets and lengths differ in size");
  } else {
    TORCH_CHECK(w.dim() == 2 && w.size(0) == b, "waveform must be [B, Nmax] when no offsets are given");
    off_host = at::arange(b, at::kLong) * w.size(1);
  }
  int64_t max_frames = 0;
  for (int i = 0; i < b; ++i) {
    const int64_t n = len_host.data_ptr<int64_t>()[i];
    TORCH_CHECK(n > (1024 - hop) / 2, "reflect padding needs more than (n_fft-hop)/2 samples");
    TORCH_CHECK(off_host.data_ptr<int64_t>()[i] >= 0 && off_host.data_ptr<int64_t>()[i] + n <= w.numel(), "utterance outside the wave buffer");
    max_frames = std::max(max_frames, n / hop);
  }
  auto len_dev = len_host.to(w.device(), true), off_dev = off_host.to(w.device(), true);
  auto mel = at::empty({b, n_mels, max_frames}, w.options());
  auto lens = at::empty({b}, w.options().dtype(at::kLong));
  int rc = b200fe_tts_forward(reinterpret_cast<b200fe_tts*>(static_cast<intptr_t>(t)), w.data_ptr<float>(), w.numel(),
                              off_dev.data_ptr<int64_t>(), len_dev.data_ptr<int64_t>(), b, max_frames, mel.data_ptr<float>(),
                              max_frames, lens.data_ptr<int64_t>(), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_tts_forward failed");
  return {mel, lens};
}

// float64 [B, 6] = {max, min, mean |x|, rms, clipping ratio, max |x|} per utterance
at::Tensor fe_audio_stats(const at::Tensor& wave, const c10::optional<at::Tensor>& offsets, const at::Tensor& lengths,
                          double clip_level) {
  TORCH_CHECK(wave.is_cuda() && wave.scalar_type() == at::kFloat, "b200fe.audio_stats: waveform must be CUDA float32");
  c10::cuda::CUDAGuard guard(wave.device());
  auto w = wave.contiguous();
  auto len_host = lengths.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)len_host.numel();
  int64_t row_stride = 0, max_len = 0;
  at::Tensor off_dev;
  if (offsets.has_value() && offsets->defined()) {
    auto off_host = offsets->to(at::kCPU, at::kLong).contiguous();
    TORCH_CHECK(off_host.numel() == b, "offsets and lengths differ in size");
    for (int i = 0; i < b; ++i)
      TORCH_CHECK(off_host.data_ptr<int64_t>()[i] >= 0 &&
                      off_host.data_ptr<int64_t>()[i] + len_host.data_ptr<int64_t>()[i] <= w.numel(),
                  "utterance outside the wave buffer");
    off_dev = off_host.to(w.device(), true);
  } else {
    TORCH_CHECK(w.dim() == 2 && w.size(0) == b, "waveform must be [B, Nmax] when no offsets are given");
    row_stride = w.size(1);
  }
  for (int i = 0; i < b; ++i) {
    const int64_t n = len_host.data_ptr<int64_t>()[i];
    TORCH_CHECK(n >= 0 && (off_dev.defined() || n <= row_stride), "bad length");
    max_len = std::max(max_len, n);
  }
  auto len_dev = len_host.to(w.device(), true);
  auto out = at::zeros({b, 6}, w.options().dtype(at::kDouble));
  auto work = at::empty({(int64_t)b200fe_audio_stats_workspace(b)}, w.options().dtype(at::kByte));
  int rc = b200fe_audio_stats(w.data_ptr<float>(), off_dev.defined() ? off_dev.data_ptr<int64_t>() : nullptr, row_stride,
                              len_dev.data_ptr<int64_t>(), b, max_len, (float)clip_level, out.data_ptr<double>(),
                              work.data_ptr(), (size_t)work.numel(), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_audio_stats failed (", rc, ")");
  return out;
}

// pcm: CUDA uint8 / int16 / int32, interleaved channels -> float32 mono at dst_rate
at::Tensor fe_ingest_pcm_impl(const at::Tensor& pcm, int64_t channels, int64_t src_rate, int64_t dst_rate, bool fourier);
at::Tensor fe_ingest_pcm(const at::Tensor& pcm, int64_t channels, int64_t src_rate, int64_t dst_rate) {
  return fe_ingest_pcm_impl(pcm, channels, src_rate, dst_rate, false);
}
// scipy.signal.resample as the resampler (the reference's branch when scipy is installed)
at::Tensor fe_ingest_pcm_fft(const at::Tensor& pcm, int64_t channels, int64_t src_rate, int64_t dst_rate) {
  return fe_ingest_pcm_impl(pcm, channels, src_rate, dst_rate, true);
}
at::Tensor fe_ingest_pcm_impl(const at::Tensor& pcm, int64_t channels, int64_t src_rate, int64_t dst_rate, bool fourier) {
  TORCH_CHECK(pcm.is_cuda(), "b200fe.ingest_pcm: PCM must be a CUDA tensor");
  int width = 0;
  if (pcm.scalar_type() == at::kByte) width = 1;
  else if (pcm.scalar_type() == at::kShort) width = 2;
  else if (pcm.scalar_type() == at::kInt) width = 4;
  TORCH_CHECK(width != 0, "b200fe.ingest_pcm: PCM must be uint8, int16 or int32");
  TORCH_CHECK(channels >= 1 && pcm.numel() % channels == 0, "b200fe.ingest_pcm: element count is not a multiple of the channel count");
  c10::cuda::CUDAGuard guard(pcm.device());
  auto x = pcm.contiguous();
  const int64_t n_in = x.numel() / channels;
  const int64_t n_out = b200fe_ingest_length(n_in, (int)src_rate, (int)dst_rate);
  TORCH_CHECK(n_out >= 0, "b200fe.ingest_pcm: bad sample rates");
  auto out = at::empty({n_out}, x.options().dtype(at::kFloat));
  if (n_out == 0) return out;
  int rc;
  if (fourier && src_rate != dst_rate) {
    const size_t ws = b200fe_resample_fft_workspace(n_in, (int)src_rate, (int)dst_rate);
    auto work = at::empty({(int64_t)ws}, x.options().dtype(at::kByte));
    rc = b200fe_ingest_pcm_fft(x.data_ptr(), width, (int)channels, n_in, (int)src_rate, (int)dst_rate, out.data_ptr<float>(),
                               n_out, work.data_ptr(), ws, cur_stream());
  } else {
    rc = b200fe_ingest_pcm(x.data_ptr(), width, (int)channels, n_in, (int)src_rate, (int)dst_rate, out.data_ptr<float>(),
                           n_out, cur_stream());
  }
  TORCH_CHECK(rc == B200FE_OK, "b200fe_ingest_pcm failed (", rc, ")");
  return out;
}

at::Tensor ring_state(int64_t n_streams, int64_t capacity, c10::Device device) {
  TORCH_CHECK(device.is_cuda(), "the audio ring lives on a CUDA device");
  const size_t bytes = b200fe_ring_state_bytes((int)n_streams, (int)capacity);
  TORCH_CHECK(bytes > 0, "bad ring geometry");
  c10::cuda::CUDAGuard guard(device);
  auto st = at::empty({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte).device(device));
  int rc = b200fe_ring_reset(st.data_ptr(), (int)n_streams, (int)capacity, nullptr, (int)n_streams, cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_ring_reset failed");
  return st;
}

void ring_reset(at::Tensor state, int64_t n_streams, int64_t capacity, const at::Tensor& ids) {
  c10::cuda::CUDAGuard guard(state.device());
  auto i = ids.to(state.device(), at::kInt).contiguous();
  int rc = b200fe_ring_reset(state.data_ptr(), (int)n_streams, (int)capacity, i.data_ptr<int32_t>(), (int)i.numel(), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_ring_reset failed");
}

void ring_push(at::Tensor state, int64_t n_streams, int64_t capacity, const at::Tensor& chunks, const at::Tensor& lens,
               const at::Tensor& ids) {
  TORCH_CHECK(chunks.is_cuda() && chunks.scalar_type() == at::kFloat && chunks.dim() == 2, "chunks must be CUDA float32 [n, len]");
  c10::cuda::CUDAGuard guard(state.device());
  auto c = chunks.contiguous();
  auto l = lens.to(state.device(), at::kInt).contiguous(), i = ids.to(state.device(), at::kInt).contiguous();
  TORCH_CHECK(l.numel() == c.size(0) && i.numel() == c.size(0), "one length and one stream id per chunk");
  int rc = b200fe_ring_push(state.data_ptr(), (int)n_streams, (int)capacity, c.data_ptr<float>(), c.size(1),
                            l.data_ptr<int32_t>(), i.data_ptr<int32_t>(), (int)c.size(0), (int)c.size(1), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_ring_push failed");
}

std::tuple<at::Tensor, at::Tensor> ring_window(const at::Tensor& state, int64_t n_streams, int64_t capacity, const at::Tensor& ids) {
  c10::cuda::CUDAGuard guard(state.device());
  auto i = ids.to(state.device(), at::kInt).contiguous();
  auto out = at::empty({i.numel(), capacity}, at::TensorOptions().dtype(at::kFloat).device(state.device()));
  auto lens = at::empty({i.numel()}, at::TensorOptions().dtype(at::kLong).device(state.device()));
  int rc = b200fe_ring_window(state.data_ptr(), (int)n_streams, (int)capacity, i.data_ptr<int32_t>(), (int)i.numel(),
                              out.data_ptr<float>(), lens.data_ptr<int64_t>(), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_ring_window failed");
  return {out, lens};
}

void fe_subtract_column_mean(at::Tensor feats, const at::Tensor& n_rows) {
  TORCH_CHECK(feats.is_cuda() && feats.scalar_type() == at::kFloat && feats.dim() == 3 && feats.is_contiguous(),
              "b200fe.subtract_column_mean: features must be a contiguous CUDA float32 [B, T, D] tensor");
  c10::cuda::CUDAGuard guard(feats.device());
  auto nr = n_rows.to(feats.device(), at::kLong).contiguous();
  TORCH_CHECK(nr.numel() == feats.size(0), "n_rows must have one entry per utterance");
  int rc = b200fe_subtract_column_mean(feats.data_ptr<float>(), feats.size(1), (int)feats.size(2), nr.data_ptr<int64_t>(),
                                       (int)feats.size(0), cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_subtract_column_mean failed (", rc, ")");
}

// Host ingest: utterances (CPU tensors, float32 or int16, 1-D contiguous; e.g. torch.from_numpy views of the arrays the
// reference hands to funasr) -> pinned staging -> device, multi-threaded gather pipelined with the PCIe copy, on the
// current stream of `wave_dev`'s device.
void fe_host_ingest(at::TensorList waves, const at::Tensor& lengths, const at::Tensor& offsets, at::Tensor staging,
                    at::Tensor wave_dev, int64_t groups, int64_t threads) {
  TORCH_CHECK(wave_dev.is_cuda() && wave_dev.is_contiguous(), "b200fe.host_ingest: destination must be a contiguous CUDA tensor");
  TORCH_CHECK(!staging.is_cuda() && staging.is_pinned() && staging.is_contiguous(),
              "b200fe.host_ingest: staging must be a pinned, contiguous host tensor");
  TORCH_CHECK(staging.scalar_type() == wave_dev.scalar_type() && staging.numel() >= wave_dev.numel(),
              "b200fe.host_ingest: staging and destination must have one dtype and staging must be at least as large");
  const auto dt = wave_dev.scalar_type();
  TORCH_CHECK(dt == at::kFloat || dt == at::kShort, "b200fe.host_ingest: float32 or int16 PCM");
  auto len = lengths.to(at::kCPU, at::kLong).contiguous();
  auto off = offsets.to(at::kCPU, at::kLong).contiguous();
  const int b = (int)len.numel();
  TORCH_CHECK((int)waves.size() == b && off.numel() == b, "b200fe.host_ingest: one utterance, length and offset per entry");
  std::vector<const void*> src(b);
  for (int u = 0; u < b; ++u) {
    const at::Tensor& w = waves[u];
    TORCH_CHECK(!w.is_cuda() && w.scalar_type() == dt && w.is_contiguous() && w.numel() >= len.data_ptr<int64_t>()[u],
                "b200fe.host_ingest: utterance ", u, " must be a contiguous host tensor of the destination's dtype holding its length");
    src[u] = w.data_ptr();
  }
  c10::cuda::CUDAGuard guard(wave_dev.device());
  const int rc = b200fe_host_ingest(src.data(), len.data_ptr<int64_t>(), off.data_ptr<int64_t>(), b, (int)wave_dev.element_size(),
                                    staging.data_ptr(), wave_dev.data_ptr(), wave_dev.numel(), (int)groups, (int)threads,
                                    cur_stream());
  TORCH_CHECK(rc == B200FE_OK, "b200fe_host_ingest failed (", rc, ")");
}

int64_t fe_launch_count(int64_t h) { return b200fe_launch_count(H(h)); }

void fe_select_kernel(int64_t h, int64_t which) { check(b200fe_select_kernel(H(h), (int)which), H(h), "b200fe_select_kernel"); }

void fe_profile_enable(int64_t h, int64_t every) { check(b200fe_profile_enable(H(h), (int)every), H(h), "b200fe_profile_enable"); }

std::tuple<double, int64_t> fe_profile_collect(int64_t h) {
  double ms = 0.0;
  int64_t n = 0;
  check(b200fe_profile_collect(H(h), &ms, &n), H(h), "b200fe_profile_collect");
  return {ms, n};
}

}  // namespace

TORCH_LIBRARY(b200fe, m) {
  m.def("create(int fs, float frame_length, float frame_shift, int n_mels, int window_type, int lfr_m, int lfr_n, "
        "float dither, bool snip_edges, bool upscale, float preemph, bool remove_dc, float low_freq, float high_freq, "
        "float blackman_coeff, Tensor? cmvn) -> int", fe_create);
  m.def("destroy(int h) -> ()", fe_destroy);
  m.def("plan(int h, Tensor lengths) -> (Tensor, Tensor)", fe_plan);
  m.def("forward(int h, Tensor wave, Tensor? offsets, Tensor lengths, int rows_cap, Tensor? stats, int seed) -> (Tensor, Tensor)",
        fe_forward);
  m.def("lfr_cmvn(int h, Tensor fbank, Tensor n_frames) -> (Tensor, Tensor)", fe_lfr_cmvn);
  m.def("get_tables(int h) -> (Tensor, Tensor)", fe_get_tables);
  m.def("geometry(int h) -> int[]", fe_geometry);
  m.def("stream_state(int h, int n_streams, int max_chunk, Device device) -> Tensor", fe_stream_state);
  m.def("stream_max_rows(int h, int max_chunk) -> int", fe_stream_max_rows);
  m.def("stream_reset(int h, Tensor state, int n_streams, int max_chunk, Tensor? ids) -> ()", fe_stream_reset);
  m.def("stream_push(int h, Tensor state, int n_streams, int max_chunk, Tensor chunks, Tensor chunk_lens, Tensor stream_ids, "
        "Tensor? is_final) -> (Tensor, Tensor)", fe_stream_push);
  m.def("stream_push_stats(int h, Tensor state, int n_streams, int max_chunk, Tensor chunks, Tensor chunk_lens, "
        "Tensor stream_ids, Tensor? is_final) -> (Tensor, Tensor, Tensor)", fe_stream_push_stats);
  m.def("synth_uniform(Tensor wave, Tensor offsets, Tensor lengths, int seed, float amp) -> ()", fe_synth_uniform);
  m.def("synth_uniform_ids(Tensor wave, Tensor offsets, Tensor lengths, Tensor? utt_ids, int seed, float amp) -> ()",
        fe_synth_uniform_ids);
  m.def("tts_create(int sample_rate, int n_fft, int hop, int n_mels, float f_min, float f_max) -> int", tts_create);
  m.def("tts_destroy(int t) -> ()", tts_destroy);
  m.def("tts_forward(int t, Tensor wave, Tensor? offsets, Tensor lengths, int hop, int n_mels) -> (Tensor, Tensor)", tts_forward);
  m.def("audio_stats(Tensor wave, Tensor? offsets, Tensor lengths, float clip_level) -> Tensor", fe_audio_stats);
  m.def("subtract_column_mean(Tensor(a!) feats, Tensor n_rows) -> ()", fe_subtract_column_mean);
  m.def("ingest_pcm(Tensor pcm, int channels, int src_rate, int dst_rate) -> Tensor", fe_ingest_pcm);
  m.def("ingest_pcm_fft(Tensor pcm, int channels, int src_rate, int dst_rate) -> Tensor", fe_ingest_pcm_fft);
  m.def("ring_state(int n_streams, int capacity, Device device) -> Tensor", ring_state);
  m.def("ring_reset(Tensor(a!) state, int n_streams, int capacity, Tensor ids) -> ()", ring_reset);
  m.def("ring_push(Tensor(a!) state, int n_streams, int capacity, Tensor chunks, Tensor lens, Tensor ids) -> ()", ring_push);
  m.def("ring_window(Tensor state, int n_streams, int capacity, Tensor ids) -> (Tensor, Tensor)", ring_window);
  m.def("host_ingest(Tensor[] waves, Tensor lengths, Tensor offsets, Tensor(a!) staging, Tensor(b!) wave_dev, int groups, "
        "int threads) -> ()", fe_host_ingest);
  m.def("launch_count(int h) -> int", fe_launch_count);
  m.def("select_kernel(int h, int which) -> ()", fe_select_kernel);
  m.def("profile_enable(int h, int every) -> ()", fe_profile_enable);
  m.def("profile_collect(int h) -> (float, int)", fe_profile_collect);
}
