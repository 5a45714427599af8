// C ABI of csrc/extras.cuh (included by b200fe.cu).
extern "C" {

size_t b200fe_audio_stats_workspace(int batch) { return (size_t)(batch > 0 ? batch : 0) * sizeof(AudioStatsAcc) + 256; }

int b200fe_audio_stats(const float* wave_dev, const int64_t* offsets_dev, int64_t row_stride, const int64_t* lengths_dev,
                       int batch, int64_t max_length, float clip_level, double* out_dev, void* workspace_dev,
                       size_t workspace_bytes, void* stream) {
  if (batch == 0) return B200FE_OK;
  if (!wave_dev || !lengths_dev || !out_dev || !workspace_dev || batch < 0) return B200FE_E_INVALID;
  if (workspace_bytes < b200fe_audio_stats_workspace(batch)) return B200FE_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  AudioStatsAcc* acc = reinterpret_cast<AudioStatsAcc*>(workspace_dev);
  audio_stats_init_kernel<<<(batch + 255) / 256, 256, 0, st>>>(acc, batch);
  long long chunks = (max_length + 256 * 16 - 1) / (256 * 16);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  audio_stats_kernel<<<dim3((unsigned)chunks, batch), 256, 0, st>>>(wave_dev, (const long long*)offsets_dev,
                                                                      (const long long*)lengths_dev, row_stride, clip_level, acc);
  audio_stats_final_kernel<<<(batch + 255) / 256, 256, 0, st>>>(acc, (const long long*)lengths_dev, batch, out_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

int64_t b200fe_ingest_length(int64_t n_frames_in, int src_rate, int dst_rate) {
  if (n_frames_in < 0 || src_rate <= 0 || dst_rate <= 0) return B200FE_E_INVALID;
  if (src_rate == dst_rate) return n_frames_in;
  return (int64_t)((double)(n_frames_in * (int64_t)dst_rate) / (double)src_rate);   // int(len * 16000 / orig_sr), :1026
}

int b200fe_ingest_pcm(const void* pcm_dev, int sample_width, int channels, int64_t n_frames_in, int src_rate, int dst_rate,
                      float* out_dev, int64_t out_capacity, void* stream) {
  if (!pcm_dev || !out_dev || channels < 1 || (sample_width != 1 && sample_width != 2 && sample_width != 4))
    return B200FE_E_INVALID;
  const int64_t n_out = b200fe_ingest_length(n_frames_in, src_rate, dst_rate);
  if (n_out < 0 || n_out > out_capacity) return B200FE_E_INVALID;
  if (n_out == 0) return B200FE_OK;
  if (n_frames_in < 1) return B200FE_E_INVALID;
  // np.linspace(0, n-1, m): step = (n-1)/(m-1) (m == 1: the single point is 0)
  const double step = src_rate == dst_rate ? 1.0 : (n_out > 1 ? (double)(n_frames_in - 1) / (double)(n_out - 1) : 0.0);
  long long blocks = (n_out + 255) / 256;
  blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
  ingest_pcm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pcm_dev, sample_width, channels, n_frames_in,
                                                                         src_rate == dst_rate ? n_frames_in : n_out, step, out_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

size_t b200fe_resample_fft_workspace(int64_t n_frames_in, int src_rate, int dst_rate) {
  const int64_t n_out = b200fe_ingest_length(n_frames_in, src_rate, dst_rate);
  if (n_out <= 0 || n_frames_in <= 0) return 0;
  const int64_t m = n_out < n_frames_in ? n_out : n_frames_in;
  return (size_t)n_frames_in * sizeof(double) + (size_t)(m / 2 + 1) * sizeof(double2) + 256;
}

int b200fe_ingest_pcm_fft(const void* pcm_dev, int sample_width, int channels, int64_t n_frames_in, int src_rate,
                          int dst_rate, float* out_dev, int64_t out_capacity, void* workspace_dev, size_t workspace_bytes,
                          void* stream) {
  if (src_rate == dst_rate)
    return b200fe_ingest_pcm(pcm_dev, sample_width, channels, n_frames_in, src_rate, dst_rate, out_dev, out_capacity, stream);
  if (!pcm_dev || !out_dev || !workspace_dev || channels < 1 || (sample_width != 1 && sample_width != 2 && sample_width != 4))
    return B200FE_E_INVALID;
  const int64_t n_out = b200fe_ingest_length(n_frames_in, src_rate, dst_rate);
  if (n_out < 0 || n_out > out_capacity) return B200FE_E_INVALID;
  if (n_out == 0) return B200FE_OK;
  if (n_frames_in < 1 || workspace_bytes < b200fe_resample_fft_workspace(n_frames_in, src_rate, dst_rate)) return B200FE_E_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  double* mono = reinterpret_cast<double*>(workspace_dev);
  double2* X = reinterpret_cast<double2*>(reinterpret_cast<char*>(workspace_dev) + (((size_t)n_frames_in * sizeof(double) + 255) & ~(size_t)255));
  const int64_t m = n_out < n_frames_in ? n_out : n_frames_in;
  const int n_bins = (int)(m / 2 + 1);
  long long blocks = (n_frames_in + 255) / 256;
  blocks = blocks > 148 * 8 ? 148 * 8 : blocks;
  resample_mono_kernel<<<(unsigned)blocks, 256, 0, st>>>(pcm_dev, sample_width, channels, n_frames_in, mono);
  resample_dft_kernel<<<n_bins, 256, 0, st>>>(mono, n_frames_in, n_bins, X);
  resample_idft_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>(X, n_frames_in, n_out, out_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

size_t b200fe_ring_state_bytes(int n_streams, int capacity_samples) {
  if (n_streams <= 0 || capacity_samples <= 0) return 0;
  RingLayout lay{n_streams, capacity_samples};
  return lay.total_bytes();
}

int b200fe_ring_reset(void* state_dev, int n_streams, int capacity_samples, const int32_t* stream_ids_dev_or_null, int n,
                      void* stream) {
  if (!state_dev || n_streams <= 0 || capacity_samples <= 0 || n < 0) return B200FE_E_INVALID;
  if (n == 0) return B200FE_OK;
  RingLayout lay{n_streams, capacity_samples};
  ring_reset_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(state_dev, lay, stream_ids_dev_or_null, n);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

int b200fe_ring_push(void* state_dev, int n_streams, int capacity_samples, const float* chunks_dev, int64_t chunk_stride,
                     const int32_t* chunk_lens_dev, const int32_t* stream_ids_dev, int n, int max_chunk_samples,
                     void* stream) {
  if (!state_dev || !chunks_dev || !chunk_lens_dev || !stream_ids_dev || n_streams <= 0 || capacity_samples <= 0 || n < 0 ||
      max_chunk_samples <= 0)
    return B200FE_E_INVALID;
  if (n == 0) return B200FE_OK;
  RingLayout lay{n_streams, capacity_samples};
  int gx = (max_chunk_samples + 1023) / 1024;
  gx = gx < 1 ? 1 : (gx > 32 ? 32 : gx);
  cudaStream_t st = (cudaStream_t)stream;
  ring_push_kernel<<<dim3(gx, n), 256, 0, st>>>(state_dev, lay, chunks_dev, chunk_stride, chunk_lens_dev, stream_ids_dev,
                                                max_chunk_samples);
  ring_commit_kernel<<<(n + 255) / 256, 256, 0, st>>>(state_dev, lay, chunk_lens_dev, stream_ids_dev, n, max_chunk_samples);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

int b200fe_ring_window(const void* state_dev, int n_streams, int capacity_samples, const int32_t* stream_ids_dev, int n,
                       float* out_dev, int64_t* lens_dev, void* stream) {
  if (!state_dev || !stream_ids_dev || !out_dev || !lens_dev || n_streams <= 0 || capacity_samples <= 0 || n < 0)
    return B200FE_E_INVALID;
  if (n == 0) return B200FE_OK;
  RingLayout lay{n_streams, capacity_samples};
  int gx = (capacity_samples + 1023) / 1024;
  gx = gx < 1 ? 1 : (gx > 32 ? 32 : gx);
  ring_window_kernel<<<dim3(gx, n), 256, 0, (cudaStream_t)stream>>>(state_dev, lay, stream_ids_dev, out_dev,
                                                                     (long long*)lens_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

int b200fe_subtract_column_mean(float* feats_dev, int64_t rows_cap, int dim, const int64_t* n_rows_dev, int batch,
                                void* stream) {
  if (batch == 0) return B200FE_OK;
  if (!feats_dev || !n_rows_dev || batch < 0 || dim < 1 || rows_cap < 0) return B200FE_E_INVALID;
  column_mean_kernel<<<dim3((dim + 31) / 32, batch), 256, 0, (cudaStream_t)stream>>>(feats_dev, rows_cap, dim,
                                                                                      (const long long*)n_rows_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

}  // extern "C"
