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

int b200fe_subtract_column_mean(float* feats_dev, int64_t rows_cap, int dim, const int64_t* n_rows_dev, int batch,
                                void* stream) {
  if (batch == 0) return B200FE_OK;
  if (!feats_dev || !n_rows_dev || batch < 0 || dim < 1 || rows_cap < 0) return B200FE_E_INVALID;
  column_mean_kernel<<<dim3((dim + 31) / 32, batch), 256, 0, (cudaStream_t)stream>>>(feats_dev, rows_cap, dim,
                                                                                      (const long long*)n_rows_dev);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

}  // extern "C"
