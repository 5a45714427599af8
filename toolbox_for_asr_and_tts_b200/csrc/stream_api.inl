// Streaming entry points of the C ABI (included at the end of b200fe.cu).
#include <stdlib.h>
namespace {

int stream_layout(const b200fe_handle* h, int n_streams, int max_chunk, StreamLayout& lay, int& nf_max, int& e_cap,
                  size_t& smem) {
  if (!h || n_streams <= 0 || max_chunk <= 0) return B200FE_E_INVALID;
  lay.n_streams = n_streams;
  lay.carry_cap = (h->L + 3) & ~3;
  lay.cache_cap = h->cfg.lfr_m - 1 > 1 ? h->cfg.lfr_m - 1 : 1;
  lay.n_mels = h->cfg.n_mels;
  nf_max = (max_chunk - 1) / h->S + 1;
  lay.frames_cap = lay.cache_cap + nf_max;
  lay.q_max = (nf_max + 3) / 4;
  e_cap = ((h->L - 1 + max_chunk + 8) + 3) & ~3;
  smem = stream_smem_bytes(e_cap, warp_kernel_fits(h->L, h->S));
  return 0;
}

template <int NROWS, bool EXACT, class MELS>
int launch_stream(b200fe_handle* h, const StreamParams& p, size_t smem, bool dither, cudaStream_t st) {
#define LAUNCHS(DI, PQ)                                                                               \
  do {                                                                                                \
    auto k = stream_push_kernel<NROWS, EXACT, DI, MELS, PQ>;                                          \
    CUDA_TRY(h, allow_dynamic_smem((const void*)k, smem));                                            \
    k<<<p.n, kCtaThreads, smem, st>>>(p);                                                             \
  } while (0)
#define LAUNCHQ(DI)                                                                                   \
  do {                                                                                                \
    auto k = stream_quad_kernel<NROWS, EXACT, DI, MELS>;                                              \
    CUDA_TRY(h, allow_dynamic_smem((const void*)k, smem));                                            \
    stream_tick_prep_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p);                                    \
    const long long items = (long long)p.n * p.lay.q_max;                                             \
    const long long ctas = std::min<long long>(4ll * h->n_sms, (items + kWarps - 1) / kWarps);        \
    k<<<(unsigned)ctas, kCtaThreads, smem, st>>>(p);                                                  \
    stream_tick_finish_kernel<<<(p.n + kWarps - 1) / kWarps, kCtaThreads, 0, st>>>(p);                \
  } while (0)
  // One CTA per stream is the shipped tick.  B200FE_STREAM_KERNEL=quad selects the quad-level alternative (three
  // launches: descriptors, work items = (chunk, quad) on persistent warps, state update), which needs a quad to fit the
  // warp buffer.  Measured on B200: 52 us against 46 us per tick of 512 streams - the tick is too
  // small (19 us of quads) for the flat work list to pay for its two extra launches.
  const bool per_quad = warp_kernel_fits(h->L, h->S);
  const char* force = getenv("B200FE_STREAM_KERNEL");
  const bool quad_level = per_quad && force && force[0] == 'q';
#ifdef B200FE_BENCH_ONLY
  if (!per_quad || dither) return fail(h, B200FE_E_UNSUPPORTED, "bench-only build");
  if (quad_level) LAUNCHQ(false); else LAUNCHS(false, true);
#else
  if (quad_level) { if (dither) LAUNCHQ(true); else LAUNCHQ(false); }
  else if (per_quad) { if (dither) LAUNCHS(true, true); else LAUNCHS(false, true); }
  else          { if (dither) LAUNCHS(true, false); else LAUNCHS(false, false); }
#endif
#undef LAUNCHQ
#undef LAUNCHS
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return 0;
}

}  // namespace

extern "C" {

int b200fe_stream_state_bytes(const b200fe_handle* h, int n_streams, int max_chunk_samples, size_t* bytes) {
  StreamLayout lay; int nf_max, e_cap; size_t smem;
  int rc = stream_layout(h, n_streams, max_chunk_samples, lay, nf_max, e_cap, smem);
  if (rc) return rc;
  if (smem > 227 * 1024) return fail(h, B200FE_E_UNSUPPORTED, "max_chunk_samples too large for one shared-memory tile");
  if (bytes) *bytes = lay.total_bytes();
  return B200FE_OK;
}

int b200fe_stream_max_rows(const b200fe_handle* h, int max_chunk_samples) {
  if (!h || max_chunk_samples <= 0) return B200FE_E_INVALID;
  const int nf_max = (max_chunk_samples - 1) / h->S + 1;
  return (nf_max + h->cfg.lfr_m + h->cfg.lfr_n - 1) / h->cfg.lfr_n + 1;
}

int b200fe_stream_reset(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples,
                        const int32_t* stream_ids_dev_or_null, int n, void* stream) {
  StreamLayout lay; int nf_max, e_cap; size_t smem;
  int rc = stream_layout(h, n_streams, max_chunk_samples, lay, nf_max, e_cap, smem);
  if (rc) return rc;
  if (!state_dev) return fail(h, B200FE_E_INVALID, "null state");
  std::lock_guard<std::mutex> lock(h->mu);     // launches / err are shared with the other entry points
  const int cnt = stream_ids_dev_or_null ? n : n_streams;
  if (cnt <= 0) return B200FE_OK;
  stream_reset_kernel<<<(cnt + 255) / 256, 256, 0, (cudaStream_t)stream>>>(state_dev, lay, stream_ids_dev_or_null, cnt);
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return B200FE_OK;
}

int b200fe_stream_push(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples, const float* chunks_dev,
                       int64_t chunk_stride, const int32_t* chunk_lens_dev, const int32_t* stream_ids_dev,
                       const uint8_t* is_final_dev, int n, float* feats_dev, int64_t rows_cap, int32_t* rows_out_dev,
                       void* stream) {
  return b200fe_stream_push_stats(h, state_dev, n_streams, max_chunk_samples, chunks_dev, chunk_stride, chunk_lens_dev,
                                  stream_ids_dev, is_final_dev, n, feats_dev, rows_cap, rows_out_dev, nullptr, stream);
}

int b200fe_stream_push_stats(b200fe_handle* h, void* state_dev, int n_streams, int max_chunk_samples,
                             const float* chunks_dev, int64_t chunk_stride, const int32_t* chunk_lens_dev,
                             const int32_t* stream_ids_dev, const uint8_t* is_final_dev, int n, float* feats_dev,
                             int64_t rows_cap, int32_t* rows_out_dev, float* chunk_stats_dev, void* stream) {
  StreamLayout lay; int nf_max, e_cap; size_t smem;
  int rc = stream_layout(h, n_streams, max_chunk_samples, lay, nf_max, e_cap, smem);
  if (rc) return rc;
  if (n == 0) return B200FE_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!state_dev || !chunks_dev || !chunk_lens_dev || !stream_ids_dev || !feats_dev || !rows_out_dev || n < 0)
    return fail(h, B200FE_E_INVALID, "null argument");
  if (smem > 227 * 1024) return fail(h, B200FE_E_UNSUPPORTED, "max_chunk_samples too large for one shared-memory tile");
  if (rows_cap < b200fe_stream_max_rows(h, max_chunk_samples)) return fail(h, B200FE_E_INVALID, "rows_cap < b200fe_stream_max_rows");
  StreamParams p;
  p.state = state_dev; p.lay = lay; p.chunks = chunks_dev; p.chunk_stride = chunk_stride; p.chunk_lens = chunk_lens_dev;
  p.stream_ids = stream_ids_dev; p.is_final = is_final_dev; p.n = n; p.max_chunk = max_chunk_samples; p.nf_max = nf_max;
  p.feats = feats_dev; p.rows_cap = rows_cap; p.rows_out = rows_out_dev; p.chunk_stats = chunk_stats_dev;
  p.frame_len = h->L; p.frame_shift = h->S; p.n_mels = h->cfg.n_mels; p.lfr_m = h->cfg.lfr_m; p.lfr_n = h->cfg.lfr_n;
  p.e_cap = e_cap; p.preemph = h->cfg.preemphasis; p.remove_dc = h->cfg.remove_dc_offset; p.log_floor = h->cfg.log_floor;
  p.dither = h->cfg.dither / (h->cfg.upscale_samples ? 32768.f : 1.f); p.seed = 0;
  p.window = h->d_window; p.twiddle = h->d_twiddle_stream; p.mel_w = h->d_mel_w; p.mel_lo = h->d_mel_lo; p.mel_rounds = h->mel_rounds;
  for (int r = 0; r < kMelRounds; ++r) { p.mel_cnt[r] = h->mel_cnt[r]; p.mel_base[r] = h->mel_base[r]; } p.cmvn = h->d_cmvn;
  const bool dither = h->cfg.dither != 0.f;
  if (h->L == 400 && h->mel_paraformer) return launch_stream<25, true, MelShapeParaformer>(h, p, smem, dither, (cudaStream_t)stream);
#ifdef B200FE_BENCH_ONLY
  return fail(h, B200FE_E_UNSUPPORTED, "bench-only build");
#endif
  if (h->L == 400) return launch_stream<25, true, MelShapeRuntime>(h, p, smem, dither, (cudaStream_t)stream);
  return launch_stream<32, false, MelShapeRuntime>(h, p, smem, dither, (cudaStream_t)stream);
}

}  // extern "C"
