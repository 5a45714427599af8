// TTS log-mel entry points of the C ABI (included at the end of b200fe.cu).
struct b200fe_tts {
  int sample_rate = 0, n_fft = 0, hop = 0, n_mels = 0;
  float2* d_twiddle = nullptr;
  float2* d_w1024 = nullptr;
  float2* d_mel_w = nullptr;
  int* d_mel_lo = nullptr;
  int mel_rounds = 0, mel_cnt[kMelRounds] = {0}, mel_base[kMelRounds] = {0};
  size_t smem = 0;
  int mel_slots = 0;            // rows of 32 weights in d_mel_w
  bool mel_fixed = false;       // the interval table has the shape the kernel is specialised for (MelShapeTts)
  int n_sms = 0, ctas_per_sm = B200FE_TTS_CTAS;
  TtsUtt* d_utts = nullptr;     // launch workspace, grown on demand: [cap] clip descriptors, [cap + 1] pair prefix sums
  int* d_pair_begin = nullptr;
  int cap = 0;
  std::mutex mu;                // one forward at a time per handle: the workspace is shared
  std::string err;
};

namespace {

double slaney_hz_to_mel(double f) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double slaney_mel_to_hz(double m) {
  const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// Slaney-scale, Slaney-normalised triangular filters in the Hz domain: dense [n_mels, n_freqs]
void build_slaney_bank(int n_freqs, double f_min, double f_max, int n_mels, int sample_rate, std::vector<float>& bank) {
  std::vector<double> f_pts(n_mels + 2);
  const double m_lo = slaney_hz_to_mel(f_min), m_hi = slaney_hz_to_mel(f_max);
  for (int i = 0; i < n_mels + 2; ++i) f_pts[i] = slaney_mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
  bank.assign((size_t)n_mels * n_freqs, 0.f);
  for (int k = 0; k < n_freqs; ++k) {
    const double f = (double)(sample_rate / 2) * k / (n_freqs - 1);
    for (int m = 0; m < n_mels; ++m) {
      const double down = (f - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - f) / (f_pts[m + 2] - f_pts[m + 1]);
      const double v = fmax(0.0, fmin(down, up)) * (2.0 / (f_pts[m + 2] - f_pts[m]));
      bank[(size_t)m * n_freqs + k] = (float)v;
    }
  }
}

thread_local std::string g_tts_error;

}  // namespace

extern "C" {

int b200fe_tts_create(int sample_rate, int n_fft, int hop, int n_mels, float f_min, float f_max, b200fe_tts** out) {
  if (!out) return B200FE_E_INVALID;
  *out = nullptr;
  auto failc = [&](int code, const std::string& m) { g_create_error = m; return code; };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return failc(B200FE_E_CUDA, "no CUDA device: no CPU fallback");
  if (n_fft != kTtsNfft) return failc(B200FE_E_UNSUPPORTED, "only n_fft = 1024 is implemented");
  if (hop <= 0 || (hop & 1) || hop > n_fft || (n_fft - hop) % 2) return failc(B200FE_E_INVALID, "hop must be even and <= n_fft");
  if (n_mels < 4 || n_mels > kMaxMels) return failc(B200FE_E_UNSUPPORTED, "n_mels must be in [4, 128]");
  if (!(f_min >= 0 && f_max > f_min && f_max <= sample_rate / 2.0f)) return failc(B200FE_E_INVALID, "bad f_min / f_max");
  b200fe_tts* t = new b200fe_tts();
  t->sample_rate = sample_rate; t->n_fft = n_fft; t->hop = hop; t->n_mels = n_mels;
  {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { delete t; return failc(B200FE_E_CUDA, "cudaGetDeviceProperties failed"); }
    t->n_sms = prop.multiProcessorCount;
  }
  std::vector<float> bank;
  build_slaney_bank(n_fft / 2 + 1, f_min, f_max, n_mels, sample_rate, bank);
  std::vector<float2> mw;
  std::vector<int> mlo;
  std::string why;
  if (build_interval_table(bank, n_mels, n_fft / 2, n_fft / 2 + 1, 1.0f, mw, mlo, t->mel_rounds, t->mel_cnt, t->mel_base, why, 16) != 0) {
    delete t;
    return failc(B200FE_E_UNSUPPORTED, why);
  }
  t->mel_slots = t->mel_base[t->mel_rounds - 1] + t->mel_cnt[t->mel_rounds - 1];
  t->smem = tts_smem_bytes(hop, t->mel_slots);
  // resident CTAs per SM: what the kernel is compiled for (registers), fewer when the sample buffers of a large hop or a
  // large mel table need the shared memory
  t->ctas_per_sm = (int)std::min<size_t>(B200FE_TTS_CTAS, (227 * 1024) / (t->smem + 1024));
  if (t->ctas_per_sm < 1) { delete t; return failc(B200FE_E_UNSUPPORTED, "hop too large for the per-warp sample buffers"); }
  t->mel_fixed = t->mel_rounds == MelShapeTts::kRounds;
  for (int r = 0; r < MelShapeTts::kRounds; ++r) t->mel_fixed = t->mel_fixed && t->mel_cnt[r] == MelShapeTts::cnt(r) && t->mel_base[r] == MelShapeTts::base(r);
  for (int m = 0; m < n_mels; ++m)
    if (bank[(size_t)m * (n_fft / 2 + 1) + n_fft / 2] != 0.f) { delete t; return failc(B200FE_E_UNSUPPORTED, "the Nyquist bin must not carry mel weight"); }
  // stage-2 twiddles + column-0 table of the packed 512-point real FFT (same tables as b200fe_create), W1024^col
  std::vector<float2> tw(kTw2Total, make_float2(0.f, 0.f)), w1024(17);
  for (int k1 = 1; k1 <= 16; ++k1)
    for (int c = 0; c < 16; ++c) {
      const int ph = (c * k1) % 512;
      const double sc = (k1 == 8 || k1 == 16) ? 2.0 : 1.0;
      tw[(k1 - 1) * kTwPitch + c] = make_float2((float)(sc * cos(2.0 * M_PI * ph / 512.0)), (float)(-sc * sin(2.0 * M_PI * ph / 512.0)));
    }
  for (int tt = 0; tt < 8; ++tt)
    for (int c = 0; c < 8; ++c) {
      const int ph = (c * tt) % 16;
      tw[kTw2Table + tt * kC0Pitch + c] = make_float2((float)(2.0 * cos(2.0 * M_PI * ph / 16.0)), (float)(-2.0 * sin(2.0 * M_PI * ph / 16.0)));
    }
  for (int jj = 0; jj <= 16; ++jj) w1024[jj] = make_float2((float)cos(2.0 * M_PI * jj / 1024.0), (float)(-sin(2.0 * M_PI * jj / 1024.0)));
  bool ok = cudaMalloc(&t->d_twiddle, tw.size() * 8) == cudaSuccess && cudaMalloc(&t->d_w1024, w1024.size() * 8) == cudaSuccess &&
            cudaMalloc(&t->d_mel_w, mw.size() * 8) == cudaSuccess && cudaMalloc(&t->d_mel_lo, mlo.size() * 4) == cudaSuccess;
  ok = ok && cudaMemcpy(t->d_twiddle, tw.data(), tw.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->d_w1024, w1024.data(), w1024.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->d_mel_w, mw.data(), mw.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(t->d_mel_lo, mlo.data(), mlo.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
  if (!ok) { b200fe_tts_destroy(t); return failc(B200FE_E_CUDA, "device allocation failed"); }
  *out = t;
  return B200FE_OK;
}

void b200fe_tts_destroy(b200fe_tts* t) {
  if (!t) return;
  cudaFree(t->d_twiddle); cudaFree(t->d_w1024); cudaFree(t->d_mel_w); cudaFree(t->d_mel_lo);
  cudaFree(t->d_utts); cudaFree(t->d_pair_begin);
  delete t;
}

int b200fe_tts_forward(b200fe_tts* t, const float* wave_dev, int64_t wave_total, const int64_t* offsets_dev,
                       const int64_t* lengths_dev, int batch, int64_t max_frames, float* mel_dev, int64_t frames_cap,
                       int64_t* mel_lens_dev, void* stream) {
  if (!t) return B200FE_E_INVALID;
  if (batch == 0) return B200FE_OK;
  if (!wave_dev || !offsets_dev || !lengths_dev || !mel_dev || batch < 0 || max_frames > frames_cap) return B200FE_E_INVALID;
  if (batch > 65535) { g_tts_error = "at most 65535 clips per b200fe_tts_forward call"; return B200FE_E_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(t->mu);
  if (batch > t->cap) {   // grow the workspace (synchronises: earlier launches may still read the old one)
    if (cudaDeviceSynchronize() != cudaSuccess) return B200FE_E_CUDA;
    cudaFree(t->d_utts); cudaFree(t->d_pair_begin);
    t->d_utts = nullptr; t->d_pair_begin = nullptr; t->cap = 0;
    const int cap = batch + batch / 2 + 16;
    if (cudaMalloc(&t->d_utts, (size_t)cap * sizeof(TtsUtt)) != cudaSuccess || cudaMalloc(&t->d_pair_begin, ((size_t)cap + 1) * 4) != cudaSuccess)
      return B200FE_E_CUDA;
    t->cap = cap;
  }
  TtsParams p;
  p.wave = wave_dev; p.wave_total = wave_total;
  p.batch = batch; p.hop = t->hop; p.n_mels = t->n_mels; p.mel = mel_dev; p.frames_cap = frames_cap;
  p.mag_eps = 1e-9f; p.log_floor = 1e-5f;
  p.twiddle = t->d_twiddle; p.w1024 = t->d_w1024;
  p.mel_tab.w = t->d_mel_w; p.mel_tab.lo = t->d_mel_lo; p.mel_tab.rounds = t->mel_rounds;
  for (int r = 0; r < kMelRounds; ++r) { p.mel_tab.cnt[r] = t->mel_cnt[r]; p.mel_tab.base[r] = t->mel_base[r]; }
  p.utts = t->d_utts; p.pair_begin = t->d_pair_begin; p.mel_slots = t->mel_slots;
  if (allow_dynamic_smem((const void*)tts_mel_kernel<MelShapeTts>, t->smem) != cudaSuccess ||
      allow_dynamic_smem((const void*)tts_mel_kernel<MelShapeRuntime>, t->smem) != cudaSuccess) return B200FE_E_CUDA;
  tts_prep_kernel<<<1, 1024, 0, st>>>((const long long*)offsets_dev, (const long long*)lengths_dev, batch, t->hop, t->d_utts,
                                      t->d_pair_begin, (long long*)mel_lens_dev);
  tts_pad_kernel<<<dim3((t->n_mels + 7) / 8, batch), 256, 0, st>>>((const long long*)lengths_dev, t->hop, t->n_mels, mel_dev, frames_cap);
  if (max_frames > 0) {
    // persistent warps: as many CTAs as stay resident, but not more warps than an upper bound of the pair count
    const long long pairs_ub = (long long)batch * ((max_frames + 1) / 2);
    const long long ctas = std::min<long long>((long long)t->n_sms * t->ctas_per_sm, (pairs_ub + kWarps - 1) / kWarps);
    const unsigned grid = (unsigned)std::max<long long>(ctas, 1);
    if (t->mel_fixed) tts_mel_kernel<MelShapeTts><<<grid, kCtaThreads, t->smem, st>>>(p);
    else tts_mel_kernel<MelShapeRuntime><<<grid, kCtaThreads, t->smem, st>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

}  // extern "C"
