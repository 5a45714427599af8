// TTS-side log-mel (BASELINE.json configs[4]): 24 kHz, n_fft = win = 1024, hop 256, periodic Hann, reflect padding
// (n_fft-hop)/2 per side (frames = N / hop), magnitude sqrt(re^2 + im^2 + 1e-9), 80 Slaney filters 0..12 kHz,
// log(clamp(., 1e-5)), output [B, n_mels, frames] (mel-major).  Definition frozen in oracle/tts_mel_np.py (the
// reference has no audio->mel code: parity unpinned).
//
// One 16-thread group transforms ONE frame: the 1024 real samples are packed as 512 complex points
// z[m] = x[2m] + i x[2m+1], sent through the same 512-point FFT core as the ASR front-end (fft512_columns), and the
// even/odd spectra are recombined with the W1024 twiddles; because Z[k] and Z[512-k] sit in the same thread that step
// is thread-local and yields bins k and 512-k at once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fbank_tile.cuh"
#include "fft512_complex.cuh"

namespace b200fe {

constexpr int kTtsNfft = 1024;
constexpr int kTtsFrames = 16;          // frames per tile
constexpr int kTtsBins = 513;
constexpr int kTtsPRow = 516;           // floats per frame in the magnitude buffer

struct TtsParams {
  const float* wave;
  long long wave_total;
  const long long* offsets;    // [batch] device
  const long long* lengths;    // [batch] device
  int batch;
  int hop;                     // 256
  int n_mels;
  float* mel;                  // [batch, n_mels, frames_cap]
  long long frames_cap;
  long long* mel_lens;         // [batch] or nullptr
  float mag_eps;               // 1e-9
  float log_floor;             // 1e-5
  const float* window;         // [1024] periodic Hann
  const float2* twiddle;       // [kTwTable] table 0 of the ASR front-end
  const float2* w1024;         // [16] W1024^j
  MelTab mel_tab;              // interval table over 512 bins (bin 512 carries no weight)
};

__host__ __device__ inline size_t tts_smem_bytes(int hop, int n_mels) {
  size_t b = 0;
  b += (size_t)((kTtsFrames - 1) * hop + kTtsNfft) * 4;     // staged samples (reflection resolved)
  b += kTtsNfft * 4;                                        // window
  b += (size_t)kWarps * 2 * kXGroupFloat2 * 8;              // transpose buffers / magnitude spectra
  b += (size_t)kTtsFrames * n_mels * 4;                     // log-mel tile
  b += kTwTable * 8;
  return b;
}

// 154 registers and 68 KB of shared memory per CTA: three CTAs (12 warps) per SM
__global__ void __launch_bounds__(kCtaThreads, 3)
tts_mel_kernel(const TtsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = p.hop, M = p.n_mels;
  const int ncap = (kTtsFrames - 1) * hop + kTtsNfft;
  float* xs = reinterpret_cast<float*>(smem_raw);
  float* win_s = xs + ncap;
  float2* xbuf = reinterpret_cast<float2*>(win_s + kTtsNfft);
  float* logmel_s = reinterpret_cast<float*>(xbuf + kWarps * 2 * kXGroupFloat2);
  float2* tw_s = reinterpret_cast<float2*>(logmel_s + kTtsFrames * M);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid & (kGroup - 1), g = lane >> 4;
  const int u = blockIdx.y;
  const long long N = p.lengths[u];
  const int T = (int)(N / hop);
  if (blockIdx.x == 0 && tid == 0 && p.mel_lens) p.mel_lens[u] = T;
  const int f0 = blockIdx.x * kTtsFrames;
  if (f0 >= T) return;
  const int F = min(kTtsFrames, T - f0);
  const float* x = p.wave + p.offsets[u];
  const int pad = (kTtsNfft - hop) / 2;

  for (int i = tid; i < kTtsNfft; i += kCtaThreads) win_s[i] = p.window[i];
  for (int i = tid; i < kTwTable; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  // stage with reflect padding: padded[s] = x[reflect(s - pad)]
  {
    const int n_s = (F - 1) * hop + kTtsNfft;
    const long long s0 = (long long)f0 * hop - pad;
    for (int i = tid; i < n_s; i += kCtaThreads) {
      long long s = s0 + i;
      if (s < 0) s = -s;
      if (s >= N) s = 2 * (N - 1) - s;
      s = s < 0 ? 0 : s;
      xs[i] = x[s];
    }
  }
  __syncthreads();

  float2* xg = xbuf + (warp * 2 + g) * kXGroupFloat2;
  float* pbuf = reinterpret_cast<float*>(xbuf + warp * 2 * kXGroupFloat2);   // this warp's [2 frames][kTtsPRow] magnitudes
  const float2 cj = p.w1024[j];          // W1024^j

  for (int pair = warp; 2 * pair < F; pair += kWarps) {
    const int f = 2 * pair + g;          // tile-local frame of this group
    const bool valid = f < F;
    float re[32], im[32];
    {
      const float2* xf = reinterpret_cast<const float2*>(xs + f * hop);   // hop is even: 8-byte aligned
      const float2* wf = reinterpret_cast<const float2*>(win_s);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int m = 16 * i + j;
        float2 v = make_float2(0.f, 0.f);
        if (valid) v = xf[m];
        const float2 w = wf[m];
        re[i] = v.x * w.x;
        im[i] = v.y * w.y;
      }
    }
    float ar[16], ai[16], br[16], bi[16];
    fft512_columns<32>(re, im, xg, tw_s, j, ar, ai, br, bi);

    // even/odd recombination: with s = Z[k] + conj Z[512-k] = 2E and d = Z[k] - conj Z[512-k] = 2iO,
    // 2X[k] = s + W1024^k (d/i),  2X[512-k] = conj(s - W1024^k (d/i))
    {
      const bool t0 = (j == 0);
      float* pf = pbuf + g * kTtsPRow;
#define A_RE(k) ar[bitrev<16>(k)]
#define A_IM(k) ai[bitrev<16>(k)]
#define B_RE(k) br[bitrev<16>(((k) + 1) & 15)]
#define B_IM(k) bi[bitrev<16>(((k) + 1) & 15)]
      // thread 0, slot 0 holds (Z[0], Z[256]): bin 256 pairs with itself, |X[256]| = |Z[256]|; bins 0 and 512 come from
      // the generic formula with Z[512-0] := Z[0]
      const float nyq = sqrtf(fmaf(B_RE(15), B_RE(15), B_IM(15) * B_IM(15)) + p.mag_eps);
      static_for<0, 16>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        const float ur = A_RE(i), ui = A_IM(i);
        float vr = B_RE(15 - i), vi = B_IM(15 - i);
        if (i == 0) { vr = t0 ? ur : vr; vi = t0 ? ui : vi; }
        const float sr = ur + vr, si = ui - vi;          // 2E
        const float or_ = ui + vi, oi = vr - ur;         // 2O = d / i
        // k = cA + 32 i (thread 0, i >= 8: k = 16 + 32 i);  W1024^k = W1024^c * W32^i
        constexpr float wr32 = (float)ct_cos2pi(i, 32), wi32 = (float)(-ct_sin2pi(i, 32));
        float cr = cj.x, ci = cj.y;
        if (i >= 8 && t0) { cr = (float)ct_cos2pi(16, 1024); ci = (float)(-ct_sin2pi(16, 1024)); }
        const float wr = cr * wr32 - ci * wi32, wi = cr * wi32 + ci * wr32;
        const float tr = fmaf(or_, wr, -(oi * wi)), ti = fmaf(or_, wi, oi * wr);
        const float ar2 = sr + tr, ai2 = si + ti;        // 2 X[k]
        const float br2 = sr - tr, bi2 = si - ti;        // 2 conj X[512-k]
        const float mk = sqrtf(fmaf(0.25f, fmaf(ar2, ar2, ai2 * ai2), p.mag_eps));
        const float mm = sqrtf(fmaf(0.25f, fmaf(br2, br2, bi2 * bi2), p.mag_eps));
        const int k = (i >= 8 && t0) ? 16 + 32 * i : j + 32 * i;
        pf[k] = mk;
        pf[512 - k] = mm;
      });
      if (t0) pf[256] = nyq;
#undef A_RE
#undef A_IM
#undef B_RE
#undef B_IM
    }
    __syncwarp();

    // mel over magnitudes: lane <-> interval, both frames of the warp at once
    {
      const MelTab& mel = p.mel_tab;
#pragma unroll 1
      for (int r = 0; r < mel.rounds; ++r) {
        int cnt = mel.cnt[0], base = mel.base[0];
#pragma unroll
        for (int t = 1; t < kMelRounds; ++t)
          if (r == t) { cnt = mel.cnt[t]; base = mel.base[t]; }
        const int iv = lane + 31 * r;
        const int lo = __ldg(mel.lo + 32 * r + lane) & 0xfff;   // identity lane layout (build_interval_table, bank_mod 0)
        const float2* wt = mel.w + (base * 32 + lane);
        float up0 = 0.f, up1 = 0.f, dn0 = 0.f, dn1 = 0.f;
#pragma unroll 4
        for (int q = 0; q < cnt; ++q) {
          const float2 w = __ldg(wt + 32 * q);
          const float s0 = pbuf[lo + q], s1 = pbuf[kTtsPRow + lo + q];
          up0 = fmaf(w.x, s0, up0); up1 = fmaf(w.x, s1, up1);
          dn0 = fmaf(w.y, s0, dn0); dn1 = fmaf(w.y, s1, dn1);
        }
        const float e0 = up0 + __shfl_down_sync(0xffffffffu, dn0, 1);
        const float e1 = up1 + __shfl_down_sync(0xffffffffu, dn1, 1);
        if (lane < 31 && iv < M) {
          const int fr = 2 * pair;
          if (fr < F) logmel_s[fr * M + iv] = logf(fmaxf(e0, p.log_floor));
          if (fr + 1 < F) logmel_s[(fr + 1) * M + iv] = logf(fmaxf(e1, p.log_floor));
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();

  // mel-major output: out[u][m][f0 + f]
  float* out = p.mel + (long long)u * M * p.frames_cap + f0;
  for (int idx = tid; idx < M * kTtsFrames; idx += kCtaThreads) {
    const int m = idx / kTtsFrames, f = idx - m * kTtsFrames;
    if (f < F) out[(long long)m * p.frames_cap + f] = logmel_s[f * M + m];
  }
}

// frames beyond an utterance's length are zeroed (padded batch)
__global__ void tts_pad_kernel(const long long* lengths, int hop, int M, float* mel, long long frames_cap) {
  const int u = blockIdx.y;
  const long long T = lengths[u] / hop;
  const long long per = frames_cap - T;
  if (per <= 0) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per * M; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / per, f = T + (i - m * per);
    mel[((long long)u * M + m) * frames_cap + f] = 0.f;
  }
}

}  // namespace b200fe
