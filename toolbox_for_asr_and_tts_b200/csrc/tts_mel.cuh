// TTS-side log-mel (BASELINE.json configs[4]): 24 kHz, n_fft = win = 1024, hop 256, periodic Hann, reflect padding
// (n_fft-hop)/2 per side (frames = N / hop), magnitude sqrt(re^2 + im^2 + 1e-9), 80 Slaney filters 0..12 kHz,
// log(clamp(., 1e-5)), output [B, n_mels, frames] (mel-major).  Definition frozen in oracle/tts_mel_np.py (the
// reference has no audio->mel code: parity unpinned).
//
// Round 2: on the packed f32x2 real-FFT core of the ASR front-end, and warp-autonomous like fbank_warp_kernel.
//   * FFT: a 1024-point real FFT is two 512-point real FFTs, of the even and of the odd samples:
//     X[k] = E[k] + W1024^k O[k].  That is exactly what the ASR core transforms at once - two real 512-point sequences,
//     one per lane of the packed registers - so a 16-thread group runs ONE TTS frame with lane .x = even samples,
//     lane .y = odd samples (one 64-bit shared load fetches both), through the same real 32-point stage, transpose,
//     twiddle and 16-point stage (quad_stage2's data flow), and the recombination is thread-local because E[k] and
//     O[k] sit in the two lanes of one register:  2 X[k] = 2E + W (2O),  2 X[512-k] = conj(2E - W (2O)).
//     The periodic Hann window needs no table: w[32 i + 2 j (+1)] follows from cos(a_i + b), a_i a compile-time constant.
//   * work: one PAIR of consecutive frames (n_fft + hop samples) per warp and iteration; persistent warps stride over
//     the launch-wide pair list (tts_prep_kernel: prefix sums of ceil(T/2)), so neighbouring warps work on neighbouring
//     pairs and share their 768 overlapping samples through L2.  The pair's samples arrive by ONE bulk copy
//     (cp.async.bulk, SASS UBLKCP) into the warp's buffer, issued a whole pair ahead and completed on the warp's own
//     mbarrier; pairs that touch the reflected ends of a clip (or a clip that does not start on an 8-byte boundary)
//     are filled by the lanes.  No CTA-wide barrier in the loop.
//   * output: the lane that holds filter m of the pair's two frames stores them as one 8-byte word of row m.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fbank_tile.cuh"
#include "fbank_warp.cuh"

#ifndef B200FE_TTS_CTAS      // resident CTAs per SM the TTS kernel is compiled and launched for
#define B200FE_TTS_CTAS 3
#endif

namespace b200fe {

constexpr int kTtsNfft = 1024;
constexpr int kTtsBins = 513;
constexpr int kTtsMagF2 = 520;          // float2 (frame of group 0, frame of group 1) per bin, padded

struct __align__(16) TtsUtt {   // built by tts_prep_kernel
  long long off;      // first sample of the clip inside the wave buffer
  long long N;        // samples
  int T;              // frames = N / hop
  int pair_begin;     // index of the clip's first pair in the launch-wide pair list
  int r0, r1;
};

struct TtsParams {
  const float* wave;
  long long wave_total;
  int batch;
  int hop;                     // 256
  int n_mels;
  float* mel;                  // [batch, n_mels, frames_cap]
  long long frames_cap;
  float mag_eps;               // 1e-9
  float log_floor;             // 1e-5
  const float2* twiddle;       // [kTw2Total]: the ASR front-end's stage-2 twiddles + column-0 table (TileParams::twiddle)
  const float2* w1024;         // [17] W1024^col, col = 0..16
  MelTab mel_tab;              // interval table over 512 bins (bin 512 carries no weight), bank-matched lanes
  int mel_slots;               // rows of 32 (up, down) weights in mel_tab.w: copied into shared memory by every CTA
  const TtsUtt* utts;          // [batch]
  const int* pair_begin;       // [batch + 1]: exclusive prefix sums of ceil(T / 2); the last entry is the pair count
};

// floats of a warp's sample buffer: a pair's n_fft + hop samples, placed up to 2 floats behind a 16-byte boundary
__host__ __device__ inline int tts_buf_floats(int hop) { return ((hop + kTtsNfft + 3) & ~3) + 4; }
constexpr int kTtsThreadConsts = 4;   // float2 per thread: Hann angle-addition pair (cos b, sin b), W1024^col, W32^t0
__host__ __device__ inline size_t tts_smem_bytes(int hop, int mel_slots) {
  return (size_t)kWarps * tts_buf_floats(hop) * 4 + (size_t)kWarps * kYWarpF4 * 16 + (size_t)kTw2Total * 8 + (size_t)kWarps * 8 +
         (size_t)kCtaThreads * kTtsThreadConsts * 8 + (size_t)mel_slots * 32 * 8;
}
using MelShapeTts = MelShapeFixed<3, 3, 10, 21>;   // 80 Slaney filters 0..12 kHz over the 513 bins of a 1024-point FFT at 24 kHz

// One block: frames, pair prefix sums and clip descriptors of the batch; also the frame counts handed to the caller.
__global__ void __launch_bounds__(1024)
tts_prep_kernel(const long long* offsets, const long long* lengths, int batch, int hop, TtsUtt* utts, int* pair_begin,
                long long* mel_lens) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry = 0;
  for (int base = 0; base < batch; base += 1024) {
    const int u = base + threadIdx.x;
    const long long N = u < batch ? lengths[u] : 0;
    const int T = (int)(N / hop);
    const int v = (T + 1) >> 1;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int excl = carry + (warp ? warp_sums[warp - 1] : 0) + x - v;
    if (u < batch) {
      pair_begin[u] = excl;
      TtsUtt d;
      d.off = offsets[u]; d.N = N; d.T = T; d.pair_begin = excl; d.r0 = d.r1 = 0;
      utts[u] = d;
      if (mel_lens) mel_lens[u] = T;
    }
    carry += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) pair_begin[batch] = carry;
}

__device__ __forceinline__ float fast_sqrt(float x) {   // x >= mag_eps > 0
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Stage 2 of one group's (even, odd) pair, quad_stage2's data flow with complex outputs kept: on return ar/ai[k2] hold
// 2 x the 512-point spectra (lane .x: even samples, lane .y: odd samples) at bin col + 32 k2, and (c0r, c0i) the same at
// bin 32 t0 for the lane's own group; u_alt = sum_c (-1)^c Y_c[0] = bin 256 of both sequences (unscaled, real).
__device__ __forceinline__ void tts_stage2(const f2 (&zr)[16], const f2 (&zi)[16], f2 y0, f2 y16, float4* yg,
                                           const float2* tw_row, const float2* c0_row, int j, int grp_in_warp,
                                           f2 (&ar)[16], f2 (&ai)[16], f2& c0r, f2& c0i, f2& u_alt) {
  {
    float2* yg2 = reinterpret_cast<float2*>(yg) + 2 * j;
    const int hr = j >> 3, hi = hr ^ 1;
    static_for<1, 16>([&](auto ic) {
      constexpr int k1 = decltype(ic)::value;
      yg2[(k1 - 1) * 2 * kYPitch + hr] = zr[k1];
      yg2[(k1 - 1) * 2 * kYPitch + hi] = zi[k1];
    });
    yg2[15 * 2 * kYPitch + hr] = y16;
    yg2[15 * 2 * kYPitch + hi] = make_float2(0.f, 0.f);
  }
  reinterpret_cast<float2*>(yg + 16 * kYPitch)[j] = y0;
  __syncwarp();
  const int lane = (grp_in_warp << 4) | j;
  const int c = lane >> 1, g2 = lane & 1;
  const int col = c == 0 ? 16 : c;
  float4* const ywarp = yg - grp_in_warp * kYGroupF4;
  {
    const float4* rowp = ywarp + g2 * kYGroupF4 + (col - 1) * kYPitch;
    static_for<0, 8>([&](auto ic) {
      constexpr int h = decltype(ic)::value;
      const float2 ta = tw_row[2 * h], tb = tw_row[2 * h + 1];
      const float4 v0 = rowp[2 * h], v1 = rowp[2 * h + 1];
      constexpr bool sw = 2 * h >= 8;   // slots 8..15 are stored (im, re)
      {
        const f2 yr = sw ? make_float2(v0.z, v0.w) : make_float2(v0.x, v0.y);
        const f2 yi = sw ? make_float2(v0.x, v0.y) : make_float2(v0.z, v0.w);
        ar[bitrev<16>(2 * h)] = fma2s(yr, ta.x, neg2(mul2s(yi, ta.y)));
        ai[bitrev<16>(2 * h)] = fma2s(yr, ta.y, mul2s(yi, ta.x));
      }
      {
        const f2 yr = sw ? make_float2(v1.z, v1.w) : make_float2(v1.x, v1.y);
        const f2 yi = sw ? make_float2(v1.x, v1.y) : make_float2(v1.z, v1.w);
        ar[bitrev<16>(2 * h + 1)] = fma2s(yr, tb.x, neg2(mul2s(yi, tb.y)));
        ai[bitrev<16>(2 * h + 1)] = fma2s(yr, tb.y, mul2s(yi, tb.x));
      }
    });
  }
  // column 0 of the lane's OWN group: bin 32 t0 (table entries carry the factor 2) and the alternating sum (bin 256)
  {
    const int t0 = (lane >> 1) & 7;
    const float4* u4 = reinterpret_cast<const float4*>(yg + 16 * kYPitch);   // 16 x (even, odd)
    const float4* w4 = reinterpret_cast<const float4*>(c0_row);
    const float sgn = (t0 & 1) ? -1.f : 1.f;
    f2 cr = make_float2(0.f, 0.f), ci = make_float2(0.f, 0.f), alt = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float4 ua = u4[h], ub = u4[h + 4], w = w4[h];
      const f2 v0 = fma2s(make_float2(ub.x, ub.y), sgn, make_float2(ua.x, ua.y));
      const f2 v1 = fma2s(make_float2(ub.z, ub.w), sgn, make_float2(ua.z, ua.w));
      cr = fma2s(v0, w.x, cr);
      ci = fma2s(v0, w.y, ci);
      cr = fma2s(v1, w.z, cr);
      ci = fma2s(v1, w.w, ci);
      alt = add2(alt, sub2(add2(make_float2(ua.x, ua.y), make_float2(ub.x, ub.y)), add2(make_float2(ua.z, ua.w), make_float2(ub.z, ub.w))));
    }
    c0r = cr; c0i = ci; u_alt = alt;
  }
  fft_dit2<16>(ar, ai);
  __syncwarp();   // every lane has consumed the transpose buffers: the magnitudes may overwrite them
}

// One round of the interval mel over the pair's magnitudes (float2: frame of group 0, frame of group 1); the lane that
// holds filter iv stores (frame 2 pair, frame 2 pair + 1) of row iv.  CNT >= 0: compile-time trip count, fully unrolled.
template <int CNT>
__device__ __forceinline__ void tts_mel_round(const MelTab& mel, int r, int cnt_rt, int base, const float2* mag, int lane,
                                              float log_floor, float* out_u, long long frames_cap, bool both, bool second) {
  const int cnt = CNT >= 0 ? CNT : cnt_rt;
  const unsigned word = (unsigned)__ldg(mel.lo + 32 * r + lane);
  const int lo = (int)(word & 0xfffu), partner = (int)((word >> 12) & 31u), iv = (int)((word >> 17) & 0xffu);
  const float2* wt = mel.w + (base * 32 + lane);
  const float2* p0 = mag + lo;
  f2 up = make_float2(0.f, 0.f), dn = make_float2(0.f, 0.f);
  auto body = [&](int q) {
    const float2 wq = wt[32 * q];     // shared memory (tts_mel_kernel copies the table)
    const float2 sv = p0[q];
    up = fma2s(sv, wq.x, up);
    dn = fma2s(sv, wq.y, dn);
  };
  if constexpr (CNT >= 0) {
#pragma unroll
    for (int q = 0; q < CNT; ++q) body(q);
  } else {
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) body(q);
  }
  const float e0 = up.x + __shfl_sync(0xffffffffu, dn.x, partner);
  const float e1 = up.y + __shfl_sync(0xffffffffu, dn.y, partner);
  if (word >> 31) {
    const float l0 = fast_ln(fmaxf(e0, log_floor)), l1 = fast_ln(fmaxf(e1, log_floor));
    float* o = out_u + (long long)iv * frames_cap;
    if (both) {
      *reinterpret_cast<float2*>(o) = make_float2(l0, l1);
    } else {
      o[0] = l0;
      if (second) o[1] = l1;
    }
  }
}

// Lanes of a warp: the samples of pair `pair` of a clip into buf (buf[i] = padded[2 pair hop + i], reflection resolved)
__device__ __forceinline__ void tts_fill_generic(const float* x, long long N, long long s0, int n, int lane, float* buf) {
  for (int i = lane; i < n; i += 32) {
    long long s = s0 + i;
    if (s < 0) s = -s;
    if (s >= N) s = 2 * (N - 1) - s;
    s = s < 0 ? 0 : (s >= N ? N - 1 : s);
    buf[i] = __ldg(x + s);
  }
}

// Where the bulk copy of a pair's samples starts (absolute index in the wave buffer) when the pair lies inside the clip
// and starts on an 8-byte boundary; -1 when the lanes have to fill the buffer (reflected ends, odd offsets).
__device__ __forceinline__ long long tts_tma_source(const TtsParams& p, const TtsUtt& ut, int pair) {
  const int pad = (kTtsNfft - p.hop) / 2, n = p.hop + kTtsNfft;
  const long long s0 = 2ll * pair * p.hop - pad;
  if (s0 < 0 || s0 + n > ut.N) return -1;
  const long long g0 = ut.off + s0;
  return (quad_a_off<float>(p.wave, g0) & 1) ? -1 : g0;
}
// The whole warp: issue the copy (lane 0).  Returns 1 + the buffer offset (0 or 2 floats) of the pair's first sample,
// or 0 when there is no copy in flight.
__device__ __forceinline__ int tts_issue_copy(const TtsParams& p, long long g0, int lane, float* buf, unsigned long long* bar) {
  if (g0 < 0) return 0;
  int ok = 0;
  if (lane == 0) ok = quad_fill_tma(p.wave, p.wave_total, g0, p.hop + kTtsNfft, buf, bar) ? 1 : 0;
  ok = __shfl_sync(0xffffffffu, ok, 0);
  return ok ? 1 + quad_a_off<float>(p.wave, g0) : 0;
}

template <class MELS>
__global__ void __launch_bounds__(kCtaThreads, B200FE_TTS_CTAS)
tts_mel_kernel(const TtsParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int hop = p.hop;
  const int nbuf = tts_buf_floats(hop);
  float* bufs = reinterpret_cast<float*>(smem_raw);
  float4* xbuf = reinterpret_cast<float4*>(bufs + kWarps * nbuf);
  float2* tw_s = reinterpret_cast<float2*>(xbuf + kWarps * kYWarpF4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tw_s + kTw2Total);
  // per-thread constants live in shared memory and are re-read where they are used (volatile): kept in registers across
  // the loop they are spilled, and the local-memory footprint of an SM does not stay in L1
  volatile float2* tc = reinterpret_cast<volatile float2*>(bars + kWarps) + threadIdx.x;
  float2* melw_s = reinterpret_cast<float2*>(bars + kWarps) + kCtaThreads * kTtsThreadConsts;   // the mel weights

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid & (kGroup - 1), g = lane >> 4;
  const int c = lane >> 1, g2 = lane & 1, col = c == 0 ? 16 : c, t0 = (lane >> 1) & 7;

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  for (int i = tid; i < p.mel_slots * 32; i += kCtaThreads) melw_s[i] = p.mel_tab.w[i];
  if (lane == 0) {
    mbar_init(bars + warp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // periodic Hann of this thread's samples n = 32 i + 2 j (+ 1): cos(2 pi n / 1024) = cos(a_i) cos(b) - sin(a_i) sin(b)
  {
    float s0, c0, s1, c1, sn, cs;
    sincospif(2.0f * (float)(2 * j) / 1024.0f, &s0, &c0);
    sincospif(2.0f * (float)(2 * j + 1) / 1024.0f, &s1, &c1);
    sincospif(2.0f * (float)t0 / 32.0f, &sn, &cs);
    const float2 wc = __ldg(p.w1024 + col);                  // W1024^col = (cos, -sin)
    tc[0 * kCtaThreads].x = c0; tc[0 * kCtaThreads].y = c1;   // cos b (even sample, odd sample)
    tc[1 * kCtaThreads].x = s0; tc[1 * kCtaThreads].y = s1;   // sin b
    tc[2 * kCtaThreads].x = wc.x; tc[2 * kCtaThreads].y = wc.y;
    tc[3 * kCtaThreads].x = cs; tc[3 * kCtaThreads].y = -sn;  // W32^t0 for the column-0 bins
  }
  __syncthreads();   // the only CTA-wide barrier: the twiddle tables and the mbarriers

  float* buf = bufs + warp * nbuf;
  unsigned long long* bar = bars + warp;
  float4* yg = xbuf + warp * kYWarpF4 + g * kYGroupF4;
  float2* mag = reinterpret_cast<float2*>(xbuf + warp * kYWarpF4);   // this warp's magnitudes (frame of group 0, of group 1)
  const float2* tw_row = fft_twiddle_row<32>(tw_s, j, g);
  const float2* c0_row = fft_c0_row(tw_s, j);

  // Work distribution: warp w of the grid takes pairs w, w + W, w + 2 W, ...  The clip of a pair is found by walking
  // the prefix sums forward, 32 entries per step (one load per lane and a ballot).
  const int W = gridDim.x * kWarps;
  const int total = __ldg(p.pair_begin + p.batch);
  int w = blockIdx.x * kWarps + warp;
  if (w >= total) return;
  int u = 0, pb_u = 0;
  auto seek = [&](int wq, int& uu, int& pb) {   // uu, pb: clip with pair_begin <= wq < next pair_begin
    while (true) {
      const int idx = uu + 1 + lane;
      const int v = __ldg(p.pair_begin + min(idx, p.batch));
      const unsigned m = __ballot_sync(0xffffffffu, idx <= p.batch && wq >= v);
      const int n = __popc(m);
      if (n) pb = __shfl_sync(0xffffffffu, v, n - 1);
      uu += n;
      if (n < 32) break;
    }
  };
  seek(w, u, pb_u);
  int in_flight;
  {
    const TtsUtt ut = p.utts[u];
    in_flight = tts_issue_copy(p, tts_tma_source(p, ut, w - pb_u), lane, buf, bar);
  }
  unsigned phase = 0;
  const int pad = (kTtsNfft - hop) / 2;
  const bool cap_even = (p.frames_cap & 1) == 0;

  while (true) {
    const TtsUtt ut = p.utts[u];
    const int pair = w - pb_u;
    // the next pair of this warp: the first window of prefix sums is loaded now and looked at after stage 1
    const int wn = w + W;
    const bool have_next = wn < total;
    const int v_next = __ldg(p.pair_begin + min(u + 1 + lane, p.batch));

    int un = u, pbn = pb_u;
    TtsUtt utn = ut;
    int a_off = 0;
    if (in_flight) {
      mbar_wait(bar, phase);
      phase ^= 1u;
      a_off = in_flight - 1;
    } else {
      tts_fill_generic(p.wave + ut.off, ut.N, 2ll * pair * hop - pad, hop + kTtsNfft, lane, buf);
      __syncwarp();
    }
    f2 ar[16], ai[16], c0r, c0i, u_alt;
    {
      f2 zr[16], zi[16], y0, y16;
      {
        f2 y[32];
        const float2* xf = reinterpret_cast<const float2*>(buf + a_off + g * hop) + j;   // hop, a_off even: 8-byte aligned
        const f2 cb = make_float2(tc[0].x, tc[0].y), sb = make_float2(tc[kCtaThreads].x, tc[kCtaThreads].y);
        static_for<0, 32>([&](auto ic) {
          constexpr int i = decltype(ic)::value;
          constexpr float ca = (float)(0.5 * ct_cos2pi(i, 32)), sa = (float)(0.5 * ct_sin2pi(i, 32));
          const f2 wv = fma2s(sb, sa, fma2s(cb, -ca, make_float2(0.5f, 0.5f)));
          y[i] = mul2(xf[16 * i], wv);
        });
        static_for<0, 16>([&](auto ic) {
          constexpr int m = decltype(ic)::value;
          zr[bitrev<16>(m)] = y[2 * m];
          zi[bitrev<16>(m)] = y[2 * m + 1];
        });
      }
      __syncwarp();   // every lane holds its samples (the buffer is free for the next pair's copy) and is done with the
                      // previous pair's magnitudes, which alias the transpose buffers
      fft_dit2<16>(zr, zi);
      real32_split2(zr, zi, y0, y16);
      // clip of the next pair (its descriptor load overlaps stage 2)
      if (have_next) {
        const unsigned mm = __ballot_sync(0xffffffffu, u + 1 + lane <= p.batch && wn >= v_next);
        const int n = __popc(mm);
        if (n) pbn = __shfl_sync(0xffffffffu, v_next, n - 1);
        un = u + n;
        if (n == 32) seek(wn, un, pbn);
        utn = p.utts[un];
      }
      tts_stage2(zr, zi, y0, y16, yg, tw_row, c0_row, j, g, ar, ai, c0r, c0i, u_alt);
      // the next pair's samples start their way from HBM: the buffer has been free since the barrier above
      in_flight = tts_issue_copy(p, have_next ? tts_tma_source(p, utn, wn - pbn) : -1, lane, buf, bar);
    }
    // ---- recombination: lane .x = 2E[k], lane .y = 2O[k] at k = col + 32 k2;  2X[k] = 2E + W^k 2O,
    //      2X[512-k] = conj(2E - W^k 2O);  magnitudes into component g2 of the warp's buffer
    {
      float* mg = reinterpret_cast<float*>(mag) + g2;
      const float2 wcol = make_float2(tc[2 * kCtaThreads].x, tc[2 * kCtaThreads].y);
      static_for<0, 16>([&](auto ic) {
        constexpr int k2 = decltype(ic)::value;
        constexpr float c32 = (float)ct_cos2pi(k2, 32), s32 = (float)(-ct_sin2pi(k2, 32));   // W32^k2
        const float wr = wcol.x * c32 - wcol.y * s32, wi = wcol.x * s32 + wcol.y * c32;      // W1024^(col + 32 k2)
        const float er = ar[k2].x, ei = ai[k2].x, orr = ar[k2].y, oi = ai[k2].y;
        const float tr = fmaf(orr, wr, -(oi * wi)), ti = fmaf(orr, wi, oi * wr);
        const float pr = er + tr, pi = ei + ti, qr = er - tr, qi = ei - ti;
        const int k = col + 32 * k2;
        mg[2 * k] = fast_sqrt(fmaf(0.25f, fmaf(pr, pr, pi * pi), p.mag_eps));
        if (c != 0) mg[2 * (512 - k)] = fast_sqrt(fmaf(0.25f, fmaf(qr, qr, qi * qi), p.mag_eps));
      });
      if ((lane & 1) == 0) {   // column 0 of the lane's own group: bins 32 t0 and 512 - 32 t0, and bin 256
        float* mo = reinterpret_cast<float*>(mag) + g;
        const float w32r = tc[3 * kCtaThreads].x, w32i = tc[3 * kCtaThreads].y;
        const float er = c0r.x, ei = c0i.x, orr = c0r.y, oi = c0i.y;
        const float tr = fmaf(orr, w32r, -(oi * w32i)), ti = fmaf(orr, w32i, oi * w32r);
        const float pr = er + tr, pi = ei + ti, qr = er - tr, qi = ei - ti;
        mo[2 * (32 * t0)] = fast_sqrt(fmaf(0.25f, fmaf(pr, pr, pi * pi), p.mag_eps));
        mo[2 * (512 - 32 * t0)] = fast_sqrt(fmaf(0.25f, fmaf(qr, qr, qi * qi), p.mag_eps));
        if (t0 == 0) mo[2 * 256] = fast_sqrt(fmaf(u_alt.x, u_alt.x, u_alt.y * u_alt.y) + p.mag_eps);   // E[256] - i O[256]
      }
    }
    __syncwarp();

    // ---- mel over magnitudes: lane <-> interval (bank-matched lanes, as in the ASR mel stage), both frames at once;
    //      the lane that holds filter iv stores (frame 2 pair, frame 2 pair + 1) of row iv
    {
      MelTab mel = p.mel_tab;
      mel.w = melw_s;
      float* out_u = p.mel + (long long)u * p.n_mels * p.frames_cap + 2 * pair;
      const bool second = 2 * pair + 1 < ut.T, both = cap_even && second;
      if constexpr (MELS::kFixed) {
        static_for<0, MELS::kRounds>([&](auto ic) {
          constexpr int r = decltype(ic)::value;
          tts_mel_round<MELS::cnt(r)>(mel, r, 0, MELS::base(r), mag, lane, p.log_floor, out_u, p.frames_cap, both, second);
        });
      } else {
#pragma unroll 1
        for (int r = 0; r < mel.rounds; ++r) {
          int cnt = mel.cnt[0], base = mel.base[0];
#pragma unroll
          for (int t = 1; t < kMelRounds; ++t)
            if (r == t) { cnt = mel.cnt[t]; base = mel.base[t]; }
          tts_mel_round<-1>(mel, r, cnt, base, mag, lane, p.log_floor, out_u, p.frames_cap, both, second);
        }
      }
    }
    if (!have_next) break;
    w = wn; u = un; pb_u = pbn;
  }
}

// frames beyond a clip's length are zeroed (padded batch): one warp per (clip, mel row), 128-bit stores on the 16-byte grid
__global__ void __launch_bounds__(256)
tts_pad_kernel(const long long* lengths, int hop, int M, float* mel, long long frames_cap) {
  const int u = blockIdx.y, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  const long long T = lengths[u] / hop;
  const long long n = frames_cap - T;
  if (n <= 0) return;
  float* row = mel + ((long long)u * M + m) * frames_cap + T;
  const long long head = min(n, (long long)((4 - ((reinterpret_cast<uintptr_t>(row) >> 2) & 3)) & 3));
  const long long n4 = (n - head) >> 2;
  if (lane < head) row[lane] = 0.f;
  float4* row4 = reinterpret_cast<float4*>(row + head);
  for (long long i = lane; i < n4; i += 32) row4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < n - head - 4 * n4) row[head + 4 * n4 + lane] = 0.f;
}

}  // namespace b200fe
