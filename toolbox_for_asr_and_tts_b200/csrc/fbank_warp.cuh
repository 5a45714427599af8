// Warp-autonomous front-end kernel for sm_100a: every warp owns one QUAD (4 consecutive frames of one utterance) at a
// time, from HBM samples to the LFR-stacked, CMVN-normalised output rows, with no CTA-wide barrier in the loop.
//
//   * work list  = QuadDesc[n_quads], built on the device (build_quads_kernel); persistent warps stride over it, so
//                  the 4 warps of a CTA work on neighbouring quads and share their 240 overlapping samples through L1.
//   * samples    = the quad's 880 samples are prefetched into L2 one quad AHEAD (one line per lane), then read with
//                  128-bit loads, pre-emphasised in registers and stored to the warp's private buffer; the other 15
//                  warps of the SM cover the L2 latency.
//   * FFT / mel  = quad_stage1 / quad_stage2 / mel_stage of fbank_tile.cuh (packed f32x2 real FFT, interval mel).
//   * LFR + CMVN = closed form, per quad: the 4 log-mel frames go through a 1.25 KB per-warp staging tile and are
//                  written as 128-bit rows wherever clamp(n*i + jj - left, 0, T-1) == f  (VF:40-60: frame f is slot jj
//                  of row i), at most two places away from utterance edges; (x + shift) * scale in the reference's
//                  operation order (VF:23-37).  No halo frames, no CTA-wide second pass.
//
// The tile kernel (fbank_tile.cuh) remains for the CMVN-statistics pass and for frame shifts whose quad does not fit
// the 896-float buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fbank_tile.cuh"

namespace b200fe {

constexpr int kQuadVecs = 7;                      // float4 per lane and quad
constexpr int kQuadBuf = kQuadVecs * 32 * 4;      // 896 floats per warp
constexpr unsigned kNoTarget = 0xFFFFFFFFu;
constexpr int kTargetOffBits = 27;                // output offsets inside one utterance must stay below 2^27 floats

// One unit of work, fully resolved by build_quads_kernel (64 bytes, read one quad ahead).
struct __align__(16) QuadDesc {
  long long g0;       // absolute index (in the wave buffer) of the first sample of the quad's first frame
  int utt;
  int f0;             // first frame of the quad
  int nF;             // bits 0..7: frames in the quad (1..4); bits 8..11: frame t takes the generic LFR path
  int T;              // frames of the utterance
  int rows;           // LFR rows of the utterance
  int pad;
  unsigned tgt[8];    // tgt[2 t + k]: where frame f0 + t goes: (jj << 27) | (row * D + jj * M), or kNoTarget
};

struct QuadParams {
  const void* wave;       // float32 PCM in [-1, 1] or int16 PCM (sample / 32768, R:voice_interface.py:1008-1013)
  long long wave_total;   // samples addressable behind `wave`
  const QuadDesc* quads;
  int n_quads;
  int* next_quad;         // work counter (zeroed by build_quads_kernel): quads beyond the first wave are claimed dynamically
  float* feats;           // [batch, rows_cap, out_dim]
  long long rows_cap;
  int frame_len, frame_shift, n_mels, lfr_m, lfr_n;
  float preemph;
  int remove_dc;
  float log_floor;
  float dither;
  unsigned long long seed;
  const float* window;
  const float2* twiddle;
  const float2* mel_w;
  const int* mel_lo;
  int mel_rounds;
  int mel_cnt[kMelRounds];
  int mel_base[kMelRounds];
  const float* cmvn;      // nullptr or [2][out_dim]
};

__host__ __device__ inline size_t warp_smem_bytes() {
  return (size_t)kWarps * kQuadBuf * 4 + (size_t)kWarps * kYWarpF4 * 16 + (size_t)kTw2Total * 8;
}
// the quad of a (frame_len, frame_shift) pair fits the per-warp buffer (3 floats of alignment slack)
__host__ __device__ inline bool warp_kernel_fits(int L, int S) { return 3 + 3 * S + L <= kQuadBuf; }

// Where frame f of an utterance with T frames / `rows` LFR rows goes (VF:40-60): slot jj of row i wherever
// n*i + jj - left == f.  With (i_hi, jj1) the solution of largest i, the others are (i_hi - k, jj1 + k n).  Two are
// encoded for the inline path; utterance edges (frames replicated by the LFR padding) and lfr_m > 2 lfr_n are flagged
// for the generic loop.  Runs in build_quads_kernel, so the integer divisions stay out of the hot loop.
__host__ __device__ inline bool quad_targets(int f, int T, int rows, int lfr_m, int lfr_n, int M, unsigned* t2) {
  const int left = (lfr_m - 1) / 2, D = lfr_m * M;
  const int i_hi = (f + left) / lfr_n, jj1 = (f + left) % lfr_n;
  t2[0] = t2[1] = kNoTarget;
  if (f == 0 || f == T - 1 || jj1 + 2 * lfr_n < lfr_m) return true;
  if (i_hi < rows && jj1 < lfr_m) t2[0] = ((unsigned)jj1 << kTargetOffBits) | (unsigned)(i_hi * D + jj1 * M);
  if (i_hi >= 1 && i_hi - 1 < rows && jj1 + lfr_n < lfr_m)
    t2[1] = ((unsigned)(jj1 + lfr_n) << kTargetOffBits) | (unsigned)((i_hi - 1) * D + (jj1 + lfr_n) * M);
  return false;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Samples of a quad: HBM/L2 -> registers (128-bit streaming loads on the 16-byte grid of the wave buffer; the samples
// must not evict the mel / CMVN tables from the few KB of L1 left beside 4 CTAs) -> pre-emphasis -> the warp's buffer:
// buf[a_off + n] = x[n] - preemph * x[n-1] for sample n of the quad's first frame (the predecessor of a vector's first
// sample comes from the previous lane through one shuffle).  Also returns the raw first / last sample of the group's
// two frames (x0 / xl, for the DC correction), read by lanes 0..7 and shuffled.  Quads touching the ends of the wave
// buffer load element-wise.  Returns a_off.
template <class SampleT>
__device__ __forceinline__ int quad_samples(const QuadParams& p, long long g0, int nF, int lane, int grp_in_warp,
                                            float* buf, f2& x0, f2& xl);

template <>
__device__ __forceinline__ int quad_samples<float>(const QuadParams& p, long long g0, int nF, int lane, int grp_in_warp,
                                                   float* buf, f2& x0, f2& xl) {
  const float* wave = static_cast<const float*>(p.wave);
  const unsigned wave_mis = (unsigned)((reinterpret_cast<uintptr_t>(wave) >> 2) & 3);
  const int a_off = (int)((wave_mis + (unsigned)(g0 & 3)) & 3);
  const long long ga = g0 - a_off;
  const int nv = (a_off + (nF - 1) * p.frame_shift + p.frame_len + 3) >> 2;
  float4 x[kQuadVecs];
  if (ga >= 0 && ga + 4ll * nv <= p.wave_total) {
    const float* src = wave + ga + 4 * lane;
#pragma unroll
    for (int u = 0; u < kQuadVecs; ++u)
      x[u] = lane + 32 * u < nv ? ldg_stream4(src + 128 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
#pragma unroll
    for (int u = 0; u < kQuadVecs; ++u) {
      const long long ab = ga + 4ll * (lane + 32 * u);
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane + 32 * u < nv) {
        if (ab >= 0 && ab < p.wave_total) x[u].x = wave[ab];
        if (ab + 1 >= 0 && ab + 1 < p.wave_total) x[u].y = wave[ab + 1];
        if (ab + 2 >= 0 && ab + 2 < p.wave_total) x[u].z = wave[ab + 2];
        if (ab + 3 >= 0 && ab + 3 < p.wave_total) x[u].w = wave[ab + 3];
      }
    }
  }
  float cap = 0.f;
  if (lane < 8 && (lane & 3) < nF)
    cap = __ldg(wave + g0 + (long long)(lane & 3) * p.frame_shift + (lane < 4 ? 0 : p.frame_len - 1));
  float4* buf4 = reinterpret_cast<float4*>(buf);
  float below = 0.f;   // lane 31's last sample of the previous vector row (lane 0's predecessor)
#pragma unroll
  for (int u = 0; u < kQuadVecs; ++u) {
    const float rot = __shfl_sync(0xffffffffu, x[u].w, (lane + 31) & 31);
    const float pv = lane == 0 ? below : rot;
    below = rot;   // only lane 0 uses it: there rot is lane 31's value
    if (lane + 32 * u < nv) {
      float4 e;
      e.x = fmaf(-p.preemph, pv, x[u].x);
      e.y = fmaf(-p.preemph, x[u].x, x[u].y);
      e.z = fmaf(-p.preemph, x[u].y, x[u].z);
      e.w = fmaf(-p.preemph, x[u].z, x[u].w);
      buf4[lane + 32 * u] = e;
    }
  }
  const int fa = 2 * grp_in_warp;
  x0.x = __shfl_sync(0xffffffffu, cap, fa);
  x0.y = __shfl_sync(0xffffffffu, cap, fa + 1);
  xl.x = __shfl_sync(0xffffffffu, cap, fa + 4);
  xl.y = __shfl_sync(0xffffffffu, cap, fa + 5);
  return a_off;
}

// int16 PCM: 8 samples per 16-byte vector, converted with the reference's own rule float(s) / 32768 (exact), so the
// result is bit-identical to handing the converted float32 buffer to the float path - at half the HBM / PCIe bytes.
constexpr int kQuadVecs16 = 4;
template <>
__device__ __forceinline__ int quad_samples<short>(const QuadParams& p, long long g0, int nF, int lane, int grp_in_warp,
                                                   float* buf, f2& x0, f2& xl) {
  const short* wave = static_cast<const short*>(p.wave);
  const unsigned wave_mis = (unsigned)((reinterpret_cast<uintptr_t>(wave) >> 1) & 7);
  const int a_off = (int)((wave_mis + (unsigned)(g0 & 7)) & 7);
  const long long ga = g0 - a_off;
  const int nv = (a_off + (nF - 1) * p.frame_shift + p.frame_len + 7) >> 3;
  constexpr float kScale = 1.0f / 32768.0f;
  float x[kQuadVecs16][8];
  if (ga >= 0 && ga + 8ll * nv <= p.wave_total) {
    const short* src = wave + ga + 8 * lane;
#pragma unroll
    for (int u = 0; u < kQuadVecs16; ++u) {
      uint4 r = make_uint4(0u, 0u, 0u, 0u);
      if (lane + 32 * u < nv) r = ldg_stream_u4(src + 256 * u);
      const unsigned w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x[u][2 * k] = (float)(short)(w[k] & 0xffffu) * kScale;
        x[u][2 * k + 1] = (float)(short)(w[k] >> 16) * kScale;
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < kQuadVecs16; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const long long ab = ga + 8ll * (lane + 32 * u) + k;
        x[u][k] = (lane + 32 * u < nv && ab >= 0 && ab < p.wave_total) ? (float)wave[ab] * kScale : 0.f;
      }
  }
  float cap = 0.f;
  if (lane < 8 && (lane & 3) < nF)
    cap = (float)wave[g0 + (long long)(lane & 3) * p.frame_shift + (lane < 4 ? 0 : p.frame_len - 1)] * kScale;
  float4* buf4 = reinterpret_cast<float4*>(buf);
  float below = 0.f;
#pragma unroll
  for (int u = 0; u < kQuadVecs16; ++u) {
    const float rot = __shfl_sync(0xffffffffu, x[u][7], (lane + 31) & 31);
    const float pv = lane == 0 ? below : rot;
    below = rot;
    if (lane + 32 * u < nv) {
      float4 e0, e1;
      e0.x = fmaf(-p.preemph, pv, x[u][0]);
      e0.y = fmaf(-p.preemph, x[u][0], x[u][1]);
      e0.z = fmaf(-p.preemph, x[u][1], x[u][2]);
      e0.w = fmaf(-p.preemph, x[u][2], x[u][3]);
      e1.x = fmaf(-p.preemph, x[u][3], x[u][4]);
      e1.y = fmaf(-p.preemph, x[u][4], x[u][5]);
      e1.z = fmaf(-p.preemph, x[u][5], x[u][6]);
      e1.w = fmaf(-p.preemph, x[u][6], x[u][7]);
      buf4[2 * (lane + 32 * u)] = e0;
      buf4[2 * (lane + 32 * u) + 1] = e1;
    }
  }
  const int fa = 2 * grp_in_warp;
  x0.x = __shfl_sync(0xffffffffu, cap, fa);
  x0.y = __shfl_sync(0xffffffffu, cap, fa + 1);
  xl.x = __shfl_sync(0xffffffffu, cap, fa + 4);
  xl.y = __shfl_sync(0xffffffffu, cap, fa + 5);
  return a_off;
}

// SR: frame shift in 16-sample rows when it is a whole number of rows and known at compile time (10 for 400/160):
// the two frames of a pair then share their overlapping sample loads.  0 = generic.
template <int NROWS, bool EXACT, bool DITHER, class MELS, int SR, class SampleT>
__global__ void __launch_bounds__(kCtaThreads, 4)
fbank_warp_kernel(const QuadParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* bufs = reinterpret_cast<float*>(smem_raw);
  float4* xbuf = reinterpret_cast<float4*>(bufs + kWarps * kQuadBuf);
  float2* tw_s = reinterpret_cast<float2*>(xbuf + kWarps * kYWarpF4);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int j = tid & (kGroup - 1);
  const int grp_in_warp = lane >> 4;
  const int g = NROWS < 32 ? grp_in_warp : 0;
  const int L = p.frame_len, S = p.frame_shift, M = p.n_mels;
  const int lfr_m = p.lfr_m, lfr_n = p.lfr_n;
  const int D = lfr_m * M;
  const int lfr_left = (lfr_m - 1) / 2;

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);
  __syncthreads();   // the only CTA-wide barrier: the twiddle tables
  // Launched as a programmatic dependent of prep_warp_kernel (b200fe.cu: launch_warp): everything above reads
  // per-handle constants only and overlaps the tail of that kernel; the quad list and the work counter are its
  // output.  (A no-op when the launch carries no such attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  float* buf = bufs + warp * kQuadBuf;
  float4* yg = xbuf + warp * kYWarpF4 + grp_in_warp * kYGroupF4;
  float4* pbuf4 = xbuf + warp * kYWarpF4;
  float* lm_s = reinterpret_cast<float*>(pbuf4 + kSpecF4);   // log-mel staging tile, behind the warp's spectra
  const float2* tw_row = fft_twiddle_row<NROWS>(tw_s, j, grp_in_warp);
  const float2* c0_row = fft_c0_row(tw_s, j);
  const int M4 = M >> 2;
  const bool act = lane < M4;      // lanes that move a float4 of a log-mel row (n_mels <= 128)

  // The 64-byte descriptor is read in pieces, each just before it is needed, so that it never occupies 16 registers:
  // {g0, utt, f0} and {nF, T, rows} at the top of a quad, the targets before the mel stage, and the next quad's
  // {g0, nF} (for its sample loads) one quad ahead.
  // Work distribution: the first quad of every warp is static (neighbouring warps start on neighbouring quads), all
  // later ones are claimed from a global counter one quad ahead, so that no SM idles while another still has a queue.
  const int first_wave = gridDim.x * kWarps;
  int q = blockIdx.x * kWarps + warp;
  if (q >= p.n_quads) return;
  long long g0_cur = __ldg(&p.quads[q].g0);
  int nf_cur = __ldg(&p.quads[q].nF);

  while (true) {
    int qn = 0;
    if (lane == 0) qn = first_wave + atomicAdd(p.next_quad, 1);
    qn = __shfl_sync(0xffffffffu, qn, 0);
    const bool have_next = qn < p.n_quads;
    long long g0_next = 0;
    int nf_next = 0;
    if (have_next) {
      g0_next = __ldg(&p.quads[qn].g0);
      nf_next = __ldg(&p.quads[qn].nF);
    }
    const int4 hd = __ldg(reinterpret_cast<const int4*>(p.quads + q));        // g0 (2 words), utt, f0
    const int4 hd1 = __ldg(reinterpret_cast<const int4*>(p.quads + q) + 1);   // nF | slow << 8, T, rows, pad
    const int utt = hd.z, f0 = hd.w;
    const int nF = nf_cur & 0xff, slow = (nf_cur >> 8) & 0xf, T = hd1.y, rows = hd1.z;

    // ---- this quad's samples: HBM -> L2 was started a whole quad ago (prefetch below), so these loads hit L2;
    //      pre-emphasis on the way from registers to the warp's buffer
    f2 x0, xl;
    const int a_off = quad_samples<SampleT>(p, g0_cur, nF, lane, grp_in_warp, buf, x0, xl);
    __syncwarp();

    // ---- stage 1 (samples -> registers -> real 32-point FFT) and stage 2 (transpose, 16-point FFT, power spectra)
    const int fA = 2 * grp_in_warp;
    const bool vA = fA < nF, vB = fA + 1 < nF;
    {
      f2 zr[16], zi[16], y0, y16;
      quad_stage1<NROWS, EXACT, DITHER, SR>(buf + a_off + fA * S, x0, xl, vA, vB, S, L, win, p.preemph, p.remove_dc,
                                            p.dither, p.seed, (unsigned)utt, (unsigned)(f0 + fA), j, g, zr, zi, y0, y16);
      __syncwarp();   // every lane is done with the sample buffer and with the previous quad's staging tile
      quad_stage2(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
    }

    // ---- the next quad's samples start their way from HBM to L2 now (one 128-byte line per lane)
    if (have_next) {
      const char* base = static_cast<const char*>(p.wave);
      const long long first = (g0_next * (long long)sizeof(SampleT) - 16) & ~127ll;
      const long long idx = first + 128ll * lane;
      const long long last = (g0_next + 3 * S + L) * (long long)sizeof(SampleT);
      if (idx >= 0 && idx < last && idx < p.wave_total * (long long)sizeof(SampleT))
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + idx));
    }
    const uint4 tg0 = __ldg(reinterpret_cast<const uint4*>(p.quads + q) + 2);   // targets of frames 0, 1
    const uint4 tg1 = __ldg(reinterpret_cast<const uint4*>(p.quads + q) + 3);   // targets of frames 2, 3

    // ---- log-mel of the 4 frames into the warp's staging tile
    mel_stage<MELS>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b, float c, float d) {
      lm_s[iv] = a;
      lm_s[M + iv] = b;
      lm_s[2 * M + iv] = c;
      lm_s[3 * M + iv] = d;
    });
    __syncwarp();

    // ---- LFR + CMVN: each frame's row of n_mels goes, as 128-bit pieces, to the (row, slot) pairs of its descriptor
    {
      float* out_l = p.feats + (long long)utt * p.rows_cap * D + 4 * lane;
      const unsigned tgt[8] = {tg0.x, tg0.y, tg0.z, tg0.w, tg1.x, tg1.y, tg1.z, tg1.w};
      const float* cm_l = p.cmvn ? p.cmvn + 4 * lane : nullptr;
      const float4* lm4 = reinterpret_cast<const float4*>(lm_s) + lane;
      auto cmvn4 = [&](const float4& v, const float4& sh, const float4& sc) {   // (x + shift) * scale, VF:34-35
        return make_float4((v.x + sh.x) * sc.x, (v.y + sh.y) * sc.y, (v.z + sh.z) * sc.z, (v.w + sh.w) * sc.w);
      };
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      // primary slot of the 4 frames: all loads first (one exposed latency), then arithmetic and stores
      float4 v[4], sh[4], sc[4];
      bool ok[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        ok[t] = tgt[2 * t] != kNoTarget && act;
        v[t] = act ? lm4[t * M4] : zero4;
        sh[t] = sc[t] = zero4;
        if (cm_l && ok[t]) {
          const int jm = (int)(tgt[2 * t] >> kTargetOffBits) * M;
          sh[t] = __ldg(reinterpret_cast<const float4*>(cm_l + jm));
          sc[t] = __ldg(reinterpret_cast<const float4*>(cm_l + D + jm));
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (ok[t]) stg_stream4(out_l + (int)(tgt[2 * t] & ((1u << kTargetOffBits) - 1)), cm_l ? cmvn4(v[t], sh[t], sc[t]) : v[t]);
      // second slot (one frame in lfr_n when lfr_m = lfr_n + 1): warp-uniform branch
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned code = tgt[2 * t + 1];
        if (code != kNoTarget && act) {
          float4 o = v[t];
          if (cm_l) {
            const int jm = (int)(code >> kTargetOffBits) * M;
            o = cmvn4(o, __ldg(reinterpret_cast<const float4*>(cm_l + jm)), __ldg(reinterpret_cast<const float4*>(cm_l + D + jm)));
          }
          stg_stream4(out_l + (int)(code & ((1u << kTargetOffBits) - 1)), o);
        }
      }
      if (slow) {   // first / last frame of the utterance (replicated by the LFR padding), or lfr_m > 2 lfr_n
#pragma unroll 1
        for (int t = 0; t < nF; ++t) {
          if (!((slow >> t) & 1)) continue;
          const int f = f0 + t;
          const int num = f + lfr_left - (lfr_m - 1);
          const int i_lo = (f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n;
          const int i_top = f == T - 1 ? rows - 1 : min((f + lfr_left) / lfr_n, rows - 1);
          const float4 v = act ? lm4[t * M4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int i = i_lo; i <= i_top; ++i)
#pragma unroll 1
            for (int jj = 0; jj < lfr_m; ++jj)
              if (act && min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) {
                float4 o = v;
                if (cm_l) {
                  const float4 sh = __ldg(reinterpret_cast<const float4*>(cm_l + jj * M));
                  const float4 sc = __ldg(reinterpret_cast<const float4*>(cm_l + D + jj * M));
                  o.x = (v.x + sh.x) * sc.x;
                  o.y = (v.y + sh.y) * sc.y;
                  o.z = (v.z + sh.z) * sc.z;
                  o.w = (v.w + sh.w) * sc.w;
                }
                stg_stream4(out_l + (long long)i * D + jj * M, o);
              }
        }
      }
    }

    if (!have_next) break;
    q = qn;
    g0_cur = g0_next;
    nf_cur = nf_next;
  }
}

}  // namespace b200fe
