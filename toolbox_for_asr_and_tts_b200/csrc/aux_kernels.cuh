// Small kernels around the fused tile kernel: padding rows, LFR+CMVN of given features, utterances shorter than
// one frame (VF:147 quirk), synthetic PCM for benchmarks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fbank_tile.cuh"
#include "fbank_warp.cuh"

namespace b200fe {

// pad_sequence(padding_value=0.0) (VF:163-166): rows [n_rows[u], rows_cap) of every utterance are zeroed, and
// feat_lens[u] = n_rows[u] (VF:160).  grid = (chunks, batch).
__device__ __forceinline__ void pad_rows_block(const UttDesc* utts, float* feats, long long rows_cap, int D,
                                               long long* feat_lens, int u, int bx, int nbx) {
  const int n_rows = utts[u].n_rows;
  if (bx == 0 && threadIdx.x == 0 && feat_lens) feat_lens[u] = n_rows;
  float* base = feats + ((long long)u * rows_cap + n_rows) * D;
  const long long total = (rows_cap - n_rows) * (long long)D;
  const long long stride = (long long)nbx * blockDim.x;
  const long long i0 = (long long)bx * blockDim.x + threadIdx.x;
  if ((D & 3) == 0) {  // row starts stay 16-byte aligned
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i = i0; i < (total >> 2); i += stride) stg_stream4(base + 4 * i, z);
  } else {
    for (long long i = i0; i < total; i += stride) base[i] = 0.f;
  }
}

__global__ void pad_rows_kernel(const UttDesc* utts, float* feats, long long rows_cap, int D, long long* feat_lens) {
  pad_rows_block(utts, feats, rows_cap, D, feat_lens, blockIdx.y, blockIdx.x, gridDim.x);
}

// WavFrontend.forward_lfr_cmvn (VF:198-218): out[u, i, j*M + d] = (in[u, clamp(n*i + j - left, 0, T-1), d] + shift) * scale
__global__ void lfr_cmvn_kernel(const float* fbank, long long frames_cap, const UttDesc* utts, int M, int lfr_m,
                                int lfr_n, const float* cmvn, float* feats, long long rows_cap) {
  const int u = blockIdx.y;
  const int T = utts[u].n_frames, R = utts[u].n_rows;
  const int D = lfr_m * M, left = (lfr_m - 1) / 2;
  const float* in = fbank + (long long)u * frames_cap * M;
  float* out = feats + (long long)u * rows_cap * D;
  const long long total = (long long)R * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i - (long long)r * D);
    const int jj = c / M, d = c - jj * M;
    const int f = min(max(lfr_n * r + jj - left, 0), T - 1);
    float v = in[(long long)f * M + d];
    if (cmvn) v = (v + cmvn[c]) * cmvn[D + c];
    out[i] = v;
  }
}

// Utterances with fewer samples than one frame: the reference shrinks the frame to the utterance
// (frame_length = min(frame_length, len/fs*1000), VF:147), so window size, FFT size and the mel bank change per
// utterance.  Rare and tiny: one CTA per utterance, direct DFT with exact integer phase reduction.
struct ShortDesc {
  long long wave_off;
  int utt;          // batch index
  int n_samples;
  int win;          // window_size  (TA:138)
  int nfft;         // padded_window_size (TA:139), power of two <= 512
  int n_frames;
  int n_rows;
  int mel_off;      // offset (floats) of this nfft's dense [n_mels, nfft/2] bank inside short_mel
  int row_begin;    // first row of the utterance in a rows-packed output
};

__global__ void __launch_bounds__(256)
short_utt_kernel(const float* wave, const ShortDesc* descs, const float* short_mel, int S, int M, int lfr_m, int lfr_n,
                 int window_type, float blackman_coeff, float preemph, int remove_dc, float upscale, float log_floor,
                 const float* cmvn, float* feats, long long rows_cap) {
  __shared__ float z[512];
  __shared__ float pw[257];
  __shared__ float red[8];
  __shared__ float lm[kMaxMels];
  const ShortDesc sd = descs[blockIdx.x];
  const int tid = threadIdx.x;
  const int W = sd.win, N = sd.nfft, D = lfr_m * M, left = (lfr_m - 1) / 2;
  const float* x = wave + sd.wave_off;
  float* out = feats + (rows_cap >= 0 ? (long long)sd.utt * rows_cap : (long long)sd.row_begin) * D;
  // n_frames is 1 unless frame_shift is tiny; LFR rows are emitted per frame window in order, recomputing frames.
  for (int r = 0; r < sd.n_rows; ++r) {
    for (int jj = 0; jj < lfr_m; ++jj) {
      const int f = min(max(lfr_n * r + jj - left, 0), sd.n_frames - 1);
      const float* xf = x + (long long)f * S;
      // mean over the frame
      float part = 0.f;
      for (int n = tid; n < W; n += blockDim.x) part += xf[n] * upscale;
      for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((tid & 31) == 0) red[tid >> 5] = part;
      __syncthreads();
      float mean = 0.f;
      if (remove_dc) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mean += red[w];
        mean /= (float)W;
      }
      for (int n = tid; n < N; n += blockDim.x) {
        float v = 0.f;
        if (n < W) {
          const float cur = xf[n] * upscale - mean;
          const float prv = xf[n > 0 ? n - 1 : 0] * upscale - mean;
          const float y = cur - preemph * prv;
          const double a = 2.0 * n / (double)(W - 1);   // in units of pi
          float wv;
          if (window_type == 0) wv = (float)(0.54 - 0.46 * cospi(a));
          else if (window_type == 1) wv = (float)(0.5 - 0.5 * cospi(a));
          else if (window_type == 2) wv = powf((float)(0.5 - 0.5 * cospi(a)), 0.85f);
          else if (window_type == 3) wv = 1.f;
          else wv = (float)(blackman_coeff - 0.5 * cospi(a) + (0.5 - blackman_coeff) * cospi(2.0 * a));
          v = y * wv;
        }
        z[n] = v;
      }
      __syncthreads();
      for (int k = tid; k <= N / 2; k += blockDim.x) {
        float sr = 0.f, si = 0.f;
        for (int n = 0; n < W; ++n) {
          const int ph = (k * n) & (N - 1);
          float s, c;
          sincospif(2.0f * (float)ph / (float)N, &s, &c);
          sr = fmaf(z[n], c, sr);
          si = fmaf(z[n], -s, si);
        }
        pw[k] = sr * sr + si * si;
      }
      __syncthreads();
      const float* bank = short_mel + sd.mel_off;
      for (int m = tid; m < M; m += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < N / 2; ++k) acc = fmaf(bank[m * (N / 2) + k], pw[k], acc);
        lm[m] = logf(fmaxf(acc, log_floor));
      }
      __syncthreads();
      for (int m = tid; m < M; m += blockDim.x) {
        const int c = jj * M + m;
        float v = lm[m];
        if (cmvn) v = (v + cmvn[c]) * cmvn[D + c];
        out[(long long)r * D + c] = v;
      }
      __syncthreads();
    }
  }
}

__global__ void add_count_kernel(double* dst, double v) { *dst += v; }

// Expands the per-utterance descriptors into the launch-wide tile list (one thread per tile).
__global__ void build_tiles_kernel(const UttDesc* utts, int batch, int n_tiles, int rows_per_tile, int lfr_m, int lfr_n,
                                   int S, TileDesc* tiles, int* next_tile) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) *next_tile = 0;   // the tile kernel's work counter
  if (t >= n_tiles) return;
  int lo = 0, hi = batch - 1;   // last utterance whose tile_begin <= t (utterances without tiles share the next begin)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (utts[mid].tile_begin <= t) lo = mid; else hi = mid - 1;
  }
  const UttDesc ud = utts[lo];
  const int left = (lfr_m - 1) / 2;
  TileDesc d;
  d.utt = lo;
  d.T = ud.n_frames;
  d.row0 = (t - ud.tile_begin) * rows_per_tile;
  d.nrow = min(rows_per_tile, ud.n_rows - d.row0);
  d.f_lo = min(max(lfr_n * d.row0 - left, 0), d.T - 1);
  const int f_hi = min(max(lfr_n * (d.row0 + d.nrow - 1) - left + lfr_m - 1, 0), d.T - 1);
  d.F = f_hi - d.f_lo + 1;
  d.g0 = ud.wave_off + (long long)d.f_lo * S;
  tiles[t] = d;
}

// int16 PCM -> float32 with the reference's rule s / 32768 (R:voice_interface.py:1008-1013); used for the few samples
// of utterances shorter than one frame when the batch arrives as int16.
__global__ void pcm16_to_float_kernel(const short* src, int n, float* dst) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = (float)src[i] * (1.0f / 32768.0f);
}

// Expands the per-utterance descriptors into the launch-wide quad list of the warp kernel: quad q of utterance u covers
// frames 4 (q - quad_begin[u]) .. +3.  One thread per quad; utterances without frames own no quads.
__device__ __forceinline__ void build_quad(const UttDesc* utts, int batch, int q, int S, int lfr_m, int lfr_n, int M,
                                           QuadDesc* quads) {
  int lo = 0, hi = batch - 1;
  while (lo < hi) {   // last utterance whose quad_begin <= q (empty utterances share the next one's quad_begin)
    const int mid = (lo + hi + 1) >> 1;
    if (utts[mid].quad_begin <= q) lo = mid; else hi = mid - 1;
  }
  const UttDesc ud = utts[lo];
  QuadDesc d;
  d.utt = lo;
  d.T = ud.n_frames;
  d.rows = ud.n_rows;
  d.f0 = 4 * (q - ud.quad_begin);
  const int nF = min(4, ud.n_frames - d.f0);
  d.g0 = ud.wave_off + (long long)d.f0 * S;
  d.row_begin = ud.row_begin;
  int slow = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    d.tgt[2 * t] = d.tgt[2 * t + 1] = kNoTarget;
    if (t < nF && quad_targets(d.f0 + t, d.T, d.rows, lfr_m, lfr_n, M, d.tgt + 2 * t)) slow |= 1 << t;
  }
  d.nF = nF | (slow << 8);
  quads[q] = d;
}

// One launch in front of the warp kernel: blocks [0, quad_blocks) build the quad list (latency-bound: a binary search
// per quad), the next table_blocks write feat_lens (and copy the utterance table to global memory when it arrived in
// the launch parameters), the remaining pad_bx * batch blocks clear the padding rows (bandwidth-bound; pad_bx = 0
// when the warp kernel writes them itself, B200FE_PAD_MODE).
__device__ __forceinline__ void prep_warp_body(const UttDesc* utts, int batch, int n_quads, int quad_blocks, int table_blocks,
                                               int S, int lfr_m, int lfr_n, int M, QuadDesc* quads, int* next_quad,
                                               float* feats, long long rows_cap, long long* feat_lens, UttDesc* utts_out,
                                               int pad_bx) {
  // the warp kernel that follows may start its (input-independent) prologue while this grid drains; it waits for this
  // grid's completion (griddepcontrol.wait) before it touches the quad list
  asm volatile("griddepcontrol.launch_dependents;");
  const int b = blockIdx.x;
  if (b < quad_blocks) {
    const int q = b * blockDim.x + threadIdx.x;
    if (q == 0) *next_quad = 0;   // the warp kernel's work counter
    if (q < n_quads) build_quad(utts, batch, q, S, lfr_m, lfr_n, M, quads);
  } else if (b < quad_blocks + table_blocks) {
    const int u = (b - quad_blocks) * blockDim.x + threadIdx.x;
    if (u < batch) {
      if (feat_lens) feat_lens[u] = utts[u].n_rows;
      if (utts_out) utts_out[u] = utts[u];
    }
  } else {
    const int k = b - quad_blocks - table_blocks;
    pad_rows_block(utts, feats, rows_cap, lfr_m * M, nullptr, k / pad_bx, k % pad_bx, pad_bx);
  }
}

__global__ void prep_warp_kernel(const UttDesc* utts, int batch, int n_quads, int quad_blocks, int table_blocks, int S,
                                 int lfr_m, int lfr_n, int M, QuadDesc* quads, int* next_quad, float* feats,
                                 long long rows_cap, long long* feat_lens, int pad_bx) {
  prep_warp_body(utts, batch, n_quads, quad_blocks, table_blocks, S, lfr_m, lfr_n, M, quads, next_quad, feats, rows_cap,
                 feat_lens, nullptr, pad_bx);
}

// The same with the utterance table carried in the launch itself (kernel parameters, 8 KB of the 32 KB a launch may
// carry): batches of up to kParamUtts utterances need no host -> device copy in front of the two kernels.
constexpr int kParamUtts = 256;
struct UttTable { UttDesc u[kParamUtts]; };
__global__ void prep_warp_kernel_tab(const __grid_constant__ UttTable tab, int batch, int n_quads, int quad_blocks,
                                     int table_blocks, int S, int lfr_m, int lfr_n, int M, QuadDesc* quads, int* next_quad,
                                     float* feats, long long rows_cap, long long* feat_lens, UttDesc* utts_out, int pad_bx) {
  prep_warp_body(tab.u, batch, n_quads, quad_blocks, table_blocks, S, lfr_m, lfr_n, M, quads, next_quad, feats, rows_cap,
                 feat_lens, utts_out, pad_bx);
}

// Counter-based synthetic PCM, bit-identical to toolbox_for_asr_and_tts_b200/synth.py::uniform_pcm.
__device__ __forceinline__ unsigned long long synth_mix(unsigned long long seed, unsigned long long u,
                                                        unsigned long long n) {
  unsigned long long z = seed * 0x9E3779B97F4A7C15ull + u * 0xBF58476D1CE4E5B9ull + n;
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

// utt_ids (optional): the generator's utterance id of batch entry u (default: u itself), so that a rank can synthesise
// an arbitrary subset of a corpus (sharding.partition_utterances) bit-identically to synth.uniform_pcm(seed, id, n).
__global__ void synth_uniform_kernel(float* wave, const long long* offsets, const long long* lengths,
                                     const long long* utt_ids, int batch, unsigned long long seed, float amp) {
  for (int u = blockIdx.y; u < batch; u += gridDim.y) {
  const long long n_u = lengths[u];
  const unsigned long long id = utt_ids ? (unsigned long long)utt_ids[u] : (unsigned long long)u;
  float* dst = wave + offsets[u];
  for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n_u;
       n += (long long)gridDim.x * blockDim.x) {
    const unsigned k = (unsigned)(synth_mix(seed, id, (unsigned long long)n) >> 40);
    const float c = (float)((int)k - (1 << 23)) * (1.0f / 8388608.0f);   // exact: 2*U01 - 1
    dst[n] = __fmul_rn(amp, c);
  }
  }
}

}  // namespace b200fe
