// Chunked-streaming front-end: one CTA per (stream, tick).  The per-stream state (sample carry = upstream
// WavFrontendOnline `input_cache`, LFR splice frames = `lfr_splice_cache`, counters) lives in one HBM slab, so the
// host never concatenates audio (cf. the np.concatenate growth at R:voice-service/app/services/
// voice_interface.py:1304-1311,1688-1746) and nothing crosses PCIe but the new chunk and the new rows.
//
// Semantics (oracle/wav_frontend_np.py::OnlineFrontend): frames are cut from [carry | chunk]; LFR row i is emitted as
// soon as frame lfr_n*i + lfr_m-1-(lfr_m-1)/2 exists; `is_final` flushes the remaining rows up to ceil(T/lfr_n) with
// the last frame replicated and resets the stream.  Concatenated outputs equal the offline output.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fbank_tile.cuh"
#include "fbank_warp.cuh"

namespace b200fe {

struct StreamLayout {
  int n_streams;
  int carry_cap;    // floats per stream (>= frame_len - 1, multiple of 4)
  int cache_cap;    // frames per stream: max(lfr_m - 1, 1)
  int n_mels;
  __host__ __device__ size_t counters_bytes() const { return ((size_t)4 * n_streams * sizeof(int) + 255) & ~(size_t)255; }
  __host__ __device__ size_t carry_bytes() const { return ((size_t)n_streams * carry_cap * sizeof(float) + 255) & ~(size_t)255; }
  __host__ __device__ size_t cache_bytes() const { return ((size_t)n_streams * cache_cap * n_mels * sizeof(float) + 255) & ~(size_t)255; }
  __host__ __device__ size_t total_bytes() const { return counters_bytes() + carry_bytes() + cache_bytes(); }
  __host__ __device__ int* counters(void* base) const { return reinterpret_cast<int*>(base); }
  __host__ __device__ float* carry(void* base) const { return reinterpret_cast<float*>((char*)base + counters_bytes()); }
  __host__ __device__ float* cache(void* base) const { return reinterpret_cast<float*>((char*)base + counters_bytes() + carry_bytes()); }
};

struct StreamParams {
  void* state;
  StreamLayout lay;
  const float* chunks;
  long long chunk_stride;
  const int* chunk_lens;
  const int* stream_ids;
  const unsigned char* is_final;
  int n;
  int max_chunk;
  int nf_max;             // frames one push can create
  float* feats;           // [n, rows_cap, D]
  long long rows_cap;
  int* rows_out;          // [n]
  float* chunk_stats;     // nullptr or [n, 2]: mean |x| and max |x| of each pushed chunk (the reference's energy gate)
  int frame_len, frame_shift, n_mels, lfr_m, lfr_n;
  int e_cap;
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
  const float* cmvn;
};

// per_quad: every warp fetches the samples of its quads itself (3.5 KB per warp) instead of the CTA staging the whole
// [carry | chunk] buffer: 74 KB instead of 100 KB for a 600 ms chunk, i.e. 3 stream-CTAs per SM instead of 2.
__host__ __device__ inline size_t stream_smem_bytes(int e_cap, int nf_max, int cache_cap, int n_mels, bool per_quad,
                                                    int warps = kWarps) {
  size_t b = 0;
  if (per_quad) {
    b += (size_t)warps * kQuadBuf * 4;
  } else {
    b += (size_t)e_cap * 4;
  }
  b += (size_t)warps * kYWarpF4 * 16;
  b += (size_t)(cache_cap + nf_max) * n_mels * 4;
  b += kTw2Total * 8;
  return b;
}

// Samples of one quad of a stream: frames start at virtual index `base` of [carry | chunk].  Vectors that lie inside
// the chunk are read with aligned 128-bit loads (the load grid is aligned to the chunk's address), vectors that touch
// the carry or the ends element-wise; raw samples go to the warp's buffer (quad_stage1 does the rest).  Returns a_off.
__device__ __forceinline__ int stream_quad_samples(const float* carry, int carry_len, const float* chunk, int n, int base,
                                                   int nF, int S, int L, int lane, float* buf) {
  auto at = [&](int i) { return (i < 0 || i >= n) ? 0.f : (i < carry_len ? carry[i] : chunk[i - carry_len]); };
  const int chunk_mis = (int)((reinterpret_cast<uintptr_t>(chunk) >> 2) & 3);
  const int a_off = (chunk_mis + base - carry_len) & 3;
  const int s0 = base - a_off;
  const int nv = (a_off + (nF - 1) * S + L + 3) >> 2;
  float4 x[kQuadVecs];
#pragma unroll
  for (int u = 0; u < kQuadVecs; ++u) {
    const int s = s0 + 4 * (lane + 32 * u);
    x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane + 32 * u < nv) {
      if (s >= carry_len && s + 3 < n) x[u] = __ldg(reinterpret_cast<const float4*>(chunk + (s - carry_len)));
      else x[u] = make_float4(at(s), at(s + 1), at(s + 2), at(s + 3));
    }
  }
  float4* buf4 = reinterpret_cast<float4*>(buf);
#pragma unroll
  for (int u = 0; u < kQuadVecs; ++u)
    if (lane + 32 * u < nv) buf4[lane + 32 * u] = x[u];
  return a_off;
}

__global__ void stream_reset_kernel(void* state, StreamLayout lay, const int* ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = ids ? ids[i] : i;
  if (s < 0 || s >= lay.n_streams) return;
  int* c = lay.counters(state);
  c[s] = 0;
  c[lay.n_streams + s] = 0;
  c[2 * lay.n_streams + s] = 0;
  c[3 * lay.n_streams + s] = 0;
}

// WARPS: warps per stream-CTA.  4 is the default; 2 halves the CTA's shared memory (49 KB for 600 ms chunks: 4 CTAs per SM)
// so that 445..592 streams still run as ONE wave on 148 SMs instead of a full wave plus a mostly empty one.
template <int NROWS, bool EXACT, bool DITHER, class MELS, bool PERQUAD, int WARPS = kWarps>
__global__ void __launch_bounds__(32 * WARPS, PERQUAD ? (WARPS == 2 ? 4 : 3) : 2)
stream_push_kernel(const StreamParams p) {
  constexpr int kWarps = WARPS;                 // shadows the namespace constants inside this kernel
  constexpr int kCtaThreads = 32 * WARPS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* e_s = reinterpret_cast<float*>(smem_raw);                       // PERQUAD: [kWarps][kQuadBuf] sample buffers
  float4* xbuf = reinterpret_cast<float4*>(e_s + (PERQUAD ? kWarps * kQuadBuf : p.e_cap));
  float* logmel_s = reinterpret_cast<float*>(xbuf + kWarps * kYWarpF4);
  float2* tw_s = reinterpret_cast<float2*>(logmel_s + (p.lay.cache_cap + p.nf_max) * p.n_mels);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid & (kGroup - 1), grp_in_warp = lane >> 4;
  const int L = p.frame_len, S = p.frame_shift, M = p.n_mels;
  const int D = p.lfr_m * M, lfr_left = (p.lfr_m - 1) / 2;
  const int b = blockIdx.x;
  const int sid = p.stream_ids[b];
  if (sid < 0 || sid >= p.lay.n_streams) {
    if (tid == 0) {
      p.rows_out[b] = 0;
      if (p.chunk_stats) p.chunk_stats[2 * b] = p.chunk_stats[2 * b + 1] = 0.f;
    }
    return;
  }
  int* cnt = p.lay.counters(p.state);
  const int NS = p.lay.n_streams;
  const int carry_len = cnt[sid], t_seen = cnt[NS + sid], rows_done = cnt[2 * NS + sid], cache_len = cnt[3 * NS + sid];
  float* carry = p.lay.carry(p.state) + (size_t)sid * p.lay.carry_cap;
  float* cache = p.lay.cache(p.state) + (size_t)sid * p.lay.cache_cap * M;
  const float* chunk = p.chunks + (long long)b * p.chunk_stride;
  const int n_new = min(max(p.chunk_lens[b], 0), p.max_chunk);
  const bool fin = p.is_final && p.is_final[b];

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);

  // splice frames of earlier ticks come first in the tile's log-mel buffer
  for (int i = tid; i < cache_len * M; i += kCtaThreads) logmel_s[i] = cache[i];

  const int n = carry_len + n_new;
  const int nf = n >= L ? (n - L) / S + 1 : 0;
  if constexpr (!PERQUAD) {
    // stage [carry | chunk], raw
    for (int i = tid; i < n; i += kCtaThreads) e_s[i] = i < carry_len ? carry[i] : chunk[i - carry_len];
  }
  // samples that stay behind for the next tick (read before anything overwrites the carry)
  const int new_carry = n - nf * S;
  constexpr int kKeep = (512 + kCtaThreads - 1) / kCtaThreads;   // the carry is shorter than one frame (<= 512 samples)
  float keep[kKeep];
#pragma unroll
  for (int q = 0; q < kKeep; ++q) {
    const int k = tid + q * kCtaThreads;
    keep[q] = 0.f;
    if (k < new_carry) {
      const int src = nf * S + k;
      keep[q] = src < carry_len ? carry[src] : chunk[src - carry_len];
    }
  }
  __syncthreads();

  float4* yg = xbuf + warp * kYWarpF4 + grp_in_warp * kYGroupF4;
  float4* pbuf4 = xbuf + warp * kYWarpF4;
  const float2* tw_row = fft_twiddle_row<NROWS>(tw_s, j, grp_in_warp);
  const float2* c0_row = fft_c0_row(tw_s, j);
  if constexpr (PERQUAD) {
    // every warp fetches, transforms and mel-projects its own quads: no CTA-wide staging
    float* buf = e_s + warp * kQuadBuf;
    const int g = NROWS < 32 ? grp_in_warp : 0;
    for (int quad = warp; 4 * quad < nf; quad += kWarps) {
      const int nFq = min(4, nf - 4 * quad);
      const int a_off = stream_quad_samples(carry, carry_len, chunk, n, 4 * quad * S, nFq, S, L, lane, buf);
      __syncwarp();
      const int fA = 2 * grp_in_warp;
      const bool vA = fA < nFq, vB = fA + 1 < nFq;
      {
        f2 zr[16], zi[16], y0, y16;
        quad_stage1<NROWS, EXACT, DITHER>(buf + a_off + fA * S, vA, vB, S, L, win, p.preemph, p.remove_dc, p.dither,
                                          p.seed, (unsigned)sid, (unsigned)(t_seen + 4 * quad + fA), j, g, zr, zi, y0, y16);
        __syncwarp();
        quad_stage2(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
      }
      float* dst = logmel_s + (cache_len + 4 * quad) * M;
      mel_stage<MELS>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b2, float c, float d) {
        if (nFq > 0) dst[iv] = a;
        if (nFq > 1) dst[M + iv] = b2;
        if (nFq > 2) dst[2 * M + iv] = c;
        if (nFq > 3) dst[3 * M + iv] = d;
      });
      __syncwarp();
    }
  } else {
    for (int quad = warp; 4 * quad < nf; quad += kWarps)
      fbank_quad<NROWS, EXACT, DITHER, MELS>(e_s, nf, quad, S, L, win, yg, pbuf4, tw_row, c0_row, mel, M,
                                             p.preemph, p.remove_dc, p.log_floor, p.dither, p.seed, (unsigned)sid,
                                             (unsigned)t_seen, logmel_s + cache_len * M, j, grp_in_warp, lane);
  }
  __syncthreads();

#pragma unroll
  for (int q = 0; q < kKeep; ++q) {
    const int k = tid + q * kCtaThreads;
    if (k < new_carry) carry[k] = keep[q];
  }

  // rows that became available
  const int T = t_seen + nf;
  const int base_abs = t_seen - cache_len;         // absolute index of logmel_s row 0
  const int need = p.lfr_m - 1 - lfr_left;
  const int rows_all = T > 0 ? (T + p.lfr_n - 1) / p.lfr_n : 0;
  int rows_total = fin ? rows_all : (T - 1 >= need ? (T - 1 - need) / p.lfr_n + 1 : 0);
  rows_total = min(rows_total, rows_all);
  rows_total = max(rows_total, rows_done);
  int n_emit = rows_total - rows_done;
  // b200fe_stream_push refuses rows_cap < b200fe_stream_max_rows, so this cannot trigger through the ABI; if it ever
  // does, the rows that do not fit stay pending (the counters below follow n_emit) instead of being lost
  const bool truncated = n_emit > p.rows_cap;
  if (truncated) {
    n_emit = (int)p.rows_cap;
    rows_total = rows_done + n_emit;
  }
  {
    float* out = p.feats + (long long)b * p.rows_cap * D;
    const int D4 = D >> 2, M4 = M >> 2;
    for (int c4 = tid; c4 < D4; c4 += kCtaThreads) {
      const int cj = c4 / M4, cd = c4 - cj * M4;
      float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f);
      if (p.cmvn) {
        sh = *reinterpret_cast<const float4*>(p.cmvn + 4 * c4);
        sc = *reinterpret_cast<const float4*>(p.cmvn + D + 4 * c4);
      }
      for (int r = 0; r < n_emit; ++r) {
        int f = p.lfr_n * (rows_done + r) + cj - lfr_left;
        f = min(max(f, 0), T - 1) - base_abs;
        const float4 v = *reinterpret_cast<const float4*>(logmel_s + f * M + 4 * cd);
        float4 o;
        o.x = (v.x + sh.x) * sc.x;
        o.y = (v.y + sh.y) * sc.y;
        o.z = (v.z + sh.z) * sc.z;
        o.w = (v.w + sh.w) * sc.w;
        *reinterpret_cast<float4*>(out + (long long)r * D + 4 * c4) = o;
      }
    }
  }
  // frames the next rows still need (always at least the newest frame, for right replication on the final flush)
  int keep_from = max(rows_total * p.lfr_n - lfr_left, 0);
  keep_from = min(keep_from, max(T - 1, 0));
  keep_from = max(keep_from, base_abs);
  keep_from = max(keep_from, T - p.lay.cache_cap);
  const int new_cache = T > 0 ? T - keep_from : 0;
  for (int i = tid; i < new_cache * M; i += kCtaThreads) cache[i] = logmel_s[(keep_from - base_abs) * M + i];

  // The reference's per-chunk energy gate (R:voice-service/app/services/voice_interface.py:1569-1578: mean |x| and
  // max |x| of the chunk against two thresholds) as a by-product of the tick: the chunk was just read by the framing
  // pass (L2-resident), the CTA reduces it once more and returns both numbers next to the rows.
  if (p.chunk_stats) {
    float sa = 0.f, mx = 0.f;
    for (int i = tid; i < n_new; i += kCtaThreads) {
      const float a = fabsf(chunk[i]);
      sa += a;
      mx = fmaxf(mx, a);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __syncthreads();                       // logmel_s is free: every thread is past the row / splice copies above
    if (lane == 0) { logmel_s[warp] = sa; logmel_s[kWarps + warp] = mx; }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      float m = 0.f;
      for (int w = 0; w < kWarps; ++w) { tot += (double)logmel_s[w]; m = fmaxf(m, logmel_s[kWarps + w]); }
      p.chunk_stats[2 * b] = n_new > 0 ? (float)(tot / (double)n_new) : 0.f;
      p.chunk_stats[2 * b + 1] = m;
    }
  }

  if (tid == 0) {
    p.rows_out[b] = n_emit;
    if (fin && !truncated) {
      cnt[sid] = 0; cnt[NS + sid] = 0; cnt[2 * NS + sid] = 0; cnt[3 * NS + sid] = 0;
    } else {
      cnt[sid] = new_carry; cnt[NS + sid] = T; cnt[2 * NS + sid] = rows_total; cnt[3 * NS + sid] = new_cache;
    }
  }
}

}  // namespace b200fe
