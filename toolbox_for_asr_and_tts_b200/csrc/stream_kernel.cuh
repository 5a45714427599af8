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
  int cache_cap;    // frames a stream carries from tick to tick: max(lfr_m - 1, 1)
  int frames_cap;   // frames of a stream's log-mel buffer: cache_cap + the frames one push can create
  int n_mels;
  __host__ __device__ size_t counters_bytes() const { return ((size_t)4 * n_streams * sizeof(int) + 255) & ~(size_t)255; }
  __host__ __device__ size_t carry_bytes() const { return ((size_t)n_streams * carry_cap * sizeof(float) + 255) & ~(size_t)255; }
  __host__ __device__ size_t cache_bytes() const { return ((size_t)n_streams * frames_cap * n_mels * sizeof(float) + 255) & ~(size_t)255; }
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
// [carry | chunk] buffer.  The log-mel frames of the tick go to the stream's buffer in the state slab (L2-resident), not
// to shared memory: 53 KB per 4-warp CTA, i.e. 4 stream-CTAs (16 warps) per SM and 592 streams per wave.
__host__ __device__ inline size_t stream_smem_bytes(int e_cap, bool per_quad) {
  size_t b = 0;
  if (per_quad) {
    b += (size_t)kWarps * kQuadBuf * 4;
  } else {
    b += (size_t)e_cap * 4;
  }
  b += (size_t)kWarps * kYWarpF4 * 16;
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

template <int NROWS, bool EXACT, bool DITHER, class MELS, bool PERQUAD>
__global__ void __launch_bounds__(kCtaThreads, PERQUAD ? 4 : 2)
stream_push_kernel(const StreamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* e_s = reinterpret_cast<float*>(smem_raw);                       // PERQUAD: [kWarps][kQuadBuf] sample buffers
  float4* xbuf = reinterpret_cast<float4*>(e_s + (PERQUAD ? kWarps * kQuadBuf : p.e_cap));
  float2* tw_s = reinterpret_cast<float2*>(xbuf + kWarps * kYWarpF4);
  __shared__ float red_s[2 * kWarps];

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
  // the stream's log-mel buffer: rows 0 .. cache_len-1 are the splice frames of earlier ticks, this tick's frames follow
  float* logmel_s = p.lay.cache(p.state) + (size_t)sid * p.lay.frames_cap * M;
  const float* chunk = p.chunks + (long long)b * p.chunk_stride;
  const int n_new = min(max(p.chunk_lens[b], 0), p.max_chunk);
  const bool fin = p.is_final && p.is_final[b];
  // the whole chunk is on its way into L2 while the CTA sets up: the quads' sample fetches then see L2 latency only
  for (int i = 32 * tid; i < n_new; i += 32 * kCtaThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(chunk + i));

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);

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
    // (row, 128-bit piece) pairs dealt to the threads, four at a time: all loads of a batch are issued before its
    // arithmetic and stores (the log-mel buffer is in L2, not in shared memory)
    float* out = p.feats + (long long)b * p.rows_cap * D;
    const int D4 = D >> 2, M4 = M >> 2;
    const int total4 = n_emit * D4;
    constexpr int kBatch = 4;
    for (int i0 = tid; i0 < total4; i0 += kBatch * kCtaThreads) {
      float4 v[kBatch], sh[kBatch], sc[kBatch];
      int dst[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const int idx = i0 + k * kCtaThreads;
        dst[k] = -1;
        if (idx < total4) {
          const int r = idx / D4, c4 = idx - r * D4;
          const int cj = c4 / M4, cd = c4 - cj * M4;
          int f = p.lfr_n * (rows_done + r) + cj - lfr_left;
          f = min(max(f, 0), T - 1) - base_abs;
          v[k] = __ldcg(reinterpret_cast<const float4*>(logmel_s + f * M + 4 * cd));
          sh[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          sc[k] = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.cmvn) {
            sh[k] = __ldg(reinterpret_cast<const float4*>(p.cmvn + 4 * c4));
            sc[k] = __ldg(reinterpret_cast<const float4*>(p.cmvn + D + 4 * c4));
          }
          dst[k] = r * D + 4 * c4;
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k)
        if (dst[k] >= 0) {
          float4 o;
          o.x = (v[k].x + sh[k].x) * sc[k].x;
          o.y = (v[k].y + sh[k].y) * sc[k].y;
          o.z = (v[k].z + sh[k].z) * sc[k].z;
          o.w = (v[k].w + sh[k].w) * sc[k].w;
          *reinterpret_cast<float4*>(out + dst[k]) = o;
        }
    }
  }
  // frames the next rows still need (always at least the newest frame, for right replication on the final flush)
  int keep_from = max(rows_total * p.lfr_n - lfr_left, 0);
  keep_from = min(keep_from, max(T - 1, 0));
  keep_from = max(keep_from, base_abs);
  keep_from = max(keep_from, T - p.lay.cache_cap);
  const int new_cache = T > 0 ? T - keep_from : 0;
  // ... move to the front of the buffer, one CTA-wide slice at a time (the slices a step writes were read before its
  // barrier, the ones later steps read lie behind it; the first barrier also orders the row reads above before any write)
  if (keep_from > base_abs)
    for (int i0 = 0; i0 < new_cache * M; i0 += kCtaThreads) {
      const int i = i0 + tid;
      float v = 0.f;
      if (i < new_cache * M) v = __ldcg(logmel_s + (keep_from - base_abs) * M + i);
      __syncthreads();
      if (i < new_cache * M) logmel_s[i] = v;
    }

  // The reference's per-chunk energy gate (R:voice-service/app/services/voice_interface.py:1569-1578: mean |x| and
  // max |x| of the chunk against two thresholds) as a by-product of the tick: the chunk was just read by the framing
  // pass (L2-resident), the CTA reduces it once more and returns both numbers next to the rows.
  if (p.chunk_stats) {
    float sa = 0.f, mx = 0.f;
    const int head = min(n_new, (int)((4 - ((reinterpret_cast<uintptr_t>(chunk) >> 2) & 3)) & 3));   // up to the 16-byte grid
    const int n4 = (n_new - head) >> 2;
    auto acc = [&](float x) { const float a = fabsf(x); sa += a; mx = fmaxf(mx, a); };
    if (tid < head) acc(chunk[tid]);
    if (tid < n_new - head - 4 * n4) acc(chunk[head + 4 * n4 + tid]);
    const float4* c4p = reinterpret_cast<const float4*>(chunk + head);
    constexpr int kU = 8;
    for (int i0 = tid; i0 < n4; i0 += kU * kCtaThreads) {
      float4 x[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int i = i0 + k * kCtaThreads;
        x[k] = i < n4 ? __ldg(c4p + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) { acc(x[k].x); acc(x[k].y); acc(x[k].z); acc(x[k].w); }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { red_s[warp] = sa; red_s[kWarps + warp] = mx; }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      float m = 0.f;
      for (int w = 0; w < kWarps; ++w) { tot += (double)red_s[w]; m = fmaxf(m, red_s[kWarps + w]); }
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
