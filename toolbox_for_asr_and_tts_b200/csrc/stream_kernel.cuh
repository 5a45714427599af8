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
  int q_max;        // quads one push can create (>= 1): work items per stream and tick of stream_quad_kernel
  // counters: [0] carry length, [1] frames seen, [2] rows emitted, [3] splice frames kept
  __host__ __device__ size_t counters_bytes() const { return ((size_t)4 * n_streams * sizeof(int) + 255) & ~(size_t)255; }
  __host__ __device__ size_t partial_bytes() const { return ((size_t)n_streams * q_max * sizeof(float2) + 255) & ~(size_t)255; }
  __host__ __device__ size_t carry_bytes() const { return ((size_t)n_streams * carry_cap * sizeof(float) + 255) & ~(size_t)255; }
  __host__ __device__ size_t cache_bytes() const { return ((size_t)n_streams * frames_cap * n_mels * sizeof(float) + 255) & ~(size_t)255; }
  __host__ __device__ size_t tick_bytes() const { return ((size_t)n_streams * 32 + 255) & ~(size_t)255; }
  __host__ __device__ size_t total_bytes() const { return counters_bytes() + carry_bytes() + cache_bytes() + partial_bytes() + tick_bytes(); }
  // per pushed chunk of the current tick: two int4 written by stream_tick_prep_kernel (StreamTick below)
  __host__ __device__ int4* tick(void* base) const {
    return reinterpret_cast<int4*>((char*)base + counters_bytes() + carry_bytes() + cache_bytes() + partial_bytes());
  }
  // per work item: (sum |x|, max |x|) of its share of the chunk, reduced in a fixed order by stream_tick_finish_kernel
  __host__ __device__ float2* partial(void* base) const {
    return reinterpret_cast<float2*>((char*)base + counters_bytes() + carry_bytes() + cache_bytes());
  }
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
  b += (size_t)kWarps * 8;      // one mbarrier per warp (stream_quad_kernel's bulk copies)
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
  mel_preload(mel, threadIdx.x & 31);
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
  const float2* c0_row = fft_c0_row(tw_s, j);   // StreamParams::twiddle carries the [8][kC0Pitch] column-0 table
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
        quad_stage2<false>(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
      }
      float* dst = logmel_s + (cache_len + 4 * quad) * M;
      mel_stage<MELS, false>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b2, float c, float d) {
        if (nFq > 0) dst[iv] = a;
        if (nFq > 1) dst[M + iv] = b2;
        if (nFq > 2) dst[2 * M + iv] = c;
        if (nFq > 3) dst[3 * M + iv] = d;
      });
      __syncwarp();
    }
  } else {
    for (int quad = warp; 4 * quad < nf; quad += kWarps)
      fbank_quad<NROWS, EXACT, DITHER, MELS, false>(e_s, nf, quad, S, L, win, yg, pbuf4, tw_row, c0_row, mel, M,
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

// ---------------------------------------------------------------------------------------------------------------------
// Quad-level streaming tick (the shipped path whenever a quad fits the warp buffer): three launches.  The CTA-per-stream
// kernel above is latency-bound: all stream-CTAs start together and walk through the same phases (state reads, quads,
// row pass, state update), so nothing covers the phases' latencies.  Here
//   1. stream_tick_prep_kernel resolves every chunk's geometry once (stream id -> counters -> frames, rows: a chain of
//      dependent loads and integer divisions) into a 32-byte descriptor;
//   2. stream_quad_kernel treats the tick as a flat list of work items (chunk b, quad q), q < q_max, over which
//      persistent warps stride exactly like fbank_warp_kernel's warps stride over an utterance batch.  An item reads its
//      descriptor one item ahead (cp.async into shared memory), fetches its quad's samples from [carry | chunk] - by one
//      bulk copy (cp.async.bulk) issued an item ahead when the quad lies inside the chunk, else by the lanes -, runs
//      stage 1 / stage 2 / mel, appends its log-mel frames to the stream's buffer in the state slab and SCATTERS them
//      straight into the rows of this tick (frame f is slot jj of row i wherever clamp(n i + jj - left, 0, T-1) == f):
//      there is no row pass.  Item 0 also scatters the splice frames kept from earlier ticks and owns the chunk when no
//      frame completes; every item leaves (sum |x|, max |x|) of its share of the chunk, taken from the shared-memory
//      copy of the samples.  The state is only read;
//   3. stream_tick_finish_kernel (one warp per chunk) updates the state: new sample carry, splice frames moved to the
//      front, counters, the energy gate reduced in a fixed order (deterministic), the row count.
// Streams must be distinct within one call (as before).
struct StreamTick {   // two int4
  int sid;          // < 0: invalid stream id
  int t_seen;       // frames of earlier ticks
  int rows_done;
  int n_new;        // samples of the chunk (clamped to max_chunk)
  int lens;         // carry_len | cache_len << 16
  int nf;           // frames completed by this tick
  int rows_total;   // rows emitted after this tick
  int flags;        // bit 0: final, bit 1: truncated (rows_cap too small: cannot happen through the ABI)
};

__global__ void stream_tick_prep_kernel(const StreamParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.n) return;
  const int NS = p.lay.n_streams;
  StreamTick t;
  t.sid = p.stream_ids[b];
  t.t_seen = t.rows_done = t.n_new = t.lens = t.nf = t.rows_total = t.flags = 0;
  if (t.sid >= 0 && t.sid < NS) {
    const int* cnt = p.lay.counters(p.state);
    const int carry_len = cnt[t.sid], cache_len = cnt[3 * NS + t.sid];
    t.t_seen = cnt[NS + t.sid];
    t.rows_done = cnt[2 * NS + t.sid];
    t.n_new = min(max(p.chunk_lens[b], 0), p.max_chunk);
    const bool fin = p.is_final && p.is_final[b];
    const int n = carry_len + t.n_new;
    t.nf = n >= p.frame_len ? (n - p.frame_len) / p.frame_shift + 1 : 0;
    const int T = t.t_seen + t.nf;
    const int need = p.lfr_m - 1 - (p.lfr_m - 1) / 2;
    const int rows_all = T > 0 ? (T + p.lfr_n - 1) / p.lfr_n : 0;
    int rows_total = fin ? rows_all : (T - 1 >= need ? (T - 1 - need) / p.lfr_n + 1 : 0);
    rows_total = min(rows_total, rows_all);
    rows_total = max(rows_total, t.rows_done);
    const bool truncated = rows_total - t.rows_done > p.rows_cap;
    if (truncated) rows_total = t.rows_done + (int)p.rows_cap;
    t.rows_total = rows_total;
    t.lens = carry_len | (cache_len << 16);
    t.flags = (fin ? 1 : 0) | (truncated ? 2 : 0);
  } else {
    t.sid = -1;
    p.rows_out[b] = 0;
    if (p.chunk_stats) p.chunk_stats[2 * b] = p.chunk_stats[2 * b + 1] = 0.f;
  }
  int4* d = p.lay.tick(p.state) + 2 * b;
  d[0] = make_int4(t.sid, t.t_seen, t.rows_done, t.n_new);
  d[1] = make_int4(t.lens, t.nf, t.rows_total, t.flags);
}

// The geometry of one work item, re-read from the warp's shared-memory copy of the descriptor wherever it is needed
// (kept in registers across the FFT it is spilled, and spill reloads miss L1: the local memory of an SM does not fit).
struct TickView {
  int sid, t_seen, rows_done, n_new, carry_len, cache_len, nf, rows_total, fin, truncated;
  template <class P>
  __device__ __forceinline__ explicit TickView(P d) {
    sid = d[0]; t_seen = d[1]; rows_done = d[2]; n_new = d[3];
    const int lens = d[4];
    carry_len = lens & 0xffff; cache_len = lens >> 16;
    nf = d[5]; rows_total = d[6];
    const int fl = d[7];
    fin = fl & 1; truncated = (fl >> 1) & 1;
  }
};

// State update after the quads of a tick, one warp per chunk: new sample carry, splice frames to the front of the log-mel
// buffer, the energy gate reduced in a fixed order, counters, row count.
__global__ void __launch_bounds__(kCtaThreads)
stream_tick_finish_kernel(const StreamParams p) {
  const int b = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= p.n) return;
  int dl[8];
  {
    const int4 da = p.lay.tick(p.state)[2 * b], db = p.lay.tick(p.state)[2 * b + 1];
    dl[0] = da.x; dl[1] = da.y; dl[2] = da.z; dl[3] = da.w; dl[4] = db.x; dl[5] = db.y; dl[6] = db.z; dl[7] = db.w;
  }
  const TickView t(dl);
  if (t.sid < 0) return;
  const int S = p.frame_shift, M = p.n_mels, NS = p.lay.n_streams;
  const int lfr_left = (p.lfr_m - 1) / 2;
  int* cnt = p.lay.counters(p.state);
  float* carry = p.lay.carry(p.state) + (size_t)t.sid * p.lay.carry_cap;
  float* logmel = p.lay.cache(p.state) + (size_t)t.sid * p.lay.frames_cap * M;
  const float* chunk = p.chunks + (long long)b * p.chunk_stride;
  const int n = t.carry_len + t.n_new, T = t.t_seen + t.nf, base_abs = t.t_seen - t.cache_len;
  // samples that stay behind for the next tick (read completely before the carry is overwritten)
  const int new_carry = n - t.nf * S;
  constexpr int kKeep = 512 / 32;                    // the carry is shorter than one frame (<= 512 samples)
  float keep[kKeep];
#pragma unroll
  for (int k = 0; k < kKeep; ++k) {
    const int i = lane + 32 * k;
    keep[k] = 0.f;
    if (i < new_carry) {
      const int src = t.nf * S + i;
      keep[k] = src < t.carry_len ? __ldcg(carry + src) : __ldg(chunk + (src - t.carry_len));
    }
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < kKeep; ++k) {
    const int i = lane + 32 * k;
    if (i < new_carry) carry[i] = keep[k];
  }
  // frames the next rows still need (always at least the newest frame, for right replication on the final flush)
  int keep_from = max(t.rows_total * p.lfr_n - lfr_left, 0);
  keep_from = min(keep_from, max(T - 1, 0));
  keep_from = max(keep_from, base_abs);
  keep_from = max(keep_from, T - p.lay.cache_cap);
  const int new_cache = T > 0 ? T - keep_from : 0;
  if (keep_from > base_abs) {   // to the front, 8 x 32 floats at a time (a step's sources lie behind its destinations)
    const float* src = logmel + (size_t)(keep_from - base_abs) * M;
#pragma unroll 1
    for (int i0 = 0; i0 < new_cache * M; i0 += 256) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + lane + 32 * k;
        v[k] = i < new_cache * M ? __ldcg(src + i) : 0.f;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + lane + 32 * k;
        if (i < new_cache * M) logmel[i] = v[k];
      }
    }
  }
  if (p.chunk_stats) {          // fixed-order reduction of the items' partial sums
    const int items_b = max((t.nf + 3) >> 2, 1);
    const float2* part = p.lay.partial(p.state) + (size_t)t.sid * p.lay.q_max;
    double tot = 0.0;
    float m = 0.f;
    for (int k = 0; k < items_b; ++k) {
      const float2 pk = __ldcg(part + k);
      tot += (double)pk.x;
      m = fmaxf(m, pk.y);
    }
    if (lane == 0) {
      p.chunk_stats[2 * b] = t.n_new > 0 ? (float)(tot / (double)t.n_new) : 0.f;
      p.chunk_stats[2 * b + 1] = m;
    }
  }
  if (lane == 0) {
    p.rows_out[b] = t.rows_total - t.rows_done;
    if (t.fin && !t.truncated) {
      cnt[t.sid] = 0; cnt[NS + t.sid] = 0; cnt[2 * NS + t.sid] = 0; cnt[3 * NS + t.sid] = 0;
    } else {
      cnt[t.sid] = new_carry; cnt[NS + t.sid] = T; cnt[2 * NS + t.sid] = t.rows_total; cnt[3 * NS + t.sid] = new_cache;
    }
  }
}

// Frame f (absolute index) with this lane's 128-bit piece v -> its (row, slot) places among rows_done .. rows_total-1.
__device__ __forceinline__ void stream_emit_frame(int f, const float4& v, int T, int fin, int rows_done, int rows_total,
                                                  int lfr_m, int lfr_n, int M, const float* cm_l, float* out_l) {
  const int lfr_left = (lfr_m - 1) / 2, D = lfr_m * M;
  auto put = [&](int i, int jj) {
    float4 o = v;
    if (cm_l) {
      const float4 sh = __ldg(reinterpret_cast<const float4*>(cm_l + jj * M));
      const float4 sc = __ldg(reinterpret_cast<const float4*>(cm_l + D + jj * M));
      o = make_float4((v.x + sh.x) * sc.x, (v.y + sh.y) * sc.y, (v.z + sh.z) * sc.z, (v.w + sh.w) * sc.w);
    }
    stg_stream4(out_l + (long long)(i - rows_done) * D + jj * M, o);
  };
  if (f == 0 || (fin && f == T - 1)) {   // replicated by the LFR padding: every place that clamps to f
    const int num = f + lfr_left - (lfr_m - 1);
    const int i_lo = max((f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n, rows_done);
    const int i_hi = (f == T - 1 ? rows_total - 1 : min((f + lfr_left) / lfr_n, rows_total - 1));
#pragma unroll 1
    for (int i = i_lo; i <= i_hi; ++i)
#pragma unroll 1
      for (int jj = 0; jj < lfr_m; ++jj)
        if (min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) put(i, jj);
  } else {                              // slot jj = f + left - n i of rows i = i_top, i_top - 1, ...
    const int i_top = (f + lfr_left) / lfr_n;
    int jj = f + lfr_left - lfr_n * i_top;
#pragma unroll 1
    for (int i = i_top; jj < lfr_m && i >= rows_done; --i, jj += lfr_n)
      if (i < rows_total) put(i, jj);
  }
}

template <int NROWS, bool EXACT, bool DITHER, class MELS>
__global__ void __launch_bounds__(kCtaThreads, 4)
stream_quad_kernel(const StreamParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* bufs = reinterpret_cast<float*>(smem_raw);
  float4* xbuf = reinterpret_cast<float4*>(bufs + kWarps * kQuadBuf);
  float2* tw_s = reinterpret_cast<float2*>(xbuf + kWarps * kYWarpF4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tw_s + kTw2Total);
  __shared__ __align__(16) int desc_s[kWarps][2][8];   // per warp: the descriptors of the current and of the next item

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid & (kGroup - 1), grp_in_warp = lane >> 4;
  const int g = NROWS < 32 ? grp_in_warp : 0;
  const int L = p.frame_len, S = p.frame_shift, M = p.n_mels;

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  if (lane == 0) {
    mbar_init(bars + warp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
  mel_preload(mel, threadIdx.x & 31);
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);
  __syncthreads();   // the only CTA-wide barrier: the twiddle tables and the mbarriers

  float* buf = bufs + warp * kQuadBuf;
  unsigned long long* bar = bars + warp;
  float4* yg = xbuf + warp * kYWarpF4 + grp_in_warp * kYGroupF4;
  float4* pbuf4 = xbuf + warp * kYWarpF4;
  float* lm_s = reinterpret_cast<float*>(pbuf4 + kSpecF4);   // log-mel staging tile, behind the warp's spectra
  const float2* tw_row = fft_twiddle_row<NROWS>(tw_s, j, grp_in_warp);
  const float2* c0_row = fft_c0_row(tw_s, j);   // StreamParams::twiddle carries the [8][kC0Pitch] column-0 table
  const int4* ticks = p.lay.tick(p.state);

  const int q_max = p.lay.q_max;
  const int n_items = p.n * q_max;
  const int W = gridDim.x * kWarps;
  int item = blockIdx.x * kWarps + warp;
  if (item >= n_items) return;

  // The whole warp: start the bulk copy of (chunk b, quad q) described by d when the quad lies inside the chunk.
  // Returns 1 + a_off when a copy is in flight, 0 when the lanes have to fill the buffer (the quad touches the carry)
  // or there is nothing to fetch.
  auto issue_copy = [&](int b, int q, const volatile int* d) -> int {
    const int sid = d[0], carry_len = d[4] & 0xffff, nf = d[5];
    if (sid < 0 || 4 * q >= nf || 4 * q * S < carry_len) return 0;
    const long long g0 = (long long)b * p.chunk_stride + (4 * q * S - carry_len);
    int ok = 0;
    if (lane == 0)
      ok = quad_fill_tma(p.chunks, (long long)p.n * p.chunk_stride, g0, (min(4, nf - 4 * q) - 1) * S + L, buf, bar) ? 1 : 0;
    ok = __shfl_sync(0xffffffffu, ok, 0);
    return ok ? 1 + quad_a_off<float>(p.chunks, g0) : 0;
  };

  int cur = 0;
  int b = item / q_max, q = item - b * q_max;
  if (lane < 2) reinterpret_cast<int4*>(desc_s[warp][0])[lane] = __ldg(ticks + 2 * b + lane);
  __syncwarp();
  unsigned phase = 0;
  int in_flight = issue_copy(b, q, desc_s[warp][0]);
  const int lfr_m = p.lfr_m, lfr_n = p.lfr_n, lfr_left = (lfr_m - 1) / 2, D = lfr_m * M, M4 = M >> 2;
  const bool act = lane < M4;      // lanes that move a float4 of a log-mel row

#pragma unroll 1
  while (true) {
    const volatile int* d = desc_s[warp][cur];
    const volatile int* dn = desc_s[warp][cur ^ 1];
    // the next item of this warp: its descriptor travels into shared memory now (cp.async: no registers) and is looked at
    // after stage 1
    const int item_n = item + W;
    const bool have_next = item_n < n_items;
    const int bn = have_next ? item_n / q_max : b, qn = have_next ? item_n - bn * q_max : q;
    if (have_next && lane < 2) {
      const unsigned dst = smem_u32(const_cast<const int*>(dn) + 4 * lane);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(ticks + 2 * bn + lane) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    auto next_copy = [&]() -> int {   // called by the whole warp once every lane is done with the sample buffer
      if (!have_next) return 0;
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      return issue_copy(bn, qn, dn);
    };

    bool work, frames;
    {
      const int sid = d[0], nf = d[5];
      work = sid >= 0 && q < max((nf + 3) >> 2, 1);
      frames = work && 4 * q < nf;
    }
    if (frames) {
      int a_off, nFq;
      {
        const TickView t(d);
        nFq = min(4, t.nf - 4 * q);
        const int n = t.carry_len + t.n_new;
        const float* chunk = p.chunks + (long long)b * p.chunk_stride;
        if (in_flight) {
          mbar_wait(bar, phase);
          phase ^= 1u;
          a_off = in_flight - 1;
        } else {
          const float* carry = p.lay.carry(p.state) + (size_t)t.sid * p.lay.carry_cap;
          a_off = stream_quad_samples(carry, t.carry_len, chunk, n, 4 * q * S, nFq, S, L, lane, buf);
          __syncwarp();
        }
        // this quad's share of the energy gate: virtual samples [4 q S, 4 (q+1) S), the last quad up to the end; all but
        // the last (< S) samples of the last quad's share are in the buffer
        if (p.chunk_stats) {
          float sa = 0.f, mx = 0.f;
          auto acc = [&](float x) { const float a = fabsf(x); sa += a; mx = fmaxf(mx, a); };
          const int v0 = 4 * q * S;
          const int v_lo = max(v0, t.carry_len), v_hi = 4 * (q + 1) >= t.nf ? n : min(v0 + 4 * S, n);
          const int v_buf = min(v_hi, v0 + (nFq - 1) * S + L);
          for (int v = v_lo + lane; v < v_buf; v += 32) acc(buf[a_off + (v - v0)]);
          for (int v = max(v_buf, v_lo) + lane; v < v_hi; v += 32) acc(__ldg(chunk + (v - t.carry_len)));
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          }
          if (lane == 0) p.lay.partial(p.state)[(size_t)t.sid * q_max + q] = make_float2(sa, mx);
        }
      }
      const int fA = 2 * grp_in_warp;
      const bool vA = fA < nFq, vB = fA + 1 < nFq;
      {
        f2 zr[16], zi[16], y0, y16;
        unsigned dsid = 0, dframe = 0;
        if constexpr (DITHER) { dsid = (unsigned)d[0]; dframe = (unsigned)(d[1] + 4 * q + fA); }
        quad_stage1<NROWS, EXACT, DITHER>(buf + a_off + fA * S, vA, vB, S, L, win, p.preemph, p.remove_dc, p.dither,
                                          p.seed, dsid, dframe, j, g, zr, zi, y0, y16);
        __syncwarp();   // every lane is done with the sample buffer and with the previous item's staging tile
        in_flight = next_copy();
        quad_stage2<false>(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
      }
      mel_stage<MELS, false>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b2, float c, float dd) {
        lm_s[iv] = a;
        lm_s[M + iv] = b2;
        lm_s[2 * M + iv] = c;
        lm_s[3 * M + iv] = dd;
      });
      __syncwarp();
      // ---- LFR + CMVN: where the 4 frames go among the rows of this tick (as QuadDesc::tgt of the offline kernel, but
      //      resolved here: one division per quad), then all loads, then arithmetic and stores
      {
        const TickView t(d);
        const int T = t.t_seen + t.nf, f0 = t.t_seen + 4 * q;
        unsigned tgt[8];
        int slow = 0;
        {
          int i_top = (f0 + lfr_left) / lfr_n, jj = f0 + lfr_left - lfr_n * i_top;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tgt[2 * k] = tgt[2 * k + 1] = kNoTarget;
            if (k < nFq) {
              const int f = f0 + k;
              if (f == 0 || (t.fin && f == T - 1) || jj + 2 * lfr_n < lfr_m) {
                slow |= 1 << k;
              } else {
                if (i_top >= t.rows_done && i_top < t.rows_total)
                  tgt[2 * k] = ((unsigned)jj << kTargetOffBits) | (unsigned)((i_top - t.rows_done) * D + jj * M);
                if (i_top - 1 >= t.rows_done && i_top - 1 < t.rows_total && jj + lfr_n < lfr_m)
                  tgt[2 * k + 1] = ((unsigned)(jj + lfr_n) << kTargetOffBits) | (unsigned)((i_top - 1 - t.rows_done) * D + (jj + lfr_n) * M);
              }
            }
            if (++jj == lfr_n) { jj = 0; ++i_top; }
          }
        }
        const float* cm_l = p.cmvn ? p.cmvn + 4 * lane : nullptr;
        float* out_l = p.feats + (long long)b * p.rows_cap * D + 4 * lane;
        const float4* lm4 = reinterpret_cast<const float4*>(lm_s) + lane;
        auto cmvn4 = [&](const float4& v, const float4& sh, const float4& sc) {   // (x + shift) * scale, VF:34-35
          return make_float4((v.x + sh.x) * sc.x, (v.y + sh.y) * sc.y, (v.z + sh.z) * sc.z, (v.w + sh.w) * sc.w);
        };
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        constexpr unsigned kOffMask = (1u << kTargetOffBits) - 1;
        float4 v[4], sh[4], sc[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ok[k] = tgt[2 * k] != kNoTarget && act;
          v[k] = act ? lm4[k * M4] : zero4;
          sh[k] = sc[k] = zero4;
          if (cm_l && ok[k]) {
            const int jm = (int)(tgt[2 * k] >> kTargetOffBits) * M;
            sh[k] = __ldg(reinterpret_cast<const float4*>(cm_l + jm));
            sc[k] = __ldg(reinterpret_cast<const float4*>(cm_l + D + jm));
          }
        }
        if (act) {   // the frames themselves, kept for later ticks (splice frames): stream_tick_finish_kernel moves them
          float4* dst = reinterpret_cast<float4*>(p.lay.cache(p.state) + ((size_t)t.sid * p.lay.frames_cap + t.cache_len + 4 * q) * M) + lane;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < nFq) dst[k * M4] = v[k];
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ok[k]) stg_stream4(out_l + (int)(tgt[2 * k] & kOffMask), cm_l ? cmvn4(v[k], sh[k], sc[k]) : v[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const unsigned code = tgt[2 * k + 1];
          if (code != kNoTarget && act) {
            float4 o = v[k];
            if (cm_l) {
              const int jm = (int)(code >> kTargetOffBits) * M;
              o = cmvn4(o, __ldg(reinterpret_cast<const float4*>(cm_l + jm)), __ldg(reinterpret_cast<const float4*>(cm_l + D + jm)));
            }
            stg_stream4(out_l + (int)(code & kOffMask), o);
          }
        }
        if (slow && act) {   // first frame of the stream / last frame of a final flush (replicated), or lfr_m > 2 lfr_n
#pragma unroll 1
          for (int k = 0; k < nFq; ++k)
            if ((slow >> k) & 1)
              stream_emit_frame(f0 + k, lm4[k * M4], T, t.fin, t.rows_done, t.rows_total, lfr_m, lfr_n, M, cm_l, out_l);
        }
      }
    } else {
      if (work && p.chunk_stats) {   // no frame completes in this tick: item 0 owns the whole chunk
        const TickView t(d);
        const float* chunk = p.chunks + (long long)b * p.chunk_stride;
        float sa = 0.f, mx = 0.f;
        for (int c = lane; c < t.n_new; c += 32) { const float a = fabsf(__ldg(chunk + c)); sa += a; mx = fmaxf(mx, a); }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) p.lay.partial(p.state)[(size_t)t.sid * q_max + q] = make_float2(sa, mx);
      }
      __syncwarp();
      in_flight = next_copy();        // nothing was copied for this item: the buffer goes to the next one
    }
    if (work && q == 0 && act) {   // splice frames of earlier ticks -> this tick's rows, four loads at a time
      const TickView t(d);
      const int T = t.t_seen + t.nf, base_abs = t.t_seen - t.cache_len;
      const float* cm_l = p.cmvn ? p.cmvn + 4 * lane : nullptr;
      float* out_l = p.feats + (long long)b * p.rows_cap * D + 4 * lane;
      const float* logmel = p.lay.cache(p.state) + (size_t)t.sid * p.lay.frames_cap * M;
#pragma unroll 1
      for (int c0 = 0; c0 < t.cache_len; c0 += 4) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < t.cache_len) v[k] = __ldcg(reinterpret_cast<const float4*>(logmel + (size_t)(c0 + k) * M) + lane);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < t.cache_len)
            stream_emit_frame(base_abs + c0 + k, v[k], T, t.fin, t.rows_done, t.rows_total, lfr_m, lfr_n, M, cm_l, out_l);
      }
    }
    if (!have_next) break;
    __syncwarp();   // the staging tile and the descriptor are read: the next item may overwrite them
    item = item_n; b = bn; q = qn;
    cur ^= 1;
  }
}

}  // namespace b200fe
