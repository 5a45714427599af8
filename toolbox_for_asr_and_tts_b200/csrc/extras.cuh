// By-products of the front-end pass that the reference computes on the host around its funasr calls:
//   * audio_stats_kernel   _log_audio_statistics / the per-chunk energy gate (R:voice-service/app/services/
//                          voice_interface.py:873-939, 1298-1300, 1569-1570): max, min, mean |x|, RMS, clipping ratio
//   * ingest_pcm_kernel    base64_to_audio_np after the WAV header (R:voice_interface.py:1004-1034): sample width
//                          normalisation, channel mean, linear-interpolation resampling (its numpy branch), float32
//   * ring_push / ring_window   the per-session sliding audio buffers (KWS 1.6 s window, 0.4 s pre-speech guard:
//                          buffer = np.concatenate([buffer, chunk])[-target:], R:voice_interface.py:1304-1311,1742-1746)
//                          as a device-resident circular buffer per stream
//   * column_mean_kernel   Kaldi subtract_mean (TA:642-644, _subtract_column_mean) = the utterance mean
//                          normalisation of the CAM++ speaker-verification features (R:voice_interface.py:2430,2520,2558)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200fe {

// order-preserving float <-> unsigned map, so that max / min reduce with integer atomics
__device__ __forceinline__ unsigned float_to_ordered(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct AudioStatsAcc {     // one per utterance, zero-initialised by audio_stats_init_kernel
  double sum_abs, sum_sq;
  unsigned long long n_clip;
  unsigned max_ord, min_ord;
};

__global__ void audio_stats_init_kernel(AudioStatsAcc* acc, int batch) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= batch) return;
  acc[u].sum_abs = 0.0;
  acc[u].sum_sq = 0.0;
  acc[u].n_clip = 0ull;
  acc[u].max_ord = 0u;            // below every float
  acc[u].min_ord = 0xffffffffu;   // above every float
}

// grid = (chunks, batch): every CTA reduces a strided slice of one utterance and adds it with a handful of atomics.
__global__ void __launch_bounds__(256)
audio_stats_kernel(const float* wave, const long long* offsets, const long long* lengths, long long row_stride,
                   float clip_level, AudioStatsAcc* acc) {
  const int u = blockIdx.y;
  const long long n = lengths[u];
  const float* x = wave + (offsets ? offsets[u] : (long long)u * row_stride);
  double s_abs = 0.0, s_sq = 0.0;
  unsigned n_clip = 0;
  float mx = -INFINITY, mn = INFINITY;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i], a = fabsf(v);
    s_abs += a;
    s_sq += (double)v * v;
    n_clip += a >= clip_level;
    mx = fmaxf(mx, v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    s_abs += __shfl_xor_sync(0xffffffffu, s_abs, o);
    s_sq += __shfl_xor_sync(0xffffffffu, s_sq, o);
    n_clip += __shfl_xor_sync(0xffffffffu, n_clip, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0 && (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31) < n) {
    atomicAdd(&acc[u].sum_abs, s_abs);
    atomicAdd(&acc[u].sum_sq, s_sq);
    atomicAdd(&acc[u].n_clip, (unsigned long long)n_clip);
    atomicMax(&acc[u].max_ord, float_to_ordered(mx));
    atomicMin(&acc[u].min_ord, float_to_ordered(mn));
  }
}

// out[u] = {max, min, mean |x|, rms, clipping ratio, max |x|}; an empty utterance yields zeros.
__global__ void audio_stats_final_kernel(const AudioStatsAcc* acc, const long long* lengths, int batch, double* out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= batch) return;
  const long long n = lengths[u];
  double* o = out + 6 * u;
  if (n <= 0) {
    for (int k = 0; k < 6; ++k) o[k] = 0.0;
    return;
  }
  const double mx = ordered_to_float(acc[u].max_ord), mn = ordered_to_float(acc[u].min_ord);
  o[0] = mx;
  o[1] = mn;
  o[2] = acc[u].sum_abs / (double)n;
  o[3] = sqrt(acc[u].sum_sq / (double)n);
  o[4] = (double)acc[u].n_clip / (double)n;
  o[5] = fmax(fabs(mx), fabs(mn));
}

// Wire PCM -> float32 mono at the target rate, in the reference's float64 arithmetic and operation order so that the
// float32 result is bit-identical to numpy's (R:voice_interface.py:1004-1034):
//   width 1: ((u8 - 128) mod 256) / 128.0   (uint8 arithmetic wraps in numpy, :1007)      width 2: s16 / 32768.0
//   width 4: s32 / 2147483648.0             channels: np.mean(x.reshape(-1, ch), axis=1) = sequential sum / ch
//   resample: np.interp(np.linspace(0, n-1, m), np.arange(n), x): x_i = i * step, last = n - 1 exactly,
//             y = (x[j+1] - x[j]) * (x_i - j) + x[j] with separate multiply and add (no contraction).
__device__ __forceinline__ double ingest_mono(const void* pcm, int width, int channels, long long k) {
  double acc = 0.0;
  for (int c = 0; c < channels; ++c) {
    const long long i = k * channels + c;
    double v;
    if (width == 1) v = (double)(unsigned char)(static_cast<const unsigned char*>(pcm)[i] - 128u) / 128.0;
    else if (width == 2) v = (double)static_cast<const short*>(pcm)[i] / 32768.0;
    else v = (double)static_cast<const int*>(pcm)[i] / 2147483648.0;
    acc = c == 0 ? v : __dadd_rn(acc, v);
  }
  return channels > 1 ? acc / (double)channels : acc;
}

__global__ void ingest_pcm_kernel(const void* pcm, int width, int channels, long long n_in, long long n_out, double step,
                                  float* out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += (long long)gridDim.x * blockDim.x) {
    double y;
    if (n_out == n_in && step == 1.0) {
      y = ingest_mono(pcm, width, channels, i);
    } else {
      // np.linspace: the last point is `stop` exactly - unless it is the only point, which is `start`
      const double x = (i == n_out - 1 && n_out > 1) ? (double)(n_in - 1) : __dmul_rn((double)i, step);
      long long j = (long long)x;                 // x >= 0: truncation = floor
      if (j >= n_in - 1) {
        y = ingest_mono(pcm, width, channels, n_in - 1);
      } else {
        const double f0 = ingest_mono(pcm, width, channels, j), f1 = ingest_mono(pcm, width, channels, j + 1);
        y = (double)j == x ? f0 : __dadd_rn(__dmul_rn(__dadd_rn(f1, -f0), __dadd_rn(x, -(double)j)), f0);
      }
    }
    out[i] = (float)y;
  }
}

// Sliding audio windows, one circular buffer of `cap` samples per stream in one slab: [n_streams] int64 totals (samples
// ever pushed) followed by [n_streams][cap] float32.  Sample number t of a stream lives at slot t % cap, so a push never
// moves old data and the window of the newest min(total, cap) samples is a gather.
struct RingLayout {
  int n_streams;
  int cap;
  __host__ __device__ size_t totals_bytes() const { return ((size_t)n_streams * sizeof(long long) + 255) & ~(size_t)255; }
  __host__ __device__ size_t total_bytes() const { return totals_bytes() + (size_t)n_streams * cap * sizeof(float); }
  __host__ __device__ long long* totals(void* base) const { return reinterpret_cast<long long*>(base); }
  __host__ __device__ float* data(void* base) const { return reinterpret_cast<float*>((char*)base + totals_bytes()); }
};

__global__ void ring_reset_kernel(void* state, RingLayout lay, const int* ids, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = ids ? ids[i] : i;
  if (s >= 0 && s < lay.n_streams) lay.totals(state)[s] = 0;
}

// grid = (chunks of the copy, n): appends chunk b to stream ids[b]; only the newest `cap` samples of a long chunk land.
// Stream ids must be distinct within one call.  The total is advanced by a second tiny kernel (ring_commit_kernel) so
// that every block of this one sees the same starting total.
__global__ void ring_push_kernel(void* state, RingLayout lay, const float* chunks, long long chunk_stride,
                                 const int* chunk_lens, const int* ids, int max_chunk) {
  const int b = blockIdx.y, s = ids[b];
  if (s < 0 || s >= lay.n_streams) return;
  const int len = min(max(chunk_lens[b], 0), max_chunk);
  const long long total = lay.totals(state)[s];
  const int skip = len > lay.cap ? len - lay.cap : 0;       // older part of an over-long chunk never becomes visible
  float* ring = lay.data(state) + (size_t)s * lay.cap;
  const float* src = chunks + (long long)b * chunk_stride;
  for (int i = skip + blockIdx.x * blockDim.x + threadIdx.x; i < len; i += gridDim.x * blockDim.x)
    ring[(total + i) % lay.cap] = src[i];
}
__global__ void ring_commit_kernel(void* state, RingLayout lay, const int* chunk_lens, const int* ids, int n, int max_chunk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  const int s = ids[b];
  if (s >= 0 && s < lay.n_streams) lay.totals(state)[s] += min(max(chunk_lens[b], 0), max_chunk);
}

// out[b, 0 : m] = the newest m = min(total, cap) samples of stream ids[b], oldest first; out[b, m : cap] = 0; lens[b] = m.
__global__ void ring_window_kernel(const void* state, RingLayout lay, const int* ids, float* out, long long* lens) {
  const int b = blockIdx.y, s = ids[b];
  float* dst = out + (size_t)b * lay.cap;
  if (s < 0 || s >= lay.n_streams) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lay.cap; i += gridDim.x * blockDim.x) dst[i] = 0.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) lens[b] = 0;
    return;
  }
  const long long total = lay.totals(const_cast<void*>(state))[s];
  const int m = (int)(total < lay.cap ? total : lay.cap);
  const long long first = total - m;
  const float* ring = lay.data(const_cast<void*>(state)) + (size_t)s * lay.cap;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < lay.cap; i += gridDim.x * blockDim.x)
    dst[i] = i < m ? ring[(first + i) % lay.cap] : 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) lens[b] = m;
}

// feats[u, t, d] -= mean_t feats[u, :T_u, d]  (TA:642-644).  One CTA per (utterance, 32 columns): 8 row-lanes x 32
// column-lanes, float64 column sums (the reference sums in float32; the difference is below its own rounding).
__global__ void __launch_bounds__(256)
column_mean_kernel(float* feats, long long rows_cap, int D, const long long* n_rows) {
  __shared__ double part[8][33];
  const int u = blockIdx.y;
  const int d = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = threadIdx.x >> 5;
  const long long T = n_rows[u];
  float* base = feats + (long long)u * rows_cap * D;
  double s = 0.0;
  if (d < D)
    for (long long t = r0; t < T; t += 8) s += base[t * D + d];
  part[r0][threadIdx.x & 31] = s;
  __syncthreads();
  if (d >= D || T <= 0) return;
  double tot = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += part[k][threadIdx.x & 31];
  const float mean = (float)(tot / (double)T);
  for (long long t = r0; t < T; t += 8) base[t * D + d] -= mean;
}

}  // namespace b200fe
