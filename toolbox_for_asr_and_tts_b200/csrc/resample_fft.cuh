// scipy.signal.resample (the Fourier method) on the GPU: the resampler the reference runs on every chunk whose rate is
// not 16 kHz when scipy is installed (R:voice-service/app/services/voice_interface.py:1022-1027; its np.interp fallback
// is ingest_pcm_kernel in extras.cuh).  Semantics restated from scipy 1.18 (scipy/signal/_signaltools.py::resample,
// real input, window=None):
//     X = rfft(x);  m = min(num, n);  keep bins 0 .. m/2;  if m is even: bin m/2 *= 2 (num < n) or 0.5 (num > n)
//     y = irfft(X / (n / num), n=num)          -> float64, cast to float32 by the caller (:1045)
// irfft uses bin 0 and (num even) bin num/2 with weight 1 and their real parts only, every other bin with weight 2.
//
// The lengths are arbitrary (a 240-400 ms chunk at 44.1 / 48 kHz: n = 10 584 .. 19 200, num = 3 840 .. 6 400), so this
// is not an FFT but the two DFT sums themselves, in float64, O(n * m/2 + num * m/2) ~ 10^8 multiply-adds per chunk -
// microseconds on a B200.  Phases: every thread seeds its twiddle with an exact, integer-reduced sincospi and advances
// it by complex rotation, re-seeding every 64 steps, so the float64 result differs from pocketfft's by ~1e-15 relative
// and the float32 cast agrees with scipy's except where float64 noise straddles a float32 rounding boundary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200fe {

constexpr int kResampleReseed = 64;

// exp(-2*pi*i*(a mod n)/n) with the phase reduced on integers
__device__ __forceinline__ void unit_phase(long long a, long long n, double sign, double& c, double& s) {
  const long long r = a % n;
  sincospi(2.0 * (double)r / (double)n, &s, &c);
  s *= sign;
}

// mono[n] = width normalisation + channel mean of wire PCM (ingest_mono, extras.cuh), float64
__global__ void resample_mono_kernel(const void* pcm, int width, int channels, long long n_in, double* mono) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_in; i += (long long)gridDim.x * blockDim.x)
    mono[i] = ingest_mono(pcm, width, channels, i);
}

// X[k] = sum_n x[n] exp(-2 pi i k n / n_in), k = blockIdx.x < n_bins; one block per bin, threads own contiguous n ranges
__global__ void __launch_bounds__(256)
resample_dft_kernel(const double* __restrict__ x, long long n_in, int n_bins, double2* __restrict__ X) {
  __shared__ double red_r[8], red_i[8];
  const int k = blockIdx.x;
  const long long per = (n_in + blockDim.x - 1) / blockDim.x;
  const long long n0 = (long long)threadIdx.x * per, n1 = n0 + per < n_in ? n0 + per : n_in;
  double ar = 0.0, ai = 0.0;
  double wr, wi;
  unit_phase(k, n_in, -1.0, wr, wi);                    // w^k
  double cr = 1.0, ci = 0.0;
  for (long long n = n0; n < n1; ++n) {
    if (((n - n0) & (kResampleReseed - 1)) == 0) unit_phase((long long)k * n, n_in, -1.0, cr, ci);
    const double v = x[n];
    ar = fma(v, cr, ar);
    ai = fma(v, ci, ai);
    const double tr = fma(cr, wr, -ci * wi), ti = fma(cr, wi, ci * wr);
    cr = tr;
    ci = ti;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ai += __shfl_xor_sync(0xffffffffu, ai, o);
  }
  if ((threadIdx.x & 31) == 0) { red_r[threadIdx.x >> 5] = ar; red_i[threadIdx.x >> 5] = ai; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sr = 0.0, si = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sr += red_r[w]; si += red_i[w]; }
    X[k] = make_double2(sr, si);
  }
}

// y[j] = irfft(X * num / n_in, n = num)[j], one thread per output sample; cast to float32 like the reference (:1045)
__global__ void __launch_bounds__(128)
resample_idft_kernel(const double2* __restrict__ X, long long n_in, long long num, float* __restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num) return;
  const long long m = num < n_in ? num : n_in;
  const long long last = m / 2;                          // highest kept bin
  const bool m_even = (m & 1) == 0;
  const bool lone = m_even && num < n_in;                // bin m/2 is the output's own Nyquist bin: weight 1, real part
  const double scale = (double)num / (double)n_in;       // 1 / s_fac
  double acc = X[0].x * scale;
  double wr, wi;
  unit_phase(j, num, 1.0, wr, wi);                       // e^{+2 pi i j / num}
  double cr = 1.0, ci = 0.0;
  const long long k_hi = lone ? last - 1 : last;
  for (long long k = 1; k <= k_hi; ++k) {
    if (((k - 1) & (kResampleReseed - 1)) == 0) unit_phase(j * k, num, 1.0, cr, ci);
    double2 v = X[k];
    double wgt = 2.0 * scale;
    if (m_even && k == last) wgt *= 0.5;                 // up-sampling: the input's Nyquist bin is split (x 0.5)
    acc = fma(wgt, fma(v.x, cr, -v.y * ci), acc);        // Re(X_k e^{i theta})
    const double tr = fma(cr, wr, -ci * wi), ti = fma(cr, wi, ci * wr);
    cr = tr;
    ci = ti;
  }
  if (lone) acc = fma(2.0 * scale * X[last].x, (j & 1) ? -1.0 : 1.0, acc);   // (X[m/2] *= 2), weight 1, (-1)^j
  out[j] = (float)(acc / (double)num);
}

}  // namespace b200fe
