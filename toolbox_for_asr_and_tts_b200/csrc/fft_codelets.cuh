// Register-resident radix-2 DIF FFT codelets (N = 16, 32) for sm_100a.
//
// Everything is unrolled at compile time (template recursion, no runtime-indexed arrays), so the complex
// work arrays live in registers and every twiddle is an immediate operand.  Twiddles are evaluated at
// compile time in double precision and rounded once to float.
//
// Output order: after fft_dif<N>(re, im) the DFT bin k is stored at position bitrev<N>(k).
#pragma once
#include <type_traits>

namespace b200fe {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// ---- compile-time cos/sin(2*pi*q/n): range-reduced Taylor series in double
constexpr double kPi = 3.14159265358979323846264338327950288;

__host__ __device__ constexpr double ct_sin_taylor(double x) {  // |x| <= pi/4
  double x2 = x * x, term = x, sum = x;
  for (int k = 1; k < 14; ++k) {
    term *= -x2 / double((2 * k) * (2 * k + 1));
    sum += term;
  }
  return sum;
}
__host__ __device__ constexpr double ct_cos_taylor(double x) {  // |x| <= pi/4
  double x2 = x * x, term = 1.0, sum = 1.0;
  for (int k = 1; k < 14; ++k) {
    term *= -x2 / double((2 * k - 1) * (2 * k));
    sum += term;
  }
  return sum;
}
// cos(2*pi*q/n) and sin(2*pi*q/n) for 0 <= q < n, exact symmetries resolved on the integer phase
__host__ __device__ constexpr double ct_cos2pi(int q, int n) {
  q %= n;
  if (q < 0) q += n;
  if (8 * q <= n) return ct_cos_taylor(2.0 * kPi * q / n);
  if (8 * q <= 3 * n) return -ct_sin_taylor(2.0 * kPi * q / n - kPi / 2);   // around pi/2
  if (8 * q <= 5 * n) return -ct_cos_taylor(2.0 * kPi * q / n - kPi);       // around pi
  if (8 * q <= 7 * n) return ct_sin_taylor(2.0 * kPi * q / n - 1.5 * kPi);  // around 3pi/2
  return ct_cos_taylor(2.0 * kPi * q / n - 2.0 * kPi);
}
__host__ __device__ constexpr double ct_sin2pi(int q, int n) { return ct_cos2pi(4 * q - n, 4 * n); }  // sin(x) = cos(x - pi/2)

template <int N>
__host__ __device__ constexpr int bitrev(int k) {
  int r = 0;
  for (int b = 1; b < N; b <<= 1) {
    r = (r << 1) | (k & 1);
    k >>= 1;
  }
  return r;
}

// One DIF butterfly: (a, c) <- (a + c, (a - c) * W_{2*SPAN}^Q),  W = exp(-2*pi*i*Q/(2*SPAN)).
// C_IS_ZERO prunes butterflies whose second input is a structural zero (zero-padded frames).
template <int SPAN, int Q, bool C_IS_ZERO>
__device__ __forceinline__ void dif_butterfly(float& ar, float& ai, float& cr, float& ci) {
  float dr, di;
  if constexpr (C_IS_ZERO) {
    dr = ar;
    di = ai;
  } else {
    const float ur = ar + cr, ui = ai + ci;
    dr = ar - cr;
    di = ai - ci;
    ar = ur;
    ai = ui;
  }
  if constexpr (Q == 0) {
    cr = dr;
    ci = di;
  } else if constexpr (2 * Q == SPAN) {  // W = -i
    cr = di;
    ci = -dr;
  } else if constexpr (4 * Q == SPAN) {  // W = (1 - i)/sqrt2
    constexpr float c = 0.70710678118654752440f;
    cr = (dr + di) * c;
    ci = (di - dr) * c;
  } else if constexpr (4 * Q == 3 * SPAN) {  // W = (-1 - i)/sqrt2
    constexpr float c = 0.70710678118654752440f;
    cr = (di - dr) * c;
    ci = -(dr + di) * c;
  } else {
    constexpr float wr = (float)ct_cos2pi(Q, 2 * SPAN);
    constexpr float wi = (float)(-ct_sin2pi(Q, 2 * SPAN));
    cr = fmaf(dr, wr, -(di * wi));
    ci = fmaf(dr, wi, di * wr);
  }
}

template <int N, int SPAN, int LIVE>
__device__ __forceinline__ void dif_stage(float (&re)[N], float (&im)[N]) {
  static_for<0, N / 2>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int q = i % SPAN;
    constexpr int a = (i / SPAN) * 2 * SPAN + q;
    constexpr int c = a + SPAN;
    dif_butterfly<SPAN, q, (c >= LIVE)>(re[a], im[a], re[c], im[c]);
  });
  if constexpr (SPAN > 1) dif_stage<N, SPAN / 2, N>(re, im);
}

// In-place forward DFT of N complex points held in registers.  Inputs at positions >= LIVE are structural zeros
// (only the first stage can exploit that; LIVE must be > N/2 or == N).
template <int N, int LIVE = N>
__device__ __forceinline__ void fft_dif(float (&re)[N], float (&im)[N]) {
  static_assert(LIVE == N || LIVE > N / 2, "LIVE must cover the first half");
  dif_stage<N, N / 2, LIVE>(re, im);
}

}  // namespace b200fe
