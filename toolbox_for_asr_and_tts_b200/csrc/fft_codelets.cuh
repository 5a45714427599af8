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

// ------------------------------------------------------------------------------------------------------------
// Packed (f32x2) codelets: every value is a float2 whose two lanes belong to two independent transforms (two
// audio frames).  sm_100a issues FFMA2 / FADD2 / FMUL2 at half the rate of the scalar forms but each does two lanes,
// so the FP32 pipe throughput is unchanged while the issue slots per flop halve (tools/microbench/fp32_pipes.cu).
// Constants are the same in both lanes, which ptxas encodes as a replicated immediate; a scalar register broadcasts
// for free (operand modifier .F32), and so do negation and the lane swap.
namespace b200fe {

using f2 = float2;
__device__ __forceinline__ f2 neg2(f2 a) { return make_float2(-a.x, -a.y); }   // folds into an operand modifier
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, neg2(b)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 mul2s(f2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 fma2s(f2 a, float s, f2 c) { return __ffma2_rn(a, make_float2(s, s), c); }   // a*s + c

// One DIT butterfly on packed values: (a, b) <- (a + w b, a - w b), w = exp(-2*pi*i*Q/(2*SPAN)).
// The twiddle product is taken in ratio form (6 FMAs instead of 4 + 4): with |wr| >= |wi| and r = wi/wr,
//   w b = wr * (br - r bi, bi + r br);  otherwise, with r = wr/wi,  w b = wi * (r br - bi, br + r bi).
// B_IS_ZERO prunes butterflies whose second input is a structural zero.
template <int SPAN, int Q, bool B_IS_ZERO>
__device__ __forceinline__ void dit_butterfly2(f2& ar, f2& ai, f2& br, f2& bi) {
  if constexpr (B_IS_ZERO) {
    static_assert(Q == 0, "zero pruning is only used in the first stage");
    br = ar;
    bi = ai;
  } else if constexpr (Q == 0) {
    const f2 xr = add2(ar, br), xi = add2(ai, bi);
    br = sub2(ar, br);
    bi = sub2(ai, bi);
    ar = xr;
    ai = xi;
  } else if constexpr (2 * Q == SPAN) {   // w = -i:  w b = (bi, -br)
    const f2 xr = add2(ar, bi), xi = sub2(ai, br);
    const f2 yr = sub2(ar, bi), yi = add2(ai, br);
    ar = xr; ai = xi; br = yr; bi = yi;
  } else {
    constexpr double cwr = ct_cos2pi(Q, 2 * SPAN), cwi = -ct_sin2pi(Q, 2 * SPAN);
    constexpr bool kBigR = (cwr < 0 ? -cwr : cwr) >= (cwi < 0 ? -cwi : cwi);
    f2 p, q;
    float m;
    // the ratio is taken against the ROUNDED multiplier, so that m*r carries one rounding error, like m itself
    if constexpr (kBigR) {
      constexpr float mm = (float)cwr;
      constexpr float r = (float)(cwi / (double)mm);
      m = mm;
      p = fma2s(bi, -r, br);
      q = fma2s(br, r, bi);
    } else {
      constexpr float mm = (float)cwi;
      constexpr float r = (float)(cwr / (double)mm);
      m = mm;
      p = fma2s(br, r, neg2(bi));
      q = fma2s(bi, r, br);
    }
    const f2 xr = fma2s(p, m, ar), xi = fma2s(q, m, ai);
    br = fma2s(p, -m, ar);
    bi = fma2s(q, -m, ai);
    ar = xr;
    ai = xi;
  }
}

template <int N, int SPAN, int LIVE>
__device__ __forceinline__ void dit_stage2(f2 (&re)[N], f2 (&im)[N]) {
  static_for<0, N / 2>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int q = i % SPAN;
    constexpr int a = (i / SPAN) * 2 * SPAN + q;
    constexpr int b = a + SPAN;
    // position b of the first stage holds input index bitrev(b) = bitrev(a) + N/2
    constexpr bool zero = SPAN == 1 && (bitrev<N>(b) >= LIVE);
    dit_butterfly2<SPAN, q, zero>(re[a], im[a], re[b], im[b]);
  });
  if constexpr (2 * SPAN < N) dit_stage2<N, 2 * SPAN, LIVE>(re, im);
}

// In-place forward DFT of N packed complex points.  On entry position p holds input bitrev<N>(p); on exit position k
// holds bin k.  Inputs with index >= LIVE are structural zeros (LIVE > N/2 or == N; their registers are not read).
template <int N, int LIVE = N>
__device__ __forceinline__ void fft_dit2(f2 (&re)[N], f2 (&im)[N]) {
  static_assert(LIVE == N || LIVE > N / 2, "LIVE must cover the first half");
  dit_stage2<N, 1, LIVE>(re, im);
}

// Real 32-point DFT of two packed real sequences from the 16-point complex DFT Z of z[m] = y[2m] + i y[2m+1]
// (zr/zi, bins in natural order).  Writes bins 1..15 in place, SCALED BY 2:  2 Y[k] = S + T,  2 Y[16-k] = conj(S - T),
// with A = Z[k], B = Z[16-k], S = A + conj B, T = W32^k (-i)(A - conj B).  Returns the real bins Y[0] and Y[16]
// (NOT scaled) in y0 / y16; bin 8 is conj Z[8] (not scaled).
__device__ __forceinline__ void real32_split2(f2 (&zr)[16], f2 (&zi)[16], f2& y0, f2& y16) {
  y0 = add2(zr[0], zi[0]);
  y16 = sub2(zr[0], zi[0]);
  zi[8] = neg2(zi[8]);
  static_for<1, 8>([&](auto ic) {
    constexpr int k = decltype(ic)::value;
    const f2 sr = add2(zr[k], zr[16 - k]), si = sub2(zi[k], zi[16 - k]);
    const f2 orr = add2(zi[k], zi[16 - k]), oi = sub2(zr[16 - k], zr[k]);
    constexpr double cwr = ct_cos2pi(k, 32), cwi = -ct_sin2pi(k, 32);
    constexpr bool kBigR = (cwr < 0 ? -cwr : cwr) >= (cwi < 0 ? -cwi : cwi);
    f2 p, q;
    float m;
    if constexpr (kBigR) {
      constexpr float mm = (float)cwr;
      constexpr float r = (float)(cwi / (double)mm);
      m = mm;
      p = fma2s(oi, -r, orr);
      q = fma2s(orr, r, oi);
    } else {
      constexpr float mm = (float)cwi;
      constexpr float r = (float)(cwr / (double)mm);
      m = mm;
      p = fma2s(orr, r, neg2(oi));
      q = fma2s(oi, r, orr);
    }
    zr[k] = fma2s(p, m, sr);
    zi[k] = fma2s(q, m, si);
    zr[16 - k] = fma2s(p, -m, sr);
    zi[16 - k] = fma2s(q, m, neg2(si));
  });
}

}  // namespace b200fe
