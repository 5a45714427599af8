// 512-point COMPLEX FFT core (16 threads, one transpose) used by the TTS log-mel kernel, whose 1024-point real FFT is
// a 512-point complex FFT of z[m] = x[2m] + i x[2m+1] plus an even/odd recombination.  (The ASR front-end uses the
// packed real FFT of fbank_tile.cuh instead.)
#pragma once
#include <cuda_runtime.h>

#include "fft_codelets.cuh"

namespace b200fe {

constexpr int kXGroupFloat2 = 32 * 18;   // transpose buffer of one group: 32 rows of 16 + 2 (pad) float2
constexpr int kTwRows = 17;              // twiddle rows 0..16 (row 16 serves thread 0's second column)
constexpr int kTwTable = kTwRows * 18;

// 512-point complex FFT by a group of 16 threads.  In: re/im[i] = z[16 i + j] (thread j, rows >= LIVE are zero).
// A 32-point register FFT per thread, ONE shared-memory transpose (rows padded to 18 float2), the W512 twiddles, then
// two 16-point register FFTs on columns cA = j and cB = 32 - rowB (rowB = j, or 16 for thread 0, whose columns are 0
// and 16).  Out, bit-reversed positions: Z[cA + 32 k2] = (ar, ai)[bitrev16(k2)], Z[cB + 32 k2] = (br, bi)[bitrev16((k2+1)&15)]
// (column B is multiplied by the conjugated twiddle row, which rotates its spectrum by one bin).
// After the call every lane of the warp has passed a __syncwarp: the transpose buffer may be reused.
// Finally thread 0's registers are permuted so that for EVERY thread "slot i" = (A[i], B[15-i]) holds the conjugate
// pair (Z[k], Z[512-k]) with k = cA + 32 i for i < 8 (thread 0: i in 1..7 pair inside column 0, i in 8..15 inside
// column 16; its slot 0 is (Z[0], Z[256])).
template <int LIVE>
__device__ __forceinline__ void fft512_columns(float (&re)[32], float (&im)[32], float2* xg, const float2* tw_g, int j,
                                               float (&ar)[16], float (&ai)[16], float (&br)[16], float (&bi)[16]) {
  fft_dif<32, LIVE>(re, im);
  __syncwarp();   // earlier readers of the (aliased) buffer are done
  static_for<0, 32>([&](auto ic) {
    constexpr int k1 = decltype(ic)::value;
    constexpr int pos = bitrev<32>(k1);
    xg[k1 * 18 + j] = make_float2(re[pos], im[pos]);
  });
  __syncwarp();
  const int rowB = j == 0 ? 16 : j;
  const bool t0 = (j == 0);
  {
    const float4* rowA4 = reinterpret_cast<const float4*>(xg + j * 18);
    const float4* rowB4 = reinterpret_cast<const float4*>(xg + (32 - rowB) * 18);
    const float4* twA4 = reinterpret_cast<const float4*>(tw_g + j * 18);
    const float4* tw16 = reinterpret_cast<const float4*>(tw_g + 16 * 18);   // same address for the whole group
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float4 ya = rowA4[h], yb = rowB4[h], ta = twA4[h], t16 = tw16[h];
      // columns j and 32-j share one twiddle row (conjugated); only thread 0 (columns 0 and 16) needs row 16
      float4 tb;
      tb.x = t0 ? t16.x : ta.x; tb.y = t0 ? t16.y : ta.y; tb.z = t0 ? t16.z : ta.z; tb.w = t0 ? t16.w : ta.w;
      ar[2 * h] = fmaf(ya.x, ta.x, -(ya.y * ta.y));
      ai[2 * h] = fmaf(ya.x, ta.y, ya.y * ta.x);
      ar[2 * h + 1] = fmaf(ya.z, ta.z, -(ya.w * ta.w));
      ai[2 * h + 1] = fmaf(ya.z, ta.w, ya.w * ta.z);
      br[2 * h] = fmaf(yb.x, tb.x, yb.y * tb.y);
      bi[2 * h] = fmaf(yb.y, tb.x, -(yb.x * tb.y));
      br[2 * h + 1] = fmaf(yb.z, tb.z, yb.w * tb.w);
      bi[2 * h + 1] = fmaf(yb.w, tb.z, -(yb.z * tb.w));
    }
  }
  fft_dif<16>(ar, ai);
  fft_dif<16>(br, bi);
  __syncwarp();   // every lane has consumed the transpose buffer
#define A_RE(k) ar[bitrev<16>(k)]
#define A_IM(k) ai[bitrev<16>(k)]
#define B_RE(k) br[bitrev<16>(((k) + 1) & 15)]
#define B_IM(k) bi[bitrev<16>(((k) + 1) & 15)]
  const float z256r = A_RE(8), z256i = A_IM(8);   // thread 0: Z[256], which pairs with itself
  static_for<8, 15>([&](auto ic) {
    constexpr int q = decltype(ic)::value;
    const float tr = B_RE(q), ti = B_IM(q);
    B_RE(q) = t0 ? A_RE(q + 1) : tr;
    B_IM(q) = t0 ? A_IM(q + 1) : ti;
    A_RE(q) = t0 ? tr : A_RE(q);
    A_IM(q) = t0 ? ti : A_IM(q);
  });
  {
    const float tr = B_RE(15), ti = B_IM(15);
    B_RE(15) = t0 ? z256r : tr;        // thread 0, slot 0: (Z[0], Z[256])
    B_IM(15) = t0 ? z256i : ti;
    A_RE(15) = t0 ? tr : A_RE(15);
    A_IM(15) = t0 ? ti : A_IM(15);
  }
#undef A_RE
#undef A_IM
#undef B_RE
#undef B_IM
}

}  // namespace b200fe
