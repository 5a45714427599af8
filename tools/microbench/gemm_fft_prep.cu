// Microbenchmark (B200, sm_100a): the CUDA-core work a tcgen05 GEMM-FFT front-end would still have to do AROUND its MMAs,
// measured on the frame count of BASELINE configs[1] (412 k frames).  Evidence for DESIGN.md section 4.4: is there room
// for a tensor-core design below the shipped kernel's 0.241 ms?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/gemm_fft_prep tools/microbench/gemm_fft_prep.cu
//
// A two-level GEMM-FFT (n = 16 i + j; stage A: 32-point real DFT over i, stage B: 16-point complex DFT over j) with the
// only operand format that meets the stated tolerance (three-way bf16 split, profiles/r2_gemm_fft_accuracy.txt) needs,
// per frame, on the CUDA cores:
//   (a) 400 samples: load x[n] and x[n-1], pre-emphasis, window (2 FMAs, as the shipped kernel), split into bf16
//       hi / mid / lo, store 3 x 400 x 2 B as the stage-A operand;
//   (b) the 17 x 32 non-redundant complex outputs of stage A (544 complex = 1 088 floats): twiddle multiply, split re and
//       im into bf16 hi / mid / lo, store 3 x 1 088 x 2 B as the stage-B operand;
//   (c) power, mel, log, LFR, CMVN, output - as today.
// This kernel runs (a) and (b) only, and generously: the stage-A outputs "come from TMEM" for free (made up from a
// register seed, no tcgen05.ld), the twiddle is one complex constant per lane (no table loads), there is no mbarrier or
// MMA-issue logic, operand tiles are written with 128-bit stores to a small recycled shared-memory buffer (no layout
// arithmetic beyond a linear index), and (c) is not run at all.  Work distribution and sample fetch are the shipped kernel's (a warp per
// quad of 4 frames, persistent CTAs, 880 samples per quad through shared memory).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kWarps = 4, kThreads = 128, kL = 400, kS = 160;   // a quad: 3 * 160 + 400 = 880 samples

__device__ __forceinline__ unsigned pack_bf16x2(float a, float b) {
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float lo_as_float(unsigned p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float hi_as_float(unsigned p) { return __uint_as_float(p & 0xffff0000u); }

// three-way bf16 split of two floats: hi, mid, lo packed pairwise (what 6 MMAs per product consume).
// Default: round-to-nearest conversions (cvt.rn.bf16x2.f32, SASS F2FP).  -DTRUNC_SPLIT: truncation - the upper halves of the
// two floats packed by one byte permute (PRMT), remainders by AND + subtract: exact (3 x 8 mantissa bits), no conversion
// instructions, the same 11 instructions per pair.
#ifdef TRUNC_SPLIT
__device__ __forceinline__ unsigned pack_hi(float a, float b) { return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632); }
__device__ __forceinline__ float rem(float a) { return a - __uint_as_float(__float_as_uint(a) & 0xffff0000u); }
__device__ __forceinline__ void split3(float a, float b, unsigned& h, unsigned& m, unsigned& l) {
  h = pack_hi(a, b);
  const float ra = rem(a), rb = rem(b);
  m = pack_hi(ra, rb);
  l = pack_hi(rem(ra), rem(rb));
}
#else
__device__ __forceinline__ void split3(float a, float b, unsigned& h, unsigned& m, unsigned& l) {
  h = pack_bf16x2(a, b);
  const float ra = a - lo_as_float(h), rb = b - hi_as_float(h);
  m = pack_bf16x2(ra, rb);
  const float sa = ra - lo_as_float(m), sb = rb - hi_as_float(m);
  l = pack_bf16x2(sa, sb);
}
#endif

__global__ void __launch_bounds__(kThreads, 4)
prep_kernel(const float* __restrict__ wave, long long n_quads, const float* __restrict__ window, int do_a, int do_b,
            unsigned* __restrict__ sink) {
  __shared__ __align__(16) float samples[kWarps][896];
  __shared__ __align__(16) unsigned tile[kWarps][3][512];   // recycled operand buffer: 3 planes x 2 KB per warp
  __shared__ float win_s[kL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kL; i += kThreads) win_s[i] = window[i];
  __syncthreads();
  const unsigned th = (unsigned)__cvta_generic_to_shared(tile[warp][0]), tm = (unsigned)__cvta_generic_to_shared(tile[warp][1]),
                 tl = (unsigned)__cvta_generic_to_shared(tile[warp][2]);
  auto sts128 = [](unsigned addr, const unsigned (&v)[4]) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
  };
  auto store3 = [&](int o, const unsigned (&h)[4], const unsigned (&m)[4], const unsigned (&l)[4]) {   // three 128-bit stores
    sts128(th + 16 * o, h);
    sts128(tm + 16 * o, m);
    sts128(tl + 16 * o, l);
  };
  const float twr = 0.92387953f + 1e-3f * lane, twi = -0.38268343f + 1e-3f * lane;   // "the lane's twiddle"
  unsigned acc = 0;
  const long long W = (long long)gridDim.x * kWarps;
  for (long long q = (long long)blockIdx.x * kWarps + warp; q < n_quads; q += W) {
    // ---- the quad's 880 samples through shared memory (128-bit loads and stores, as round 1 of the shipped kernel)
    const float4* src = reinterpret_cast<const float4*>(wave + q * (4 * kS));
    float4* dst = reinterpret_cast<float4*>(samples[warp]);
#pragma unroll
    for (int u = 0; u < 7; ++u)
      if (lane + 32 * u < 220) dst[lane + 32 * u] = __ldg(src + lane + 32 * u);
    __syncwarp();
    if (do_a) {
      // ---- (a) 4 frames x 400 samples = 200 chunks of 8 consecutive samples, 6.25 per lane: 128-bit loads of the samples
      //      and of the window, 16 FMAs, four pairwise splits, three 128-bit stores per chunk
      const float* x = samples[warp];
#pragma unroll 2
      for (int c = lane; c < 200; c += 32) {
        const int f = c / 50, n = 8 * (c - 50 * f);
        const float* xf = x + f * kS + n;
        const float4 xa = *reinterpret_cast<const float4*>(xf), xb = *reinterpret_cast<const float4*>(xf + 4);
        const float4 wa = *reinterpret_cast<const float4*>(win_s + n), wb = *reinterpret_cast<const float4*>(win_s + n + 4);
        const float xm = n ? xf[-1] : xa.x;
        const float xv[9] = {xm, xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        unsigned h[4], m[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float y0 = fmaf(-0.97f, xv[2 * k], xv[2 * k + 1]) * wv[2 * k];          // (the DC term: one more FMA each)
          const float y1 = fmaf(-0.97f, xv[2 * k + 1], xv[2 * k + 2]) * wv[2 * k + 1];
          split3(y0, y1, h[k], m[k], l[k]);
        }
        store3(c & 127, h, m, l);
      }
    }
    if (do_b) {
      // ---- (b) 4 frames x 544 complex stage-A outputs = 544 chunks of 4 complex, 17 per lane; values made up from a
      //      register seed (stands in for the tcgen05.ld, free in this model): twiddle, splits, three 128-bit stores
      float sr = samples[warp][lane] + 1.0f, si = samples[warp][lane + 32] - 1.0f;
#pragma unroll 2
      for (int c = lane; c < 544; c += 32) {
        unsigned h[4], m[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          sr = fmaf(sr, 1.0001f, 0.001f);
          si = fmaf(si, 0.9999f, -0.001f);
          const float ar = fmaf(sr, twr, -(si * twi)), ai = fmaf(sr, twi, si * twr);
          split3(ar, ai, h[k], m[k], l[k]);
        }
        store3(c & 127, h, m, l);
      }
    }
    __syncwarp();
    acc += tile[warp][0][lane] ^ tile[warp][2][lane + 64];
  }
  if (acc == 0x12345678u) sink[0] = acc;   // keeps the work alive
}

int main(int argc, char** argv) {
  const long long n_frames = 411900, n_quads = (n_frames + 3) / 4;
  const long long n_samples = n_quads * 4 * 160 + 1024;
  std::vector<float> h(n_samples), win(400);
  for (long long i = 0; i < n_samples; ++i) h[i] = (float)((i * 2654435761u) & 0xffff) / 65536.0f * 0.6f - 0.3f;
  for (int i = 0; i < 400; ++i) win[i] = 0.54f - 0.46f * cosf(6.283185307f * i / 399.0f);
  float *d_wave, *d_win;
  unsigned* d_sink;
  cudaMalloc(&d_wave, n_samples * 4);
  cudaMalloc(&d_win, 400 * 4);
  cudaMalloc(&d_sink, 4);
  cudaMemcpy(d_wave, h.data(), n_samples * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_win, win.data(), 400 * 4, cudaMemcpyHostToDevice);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[3] = {"(a) stage-A operand: pre-emphasis, window, 3-way bf16 split, stores",
                          "(b) stage-B operand: twiddle, 3-way bf16 split of re and im, stores", "(a) + (b)"};
  for (int mode = 0; mode < 3; ++mode) {
    const int a = mode != 1, b = mode != 0;
    for (int i = 0; i < 3; ++i) prep_kernel<<<grid, kThreads>>>(d_wave, n_quads, d_win, a, b, d_sink);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) prep_kernel<<<grid, kThreads>>>(d_wave, n_quads, d_win, a, b, d_sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-78s %.4f ms per %lld frames (%d CTAs x %d threads)\n", names[mode], ms / reps, n_frames, grid, kThreads);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
#ifdef TRUNC_SPLIT
  printf("(split by truncation: PRMT / AND / subtract, no conversion instructions)\n");
#else
  printf("(split by round-to-nearest conversions: cvt.rn.bf16x2.f32)\n");
#endif
  printf("shipped fused kernel, everything included: 0.241 ms for the same frames (profiles/r2_bench_n1_steps200.json)\n");
  return 0;
}
