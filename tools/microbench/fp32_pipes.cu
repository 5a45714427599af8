// Microbenchmark (B200, sm_100a): issue rate of scalar vs packed FP32 instructions, alone and mixed with shared-memory
// loads.  Evidence for the instruction-selection choices in DESIGN.md (is fma.rn.f32x2 worth restructuring the FFT?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/fp32_pipes tools/microbench/fp32_pipes.cu
// Prints, per kernel, warp-instructions issued per cycle per SM (4 = one per scheduler per cycle) and scalar-equivalent
// FP32 operations per cycle per SM.
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 16
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) { float2 v = make_float2(a, b); return *reinterpret_cast<u64*>(&v); }

enum Kind { FFMA, FADD, FMUL, FFMA2, FADD2, FMUL2, MIX_FFMA_FADD, MIX_FFMA2_FADD2, FFMA_LDS, FFMA2_LDS, FFMA_IADD, FFMA_LDS128 };

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* cyc) {
  __shared__ float sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i * 1e-3f;
  __syncthreads();
  float a[CHAINS];
  u64 p[CHAINS / 2];
  int ia[4] = {1, 2, 3, 4};
  for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < CHAINS / 2; ++i) p[i] = pack(threadIdx.x * 0.001f + i, i * 0.5f);
  float b = 1.0001f + out[0] * 0.f, c = 0.5f + out[1] * 0.f;
  u64 b2 = pack(b, b + 1e-4f), c2 = pack(c, c * 0.5f);
  const float* sp = sm + threadIdx.x;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (KIND == FFMA) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
    } else if constexpr (KIND == FADD) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
    } else if constexpr (KIND == FMUL) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
    } else if constexpr (KIND == FFMA2) {
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
    } else if constexpr (KIND == FADD2) {
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
    } else if constexpr (KIND == FMUL2) {
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(b2));
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(b2));
    } else if constexpr (KIND == MIX_FFMA_FADD) {
#pragma unroll
      for (int i = 0; i < CHAINS; i += 2) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(c));
      }
    } else if constexpr (KIND == MIX_FFMA2_FADD2) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < CHAINS / 2; i += 2) {
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(b2), "l"(c2));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i + 1]) : "l"(c2));
        }
    } else if constexpr (KIND == FFMA_LDS) {   // 12 FFMA + 4 LDS.32 (conflict-free)
#pragma unroll
      for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += sp[((it + i) & 7) * 256];
    } else if constexpr (KIND == FFMA2_LDS) {  // 12 FFMA2 + 4 LDS.32
#pragma unroll
      for (int r = 0; r < 12; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[r & 7]) : "l"(b2), "l"(c2));
#pragma unroll
      for (int i = 0; i < 4; ++i) acc += sp[((it + i) & 7) * 256];
    } else if constexpr (KIND == FFMA_IADD) {  // 12 FFMA + 4 integer ALU ops
#pragma unroll
      for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(ia[i]) : "r"(it));
    } else if constexpr (KIND == FFMA_LDS128) {  // 12 FFMA + 1 LDS.128 (4 wavefronts)
#pragma unroll
      for (int i = 0; i < 12; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
      const float4 v = *reinterpret_cast<const float4*>(sm + 4 * ((threadIdx.x + it) & 255));
      acc += v.x + v.y + v.z + v.w;
    }
  }
  long long t1 = clock64();
  float s = acc + ia[0] + ia[1] + ia[2] + ia[3];
  for (int i = 0; i < CHAINS; ++i) s += a[i];
  for (int i = 0; i < CHAINS / 2; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
void run(const char* name, float* out, long long* cyc, int warps_per_sm, double instr_per_iter, double flop_per_iter) {
  const int iters = 4000;
  const int threads = 256, blocks = 148 * (warps_per_sm * 32 / threads);
  k<KIND><<<blocks, threads>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<KIND><<<blocks, threads>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  static long long h[148 * 8];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < blocks; ++i) mx = h[i] > mx ? h[i] : mx;
  const double wi = (double)iters * instr_per_iter * warps_per_sm;   // warp-instructions per SM
  printf("%-18s warps/SM=%2d  %.3f ms  cycles=%.0f  warp-instr/clk/SM=%.2f  fp32-lane-ops/clk/SM=%.1f\n", name, warps_per_sm, ms, mx,
         wi / mx, (double)iters * flop_per_iter * warps_per_sm * 32 / mx);
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  cudaMemset(out, 0, 148 * 8 * 256 * sizeof(float));
  cudaMalloc(&cyc, 148 * 8 * sizeof(long long));
  for (int w : {8, 16, 32}) {
    run<FFMA>("FFMA", out, cyc, w, 16, 16);
    run<FADD>("FADD", out, cyc, w, 16, 16);
    run<FMUL>("FMUL", out, cyc, w, 16, 16);
    run<FFMA2>("FFMA2", out, cyc, w, 16, 32);
    run<FADD2>("FADD2", out, cyc, w, 16, 32);
    run<FMUL2>("FMUL2", out, cyc, w, 16, 32);
    run<MIX_FFMA_FADD>("FFMA+FADD", out, cyc, w, 16, 16);
    run<MIX_FFMA2_FADD2>("FFMA2+FADD2", out, cyc, w, 16, 32);
    run<FFMA_LDS>("12FFMA+4LDS", out, cyc, w, 20, 12);
    run<FFMA2_LDS>("12FFMA2+4LDS", out, cyc, w, 20, 24);
    run<FFMA_IADD>("12FFMA+4XOR", out, cyc, w, 16, 12);
    run<FFMA_LDS128>("12FFMA+LDS128", out, cyc, w, 17, 12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
