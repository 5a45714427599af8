// Microbenchmark (B200, sm_100a): how many data-pipe cycles a shared-memory load costs when lanes repeat addresses.
// Evidence for DESIGN.md section 7 (tables that both 16-thread groups of a warp read: do the duplicates merge?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/smem_wavefronts tools/microbench/smem_wavefronts.cu
// One CTA of 128 threads per SM issues UNROLL independent loads per iteration; cycles per warp-load on one SM
// (4 warps share the pipe) approximate the wavefronts of that access pattern.
#include <cstdio>
#include <cuda_runtime.h>

enum Pattern { DISTINCT, HALF_DUP, PAIR_DUP, ALL_SAME, HALF_DUP_PITCH17 };
static const char* kNames[] = {"32 lanes distinct, contiguous", "lanes 16-31 repeat lanes 0-15", "lanes 2j, 2j+1 share an address",
                               "all lanes one address", "lanes 16-31 repeat lanes 0-15, 16 rows of pitch 17/18"};

template <int WIDTH, int PATTERN>
__global__ void __launch_bounds__(128) k(float* out, int iters, long long* cyc) {
  __shared__ __align__(16) float sm[8192];   // 4 warps x 4 KB + the rotating 0..3 KB offset stays inside 32 KB
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx;   // in units of the access width
  if (PATTERN == DISTINCT) idx = lane;
  else if (PATTERN == HALF_DUP) idx = lane & 15;
  else if (PATTERN == PAIR_DUP) idx = lane >> 1;
  else if (PATTERN == ALL_SAME) idx = 0;
  else idx = (lane & 15) * (WIDTH == 4 ? 9 : 17);   // 128-bit: rows of 18 float2 = 9 float4; 64-bit: rows of 17 float2
  const char* base = reinterpret_cast<const char*>(sm) + (size_t)idx * WIDTH * 4 + (threadIdx.x >> 5) * 4096;   // rows of pattern 4 span 16 * 144 B = 2.3 KB
  constexpr int UNROLL = 8;
  float accs[UNROLL] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // independent chains: the loads, not the adds, set the pace
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      // a different (bank-equivalent) address every load, so that nothing can be hoisted or merged
      unsigned a = (unsigned)__cvta_generic_to_shared(base) + (((unsigned)it + u) & 3u) * 1024u;
      asm volatile("mov.u32 %0, %0;" : "+r"(a));   // opaque: the compiler cannot prove two loads equal
      if constexpr (WIDTH == 1) {
        float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
        accs[u] += v;
      } else if constexpr (WIDTH == 2) {
        float v, w; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v), "=f"(w) : "r"(a) : "memory");
        accs[u] += v + w;
      } else {
        float v, w, x, y; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v), "=f"(w), "=f"(x), "=f"(y) : "r"(a) : "memory");
        accs[u] += (v + w) + (x + y);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float acc = 0.f;
  for (int u = 0; u < UNROLL; ++u) acc += accs[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int WIDTH, int PATTERN>
void run(float* out, long long* cyc, int sms) {
  const int iters = 20000;
  k<WIDTH, PATTERN><<<sms, 128>>>(out, 100, cyc);
  k<WIDTH, PATTERN><<<sms, 128>>>(out, iters, cyc);
  cudaDeviceSynchronize();
  long long h[512];
  cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; ++i) avg += (double)h[i];
  avg /= sms;
  // 4 warps x 8 loads per iteration on one SM
  printf("LDS.%-3d %-55s %6.2f cycles per warp-load\n", WIDTH * 32, kNames[PATTERN], avg / ((double)iters * 8 * 4));
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 128);
  cudaMalloc(&cyc, sizeof(long long) * 512);
  run<1, DISTINCT>(out, cyc, sms); run<1, HALF_DUP>(out, cyc, sms); run<1, PAIR_DUP>(out, cyc, sms); run<1, ALL_SAME>(out, cyc, sms);
  run<2, DISTINCT>(out, cyc, sms); run<2, HALF_DUP>(out, cyc, sms); run<2, PAIR_DUP>(out, cyc, sms); run<2, ALL_SAME>(out, cyc, sms);
  run<2, HALF_DUP_PITCH17>(out, cyc, sms);
  run<4, DISTINCT>(out, cyc, sms); run<4, HALF_DUP>(out, cyc, sms); run<4, PAIR_DUP>(out, cyc, sms); run<4, ALL_SAME>(out, cyc, sms);
  run<4, HALF_DUP_PITCH17>(out, cyc, sms);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
