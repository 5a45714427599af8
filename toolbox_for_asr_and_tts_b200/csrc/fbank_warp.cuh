// Warp-autonomous front-end kernel for sm_100a: every warp owns one QUAD (4 consecutive frames of one utterance) at a
// time, from HBM samples to the LFR-stacked, CMVN-normalised output rows, with no CTA-wide barrier in the loop.
//
//   * work list  = QuadDesc[n_quads], built on the device (build_quads_kernel); persistent warps stride over it, so
//                  the 4 warps of a CTA work on neighbouring quads and share their 240 overlapping samples through L1.
//   * samples    = the quad's 880 raw samples arrive by ONE bulk copy (cp.async.bulk, 1-D TMA) per quad into the warp's
//                  private 3.5 KB buffer, issued by lane 0 a whole quad ahead and completed on the warp's own mbarrier:
//                  no registers, no LSU-pipe traffic, no exposed L2 latency.  (Ends of the wave buffer and int16 PCM:
//                  128-bit loads -> registers -> shared, prefetched into L2 one quad ahead.)
//   * FFT / mel  = quad_stage1 / quad_stage2 / mel_stage of fbank_tile.cuh (packed f32x2 real FFT, interval mel).
//   * LFR + CMVN = closed form, per quad: the 4 log-mel frames go through a 1.25 KB per-warp staging tile and are
//                  written as 128-bit rows wherever clamp(n*i + jj - left, 0, T-1) == f  (VF:40-60: frame f is slot jj
//                  of row i), at most two places away from utterance edges; (x + shift) * scale in the reference's
//                  operation order (VF:23-37).  No halo frames, no CTA-wide second pass.
//
// The tile kernel (fbank_tile.cuh) remains for the CMVN-statistics pass and for frame shifts whose quad does not fit
// the 896-float buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fbank_tile.cuh"

#ifndef B200FE_OUT_LEAN      // 1: one divergent region, loads next to their stores; 0: all CMVN loads first (measured faster)
#define B200FE_OUT_LEAN 0
#endif
#ifndef B200FE_OUT_DIRECT    // 1: the mel epilogue writes the output rows itself (32-bit stores, no staging tile)
#define B200FE_OUT_DIRECT 0
#endif
#ifndef B200FE_OUT_V2        // 1: output stage on the interleaved CMVN table (QuadParams::cmvn_il): unconditional loads,
#define B200FE_OUT_V2 1      //    packed arithmetic, one 32-bit offset per store, predicated stores (no branches, no zeroed registers)
#endif
#ifndef B200FE_TMA_FENCE_ALWAYS   // 1: proxy fence in front of every bulk copy; 0: only after a generic fill wrote the buffer
#define B200FE_TMA_FENCE_ALWAYS 0
#endif
#ifndef B200FE_CLAIM_LATE    // 1 (needs DESC_SMEM): the work counter's atomic is issued at the top of a quad but its result is
#define B200FE_CLAIM_LATE 1  //    first read after stage 1; the next quad's head is fetched behind stage 2 and its bulk copy
#endif                       //    issued after stage 2: neither the atomic's nor the head's latency is exposed
#ifndef B200FE_CMVN_EARLY    // 1 (OUT_V2): the CMVN entries of the quad's primary targets are loaded BEFORE the mel stage, so
#define B200FE_CMVN_EARLY 0  //    their L1 latency hides behind it instead of sitting in front of the row stores
#endif
#ifndef B200FE_TMA_UNIFORM   // 1: every lane evaluates whether the next quad takes the bulk copy (no broadcast from lane 0)
#define B200FE_TMA_UNIFORM 1
#endif
#ifndef B200FE_CLAIM_ASM     // 1 (CLAIM_LATE): the counter's atomic as plain PTX, not atomicAdd() (see the call site)
#define B200FE_CLAIM_ASM 1
#endif
#ifndef B200FE_CLAIM2        // 1: quads claimed two ahead (atomic read at the end of the iteration); 0: one ahead
#define B200FE_CLAIM2 0
#endif
#ifndef B200FE_HEAD_PRED     // 1: round-2a behaviour, the next quad's head is loaded under a predicate (experiment baseline)
#define B200FE_HEAD_PRED 0
#endif
#ifndef B200FE_DESC_SMEM     // 1: the quad descriptors travel through shared memory (cp.async) instead of registers:
#define B200FE_DESC_SMEM 1   //    nothing is held across the FFT, no spills (measured -0.8 %); 0: __ldg into registers
#endif
#ifndef B200FE_CLAIM_STATIC  // 1: no work counter, warp w takes quads w, w + W, w + 2 W, ... (W = warps of the grid)
#define B200FE_CLAIM_STATIC 0
#endif
#ifndef B200FE_WARP_CTAS     // resident CTAs per SM the warp kernel is compiled and launched for
#define B200FE_WARP_CTAS 4
#endif
#ifndef B200FE_PAD_MODE      // 1: padding rows by bulk stores inside the fused kernel; 0: by the prep kernel in front of it
#define B200FE_PAD_MODE 0
#endif

namespace b200fe {

constexpr int kQuadVecs = 7;                      // float4 per lane and quad
constexpr int kQuadBuf = kQuadVecs * 32 * 4;      // 896 floats per warp
constexpr unsigned kNoTarget = 0xFFFFFFFFu;
constexpr int kTargetOffBits = 27;                // output offsets inside one utterance must stay below 2^27 floats
constexpr int kCmvnSlots = 1 << (32 - kTargetOffBits);   // slots of the interleaved CMVN table: every slot field indexes it

// One unit of work, fully resolved by build_quads_kernel (64 bytes).  The first 16 bytes are everything the sample
// fetch of a quad needs and are read one quad ahead; the second 16 bytes and the targets are read by the quad itself.
struct __align__(16) QuadDesc {
  long long g0;       // absolute index (in the wave buffer) of the first sample of the quad's first frame
  int nF;             // bits 0..7: frames in the quad (1..4); bits 8..11: frame t takes the generic LFR path
  int utt;
  int f0;             // first frame of the quad
  int T;              // frames of the utterance
  int rows;           // LFR rows of the utterance
  int row_begin;      // first row of the utterance in a rows-packed output (QuadParams::rows_cap < 0)
  unsigned tgt[8];    // tgt[2 t + k]: where frame f0 + t goes: (jj << 27) | (row * D + jj * M), or kNoTarget
};

struct QuadParams {
  const void* wave;       // float32 PCM in [-1, 1] or int16 PCM (sample / 32768, R:voice_interface.py:1008-1013)
  long long wave_total;   // samples addressable behind `wave`
  const QuadDesc* quads;
  int n_quads;
  int* next_quad;         // work counter (zeroed by build_quads_kernel): quads beyond the first wave are claimed dynamically
  float* feats;           // [batch, rows_cap, out_dim], or rows-packed [sum of rows, out_dim] when rows_cap < 0
  long long rows_cap;
  int frame_len, frame_shift, n_mels, lfr_m, lfr_n;
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
  const float* cmvn;      // nullptr or [2][out_dim]
  const float4* cmvn_il;  // [kCmvnSlots][n_mels / 4][2]: (shift4, scale4) of slot jj, piece l; identity when cmvn == nullptr
  const UttDesc* utts;    // [batch] in global memory: the padding fill reads n_rows from it
  int batch;
};

// Padding rows (pad_sequence zeros, VF:163-166) are written by the fused kernel itself, as bulk stores
// (cp.async.bulk shared -> global, SASS UBLKCP) of a zeroed shared-memory block: no registers, no LSU traffic, no
// separate pass over 133 MB (configs[1]) in front of the kernel - the stores ride on DRAM bandwidth the kernel leaves
// idle.  Every utterance's padding region is cut into kPadPieces pieces; piece i belongs to quad i (mod n_quads).
constexpr int kPadPieceShift = 6;
constexpr int kPadPieces = 1 << kPadPieceShift;
constexpr int kPadChunk = 2048;                   // bytes per bulk store = size of the zeroed block

__host__ __device__ inline size_t warp_smem_bytes() {
  return (size_t)kWarps * kQuadBuf * 4 + (size_t)kWarps * kYWarpF4 * 16 + (size_t)kTw2Total * 8 + (size_t)kWarps * 8 +
         (B200FE_PAD_MODE ? kPadChunk : 0);
}
// The quad of a (frame_len, frame_shift) pair fits the per-warp buffer: whole 16-byte vectors are stored, starting up
// to one vector minus one sample before the quad's first sample (3 floats for float32 PCM, 7 for int16 PCM).
__host__ __device__ inline bool warp_kernel_fits(int L, int S, bool pcm16 = false) {
  const int per = pcm16 ? 8 : 4;
  return ((per - 1 + 3 * S + L + per - 1) / per) * per <= kQuadBuf;
}

// Where frame f of an utterance with T frames / `rows` LFR rows goes (VF:40-60): slot jj of row i wherever
// n*i + jj - left == f.  With (i_hi, jj1) the solution of largest i, the others are (i_hi - k, jj1 + k n).  Two are
// encoded for the inline path; utterance edges (frames replicated by the LFR padding) and lfr_m > 2 lfr_n are flagged
// for the generic loop.  Runs in build_quads_kernel, so the integer divisions stay out of the hot loop.
__host__ __device__ inline bool quad_targets(int f, int T, int rows, int lfr_m, int lfr_n, int M, unsigned* t2) {
  const int left = (lfr_m - 1) / 2, D = lfr_m * M;
  const int i_hi = (f + left) / lfr_n, jj1 = (f + left) % lfr_n;
  t2[0] = t2[1] = kNoTarget;
  if (f == 0 || f == T - 1 || jj1 + 2 * lfr_n < lfr_m) return true;
  if (i_hi < rows && jj1 < lfr_m) t2[0] = ((unsigned)jj1 << kTargetOffBits) | (unsigned)(i_hi * D + jj1 * M);
  if (i_hi >= 1 && i_hi - 1 < rows && jj1 + lfr_n < lfr_m)
    t2[1] = ((unsigned)(jj1 + lfr_n) << kTargetOffBits) | (unsigned)((i_hi - 1) * D + (jj1 + lfr_n) * M);
  return false;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// ---- samples of a quad: HBM -> the warp's buffer, raw (pre-emphasis and DC removal happen in quad_stage1) --------
// buf[a_off + n] = sample n of the quad's first frame; a_off = position of that sample on the 16-byte grid of the
// wave buffer (0 for aligned offsets).
//
// float32 PCM, interior quads: ONE bulk copy (cp.async.bulk = 1-D TMA, SASS UBLKCP) per quad, issued by lane 0 a whole
// quad ahead and completed on the warp's own mbarrier, so the samples never pass through registers or the LSU pipe
// and no L2 latency is exposed.  The buffer is single: it is only read by the stage-1 loads at the top of a quad, and
// the next quad's copy is issued right after them.
// Quads that touch the ends of the wave buffer (a bulk copy needs 16-byte aligned, in-bounds ends) and int16 PCM take
// the generic fill below: vector loads on the 16-byte grid -> registers -> shared.
template <class SampleT>
__device__ __forceinline__ int quad_a_off(const void* wave, long long g0) {
  constexpr int kPer = 16 / (int)sizeof(SampleT);   // samples per 16-byte vector
  const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(wave) / sizeof(SampleT)) & (kPer - 1));
  return (int)((mis + (unsigned)(g0 & (kPer - 1))) & (kPer - 1));
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Every lane: the quad's samples can be fetched with one bulk copy (16-byte aligned, in-bounds ends).
__device__ __forceinline__ bool quad_tma_ok(const float* wave, long long wave_total, long long g0, int n_samples) {
  const int a_off = quad_a_off<float>(wave, g0);
  const long long ga = g0 - a_off;
  const int nv = (a_off + n_samples + 3) >> 2;
  return !(ga < 0 || ga + 4ll * nv > wave_total);
}

// Lane 0 only.  True when the quad's samples can be fetched with one bulk copy; then the copy is in flight.
__device__ __forceinline__ bool quad_fill_tma(const float* wave, long long wave_total, long long g0, int n_samples,
                                              float* buf, unsigned long long* bar, bool generic_wrote = true) {
  const int a_off = quad_a_off<float>(wave, g0);
  const long long ga = g0 - a_off;
  const int nv = (a_off + n_samples + 3) >> 2;
  if (ga < 0 || ga + 4ll * nv > wave_total) return false;
  const unsigned bytes = 16u * (unsigned)nv;
  // earlier generic-proxy accesses of the buffer (the generic fill's stores) are ordered before the async-proxy write
  // (a buffer that was only READ through the generic proxy since the last bulk copy needs no fence: the reads were
  // consumed by the FFT registers before the __syncwarp in front of this call)
  if (B200FE_TMA_FENCE_ALWAYS || generic_wrote) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(buf)), "l"(wave + ga), "r"(bytes), "r"(smem_u32(bar)) : "memory");
  return true;
}

// Lane 0 only: zero piece `item` of the padding regions (see kPadPieces).  n_rows = valid rows of utterance item >> 6.
__device__ __forceinline__ void pad_piece_store(float* feats, long long rows_cap, int D, int item, int n_rows,
                                                const void* zero_block) {
  const int u = item >> kPadPieceShift, piece = item & (kPadPieces - 1);
  const long long region = (rows_cap - n_rows) * (long long)D * 4;            // bytes, a multiple of 16
  if (region <= 0) return;
  const long long per = ((region / 16 + kPadPieces - 1) >> kPadPieceShift) * 16;
  const long long start = piece * per;
  if (start >= region) return;
  const long long len = region - start < per ? region - start : per;
  char* dst = reinterpret_cast<char*>(feats + ((long long)u * rows_cap + n_rows) * D) + start;
  const unsigned src = smem_u32(zero_block);
  for (long long off = 0; off < len; off += kPadChunk) {
    const unsigned bytes = (unsigned)(len - off < kPadChunk ? len - off : kPadChunk);
#if B200FE_PAD_MODE == 1
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src), "r"(bytes) : "memory");
#elif B200FE_PAD_MODE == 3   // experiment: same stores, L2-resident destination
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<char*>(feats) + (((unsigned long long)(dst + off) >> 3) & 0xFF800ull)), "r"(src), "r"(bytes) : "memory");
#else                        // experiment: address arithmetic only
    asm volatile("" ::"l"(dst + off), "r"(src), "r"(bytes) : "memory");
#endif
  }
}

template <class SampleT>
__device__ __forceinline__ void quad_fill_generic(const void* wave_any, long long wave_total, long long g0, int n_samples,
                                                  int lane, float* buf);

template <>
__device__ __forceinline__ void quad_fill_generic<float>(const void* wave_any, long long wave_total, long long g0,
                                                         int n_samples, int lane, float* buf) {
  const float* wave = static_cast<const float*>(wave_any);
  const int a_off = quad_a_off<float>(wave, g0);
  const long long ga = g0 - a_off;
  const int nv = (a_off + n_samples + 3) >> 2;
  float4* buf4 = reinterpret_cast<float4*>(buf);
#pragma unroll
  for (int u = 0; u < kQuadVecs; ++u) {
    if (lane + 32 * u >= nv) continue;
    const long long ab = ga + 4ll * (lane + 32 * u);
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ab >= 0 && ab + 4 <= wave_total) {
      x = ldg_stream4(wave + ab);
    } else {
      if (ab >= 0 && ab < wave_total) x.x = wave[ab];
      if (ab + 1 >= 0 && ab + 1 < wave_total) x.y = wave[ab + 1];
      if (ab + 2 >= 0 && ab + 2 < wave_total) x.z = wave[ab + 2];
      if (ab + 3 >= 0 && ab + 3 < wave_total) x.w = wave[ab + 3];
    }
    buf4[lane + 32 * u] = x;
  }
}

// int16 PCM: 8 samples per 16-byte vector, converted with the reference's own rule float(s) / 32768 (exact,
// R:voice_interface.py:1008-1013), so the result is bit-identical to handing the converted float32 buffer to the
// float path - at half the HBM / PCIe bytes.
constexpr int kQuadVecs16 = 4;
template <>
__device__ __forceinline__ void quad_fill_generic<short>(const void* wave_any, long long wave_total, long long g0,
                                                         int n_samples, int lane, float* buf) {
  const short* wave = static_cast<const short*>(wave_any);
  const int a_off = quad_a_off<short>(wave, g0);
  const long long ga = g0 - a_off;
  const int nv = (a_off + n_samples + 7) >> 3;
  constexpr float kScale = 1.0f / 32768.0f;
  float4* buf4 = reinterpret_cast<float4*>(buf);
  uint4 r[kQuadVecs16];
  bool fast[kQuadVecs16];
#pragma unroll
  for (int u = 0; u < kQuadVecs16; ++u) {   // all vector loads first: one exposed latency
    const long long ab = ga + 8ll * (lane + 32 * u);
    fast[u] = lane + 32 * u < nv && ab >= 0 && ab + 8 <= wave_total;
    r[u] = fast[u] ? ldg_stream_u4(wave + ab) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int u = 0; u < kQuadVecs16; ++u) {
    if (lane + 32 * u >= nv) continue;
    float x[8];
    if (fast[u]) {
      const unsigned w[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        x[2 * k] = (float)(short)(w[k] & 0xffffu) * kScale;
        x[2 * k + 1] = (float)(short)(w[k] >> 16) * kScale;
      }
    } else {
      const long long ab = ga + 8ll * (lane + 32 * u);
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = (ab + k >= 0 && ab + k < wave_total) ? (float)wave[ab + k] * kScale : 0.f;
    }
    buf4[2 * (lane + 32 * u)] = make_float4(x[0], x[1], x[2], x[3]);
    buf4[2 * (lane + 32 * u) + 1] = make_float4(x[4], x[5], x[6], x[7]);
  }
}

// SR: frame shift in 16-sample rows when it is a whole number of rows and known at compile time (10 for 400/160):
// the two frames of a pair then share their overlapping sample loads.  0 = generic.
// PACKED: rows-packed output (utterance u starts at row QuadDesc::row_begin) instead of the padded [B, rows_cap, D]; a
// separate instantiation so that the padded kernel's address arithmetic stays as it is.
template <int NROWS, bool EXACT, bool DITHER, class MELS, int SR, class SampleT, bool PACKED = false>
__global__ void __launch_bounds__(kCtaThreads, B200FE_WARP_CTAS)
fbank_warp_kernel(const QuadParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* bufs = reinterpret_cast<float*>(smem_raw);
  float4* xbuf = reinterpret_cast<float4*>(bufs + kWarps * kQuadBuf);
  float2* tw_s = reinterpret_cast<float2*>(xbuf + kWarps * kYWarpF4);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(tw_s + kTw2Total);
  float4* zero_block = reinterpret_cast<float4*>(bars + kWarps);   // kPadChunk bytes of zeros: source of the padding stores
  constexpr bool kTma = std::is_same<SampleT, float>::value;
#if B200FE_DESC_SMEM
  // per warp: [0] head of the next quad, [1..3] the rest of the current quad's descriptor ({f0, T, rows, row_begin}, targets)
  __shared__ __align__(16) int4 desc_s[kWarps][4];
#endif
#if B200FE_CLAIM_ASM
  __shared__ int claim_zero[kWarps];   // zeros, see the work claim
  if (threadIdx.x < kWarps) claim_zero[threadIdx.x] = 0;
#endif

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int j = tid & (kGroup - 1);
  const int grp_in_warp = lane >> 4;
  const int g = NROWS < 32 ? grp_in_warp : 0;
  const int L = p.frame_len, S = p.frame_shift, M = p.n_mels;
  const int lfr_m = p.lfr_m, lfr_n = p.lfr_n;
  const int D = lfr_m * M;
  const int lfr_left = (lfr_m - 1) / 2;

  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  if (B200FE_PAD_MODE)
    for (int i = tid; i < kPadChunk / 16; i += kCtaThreads) zero_block[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kTma && lane == 0) {
    mbar_init(bars + warp, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy
  }
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
  mel_preload(mel, threadIdx.x & 31);
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);
  __syncthreads();   // the only CTA-wide barrier: the twiddle tables (and the mbarriers)
  // Launched as a programmatic dependent of prep_warp_kernel (b200fe.cu: launch_warp): everything above reads
  // per-handle constants only and overlaps the tail of that kernel; the quad list and the work counter are its
  // output.  (A no-op when the launch carries no such attribute.)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  float* buf = bufs + warp * kQuadBuf;
  unsigned long long* bar = bars + warp;
  float4* yg = xbuf + warp * kYWarpF4 + grp_in_warp * kYGroupF4;
  float4* pbuf4 = xbuf + warp * kYWarpF4;
  float* lm_s = reinterpret_cast<float*>(pbuf4 + kSpecF4);   // log-mel staging tile, behind the warp's spectra
  const float2* tw_row = fft_twiddle_row<NROWS>(tw_s, j, grp_in_warp);
  const float2* c0_row = fft_c0s_row(tw_s, j);
  const int M4 = M >> 2;
  const bool act = lane < M4;      // lanes that move a float4 of a log-mel row (n_mels <= 128)
#if B200FE_OUT_V2
  // lanes beyond the row repeat its last piece (their stores are predicated off), so that every address below is valid
  // without a predicate
  const int lane_c = act ? lane : M4 - 1;
  const float4* const cm_il = p.cmvn_il + 2 * lane_c;
  const unsigned cm_slot_bytes = 32u * (unsigned)M4;
#endif

  // Work distribution: the first quad of every warp is static (neighbouring warps start on neighbouring quads), all
  // later ones are claimed from a global counter one quad ahead.  The 16-byte head {g0, nF, utt} of the claimed quad and
  // the rest of the current quad's descriptor ({f0, T, rows}, the LFR targets) travel by cp.async into four 16-byte
  // shared-memory slots of the warp at the top of the quad; the head is read after stage 1 (for the bulk copy of the
  // next quad's samples), the rest by the output stage - no descriptor register lives across the FFT.
  const int first_wave = gridDim.x * kWarps;
  const int last_quad = p.n_quads - 1;
  const int n_pad_items = p.batch << kPadPieceShift;
  int q = blockIdx.x * kWarps + warp;
  if (q >= p.n_quads) return;
  if (B200FE_PAD_MODE && lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero block
  long long g0_cur;
  int nF, slow, utt;
  {
    const int4 hd = __ldg(reinterpret_cast<const int4*>(p.quads + q));
    g0_cur = ((long long)hd.y << 32) | (unsigned)hd.x;
    nF = hd.z & 0xff; slow = (hd.z >> 8) & 0xf; utt = hd.w;
  }
#if B200FE_CLAIM2
  int qn = 0;
  if (lane == 0) qn = first_wave + atomicAdd(p.next_quad, 1);
  qn = __shfl_sync(0xffffffffu, qn, 0);
  bool have_next = qn < p.n_quads;
  int4 hdn = __ldg(reinterpret_cast<const int4*>(p.quads + min(qn, last_quad)));
#endif
  // the first quad's samples: bulk copy if possible (in_flight), else filled at the top of the loop
  unsigned phase = 0;
  int in_flight = 0;
  bool dirty = false;   // the generic fill has written the sample buffer since the last bulk copy (warp-uniform)
  if constexpr (kTma) {
    if (lane == 0)
      in_flight = quad_fill_tma(static_cast<const float*>(p.wave), p.wave_total, g0_cur, (nF - 1) * S + L, buf, bar, false);
    in_flight = __shfl_sync(0xffffffffu, in_flight, 0);
  }

  while (true) {
#if B200FE_CLAIM2
    int claim = 0;
    if (lane == 0) claim = atomicAdd(p.next_quad, 1);      // quad k+2; read at the end of this iteration
#elif B200FE_CLAIM_STATIC
    const int qn = q + first_wave;
    const bool have_next = qn < p.n_quads;
    const int4 hdn = __ldg(reinterpret_cast<const int4*>(p.quads + min(qn, last_quad)));
#elif B200FE_CLAIM_LATE && B200FE_DESC_SMEM
    int qn = 0;
#if B200FE_CLAIM_ASM
    // atom.inc, not atom.add / atomicAdd(): ptxas turns an add under a divergent branch into a warp-aggregated sequence (vote, leader atomic,
    // SHFL of the result) whose shuffle waits for the atomic on the spot - the round trip this variant is about
    if (lane == 0) {
      // the address goes through a value ptxas cannot prove warp-uniform (a zero read back from shared memory): an
      // atomic on a provably uniform address is aggregated whatever its width or spelling (add, inc, inline PTX)
      const int* ctr = p.next_quad + *const_cast<const volatile int*>(&claim_zero[warp]);
      asm volatile("atom.global.add.u32 %0, [%1], 1;" : "=r"(qn) : "l"(ctr) : "memory");
    }
#else
    if (lane == 0) qn = atomicAdd(p.next_quad, 1);   // consumed after stage 1
#endif
    bool have_next = false;
    if (lane >= 1 && lane < 4) {   // the rest of this quad's descriptor
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&desc_s[warp][lane])),
                   "l"(reinterpret_cast<const int4*>(p.quads + q) + lane) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
#else
    int qn = 0;
    if (lane == 0) qn = first_wave + atomicAdd(p.next_quad, 1);
    qn = __shfl_sync(0xffffffffu, qn, 0);
    const bool have_next = qn < p.n_quads;
    // unconditional (clamped) load: a predicated one makes the compiler copy the fields right behind the load, which
    // exposes its whole L2 latency at the top of the quad
#if B200FE_HEAD_PRED
    int4 hdn = make_int4(0, 0, 0, 0);
    if (have_next) hdn = __ldg(reinterpret_cast<const int4*>(p.quads + qn));
#elif B200FE_DESC_SMEM
    // lanes 1..3: the rest of this quad's descriptor, lane 0: the head of the next quad - all by cp.async into the warp's
    // shared-memory slots (no registers held across the FFT); waited for after stage 1
    {
      const int4* src = lane == 0 ? reinterpret_cast<const int4*>(p.quads + min(qn, last_quad))
                                  : reinterpret_cast<const int4*>(p.quads + q) + lane;
      if (lane < 4) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&desc_s[warp][lane])), "l"(src) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
#else
    const int4 hdn = __ldg(reinterpret_cast<const int4*>(p.quads + min(qn, last_quad)));
#endif
#endif
#if B200FE_DESC_SMEM
    const volatile int4* ds = desc_s[warp];
#define hd1 (*const_cast<const int4*>(ds + 1))
#else
    const int4 hd1 = __ldg(reinterpret_cast<const int4*>(p.quads + q) + 1);   // f0, T, rows, pad: used by the output stage
#endif
    int pad_rows = 0;                                                          // padding piece q: read now, stored mid-quad
    if (B200FE_PAD_MODE && lane == 0 && q < n_pad_items) pad_rows = __ldg(&p.utts[q >> kPadPieceShift].n_rows);

    // ---- this quad's samples: the bulk copy was issued a whole quad ago; otherwise (ends of the wave buffer, int16
    //      PCM: prefetched into L2 a quad ago) fill the buffer now
    if (in_flight) {
      mbar_wait(bar, phase);
      phase ^= 1u;
    } else {
      quad_fill_generic<SampleT>(p.wave, p.wave_total, g0_cur, (nF - 1) * S + L, lane, buf);
      dirty = true;
      __syncwarp();
    }
    const int a_off = quad_a_off<SampleT>(p.wave, g0_cur);

    // ---- stage 1 (samples -> registers -> real 32-point FFT) and stage 2 (transpose, 16-point FFT, power spectra)
    const int fA = 2 * grp_in_warp;
    const bool vA = fA < nF, vB = fA + 1 < nF;
    long long g0_next = 0;
    int nF_next = 0, slow_next = 0, utt_next = 0;
    {
      f2 zr[16], zi[16], y0, y16;
      int f0_dither = 0;
#if B200FE_DESC_SMEM
      if constexpr (DITHER) {   // the frame index seeds the dither: the descriptor is needed before stage 1
        if (lane < 4) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
      }
#endif
      if constexpr (DITHER) f0_dither = hd1.x;
      quad_stage1<NROWS, EXACT, DITHER, SR>(buf + a_off + fA * S, vA, vB, S, L, win, p.preemph, p.remove_dc, p.dither,
                                            p.seed, (unsigned)utt, (unsigned)(f0_dither + fA), j, g, zr, zi, y0, y16);
      __syncwarp();   // every lane is done with the sample buffer and with the previous quad's staging tile
      // ---- the next quad's samples start their way from HBM now: one bulk copy into the (now free) buffer, or one L2
      //      prefetch per 128-byte line for the generic fill
      in_flight = 0;
#if B200FE_CLAIM_LATE && B200FE_DESC_SMEM
      // the claim's result is read here; the claimed quad's head travels while stage 2 runs
      qn = first_wave + __shfl_sync(0xffffffffu, qn, 0);
      have_next = qn < p.n_quads;
      if (lane == 0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&desc_s[warp][0])),
                     "l"(reinterpret_cast<const int4*>(p.quads + min(qn, last_quad))) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      quad_stage2(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
#endif
#if B200FE_DESC_SMEM
      if (lane < 4) asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      const int4 hdn = *const_cast<const int4*>(ds);
#endif
      if (have_next) {
        g0_next = ((long long)hdn.y << 32) | (unsigned)hdn.x;
        nF_next = hdn.z & 0xff; slow_next = (hdn.z >> 8) & 0xf; utt_next = hdn.w;
        if constexpr (kTma) {
#if B200FE_TMA_UNIFORM
          in_flight = quad_tma_ok(static_cast<const float*>(p.wave), p.wave_total, g0_next, (nF_next - 1) * S + L);
          if (lane == 0 && in_flight)
            quad_fill_tma(static_cast<const float*>(p.wave), p.wave_total, g0_next, (nF_next - 1) * S + L, buf, bar, dirty);
#else
          if (lane == 0)
            in_flight = quad_fill_tma(static_cast<const float*>(p.wave), p.wave_total, g0_next, (nF_next - 1) * S + L, buf, bar, dirty);
          in_flight = __shfl_sync(0xffffffffu, in_flight, 0);
#endif
          dirty = false;
        } else {
          const char* base = static_cast<const char*>(p.wave);
          const long long first = (g0_next * (long long)sizeof(SampleT) - 16) & ~127ll;
          const long long idx = first + 128ll * lane;
          const long long last = (g0_next + 3 * S + L) * (long long)sizeof(SampleT);
          if (idx >= 0 && idx < last && idx < p.wave_total * (long long)sizeof(SampleT))
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + idx));
        }
      }
      // ---- this quad's share of the padding rows: bulk stores of the zero block
      if (B200FE_PAD_MODE != 0 && lane == 0 && q < n_pad_items) {
        pad_piece_store(p.feats, p.rows_cap, D, q, pad_rows, zero_block);
        for (int it = q + p.n_quads; it < n_pad_items; it += p.n_quads)      // fewer quads than pieces: rare
          pad_piece_store(p.feats, p.rows_cap, D, it, __ldg(&p.utts[it >> kPadPieceShift].n_rows), zero_block);
      }
#if !(B200FE_CLAIM_LATE && B200FE_DESC_SMEM)
      quad_stage2(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
#endif
    }
#if B200FE_DESC_SMEM
#define tg0 (*reinterpret_cast<const uint4*>(const_cast<const int4*>(ds + 2)))
#define tg1 (*reinterpret_cast<const uint4*>(const_cast<const int4*>(ds + 3)))
#else
    const uint4 tg0 = __ldg(reinterpret_cast<const uint4*>(p.quads + q) + 2);   // targets of frames 0, 1
    const uint4 tg1 = __ldg(reinterpret_cast<const uint4*>(p.quads + q) + 3);   // targets of frames 2, 3
#endif

#if B200FE_OUT_DIRECT
    // ---- mel + log + LFR + CMVN in one go: the lane that holds filter iv of the quad's 4 frames writes it straight to
    //      the (row, slot) targets of each frame - out[target + iv] = (x + shift[slot][iv]) * scale[slot][iv]
    //      (VF:34-35, VF:40-60) - as 32-bit stores that a warp coalesces into one or two 128-byte lines.  No staging tile,
    //      no regrouping into 128-bit pieces, the CMVN entries are per-lane scalar loads.
    {
      float* out_u = p.feats + (PACKED ? (long long)hd1.w : (long long)utt * p.rows_cap) * D;
      const float* cm = p.cmvn;
      const unsigned tgt[8] = {tg0.x, tg0.y, tg0.z, tg0.w, tg1.x, tg1.y, tg1.z, tg1.w};
      constexpr unsigned kOffMask = (1u << kTargetOffBits) - 1;
      mel_stage<MELS>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b, float c, float d) {
        const float val[4] = {a, b, c, d};
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const unsigned cd = tgt[2 * t + k];
            if (cd == kNoTarget) continue;                       // warp-uniform
            float x = val[t];
            if (cm) {
              const int jm = (int)(cd >> kTargetOffBits) * M + iv;
              x = (x + __ldg(cm + jm)) * __ldg(cm + D + jm);
            }
            out_u[(cd & kOffMask) + iv] = x;
          }
        if (slow) {   // first / last frame of the utterance (replicated by the LFR padding), or lfr_m > 2 lfr_n
          const int f0 = hd1.x, T = hd1.y, rows = hd1.z;
#pragma unroll 1
          for (int t = 0; t < nF; ++t) {
            if (!((slow >> t) & 1)) continue;
            const int f = f0 + t;
            const int num = f + lfr_left - (lfr_m - 1);
            const int i_lo = (f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n;
            const int i_top = f == T - 1 ? rows - 1 : min((f + lfr_left) / lfr_n, rows - 1);
            const float v = t == 0 ? a : (t == 1 ? b : (t == 2 ? c : d));
#pragma unroll 1
            for (int i = i_lo; i <= i_top; ++i)
#pragma unroll 1
              for (int jj = 0; jj < lfr_m; ++jj)
                if (min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) {
                  float x = v;
                  if (cm) x = (x + __ldg(cm + jj * M + iv)) * __ldg(cm + D + jj * M + iv);
                  out_u[(long long)i * D + jj * M + iv] = x;
                }
          }
        }
      });
    }
#else
#if B200FE_OUT_V2 && B200FE_CMVN_EARLY
    float4 sh_e[4], sc_e[4];
    {
      const unsigned t4[4] = {tg0.x, tg0.z, tg1.x, tg1.z};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4* c = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(cm_il) +
                                                          (size_t)((t4[t] >> kTargetOffBits) * cm_slot_bytes));
        sh_e[t] = __ldg(c);
        sc_e[t] = __ldg(c + 1);
      }
    }
#endif
    // ---- log-mel of the 4 frames into the warp's staging tile
    mel_stage<MELS>(mel, pbuf4, lane, M, p.log_floor, [&](int iv, float a, float b, float c, float d) {
      lm_s[iv] = a;
      lm_s[M + iv] = b;
      lm_s[2 * M + iv] = c;
      lm_s[3 * M + iv] = d;
    });
    __syncwarp();
#endif

#if B200FE_OUT_DIRECT
#elif B200FE_OUT_V2
    // ---- LFR + CMVN: each frame's row of n_mels goes, as 128-bit pieces, to the (row, slot) pairs of its descriptor.
    //      The slot field of a target code indexes the interleaved CMVN table directly (kNoTarget reads its last, unused
    //      slot), so all loads are unconditional and issued first; the arithmetic is packed; a store costs one 32-bit
    //      offset and a predicate.  Same operation order as the reference, (x + shift) * scale (VF:34-35): bit-identical
    //      to the scalar form.
    {
      char* const out_b = reinterpret_cast<char*>(p.feats + (PACKED ? (long long)hd1.w : (long long)utt * p.rows_cap) * D + 4 * lane_c);
      const unsigned tgt[8] = {tg0.x, tg0.y, tg0.z, tg0.w, tg1.x, tg1.y, tg1.z, tg1.w};
      const float4* lm4 = reinterpret_cast<const float4*>(lm_s) + lane_c;
      constexpr unsigned kOffMask = (1u << kTargetOffBits) - 1;
      auto cm_at = [&](unsigned code) {
        return reinterpret_cast<const float4*>(reinterpret_cast<const char*>(cm_il) + (size_t)((code >> kTargetOffBits) * cm_slot_bytes));
      };
      auto cmvn4 = [&](const float4& v, const float4& sh, const float4& sc) {
        const f2 a = mul2(add2(make_float2(v.x, v.y), make_float2(sh.x, sh.y)), make_float2(sc.x, sc.y));
        const f2 b = mul2(add2(make_float2(v.z, v.w), make_float2(sh.z, sh.w)), make_float2(sc.z, sc.w));
        return make_float4(a.x, a.y, b.x, b.y);
      };
      auto store_if = [&](bool on, unsigned code, const float4& o) {
        float* dst = reinterpret_cast<float*>(out_b + (size_t)(4u * (code & kOffMask)));
        asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t"
                     "@q st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};\n\t}"
                     :: "l"(dst), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w), "r"((unsigned)on) : "memory");
      };
      float4 v[4], sh[4], sc[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        v[t] = lm4[t * M4];
#if B200FE_CMVN_EARLY
        sh[t] = sh_e[t];
        sc[t] = sc_e[t];
#else
        const float4* c = cm_at(tgt[2 * t]);
        sh[t] = __ldg(c);
        sc[t] = __ldg(c + 1);
#endif
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) store_if(act && tgt[2 * t] != kNoTarget, tgt[2 * t], cmvn4(v[t], sh[t], sc[t]));
      // second slot (one frame in lfr_n when lfr_m = lfr_n + 1): warp-uniform branch
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned code = tgt[2 * t + 1];
        if (code != kNoTarget) {
          const float4* c = cm_at(code);
          store_if(act, code, cmvn4(v[t], __ldg(c), __ldg(c + 1)));
        }
      }
      if (slow) {   // first / last frame of the utterance (replicated by the LFR padding), or lfr_m > 2 lfr_n
        const int f0 = hd1.x, T = hd1.y, rows = hd1.z;
#pragma unroll 1
        for (int t = 0; t < nF; ++t) {
          if (!((slow >> t) & 1)) continue;
          const int f = f0 + t;
          const int num = f + lfr_left - (lfr_m - 1);
          const int i_lo = (f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n;
          const int i_top = f == T - 1 ? rows - 1 : min((f + lfr_left) / lfr_n, rows - 1);
          const float4 vv = lm4[t * M4];
#pragma unroll 1
          for (int i = i_lo; i <= i_top; ++i)
#pragma unroll 1
            for (int jj = 0; jj < lfr_m; ++jj)
              if (min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) {
                const float4* c = cm_at((unsigned)jj << kTargetOffBits);
                const float4 o = cmvn4(vv, __ldg(c), __ldg(c + 1));
                if (act) stg_stream4(reinterpret_cast<float*>(out_b) + (long long)i * D + jj * M, o);
              }
        }
      }
    }

#elif B200FE_OUT_LEAN
    // ---- LFR + CMVN: each frame's row of n_mels goes, as 128-bit pieces, to the (row, slot) pairs of its descriptor.
    //      One divergent region for the lanes that carry a piece; inside it every branch is warp-uniform.
    if (act) {
      const int f0 = hd1.x, T = hd1.y, rows = hd1.z;
      float* out_l = p.feats + (PACKED ? (long long)hd1.w : (long long)utt * p.rows_cap) * D + 4 * lane;
      const float* cm_l = p.cmvn ? p.cmvn + 4 * lane : nullptr;
      const float4* lm4 = reinterpret_cast<const float4*>(lm_s) + lane;
      const unsigned tgt[8] = {tg0.x, tg0.y, tg0.z, tg0.w, tg1.x, tg1.y, tg1.z, tg1.w};
      auto emit = [&](const float4& v, int jm, float* dst) {   // (x + shift) * scale in the reference's order, VF:34-35
        float4 o = v;
        if (cm_l) {
          const float4 sh = __ldg(reinterpret_cast<const float4*>(cm_l + jm));
          const float4 sc = __ldg(reinterpret_cast<const float4*>(cm_l + D + jm));
          o = make_float4((v.x + sh.x) * sc.x, (v.y + sh.y) * sc.y, (v.z + sh.z) * sc.z, (v.w + sh.w) * sc.w);
        }
        stg_stream4(dst, o);
      };
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned c0 = tgt[2 * t], c1 = tgt[2 * t + 1];
        if (c0 == kNoTarget && c1 == kNoTarget) continue;          // invalid frame, or the generic path below
        const float4 v = lm4[t * M4];
        if (c0 != kNoTarget) emit(v, (int)(c0 >> kTargetOffBits) * M, out_l + (c0 & ((1u << kTargetOffBits) - 1)));
        if (c1 != kNoTarget) emit(v, (int)(c1 >> kTargetOffBits) * M, out_l + (c1 & ((1u << kTargetOffBits) - 1)));
      }
      if (slow) {   // first / last frame of the utterance (replicated by the LFR padding), or lfr_m > 2 lfr_n
#pragma unroll 1
        for (int t = 0; t < nF; ++t) {
          if (!((slow >> t) & 1)) continue;
          const int f = f0 + t;
          const int num = f + lfr_left - (lfr_m - 1);
          const int i_lo = (f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n;
          const int i_top = f == T - 1 ? rows - 1 : min((f + lfr_left) / lfr_n, rows - 1);
          const float4 v = lm4[t * M4];
#pragma unroll 1
          for (int i = i_lo; i <= i_top; ++i)
#pragma unroll 1
            for (int jj = 0; jj < lfr_m; ++jj)
              if (min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) emit(v, jj * M, out_l + (long long)i * D + jj * M);
        }
      }
    }

#else
    // ---- LFR + CMVN: each frame's row of n_mels goes, as 128-bit pieces, to the (row, slot) pairs of its descriptor
    {
      const int f0 = hd1.x, T = hd1.y, rows = hd1.z;
      float* out_l = p.feats + (PACKED ? (long long)hd1.w : (long long)utt * p.rows_cap) * D + 4 * lane;
      const unsigned tgt[8] = {tg0.x, tg0.y, tg0.z, tg0.w, tg1.x, tg1.y, tg1.z, tg1.w};
      const float* cm_l = p.cmvn ? p.cmvn + 4 * lane : nullptr;
      const float4* lm4 = reinterpret_cast<const float4*>(lm_s) + lane;
      auto cmvn4 = [&](const float4& v, const float4& sh, const float4& sc) {   // (x + shift) * scale, VF:34-35
        return make_float4((v.x + sh.x) * sc.x, (v.y + sh.y) * sc.y, (v.z + sh.z) * sc.z, (v.w + sh.w) * sc.w);
      };
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      // primary slot of the 4 frames: all loads first (one exposed latency), then arithmetic and stores
      float4 v[4], sh[4], sc[4];
      bool ok[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        ok[t] = tgt[2 * t] != kNoTarget && act;
        v[t] = act ? lm4[t * M4] : zero4;
        sh[t] = sc[t] = zero4;
        if (cm_l && ok[t]) {
          const int jm = (int)(tgt[2 * t] >> kTargetOffBits) * M;
          sh[t] = __ldg(reinterpret_cast<const float4*>(cm_l + jm));
          sc[t] = __ldg(reinterpret_cast<const float4*>(cm_l + D + jm));
        }
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (ok[t]) stg_stream4(out_l + (int)(tgt[2 * t] & ((1u << kTargetOffBits) - 1)), cm_l ? cmvn4(v[t], sh[t], sc[t]) : v[t]);
      // second slot (one frame in lfr_n when lfr_m = lfr_n + 1): warp-uniform branch
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const unsigned code = tgt[2 * t + 1];
        if (code != kNoTarget && act) {
          float4 o = v[t];
          if (cm_l) {
            const int jm = (int)(code >> kTargetOffBits) * M;
            o = cmvn4(o, __ldg(reinterpret_cast<const float4*>(cm_l + jm)), __ldg(reinterpret_cast<const float4*>(cm_l + D + jm)));
          }
          stg_stream4(out_l + (int)(code & ((1u << kTargetOffBits) - 1)), o);
        }
      }
      if (slow) {   // first / last frame of the utterance (replicated by the LFR padding), or lfr_m > 2 lfr_n
#pragma unroll 1
        for (int t = 0; t < nF; ++t) {
          if (!((slow >> t) & 1)) continue;
          const int f = f0 + t;
          const int num = f + lfr_left - (lfr_m - 1);
          const int i_lo = (f == 0 || num <= 0) ? 0 : (num + lfr_n - 1) / lfr_n;
          const int i_top = f == T - 1 ? rows - 1 : min((f + lfr_left) / lfr_n, rows - 1);
          const float4 v = act ? lm4[t * M4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int i = i_lo; i <= i_top; ++i)
#pragma unroll 1
            for (int jj = 0; jj < lfr_m; ++jj)
              if (act && min(max(lfr_n * i + jj - lfr_left, 0), T - 1) == f) {
                float4 o = v;
                if (cm_l) {
                  const float4 sh = __ldg(reinterpret_cast<const float4*>(cm_l + jj * M));
                  const float4 sc = __ldg(reinterpret_cast<const float4*>(cm_l + D + jj * M));
                  o.x = (v.x + sh.x) * sc.x;
                  o.y = (v.y + sh.y) * sc.y;
                  o.z = (v.z + sh.z) * sc.z;
                  o.w = (v.w + sh.w) * sc.w;
                }
                stg_stream4(out_l + (long long)i * D + jj * M, o);
              }
        }
      }
    }

#endif
    if (!have_next) break;
#if B200FE_DESC_SMEM
    __syncwarp();   // the descriptor slots are read: the next quad's copies may overwrite them
#undef hd1
#undef tg0
#undef tg1
#endif
    q = qn;
    g0_cur = g0_next;
    nF = nF_next; slow = slow_next; utt = utt_next;
#if B200FE_CLAIM2
    qn = first_wave + __shfl_sync(0xffffffffu, claim, 0);
    have_next = qn < p.n_quads;
    hdn = __ldg(reinterpret_cast<const int4*>(p.quads + min(qn, last_quad)));
#endif
  }
  // the zero block must outlive the bulk stores that read it
  if (B200FE_PAD_MODE && lane == 0) {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

}  // namespace b200fe
