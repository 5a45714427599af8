// Fused front-end kernel for sm_100a: framing + DC removal + pre-emphasis + window -> 512-point power
// spectrum -> sparse Kaldi mel -> log -> LFR stack -> CMVN, one persistent CTA per tile of LFR rows.
//
// Replaces, for a whole ragged batch in one launch, what the reference runs per utterance on the CPU:
//   TA:154-217 (_get_window), TA:616-618 (rfft/abs/pow), TA:621-633 (mel mm + log), VF:40-60 (apply_lfr),
//   VF:23-37 (apply_cmvn), VF:163-166 (pad_sequence).
//
// Work decomposition ("fbank tile kernel"):
//   * tile  = rows_per_tile LFR rows of one utterance = F <= kFMax consecutive frames (1 frame of halo is recomputed
//             between neighbouring tiles);  persistent CTAs stride over the tile list.
//   * stage = the tile's raw samples are read ONCE from HBM with 128-bit loads and kept in shared memory (each sample
//             feeds 2.5 frames); pre-emphasis and DC removal happen on the way into the FFT registers (quad_stage1).
//   * FFT   = a group of 16 threads transforms two real frames at once, one frame per lane of packed f32x2 registers
//             (FFMA2/FADD2/FMUL2: half the issue slots per flop).  Thread j owns samples n = 16 i + j: a REAL 32-point
//             register FFT per thread (16-point complex FFT of even/odd samples + split), ONE shared-memory
//             transpose of the 17 non-redundant columns, then one complex 16-point register FFT per thread on column
//             j (column 16 for thread 0), whose 16 bins cover columns j and 32-j of the real spectrum.  Column 0
//             (bins 0, 32, .., 224) is a direct 16-term sum by threads 0..7.  No shuffles on the FFT path; a warp
//             (2 groups = 4 frames) never waits for another warp inside a tile.
//   * mel   = power spectra of the warp's 4 frames are interleaved [bin][4] so one 128-bit shared load feeds 4 FMAs;
//             the filterbank is stored sparse (<= 2 non-zeros per FFT bin).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_codelets.cuh"

#ifndef B200FE_S1_UNCOND
#define B200FE_S1_UNCOND 1
#endif
#ifndef B200FE_DC_REDUCE4    // 1: the frame sums are reduced over the 16 lanes in TWO dependent levels of three independent shuffles
#define B200FE_DC_REDUCE4 0  //    (lanes j^1, j^2, j^3, then j^4, j^8, j^12) instead of four levels of one: 12 shuffles, half the chain
#endif
#ifndef B200FE_DC_LATE       // 1: stage 1 multiplies by the window BEFORE the frame-mean shuffle reduction and adds -(1-p) mean w[n]
#define B200FE_DC_LATE 0     //    afterwards (one FMA): the 25 multiplies run while the 4 shuffle levels are in flight
#endif
#ifndef B200FE_C0_SHFL       // 1: column 0 of the spectrum (bins 32 t) by a 16-point radix-2 FFT ACROSS the 16 lanes of a group (14
#define B200FE_C0_SHFL 1     //    shuffles, 3 table loads) instead of 8 broadcast row loads + 4 table loads per lane through shared memory
#endif
#ifndef B200FE_S2_TW_EARLY   // 1: stage 2 loads its twiddle row and column-0 weights BEFORE the barrier that follows the transpose
#define B200FE_S2_TW_EARLY 0 //    stores (they do not depend on them), so only the transposed rows are loaded behind it
#endif
#ifndef B200FE_MEL_WORD_REGS // 1: the lane's interval words of a fixed-shape bank live in registers for the whole kernel
#define B200FE_MEL_WORD_REGS 1
#endif
#ifndef B200FE_MEL_COMPACT   // 1: the fixed-shape mel stage reads ONE float per (bin, lane) - the up-slope weight - and derives
#define B200FE_MEL_COMPACT 1 //    the down-slope weight (they sum to the bank's scale): half the weight bytes through L1
#endif

namespace b200fe {

constexpr int kGroup = 16;          // threads per frame pair
constexpr int kCtaThreads = 128;    // 4 warps, 8 groups
constexpr int kWarps = kCtaThreads / 32;
constexpr int kFMax = 32;           // frames per tile (upper bound)
constexpr int kXRow = 18;           // float2 per twiddle row: 16 + 2 pad -> 144 B, conflict-free 128-bit reads
// Transpose buffer of the packed real FFT, per 16-thread group: rows k1 = 1..16 of 16 float4 (re A, re B, im A, im B),
// pitch 17 float4 (272 B: 128-bit row reads by 16 threads are conflict-free), then row 0 as 16 float2 (real bins).
constexpr int kYPitch = 17;
// The group pitch is 16 banks (mod 32) so that a quarter-warp reading 4 columns of BOTH groups (stage 2 assigns lanes
// 2c, 2c+1 to column c of group 0 / group 1) touches 32 distinct banks.
constexpr int kYGroupF4 = 16 * kYPitch + 12;      // 284 float4 = 4544 B
constexpr int kYWarpF4 = 2 * kYGroupF4;           // 9088 B per warp, aliased by the warp's power spectra
constexpr int kSpecPitch = 256 + 8;               // float2 between the two groups' power spectra: again 16 banks apart
constexpr int kSpecF4 = kSpecPitch;               // float4 taken by both spectra (2 * kSpecPitch float2)
// Twiddles of the packed FFT: [16 rows k1 = 1..16][kXRow float2] then the column-0 table [8][kC0Pitch].
constexpr int kTwPitch = 17;         // float2 per row: 136 B, so the 16 lanes of a group read 32 distinct banks with 64-bit loads
constexpr int kTw2Table = 16 * kTwPitch;
constexpr int kC0Pitch = 10;                      // float2; 80 B rows: conflict-free 128-bit reads by 8 threads
constexpr int kTw2Total = kTw2Table + 8 * kC0Pitch;       // float2
constexpr int kC0sPitch = 5;                      // B200FE_C0_SHFL: [16 lanes][kC0sPitch] float2 in the same 80 float2: the lane's
                                                  // twiddles of butterfly stages 0..2; 40 B rows: conflict-free 64-bit reads
static_assert(16 * kC0sPitch <= 8 * kC0Pitch, "the lane-FFT table must fit the column-0 table");
constexpr int kMaxMels = 128;
constexpr int kMaxInt = kMaxMels + 1;  // intervals between consecutive filter centres
constexpr int kMelRounds = (kMaxMels + 30) / 31;  // rounds of 31 filters: lane <-> interval, lane 31 only feeds lane 30
constexpr int kMelSlots = 64;       // sum over rounds of the round's widest interval (bins), upper bound
constexpr int kNfft = 512;
constexpr float kMelScale = 0.25f;  // the stored powers are 4 |X|^2: the mel weights carry the 0.25

struct UttDesc {          // built on the host by b200fe_plan/forward
  long long wave_off;     // first sample of the utterance inside the wave buffer
  int n_samples;
  int n_frames;           // T      (TA:65-70)
  int n_rows;             // T_lfr  (VF:43)
  int tile_begin;         // index of the utterance's first tile in the launch-wide tile list
  int quad_begin;         // index of the utterance's first quad in the launch-wide quad list (fbank_warp.cuh)
  int row_begin;          // rows of all earlier utterances: where this one starts in a rows-packed output
};

// One tile of the launch-wide work list, fully resolved by build_tiles_kernel so that the fused kernel never chases a
// dependent global load: 32 bytes, read one tile ahead.
struct __align__(16) TileDesc {
  long long g0;   // absolute index (in the wave buffer) of the first sample of the tile's first frame
  int utt;
  int row0;       // first LFR row of the tile
  int f_lo;       // first frame of the tile
  int F;          // frames in the tile (<= kFMax)
  int nrow;       // LFR rows in the tile
  int T;          // frames of the utterance
};

struct TileParams {
  const float* wave;
  long long wave_total;
  const UttDesc* utts;
  const TileDesc* tiles;  // [n_tiles], built on the device from utts
  int* next_tile;         // work counter (zeroed by build_tiles_kernel): tiles beyond the first wave are claimed dynamically
  int batch;
  int n_tiles;
  float* feats;           // [batch, rows_cap, out_dim]
  long long rows_cap;
  double* stats;          // nullptr or [2*out_dim + 1]
  int frame_len;          // L  (window_size, <= 512)
  int frame_shift;        // S
  int n_mels;
  int lfr_m, lfr_n, rows_per_tile;
  int e_cap;              // floats reserved for the staged samples
  float preemph;
  int remove_dc;
  float log_floor;
  float dither;
  unsigned long long seed;
  const float* window;    // [512] window * (2^15 if upscale), zero beyond L
  const float2* twiddle;  // [kTw2Total]: row k1-1, entry c = s(k1) * exp(-2*pi*i*c*k1/512), s = 2 for k1 in {8, 16}
                          // else 1.  Then the column-0 table: row t, entry c = 2 * exp(-2*pi*i*c*t/16), c < 8.
  // sparse mel bank by interval between filter centres.  Round r covers intervals 31 r .. 31 r + 31 (the last one is
  // the first of the next round: here it only supplies the down-slope sum of filter 31 r + 30), one per lane, placed by
  // build_interval_table so that the gathers are bank-conflict free.  Every round has a warp-uniform trip count
  // mel_cnt[r] = its widest interval; weights are zero-padded to it and stored lane-transposed:
  // mel_w[(mel_base[r] + q) * 32 + lane] = (up, down) weight x 0.25 of bin (mel_lo[32 r + lane] & 0xfff) + q; the rest
  // of the mel_lo word names the partner lane, the filter index and whether the lane outputs.  Read through L1.
  const float2* mel_w;
  const int* mel_lo;      // [32 * mel_rounds]
  int mel_rounds;
  int mel_cnt[kMelRounds];
  int mel_base[kMelRounds];
  const float* cmvn;      // nullptr or [2][out_dim]
};

__host__ __device__ inline size_t tile_smem_bytes(int e_cap, int n_mels) {
  size_t b = 0;
  b += (size_t)e_cap * 4;                           // staged raw samples
  b += (size_t)kWarps * kYWarpF4 * 16;              // transpose buffers (aliased by the power spectra)
  b += (size_t)kFMax * n_mels * 4;                  // log-mel of the tile
  b += kTw2Total * 8;                               // twiddles (stage 2 + column 0)
  return b;
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Counter-based N(0,1) for dither, cf. TA:179-181 (independent noise per (frame, sample)): one Box-Muller pair per
// (utterance, frame PAIR, sample-in-frame) - the cosine branch belongs to the even frame 2k, the sine branch to 2k+1 -
// so a thread that transforms frames (fa, fa+1) needs one hash / log / sqrt / sincos per sample when fa is even.
__device__ __forceinline__ unsigned mix32(unsigned x) {   // 32-bit avalanche (two multiply / xorshift rounds)
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// salt of one (seed, utterance): the 64-bit mixing is done once per call site, not per sample
__device__ __forceinline__ unsigned dither_salt(unsigned long long seed, unsigned utt) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)utt + 1);
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned)z ^ (unsigned)(z >> 32);
}
__device__ __forceinline__ float2 dither_bm(unsigned salt, unsigned frame_pair, unsigned n) {
  const unsigned k1 = mix32((frame_pair * 1024u + n) ^ salt);
  const unsigned k2 = mix32(k1 ^ 0x9E3779B9u);
  const float u1 = ((k1 >> 8) + 1u) * (1.0f / 16777216.0f);  // (0, 1]
  const float u2 = (k2 >> 8) * (1.0f / 16777216.0f);         // [0, 1)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-2.0f * __logf(u1)));
  float sn, cs;
  __sincosf(6.28318530717958647692f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
// noise of frames (fa, fa + 1) at sample n, as the two lanes of an f2
__device__ __forceinline__ float2 dither_pair(unsigned salt, unsigned fa, unsigned n) {
  const float2 p0 = dither_bm(salt, fa >> 1, n);
  if ((fa & 1u) == 0u) return p0;                       // warp-uniform: fa is the same for the whole group
  const float2 p1 = dither_bm(salt, (fa >> 1) + 1u, n);
  return make_float2(p0.y, p1.x);
}

// ------------------------------------------------------------------------------------------------------------
// fbank_quad: log-mel of 4 consecutive frames (tile-local frames 4*quad .. 4*quad+3) by one warp.
//   x_base   staged raw samples, x_base[f*S + n] = sample n of tile-local frame f
//   win      this thread's window taps (n = 16*i + j), already scaled by 2^15 when upscaling
//   xg       this 16-thread group's transpose buffer; pbuf4 = the warp's spectra, [2 groups][256 bins] x (frame A, frame B),
//            aliasing the warp's two xg's
//   logmel   shared [F][M] destination
// NROWS = ceil(frame_len / 16): rows of 16 samples that can be non-zero (25 for 400-sample frames, else 32).
// EXACT: frame_len == 16*NROWS, so no per-sample bounds predicate is needed in the stage-1 load.
struct MelTab {
  const float2* w;
  const int* lo;
  int rounds;
  int cnt[kMelRounds];
  int base[kMelRounds];
#if B200FE_MEL_WORD_REGS
  unsigned word[3];   // fixed-shape banks: this lane's interval words, loaded once per kernel (mel_preload)
#endif
};

__device__ __forceinline__ void mel_preload(MelTab& mel, int lane) {
#if B200FE_MEL_WORD_REGS
#pragma unroll
  for (int r = 0; r < 3; ++r) mel.word[r] = (unsigned)__ldg(mel.lo + 32 * r + lane);
#endif
}

__device__ __forceinline__ float fast_ln(float x) {   // x is a normal positive number (>= the log floor)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.69314718055994530942f;
}

// Compile-time shape of the interval table (rounds and per-round trip counts).  MelShapeRuntime reads them from MelTab;
// MelShapeFixed lets the compiler unroll the whole mel stage and hoist its loads (the 80-mel / 512-point / 16 kHz bank
// of the Paraformer front-end is MelShapeFixed<3, 2, 5, 8>, <3, 2, 5, 9> without B200FE_MEL_COMPACT).
struct MelShapeRuntime { static constexpr bool kFixed = false; };
template <int R, int C0, int C1 = 0, int C2 = 0>
struct MelShapeFixed {
  static constexpr bool kFixed = true;
  static constexpr int kRounds = R;
  __host__ __device__ static constexpr int cnt(int r) { return r == 0 ? C0 : (r == 1 ? C1 : C2); }
  __host__ __device__ static constexpr int base(int r) { return r == 0 ? 0 : (r == 1 ? C0 : C0 + C1); }
};
#if B200FE_MEL_COMPACT   // the last filter's peak bin sits in the last interval there (build_interval_table): 8 trips, not 9
using MelShapeParaformer = MelShapeFixed<3, 2, 5, 8>;
#else
using MelShapeParaformer = MelShapeFixed<3, 2, 5, 9>;
#endif

// One round of the interval mel: lane <-> interval 31 r + lane.  CNT >= 0: compile-time trip count (fully unrolled).
// epi(iv, a, b, c, d) receives the natural-log mel energies of filter iv for the warp's 4 frames.
template <int CNT, bool COMPACT, class EPI>
__device__ __forceinline__ void mel_round(const MelTab& mel, int r, int cnt_rt, int base, const float2* pg, int lane, int M,
                                          float log_floor, EPI&& epi) {
  const int cnt = CNT >= 0 ? CNT : cnt_rt;
#if B200FE_MEL_WORD_REGS
  const unsigned word = CNT >= 0 ? mel.word[r < 3 ? r : 0] : (unsigned)__ldg(mel.lo + 32 * r + lane);
#else
  const unsigned word = (unsigned)__ldg(mel.lo + 32 * r + lane);   // run start | partner lane | filter | outputs
#endif
  const int lo = (int)(word & 0xfffu), partner = (int)((word >> 12) & 31u), iv = (int)((word >> 17) & 0xffu);
  const float2* wt = mel.w + (base * 32 + lane);
  const float2* p0 = pg + lo;
  // packed accumulators: (frame A, frame B) of group 0 and of group 1, as the spectra are stored
  f2 up0 = make_float2(0.f, 0.f), up1 = up0, dn0 = up0, dn1 = up0;
#if B200FE_MEL_COMPACT
  // compact table behind the (up, down) table (build_compact_mel, b200fe.cu): u = up-slope weight, 0 in padding slots;
  // down = scale - u where u > 0.  min(scale - u, u * 2^100) is that without a predicate (u * 2^100 >= 2^-2 for every
  // weight a bank can hold, 0 for padding).  Only instantiated for banks the host has checked to be encodable.
  const float* wu = reinterpret_cast<const float*>(mel.w + kMelSlots * 32) + (base * 32 + lane);
#endif
  auto body = [&](int q) {
#if B200FE_MEL_COMPACT
    float2 w;
    if constexpr (CNT >= 0 && COMPACT) {
      w.x = __ldg(wu + 32 * q);
      w.y = fminf(kMelScale - w.x, w.x * 0x1p100f);
    } else {
      w = __ldg(wt + 32 * q);
    }
#else
    const float2 w = __ldg(wt + 32 * q);
#endif
    const float2 s0 = p0[q], s1 = p0[kSpecPitch + q];
    up0 = fma2s(s0, w.x, up0);
    up1 = fma2s(s1, w.x, up1);
    dn0 = fma2s(s0, w.y, dn0);
    dn1 = fma2s(s1, w.y, dn1);
  };
  if constexpr (CNT >= 0) {
#pragma unroll
    for (int q = 0; q < CNT; ++q) body(q);
  } else {
#pragma unroll 4
    for (int q = 0; q < cnt; ++q) body(q);
  }
  // energy[filter iv] = up-slope sum of interval iv + down-slope sum of interval iv + 1 (held by the partner lane)
  const float ex = up0.x + __shfl_sync(0xffffffffu, dn0.x, partner);
  const float ey = up0.y + __shfl_sync(0xffffffffu, dn0.y, partner);
  const float ez = up1.x + __shfl_sync(0xffffffffu, dn1.x, partner);
  const float ew = up1.y + __shfl_sync(0xffffffffu, dn1.y, partner);
  if (word >> 31)
    epi(iv, fast_ln(fmaxf(ex, log_floor)), fast_ln(fmaxf(ey, log_floor)), fast_ln(fmaxf(ez, log_floor)),
        fast_ln(fmaxf(ew, log_floor)));
}

// Sparse mel + log for the warp's 4 frames.  lane <-> interval between two filter centres: every FFT bin lies in
// exactly one interval and feeds the up-slope of filter j and the down-slope of filter j-1, so each bin is read once:
// energy[m] = up[m] + down[m+1].  Trip counts are warp-uniform (zero-padded weights): no divergence.
// COMPACT = false keeps the (up, down) weight pairs also for a fixed-shape bank: the streaming tick is latency-bound and
// the three dependent operations that derive the second weight cost it more than the halved weight bytes save
// (measured: 512 streams 46.65 vs 47.2 us per tick).
template <class MELS, bool COMPACT = (B200FE_MEL_COMPACT != 0), class EPI>
__device__ __forceinline__ void mel_stage(const MelTab& mel, const float4* pbuf4, int lane, int M, float log_floor, EPI&& epi) {
  const float2* pg = reinterpret_cast<const float2*>(pbuf4);
  if constexpr (MELS::kFixed) {
    static_for<0, MELS::kRounds>([&](auto ic) {
      constexpr int r = decltype(ic)::value;
      mel_round<MELS::cnt(r), COMPACT>(mel, r, 0, MELS::base(r), pg, lane, M, log_floor, epi);
    });
  } else {
#pragma unroll 1
    for (int r = 0; r < mel.rounds; ++r) {
      int cnt = mel.cnt[0], base = mel.base[0];   // selected with static indices so that they stay warp-uniform
#pragma unroll
      for (int t = 1; t < kMelRounds; ++t)
        if (r == t) { cnt = mel.cnt[t]; base = mel.base[t]; }
      mel_round<-1, false>(mel, r, cnt, base, pg, lane, M, log_floor, epi);
    }
  }
}

// This thread's window taps: register i multiplies sample row i - g (g = 1 for the rotated second group of a warp).
template <int NROWS>
__device__ __forceinline__ void load_window_taps(float (&win)[NROWS + 1], const float* window512, int j, int grp_in_warp) {
  const int g = NROWS < 32 ? grp_in_warp : 0;
#pragma unroll
  for (int i = 0; i <= NROWS; ++i) {
    const int row = i - g;
    win[i] = (row >= 0 && row < NROWS) ? window512[16 * row + j] : 0.f;
  }
}

// This thread's row of the stage-2 twiddle table.  Stage 2 is assigned across the warp (quad_stage2): lanes 2c and 2c+1
// transform column c (16 for c = 0) of group 0 and of group 1 and read the same row.  One table serves both groups: the
// one-row rotation of the second group only multiplies its column k1 by the unit-modulus constant W32^k1, which the
// power spectrum does not see.
template <int NROWS>
__device__ __forceinline__ const float2* fft_twiddle_row(const float2* tw_s, int j, int grp_in_warp) {
  const int c = ((grp_in_warp << 4) | j) >> 1;
  const int col = c == 0 ? 16 : c;
  return tw_s + (col - 1) * kTwPitch;
}
// This thread's row of the column-0 table: lanes 2t and 2t+1 of a half-warp both compute bin 32 t of their own group.
__device__ __forceinline__ const float2* fft_c0_row(const float2* tw_s, int j) {
  return tw_s + kTw2Table + ((j >> 1) & 7) * kC0Pitch;
}
// B200FE_C0_SHFL: this lane's twiddles of the 16-point FFT across the lanes of its group (fill_c0_lane_table).
__device__ __forceinline__ const float2* fft_c0s_row(const float2* tw_s, int j) {
#if B200FE_C0_SHFL
  return tw_s + kTw2Table + j * kC0sPitch;
#else
  return fft_c0_row(tw_s, j);
#endif
}
// Host: entry [j][s], s = 0..2, of that table.  Decimation in frequency over the lane index with xor distances
// D = 8, 4, 2, 1: a lane whose bit D is clear keeps a + b, the other one (a - b) W_2D^(j mod D); stage 0 carries the
// factor 2 every stored spectrum value has.  Stage 3 has no twiddle.  After the four stages lane j holds bin
// 32 * bitrev4(j).
inline void fill_c0_lane_table(float2* tab /* [16 * kC0sPitch] */) {
  for (int j = 0; j < 16; ++j)
    for (int s = 0; s < 3; ++s) {
      const int D = 8 >> s;
      const bool upper = (j & D) != 0;
      const double ang = -2.0 * 3.14159265358979323846 * (double)(j & (D - 1)) / (double)(2 * D);
      const double sc = s == 0 ? 2.0 : 1.0;
      tab[j * kC0sPitch + s] = upper ? make_float2((float)(sc * cos(ang)), (float)(sc * sin(ang))) : make_float2((float)sc, 0.f);
    }
}

// ROT: the second 16-thread group of a warp loads its samples one 16-sample row late (register i holds row i-1).
// Frames start 160 samples = 5*32 banks apart, so without this both groups of a warp would hit the same 16 banks on
// every sample load (2-way conflict).  The rotation multiplies FFT32 output k1 - hence the whole column k1 of the
// spectrum - by W32^k1, a unit-modulus constant: the power spectrum is unchanged.
//
// Packed real FFT (tools/model_s2.py is the numpy model of this dataflow): lane .x of every f2 is frame A, .y frame B.
//   stage 1  y[i] = windowed sample 16(i-g)+j;  z[m] = y[2m] + i y[2m+1];  Z = FFT16(z);  split -> Y[k1], k1 = 0..16
//            (2 Y[k1] for k1 not in {0, 8, 16}; the twiddle rows of 8 and 16 and the column-0 table carry the 2).
//   transpose rows k1 = 1..16 as float4 (re A, re B, im A, im B), row 0 (real) as float2.
//   stage 2  thread j: column col = j (16 for thread 0):  X[col + 32 k2] = FFT16_c( Y_c[col] * tw[col][c] ), k2 = 0..15;
//            bins beyond 256 are the conjugates of bins (32 - col) + 32 (15 - k2) and have the same power.
//   column 0 thread t < 8: X[32 t] = sum_c Y_c[0] W16^(c t)  (real input: 8 folded terms).
// All stored powers are 4 |X[k]|^2 (the mel weights carry the 0.25).
// Stage 1 of a group's frame pair: load RAW samples, [dither], pre-emphasis with the frame-start rule, DC removal,
// window, real 32-point FFT.
//   xA   first raw sample of frame A in shared memory (x / 32768 scale or not: the window table carries the 2^15);
//        frame B starts S samples later.  The sample before a frame's first one is never read: the reference
//        replicates the first sample there (TA:193-198), and that is what thread 0 loads instead.
// Pre-emphasis  e[n] = x[n] - p x[n-1]  is one FMA per (frame, sample) on the two loaded values.  The frame mean is
// taken from the RAW samples and enters as (1-p)*mean inside the window FMA:
//   z[n] = w[n] * ((x[n] - mean) - p (x[n-1] - mean)) = e[n] * w[n] - ((1-p) mean) * w[n]          (TA:183-204)
// Both choices are about the weakest mel bins (FFT bins 1..2, 15 nepers below the rest after DC removal and
// pre-emphasis): an error of the subtracted constant is DC-coherent over the frame, i.e. it lands exactly there.
// Summing raw samples keeps that error (1-p) = 0.03 times smaller than recovering the mean from the sum of the
// pre-emphasised samples, and subtracting inside the FMA avoids rounding e[n] - const to e's grid (the same error
// on every sample of a binade).  profiles/r2_parity_report.txt has the before / after.
// Out: zr/zi[1..15] = 2 Y[k1], zr/zi[8] = Y[8], y0 = Y[0], y16 = Y[16] (real).
// SR > 0: the frame shift is SR 16-sample rows (compile-time), so frame B's row i is frame A's row i + SR and the
// overlapping rows are loaded once.
template <int NROWS, bool EXACT, bool DITHER, int SR = 0>
__device__ __forceinline__ void quad_stage1(const float* xA, bool vA, bool vB, int S, int L,
                                            const float (&win)[NROWS + 1], float preemph, int remove_dc, float dither,
                                            unsigned long long seed, unsigned utt, unsigned frame_abs_a, int j, int g,
                                            f2 (&zr)[16], f2 (&zi)[16], f2& y0, f2& y16) {
  constexpr bool ROT = NROWS < 32;
  constexpr int NR = ROT ? NROWS + 1 : NROWS;      // register rows in use
  constexpr int LIVE = (NR + 1) / 2;               // complex points z[m] that can be non-zero
  static_assert(LIVE == 16 || LIVE > 8, "frame rows must cover more than half of the FFT");
  {
    f2 y[2 * LIVE];
    f2 s;   // raw frame sums of this thread's samples
    // thread j owns samples n = 16*row + j of both frames; register i holds row i - g
    xA += j - 16 * g;
    auto row_in = [&](int i) {
      if constexpr (EXACT) return ROT ? (i == 0 ? g == 0 : (i == NROWS ? g == 1 : i < NROWS)) : true;
      else return (i - g >= 0) && (i < NR) && (16 * (i - g) + j < L);
    };
    const bool j0 = j == 0;
    if constexpr (SR > 0 && EXACT && ROT) {
      static_assert(SR < NROWS, "frames of a pair must overlap");
      constexpr int NREG = 2 * LIVE + SR;
      float r[NREG], pr[NREG];
      // Unconditional loads: registers outside a frame (the rotation's edge rows, rows of a frame beyond the quad's
      // last one) read whatever the buffer holds - finite stale samples - and only ever reach frames whose results are
      // discarded, or are masked below (row_in).  No predicates, no zero-initialised registers.
#pragma unroll
      for (int i = 0; i < NREG; ++i) {
        const bool first = j0 && i == g;           // frame A's n = 0: its predecessor is the sample itself
#if B200FE_S1_UNCOND
        r[i] = xA[16 * i];
        pr[i] = xA[16 * i - (first ? 0 : 1)];
#else
        const bool need = (i < 2 * LIVE && vA && row_in(i)) || (i >= SR && vB && row_in(i - SR));
        r[i] = need ? xA[16 * i] : 0.f;
        pr[i] = need ? xA[16 * i - (first ? 0 : 1)] : 0.f;
#endif
      }
#pragma unroll
      for (int i = 0; i < 2 * LIVE; ++i) {
        const float pb = (j0 && i == g) ? r[i + SR] : pr[i + SR];   // frame B's n = 0
        y[i].x = row_in(i) ? fmaf(-preemph, pr[i], r[i]) : 0.f;
        y[i].y = row_in(i) ? fmaf(-preemph, pb, r[i + SR]) : 0.f;
      }
      // raw sums: frame A = registers g .. g+NROWS-1, frame B = the same + SR; registers SR+1 .. NROWS-1 are common
      float c4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = SR + 1; k < NROWS; ++k) c4[(k - SR - 1) & 3] += r[k];
      float a2[2] = {g == 0 ? r[0] : r[NROWS], 0.f}, b2[2] = {g == 0 ? r[SR] : r[NROWS + SR], 0.f};
#pragma unroll
      for (int k = 1; k <= SR; ++k) a2[k & 1] += r[k];
#pragma unroll
      for (int k = NROWS; k < NROWS + SR; ++k) b2[k & 1] += r[k];
      const float common = (c4[0] + c4[1]) + (c4[2] + c4[3]);
      s = make_float2(common + (a2[0] + a2[1]), common + (b2[0] + b2[1]));
    } else {
      const float* xB = xA + S;
      f2 s4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) s4[i] = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 2 * LIVE; ++i) {
        const bool in = row_in(i);
        const int back = (j0 && i == g) ? 0 : 1;
        f2 x = make_float2(0.f, 0.f), px = make_float2(0.f, 0.f);
        if (vA && in) { x.x = xA[16 * i]; px.x = xA[16 * i - back]; }
        if (vB && in) { x.y = xB[16 * i]; px.y = xB[16 * i - back]; }
        y[i] = fma2s(px, -preemph, x);
        s4[i & 3] = add2(s4[i & 3], x);
      }
      s = add2(add2(s4[0], s4[1]), add2(s4[2], s4[3]));
    }
    if constexpr (DITHER) {
      // x'[n] = x[n] + dither*g(frame, n)  (TA:179-181): the pre-emphasised value picks up
      // dither*(g(n) - preemph*g(n-1)) and the raw sum dither*g(n).  Every thread draws the noise of its own samples
      // once; g(n-1) is the previous lane's value of the same row (lane 15's value of the previous row for lane 0),
      // fetched with one shuffle.
      const unsigned fa = frame_abs_a;
      const unsigned salt = dither_salt(seed, utt);
      const int src = (threadIdx.x & 16) | ((j + 15) & 15);
      f2 rot_prev = make_float2(0.f, 0.f), gs = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int n = 16 * (i - g) + j;
        const bool live = n >= 0 && n < L;
        f2 gn = make_float2(0.f, 0.f);
        if (live) gn = dither_pair(salt, fa, (unsigned)n);
        f2 rot;
        rot.x = __shfl_sync(0xffffffffu, gn.x, src);
        rot.y = __shfl_sync(0xffffffffu, gn.y, src);
        const f2 prev = n == 0 ? gn : (j == 0 ? rot_prev : rot);   // n = 0: replicate rule, g(-1) := g(0)
        rot_prev = rot;
        if (live) {
          y[i] = fma2s(fma2s(prev, -preemph, gn), dither, y[i]);
          gs = add2(gs, gn);
        }
      }
      s = fma2s(gs, dither, s);
    }
#if B200FE_DC_LATE
#pragma unroll
    for (int i = 0; i < NR; ++i) y[i] = mul2s(y[i], win[i]);
#endif
#if B200FE_DC_REDUCE4
#pragma unroll
    for (int lvl = 0; lvl < 2; ++lvl) {
      const int o = lvl ? 4 : 1;
      f2 t1, t2, t3;
      t1.x = __shfl_xor_sync(0xffffffffu, s.x, o);     t1.y = __shfl_xor_sync(0xffffffffu, s.y, o);
      t2.x = __shfl_xor_sync(0xffffffffu, s.x, 2 * o); t2.y = __shfl_xor_sync(0xffffffffu, s.y, 2 * o);
      t3.x = __shfl_xor_sync(0xffffffffu, s.x, 3 * o); t3.y = __shfl_xor_sync(0xffffffffu, s.y, 3 * o);
      s = add2(add2(s, t1), add2(t2, t3));
    }
#else
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
      f2 t;
      t.x = __shfl_xor_sync(0xffffffffu, s.x, o);
      t.y = __shfl_xor_sync(0xffffffffu, s.y, o);
      s = add2(s, t);
    }
#endif
    // -(1-p) * mean(frame), applied inside the window FMA: e*w - ((1-p) mean)*w rounds once, after the subtraction
    f2 nmean = make_float2(0.f, 0.f);
    if (remove_dc) nmean = mul2s(s, -(1.0f - preemph) / (float)L);
#pragma unroll
#if B200FE_DC_LATE
    for (int i = 0; i < NR; ++i) y[i] = fma2s(nmean, win[i], y[i]);
#else
    for (int i = 0; i < NR; ++i) y[i] = fma2s(y[i], win[i], mul2s(nmean, win[i]));
#endif
    // z[m] = y[2m] + i y[2m+1] at bit-reversed positions (decimation in time)
    static_for<0, 16>([&](auto ic) {
      constexpr int m = decltype(ic)::value;
      if constexpr (m < LIVE) {
        zr[bitrev<16>(m)] = y[2 * m];
        zi[bitrev<16>(m)] = (2 * m + 1 < NR) ? y[2 * m + 1] : make_float2(0.f, 0.f);
      } else {
        zr[bitrev<16>(m)] = make_float2(0.f, 0.f);
        zi[bitrev<16>(m)] = make_float2(0.f, 0.f);
      }
    });
  }
  fft_dit2<16, LIVE>(zr, zi);
  real32_split2(zr, zi, y0, y16);
}

// Stage 2: transpose through the group's buffer, twiddle, complex 16-point FFT of this thread's column, column 0,
// and the power spectra of the group's two frames into pbuf4 = [2 groups][256 bins] x (frame A, frame B), which
// aliases both groups' transpose buffers.  The caller has passed a __syncwarp since the last reader of that memory;
// on return every lane has passed a __syncwarp after the last write.
// C0LANE: column 0 by the FFT across the lanes of a group (c0_row = fft_c0s_row of a table filled by
// fill_c0_lane_table); false: by broadcast row loads through shared memory (c0_row = fft_c0_row of the [8][kC0Pitch]
// table).  The streaming tick keeps the latter: its CTAs walk through the quad's phases together, so the four dependent
// shuffle stages sit on its critical path (measured: 512 streams 46.5 vs 47.2 us per tick).
template <bool C0LANE = (B200FE_C0_SHFL != 0)>
__device__ __forceinline__ void quad_stage2(const f2 (&zr)[16], const f2 (&zi)[16], f2 y0, f2 y16, float4* yg, float4* pbuf4,
                                            const float2* tw_row, const float2* c0_row, int j, int grp_in_warp) {
  // 64-bit stores (the register allocator does not form the aligned quads a 128-bit store needs).  Slots 8..15 hold
  // (im, re) instead of (re, im), so the 16 lanes of a group always write 32 distinct banks.
  {
    float2* yg2 = reinterpret_cast<float2*>(yg) + 2 * j;
    const int hr = j >> 3, hi = hr ^ 1;
    static_for<1, 16>([&](auto ic) {
      constexpr int k1 = decltype(ic)::value;
      yg2[(k1 - 1) * 2 * kYPitch + hr] = zr[k1];
      yg2[(k1 - 1) * 2 * kYPitch + hi] = zi[k1];
    });
    yg2[15 * 2 * kYPitch + hr] = y16;
    yg2[15 * 2 * kYPitch + hi] = make_float2(0.f, 0.f);
  }
  if constexpr (!C0LANE) reinterpret_cast<float2*>(yg + 16 * kYPitch)[j] = y0;
#if B200FE_S2_TW_EARLY
  float2 tw_e[16];
  float4 c0w_e[4];
#pragma unroll
  for (int h = 0; h < 16; ++h) tw_e[h] = tw_row[h];
  if constexpr (!C0LANE) {
#pragma unroll
    for (int h = 0; h < 4; ++h) c0w_e[h] = reinterpret_cast<const float4*>(c0_row)[h];
  }
#endif
  __syncwarp();

  // Stage 2 is assigned across the warp, not per group: lanes 2c and 2c+1 transform column c (16 for c = 0) of group 0
  // and of group 1.  Both read the same twiddle row, and a shared load merges identical addresses of adjacent lanes
  // (profiles/r1_microbench_smem_wavefronts.txt), so the tables are read once per warp instead of once per group.
  const int lane = (grp_in_warp << 4) | j;
  const int c = lane >> 1, g2 = lane & 1;
  const int col = c == 0 ? 16 : c;
  float4* const ywarp = yg - grp_in_warp * kYGroupF4;
  f2 ar[16], ai[16];
  {
    const float4* rowp = ywarp + g2 * kYGroupF4 + (col - 1) * kYPitch;
    static_for<0, 8>([&](auto ic) {
      constexpr int h = decltype(ic)::value;
#if B200FE_S2_TW_EARLY
      const float2 ta = tw_e[2 * h], tb = tw_e[2 * h + 1];
#else
      const float2 ta = tw_row[2 * h], tb = tw_row[2 * h + 1];
#endif
      const float4 t = make_float4(ta.x, ta.y, tb.x, tb.y);
      const float4 v0 = rowp[2 * h], v1 = rowp[2 * h + 1];
      constexpr bool sw = 2 * h >= 8;   // slots 8..15 are stored (im, re)
      {
        const f2 yr = sw ? make_float2(v0.z, v0.w) : make_float2(v0.x, v0.y);
        const f2 yi = sw ? make_float2(v0.x, v0.y) : make_float2(v0.z, v0.w);
        ar[bitrev<16>(2 * h)] = fma2s(yr, t.x, neg2(mul2s(yi, t.y)));
        ai[bitrev<16>(2 * h)] = fma2s(yr, t.y, mul2s(yi, t.x));
      }
      {
        const f2 yr = sw ? make_float2(v1.z, v1.w) : make_float2(v1.x, v1.y);
        const f2 yi = sw ? make_float2(v1.x, v1.y) : make_float2(v1.z, v1.w);
        ar[bitrev<16>(2 * h + 1)] = fma2s(yr, t.z, neg2(mul2s(yi, t.w)));
        ai[bitrev<16>(2 * h + 1)] = fma2s(yr, t.w, mul2s(yi, t.z));
      }
    });
  }
  // ---- column 0 of the lane's OWN group (half-warp): lanes 2t and 2t+1 both sum the 16 real bins Y_c[0] against
  //      W16^(c t), folded to 8 terms (the pair's identical loads merge; the even lane stores)
  int t0;
  f2 p0;
  if constexpr (C0LANE) {
  // ---- column 0 of the lane's OWN group: X[32 t] = 2 sum_j Y_j[0] W16^(j t) is a 16-point DFT of the REAL values the
  //      16 lanes hold in y0: four butterfly stages over xor-shuffles (2 + 4 + 4 + 4 shuffles for the packed pair of
  //      frames), the lane's three twiddles from the table, signs from the lane's bits.  Lane j ends up with bin
  //      32 * bitrev4(j); bins 0..224 sit on the even lanes.
  t0 = (int)(__brev((unsigned)j) >> 28);      // bitrev4(j): below 8 on the even lanes, the ones that store
  {
    const unsigned jb = (unsigned)j << 28;               // bit 3 of j in the sign position
    auto sgn = [&](int s) { return __uint_as_float(0x3f800000u | ((jb << s) & 0x80000000u)); };
    auto xchg = [&](const f2& v, int d) {
      return make_float2(__shfl_xor_sync(0xffffffffu, v.x, d), __shfl_xor_sync(0xffffffffu, v.y, d));
    };
    const float2 w0 = c0_row[0], w1 = c0_row[1], w2 = c0_row[2];
    f2 d = fma2s(y0, sgn(0), xchg(y0, 8));
    f2 cr = mul2s(d, w0.x), ci = mul2s(d, w0.y);
    {
      const f2 dr = fma2s(cr, sgn(1), xchg(cr, 4)), di = fma2s(ci, sgn(1), xchg(ci, 4));
      cr = fma2s(dr, w1.x, neg2(mul2s(di, w1.y)));
      ci = fma2s(dr, w1.y, mul2s(di, w1.x));
    }
    {
      const f2 dr = fma2s(cr, sgn(2), xchg(cr, 2)), di = fma2s(ci, sgn(2), xchg(ci, 2));
      cr = fma2s(dr, w2.x, neg2(mul2s(di, w2.y)));
      ci = fma2s(dr, w2.y, mul2s(di, w2.x));
    }
    cr = fma2s(cr, sgn(3), xchg(cr, 1));
    ci = fma2s(ci, sgn(3), xchg(ci, 1));
    p0 = fma2(cr, cr, mul2(ci, ci));
  }
  } else {
  t0 = (lane >> 1) & 7;
  {
    const float4* u4 = reinterpret_cast<const float4*>(yg + 16 * kYPitch);   // 16 x (A, B)
    const float4* w4 = reinterpret_cast<const float4*>(c0_row);   // fft_c0_row: row t0
    const float sgn = (t0 & 1) ? -1.f : 1.f;
    f2 cr = make_float2(0.f, 0.f), ci = make_float2(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#if B200FE_S2_TW_EARLY
      const float4 ua = u4[h], ub = u4[h + 4], w = c0w_e[h];
#else
      const float4 ua = u4[h], ub = u4[h + 4], w = w4[h];
#endif
      const f2 v0 = fma2s(make_float2(ub.x, ub.y), sgn, make_float2(ua.x, ua.y));
      const f2 v1 = fma2s(make_float2(ub.z, ub.w), sgn, make_float2(ua.z, ua.w));
      cr = fma2s(v0, w.x, cr);
      ci = fma2s(v0, w.y, ci);
      cr = fma2s(v1, w.z, cr);
      ci = fma2s(v1, w.w, ci);
    }
    p0 = fma2(cr, cr, mul2(ci, ci));
  }
  }
  fft_dit2<16>(ar, ai);
  __syncwarp();   // every lane has consumed the transpose buffers: the spectra may overwrite them
  {
    float2* pb2 = reinterpret_cast<float2*>(pbuf4) + g2 * kSpecPitch;
    static_for<0, 16>([&](auto ic) {
      constexpr int k2 = decltype(ic)::value;
      const f2 pw = fma2(ar[k2], ar[k2], mul2(ai[k2], ai[k2]));
      if constexpr (k2 < 8) pb2[col + 32 * k2] = pw;
      else if (c != 0) pb2[(32 - col) + 32 * (15 - k2)] = pw;
    });
    // bin 0 carries no mel weight
    if ((lane & 1) == 0) reinterpret_cast<float2*>(pbuf4)[grp_in_warp * kSpecPitch + 32 * t0] = p0;
  }
  __syncwarp();
}

// Log-mel of tile-local frames 4*quad .. 4*quad+3 into logmel[F][M] (shared memory): tile and streaming kernels.
template <int NROWS, bool EXACT, bool DITHER, class MELS, bool LANE_OPT = true>
__device__ __forceinline__ void fbank_quad(const float* x_base, int F, int quad,
                                           int S, int L, const float (&win)[NROWS + 1], float4* yg, float4* pbuf4,
                                           const float2* tw_row, const float2* c0_row, const MelTab& mel, int M,
                                           float preemph, int remove_dc, float log_floor, float dither,
                                           unsigned long long seed, unsigned utt, unsigned frame_abs0, float* logmel,
                                           int j, int grp_in_warp, int lane) {
  const int g = NROWS < 32 ? grp_in_warp : 0;
  const int fA = 4 * quad + 2 * grp_in_warp;   // tile-local frame of lane .x; fA + 1 is lane .y
  const bool vA = fA < F, vB = fA + 1 < F;
  f2 zr[16], zi[16], y0, y16;
  quad_stage1<NROWS, EXACT, DITHER>(x_base + fA * S, vA, vB, S, L, win, preemph, remove_dc, dither, seed, utt,
                                    frame_abs0 + (unsigned)fA, j, g, zr, zi, y0, y16);
  __syncwarp();   // earlier readers of the (aliased) buffer are done
  quad_stage2<LANE_OPT && (B200FE_C0_SHFL != 0)>(zr, zi, y0, y16, yg, pbuf4, tw_row, c0_row, j, grp_in_warp);
  const int fr = 4 * quad;
  mel_stage<MELS, LANE_OPT && (B200FE_MEL_COMPACT != 0)>(mel, pbuf4, lane, M, log_floor, [&](int iv, float a, float b, float c, float d) {
    float* dst = logmel + fr * M + iv;
    if (fr < F) dst[0] = a;
    if (fr + 1 < F) dst[M] = b;
    if (fr + 2 < F) dst[2 * M] = c;
    if (fr + 3 < F) dst[3 * M] = d;
  });
}

template <bool STATS>
struct StatsAcc {
  double sum[2][4], sq[2][4];
};
template <>
struct StatsAcc<false> {};

template <int NROWS, bool EXACT, bool DITHER, bool STATS, class MELS>
__global__ void __launch_bounds__(kCtaThreads, 3)
fbank_lfr_cmvn_tile_kernel(const TileParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* e_s = reinterpret_cast<float*>(smem_raw);              // the tile's raw samples
  float4* xbuf = reinterpret_cast<float4*>(e_s + p.e_cap);
  float* logmel_s = reinterpret_cast<float*>(xbuf + kWarps * kYWarpF4);
  float2* tw_s = reinterpret_cast<float2*>(logmel_s + kFMax * p.n_mels);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int j = tid & (kGroup - 1);     // position inside the 16-thread group: n2 in stage 1, column id in stage 2
  const int grp_in_warp = lane >> 4;    // 0 / 1
  const int L = p.frame_len, S = p.frame_shift, M = p.n_mels;
  const int D = p.lfr_m * M;
  const int lfr_left = (p.lfr_m - 1) / 2;

  // ---- per-CTA constants: tables to shared memory, this thread's window taps to registers
  for (int i = tid; i < kTw2Total; i += kCtaThreads) tw_s[i] = p.twiddle[i];
  MelTab mel;
  mel.w = p.mel_w; mel.lo = p.mel_lo; mel.rounds = p.mel_rounds;
  mel_preload(mel, threadIdx.x & 31);
#pragma unroll
  for (int r = 0; r < kMelRounds; ++r) { mel.cnt[r] = p.mel_cnt[r]; mel.base[r] = p.mel_base[r]; }
  float win[NROWS + 1];
  load_window_taps<NROWS>(win, p.window, j, grp_in_warp);

  // output columns handled by this thread in the LFR/CMVN phase (float4 granularity), fixed for the CTA's lifetime
  const int D4 = D >> 2, M4 = M >> 2;
  float4 cm_shift[2], cm_scale[2];
  int col_j[2], col_d[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int c4 = tid + c * kCtaThreads;
    cm_shift[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    cm_scale[c] = make_float4(1.f, 1.f, 1.f, 1.f);
    col_j[c] = 0;
    col_d[c] = 0;
    if (c4 < D4) {
      col_j[c] = c4 / M4;
      col_d[c] = c4 - col_j[c] * M4;
      if (p.cmvn) {
        cm_shift[c] = *reinterpret_cast<const float4*>(p.cmvn + 4 * c4);
        cm_scale[c] = *reinterpret_cast<const float4*>(p.cmvn + D + 4 * c4);
      }
    }
  }
  StatsAcc<STATS> st;
  if constexpr (STATS) {
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) st.sum[c][k] = st.sq[c][k] = 0.0;
  }
  __syncthreads();

  const unsigned wave_mis = (unsigned)((reinterpret_cast<uintptr_t>(p.wave) >> 2) & 3);
  float4* yg = xbuf + warp * kYWarpF4 + grp_in_warp * kYGroupF4;   // this group's transpose buffer
  float4* pbuf4 = xbuf + warp * kYWarpF4;                          // this warp's power spectra (aliases both groups)
  const float2* tw_row = fft_twiddle_row<NROWS>(tw_s, j, grp_in_warp);
  const float2* c0_row = fft_c0s_row(tw_s, j);

    // Work distribution: the first tile of every CTA is static, later ones are claimed from a global counter (thread 0,
  // published through shared memory across the barrier that follows the staging).
  __shared__ int next_tile_s;
  TileDesc cur;
  int tile = blockIdx.x;
  if (tile < p.n_tiles) cur = p.tiles[tile];
  while (tile < p.n_tiles) {
    if (tid == 0) next_tile_s = (int)gridDim.x + atomicAdd(p.next_tile, 1);
    const int utt = cur.utt, row0 = cur.row0, nrow = cur.nrow, T = cur.T, f_lo = cur.f_lo, F = cur.F;
    const long long g0 = cur.g0;
    const int a_off = (int)((wave_mis + (unsigned)(g0 & 3)) & 3);
    const long long ga = g0 - a_off;                          // 16-byte aligned load grid
    const int n_s = (F - 1) * S + L;

    // ---- stage: HBM -> shared, 128-bit both ways (raw samples: pre-emphasis happens in stage 1).  All loads of a
    //      batch are issued before the first use, so a tile pays one memory latency instead of one per iteration.
    {
      const int nv = (a_off + n_s + 3) >> 2;
      const bool interior = ga >= 0 && ga + 4ll * nv <= p.wave_total;   // no per-element bounds checks needed
      const float* src = p.wave + ga;
      constexpr int kBatch = 11;   // 11 x 128 float4 cover a 400/160 tile in one batch: one exposed memory latency
      for (int vb = 0; vb < nv; vb += kBatch * kCtaThreads) {
        float4 x[kBatch];
        if (interior) {
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int v = vb + u * kCtaThreads + tid;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nv) x[u] = ldg_stream4(src + 4 * v);
          }
        } else {
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int v = vb + u * kCtaThreads + tid;
            const long long ab = ga + 4ll * v;
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < nv) {
              if (ab >= 0 && ab < p.wave_total) x[u].x = p.wave[ab];
              if (ab + 1 >= 0 && ab + 1 < p.wave_total) x[u].y = p.wave[ab + 1];
              if (ab + 2 >= 0 && ab + 2 < p.wave_total) x[u].z = p.wave[ab + 2];
              if (ab + 3 >= 0 && ab + 3 < p.wave_total) x[u].w = p.wave[ab + 3];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int v = vb + u * kCtaThreads + tid;
          if (v < nv) *reinterpret_cast<float4*>(e_s + 4 * v) = x[u];
        }
      }
    }
    __syncthreads();
    const int tile_next = next_tile_s;   // rewritten by thread 0 only after the next barrier
    TileDesc nxt = cur;
    if (tile_next < p.n_tiles) nxt = p.tiles[tile_next];

    // ---- pull the next tile's samples into L2 while this tile computes (one prefetch per 128-byte line)
    if (tile_next < p.n_tiles) {
      const long long nb = nxt.g0 & ~31ll;
      const int lines = (((nxt.F - 1) * S + L + 32 + 31) >> 5);
      for (int ln = tid; ln < lines; ln += kCtaThreads) {
        const long long idx = nb + 32ll * ln;
        if (idx >= 0 && idx < p.wave_total) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.wave + idx));
      }
    }

    // ---- per warp: quads of 4 frames (2 groups x 2 frames), no CTA-wide sync inside
    for (int quad = warp; 4 * quad < F; quad += kWarps)
      fbank_quad<NROWS, EXACT, DITHER, MELS>(e_s + a_off, F, quad, S, L, win, yg, pbuf4, tw_row, c0_row, mel, M,
                                       p.preemph, p.remove_dc, p.log_floor, p.dither, p.seed, (unsigned)utt,
                                       (unsigned)f_lo, logmel_s, j, grp_in_warp, lane);
    __syncthreads();

    // ---- LFR stack + CMVN, written straight into the padded [batch, rows_cap, D] output
    {
      float* out_u = p.feats + ((long long)utt * p.rows_cap + row0) * D;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c4 = tid + c * kCtaThreads;
        if (c4 < D4) {
          for (int r = 0; r < nrow; ++r) {
            int f = p.lfr_n * (row0 + r) + col_j[c] - lfr_left;
            f = min(max(f, 0), T - 1) - f_lo;
            const float4 v = *reinterpret_cast<const float4*>(logmel_s + f * M + 4 * col_d[c]);
            if constexpr (STATS) {
              st.sum[c][0] += v.x; st.sum[c][1] += v.y; st.sum[c][2] += v.z; st.sum[c][3] += v.w;
              st.sq[c][0] += (double)v.x * v.x; st.sq[c][1] += (double)v.y * v.y;
              st.sq[c][2] += (double)v.z * v.z; st.sq[c][3] += (double)v.w * v.w;
            }
            float4 o;
            o.x = (v.x + cm_shift[c].x) * cm_scale[c].x;
            o.y = (v.y + cm_shift[c].y) * cm_scale[c].y;
            o.z = (v.z + cm_shift[c].z) * cm_scale[c].z;
            o.w = (v.w + cm_shift[c].w) * cm_scale[c].w;
            stg_stream4(out_u + (long long)r * D + 4 * c4, o);
          }
        }
      }
    }
    cur = nxt;
    tile = tile_next;
    // no barrier here: the next tile's staging only writes e_s, whose last readers finished before the
    // barrier above, and logmel_s is not written again before the barrier that follows the next staging.
  }

  if constexpr (STATS) {
    if (p.stats) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c4 = tid + c * kCtaThreads;
        if (c4 < D4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            atomicAdd(p.stats + 4 * c4 + k, st.sum[c][k]);
            atomicAdd(p.stats + D + 4 * c4 + k, st.sq[c][k]);
          }
        }
      }
    }
  }
}

}  // namespace b200fe
