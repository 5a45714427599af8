// libb200fe.so - host side of the C ABI declared in include/b200fe.h.
//
// No torch types here: plain pointers, sizes and a cudaStream_t.  There is no CPU fallback: every entry point that
// computes launches sm_100a kernels, and create() fails when no CUDA device is present.
#include "../../include/b200fe.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "extras.cuh"
#include "resample_fft.cuh"
#include "fbank_tile.cuh"
#include "fbank_warp.cuh"
#include "stream_kernel.cuh"
#include "tts_mel.cuh"

using namespace b200fe;

namespace {

thread_local std::string g_create_error;

struct PinnedSlot {
  void* ptr = nullptr;
  size_t cap = 0;
  cudaEvent_t ev = nullptr;
  bool busy = false;
};

}  // namespace

struct b200fe_handle {
  b200fe_config cfg;
  int L = 0, S = 0, nfft = 0, D = 0, rows_per_tile = 0, e_cap = 0;
  size_t smem_bytes = 0;
  int n_sms = 0;
  bool has_cmvn = false;
  std::vector<float> window_host;   // [L], without upscale
  std::vector<float> mel_host;      // [n_mels, nfft/2]
  float* d_window = nullptr;        // [512]
  float2* d_twiddle = nullptr;      // [kTw2Total]
  float2* d_mel_w = nullptr;        // [kMelSlots * 32] lane-transposed (up, down) weights
  int* d_mel_lo = nullptr;          // [32 * kMelRounds] first bin of every interval's padded run
  int mel_rounds = 0, mel_cnt[kMelRounds] = {0}, mel_base[kMelRounds] = {0};
  bool mel_paraformer = false;      // table shape == MelShapeParaformer: fully unrolled mel stage
  float* d_cmvn = nullptr;          // [2*D]
  float2* d_twiddle_stream = nullptr;   // the same twiddles with the [8][kC0Pitch] column-0 table (stream kernels)
  float4* d_cmvn_il = nullptr;      // warp kernel: [32 slots][n_mels/4][shift4, scale4], identity without CMVN (fbank_warp.cuh)
  // dense mel banks for shrunken frames (VF:147), keyed by fft size
  std::map<int, int> short_mel_off;
  float* d_short_mel = nullptr;
  size_t short_mel_floats = 0;
  std::mutex mu;
  PinnedSlot slots[4];
  int next_slot = 0;
  long long launches = 0;
  int profile_every = 0, profile_tick = 0;   // b200fe_profile_enable: time every n-th fused-kernel launch
  bool force_tile = false;           // b200fe_select_kernel(h, 1): keep the tile kernel (A/B measurements, tests)
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  std::vector<cudaEvent_t> event_pool;   // recycled by b200fe_profile_collect
  UttTable utt_tab;                  // staging for the prep launch's by-value utterance table (under mu)
  mutable std::string err;
};

namespace {

int fail(const b200fe_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

#define CUDA_TRY(h, expr)                                                                      \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(h, B200FE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
  } while (0)

int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return x == 0 ? 1 : p;
}

// TA:86-113, evaluated in double and rounded once
void build_window(int type, int n, double blackman, std::vector<float>& w) {
  w.resize(n);
  const double a = 2.0 * M_PI / (n - 1);
  for (int i = 0; i < n; ++i) {
    double v;
    switch (type) {
      case B200FE_WIN_HAMMING: v = 0.54 - 0.46 * cos(a * i); break;
      case B200FE_WIN_HANNING: v = 0.5 - 0.5 * cos(a * i); break;
      case B200FE_WIN_POVEY: v = pow((double)(float)(0.5 - 0.5 * cos(a * i)), 0.85); break;
      case B200FE_WIN_RECTANGULAR: v = 1.0; break;
      default: v = blackman - 0.5 * cos(a * i) + (0.5 - blackman) * cos(2 * a * i); break;
    }
    w[i] = (float)v;
  }
}

// TA:436-511 (vtln_warp = 1): triangular filters in the mel domain, [n_mels, nfft/2]
int build_mel(int n_mels, int nfft, double fs, double low, double high, std::vector<float>& bank) {
  const int nb = nfft / 2;
  const double nyq = 0.5 * fs;
  if (high <= 0.0) high += nyq;
  if (!(0.0 <= low && low < nyq && 0.0 < high && high <= nyq && low < high)) return -1;
  const double bw = fs / nfft;
  const double mlo = 1127.0 * log(1.0 + low / 700.0), mhi = 1127.0 * log(1.0 + high / 700.0);
  const double delta = (mhi - mlo) / (n_mels + 1);
  bank.assign((size_t)n_mels * nb, 0.f);
  for (int m = 0; m < n_mels; ++m) {
    const double left = mlo + m * delta, center = mlo + (m + 1.0) * delta, right = mlo + (m + 2.0) * delta;
    for (int k = 0; k < nb; ++k) {
      const double mel = 1127.0 * log(1.0 + bw * k / 700.0);
      const double up = (mel - left) / (center - left), down = (right - mel) / (right - center);
      const double v = fmax(0.0, fmin(up, down));
      bank[(size_t)m * nb + k] = (float)v;
    }
  }
  return 0;
}

// Sparse mel bank by interval between filter centres (see TileParams::mel_w).  `bank` is dense [nm, ld] (row stride ld),
// only columns < nb are used.  A bin must feed at most two adjacent filters (triangular banks).
//
// Round r holds intervals 31 r .. 31 r + 31 (the last one belongs to the next round too: here it only supplies the
// down-slope sum of filter 31 r + 30).  Which LANE an interval sits on is free, and so is the start of its padded run
// inside the slack the round's trip count leaves: with bank_mod = 16 both are chosen (bipartite matching) so that the 16
// lanes of each half-warp start on 16 different residues mod 16, which makes the 64-bit gathers of the mel stage
// bank-conflict free for every step (all lanes advance together).  bank_mod = 0 keeps lane l on interval 31 r + l.
// mlo word: bits 0..11 first bin of the run, 12..16 lane that holds the NEXT interval (its down-slope sum completes
// this lane's filter), 17..24 filter index, bit 31 = this lane outputs a filter.
int build_interval_table(const std::vector<float>& bank, int nm, int nb, int ld, float scale, std::vector<float2>& mw,
                         std::vector<int>& mlo, int& rounds, int* cnt_out, int* base_out, std::string& why,
                         int bank_mod = 0, bool last_peak_down = false) {
  mw.assign((size_t)kMelSlots * 32, make_float2(0.f, 0.f));
  mlo.assign(32 * kMelRounds, 1);
  auto W = [&](int m, int k) { return bank[(size_t)m * ld + k]; };
  std::vector<int> iv_of(nb, -1);
  int prev = 0;
  for (int k = 0; k < nb; ++k) {
    int first = -1, last = -1, nz = 0;
    for (int m = 0; m < nm; ++m)
      if (W(m, k) > 0.f) { if (first < 0) first = m; last = m; ++nz; }
    if (nz == 0) continue;
    if (k == 0) { why = "FFT bin 0 must not carry mel weight"; return -1; }
    if (nz > 2 || last - first > 1) { why = "mel filterbank is not a 2-banded triangle bank"; return -1; }
    // two weights: up-slope of `last`, down-slope of `first` -> interval `last`.  One weight (interval 0, the last
    // interval, or a bin exactly on a centre): it is an up-slope weight up to the filter's peak, a down-slope weight
    // after it; on the peak either neighbour interval yields the same sum, so keep the run monotone.
    int iv;
    if (nz == 2) iv = last;
    else {
      int peak = 0;
      for (int kk = 1; kk < nb; ++kk)
        if (W(first, kk) > W(first, peak)) peak = kk;
      iv = k > peak ? first + 1 : first;
      // last_peak_down: the peak bin of the LAST filter, when it carries that one weight only, goes to the last interval
      // as a down-slope weight (same sum); the compact table (build_compact_mel) cannot hold a lone up-slope weight on
      // a lane whose down-slope sum is read
      if (last_peak_down && first == nm - 1 && k == peak) iv = first + 1;
      if (iv < prev && prev <= first + 1) iv = prev;
    }
    if (iv < prev) { why = "mel filterbank intervals are not monotone"; return -1; }
    iv_of[k] = iv;
    prev = iv;
  }
  std::vector<int> ilo(nm + 1, -1), icnt(nm + 1, 0);
  for (int k = 0; k < nb; ++k) {
    if (iv_of[k] < 0) continue;
    const int iv = iv_of[k];
    if (ilo[iv] < 0) ilo[iv] = k;
    if (k != ilo[iv] + icnt[iv]) { why = "mel filterbank intervals are not contiguous"; return -1; }
    icnt[iv]++;
  }
  if (nb > 4096 || nm > 255) { why = "mel filterbank does not fit the packed interval word"; return -1; }
  rounds = (nm + 30) / 31;
  int base = 0;
  for (int r = 0; r < rounds; ++r) {
    int c = 0;
    for (int l = 0; l < 32; ++l) {
      const int iv = 31 * r + l;
      if (iv <= nm && icnt[iv] > c) c = icnt[iv];
    }
    cnt_out[r] = c;
    base_out[r] = base;
    if (base + c > kMelSlots) { why = "mel filterbank does not fit the sparse layout"; return -1; }
    // items of the round and the run starts each may take
    int n_items = 0, item_iv[32], lo_min[32], lo_max[32];
    for (int l = 0; l < 32; ++l) {
      const int iv = 31 * r + l;
      if (iv > nm) break;
      item_iv[n_items] = iv;
      if (icnt[iv]) {
        lo_max[n_items] = ilo[iv] + c > nb ? nb - c : ilo[iv];
        lo_min[n_items] = ilo[iv] + icnt[iv] - c < 0 ? 0 : ilo[iv] + icnt[iv] - c;
        if (lo_min[n_items] > lo_max[n_items]) lo_min[n_items] = lo_max[n_items];
      } else {            // no bins: all weights are zero, any run will do
        lo_min[n_items] = 0;
        lo_max[n_items] = nb - c;
      }
      ++n_items;
    }
    int lane_of[32], lo_of[32], slot_item[32];
    for (int i = 0; i < 32; ++i) { lane_of[i] = -1; lo_of[i] = 0; slot_item[i] = -1; }
    if (bank_mod == 16) {
      // Kuhn's augmenting paths: item -> slot (half h, residue rho) = lane 16 h + rho
      auto can = [&](int it, int slot) {
        const int rho = slot & 15;
        for (int lo = lo_max[it]; lo >= lo_min[it]; --lo)
          if ((lo & 15) == rho) return lo;
        return -1;
      };
      std::vector<char> seen(32);
      std::function<bool(int)> aug = [&](int it) {
        for (int slot = 0; slot < 32; ++slot) {
          if (seen[slot] || can(it, slot) < 0) continue;
          seen[slot] = 1;
          if (slot_item[slot] < 0 || aug(slot_item[slot])) { slot_item[slot] = it; return true; }
        }
        return false;
      };
      // widest (least slack) items first
      std::vector<int> order(n_items);
      for (int i = 0; i < n_items; ++i) order[i] = i;
      std::sort(order.begin(), order.end(), [&](int a, int b2) { return lo_max[a] - lo_min[a] < lo_max[b2] - lo_min[b2]; });
      for (int it : order) { std::fill(seen.begin(), seen.end(), 0); aug(it); }
      for (int slot = 0; slot < 32; ++slot)
        if (slot_item[slot] >= 0) { lane_of[slot_item[slot]] = slot; lo_of[slot_item[slot]] = can(slot_item[slot], slot); }
      for (int it = 0; it < n_items; ++it) {     // unmatched: any free lane, natural start (some conflicts remain)
        if (lane_of[it] >= 0) continue;
        for (int slot = 0; slot < 32; ++slot)
          if (slot_item[slot] < 0) { slot_item[slot] = it; lane_of[it] = slot; lo_of[it] = lo_max[it]; break; }
      }
    } else {
      for (int it = 0; it < n_items; ++it) { lane_of[it] = it; lo_of[it] = icnt[item_iv[it]] ? lo_max[it] : (1 + c > nb ? nb - c : 1); slot_item[it] = it; }
    }
    // idle lanes read the same run as an active lane of their half-warp (a broadcast, not a conflict), partner = self
    for (int l = 0; l < 32; ++l) {
      int lo = 0;
      for (int it = 0; it < n_items; ++it)
        if ((lane_of[it] >> 4) == (l >> 4)) { lo = lo_of[it]; break; }
      mlo[32 * r + l] = lo | (l << 12);
    }
    for (int it = 0; it < n_items; ++it) {
      const int iv = item_iv[it], l = lane_of[it], lo = lo_of[it];
      const bool outputs = iv < nm && iv <= 31 * r + 30;
      const int partner = (it + 1 < n_items) ? lane_of[it + 1] : l;
      mlo[32 * r + l] = lo | (partner << 12) | ((iv & 0xff) << 17) | (outputs ? (int)0x80000000u : 0);
      for (int q = 0; q < c; ++q) {
        const int k = lo + q;
        if (k < 0 || k >= nb || iv_of[k] != iv) continue;   // padding slot: weight 0
        const float up = iv < nm ? W(iv, k) : 0.f;
        const float dn = iv >= 1 ? W(iv - 1, k) : 0.f;
        mw[(size_t)(base + q) * 32 + l] = make_float2(scale * up, scale * dn);
      }
    }
    base += c;
  }
  for (int r = rounds; r < kMelRounds; ++r) { cnt_out[r] = 0; base_out[r] = base; }
  return 0;
}

// Compact form of the interval table for the fixed-shape mel stage (fbank_tile.cuh, B200FE_MEL_COMPACT): one float per
// slot, appended behind the kMelSlots * 32 (up, down) pairs.  u = up weight; the kernel derives down = scale - u
// wherever u > 0.  Encodable: padding (0, 0) -> 0; up + down == scale (to float rounding) -> up; a lone up weight whose
// lane's down-slope sum nobody reads (interval 0) or that equals the scale -> up; a lone down weight on a lane whose
// up-slope sum nobody reads (it outputs no filter) -> scale - down, or 2^-90 when down == scale.  Returns false if a
// slot is not (then the fixed-shape instantiations are not used).
bool build_compact_mel(std::vector<float2>& mw, const std::vector<int>& mlo, int rounds, const int* cnt, const int* base,
                       float scale) {
  std::vector<float> cu((size_t)kMelSlots * 32, 0.f);
  bool ok = true;
  for (int r = 0; r < rounds; ++r)
    for (int l = 0; l < 32; ++l) {
      const unsigned word = (unsigned)mlo[32 * r + l];
      const int iv = (int)((word >> 17) & 0xffu);
      const bool outputs = (word >> 31) != 0;
      for (int q = 0; q < cnt[r]; ++q) {
        const float2 w = mw[(size_t)(base[r] + q) * 32 + l];
        float u = 0.f;
        if (w.x == 0.f && w.y == 0.f) u = 0.f;
        else if (w.x > 0.f && w.y > 0.f) { u = w.x; ok = ok && fabsf(w.x + w.y - scale) <= 4e-7f * scale; }
        else if (w.y == 0.f) { u = w.x; ok = ok && (iv == 0 || w.x == scale); }
        else { u = w.y < scale ? scale - w.y : 0x1p-90f; ok = ok && !outputs && w.y <= scale; }
        ok = ok && (u == 0.f || u >= 0x1p-100f);
        cu[(size_t)(base[r] + q) * 32 + l] = u;
      }
    }
  mw.resize((size_t)kMelSlots * 32 + (size_t)kMelSlots * 16, make_float2(0.f, 0.f));
  memcpy(mw.data() + (size_t)kMelSlots * 32, cu.data(), cu.size() * sizeof(float));
  return ok;
}

int frame_count(long long n, int win, int shift) { return n < win ? 0 : (int)(1 + (n - win) / shift); }
int ceil_div(int a, int b) { return (a + b - 1) / b; }

// VF:147 + TA:137-139: window size of an utterance shorter than the configured frame
int short_window(const b200fe_config& c, long long n) {
  const double fl = (double)n / (double)c.sample_rate * 1000.0;
  return (int)((double)c.sample_rate * fl * 0.001);
}

struct Plan {
  std::vector<UttDesc> utts;
  std::vector<ShortDesc> shorts;
  int n_tiles = 0;
  int n_quads = 0;
  long long max_rows = 0;
  long long total_rows = 0;
  long long stat_rows = 0;   // rows of regular utterances: the ones the statistics pass accumulates
};

int make_plan(b200fe_handle* h, const int64_t* lengths, const int64_t* offsets, int64_t row_stride, int batch,
              Plan& pl) {
  pl.utts.resize(batch);
  pl.shorts.clear();
  int tile = 0, quad = 0;
  for (int u = 0; u < batch; ++u) {
    const long long n = lengths[u];
    if (n < 0 || n > 0x7fffffffll) return fail(h, B200FE_E_INVALID, "utterance length outside [0, 2^31)");
    UttDesc d;
    d.wave_off = offsets ? offsets[u] : (long long)u * row_stride;
    d.n_samples = (int)n;
    d.tile_begin = tile;
    d.quad_begin = quad;
    d.row_begin = (int)pl.total_rows;
    if (n >= h->L) {
      d.n_frames = frame_count(n, h->L, h->S);
      d.n_rows = ceil_div(d.n_frames, h->cfg.lfr_n);
      tile += ceil_div(d.n_rows, h->rows_per_tile);
      quad += ceil_div(d.n_frames, 4);
      pl.stat_rows += d.n_rows;
    } else {
      const int win = short_window(h->cfg, n);
      if (win < 2 || win > n) return fail(h, B200FE_E_SHORT, "choose a window size " + std::to_string(win) +
                                          " that is [2, " + std::to_string(n) + "]");
      ShortDesc s;
      s.wave_off = d.wave_off;
      s.utt = u;
      s.n_samples = (int)n;
      s.win = win;
      s.nfft = next_pow2(win);
      s.n_frames = frame_count(n, win, h->S);
      s.n_rows = ceil_div(s.n_frames, h->cfg.lfr_n);
      s.mel_off = 0;
      s.row_begin = (int)pl.total_rows;
      pl.shorts.push_back(s);
      d.n_frames = 0;          // the tile kernel skips it (no tiles); rows come from the short path
      d.n_rows = s.n_rows;
    }
    pl.max_rows = d.n_rows > pl.max_rows ? d.n_rows : pl.max_rows;
    pl.total_rows += d.n_rows;
    pl.utts[u] = d;
  }
  pl.n_tiles = tile;
  pl.n_quads = quad;
  return 0;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// the tile list (statistics pass) and the quad list (warp kernel) share one region: a launch uses one of them
// then 512 floats per utterance shorter than one frame (scratch for int16 input: the short path reads float32)
size_t workspace_need(int batch, int n_tiles = 0, int n_quads = 0, int n_short = 0) {
  const size_t lists = std::max((size_t)n_tiles * sizeof(TileDesc), (size_t)n_quads * sizeof(QuadDesc));
  return 256 /* work counter */ + align256((size_t)batch * sizeof(UttDesc)) + align256((size_t)batch * sizeof(ShortDesc)) + align256(lists) +
         align256((size_t)n_short * 512 * sizeof(float));
}

// stream-ordered upload through a small ring of pinned buffers
int upload(b200fe_handle* h, const void* src, size_t bytes, void* dst, cudaStream_t st) {
  if (bytes == 0) return 0;
  PinnedSlot& s = h->slots[h->next_slot];
  h->next_slot = (h->next_slot + 1) & 3;
  if (s.busy) {
    CUDA_TRY(h, cudaEventSynchronize(s.ev));
    s.busy = false;
  }
  if (s.cap < bytes) {
    if (s.ptr) cudaFreeHost(s.ptr);
    s.ptr = nullptr;
    s.cap = 0;
    size_t cap = bytes < 65536 ? 65536 : bytes * 2;
    CUDA_TRY(h, cudaMallocHost(&s.ptr, cap));
    s.cap = cap;
  }
  if (!s.ev) CUDA_TRY(h, cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
  memcpy(s.ptr, src, bytes);
  CUDA_TRY(h, cudaMemcpyAsync(dst, s.ptr, bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaEventRecord(s.ev, st));
  s.busy = true;
  return 0;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is set once per (kernel, device) and grown when a larger size is asked for
cudaError_t allow_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> done;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[{func, dev}];
  if (have >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

// event pairs of the timed launches come from a per-handle pool (no cudaEventCreate on the launch path after warm-up)
int timed_events(b200fe_handle* h, cudaEvent_t& e0, cudaEvent_t& e1) {
  if (h->event_pool.size() < 2) {
    cudaEvent_t a = nullptr, b = nullptr;
    CUDA_TRY(h, cudaEventCreate(&a));
    CUDA_TRY(h, cudaEventCreate(&b));
    h->event_pool.push_back(a);
    h->event_pool.push_back(b);
  }
  e1 = h->event_pool.back(); h->event_pool.pop_back();
  e0 = h->event_pool.back(); h->event_pool.pop_back();
  return 0;
}

template <int NROWS, bool EXACT, class MELS>
int launch_tile(b200fe_handle* h, const TileParams& p, int grid, bool dither, bool stats, cudaStream_t st) {
#define LAUNCH(DI, STT)                                                                                         \
  do {                                                                                                          \
    auto k = fbank_lfr_cmvn_tile_kernel<NROWS, EXACT, DI, STT, MELS>;                                                 \
    CUDA_TRY(h, allow_dynamic_smem((const void*)k, h->smem_bytes));                                             \
    k<<<grid, kCtaThreads, h->smem_bytes, st>>>(p);                                                             \
  } while (0)
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = h->profile_every > 0 && (h->profile_tick++ % h->profile_every) == 0;
  if (timed) {
    if (int rc_ev = timed_events(h, e0, e1)) return rc_ev;
    CUDA_TRY(h, cudaEventRecord(e0, st));
  }
#ifdef B200FE_BENCH_ONLY
  if (dither || stats) return fail(h, B200FE_E_UNSUPPORTED, "bench-only build");
  LAUNCH(false, false);
#else
  if (dither) { if (stats) LAUNCH(true, true); else LAUNCH(true, false); }
  else        { if (stats) LAUNCH(false, true); else LAUNCH(false, false); }
#endif
#undef LAUNCH
  if (timed) {
    CUDA_TRY(h, cudaEventRecord(e1, st));
    h->prof_events.emplace_back(e0, e1);
  }
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return 0;
}

template <int NROWS, bool EXACT, class MELS, int SR, class SampleT>
int launch_warp(b200fe_handle* h, const QuadParams& p, int grid, bool dither, cudaStream_t st) {
  const size_t smem = warp_smem_bytes();
#define LAUNCHW(DI)                                                                                   \
  do {                                                                                                \
    auto k = p.rows_cap < 0 ? fbank_warp_kernel<NROWS, EXACT, DI, MELS, SR, SampleT, true>                      \
                            : fbank_warp_kernel<NROWS, EXACT, DI, MELS, SR, SampleT, false>;                     \
    CUDA_TRY(h, allow_dynamic_smem((const void*)k, smem));                                            \
    cudaLaunchConfig_t lc = {};                                                                       \
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kCtaThreads); lc.dynamicSmemBytes = smem; lc.stream = st; \
    cudaLaunchAttribute at[1];                                                                        \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                    \
    at[0].val.programmaticStreamSerializationAllowed = 1;                                             \
    lc.attrs = at; lc.numAttrs = timed ? 0 : 1;                                                                 \
    CUDA_TRY(h, cudaLaunchKernelEx(&lc, k, p));                                                       \
  } while (0)
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = h->profile_every > 0 && (h->profile_tick++ % h->profile_every) == 0;
  if (timed) {
    if (int rc_ev = timed_events(h, e0, e1)) return rc_ev;
    CUDA_TRY(h, cudaEventRecord(e0, st));
  }
#ifdef B200FE_BENCH_ONLY   // experiment builds (tools/build_variant.py): only what bench.py launches is instantiated
  if (dither) return fail(h, B200FE_E_UNSUPPORTED, "bench-only build");
  LAUNCHW(false);
#else
  if (dither) LAUNCHW(true); else LAUNCHW(false);
#endif
#undef LAUNCHW
  if (timed) {
    CUDA_TRY(h, cudaEventRecord(e1, st));
    h->prof_events.emplace_back(e0, e1);
  }
  CUDA_TRY(h, cudaGetLastError());
  h->launches++;
  return 0;
}

int ensure_short_banks(b200fe_handle* h, Plan& pl, cudaStream_t st) {
  (void)st;
  bool grow = false;
  for (auto& s : pl.shorts)
    if (!h->short_mel_off.count(s.nfft)) grow = true;
  if (grow) {
    // all banks for fft sizes 2..512; 80 * (1+2+..+256) floats = 164 KB, once per handle.  Built into locals and
    // committed only after the (blocking) upload succeeded, so a failed call leaves the handle as it was.
    std::vector<float> all;
    std::map<int, int> off;
    for (int nfft = 2; nfft <= 512; nfft <<= 1) {
      std::vector<float> b;
      if (build_mel(h->cfg.n_mels, nfft, h->cfg.sample_rate, h->cfg.low_freq, h->cfg.high_freq, b) != 0)
        return fail(h, B200FE_E_INVALID, "bad low/high frequency");
      off[nfft] = (int)all.size();
      all.insert(all.end(), b.begin(), b.end());
    }
    float* dev = nullptr;
    CUDA_TRY(h, cudaMalloc(&dev, all.size() * sizeof(float)));
    const cudaError_t e = cudaMemcpy(dev, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      cudaFree(dev);
      return fail(h, B200FE_E_CUDA, std::string("short-utterance mel banks: ") + cudaGetErrorString(e));
    }
    h->d_short_mel = dev;
    h->short_mel_floats = all.size();
    h->short_mel_off.swap(off);
  }
  for (auto& s : pl.shorts) s.mel_off = h->short_mel_off[s.nfft];
  return 0;
}

}  // namespace

extern "C" {

void b200fe_default_config(b200fe_config* c) {
  memset(c, 0, sizeof(*c));
  c->struct_size = (int32_t)sizeof(b200fe_config);
  c->sample_rate = 16000;
  c->frame_length_ms = 25.f;
  c->frame_shift_ms = 10.f;
  c->n_mels = 80;
  c->window_type = B200FE_WIN_HAMMING;
  c->lfr_m = 1;
  c->lfr_n = 1;
  c->dither = 1.0f;
  c->snip_edges = 1;
  c->upscale_samples = 1;
  c->preemphasis = 0.97f;
  c->remove_dc_offset = 1;
  c->low_freq = 20.f;
  c->high_freq = 0.f;
  c->blackman_coeff = 0.42f;
  c->log_floor = 1.1920928955078125e-07f;
}

int b200fe_create(const b200fe_config* cfg, const float* cmvn_host, b200fe_handle** out) {
  if (!cfg || !out) return fail(nullptr, B200FE_E_INVALID, "null argument");
  if (cfg->struct_size != (int32_t)sizeof(b200fe_config)) return fail(nullptr, B200FE_E_INVALID, "b200fe_config size mismatch");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, B200FE_E_CUDA, "no CUDA device: the B200 front-end has no CPU fallback");
  const int L = (int)((double)cfg->sample_rate * (double)cfg->frame_length_ms * 0.001);
  const int S = (int)((double)cfg->sample_rate * (double)cfg->frame_shift_ms * 0.001);
  if (L < 2 || S <= 0) return fail(nullptr, B200FE_E_INVALID, "bad frame length / shift");
  const int nfft = next_pow2(L);
  if (nfft != kNfft) return fail(nullptr, B200FE_E_UNSUPPORTED, "only frames that pad to a 512-point FFT are implemented (257..512 samples)");
  if (cfg->n_mels < 4 || cfg->n_mels > kMaxMels || (cfg->n_mels & 3)) return fail(nullptr, B200FE_E_UNSUPPORTED, "n_mels must be a multiple of 4 in [4, 128]");
  if (cfg->lfr_m < 1 || cfg->lfr_n < 1 || cfg->lfr_m > kFMax) return fail(nullptr, B200FE_E_INVALID, "bad lfr_m / lfr_n");
  if (cfg->lfr_m * cfg->n_mels > 8 * kCtaThreads) return fail(nullptr, B200FE_E_UNSUPPORTED, "n_mels*lfr_m > 1024");
  if (!cfg->snip_edges) return fail(nullptr, B200FE_E_UNSUPPORTED, "snip_edges=False is not implemented");
  if (cfg->window_type < 0 || cfg->window_type > B200FE_WIN_BLACKMAN) return fail(nullptr, B200FE_E_INVALID, "Invalid window type");
  if (!(cfg->preemphasis >= 0.f && cfg->preemphasis <= 1.f)) return fail(nullptr, B200FE_E_INVALID, "`preemphasis_coefficient` must be between [0,1]");

  b200fe_handle* h = new b200fe_handle();
  h->cfg = *cfg;
  h->L = L; h->S = S; h->nfft = nfft;
  h->D = cfg->lfr_m * cfg->n_mels;
  h->rows_per_tile = (kFMax - cfg->lfr_m) / cfg->lfr_n + 1;
  h->e_cap = (((kFMax - 1) * S + L + 8) + 3) & ~3;
  h->smem_bytes = tile_smem_bytes(h->e_cap, cfg->n_mels);
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, dev);
  h->n_sms = prop.multiProcessorCount;
  if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
    delete h;
    return fail(nullptr, B200FE_E_UNSUPPORTED, "frame shift too large for the shared-memory tile");
  }
  auto bail = [&](int code, const std::string& m) { b200fe_destroy(h); return fail(nullptr, code, m); };
#define CK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return bail(B200FE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } while (0)

  build_window(cfg->window_type, L, cfg->blackman_coeff, h->window_host);
  std::vector<float> win512(512, 0.f);
  const float up = cfg->upscale_samples ? 32768.f : 1.f;
  for (int i = 0; i < L; ++i) win512[i] = h->window_host[i] * up;   // power-of-two scale: exact
  if (build_mel(cfg->n_mels, nfft, cfg->sample_rate, cfg->low_freq, cfg->high_freq, h->mel_host) != 0)
    return bail(B200FE_E_INVALID, "Bad values in options: low-freq / high-freq vs. nyquist");

  // sparse filterbank by interval: interval iv = bins whose mel lies in [centre(iv-1), centre(iv)); such a bin feeds
  // the up-slope of filter iv and the down-slope of filter iv-1 and nothing else (TA:494-499, triangles in mel domain)
  std::vector<float2> mw;
  std::vector<int> mlo;
  {
    std::string why;
    if (build_interval_table(h->mel_host, cfg->n_mels, nfft / 2, nfft / 2, kMelScale, mw, mlo, h->mel_rounds, h->mel_cnt,
                             h->mel_base, why, 16, B200FE_MEL_COMPACT != 0) != 0)
      return bail(B200FE_E_UNSUPPORTED, why);
    const bool compact = build_compact_mel(mw, mlo, h->mel_rounds, h->mel_cnt, h->mel_base, kMelScale);
    h->mel_paraformer = h->mel_rounds == MelShapeParaformer::kRounds && (compact || !B200FE_MEL_COMPACT);
    for (int r = 0; r < MelShapeParaformer::kRounds && h->mel_paraformer; ++r)
      h->mel_paraformer = h->mel_cnt[r] == MelShapeParaformer::cnt(r);
  }
  // stage-2 twiddles of the packed real FFT (TileParams::twiddle) and the column-0 table, evaluated in double
  std::vector<float2> tw(kTw2Total, make_float2(0.f, 0.f));
  for (int k1 = 1; k1 <= 16; ++k1)
    for (int c = 0; c < 16; ++c) {
      const int ph = (c * k1) % 512;
      const double sc = (k1 == 8 || k1 == 16) ? 2.0 : 1.0;
      tw[(k1 - 1) * kTwPitch + c] =
          make_float2((float)(sc * cos(2.0 * M_PI * ph / 512.0)), (float)(-sc * sin(2.0 * M_PI * ph / 512.0)));
    }
  for (int t = 0; t < 8; ++t)
    for (int c = 0; c < 8; ++c) {
      const int ph = (c * t) % 16;
      tw[kTw2Table + t * kC0Pitch + c] =
          make_float2((float)(2.0 * cos(2.0 * M_PI * ph / 16.0)), (float)(-2.0 * sin(2.0 * M_PI * ph / 16.0)));
    }
  CK(cudaMalloc(&h->d_window, 512 * sizeof(float)));
  CK(cudaMalloc(&h->d_twiddle, tw.size() * sizeof(float2)));
  CK(cudaMalloc(&h->d_twiddle_stream, tw.size() * sizeof(float2)));
  CK(cudaMalloc(&h->d_mel_w, mw.size() * sizeof(float2)));
  CK(cudaMalloc(&h->d_mel_lo, mlo.size() * sizeof(int)));
  CK(cudaMemcpy(h->d_window, win512.data(), 512 * sizeof(float), cudaMemcpyHostToDevice));
  // tw carries the [8][kC0Pitch] column-0 table here: that copy is the stream kernels'; the offline kernels' copy gets
  // the lane-FFT twiddles in its place (B200FE_C0_SHFL)
  CK(cudaMemcpy(h->d_twiddle_stream, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
#if B200FE_C0_SHFL
  for (int i = kTw2Table; i < kTw2Total; ++i) tw[i] = make_float2(0.f, 0.f);
  fill_c0_lane_table(tw.data() + kTw2Table);
#endif
  CK(cudaMemcpy(h->d_twiddle, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_mel_w, mw.data(), mw.size() * sizeof(float2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(h->d_mel_lo, mlo.data(), mlo.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (cmvn_host) {
    CK(cudaMalloc(&h->d_cmvn, 2 * h->D * sizeof(float)));
    CK(cudaMemcpy(h->d_cmvn, cmvn_host, 2 * h->D * sizeof(float), cudaMemcpyHostToDevice));
    h->has_cmvn = true;
  }
  if (h->cfg.n_mels % 4 == 0 && h->cfg.lfr_m <= kCmvnSlots) {
    // the warp kernel's view of the CMVN table: per (slot, float4 piece of the mel row) the shift and the scale side by
    // side, 32 slots so that the slot field of any target code (kNoTarget included) indexes it; (x + 0) * 1 without CMVN
    const int M = h->cfg.n_mels, M4 = M / 4;
    std::vector<float> il((size_t)kCmvnSlots * M4 * 8, 0.f);
    for (int slot = 0; slot < kCmvnSlots; ++slot)
      for (int l = 0; l < M4; ++l)
        for (int k = 0; k < 4; ++k) {
          const bool in = slot < h->cfg.lfr_m;
          const int col = slot * M + 4 * l + k;
          il[((size_t)slot * M4 + l) * 8 + k] = (in && cmvn_host) ? cmvn_host[col] : 0.f;
          il[((size_t)slot * M4 + l) * 8 + 4 + k] = (in && cmvn_host) ? cmvn_host[h->D + col] : 1.f;
        }
    CK(cudaMalloc(&h->d_cmvn_il, il.size() * sizeof(float)));
    CK(cudaMemcpy(h->d_cmvn_il, il.data(), il.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
#undef CK
  *out = h;
  return B200FE_OK;
}

void b200fe_destroy(b200fe_handle* h) {
  if (!h) return;
  cudaFree(h->d_window); cudaFree(h->d_twiddle); cudaFree(h->d_twiddle_stream); cudaFree(h->d_mel_w); cudaFree(h->d_mel_lo);
  cudaFree(h->d_cmvn); cudaFree(h->d_cmvn_il); cudaFree(h->d_short_mel);
  for (auto& pr : h->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  for (auto& e : h->event_pool) cudaEventDestroy(e);
  for (auto& s : h->slots) {
    if (s.ev) { cudaEventSynchronize(s.ev); cudaEventDestroy(s.ev); }
    if (s.ptr) cudaFreeHost(s.ptr);
  }
  delete h;
}

const char* b200fe_last_error(const b200fe_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
int b200fe_output_dim(const b200fe_handle* h) { return h ? h->D : B200FE_E_INVALID; }
int b200fe_frame_samples(const b200fe_handle* h) { return h ? h->L : B200FE_E_INVALID; }
int b200fe_shift_samples(const b200fe_handle* h) { return h ? h->S : B200FE_E_INVALID; }
int b200fe_fft_size(const b200fe_handle* h) { return h ? h->nfft : B200FE_E_INVALID; }
int64_t b200fe_launch_count(const b200fe_handle* h) { return h ? h->launches : 0; }

int b200fe_lfr_targets(int frame, int n_frames, int n_rows, int lfr_m, int lfr_n, int n_mels, uint32_t targets_out[2]) {
  if (!targets_out || frame < 0 || frame >= n_frames || n_rows < 1 || lfr_m < 1 || lfr_n < 1 || n_mels < 1) return B200FE_E_INVALID;
  unsigned t2[2];
  const bool slow = quad_targets(frame, n_frames, n_rows, lfr_m, lfr_n, n_mels, t2);
  targets_out[0] = t2[0];
  targets_out[1] = t2[1];
  return slow ? 1 : 0;
}

int b200fe_select_kernel(b200fe_handle* h, int which) {
  if (!h || which < 0 || which > 1) return B200FE_E_INVALID;
  std::lock_guard<std::mutex> lock(h->mu);
  h->force_tile = which == 1;
  return B200FE_OK;
}

int b200fe_profile_enable(b200fe_handle* h, int every) {
  if (!h) return B200FE_E_INVALID;
  std::lock_guard<std::mutex> lock(h->mu);
  h->profile_every = every > 0 ? every : 0;
  h->profile_tick = 0;
  return B200FE_OK;
}

int b200fe_profile_collect(b200fe_handle* h, double* total_ms, int64_t* n_launches) {
  if (!h) return B200FE_E_INVALID;
  std::lock_guard<std::mutex> lock(h->mu);
  double tot = 0.0;
  for (auto& pr : h->prof_events) {
    CUDA_TRY(h, cudaEventSynchronize(pr.second));
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, pr.first, pr.second));
    tot += ms;
    h->event_pool.push_back(pr.first);
    h->event_pool.push_back(pr.second);
  }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = (int64_t)h->prof_events.size();
  h->prof_events.clear();
  return B200FE_OK;
}

int b200fe_get_tables(const b200fe_handle* h, float* window_out, float* mel_out) {
  if (!h) return B200FE_E_INVALID;
  if (window_out) memcpy(window_out, h->window_host.data(), h->window_host.size() * sizeof(float));
  if (mel_out) memcpy(mel_out, h->mel_host.data(), h->mel_host.size() * sizeof(float));
  return B200FE_OK;
}

int b200fe_plan(b200fe_handle* h, const int64_t* lengths_host, int batch, int64_t* n_frames_out, int64_t* n_rows_out,
                int64_t* max_rows_out, size_t* workspace_bytes) {
  if (!h || !lengths_host || batch < 0) return fail(h, B200FE_E_INVALID, "null argument");
  Plan pl;
  int rc = make_plan(h, lengths_host, nullptr, 0, batch, pl);
  if (rc) return rc;
  size_t si = 0;
  for (int u = 0; u < batch; ++u) {
    int nf = pl.utts[u].n_frames;
    if (lengths_host[u] < h->L) nf = pl.shorts[si++].n_frames;
    if (n_frames_out) n_frames_out[u] = nf;
    if (n_rows_out) n_rows_out[u] = pl.utts[u].n_rows;
  }
  if (max_rows_out) *max_rows_out = pl.max_rows;
  if (workspace_bytes) *workspace_bytes = workspace_need(batch, pl.n_tiles, pl.n_quads, (int)pl.shorts.size());
  return B200FE_OK;
}

}  // extern "C"

namespace {

template <class SampleT>
int dispatch_warp(b200fe_handle* h, const QuadParams& p, int grid, bool dither, cudaStream_t st) {
  if (h->L == 400 && h->S == 160 && h->mel_paraformer) return launch_warp<25, true, MelShapeParaformer, 10, SampleT>(h, p, grid, dither, st);
#ifdef B200FE_BENCH_ONLY
  return fail(h, B200FE_E_UNSUPPORTED, "bench-only build");
#else
  if (h->L == 400 && h->S == 160) return launch_warp<25, true, MelShapeRuntime, 10, SampleT>(h, p, grid, dither, st);
  if (h->L == 400) return launch_warp<25, true, MelShapeRuntime, 0, SampleT>(h, p, grid, dither, st);
  return launch_warp<32, false, MelShapeRuntime, 0, SampleT>(h, p, grid, dither, st);
#endif
}

// b200fe_forward / b200fe_forward_pcm16.  pcm16: `wave_dev` is int16 PCM, sample value = s / 32768.
int forward_impl(b200fe_handle* h, const void* wave_any, bool pcm16, int64_t wave_total, const int64_t* offsets_host,
                 int64_t row_stride, const int64_t* lengths_host, int batch, float* feats_dev, int64_t rows_cap,
                 int64_t* feat_lens_dev, double* stats_dev, uint64_t dither_seed, void* workspace_dev,
                 size_t workspace_bytes, void* stream) {
  const float* wave_dev = static_cast<const float*>(wave_any);   // only dereferenced when !pcm16
  if (!h) return B200FE_E_INVALID;
  if (batch == 0) return B200FE_OK;
  if (!wave_dev || !lengths_host || !feats_dev || !workspace_dev || batch < 0)
    return fail(h, B200FE_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(h->mu);
  Plan pl;
  int rc = make_plan(h, lengths_host, offsets_host, row_stride, batch, pl);
  if (rc) return rc;
  if (workspace_bytes < workspace_need(batch, pl.n_tiles, pl.n_quads, (int)pl.shorts.size())) return fail(h, B200FE_E_WORKSPACE, "workspace too small");
  const bool rows_packed = rows_cap == B200FE_ROWS_PACKED;   // feats_dev = [sum of rows, D], utterance u at row_begin[u]
  if (!rows_packed && pl.max_rows > rows_cap) return fail(h, B200FE_E_INVALID, "rows_cap smaller than the longest utterance's row count");
  if (pl.total_rows > 0x7fffffffll) return fail(h, B200FE_E_INVALID, "more than 2^31 output rows in one call");
  for (int u = 0; u < batch; ++u) {
    const long long end = pl.utts[u].wave_off + lengths_host[u];
    if (pl.utts[u].wave_off < 0 || end > wave_total) return fail(h, B200FE_E_INVALID, "utterance outside the wave buffer");
  }
  int* d_counter = reinterpret_cast<int*>(workspace_dev);
  UttDesc* d_utts = reinterpret_cast<UttDesc*>((char*)workspace_dev + 256);
  ShortDesc* d_shorts = reinterpret_cast<ShortDesc*>((char*)d_utts + align256((size_t)batch * sizeof(UttDesc)));
  TileDesc* d_tiles = reinterpret_cast<TileDesc*>((char*)d_shorts + align256((size_t)batch * sizeof(ShortDesc)));
  QuadDesc* d_quads = reinterpret_cast<QuadDesc*>(d_tiles);   // same region: a launch uses one of the two lists
  // the warp kernel does everything except the CMVN statistics (which need the row-major tile pass)
  const bool use_warp = stats_dev == nullptr && warp_kernel_fits(h->L, h->S, pcm16) && (!h->force_tile || pcm16) &&
                        pl.max_rows * (long long)h->D < (1ll << kTargetOffBits) - 1;   // target offsets are relative to the utterance
  // utterance table: inside the prep launch's parameters when it fits (warp path, no short utterances, which read
  // d_utts-independent descriptors of their own), else one pinned-buffer upload
  const bool utts_in_params = use_warp && pl.n_quads > 0 && batch <= kParamUtts;
  if (!utts_in_params && (rc = upload(h, pl.utts.data(), (size_t)batch * sizeof(UttDesc), d_utts, st))) return rc;
  if (rows_packed && !(use_warp && pl.n_quads > 0))
    return fail(h, B200FE_E_UNSUPPORTED, "the rows-packed output is written by the warp kernel only (no statistics pass, "
                                         "frame shifts whose quad fits its buffer, at least one full frame in the batch)");
  if (!(use_warp && pl.n_quads > 0) && batch > 65535)
    return fail(h, B200FE_E_UNSUPPORTED, "more than 65535 utterances per call are implemented in the warp kernel only "
                                         "(the tile / statistics path puts the batch in grid.y)");
  if (pcm16 && !use_warp)
    return fail(h, B200FE_E_UNSUPPORTED, "int16 input is implemented in the warp kernel only (no statistics pass, "
                                         "frame shifts whose quad fits its buffer)");
  const long long per_utt = rows_cap * (long long)h->D / 4;
  int gx = (int)((per_utt + 256 * 8 - 1) / (256 * 8));
  gx = gx < 1 ? 1 : (gx > 64 ? 64 : gx);
  if (use_warp && pl.n_quads > 0) {
    // 1. one launch: quad list (+ work counter), feat_lens, and the padding rows unless the warp kernel writes them
    const int qb = (pl.n_quads + 255) / 256, ub = (batch + 255) / 256;
    const int pad_bx = (B200FE_PAD_MODE == 1 || rows_packed) ? 0 : gx;
    const int grid0 = qb + ub + pad_bx * batch;
    if (utts_in_params) {
      static_assert(sizeof(UttTable) <= 16384, "kernel parameters");
      UttTable& tab = h->utt_tab;
      memcpy(tab.u, pl.utts.data(), (size_t)batch * sizeof(UttDesc));
      prep_warp_kernel_tab<<<grid0, 256, 0, st>>>(tab, batch, pl.n_quads, qb, ub, h->S, h->cfg.lfr_m, h->cfg.lfr_n, h->cfg.n_mels,
                                                  d_quads, d_counter, feats_dev, rows_cap, (long long*)feat_lens_dev,
                                                  B200FE_PAD_MODE == 1 ? d_utts : nullptr, pad_bx);
    } else {
      prep_warp_kernel<<<grid0, 256, 0, st>>>(d_utts, batch, pl.n_quads, qb, ub, h->S, h->cfg.lfr_m, h->cfg.lfr_n, h->cfg.n_mels,
                                              d_quads, d_counter, feats_dev, rows_cap, (long long*)feat_lens_dev, pad_bx);
    }
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
  } else {
    if (pl.n_tiles > 0) {
      build_tiles_kernel<<<(pl.n_tiles + 255) / 256, 256, 0, st>>>(d_utts, batch, pl.n_tiles, h->rows_per_tile, h->cfg.lfr_m,
                                                                   h->cfg.lfr_n, h->S, d_tiles, d_counter);
      CUDA_TRY(h, cudaGetLastError());
      h->launches++;
    }
    // 1. padding rows + feat_lens
    pad_rows_kernel<<<dim3(gx, batch), 256, 0, st>>>(d_utts, feats_dev, rows_cap, h->D, (long long*)feat_lens_dev);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
  }
  // 2. the fused kernel over all regular utterances: warp kernel, or tile kernel for the statistics pass
  if (use_warp && pl.n_quads > 0) {
    QuadParams p;
    p.wave = wave_any; p.wave_total = wave_total; p.quads = d_quads; p.n_quads = pl.n_quads; p.next_quad = d_counter;
    p.feats = feats_dev; p.rows_cap = rows_cap;
    p.frame_len = h->L; p.frame_shift = h->S; p.n_mels = h->cfg.n_mels; p.lfr_m = h->cfg.lfr_m; p.lfr_n = h->cfg.lfr_n;
    p.preemph = h->cfg.preemphasis; p.remove_dc = h->cfg.remove_dc_offset; p.log_floor = h->cfg.log_floor;
    p.dither = h->cfg.dither / (h->cfg.upscale_samples ? 32768.f : 1.f);   // TA:179 adds it after the 2^15 upscale
    p.seed = dither_seed;
    p.window = h->d_window; p.twiddle = h->d_twiddle; p.mel_w = h->d_mel_w; p.mel_lo = h->d_mel_lo; p.mel_rounds = h->mel_rounds;
    for (int r = 0; r < kMelRounds; ++r) { p.mel_cnt[r] = h->mel_cnt[r]; p.mel_base[r] = h->mel_base[r]; }
    p.cmvn = h->d_cmvn; p.cmvn_il = h->d_cmvn_il;
    p.utts = d_utts; p.batch = batch;
    const int ctas = (pl.n_quads + kWarps - 1) / kWarps;
    const int grid = ctas < B200FE_WARP_CTAS * h->n_sms ? ctas : B200FE_WARP_CTAS * h->n_sms;
    const bool dither = h->cfg.dither != 0.f;
  #ifdef B200FE_BENCH_ONLY
    rc = pcm16 ? fail(h, B200FE_E_UNSUPPORTED, "bench-only build") : dispatch_warp<float>(h, p, grid, dither, st);
#else
    rc = pcm16 ? dispatch_warp<short>(h, p, grid, dither, st) : dispatch_warp<float>(h, p, grid, dither, st);
#endif
    if (rc) return rc;
  } else if (pl.n_tiles > 0) {
    TileParams p;
    p.wave = wave_dev; p.wave_total = wave_total; p.utts = d_utts; p.tiles = d_tiles; p.batch = batch; p.n_tiles = pl.n_tiles;
    p.next_tile = d_counter;
    p.feats = feats_dev; p.rows_cap = rows_cap; p.stats = stats_dev;
    p.frame_len = h->L; p.frame_shift = h->S; p.n_mels = h->cfg.n_mels;
    p.lfr_m = h->cfg.lfr_m; p.lfr_n = h->cfg.lfr_n; p.rows_per_tile = h->rows_per_tile; p.e_cap = h->e_cap;
    p.preemph = h->cfg.preemphasis; p.remove_dc = h->cfg.remove_dc_offset; p.log_floor = h->cfg.log_floor;
    p.dither = h->cfg.dither / (h->cfg.upscale_samples ? 32768.f : 1.f);   // TA:179 adds it after the 2^15 upscale
    p.seed = dither_seed;
    p.window = h->d_window; p.twiddle = h->d_twiddle; p.mel_w = h->d_mel_w; p.mel_lo = h->d_mel_lo; p.mel_rounds = h->mel_rounds;
    for (int r = 0; r < kMelRounds; ++r) { p.mel_cnt[r] = h->mel_cnt[r]; p.mel_base[r] = h->mel_base[r]; }
    p.cmvn = h->d_cmvn;
    const int grid = pl.n_tiles < 3 * h->n_sms ? pl.n_tiles : 3 * h->n_sms;
    const bool dither = h->cfg.dither != 0.f, stats = stats_dev != nullptr;
    if (h->L == 400 && h->mel_paraformer) rc = launch_tile<25, true, MelShapeParaformer>(h, p, grid, dither, stats, st);
#ifndef B200FE_BENCH_ONLY
    else if (h->L == 400) rc = launch_tile<25, true, MelShapeRuntime>(h, p, grid, dither, stats, st);
    else rc = launch_tile<32, false, MelShapeRuntime>(h, p, grid, dither, stats, st);
#endif
    if (rc) return rc;
  }
  // 3. utterances shorter than one frame (VF:147)
  if (!pl.shorts.empty()) {
    if (h->cfg.dither != 0.f) return fail(h, B200FE_E_UNSUPPORTED, "dither on utterances shorter than one frame");
    if ((rc = ensure_short_banks(h, pl, st))) return rc;
    const float* short_wave = wave_dev;
    if (pcm16) {   // the short path reads float32: convert those few samples into the workspace scratch
      float* scratch = reinterpret_cast<float*>((char*)d_tiles + align256(std::max((size_t)pl.n_tiles * sizeof(TileDesc),
                                                                                 (size_t)pl.n_quads * sizeof(QuadDesc))));
      std::vector<long long> src_off(pl.shorts.size());
      for (size_t k = 0; k < pl.shorts.size(); ++k) { src_off[k] = pl.shorts[k].wave_off; pl.shorts[k].wave_off = (long long)k * 512; }
      if ((rc = upload(h, pl.shorts.data(), pl.shorts.size() * sizeof(ShortDesc), d_shorts, st))) return rc;
      for (size_t k = 0; k < pl.shorts.size(); ++k) {
        pcm16_to_float_kernel<<<1, 256, 0, st>>>(static_cast<const short*>(wave_any) + src_off[k], pl.shorts[k].n_samples,
                                                 scratch + k * 512);
        h->launches++;
      }
      CUDA_TRY(h, cudaGetLastError());
      short_wave = scratch;
    } else if ((rc = upload(h, pl.shorts.data(), pl.shorts.size() * sizeof(ShortDesc), d_shorts, st))) {
      return rc;
    }
    short_utt_kernel<<<(int)pl.shorts.size(), 256, 0, st>>>(
        short_wave, d_shorts, h->d_short_mel, h->S, h->cfg.n_mels, h->cfg.lfr_m, h->cfg.lfr_n, h->cfg.window_type,
        h->cfg.blackman_coeff, h->cfg.preemphasis, h->cfg.remove_dc_offset, h->cfg.upscale_samples ? 32768.f : 1.f,
        h->cfg.log_floor, h->d_cmvn, feats_dev, rows_cap);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
  }
  if (stats_dev) {
    // utterances shorter than one frame (shrunken frame, other filterbank: VF:147) stay out of the statistics: their
    // rows are neither summed nor counted
    add_count_kernel<<<1, 1, 0, st>>>(stats_dev + 2 * h->D, (double)pl.stat_rows);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
  }
  return B200FE_OK;
}

}  // namespace

extern "C" {

int b200fe_forward(b200fe_handle* h, const float* wave_dev, int64_t wave_total, const int64_t* offsets_host,
                   int64_t row_stride, const int64_t* lengths_host, int batch, float* feats_dev, int64_t rows_cap,
                   int64_t* feat_lens_dev, double* stats_dev, uint64_t dither_seed, void* workspace_dev,
                   size_t workspace_bytes, void* stream) {
  return forward_impl(h, wave_dev, false, wave_total, offsets_host, row_stride, lengths_host, batch, feats_dev, rows_cap,
                      feat_lens_dev, stats_dev, dither_seed, workspace_dev, workspace_bytes, stream);
}

int b200fe_forward_pcm16(b200fe_handle* h, const int16_t* wave_dev, int64_t wave_total, const int64_t* offsets_host,
                         int64_t row_stride, const int64_t* lengths_host, int batch, float* feats_dev, int64_t rows_cap,
                         int64_t* feat_lens_dev, uint64_t dither_seed, void* workspace_dev, size_t workspace_bytes,
                         void* stream) {
  return forward_impl(h, wave_dev, true, wave_total, offsets_host, row_stride, lengths_host, batch, feats_dev, rows_cap,
                      feat_lens_dev, nullptr, dither_seed, workspace_dev, workspace_bytes, stream);
}

int b200fe_lfr_cmvn(b200fe_handle* h, const float* fbank_dev, int64_t frames_cap, const int64_t* n_frames_host, int batch,
                    float* feats_dev, int64_t rows_cap, int64_t* feat_lens_dev, void* workspace_dev,
                    size_t workspace_bytes, void* stream) {
  if (!h) return B200FE_E_INVALID;
  if (batch == 0) return B200FE_OK;
  if (!fbank_dev || !n_frames_host || !feats_dev || !workspace_dev) return fail(h, B200FE_E_INVALID, "null argument");
  if (workspace_bytes < workspace_need(batch)) return fail(h, B200FE_E_WORKSPACE, "workspace too small");
  if (batch > 65535) return fail(h, B200FE_E_UNSUPPORTED, "at most 65535 utterances per b200fe_lfr_cmvn call");
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(h->mu);
  std::vector<UttDesc> utts(batch);
  long long max_rows = 0;
  for (int u = 0; u < batch; ++u) {
    UttDesc d{};
    d.n_frames = (int)n_frames_host[u];
    if (d.n_frames < 0 || d.n_frames > frames_cap) return fail(h, B200FE_E_INVALID, "bad frame count");
    d.n_rows = ceil_div(d.n_frames, h->cfg.lfr_n);
    max_rows = d.n_rows > max_rows ? d.n_rows : max_rows;
    utts[u] = d;
  }
  if (max_rows > rows_cap) return fail(h, B200FE_E_INVALID, "rows_cap too small");
  UttDesc* d_utts = reinterpret_cast<UttDesc*>(workspace_dev);
  int rc = upload(h, utts.data(), (size_t)batch * sizeof(UttDesc), d_utts, st);
  if (rc) return rc;
  pad_rows_kernel<<<dim3(16, batch), 256, 0, st>>>(d_utts, feats_dev, rows_cap, h->D, (long long*)feat_lens_dev);
  CUDA_TRY(h, cudaGetLastError());
  lfr_cmvn_kernel<<<dim3(32, batch), 256, 0, st>>>(fbank_dev, frames_cap, d_utts, h->cfg.n_mels, h->cfg.lfr_m,
                                                   h->cfg.lfr_n, h->d_cmvn, feats_dev, rows_cap);
  CUDA_TRY(h, cudaGetLastError());
  h->launches += 2;
  return B200FE_OK;
}

int b200fe_synth_uniform_ids(float* wave_dev, const int64_t* offsets_dev, const int64_t* lengths_dev,
                             const int64_t* utt_ids_dev_or_null, int batch, uint64_t seed, float amp, void* stream) {
  if (!wave_dev || !offsets_dev || !lengths_dev || batch <= 0) return B200FE_E_INVALID;
  synth_uniform_kernel<<<dim3(64, batch < 65535 ? batch : 65535), 256, 0, (cudaStream_t)stream>>>(
      wave_dev, (const long long*)offsets_dev, (const long long*)lengths_dev, (const long long*)utt_ids_dev_or_null, batch, seed,
      amp);
  return cudaGetLastError() == cudaSuccess ? B200FE_OK : B200FE_E_CUDA;
}

int b200fe_synth_uniform(float* wave_dev, const int64_t* offsets_dev, const int64_t* lengths_dev, int batch, uint64_t seed,
                         float amp, void* stream) {
  return b200fe_synth_uniform_ids(wave_dev, offsets_dev, lengths_dev, nullptr, batch, seed, amp, stream);
}

}  // extern "C"

#include "extras_api.inl"
#include "stream_api.inl"
#include "tts_api.inl"
#include "host_ingest.inl"
