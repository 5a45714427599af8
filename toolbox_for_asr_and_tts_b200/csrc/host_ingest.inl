// Host ingest (included at the end of b200fe.cu): the gather that funasr performs on the host in front of the
// front-end call - a list of np.float32 utterances padded into one [B, Nmax] tensor (pad_sequence in extract_fbank,
// UPSTREAM-RECALLED; the reference hands the list over at R:voice-service/app/services/voice_interface.py:2049-2053) -
// replaced by a multi-threaded gather into a length-packed PINNED staging buffer that is pipelined with the
// host -> device copy: while group g crosses PCIe, group g+1 is gathered.  Nothing here computes features.
#include <atomic>
#include <condition_variable>
#include <thread>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

// Copy into the pinned staging buffer with non-temporal stores: the destination is written once and next read by the
// PCIe DMA engine, so pulling its lines into the cache first (read-for-ownership) would only cost host memory bandwidth.
inline void stream_copy(char* dst, const char* src, size_t n) {
  size_t i = 0;
#if defined(__SSE2__)
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (; i + 64 <= n; i += 64) {
      const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
      const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
      const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
      const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
      _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
    }
  }
#endif
  if (i < n) memcpy(dst + i, src + i, n - i);
}

class HostPool {
 public:
  explicit HostPool(int n) : n_(n) {
    for (int t = 0; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return n_; }
  // runs job(t) on every worker and returns when all are done
  void run(const std::function<void(int)>& job) {
    std::unique_lock<std::mutex> lk(mu_);
    job_ = &job;
    pending_ = n_;
    ++gen_;
    cv_.notify_all();
    done_.wait(lk, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop(int t) {
    unsigned long long seen = 0;
    while (true) {
      const std::function<void(int)>* job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      (*job)(t);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  unsigned long long gen_ = 0;
  int pending_ = 0;
  bool stop_ = false;
};

std::mutex g_pool_mu;
std::unique_ptr<HostPool> g_pool;

HostPool& host_pool(int threads) {
  if (!g_pool || g_pool->size() != threads) g_pool.reset(new HostPool(threads));
  return *g_pool;
}

}  // namespace

extern "C" {

int b200fe_host_threads(void) {
  const unsigned hc = std::thread::hardware_concurrency();
  return hc == 0 ? 1 : (int)(hc > 32 ? 32 : hc);
}

int b200fe_host_ingest(const void* const* utterances_host, const int64_t* lengths, const int64_t* dst_offsets, int batch,
                       int elem_size, void* staging_pinned, void* wave_dev, int64_t capacity_elems, int groups, int threads,
                       void* stream) {
  if (batch == 0) return B200FE_OK;
  if (!utterances_host || !lengths || !dst_offsets || !staging_pinned || !wave_dev || batch < 0 ||
      (elem_size != 2 && elem_size != 4))
    return B200FE_E_INVALID;
  for (int u = 0; u < batch; ++u) {
    if (lengths[u] < 0 || dst_offsets[u] < 0 || dst_offsets[u] + lengths[u] > capacity_elems) return B200FE_E_INVALID;
    if (u && dst_offsets[u] < dst_offsets[u - 1] + lengths[u - 1]) return B200FE_E_INVALID;   // ascending, disjoint
    if (lengths[u] > 0 && !utterances_host[u]) return B200FE_E_INVALID;
  }
  if (threads <= 0) threads = b200fe_host_threads();
  if (groups <= 0) groups = 8;
  if (groups > batch) groups = batch;
  cudaStream_t st = (cudaStream_t)stream;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  HostPool& pool = host_pool(threads);
  // groups of consecutive utterances with about equal byte counts
  long long total = 0;
  for (int u = 0; u < batch; ++u) total += lengths[u];
  constexpr long long kChunk = 64 * 1024;   // elements per unit of work
  struct Piece { const char* src; char* dst; size_t bytes; };
  std::vector<Piece> pieces;
  int u0 = 0;
  long long done = 0;
  for (int g = 0; g < groups && u0 < batch; ++g) {
    const long long goal = total * (g + 1) / groups;
    int u1 = u0;
    long long acc = done;
    while (u1 < batch && (acc < goal || u1 == u0)) acc += lengths[u1++];
    if (g == groups - 1) u1 = batch;
    pieces.clear();
    for (int u = u0; u < u1; ++u)
      for (long long o = 0; o < lengths[u]; o += kChunk) {
        const long long n = lengths[u] - o < kChunk ? lengths[u] - o : kChunk;
        pieces.push_back({static_cast<const char*>(utterances_host[u]) + o * elem_size,
                          static_cast<char*>(staging_pinned) + (dst_offsets[u] + o) * elem_size, (size_t)n * elem_size});
      }
    std::atomic<size_t> next{0};
    const std::function<void(int)> job = [&](int) {
      for (size_t i = next.fetch_add(1); i < pieces.size(); i = next.fetch_add(1)) stream_copy(pieces[i].dst, pieces[i].src, pieces[i].bytes);
#if defined(__SSE2__)
      _mm_sfence();   // the streamed lines are globally visible before the copy that reads them is enqueued
#endif
    };
    pool.run(job);
    const long long b0 = dst_offsets[u0] * elem_size, b1 = (dst_offsets[u1 - 1] + lengths[u1 - 1]) * elem_size;
    if (b1 > b0 &&
        cudaMemcpyAsync(static_cast<char*>(wave_dev) + b0, static_cast<char*>(staging_pinned) + b0, (size_t)(b1 - b0),
                        cudaMemcpyHostToDevice, st) != cudaSuccess)
      return B200FE_E_CUDA;
    done = acc;
    u0 = u1;
  }
  return B200FE_OK;
}

}  // extern "C"
