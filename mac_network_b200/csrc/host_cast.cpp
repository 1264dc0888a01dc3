// Host side of the inference front end: fp32 -> bf16 conversion of the knowledge base on the CPU, so that the
// host-buffer path copies 2 bytes per KB element over PCIe instead of 4 (the bf16 read unit never touches the fp32
// copy: mac_cast_bf16 on the device would produce exactly these bits).  Round-to-nearest-even, NaN kept quiet --
// bit-identical to __float2bfloat16_rn for every finite input.
//
// A small persistent thread pool (no OpenMP runtime: torch ships its own libgomp and two runtimes in one process
// is asking for trouble).  One job at a time; the caller blocks until it is done (ctypes releases the GIL).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mac_b200.h"

namespace {

__attribute__((target_clones("avx512f", "avx2", "default")))
void cast_range(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  for (long long i = 0; i < n; ++i) {
    const uint32_t u = s[i];
    const uint32_t rounded = u + 0x7fffu + ((u >> 16) & 1u);
    const bool is_nan = (u & 0x7fffffffu) > 0x7f800000u;
    dst[i] = is_nan ? (uint16_t)((u >> 16) | 0x0040u) : (uint16_t)(rounded >> 16);
  }
}

class Pool {
 public:
  explicit Pool(int nthreads) : stop_(false), gen_(0), pending_(0) {
    for (int t = 0; t < nthreads; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++gen_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return (int)workers_.size(); }
  // post a job and return; at most one job is in flight (a second begin() first waits for the previous one)
  void begin(const float* src, uint16_t* dst, long long n) {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
    src_ = src; dst_ = dst; n_ = n;
    pending_ = size();
    ++gen_;
    cv_.notify_all();
  }
  void end() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }
  void run(const float* src, uint16_t* dst, long long n) {
    begin(src, dst, n);
    end();
  }

 private:
  void loop(int t) {
    unsigned long long seen = 0;
    for (;;) {
      const float* src; uint16_t* dst; long long n;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        src = src_; dst = dst_; n = n_;
      }
      // contiguous, 64-byte aligned slices
      const long long per = ((n + size() - 1) / size() + 31) & ~31LL;
      const long long lo = per * t, hi = lo + per < n ? lo + per : n;
      if (lo < hi) cast_range(src + lo, dst + lo, hi - lo);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  bool stop_;
  unsigned long long gen_;
  int pending_;
  const float* src_ = nullptr;
  uint16_t* dst_ = nullptr;
  long long n_ = 0;
};

std::mutex g_mu;
Pool* g_pool = nullptr;

}  // namespace

static Pool* pool_for(int nthreads) {       // g_mu held
  if (!g_pool || g_pool->size() != nthreads) {
    if (g_pool) g_pool->end();
    delete g_pool;
    g_pool = new Pool(nthreads);
  }
  return g_pool;
}

// asynchronous form: begin() returns as soon as the job is posted to the pool, end() waits for it
extern "C" int mac_host_cast_bf16_begin(const float* src, void* dst_bf16, long long n, int nthreads) {
  if (!src || !dst_bf16 || n < 0) return MAC_ERR_INVALID;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 64) nthreads = 64;
  std::lock_guard<std::mutex> lk(g_mu);
  pool_for(nthreads)->begin(src, reinterpret_cast<uint16_t*>(dst_bf16), n);
  return MAC_OK;
}

extern "C" int mac_host_cast_bf16_end(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_pool) g_pool->end();
  return MAC_OK;
}

extern "C" int mac_host_cast_bf16(const float* src, void* dst_bf16, long long n, int nthreads) {
  if (!src || !dst_bf16 || n < 0) return MAC_ERR_INVALID;
  if (n == 0) return MAC_OK;
  if (nthreads <= 1 || n < (1 << 16)) {
    cast_range(src, reinterpret_cast<uint16_t*>(dst_bf16), n);
    return MAC_OK;
  }
  if (nthreads > 64) nthreads = 64;
  std::lock_guard<std::mutex> lk(g_mu);           // one job at a time
  pool_for(nthreads)->run(src, reinterpret_cast<uint16_t*>(dst_bf16), n);
  return MAC_OK;
}

// CRC-32C (Castagnoli), the checksum of TensorFlow checkpoint ("tensor bundle") tensors and index blocks
// (mac_network_b200/tf_bundle.py): hardware instruction where the compiler targets SSE4.2, else a byte table.
#if defined(__SSE4_2__)
#include <nmmintrin.h>
#endif
extern "C" uint32_t mac_host_crc32c(const void* data, long long n, uint32_t crc) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
#if defined(__SSE4_2__)
  uint64_t c64 = c;
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    c64 = _mm_crc32_u64(c64, v);
    p += 8;
    n -= 8;
  }
  c = (uint32_t)c64;
  while (n-- > 0) c = _mm_crc32_u8(c, *p++);
#else
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t x = i;
      for (int k = 0; k < 8; ++k) x = (x >> 1) ^ ((x & 1u) ? 0x82F63B78u : 0u);
      table[i] = x;
    }
    init = true;
  }
  while (n-- > 0) c = table[(c ^ *p++) & 0xFFu] ^ (c >> 8);
#endif
  return c ^ 0xFFFFFFFFu;
}
