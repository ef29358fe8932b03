// copy_pool.hpp — a few helper threads that copy a pageable host cloud into pinned staging memory in
// parallel (one core's memcpy, ~10 GB/s, is what bounds the latency of a pageable frame).
#pragma once
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace cp {

// memcpy with non-temporal stores: the staging buffer is read next by the GPU's DMA engine, not by this core,
// so the lines should land in DRAM instead of sitting dirty in a core's cache (where the DMA has to snoop them
// out, one core after another when several helpers wrote the buffer).
inline void copy_streaming(uint8_t* dst, const uint8_t* src, size_t len) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0 || len < 256) {
    std::memcpy(dst, src, len);
    return;
  }
  const size_t body = len & ~(size_t)63;
  for (size_t i = 0; i < body; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (body < len) std::memcpy(dst + body, src + body, len - body);
  _mm_sfence();
}

// `len` bytes (a multiple of 32) of value-initialised pcl::PointXYZI records: x = y = z = 0, data[3] = 1.0f,
// intensity = 0, padding 0 — what GroundRemover's cloud->points.resize(N) appends (src/ground_removal.cpp:79).
// Non-temporal stores when the destination allows it: 4 MB of padding should not evict a node's working set.
inline void fill_pad_points(uint8_t* dst, size_t len) {
  const __m128i lo = _mm_castps_si128(_mm_set_ps(1.0f, 0.0f, 0.0f, 0.0f)), hi = _mm_setzero_si128();
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) != 0) {
    for (size_t i = 0; i + 32 <= len; i += 32) {
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), lo);
      _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i + 16), hi);
    }
    return;
  }
  for (size_t i = 0; i + 32 <= len; i += 32) {
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), lo);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), hi);
  }
  _mm_sfence();
}

class CopyPool {
 public:
  struct Batch {
    uint8_t* dst = nullptr;
    const uint8_t* src = nullptr;
    size_t total = 0, piece = 0;
    bool streaming = true;
    uint32_t n_pieces = 0, pieces_per_group = 1, n_groups = 0;
    std::atomic<uint32_t> next{0};
    std::unique_ptr<std::atomic<uint32_t>[]> group_left;  // pieces of the group still to copy
    // copies pieces until none is left; returns after the last piece this thread took
    void work() {
      for (;;) {
        const uint32_t j = next.fetch_add(1, std::memory_order_relaxed);
        if (j >= n_pieces) return;
        copy_piece(j);
      }
    }
    bool take_one() {
      const uint32_t j = next.fetch_add(1, std::memory_order_relaxed);
      if (j >= n_pieces) return false;
      copy_piece(j);
      return true;
    }
    bool group_done(uint32_t g) const { return group_left[g].load(std::memory_order_acquire) == 0; }

   private:
    void copy_piece(uint32_t j) {
      const size_t off = (size_t)j * piece;
      const size_t len = off + piece <= total ? piece : total - off;
      if (!src) fill_pad_points(dst + off, len);          // src == NULL: a fill with PCL's padding point
      else if (streaming) copy_streaming(dst + off, src + off, len);
      else std::memcpy(dst + off, src + off, len);
      group_left[j / pieces_per_group].fetch_sub(1, std::memory_order_release);
    }
  };

  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) th_.emplace_back([this] { loop(); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  CopyPool(const CopyPool&) = delete;
  CopyPool& operator=(const CopyPool&) = delete;

  // pieces of `piece` bytes, `pieces_per_group` of them per group (a group = one DMA of the caller)
  std::shared_ptr<Batch> start(uint8_t* dst, const uint8_t* src, size_t total, size_t piece, uint32_t pieces_per_group,
                               bool streaming = true) {
    auto b = std::make_shared<Batch>();
    b->streaming = streaming;
    b->dst = dst;
    b->src = src;
    b->total = total;
    b->piece = piece;
    b->n_pieces = (uint32_t)((total + piece - 1) / piece);
    b->pieces_per_group = pieces_per_group;
    b->n_groups = (b->n_pieces + pieces_per_group - 1) / pieces_per_group;
    b->group_left.reset(new std::atomic<uint32_t>[b->n_groups]);
    for (uint32_t g = 0; g < b->n_groups; ++g) {
      const uint32_t first = g * pieces_per_group;
      b->group_left[g].store(std::min(pieces_per_group, b->n_pieces - first), std::memory_order_relaxed);
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      cur_ = b;
      ++gen_;
      taken_ = 0;
    }
    // wake ONE helper; each helper wakes the next before it starts copying.  Waking a halted core costs the
    // waker tens of microseconds (an IPI, more under a hypervisor), so the caller pays for one, not for all.
    // `to_wake_` counts the helpers that have not picked this generation up yet: a helper that is woken passes the
    // wake on while any are left, so the chain cannot die on a helper whose predicate is already false.
    cv_.notify_one();
    return b;
  }
  int workers() const { return (int)th_.size(); }

 private:
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      std::shared_ptr<Batch> b;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        b = cur_;
        ++taken_;
      }
      // pass the wake on.  notify_one may land on a helper that has already taken this generation (its predicate
      // is false, it goes back to sleep without passing anything on), so the chain is re-armed by every helper
      // that still finds work: a helper keeps waking others until the batch is handed out.
      cv_.notify_one();
      while (b->take_one()) {
        if (!chain_done(seen)) cv_.notify_one();
      }
    }
  }
  // true once every helper has picked generation `gen` up (then nobody is left to wake)
  bool chain_done(uint64_t gen) {
    std::lock_guard<std::mutex> lk(m_);
    return gen != gen_ || taken_ >= th_.size();
  }
  std::vector<std::thread> th_;
  size_t taken_ = 0;  // helpers that have picked the current generation up (guarded by m_)
  std::mutex m_;
  std::condition_variable cv_;
  std::shared_ptr<Batch> cur_;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

}  // namespace cp
