// CPU test of csrc/copy_pool.hpp (the parallel pageable -> pinned staging copy): many batches of odd sizes,
// every byte must arrive, every group must be reported complete exactly when its pieces are.
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../cones_perception_b200/csrc/copy_pool.hpp"

int main(int argc, char** argv) {
  const int workers = argc > 1 ? std::atoi(argv[1]) : 3;
  cp::CopyPool pool(workers);
  std::vector<uint8_t> src(9u << 20), dst((9u << 20) + 64);
  uint64_t s = 0x9E3779B97F4A7C15ull;
  for (auto& b : src) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    b = (uint8_t)s;
  }
  const size_t sizes[] = {1, 63, 64, 4097, 65536, 65537, 524288, 1000003, 2097152, 5000011, 9u << 20};
  for (int rep = 0; rep < 40; ++rep)
    for (size_t total : sizes)
      for (int streaming = 0; streaming < 2; ++streaming) {
        uint8_t* d = dst.data() + ((reinterpret_cast<uintptr_t>(dst.data()) + 63) & ~(uintptr_t)63) - reinterpret_cast<uintptr_t>(dst.data());
        std::memset(d, 0xAB, total);
        const size_t off = (size_t)(rep * 131) % 4096;  // unaligned sources, like a message's std::vector
        if (off + total > src.size()) continue;
        auto b = pool.start(d, src.data() + off, total, 64 << 10, 4, streaming != 0);
        uint32_t issued = 0;
        while (b->take_one())
          while (issued < b->n_groups && b->group_done(issued)) ++issued;
        while (issued < b->n_groups) {
          if (b->group_done(issued)) ++issued;
          else std::this_thread::yield();
        }
        if (std::memcmp(d, src.data() + off, total) != 0) {
          std::printf("MISMATCH total=%zu rep=%d streaming=%d\n", total, rep, streaming);
          return 1;
        }
      }
  // fill mode (src == NULL): PCL's value-initialised PointXYZI, 32 bytes per point, aligned and unaligned targets
  for (size_t npts : {(size_t)1, (size_t)2047, (size_t)16384, (size_t)131072, (size_t)250001})
    for (size_t mis : {(size_t)0, (size_t)4}) {
      uint8_t* d = dst.data() + ((reinterpret_cast<uintptr_t>(dst.data()) + 63) & ~(uintptr_t)63) - reinterpret_cast<uintptr_t>(dst.data()) + mis;
      const size_t total = npts * 32;
      std::memset(d, 0xAB, total + 8);
      auto b = pool.start(d, nullptr, total, 64 << 10, 1, true);
      b->work();
      for (uint32_t g = 0; g < b->n_groups; ++g)
        while (!b->group_done(g)) std::this_thread::yield();
      const float one = 1.0f;
      for (size_t i = 0; i < npts; ++i) {
        uint8_t expect[32] = {0};
        std::memcpy(expect + 12, &one, 4);
        if (std::memcmp(d + 32 * i, expect, 32) != 0) {
          std::printf("FILL MISMATCH npts=%zu mis=%zu at %zu\n", npts, mis, i);
          return 1;
        }
      }
      if (d[total] != 0xAB) {
        std::printf("FILL OVERRUN npts=%zu\n", npts);
        return 1;
      }
    }
  std::printf("ok workers=%d\n", workers);
  return 0;
}
