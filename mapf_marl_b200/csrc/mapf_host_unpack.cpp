// Host side of the bit-packed PCIe transport of mapf_step_observe_host (include/mapf_b200.h).
//
// The field-of-view observation is 0/1 cells.  The device keeps them as bits until the final store, so the host entry
// point ships the BITS (8x fewer bytes than the uint8 tensor the caller asked for) and these worker threads expand
// them into the caller's buffer while the next chunk is still crossing PCIe.  This is a transport decode of values the
// GPU computed, not a computation of the path: there is no CPU implementation of step / observe in this library.
//
// Bit i of the stream == element i of the output (uint8 0/1 or float32 0.0/1.0; little-endian bit order inside
// 32-bit words).
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define MAPF_X86 1
#else
#define MAPF_X86 0
#endif

namespace {

// 32 bits -> 32 bytes per word, any alignment of dst, plain stores.
void unpack_scalar(const uint32_t* bits, uint8_t* dst, size_t nwords) {
  for (size_t i = 0; i < nwords; ++i) {
    const uint32_t w = bits[i];
    for (int k = 0; k < 4; ++k) {
      const uint64_t b = (w >> (8 * k)) & 0xffu;
      // bit j of b -> byte j: multiply spreads the byte to every byte lane, the mask keeps bit j in lane j,
      // adding 0x7f.. carries any kept bit into bit 7 of its lane
      uint64_t r = (b * 0x0101010101010101ULL) & 0x8040201008040201ULL;
      r = ((r + 0x7f7f7f7f7f7f7f7fULL) >> 7) & 0x0101010101010101ULL;
      memcpy(dst + 32 * i + 8 * k, &r, 8);
    }
  }
}

#if MAPF_X86
__attribute__((target("avx2"))) void unpack_avx2(const uint32_t* bits, uint8_t* dst, size_t nwords) {
  const __m256i sel = _mm256_setr_epi8(0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3,
                                       3, 3, 3, 3);
  const __m256i bitm = _mm256_setr_epi8(1, 2, 4, 8, 16, 32, 64, -128, 1, 2, 4, 8, 16, 32, 64, -128, 1, 2, 4, 8, 16, 32,
                                        64, -128, 1, 2, 4, 8, 16, 32, 64, -128);
  const __m256i one = _mm256_set1_epi8(1);
  const bool aligned = (((uintptr_t)dst) & 31) == 0;
  for (size_t i = 0; i < nwords; ++i) {
    __m256i v = _mm256_set1_epi32((int)bits[i]);
    v = _mm256_shuffle_epi8(v, sel);                 // byte lane b holds source byte b / 8
    v = _mm256_cmpeq_epi8(_mm256_and_si256(v, bitm), bitm);
    v = _mm256_and_si256(v, one);
    if (aligned) _mm256_stream_si256((__m256i*)(dst + 32 * i), v);   // written once, read by somebody else later
    else _mm256_storeu_si256((__m256i*)(dst + 32 * i), v);
  }
  if (aligned) _mm_sfence();
}
#endif

// The same for float32 cells (0.0f / 1.0f): 32 bits -> 128 bytes per word.
void unpack_scalar_f32(const uint32_t* bits, uint8_t* dst8, size_t nwords) {
  float* dst = (float*)dst8;
  for (size_t i = 0; i < nwords; ++i) {
    const uint32_t w = bits[i];
    for (int b = 0; b < 32; ++b) dst[32 * i + b] = (float)((w >> b) & 1u);
  }
}

#if MAPF_X86
__attribute__((target("avx2"))) void unpack_avx2_f32(const uint32_t* bits, uint8_t* dst8, size_t nwords) {
  float* dst = (float*)dst8;
  const __m256i bitm = _mm256_setr_epi32(1, 2, 4, 8, 16, 32, 64, 128);
  const __m256i onef = _mm256_set1_epi32(0x3f800000);
  const bool aligned = (((uintptr_t)dst) & 31) == 0;
  for (size_t i = 0; i < nwords; ++i) {
    const uint32_t w = bits[i];
#pragma GCC unroll 4
    for (int k = 0; k < 4; ++k) {
      __m256i v = _mm256_set1_epi32((int)((w >> (8 * k)) & 0xffu));
      v = _mm256_and_si256(_mm256_cmpeq_epi32(_mm256_and_si256(v, bitm), bitm), onef);
      if (aligned) _mm256_stream_si256((__m256i*)(dst + 32 * i + 8 * k), v);
      else _mm256_storeu_si256((__m256i*)(dst + 32 * i + 8 * k), v);
    }
  }
  if (aligned) _mm_sfence();
}
#endif

typedef void (*unpack_fn)(const uint32_t*, uint8_t*, size_t);

unpack_fn pick_unpack(int elem) {
#if MAPF_X86
  if (__builtin_cpu_supports("avx2")) return elem == 4 ? unpack_avx2_f32 : unpack_avx2;
#endif
  return elem == 4 ? unpack_scalar_f32 : unpack_scalar;
}

}  // namespace

// A fixed pool of worker threads; run() splits [0, nwords) statically and returns when every part is done.
struct MapfUnpackPool {
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  const uint32_t* src = nullptr;
  uint8_t* dst = nullptr;
  size_t nwords = 0, ncells = 0;
  int elem = 1;                     // bytes per output cell: 1 (uint8) or 4 (float32)
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;

  explicit MapfUnpackPool(int n) {
    for (int t = 0; t < n; ++t) workers.emplace_back([this, t, n]() { loop(t, n); });
  }

  ~MapfUnpackPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_work.notify_all();
    for (auto& w : workers) w.join();
  }

  void loop(int t, int n) {
    uint64_t seen = 0;
    for (;;) {
      const uint32_t* s;
      uint8_t* d;
      size_t lo, hi, nc;
      int el;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&]() { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        s = src;
        d = dst;
        nc = ncells;
        el = elem;
        // parts are multiples of 2 words so that every part but the first keeps the 64-byte phase of dst
        const size_t per = ((nwords + n - 1) / n + 1) & ~(size_t)1;
        lo = per * t < nwords ? per * t : nwords;
        hi = lo + per < nwords ? lo + per : nwords;
      }
      if (hi > lo) {
        const unpack_fn fn = pick_unpack(el);
        const size_t full = (hi == nwords && (nc & 31)) ? hi - lo - 1 : hi - lo;   // the very last word may be partial
        fn(s + lo, d + 32 * lo * el, full);
        if (full != hi - lo) {
          uint8_t tail[128];
          (el == 4 ? unpack_scalar_f32 : unpack_scalar)(s + hi - 1, tail, 1);
          memcpy(d + 32 * (hi - 1) * el, tail, (nc & 31) * el);
        }
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }

  // Expands `cells` bits starting at bits[0] into `cells` output elements of `elem_bytes` bytes each.
  void run(const uint32_t* bits, uint8_t* out, size_t cells, int elem_bytes) {
    std::unique_lock<std::mutex> lk(mu);
    src = bits;
    dst = out;
    ncells = cells;
    elem = elem_bytes;
    nwords = (cells + 31) / 32;
    pending = (int)workers.size();
    ++generation;
    cv_work.notify_all();
    cv_done.wait(lk, [&]() { return pending == 0; });
  }
};

extern "C" {

MapfUnpackPool* mapf_unpack_pool_create(int threads) {
  if (threads <= 0) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 4;
    if (threads > 32) threads = 32;
  }
  try {
    return new MapfUnpackPool(threads);
  } catch (...) {
    return nullptr;
  }
}

void mapf_unpack_pool_destroy(MapfUnpackPool* p) { delete p; }

void mapf_unpack_pool_run(MapfUnpackPool* p, const uint32_t* bits, void* out, size_t cells, int elem_bytes) {
  p->run(bits, (uint8_t*)out, cells, elem_bytes);
}

int mapf_unpack_pool_threads(const MapfUnpackPool* p) { return p ? (int)p->workers.size() : 0; }

}  // extern "C"
