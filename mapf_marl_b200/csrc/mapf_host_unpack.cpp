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
#include <stdlib.h>
#include <string.h>
#if defined(__linux__)
#include <sched.h>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
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

#if MAPF_X86
// AVX-512BW: one mask move expands 64 bits into a full cache line, written with one non-temporal store.
__attribute__((target("avx512f,avx512bw"))) void unpack_avx512(const uint32_t* bits, uint8_t* dst, size_t nwords) {
  const __m512i one = _mm512_set1_epi8(1);
  size_t i = 0;
  if ((((uintptr_t)dst) & 31) == 0) {
    if ((((uintptr_t)dst) & 63) != 0 && nwords) {            // 32-byte phase: one half line first
      _mm256_stream_si256((__m256i*)dst, _mm512_castsi512_si256(_mm512_maskz_mov_epi8((__mmask64)bits[0], one)));
      i = 1;
    }
    for (; i + 2 <= nwords; i += 2) {
      uint64_t m;
      memcpy(&m, bits + i, 8);
      _mm512_stream_si512((__m512i*)(dst + 32 * i), _mm512_maskz_mov_epi8((__mmask64)m, one));
    }
    if (i < nwords)
      _mm256_stream_si256((__m256i*)(dst + 32 * i),
                          _mm512_castsi512_si256(_mm512_maskz_mov_epi8((__mmask64)bits[i], one)));
    _mm_sfence();
    return;
  }
  for (; i + 2 <= nwords; i += 2) {
    uint64_t m;
    memcpy(&m, bits + i, 8);
    _mm512_storeu_si512((void*)(dst + 32 * i), _mm512_maskz_mov_epi8((__mmask64)m, one));
  }
  if (i < nwords)
    _mm256_storeu_si256((__m256i*)(dst + 32 * i), _mm512_castsi512_si256(_mm512_maskz_mov_epi8((__mmask64)bits[i], one)));
}

__attribute__((target("avx512f,avx512bw"))) void unpack_avx512_f32(const uint32_t* bits, uint8_t* dst8, size_t nwords) {
  float* dst = (float*)dst8;
  const __m512 onef = _mm512_set1_ps(1.0f);
  const bool aligned = (((uintptr_t)dst) & 63) == 0;       // 32 floats per word = 128 bytes: the phase never changes
  for (size_t i = 0; i < nwords; ++i) {
    const uint32_t w = bits[i];
    const __m512 lo = _mm512_maskz_mov_ps((__mmask16)(w & 0xffffu), onef);
    const __m512 hi = _mm512_maskz_mov_ps((__mmask16)(w >> 16), onef);
    if (aligned) {
      _mm512_stream_ps(dst + 32 * i, lo);
      _mm512_stream_ps(dst + 32 * i + 16, hi);
    } else {
      _mm512_storeu_ps(dst + 32 * i, lo);
      _mm512_storeu_ps(dst + 32 * i + 16, hi);
    }
  }
  if (aligned) _mm_sfence();
}
#endif

typedef void (*unpack_fn)(const uint32_t*, uint8_t*, size_t);

unpack_fn pick_unpack(int elem) {
#if MAPF_X86
  if (__builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f"))
    return elem == 4 ? unpack_avx512_f32 : unpack_avx512;
  if (__builtin_cpu_supports("avx2")) return elem == 4 ? unpack_avx2_f32 : unpack_avx2;
#endif
  return elem == 4 ? unpack_scalar_f32 : unpack_scalar;
}

inline void cpu_relax() {
#if MAPF_X86
  _mm_pause();
#else
  std::this_thread::yield();
#endif
}

}  // namespace

// A fixed pool of worker threads around ONE job at a time.  A job is the expansion of `cells` bits; it is cut into
// blocks of kBlockWords words that the workers (and the calling thread, inside finish()) claim in order from an atomic
// counter, so a worker that shares its core with somebody else delays nobody.  The words of a job become readable
// progressively (they are still crossing PCIe when the job begins): publish(n) says "the first n words are in
// memory"; a worker that claimed a block beyond that spins until it is.  The count can also live in a word the DEVICE
// writes (begin(..., ext_ready): a 4-byte copy queued behind every chunk on the same stream), so that no host thread
// has to wait on CUDA events and forward them: on a box with as many workers as cores that forwarding thread would
// share a core with a spinning worker.  The workers sleep on a condition variable only BETWEEN jobs, so a job costs
// one wake-up, not one per PCIe chunk.
struct MapfUnpackPool {
  static constexpr size_t kBlockWords = 1024;        // 4 KB of bits -> 32 KB of uint8 / 128 KB of float32
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  const uint32_t* src = nullptr;
  uint8_t* dst = nullptr;
  size_t nwords = 0, ncells = 0, nblocks = 0;
  int elem = 1;                     // bytes per output cell: 1 (uint8) or 4 (float32)
  unpack_fn fn = nullptr;
  uint64_t generation = 0;
  int active = 0;                   // workers that have not yet left the current job
  bool stop = false;
  std::atomic<size_t> next_block{0}, ready_words{0}, done_blocks{0};
  const volatile uint32_t* ext_ready = nullptr;   // words in memory, written by the device (pinned host word)
  std::atomic<bool> aborted{false};               // the caller gave up on the transfer: drain without expanding

  explicit MapfUnpackPool(int n) {
    for (int t = 0; t < n; ++t) workers.emplace_back([this]() { loop(); });
  }

  ~MapfUnpackPool() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_work.notify_all();
    for (auto& w : workers) w.join();
  }

  size_t words_ready() const {
    if (ext_ready) {
      const size_t w = *ext_ready;
      std::atomic_thread_fence(std::memory_order_acquire);
      return w;
    }
    return ready_words.load(std::memory_order_acquire);
  }

  // Claims and expands blocks until none is left.  `poll` (caller thread only) is asked now and then while waiting
  // for words; a non-zero answer aborts the job.
  void work(int (*poll)(void*) = nullptr, void* poll_arg = nullptr) {
    for (;;) {
      const size_t b = next_block.fetch_add(1, std::memory_order_relaxed);
      if (b >= nblocks) return;
      const size_t lo = b * kBlockWords;
      const size_t hi = lo + kBlockWords < nwords ? lo + kBlockWords : nwords;
      for (unsigned spins = 0; words_ready() < hi && !aborted.load(std::memory_order_relaxed); ++spins) {
        cpu_relax();
        if ((spins & 63) == 63) {
          if (poll && poll(poll_arg)) aborted.store(true, std::memory_order_relaxed);
          std::this_thread::yield();             // oversubscribed boxes (several ranks per host): let the others run
        }
      }
      if (aborted.load(std::memory_order_relaxed)) {
        done_blocks.fetch_add(1, std::memory_order_release);
        continue;
      }
      const size_t full = (hi == nwords && (ncells & 31)) ? hi - lo - 1 : hi - lo;   // the very last word may be partial
      fn(src + lo, dst + 32 * lo * elem, full);
      if (full != hi - lo) {
        uint8_t tail[128];
        (elem == 4 ? unpack_scalar_f32 : unpack_scalar)(src + hi - 1, tail, 1);
        memcpy(dst + 32 * (hi - 1) * elem, tail, (ncells & 31) * elem);
      }
      done_blocks.fetch_add(1, std::memory_order_release);
    }
  }

  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&]() { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--active == 0) cv_done.notify_one();
      }
    }
  }

  // Starts the expansion of `cells` bits at bits[0] into `cells` output elements of `elem_bytes` bytes each; no word
  // is readable yet.
  void begin(const uint32_t* bits, uint8_t* out, size_t cells, int elem_bytes, const volatile uint32_t* ext = nullptr) {
    std::lock_guard<std::mutex> lk(mu);
    ext_ready = ext;
    aborted.store(false, std::memory_order_relaxed);
    src = bits;
    dst = out;
    ncells = cells;
    elem = elem_bytes;
    fn = pick_unpack(elem_bytes);
    nwords = (cells + 31) / 32;
    nblocks = (nwords + kBlockWords - 1) / kBlockWords;
    next_block.store(0, std::memory_order_relaxed);
    done_blocks.store(0, std::memory_order_relaxed);
    ready_words.store(0, std::memory_order_relaxed);
    active = (int)workers.size();
    ++generation;
    cv_work.notify_all();
  }

  // The first `cells_ready` bits of the job are in memory (monotonic).
  void publish(size_t cells_ready) {
    const size_t w = cells_ready >= ncells ? nwords : cells_ready / 32;
    ready_words.store(w, std::memory_order_release);
  }

  // The caller helps with the remaining blocks, then waits until every block is written and every worker has left the
  // job (so that the next begin() cannot race with a late worker).
  bool finish(int (*poll)(void*) = nullptr, void* poll_arg = nullptr) {
    work(poll, poll_arg);
    std::unique_lock<std::mutex> lk(mu);
    // Timed wait: once the caller has no block left to claim it is the only thread that can notice a failed stream
    // (the workers would spin on words that never arrive), so it keeps polling until every worker has left the job.
    while (!cv_done.wait_for(lk, std::chrono::milliseconds(2), [&]() { return active == 0; })) {
      if (poll && !aborted.load(std::memory_order_relaxed)) {
        lk.unlock();
        const bool failed = poll(poll_arg) != 0;
        lk.lock();
        if (failed) aborted.store(true, std::memory_order_relaxed);
      }
    }
    return !aborted.load(std::memory_order_relaxed);
  }
};

extern "C" {

// threads <= 0: MAPF_HOST_THREADS if set, else the CPUs this process may run on divided by the number of ranks that
// share the host (LOCAL_WORLD_SIZE, the torchrun convention): eight ranks with one pool each must not put eight
// times the core count of spinning workers on the box.  The calling thread works too, hence the "- 1".
MapfUnpackPool* mapf_unpack_pool_create(int threads) {
  if (threads <= 0) {
    const char* env = getenv("MAPF_HOST_THREADS");
    if (env && atoi(env) > 0) {
      threads = atoi(env);
    } else {
      int cpus = 0;
#if defined(__linux__)
      cpu_set_t set;
      if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = CPU_COUNT(&set);
#endif
      if (cpus <= 0) cpus = (int)std::thread::hardware_concurrency();
      if (cpus <= 0) cpus = 4;
      const char* lws = getenv("LOCAL_WORLD_SIZE");
      const int ranks = (lws && atoi(lws) > 0) ? atoi(lws) : 1;
      threads = ranks > 1 ? cpus / ranks - 1 : cpus;
      if (threads < 1) threads = 1;
    }
    if (threads > 64) threads = 64;
  }
  try {
    return new MapfUnpackPool(threads);
  } catch (...) {
    return nullptr;
  }
}

void mapf_unpack_pool_destroy(MapfUnpackPool* p) { delete p; }

void mapf_unpack_pool_begin(MapfUnpackPool* p, const uint32_t* bits, void* out, size_t cells, int elem_bytes) {
  p->begin(bits, (uint8_t*)out, cells, elem_bytes);
}

void mapf_unpack_pool_publish(MapfUnpackPool* p, size_t cells_ready) { p->publish(cells_ready); }

void mapf_unpack_pool_finish(MapfUnpackPool* p) { p->finish(); }

// The job of one mapf_step_observe_host call: `ready_words` is a pinned host word the device updates behind every
// chunk; `poll` lets the calling thread notice a failed stream.  Returns 0 when the job was aborted.
int mapf_unpack_pool_expand_streamed(MapfUnpackPool* p, const uint32_t* bits, void* out, size_t cells, int elem_bytes,
                                     const volatile uint32_t* ready_words, int (*poll)(void*), void* poll_arg) {
  p->begin(bits, (uint8_t*)out, cells, elem_bytes, ready_words);
  return p->finish(poll, poll_arg) ? 1 : 0;
}

void mapf_unpack_pool_run(MapfUnpackPool* p, const uint32_t* bits, void* out, size_t cells, int elem_bytes) {
  p->begin(bits, (uint8_t*)out, cells, elem_bytes);
  p->publish(cells);
  p->finish();
}

int mapf_unpack_pool_threads(const MapfUnpackPool* p) { return p ? (int)p->workers.size() : 0; }

// include/mapf_b200.h: mapf_host_unpack -- expands cells [first_cell, first_cell + n_cells) of a host bit stream on the
// CALLING thread (a CPU consumer of the MAPF_BITS host output expands only the environments it reads).
// elem_bytes: 1 (uint8 0/1) or 4 (float32 0.0/1.0).
int mapf_host_unpack_range(const uint32_t* bits, uint64_t first_cell, uint64_t n_cells, void* out, int elem_bytes) {
  if (!bits || !out || (elem_bytes != 1 && elem_bytes != 4)) return -1;
  uint8_t* dst = (uint8_t*)out;
  uint64_t c = first_cell, done = 0;
  auto scalar_cells = [&](uint64_t cnt) {          // cell by cell: the unaligned head / tail of the range
    for (uint64_t k = 0; k < cnt; ++k, ++c, ++done) {
      const uint32_t b = (bits[c >> 5] >> (c & 31)) & 1u;
      if (elem_bytes == 1) dst[done] = (uint8_t)b;
      else ((float*)dst)[done] = (float)b;
    }
  };
  if (c & 31) scalar_cells(std::min<uint64_t>(n_cells, 32 - (c & 31)));
  const uint64_t words = (n_cells - done) >> 5;
  if (words) {
    pick_unpack(elem_bytes)(bits + (c >> 5), dst + done * elem_bytes, (size_t)words);
    c += words << 5;
    done += words << 5;
  }
  scalar_cells(n_cells - done);
  return 0;
}

}  // extern "C"
