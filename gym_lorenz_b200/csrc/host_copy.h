// Host-side staging of a caller-owned action array into the pinned buffer the step kernel reads
// (cl_step_host_async, streamed mode): streaming-store copy, optional second copy thread, monotonic
// publication of the "slices staged" word.  Plain C++ (no CUDA) so that the protocol can be stress-tested
// on the CPU (tests/test_host_copy_cpu.py).
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

// ---- second staging thread ------------------------------------------------------------------
// A single core copies the caller's action array into pinned memory at 12-16 GB/s (50 us for the 786 KB
// of 65,536 Lorenz envs): in streamed mode that copy, not PCIe, is what the step waits for.  A helper
// thread can take the upper half of the slices (CHAOS_B200_COPY_THREADS=2, see copy_helper_start).  It spins (pause) while steps keep coming, naps in 100 us sleeps once the env has
// been idle for 2 ms, and is joined by cl_destroy.  Never started when the process may use fewer than
// 4 cores or for batches whose actions are under 64 KB.
struct CopyHelper {
  pthread_t th;
  bool started;
  volatile uint32_t job_gen;     // bumped by the stepping thread to start a job
  volatile uint32_t quit;
  // job (written before job_gen, read after): slices [first_helper_slice, nsl) as one piece
  unsigned char* dst;
  const unsigned char* src;
  size_t per_bytes, total_bytes;
  uint32_t first_helper_slice;
  volatile uint32_t helper_done;   // 1 once the helper's piece is staged
};

static void stage_copy_bytes(void* dst, const void* src, size_t n);

static void* copy_helper_main(void* arg) {
  CopyHelper* c = (CopyHelper*)arg;
  uint32_t seen = 0;
  uint64_t idle = 0;
  while (!__atomic_load_n(&c->quit, __ATOMIC_ACQUIRE)) {
    const uint32_t g = __atomic_load_n(&c->job_gen, __ATOMIC_ACQUIRE);
    if (g == seen) {
      if (++idle < 400000) { __builtin_ia32_pause(); }
      else { struct timespec ts = {0, 100000}; nanosleep(&ts, nullptr); }
      continue;
    }
    seen = g;
    idle = 0;
    {   // the upper half of the slices in one piece; reported once (no shared traffic while copying)
      const size_t b = (size_t)c->first_helper_slice * c->per_bytes;
      if (b < c->total_bytes) stage_copy_bytes(c->dst + b, c->src + b, c->total_bytes - b);
      __atomic_store_n(&c->helper_done, 1u, __ATOMIC_RELEASE);
    }
  }
  return nullptr;
}

static CopyHelper* copy_helper_start(size_t action_bytes) {
  // Measured on the GPU box at 65,536 envs, streamed mode, 32 slices (profiles/r02_e2e_copy_threads.jsonl):
  // lorenz_rk4 94.9 -> 83.1 us per step with a warm source array, 97.0 -> 84.3 us with a cold one; hr_sync
  // 100.5 -> 93.5 / 88.1 -> 82.4.  (A first version that interleaved the two threads slice by slice and
  // had both advance the publication word was SLOWER than one thread: 94.6 vs 80.3 us.)
  // CHAOS_B200_COPY_THREADS=1 turns the helper off.
  int threads = 2;
  if (const char* ov = getenv("CHAOS_B200_COPY_THREADS")) threads = atoi(ov);
  cpu_set_t set;
  CPU_ZERO(&set);
  const int cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 1;
  if (threads < 2 || cores < 4 || action_bytes < 64 * 1024) return nullptr;
  CopyHelper* c = (CopyHelper*)calloc(1, sizeof(CopyHelper));
  if (!c) return nullptr;
  if (pthread_create(&c->th, nullptr, copy_helper_main, c) != 0) { free(c); return nullptr; }
  c->started = true;
  return c;
}

static void copy_helper_stop(CopyHelper* c) {
  if (!c) return;
  __atomic_store_n(&c->quit, 1u, __ATOMIC_RELEASE);
  if (c->started) pthread_join(c->th, nullptr);
  free(c);
}

// Staging copy caller array -> pinned buffer.  Optional variant with non-temporal (streaming) stores (AVX2;
// memcpy for the unaligned edges; ends with a store fence: the "slice staged" word that follows must not
// overtake the weakly ordered streaming stores) -- measured slower than memcpy here, kept for A/B.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void stream_copy_avx2(unsigned char* d, const unsigned char* s, size_t n) {
  size_t head = (32 - ((uintptr_t)d & 31)) & 31;
  if (head > n) head = n;
  if (head) { memcpy(d, s, head); d += head; s += head; n -= head; }
  size_t k = 0;
  for (; k + 128 <= n; k += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(s + k)), b = _mm256_loadu_si256((const __m256i*)(s + k + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(s + k + 64)), e = _mm256_loadu_si256((const __m256i*)(s + k + 96));
    _mm256_stream_si256((__m256i*)(d + k), a); _mm256_stream_si256((__m256i*)(d + k + 32), b);
    _mm256_stream_si256((__m256i*)(d + k + 64), c); _mm256_stream_si256((__m256i*)(d + k + 96), e);
  }
  for (; k + 32 <= n; k += 32) _mm256_stream_si256((__m256i*)(d + k), _mm256_loadu_si256((const __m256i*)(s + k)));
  if (k < n) memcpy(d + k, s + k, n - k);
  _mm_sfence();
}
#endif
static void stage_copy_bytes(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
  // default memcpy: measured on the GPU box at 65,536 envs, streamed mode, 16 / 32 slices: 89.5 / 89.5 us
  // per step with memcpy vs 94.1 / 92.3 us with streaming stores (the GPU reads the freshly written lines
  // out of the CPU's last-level cache; streamed stores send it to DRAM instead).  CHAOS_B200_STAGE_COPY=stream
  // selects the streaming-store copy.
  static int mode = -1;   // 1: AVX2 streaming stores, 0: memcpy
  if (mode < 0) {
    const char* ov = getenv("CHAOS_B200_STAGE_COPY");
    mode = (ov && !strcmp(ov, "stream") && __builtin_cpu_supports("avx2")) ? 1 : 0;
  }
  if (mode == 1 && n >= 4096) { stream_copy_avx2((unsigned char*)dst, (const unsigned char*)src, n); return; }
#endif
  memcpy(dst, src, n);
}


// Stage `total_bytes` from src to dst in `nsl` slices of `per_bytes`, publishing (gen << 8) | slices-staged
// in *word after every slice.  With a helper: even slices here, odd slices on the helper thread.
static void stage_slices(CopyHelper* c, unsigned char* dst, const unsigned char* src, size_t per_bytes,
                         size_t total_bytes, uint32_t nsl, uint32_t gen, uint32_t* word) {
  if (c && nsl >= 2) {
    // lower half here, slice by slice with publication (its blocks start while the copy goes on); upper
    // half on the helper in one piece; the full count is published once both are done
    const uint32_t half = nsl / 2;
    c->dst = dst; c->src = src; c->per_bytes = per_bytes; c->total_bytes = total_bytes;
    c->first_helper_slice = half; c->helper_done = 0;
    __atomic_store_n(&c->job_gen, c->job_gen + 1, __ATOMIC_RELEASE);
    for (uint32_t j = 0; j < half; ++j) {
      const size_t b = (size_t)j * per_bytes;
      stage_copy_bytes(dst + b, src + b, per_bytes);
      __atomic_store_n(word, (gen << 8) | (j + 1), __ATOMIC_RELEASE);
    }
    while (!__atomic_load_n(&c->helper_done, __ATOMIC_ACQUIRE)) __builtin_ia32_pause();
    __atomic_store_n(word, (gen << 8) | nsl, __ATOMIC_RELEASE);
  } else {
    uint32_t j = 0;
    for (size_t b = 0; b < total_bytes; b += per_bytes) {
      const size_t e = b + per_bytes < total_bytes ? b + per_bytes : total_bytes;
      stage_copy_bytes(dst + b, src + b, e - b);
      __atomic_store_n(word, (gen << 8) | ++j, __ATOMIC_RELEASE);
    }
  }
}
