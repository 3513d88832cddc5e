// Host-side staging of a caller-owned action array into the pinned buffer the step kernel reads
// (cl_step_host_async, streamed mode): staging lanes (one copy thread each, up to four), each publishing its own
// monotonic "slices staged" word; optional streaming-store copy.  Plain C++ (no CUDA) so that the protocol can be stress-tested
// on the CPU (tests/test_host_copy_cpu.py).
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

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



// ---- staging lanes ----------------------------------------------------------------------------
// A single core copies the caller's action array into pinned memory at 12-16 GB/s (50 us for the 786 KB
// of 65,536 Lorenz envs): in streamed mode that copy, not PCIe, is what the step waits for.  The slices are
// therefore dealt out to `lanes` CONTIGUOUS ranges ("lanes"), lane 0 copied by the stepping thread, lanes
// 1.. by helper threads, and every lane publishes its own progress word -- (gen << 8) | slices of THIS lane
// staged so far -- on a cache line of its own, so the copying threads share nothing and the blocks of every
// lane start as soon as their slice is there.  (Round-2 history, profiles/r02_e2e_copy_threads*.jsonl: two
// threads interleaved slice by slice advancing ONE word were slower than one thread, 94.6 vs 80.3 us; a
// helper that copied the upper half in one piece and reported once: 83-93 us against 94-106 with one thread,
// but its half of the transfers could not start before the whole half was staged.)
// Helper threads spin (pause) while steps keep coming, nap in 100 us sleeps once the env has been idle for a
// few ms, and are joined by cl_destroy.
#define CL_STAGE_MAX_LANES 4
#define CL_STAGE_WORD_STRIDE 16      // uint32 words between the progress words of two lanes (64 bytes)

struct CopyWorker {
  pthread_t th;
  bool started;
  volatile uint32_t job_gen;     // bumped by the stepping thread to start a job
  volatile uint32_t quit;
  // job (written before job_gen, read after): slices [first_slice, first_slice + n_slices) of per_bytes each
  unsigned char* dst;
  const unsigned char* src;
  size_t per_bytes, total_bytes;
  uint32_t first_slice, n_slices, gen;
  uint32_t* word;                  // this lane's progress word
  volatile uint32_t done;          // 1 once the lane is staged and published
  char pad[64];
};
struct CopyHelper {
  int n_workers;                   // helper threads = lanes - 1
  CopyWorker w[CL_STAGE_MAX_LANES - 1];
};

// One lane: slices [first, first + n) copied one by one, the lane's word advanced after each.
static void stage_lane(unsigned char* dst, const unsigned char* src, size_t per_bytes, size_t total_bytes,
                       uint32_t first, uint32_t n, uint32_t gen, uint32_t* word) {
  for (uint32_t j = 0; j < n; ++j) {
    const size_t b = (size_t)(first + j) * per_bytes;
    if (b < total_bytes) {
      const size_t e = b + per_bytes < total_bytes ? b + per_bytes : total_bytes;
      stage_copy_bytes(dst + b, src + b, e - b);
    }
    __atomic_store_n(word, (gen << 8) | (j + 1), __ATOMIC_RELEASE);
  }
}

static void* copy_worker_main(void* arg) {
  CopyWorker* c = (CopyWorker*)arg;
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
    stage_lane(c->dst, c->src, c->per_bytes, c->total_bytes, c->first_slice, c->n_slices, c->gen, c->word);
    __atomic_store_n(&c->done, 1u, __ATOMIC_RELEASE);
  }
  return nullptr;
}

// Copy threads this process should use for one env batch: CHAOS_B200_COPY_THREADS=1..4 decides; otherwise by
// the cores this process may run on, shared between the ranks of the node (LOCAL_WORLD_SIZE, torchrun): three
// threads from 8 cores per rank, two from 4, else one.  Measured at 65,536 envs, 16 cores, in-process A/B
// (profiles/r02l_e2e_lanes_*.jsonl): lorenz_rk4 84-88 / 70.5-74 / 69.4-71.1 / 69.0-71.7 us per step with
// 1 / 2 / 3 / 4 lanes (actions already pinned: 63.5); hr_sync 70-72 with 2 or 4; pmsm_sync 69-71 with 2 or 4 --
// a fourth thread buys nothing.
static int copy_threads_default(void) {
  if (const char* ov = getenv("CHAOS_B200_COPY_THREADS")) {
    const int t = atoi(ov);
    return t < 1 ? 1 : (t > CL_STAGE_MAX_LANES ? CL_STAGE_MAX_LANES : t);
  }
  cpu_set_t set;
  CPU_ZERO(&set);
  int cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 1;
  // an unpinned rank shares the machine with the other local ranks; a rank already pinned to its share
  // (distributed.pin_rank_to_cores) reports that share itself
  const long online = sysconf(_SC_NPROCESSORS_ONLN);
  if (const char* lw = getenv("LOCAL_WORLD_SIZE")) { const int k = atoi(lw); if (k > 1 && (long)cores >= online) cores /= k; }
  return cores >= 8 ? 3 : (cores >= 4 ? 2 : 1);
}

// nullptr: single-threaded staging (one lane).  Never started for batches whose actions are under 64 KB.
static CopyHelper* copy_helper_start_n(int threads) {
  if (threads > CL_STAGE_MAX_LANES) threads = CL_STAGE_MAX_LANES;
  if (threads < 2) return nullptr;
  CopyHelper* c = (CopyHelper*)calloc(1, sizeof(CopyHelper));
  if (!c) return nullptr;
  for (int k = 0; k < threads - 1; ++k) {
    if (pthread_create(&c->w[k].th, nullptr, copy_worker_main, &c->w[k]) != 0) break;
    c->w[k].started = true;
    c->n_workers = k + 1;
  }
  if (c->n_workers == 0) { free(c); return nullptr; }
  return c;
}

static CopyHelper* copy_helper_start(size_t action_bytes) {
  if (action_bytes < 64 * 1024) return nullptr;
  return copy_helper_start_n(copy_threads_default());
}

static void copy_helper_stop(CopyHelper* c) {
  if (!c) return;
  for (int k = 0; k < c->n_workers; ++k) __atomic_store_n(&c->w[k].quit, 1u, __ATOMIC_RELEASE);
  for (int k = 0; k < c->n_workers; ++k) if (c->w[k].started) pthread_join(c->w[k].th, nullptr);
  free(c);
}

// How `nsl` slices are dealt out: *spl slices per lane (the last lane may hold fewer); returns the lanes used.
static uint32_t stage_plan(const CopyHelper* c, uint32_t nsl, uint32_t* spl) {
  uint32_t lanes = 1u + (c ? (uint32_t)c->n_workers : 0u);
  if (lanes > nsl) lanes = nsl ? nsl : 1u;
  *spl = (nsl + lanes - 1u) / lanes;
  if (*spl == 0u) *spl = 1u;
  const uint32_t used = (nsl + *spl - 1u) / *spl;
  return used ? used : 1u;
}

static inline uint32_t stage_lane_count(uint32_t lane, uint32_t spl, uint32_t nsl) {
  const uint32_t first = lane * spl;
  return first >= nsl ? 0u : (nsl - first < spl ? nsl - first : spl);
}

// Stage `total_bytes` from src to dst in `nsl` slices of `per_bytes`; lane k publishes (gen << 8) | slices of
// lane k staged in words[k * CL_STAGE_WORD_STRIDE] after every slice.  Returns once every lane is complete.
static void stage_slices(CopyHelper* c, unsigned char* dst, const unsigned char* src, size_t per_bytes,
                         size_t total_bytes, uint32_t nsl, uint32_t gen, uint32_t* words) {
  uint32_t spl = 1;
  const uint32_t lanes = stage_plan(c, nsl, &spl);
  for (uint32_t k = 1; k < lanes; ++k) {
    CopyWorker* w = &c->w[k - 1];
    w->dst = dst; w->src = src; w->per_bytes = per_bytes; w->total_bytes = total_bytes;
    w->first_slice = k * spl; w->n_slices = stage_lane_count(k, spl, nsl); w->gen = gen;
    w->word = words + (size_t)k * CL_STAGE_WORD_STRIDE; w->done = 0;
    __atomic_store_n(&w->job_gen, w->job_gen + 1, __ATOMIC_RELEASE);
  }
  stage_lane(dst, src, per_bytes, total_bytes, 0, stage_lane_count(0, spl, nsl), gen, words);
  for (uint32_t k = 1; k < lanes; ++k)
    while (!__atomic_load_n(&c->w[k - 1].done, __ATOMIC_ACQUIRE)) __builtin_ia32_pause();
}
