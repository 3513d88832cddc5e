"""CPU stress test of the streamed host mode's staging protocol (csrc/host_copy.h, plain C++): the
stepping thread and the helper thread copy alternate slices of a caller-owned array into the staging
buffer and advance the published "slices staged" word monotonically; a checker thread playing the GPU
side verifies, every time it sees the word move, that all slices below the published count already
hold the new data (release/acquire + store fence after the streaming stores)."""
import os
import shutil
import subprocess
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = textwrap.dedent(r'''
    #include <stdio.h>
    #include <atomic>
    #include <thread>
    #include <vector>
    #include "host_copy.h"

    int main(int argc, char** argv) {
      const int threads = argc > 1 ? atoi(argv[1]) : 2;
      const size_t total = 786432 + 12 * 37;          // not a multiple of the slice size
      unsigned char* dst = (unsigned char*)aligned_alloc(4096, (total + 4095) / 4096 * 4096);
      std::vector<std::vector<unsigned char>> srcs(4, std::vector<unsigned char>(total));
      uint32_t* word = (uint32_t*)aligned_alloc(64, 64);
      *word = 0;
      CopyHelper* helper = nullptr;
      if (threads >= 2) {
        helper = (CopyHelper*)calloc(1, sizeof(CopyHelper));
        if (pthread_create(&helper->th, nullptr, copy_helper_main, helper) != 0) return 2;
        helper->started = true;
      }
      std::atomic<uint32_t> cur_gen{0}, cur_nsl{0}, errors{0}, stop{0};
      std::atomic<size_t> cur_per{0};
      std::atomic<int> cur_src{0};
      std::thread checker([&] {
        uint32_t last_word = 0;
        while (!stop.load()) {
          const uint32_t w = __atomic_load_n(word, __ATOMIC_ACQUIRE);
          if (w == last_word) continue;
          if ((w >> 8) == (last_word >> 8) && (w & 255u) < (last_word & 255u)) errors++;   // went backwards
          last_word = w;
          const uint32_t gen = w >> 8, cnt = w & 255u;
          if (gen != cur_gen.load() || cnt == 0) continue;
          const size_t per = cur_per.load();
          const unsigned char* s = srcs[cur_src.load()].data();
          const size_t upto = (size_t)cnt * per < total ? (size_t)cnt * per : total;
          if (gen == cur_gen.load() && memcmp(dst, s, upto) != 0 && gen == cur_gen.load()) errors++;
        }
      });
      uint32_t gen = 0;
      for (int it = 0; it < 3000; ++it) {
        const uint32_t nsl_want = 1 + (uint32_t)(it * 7 % 64);
        size_t per = ((total + nsl_want - 1) / nsl_want + 3071) / 3072 * 3072;
        const uint32_t nsl = (uint32_t)((total + per - 1) / per);
        const int si = it & 3;
        for (size_t k = 0; k < total; k += 97) srcs[si][k] = (unsigned char)(it + k);
        gen = (gen + 1) & 0x00FFFFFFu; if (gen == 0) gen = 1;
        cur_src = si; cur_per = per; cur_nsl = nsl;
        cur_gen = gen;
        __atomic_store_n(word, gen << 8, __ATOMIC_RELEASE);
        stage_slices(helper, dst, srcs[si].data(), per, total, nsl, gen, word);
        __atomic_store_n(word, (gen << 8) | nsl, __ATOMIC_RELEASE);
        if (memcmp(dst, srcs[si].data(), total) != 0) { printf("final mismatch at %d\n", it); return 1; }
        if ((*word & 255u) != nsl) { printf("count %u != %u\n", *word & 255u, nsl); return 1; }
      }
      stop = 1;
      checker.join();
      if (helper) copy_helper_stop(helper);
      printf("errors=%u\n", errors.load());
      return errors.load() ? 1 : 0;
    }
''')


@pytest.mark.parametrize("threads", [1, 2])
def test_staging_protocol_under_stress(tmp_path, threads):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "gym_lorenz_b200", "csrc"),
                    str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), str(threads)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "errors=0" in r.stdout
