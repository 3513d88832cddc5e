"""CPU stress test of the streamed host mode's staging protocol (csrc/host_copy.h, plain C++): the
slices of a caller-owned array are dealt out to up to four staging lanes (stepping thread + helper threads),
each lane copies its contiguous range slice by slice and advances its own published "slices staged" word
monotonically; a checker thread playing the GPU side verifies, every time it sees a lane's word move, that
all slices of that lane below the published count already hold the new data (release/acquire + store fence
after the streaming stores)."""
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
      const int W = CL_STAGE_MAX_LANES * CL_STAGE_WORD_STRIDE;
      uint32_t* words = (uint32_t*)aligned_alloc(64, W * sizeof(uint32_t));
      for (int k = 0; k < W; ++k) words[k] = 0;
      CopyHelper* helper = copy_helper_start_n(threads);
      if (threads >= 2 && (!helper || helper->n_workers != threads - 1)) return 2;
      std::atomic<uint32_t> cur_gen{0}, cur_nsl{0}, cur_spl{1}, errors{0}, stop{0};
      std::atomic<size_t> cur_per{0};
      std::atomic<int> cur_src{0};
      std::thread checker([&] {          // plays the GPU side: every lane's word on its own
        uint32_t last_word[CL_STAGE_MAX_LANES] = {0, 0, 0, 0};
        while (!stop.load()) {
          for (int lane = 0; lane < CL_STAGE_MAX_LANES; ++lane) {
            const uint32_t w = __atomic_load_n(words + lane * CL_STAGE_WORD_STRIDE, __ATOMIC_ACQUIRE);
            if (w == last_word[lane]) continue;
            if ((w >> 8) == (last_word[lane] >> 8) && (w & 255u) < (last_word[lane] & 255u)) errors++;   // went backwards
            last_word[lane] = w;
            const uint32_t gen = w >> 8, cnt = w & 255u;
            if (gen != cur_gen.load() || cnt == 0) continue;
            const size_t per = cur_per.load();
            const uint32_t spl = cur_spl.load();
            const unsigned char* s = srcs[cur_src.load()].data();
            const size_t from = (size_t)lane * spl * per;
            size_t upto = ((size_t)lane * spl + cnt) * per;
            if (upto > total) upto = total;
            if (from >= upto) continue;
            if (gen == cur_gen.load() && memcmp(dst + from, s + from, upto - from) != 0 && gen == cur_gen.load()) errors++;
          }
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
        uint32_t spl = 1;
        const uint32_t lanes = stage_plan(helper, nsl, &spl);
        if (lanes < 1 || lanes > (uint32_t)threads || (uint64_t)lanes * spl < nsl || (uint64_t)(lanes - 1) * spl >= nsl) {
          printf("bad plan: nsl %u lanes %u spl %u\n", nsl, lanes, spl); return 1;
        }
        cur_src = si; cur_per = per; cur_nsl = nsl; cur_spl = spl;
        cur_gen = gen;
        for (uint32_t k = 0; k < lanes; ++k) __atomic_store_n(words + k * CL_STAGE_WORD_STRIDE, gen << 8, __ATOMIC_RELEASE);
        stage_slices(helper, dst, srcs[si].data(), per, total, nsl, gen, words);
        if (memcmp(dst, srcs[si].data(), total) != 0) { printf("final mismatch at %d\n", it); return 1; }
        uint32_t sum = 0;
        for (uint32_t k = 0; k < lanes; ++k) {
          const uint32_t w = words[k * CL_STAGE_WORD_STRIDE];
          if ((w >> 8) != gen || (w & 255u) != stage_lane_count(k, spl, nsl)) { printf("lane %u count %u\n", k, w & 255u); return 1; }
          sum += w & 255u;
        }
        if (sum != nsl) { printf("count %u != %u\n", sum, nsl); return 1; }
      }
      stop = 1;
      checker.join();
      if (helper) copy_helper_stop(helper);
      printf("errors=%u\n", errors.load());
      return errors.load() ? 1 : 0;
    }
''')


@pytest.mark.parametrize("threads", [1, 2, 3, 4])
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
