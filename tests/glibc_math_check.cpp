// TEST HELPER (host only): runs the restated glibc sinf / cosf of csrc/glibc_math.cuh against the C library of this
// machine.  usage: glibc_math_check <stride>   (stride 1 = every float with |x| < 120)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <thread>
#include <vector>
#include <atomic>
#include "../creating-2d-laser-slam-from-scratch_b200/csrc/glibc_math.cuh"

int main(int argc, char **argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 97u;
  const int variant = b2s::glibc_sincosf_variant_of_host();
  const uint32_t last = 0x42f00000u;  // 120.0f
  const int nt = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
  std::atomic<long> bad[2] = {{0}, {0}}, total{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      long b0 = 0, b1 = 0, n = 0;
      for (uint64_t u = (uint64_t)t * stride; u < last; u += (uint64_t)stride * nt) {
        for (int sg = 0; sg < 2; sg++) {
          uint32_t bits = (uint32_t)u | (sg ? 0x80000000u : 0u);
          float y;
          memcpy(&y, &bits, 4);
          volatile float vy = y;
          const float hs = sinf(vy), hc = cosf(vy);
          float hs2, hc2;
          sincosf(vy, &hs2, &hc2);
          for (int v = 0; v < 2; v++) {
            const float ms = b2s::glibc_sinf(y, v), mc = b2s::glibc_cosf(y, v);
            float ms2, mc2;
            b2s::glibc_sincosf(y, v, &ms2, &mc2);
            const bool miss = memcmp(&ms, &hs, 4) || memcmp(&mc, &hc, 4) || memcmp(&ms, &hs2, 4) || memcmp(&mc, &hc2, 4) ||
                              memcmp(&ms2, &hs, 4) || memcmp(&mc2, &hc, 4);
            if (miss) (v ? b1 : b0)++;
          }
          n++;
        }
      }
      bad[0] += b0; bad[1] += b1; total += n;
    });
  for (auto &x : th) x.join();
  printf("{\"host_variant\": %d, \"inputs\": %ld, \"mismatch_plain\": %ld, \"mismatch_fma\": %ld}\n", variant, total.load(),
         bad[0].load(), bad[1].load());
  return 0;
}
