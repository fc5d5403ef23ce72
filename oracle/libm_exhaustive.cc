// TEST INFRASTRUCTURE ONLY.  Exhaustive check of dspnet_b200/csrc/libm_compat.h (host build of the exact
// routine the CUDA kernels run) against the platform libm over all 2^32 binary32 inputs.
//   usage: libm_exhaustive <fma_build 0|1> [stride]
// glibc's variant is chosen at load time; run with GLIBC_TUNABLES=glibc.cpu.hwcaps=-AVX2,-FMA to make libm use
// its non-FMA build and check fma_build=0 against it.
// Prints: "<fn> mismatches <count> first <hex>" per function; exit code 0 iff both counts are 0.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include "../dspnet_b200/csrc/libm_compat.h"

int main(int argc, char **argv) {
  const bool fma_build = argc > 1 ? atoi(argv[1]) != 0 : true;
  const uint64_t stride = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
  const unsigned nthreads = std::max(1u, std::thread::hardware_concurrency());
  std::atomic<uint64_t> bad_exp{0}, bad_log{0};
  std::atomic<uint64_t> first_exp{~0ull}, first_log{~0ull};
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nthreads; ++t)
    pool.emplace_back([&, t] {
      uint64_t be = 0, bl = 0;
      for (uint64_t u = t * stride; u < (1ull << 32); u += nthreads * stride) {
        float x = dspmb::libm::ffrom((uint32_t)u);
        float a = expf(x), b = dspmb::libm::expf_glibc(x, fma_build);
        uint32_t ua = dspmb::libm::fbits(a), ub = dspmb::libm::fbits(b);
        if (ua != ub && !(a != a && b != b)) {
          if (!be++) { uint64_t e = first_exp.load(); while (u < e && !first_exp.compare_exchange_weak(e, u)) {} }
        }
        a = logf(x); b = dspmb::libm::logf_glibc(x, fma_build);
        ua = dspmb::libm::fbits(a); ub = dspmb::libm::fbits(b);
        if (ua != ub && !(a != a && b != b)) {
          if (!bl++) { uint64_t e = first_log.load(); while (u < e && !first_log.compare_exchange_weak(e, u)) {} }
        }
      }
      bad_exp += be; bad_log += bl;
    });
  for (auto &th : pool) th.join();
  printf("expf mismatches %llu first %llx\n", (unsigned long long)bad_exp.load(), (unsigned long long)first_exp.load());
  printf("logf mismatches %llu first %llx\n", (unsigned long long)bad_log.load(), (unsigned long long)first_log.load());
  return (bad_exp.load() || bad_log.load()) ? 1 : 0;
}
