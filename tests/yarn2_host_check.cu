// Host-side check program of tests/test_yarn2_cpu.py: prints the uniforms of chain `k` as (1) the product's generator arithmetic
// (csrc/yarn2.cuh: yarn2_jump / yarn2_mod / the two-table output map, executed on the HOST -- no kernel is launched) and (2) the
// restatement in the library's class shape that the reference CUDA drivers are compiled against (baseline/shim_yarn2).
// usage: yarn2_host_check seed seed_distance chain skip ndraws
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../neural_network_quantum_state_b200/csrc/yarn2.cuh"
#include "../baseline/shim_yarn2/trng/yarn2.hpp"
#include "../baseline/shim_yarn2/trng/uniform01_dist.hpp"

int main(int argc, char ** argv)
{
  if (argc != 6) return 2;
  const unsigned long long seed = std::strtoull(argv[1], nullptr, 10), dist = std::strtoull(argv[2], nullptr, 10),
    chain = std::strtoull(argv[3], nullptr, 10), skip = std::strtoull(argv[4], nullptr, 10);
  const int n = std::atoi(argv[5]);
  using namespace nqs;
  std::vector<uint32_t> tab((size_t)YARN2_TAB0+YARN2_TAB1);
  { // what yarn2_table_kernel writes, by running products (checked against yarn2_powmod at a few places)
    uint32_t x = 1u;
    for (int i = 0; i < YARN2_TAB0; ++i) { tab[i] = x; x = yarn2_mulmod(x, YARN2_GEN); }
    const uint32_t g16 = x;
    x = 1u;
    for (int i = 0; i < YARN2_TAB1; ++i) { tab[YARN2_TAB0+i] = x; x = yarn2_mulmod(x, g16); }
    for (int i : {0, 1, 2, 77, 65535}) if (tab[i] != yarn2_powmod(YARN2_GEN, (uint32_t)i)) return 3;
    for (int i : {0, 1, 2, 77, 32767}) if (tab[YARN2_TAB0+i] != yarn2_powmod(YARN2_GEN, (uint32_t)i<<16)) return 3;
  }
  uint32_t r0 = yarn2_seed_state(seed), r1 = 1u;
  yarn2_jump(r0, r1, 2ull*dist*chain);
  yarn2_jump(r0, r1, skip);
  trng::yarn2 e;
  e.seed((unsigned long)seed);
  e.jump(2ull*dist*chain);
  e.jump(skip);
  trng::uniform01_dist<double> u01;
  for (int t = 0; t < n; ++t)
  {
    const uint32_t nr = yarn2_mod((uint64_t)YARN2_A0*r0+(uint64_t)YARN2_A1*r1);
    r1 = r0; r0 = nr;
    const uint32_t x = (r0 == 0u) ? 0u : yarn2_mulmod(tab[YARN2_TAB0+(r0>>16)], tab[r0&0xffffu]);
    std::printf("%a %a\n", (double)x*(1.0/2147483647.0), u01(e));
  }
  return 0;
}
