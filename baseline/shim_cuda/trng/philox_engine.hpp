// Stand-in for TRNG4 (v4.22, pinned by the reference's cmake/FindTRNG4.cmake:46-48; third-party, absent from /root/reference and
// not installable offline) for the SAME-BOX REFERENCE CUDA BUILD of bench.py's `gpu_reference` block (baseline/Makefile).
// TEST / BENCHMARK INFRASTRUCTURE ONLY: nothing under neural_network_quantum_state_b200/ includes this.
//
// The reference uses of the generator on this path (gpu/include/trng4cuda.cuh:14-65):
//   host:   rng[k].seed(seedNumber); rng[k].jump(2ul*seedDistance*k);        one engine per Markov chain
//   device: trng::uniform01_dist<FloatType>()(rng[idx])                     one draw per chain and proposal
// The engine here is Philox4x32-10 keyed by (seed, chain k, draw index) -- the function nqs::philox_uniform of
// neural_network_quantum_state_b200/csrc/device_math.cuh restated -- so the reference binary and libnqs_b200.so draw the SAME
// uniforms for the same seed and their accept/reject decisions can be compared chain by chain.  The yarn2 stream itself is not
// reproduced (parity unpinned at the RNG, as in oracle/shim/trng).
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define NQS_SHIM_HD __host__ __device__
#else
#define NQS_SHIM_HD
#endif
namespace trng
{
inline unsigned long long & shim_jump_unit() { static unsigned long long u = 0; return u; }

struct philox_engine
{
  unsigned long long seed_ = 0, chain_ = 0, ndraw_ = 0;
  void seed(unsigned long s) { seed_ = s; chain_ = 0; ndraw_ = 0; }
  void jump(unsigned long long s)
  { // called as jump(2*seedDistance*k), k = 0, 1, 2, ... in this order: the first non-zero argument is the stride between chains
    if (s == 0) { chain_ = 0; return; }
    if (shim_jump_unit() == 0) shim_jump_unit() = s;
    chain_ = s/shim_jump_unit();
  }
  NQS_SHIM_HD static void round(uint32_t c[4], uint32_t k0, uint32_t k1)
  {
    const uint64_t p0 = (uint64_t)0xD2511F53u*c[0], p1 = (uint64_t)0xCD9E8D57u*c[2];
    const uint32_t hi0 = (uint32_t)(p0>>32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1>>32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1^c[1]^k0, n2 = hi0^c[3]^k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  NQS_SHIM_HD double next01()
  {
    const unsigned long long step = ndraw_++;
    uint32_t c[4] = {(uint32_t)chain_, (uint32_t)(chain_>>32), (uint32_t)step, (uint32_t)(step>>32)};
    uint32_t k0 = (uint32_t)seed_, k1 = (uint32_t)(seed_>>32);
    for (int r = 0; r < 10; ++r)
    {
      round(c, k0, k1);
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return ((double)(c[0]>>5)*67108864.0+(double)(c[1]>>6))*(1.0/9007199254740992.0);
  }
};
struct yarn2 : public philox_engine {};
struct yarn5 : public philox_engine {};
struct yarn5s : public philox_engine {};
} // namespace trng
